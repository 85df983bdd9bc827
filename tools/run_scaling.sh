#!/bin/bash
# Every BASELINE.json config at N GPUs of one box (round 2 table: profiles/r2_scaling.md).
#   tools/run_scaling.sh N [TTS_SIZE] [OUTDIR]
# cfg2 (strips and the 2-D px x py partition, both with a to-tolerance leg at TTS_SIZE^2),
# cfg2 in the inexact regime of the reference's run_script, cfg3 (ani4 / ani3), cfg4, and cfg5
# through the drop-in bench_ras.  One JSON line per run in OUTDIR, a summary table at the end.
set -u
N=${1:-1}
TTS=${2:-1024}
OUT=${3:-gpurun_out/scale_n$N}
mkdir -p "$OUT"
port=29700
run() {
    name=$1; shift
    # RUNS="cfg2_strips cfg3_ani4": only those
    if [ -n "${RUNS:-}" ] && [[ " $RUNS " != *" $name "* ]]; then return; fi
    port=$((port + 1))
    if [ "$N" -gt 1 ]; then
        timeout ${TMO:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
            --master-addr 127.0.0.1 --master-port $port bench.py --gpus "$N" "$@" \
            > "$OUT/$name.json" 2> "$OUT/$name.err"
    else
        timeout ${TMO:-900} python bench.py --gpus 1 "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"
    fi
    echo "$name rc=$?"
}
run cfg2_strips --steps 10 --warmup 3 --no-cpu-baseline --tts-size "$TTS"
# the reference's own regular2d rule: 2 x 2 blocks (4 subdomains)
run cfg2_2x2 --subdomains 4 --partition regular2d --steps 10 --warmup 3 --no-cpu-baseline --tts-size "$TTS"
run cfg2_blocks2d --steps 10 --warmup 3 --no-cpu-baseline --partition regular2d --tts-size "$TTS"
run cfg2_inexact --steps 10 --warmup 3 --no-cpu-baseline --tts-size 0 --local-tol 0.1 --local-iters 70
run cfg3_ani4 --matrix ani4 --steps 100 --warmup 10 --no-cpu-baseline
run cfg3_ani3 --matrix ani3 --steps 50 --warmup 10 --no-cpu-baseline    # converges at 99 (P = 8)
TMO=1500 run cfg4_3d512_onesided --dim 3 --size 512 --onesided --steps 5 --warmup 3 --no-cpu-baseline --tts-size 0
if [ -n "${TTS_BIG:-}" ]; then
    # a to-tolerance run of the strips workload at TTS_BIG^2 (recorded as time_to_solution_full)
    TMO=1500 run cfg2_tts_big --size "$TTS_BIG" --steps 10 --warmup 3 --to-tolerance 200000 --tts-size 0 --no-cpu-baseline
fi
if [ "${SKIP_CFG5:-0}" != "1" ] && { [ -z "${RUNS:-}" ] || [[ " $RUNS " == *" cfg5 "* ]]; }; then
    NUM_DEVICES=$N ONLY=cfg5 TMO=1500 tools/run_configs.sh "$OUT/bench_ras" > "$OUT/cfg5.log" 2>&1
    grep -E "Rank 0 |Time taken|relative residual|real" "$OUT/cfg5.log" | head -6
fi
python - "$OUT" "$N" <<'PY'
import glob, json, os, re, sys
out, N = sys.argv[1], sys.argv[2]
print("| config | N | outer iters/s | ms/iter | e2e | halo push us (NVLink GB/s) | to tolerance |")
print("|---|---|---|---|---|---|---|")
for f in sorted(glob.glob(os.path.join(out, "*.json"))):
    name = os.path.basename(f)[:-5]
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print("| %s | %s | failed: %s |" % (name, N, e))
        continue
    h = d.get("halo") or {}
    nv = (h.get("nvlink") or {}).get("GB/s")
    t = d.get("time_to_solution")
    tt = ("%dx%d: %d its, %.2f s%s" % (t["n"], t["n"], t["outer_iterations"], t["time_to_solution_s"],
                                       "" if t["converged"] else " (NOT converged)")) if t else "-"
    tf = d.get("time_to_solution_full")
    if tf:
        tt = "%dx%d: %d its, %.2f s%s, true rel. residual %.2e" % (
            d["config"]["n"], d["config"]["n"], tf["outer_iterations"], tf["time_to_solution_s"],
            "" if tf["converged"] else " (NOT converged)", tf["true_relative_residual"])
    print("| %s | %s | %.2f | %.3f | %s | %.1f (%s) | %s |"
          % (name, N, d["value"], d["ms_per_step"],
             "%.2f" % d["e2e"]["value"] if d.get("e2e") else "-",
             1e3 * h.get("push_ms", float("nan")), "%.1f" % nv if nv else "-", tt))
logs = glob.glob(os.path.join(out, "bench_ras", "cfg5*", "run.log"))
if logs:
    s = open(logs[0]).read()
    it = re.search(r"Rank 0 converged in (\d+) iterations", s)
    tm = re.search(r"Time taken for solve ([0-9.eE+-]+)", s)
    if it and tm:
        print("| cfg5 (bench_ras, 64 subdomains) | %s | %.2f | %.3f | - | - | %d its, %.2f s |"
              % (N, int(it.group(1)) / float(tm.group(1)), 1e3 * float(tm.group(1)) / int(it.group(1)),
                 int(it.group(1)), float(tm.group(1))))
    else:
        # fixed budget (--num_iters=100, not converged): per-iteration time = the stage totals
        # of rank 0's timing CSV (source/schwarz_base.cpp's timers) / iterations
        import csv
        tfile = glob.glob(os.path.join(out, "bench_ras", "cfg5*", "timings_00.csv"))
        nit = re.search(r"did not converge in (\d+) iterations", s)
        if tfile and nit:
            tot = sum(float(r[1]) for r in list(csv.reader(open(tfile[0])))[1:])
            k = int(nit.group(1))
            print("| cfg5 (bench_ras, 64 subdomains, stage by stage) | %s | %.2f | %.3f | - | - | %d its (budget) |"
                  % (N, k / tot, 1e3 * tot / k, k))
PY
