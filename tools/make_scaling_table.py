"""profiles/r2_scaling/*_n{N}.json (the bench.py lines tools/run_scaling.sh wrote at N = 1, 2, 4, 8 GPUs)
-> the markdown table of profiles/r2_scaling.md.

    python tools/make_scaling_table.py > /tmp/table.md
"""
import csv
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "profiles", "r2_scaling")
ORDER = ["cfg2_strips", "cfg2_blocks2d", "cfg2_2x2", "cfg2_inexact", "cfg2_tts_big", "cfg3_ani4", "cfg3_ani3",
         "cfg4_3d512_onesided"]
rows = {}
for f in glob.glob(os.path.join(D, "*.json")):
    m = re.match(r"(.*)_n(\d+)\.json", os.path.basename(f))
    if not m:
        continue
    lines = [l for l in open(f) if l.startswith("{")]
    if lines:
        rows[(m.group(1), int(m.group(2)))] = json.loads(lines[-1])
Ns = sorted({n for _, n in rows})
print("| config | " + " | ".join("N = %d" % n for n in Ns) + " | speed-up 1 -> %d |" % Ns[-1])
print("|---|" + "---|" * (len(Ns) + 1))
for name in ORDER:
    cells = []
    for n in Ns:
        d = rows.get((name, n))
        cells.append("%.2f it/s (%.3f ms; e2e %s)" % (d["value"], d["ms_per_step"],
                                                       "%.2f" % d["e2e"]["value"] if d.get("e2e") else "-")
                     if d else "-")
    a, b = rows.get((name, Ns[0])), rows.get((name, Ns[-1]))
    print("| %s | %s | %s |" % (name, " | ".join(cells), "%.2fx" % (b["value"] / a["value"]) if a and b else "-"))
# cfg5 through bench_ras: stage timers of rank 0 / 100 iterations
cells = []
for n in Ns:
    f = os.path.join(D, "cfg5_bench_ras_timings_rank0_n%d.csv" % n)
    if os.path.exists(f):
        tot = sum(float(r[1]) for r in list(csv.reader(open(f)))[1:])
        cells.append("%.2f it/s (%.3f ms)" % (100 / tot, 10 * tot))
    else:
        cells.append("-")
print("| cfg5 (bench_ras, 64 subdomains, 100 iterations) | %s | - |" % " | ".join(cells))
print()
print("| to tolerance (set_tol 1e-6, zero start, host buffers in and out) | N | outer iterations | seconds | g/g0 | true relative residual | CPU arm, same run |")
print("|---|---|---|---|---|---|---|")
for (name, n), d in sorted(rows.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    for key in ("time_to_solution", "time_to_solution_full"):
        t = d.get(key)
        if not t:
            continue
        size = t.get("n", d["config"].get("n"))
        print("| %s %dx%d | %d | %d%s | %.2f | %.3e | %.2e | %s |"
              % (name, size, size, n, t["outer_iterations"], "" if t["converged"] else " (not converged)",
                 t["time_to_solution_s"], t["global_resnorm_ratio"] or float("nan"),
                 t.get("true_relative_residual", float("nan")),
                 "%.0f s (%s)" % (t["cpu_baseline_s"], t["cpu_baseline_how"]) if t.get("cpu_baseline_s") else "-"))
print()
print("| halo exchange | N | push (pack + peer stores + publish) | unpack | payload over NVLink per push | GB/s |")
print("|---|---|---|---|---|---|")
for (name, n), d in sorted(rows.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    h = d.get("halo") or {}
    nv = h.get("nvlink") or {}
    if name in ("cfg2_strips", "cfg4_3d512_onesided"):
        print("| %s | %d | %.1f us | %.1f us | %d B | %s |"
              % (name, n, 1e3 * h["push_ms"], 1e3 * h["unpack_ms"], nv.get("payload_bytes", 0),
                 "%.1f" % nv["GB/s"] if nv.get("GB/s") else "-"))
