TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8_r1n.json 2> gpurun_out/bench_n8_r1n.err; echo cfg2 rc=$?
$TR --master-port 29602 bench.py --gpus 8 --steps 5 --warmup 3 --dim 3 --size 512 --onesided > gpurun_out/bench_cfg4_n8_r1n.json 2> gpurun_out/bench_cfg4_n8_r1n.err; echo cfg4 rc=$?
$TR --master-port 29603 bench.py --gpus 8 --matrix ani4 --steps 50 --warmup 5 > gpurun_out/bench_cfg3_n8_r1n.json 2> gpurun_out/bench_cfg3_n8_r1n.err; echo cfg3 rc=$?
$TR --master-port 29604 bench.py --gpus 8 --size 2048 --steps 10 --warmup 3 --to-tolerance 40000 > gpurun_out/bench_tts2048_n8_r1n.json 2> gpurun_out/bench_tts2048_n8_r1n.err; echo tts rc=$?
for f in bench_n8_r1n bench_cfg4_n8_r1n bench_cfg3_n8_r1n bench_tts2048_n8_r1n; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").readlines()[-1])
    print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None, d.get("halo",{}).get("nvlink"), d.get("halo",{}).get("push_ms"), d.get("time_to_solution"))
except Exception as e:
    print("$f failed", e)
PY
done
