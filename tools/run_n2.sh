# N=2 data points of round 1: cfg2 (scaling), cfg4 (one remote neighbour per GPU boundary: push timing)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_r1n.json 2> gpurun_out/bench_n2_r1n.err; echo cfg2 rc=$?
$TR --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --dim 3 --size 512 --onesided > gpurun_out/bench_cfg4_n2_r1n.json 2> gpurun_out/bench_cfg4_n2_r1n.err; echo cfg4 rc=$?
for f in bench_n2_r1n bench_cfg4_n2_r1n; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").readlines()[-1])
    print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None, d.get("halo"))
except Exception as e:
    print("$f failed", e)
PY
done
