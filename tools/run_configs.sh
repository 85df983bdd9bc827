#!/bin/bash
# Full-size runs of BASELINE.json configs[2..4] (and cfg2 through the C++ driver) with the
# drop-in bench_ras on one B200 (all subdomains on GPU 0 unless NUM_DEVICES is set).
# Usage: tools/run_configs.sh <outdir> ; each config gets <outdir>/<name>.log + timing CSVs.
set -u
OUT=${1:-gpurun_out/configs}
ND=${NUM_DEVICES:-1}
BIN=$(dirname "$0")/../schwarz-lib_b200/bin/bench_ras
BIN=$(readlink -f "$BIN")
mkdir -p "$OUT"
run() {
    name=$1; shift
    if [[ -n "${ONLY:-}" && "$name" != *"$ONLY"* ]]; then return; fi
    mkdir -p "$OUT/$name"
    ( cd "$OUT/$name" && { time timeout ${TMO:-600} "$BIN" --executor=cuda --num_devices=$ND \
        --timings_file=timings "$@" > run.log 2> run.err; echo "rc=$?" >> run.log; } 2> wall.txt )
    echo "== $name"; grep -E "converged|did not converge|Time taken|relative residual|rc=|local problem size" "$OUT/$name/run.log" | head -12
    grep real "$OUT/$name/wall.txt"; head -c 600 "$OUT/$name/run.err"
}
# the .mtx of cfg3 is rebuilt from the committed fixture (/root/reference is not on the GPU box)
python - "$OUT/ani4_crop.mtx" <<'PY'
import sys, numpy as np, os
z = np.load(os.path.join(os.path.dirname(os.path.abspath(sys.argv[0])) if False else "tests/golden", "ani4_crop.npz"))
rp, ci, v = z["rowptr"], z["col"], z["val"]
n = len(rp) - 1
rows = np.repeat(np.arange(n), np.diff(rp))
with open(sys.argv[1], "w") as f:
    f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(ci)))
    for r, c, x in zip(rows, ci, v):
        f.write("%d %d %.17g\n" % (r + 1, c + 1, x))
PY
MTX=$(readlink -f "$OUT/ani4_crop.mtx")
# cfg3: ani4_crop, METIS, GMRES(30), overlap 2, P = 2 / 4 / 8
for P in 2 4 8; do
  run cfg3_ani4_P$P --matrix_filename=$MTX --partition=metis --overlap=2 \
      --non_symmetric_matrix --restart_iter=30 --local_solver=iterative-ginkgo --enable_global_check \
      --set_tol=1e-6 --local_tol=1e-12 --num_iters=800 --num_subdomains=$P
done
# cfg2 through the C++ SolverRAS (same workload as bench.py): 8192^2, 8 strips, CG 50
run cfg2_lap8192_P8 --explicit_laplacian --set_1d_laplacian_size=8192 --partition=regular --overlap=2 \
    --local_solver=iterative-ginkgo --local_max_iters=50 --enable_global_check --set_tol=1e-6 \
    --local_tol=1e-12 --num_iters=20 --num_subdomains=8
# cfg5: 4096^2, regular2d 8x8 = 64 subdomains, factorised local solve
TMO=1200 run cfg5_lap4096_P64_direct --explicit_laplacian --set_1d_laplacian_size=4096 --partition=regular2d --overlap=2 \
    --local_solver=direct-ginkgo --local_factorization=cholmod --enable_global_check --set_tol=1e-6 \
    --num_iters=100 --num_subdomains=64
# cfg4: 3-D 7-pt 512^3, 8 slabs, one-sided Put gathered, decentralised convergence
TMO=1200 run cfg4_lap3d512_P8_onesided --explicit_laplacian --laplacian_dim=3 --set_1d_laplacian_size=512 \
    --partition=regular --overlap=2 --local_solver=iterative-ginkgo --local_max_iters=50 \
    --enable_onesided --remote_comm_type=put --global_convergence_type=decentralized \
    --set_tol=1e-6 --local_tol=1e-12 --num_iters=20 --num_subdomains=8
