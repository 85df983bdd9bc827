"""Timing of the local preconditioners on one interior strip of configuration 2
(8192^2 / 8 strips: 8 404 992 rows): one application of each kind, and a
preconditioned CG iteration, CUDA events on the context's stream.

    python tools/prof_precond.py [n=8192] [reps=10] [kinds=block-jacobi,isai,ilu]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
import schwz_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kinds = (sys.argv[3] if len(sys.argv) > 3 else "block-jacobi,isai,ilu").split(",")
setup = S.Setup(("laplacian2d", n), 8)
rp, ci, v = setup.local_matrix(1)
rows = len(rp) - 1
ctx = S.Context(0)
A = S.Csr(ctx, rp, ci, v)
r = ctx.to_device(np.random.default_rng(0).standard_normal(rows))
z = ctx.zeros(rows)
d = ctx.zeros(1)
b = ctx.to_device(np.ones(rows))
out = {"rows": rows, "nnz": int(rp[-1])}
for kind in kinds:
    t0 = time.perf_counter()
    M = S.Precond(ctx, rp, ci, v, kind, 16)
    gen = time.perf_counter() - t0
    M.apply(r, z, d)
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        M.apply(r, z, d)
    ms = ctx.timer_stop() / reps
    by = M.bytes_per_apply()
    # K preconditioned CG iterations (fixed budget), against K plain ones
    K = 20
    res = {}
    for name, pc in (("plain", None), (kind, M)):
        cg = S.Cg(ctx, A, precond=pc)
        x = ctx.zeros(rows)
        cg.solve(b, x, K, 1e-300)
        ctx.sync()
        ctx.h2d(x, np.zeros(rows))
        ctx.timer_start()
        cg.solve(b, x, K, 1e-300)
        t = ctx.timer_stop()
        it, rn, r0 = cg.result()
        res[name] = {"ms_per_iteration": t / K, "relative_residual_after_%d" % K: rn / r0}
        ctx.free(x)
        cg.close()
    out[kind] = {"generate_s": gen, "apply_us": ms * 1e3, "bytes_per_apply": by,
                 "GB/s": by / ms / 1e6, "cg": res}
    print("%-13s generate %6.2f s   apply %9.1f us   %7.1f GB/s (%d B)   CG it: %.3f ms vs plain %.3f ms"
          % (kind, gen, ms * 1e3, by / ms / 1e6, by, res[kind]["ms_per_iteration"],
             res["plain"]["ms_per_iteration"]), flush=True)
    M.close()
print(json.dumps(out))
