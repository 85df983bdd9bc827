"""Timing of the one-CTA local solvers (small_solvers.cu) at the sizes of BASELINE.json
configs[0] (100^2 / 2 strips: 5100 rows, CG) and configs[2] (ani4_crop / METIS / 2, 4, 8
subdomains: 1616 / 830 / 457 rows, GMRES(30)): K iterations with a fixed budget, CUDA events."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
import schwz_b200 as S

ctx = S.Context(0)
z = np.load(os.path.join(ROOT, "tests", "golden", "ani4_crop.npz"))
ani4 = (z["rowptr"], z["col"], z["val"])


def run(name, mat, solver, K):
    rp, ci, v = mat
    n = len(rp) - 1
    A = S.Csr(ctx, rp, ci, v)
    b = ctx.to_device(np.ones(n))
    x = ctx.zeros(n)
    solver_obj = S.Gmres(ctx, A, 30) if solver == "gmres" else S.Cg(ctx, A)
    solver_obj.solve(b, x, K, 1e-300)
    ctx.sync()
    best = 1e9
    for _ in range(3):
        ctx.h2d(x, np.zeros(n))
        ctx.timer_start()
        solver_obj.solve(b, x, K, 1e-300)
        best = min(best, ctx.timer_stop())
    it, rn, r0 = solver_obj.result()
    print("%-28s n=%5d  %s  %4d iterations  %8.3f ms  %7.2f us/iteration  (rel. residual %.2e)"
          % (name, n, solver, it, best, 1e3 * best / max(it, 1), rn / r0), flush=True)
    solver_obj.close(); ctx.free(b); ctx.free(x); A.close()


for P in (8, 4, 2):
    part = S.partition_metis(ani4[0], ani4[1], P)
    setup = S.Setup(ani4, P, part=part)
    run("cfg3 ani4 METIS P=%d sub 1" % P, setup.local_matrix(1), "gmres", 300)
setup = S.Setup(("laplacian2d", 100), 2)
run("cfg1 100^2 strips P=2 sub 0", setup.local_matrix(0), "cg", 300)
ctx.close()
