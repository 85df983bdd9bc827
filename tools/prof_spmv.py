"""Profiling driver: a handful of launches of the CG SpMV (q = A p with the
p.q reduction fused) on one interior strip of configuration 2, for
`ncu --set full -k regex:csr_spmv`.  Prints the CUDA-event time as well."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
import schwz_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
setup = S.Setup(("laplacian2d", n), 8)
ctx = S.Context(0)
sub = S.Ras(ctx, setup, 1, local_max_iters=50)
# one real local solve first, so that the CG scalars (rho, beta, iteration count) hold live
# values: with the zero-initialised ones the vector kernels take their degenerate branches
# (beta == 0 skips the x/r update, iteration 0 turns the p update into a copy)
sub.update_boundary()
sub.local_residual()
sub.local_solve()
sub.sync()
for kind, name in ((0, "spmv+dot"), (3, "residual spmv+norm"), (1, "cg r update"), (2, "cg x/p update")):
    ms = sub.kernel_time_ms(kind, reps)
    b = sub.kernel_bytes(kind)
    print("%-20s %8.3f us  %8.1f GB/s  (%d bytes)" % (name, ms * 1e3, b / ms / 1e6, b))
sub.close()
ctx.close()
