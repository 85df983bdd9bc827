"""CPU only, needs /root/reference (oracle/_ref): time per outer iteration of the REFERENCE'S OWN loop
(source/*.cpp compiled against the stand-ins, 8 rank threads, sequential stand-in Ginkgo per rank)
next to the oracle port on the same workload (n^2 5-pt Laplacian, 8 strips, CG with
local_max_iters = 50, synchronous exchange, global check).  Backs the statement in DESIGN.md section 6
that the port, which is what bench.py's CPU arm times, does not flatter the GPU.

    python tools/ref_vs_port.py [n=1024]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
import oracle as O
import ref as R
import schwz_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ts = {}
for K in (10, 30):      # max_iters must be >= the number of ranks (a window sizing quirk upstream)
    t = time.time()
    R.Run(8, laplacian_n=n, max_iters=K, local_max_iters=50, enable_global_check=True,
          tolerance=1e-30)
    ts[K] = time.time() - t
print("reference (8 rank threads): %.4f s per outer iteration (setup + 10 iterations %.2f s)"
      % ((ts[30] - ts[10]) / 20, ts[10]))
setup = S.Setup(("laplacian2d", n), 8)
rp, ci, v = setup.local_matrix(1)
nn = len(rp) - 1
for th in (1, O.max_threads()):
    O.set_threads(th)
    b = np.ones(nn)
    x = np.zeros(nn)
    x, _ = O.cg(rp, ci, v, b, x, 50, 1e-12)
    t = time.time()
    for _ in range(3):
        O.spmv(rp, ci, v, x, -1.0, 1.0, b)
        x, _ = O.cg(rp, ci, v, b, x, 50, 1e-12)
    dt = (time.time() - t) / 3
    print("port, one strip with %d thread(s): %.4f s per (residual + 50 CG) -> %.4f s per outer "
          "iteration with the 8 strips %s"
          % (th, dt, dt if th == 1 else 8 * dt,
             "side by side on 8 cores" if th == 1 else "one after the other"))
