"""Timing of the level-scheduled triangular solves on the cfg5 factor (4096^2, regular2d 8x8:
one 264 192-row subdomain, nested-dissection Cholesky, nnz(L) = 7.9 M): one L solve + one
U = L^T solve, (a) alone on the GPU and (b) S solves side by side on S streams (the
oversubscribed regime of cfg5: 8 subdomains per GPU on 8 GPUs, 64 on one).

    python tools/prof_trs.py [S=16] [reps=20] [ilu]

SCHWZ_B200_TRS_LEVELS=1 selects the level-per-launch graph (round 1), default is the
dependency-driven one-kernel solve.  With a third argument "ilu": the ILU(0) factors of a
cfg2 strip (8192 x 1026 grid, natural order: ~9 200 wavefront levels) instead.
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
import schwz_b200 as S

nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ilu = len(sys.argv) > 3 and sys.argv[3] == "ilu"
if ilu:
    # ILU(0) of the 5-pt Laplacian keeps the pattern of tril / triu; the values do not matter
    # for the timing, a diagonally dominant pair with that pattern is used
    n, P = 8192, 8
    setup = S.Setup(("laplacian2d", n), P)
    rp, ci, v = setup.local_matrix(1)
    rows = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(rows, rows))
    Lm = sp.tril(A).tocsr(); Lm.sort_indices()
    Um = sp.triu(A).tocsr(); Um.sort_indices()
    Lrp, Lci, Lv = Lm.indptr.astype(np.int32), Lm.indices.astype(np.int32), Lm.data
    Urp, Uci, Uv = Um.indptr.astype(np.int32), Um.indices.astype(np.int32), Um.data
else:
    n, P = 4096, 64
    setup = S.Setup(("laplacian2d", n), P, part=S.partition_regular2d(n * n, P))
    rp, ci, v = setup.local_matrix(9)
    rows = len(rp) - 1
    perm = S.nd_ordering(rp, ci)
    Lrp, Lci, Lv = S.host_cholesky(rp, ci, v, perm)
    U = sp.csr_matrix((Lv, Lci, Lrp), shape=(rows, rows)).T.tocsr()
    U.sort_indices()
    Urp, Uci, Uv = U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data
ctxs = [S.Context(0) for _ in range(nsub)]
plans = []
b = np.random.default_rng(0).standard_normal(rows)
for c in ctxs:
    tl = S.Trs(c, Lrp, Lci, Lv, upper=False)
    tu = S.Trs(c, Urp, Uci, Uv, upper=True)
    plans.append((c, tl, tu, c.to_device(b), c.zeros(rows), c.zeros(rows)))
c, tl, tu, db, dy, dz = plans[0]
tl.solve(db, dy); tu.solve(dy, dz); c.sync()
z = c.to_host(dz, rows)
Lm = sp.csr_matrix((Lv, Lci, Lrp), shape=(rows, rows))
Um = sp.csr_matrix((Uv, Uci, Urp), shape=(rows, rows))
print("residual |L U z - b| / |b| = %.2e, levels %d, error words %d %d, mode %s"
      % (np.linalg.norm(Lm @ (Um @ z) - b) / np.linalg.norm(b), tl.levels(), tl.error(), tu.error(),
         "levels" if os.environ.get("SCHWZ_B200_TRS_LEVELS") == "1" else "one-kernel"))
# (a) alone
c.timer_start()
for _ in range(reps):
    tl.solve(db, dy); tu.solve(dy, dz)
ms = c.timer_stop() / reps
byts = 12 * (int(Lrp[-1]) + int(Urp[-1])) + 2 * 20 * rows
print("alone: L + U solve %.1f us, %.1f GB/s algorithmic" % (ms * 1e3, byts / ms / 1e6))
# (b) nsub side by side
for (c, tl, tu, db, dy, dz) in plans:
    tl.solve(db, dy); tu.solve(dy, dz)
for (c, *_rest) in plans:
    c.sync()
t0 = time.perf_counter()
for _ in range(reps):
    for (c, tl, tu, db, dy, dz) in plans:
        tl.solve(db, dy); tu.solve(dy, dz)
for (c, *_rest) in plans:
    c.sync()
dt = (time.perf_counter() - t0) / reps
print("%d side by side: %.1f us per (L + U) pair, %.2f ms per sweep over all, %.1f GB/s aggregate"
      % (nsub, dt / nsub * 1e6, dt * 1e3, nsub * byts / dt / 1e9))

# (c) the same with one host thread per subdomain (what bench_ras does: ranks are threads), to
# separate the host's graph-launch cost from what the GPU can overlap
from concurrent.futures import ThreadPoolExecutor


def worker(pl):
    c, tl, tu, db, dy, dz = pl
    for _ in range(reps):
        tl.solve(db, dy); tu.solve(dy, dz)
    c.sync()


t0 = time.perf_counter()
with ThreadPoolExecutor(nsub) as ex:
    list(ex.map(worker, plans))
dt = (time.perf_counter() - t0) / reps
print("%d side by side, one host thread each: %.1f us per (L + U) pair, %.2f ms per sweep, %.1f GB/s aggregate"
      % (nsub, dt / nsub * 1e6, dt * 1e3, nsub * byts / dt / 1e9))
t0 = time.perf_counter()
for _ in range(reps):
    tl.solve(db, dy); tu.solve(dy, dz)
host = (time.perf_counter() - t0) / reps
c.sync()
print("host time to enqueue one (L + U) pair: %.1f us (%d levels)" % (host * 1e6, tl.levels()))
