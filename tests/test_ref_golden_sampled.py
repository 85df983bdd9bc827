"""The bench workload (BASELINE.json configs[1]: strips, CG with local_max_iters = 50,
synchronous exchange, global check) pinned to the reference at 1024^2: tests/golden/refs_*.npz
hold what oracle/_ref (the reference's own sources) computed in its first 13 outer iterations -
every local residual norm, and at the recorded iterations the norm of each subdomain's iterate
plus its values at a fixed sample of positions (tests/golden/make_ref_golden.py, main_sampled).

* CPU: the oracle restatement reproduces those numbers BIT FOR BIT with one thread per
  subdomain, and within 1e-10 with an OpenMP team (the dot products then sum in another order:
  the measured sensitivity of this truncated, warm-started CG is ~2e-11 per outer iteration and
  does not grow over these iterations).
* GPU: the CUDA path holds the 1e-10 contract of BASELINE.json on the same numbers.
"""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, os.path.join(GOLDEN))
CASE = "cfg2_lap1024_P8_cg50"
N1D, P, ITERS = 1024, 8, 13
TOL_ITERATE = 1e-10


def _golden():
    return np.load(os.path.join(GOLDEN, "refs_%s.npz" % CASE))


def _positions(g, r):
    own_stride, ext_stride = (int(v) for v in g["strides"])
    ls = int(g["sizes_%d" % r][0])
    nk = int(g["n_known_%d" % r][0])
    return np.concatenate([np.arange(0, ls, own_stride), np.arange(ls, nk, ext_stride)])


def test_oracle_reproduces_the_reference_bit_for_bit(orc):
    g = _golden()
    ob = orc.Problem(*orc.laplacian2d(N1D), P)
    ob.configure(tolerance=1e-30, local_tol=1e-12, local_max_iters=50, max_iters=100,
                 enable_global_check=True)
    orc.set_threads(1)
    orc.set_rank_threads(min(P, len(os.sched_getaffinity(0))))
    try:
        snap = set(int(k) for k in g["snap"])
        for it in range(ITERS):
            ob.step()
            # ob.x after step `it` holds the solve of `it`; the reference's snapshot k is taken
            # right after the exchange of iteration k, i.e. the oracle's state after step k - 1
            # plus that exchange - compared through the residual history instead, and through
            # the iterates on the GPU side where the stages are stepped one by one
        for r in range(P):
            assert np.array_equal(ob.history(r)[0][:ITERS], g["local_res_%d" % r][:ITERS])
            s = ob.sizes(r)
            assert [s[k] for k in ("local_size", "local_size_x", "overlap_size", "nnz_local",
                                   "nnz_interface")] == g["sizes_%d" % r].tolist()
            l2g = ob.l2g(r)
            assert len(l2g) == int(g["n_known_%d" % r][0])
            assert np.array_equal(l2g[_positions(g, r)], g["l2g_sample_%d" % r])
            chk = [int(l2g.astype(np.int64).sum()),
                   int((l2g.astype(np.int64) * (np.arange(len(l2g)) % 1009)).sum())]
            assert chk == g["l2g_checksum_%d" % r].tolist()
    finally:
        orc.set_rank_threads(1)


def test_product_index_sets_match_the_sampled_reference(sz):
    g = _golden()
    setup = sz.Setup(("laplacian2d", N1D), P)
    assert np.array_equal(setup.first_row(), g["first_row"])
    for r in range(P):
        s = setup.sizes(r)
        assert [s[k] for k in ("local_size", "local_size_x", "overlap_size", "nnz_local",
                               "nnz_interface")] == g["sizes_%d" % r].tolist()
        l2g = setup.l2g(r)
        assert np.array_equal(l2g[_positions(g, r)], g["l2g_sample_%d" % r])
        chk = [int(l2g.astype(np.int64).sum()),
               int((l2g.astype(np.int64) * (np.arange(len(l2g)) % 1009)).sum())]
        assert chk == g["l2g_checksum_%d" % r].tolist()
        nin, nout = setup.neighbors(r)
        assert np.array_equal(nin, g["nbr_in_%d" % r]) and np.array_equal(nout, g["nbr_out_%d" % r])


@pytest.mark.gpu
def test_cuda_path_holds_1e10_on_the_bench_workload(sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    g = _golden()
    setup = sz.Setup(("laplacian2d", N1D), P)
    ctxs = [sz.Context(0) for _ in range(P)]
    subs = [sz.Ras(ctxs[r], setup, r, local_tol=1e-12, local_max_iters=50) for r in range(P)]
    sz.connect_local(subs, setup)
    snap = set(int(k) for k in g["snap"])
    pos = [_positions(g, r) for r in range(P)]
    worst_x = worst_r = 0.0
    for it in range(ITERS):
        for s in subs:
            s.exchange_push(it)
        for s in subs:
            nin, _ = s.neighbors()
            for p in nin:
                s.wait_push_of(subs[int(p)])
            s.exchange_unpack(it)
        if it in snap:
            for r in range(P):
                x = subs[r].x()
                want_norm = float(g["xnorm_%d_%d" % (r, it)][0])
                ex = np.linalg.norm(x[pos[r]] - g["x_%d_%d" % (r, it)]) / np.linalg.norm(
                    g["x_%d_%d" % (r, it)])
                en = abs(np.linalg.norm(x) - want_norm) / want_norm
                worst_x = max(worst_x, ex, en)
                assert ex <= TOL_ITERATE and en <= TOL_ITERATE, (it, r, ex, en)
        for s in subs:
            s.update_boundary()
            s.local_residual()
        norms = [s.residual_norm() for s in subs]
        for r in range(P):
            ref = g["local_res_%d" % r]
            er = abs(norms[r] - ref[it]) / ref[0]
            worst_r = max(worst_r, er)
            assert er <= 1e-10, (it, r, er)
        for s in subs:
            s.local_solve()
            s.restrict()
    print("worst relative deviation: iterates %.2e, residual norms %.2e" % (worst_x, worst_r))
    # the run-ahead loop computes the same residual history
    for s in subs:
        s.reset()
    out = sz.ras_run(subs, P, ITERS, tolerance=1e-30, enable_global_check=True, history=True)
    for r in range(P):
        ref = g["local_res_%d" % r]
        assert np.max(np.abs(out["history"][:, r] - ref[:ITERS])) <= 1e-10 * ref[0]
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()
