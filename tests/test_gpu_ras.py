"""Parity of the RAS iteration on the GPU against the CPU oracle, through the
C ABI.  Stated tolerance (BASELINE.json north_star / SURVEY.md 8c): in
synchronous mode ||x_k(gpu) - x_k(oracle)|| / ||x_k(oracle)|| <= 1e-10 at every
outer iteration and the same outer-iteration count; in asynchronous mode the
final true relative residual meets the oracle's threshold."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
E = json.load(open(os.path.join(GOLDEN, "appendix_e.json")))
TOL_ITERATE = 1e-10


@pytest.fixture(autouse=True)
def _need_gpu(sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")


def _make(sz, ctxs, setup, P, **kw):
    subs = [sz.Ras(ctxs[r % len(ctxs)], setup, r, **kw) for r in range(P)]
    sz.connect_local(subs, setup)
    return subs


def _manual_step(subs, it, P):
    """One pass of the loop body, stage by stage (the calls SolverRAS makes)."""
    if P > 1:
        for s in subs:
            s.exchange_push(it)
        for s in subs:
            nin, _ = s.neighbors()
            for p in nin:
                s.wait_push_of(subs[int(p)])
            s.exchange_unpack(it)
    for s in subs:
        s.update_boundary()
        s.local_residual()
    norms = [s.residual_norm() for s in subs]
    for s in subs:
        s.local_solve()
        s.restrict()
    return norms


def _gpu_x_global(sub, setup, r, N):
    """scatter the compact device vector back to the oracle's length-N layout"""
    l2g = setup.l2g(r)
    out = np.zeros(N)
    out[l2g] = sub.x()
    return out


def _fresh_ctxs(sz, n):
    return [sz.Context(0) for _ in range(n)]


@pytest.mark.parametrize("case", ["strips2", "strips4", "regular2d4"])
def test_sync_iterates_match_oracle_every_iteration(sz, orc, case):
    n, P, part = {"strips2": (40, 2, None), "strips4": (32, 4, None),
                  "regular2d4": (24, 4, "2d")}[case]
    mat = orc.laplacian2d(n)
    N = n * n
    pv = orc.partition_regular2d(N, P) if part else None
    ob = orc.Problem(*mat, P, part=pv)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=1000, enable_global_check=True)
    setup = sz.Setup(("laplacian2d", n), P, part=pv)
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P, local_tol=1e-12)
    for it in range(12):
        norms = _manual_step(subs, it, P)
        ob.step()
        for r in range(P):
            st = ob.status(r)
            assert norms[r] == pytest.approx(st["resnorm"], rel=1e-10)
            xo = ob.x(r)
            xg = _gpu_x_global(subs[r], setup, r, N)
            # compare where the subdomain holds data (own + overlap + halo)
            l2g = setup.l2g(r)
            denom = np.linalg.norm(xo[l2g])
            assert np.linalg.norm(xg[l2g] - xo[l2g]) <= TOL_ITERATE * max(denom, 1e-300), (it, r)
            np.testing.assert_allclose(subs[r].local_solution(), ob.local_solution(r),
                                       rtol=0, atol=TOL_ITERATE * max(denom, 1.0))
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def test_cfg1_outer_loop_matches_appendix_e(sz, orc):
    """BASELINE.json configs[0]: 100x100, 2 subdomains, CG local solve, sync."""
    g = E["cfg1"]
    setup = sz.Setup(("laplacian2d", 100), 2)
    ctxs = _fresh_ctxs(sz, 2)
    subs = _make(sz, ctxs, setup, 2, local_tol=1e-12)
    out = sz.ras_run(subs, 2, 300, tolerance=1e-6, enable_global_check=True, history=True)
    assert out["converged"] and out["iters"] == g["stop_iter"]
    h = out["history"]
    assert h[0, 0] == pytest.approx(g["rho0_local"], rel=1e-13)
    gsum = h.sum(axis=1)
    for k, want in g["ratios"].items():
        assert gsum[int(k)] / gsum[0] == pytest.approx(want, rel=2e-7)
    fix = json.load(open(os.path.join(GOLDEN, "cfg1_history.json")))
    # residual norms inherit the absolute error of the iterates (1e-10 * ||A|| ||x||
    # is the contract; the observed difference is ~1e-12)
    np.testing.assert_allclose(gsum, fix["global_resnorm"], rtol=1e-9, atol=1e-9)
    # solution: own parts, and the distributed true residual
    x = np.concatenate([s.x()[:s.local_size] for s in subs])
    assert np.linalg.norm(x) == pytest.approx(g["sol_norm"], rel=1e-9)
    assert x[0] == pytest.approx(g["x0"], rel=1e-9)
    assert x[5050] == pytest.approx(g["x5050"], rel=1e-9)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def test_regular2d_and_strip_iteration_counts(sz, orc):
    g = E["cfg1"]["outer_iters_exact_local"]
    for n, P, part, want in ((16, 4, None, g["lap16_P4_strips"]),
                             (64, 4, "2d", g["lap64_P4_regular2d"])):
        pv = sz.partition_regular2d(n * n, P) if part else None
        setup = sz.Setup(("laplacian2d", n), P, part=pv)
        ctxs = _fresh_ctxs(sz, P)
        subs = _make(sz, ctxs, setup, P)
        out = sz.ras_run(subs, P, 400, enable_global_check=True)
        assert out["converged"] and out["iters"] == want
        for s in subs:
            s.close()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("P", [2, 4, 8])
def test_cfg3_ani4_metis_gmres(sz, orc, ani4, P):
    """BASELINE.json configs[2]: ani4_crop, METIS partition, GMRES local solve."""
    part = sz.partition_metis(ani4[0], ani4[1], P)
    ob = orc.Problem(*ani4, P, part=part)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=800, enable_global_check=True,
                 non_symmetric=True, restart_iter=30)
    iters_o = ob.run()
    setup = sz.Setup(ani4, P, part=part)
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P, local_tol=1e-12, non_symmetric=True, restart_iter=30)
    out = sz.ras_run(subs, P, 800, tolerance=1e-6, enable_global_check=True, history=True)
    _, gres = ob.history(0)
    hs = out["history"].sum(axis=1)
    assert out["converged"] and out["iters"] == iters_o, (out["iters"], iters_o, hs[-3:], gres[-3:])
    np.testing.assert_allclose(hs, gres, rtol=1e-8, atol=1e-9)
    xo, fr = ob.final_residual()
    fr_first = setup.first_row()
    for r in range(P):
        own = subs[r].x()[:subs[r].local_size]
        ref = xo[fr_first[r]:fr_first[r + 1]]
        assert np.linalg.norm(own - ref) <= 1e-9 * np.linalg.norm(xo)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def test_direct_local_solve_matches_oracle(sz, orc):
    """BASELINE.json configs[4] in miniature: factorised local solve, regular2d,
    more subdomains than GPUs."""
    n, P = 32, 16
    pv = sz.partition_regular2d(n * n, P)
    setup = sz.Setup(("laplacian2d", n), P, part=pv)
    perms = []
    ctxs = _fresh_ctxs(sz, 4)
    subs = _make(sz, ctxs, setup, P, local_solver="direct-ginkgo")
    for r in range(P):
        rp, ci, v = setup.local_matrix(r)
        perm = sz.nd_ordering(rp, ci)
        perms.append(perm)
        subs[r].set_factors(*sz.host_cholesky(rp, ci, v, perm), perm)
    ob = orc.Problem(*orc.laplacian2d(n), P, part=pv)
    ob.configure(max_iters=600, enable_global_check=True, local_solver="direct-ginkgo",
                 factor_perms=perms)
    for it in range(5):
        norms = _manual_step(subs, it, P)
        ob.step()
        for r in range(P):
            assert norms[r] == pytest.approx(ob.status(r)["resnorm"], rel=1e-10, abs=1e-13)
            np.testing.assert_allclose(subs[r].local_solution(), ob.local_solution(r),
                                       rtol=1e-10, atol=1e-12)
    for s in subs:
        s.close()
    ctxs2 = _fresh_ctxs(sz, 4)
    subs = _make(sz, ctxs2, setup, P, local_solver="direct-ginkgo")
    for r in range(P):
        rp, ci, v = setup.local_matrix(r)
        subs[r].set_factors(*sz.host_cholesky(rp, ci, v, perms[r]), perms[r])
    ob = orc.Problem(*orc.laplacian2d(n), P, part=pv)
    ob.configure(max_iters=600, enable_global_check=True, local_solver="direct-ginkgo",
                 factor_perms=perms)
    out = sz.ras_run(subs, P, 600, enable_global_check=True)
    assert out["converged"] and out["iters"] == ob.run()
    for s in subs:
        s.close()
    for c in ctxs + ctxs2:
        c.close()


def test_async_onesided_decentralized_meets_threshold(sz, orc):
    """BASELINE.json configs[3] in miniature: 3-D 7-pt Laplacian, slabs, one-sided
    exchange (no waits), decentralised flag convergence."""
    n, P = 12, 4
    mat = orc.laplacian3d(n)
    ob = orc.Problem(*mat, P)
    ob.configure(tolerance=1e-6, max_iters=2000, enable_onesided=True, remote_comm_type="put",
                 global_convergence_type="decentralized")
    ob.run()
    _, fro = ob.final_residual()
    setup = sz.Setup(("laplacian3d", n), P)
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P)
    out = sz.ras_run(subs, P, 2000, tolerance=1e-6, enable_onesided=True, conv_decentralized=True)
    assert out["converged"]
    # distributed true residual after a final exchange
    for s in subs:
        s.exchange_push(0)
    for s in subs:
        s.sync()
    for s in subs:
        s.exchange_unpack(0)
    rsq = sum(s.true_residual_sq() for s in subs)
    rel = np.sqrt(rsq) / np.sqrt(n ** 3)
    assert rel <= max(fro["relative"] * 10, 1e-5)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def test_true_residual_matches_host(sz, orc):
    n, P = 20, 2
    mat = orc.laplacian2d(n)
    setup = sz.Setup(mat, P)
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n * n)
    fr = setup.first_row()
    for r in range(P):
        subs[r].set_x_own(x[fr[r]:fr[r + 1]])
    for s in subs:
        s.exchange_push(0)
    for s in subs:
        s.sync()
    for s in subs:
        s.exchange_unpack(0)
    rsq = sum(s.true_residual_sq() for s in subs)
    import scipy.sparse as sp
    A = sp.csr_matrix((mat[2], mat[1], mat[0]), shape=(n * n, n * n))
    want = np.linalg.norm(np.ones(n * n) - A @ x) ** 2
    assert rsq == pytest.approx(want, rel=1e-12)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def test_full_size_cfg2_properties(sz):
    """BASELINE.json configs[1] at full size (8192 x 8192, 8 strips on one GPU):
    size-independent properties instead of an oracle run.
      * iteration 0: x = 0, so every local residual norm is sqrt(local_size_x)
        exactly (rhs = ones) — a checksum of the index sets and of the SpMV;
      * after an exchange, every overlap/halo slot of a subdomain equals the
        owner's value (checksum of checksums over the halo lists);
      * the residual norms of a second run from the same state are
        bit-identical (determinism of the fused reductions)."""
    n, P = 8192, 8
    setup = sz.Setup(("laplacian2d", n), P)
    ctxs = _fresh_ctxs(sz, P)
    subs = []
    for r in range(P):
        subs.append(sz.Ras(ctxs[r], setup, r, local_max_iters=5))
        setup.release(r)
    sz.connect_local(subs, setup)
    out = sz.ras_run(subs, P, 3, tolerance=1e-6, enable_global_check=True, history=True)
    h = out["history"] if out["history"].shape[0] == 3 else None
    assert h is not None
    for r in range(P):
        assert h[0, r] == np.sqrt(float(subs[r].local_size_x))
    assert np.all(h[1] > 0) and np.all(np.isfinite(h))
    # halo consistency after one more exchange
    sz.refresh_halo(subs, P)
    fr = setup.first_row()
    xs = [s.x() for s in subs]
    for r in (0, 3, 7):
        l2g = setup.l2g(r)
        ls = subs[r].local_size
        ext = l2g[ls:]
        owner = np.searchsorted(fr, ext, side="right") - 1
        want = np.empty(len(ext))
        for q in np.unique(owner):
            m = owner == q
            want[m] = xs[q][ext[m] - fr[q]]
        assert np.array_equal(xs[r][ls:], want)
    # determinism: zero state again, same three iterations -> identical norms and iterates
    for s in subs:
        s.reset()
    out2 = sz.ras_run(subs, P, 3, tolerance=1e-6, enable_global_check=True, history=True)
    assert np.array_equal(out2["history"], h)
    assert out2["global_resnorm"] == out["global_resnorm"]
    sz.refresh_halo(subs, P)
    for r in (0, 3, 7):
        assert np.array_equal(subs[r].x(), xs[r])
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


def test_local_iteration_cap_can_be_reset_mid_run(sz):
    """settings.reset_local_crit_iter / metadata.updated_max_iters (source/solve.cpp:721-741): past
    a given outer iteration the local solver's iteration cap is replaced."""
    P = 2
    setup = sz.Setup(("laplacian2d", 200), P)          # 20 200 rows: the multi-kernel CG
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P, local_max_iters=5)
    _manual_step(subs, 0, P)
    assert [s.last_local_iters() for s in subs] == [5, 5]
    for s in subs:
        s.set_local_max_iters(2)
    _manual_step(subs, 1, P)
    assert [s.last_local_iters() for s in subs] == [2, 2]
    for s in subs:
        s.set_local_max_iters(-1)                      # -1: the local size, i.e. to local_tol
    _manual_step(subs, 2, P)
    assert all(5 < s.last_local_iters() < s.local_size_x for s in subs)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("n,seed,P,spd", [(400, 3, 5, True), (260, 6, 6, False)])
def test_random_matrix_iterates_match_oracle(sz, orc, n, seed, P, spd):
    """Random sparse matrices (symmetric pattern, ragged rows, hub rows), METIS partition - the
    inputs on which tests/test_ref_pinning_random.py pins the oracle to the reference run: the
    CUDA path reproduces the oracle's iterates within 1e-10 at every outer iteration."""
    from test_ref_pinning_random import random_matrix
    mat = random_matrix(n, seed, spd)
    part = sz.partition_metis(mat[0], mat[1], P)
    kw = {} if spd else dict(non_symmetric=True, restart_iter=20)
    ob = orc.Problem(*mat, P, part=part)
    ob.configure(tolerance=1e-10, local_tol=1e-12, max_iters=100, enable_global_check=True, **kw)
    setup = sz.Setup(mat, P, part=part)
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P, local_tol=1e-12, **kw)
    for it in range(10):
        norms = _manual_step(subs, it, P)
        ob.step()
        for r in range(P):
            assert norms[r] == pytest.approx(ob.status(r)["resnorm"], rel=1e-9, abs=1e-13)
            l2g = setup.l2g(r)
            xo = ob.x(r)[l2g]
            xg = _gpu_x_global(subs[r], setup, r, n)[l2g]
            assert np.linalg.norm(xg - xo) <= TOL_ITERATE * max(np.linalg.norm(xo), 1e-300), (it, r)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()
