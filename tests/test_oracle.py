"""Independent anchors of the CPU oracle: (i) SURVEY Appendix E known answers (tests/golden/
appendix_e.json), (ii) scipy cross-checks of its local solves, (iii) the fixed
point of the iteration.  The reference itself ships no tests or golden vectors
(TESTING.md:1-2); the pin against the reference's own code is tests/test_ref_pinning*.py and
tests/test_precond_pinning.py (oracle/_ref), these tests are the cross-checks beside it."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import GOLDEN

E = json.load(open(os.path.join(GOLDEN, "appendix_e.json")))


def _check_sizes(pb, key):
    g = E["sizes"][key]
    assert pb.first_row().tolist() == g["first_row"]
    for r, want in enumerate(g.get("ranks", [])):
        s = pb.sizes(r)
        assert [s["local_size"], s["local_size_x"], s["nnz_local"], s["nnz_interface"]] == want
    for r, want in enumerate(g.get("halo_in", [])):
        nin, _ = pb.neighbors(r)
        got = {str(int(p)): len(pb.get_list(r, j)) for j, p in enumerate(nin)}
        assert got == want


def test_laplacian_matches_closed_form(orc):
    for n, nnz in E["laplacian_nnz"].items():
        rp, ci, v = orc.laplacian2d(int(n))
        assert len(ci) == nnz == 5 * int(n) ** 2 - 4 * int(n)
    n = 7
    rp, ci, v = orc.laplacian2d(n)
    A = sp.csr_matrix((v, ci, rp), shape=(n * n, n * n))
    T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(n, n))
    K = sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n))
    assert abs(A - K).max() == 0
    assert all(np.all(np.diff(ci[rp[i]:rp[i + 1]]) > 0) for i in range(n * n))


def test_index_set_sizes_appendix_e(orc, ani4):
    _check_sizes(orc.Problem(*orc.laplacian2d(100), 2), "lap100_P2_regular")
    lap16 = orc.laplacian2d(16)
    _check_sizes(orc.Problem(*lap16, 4, part=orc.partition_regular2d(256, 4)), "lap16_P4_regular2d")
    pb = orc.Problem(*lap16, 8, part=orc.partition_regular2d(256, 8))
    assert pb.first_row().tolist() == E["sizes"]["lap16_P8_regular2d"]["first_row"]
    _check_sizes(orc.Problem(*ani4, 2), "ani4_P2_regular")
    _check_sizes(orc.Problem(*ani4, 4), "ani4_P4_regular")
    pb = orc.Problem(*ani4, 8)
    g = E["sizes"]["ani4_P8_regular"]
    for r, key in ((0, "rank0"), (7, "rank7")):
        s = pb.sizes(r)
        assert [s["local_size"], s["local_size_x"], s["nnz_local"], s["nnz_interface"]] == g[key]
    assert len(pb.get_list(0, 0)) == 123 and len(pb.get_list(7, 0)) == 72


def test_ordered_index_sets_appendix_e(orc):
    g = E["ordered"]["lap8_P4_regular2d"]
    pb = orc.Problem(*orc.laplacian2d(8), 4, part=orc.partition_regular2d(64, 4))
    for r, key in ((0, "rank0"), (3, "rank3")):
        s = pb.sizes(r)
        l2g = pb.l2g(r)
        assert l2g[s["local_size"]:s["local_size_x"]].tolist() == g[key]["overlap_row"]
        assert l2g[s["local_size_x"]:].tolist() == g[key]["halo"]
        nin, _ = pb.neighbors(r)
        assert {str(int(p)): pb.get_list(r, j).tolist() for j, p in enumerate(nin)} == g[key]["get"]
    g = E["ordered"]["lap6_P2_regular_overlap3"]["rank1"]
    pb = orc.Problem(*orc.laplacian2d(6), 2, overlap=3)
    s = pb.sizes(1)
    assert (s["local_size"], s["local_size_x"]) == (g["local_size"], g["local_size_x"])
    l2g = pb.l2g(1)
    assert l2g[18:30].tolist() == g["overlap_row"] and l2g[30:].tolist() == g["halo"]
    assert pb.get_list(1, 0).tolist() == g["get"]["0"]


def test_put_lists_and_displacements_are_consistent(orc, ani4):
    pb = orc.Problem(*ani4, 4)
    for r in range(4):
        nin, nout = pb.neighbors(r)
        pd, gd = pb.displacements(r)
        for j, q in enumerate(nout):
            qin, _ = pb.neighbors(int(q))
            k = qin.tolist().index(r)
            assert np.array_equal(pb.put_list(r, j), pb.get_list(int(q), k))
            # my block inside q's receive buffer starts after q's earlier in-lists
            assert pd[q] == sum(len(pb.get_list(int(q), kk)) for kk in range(k))


def test_cfg1_residual_history_appendix_e(orc):
    g = E["cfg1"]
    pb = orc.Problem(*orc.laplacian2d(100), 2)
    pb.configure(tolerance=1e-6, local_tol=1e-12, max_iters=300, enable_global_check=True)
    iters = pb.run()
    assert iters == g["stop_iter"]
    res, gres = pb.history(0)
    assert res[0] == pytest.approx(g["rho0_local"], rel=1e-14)
    assert gres[0] == pytest.approx(g["g0"], rel=1e-14)
    for k, want in g["ratios"].items():
        # CG at local_tol 1e-12 reproduces the exact-local-solve history to ~1e-8
        assert gres[int(k)] / gres[0] == pytest.approx(want, rel=2e-7)
    x, fr = pb.final_residual()
    assert fr["relative"] == pytest.approx(g["final_relative_residual"], rel=1e-3)
    assert fr["sol_norm"] == pytest.approx(g["sol_norm"], rel=1e-9)
    assert x[0] == pytest.approx(g["x0"], rel=1e-9)
    assert x[5050] == pytest.approx(g["x5050"], rel=1e-9)
    fix = json.load(open(os.path.join(GOLDEN, "cfg1_history.json")))
    assert iters == fix["iters"]
    np.testing.assert_allclose(gres, fix["global_resnorm"], rtol=1e-12)


def test_outer_iteration_counts_appendix_e(orc):
    g = E["cfg1"]["outer_iters_exact_local"]
    pb = orc.Problem(*orc.laplacian2d(16), 4)
    pb.configure(max_iters=400, enable_global_check=True)
    assert pb.run() == g["lap16_P4_strips"]
    pb = orc.Problem(*orc.laplacian2d(64), 4, part=orc.partition_regular2d(64 * 64, 4))
    pb.configure(max_iters=400, enable_global_check=True)
    assert pb.run() == g["lap64_P4_regular2d"]


def test_cg_against_scipy(orc, ani4):
    rp, ci, v = ani4
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    x, it = orc.cg(rp, ci, v, b, np.zeros(n), n, 1e-12)
    xs = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-10
    assert np.linalg.norm(b - A @ x) < 1e-12 * np.linalg.norm(b) * 1.01
    # Iteration(max): exactly K updates are applied
    x5, it5 = orc.cg(rp, ci, v, b, np.zeros(n), 5, 1e-30)
    assert it5 == 5
    # textbook CG, same recurrences
    xr = np.zeros(n); r = b.copy(); p = np.zeros(n); rho_prev = 1.0
    for _ in range(5):
        rho = r @ r
        p = r + (rho / rho_prev) * p
        q = A @ p
        a = rho / (p @ q)
        xr += a * p; r -= a * q; rho_prev = rho
    np.testing.assert_allclose(x5, xr, rtol=1e-12, atol=1e-14)


def test_gmres_against_scipy(orc, ani4):
    rp, ci, v = ani4
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(1)
    b = rng.standard_normal(n)
    x, it = orc.gmres(rp, ci, v, b, np.zeros(n), 2000, 1e-10, 30)
    assert np.linalg.norm(b - A @ x) <= 1.0001e-10 * np.linalg.norm(b) * 10
    xs = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-8
    # one full cycle of GMRES(m) == scipy's gmres with restart m, maxiter 1
    m = 10
    x1, it1 = orc.gmres(rp, ci, v, b, np.zeros(n), m, 1e-30, m)
    xs1, _ = spla.gmres(A, b, x0=np.zeros(n), restart=m, maxiter=1, rtol=1e-30, atol=0.0)
    assert np.linalg.norm(x1 - xs1) / np.linalg.norm(xs1) < 1e-10


def test_cholesky_and_trs_against_scipy(orc):
    rp, ci, v = orc.laplacian2d(12)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(2)
    perm = rng.permutation(n).astype(np.int32)
    Lrp, Lci, Lv = orc.cholesky(rp, ci, v, perm)
    L = sp.csr_matrix((Lv, Lci, Lrp), shape=(n, n))
    B = A[perm][:, perm]
    assert abs(L @ L.T - B).max() < 1e-12
    b = rng.standard_normal(n)
    y = orc.trs(Lrp, Lci, Lv, b[perm], upper=False)
    U = sp.csr_matrix(L.T)
    U.sort_indices()
    z = orc.trs(U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data, y, upper=True)
    x = np.zeros(n)
    x[perm] = z
    np.testing.assert_allclose(x, spla.spsolve(A.tocsc(), b), rtol=1e-10)


def test_direct_local_solve_equals_exact_fixed_point(orc):
    pb = orc.Problem(*orc.laplacian2d(16), 4)
    pb.configure(max_iters=400, enable_global_check=True, local_solver="direct-ginkgo")
    assert pb.run() == E["cfg1"]["outer_iters_exact_local"]["lap16_P4_strips"]
    x, fr = pb.final_residual()
    assert fr["relative"] < 2e-6


def test_onesided_put_and_get_converge(orc):
    for kind in ("put", "get"):
        for one in (False, True):
            pb = orc.Problem(*orc.laplacian2d(16), 4)
            pb.configure(max_iters=500, enable_onesided=True, remote_comm_type=kind,
                         enable_one_by_one=one, global_convergence_type="decentralized")
            pb.run()
            assert all(pb.status(r)["finished"] == 1.0 for r in range(4))
            x, fr = pb.final_residual()
            assert fr["relative"] < 1e-4


def test_twosided_without_global_check_never_converges(orc):
    pb = orc.Problem(*orc.laplacian2d(8), 2)
    pb.configure(max_iters=60, enable_global_check=False)
    assert pb.run() == 60


def test_rank_parallel_stepping_is_bit_identical(orc):
    """bench.py's CPU arm steps the subdomains side by side (orc.set_rank_threads): every rank's
    arithmetic is the same sequence as in the serial sweep, so iterates and residual histories
    are identical doubles"""
    import numpy as np
    n, P = 96, 4
    mat = orc.laplacian2d(n)

    def run(rank_threads):
        ob = orc.Problem(*mat, P)
        ob.configure(tolerance=1e-8, local_tol=1e-12, local_max_iters=20, max_iters=100,
                     enable_global_check=True)
        orc.set_threads(1)
        orc.set_rank_threads(rank_threads)
        try:
            for _ in range(6):
                ob.step()
        finally:
            orc.set_rank_threads(1)
        return [ob.history(r)[0] for r in range(P)], [ob.x(r) for r in range(P)]

    h1, x1 = run(1)
    h4, x4 = run(4)
    for r in range(P):
        assert np.array_equal(h1[r], h4[r]) and np.array_equal(x1[r], x4[r])
