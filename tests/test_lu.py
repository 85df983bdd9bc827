"""The unsymmetric factorised local solve (UMFPACK branch of the reference:
source/solve.cpp:145-171 factorise, :322-385 extract P, Q, L, U, :709-720 + solver_tools.hpp:69-87
solve x = Q U^-1 L^-1 P b with triangular solvers).  UMFPACK is absent here (and cannot run in
oracle/_ref), so the pin is mathematical: P A Q = L U exactly on the product's host factors, and the
RAS iterates agree with the oracle, whose own LU is an independent dense factorisation."""
import re

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from test_precond_pinning import unsym


def _check_factors(sz, rp, ci, v, q):
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    lu = sz.HostLu(rp, ci, v, q)
    L, U, p = lu.factors()
    Lm = sp.csr_matrix(L[::-1], shape=(n, n))
    Um = sp.csr_matrix(U[::-1], shape=(n, n))
    qq = np.arange(n) if q is None else q
    assert sorted(p) == list(range(n))
    assert (Lm.diagonal() == 1.0).all() and sp.triu(Lm, 1).nnz == 0 and sp.tril(Um, -1).nnz == 0
    assert Lm.has_sorted_indices and Um.has_sorted_indices
    assert abs(Lm @ Um - A[p][:, qq]).max() <= 1e-13 * abs(A).max() * max(1.0, abs(Um).max())
    assert abs(Lm).max() <= 1e3 + 1e-9      # threshold pivoting bounds the multipliers
    b = np.random.default_rng(0).standard_normal(n)
    z = spl.spsolve_triangular(Um, spl.spsolve_triangular(Lm, b[p], lower=True), lower=False)
    x = np.zeros(n)
    x[qq] = z
    assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b) * n
    lu.close()
    return p, qq


def test_host_sparse_lu_properties(sz, orc, ani4):
    for rp, ci, v in (ani4, orc.laplacian2d(20), unsym(200, 8)):
        p, q = _check_factors(sz, rp, ci, v, None)
        assert np.array_equal(p, q)                    # dominant diagonals: no off-diagonal pivot
        _check_factors(sz, rp, ci, v, sz.nd_ordering(rp, ci))


def test_host_sparse_lu_pivots_off_the_diagonal_when_it_must(sz):
    """rows shuffled: the diagonal of the input is (mostly) structurally zero or tiny"""
    rp, ci, v = unsym(120, 9)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    shuffle = np.random.default_rng(1).permutation(n)
    B = sp.csr_matrix(A[shuffle])
    B.sort_indices()
    mat = (B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data)
    p, q = _check_factors(sz, *mat, None)
    assert (p != q).sum() > n // 2
    # singular input is reported, not factorised
    S0 = sp.csr_matrix(np.array([[1.0, 2.0], [2.0, 4.0]]))
    with pytest.raises(sz.SchwzError, match="singular"):
        sz.HostLu(S0.indptr.astype(np.int32), S0.indices.astype(np.int32), S0.data)


@pytest.mark.gpu
def test_direct_lu_local_solve_matches_oracle(sz, orc, ani4):
    """ani4_crop (unsymmetric values), 4 subdomains, LU-factorised local solve: local solutions
    and residual norms within 1e-10 of the oracle at every outer iteration, same stopping
    iteration through the loop."""
    from test_gpu_ras import _fresh_ctxs, _make, _manual_step
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    P = 4
    part = sz.partition_metis(ani4[0], ani4[1], P)
    setup = sz.Setup(ani4, P, part=part)

    def build(ctxs):
        subs = _make(sz, ctxs, setup, P, local_solver="direct-ginkgo")
        for r in range(P):
            rp, ci, v = setup.local_matrix(r)
            lu = sz.HostLu(rp, ci, v, sz.nd_ordering(rp, ci))
            subs[r].set_lu_factors(lu)
            lu.close()
        return subs

    def oracle():
        ob = orc.Problem(*ani4, P, part=part)
        ob.configure(max_iters=600, enable_global_check=True, local_solver="direct-ginkgo",
                     local_factorization="umfpack")
        return ob

    ctxs = _fresh_ctxs(sz, P)
    subs, ob = build(ctxs), oracle()
    for it in range(6):
        norms = _manual_step(subs, it, P)
        ob.step()
        for r in range(P):
            assert norms[r] == pytest.approx(ob.status(r)["resnorm"], rel=1e-9, abs=1e-12)
            want = ob.local_solution(r)
            np.testing.assert_allclose(subs[r].local_solution(), want, rtol=0,
                                       atol=1e-10 * np.linalg.norm(want))
    for s in subs:
        s.close()
    ctxs2 = _fresh_ctxs(sz, P)
    subs, ob = build(ctxs2), oracle()
    out = sz.ras_run(subs, P, 600, enable_global_check=True)
    assert out["converged"] and out["iters"] == ob.run()
    for s in subs:
        s.close()
    for c in ctxs + ctxs2:
        c.close()


@pytest.mark.gpu
def test_bench_ras_local_factorization_umfpack(tmp_path, sz, orc, ani4):
    """the drop-in driver: --local_solver=direct-ginkgo --local_factorization=umfpack"""
    import subprocess
    from test_gpu_bench_ras import BIN
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    rp, ci, v = ani4
    n = len(rp) - 1
    path = tmp_path / "ani4_crop.mtx"
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(ci)))
        rows = np.repeat(np.arange(n), np.diff(rp))
        for r, c, x in zip(rows, ci, v):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, x))
    p = subprocess.run([BIN, "--executor=cuda", "--matrix_filename=%s" % path, "--partition=metis",
                        "--local_solver=direct-ginkgo", "--local_factorization=umfpack",
                        "--enable_global_check", "--num_iters=600", "--num_subdomains=4"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    part = sz.partition_metis(rp, ci, 4)
    ob = orc.Problem(rp, ci, v, 4, part=part)
    ob.configure(max_iters=600, enable_global_check=True, local_solver="direct-ginkgo",
                 local_factorization="umfpack")
    assert " Rank 2 converged in %d iterations" % ob.run() in p.stdout
    assert "sparse LU" in p.stdout
    m = re.search(r"relative residual norm of solution ([0-9.eE+-]+)", p.stdout)
    assert m and float(m.group(1)) < 2e-6
