"""Local preconditioners of the iterative local solve (the reference's call sites are
source/solve.cpp:486-652): block-Jacobi, ParILU + triangular solves, ISAI.

Two things are pinned here, on the CPU:
 1. the oracle restatement (oracle/schwz_oracle.cpp) against the Ginkgo stand-in that the
    reference's own solve.cpp is linked with in oracle/_ref - block pointers, inverse blocks, L / U
    factors, approximate inverses, one application and whole preconditioned CG / GMRES solves,
    all bit for bit, and then full reference runs (SolverRAS::run with --local_precond=...);
 2. the mathematics, independently: inverse blocks times blocks = I, L U = A on the pattern of A
    (ILU(0)), (M T = I) on the pattern of T (ISAI), scipy's spilu-free cross-checks.
The arithmetic itself lives in upstream Ginkgo (not in the tree): parity with upstream's bits is
unpinned, see DESIGN.md section 5.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from test_ref_pinning import ref, same_history, same_setup, write_mtx  # noqa: F401

KINDS = ["block-jacobi", "ilu", "isai"]


def unsym(n, seed):
    """a diagonally dominant, structurally unsymmetric matrix with ragged rows"""
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=0.06, random_state=seed, format="csr")
    A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0 + rng.random(n))
    A = sp.csr_matrix(A)
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def blockish(nb, seed):
    """rows with repeated column patterns (natural blocks of sizes 1..5) for find_blocks"""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 6, nb)
    n = int(sizes.sum())
    start = np.concatenate([[0], np.cumsum(sizes)])
    rows = []
    for b in range(nb):
        cols = set(range(start[b], start[b + 1]))
        for o in rng.integers(0, nb, 2):
            cols |= set(range(start[o], start[o + 1]))
        cols = sorted(cols)
        for i in range(start[b], start[b + 1]):
            vals = rng.standard_normal(len(cols))
            vals[cols.index(i)] = 8.0 + len(cols)
            rows.append((cols, vals))
    rp = np.zeros(n + 1, np.int32)
    ci, v = [], []
    for i, (c, x) in enumerate(rows):
        ci += c
        v += list(x)
        rp[i + 1] = len(ci)
    return rp, np.array(ci, np.int32), np.array(v)


def matrices(orc, ani4):
    return {"lap12": orc.laplacian2d(12), "lap3d5": orc.laplacian3d(5), "ani4": ani4,
            "unsym": unsym(150, 3), "blockish": blockish(40, 5)}


@pytest.mark.parametrize("name", ["lap12", "lap3d5", "ani4", "unsym", "blockish"])
@pytest.mark.parametrize("mbs", [1, 4, 16, 32])
def test_block_jacobi_matches_the_reference_stand_in(ref, orc, ani4, name, mbs):
    rp, ci, v = matrices(orc, ani4)[name]
    a = ref.Precond(rp, ci, v, "block-jacobi", mbs)
    b = orc.Precond(rp, ci, v, "block-jacobi", mbs)
    bp = b.block_ptrs()
    assert np.array_equal(a.block_ptrs(), bp)
    assert bp[0] == 0 and bp[-1] == len(rp) - 1 and (np.diff(bp) <= mbs).all() and (np.diff(bp) > 0).all()
    assert np.array_equal(a.blocks(), b.blocks())
    r = np.random.default_rng(1).standard_normal(len(rp) - 1)
    assert np.array_equal(a.apply(r), b.apply(r))
    # independent: every stored block is the inverse of the diagonal block
    A = sp.csr_matrix((v, ci, rp))
    blk, off = b.blocks(), 0
    for k in range(len(bp) - 1):
        bs = bp[k + 1] - bp[k]
        inv = blk[off:off + bs * bs].reshape(bs, bs).T
        D = A[bp[k]:bp[k + 1], bp[k]:bp[k + 1]].toarray()
        assert abs(inv @ D - np.eye(bs)).max() < 1e-12
        off += bs * bs


def test_block_detection_groups_equal_patterns(orc):
    rp, ci, v = blockish(40, 5)
    bp1 = orc.Precond(rp, ci, v, "block-jacobi", 32).block_ptrs()
    # rows of one natural block never end up in different blocks unless the cap splits them
    pat = [tuple(ci[rp[i]:rp[i + 1]]) for i in range(len(rp) - 1)]
    owner = np.searchsorted(bp1, np.arange(len(pat)), side="right") - 1
    for i in range(1, len(pat)):
        if pat[i] == pat[i - 1]:
            assert owner[i] == owner[i - 1]
    # the 5-pt Laplacian has no two equal rows: blocks of exactly max_block_size rows
    rp, ci, v = orc.laplacian2d(10)
    assert np.array_equal(orc.Precond(rp, ci, v, "block-jacobi", 16).block_ptrs(),
                          [0, 16, 32, 48, 64, 80, 96, 100])


@pytest.mark.parametrize("name", ["lap12", "lap3d5", "ani4", "unsym"])
def test_par_ilu_and_isai_match_the_reference_stand_in(ref, orc, ani4, name):
    rp, ci, v = matrices(orc, ani4)[name]
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    a, b = ref.Precond(rp, ci, v, "ilu"), orc.Precond(rp, ci, v, "ilu")
    for which in (0, 1):
        for x, y in zip(a.csr(which), b.csr(which)):
            assert np.array_equal(x, y)
    L = sp.csr_matrix(b.csr(0)[::-1], shape=(n, n))
    U = sp.csr_matrix(b.csr(1)[::-1], shape=(n, n))
    assert (L.diagonal() == 1.0).all() and sp.triu(L, 1).nnz == 0 and sp.tril(U, -1).nnz == 0
    pat = A.toarray() != 0
    assert abs(((L @ U).toarray() - A.toarray())[pat]).max() < 1e-12 * abs(v).max()   # ILU(0)
    r = np.random.default_rng(2).standard_normal(n)
    assert np.array_equal(a.apply(r), b.apply(r))
    z = spl.spsolve_triangular(U, spl.spsolve_triangular(L, r, lower=True), lower=False)
    np.testing.assert_allclose(b.apply(r), z, rtol=1e-10, atol=1e-12)

    a, b = ref.Precond(rp, ci, v, "isai"), orc.Precond(rp, ci, v, "isai")
    for which in (2, 3):
        for x, y in zip(a.csr(which), b.csr(which)):
            assert np.array_equal(x, y)
    Li = sp.csr_matrix(b.csr(2)[::-1], shape=(n, n))
    Ui = sp.csr_matrix(b.csr(3)[::-1], shape=(n, n))
    assert abs(((Li @ L).toarray() - np.eye(n))[L.toarray() != 0]).max() < 1e-12
    assert abs(((Ui @ U).toarray() - np.eye(n))[U.toarray() != 0]).max() < 1e-12
    assert np.array_equal(a.apply(r), b.apply(r))
    np.testing.assert_array_equal(b.apply(r), Ui @ (Li @ r))   # two SpMVs, sequential row sums


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("solver", ["cg", "gmres"])
def test_preconditioned_krylov_matches_the_reference_stand_in(ref, orc, ani4, kind, solver):
    orc.set_threads(1)
    rp, ci, v = orc.laplacian2d(14) if solver == "cg" else ani4
    n = len(rp) - 1
    rng = np.random.default_rng(7)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    pa, pb = ref.Precond(rp, ci, v, kind, 8), orc.Precond(rp, ci, v, kind, 8)
    for budget in (1, 7, 45, 1500):
        if solver == "cg":
            xa = ref.krylov_solve(rp, ci, v, b, x0, budget, 1e-10, precond=pa)
            xb, it = orc.cg(rp, ci, v, b, x0, budget, 1e-10, precond=pb)
        else:
            xa = ref.krylov_solve(rp, ci, v, b, x0, budget, 1e-10, gmres=True, restart=30,
                                  precond=pa)
            xb, it = orc.gmres(rp, ci, v, b, x0, budget, 1e-10, 30, precond=pb)
        assert np.array_equal(xa, xb), (budget, it)
    A = sp.csr_matrix((v, ci, rp))
    assert it < 1500 and np.linalg.norm(b - A @ xb) <= 2e-10 * np.linalg.norm(b - A @ x0)
    # and the preconditioner pays: fewer iterations than the plain solver
    _, it0 = (orc.cg(rp, ci, v, b, x0, 1500, 1e-10) if solver == "cg"
              else orc.gmres(rp, ci, v, b, x0, 1500, 1e-10, 30))
    assert it < it0


@pytest.mark.parametrize("kind", KINDS)
def test_reference_run_with_local_preconditioner_cg(ref, orc, kind):
    """SolverRAS::run of the reference with --local_precond, truncated local solves (so that the
    preconditioner shapes the iterates), against the oracle: iterates and residual norms bit for
    bit at every outer iteration."""
    orc.set_threads(1)
    kw = dict(max_iters=25, tolerance=1e-8, local_tol=1e-12, local_max_iters=6,
              enable_global_check=True, local_precond=kind, precond_max_block_size=8)
    rr = ref.Run(4, laplacian_n=16, partition="regular", overlap=2, record_iterates=True, **kw)
    ob = orc.Problem(*orc.laplacian2d(16), 4)
    ob.configure(**kw)
    same_setup(rr, ob, 4, False)
    same_history(rr, ob, 4)


@pytest.mark.parametrize("kind", KINDS)
def test_reference_run_with_local_preconditioner_gmres(ref, orc, sz, ani4, tmp_path, kind):
    orc.set_threads(1)
    path = write_mtx(tmp_path / "ani4.mtx", ani4)
    kw = dict(max_iters=20, tolerance=1e-6, local_tol=1e-12, local_max_iters=12,
              non_symmetric=True, restart_iter=5, enable_global_check=True, local_precond=kind,
              precond_max_block_size=16)
    rr = ref.Run(4, matrix_file=path, partition="metis", overlap=2, record_iterates=True, **kw)
    part = sz.partition_metis(ani4[0], ani4[1], 4)
    ob = orc.Problem(*ani4, 4, part=part)
    ob.configure(**kw)
    same_setup(rr, ob, 4, True)
    same_history(rr, ob, 4)
