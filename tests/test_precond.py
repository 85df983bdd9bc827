"""Local preconditioners of the product (schwarz-lib_b200/csrc/precond.cu) against the oracle.

CPU part (no GPU): what the host generation produces - block pointers, inverse blocks, ILU(0)
factors, ISAI approximate inverses - is bit-identical to the oracle restatement, which
tests/test_precond_pinning.py pins to the Ginkgo stand-in of oracle/_ref.
GPU part: one application (bit-exact for block-Jacobi and ISAI, whose row sums run in the same
order; 1e-13 for the triangular solves, whose warp sums do not), preconditioned CG / GMRES with a
fixed budget and to tolerance, and the RAS iteration with --local_precond at every outer
iteration (tolerance 1e-10).
"""
import numpy as np
import pytest

from test_precond_pinning import blockish, unsym

KINDS = ["block-jacobi", "ilu", "isai"]


def _mats(orc, ani4):
    return {"lap12": orc.laplacian2d(12), "lap3d5": orc.laplacian3d(5), "ani4": ani4,
            "unsym": unsym(150, 3), "blockish": blockish(40, 5)}


# ------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("name", ["lap12", "lap3d5", "ani4", "unsym", "blockish"])
def test_host_generation_equals_the_oracle(sz, orc, ani4, name):
    rp, ci, v = _mats(orc, ani4)[name]
    for mbs in (1, 5, 16, 32):
        a = sz.Precond(None, rp, ci, v, "block-jacobi", mbs)
        b = orc.Precond(rp, ci, v, "block-jacobi", mbs)
        assert np.array_equal(a.block_ptrs(), b.block_ptrs())
        assert np.array_equal(a.blocks(), b.blocks())
        a.close()
    for kind, which in (("ilu", (0, 1)), ("isai", (0, 1, 2, 3))):
        a = sz.Precond(None, rp, ci, v, kind)
        b = orc.Precond(rp, ci, v, kind)
        for w in which:
            for x, y in zip(a.csr(w), b.csr(w)):
                assert np.array_equal(x, y), (kind, w)
        a.close()


def test_host_only_handle_and_bad_arguments_fail_loudly(sz, orc):
    rp, ci, v = orc.laplacian2d(6)
    h = sz.Precond(None, rp, ci, v, "isai")
    with pytest.raises(sz.SchwzError, match="host-only"):
        h.apply(None, None)
    h.close()
    with pytest.raises(sz.SchwzError, match="max_block_size"):
        sz.Precond(None, rp, ci, v, "block-jacobi", 33)
    with pytest.raises(sz.SchwzError, match="Unsupported preconditioner"):
        sz.PRECOND["bogus"] = 9
        try:
            sz.Precond(None, rp, ci, v, "bogus")
        finally:
            del sz.PRECOND["bogus"]


# ------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_apply_matches_oracle(gpu, sz, orc, ani4, kind):
    rng = np.random.default_rng(11)
    mats = _mats(orc, ani4)
    mats["lap70"] = orc.laplacian2d(70)          # several CTA tiles, blocks straddling tile edges
    for name, (rp, ci, v) in mats.items():
        n = len(rp) - 1
        for mbs in ((3, 16, 32) if kind == "block-jacobi" else (16,)):
            M = sz.Precond(gpu, rp, ci, v, kind, mbs)
            Mo = orc.Precond(rp, ci, v, kind, mbs)
            r = rng.standard_normal(n)
            dr, dz, dd = gpu.to_device(r), gpu.zeros(n), gpu.zeros(1)
            M.apply(dr, dz, dd)
            z = gpu.to_host(dz, n)
            zo = Mo.apply(r)
            if kind == "ilu":
                np.testing.assert_allclose(z, zo, rtol=1e-12, atol=1e-13 * abs(zo).max())
            else:
                assert np.array_equal(z, zo), (name, mbs)
            assert gpu.to_host(dd, 1)[0] == pytest.approx(float(r @ zo), rel=1e-12, abs=1e-12)
            M.apply(dr, dz)                      # without the fused dot
            assert np.array_equal(gpu.to_host(dz, n), z)
            assert M.bytes_per_apply() > 16 * n
            for p in (dr, dz, dd):
                gpu.free(p)
            M.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_preconditioned_cg_matches_oracle(gpu, sz, orc, kind):
    rng = np.random.default_rng(12)
    for nlap in (30, 125):                       # 125^2 rows: the multi-kernel path either way
        rp, ci, v = orc.laplacian2d(nlap)
        n = len(rp) - 1
        A = sz.Csr(gpu, rp, ci, v)
        M = sz.Precond(gpu, rp, ci, v, kind, 16)
        Mo = orc.Precond(rp, ci, v, kind, 16)
        cg = sz.Cg(gpu, A, precond=M)
        b, x0 = rng.standard_normal(n), rng.standard_normal(n) * 0.1
        db = gpu.to_device(b)
        for K in (0, 1, 6, 30):
            dx = gpu.to_device(x0)
            cg.solve(db, dx, K, 1e-300)
            it, rn, r0 = cg.result()
            xo, ito = orc.cg(rp, ci, v, b, x0, K, 1e-300, precond=Mo)
            assert it == ito == K
            np.testing.assert_allclose(gpu.to_host(dx, n), xo, rtol=1e-10, atol=1e-12)
            gpu.free(dx)
        dx = gpu.to_device(x0)
        cg.solve(db, dx, n, 1e-12)
        it, rn, r0 = cg.result()
        xo, ito = orc.cg(rp, ci, v, b, x0, n, 1e-12, precond=Mo)
        _, it_plain = orc.cg(rp, ci, v, b, x0, n, 1e-12)
        assert abs(it - ito) <= 1 and it < it_plain
        assert rn < 1e-12 * r0
        got = gpu.to_host(dx, n)
        assert np.linalg.norm(got - xo) / np.linalg.norm(xo) < 1e-10
        gpu.free(dx); gpu.free(db); cg.close(); M.close(); A.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_preconditioned_gmres_matches_oracle(gpu, sz, orc, ani4, kind):
    rng = np.random.default_rng(13)
    rp, ci, v = ani4
    n = len(rp) - 1
    A = sz.Csr(gpu, rp, ci, v)
    M = sz.Precond(gpu, rp, ci, v, kind, 16)
    Mo = orc.Precond(rp, ci, v, kind, 16)
    b = rng.standard_normal(n)
    db = gpu.to_device(b)
    for m, K in ((1, 4), (5, 12), (30, 30), (30, 70)):
        g = sz.Gmres(gpu, A, m, precond=M)
        dx = gpu.zeros(n)
        g.solve(db, dx, K, 1e-300)
        it, rn, r0 = g.result()
        xo, ito = orc.gmres(rp, ci, v, b, np.zeros(n), K, 1e-300, m, precond=Mo)
        assert it == ito == K
        np.testing.assert_allclose(gpu.to_host(dx, n), xo, rtol=1e-8, atol=1e-10)
        gpu.free(dx); g.close()
    g = sz.Gmres(gpu, A, 30, precond=M)
    dx = gpu.zeros(n)
    g.solve(db, dx, 3000, 1e-10)
    it, rn, r0 = g.result()
    xo, ito = orc.gmres(rp, ci, v, b, np.zeros(n), 3000, 1e-10, 30, precond=Mo)
    _, it_plain = orc.gmres(rp, ci, v, b, np.zeros(n), 3000, 1e-10, 30)
    assert abs(it - ito) <= 2 and it < it_plain
    got = gpu.to_host(dx, n)
    assert np.linalg.norm(got - xo) / np.linalg.norm(xo) < 1e-8
    gpu.free(dx); gpu.free(db); g.close(); M.close(); A.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("solver", ["cg", "gmres"])
def test_ras_with_local_preconditioner_matches_oracle_every_iteration(sz, orc, ani4, kind, solver):
    """--local_precond through the RAS object: iterates within 1e-10 of the oracle (itself equal
    to the reference run, tests/test_precond_pinning.py) at every outer iteration."""
    from test_gpu_ras import _fresh_ctxs, _gpu_x_global, _make, _manual_step
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    P = 4
    if solver == "cg":
        mat, msetup, kw = orc.laplacian2d(32), ("laplacian2d", 32), {}
    else:
        mat, msetup, kw = ani4, ani4, dict(non_symmetric=True, restart_iter=30)
    N = len(mat[0]) - 1
    ob = orc.Problem(*mat, P)
    ob.configure(tolerance=1e-8, local_tol=1e-12, max_iters=1000, enable_global_check=True,
                 local_precond=kind, precond_max_block_size=8, **kw)
    setup = sz.Setup(msetup, P)
    ctxs = _fresh_ctxs(sz, P)
    subs = _make(sz, ctxs, setup, P, local_tol=1e-12, local_precond=kind,
                 precond_max_block_size=8, **kw)
    for it in range(8):
        norms = _manual_step(subs, it, P)
        ob.step()
        for r in range(P):
            assert norms[r] == pytest.approx(ob.status(r)["resnorm"], rel=1e-9)
            l2g = setup.l2g(r)
            xo = ob.x(r)[l2g]
            xg = _gpu_x_global(subs[r], setup, r, N)[l2g]
            assert np.linalg.norm(xg - xo) <= 1e-10 * max(np.linalg.norm(xo), 1e-300), (it, r)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()


@pytest.mark.gpu
def test_full_size_strip_properties(gpu, sz):
    """One interior strip of a 2048^2 / 8 problem (526 336 rows), no oracle run: block-Jacobi
    undoes the block-diagonal part of A (D z = r block by block), ISAI's application is the two
    sequential-row-sum SpMVs (bit-exact against scipy's CSR product), its factors satisfy the
    ILU(0) identity on a row sample, and preconditioned CG beats plain CG at equal budget."""
    import scipy.sparse as sp
    setup = sz.Setup(("laplacian2d", 2048), 8)
    rp, ci, v = setup.local_matrix(1)
    n = len(rp) - 1
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(21)
    r = rng.standard_normal(n)
    dr, dz = gpu.to_device(r), gpu.zeros(n)
    # block-Jacobi, blocks of 16 consecutive rows (no two rows of a 5-pt matrix are alike)
    M = sz.Precond(gpu, rp, ci, v, "block-jacobi", 16)
    bp = M.block_ptrs()
    assert np.array_equal(bp, np.minimum(np.arange(0, n + 16, 16), n)[:len(bp)]) and bp[-1] == n
    M.apply(dr, dz)
    z = gpu.to_host(dz, n)
    rows = np.repeat(np.arange(n), np.diff(rp))
    same_block = (rows // 16) == (ci // 16)
    D = sp.csr_matrix((v[same_block], (rows[same_block], ci[same_block])), shape=(n, n))
    assert np.linalg.norm(D @ z - r) <= 1e-13 * np.linalg.norm(r) * 16
    M.close()
    # ISAI
    M = sz.Precond(gpu, rp, ci, v, "isai")
    L, U = (sp.csr_matrix(M.csr(w)[::-1], shape=(n, n)) for w in (0, 1))
    Li, Ui = (sp.csr_matrix(M.csr(w)[::-1], shape=(n, n)) for w in (2, 3))
    M.apply(dr, dz)
    assert np.array_equal(gpu.to_host(dz, n), Ui @ (Li @ r))
    sample = rng.integers(0, n, 2000)
    LU = (L[sample] @ U).tocsr()
    As = A[sample]
    assert abs(LU.multiply(As != 0) - As).max() < 1e-12        # ILU(0): L U = A on the pattern
    # 30 preconditioned iterations reduce the residual further than 30 plain ones
    Acsr = sz.Csr(gpu, rp, ci, v)
    db = gpu.to_device(np.ones(n))
    red = {}
    for name, pc in (("plain", None), ("isai", M)):
        cg = sz.Cg(gpu, Acsr, precond=pc)
        dx = gpu.zeros(n)
        cg.solve(db, dx, 30, 1e-300)
        it, rn, r0 = cg.result()
        x = gpu.to_host(dx, n)
        assert it == 30 and np.linalg.norm(np.ones(n) - A @ x) == pytest.approx(rn, rel=1e-8)
        red[name] = rn / r0
        gpu.free(dx); cg.close()
    assert red["isai"] < red["plain"]
    for p in (dr, dz, db):
        gpu.free(p)
    M.close(); Acsr.close()
