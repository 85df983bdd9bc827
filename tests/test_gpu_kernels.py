"""Parity of the sm_100a kernels against the CPU oracle, called through the C
ABI.  Integer/index work and the row-sequential SpMV are bit-exact; solver
results are compared at the tolerances stated in each test."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _off(p, nbytes):
    return C.c_void_p(p.value + nbytes)


def _rand_csr(rng, n, m, max_row, long_row=None):
    lens = rng.integers(0, max_row + 1, size=n)
    if long_row is not None:
        lens[long_row[0]] = long_row[1]
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum(lens)
    ci = np.concatenate([np.sort(rng.choice(m, size=l, replace=False)) for l in lens] +
                        [np.zeros(0, np.int64)]).astype(np.int32)
    v = rng.standard_normal(rp[-1])
    return rp, ci, v


def test_spmv_bit_exact_laplacian(gpu, sz, orc):
    rng = np.random.default_rng(0)
    for mat in (orc.laplacian2d(50), orc.laplacian3d(11)):
        rp, ci, v = mat
        n = len(rp) - 1
        A = sz.Csr(gpu, rp, ci, v)
        x = rng.standard_normal(n)
        y0 = rng.standard_normal(n)
        dx = gpu.to_device(x)
        for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.37, -2.5)):
            dy = gpu.to_device(y0)
            A.spmv(dx, dy, alpha, beta)
            got = gpu.to_host(dy, n)
            want = orc.spmv(rp, ci, v, x, alpha, beta, y0)
            assert np.array_equal(got, want), (alpha, beta)
            gpu.free(dy)
        gpu.free(dx)
        A.close()


def test_spmv_bit_exact_ragged_and_edge_cases(gpu, sz, orc):
    rng = np.random.default_rng(1)
    cases = [
        _rand_csr(rng, 1000, 777, 12),                       # ragged, empty rows, rectangular
        _rand_csr(rng, 300, 5000, 40),                        # rows longer than a warp
        _rand_csr(rng, 1, 10, 3),                             # one row
        (np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0)),   # all rows empty
    ]
    for rp, ci, v in cases:
        n = len(rp) - 1
        m = int(ci.max()) + 1 if len(ci) else 4
        A = sz.Csr(gpu, rp, ci, v, ncols=m)
        x = rng.standard_normal(m)
        y0 = rng.standard_normal(n)
        dx = gpu.to_device(x)
        dy = gpu.to_device(y0)
        A.spmv(dx, dy, -1.0, 1.0)
        assert np.array_equal(gpu.to_host(dy, n), orc.spmv(rp, ci, v, x, -1.0, 1.0, y0))
        # beta == 0 must not read y (NaN in y stays out of the result)
        gpu.h2d(dy, np.full(n, np.nan))
        A.spmv(dx, dy, 1.0, 0.0)
        assert np.array_equal(gpu.to_host(dy, n), orc.spmv(rp, ci, v, x, 1.0, 0.0))
        gpu.free(dx); gpu.free(dy); A.close()


@pytest.mark.parametrize("col32", ["0", "1"])
def test_spmv_compact_column_stream_mixed_tiles(sz, orc, monkeypatch, col32):
    """the pipelined SpMV reads a row tile's columns as 16-bit offsets when they span < 2^16 and
    as 32-bit indices otherwise, tile by tile: a banded matrix of 150 000 rows (narrow tiles)
    whose last rows couple to both ends (wide tiles, as the overlap rows of a strip do), bit for
    bit against the oracle with the compact stream on and off, incl. the fused reductions"""
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    monkeypatch.setenv("SCHWZ_B200_SPMV_COL32", col32)
    import scipy.sparse as sp
    rng = np.random.default_rng(9)
    n = 150000
    offs = [-900, -3, -1, 0, 1, 2, 700]
    B = sp.diags([rng.standard_normal(n - abs(o)) for o in offs], offs, shape=(n, n), format="lil")
    for r in range(n - 600, n):                       # wide rows: both ends of the index range
        B[r, r - (n - 600)] = rng.standard_normal()
        B[r, 70000 + (r % 1000)] = rng.standard_normal()
    B = B.tocsr(); B.sort_indices()
    rp, ci, v = B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data
    gpu = sz.Context(0)                               # the switch is read when a context is made
    A = sz.Csr(gpu, rp, ci, v)
    x = rng.standard_normal(n)
    y0 = rng.standard_normal(n)
    dx = gpu.to_device(x)
    orc.set_threads(orc.max_threads())
    try:
        for alpha, beta in ((1.0, 0.0), (-1.0, 1.0)):
            dy = gpu.to_device(y0)
            A.spmv(dx, dy, alpha, beta)
            assert np.array_equal(gpu.to_host(dy, n), orc.spmv(rp, ci, v, x, alpha, beta, y0))
            gpu.free(dy)
    finally:
        orc.set_threads(1)
    gpu.free(dx)
    A.close()
    gpu.close()


def test_spmv_long_row_path(gpu, sz, orc):
    rng = np.random.default_rng(2)
    rp, ci, v = _rand_csr(rng, 40, 9000, 6, long_row=(17, 5000))   # > one CTA tile
    A = sz.Csr(gpu, rp, ci, v, ncols=9000)
    x = rng.standard_normal(9000)
    dx = gpu.to_device(x)
    dy = gpu.zeros(40)
    A.spmv(dx, dy, 1.0, 0.0)
    got = gpu.to_host(dy, 40)
    want = orc.spmv(rp, ci, v, x)
    mask = np.arange(40) != 17
    assert np.array_equal(got[mask], want[mask])
    assert got[17] == pytest.approx(want[17], rel=1e-13)      # tree sum, not sequential
    gpu.free(dx); gpu.free(dy); A.close()


def test_spmv_linearity_large(gpu, sz):
    # size-independent property at a size the oracle is not asked to run
    n = 1024
    S = sz.Setup(("laplacian2d", n), 1)
    rp, ci, v = S.local_matrix(0)
    A = sz.Csr(gpu, rp, ci, v)
    N = n * n
    rng = np.random.default_rng(3)
    x1 = rng.standard_normal(N); x2 = rng.standard_normal(N)
    d1 = gpu.to_device(x1); d2 = gpu.to_device(x2); d3 = gpu.to_device(x1 + 2.0 * x2)
    y1 = gpu.zeros(N); y2 = gpu.zeros(N); y3 = gpu.zeros(N)
    A.spmv(d1, y1); A.spmv(d2, y2); A.spmv(d3, y3)
    h1, h2, h3 = (gpu.to_host(y, N) for y in (y1, y2, y3))
    np.testing.assert_allclose(h3, h1 + 2.0 * h2, rtol=0, atol=1e-12)
    # A * ones = boundary indicator of the 5-pt Laplacian: sum = 4n
    gpu.h2d(d1, np.ones(N)); A.spmv(d1, y1)
    assert gpu.to_host(y1, N).sum() == 4.0 * n
    for p in (d1, d2, d3, y1, y2, y3):
        gpu.free(p)
    A.close()


def test_blas1_and_gather_scatter(gpu, sz):
    rng = np.random.default_rng(4)
    for n in (1, 7, 1000, 300001):
        a = rng.standard_normal(n); b = rng.standard_normal(n)
        da = gpu.to_device(a); db = gpu.to_device(b)
        assert gpu.dot(n, da, db) == pytest.approx(float(a @ b), rel=1e-12, abs=1e-12)
        assert gpu.nrm2(n, da) == pytest.approx(float(np.linalg.norm(a)), rel=1e-13)
        gpu.axpy(n, -0.75, da, db)
        assert np.array_equal(gpu.to_host(db, n), b + (-0.75) * a)
        gpu.free(da); gpu.free(db)
    n, m = 5000, 20000
    idx = rng.permutation(m)[:n].astype(np.int32)
    frm = rng.standard_normal(m); into0 = rng.standard_normal(n)
    di = gpu.to_device(idx); df = gpu.to_device(frm)
    # include/gather.hpp:86-107 (OpenMP path semantics)
    ref = {1: frm[idx], 0: frm[idx] + into0, 2: frm[idx] - into0, 3: (frm[idx] + into0) / 2}
    for op, want in ref.items():
        dt = gpu.to_device(into0)
        gpu.gather(n, di, df, dt, op)
        assert np.array_equal(gpu.to_host(dt, n), want)
        gpu.free(dt)
    # include/scatter.hpp:86-108
    src = rng.standard_normal(n); tgt0 = rng.standard_normal(m)
    ds = gpu.to_device(src)
    for op in (1, 0, 2, 3):
        want = tgt0.copy()
        want[idx] = {1: src, 0: src + tgt0[idx], 2: src - tgt0[idx], 3: (src + tgt0[idx]) / 2}[op]
        dt = gpu.to_device(tgt0)
        gpu.scatter(n, di, ds, dt, op)
        assert np.array_equal(gpu.to_host(dt, m), want)
        gpu.free(dt)
    # permutation: out[i] = in[perm[i]] / out[perm[i]] = in[i]
    perm = rng.permutation(n).astype(np.int32)
    dp = gpu.to_device(perm); dv = gpu.to_device(src); do = gpu.zeros(n)
    gpu.permute(n, dp, False, dv, do)
    assert np.array_equal(gpu.to_host(do, n), src[perm])
    gpu.permute(n, dp, True, dv, do)
    want = np.zeros(n); want[perm] = src
    assert np.array_equal(gpu.to_host(do, n), want)
    for p in (di, df, ds, dp, dv, do):
        gpu.free(p)


def test_cg_matches_oracle(gpu, sz, orc, ani4):
    """40x40 and ani4 run the one-CTA solver (small_solvers.cu); 120x120 (14 400
    rows) is above its shared-memory limit and runs the multi-kernel solver."""
    rng = np.random.default_rng(5)
    for (rp, ci, v) in (orc.laplacian2d(40), ani4, orc.laplacian2d(120)):
        n = len(rp) - 1
        A = sz.Csr(gpu, rp, ci, v)
        cg = sz.Cg(gpu, A)
        b = rng.standard_normal(n)
        x0 = rng.standard_normal(n) * 0.1
        db = gpu.to_device(b)
        # (a) fixed budget: exactly K updates, iterates agree to rounding
        for K in (0, 1, 7, 50):
            dx = gpu.to_device(x0)
            cg.solve(db, dx, K, 1e-300)
            it, rn, r0 = cg.result()
            xo, ito = orc.cg(rp, ci, v, b, x0, K, 1e-300)
            assert it == ito == K
            np.testing.assert_allclose(gpu.to_host(dx, n), xo, rtol=1e-11, atol=1e-13)
            gpu.free(dx)
        # (b) to tolerance: same stopping iteration (+-1 at the threshold), same solution
        dx = gpu.to_device(x0)
        cg.solve(db, dx, n, 1e-12)
        it, rn, r0 = cg.result()
        xo, ito = orc.cg(rp, ci, v, b, x0, n, 1e-12)
        assert abs(it - ito) <= 1
        assert rn < 1e-12 * r0
        got = gpu.to_host(dx, n)
        assert np.linalg.norm(got - xo) / np.linalg.norm(xo) < 1e-10
        gpu.free(dx); gpu.free(db); cg.close(); A.close()


def test_cg_exact_initial_guess_is_a_no_op(gpu, sz, orc):
    rp, ci, v = orc.laplacian2d(10)
    n = 100
    A = sz.Csr(gpu, rp, ci, v)
    cg = sz.Cg(gpu, A)
    db = gpu.zeros(n); dx = gpu.zeros(n)
    cg.solve(db, dx, 20, 1e-12)      # r0 = 0: 0 < 0 is false -> runs to the cap, x untouched
    it, rn, r0 = cg.result()
    assert it == 20 and rn == 0.0
    assert np.array_equal(gpu.to_host(dx, n), np.zeros(n))
    gpu.free(db); gpu.free(dx); cg.close(); A.close()


def test_gmres_matches_oracle(gpu, sz, orc, ani4):
    rng = np.random.default_rng(6)
    rp, ci, v = ani4
    n = len(rp) - 1
    A = sz.Csr(gpu, rp, ci, v)
    b = rng.standard_normal(n)
    db = gpu.to_device(b)
    for m, K in ((1, 5), (5, 12), (30, 30), (30, 75)):
        g = sz.Gmres(gpu, A, m)
        dx = gpu.zeros(n)
        g.solve(db, dx, K, 1e-300)
        it, rn, r0 = g.result()
        xo, ito = orc.gmres(rp, ci, v, b, np.zeros(n), K, 1e-300, m)
        assert it == ito == K
        np.testing.assert_allclose(gpu.to_host(dx, n), xo, rtol=1e-9, atol=1e-11)
        gpu.free(dx); g.close()
    # above the one-CTA limit (16 384 rows): the multi-kernel GMRES
    rpL, ciL, vL = orc.laplacian2d(130)
    nL = len(rpL) - 1
    AL = sz.Csr(gpu, rpL, ciL, vL)
    bL = rng.standard_normal(nL)
    dbL = gpu.to_device(bL)
    for m, K in ((5, 12), (20, 45)):
        g = sz.Gmres(gpu, AL, m)
        dx = gpu.zeros(nL)
        g.solve(dbL, dx, K, 1e-300)
        it, rn, r0 = g.result()
        xo, ito = orc.gmres(rpL, ciL, vL, bL, np.zeros(nL), K, 1e-300, m)
        assert it == ito == K
        np.testing.assert_allclose(gpu.to_host(dx, nL), xo, rtol=1e-9, atol=1e-11)
        gpu.free(dx); g.close()
    gpu.free(dbL); AL.close()
    g = sz.Gmres(gpu, A, 30)
    dx = gpu.zeros(n)
    g.solve(db, dx, 3000, 1e-10)
    it, rn, r0 = g.result()
    xo, ito = orc.gmres(rp, ci, v, b, np.zeros(n), 3000, 1e-10, 30)
    assert abs(it - ito) <= 2
    got = gpu.to_host(dx, n)
    assert np.linalg.norm(got - xo) / np.linalg.norm(xo) < 1e-8
    gpu.free(dx); gpu.free(db); g.close(); A.close()


def test_sptrsv_matches_oracle(gpu, sz, orc):
    rng = np.random.default_rng(7)
    rp, ci, v = orc.laplacian2d(30)
    n = len(rp) - 1
    for perm in (None, sz.nd_ordering(rp, ci)):
        Lrp, Lci, Lv = sz.host_cholesky(rp, ci, v, perm)
        import scipy.sparse as sp
        U = sp.csr_matrix((Lv, Lci, Lrp), shape=(n, n)).T.tocsr(); U.sort_indices()
        Urp, Uci, Uv = U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data
        b = rng.standard_normal(n)
        db = gpu.to_device(b); dy = gpu.zeros(n); dz = gpu.zeros(n)
        tl = sz.Trs(gpu, Lrp, Lci, Lv, upper=False)
        tu = sz.Trs(gpu, Urp, Uci, Uv, upper=True)
        assert 1 <= tl.levels() <= n and tl.levels() == tu.levels()
        tl.solve(db, dy); tu.solve(dy, dz)
        yo = orc.trs(Lrp, Lci, Lv, b, upper=False)
        zo = orc.trs(Urp, Uci, Uv, yo, upper=True)
        np.testing.assert_allclose(gpu.to_host(dy, n), yo, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(gpu.to_host(dz, n), zo, rtol=1e-11, atol=1e-13)
        # replay of the captured level graph gives the same answer
        tl.solve(db, dy)
        np.testing.assert_allclose(gpu.to_host(dy, n), yo, rtol=1e-12, atol=1e-14)
        for p in (db, dy, dz):
            gpu.free(p)
        tl.close(); tu.close()


def test_fused_reductions_beyond_65536_row_tiles(gpu, sz, orc):
    """cfg4's slabs have 17.3 M rows = 67.6 k row tiles: the persistent SpMV keeps one partial
    per resident CTA, so the fused dot / norm must work past the 65 536-entry scratch (this
    used to be rejected).  3-D 7-pt Laplacian 262^3 (18.0 M rows), CG against the row-sum
    identity: A * 1 is the count of missing neighbours, checked bit-exactly against the
    oracle's SpMV and through 3 CG iterations (fused p.q and ||r||) against the oracle's CG."""
    n = 262
    rp, ci, v = sz.laplacian3d(n)
    N = n ** 3
    assert N // 256 > 65536
    A = sz.Csr(gpu, rp, ci, v)
    ones = np.ones(N)
    dx = gpu.to_device(ones)
    dy = gpu.zeros(N)
    A.spmv(dx, dy, 1.0, 0.0)
    got = gpu.to_host(dy, N)
    cg = sz.Cg(gpu, A)
    b = gpu.to_device(got)
    x = gpu.zeros(N)
    cg.solve(b, x, 3, 1e-30)
    it, res, res0 = cg.result()
    orc.set_threads(orc.max_threads())      # 126 M non-zeros on the CPU: use the cores ...
    try:
        assert np.array_equal(got, orc.spmv(rp, ci, v, ones))
        xo, ito = orc.cg(rp, ci, v, got, np.zeros(N), 3, 1e-30)
    finally:
        orc.set_threads(1)                  # ... but do not leave them on for the tiny cases
    assert it == ito == 3
    xg = gpu.to_host(x, N)
    assert np.linalg.norm(xg - xo) <= 1e-12 * np.linalg.norm(xo)
    assert res0 == pytest.approx(np.linalg.norm(got), rel=1e-13)
    for p in (dx, dy, b, x):
        gpu.free(p)
    cg.close()
    A.close()


def test_sptrsv_blocked_dense_top_of_a_nested_dissection_factor(gpu, sz, orc):
    """A factor with both regimes: wide levels (one launch each) and long runs of small levels
    (the dense separator triangles), which are solved in blocks of <= 128 rows through the
    explicit inverse of the block's own triangle.  L then U = L^T against the oracle's serial
    substitution; the residual L (L^T z) = b closes the loop independently of the oracle."""
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    rp, ci, v = sz.laplacian2d(120)
    n = len(rp) - 1
    perm = sz.nd_ordering(rp, ci)
    Lrp, Lci, Lv = sz.host_cholesky(rp, ci, v, perm)
    Lm = sp.csr_matrix((Lv, Lci, Lrp), shape=(n, n))
    U = Lm.T.tocsr(); U.sort_indices()
    Urp, Uci, Uv = U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data
    tl = sz.Trs(gpu, Lrp, Lci, Lv, upper=False)
    tu = sz.Trs(gpu, Urp, Uci, Uv, upper=True)
    assert tl.levels() > 300          # long dependency chain ...
    b = rng.standard_normal(n)
    db = gpu.to_device(b); dy = gpu.zeros(n); dz = gpu.zeros(n)
    for _ in range(2):                # second pass replays the captured graph
        tl.solve(db, dy); tu.solve(dy, dz)
        y = gpu.to_host(dy, n); z = gpu.to_host(dz, n)
        yo = orc.trs(Lrp, Lci, Lv, b, upper=False)
        zo = orc.trs(Urp, Uci, Uv, yo, upper=True)
        assert np.linalg.norm(y - yo) <= 1e-12 * np.linalg.norm(yo)
        assert np.linalg.norm(z - zo) <= 1e-11 * np.linalg.norm(zo)
        assert np.linalg.norm(Lm @ (Lm.T @ z) - b) <= 1e-11 * np.linalg.norm(b)
    for p in (db, dy, dz):
        gpu.free(p)
    tl.close(); tu.close()


def test_sptrsv_one_kernel_solve_equals_the_level_graph(gpu, sz, orc, monkeypatch):
    """the dependency-driven solve (one persistent kernel, entries of x as their own ready flags)
    and the level-per-launch graph agree to rounding on a nested-dissection factor (wide
    levels, thread-per-row levels, inverted blocks) and on an ILU(0)-like wavefront factor; the
    error word of the solve stays clear; repeated solves on the same buffers are reproducible"""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    cases = []
    rp, ci, v = sz.laplacian2d(150)
    n = len(rp) - 1
    perm = sz.nd_ordering(rp, ci)
    cases.append(sz.host_cholesky(rp, ci, v, perm))
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    Lw = sp.tril(A).tocsr(); Lw.sort_indices()      # natural-order wavefronts: 2 n - 1 levels
    cases.append((Lw.indptr.astype(np.int32), Lw.indices.astype(np.int32), Lw.data))
    for Lrp, Lci, Lv in cases:
        Lm = sp.csr_matrix((Lv, Lci, Lrp), shape=(n, n))
        U = Lm.T.tocsr(); U.sort_indices()
        Urp, Uci, Uv = U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data
        b = rng.standard_normal(n)
        db = gpu.to_device(b); dy = gpu.zeros(n); dz = gpu.zeros(n)
        out = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("SCHWZ_B200_TRS_LEVELS", mode)
            tl = sz.Trs(gpu, Lrp, Lci, Lv, upper=False)
            tu = sz.Trs(gpu, Urp, Uci, Uv, upper=True)
            for rep in range(3):
                tl.solve(db, dy); tu.solve(dy, dz)
                y, z = gpu.to_host(dy, n), gpu.to_host(dz, n)
                if rep:
                    assert np.array_equal(y, out[mode][0]) and np.array_equal(z, out[mode][1])
                out[mode] = (y, z)
            assert tl.error() == 0 and tu.error() == 0
            tl.close(); tu.close()
        # same row sums up to the association inside the block rows (4-way vs 1-way partials)
        for k in (0, 1):
            assert np.linalg.norm(out["0"][k] - out["1"][k]) <= 1e-13 * np.linalg.norm(out["1"][k])
        yo = orc.trs(Lrp, Lci, Lv, b, upper=False)
        assert np.linalg.norm(out["0"][0] - yo) <= 1e-12 * np.linalg.norm(yo)
        assert np.linalg.norm(Lm @ (Lm.T @ out["0"][1]) - b) <= 1e-10 * np.linalg.norm(b)
        for p in (db, dy, dz):
            gpu.free(p)
