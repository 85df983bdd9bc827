"""The px x py rectangular extension of regular2d (BASELINE.json configs[1] asks for a "regular 2D
partition" of 8 subdomains; the reference's rule, include/partition_tools.hpp:70-94, truncates
sqrt(8) to 2 and leaves four subdomains without rows - SURVEY F5).  For perfect squares the
extension IS the reference's rule (checked against the reference's own run, tests/golden/ref_*);
otherwise the oracle and the product agree bit for bit on every index set it leads to, and the
CUDA path holds the 1e-10 contract against the oracle on it."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_setup import _same


@pytest.mark.parametrize("n,P", [(16, 4), (24, 9), (32, 16), (64, 64)])
def test_equals_the_reference_rule_for_perfect_squares(orc, sz, n, P):
    ref = orc.partition_regular2d(n * n, P)
    assert np.array_equal(orc.partition_regular2d_rect(n * n, P), ref)
    assert np.array_equal(sz.partition_regular2d_rect(n * n, P), ref)


def test_equals_the_reference_run_itself():
    g = np.load(os.path.join(GOLDEN, "ref_lap16_P4_regular2d_cg.npz"))
    import oracle as O
    assert np.array_equal(O.partition_regular2d_rect(256, 4), g["partition_indices"])


@pytest.mark.parametrize("n,P,shape", [(16, 8, (2, 4)), (32, 8, (2, 4)), (30, 8, (2, 4)),
                                       (32, 2, (1, 2)), (17, 6, (2, 3)), (24, 12, (3, 4))])
def test_rectangular_blocks(orc, sz, n, P, shape):
    a = orc.partition_regular2d_rect(n * n, P)
    assert np.array_equal(a, sz.partition_regular2d_rect(n * n, P))
    px, py = shape
    assert np.array_equal(a, orc.partition_regular2d_rect(n * n, P, px, py))
    grid = a.reshape(n, n)
    counts = np.bincount(a, minlength=P)
    assert counts.min() > 0 and counts.sum() == n * n          # nobody is left without rows
    for j1 in range(px):
        for j2 in range(py):
            r0, r1 = j1 * n // px, (j1 + 1) * n // px
            c0, c1 = j2 * n // py, (j2 + 1) * n // py
            assert (grid[r0:r1, c0:c1] == py * j1 + j2).all()   # numbering of partition_tools.hpp:84


def test_rejects_impossible_shapes(orc, sz):
    with pytest.raises(ValueError):
        orc.partition_regular2d_rect(15, 4)                     # N is not a square
    with pytest.raises(sz.SchwzError):
        sz.partition_regular2d_rect(256, 8, 3, 3)               # px * py != P


@pytest.mark.parametrize("n,P", [(16, 8), (24, 6), (20, 2)])
def test_index_sets_on_the_extension_are_bit_exact(orc, sz, n, P):
    part = orc.partition_regular2d_rect(n * n, P)
    mat = orc.laplacian2d(n)
    _same(orc.Problem(*mat, P, part=part), sz.Setup(mat, P, part=part), P, True)
    _same(orc.Problem(*mat, P, part=part), sz.Setup(("laplacian2d", n), P, part=part), P, True)
    # 2 x 4 blocks: edge neighbours exchange a block side (16 / 8 values), diagonal neighbours a
    # single corner value (the halo layer of the overlap rows reaches round the corner,
    # SURVEY section 8e)
    if (n, P) == (16, 8):
        ob = orc.Problem(*mat, P, part=part)
        assert sorted(len(ob.neighbors(r)[0]) for r in range(P)) == [3, 3, 3, 3, 5, 5, 5, 5]
        assert [len(ob.get_list(1, j)) for j in range(5)] == [16, 16, 1, 8, 1]


@pytest.mark.gpu
def test_cuda_iterates_on_2x4_blocks_match_the_oracle(orc, sz):
    if sz.device_count() < 1:
        pytest.skip("no CUDA device")
    n, P = 32, 8
    part = sz.partition_regular2d_rect(n * n, P)
    mat = orc.laplacian2d(n)
    ob = orc.Problem(*mat, P, part=part)
    ob.configure(tolerance=1e-8, local_tol=1e-12, max_iters=400, enable_global_check=True)
    setup = sz.Setup(("laplacian2d", n), P, part=part)
    ctxs = [sz.Context(0) for _ in range(P)]
    subs = [sz.Ras(ctxs[r], setup, r, local_tol=1e-12) for r in range(P)]
    sz.connect_local(subs, setup)
    out = sz.ras_run(subs, P, 400, tolerance=1e-8, enable_global_check=True, history=True)
    ob.run()
    assert out["converged"] and out["iters"] == int(ob.status(0)["finished_iter"])
    for r in range(P):
        ref = ob.history(r)[0]
        assert np.max(np.abs(out["history"][:, r] - ref[:len(out["history"])])) <= 1e-9 * ref[0]
        l2g = setup.l2g(r)[:subs[r].local_size]
        xo = ob.x(r)[l2g]
        assert np.linalg.norm(subs[r].x()[:subs[r].local_size] - xo) <= 1e-10 * np.linalg.norm(xo)
    for s in subs:
        s.close()
    for c in ctxs:
        c.close()
