"""The product's host index sets (schwarz-lib_b200/csrc/setup.cpp, reached
through the C ABI) against the oracle, bit-exact, and against the Appendix E
known answers.  No GPU needed."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

E = json.load(open(os.path.join(GOLDEN, "appendix_e.json")))


def _same(ob, sb, P, permuted):
    assert np.array_equal(ob.first_row(), sb.first_row())
    if permuted:
        for a, b in zip(ob.permutation(), sb.permutation()):
            assert np.array_equal(a, b)
    for r in range(P):
        so = ob.sizes(r)
        assert so == sb.sizes(r)
        assert np.array_equal(ob.l2g(r), sb.l2g(r))
        for a, b in zip(ob.local_matrix(r), sb.local_matrix(r)):
            assert np.array_equal(a, b)
        for a, b in zip(ob.interface_matrix(r), sb.interface_matrix(r)):
            assert np.array_equal(a, b)
        for a, b in zip(ob.neighbors(r), sb.neighbors(r)):
            assert np.array_equal(a, b)
        for j in range(so["num_neighbors_in"]):
            assert np.array_equal(ob.get_list(r, j), sb.get_list(r, j))
        for j in range(so["num_neighbors_out"]):
            assert np.array_equal(ob.put_list(r, j), sb.put_list(r, j))
        for a, b in zip(ob.displacements(r), sb.displacements(r)):
            assert np.array_equal(a, b)


def test_generators_bit_exact(orc, sz):
    for n in (1, 2, 3, 7, 16, 100):
        for a, b in zip(orc.laplacian2d(n), sz.laplacian2d(n)):
            assert np.array_equal(a, b)
    for n in (1, 2, 5, 9):
        for a, b in zip(orc.laplacian3d(n), sz.laplacian3d(n)):
            assert np.array_equal(a, b)
    for N, P in ((256, 4), (256, 8), (64, 4), (4096, 16), (100, 9)):
        assert np.array_equal(orc.partition_regular2d(N, P), sz.partition_regular2d(N, P))


@pytest.mark.parametrize("n,P,overlap", [(100, 2, 2), (16, 4, 2), (6, 2, 3), (6, 3, 4), (9, 3, 2),
                                         (20, 7, 2), (12, 1, 2)])
def test_regular_partition(orc, sz, n, P, overlap):
    mat = orc.laplacian2d(n)
    _same(orc.Problem(*mat, P, overlap=overlap), sz.Setup(mat, P, overlap=overlap), P, False)
    # generated rows (no stored global matrix) give the same arrays
    _same(orc.Problem(*mat, P, overlap=overlap), sz.Setup(("laplacian2d", n), P, overlap=overlap),
          P, False)


@pytest.mark.parametrize("n,P", [(16, 4), (16, 8), (8, 4), (24, 9), (32, 16)])
def test_regular2d_partition(orc, sz, n, P):
    mat = orc.laplacian2d(n)
    part = orc.partition_regular2d(n * n, P)
    _same(orc.Problem(*mat, P, part=part), sz.Setup(("laplacian2d", n), P, part=part), P, True)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_ani4_regular_and_metis(orc, sz, ani4, P):
    _same(orc.Problem(*ani4, P), sz.Setup(ani4, P), P, False)
    part = sz.partition_metis(ani4[0], ani4[1], P)
    # same METIS build as the survey probed (toolkit static lib); informational pin
    assert np.bincount(part).tolist() == E["metis_part_sizes"][str(P)]
    _same(orc.Problem(*ani4, P, part=part), sz.Setup(ani4, P, part=part), P, True)


def test_laplacian3d_slabs(orc, sz):
    mat = orc.laplacian3d(6)
    _same(orc.Problem(*mat, 3), sz.Setup(("laplacian3d", 6), 3), 3, False)


def test_appendix_e_goldens_through_the_c_abi(sz, ani4):
    g = E["ordered"]["lap8_P4_regular2d"]
    sb = sz.Setup(("laplacian2d", 8), 4, part=sz.partition_regular2d(64, 4))
    for r, key in ((0, "rank0"), (3, "rank3")):
        s = sb.sizes(r)
        l2g = sb.l2g(r)
        assert l2g[s["local_size"]:s["local_size_x"]].tolist() == g[key]["overlap_row"]
        assert l2g[s["local_size_x"]:].tolist() == g[key]["halo"]
        nin, _ = sb.neighbors(r)
        assert {str(int(p)): sb.get_list(r, j).tolist() for j, p in enumerate(nin)} == g[key]["get"]
    sb = sz.Setup(ani4, 4)
    want = E["sizes"]["ani4_P4_regular"]
    assert sb.first_row().tolist() == want["first_row"]
    for r in range(4):
        s = sb.sizes(r)
        assert [s["local_size"], s["local_size_x"], s["nnz_local"], s["nnz_interface"]] == want["ranks"][r]


def test_read_mtx_matches_fixture(sz, ani4, tmp_path):
    # write the fixture back as MatrixMarket and parse it with the product reader
    rp, ci, v = ani4
    n = len(rp) - 1
    p = tmp_path / "m.mtx"
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(ci)))
        rng = np.random.default_rng(0)
        rows = np.repeat(np.arange(n), np.diff(rp))
        order = rng.permutation(len(ci))   # unsorted on disk
        for k in order:
            f.write("%d %d %.17g\n" % (rows[k] + 1, ci[k] + 1, v[k]))
    a = sz.read_mtx(str(p))
    for x, y in zip(a, ani4):
        assert np.array_equal(x, y)


def test_host_cholesky_matches_oracle(orc, sz):
    rp, ci, v = orc.laplacian2d(14)
    perm = sz.nd_ordering(rp, ci)
    assert sorted(perm.tolist()) == list(range(len(rp) - 1))
    Lo = orc.cholesky(rp, ci, v, perm)
    Ls = sz.host_cholesky(rp, ci, v, perm)
    assert np.array_equal(Lo[0], Ls[0]) and np.array_equal(Lo[1], Ls[1])
    np.testing.assert_allclose(Ls[2], Lo[2], rtol=1e-13)
    # the ordering pays: fewer non-zeros than the natural order
    assert len(Ls[1]) < len(sz.host_cholesky(rp, ci, v, None)[1])
