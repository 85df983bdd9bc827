"""Randomised pinning of the oracle AND of the product's host index sets against the reference
itself (oracle/_ref): seeded random sparse matrices with a symmetric pattern (ragged rows, a few
dense-ish rows, unsymmetric values), the reference's own METIS / regular partitioning, overlaps
1..3, CG or GMRES local solves.  Everything integer must be bit-identical (partition, permutation,
permuted matrix, local / interface matrices, halo lists, displacement tables), and so must the
iterates after every exchange and the residual histories of the oracle.  No GPU needed."""
import numpy as np
import pytest
import scipy.sparse as sp

from test_ref_pinning import ref, same_history, same_setup, write_mtx  # noqa: F401


def random_matrix(n, seed, spd, band=0):
    """band > 0: entries within `band` of the diagonal only (neighbourhoods grow linearly, so
    two overlap layers do not swallow the domain - which the reference does not survive,
    SURVEY Appendix D); band == 0: a random graph with a few hub rows."""
    rng = np.random.default_rng(seed)
    if band:
        rows = rng.integers(0, n, 3 * n)
        cols = np.clip(rows + rng.integers(-band, band + 1, 3 * n), 0, n - 1)
        R = sp.csr_matrix((rng.random(3 * n), (rows, cols)), shape=(n, n))
        H = sp.csr_matrix((n, n))
    else:
        R = sp.random(n, n, density=3.0 / n, random_state=seed, format="csr")
        hub = rng.integers(0, n, 3)                   # a few long rows / columns
        H = sp.csr_matrix((rng.random(3 * 12), (np.repeat(hub, 12), rng.integers(0, n, 36))),
                          shape=(n, n))
    # a ring keeps the graph connected (METIS, and every subdomain gets neighbours)
    ring = sp.csr_matrix((np.full(n, 0.5), (np.arange(n), (np.arange(n) + 1) % n)), shape=(n, n))
    S = R + H + ring
    S = S + S.T
    if not spd:                                       # same pattern, unsymmetric values
        S = S.multiply(sp.csr_matrix((1.0 + 0.3 * rng.random(S.nnz), S.nonzero()), shape=(n, n)))
    S = sp.csr_matrix(S)
    A = sp.csr_matrix(-abs(S) + sp.diags(np.asarray(abs(S).sum(axis=1)).ravel() + 0.5 + rng.random(n)))
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


CASES = [  # n, seed, P, partition, overlap, spd  (overlap 3 runs on a banded matrix)
    (180, 1, 2, "metis", 2, True),
    (240, 2, 3, "metis", 2, False),
    (400, 3, 5, "metis", 2, True),
    (210, 4, 4, "regular", 2, False),
    (160, 5, 3, "regular", 1, True),
    (260, 6, 6, "metis", 2, False),
    (400, 8, 3, "regular", 3, True),
]


@pytest.mark.parametrize("n,seed,P,partition,overlap,spd", CASES)
def test_random_matrix_against_the_reference(ref, orc, sz, tmp_path, n, seed, P, partition, overlap,
                                             spd):
    orc.set_threads(1)
    mat = random_matrix(n, seed, spd, band=12 if overlap == 3 else 0)
    path = write_mtx(tmp_path / "rand.mtx", mat)
    kw = dict(max_iters=12, tolerance=1e-10, local_tol=1e-12, enable_global_check=True)
    if not spd:
        kw.update(non_symmetric=True, restart_iter=20)
    rr = ref.Run(P, matrix_file=path, partition=partition, overlap=overlap, record_iterates=True,
                 **kw)
    part = None
    if partition == "metis":
        part = sz.partition_metis(mat[0], mat[1], P)          # the product's METIS call sequence
        assert np.array_equal(rr.vec("partition_indices", 0), part)
    ob = orc.Problem(*mat, P, part=part, overlap=overlap)
    ob.configure(**kw)
    same_setup(rr, ob, P, part is not None)
    same_history(rr, ob, P)
    # the product's host index sets (what the GPU path is built from) against the reference too
    setup = sz.Setup(mat, P, part=part, overlap=overlap)
    assert np.array_equal(setup.first_row(), rr.vec("first_row", 0))
    for r in range(P):
        s, sr = setup.sizes(r), rr.sizes(r)
        for k in ("local_size", "local_size_x", "overlap_size", "nnz_local", "nnz_interface",
                  "num_neighbors_in", "num_neighbors_out"):
            assert s[k] == sr[k], (r, k)
        for a, b in zip(setup.local_matrix(r), rr.local_matrix(r)):
            assert np.array_equal(a, b), r
        l2g = setup.l2g(r)
        assert np.array_equal(rr.vec("l2g", r)[:len(l2g)], l2g)
        nin, nout = setup.neighbors(r)
        assert np.array_equal(nin, rr.vec("neighbors_in", r)[:len(nin)])
        assert np.array_equal(nout, rr.vec("neighbors_out", r)[:len(nout)])
        for j in range(len(nin)):
            assert np.array_equal(setup.get_list(r, j), rr.get_list(r, j))
        for j in range(len(nout)):
            assert np.array_equal(setup.put_list(r, j), rr.put_list(r, j))
        pd, gd = setup.displacements(r)
        assert np.array_equal(pd, rr.vec("put_displacements", r))
        assert np.array_equal(gd, rr.vec("get_displacements", r))
