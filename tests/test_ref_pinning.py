"""Pins the CPU oracle (oracle/schwz_oracle.cpp) against the REFERENCE ITSELF.

oracle/_ref/libschwz_ref.so is the reference's own source/*.cpp compiled unmodified
(oracle/Makefile) against stand-ins for MPI (threads), Ginkgo (host-only subset) and the
generated config header (oracle/ref_shim/).  Everything integer that comes out of it -
partition, permutation, overlap/halo numbering, local and interface matrices, get/put
lists, displacement tables - and the whole outer-iteration orchestration (exchange,
boundary update, convergence protocols) is produced by the reference's code; the tests
below require the restated oracle to reproduce all of it bit for bit, and the residual
histories / iterates too (they agree exactly because both sum in the same order).

No GPU needed.  Skipped only when the library can be neither found nor built.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def ref():
    import ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    R.build()
    return R


def write_mtx(path, mat):
    rp, ci, v = mat
    n = len(rp) - 1
    rows = np.repeat(np.arange(n), np.diff(rp))
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n" % (n, n, len(ci)))
        for r, c, x in zip(rows, ci, v):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, x))
    return str(path)


def same_setup(rr, ob, P, permuted):
    """reference run `rr` vs oracle problem `ob`: every index set, bit-exact."""
    assert np.array_equal(rr.vec("first_row", 0), ob.first_row())
    if permuted:
        perm, iperm = ob.permutation()
        assert np.array_equal(rr.vec("permutation", 0), perm)
        assert np.array_equal(rr.vec("i_permutation", 0), iperm)
    for a, b in zip(rr.global_matrix(0), ob.global_matrix()):
        assert np.array_equal(a, b)
    for r in range(P):
        so, sr = ob.sizes(r), rr.sizes(r)
        for k in ("local_size", "local_size_x", "overlap_size", "nnz_local", "nnz_interface",
                  "num_neighbors_in", "num_neighbors_out"):
            assert so[k] == sr[k], (r, k)
        assert np.array_equal(rr.vec("first_row", r), ob.first_row())
        l2g = ob.l2g(r)
        l2g_ref = rr.vec("l2g", r)
        assert np.array_equal(l2g_ref[:len(l2g)], l2g)          # own | overlap | halo sweep
        assert not l2g_ref[len(l2g):].any()
        assert np.array_equal(rr.vec("g2l", r), ob.g2l(r))
        assert np.array_equal(rr.vec("overlap_row", r),
                              l2g[so["local_size"]:so["local_size_x"]])
        for a, b in zip(rr.local_matrix(r), ob.local_matrix(r)):
            assert np.array_equal(a, b)
        if so["nnz_interface"]:
            for a, b in zip(rr.interface_matrix(r), ob.interface_matrix(r)):
                assert np.array_equal(a, b)
        nin, nout = ob.neighbors(r)
        assert np.array_equal(rr.vec("neighbors_in", r), nin)
        assert np.array_equal(rr.vec("neighbors_out", r), nout)
        for j in range(so["num_neighbors_in"]):
            assert np.array_equal(rr.get_list(r, j), ob.get_list(r, j))
        for j in range(so["num_neighbors_out"]):
            assert np.array_equal(rr.put_list(r, j), ob.put_list(r, j))
        assert np.array_equal(rr.vec("local_rhs", r), ob.local_rhs(r))


def same_history(rr, ob, P):
    """sync mode, EVERY outer iteration: the oracle is stepped one pass of the loop of
    source/schwarz_base.cpp:387-452 at a time; after pass k its x (own rows: k+1 local solves;
    halo entries: what the exchange of pass k scattered) must equal what the reference held
    when it entered update_boundary in pass k / k+1, and the local residual norms pushed by
    check_convergence must be the same doubles.  Same outer iteration count at the end."""
    fr = ob.first_row()
    n_ref = [rr.num_iterates(r) for r in range(P)]
    k = 0
    alive = P
    while k < ob.max_iters and alive > 0:
        x_before = [ob.x(r) for r in range(P)]
        alive -= ob.step()
        for r in range(P):
            own = slice(fr[r], fr[r + 1])
            if k < n_ref[r]:
                # reference snapshot k = its x right after exchange k: own rows are the result
                # of k local solves (= oracle before this pass), the rest holds halo values
                ref_x = rr.iterate(r, k)
                assert np.array_equal(ref_x[own], x_before[r][own]), (r, k)
                halo = np.ones(len(ref_x), bool)
                halo[own] = False
                assert np.array_equal(ref_x[halo], ob.x(r)[halo]), (r, k)
        k += 1
    for r in range(P):
        res, _ = ob.history(r)
        ref_res = rr.vec("local_residuals", r)
        assert len(res) == len(ref_res) == n_ref[r]
        assert np.array_equal(res, ref_res)
        st = ob.status(r)
        ref_iters = rr.iter_count(r)
        assert (int(st["finished_iter"]) if st["finished"] else ob.iter_count()) == ref_iters
        pd, gd = ob.displacements(r)
        assert np.array_equal(rr.vec("put_displacements", r), pd)
        assert np.array_equal(rr.vec("get_displacements", r), gd)


def test_cfg1_laplacian100_two_strips_cg(ref, orc):
    """BASELINE.json configs[0]: 100x100, regular, 2 ranks, CG to 1e-12, sync, global check."""
    orc.set_threads(1)
    rr = ref.Run(2, laplacian_n=100, partition="regular", overlap=2, max_iters=300,
                 tolerance=1e-6, local_tol=1e-12, enable_global_check=True, record_iterates=True)
    ob = orc.Problem(*orc.laplacian2d(100), 2)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=300, enable_global_check=True)
    same_setup(rr, ob, 2, False)
    same_history(rr, ob, 2)
    assert rr.iter_count(0) == rr.iter_count(1) == 150
    assert " relative residual norm of solution 6.67266e-07" in rr.log
    # the solution gathered by compute_residual_norm (source/solve.cpp:1025-1085), rank 0
    x, fin = ob.final_residual()
    np.testing.assert_array_equal(rr.vec("solution", 0), x)


# (6, 3, 4) is deliberately absent: there the overlap of the middle strip swallows the whole
# domain, nnz_interface is 0 and the reference skips filling the overlap rows of its local
# matrix altogether (restricted_schwarz.cpp:263-284) - a broken matrix the
# oracle does not imitate.
@pytest.mark.parametrize("n,P,overlap", [(16, 4, 2), (6, 2, 3), (12, 3, 3), (9, 3, 2), (20, 7, 2),
                                         (10, 5, 1)])
def test_regular_strips(ref, orc, n, P, overlap):
    orc.set_threads(1)
    rr = ref.Run(P, laplacian_n=n, partition="regular", overlap=overlap, max_iters=40,
                 tolerance=1e-8, local_tol=1e-12, enable_global_check=True, record_iterates=True)
    ob = orc.Problem(*orc.laplacian2d(n), P, overlap=overlap)
    ob.configure(tolerance=1e-8, local_tol=1e-12, max_iters=40, enable_global_check=True)
    same_setup(rr, ob, P, False)
    same_history(rr, ob, P)


@pytest.mark.parametrize("n,P", [(16, 4), (8, 4), (24, 9), (32, 16)])
def test_regular2d(ref, orc, n, P):
    orc.set_threads(1)
    rr = ref.Run(P, laplacian_n=n, partition="regular2d", overlap=2, max_iters=30,
                 tolerance=1e-8, local_tol=1e-12, enable_global_check=True, record_iterates=True)
    part = orc.partition_regular2d(n * n, P)
    assert np.array_equal(rr.vec("partition_indices", 0), part)
    ob = orc.Problem(*orc.laplacian2d(n), P, part=part)
    ob.configure(tolerance=1e-8, local_tol=1e-12, max_iters=30, enable_global_check=True)
    same_setup(rr, ob, P, True)
    same_history(rr, ob, P)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_cfg3_ani4_metis_gmres(ref, orc, sz, ani4, tmp_path, P):
    """BASELINE.json configs[2]: the reference reads the .mtx, calls METIS itself, GMRES(30)."""
    orc.set_threads(1)
    path = write_mtx(tmp_path / "ani4.mtx", ani4)
    rr = ref.Run(P, matrix_file=path, partition="metis", overlap=2, max_iters=60,
                 tolerance=1e-6, local_tol=1e-12, non_symmetric=True, restart_iter=30,
                 enable_global_check=True, record_iterates=True)
    part = sz.partition_metis(ani4[0], ani4[1], P)      # the product's METIS call sequence
    assert np.array_equal(rr.vec("partition_indices", 0), part)
    ob = orc.Problem(*ani4, P, part=part)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=60, non_symmetric=True,
                 restart_iter=30, enable_global_check=True)
    same_setup(rr, ob, P, True)
    same_history(rr, ob, P)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_cfg3_ani3_metis_gmres(ref, orc, sz, tmp_path, P):
    """the other shipped matrix, matrices/ani3_crop.mtx: same configuration"""
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "ani3_crop.npz"))
    ani3 = (z["rowptr"], z["col"], z["val"])
    orc.set_threads(1)
    path = write_mtx(tmp_path / "ani3.mtx", ani3)
    rr = ref.Run(P, matrix_file=path, partition="metis", overlap=2, max_iters=60,
                 tolerance=1e-6, local_tol=1e-12, non_symmetric=True, restart_iter=30,
                 enable_global_check=True, record_iterates=True)
    part = sz.partition_metis(ani3[0], ani3[1], P)
    assert np.array_equal(rr.vec("partition_indices", 0), part)
    ob = orc.Problem(*ani3, P, part=part)
    ob.configure(tolerance=1e-6, local_tol=1e-12, max_iters=60, non_symmetric=True,
                 restart_iter=30, enable_global_check=True)
    same_setup(rr, ob, P, True)
    same_history(rr, ob, P)


def test_ani4_regular_cg_fixed_budget(ref, orc, ani4, tmp_path):
    """fixed local budget (local_max_iters) instead of a converged local solve."""
    orc.set_threads(1)
    path = write_mtx(tmp_path / "ani4.mtx", ani4)
    rr = ref.Run(4, matrix_file=path, partition="regular", overlap=2, max_iters=25,
                 tolerance=1e-10, local_tol=1e-12, local_max_iters=20, non_symmetric=True,
                 restart_iter=10, enable_global_check=True, record_iterates=True)
    ob = orc.Problem(*ani4, 4)
    ob.configure(tolerance=1e-10, local_tol=1e-12, max_iters=25, local_max_iters=20,
                 non_symmetric=True, restart_iter=10, enable_global_check=True)
    same_setup(rr, ob, 4, False)
    same_history(rr, ob, 4)


def test_sync_without_global_check_never_stops_early(ref, orc):
    """two-sided, --enable_global_check off: converged_all_local is only ever raised inside the
    global-check branch (source/solve.cpp:888-912), so the MPI_Allreduce of flags at :949-953
    sums zeros and the loop always runs to --num_iters.  A reference quirk the oracle keeps."""
    orc.set_threads(1)
    rr = ref.Run(3, laplacian_n=12, partition="regular", overlap=2, max_iters=60,
                 tolerance=1e-5, local_tol=1e-12, enable_global_check=False, record_iterates=True)
    ob = orc.Problem(*orc.laplacian2d(12), 3)
    ob.configure(tolerance=1e-5, local_tol=1e-12, max_iters=60, enable_global_check=False)
    same_setup(rr, ob, 3, False)
    same_history(rr, ob, 3)
    assert rr.iter_count(0) == 60 and ob.iter_count() == 60


@pytest.mark.parametrize("comm,one_by_one,conv", [
    ("put", False, "decentralized"), ("get", False, "decentralized"),
    ("get", True, "decentralized"), ("put", False, "centralized-tree")])
def test_onesided_modes_reach_the_threshold(ref, orc, comm, one_by_one, conv):
    """One-sided runs are asynchronous (threads race like MPI ranks do), so only what the
    north star asks of async mode is compared: the same stopping threshold is met, with the
    same exchange lists and window displacements."""
    n, P, tol = 24, 4, 1e-6
    rr = ref.Run(P, laplacian_n=n, partition="regular", overlap=2, max_iters=4000,
                 tolerance=tol, local_tol=1e-12, enable_onesided=True, remote_comm_type=comm,
                 enable_one_by_one=one_by_one, global_convergence_type=conv)
    ob = orc.Problem(*orc.laplacian2d(n), P)
    ob.configure(tolerance=tol, local_tol=1e-12, max_iters=4000, enable_onesided=True,
                 remote_comm_type=comm, enable_one_by_one=one_by_one, global_convergence_type=conv)
    same_setup(rr, ob, P, False)
    ob.run()
    for r in range(P):
        assert rr.iter_count(r) < 4000
        pd, gd = ob.displacements(r)
        assert np.array_equal(rr.vec("put_displacements", r), pd)
        assert np.array_equal(rr.vec("get_displacements", r), gd)
    # true residual of the gathered solution (rank 0 of the reference)
    rp, ci, v = orc.laplacian2d(n)
    x = rr.vec("solution", 0)
    b = np.ones(n * n)
    rel = np.linalg.norm(b - orc.spmv(rp, ci, v, x)) / np.linalg.norm(b)
    xo, fin = ob.final_residual()
    assert rel < 50 * tol and fin["relative"] < 50 * tol


def test_onesided_put_one_by_one_is_broken_upstream(ref, orc):
    """Found by running the reference: with --enable_one_by_one and remote_comm_type=put the
    element-wise MPI_Put lands in the neighbour's x, but the neighbour then still runs the
    "unpack receive buffer" loop (restricted_schwarz.cpp:788-799 sits outside the one-by-one
    branch) and overwrites those halo entries with its never-written recv_buffer.  The
    reference does not converge; the oracle (and the CUDA path) implement the intended
    semantics - the Put IS the update - and do."""
    n, P, tol = 16, 2, 1e-6
    rr = ref.Run(P, laplacian_n=n, partition="regular", overlap=2, max_iters=400, tolerance=tol,
                 local_tol=1e-12, enable_onesided=True, remote_comm_type="put",
                 enable_one_by_one=True, global_convergence_type="decentralized")
    ob = orc.Problem(*orc.laplacian2d(n), P)
    ob.configure(tolerance=tol, local_tol=1e-12, max_iters=400, enable_onesided=True,
                 remote_comm_type="put", enable_one_by_one=True,
                 global_convergence_type="decentralized")
    ob.run()
    assert rr.iter_count(0) == 400 and rr.iter_count(1) == 400
    assert ob.iter_count() < 400


@pytest.mark.parametrize("onesided", [False, True])
def test_mixed_precision_halo(ref, orc, onesided):
    """settings.use_mixed_precision with MixedValueType = float (what the deal.II drivers of the
    reference instantiate; bench_ras never sets it): halo values are rounded to float on the
    wire (restricted_schwarz.cpp:483-603, 769-787, 898-903).  The reference is run as
    SolverRAS<double, int32, float>."""
    orc.set_threads(1)
    n, P = 16, 4
    kw = dict(max_iters=300, tolerance=1e-6, local_tol=1e-12)
    part = orc.partition_regular2d(n * n, P)
    if not onesided:
        rr = ref.Run(P, laplacian_n=n, partition="regular2d", overlap=2, enable_global_check=True,
                     record_iterates=True, use_mixed_precision=True, **kw)
        ob = orc.Problem(*orc.laplacian2d(n), P, part=part)
        ob.configure(enable_global_check=True, use_mixed_precision=True, **kw)
        same_setup(rr, ob, P, True)
        same_history(rr, ob, P)
        # and it is a different run from the fp64 one
        ob64 = orc.Problem(*orc.laplacian2d(n), P, part=part)
        ob64.configure(enable_global_check=True, **kw)
        ob64.run()
        assert not np.array_equal(ob64.history(0)[0][:10], ob.history(0)[0][:10])
    else:
        rr = ref.Run(P, laplacian_n=n, partition="regular2d", overlap=2, enable_onesided=True,
                     remote_comm_type="put", global_convergence_type="decentralized",
                     use_mixed_precision=True, **kw)
        ob = orc.Problem(*orc.laplacian2d(n), P, part=part)
        ob.configure(enable_onesided=True, remote_comm_type="put",
                     global_convergence_type="decentralized", use_mixed_precision=True, **kw)
        same_setup(rr, ob, P, True)
        ob.run()
        assert all(rr.iter_count(r) < 300 for r in range(P)) and ob.iter_count() < 300
