"""Golden vectors produced by THE REFERENCE ITSELF (oracle/_ref: /root/reference/source/*.cpp
compiled unmodified against the stand-ins of oracle/ref_shim/, ranks as threads).

Run in the build container (needs /root/reference):
    python tests/golden/make_ref_golden.py

Writes tests/golden/ref_<case>.npz.  Each file holds, per rank r:
  first_row, partition_indices (when a partitioner ran), sizes_r, l2g_r (own | overlap | halo
  sweep), nbr_in_r / nbr_out_r, get_<r>_<j> / put_<r>_<j>, local_res_r (the local residual
  norm pushed at every outer iteration), iters, and x_<r>_<k>: the values the reference held
  at positions l2g_r right after the exchange of outer iteration k (k in `snap`).
The CPU suite (tests/test_ref_pinning.py) compares the oracle with the live library; the GPU
suite (tests/test_gpu_ref_golden.py) compares the CUDA path with these files, because
/root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    # BASELINE.json configs[0]
    "cfg1_lap100_P2_cg": dict(P=2, laplacian_n=100, partition="regular", max_iters=300,
                              tolerance=1e-6, snap=[1, 2, 5, 50, 150]),
    "lap16_P4_regular2d_cg": dict(P=4, laplacian_n=16, partition="regular2d", max_iters=200,
                                  tolerance=1e-8, snap=list(range(0, 12))),
    "lap32_P4_strips_cg_budget": dict(P=4, laplacian_n=32, partition="regular", max_iters=30,
                                      tolerance=1e-12, local_max_iters=15, snap=[1, 2, 10, 29]),
    # BASELINE.json configs[2]
    "cfg3_ani4_metis_P2_gmres": dict(P=2, matrix="ani4_crop", partition="metis", max_iters=800,
                                     tolerance=1e-6, non_symmetric=True, restart_iter=30,
                                     snap=[1, 3, 10]),
    "cfg3_ani4_metis_P4_gmres": dict(P=4, matrix="ani4_crop", partition="metis", max_iters=800,
                                     tolerance=1e-6, non_symmetric=True, restart_iter=30,
                                     snap=[1, 3, 10]),
    "cfg3_ani4_metis_P8_gmres": dict(P=8, matrix="ani4_crop", partition="metis", max_iters=800,
                                     tolerance=1e-6, non_symmetric=True, restart_iter=30,
                                     snap=[1, 3, 10]),
    # the other shipped matrix, matrices/ani3_crop.mtx (N = 741)
    "cfg3_ani3_metis_P2_gmres": dict(P=2, matrix="ani3_crop", partition="metis", max_iters=800,
                                     tolerance=1e-6, non_symmetric=True, restart_iter=30,
                                     snap=[1, 3, 10]),
    "cfg3_ani3_metis_P4_gmres": dict(P=4, matrix="ani3_crop", partition="metis", max_iters=800,
                                     tolerance=1e-6, non_symmetric=True, restart_iter=30,
                                     snap=[1, 3, 10]),
    "cfg3_ani3_metis_P8_gmres": dict(P=8, matrix="ani3_crop", partition="metis", max_iters=800,
                                     tolerance=1e-6, non_symmetric=True, restart_iter=30,
                                     snap=[1, 3, 10]),
}


# Sampled goldens: the same reference runs at sizes whose full vectors would not be "small
# fixtures".  Stored per rank: sizes, neighbour lists, the local residual norm of every outer
# iteration, and for the iterations in `snap` the norm of the iterate over l2g plus its values at
# a fixed set of positions of l2g (every OWN_STRIDE-th own entry, every EXT_STRIDE-th overlap /
# halo entry - the part the exchange writes).
OWN_STRIDE, EXT_STRIDE = 251, 7
SAMPLED_CASES = {
    # the bench workload (BASELINE.json configs[1]: strips, CG with local_max_iters = 50,
    # synchronous exchange, global check) at 1024^2: first 13 outer iterations
    "cfg2_lap1024_P8_cg50": dict(P=8, laplacian_n=1024, partition="regular", max_iters=13,
                                 tolerance=1e-30, local_max_iters=50, snap=[1, 2, 4, 8, 12]),
}


def sample_positions(local_size, n_known):
    own = np.arange(0, local_size, OWN_STRIDE)
    ext = np.arange(local_size, n_known, EXT_STRIDE)
    return np.concatenate([own, ext]).astype(np.int64)


def main_sampled():
    import ref as R
    for name, c in SAMPLED_CASES.items():
        c = dict(c)
        P = c.pop("P")
        snap = c.pop("snap")
        rr = R.Run(P, overlap=2, local_tol=1e-12, enable_global_check=True, record_iterates=True, **c)
        out = {"first_row": rr.vec("first_row", 0), "snap": np.array(snap, np.int32),
               "iters": np.array([rr.iter_count(r) for r in range(P)], np.int32),
               "strides": np.array([OWN_STRIDE, EXT_STRIDE], np.int32)}
        for r in range(P):
            s = rr.sizes(r)
            out["sizes_%d" % r] = np.array([s[k] for k in ("local_size", "local_size_x", "overlap_size",
                                                          "nnz_local", "nnz_interface")], np.int64)
            g2l = rr.vec("g2l", r)
            n_known = int(g2l.max())
            l2g = rr.vec("l2g", r)[:n_known]
            pos = sample_positions(s["local_size"], n_known)
            out["n_known_%d" % r] = np.array([n_known], np.int64)
            out["l2g_sample_%d" % r] = l2g[pos]
            out["l2g_checksum_%d" % r] = np.array([int(l2g.astype(np.int64).sum()),
                                                    int((l2g.astype(np.int64) * (np.arange(n_known) % 1009)).sum())],
                                                   np.int64)
            out["nbr_in_%d" % r] = rr.vec("neighbors_in", r)
            out["nbr_out_%d" % r] = rr.vec("neighbors_out", r)
            out["local_res_%d" % r] = rr.vec("local_residuals", r)
            for k in snap:
                if k < rr.num_iterates(r):
                    x = rr.iterate(r, k)[l2g]
                    out["xnorm_%d_%d" % (r, k)] = np.array([np.linalg.norm(x)])
                    out["x_%d_%d" % (r, k)] = x[pos]
        path = os.path.join(HERE, "refs_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(name, "iters", out["iters"].tolist(), os.path.getsize(path), "bytes")


def main(only=None):
    import ref as R
    for name, c in CASES.items():
        if only and only not in name:
            continue
        c = dict(c)
        P = c.pop("P")
        snap = c.pop("snap")
        matrix = c.pop("matrix", None)
        if matrix:
            c["matrix_file"] = "/root/reference/matrices/%s.mtx" % matrix
        rr = R.Run(P, overlap=2, local_tol=1e-12, enable_global_check=True, record_iterates=True, **c)
        out = {"first_row": rr.vec("first_row", 0), "snap": np.array(snap, np.int32),
               "partition_indices": rr.vec("partition_indices", 0),
               "iters": np.array([rr.iter_count(r) for r in range(P)], np.int32)}
        for r in range(P):
            s = rr.sizes(r)
            out["sizes_%d" % r] = np.array([s[k] for k in ("local_size", "local_size_x", "overlap_size",
                                                          "nnz_local", "nnz_interface")], np.int64)
            g2l = rr.vec("g2l", r)
            n_known = int(g2l.max())
            l2g = rr.vec("l2g", r)[:n_known]
            out["l2g_%d" % r] = l2g
            out["nbr_in_%d" % r] = rr.vec("neighbors_in", r)
            out["nbr_out_%d" % r] = rr.vec("neighbors_out", r)
            for j in range(s["num_neighbors_in"]):
                out["get_%d_%d" % (r, j)] = rr.get_list(r, j)
            for j in range(s["num_neighbors_out"]):
                out["put_%d_%d" % (r, j)] = rr.put_list(r, j)
            out["local_res_%d" % r] = rr.vec("local_residuals", r)
            for k in snap:
                if k < rr.num_iterates(r):
                    out["x_%d_%d" % (r, k)] = rr.iterate(r, k)[l2g]
        sol = rr.vec("solution", 0)
        out["solution_norm"] = np.array([np.linalg.norm(sol)])
        out["solution_head"] = sol[:64]
        path = os.path.join(HERE, "ref_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(name, "iters", out["iters"].tolist(), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sampled":
        main_sampled()
    elif len(sys.argv) > 1:
        main(only=sys.argv[1])     # e.g. "ani3": only the cases whose name contains it
    else:
        main()
        main_sampled()
