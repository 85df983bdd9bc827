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
}


def main():
    import ref as R
    for name, c in CASES.items():
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
    main()
