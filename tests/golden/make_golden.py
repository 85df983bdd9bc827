"""Regenerates the fixtures under tests/golden/.

Run in the build container (needs /root/reference for the .mtx files):
    python tests/golden/make_golden.py

* ani{3,4}_crop.npz — the reference's shipped matrices
  (/root/reference/matrices/ani{3,4}_crop.mtx) parsed with scipy.io.mmread and
  sorted by column, i.e. what gko::read + sort_by_column_index produce
  (source/initialization.cpp:210-212).  Stored as CSR so the GPU box, which has
  no /root/reference, can run configuration 3.
* appendix_e.json is NOT generated: it holds the known answers of SURVEY.md
  Appendix E, which were derived at survey time by an independent restatement
  of the reference algorithm.  The oracle and the product are both checked
  against it.
* cfg1_history.json — residual history of configuration 1 as produced by the
  oracle (oracle/schwz_oracle.cpp); its first entries and the stopping
  iteration coincide with Appendix E, the rest extends the fixture.
"""
import json
import os
import sys

import numpy as np
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    for name in ("ani3_crop", "ani4_crop"):
        M = sp.csr_matrix(scipy.io.mmread("/root/reference/matrices/%s.mtx" % name))
        M.sort_indices()
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            rowptr=M.indptr.astype(np.int32), col=M.indices.astype(np.int32),
                            val=M.data.astype(np.float64))
    import oracle as O
    rp, ci, v = O.laplacian2d(100)
    pb = O.Problem(rp, ci, v, 2)
    pb.configure(tolerance=1e-6, local_tol=1e-12, max_iters=300, enable_global_check=True)
    iters = pb.run()
    res, gres = pb.history(0)
    x, fr = pb.final_residual()
    json.dump({"iters": iters, "local_resnorm_rank0": res.tolist(), "global_resnorm": gres.tolist(),
               "final_relative_residual": fr["relative"], "sol_norm": fr["sol_norm"],
               "x0": x[0], "x5050": x[5050]},
              open(os.path.join(HERE, "cfg1_history.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
