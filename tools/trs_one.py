import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, "schwarz-lib_b200")
import schwz_b200 as S
n, P = 4096, 64
setup = S.Setup(("laplacian2d", n), P, part=S.partition_regular2d(n * n, P))
rp, ci, v = setup.local_matrix(9)
rows = len(rp) - 1
perm = S.nd_ordering(rp, ci)
Lrp, Lci, Lv = S.host_cholesky(rp, ci, v, perm)
U = sp.csr_matrix((Lv, Lci, Lrp), shape=(rows, rows)).T.tocsr(); U.sort_indices()
c = S.Context(0)
tl = S.Trs(c, Lrp, Lci, Lv, upper=False)
tu = S.Trs(c, U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data, upper=True)
b = c.to_device(np.ones(rows)); y = c.zeros(rows); z = c.zeros(rows)
for _ in range(2):
    tl.solve(b, y); c.sync(); tu.solve(y, z); c.sync()
