// Hand-declared subset of metis.h for the header-less METIS 5 static library
// inside the CUDA toolkit (idx_t = int64, real_t = float).
#ifndef METIS_H_SHIM
#define METIS_H_SHIM
#include <stdint.h>
typedef int64_t idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
enum { METIS_OK = 1, METIS_ERROR_INPUT = -2, METIS_ERROR_MEMORY = -3, METIS_ERROR = -4 };
enum { METIS_OPTION_PTYPE = 0, METIS_OPTION_OBJTYPE = 1, METIS_OPTION_SEED = 8, METIS_OPTION_NUMBERING = 17 };
enum { METIS_OBJTYPE_CUT = 0, METIS_OBJTYPE_VOL = 1, METIS_OBJTYPE_NODE = 2 };
#ifdef __cplusplus
extern "C" {
#endif
int METIS_SetDefaultOptions(idx_t *options);
int METIS_PartGraphRecursive(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                             idx_t *vsize, idx_t *adjwgt, idx_t *nparts, real_t *tpwgts,
                             real_t *ubvec, idx_t *options, idx_t *objval, idx_t *part);
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                        idx_t *vsize, idx_t *adjwgt, idx_t *nparts, real_t *tpwgts, real_t *ubvec,
                        idx_t *options, idx_t *objval, idx_t *part);
#ifdef __cplusplus
}
#endif
#endif
