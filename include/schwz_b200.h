/* =============================================================================
 * schwz_b200.h — C ABI of the B200-native restricted additive Schwarz (RAS)
 * hot path.  Plain pointers and sizes only; every entry point returns 0 on
 * success and a non-zero code on failure (text via schwz_b200_last_error()).
 *
 * Each group cites the interface of pratikvn/schwarz-lib it replaces
 * (file:line relative to the reference tree).  The reference binds nothing
 * through an FFI — it is one C++ library — so "what the reference would bind"
 * is the set of calls its SolverRAS / Solve / Communicate classes make into
 * Ginkgo, MPI and its two CUDA kernels on this path.  INTEGRATION.md shows the
 * reference-side stubs.
 *
 * Conventions: ValueType = double, IndexType = int32_t
 * (benchmarking/bench_ras.cpp:204).  "dev" pointers are device addresses valid
 * in the calling process; "host" pointers are ordinary host memory.  All
 * device work is enqueued on the context's stream; calls are asynchronous
 * unless documented otherwise.
 * ========================================================================== */
#ifndef SCHWZ_B200_H
#define SCHWZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct schwz_ctx schwz_ctx;         /* one device + one stream            */
typedef struct schwz_csr schwz_csr;         /* device CSR matrix                  */
typedef struct schwz_cg schwz_cg;           /* device-resident CG workspace       */
typedef struct schwz_gmres schwz_gmres;     /* device-resident GMRES(m) workspace */
typedef struct schwz_trs schwz_trs;         /* level-scheduled triangular solve   */
typedef struct schwz_precond schwz_precond; /* local preconditioner of CG / GMRES */
typedef struct schwz_setup schwz_setup;     /* host index sets of one problem     */
typedef struct schwz_ras schwz_ras;         /* one subdomain of the RAS iteration */
typedef struct schwz_comm schwz_comm;       /* NCCL communicator (residual-norm allgather only) */

const char *schwz_b200_last_error(void);
int schwz_b200_version(void);

/* ---- context / memory --------------------------------------------------------
 * replaces: gko::CudaExecutor::create + schwz::device_guard
 * (source/schwarz_base.cpp:93-116, include/device_guard.hpp:64-98). */
int schwz_b200_device_count(int *count);
int schwz_b200_ctx_create(int device, schwz_ctx **out);
int schwz_b200_ctx_destroy(schwz_ctx *ctx);
int schwz_b200_ctx_stream(schwz_ctx *ctx, void **cuda_stream);
int schwz_b200_ctx_sync(schwz_ctx *ctx);
int schwz_b200_malloc(schwz_ctx *ctx, size_t bytes, void **dev);
int schwz_b200_free(schwz_ctx *ctx, void *dev);
int schwz_b200_memset(schwz_ctx *ctx, void *dev, int byte, size_t bytes);
int schwz_b200_h2d(schwz_ctx *ctx, void *dev, const void *host, size_t bytes);   /* synchronous */
int schwz_b200_d2h(schwz_ctx *ctx, void *host, const void *dev, size_t bytes);   /* synchronous */
int schwz_b200_d2d(schwz_ctx *ctx, void *dst, const void *src, size_t bytes);
/* CUDA IPC for peer-mapped halo mailboxes across processes (one process per
 * GPU); replaces MPI_Win_create (source/restricted_schwarz.cpp:488-499,
 * 556-567, 663-691).  handle = 64 bytes. */
int schwz_b200_ipc_export(schwz_ctx *ctx, void *dev, void *handle64);
int schwz_b200_ipc_import(schwz_ctx *ctx, const void *handle64, void **dev);
int schwz_b200_ipc_close(schwz_ctx *ctx, void *dev);
/* cudaDeviceEnablePeerAccess between every pair of `n` contexts of one
 * process. */
int schwz_b200_enable_peers(schwz_ctx **ctxs, int n);
/* event timing on the context's stream */
int schwz_b200_timer_start(schwz_ctx *ctx);
int schwz_b200_timer_stop(schwz_ctx *ctx, float *ms);   /* synchronises */
/* number of kernels this library has launched in this process */
int64_t schwz_b200_launch_count(void);

/* ---- CSR SpMV ----------------------------------------------------------------
 * replaces: gko::matrix::Csr::apply / advanced apply
 * (source/restricted_schwarz.cpp:1014, source/solve.cpp:834, 1019, 1079 and
 * every SpMV inside gko::solver::Cg/Gmres, source/solve.cpp:753). */
int schwz_b200_csr_upload(schwz_ctx *ctx, int32_t n_rows, int32_t n_cols,
                          const int32_t *host_rowptr, const int32_t *host_col,
                          const double *host_val, schwz_csr **out);
int schwz_b200_csr_destroy(schwz_csr *A);
/* y = alpha*A*x + beta*y; rows summed sequentially in stored column order */
int schwz_b200_spmv(schwz_ctx *ctx, const schwz_csr *A, double alpha,
                    const double *dev_x, double beta, double *dev_y);
/* algorithmic bytes of one SpMV launch */
int64_t schwz_b200_spmv_bytes(const schwz_csr *A, int beta_nonzero);

/* ---- BLAS-1 (device scalars; used by tests and the residual check) ----------
 * replaces: gko::matrix::Dense::compute_norm2 / compute_dot / add_scaled
 * (source/solve.cpp:841, 1069-1082; source/communicate.cpp:89-92). */
int schwz_b200_dot(schwz_ctx *ctx, int64_t n, const double *dev_a,
                   const double *dev_b, double *host_result);   /* synchronises */
int schwz_b200_nrm2(schwz_ctx *ctx, int64_t n, const double *dev_a,
                    double *host_result);                        /* synchronises */
int schwz_b200_axpy(schwz_ctx *ctx, int64_t n, double alpha, const double *dev_x,
                    double *dev_y);

/* ---- gather / scatter --------------------------------------------------------
 * replaces: schwz::Gather / schwz::Scatter gko::Operations and
 * source/gather_kernel.cu:46-109, source/scatter_kernel.cu:43-107.
 * op: 0 add, 1 copy, 2 diff, 3 avg (include/collective_common.hpp:37);
 * semantics of the reference's OpenMP path (include/gather.hpp:86-107,
 * include/scatter.hpp:86-108). */
int schwz_b200_gather(schwz_ctx *ctx, int32_t n, const int32_t *dev_idx,
                      const double *dev_from, double *dev_into, int op);
int schwz_b200_scatter(schwz_ctx *ctx, int32_t n, const int32_t *dev_idx,
                       const double *dev_from, double *dev_into, int op);

/* ---- local iterative solves --------------------------------------------------
 * replaces: gko::solver::Cg / Gmres with Combined(Iteration, ResidualNorm-
 * Reduction) (source/solve.cpp:469-478, 486-652, 746-754;
 * include/solver_tools.hpp:91-98).  x is the warm start on entry. */
int schwz_b200_cg_create(schwz_ctx *ctx, const schwz_csr *A, schwz_cg **out);
int schwz_b200_cg_destroy(schwz_cg *cg);
int schwz_b200_cg_solve(schwz_cg *cg, const double *dev_b, double *dev_x,
                        int32_t max_iters, double rel_tol);      /* asynchronous */
int schwz_b200_cg_result(schwz_cg *cg, int32_t *iters, double *resnorm,
                         double *resnorm0);                      /* synchronises */
int schwz_b200_gmres_create(schwz_ctx *ctx, const schwz_csr *A, int32_t restart,
                            schwz_gmres **out);
int schwz_b200_gmres_destroy(schwz_gmres *g);
int schwz_b200_gmres_solve(schwz_gmres *g, const double *dev_b, double *dev_x,
                           int32_t max_iters, double rel_tol);
int schwz_b200_gmres_result(schwz_gmres *g, int32_t *iters, double *resnorm,
                            double *resnorm0);

/* ---- local preconditioners of the iterative local solve ------------------------
 * replaces: gko::preconditioner::Jacobi(max_block_size), gko::factorization::ParIlu +
 * gko::preconditioner::Ilu<LowerTrs, UpperTrs>, gko::preconditioner::Ilu<LowerIsai,
 * UpperIsai> as built at source/solve.cpp:496-505, 513-532, 540-556 (GMRES) and
 * :581-589, 598-617, 625-638 (CG); metadata.local_precond / precond_max_block_size
 * (include/settings.hpp, --local_precond / --precond_max_block_size of bench_ras).
 * kind: 1 block-jacobi, 2 ilu, 3 isai.  Generated on the host from the host CSR of the
 * local matrix (setup), applied on the device (every Krylov iteration).  apply: z = M^-1 r,
 * optionally *dev_dot = r.z.  The getters return host copies of what was generated
 * (block pointers; inverse blocks column-major; which = 0 L, 1 U, 2 ISAI(L), 3 ISAI(U));
 * pass NULL to query the length.  ctx == NULL generates a host-only handle (no device, no
 * apply) - used to check the generation without a GPU. */
enum { SCHWZ_PRECOND_NONE = 0, SCHWZ_PRECOND_BLOCK_JACOBI = 1, SCHWZ_PRECOND_ILU = 2,
       SCHWZ_PRECOND_ISAI = 3 };
int schwz_b200_precond_create(schwz_ctx *ctx, int32_t n, const int32_t *host_rowptr,
                              const int32_t *host_col, const double *host_val, int32_t kind,
                              int32_t max_block_size, schwz_precond **out);
int schwz_b200_precond_destroy(schwz_precond *p);
int schwz_b200_precond_apply(schwz_precond *p, const double *dev_r, double *dev_z,
                             double *dev_dot_or_null);
int64_t schwz_b200_precond_bytes_per_apply(schwz_precond *p);
int64_t schwz_b200_precond_block_ptrs(schwz_precond *p, int32_t *host_out);
int64_t schwz_b200_precond_blocks(schwz_precond *p, double *host_out);
int64_t schwz_b200_precond_csr(schwz_precond *p, int32_t which, int32_t *host_rowptr,
                               int32_t *host_col, double *host_val);
/* with_preconditioner / with_generated_preconditioner of the solver factories; NULL detaches.
 * The preconditioner must outlive the solver. */
int schwz_b200_cg_set_precond(schwz_cg *cg, schwz_precond *p);
int schwz_b200_gmres_set_precond(schwz_gmres *g, schwz_precond *p);

/* ---- factorised direct variant ----------------------------------------------
 * replaces: gko::solver::LowerTrs / UpperTrs generate+apply and
 * gko::matrix::Permutation::apply (source/solve.cpp:391-399, 717-720;
 * include/solver_tools.hpp:69-87); host factorisation replaces CHOLMOD
 * (source/solve.cpp:75-143). */
int schwz_b200_trs_analyze(schwz_ctx *ctx, int32_t n, const int32_t *host_rowptr,
                           const int32_t *host_col, const double *host_val,
                           int upper, schwz_trs **out);
int schwz_b200_trs_destroy(schwz_trs *t);
int schwz_b200_trs_solve(schwz_trs *t, const double *dev_b, double *dev_x);
int schwz_b200_trs_levels(const schwz_trs *t, int32_t *num_levels);
/* synchronises; non-zero: a dependency wait of the one-kernel solve ran into its time bound
 * (the solve is then void).  SCHWZ_B200_TRS_LEVELS=1 selects the level-per-launch graph. */
int schwz_b200_trs_error(schwz_trs *t, int32_t *err);
/* out[i] = in[perm[i]] (inverse == 0) or out[perm[i]] = in[i] (inverse != 0) */
int schwz_b200_permute(schwz_ctx *ctx, int32_t n, const int32_t *dev_perm,
                       int inverse, const double *dev_in, double *dev_out);
/* host: simplicial LL^T of P A P^T; perm may be NULL (natural order).  Two-call
 * protocol: pass L arrays == NULL to get nnz(L).  Returns nnz(L) or <0. */
int64_t schwz_b200_host_cholesky(int32_t n, const int32_t *rowptr,
                                 const int32_t *col, const double *val,
                                 const int32_t *perm, int32_t *L_rowptr,
                                 int32_t *L_col, double *L_val);
/* host: fill-reducing ordering (METIS_NodeND of the toolkit's static METIS) of the pattern
 * of A + A^T */
int schwz_b200_host_nd_ordering(int32_t n, const int32_t *rowptr,
                                const int32_t *col, int32_t *perm);

/* host: sparse LU with threshold partial pivoting, P A Q = L U, L unit lower / U upper CSR with
 * sorted rows; replaces umfpack_di_symbolic / _numeric / _get_numeric of the UMFPACK branch
 * (source/solve.cpp:145-171, 322-385; --local_factorization=umfpack).  col_perm = Q (NULL =
 * natural), the row order P is found (row_perm[k] = row of A that became pivot row k); the
 * diagonal entry is preferred as pivot while |a_diag| >= diag_pivot_tol * max|column| (UMFPACK's
 * symmetric strategy uses 1e-3).  No row scaling (the reference fetches UMFPACK's and never
 * applies it). */
typedef struct schwz_lu schwz_lu;
int schwz_b200_host_lu_create(int32_t n, const int32_t *rowptr, const int32_t *col,
                              const double *val, const int32_t *col_perm, double diag_pivot_tol,
                              schwz_lu **out);
int schwz_b200_host_lu_destroy(schwz_lu *lu);
int schwz_b200_host_lu_nnz(const schwz_lu *lu, int64_t *nnz_l, int64_t *nnz_u);
int schwz_b200_host_lu_get(const schwz_lu *lu, int32_t *L_rowptr, int32_t *L_col, double *L_val,
                           int32_t *U_rowptr, int32_t *U_col, double *U_val, int32_t *row_perm);

/* ---- host index sets ---------------------------------------------------------
 * replaces: Initialize::setup_global_matrix / partition
 * (source/initialization.cpp:197-329), PartitionTools
 * (include/partition_tools.hpp:59-202), SolverRAS::setup_local_matrices /
 * setup_comm_buffers / setup_windows (source/restricted_schwarz.cpp:56-711).
 * Results are bit-identical to the reference's arrays. */
int64_t schwz_b200_laplacian2d(int32_t n, int32_t *rowptr, int32_t *col, double *val);
int64_t schwz_b200_laplacian3d(int32_t n, int32_t *rowptr, int32_t *col, double *val);
int schwz_b200_read_mtx(const char *path, int32_t *n_rows, int64_t *nnz,
                        int32_t **rowptr, int32_t **col, double **val);  /* free with schwz_b200_host_free */
void schwz_b200_host_free(void *p);
int schwz_b200_partition_regular2d(int64_t N, int32_t P, uint32_t *part);
/* px x py rectangular extension of PartitionRegular2D (include/partition_tools.hpp:70-94) for
 * subdomain counts that are not perfect squares (8 -> 2 x 4); px = py = 0 picks the most square
 * factorisation.  Identical to the rule above for perfect squares. */
int schwz_b200_partition_regular2d_rect(int64_t N, int32_t P, int32_t px, int32_t py,
                                        uint32_t *part);
int schwz_b200_partition_metis(int32_t N, const int32_t *rowptr, const int32_t *col,
                               int32_t P, const char *objtype, uint32_t *part);

/* matrix_kind: 0 stored CSR (rowptr/col/val, N rows), 1 generated 2-D 5-pt
 * Laplacian (N = n*n, arrays ignored), 2 generated 3-D 7-pt Laplacian.
 * partition_kind: 0 regular (1-D split), 1 permute by `part` (metis/regular2d). */
int schwz_b200_setup_create(int32_t matrix_kind, int32_t grid_n, int32_t N,
                            const int32_t *rowptr, const int32_t *col,
                            const double *val, int32_t P, int32_t partition_kind,
                            const uint32_t *part, int32_t overlap,
                            schwz_setup **out);
int schwz_b200_setup_destroy(schwz_setup *s);
int schwz_b200_setup_first_row(const schwz_setup *s, int32_t *out /* P+1 */);
int schwz_b200_setup_permutation(const schwz_setup *s, int32_t *perm, int32_t *iperm);
/* out[8]: local_size, local_size_x, overlap_size, nnz_local, nnz_interface,
 * n_halo, num_neighbors_in, num_neighbors_out */
int schwz_b200_setup_sizes(schwz_setup *s, int32_t rank, int64_t *out);
int schwz_b200_setup_l2g(schwz_setup *s, int32_t rank, int32_t *out);
int schwz_b200_setup_local_matrix(schwz_setup *s, int32_t rank, int32_t *rowptr,
                                  int32_t *col, double *val);
/* interface matrix with GLOBAL column indices (reference layout) */
int schwz_b200_setup_interface_matrix(schwz_setup *s, int32_t rank, int32_t *rowptr,
                                      int32_t *col, double *val);
int schwz_b200_setup_neighbors(schwz_setup *s, int32_t rank, int32_t *nbr_in,
                               int32_t *nbr_out);
int schwz_b200_setup_get_list(schwz_setup *s, int32_t rank, int32_t j, int32_t *out);
int schwz_b200_setup_get_count(schwz_setup *s, int32_t rank, int32_t j, int32_t *count);
int schwz_b200_setup_put_list(schwz_setup *s, int32_t rank, int32_t j, int32_t *out);
int schwz_b200_setup_put_count(schwz_setup *s, int32_t rank, int32_t j, int32_t *count);
int schwz_b200_setup_displacements(schwz_setup *s, int32_t rank, int32_t *put_disp,
                                   int32_t *get_disp /* P+1 each */);
/* drop the cached host arrays of a rank once it has been uploaded */
int schwz_b200_setup_release_rank(schwz_setup *s, int32_t rank);

/* ---- one RAS subdomain on a device ------------------------------------------
 * replaces: the per-rank state and loop stages of SchwarzBase::run
 * (source/schwarz_base.cpp:323-506): exchange_boundary
 * (source/restricted_schwarz.cpp:715-988, include/comm_helpers.hpp:58-177),
 * update_boundary (:992-1017), check_local_convergence
 * (source/solve.cpp:796-856), local_solve (:667-792), local_to_global_vector
 * (source/communicate.cpp:65-94). */
typedef struct {
    double tolerance;          /* metadata.tolerance            */
    double local_tol;          /* metadata.local_solver_tolerance */
    int32_t local_max_iters;   /* -1 => local_size_x             */
    int32_t local_solver;      /* 2 iterative (CG/GMRES), 1 direct (TRS) */
    int32_t non_symmetric;     /* GMRES instead of CG            */
    int32_t restart_iter;
    int32_t overlap;
    int32_t use_mixed_precision; /* settings.use_mixed_precision, MixedValueType = float: halo
                                  * values travel as floats (restricted_schwarz.cpp:483-603) */
    int32_t local_precond;       /* metadata.local_precond: SCHWZ_PRECOND_* (iterative only) */
    int32_t precond_max_block_size; /* metadata.precond_max_block_size (block-jacobi)        */
} schwz_ras_options;

int schwz_b200_ras_create(schwz_ctx *ctx, schwz_setup *s, int32_t rank,
                          const double *host_rhs_global /* N, permuted numbering; NULL = ones */,
                          const schwz_ras_options *opt, schwz_ras **out);
int schwz_b200_ras_destroy(schwz_ras *r);
/* direct variant: upload host factors (L lower CSR incl. diagonal, U = L^T) and
 * the ordering; analysis builds the level sets */
int schwz_b200_ras_set_factors(schwz_ras *r, const int32_t *L_rowptr,
                               const int32_t *L_col, const double *L_val,
                               const int32_t *perm);
/* the cap of the iterative local solve from now on (-1 = local_size_x): what re-building the
 * stopping criterion with metadata.updated_max_iters past settings.reset_local_crit_iter does
 * (source/solve.cpp:721-741) */
int schwz_b200_ras_set_local_max_iters(schwz_ras *r, int32_t local_max_iters);
/* direct variant, unsymmetric: the factors of schwz_b200_host_lu_create and the column order it
 * was given; local solve = Q U^-1 L^-1 P b (local_perm = P, local_inv_perm = Q of
 * source/solve.cpp:342-353) */
int schwz_b200_ras_set_lu_factors(schwz_ras *r, const schwz_lu *lu, const int32_t *col_perm);
/* mailbox = peer-visible block holding the two receive buffers (epoch
 * parity), the epoch flags and the convergence flags of a subdomain; this is
 * what replaces the MPI windows.  The layout travels with the base pointer. */
typedef struct {
    int64_t recv_stride;   /* bytes between the two receive buffers */
    int64_t flags_off;     /* u64 epoch word per in-neighbour       */
    int64_t conv_off;      /* P int32 convergence flags             */
    int64_t err_off;       /* int32 error word                      */
    int64_t send_off;      /* send buffer, sum(out-list lengths) doubles: window_send_buffer
                            * of the reference (restricted_schwarz.cpp:549-568)            */
    int64_t slots_off;     /* int32 per received element: its slot in x (local_get + 1)    */
    int64_t x_off;         /* x = [own | overlap | halo] itself: window_x (:661-667)       */
    int64_t bytes;
} schwz_mailbox_layout;
int schwz_b200_ras_mailbox(schwz_ras *r, void **dev_base, schwz_mailbox_layout *layout);
/* host only: the layout implied by the index-set sizes (in_total / out_total = sum of
 * the in- / out-list lengths, n_in = num_neighbors_in, x_len = local_size_x + n_halo) */
int schwz_b200_mailbox_layout(int64_t in_total, int32_t n_in, int32_t P, int64_t out_total,
                              int64_t x_len, schwz_mailbox_layout *layout);
/* sizes a peer needs: out[0] = num_neighbors_in, out[1] = num_neighbors_out,
 * out[2] = local_size, out[3] = local_size_x, out[4] = n_halo, out[5] = nnz_local */
int schwz_b200_ras_info(schwz_ras *r, int64_t *out);
int schwz_b200_ras_neighbors(schwz_ras *r, int32_t *nbr_in, int32_t *nbr_out);
/* tell subdomain r where out-neighbour j's mailbox lives (pointer valid in
 * this process: the neighbour's own base when it is local, an IPC-imported
 * base otherwise).  peer_recv_offset_elems = put_displacements[neighbour]
 * (source/restricted_schwarz.cpp:624-640); peer_flag_slot = my position in
 * the neighbour's neighbors_in list.  same_process != 0: ordering by stream
 * events; == 0: ordering by the device epoch flags. */
int schwz_b200_ras_connect(schwz_ras *r, int32_t j_out, void *peer_mailbox_base,
                           const schwz_mailbox_layout *peer_layout,
                           int32_t peer_recv_offset_elems, int32_t peer_flag_slot,
                           int32_t same_process);
/* Get variants: tell subdomain r where in-neighbour j's mailbox lives;
 * peer_send_offset_elems = get_displacements[neighbour]
 * (source/restricted_schwarz.cpp:642-658). */
int schwz_b200_ras_connect_in(schwz_ras *r, int32_t j_in, void *peer_mailbox_base,
                              const schwz_mailbox_layout *peer_layout,
                              int32_t peer_send_offset_elems);
/* centralised-tree convergence: parent and children need not be halo neighbours */
int schwz_b200_ras_connect_conv(schwz_ras *r, int32_t peer_rank, void *peer_mailbox_base,
                                const schwz_mailbox_layout *peer_layout);
/* Settings::comm_settings enable_put/enable_get x enable_one_by_one
 * (source/restricted_schwarz.cpp:753-851): 0 Put gathered (also the synchronous
 * exchange), 1 Get gathered, 2 Put one-by-one, 3 Get one-by-one. */
int schwz_b200_ras_set_exchange_mode(schwz_ras *r, int32_t mode);
/* 1: one-sided semantics for the Put-gathered exchange (comm_settings.enable_onesided,
 * source/restricted_schwarz.cpp:715-852): a single receive buffer, no epoch flags, the receiver
 * scatters whatever it holds; stale values of a run before schwz_b200_ras_reset are cleared.
 * 0 (default): synchronous epochs (double-buffered, flag per in-neighbour). */
int schwz_b200_ras_set_onesided(schwz_ras *r, int32_t onesided);
/* connects every pair of subdomains living in this process (out, in and conv) */
int schwz_b200_ras_connect_local(schwz_ras **subdomains, int32_t n_local, schwz_setup *s);
/* loop stages (asynchronous on the subdomain's stream) */
int schwz_b200_ras_exchange_push(schwz_ras *r, int32_t iter);
int schwz_b200_ras_exchange_unpack(schwz_ras *r, int32_t iter, int32_t wait_flags);
int schwz_b200_ras_update_boundary(schwz_ras *r);
int schwz_b200_ras_local_residual(schwz_ras *r);              /* -> device scalar */
int schwz_b200_ras_residual_norm(schwz_ras *r, double *host_norm); /* synchronises */
int schwz_b200_ras_residual_norm_dev(schwz_ras *r, double **dev_norm);
int schwz_b200_ras_local_solve(schwz_ras *r);
int schwz_b200_ras_restrict(schwz_ras *r);
int schwz_b200_ras_last_local_iters(schwz_ras *r, int32_t *iters); /* synchronises */
/* event recorded after the latest push / waited on before unpack (same-process
 * neighbours) */
int schwz_b200_ras_wait_push_of(schwz_ras *r, schwz_ras *neighbour);
int schwz_b200_ras_sync(schwz_ras *r);
/* state access for parity tests (synchronous copies) */
int schwz_b200_ras_get_x(schwz_ras *r, double *host_out /* local_size_x + n_halo */);
int schwz_b200_ras_get_local_solution(schwz_ras *r, double *host_out /* local_size_x */);
int schwz_b200_ras_set_x_own(schwz_ras *r, const double *host_in /* local_size */);
/* host-buffer path of the plugin call (SolverRAS::initialize uploads the rhs,
 * source/initialization.cpp:345-355; SchwarzBase::run returns the solution,
 * source/schwarz_base.cpp:501-503).  rhs / solution are length-N host vectors
 * in the (permuted) global numbering; copies are asynchronous on the
 * subdomain's stream (schwz_b200_ras_sync completes them). */
int schwz_b200_ras_upload_rhs(schwz_ras *r, const double *host_rhs_global);
int schwz_b200_ras_download_solution(schwz_ras *r, double *host_solution_global);
int schwz_b200_ras_reset(schwz_ras *r);   /* x, init_guess <- 0, norms unlatched */
/* measurement aid: average duration of one launch of a hot kernel on this
 * subdomain's data (CUDA events on its stream) and its algorithmic bytes.
 * kind: 0 SpMV+dot of CG, 1 CG r update, 2 CG x/p update, 3 residual
 * SpMV+norm, 4 halo push+unpack, 5 push only, 6 unpack only (call 5 and 6 with
 * the same reps: they advance the exchange epoch) */
int schwz_b200_ras_kernel_time(schwz_ras *r, int32_t kind, int32_t reps, float *ms);
int64_t schwz_b200_ras_kernel_bytes(schwz_ras *r, int32_t kind);
/* distributed true residual ||b_own - (A x)_own||^2 (needs fresh overlap) */
int schwz_b200_ras_true_residual_sq(schwz_ras *r, double *host_out);
/* decentralised convergence flags (include/conv_tools.hpp:213-275) */
int schwz_b200_ras_conv_set_local(schwz_ras *r, int32_t converged_all_local);
int schwz_b200_ras_conv_forward(schwz_ras *r);
/* accumulate variant (include/conv_tools.hpp:230-247, --enable_decentralized_accumulate): adds 1
 * to word 0 of every subdomain's flags while locally converged; needs every subdomain's flags
 * connected (schwz_b200_ras_connect_conv / _connect_local); conv_count returns word 0 */
int schwz_b200_ras_conv_accumulate(schwz_ras *r, int32_t converged_all_local);
/* centralised binary tree (include/conv_tools.hpp:147-209): push up to the parent once the
 * children have, the root pushes down; conv_count then returns P or 0 */
int schwz_b200_ras_conv_tree(schwz_ras *r, int32_t converged_all_local);
int schwz_b200_ras_conv_count(schwz_ras *r, int32_t *num_converged);  /* synchronises */

/* ---- whole outer loop over the subdomains of this process --------------------
 * replaces: the for-loop of SchwarzBase::run (source/schwarz_base.cpp:387-452)
 * and Solve::check_global_convergence (source/solve.cpp:860-955). */
typedef struct {
    int32_t num_subdomains;        /* P, all processes                    */
    int32_t max_iters;
    double tolerance;
    int32_t enable_onesided;       /* async: no waits, decentralised flags */
    int32_t enable_global_check;
    int32_t conv_decentralized;    /* 0 centralised tree, 1 flag flooding, 2 accumulate */
    int32_t iter_offset;
    int32_t exchange_mode;         /* one-sided only, see schwz_b200_ras_set_exchange_mode */
    /* cross-process allgather of the residual norms (ncclAllGather); NULL when
     * every subdomain lives in this process.  Subdomain ids must be spread
     * contiguously and evenly: process r owns ids [r*n_local, (r+1)*n_local). */
    schwz_comm *comm;
} schwz_loop_options;

/* The loop runs ahead of the host: stages are enqueued a few outer iterations deep, the
 * convergence decision is taken on the device and turns whatever was enqueued past the break
 * point into no-ops; the host looks at the loop state once per chunk of iterations
 * (SCHWZ_B200_OUTER_CHUNK, default 4).  A halo wait that exceeds SCHWZ_B200_HALO_TIMEOUT_MS
 * (default 20000) makes the call fail with "halo exchange timed out". */
typedef struct {
    int32_t iters;                 /* metadata.iter_count at exit */
    int32_t converged;
    double global_resnorm, global_resnorm0;
    double elapsed_s;              /* steady_clock window of the reference */
    /* host-blocking CUDA calls the loop made: stream synchronisations (all of them before the
     * first / after the last iteration is enqueued) and event waits (one per subdomain per chunk
     * of iterations, on a snapshot two chunks old) */
    int32_t host_stream_syncs, host_event_waits;
} schwz_loop_result;

int schwz_b200_ras_run(schwz_ras **subdomains, int32_t n_local,
                       const schwz_loop_options *opt, schwz_loop_result *res,
                       double *host_resnorm_history /* max_iters*n_local or NULL */);

/* One synchronous halo exchange outside the loop (collective over all processes): afterwards
 * the overlap / halo entries of every x hold the neighbours' current values - what the final
 * residual of Solve::compute_residual_norm (source/solve.cpp:1025-1085) needs when the loop
 * ended on its iteration budget. */
int schwz_b200_ras_refresh_halo(schwz_ras **subdomains, int32_t n_local, int32_t num_subdomains);

/* ---- NCCL communicator for the residual-norm allgather ----------------------
 * replaces: MPI_Allgather (source/solve.cpp:890-891).  id128 is an
 * ncclUniqueId created on one process and distributed by the launcher
 * (the process-group plumbing of the host program). */
int schwz_b200_comm_unique_id(void *id128);
int schwz_b200_comm_create(schwz_ctx *ctx, const void *id128, int32_t nranks,
                           int32_t rank, schwz_comm **out);
int schwz_b200_comm_destroy(schwz_comm *c);
int schwz_b200_comm_allgather_f64(schwz_comm *c, const double *dev_in, int32_t count,
                                  double *dev_out);

#ifdef __cplusplus
}
#endif
#endif /* SCHWZ_B200_H */
