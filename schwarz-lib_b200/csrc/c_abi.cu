// extern "C" boundary of libschwz_b200.so (include/schwz_b200.h).  Exceptions
// become status codes; the text is kept per thread for
// schwz_b200_last_error().
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/schwz_b200.h"
#include "engine.hpp"

using namespace schwz_b200;

struct schwz_ctx { Ctx impl; explicit schwz_ctx(int d) : impl(d) {} };
struct schwz_csr { std::unique_ptr<DeviceCsr> impl; };
struct schwz_cg { std::unique_ptr<CgSolver> impl; };
struct schwz_gmres { std::unique_ptr<GmresSolver> impl; };
struct schwz_precond {
    std::unique_ptr<Preconditioner> impl;   // null for a host-only handle (ctx == NULL)
    PrecondData host_only;
    const PrecondData &data() const { return impl ? impl->host : host_only; }
};
struct schwz_trs { std::unique_ptr<TrsPlan> impl; };
struct schwz_setup { std::unique_ptr<Setup> impl; };
struct schwz_ras { std::unique_ptr<Ras> impl; };
struct schwz_comm { std::unique_ptr<Comm> impl; };

static thread_local std::string g_err;

#define ABI_BEGIN try {
#define ABI_END                         \
    return 0;                           \
    }                                   \
    catch (const CudaFailure &e) {      \
        g_err = e.what();               \
        return 2;                       \
    }                                   \
    catch (const std::exception &e) {   \
        g_err = e.what();               \
        return 1;                       \
    }                                   \
    catch (...) {                       \
        g_err = "unknown error";        \
        return 3;                       \
    }

extern "C" {

const char *schwz_b200_last_error(void) { return g_err.c_str(); }
int schwz_b200_version(void) { return 100; }

// ---- context / memory ---------------------------------------------------------
int schwz_b200_device_count(int *count)
{
    ABI_BEGIN
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    ABI_END
}
int schwz_b200_ctx_create(int device, schwz_ctx **out)
{
    ABI_BEGIN
    const char *e = std::getenv("SCHWZ_B200_SIMPLE_SPMV");
    g_force_simple_spmv = e && e[0] == '1';
    const char *vr = std::getenv("SCHWZ_B200_SPMV_VARIANT");
    if (vr) g_spmv_variant = std::atoi(vr);
    const char *ns = std::getenv("SCHWZ_B200_NO_SMALL");
    g_use_small_solvers = !(ns && ns[0] == '1');
    const char *ng = std::getenv("SCHWZ_B200_NO_CG_GRAPH");
    g_use_cg_graph = !(ng && ng[0] == '1');
    const char *c32 = std::getenv("SCHWZ_B200_SPMV_COL32");
    g_spmv_col16 = !(c32 && c32[0] == '1');
    const char *cpdl = std::getenv("SCHWZ_B200_CG_PDL");
    g_cg_pdl = cpdl && cpdl[0] == '1';
    const char *npdl = std::getenv("SCHWZ_B200_TRS_NO_PDL");
    g_trs_pdl = !(npdl && npdl[0] == '1');
    const char *mgs = std::getenv("SCHWZ_B200_GMRES_MGS");
    g_gmres_cgs2 = !(mgs && mgs[0] == '1');
    const char *ht = std::getenv("SCHWZ_B200_HALO_TIMEOUT_MS");
    if (ht && std::atoll(ht) > 0) g_halo_timeout_ns = std::atoll(ht) * 1000000ll;
    *out = new schwz_ctx(device);
    ABI_END
}
int schwz_b200_ctx_destroy(schwz_ctx *ctx)
{
    ABI_BEGIN
    delete ctx;
    ABI_END
}
int schwz_b200_ctx_stream(schwz_ctx *ctx, void **s)
{
    ABI_BEGIN
    *s = (void *)ctx->impl.stream;
    ABI_END
}
int schwz_b200_ctx_sync(schwz_ctx *ctx)
{
    ABI_BEGIN
    ctx->impl.sync();
    ABI_END
}
int schwz_b200_malloc(schwz_ctx *ctx, size_t bytes, void **dev)
{
    ABI_BEGIN
    *dev = ctx->impl.alloc<char>(bytes);
    ABI_END
}
int schwz_b200_free(schwz_ctx *ctx, void *dev)
{
    ABI_BEGIN
    ctx->impl.release(dev);
    ABI_END
}
int schwz_b200_memset(schwz_ctx *ctx, void *dev, int byte, size_t bytes)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaMemsetAsync(dev, byte, bytes, ctx->impl.stream));
    ABI_END
}
int schwz_b200_h2d(schwz_ctx *ctx, void *dev, const void *host, size_t bytes)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->impl.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx->impl.stream));
    ABI_END
}
int schwz_b200_d2h(schwz_ctx *ctx, void *host, const void *dev, size_t bytes)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->impl.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx->impl.stream));
    ABI_END
}
int schwz_b200_d2d(schwz_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->impl.stream));
    ABI_END
}
int schwz_b200_ipc_export(schwz_ctx *ctx, void *dev, void *handle64)
{
    ABI_BEGIN
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ctx->impl.use();
    cudaIpcMemHandle_t h;
    SCHWZ_CUDA(cudaIpcGetMemHandle(&h, dev));
    std::memcpy(handle64, &h, 64);
    ABI_END
}
int schwz_b200_ipc_import(schwz_ctx *ctx, const void *handle64, void **dev)
{
    ABI_BEGIN
    ctx->impl.use();
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    SCHWZ_CUDA(cudaIpcOpenMemHandle(dev, h, cudaIpcMemLazyEnablePeerAccess));
    ABI_END
}
int schwz_b200_ipc_close(schwz_ctx *ctx, void *dev)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaIpcCloseMemHandle(dev));
    ABI_END
}
int schwz_b200_enable_peers(schwz_ctx **ctxs, int n)
{
    ABI_BEGIN
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const int a = ctxs[i]->impl.device, b = ctxs[j]->impl.device;
            if (a == b) continue;
            int can = 0;
            SCHWZ_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
            SCHWZ_REQUIRE(can, "devices cannot access each other");
            SCHWZ_CUDA(cudaSetDevice(a));
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else SCHWZ_CUDA(e);
        }
    ABI_END
}
int schwz_b200_timer_start(schwz_ctx *ctx)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaEventRecord(ctx->impl.ev_start, ctx->impl.stream));
    ABI_END
}
int schwz_b200_timer_stop(schwz_ctx *ctx, float *ms)
{
    ABI_BEGIN
    ctx->impl.use();
    SCHWZ_CUDA(cudaEventRecord(ctx->impl.ev_stop, ctx->impl.stream));
    SCHWZ_CUDA(cudaEventSynchronize(ctx->impl.ev_stop));
    SCHWZ_CUDA(cudaEventElapsedTime(ms, ctx->impl.ev_start, ctx->impl.ev_stop));
    ABI_END
}
int64_t schwz_b200_launch_count(void) { return g_launches.load(); }

// ---- CSR ----------------------------------------------------------------------
int schwz_b200_csr_upload(schwz_ctx *ctx, int32_t n_rows, int32_t n_cols, const int32_t *rp,
                          const int32_t *ci, const double *v, schwz_csr **out)
{
    ABI_BEGIN
    auto *h = new schwz_csr();
    h->impl.reset(csr_upload(ctx->impl, n_rows, n_cols, rp, ci, v));
    *out = h;
    ABI_END
}
int schwz_b200_csr_destroy(schwz_csr *A)
{
    ABI_BEGIN
    delete A;
    ABI_END
}
int schwz_b200_spmv(schwz_ctx *ctx, const schwz_csr *A, double alpha, const double *x, double beta,
                    double *y)
{
    ABI_BEGIN
    launch_spmv(ctx->impl, *A->impl, alpha, x, beta, y, y, EPI_NONE, nullptr, nullptr, 0, nullptr);
    ABI_END
}
int64_t schwz_b200_spmv_bytes(const schwz_csr *A, int beta_nonzero)
{
    const DeviceCsr &M = *A->impl;
    return 12 * M.nnz + 4 * ((int64_t)M.nrows + 1) + 8 * (int64_t)M.nrows + 8 * (int64_t)M.ncols +
           (beta_nonzero ? 8 * (int64_t)M.nrows : 0);
}

// ---- BLAS-1 -------------------------------------------------------------------
int schwz_b200_dot(schwz_ctx *ctx, int64_t n, const double *a, const double *b, double *out)
{
    ABI_BEGIN
    launch_dot(ctx->impl, n, a, b, ctx->impl.dev_scalars, false);
    SCHWZ_CUDA(cudaMemcpyAsync(out, ctx->impl.dev_scalars, 8, cudaMemcpyDeviceToHost, ctx->impl.stream));
    ctx->impl.sync();
    ABI_END
}
int schwz_b200_nrm2(schwz_ctx *ctx, int64_t n, const double *a, double *out)
{
    ABI_BEGIN
    launch_dot(ctx->impl, n, a, a, ctx->impl.dev_scalars, true);
    SCHWZ_CUDA(cudaMemcpyAsync(out, ctx->impl.dev_scalars, 8, cudaMemcpyDeviceToHost, ctx->impl.stream));
    ctx->impl.sync();
    ABI_END
}
int schwz_b200_axpy(schwz_ctx *ctx, int64_t n, double alpha, const double *x, double *y)
{
    ABI_BEGIN
    launch_axpy(ctx->impl, n, alpha, x, y);
    ABI_END
}
int schwz_b200_gather(schwz_ctx *ctx, int32_t n, const int32_t *idx, const double *from,
                      double *into, int op)
{
    ABI_BEGIN
    launch_gather(ctx->impl, n, idx, from, into, op);
    ABI_END
}
int schwz_b200_scatter(schwz_ctx *ctx, int32_t n, const int32_t *idx, const double *from,
                       double *into, int op)
{
    ABI_BEGIN
    launch_scatter(ctx->impl, n, idx, from, into, op);
    ABI_END
}

// ---- local solves -------------------------------------------------------------
int schwz_b200_cg_create(schwz_ctx *ctx, const schwz_csr *A, schwz_cg **out)
{
    ABI_BEGIN
    auto *h = new schwz_cg();
    h->impl.reset(new CgSolver(ctx->impl, *A->impl));
    *out = h;
    ABI_END
}
int schwz_b200_cg_destroy(schwz_cg *cg)
{
    ABI_BEGIN
    delete cg;
    ABI_END
}
int schwz_b200_cg_solve(schwz_cg *cg, const double *b, double *x, int32_t max_iters, double tol)
{
    ABI_BEGIN
    cg->impl->solve(b, x, max_iters, tol);
    ABI_END
}
int schwz_b200_cg_result(schwz_cg *cg, int32_t *iters, double *resnorm, double *resnorm0)
{
    ABI_BEGIN
    cg->impl->result(iters, resnorm, resnorm0);
    ABI_END
}
// ---- local preconditioners ------------------------------------------------------
int schwz_b200_precond_create(schwz_ctx *ctx, int32_t n, const int32_t *rp, const int32_t *ci,
                              const double *v, int32_t kind, int32_t max_block_size,
                              schwz_precond **out)
{
    ABI_BEGIN
    auto *h = new schwz_precond();
    if (ctx)
        h->impl.reset(new Preconditioner(ctx->impl, n, rp, ci, v, kind, max_block_size));
    else
        h->host_only.generate(n, rp, ci, v, kind, max_block_size);
    *out = h;
    ABI_END
}
int schwz_b200_precond_destroy(schwz_precond *p)
{
    ABI_BEGIN
    delete p;
    ABI_END
}
int schwz_b200_precond_apply(schwz_precond *p, const double *dev_r, double *dev_z,
                             double *dev_dot_or_null)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(p->impl != nullptr, "host-only preconditioner handle (created without a context)");
    p->impl->apply(dev_r, dev_z, dev_dot_or_null, nullptr);
    ABI_END
}
int64_t schwz_b200_precond_bytes_per_apply(schwz_precond *p)
{
    return p->impl ? p->impl->bytes_per_apply() : 0;
}
int64_t schwz_b200_precond_block_ptrs(schwz_precond *p, int32_t *host_out)
{
    const auto &bp = p->data().block_ptrs;
    if (host_out) std::copy(bp.begin(), bp.end(), host_out);
    return (int64_t)bp.size();
}
int64_t schwz_b200_precond_blocks(schwz_precond *p, double *host_out)
{
    const auto &b = p->data().blocks;
    if (host_out) std::copy(b.begin(), b.end(), host_out);
    return (int64_t)b.size();
}
int64_t schwz_b200_precond_csr(schwz_precond *p, int32_t which, int32_t *rp, int32_t *ci, double *v)
{
    const PrecondData &M = p->data();
    const HostCsr &T = which == 0 ? M.L : which == 1 ? M.U : which == 2 ? M.Li : M.Ui;
    if (rp) {
        std::copy(T.rp.begin(), T.rp.end(), rp);
        std::copy(T.ci.begin(), T.ci.end(), ci);
        std::copy(T.v.begin(), T.v.end(), v);
    }
    return (int64_t)T.ci.size();
}
int schwz_b200_cg_set_precond(schwz_cg *cg, schwz_precond *p)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(!p || p->impl, "host-only preconditioner handle");
    cg->impl->set_precond(p ? p->impl.get() : nullptr);
    ABI_END
}
int schwz_b200_gmres_set_precond(schwz_gmres *g, schwz_precond *p)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(!p || p->impl, "host-only preconditioner handle");
    g->impl->set_precond(p ? p->impl.get() : nullptr);
    ABI_END
}

int schwz_b200_gmres_create(schwz_ctx *ctx, const schwz_csr *A, int32_t restart, schwz_gmres **out)
{
    ABI_BEGIN
    auto *h = new schwz_gmres();
    h->impl.reset(new GmresSolver(ctx->impl, *A->impl, restart));
    *out = h;
    ABI_END
}
int schwz_b200_gmres_destroy(schwz_gmres *g)
{
    ABI_BEGIN
    delete g;
    ABI_END
}
int schwz_b200_gmres_solve(schwz_gmres *g, const double *b, double *x, int32_t max_iters, double tol)
{
    ABI_BEGIN
    g->impl->solve(b, x, max_iters, tol);
    ABI_END
}
int schwz_b200_gmres_result(schwz_gmres *g, int32_t *iters, double *resnorm, double *resnorm0)
{
    ABI_BEGIN
    g->impl->result(iters, resnorm, resnorm0);
    ABI_END
}

// ---- direct variant -----------------------------------------------------------
int schwz_b200_trs_analyze(schwz_ctx *ctx, int32_t n, const int32_t *rp, const int32_t *ci,
                           const double *v, int upper, schwz_trs **out)
{
    ABI_BEGIN
    auto *h = new schwz_trs();
    h->impl.reset(new TrsPlan(ctx->impl, n, rp, ci, v, upper != 0));
    *out = h;
    ABI_END
}
int schwz_b200_trs_destroy(schwz_trs *t)
{
    ABI_BEGIN
    delete t;
    ABI_END
}
int schwz_b200_trs_solve(schwz_trs *t, const double *b, double *x)
{
    ABI_BEGIN
    t->impl->solve(b, x);
    ABI_END
}
int schwz_b200_trs_error(schwz_trs *t, int32_t *err)
{
    ABI_BEGIN
    *err = t->impl->error();
    ABI_END
}
int schwz_b200_trs_levels(const schwz_trs *t, int32_t *num_levels)
{
    ABI_BEGIN
    *num_levels = t->impl->num_levels();
    ABI_END
}
int schwz_b200_permute(schwz_ctx *ctx, int32_t n, const int32_t *perm, int inverse,
                       const double *in, double *out)
{
    ABI_BEGIN
    launch_permute(ctx->impl, n, perm, inverse, in, out);
    ABI_END
}
int64_t schwz_b200_host_cholesky(int32_t n, const int32_t *rp, const int32_t *ci, const double *v,
                                 const int32_t *perm, int32_t *Lrp, int32_t *Lci, double *Lv)
{
    try {
        HostCsr A, L;
        A.nrows = A.ncols = n;
        A.rp.assign(rp, rp + n + 1);
        A.ci.assign(ci, ci + rp[n]);
        A.v.assign(v, v + rp[n]);
        if (!host_cholesky(A, perm, L)) {
            g_err = "matrix is not positive definite";
            return -1;
        }
        if (Lrp) {
            std::copy(L.rp.begin(), L.rp.end(), Lrp);
            std::copy(L.ci.begin(), L.ci.end(), Lci);
            std::copy(L.v.begin(), L.v.end(), Lv);
        }
        return L.nnz();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -2;
    }
}
int schwz_b200_host_nd_ordering(int32_t n, const int32_t *rp, const int32_t *ci, int32_t *perm)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(nd_ordering_symmetrized(n, rp, ci, perm) == 0, "METIS_NodeND failed");
    ABI_END
}
struct schwz_lu {
    HostCsr L, U;
    std::vector<int32_t> p;
};
int schwz_b200_host_lu_create(int32_t n, const int32_t *rp, const int32_t *ci, const double *v,
                              const int32_t *col_perm, double diag_pivot_tol, schwz_lu **out)
{
    ABI_BEGIN
    HostCsr A;
    A.nrows = A.ncols = n;
    A.rp.assign(rp, rp + n + 1);
    A.ci.assign(ci, ci + rp[n]);
    A.v.assign(v, v + rp[n]);
    std::unique_ptr<schwz_lu> h(new schwz_lu());
    SCHWZ_REQUIRE(host_sparse_lu(A, col_perm, diag_pivot_tol, h->L, h->U, h->p),
                  "matrix is singular");
    *out = h.release();
    ABI_END
}
int schwz_b200_host_lu_destroy(schwz_lu *lu)
{
    ABI_BEGIN
    delete lu;
    ABI_END
}
int schwz_b200_host_lu_nnz(const schwz_lu *lu, int64_t *nnz_l, int64_t *nnz_u)
{
    ABI_BEGIN
    *nnz_l = lu->L.nnz();
    *nnz_u = lu->U.nnz();
    ABI_END
}
int schwz_b200_host_lu_get(const schwz_lu *lu, int32_t *Lrp, int32_t *Lci, double *Lv,
                           int32_t *Urp, int32_t *Uci, double *Uv, int32_t *row_perm)
{
    ABI_BEGIN
    std::copy(lu->L.rp.begin(), lu->L.rp.end(), Lrp);
    std::copy(lu->L.ci.begin(), lu->L.ci.end(), Lci);
    std::copy(lu->L.v.begin(), lu->L.v.end(), Lv);
    std::copy(lu->U.rp.begin(), lu->U.rp.end(), Urp);
    std::copy(lu->U.ci.begin(), lu->U.ci.end(), Uci);
    std::copy(lu->U.v.begin(), lu->U.v.end(), Uv);
    std::copy(lu->p.begin(), lu->p.end(), row_perm);
    ABI_END
}

// ---- host index sets ----------------------------------------------------------
static int64_t copy_out(const HostCsr &A, int32_t *rp, int32_t *ci, double *v)
{
    if (rp) std::copy(A.rp.begin(), A.rp.end(), rp);
    if (ci) std::copy(A.ci.begin(), A.ci.end(), ci);
    if (v) std::copy(A.v.begin(), A.v.end(), v);
    return A.nnz();
}
int64_t schwz_b200_laplacian2d(int32_t n, int32_t *rp, int32_t *ci, double *v)
{
    try {
        return copy_out(materialize(*make_laplacian2d(n)), rp, ci, v);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
int64_t schwz_b200_laplacian3d(int32_t n, int32_t *rp, int32_t *ci, double *v)
{
    try {
        return copy_out(materialize(*make_laplacian3d(n)), rp, ci, v);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
int schwz_b200_read_mtx(const char *path, int32_t *n_rows, int64_t *nnz, int32_t **rp, int32_t **ci,
                        double **v)
{
    ABI_BEGIN
    HostCsr A = read_mtx(path);
    *n_rows = A.nrows;
    *nnz = A.nnz();
    *rp = (int32_t *)std::malloc(sizeof(int32_t) * (A.nrows + 1));
    *ci = (int32_t *)std::malloc(sizeof(int32_t) * std::max<int64_t>(A.nnz(), 1));
    *v = (double *)std::malloc(sizeof(double) * std::max<int64_t>(A.nnz(), 1));
    copy_out(A, *rp, *ci, *v);
    ABI_END
}
void schwz_b200_host_free(void *p) { std::free(p); }
int schwz_b200_partition_regular2d(int64_t N, int32_t P, uint32_t *part)
{
    ABI_BEGIN
    partition_regular2d(N, P, part);
    ABI_END
}
int schwz_b200_partition_regular2d_rect(int64_t N, int32_t P, int32_t px, int32_t py,
                                        uint32_t *part)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(partition_regular2d_rect(N, P, px, py, part),
                  "regular2d: N must be a square number and px * py == P");
    ABI_END
}
int schwz_b200_partition_metis(int32_t N, const int32_t *rp, const int32_t *ci, int32_t P,
                               const char *objtype, uint32_t *part)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(partition_metis(N, rp, ci, P, objtype, part) == 0, "METIS partitioning failed");
    ABI_END
}

int schwz_b200_setup_create(int32_t matrix_kind, int32_t grid_n, int32_t N, const int32_t *rp,
                            const int32_t *ci, const double *v, int32_t P, int32_t partition_kind,
                            const uint32_t *part, int32_t overlap, schwz_setup **out)
{
    ABI_BEGIN
    std::unique_ptr<RowSource> src;
    if (matrix_kind == 1) src = make_laplacian2d(grid_n);
    else if (matrix_kind == 2) src = make_laplacian3d(grid_n);
    else src = make_stored(N, rp, ci, v, true);
    auto *h = new schwz_setup();
    h->impl.reset(new Setup(std::move(src), P, partition_kind, part, overlap));
    h->impl->build_index_sets();
    *out = h;
    ABI_END
}
int schwz_b200_setup_destroy(schwz_setup *s)
{
    ABI_BEGIN
    delete s;
    ABI_END
}
int schwz_b200_setup_first_row(const schwz_setup *s, int32_t *out)
{
    ABI_BEGIN
    std::copy(s->impl->first_row().begin(), s->impl->first_row().end(), out);
    ABI_END
}
int schwz_b200_setup_permutation(const schwz_setup *s, int32_t *perm, int32_t *iperm)
{
    ABI_BEGIN
    SCHWZ_REQUIRE(s->impl->permuted(), "no permutation for the regular partition");
    std::copy(s->impl->perm().begin(), s->impl->perm().end(), perm);
    std::copy(s->impl->iperm().begin(), s->impl->iperm().end(), iperm);
    ABI_END
}
int schwz_b200_setup_sizes(schwz_setup *s, int32_t rank, int64_t *out)
{
    ABI_BEGIN
    s->impl->build_matrices(rank);
    RankLayout &R = s->impl->rank(rank);
    out[0] = R.local_size;
    out[1] = R.local_size_x;
    out[2] = R.overlap_size;
    out[3] = R.local.nnz();
    out[4] = R.iface.nnz();
    out[5] = R.n_halo;
    out[6] = (int64_t)R.nbr_in.size();
    out[7] = (int64_t)R.nbr_out.size();
    ABI_END
}
int schwz_b200_setup_l2g(schwz_setup *s, int32_t rank, int32_t *out)
{
    ABI_BEGIN
    RankLayout &R = s->impl->rank(rank);
    std::copy(R.l2g.begin(), R.l2g.end(), out);
    ABI_END
}
int schwz_b200_setup_local_matrix(schwz_setup *s, int32_t rank, int32_t *rp, int32_t *ci, double *v)
{
    ABI_BEGIN
    s->impl->build_matrices(rank);
    copy_out(s->impl->rank(rank).local, rp, ci, v);
    ABI_END
}
int schwz_b200_setup_interface_matrix(schwz_setup *s, int32_t rank, int32_t *rp, int32_t *ci,
                                      double *v)
{
    ABI_BEGIN
    s->impl->build_matrices(rank);
    copy_out(s->impl->rank(rank).iface, rp, ci, v);
    ABI_END
}
int schwz_b200_setup_neighbors(schwz_setup *s, int32_t rank, int32_t *nin, int32_t *nout)
{
    ABI_BEGIN
    RankLayout &R = s->impl->rank(rank);
    std::copy(R.nbr_in.begin(), R.nbr_in.end(), nin);
    std::copy(R.nbr_out.begin(), R.nbr_out.end(), nout);
    ABI_END
}
int schwz_b200_setup_get_count(schwz_setup *s, int32_t rank, int32_t j, int32_t *count)
{
    ABI_BEGIN
    *count = (int32_t)s->impl->rank(rank).get.at(j).size();
    ABI_END
}
int schwz_b200_setup_get_list(schwz_setup *s, int32_t rank, int32_t j, int32_t *out)
{
    ABI_BEGIN
    auto &l = s->impl->rank(rank).get.at(j);
    std::copy(l.begin(), l.end(), out);
    ABI_END
}
int schwz_b200_setup_put_count(schwz_setup *s, int32_t rank, int32_t j, int32_t *count)
{
    ABI_BEGIN
    *count = (int32_t)s->impl->rank(rank).put.at(j).size();
    ABI_END
}
int schwz_b200_setup_put_list(schwz_setup *s, int32_t rank, int32_t j, int32_t *out)
{
    ABI_BEGIN
    auto &l = s->impl->rank(rank).put.at(j);
    std::copy(l.begin(), l.end(), out);
    ABI_END
}
int schwz_b200_setup_displacements(schwz_setup *s, int32_t rank, int32_t *pd, int32_t *gd)
{
    ABI_BEGIN
    RankLayout &R = s->impl->rank(rank);
    std::copy(R.put_disp.begin(), R.put_disp.end(), pd);
    std::copy(R.get_disp.begin(), R.get_disp.end(), gd);
    ABI_END
}
int schwz_b200_setup_release_rank(schwz_setup *s, int32_t rank)
{
    ABI_BEGIN
    s->impl->release(rank);
    ABI_END
}

// ---- RAS subdomain -----------------------------------------------------------
int schwz_b200_ras_create(schwz_ctx *ctx, schwz_setup *s, int32_t rank, const double *rhs,
                          const schwz_ras_options *o, schwz_ras **out)
{
    ABI_BEGIN
    RasOptions ro;
    ro.tolerance = o->tolerance;
    ro.local_tol = o->local_tol;
    ro.local_max_iters = o->local_max_iters;
    ro.local_solver = o->local_solver;
    ro.non_symmetric = o->non_symmetric;
    ro.restart_iter = o->restart_iter;
    ro.overlap = o->overlap;
    ro.use_mixed_precision = o->use_mixed_precision;
    ro.local_precond = o->local_precond;
    ro.precond_max_block_size = o->precond_max_block_size;
    auto *h = new schwz_ras();
    h->impl.reset(new Ras(ctx->impl, *s->impl, rank, rhs, ro));
    *out = h;
    ABI_END
}
int schwz_b200_ras_destroy(schwz_ras *r)
{
    ABI_BEGIN
    delete r;
    ABI_END
}
int schwz_b200_ras_set_factors(schwz_ras *r, const int32_t *Lrp, const int32_t *Lci,
                               const double *Lv, const int32_t *perm)
{
    ABI_BEGIN
    r->impl->set_factors(Lrp, Lci, Lv, perm);
    ABI_END
}
int schwz_b200_ras_set_local_max_iters(schwz_ras *r, int32_t local_max_iters)
{
    ABI_BEGIN
    r->impl->opt.local_max_iters = local_max_iters;
    ABI_END
}
int schwz_b200_ras_set_lu_factors(schwz_ras *r, const schwz_lu *lu, const int32_t *col_perm)
{
    ABI_BEGIN
    r->impl->set_lu_factors(lu->L.rp.data(), lu->L.ci.data(), lu->L.v.data(), lu->U.rp.data(),
                            lu->U.ci.data(), lu->U.v.data(), lu->p.data(), col_perm);
    ABI_END
}
int schwz_b200_ras_mailbox(schwz_ras *r, void **base, schwz_mailbox_layout *l)
{
    ABI_BEGIN
    *base = r->impl->mailbox;
    l->recv_stride = r->impl->mbox.recv_stride;
    l->flags_off = r->impl->mbox.flags_off;
    l->conv_off = r->impl->mbox.conv_off;
    l->err_off = r->impl->mbox.err_off;
    l->send_off = r->impl->mbox.send_off;
    l->slots_off = r->impl->mbox.slots_off;
    l->x_off = r->impl->mbox.x_off;
    l->bytes = r->impl->mbox.bytes;
    ABI_END
}
int schwz_b200_mailbox_layout(int64_t in_total, int32_t n_in, int32_t P, int64_t out_total,
                              int64_t x_len, schwz_mailbox_layout *l)
{
    ABI_BEGIN
    MailboxLayout m = MailboxLayout::make(in_total, n_in, P, out_total, x_len);
    l->recv_stride = m.recv_stride;
    l->flags_off = m.flags_off;
    l->conv_off = m.conv_off;
    l->err_off = m.err_off;
    l->send_off = m.send_off;
    l->slots_off = m.slots_off;
    l->x_off = m.x_off;
    l->bytes = m.bytes;
    ABI_END
}
int schwz_b200_ras_info(schwz_ras *r, int64_t *out)
{
    ABI_BEGIN
    out[0] = (int64_t)r->impl->nbr_in.size();
    out[1] = (int64_t)r->impl->nbr_out.size();
    out[2] = r->impl->local_size;
    out[3] = r->impl->local_size_x;
    out[4] = r->impl->n_halo;
    out[5] = r->impl->local_nnz;
    ABI_END
}
int schwz_b200_ras_neighbors(schwz_ras *r, int32_t *nin, int32_t *nout)
{
    ABI_BEGIN
    std::copy(r->impl->nbr_in.begin(), r->impl->nbr_in.end(), nin);
    std::copy(r->impl->nbr_out.begin(), r->impl->nbr_out.end(), nout);
    ABI_END
}
static MailboxLayout layout_from_abi(const schwz_mailbox_layout *l)
{
    MailboxLayout m;
    m.recv_stride = l->recv_stride;
    m.flags_off = l->flags_off;
    m.conv_off = l->conv_off;
    m.err_off = l->err_off;
    m.send_off = l->send_off;
    m.slots_off = l->slots_off;
    m.x_off = l->x_off;
    m.bytes = l->bytes;
    return m;
}
int schwz_b200_ras_connect(schwz_ras *r, int32_t j, void *peer_base, const schwz_mailbox_layout *l,
                           int32_t recv_off, int32_t flag_slot, int32_t same_process)
{
    ABI_BEGIN
    r->impl->connect(j, peer_base, layout_from_abi(l), recv_off, flag_slot, same_process != 0);
    ABI_END
}
int schwz_b200_ras_connect_in(schwz_ras *r, int32_t j, void *peer_base, const schwz_mailbox_layout *l,
                              int32_t send_off)
{
    ABI_BEGIN
    r->impl->connect_in(j, peer_base, layout_from_abi(l), send_off);
    ABI_END
}
int schwz_b200_ras_connect_conv(schwz_ras *r, int32_t peer_rank, void *peer_base,
                                const schwz_mailbox_layout *l)
{
    ABI_BEGIN
    r->impl->connect_conv(peer_rank, peer_base, layout_from_abi(l));
    ABI_END
}
int schwz_b200_ras_set_exchange_mode(schwz_ras *r, int32_t mode)
{
    ABI_BEGIN
    r->impl->set_exchange_mode(mode);
    ABI_END
}
int schwz_b200_ras_set_onesided(schwz_ras *r, int32_t onesided)
{
    ABI_BEGIN
    r->impl->set_onesided(onesided != 0);
    ABI_END
}
int schwz_b200_ras_connect_local(schwz_ras **subs, int32_t n, schwz_setup *s)
{
    ABI_BEGIN
    for (int32_t a = 0; a < n; ++a) {
        Ras &R = *subs[a]->impl;
        RankLayout &L = s->impl->rank(R.rank);
        for (size_t j = 0; j < R.nbr_out.size(); ++j) {
            const int32_t q = R.nbr_out[j];
            for (int32_t b = 0; b < n; ++b) {
                Ras &Q = *subs[b]->impl;
                if (Q.rank != q) continue;
                int32_t slot = -1;
                for (size_t k = 0; k < Q.nbr_in.size(); ++k)
                    if (Q.nbr_in[k] == R.rank) slot = (int32_t)k;
                SCHWZ_REQUIRE(slot >= 0, "asymmetric neighbour lists");
                R.connect((int32_t)j, Q.mailbox, Q.mbox, L.put_disp[q], slot, true);
            }
        }
        for (size_t j = 0; j < R.nbr_in.size(); ++j) {
            const int32_t q = R.nbr_in[j];
            for (int32_t b = 0; b < n; ++b) {
                Ras &Q = *subs[b]->impl;
                if (Q.rank == q) R.connect_in((int32_t)j, Q.mailbox, Q.mbox, L.get_disp[q]);
            }
        }
        for (int32_t b = 0; b < n; ++b) {
            Ras &Q = *subs[b]->impl;
            R.connect_conv(Q.rank, Q.mailbox, Q.mbox);
        }
    }
    ABI_END
}
int schwz_b200_ras_exchange_push(schwz_ras *r, int32_t iter)
{
    ABI_BEGIN
    r->impl->exchange_push(iter);
    ABI_END
}
int schwz_b200_ras_exchange_unpack(schwz_ras *r, int32_t iter, int32_t wait_flags)
{
    ABI_BEGIN
    r->impl->exchange_unpack(iter, wait_flags != 0);
    ABI_END
}
int schwz_b200_ras_update_boundary(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->update_boundary();
    ABI_END
}
int schwz_b200_ras_local_residual(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->local_residual();
    ABI_END
}
int schwz_b200_ras_residual_norm(schwz_ras *r, double *out)
{
    ABI_BEGIN
    // the stage-by-stage caller's one synchronisation per outer iteration: the halo wait's error
    // word comes back with the norm (a lost peer must not go unnoticed)
    r->impl->ctx.use();
    int32_t err = 0;
    SCHWZ_CUDA(cudaMemcpyAsync(out, r->impl->resnorm_dev, 8, cudaMemcpyDeviceToHost,
                               r->impl->ctx.stream));
    SCHWZ_CUDA(cudaMemcpyAsync(&err, r->impl->err_word(), 4, cudaMemcpyDeviceToHost,
                               r->impl->ctx.stream));
    r->impl->ctx.sync();
    SCHWZ_REQUIRE(err == 0, "halo exchange timed out: a neighbour's boundary values did not "
                            "arrive within SCHWZ_B200_HALO_TIMEOUT_MS");
    ABI_END
}
int schwz_b200_ras_residual_norm_dev(schwz_ras *r, double **dev)
{
    ABI_BEGIN
    *dev = r->impl->resnorm_dev;
    ABI_END
}
int schwz_b200_ras_local_solve(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->local_solve();
    ABI_END
}
int schwz_b200_ras_restrict(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->restrict_to_x();
    ABI_END
}
int schwz_b200_ras_last_local_iters(schwz_ras *r, int32_t *iters)
{
    ABI_BEGIN
    *iters = 0;
    if (r->impl->cg) r->impl->cg->result(iters, nullptr, nullptr);
    else if (r->impl->gmres) r->impl->gmres->result(iters, nullptr, nullptr);
    ABI_END
}
int schwz_b200_ras_wait_push_of(schwz_ras *r, schwz_ras *nbr)
{
    ABI_BEGIN
    r->impl->wait_push_of(*nbr->impl);
    ABI_END
}
int schwz_b200_ras_sync(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->ctx.sync();
    ABI_END
}
int schwz_b200_ras_get_x(schwz_ras *r, double *out)
{
    ABI_BEGIN
    Ras &R = *r->impl;
    R.ctx.use();
    SCHWZ_CUDA(cudaMemcpyAsync(out, R.x, sizeof(double) * ((size_t)R.local_size_x + R.n_halo),
                               cudaMemcpyDeviceToHost, R.ctx.stream));
    R.ctx.sync();
    ABI_END
}
int schwz_b200_ras_get_local_solution(schwz_ras *r, double *out)
{
    ABI_BEGIN
    Ras &R = *r->impl;
    R.ctx.use();
    // after an iterative local solve local_solution == init_guess (source/solve.cpp:781); the
    // copy is not materialised on the device
    SCHWZ_CUDA(cudaMemcpyAsync(out, R.solution_vector(), sizeof(double) * (size_t)R.local_size_x,
                               cudaMemcpyDeviceToHost, R.ctx.stream));
    R.ctx.sync();
    ABI_END
}
int schwz_b200_ras_set_x_own(schwz_ras *r, const double *in)
{
    ABI_BEGIN
    Ras &R = *r->impl;
    R.ctx.use();
    SCHWZ_CUDA(cudaMemcpyAsync(R.x, in, sizeof(double) * (size_t)R.local_size,
                               cudaMemcpyHostToDevice, R.ctx.stream));
    R.ctx.sync();
    ABI_END
}
int schwz_b200_ras_upload_rhs(schwz_ras *r, const double *host_rhs_global)
{
    ABI_BEGIN
    r->impl->upload_rhs(host_rhs_global);
    ABI_END
}
int schwz_b200_ras_download_solution(schwz_ras *r, double *host_solution_global)
{
    ABI_BEGIN
    r->impl->download_solution(host_solution_global);
    ABI_END
}
int schwz_b200_ras_reset(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->reset_state();
    ABI_END
}
int schwz_b200_ras_kernel_time(schwz_ras *r, int32_t kind, int32_t reps, float *ms)
{
    ABI_BEGIN
    *ms = r->impl->kernel_time_ms(kind, reps);
    ABI_END
}
int64_t schwz_b200_ras_kernel_bytes(schwz_ras *r, int32_t kind)
{
    Ras &R = *r->impl;
    const int64_t n = R.local_size_x, nnz = R.local_nnz;
    switch (kind) {
    case 0: return 12 * nnz + 4 * (n + 1) + 8 * n + 8 * n;            // val+col, rowptr, q, p
    case 1: return 24 * n;                                            // r,q read; r write
    case 2: return 40 * n;                                            // r,p,x read; p,x write
    case 3: return 12 * nnz + 4 * (n + 1) + 8 * n + 8 * n + 8 * n;    // + rhs read
    case 4: {                                                         // push + unpack: 20 B/element
        int64_t e = 0;
        for (int32_t c : R.in_count) e += c;
        for (int32_t c : R.out_count) e += c;
        return 20 * e;
    }
    default: return 0;
    }
}
int schwz_b200_ras_true_residual_sq(schwz_ras *r, double *out)
{
    ABI_BEGIN
    *out = r->impl->true_residual_sq();
    ABI_END
}
int schwz_b200_ras_conv_set_local(schwz_ras *r, int32_t converged_all_local)
{
    ABI_BEGIN
    r->impl->conv_forward(converged_all_local);
    ABI_END
}
int schwz_b200_ras_conv_tree(schwz_ras *r, int32_t converged_all_local)
{
    ABI_BEGIN
    r->impl->conv_tree(converged_all_local);
    ABI_END
}
int schwz_b200_ras_conv_accumulate(schwz_ras *r, int32_t converged_all_local)
{
    ABI_BEGIN
    r->impl->conv_accumulate(converged_all_local);
    ABI_END
}
int schwz_b200_ras_conv_forward(schwz_ras *r)
{
    ABI_BEGIN
    r->impl->conv_forward(0);
    ABI_END
}
int schwz_b200_ras_conv_count(schwz_ras *r, int32_t *n)
{
    ABI_BEGIN
    Ras &R = *r->impl;
    R.ctx.use();
    SCHWZ_CUDA(cudaMemcpyAsync(n, R.num_converged_dev, 4, cudaMemcpyDeviceToHost, R.ctx.stream));
    R.ctx.sync();
    ABI_END
}

int schwz_b200_ras_run(schwz_ras **subs, int32_t n_local, const schwz_loop_options *o,
                       schwz_loop_result *res, double *history)
{
    ABI_BEGIN
    std::vector<Ras *> v;
    for (int32_t i = 0; i < n_local; ++i) v.push_back(subs[i]->impl.get());
    LoopOptions lo;
    lo.num_subdomains = o->num_subdomains;
    lo.max_iters = o->max_iters;
    lo.tolerance = o->tolerance;
    lo.enable_onesided = o->enable_onesided;
    lo.enable_global_check = o->enable_global_check;
    lo.conv_decentralized = o->conv_decentralized;
    lo.iter_offset = o->iter_offset;
    lo.exchange_mode = o->exchange_mode;
    lo.conv_tree = (o->enable_onesided && !o->conv_decentralized) ? 1 : 0;
    lo.conv_accumulate = (o->enable_onesided && o->conv_decentralized == 2) ? 1 : 0;
    lo.comm = o->comm ? o->comm->impl.get() : nullptr;
    LoopResult lr;
    ras_run(v, lo, lr, history);
    res->iters = lr.iters;
    res->converged = lr.converged;
    res->global_resnorm = lr.global_resnorm;
    res->global_resnorm0 = lr.global_resnorm0;
    res->elapsed_s = lr.elapsed_s;
    res->host_stream_syncs = lr.host_stream_syncs;
    res->host_event_waits = lr.host_event_waits;
    ABI_END
}

int schwz_b200_ras_refresh_halo(schwz_ras **subs, int32_t n_local, int32_t num_subdomains)
{
    ABI_BEGIN
    std::vector<Ras *> v;
    for (int32_t i = 0; i < n_local; ++i) v.push_back(subs[i]->impl.get());
    ras_refresh_halo(v, num_subdomains);
    ABI_END
}

// ---- NCCL ----------------------------------------------------------------------
int schwz_b200_comm_unique_id(void *id128)
{
    ABI_BEGIN
    comm_unique_id(id128);
    ABI_END
}
int schwz_b200_comm_create(schwz_ctx *ctx, const void *id128, int32_t nranks, int32_t rank,
                           schwz_comm **out)
{
    ABI_BEGIN
    auto *h = new schwz_comm();
    h->impl.reset(comm_create(ctx->impl, id128, nranks, rank));
    *out = h;
    ABI_END
}
int schwz_b200_comm_destroy(schwz_comm *c)
{
    ABI_BEGIN
    delete c;
    ABI_END
}
int schwz_b200_comm_allgather_f64(schwz_comm *c, const double *in, int32_t count, double *out)
{
    ABI_BEGIN
    comm_allgather_f64(*c->impl, in, count, out);
    ABI_END
}

}  // extern "C"
