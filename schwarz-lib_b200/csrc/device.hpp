// Internal C++ layer under the C ABI: device context, CSR container, kernel
// launchers.  Nothing here is exported; include/schwz_b200.h is the boundary.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "common.cuh"

namespace schwz_b200 {

struct Ctx {
    int device = 0;
    int num_sms = 148;                 // queried at creation (B200: 148)
    int vec_grid() const { return num_sms * kVecCtasPerSM; }
    cudaStream_t stream = nullptr;
    double *partials = nullptr;        // kMaxPartials doubles
    unsigned int *tickets = nullptr;   // 16 tickets, zero-initialised
    double *dev_scalars = nullptr;     // 16 doubles of scratch
    double *pinned = nullptr;          // 16 doubles, host-pinned
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;

    explicit Ctx(int dev);
    ~Ctx();
    void use() const { SCHWZ_CUDA(cudaSetDevice(device)); }
    void sync() const
    {
        use();
        SCHWZ_CUDA(cudaStreamSynchronize(stream));
    }
    template <typename T>
    T *alloc(size_t n) const
    {
        use();
        void *p = nullptr;
        SCHWZ_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
        return (T *)p;
    }
    template <typename T>
    T *alloc_zero(size_t n) const
    {
        T *p = alloc<T>(n);
        SCHWZ_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), stream));
        return p;
    }
    template <typename T>
    T *upload(const T *host, size_t n) const
    {
        T *p = alloc<T>(n);
        if (n) SCHWZ_CUDA(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        SCHWZ_CUDA(cudaStreamSynchronize(stream));
        return p;
    }
    void release(void *p) const
    {
        if (p) {
            use();
            cudaFree(p);
        }
    }
};

// Device CSR + the row tiling used by the streaming SpMV: CTA b owns rows
// [blk_row[b], blk_row[b+1]) with at most kBlock rows and (unless it is a
// single long row) at most kSpmvTile non-zeros.
struct DeviceCsr {
    const Ctx *ctx = nullptr;
    int32_t nrows = 0, ncols = 0;
    int64_t nnz = 0;
    int32_t *rp = nullptr, *ci = nullptr;
    double *v = nullptr;
    int32_t nblocks = 0;
    int32_t *blk_row = nullptr;
    bool has_long_row = false;
    int32_t rows_per_tile = kBlock;
    // Compact column stream of the pipelined SpMV: a row tile whose columns span less than
    // 2^16 (stencil / banded local matrices: a 256-row tile of a cfg2 strip spans 16 640
    // columns) is read as 16-bit offsets from the tile's first column instead of 32-bit
    // indices: 10 instead of 12 B per non-zero from HBM.  Other tiles (tile_col0 < 0: the
    // overlap rows of a strip couple to both ends of the own block) keep their 32-bit indices.
    uint16_t *ci16 = nullptr;
    int32_t *tile_col0 = nullptr;
    ~DeviceCsr();
};

// Programmatic dependent launch inside a CG solve: while a PdlScope(true) is alive on this host
// thread, the CG step kernels (x/p update, SpMV, r update, init, flush) are launched with
// programmaticStreamSerializationAllowed: their CTAs become resident while the predecessor's
// last CTAs still run and wait in cudaGridDependencySynchronize(), which every one of those
// kernels executes before it touches anything.  MEASURED AND NOT KEPT AS THE DEFAULT
// (profiles/r2_while_graph.md): unlike the triangular solves, whose kernels have a dependent
// chain of loads to hide, these are full-width streaming kernels and the early-resident
// successor only takes SM slots from the running one - 1.19 against 1.15 ms per 50-iteration
// solve at 524 k rows, 10.47 against 9.41 ms at 8.4 M rows.  SCHWZ_B200_CG_PDL=1 turns it on.
extern bool g_cg_pdl;
extern thread_local bool t_pdl_launch;
struct PdlScope {
    bool prev;
    explicit PdlScope(bool on) : prev(t_pdl_launch) { t_pdl_launch = on; }
    ~PdlScope() { t_pdl_launch = prev; }
};
template <typename K, typename... Args>
inline void launch_maybe_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (t_pdl_launch && g_cg_pdl) ? 1 : 0;
    SCHWZ_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
}

extern bool g_spmv_col16;   // SCHWZ_B200_SPMV_COL32=1 turns the 16-bit column stream off
extern bool g_force_simple_spmv;
extern int g_spmv_variant;
DeviceCsr *csr_upload(const Ctx &ctx, int32_t nrows, int32_t ncols, const int32_t *rp,
                      const int32_t *ci, const double *v);

// scalars of one device-resident Krylov solve
struct CgScalars {
    double rho, prev_rho, beta, r0, resnorm, tol;
    int32_t iter, stop, max_iters, pending;   // pending: x += alpha * p not applied yet
    double alpha;
};

enum SpmvEpilogue { EPI_NONE = 0, EPI_DOT = 1, EPI_NRM2SQ = 2, EPI_NRM2 = 3 };

// y_out = alpha*A*x + beta*y_in  (+ fused reduction over rows < red_rows).
// stop: optional device flag; the launch is a no-op when *stop != 0.
void launch_spmv(const Ctx &ctx, const DeviceCsr &A, double alpha, const double *x,
                 double beta, const double *y_in, double *y_out, SpmvEpilogue epi,
                 const double *dot_with, double *result, int32_t red_rows,
                 const int32_t *stop);

void launch_dot(const Ctx &ctx, int64_t n, const double *a, const double *b, double *result,
                bool sqrt_result);
void launch_axpy(const Ctx &ctx, int64_t n, double alpha, const double *x, double *y);
void launch_copy(const Ctx &ctx, int64_t n, const double *src, double *dst);
// dst = src as a kernel that honours a device stop flag (no-op when *stop != 0)
void launch_copy_guarded(const Ctx &ctx, int64_t n, const double *src, double *dst,
                         const int32_t *stop);
void launch_gather(const Ctx &ctx, int32_t n, const int32_t *idx, const double *from,
                   double *into, int op);
void launch_scatter(const Ctx &ctx, int32_t n, const int32_t *idx, const double *from,
                    double *into, int op);
void launch_permute(const Ctx &ctx, int32_t n, const int32_t *perm, int inverse,
                    const double *in, double *out, const int32_t *stop = nullptr);

// CG step kernels (Ginkgo Cg semantics)
void launch_cg_init(const Ctx &ctx, CgScalars *s, int32_t max_iters, double tol,
                    const int32_t *outer_stop, cudaGraphConditionalHandle loop = 0);
void launch_cg_xp_update(const Ctx &ctx, int64_t n, const double *r_or_z, double *p, double *x,
                         const CgScalars *s);
void launch_cg_r_update(const Ctx &ctx, int64_t n, double *r, const double *q, CgScalars *s,
                        bool precond = false, cudaGraphConditionalHandle loop = 0);
void launch_cg_flush_x(const Ctx &ctx, int64_t n, double *x, const double *p, const CgScalars *s);

// halo
void launch_halo_pack_push(const Ctx &ctx, int32_t nseg, const int32_t *seg_off_dev,
                           int32_t total, const int32_t *src_idx, const double *x,
                           void *const *dst_ptrs, unsigned long long *const *flag_ptrs,
                           unsigned long long epoch, const int32_t *stop, bool f32 = false);
// flags != nullptr: wait (bounded by timeout_ns) until every in-neighbour has published `epoch`;
// on expiry *error_flag = 1 (the host raises at its next poll).  stop: no-op when set.
void launch_halo_unpack(const Ctx &ctx, int32_t nseg, int32_t total, const int32_t *dst_idx,
                        const void *recv, double *x, const unsigned long long *flags,
                        unsigned long long epoch, int32_t *error_flag, bool f32 = false,
                        const int32_t *stop = nullptr);
extern long long g_halo_timeout_ns;   // SCHWZ_B200_HALO_TIMEOUT_MS, default 20 s

void launch_halo_put_elements(const Ctx &ctx, int32_t nseg, const int32_t *seg_off_dev,
                              int32_t total, const int32_t *src_idx, const int32_t *remote_slot,
                              const double *x, double *const *peer_x,
                              const int32_t *stop = nullptr);
void launch_halo_pull(const Ctx &ctx, int32_t nseg, const int32_t *seg_off_dev, int32_t total,
                      const int32_t *dst_idx, const int32_t *src_idx,
                      const void *const *src_ptrs, double *x, bool f32 = false,
                      const int32_t *stop = nullptr);

// Device-resident bookkeeping of one subdomain's outer loop (what Solve::check_convergence keeps
// in host variables, source/solve.cpp:796-1005).  `stop` is the flag every launch of the
// subdomain honours: once set (converged, or an error) the rest of the enqueued work is no-ops,
// so the host never has to read anything back inside the loop.
struct OuterState {
    double resnorm, resnorm0, gres, gres0;
    int32_t num_converged, stop, finished_iter, iter;
    int32_t error;   // 0 ok, 1 halo wait timed out, 2 residual norm is NaN, 3 diverged
    int32_t pad[3];
};
enum OuterError { OUTER_OK = 0, OUTER_HALO_TIMEOUT = 1, OUTER_NAN = 2, OUTER_DIVERGED = 3 };

// out[i] = *norm_ptrs[i]: the local residual norms of a process, contiguous for the allgather
void launch_gather_norms(const Ctx &ctx, int32_t n, const double *const *norm_ptrs, double *out);
// Synchronous global decision (source/solve.cpp:888-912 + the break test of
// schwarz_base.cpp:432-433) for the nl subdomains of this process: all[P] = every subdomain's
// local residual norm in rank order, slot[i] = position of local subdomain i in it.
void launch_ras_decide(const Ctx &ctx, int32_t P, int32_t nl, const double *all,
                       const int32_t *slot, OuterState *const *states, double tol,
                       int32_t check, int32_t enable_global_check, int32_t iter, double *history);
// One-sided: local ratio test + flag protocol + break test on the subdomain's own stream
// (source/solve.cpp:913-943).  protocol: 0 flag flooding, 1 tree, 2 accumulate.
void launch_ras_conv_decide(const Ctx &ctx, int32_t protocol, int32_t P, int32_t me,
                            OuterState *state, const double *resnorm_dev, double tol,
                            int32_t check, int32_t iter, double *history, int32_t *conv,
                            int32_t *conv_sent, int32_t n_out, int32_t *const *out_conv,
                            int32_t *const *peer_conv, int32_t *num_converged);

// convergence flags (include/conv_tools.hpp:248-274 on peer-mapped words)
void launch_conv_forward(const Ctx &ctx, int32_t P, int32_t me, int32_t converged_all_local,
                         int32_t *conv, int32_t *conv_sent, int32_t n_out,
                         int32_t *const *peer_conv, int32_t *num_converged);
// centralised tree (include/conv_tools.hpp:147-209); peer_conv indexed by subdomain id
void launch_conv_tree(const Ctx &ctx, int32_t P, int32_t me, int32_t converged_all_local,
                      int32_t *conv, int32_t *const *peer_conv, int32_t *num_converged);

// accumulate variant of the decentralised check (include/conv_tools.hpp:230-247)
void launch_conv_accumulate(const Ctx &ctx, int32_t P, int32_t me, int32_t converged_all_local,
                            int32_t *conv, int32_t *const *peer_conv, int32_t *num_converged);

}  // namespace schwz_b200
