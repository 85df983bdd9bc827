// sm_100a kernels of the RAS hot path: streaming CSR SpMV with fused
// reductions, fused CG vector steps with device-resident scalars, halo
// pack/push + unpack with peer stores and system-scope epoch flags, indexed
// gather/scatter, permutation, convergence-flag forwarding.
//
// Everything here is HBM-bound fp64/int32 work (arithmetic intensity <= 0.17
// flop/B), so the design rules are: coalesced streaming of rowptr/col/val,
// >= 8 independent loads in flight per thread, x reused through L1/L2, fixed
// summation order (bit-reproducible), no host round trips.
#include <algorithm>
#include <cfloat>

#include "device.hpp"

namespace schwz_b200 {

std::atomic<int64_t> g_launches{0};
bool g_cg_pdl = false;              // SCHWZ_B200_CG_PDL=1: programmatic dependent launch inside a CG solve
thread_local bool t_pdl_launch = false;
bool g_spmv_col16 = true;           // SCHWZ_B200_SPMV_COL32=1: 32-bit column indices (A/B)
bool g_force_simple_spmv = false;   // SCHWZ_B200_SIMPLE_SPMV=1: one-shot kernel (A/B measurements)
int g_spmv_variant = 1;             // SCHWZ_B200_SPMV_VARIANT: launch shape of the pipelined kernel

// =============================================================================
// Context
// =============================================================================
Ctx::Ctx(int dev) : device(dev)
{
    use();
    SCHWZ_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    SCHWZ_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    partials = alloc<double>(kMaxPartials);
    tickets = alloc_zero<unsigned int>(16);
    dev_scalars = alloc_zero<double>(16);
    SCHWZ_CUDA(cudaMallocHost((void **)&pinned, 16 * sizeof(double)));
    SCHWZ_CUDA(cudaEventCreate(&ev_start));
    SCHWZ_CUDA(cudaEventCreate(&ev_stop));
    SCHWZ_CUDA(cudaStreamSynchronize(stream));
}

Ctx::~Ctx()
{
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(partials);
    cudaFree(tickets);
    cudaFree(dev_scalars);
    if (pinned) cudaFreeHost(pinned);
    if (ev_start) cudaEventDestroy(ev_start);
    if (ev_stop) cudaEventDestroy(ev_stop);
    if (stream) cudaStreamDestroy(stream);
}

DeviceCsr::~DeviceCsr()
{
    if (!ctx) return;
    ctx->release(rp);
    ctx->release(ci);
    ctx->release(v);
    ctx->release(blk_row);
    ctx->release(ci16);
    ctx->release(tile_col0);
}

DeviceCsr *csr_upload(const Ctx &ctx, int32_t nrows, int32_t ncols, const int32_t *rp,
                      const int32_t *ci, const double *v)
{
    auto *A = new DeviceCsr();
    A->ctx = &ctx;
    A->nrows = nrows;
    A->ncols = ncols;
    A->nnz = nrows > 0 ? rp[nrows] : 0;
    // Row tiling: greedy, <= kBlock rows and <= kSpmvTile nnz per CTA; a row
    // longer than the tile gets a CTA of its own (long-row path).
    const int rpt = (g_spmv_variant >= 3 && g_spmv_variant < 10) ? 2 : 1;
    A->rows_per_tile = kBlock * rpt;
    const int32_t tile_rows = A->rows_per_tile, tile_nnz = kSpmvTile * rpt;
    std::vector<int32_t> blk;
    blk.reserve(nrows / tile_rows + 2);
    blk.push_back(0);
    int32_t r = 0;
    while (r < nrows) {
        int32_t r1 = r + 1;   // always take at least one row
        const int32_t k0 = rp[r];
        while (r1 < nrows && r1 - r < tile_rows && rp[r1 + 1] - k0 <= tile_nnz) ++r1;
        blk.push_back(r1);
        r = r1;
    }
    A->nblocks = (int32_t)blk.size() - 1;
    A->has_long_row = false;
    for (int32_t b = 0; b < A->nblocks; ++b)
        if (rp[blk[b + 1]] - rp[blk[b]] > tile_nnz) A->has_long_row = true;
    if (A->has_long_row && rpt != 1) {
        // the one-shot fallback kernel works on kBlock-row tiles: retile
        A->rows_per_tile = kBlock;
        blk.assign(1, 0);
        r = 0;
        while (r < nrows) {
            int32_t r1 = r + 1;
            const int32_t k0 = rp[r];
            while (r1 < nrows && r1 - r < kBlock && rp[r1 + 1] - k0 <= kSpmvTile) ++r1;
            blk.push_back(r1);
            r = r1;
        }
        A->nblocks = (int32_t)blk.size() - 1;
    }
    // The pipelined kernel copies 16-byte aligned supersets of each tile, so
    // every array is padded by 8 elements past its end.
    auto upload_padded = [&](auto *host, size_t n, auto zero) {
        using T = decltype(zero);
        T *d = ctx.alloc_zero<T>(n + 16);
        if (n) SCHWZ_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));
        return d;
    };
    std::vector<int32_t> rp_fallback(1, 0);
    A->rp = upload_padded(nrows > 0 ? rp : rp_fallback.data(), (size_t)nrows + 1, int32_t(0));
    A->ci = upload_padded(ci, (size_t)A->nnz, int32_t(0));
    A->v = upload_padded(v, (size_t)A->nnz, double(0));
    A->blk_row = ctx.upload(blk.data(), blk.size());
    if (g_spmv_col16 && !A->has_long_row && A->nnz > 0) {
        // per tile: first column and whether the tile's columns fit 16 bits from there.  Tiles
        // that do not (the overlap rows of a strip couple to both ends of the own block) keep
        // their 32-bit indices: tile_col0 = -1.
        std::vector<int32_t> col0((size_t)A->nblocks, 0);
        std::vector<uint16_t> c16((size_t)A->nnz, 0);
        int64_t narrow_nnz = 0;
        for (int32_t b = 0; b < A->nblocks; ++b) {
            const int32_t k0 = rp[blk[b]], k1 = rp[blk[b + 1]];
            if (k0 == k1) continue;
            int32_t lo = ci[k0], hi = ci[k0];
            for (int32_t k = k0 + 1; k < k1; ++k) {
                lo = std::min(lo, ci[k]);
                hi = std::max(hi, ci[k]);
            }
            if ((int64_t)hi - lo < 65536) {
                col0[b] = lo;
                for (int32_t k = k0; k < k1; ++k) c16[k] = (uint16_t)(ci[k] - lo);
                narrow_nnz += k1 - k0;
            } else {
                col0[b] = -1;
            }
        }
        if (narrow_nnz * 10 >= A->nnz * 9) {   // worth it when (almost) every tile is narrow
            // 16-byte granules of 8 entries: pad by 16
            uint16_t *d = ctx.alloc_zero<uint16_t>((size_t)A->nnz + 16);
            SCHWZ_CUDA(cudaMemcpyAsync(d, c16.data(), c16.size() * sizeof(uint16_t),
                                       cudaMemcpyHostToDevice, ctx.stream));
            SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));
            A->ci16 = d;
            A->tile_col0 = ctx.upload(col0.data(), col0.size());
        }
    }
    return A;
}

// =============================================================================
// Streaming CSR SpMV  (replaces gko::matrix::Csr::apply; the hot kernel of the
// local solve).  One CTA = one row tile:
//   phase 1  every thread loads kSpmvUnroll (val, col) pairs of the tile with
//            fully coalesced 8 B / 4 B loads (all issued before any use), then
//            gathers x[col] through the read-only path and leaves
//            (alpha*val)*x in shared memory;
//   phase 2  thread t sums the products of row t sequentially in stored
//            column order — the same order as the CPU restatement, so results
//            are bit-identical and independent of the launch shape.
// Optional fused epilogue: sum_r y[r]*w[r] or sum_r y[r]^2 via per-CTA
// partials and a last-CTA ordered final sum.
// Algorithmic HBM bytes per launch: 12*nnz + 4*(rows+1) + 8*rows (y) +
// 8*cols (x once) [+ 8*rows if beta != 0] [+ 8*rows for the fused dot operand
// when it is not x].
// =============================================================================
template <int EPI>
__global__ void __launch_bounds__(kBlock)
    csr_spmv_stream_kernel(const int32_t *__restrict__ blk_row, const int32_t *__restrict__ rp,
                           const int32_t *__restrict__ ci, const double *__restrict__ v,
                           const double *__restrict__ x, double alpha, double beta,
                           const double *y_in, double *y_out, const double *dot_with,
                           double *partials, unsigned int *ticket, double *result,
                           int32_t red_rows, const int32_t *stop)
{
    __shared__ double s_prod[kSpmvTile];
    __shared__ int32_t s_rp[kBlock + 1];
    __shared__ double s_warp[kBlock / 32];

    if (stop != nullptr && *stop != 0) return;

    const int t = threadIdx.x;
    const int32_t r0 = blk_row[blockIdx.x];
    const int32_t nr = blk_row[blockIdx.x + 1] - r0;
    for (int i = t; i <= nr; i += kBlock) s_rp[i] = rp[r0 + i];   // nr + 1 <= kBlock + 1 entries
    __syncthreads();
    const int32_t k0 = s_rp[0];
    const int32_t nnz = s_rp[nr] - k0;

    double acc = 0.0;
    if (nnz <= kSpmvTile) {
        double vv[kSpmvUnroll];
        int32_t cc[kSpmvUnroll];
#pragma unroll
        for (int i = 0; i < kSpmvUnroll; ++i) {
            const int k = t + i * kBlock;
            if (k < nnz) {
                vv[i] = __ldcs(v + k0 + k);    // streamed once: evict-first
                cc[i] = __ldcs(ci + k0 + k);
            }
        }
#pragma unroll
        for (int i = 0; i < kSpmvUnroll; ++i) {
            const int k = t + i * kBlock;
            if (k < nnz) s_prod[k] = (alpha * vv[i]) * __ldg(x + cc[i]);
        }
        __syncthreads();
        if (t < nr) {
            const int32_t row = r0 + t;
            acc = (beta == 0.0) ? 0.0 : beta * y_in[row];
            const int32_t a = s_rp[t] - k0, b = s_rp[t + 1] - k0;
            for (int32_t k = a; k < b; ++k) acc += s_prod[k];
            y_out[row] = acc;
        }
    } else {
        // single long row: CTA-strided products, fixed-shape tree sum
        double part = 0.0;
        for (int32_t k = t; k < nnz; k += kBlock)
            part += (alpha * v[k0 + k]) * __ldg(x + ci[k0 + k]);
        part = block_sum(part, s_warp);
        if (t == 0) {
            acc = ((beta == 0.0) ? 0.0 : beta * y_in[r0]) + part;
            y_out[r0] = acc;
        }
    }

    if (EPI != EPI_NONE) {
        double c = 0.0;
        const int32_t row = r0 + t;
        if (t < nr && row < red_rows) c = (EPI == EPI_DOT) ? acc * dot_with[row] : acc * acc;
        c = block_sum(c, s_warp);
        if (t == 0) partials[blockIdx.x] = c;
        if (last_cta(ticket)) {
            double s = reduce_partials(partials, gridDim.x, s_warp);
            if (t == 0) *result = (EPI == EPI_NRM2) ? sqrt(s) : s;
        }
    }
}

// =============================================================================
// Persistent, warp-specialised, TMA-pipelined variant of the streaming SpMV —
// the production path.  Grid = kSpmvCtasPerSM CTAs per SM; each CTA walks the
// row tiles b, b + grid, ...  One producer warp keeps kSpmvStages tiles in
// flight: for every tile a single lane issues three cp.async.bulk copies
// (val, col, rowptr; 16-byte aligned supersets of the tile's ranges, L2
// evict-first) that complete on the stage's "full" mbarrier.  Eight consumer
// warps wait on that barrier, gather x through the read-only path, leave
// (alpha*val)*x in shared memory, synchronise among themselves on a named
// barrier, sum each row sequentially in stored order (bit-identical to the CPU
// restatement), store y and release the stage on its "empty" mbarrier.  The
// HBM stream of tile i+2 therefore overlaps the gather and the row sums of
// tile i, which is what the one-shot kernel above cannot do.
// =============================================================================
// Launch shapes of the pipelined kernel: RPT rows per consumer thread (a tile is
// kBlock*RPT rows), STAGES tiles in flight per CTA, CTAS resident CTAs per SM.
// DeviceCsr::rows_per_tile fixes RPT at upload time.
constexpr int kSpmvChunk = 6;     // non-zeros of a row gathered per round
constexpr int kSpmvThreads = kBlock + 32;   // 8 consumer warps + 1 producer warp

template <int RPT, typename ColT>
struct __align__(16) SpmvStage {
    double val[kSpmvTile * RPT + 16];
    int32_t col[kSpmvTile * RPT + 16];   // 32-bit indices, or 16-bit offsets in its first half
    int32_t rp[kBlock * RPT + 8];
};
template <int RPT, int STAGES, typename ColT>
struct __align__(16) SpmvSmem {
    SpmvStage<RPT, ColT> st[STAGES];
    double warp_buf[kBlock / 32];
    unsigned long long full[STAGES], empty[STAGES];
    int last;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, void *bar,
                                         unsigned long long policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(kBlock) : "memory");
}

// ColT = int32_t: column indices as stored; uint16_t: per tile either 16-bit offsets from
// tile_col0[tile] (ci16) or, where tile_col0[tile] < 0, the 32-bit indices (ci)
template <int EPI, int RPT, int STAGES, int CTAS, typename ColT>
__global__ void __launch_bounds__(kSpmvThreads, CTAS)
    csr_spmv_tma_kernel(int32_t ntiles, const int32_t *__restrict__ blk_row,
                        const int32_t *__restrict__ rp, const int32_t *__restrict__ ci,
                        const uint16_t *__restrict__ ci16, const int32_t *__restrict__ tile_col0,
                        const double *__restrict__ v, const double *__restrict__ x, double alpha,
                        double beta, const double *y_in, double *y_out, const double *dot_with,
                        double *partials, unsigned int *ticket, double *result, int32_t red_rows,
                        const int32_t *stop)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = SpmvSmem<RPT, STAGES, ColT>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    // (a no-op unless launched with programmatic stream serialisation: see PdlScope)
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (stop != nullptr && *stop != 0) return;
    constexpr bool kCol16 = sizeof(ColT) == 2;

    const int t = threadIdx.x;
    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], kBlock / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (t >= kBlock) {
        // ------------------------------ producer warp ------------------------
        if (t == kBlock) {
            unsigned long long policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int it = 0;
            for (int32_t b = blockIdx.x; b < ntiles; b += gridDim.x, ++it) {
                const int s = it % STAGES;
                const int32_t r0 = blk_row[b], r1 = blk_row[b + 1];
                const int32_t k0 = rp[r0], k1 = rp[r1];
                // 16-byte granules of the column stream: 8 entries of 16 bits, 4 of 32
                const bool narrow = kCol16 && tile_col0[b] >= 0;
                const int32_t G = narrow ? 8 : 4;
                const uint32_t cb = narrow ? 2u : 4u;
                const int32_t k0a = k0 & ~(G - 1), k1a = (k1 + G - 1) & ~(G - 1);
                const int32_t r0a = r0 & ~3, r1a = (r1 + 1 + 3) & ~3;
                const uint32_t nv = (uint32_t)(k1a - k0a), nr = (uint32_t)(r1a - r0a);
                mbar_wait(&S.empty[s], ((it / STAGES) & 1) ^ 1);
                mbar_expect_tx(&S.full[s], nv * (8u + cb) + nr * 4u);
                if (nv) {
                    bulk_g2s(S.st[s].val, v + k0a, nv * 8u, &S.full[s], policy);
                    if (narrow) bulk_g2s(S.st[s].col, ci16 + k0a, nv * 2u, &S.full[s], policy);
                    else bulk_g2s(S.st[s].col, ci + k0a, nv * 4u, &S.full[s], policy);
                }
                bulk_g2s(S.st[s].rp, rp + r0a, nr * 4u, &S.full[s], policy);
            }
        }
        return;
    }

    // -------------------------------- consumers ------------------------------
    // Thread t owns rows r0 + t + j*kBlock (j < RPT) of the tile and walks them
    // in stored order straight out of the staged tile: per non-zero one LDS.64
    // (val), one LDS.32 (col) and one x gather.  Neighbouring threads own
    // neighbouring rows, so for stencil-like matrices the i-th gathers of a
    // warp fall on consecutive addresses (coalesced) and the staged reads are
    // bank-conflict free (stride = row length).  All RPT*kSpmvChunk gathers of
    // a round are issued before the first use.  No barrier between warps: each
    // warp releases the stage as soon as its rows are done.
    double red = 0.0;
    int it = 0;
    for (int32_t b = blockIdx.x; b < ntiles; b += gridDim.x, ++it) {
        const int s = it % STAGES;
        const int32_t r0 = blk_row[b];
        const int32_t nr = blk_row[b + 1] - r0;
        mbar_wait(&S.full[s], (it / STAGES) & 1);
        const SpmvStage<RPT, ColT> &T = S.st[s];
        const int32_t *srp = T.rp + (r0 & 3);
        int32_t c0 = -1;
        if (kCol16) c0 = __ldg(tile_col0 + b);
        const bool narrow = kCol16 && c0 >= 0;
        const int32_t base = srp[0] & ~(narrow ? 7 : 3);   // first staged element
        const double *xt = narrow ? x + c0 : x;            // offsets count from the tile's column 0
        const uint16_t *col16 = reinterpret_cast<const uint16_t *>(T.col);
        int32_t k[RPT], e[RPT];
        double acc[RPT];
        bool more = false;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int32_t lr = t + j * kBlock;
            if (lr < nr) {
                k[j] = srp[lr] - base;
                e[j] = srp[lr + 1] - base;
                acc[j] = (beta == 0.0) ? 0.0 : beta * y_in[r0 + lr];
            } else {
                k[j] = e[j] = 0;
                acc[j] = 0.0;
            }
            more |= k[j] < e[j];
        }
        while (more) {
            double xv[RPT][kSpmvChunk], vv[RPT][kSpmvChunk];
#pragma unroll
            for (int j = 0; j < RPT; ++j)
#pragma unroll
                for (int i = 0; i < kSpmvChunk; ++i)
                    if (k[j] + i < e[j]) {
                        xv[j][i] = __ldg(xt + (narrow ? (int32_t)col16[k[j] + i] : T.col[k[j] + i]));
                        vv[j][i] = T.val[k[j] + i];
                    }
            more = false;
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
#pragma unroll
                for (int i = 0; i < kSpmvChunk; ++i)
                    if (k[j] + i < e[j]) acc[j] += (alpha * vv[j][i]) * xv[j][i];
                k[j] += kSpmvChunk;
                more |= k[j] < e[j];
            }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int32_t lr = t + j * kBlock;
            if (lr < nr) {
                const int32_t row = r0 + lr;
                y_out[row] = acc[j];
                if (EPI != EPI_NONE && row < red_rows)
                    red += (EPI == EPI_DOT) ? acc[j] * dot_with[row] : acc[j] * acc[j];
            }
        }
        __syncwarp();
        if ((t & 31) == 0) mbar_arrive(&S.empty[s]);
    }

    if (EPI != EPI_NONE) {
        // CTA-wide sum over the 8 consumer warps, then last-CTA ordered final sum
        red = warp_sum(red);
        if ((t & 31) == 0) S.warp_buf[t >> 5] = red;
        consumer_sync();
        if (t < 32) {
            double c = t < (kBlock / 32) ? S.warp_buf[t] : 0.0;
            c = warp_sum(c);
            if (t == 0) {
                partials[blockIdx.x] = c;
                __threadfence();
                const unsigned int tk = atomicAdd(ticket, 1u);
                S.last = (tk == gridDim.x - 1);
                if (S.last) *ticket = 0u;
            }
        }
        consumer_sync();
        if (S.last) {
            __threadfence();
            double sacc = 0.0;
            for (int i = t; i < (int)gridDim.x; i += kBlock) sacc += __ldcg(partials + i);
            sacc = warp_sum(sacc);
            if ((t & 31) == 0) S.warp_buf[t >> 5] = sacc;
            consumer_sync();
            if (t < 32) {
                double c = t < (kBlock / 32) ? S.warp_buf[t] : 0.0;
                c = warp_sum(c);
                if (t == 0) *result = (EPI == EPI_NRM2) ? sqrt(c) : c;
            }
        }
    }
}

template <int EPI, int RPT, int STAGES, int CTAS, typename ColT>
static void launch_spmv_tma_cols(const Ctx &ctx, const DeviceCsr &A, double alpha,
                                 const double *x, double beta, const double *y_in, double *y_out,
                                 const double *dot_with, double *result, int32_t red_rows,
                                 const int32_t *stop)
{
    using Smem = SpmvSmem<RPT, STAGES, ColT>;
    // per-device, set once; several host threads (bench_ras rank threads) may arrive together:
    // setting the attribute twice is harmless, the flag is an atomic
    static std::atomic<bool> configured[64];
    if (!configured[ctx.device].load(std::memory_order_acquire)) {
        SCHWZ_CUDA(cudaFuncSetAttribute(csr_spmv_tma_kernel<EPI, RPT, STAGES, CTAS, ColT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(Smem)));
        configured[ctx.device].store(true, std::memory_order_release);
    }
    const int grid = std::min<int>(A.nblocks, ctx.num_sms * CTAS);
    launch_maybe_pdl(csr_spmv_tma_kernel<EPI, RPT, STAGES, CTAS, ColT>, dim3(grid),
                     dim3(kSpmvThreads), sizeof(Smem), ctx.stream, A.nblocks,
                     (const int32_t *)A.blk_row, (const int32_t *)A.rp, (const int32_t *)A.ci,
                     (const uint16_t *)A.ci16, (const int32_t *)A.tile_col0, (const double *)A.v, x,
                     alpha, beta, y_in, y_out, dot_with, ctx.partials, ctx.tickets + 0, result,
                     red_rows, stop);
}

template <int EPI, int RPT, int STAGES, int CTAS>
static void launch_spmv_tma_cfg(const Ctx &ctx, const DeviceCsr &A, double alpha, const double *x,
                                double beta, const double *y_in, double *y_out,
                                const double *dot_with, double *result, int32_t red_rows,
                                const int32_t *stop)
{
    if (A.ci16 != nullptr)
        launch_spmv_tma_cols<EPI, RPT, STAGES, CTAS, uint16_t>(ctx, A, alpha, x, beta, y_in, y_out,
                                                               dot_with, result, red_rows, stop);
    else
        launch_spmv_tma_cols<EPI, RPT, STAGES, CTAS, int32_t>(ctx, A, alpha, x, beta, y_in, y_out,
                                                              dot_with, result, red_rows, stop);
}

template <int EPI>
static void launch_spmv_tma(const Ctx &ctx, const DeviceCsr &A, double alpha, const double *x,
                            double beta, const double *y_in, double *y_out, const double *dot_with,
                            double *result, int32_t red_rows, const int32_t *stop)
{
#define SCHWZ_TMA(RPT, ST, CT)                                                                  \
    launch_spmv_tma_cfg<EPI, RPT, ST, CT>(ctx, A, alpha, x, beta, y_in, y_out, dot_with, result, \
                                          red_rows, stop)
    // g_spmv_variant picks the launch shape for a given rows_per_tile
    if (A.rows_per_tile == kBlock) {
        switch (g_spmv_variant) {
        case 0: SCHWZ_TMA(1, 3, 2); break;
        case 2: SCHWZ_TMA(1, 4, 2); break;
        case 10: SCHWZ_TMA(1, 3, 3); break;
        case 11: SCHWZ_TMA(1, 2, 5); break;
        case 12: SCHWZ_TMA(1, 2, 3); break;
        default: SCHWZ_TMA(1, 2, 4); break;   // measured best: 0.946 of HBM peak (profiles/)
        }
    } else {
        switch (g_spmv_variant) {
        case 4: SCHWZ_TMA(2, 3, 1); break;
        default: SCHWZ_TMA(2, 2, 2); break;
        }
    }
#undef SCHWZ_TMA
}

void launch_spmv(const Ctx &ctx, const DeviceCsr &A, double alpha, const double *x,
                 double beta, const double *y_in, double *y_out, SpmvEpilogue epi,
                 const double *dot_with, double *result, int32_t red_rows,
                 const int32_t *stop)
{
    if (A.nrows == 0) {
        if (epi != EPI_NONE)
            SCHWZ_CUDA(cudaMemsetAsync(result, 0, sizeof(double), ctx.stream));
        return;
    }
    ctx.use();
    if (!A.has_long_row && !g_force_simple_spmv) {
        switch (epi) {
        case EPI_NONE: launch_spmv_tma<EPI_NONE>(ctx, A, alpha, x, beta, y_in, y_out, dot_with, result, red_rows, stop); break;
        case EPI_DOT: launch_spmv_tma<EPI_DOT>(ctx, A, alpha, x, beta, y_in, y_out, dot_with, result, red_rows, stop); break;
        case EPI_NRM2SQ: launch_spmv_tma<EPI_NRM2SQ>(ctx, A, alpha, x, beta, y_in, y_out, dot_with, result, red_rows, stop); break;
        case EPI_NRM2: launch_spmv_tma<EPI_NRM2>(ctx, A, alpha, x, beta, y_in, y_out, dot_with, result, red_rows, stop); break;
        }
        SCHWZ_CUDA(cudaGetLastError());
        count_launch();
        return;
    }
    // one partial per tile here (the persistent kernel above needs one per resident CTA only)
    SCHWZ_REQUIRE(A.nblocks <= kMaxPartials || epi == EPI_NONE,
                  "matrix too large for the fused reduction scratch of the one-shot SpMV");
    dim3 grid(A.nblocks), block(kBlock);
#define SCHWZ_SPMV_CASE(E)                                                                   \
    csr_spmv_stream_kernel<E><<<grid, block, 0, ctx.stream>>>(                               \
        A.blk_row, A.rp, A.ci, A.v, x, alpha, beta, y_in, y_out, dot_with, ctx.partials,     \
        ctx.tickets + 0, result, red_rows, stop)
    switch (epi) {
    case EPI_NONE: SCHWZ_SPMV_CASE(EPI_NONE); break;
    case EPI_DOT: SCHWZ_SPMV_CASE(EPI_DOT); break;
    case EPI_NRM2SQ: SCHWZ_SPMV_CASE(EPI_NRM2SQ); break;
    case EPI_NRM2: SCHWZ_SPMV_CASE(EPI_NRM2); break;
    }
#undef SCHWZ_SPMV_CASE
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// BLAS-1
// =============================================================================
__global__ void __launch_bounds__(kBlock)
    dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
               double *partials, unsigned int *ticket, double *result, int do_sqrt)
{
    __shared__ double s_warp[kBlock / 32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock)
        s += a[i] * b[i];
    s = block_sum(s, s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    if (last_cta(ticket)) {
        double r = reduce_partials(partials, gridDim.x, s_warp);
        if (threadIdx.x == 0) *result = do_sqrt ? sqrt(r) : r;
    }
}

static int vec_grid(const Ctx &ctx, int64_t n)
{
    int64_t need = (n + kBlock - 1) / kBlock;
    return (int)std::max<int64_t>(1, std::min<int64_t>(need, ctx.vec_grid()));
}

void launch_dot(const Ctx &ctx, int64_t n, const double *a, const double *b, double *result,
                bool sqrt_result)
{
    ctx.use();
    dot_kernel<<<vec_grid(ctx, n), kBlock, 0, ctx.stream>>>(n, a, b, ctx.partials, ctx.tickets + 1,
                                                       result, sqrt_result ? 1 : 0);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void __launch_bounds__(kBlock)
    axpy_kernel(int64_t n, double alpha, const double *__restrict__ x, double *__restrict__ y)
{
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock)
        y[i] += alpha * x[i];
}

void launch_axpy(const Ctx &ctx, int64_t n, double alpha, const double *x, double *y)
{
    if (n <= 0) return;
    ctx.use();
    axpy_kernel<<<vec_grid(ctx, n), kBlock, 0, ctx.stream>>>(n, alpha, x, y);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

void launch_copy(const Ctx &ctx, int64_t n, const double *src, double *dst)
{
    if (n <= 0) return;
    ctx.use();
    SCHWZ_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
}

__global__ void __launch_bounds__(kBlock)
    copy_guarded_kernel(int64_t n, const double *__restrict__ src, double *__restrict__ dst,
                        const int32_t *stop)
{
    if (stop != nullptr && *stop != 0) return;
    const bool vec = ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0;
    if (vec) {
        const int64_t n2 = n >> 1;
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        const int64_t stride = (int64_t)gridDim.x * kBlock * 4;
        for (int64_t i0 = (int64_t)blockIdx.x * kBlock * 4 + threadIdx.x; i0 < n2; i0 += stride) {
            double2 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + (int64_t)j * kBlock < n2) v[j] = __ldcs(s2 + i0 + (int64_t)j * kBlock);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + (int64_t)j * kBlock < n2) d2[i0 + (int64_t)j * kBlock] = v[j];
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = src[n - 1];
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
             i += (int64_t)gridDim.x * kBlock)
            dst[i] = src[i];
    }
}

void launch_copy_guarded(const Ctx &ctx, int64_t n, const double *src, double *dst,
                         const int32_t *stop)
{
    if (n <= 0) return;
    ctx.use();
    copy_guarded_kernel<<<vec_grid(ctx, (n + 7) / 8), kBlock, 0, ctx.stream>>>(n, src, dst, stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// Gather / scatter with the reference's four operations
// (include/gather.hpp:86-107, include/scatter.hpp:86-108; add 0, copy 1,
// diff 2, avg 3 — include/collective_common.hpp:37).
// =============================================================================
__device__ __forceinline__ double combine(int op, double from, double into)
{
    switch (op) {
    case 0: return from + into;
    case 2: return from - into;
    case 3: return (from + into) / 2;
    default: return from;
    }
}

__global__ void __launch_bounds__(kBlock)
    gather_kernel(int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ from,
                  double *__restrict__ into, int op)
{
    for (int32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        const double f = from[idx[i]];
        into[i] = (op == 1) ? f : combine(op, f, into[i]);
    }
}

__global__ void __launch_bounds__(kBlock)
    scatter_kernel(int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ from,
                   double *__restrict__ into, int op)
{
    for (int32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        const int32_t j = idx[i];
        const double f = from[i];
        into[j] = (op == 1) ? f : combine(op, f, into[j]);
    }
}

void launch_gather(const Ctx &ctx, int32_t n, const int32_t *idx, const double *from,
                   double *into, int op)
{
    if (n <= 0) return;
    ctx.use();
    gather_kernel<<<vec_grid(ctx, n), kBlock, 0, ctx.stream>>>(n, idx, from, into, op);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

void launch_scatter(const Ctx &ctx, int32_t n, const int32_t *idx, const double *from,
                    double *into, int op)
{
    if (n <= 0) return;
    ctx.use();
    scatter_kernel<<<vec_grid(ctx, n), kBlock, 0, ctx.stream>>>(n, idx, from, into, op);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// gko::matrix::Permutation::apply: row_permute out[i] = in[perm[i]];
// row_permute|inverse_permute out[perm[i]] = in[i]  (source/solve.cpp:717-720)
__global__ void __launch_bounds__(kBlock)
    permute_kernel(int32_t n, const int32_t *__restrict__ perm, int inverse,
                   const double *__restrict__ in, double *__restrict__ out, const int32_t *stop)
{
    if (stop != nullptr && *stop != 0) return;
    for (int32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        if (inverse) out[perm[i]] = in[i];
        else out[i] = in[perm[i]];
    }
}

void launch_permute(const Ctx &ctx, int32_t n, const int32_t *perm, int inverse,
                    const double *in, double *out, const int32_t *stop)
{
    if (n <= 0) return;
    ctx.use();
    permute_kernel<<<vec_grid(ctx, n), kBlock, 0, ctx.stream>>>(n, perm, inverse, in, out, stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// CG steps.  Ginkgo Cg ordering:
//   rho = r.r ; ++iter ; stop? ; p = r + (rho/prev_rho) p ; q = A p ;
//   beta = p.q ; x += (rho/beta) p ; r -= (rho/beta) q ; swap(prev_rho, rho)
// Three launches per iteration: cg_xp_update (x, p), the SpMV with p.q fused, cg_r_update
// (r, ||r||^2, stop test).
// All scalars and the stop decision live in CgScalars on the device; every
// kernel returns at once when stop is set, so the host may enqueue more
// iterations than are needed and never has to read a scalar back inside the
// solve.
// =============================================================================
// `loop`: handle of the WHILE node around the iteration when the solve runs as a conditional
// CUDA graph (0 otherwise): the kernels that take the stop decision also tell the graph
// whether to go round again.
__global__ void cg_init_kernel(CgScalars *s, int32_t max_iters, double tol,
                               const int32_t *outer_stop, cudaGraphConditionalHandle loop)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    // s->rho already holds ||b - A x0||^2 (fused into the residual SpMV)
    const double r0 = sqrt(s->rho);
    s->r0 = r0;
    s->resnorm = r0;
    s->prev_rho = 1.0;
    s->beta = 0.0;
    s->tol = tol;
    s->iter = 0;
    s->alpha = 0.0;
    s->pending = 0;
    s->max_iters = max_iters;
    // Iteration(max) || ResidualNormReduction(tol): ||r|| < tol * ||r0||
    int stop = (0 >= max_iters) || (r0 < tol * r0);
    if (outer_stop != nullptr && *outer_stop != 0) stop = 1;
    s->stop = stop;
    if (loop) cudaGraphSetConditional(loop, stop ? 0u : 1u);
}

void launch_cg_init(const Ctx &ctx, CgScalars *s, int32_t max_iters, double tol,
                    const int32_t *outer_stop, cudaGraphConditionalHandle loop)
{
    ctx.use();
    launch_maybe_pdl(cg_init_kernel, dim3(1), dim3(1), 0, ctx.stream, s, max_iters, tol, outer_stop,
                     loop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// The x update of step_2 is DEFERRED by one kernel: the r update below only records
// alpha = rho/beta, and the next p update - which reads p anyway - applies x += alpha * p_old
// before it overwrites p.  Same operands, same operation, same result bit for bit; it saves
// the second pass over p (8 B/row/iteration: 152 -> 144).  cg_flush_x_kernel applies the last
// pending update when the solve ends (stop set, or the budget used up).
//
// step_1 (+ pending step_2a): x += alpha p ; p = z + (rho/prev_rho) p
//   (prev_rho == 0 or first iteration: p = z; z = r without a preconditioner)
// Vector kernels keep kVecUnroll independent 16-byte loads per operand in
// flight per thread (the loads of a trip are all issued before the first use).
constexpr int kVecUnroll = 4;

__global__ void __launch_bounds__(kBlock)
    cg_xp_update_kernel(int64_t n, const double *__restrict__ r, double *__restrict__ p,
                        double *__restrict__ x, const CgScalars *__restrict__ s)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (s->stop) return;
    const bool fresh = (s->iter == 0) || (s->prev_rho == 0.0);
    const double t = fresh ? 0.0 : s->rho / s->prev_rho;
    const bool pend = s->pending != 0;
    const double a = s->alpha;
    const int64_t n2 = n >> 1;
    const double2 *r2 = reinterpret_cast<const double2 *>(r);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    double2 *x2 = reinterpret_cast<double2 *>(x);
    const int64_t stride = (int64_t)gridDim.x * kBlock * kVecUnroll;
    for (int64_t i0 = (int64_t)blockIdx.x * kBlock * kVecUnroll + threadIdx.x; i0 < n2; i0 += stride) {
        double2 rv[kVecUnroll], pv[kVecUnroll], xv[kVecUnroll];
#pragma unroll
        for (int j = 0; j < kVecUnroll; ++j) {
            const int64_t i = i0 + (int64_t)j * kBlock;
            if (i < n2) {
                rv[j] = __ldcs(r2 + i);
                if (!fresh || pend) pv[j] = p2[i];
                if (pend) xv[j] = x2[i];
            }
        }
#pragma unroll
        for (int j = 0; j < kVecUnroll; ++j) {
            const int64_t i = i0 + (int64_t)j * kBlock;
            if (i < n2) {
                if (pend) {
                    xv[j].x += a * pv[j].x;
                    xv[j].y += a * pv[j].y;
                    x2[i] = xv[j];
                }
                if (fresh) {
                    p2[i] = rv[j];
                } else {
                    pv[j].x = rv[j].x + t * pv[j].x;
                    pv[j].y = rv[j].y + t * pv[j].y;
                    p2[i] = pv[j];
                }
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double po = p[n - 1];
        if (pend) x[n - 1] += a * po;
        p[n - 1] = fresh ? r[n - 1] : r[n - 1] + t * po;
    }
}

void launch_cg_xp_update(const Ctx &ctx, int64_t n, const double *r, double *p, double *x,
                         const CgScalars *s)
{
    ctx.use();
    launch_maybe_pdl(cg_xp_update_kernel, dim3(vec_grid(ctx, (n + 1) / 2)), dim3(kBlock), 0,
                     ctx.stream, n, r, p, x, s);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// step_2b + next rho + stop test: r -= alpha q ; ||r||^2 ; alpha left pending for x
__global__ void __launch_bounds__(kBlock)
    cg_r_update_kernel(int64_t n, double *__restrict__ r, const double *__restrict__ q,
                       CgScalars *s, double *partials, unsigned int *ticket, int precond,
                       cudaGraphConditionalHandle loop)
{
    __shared__ double s_warp[kBlock / 32];
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (s->stop) return;
    const double rho = s->rho, beta = s->beta;
    const bool skip = (beta == 0.0);
    const double a = skip ? 0.0 : rho / beta;
    double acc = 0.0;
    const int64_t n2 = n >> 1;
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *q2 = reinterpret_cast<const double2 *>(q);
    const int64_t stride = (int64_t)gridDim.x * kBlock * kVecUnroll;
    for (int64_t i0 = (int64_t)blockIdx.x * kBlock * kVecUnroll + threadIdx.x; i0 < n2; i0 += stride) {
        double2 rv[kVecUnroll], qv[kVecUnroll];
#pragma unroll
        for (int j = 0; j < kVecUnroll; ++j) {
            const int64_t i = i0 + (int64_t)j * kBlock;
            if (i < n2) {
                rv[j] = r2[i];
                if (!skip) qv[j] = __ldcs(q2 + i);
            }
        }
#pragma unroll
        for (int j = 0; j < kVecUnroll; ++j) {
            const int64_t i = i0 + (int64_t)j * kBlock;
            if (i < n2) {
                if (!skip) {
                    rv[j].x -= a * qv[j].x;
                    rv[j].y -= a * qv[j].y;
                    r2[i] = rv[j];
                }
                acc += rv[j].x * rv[j].x + rv[j].y * rv[j].y;
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double rv = r[n - 1];
        if (!skip) {
            rv -= a * q[n - 1];
            r[n - 1] = rv;
        }
        acc += rv * rv;
    }
    acc = block_sum(acc, s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    if (last_cta(ticket)) {
        double rho_new = reduce_partials(partials, gridDim.x, s_warp);
        if (threadIdx.x == 0) {
            s->alpha = a;
            s->pending = skip ? 0 : 1;
            s->prev_rho = rho;
            // with a preconditioner rho = r.z comes from the next M^-1 application and this
            // sum is only ||r||^2 for the stop test
            if (!precond) s->rho = rho_new;
            const int it = s->iter + 1;
            s->iter = it;
            const double tau = sqrt(rho_new);
            s->resnorm = tau;
            const bool stop = it >= s->max_iters || tau < s->tol * s->r0;
            if (stop) s->stop = 1;
            if (loop) cudaGraphSetConditional(loop, stop ? 0u : 1u);
        }
    }
}

void launch_cg_r_update(const Ctx &ctx, int64_t n, double *r, const double *q, CgScalars *s,
                        bool precond, cudaGraphConditionalHandle loop)
{
    ctx.use();
    launch_maybe_pdl(cg_r_update_kernel, dim3(vec_grid(ctx, (n + 1) / 2)), dim3(kBlock), 0,
                     ctx.stream, n, r, q, s, ctx.partials, ctx.tickets + 2, precond ? 1 : 0, loop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// the x update still pending when the solve ends: x += alpha p
__global__ void __launch_bounds__(kBlock)
    cg_flush_x_kernel(int64_t n, double *__restrict__ x, const double *__restrict__ p,
                      const CgScalars *__restrict__ s)
{
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (!s->pending) return;
    const double a = s->alpha;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock)
        x[i] += a * p[i];
}

void launch_cg_flush_x(const Ctx &ctx, int64_t n, double *x, const double *p, const CgScalars *s)
{
    ctx.use();
    launch_maybe_pdl(cg_flush_x_kernel, dim3(vec_grid(ctx, n)), dim3(kBlock), 0, ctx.stream, n, x,
                     (const double *)p, s);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// Halo exchange.  One launch packs the values for ALL out-neighbours and
// stores them straight into the neighbours' receive buffers (peer-mapped over
// NVLink, or local when the neighbour shares the GPU) — this replaces
// Gather + MPI_Isend / pack_buffer + MPI_Put (restricted_schwarz.cpp:881-912,
// comm_helpers.hpp:92-150).  The last CTA then publishes the epoch to each
// neighbour's flag word with a system-scope release store (≙ MPI_Win_flush).
// =============================================================================
constexpr int kMaxSeg = 64;

// T = payload type on the wire: double, or float for use_mixed_precision (the float
// mirrors mixedt_send_buffer / mixedt_recv_buffer of restricted_schwarz.cpp:483-603, 769-787,
// 898-903: values are rounded to MixedValueType by the sender, widened by the receiver).
template <typename T>
__global__ void __launch_bounds__(kBlock)
    halo_pack_push_kernel(int32_t nseg, const int32_t *__restrict__ seg_off, int32_t total,
                          const int32_t *__restrict__ src_idx, const double *__restrict__ x,
                          void *const *__restrict__ dst_ptrs,
                          const int32_t *stop)
{
    __shared__ int32_t s_off[kMaxSeg + 1];
    __shared__ T *s_dst[kMaxSeg];
    if (stop != nullptr && *stop != 0) return;
    if (threadIdx.x <= nseg) s_off[threadIdx.x] = seg_off[threadIdx.x];
    if (threadIdx.x < nseg) s_dst[threadIdx.x] = (T *)dst_ptrs[threadIdx.x];
    __syncthreads();
    for (int32_t e = blockIdx.x * kBlock + threadIdx.x; e < total; e += gridDim.x * kBlock) {
        int s = 0;
        while (e >= s_off[s + 1]) ++s;
        s_dst[s][e - s_off[s]] = (T)x[src_idx[e]];
    }
}

// Publishes the epoch to every out-neighbour's flag word (the replacement for MPI_Win_flush +
// a flag Put).  Its own launch, right behind the pack/push kernel on the same stream: the
// kernel boundary orders all of that grid's peer stores before this one, and the release store
// is cumulative, so no thread of the data kernel has to execute a system-scope fence.  (ncu,
// cfg2: pack/push 5.7 us, publish 6.5 us with fence + release store; with a fence in every
// thread of the data kernel the single-kernel version took 14.0 us.  A system-scope release
// costs microseconds on this platform whoever issues it.)
__global__ void halo_publish_kernel(int32_t nseg, unsigned long long *const *__restrict__ flag_ptrs,
                                    unsigned long long epoch, const int32_t *stop)
{
    if (stop != nullptr && *stop != 0) return;
    if ((int)threadIdx.x < nseg) st_release_sys(flag_ptrs[threadIdx.x], epoch);
}

void launch_halo_pack_push(const Ctx &ctx, int32_t nseg, const int32_t *seg_off_dev,
                           int32_t total, const int32_t *src_idx, const double *x,
                           void *const *dst_ptrs, unsigned long long *const *flag_ptrs,
                           unsigned long long epoch, const int32_t *stop, bool f32)
{
    if (nseg <= 0) return;
    SCHWZ_REQUIRE(nseg <= kMaxSeg, "too many out-neighbours for one push launch");
    ctx.use();
    int grid = std::max(1, std::min((total + kBlock - 1) / kBlock, ctx.vec_grid()));
    if (f32)
        halo_pack_push_kernel<float><<<grid, kBlock, 0, ctx.stream>>>(nseg, seg_off_dev, total,
                                                                      src_idx, x, dst_ptrs, stop);
    else
        halo_pack_push_kernel<double><<<grid, kBlock, 0, ctx.stream>>>(nseg, seg_off_dev, total,
                                                                       src_idx, x, dst_ptrs, stop);
    count_launch();
    if (flag_ptrs != nullptr) {
        halo_publish_kernel<<<1, kMaxSeg, 0, ctx.stream>>>(nseg, flag_ptrs, epoch, stop);
        count_launch();
    }
    SCHWZ_CUDA(cudaGetLastError());
}

// Unpack for ALL in-neighbours in one launch (Scatter copy,
// restricted_schwarz.cpp:955-959 / comm_helpers.hpp:153-177).  With `flags`
// the kernel first waits until every in-neighbour has published `epoch`
// (synchronous semantics across processes); without, it scatters whatever the
// buffer holds (asynchronous semantics).  The wait is bounded so a lost peer
// cannot hang the GPU; on expiry error_flag is raised.
long long g_halo_timeout_ns = 20ll * 1000 * 1000 * 1000;

template <typename T>
__global__ void __launch_bounds__(kBlock)
    halo_unpack_kernel(int32_t nseg, int32_t total, const int32_t *__restrict__ dst_idx,
                       const void *recv, double *__restrict__ x,
                       const unsigned long long *flags, unsigned long long epoch,
                       int32_t *error_flag, long long timeout_ns, const int32_t *stop)
{
    if (stop != nullptr && *stop != 0) return;
    if (flags != nullptr) {
        if (threadIdx.x < nseg) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int spins = 0;
            while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
                __nanosleep(64);
                if ((++spins & 1023u) == 0 &&
                    (long long)(global_timer_ns() - t0) > timeout_ns) {
                    // a lost peer must not hang the GPU: raise the error word (the host reads
                    // it at its next poll and throws) and go on with what the buffer holds
                    if (error_flag) atomicExch(error_flag, 1);
                    break;
                }
            }
        }
        __syncthreads();
    }
    const volatile T *rv = (const volatile T *)recv;   // written by peers: never cache in L1
    for (int32_t e = blockIdx.x * kBlock + threadIdx.x; e < total; e += gridDim.x * kBlock)
        x[dst_idx[e]] = (double)rv[e];
}

void launch_halo_unpack(const Ctx &ctx, int32_t nseg, int32_t total, const int32_t *dst_idx,
                        const void *recv, double *x, const unsigned long long *flags,
                        unsigned long long epoch, int32_t *error_flag, bool f32,
                        const int32_t *stop)
{
    if (nseg <= 0 || total <= 0) return;
    SCHWZ_REQUIRE(nseg <= kBlock, "too many in-neighbours for one unpack launch");
    ctx.use();
    int grid = std::max(1, std::min((total + kBlock - 1) / kBlock, ctx.vec_grid()));
    if (f32)
        halo_unpack_kernel<float><<<grid, kBlock, 0, ctx.stream>>>(
            nseg, total, dst_idx, recv, x, flags, epoch, error_flag, g_halo_timeout_ns, stop);
    else
        halo_unpack_kernel<double><<<grid, kBlock, 0, ctx.stream>>>(
            nseg, total, dst_idx, recv, x, flags, epoch, error_flag, g_halo_timeout_ns, stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// -----------------------------------------------------------------------------
// The other three one-sided exchange variants of the reference
// (restricted_schwarz.cpp:753-851, comm_helpers.hpp:58-89):
//   * Put one-by-one: every element is stored straight into the neighbour's x
//     at the slot the neighbour keeps it in (no buffers, no unpack) — MPI_Put
//     per element on window_x becomes one peer store per element;
//   * Get gathered:   the owner packs into its own send buffer, the receiver
//     pulls its block out of it over NVLink and scatters in the same kernel
//     (MPI_Get on window_send_buffer + unpack_buffer);
//   * Get one-by-one: the receiver reads the owner's x directly (MPI_Get per
//     element on window_x).
// All three are asynchronous by construction: whatever the peer holds at that
// moment is what travels.
// -----------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
    halo_put_elements_kernel(int32_t nseg, const int32_t *__restrict__ seg_off, int32_t total,
                             const int32_t *__restrict__ src_idx,
                             const int32_t *__restrict__ remote_slot, const double *__restrict__ x,
                             double *const *__restrict__ peer_x, const int32_t *stop)
{
    __shared__ int32_t s_off[kMaxSeg + 1];
    __shared__ double *s_dst[kMaxSeg];
    if (stop != nullptr && *stop != 0) return;
    if (threadIdx.x <= nseg) s_off[threadIdx.x] = seg_off[threadIdx.x];
    if (threadIdx.x < nseg) s_dst[threadIdx.x] = peer_x[threadIdx.x];
    __syncthreads();
    for (int32_t e = blockIdx.x * kBlock + threadIdx.x; e < total; e += gridDim.x * kBlock) {
        int s = 0;
        while (e >= s_off[s + 1]) ++s;
        s_dst[s][remote_slot[e]] = x[src_idx[e]];
    }
    __threadfence_system();
}

void launch_halo_put_elements(const Ctx &ctx, int32_t nseg, const int32_t *seg_off_dev,
                              int32_t total, const int32_t *src_idx, const int32_t *remote_slot,
                              const double *x, double *const *peer_x, const int32_t *stop)
{
    if (nseg <= 0 || total <= 0) return;
    SCHWZ_REQUIRE(nseg <= kMaxSeg, "too many out-neighbours for one put launch");
    ctx.use();
    int grid = std::max(1, std::min((total + kBlock - 1) / kBlock, ctx.vec_grid()));
    halo_put_elements_kernel<<<grid, kBlock, 0, ctx.stream>>>(nseg, seg_off_dev, total, src_idx,
                                                              remote_slot, x, peer_x, stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// x[dst_idx[e]] = src[s][ src_idx ? src_idx[e] : e - seg_off[s] ]  (peer loads)
template <typename T>
__global__ void __launch_bounds__(kBlock)
    halo_pull_kernel(int32_t nseg, const int32_t *__restrict__ seg_off, int32_t total,
                     const int32_t *__restrict__ dst_idx, const int32_t *__restrict__ src_idx,
                     const void *const *__restrict__ src_ptrs, double *__restrict__ x,
                     const int32_t *stop)
{
    __shared__ int32_t s_off[kMaxSeg + 1];
    __shared__ const T *s_src[kMaxSeg];
    if (stop != nullptr && *stop != 0) return;
    if (threadIdx.x <= nseg) s_off[threadIdx.x] = seg_off[threadIdx.x];
    if (threadIdx.x < nseg) s_src[threadIdx.x] = (const T *)src_ptrs[threadIdx.x];
    __syncthreads();
    for (int32_t e = blockIdx.x * kBlock + threadIdx.x; e < total; e += gridDim.x * kBlock) {
        int s = 0;
        while (e >= s_off[s + 1]) ++s;
        const volatile T *src = s_src[s];   // peers write it: never cache in L1
        x[dst_idx[e]] = (double)src[src_idx != nullptr ? src_idx[e] : e - s_off[s]];
    }
}

void launch_halo_pull(const Ctx &ctx, int32_t nseg, const int32_t *seg_off_dev, int32_t total,
                      const int32_t *dst_idx, const int32_t *src_idx,
                      const void *const *src_ptrs, double *x, bool f32, const int32_t *stop)
{
    if (nseg <= 0 || total <= 0) return;
    SCHWZ_REQUIRE(nseg <= kMaxSeg, "too many in-neighbours for one pull launch");
    ctx.use();
    int grid = std::max(1, std::min((total + kBlock - 1) / kBlock, ctx.vec_grid()));
    if (f32)
        halo_pull_kernel<float><<<grid, kBlock, 0, ctx.stream>>>(nseg, seg_off_dev, total, dst_idx,
                                                                 src_idx, src_ptrs, x, stop);
    else
        halo_pull_kernel<double><<<grid, kBlock, 0, ctx.stream>>>(nseg, seg_off_dev, total, dst_idx,
                                                                  src_idx, src_ptrs, x, stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// Decentralised convergence flags, include/conv_tools.hpp:248-274:
//   if converged_all_local: conv[me] = 1 (sticky)
//   snapshot = conv ; num = sum(conv)
//   every out-neighbour gets a remote store of 1 for each flag newly known
//   conv_sent = snapshot
// conv lives in the peer-visible mailbox; remote MPI_Put ≙ relaxed system-scope
// store, MPI_Win_flush ≙ __threadfence_system().
// =============================================================================
// body shared by the host-driven stage kernel and the device-side decision kernel below;
// all threads of one CTA call it, the count is returned in every thread
__device__ int conv_forward_body(int32_t P, int32_t me, int32_t converged_all_local, int32_t *conv,
                                 int32_t *conv_sent, int32_t n_out, int32_t *const *peer_conv)
{
    __shared__ int s_num;
    if (threadIdx.x == 0) {
        if (converged_all_local == 1) st_relaxed_sys_i32(conv + me, 1);
        s_num = 0;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        const int c = ld_relaxed_sys_i32(conv + j);
        if (c == 1) atomicAdd(&s_num, 1);
        if (conv_sent[j] == 0 && c == 1) {
            for (int i = 0; i < n_out; ++i) st_relaxed_sys_i32(peer_conv[i] + j, 1);
            conv_sent[j] = 1;
        }
    }
    __threadfence_system();
    __syncthreads();
    return s_num;
}

__global__ void conv_forward_kernel(int32_t P, int32_t me, int32_t converged_all_local,
                                    int32_t *conv, int32_t *conv_sent, int32_t n_out,
                                    int32_t *const *peer_conv, int32_t *num_converged)
{
    const int num = conv_forward_body(P, me, converged_all_local, conv, conv_sent, n_out, peer_conv);
    if (threadIdx.x == 0) *num_converged = num;
}

void launch_conv_forward(const Ctx &ctx, int32_t P, int32_t me, int32_t converged_all_local,
                         int32_t *conv, int32_t *conv_sent, int32_t n_out,
                         int32_t *const *peer_conv, int32_t *num_converged)
{
    ctx.use();
    conv_forward_kernel<<<1, 128, 0, ctx.stream>>>(P, me, converged_all_local, conv, conv_sent,
                                                   n_out, peer_conv, num_converged);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// Centralised binary-tree convergence (Yamazaki et al. 2019), restated from
// include/conv_tools.hpp:147-209 on the same peer-visible conv words:
//   conv[0], conv[1] : "child 0 / child 1 has pushed up" (conv[0] == 2 once I pushed)
//   conv[2]          : "the root has seen everybody" (pushed down)
// peer_conv[q] = conv array of subdomain q (null when q is not connected).
// =============================================================================
// thread 0 only; returns num_converged_procs
__device__ int conv_tree_body(int32_t P, int32_t me, int32_t converged_all_local, int32_t *conv,
                              int32_t *const *peer_conv)
{
    const int c0 = ld_relaxed_sys_i32(conv + 0);
    const int c1 = ld_relaxed_sys_i32(conv + 1);
    if (((c0 == 1 && c1 == 1) || (c0 == 1 && me == P / 2 - 1) || (me >= P / 2 && c0 != 2)) &&
        converged_all_local > 0) {
        if (me == 0) {
            st_relaxed_sys_i32(conv + 2, 1);
        } else {
            const int parent = (me - 1) / 2;
            const int id = (me % 2 == 0) ? 1 : 0;
            st_relaxed_sys_i32(peer_conv[parent] + id, 1);
            __threadfence_system();
        }
        st_relaxed_sys_i32(conv + 0, 2);
    }
    if (ld_relaxed_sys_i32(conv + 2) == 1) {
        for (int p = 2 * me + 1; p <= 2 * me + 2; ++p)
            if (p < P) st_relaxed_sys_i32(peer_conv[p] + 2, 1);
        __threadfence_system();
        st_relaxed_sys_i32(conv + 1, ld_relaxed_sys_i32(conv + 1) + 1);
        return P;
    }
    return 0;
}

__global__ void conv_tree_kernel(int32_t P, int32_t me, int32_t converged_all_local, int32_t *conv,
                                 int32_t *const *peer_conv, int32_t *num_converged)
{
    if (threadIdx.x != 0) return;
    *num_converged = conv_tree_body(P, me, converged_all_local, conv, peer_conv);
}

// Decentralised, accumulate variant (include/conv_tools.hpp:230-247,
// --enable_decentralized_accumulate): while a subdomain is locally converged it adds 1 to word 0
// of EVERY other subdomain's flags (MPI_Accumulate SUM -> a system-scope atomic add over NVLink)
// and to its own; num_converged_procs = its word 0.  The counter keeps growing as long as
// subdomains stay converged - the reference's semantics, kept.
// all threads of one CTA; the count is valid in thread 0
__device__ int conv_accumulate_body(int32_t P, int32_t me, int32_t converged_all_local,
                                    int32_t *conv, int32_t *const *peer_conv)
{
    const int t = threadIdx.x;
    if (converged_all_local > 0)
        for (int j = t; j < P; j += blockDim.x)
            atomicAdd_system(j == me ? conv : peer_conv[j], 1);
    __threadfence_system();
    __syncthreads();
    return ld_relaxed_sys_i32(conv + 0);
}

__global__ void conv_accumulate_kernel(int32_t P, int32_t me, int32_t converged_all_local,
                                       int32_t *conv, int32_t *const *peer_conv,
                                       int32_t *num_converged)
{
    const int num = conv_accumulate_body(P, me, converged_all_local, conv, peer_conv);
    if (threadIdx.x == 0) *num_converged = num;
}

void launch_conv_accumulate(const Ctx &ctx, int32_t P, int32_t me, int32_t converged_all_local,
                            int32_t *conv, int32_t *const *peer_conv, int32_t *num_converged)
{
    ctx.use();
    conv_accumulate_kernel<<<1, 64, 0, ctx.stream>>>(P, me, converged_all_local, conv, peer_conv,
                                                     num_converged);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

void launch_conv_tree(const Ctx &ctx, int32_t P, int32_t me, int32_t converged_all_local,
                      int32_t *conv, int32_t *const *peer_conv, int32_t *num_converged)
{
    ctx.use();
    conv_tree_kernel<<<1, 32, 0, ctx.stream>>>(P, me, converged_all_local, conv, peer_conv,
                                               num_converged);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// =============================================================================
// Device-side convergence decisions: what Solve::check_convergence does with host variables
// (source/solve.cpp:796-1005) and the break test of SchwarzBase::run
// (source/schwarz_base.cpp:424-433), on OuterState words the rest of the loop's launches honour.
// =============================================================================
__global__ void gather_norms_kernel(int32_t n, const double *const *__restrict__ norm_ptrs,
                                    double *__restrict__ out)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = *norm_ptrs[i];
}

void launch_gather_norms(const Ctx &ctx, int32_t n, const double *const *norm_ptrs, double *out)
{
    ctx.use();
    gather_norms_kernel<<<1, 64, 0, ctx.stream>>>(n, norm_ptrs, out);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

__device__ __forceinline__ bool outer_latch_norm(OuterState *S, double r, int32_t iter,
                                                 double *history_slot)
{
    S->resnorm = r;
    if (S->resnorm0 < 0.0) S->resnorm0 = r;
    if (history_slot) *history_slot = r;
    S->iter = iter + 1;
    if (isnan(r)) {
        S->error = OUTER_NAN;
        S->stop = 1;
        return false;
    }
    return true;
}

// Two-sided: allgather (done by the caller: `all`) + ordered sum + latch + ratio test
// (source/solve.cpp:888-912), one thread per local subdomain.
__global__ void ras_decide_kernel(int32_t P, int32_t nl, const double *__restrict__ all,
                                  const int32_t *__restrict__ slot, OuterState *const *states,
                                  double tol, int32_t check, int32_t enable_global_check,
                                  int32_t iter, double *history)
{
    __shared__ double s_g;
    if (threadIdx.x == 0) {
        double g = 0.0;
        for (int j = 0; j < P; ++j) {
            if (all[j] != DBL_MAX) {
                g += all[j];
            } else {
                g = -1.0;
                break;
            }
        }
        s_g = g;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nl; i += blockDim.x) {
        OuterState *S = states[i];
        if (S->stop) continue;   // finished earlier (or failed): leave its record alone
        if (!outer_latch_norm(S, all[slot[i]], iter, history ? history + (size_t)iter * nl + i : nullptr))
            continue;
        const double r = S->resnorm, r0 = S->resnorm0;
        int num_converged_p = (tol >= 0.0 && (r * r) / (r0 * r0) < tol * tol) ? 1 : 0;
        if (check) {
            if (enable_global_check) {
                S->gres = s_g;
                if (s_g >= 0.0) {
                    if (S->gres0 < 0.0) S->gres0 = s_g;
                    if (s_g / S->gres0 <= tol) num_converged_p = P;
                }
            } else {
                num_converged_p = 0;   // never raised outside the global check
            }
            S->num_converged = num_converged_p;
        }
        if (isnan(S->gres) || S->gres > 1e12) {   // schwarz_base.cpp:424-428
            S->error = OUTER_DIVERGED;
            S->stop = 1;
            continue;
        }
        if (S->num_converged == P) {   // :432-433
            S->finished_iter = iter;
            S->stop = 1;
        }
    }
}

void launch_ras_decide(const Ctx &ctx, int32_t P, int32_t nl, const double *all,
                       const int32_t *slot, OuterState *const *states, double tol,
                       int32_t check, int32_t enable_global_check, int32_t iter, double *history)
{
    ctx.use();
    ras_decide_kernel<<<1, 64, 0, ctx.stream>>>(P, nl, all, slot, states, tol, check,
                                                enable_global_check, iter, history);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// One-sided: ratio test on the local norm, flag protocol, break test (source/solve.cpp:913-943)
__global__ void ras_conv_decide_kernel(int32_t protocol, int32_t P, int32_t me, OuterState *S,
                                       const double *resnorm_dev, double tol, int32_t check,
                                       int32_t iter, double *history_slot, int32_t *conv,
                                       int32_t *conv_sent, int32_t n_out,
                                       int32_t *const *out_conv, int32_t *const *peer_conv,
                                       int32_t *num_converged)
{
    __shared__ int s_go;
    if (S->stop) return;
    if (threadIdx.x == 0) s_go = outer_latch_norm(S, *resnorm_dev, iter, history_slot) ? 1 : 0;
    __syncthreads();
    if (!s_go || !check) return;
    const int cal = (S->resnorm / S->resnorm0 <= tol) ? 1 : 0;
    int num = 0;
    if (protocol == 1) {
        if (threadIdx.x == 0) num = conv_tree_body(P, me, cal, conv, peer_conv);
    } else if (protocol == 2) {
        num = conv_accumulate_body(P, me, cal, conv, peer_conv);
    } else {
        num = conv_forward_body(P, me, cal, conv, conv_sent, n_out, out_conv);
    }
    if (threadIdx.x == 0) {
        *num_converged = num;
        S->num_converged = num;
        if (num == P) {
            S->finished_iter = iter;
            S->stop = 1;
        }
    }
}

void launch_ras_conv_decide(const Ctx &ctx, int32_t protocol, int32_t P, int32_t me,
                            OuterState *state, const double *resnorm_dev, double tol,
                            int32_t check, int32_t iter, double *history_slot, int32_t *conv,
                            int32_t *conv_sent, int32_t n_out, int32_t *const *out_conv,
                            int32_t *const *peer_conv, int32_t *num_converged)
{
    ctx.use();
    ras_conv_decide_kernel<<<1, 128, 0, ctx.stream>>>(protocol, P, me, state, resnorm_dev, tol,
                                                      check, iter, history_slot, conv, conv_sent,
                                                      n_out, out_conv, peer_conv, num_converged);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace schwz_b200
