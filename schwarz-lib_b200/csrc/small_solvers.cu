// One-launch local solvers for small subdomains (configurations 1 and 3 of
// BASELINE.json: 400..5100 rows per subdomain).  At that size a Krylov
// iteration is a few microseconds of work, so the multi-kernel solvers in
// solvers.cu are bound by launch latency (3 launches per CG iteration,
// 2k+7 per GMRES step).  Here ONE CTA (128, 256 or 1024 threads, by size) runs the whole solve:
// vectors live in shared memory (CG) or in L2-resident global memory (the
// GMRES basis), reductions are fixed-shape CTA reductions, the matrix streams
// from L1/L2, and the recurrences / stopping rules are exactly those of
// solvers.cu (Ginkgo semantics).  Rows are summed
// sequentially in stored order, as everywhere else.
#include <algorithm>
#include <cstdlib>

#include "engine.hpp"

namespace schwz_b200 {

// SCHWZ_B200_GMRES_MGS=1: modified Gram-Schmidt (Ginkgo's variant, k + 2 dependent reductions
// per Arnoldi step) in the one-CTA GMRES instead of CGS2 (3 barriers per step)
bool g_gmres_cgs2 = true;

constexpr int kSmallMaxWarps = 32;
constexpr int kSmallBuf = 2 * kSmallMaxWarps + 2;   // doubles of reduction scratch

// CTA-wide sum with ONE barrier: the warps leave their partial sums in one of two scratch rows
// (alternating, so that a row is never overwritten while a slow warp still reads it) and every
// warp then adds the partials up itself with the same fixed shuffle tree.  These solves are a
// chain of dependent reductions (k + 2 per GMRES step), so the barrier count is the run time.
template <int THREADS>
__device__ __forceinline__ double cta_sum(double v, double *buf, int &phase)
{
    constexpr int NW = THREADS / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    double *row = buf + phase * kSmallMaxWarps;
    phase ^= 1;
    if (lane == 0) row[w] = v;
    __syncthreads();
    if (NW <= 8) {   // a short chain of adds beats a second shuffle tree
        double s = row[0];
#pragma unroll
        for (int i = 1; i < NW; ++i) s += row[i];
        return s;
    }
    double s = lane < NW ? row[lane] : 0.0;
    return warp_sum(s);
}

__device__ __forceinline__ double row_dot(const int32_t *__restrict__ rp,
                                          const int32_t *__restrict__ ci,
                                          const double *__restrict__ v, const double *xs, int32_t row)
{
    double acc = 0.0;
    for (int32_t k = rp[row]; k < rp[row + 1]; ++k) acc += __ldg(v + k) * xs[__ldg(ci + k)];
    return acc;
}

// -----------------------------------------------------------------------------
// CG: x, r, p, q in shared memory.
// -----------------------------------------------------------------------------
template <int kSmallThreads>
__global__ void __launch_bounds__(kSmallThreads, 1)
    cg_small_kernel(int32_t n, const int32_t *__restrict__ rp, const int32_t *__restrict__ ci,
                    const double *__restrict__ v, const double *__restrict__ b, double *x,
                    int32_t max_iters, double tol, CgScalars *out, const int32_t *outer_stop)
{
    extern __shared__ __align__(16) double sm[];
    double *xs = sm, *r = sm + n, *p = sm + 2 * (size_t)n, *q = sm + 3 * (size_t)n;
    double *buf = sm + 4 * (size_t)n;
    int phase = 0;
    if (outer_stop != nullptr && *outer_stop != 0) return;
    const int t = threadIdx.x;
    for (int32_t i = t; i < n; i += kSmallThreads) {
        xs[i] = x[i];
        p[i] = 0.0;
    }
    __syncthreads();
    double part = 0.0;
    for (int32_t i = t; i < n; i += kSmallThreads) {
        // r = b - A x  (same expression order as the advanced SpMV: beta*y first)
        double acc = b[i];
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k) acc += (-__ldg(v + k)) * xs[__ldg(ci + k)];
        r[i] = acc;
        part += acc * acc;
    }
    double rho = cta_sum<kSmallThreads>(part, buf, phase);
    const double r0 = sqrt(rho);
    double prev_rho = 1.0, resnorm = r0;
    int iter = 0;
    while (true) {
        if (iter >= max_iters || resnorm < tol * r0) break;
        const bool fresh = (iter == 0) || (prev_rho == 0.0);
        const double tp = fresh ? 0.0 : rho / prev_rho;
        for (int32_t i = t; i < n; i += kSmallThreads) p[i] = fresh ? r[i] : r[i] + tp * p[i];
        __syncthreads();
        part = 0.0;
        for (int32_t i = t; i < n; i += kSmallThreads) {
            const double qi = row_dot(rp, ci, v, p, i);
            q[i] = qi;
            part += qi * p[i];
        }
        const double beta = cta_sum<kSmallThreads>(part, buf, phase);
        part = 0.0;
        if (beta != 0.0) {
            const double a = rho / beta;
            for (int32_t i = t; i < n; i += kSmallThreads) {
                xs[i] += a * p[i];
                const double ri = r[i] - a * q[i];
                r[i] = ri;
                part += ri * ri;
            }
        } else {
            for (int32_t i = t; i < n; i += kSmallThreads) part += r[i] * r[i];
        }
        const double rho_new = cta_sum<kSmallThreads>(part, buf, phase);
        prev_rho = rho;
        rho = rho_new;
        ++iter;
        resnorm = sqrt(rho_new);
    }
    __syncthreads();
    for (int32_t i = t; i < n; i += kSmallThreads) x[i] = xs[i];
    if (t == 0) {
        out->rho = rho;
        out->prev_rho = prev_rho;
        out->r0 = r0;
        out->resnorm = resnorm;
        out->tol = tol;
        out->iter = iter;
        out->max_iters = max_iters;
        out->pending = 0;
        out->alpha = 0.0;
        out->stop = 1;
    }
}

bool cg_small_fits(int64_t n) { return n > 0 && (4 * n + kSmallBuf) * 8 <= 220 * 1024; }

// threads by size: few warps make the barriers cheap, many warps hide the gathers
static int small_threads(int64_t n, bool few_barriers = false)
{
    static int forced = -1;
    if (forced < 0) {
        const char *e = std::getenv("SCHWZ_B200_SMALL_THREADS");   // A/B aid: 128, 256, 512 or 1024
        forced = e ? std::atoi(e) : 0;
    }
    if (forced == 128 || forced == 256 || forced == 512 || forced == 1024) return forced;
    // measured (tools/prof_small.py, profiles/r2_cfg3.md): with one CTA-wide reduction per
    // Gram-Schmidt projection (MGS, and CG) 457 rows take 10.8 / 14.5 / 14.5 us per GMRES(30)
    // step with 256 / 128 / 1024 threads; from 827 rows up 1024 threads win.  With CGS2 (three
    // barriers per step) the kernel is bound by the instructions each warp has to issue, and
    // more warps pay: cfg3 at P = 8 (457 rows) 377 / 456 / 442 outer iterations per second
    // with 256 / 512 / 1024 threads, P = 2 (1600 rows) 92 / 121 / 131.
    if (few_barriers) return n <= 640 ? 512 : 1024;
    return n <= 640 ? 256 : 1024;
}

void launch_cg_small(const Ctx &ctx, const DeviceCsr &A, const double *b, double *x,
                     int32_t max_iters, double tol, CgScalars *out, const int32_t *outer_stop)
{
    ctx.use();
    const size_t smem = (4 * (size_t)A.nrows + kSmallBuf) * sizeof(double);
    static std::atomic<bool> configured[64];
    if (!configured[ctx.device].load(std::memory_order_acquire)) {
        for (auto *k : {cg_small_kernel<128>, cg_small_kernel<256>, cg_small_kernel<512>,
                        cg_small_kernel<1024>})
            SCHWZ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            220 * 1024));
        configured[ctx.device].store(true, std::memory_order_release);
    }
    const int th = small_threads(A.nrows);
    auto *k = th == 128 ? cg_small_kernel<128> : th == 256 ? cg_small_kernel<256>
              : th == 512 ? cg_small_kernel<512> : cg_small_kernel<1024>;
    k<<<1, th, smem, ctx.stream>>>(A.nrows, A.rp, A.ci, A.v, b, x, max_iters, tol, out, outer_stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

// -----------------------------------------------------------------------------
// GMRES(m): w and the small Hessenberg system in shared memory; the Krylov basis V
// ((m+1) x n) too when it fits (cfg3: 31 x 460 doubles = 114 KB), else in global memory
// (L2 resident at these sizes).
// -----------------------------------------------------------------------------
template <int kSmallThreads>
__global__ void __launch_bounds__(kSmallThreads, 1)
    gmres_small_kernel(int32_t n, const int32_t *__restrict__ rp_g, const int32_t *__restrict__ ci_g,
                       const double *__restrict__ v_g, const double *__restrict__ b, double *x,
                       double *V, int32_t m, int32_t max_iters, double tol, double *resnorm_out,
                       double *r0_out, int32_t *total_out, int basis_in_smem, int matrix_in_smem,
                       const int32_t *outer_stop, int cgs2)
{
    extern __shared__ __align__(16) double sm[];
    int phase = 0;
    if (outer_stop != nullptr && *outer_stop != 0) return;
    const int32_t *rp = rp_g, *ci = ci_g;
    const double *v = v_g;
    double *w = sm;                               // n
    double *H = w + n;                            // (m+1)*m, column-major
    double *cs = H + (size_t)(m + 1) * m;         // m
    double *sn = cs + m;                          // m
    double *g = sn + m;                           // m+1
    double *y = g + m + 1;                        // m
    double *buf = y + m;                          // kSmallBuf
    double *hbuf = buf + kSmallBuf;               // m + 1: the projections of one CGS pass
    double *tail = hbuf + m + 1;
    if (basis_in_smem) {
        V = tail;                                 // (m+1)*n
        tail += (size_t)(m + 1) * n;
    }
    const int t = threadIdx.x;
    if (matrix_in_smem) {
        // the matrix is read once per Arnoldi step: keep it next to the vectors (with the
        // shared-memory carve-out this large, little L1 is left for it)
        const int32_t nnz = rp_g[n];
        double *sv = tail;                                   // nnz
        int32_t *sci = reinterpret_cast<int32_t *>(sv + nnz);   // nnz
        int32_t *srp = sci + nnz;                            // n + 1
        for (int32_t k = t; k < nnz; k += kSmallThreads) {
            sv[k] = v_g[k];
            sci[k] = ci_g[k];
        }
        for (int32_t k = t; k <= n; k += kSmallThreads) srp[k] = rp_g[k];
        rp = srp;
        ci = sci;
        v = sv;
    }

    // r = b - A x ; ||r|| ; V0 = r / ||r||   (x read from global: it changes at restarts)
    auto begin_cycle = [&]() -> double {
        __syncthreads();
        double part = 0.0;
        for (int32_t i = t; i < n; i += kSmallThreads) {
            double acc = b[i];
            for (int32_t k = rp[i]; k < rp[i + 1]; ++k) acc += (-v[k]) * x[ci[k]];
            w[i] = acc;
            part += acc * acc;
        }
        const double rn = sqrt(cta_sum<kSmallThreads>(part, buf, phase));
        for (int32_t i = t; i < n; i += kSmallThreads) V[i] = rn != 0.0 ? w[i] / rn : 0.0;
        if (t == 0) {
            for (int i = 0; i <= m; ++i) g[i] = 0.0;
            g[0] = rn;
        }
        __syncthreads();
        return rn;
    };
    // x += V(:, 0:k) y  with  H(0:k,0:k) y = g(0:k)
    auto update_x = [&](int k) {
        __syncthreads();
        if (t == 0) {
            for (int i = k - 1; i >= 0; --i) {
                double s = g[i];
                for (int j = i + 1; j < k; ++j) s -= H[(size_t)j * (m + 1) + i] * y[j];
                y[i] = s / H[(size_t)i * (m + 1) + i];
            }
        }
        __syncthreads();
        for (int32_t i = t; i < n; i += kSmallThreads) {
            double xv = x[i];
            for (int j = 0; j < k; ++j) xv += y[j] * V[(size_t)j * n + i];
            x[i] = xv;
        }
        __syncthreads();
    };

    double resnorm = begin_cycle();
    const double r0 = resnorm;
    int total = -1, k = 0;
    // Givens update of Hessenberg column kk (entries 0..kk set, hn below the diagonal): the new
    // implicit residual norm goes to buf[kSmallBuf - 1]
    auto givens = [&](int kk, double hn) {
        double *col = H + (size_t)kk * (m + 1);
        col[kk + 1] = hn;
        for (int i = 0; i < kk; ++i) {
            const double tt = cs[i] * col[i] + sn[i] * col[i + 1];
            col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1];
            col[i] = tt;
        }
        const double a = col[kk], c = col[kk + 1];
        if (a == 0.0) {
            cs[kk] = 0.0;
            sn[kk] = 1.0;
        } else {
            const double sc = fabs(a) + fabs(c);
            const double hyp = sc * sqrt((a / sc) * (a / sc) + (c / sc) * (c / sc));
            cs[kk] = a / hyp;
            sn[kk] = c / hyp;
        }
        col[kk] = cs[kk] * a + sn[kk] * c;
        col[kk + 1] = 0.0;
        g[kk + 1] = -sn[kk] * g[kk];
        g[kk] = cs[kk] * g[kk];
        buf[kSmallBuf - 1] = fabs(g[kk + 1]);
    };
    if (cgs2) {
        // ---- CGS2 variant.  Per Arnoldi step: the Givens update of the PREVIOUS column (one
        // thread, a chain of dependent fp64 operations) runs beside the SpMV of this step, the
        // k + 1 projections of a Gram-Schmidt pass are taken by the warps side by side (a warp
        // sums whole dot products with 4 independent partial sums: no CTA-wide reduction), one
        // barrier ends a pass; two passes keep the basis orthogonal to working precision.
        // Stopping rule, restart schedule and the quantities handed back are those of the
        // modified variant below (Ginkgo's), the arithmetic inside a step is not.
        constexpr int NW = kSmallThreads / 32;
        const int lane = t & 31, wid = t >> 5;
        bool pending = false;   // a Givens update (column k - 1) is still to be done
        double hn_prev = 0.0;
        while (true) {
            // w = A V_k beside the pending Givens update (the last thread does that one and
            // leaves the rows to the others); the stop test below needs its result
            if (k < m) {
                const double *vk = V + (size_t)k * n;
                const int workers = pending ? kSmallThreads - 1 : kSmallThreads;
                if (t < workers)
                    for (int32_t i = t; i < n; i += workers) {
                        double acc = 0.0;
                        for (int32_t q = rp[i]; q < rp[i + 1]; ++q) acc = fma(v[q], vk[ci[q]], acc);
                        w[i] = acc;
                    }
            }
            if (pending && t == kSmallThreads - 1) givens(k - 1, hn_prev);
            __syncthreads();
            if (pending) resnorm = buf[kSmallBuf - 1];
            pending = false;
            ++total;
            if (total >= max_iters || resnorm < tol * r0) break;
            if (k == m) {
                update_x(k);
                resnorm = begin_cycle();
                k = 0;
                const double *v0 = V;
                for (int32_t i = t; i < n; i += kSmallThreads) {
                    double acc = 0.0;
                    for (int32_t q = rp[i]; q < rp[i + 1]; ++q) acc = fma(v[q], v0[ci[q]], acc);
                    w[i] = acc;
                }
                __syncthreads();
            }
            double *col = H + (size_t)k * (m + 1);
            for (int pass = 0; pass < 2; ++pass) {
                // a warp takes the projections i = wid, wid + NW, ... two at a time: every
                // element of w is read once for both, the partial sums are fused multiply-adds
                for (int i = wid; i <= k; i += 2 * NW) {
                    const int i2 = i + NW;
                    const bool two = i2 <= k;
                    const double *va = V + (size_t)i * n;
                    const double *vb = V + (size_t)(two ? i2 : i) * n;
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
                    int32_t j = lane;
                    for (; j + 32 < n; j += 64) {
                        const double w0 = w[j], w1 = w[j + 32];
                        a0 = fma(w0, va[j], a0);
                        a1 = fma(w1, va[j + 32], a1);
                        b0 = fma(w0, vb[j], b0);
                        b1 = fma(w1, vb[j + 32], b1);
                    }
                    if (j < n) {
                        const double w0 = w[j];
                        a0 = fma(w0, va[j], a0);
                        b0 = fma(w0, vb[j], b0);
                    }
                    const double pa = warp_sum(a0 + a1), pb = warp_sum(b0 + b1);
                    if (lane == 0) {
                        hbuf[i] = pa;
                        if (two) hbuf[i2] = pb;
                    }
                }
                __syncthreads();
                for (int32_t j = t; j < n; j += kSmallThreads) {
                    double a0 = w[j], a1 = 0.0;
                    int i = 0;
                    for (; i + 1 <= k; i += 2) {
                        a0 = fma(-hbuf[i], V[(size_t)i * n + j], a0);
                        a1 = fma(-hbuf[i + 1], V[(size_t)(i + 1) * n + j], a1);
                    }
                    if (i <= k) a0 = fma(-hbuf[i], V[(size_t)i * n + j], a0);
                    w[j] = a0 + a1;
                }
                if (t <= k) col[t] = pass == 0 ? hbuf[t] : col[t] + hbuf[t];
                __syncthreads();
            }
            double part = 0.0;
            for (int32_t j = t; j < n; j += kSmallThreads) part = fma(w[j], w[j], part);
            const double hn = sqrt(cta_sum<kSmallThreads>(part, buf, phase));
            const double inv = hn != 0.0 ? 1.0 / hn : 0.0;
            double *vn = V + (size_t)(k + 1) * n;
            for (int32_t j = t; j < n; j += kSmallThreads) vn[j] = w[j] * inv;
            hn_prev = hn;
            pending = true;
            ++k;
            __syncthreads();   // V_{k+1} complete (and visible when it lives in global memory)
        }
    } else
    while (true) {
        ++total;
        if (total >= max_iters || resnorm < tol * r0) break;
        if (k == m) {
            update_x(k);
            resnorm = begin_cycle();
            k = 0;
        }
        // w = A V_k  (V_k read through a shared copy: stage it in w's place is not
        // possible, so rows gather from global/L2)
        const double *vk = V + (size_t)k * n;
        for (int32_t i = t; i < n; i += kSmallThreads) {
            double acc = 0.0;
            for (int32_t q = rp[i]; q < rp[i + 1]; ++q) acc += v[q] * vk[ci[q]];
            w[i] = acc;
        }
        double *col = H + (size_t)k * (m + 1);
        for (int i = 0; i <= k; ++i) {   // modified Gram-Schmidt
            const double *vi = V + (size_t)i * n;
            double part = 0.0;
            for (int32_t j = t; j < n; j += kSmallThreads) part += w[j] * vi[j];
            const double h = cta_sum<kSmallThreads>(part, buf, phase);
            if (t == 0) col[i] = h;
            for (int32_t j = t; j < n; j += kSmallThreads) w[j] += (-h) * vi[j];
        }
        double part = 0.0;
        for (int32_t j = t; j < n; j += kSmallThreads) part += w[j] * w[j];
        const double hn = sqrt(cta_sum<kSmallThreads>(part, buf, phase));
        double *vn = V + (size_t)(k + 1) * n;
        for (int32_t j = t; j < n; j += kSmallThreads) vn[j] = hn != 0.0 ? w[j] / hn : 0.0;
        if (t == 0) givens(k, hn);
        __syncthreads();   // also makes V_{k+1} (global) visible to the whole CTA
        resnorm = buf[kSmallBuf - 1];
        ++k;
    }
    update_x(k);
    if (t == 0) {
        *resnorm_out = resnorm;
        *r0_out = r0;
        *total_out = total;
    }
}

static size_t gmres_small_smem(int64_t n, int m, bool basis, int64_t nnz_in_smem = -1)
{
    size_t bytes = ((size_t)n + (size_t)(m + 1) * m + 5 * (size_t)m + 2 + kSmallBuf + 8 +
                    (basis ? (size_t)(m + 1) * n : 0)) * sizeof(double);
    if (nnz_in_smem >= 0) bytes += 12 * (size_t)nnz_in_smem + 4 * ((size_t)n + 1) + 16;
    return bytes;
}

bool gmres_small_fits(int64_t n, int m)
{
    return n > 0 && n <= 16384 && gmres_small_smem(n, m, false) <= 220 * 1024;
}

void launch_gmres_small(const Ctx &ctx, const DeviceCsr &A, const double *b, double *x, double *V,
                        int32_t m, int32_t max_iters, double tol, double *resnorm_out,
                        double *r0_out, int32_t *total_out, const int32_t *outer_stop)
{
    ctx.use();
    static std::atomic<bool> configured[64];
    if (!configured[ctx.device].load(std::memory_order_acquire)) {
        for (auto *k : {gmres_small_kernel<128>, gmres_small_kernel<256>, gmres_small_kernel<512>,
                        gmres_small_kernel<1024>})
            SCHWZ_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            220 * 1024));
        configured[ctx.device].store(true, std::memory_order_release);
    }
    const bool basis = gmres_small_smem(A.nrows, m, true) <= 220 * 1024;
    const bool matrix = gmres_small_smem(A.nrows, m, basis, A.nnz) <= 220 * 1024;
    const int th = small_threads(A.nrows, g_gmres_cgs2);
    auto *k = th == 128 ? gmres_small_kernel<128> : th == 256 ? gmres_small_kernel<256>
              : th == 512 ? gmres_small_kernel<512> : gmres_small_kernel<1024>;
    k<<<1, th, gmres_small_smem(A.nrows, m, basis, matrix ? A.nnz : -1), ctx.stream>>>(
        A.nrows, A.rp, A.ci, A.v, b, x, V, m, max_iters, tol, resnorm_out, r0_out, total_out,
        basis ? 1 : 0, matrix ? 1 : 0, outer_stop, g_gmres_cgs2 ? 1 : 0);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace schwz_b200
