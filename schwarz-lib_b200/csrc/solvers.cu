// Device-resident local solvers: CG, restarted GMRES(m), level-scheduled
// sparse triangular solves.  No scalar ever travels to the host inside a
// solve; stopping decisions are device flags that turn the remaining launches
// into no-ops.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "engine.hpp"

namespace schwz_b200 {

// =============================================================================
// CG
// =============================================================================
bool g_use_small_solvers = true;
bool g_use_cg_graph = true;   // SCHWZ_B200_NO_CG_GRAPH=1 turns the graph replay off (A/B)

constexpr int kCgNoPoll = 160;   // up to this many iterations: enqueue all, never poll
constexpr int kCgChunk = 32;     // otherwise poll the stop flag once per chunk
constexpr int kCgWhileUnroll = 10;   // CG iterations per trip of the WHILE graph

CgSolver::CgSolver(const Ctx &ctx, const DeviceCsr &A) : ctx_(ctx), A_(A), n_(A.nrows)
{
    SCHWZ_REQUIRE(A.nrows == A.ncols, "CG needs a square matrix");
    r_ = ctx.alloc<double>(n_);
    p_ = ctx.alloc_zero<double>(n_);
    q_ = ctx.alloc<double>(n_);
    s_ = (CgScalars *)ctx.alloc_zero<char>(sizeof(CgScalars));
    SCHWZ_CUDA(cudaMallocHost((void **)&pinned_stop_, 2 * sizeof(int32_t)));
    for (auto &e : ev_) SCHWZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

CgSolver::~CgSolver()
{
    ctx_.release(r_);
    ctx_.release(p_);
    ctx_.release(q_);
    ctx_.release(z_);
    ctx_.release(s_);
    if (pinned_stop_) cudaFreeHost(pinned_stop_);
    for (auto &e : ev_)
        if (e) cudaEventDestroy(e);
    for (const Captured &c : graphs_) cudaGraphExecDestroy(c.exec);
}

void CgSolver::set_precond(Preconditioner *M)
{
    for (const Captured &c : graphs_) cudaGraphExecDestroy(c.exec);
    graphs_.clear();
    M_ = M;
    if (M_ && !z_) z_ = ctx_.alloc_zero<double>(n_);
}

int64_t CgSolver::bytes_per_iteration() const
{
    // SpMV (q = A p, p.q fused) + r update (2 reads, 1 write) + x/p update (3 reads,
    // 2 writes): SURVEY.md 8(d) minus the traffic the fusions remove
    return 12 * A_.nnz + 4 * (n_ + 1) + 16 * n_ + 24 * n_ + 40 * n_ +
           (M_ ? M_->bytes_per_apply() : 0);
}

void CgSolver::iteration(double *x, cudaGraphConditionalHandle loop)
{
    if (M_) {
        // z = M^-1 r with rho = r.z fused; the x/r update below then leaves rho alone and
        // only produces ||r|| for the stop test (which Ginkgo evaluates after rho, before
        // the p update - the order of the two does not change any number)
        M_->apply(r_, z_, &s_->rho, &s_->stop);
        launch_cg_xp_update(ctx_, n_, z_, p_, x, s_);
    } else {
        launch_cg_xp_update(ctx_, n_, r_, p_, x, s_);
    }
    launch_spmv(ctx_, A_, 1.0, p_, 0.0, nullptr, q_, EPI_DOT, p_, &s_->beta, (int32_t)n_,
                &s_->stop);
    launch_cg_r_update(ctx_, n_, r_, q_, s_, M_ != nullptr, loop);
}

void CgSolver::solve(const double *b, double *x, int32_t max_iters, double tol,
                     const int32_t *outer_stop)
{
    SCHWZ_REQUIRE(((uintptr_t)x & 15) == 0, "CG solution vector must be 16-byte aligned");
    if (g_use_small_solvers && cg_small_fits(n_) && !M_) {
        // whole solve in one launch (small_solvers.cu)
        launch_cg_small(ctx_, A_, b, x, max_iters, tol, s_, outer_stop);
        return;
    }
    auto enqueue_all = [&]() {
        // r = b - A x, rho = r.r fused (behind a full dependency on whatever produced b and x;
        // every later launch of the solve may overlap its predecessor's tail, see PdlScope -
        // with a preconditioner in between its own kernels serialise as usual)
        launch_spmv(ctx_, A_, -1.0, x, 1.0, b, r_, EPI_NRM2SQ, nullptr, &s_->rho, (int32_t)n_,
                    outer_stop);
        PdlScope pdl(true);
        launch_cg_init(ctx_, s_, max_iters, tol, outer_stop);
        for (int it = 0; it < max_iters; ++it) iteration(x);
        launch_cg_flush_x(ctx_, n_, x, p_, s_);
    };
    // Two graph shapes, both built at the second call with the same arguments:
    //  * unrolled: initial residual, max_iters iterations, flush - a fixed launch sequence
    //    whose kernels turn into no-ops once the stop flag is up.  Best when the budget is
    //    what ends the solve (local_tol = 1e-12 with local_max_iters = 50);
    //  * while:    the iteration is the body of a WHILE conditional node that the r update
    //    (which takes the stop decision) re-arms from the device: the graph leaves the loop
    //    after the last live iteration instead of running the rest of the budget as no-ops.
    //    Best for inexact local solves (local_tol = 0.1 stops after a handful of iterations).
    //    Also the only graph shape for budgets too long to unroll (local_max_iters = -1), which
    //    otherwise need the host to poll the stop flag.
    // Measured (profiles/r2_while_graph.md): a trip of the loop node costs ~5 us.  With 8
    // subdomain streams on one GPU that is invisible (cfg2: 83.8 ms per outer iteration as a WHILE
    // graph against 83.4 - 84.5 unrolled), with one subdomain per GPU it is not (524 k rows, 50
    // iterations: 1.38 against 1.13 ms; 2.1 M rows: 2.75 against 2.48).  So: a solve that its
    // BUDGET will end (tight tolerance, at most kCgNoPoll iterations) is unrolled, everything
    // else - inexact solves that stop on their tolerance, budgets too long to unroll - runs as a
    // WHILE graph with kCgWhileUnroll iterations per trip.  SCHWZ_B200_CG_WHILE=0 / 1 forces one.
    const char *force_while_s = std::getenv("SCHWZ_B200_CG_WHILE");
    const int force_while = force_while_s ? std::atoi(force_while_s) : -1;
    const bool as_while = force_while >= 0 ? (force_while != 0 || max_iters > kCgNoPoll)
                                           : (tol >= 1e-4 || max_iters > kCgNoPoll);
    // (the level-per-launch triangular solves of the ILU preconditioner are graphs of their
    // own and cannot be captured; the one-kernel solves can)
    const bool ilu_levels = M_ && M_->kind() == PRECOND_ILU && M_->uses_level_graphs();
    const bool graphable = g_use_cg_graph && !not_graphable_ && !ilu_levels &&
                           (as_while || max_iters <= kCgNoPoll);
    if (graphable) {
        // The whole solve is a launch sequence on fixed buffers (the stop decisions are
        // device flags), so from the second call on it is replayed as ONE CUDA graph: at
        // mid-size subdomains (10^5..10^6 rows, kernels of a few microseconds) the 3*max_iters
        // launches are otherwise bound by launch latency.  Not with the ILU preconditioner,
        // whose triangular solves are graphs of their own.
        if (++plain_solves_ < 2) {
            if (max_iters <= kCgNoPoll) {
                enqueue_all();
                return;
            }
        } else {
            for (const Captured &c : graphs_)
                if (c.b == b && c.x == x && c.max_iters == max_iters && c.tol == tol &&
                    c.outer_stop == outer_stop) {
                    SCHWZ_CUDA(cudaGraphLaunch(c.exec, ctx_.stream));
                    count_launch(c.launches);
                    return;
                }
            if (graphs_.size() >= 2) {
                cudaGraphExecDestroy(graphs_.front().exec);
                graphs_.erase(graphs_.begin());
            }
            ctx_.use();
            const int64_t before = g_launches.load();
            cudaGraph_t graph = nullptr;
            bool capturing = false;
            try {
                if (as_while) {
                    SCHWZ_CUDA(cudaGraphCreate(&graph, 0));
                    cudaGraphConditionalHandle loop = 0;
                    SCHWZ_CUDA(cudaGraphConditionalHandleCreate(&loop, graph, 0, 0));
                    SCHWZ_CUDA(cudaStreamBeginCaptureToGraph(ctx_.stream, graph, nullptr, nullptr, 0,
                                                             cudaStreamCaptureModeThreadLocal));
                    capturing = true;
                    launch_spmv(ctx_, A_, -1.0, x, 1.0, b, r_, EPI_NRM2SQ, nullptr, &s_->rho,
                                (int32_t)n_, outer_stop);
                    launch_cg_init(ctx_, s_, max_iters, tol, outer_stop, loop);
                    cudaStreamCaptureStatus status;
                    const cudaGraphNode_t *deps = nullptr;
                    size_t ndeps = 0;
                    cudaGraph_t cap = nullptr;
                    SCHWZ_CUDA(cudaStreamGetCaptureInfo(ctx_.stream, &status, nullptr, &cap, &deps,
                                                        &ndeps));
                    cudaGraphNodeParams np = {};
                    np.type = cudaGraphNodeTypeConditional;
                    np.conditional.handle = loop;
                    np.conditional.type = cudaGraphCondTypeWhile;
                    np.conditional.size = 1;
                    cudaGraphNode_t node;
                    SCHWZ_CUDA(cudaGraphAddNode(&node, cap, deps, ndeps, &np));
                    cudaGraph_t body = np.conditional.phGraph_out[0];
                    SCHWZ_CUDA(cudaStreamUpdateCaptureDependencies(ctx_.stream, &node, 1,
                                                                   cudaStreamSetCaptureDependencies));
                    launch_cg_flush_x(ctx_, n_, x, p_, s_);
                    SCHWZ_CUDA(cudaStreamEndCapture(ctx_.stream, &graph));
                    capturing = false;
                    SCHWZ_CUDA(cudaStreamBeginCaptureToGraph(ctx_.stream, body, nullptr, nullptr, 0,
                                                             cudaStreamCaptureModeThreadLocal));
                    capturing = true;
                    // kCgWhileUnroll iterations per trip: a trip of the loop node costs ~5 us
                    // (1.13 -> 1.38 ms per 50-iteration solve at 524 k rows with one iteration
                    // per trip, profiles/r2_while_graph.md); iterations past the stop decision
                    // inside a trip are no-ops and leave the loop condition alone
                    for (int u = 0; u < std::min(kCgWhileUnroll, std::max(max_iters, 1)); ++u)
                        iteration(x, loop);
                    cudaGraph_t ignored = nullptr;
                    SCHWZ_CUDA(cudaStreamEndCapture(ctx_.stream, &ignored));
                    capturing = false;
                } else {
                    SCHWZ_CUDA(cudaStreamBeginCapture(ctx_.stream, cudaStreamCaptureModeThreadLocal));
                    capturing = true;
                    enqueue_all();
                    SCHWZ_CUDA(cudaStreamEndCapture(ctx_.stream, &graph));
                    capturing = false;
                }
                Captured c{b, x, max_iters, tol, outer_stop, nullptr,
                           (int)(g_launches.load() - before)};
                SCHWZ_CUDA(cudaGraphInstantiate(&c.exec, graph, 0));
                cudaGraphDestroy(graph);
                graph = nullptr;
                graphs_.push_back(c);
                SCHWZ_CUDA(cudaGraphLaunch(c.exec, ctx_.stream));
                return;
            } catch (...) {
                // never leave the stream in capture mode: end the capture, drop the partial
                // graph, remember not to try again, and solve without a graph below
                if (capturing) {
                    cudaGraph_t g2 = nullptr;
                    cudaStreamEndCapture(ctx_.stream, &g2);
                    if (g2 && g2 != graph) cudaGraphDestroy(g2);
                }
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                not_graphable_ = true;
                if (max_iters <= kCgNoPoll) {
                    enqueue_all();
                    return;
                }
            }
        }
    } else if (max_iters <= kCgNoPoll) {
        enqueue_all();
        return;
    }
    // r = b - A x, rho = r.r fused
    launch_spmv(ctx_, A_, -1.0, x, 1.0, b, r_, EPI_NRM2SQ, nullptr, &s_->rho, (int32_t)n_,
                outer_stop);
    launch_cg_init(ctx_, s_, max_iters, tol, outer_stop);
    int done = 0, chunk = 0;
    pinned_stop_[0] = pinned_stop_[1] = 0;
    while (done < max_iters) {
        const int m = std::min(kCgChunk, max_iters - done);
        for (int it = 0; it < m; ++it) iteration(x);
        done += m;
        const int slot = chunk & 1;
        SCHWZ_CUDA(cudaMemcpyAsync(pinned_stop_ + slot, &s_->stop, sizeof(int32_t),
                                   cudaMemcpyDeviceToHost, ctx_.stream));
        SCHWZ_CUDA(cudaEventRecord(ev_[slot], ctx_.stream));
        if (chunk > 0) {   // look at the previous chunk while this one runs
            SCHWZ_CUDA(cudaEventSynchronize(ev_[slot ^ 1]));
            if (pinned_stop_[slot ^ 1]) break;
        }
        ++chunk;
    }
    launch_cg_flush_x(ctx_, n_, x, p_, s_);
}

void CgSolver::bench_step(int kind, double *scratch_x)
{
    // timing aid: runs one vector step on the solver's own vectors; results are meaningless.
    // kind < 0: prepare - lift the iteration cap and the tolerance and clear the stop flag
    // once, so that the timed launches run back to back without a memset in between and keep
    // taking their live branches (the scalars of the last real solve stay in place)
    if (kind < 0) {
        const int32_t cap = 0x7fffffff, zero = 0;
        const double tol = 0.0;
        SCHWZ_CUDA(cudaMemcpyAsync(&s_->max_iters, &cap, sizeof(cap), cudaMemcpyHostToDevice, ctx_.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(&s_->tol, &tol, sizeof(tol), cudaMemcpyHostToDevice, ctx_.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(&s_->stop, &zero, sizeof(zero), cudaMemcpyHostToDevice, ctx_.stream));
        // a pending x update with alpha = 0: the x/p kernel streams x as in a live iteration
        const int32_t one = 1;
        SCHWZ_CUDA(cudaMemcpyAsync(&s_->pending, &one, sizeof(one), cudaMemcpyHostToDevice, ctx_.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(&s_->alpha, &tol, sizeof(tol), cudaMemcpyHostToDevice, ctx_.stream));
        SCHWZ_CUDA(cudaStreamSynchronize(ctx_.stream));
        return;
    }
    if (kind == 1) launch_cg_r_update(ctx_, n_, r_, q_, s_, false);
    else launch_cg_xp_update(ctx_, n_, r_, p_, scratch_x, s_);
}

void CgSolver::result(int32_t *iters, double *resnorm, double *resnorm0)
{
    CgScalars h;
    ctx_.use();
    SCHWZ_CUDA(cudaMemcpyAsync(&h, s_, sizeof(h), cudaMemcpyDeviceToHost, ctx_.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (iters) *iters = h.iter;
    if (resnorm) *resnorm = h.resnorm;
    if (resnorm0) *resnorm0 = h.r0;
}

// =============================================================================
// GMRES(m): restarted, modified Gram-Schmidt, Givens rotations, implicit
// residual norm in the stopping test.  The small
// Hessenberg system lives in device memory and is advanced by one-thread
// kernels; vector work is dot / axpy kernels with device scalars.
// small_ layout: [0] resnorm [1] r0 [2] tol [3] hn [4] tmp
//                H (m+1)*m at 8, cs m, sn m, g m+1, y m ; ints at the tail
// =============================================================================
struct GmresState {
    double resnorm, r0, tol, hn, tmp;
    int32_t total, k, stop, max_iters, m, need_restart, pad0, pad1;
};

__device__ __forceinline__ double *gm_H(double *base, int m) { return base; }
__device__ __forceinline__ double *gm_cs(double *base, int m) { return base + (size_t)(m + 1) * m; }
__device__ __forceinline__ double *gm_sn(double *base, int m) { return gm_cs(base, m) + m; }
__device__ __forceinline__ double *gm_g(double *base, int m) { return gm_sn(base, m) + m; }
__device__ __forceinline__ double *gm_y(double *base, int m) { return gm_g(base, m) + m + 1; }

// after r = b - A x with ||r|| in st->tmp: start a cycle
__global__ void gmres_begin_cycle_kernel(GmresState *st, double *small, int first,
                                         int32_t max_iters, double tol, int32_t m,
                                         const int32_t *outer_stop)
{
    if (!first && st->stop) return;
    const double rn = st->tmp;
    if (first) {
        st->r0 = rn;
        st->tol = tol;
        st->total = -1;
        st->max_iters = max_iters;
        st->m = m;
        st->stop = 0;
        if (outer_stop != nullptr && *outer_stop != 0) {   // the whole solve is a no-op
            st->stop = 1;
            st->k = 0;
            st->need_restart = 0;
            return;
        }
    }
    st->resnorm = rn;
    st->k = 0;
    st->need_restart = 0;
    double *g = gm_g(small, m);
    for (int i = 0; i <= m; ++i) g[i] = 0.0;
    g[0] = rn;
}

// v0 = r / ||r||  (zero when the norm is zero)
__global__ void __launch_bounds__(kBlock)
    gmres_scale_kernel(int64_t n, const double *__restrict__ src, double *__restrict__ dst,
                       const double *norm, const GmresState *st, int check_stop)
{
    if (check_stop && st->stop) return;
    const double nv = *norm;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock)
        dst[i] = nv != 0.0 ? src[i] / nv : 0.0;
}

// top of the loop: ++total ; stop test ; restart request
__global__ void gmres_top_kernel(GmresState *st)
{
    if (st->stop) return;
    st->total += 1;
    if (st->total >= st->max_iters || st->resnorm < st->tol * st->r0) {
        st->stop = 1;
        return;
    }
    st->need_restart = (st->k == st->m) ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock)
    gmres_dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                     double *partials, unsigned int *ticket, double *result, int do_sqrt,
                     const GmresState *st)
{
    __shared__ double s_warp[kBlock / 32];
    if (st->stop) return;
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

// One modified-Gram-Schmidt step fused with the NEXT projection: w -= h v (h a device scalar the
// previous launch left), and in the same pass over w the dot product the next step needs -
// w.v_next, or ||w|| after the last projection (v_next == nullptr).  Same operands, same
// thread-to-element mapping and the same partial-sum tree as a separate w += (-h) v pass followed
// by gmres_dot_kernel (round 1), so the doubles are the same; w is streamed once instead of twice
// (24 instead of 40 B/row per step) and a step is one launch instead of two.
__global__ void __launch_bounds__(kBlock)
    gmres_axpy_dot_kernel(int64_t n, const double *h, const double *__restrict__ v,
                          double *__restrict__ w, const double *__restrict__ v_next,
                          double *partials, unsigned int *ticket, double *result,
                          const GmresState *st)
{
    __shared__ double s_warp[kBlock / 32];
    if (st->stop) return;
    const double a = -1.0 * (*h);
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock) {
        const double wi = w[i] + a * v[i];
        w[i] = wi;
        s += wi * (v_next != nullptr ? v_next[i] : wi);
    }
    s = block_sum(s, s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    if (last_cta(ticket)) {
        double r = reduce_partials(partials, gridDim.x, s_warp);
        if (threadIdx.x == 0) *result = v_next != nullptr ? r : sqrt(r);
    }
}

// Givens update of column k (Hessenberg entries already in H(0..k,k), hn in st)
__global__ void gmres_givens_kernel(GmresState *st, double *small)
{
    if (st->stop) return;
    const int m = st->m, k = st->k;
    double *H = gm_H(small, m), *cs = gm_cs(small, m), *sn = gm_sn(small, m), *g = gm_g(small, m);
    double *col = H + (size_t)k * (m + 1);
    col[k + 1] = st->hn;
    for (int i = 0; i < k; ++i) {
        const double t = cs[i] * col[i] + sn[i] * col[i + 1];
        col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1];
        col[i] = t;
    }
    const double a = col[k], c = col[k + 1];
    if (a == 0.0) {
        cs[k] = 0.0;
        sn[k] = 1.0;
    } else {
        const double sc = fabs(a) + fabs(c);
        const double hyp = sc * sqrt((a / sc) * (a / sc) + (c / sc) * (c / sc));
        cs[k] = a / hyp;
        sn[k] = c / hyp;
    }
    col[k] = cs[k] * a + sn[k] * c;
    col[k + 1] = 0.0;
    g[k + 1] = -sn[k] * g[k];
    g[k] = cs[k] * g[k];
    st->resnorm = fabs(g[k + 1]);
    st->k = k + 1;
}

// back substitution for the first k columns -> y
__global__ void gmres_backsolve_kernel(const GmresState *st, double *small, int only_if_restart)
{
    if (only_if_restart && (st->stop || !st->need_restart)) return;
    const int m = st->m, k = st->k;
    double *H = gm_H(small, m), *g = gm_g(small, m), *y = gm_y(small, m);
    for (int i = k - 1; i >= 0; --i) {
        double s = g[i];
        for (int j = i + 1; j < k; ++j) s -= H[(size_t)j * (m + 1) + i] * y[j];
        y[i] = s / H[(size_t)i * (m + 1) + i];
    }
}

// x += sum_{j<k} y[j] V_j   (column order j ascending per element)
__global__ void __launch_bounds__(kBlock)
    gmres_update_x_kernel(int64_t n, const GmresState *st, const double *small,
                          const double *__restrict__ V, double *__restrict__ x,
                          int only_if_restart)
{
    if (only_if_restart && (st->stop || !st->need_restart)) return;
    const int m = st->m, k = st->k;
    const double *y = gm_y(const_cast<double *>(small), m);
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock) {
        double xv = x[i];
        for (int j = 0; j < k; ++j) xv += y[j] * V[(size_t)j * n + i];
        x[i] = xv;
    }
}

GmresSolver::GmresSolver(const Ctx &ctx, const DeviceCsr &A, int32_t restart)
    : ctx_(ctx), A_(A), n_(A.nrows), m_(std::max(1, restart))
{
    V_ = ctx.alloc<double>((size_t)(m_ + 1) * n_);
    w_ = ctx.alloc<double>(n_);
    const size_t small = (size_t)(m_ + 1) * m_ + 4 * (size_t)m_ + 8;
    small_ = ctx.alloc_zero<double>(small + sizeof(GmresState) / sizeof(double) + 2);
    SCHWZ_CUDA(cudaMallocHost((void **)&pinned_stop_, 2 * sizeof(int32_t)));
    for (auto &e : ev_) SCHWZ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

GmresSolver::~GmresSolver()
{
    ctx_.release(V_);
    ctx_.release(w_);
    ctx_.release(pv_);
    ctx_.release(upd_);
    ctx_.release(small_);
    if (pinned_stop_) cudaFreeHost(pinned_stop_);
    for (auto &e : ev_)
        if (e) cudaEventDestroy(e);
}

void GmresSolver::set_precond(Preconditioner *M)
{
    M_ = M;
    if (M_ && !pv_) {
        pv_ = ctx_.alloc_zero<double>(n_);
        upd_ = ctx_.alloc_zero<double>(n_);
    }
}

// x += pv  (the preconditioned update M^-1 (V y)); same guard as gmres_update_x_kernel
__global__ void __launch_bounds__(kBlock)
    gmres_add_kernel(int64_t n, const GmresState *st, const double *__restrict__ pv,
                     double *__restrict__ x, int only_if_restart)
{
    if (only_if_restart && (st->stop || !st->need_restart)) return;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kBlock)
        x[i] += pv[i];
}

static int vgrid(const Ctx &ctx, int64_t n)
{
    return (int)std::max<int64_t>(1, std::min<int64_t>((n + kBlock - 1) / kBlock, ctx.vec_grid()));
}

void GmresSolver::solve(const double *b, double *x, int32_t max_iters, double tol,
                        const int32_t *outer_stop)
{
    ctx_.use();
    cudaStream_t st = ctx_.stream;
    const size_t small = (size_t)(m_ + 1) * m_ + 4 * (size_t)m_ + 8;
    GmresState *S = (GmresState *)(small_ + small);
    double *H = small_;
    const int g = vgrid(ctx_, n_);
    if (g_use_small_solvers && gmres_small_fits(n_, m_) && !M_) {
        launch_gmres_small(ctx_, A_, b, x, V_, m_, max_iters, tol, &S->resnorm, &S->r0, &S->total,
                           outer_stop);
        return;
    }

    auto begin_cycle = [&](int first) {
        // w = b - A x ; tmp = ||w|| ; V0 = w / ||w||
        launch_spmv(ctx_, A_, -1.0, x, 1.0, b, w_, EPI_NRM2, nullptr, &S->tmp, (int32_t)n_,
                    outer_stop);
        gmres_begin_cycle_kernel<<<1, 1, 0, st>>>(S, small_, first, max_iters, tol, m_, outer_stop);
        gmres_scale_kernel<<<g, kBlock, 0, st>>>(n_, w_, V_, &S->tmp, S, 1);
        count_launch(2);
    };
    // x += V y, or x += M^-1 (V y) with a preconditioner (V y accumulated from zero, columns
    // ascending, as Ginkgo's calculate_qy does)
    auto update_x = [&](int only_if_restart) {
        gmres_backsolve_kernel<<<1, 1, 0, st>>>(S, small_, only_if_restart);
        if (!M_) {
            gmres_update_x_kernel<<<g, kBlock, 0, st>>>(n_, S, small_, V_, x, only_if_restart);
            count_launch(2);
            return;
        }
        SCHWZ_CUDA(cudaMemsetAsync(upd_, 0, sizeof(double) * (size_t)n_, st));
        gmres_update_x_kernel<<<g, kBlock, 0, st>>>(n_, S, small_, V_, upd_, only_if_restart);
        M_->apply(upd_, pv_, nullptr, nullptr);
        gmres_add_kernel<<<g, kBlock, 0, st>>>(n_, S, pv_, x, only_if_restart);
        count_launch(3);
    };
    begin_cycle(1);
    pinned_stop_[0] = pinned_stop_[1] = 0;
    int chunk = 0;
    // One pass of the loop body per host iteration.  The host knows k (the
    // position in the cycle) because restarts happen on a fixed schedule.
    int k = 0;
    for (int total = 0; total <= max_iters; ++total) {
        gmres_top_kernel<<<1, 1, 0, st>>>(S);
        count_launch();
        if (total == max_iters) break;   // the test above has set stop
        if (k == m_) {
            // restart: x += V y ; new residual ; new cycle (no-ops when stopped)
            update_x(1);
            // the cycle restart must not run once stopped: guard through S->stop
            launch_spmv(ctx_, A_, -1.0, x, 1.0, b, w_, EPI_NRM2, nullptr, &S->tmp, (int32_t)n_,
                        &S->stop);
            gmres_begin_cycle_kernel<<<1, 1, 0, st>>>(S, small_, 0, max_iters, tol, m_, nullptr);
            gmres_scale_kernel<<<g, kBlock, 0, st>>>(n_, w_, V_, &S->tmp, S, 1);
            count_launch(2);
            k = 0;
        }
        // Arnoldi step k: w = A V_k ; MGS against V_0..V_k
        if (M_) {
            M_->apply(V_ + (size_t)k * n_, pv_, nullptr, &S->stop);
            launch_spmv(ctx_, A_, 1.0, pv_, 0.0, nullptr, w_, EPI_NONE, nullptr, nullptr, 0,
                        &S->stop);
        } else {
            launch_spmv(ctx_, A_, 1.0, V_ + (size_t)k * n_, 0.0, nullptr, w_, EPI_NONE, nullptr,
                        nullptr, 0, &S->stop);
        }
        double *col = H + (size_t)k * (m_ + 1);
        // h_0 = w.v_0, then per projection one fused launch: w -= h_i v_i together with the dot
        // product of the next one (or the norm after the last)
        gmres_dot_kernel<<<g, kBlock, 0, st>>>(n_, w_, V_, ctx_.partials, ctx_.tickets + 4, col, 0, S);
        for (int i = 0; i <= k; ++i)
            gmres_axpy_dot_kernel<<<g, kBlock, 0, st>>>(
                n_, col + i, V_ + (size_t)i * n_, w_, i < k ? V_ + (size_t)(i + 1) * n_ : nullptr,
                ctx_.partials, ctx_.tickets + 4, i < k ? col + i + 1 : &S->hn, S);
        gmres_scale_kernel<<<g, kBlock, 0, st>>>(n_, w_, V_ + (size_t)(k + 1) * n_, &S->hn, S, 1);
        gmres_givens_kernel<<<1, 1, 0, st>>>(S, small_);
        count_launch(k + 4);
        ++k;
        if ((total & 15) == 15 && max_iters > 64) {
            const int slot = chunk & 1;
            SCHWZ_CUDA(cudaMemcpyAsync(pinned_stop_ + slot, &S->stop, sizeof(int32_t),
                                       cudaMemcpyDeviceToHost, st));
            SCHWZ_CUDA(cudaEventRecord(ev_[slot], st));
            if (chunk > 0) {
                SCHWZ_CUDA(cudaEventSynchronize(ev_[slot ^ 1]));
                if (pinned_stop_[slot ^ 1]) break;
            }
            ++chunk;
        }
    }
    SCHWZ_CUDA(cudaGetLastError());
    // final update with the columns of the last (partial) cycle.  When the
    // device stopped earlier than the host schedule, S->k is the true count
    // and the kernels after the stop were no-ops.
    update_x(0);
    SCHWZ_CUDA(cudaGetLastError());
}

void GmresSolver::result(int32_t *iters, double *resnorm, double *resnorm0)
{
    const size_t small = (size_t)(m_ + 1) * m_ + 4 * (size_t)m_ + 8;
    GmresState h;
    ctx_.use();
    SCHWZ_CUDA(cudaMemcpyAsync(&h, small_ + small, sizeof(h), cudaMemcpyDeviceToHost, ctx_.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (iters) *iters = h.total;
    if (resnorm) *resnorm = h.resnorm;
    if (resnorm0) *resnorm0 = h.r0;
}

// =============================================================================
// Level-scheduled sparse triangular solve (replaces gko::solver::LowerTrs /
// UpperTrs, i.e. cuSPARSE csrsm2 with the level policy on the reference's CUDA
// path).  Host analysis assigns level(i) = 1 + max level of the rows that row
// i depends on; a wide level is one launch (a warp per row, lanes striding over
// the row's entries, fixed-shape shuffle reduction; one thread per row for the
// huge levels of 1..8-entry rows at the leaves), all launches of a solve in one
// CUDA graph with programmatic dependent launch between them (see trs_row).  Runs of
// small levels - the dense separator triangles at the top of a nested-dissection
// factor, a chain of 1..8-row levels holding most of the non-zeros - are cut into
// blocks of <= 512 consecutive rows and solved as x_K = Dinv_K (b_K - L[K, outside] x)
// with the explicit inverse of the block's own triangle: two launches per block
// instead of one per level.  A second variant, the dependency-driven one-kernel
// solve, follows further down (trs_flow_kernel); measurements of both in
// profiles/r2_sptrsv.md.  The solves are bound by the length of their dependency
// chain times the latency of a hop, not by bandwidth.
// Algorithmic bytes: 12*nnz + 20*rows.
// =============================================================================
bool g_trs_pdl = true;   // SCHWZ_B200_TRS_NO_PDL=1: plain stream order between the level kernels
constexpr int kTrsWarps = kBlock / 32;
constexpr int kTrsSmallLevel = 8 * kTrsWarps;   // levels of at most this many rows go into blocks

// The factor is stored in LEVEL ORDER, off-diagonal entries only: position i of the level order
// is row order[i] with entries [prp[i], prp[i+1]) and reciprocal diagonal pinv[i].  That takes
// two dependent loads (level bounds, order[i] -> rp[row]) out of the critical path of every
// launch - these solves are bound by the memory latency of a short dependent chain per level,
// not by bandwidth.
// Every kernel of a solve is launched with programmatic stream serialisation (PDL): it may
// start while its predecessor is still running, does everything that does not depend on x
// (level-order metadata, indices, values, right-hand side - the dependent chain of HBM loads
// that made a level cost 7 us in round 1) and only then waits for the predecessor's results
// with cudaGridDependencySynchronize().  b is stable during a solve: the first kernel of a
// solve is launched without the attribute, i.e. behind a full dependency on whatever produced b.
__device__ __forceinline__ void trs_row(int32_t i, int32_t e, int32_t stride,
                                        const int32_t *__restrict__ order,
                                        const int32_t *__restrict__ prp,
                                        const int32_t *__restrict__ ci,
                                        const double *__restrict__ v,
                                        const double *__restrict__ pinv,
                                        const double *__restrict__ b, volatile double *x, int lane)
{
    // prologue: the first row of this warp, up to its first round of entries
    int32_t row = 0, k0 = 0, k1 = 0, c0[4] = {-1, -1, -1, -1};
    double rhs = 0.0, d = 0.0, v0[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < e) {
        row = order[i];
        k0 = prp[i];
        k1 = prp[i + 1];
        rhs = b[row];
        d = pinv[i];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int32_t kk = k0 + lane + 32 * u;
            if (kk < k1) {
                c0[u] = ci[kk];
                v0[u] = v[kk];
            }
        }
    }
    cudaGridDependencySynchronize();
    bool first = true;
    for (; i < e; i += stride) {
        if (!first) {
            row = order[i];
            k0 = prp[i];
            k1 = prp[i + 1];
            rhs = b[row];
            d = pinv[i];
        }
        // 4 independent (index, value, x) gathers per lane in flight
        double s = 0.0;
        for (int32_t k = k0 + lane; k < k1; k += 128) {
            double t[4] = {0.0, 0.0, 0.0, 0.0};
            if (first && k == k0 + lane) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (c0[u] >= 0) t[u] = v0[u] * x[c0[u]];
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int32_t kk = k + 32 * u;
                    if (kk < k1) t[u] = v[kk] * x[ci[kk]];
                }
            }
            s += (t[0] + t[1]) + (t[2] + t[3]);
        }
        s = warp_sum(s);
        if (lane == 0) x[row] = (rhs - s) * d;
        first = false;
    }
}

// one launch = positions [a, e) of the level order (one level), a warp per row
__global__ void __launch_bounds__(kBlock)
    trs_levels_kernel(int32_t a, int32_t e, const int32_t *__restrict__ order,
                      const int32_t *__restrict__ prp, const int32_t *__restrict__ ci,
                      const double *__restrict__ v, const double *__restrict__ pinv,
                      const double *__restrict__ b, double *x, const int32_t *stop)
{
    // let the next kernel of the solve start its own prologue right away: it still waits for
    // the whole of this grid in its cudaGridDependencySynchronize()
    cudaTriggerProgrammaticLaunchCompletion();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kBlock) >> 5;
    if (stop != nullptr && *stop != 0) {   // (the flag is set outside the solve)
        cudaGridDependencySynchronize();
        return;
    }
    trs_row(a + warp, e, nwarps, order, prp, ci, v, pinv, b, x, lane);
}

// A wide level of SHORT rows (the leaves of a nested-dissection factor: 10^5 rows of 0..8
// entries): one thread per row; a warp per row would idle 24+ of its lanes and walk a dozen
// rows one after the other.
__global__ void __launch_bounds__(kBlock)
    trs_level_thread_kernel(int32_t a, int32_t e, const int32_t *__restrict__ order,
                            const int32_t *__restrict__ prp, const int32_t *__restrict__ ci,
                            const double *__restrict__ v, const double *__restrict__ pinv,
                            const double *__restrict__ b, double *x, const int32_t *stop)
{
    cudaTriggerProgrammaticLaunchCompletion();
    const bool off = stop != nullptr && *stop != 0;
    int32_t i = a + blockIdx.x * kBlock + threadIdx.x;
    int32_t row = 0, k0 = 0, k1 = 0;
    double rhs = 0.0, d = 0.0;
    if (!off && i < e) {
        row = order[i];
        k0 = prp[i];
        k1 = prp[i + 1];
        rhs = b[row];
        d = pinv[i];
    }
    cudaGridDependencySynchronize();
    if (off) return;
    bool first = true;
    for (; i < e; i += gridDim.x * kBlock) {
        if (!first) {
            row = order[i];
            k0 = prp[i];
            k1 = prp[i + 1];
            rhs = b[row];
            d = pinv[i];
        }
        double s = 0.0;
        for (int32_t k = k0; k < k1; ++k) s += v[k] * x[ci[k]];
        x[row] = (rhs - s) * d;
        first = false;
    }
}

// One block of a small-level run: warps form t_i = b_i - sum over the entries OUTSIDE the
// block (all of them already solved by earlier launches); a second launch multiplies by the
// block's explicit inverse: x_K = Dinv_K t.
constexpr int kTrsBlock = 512;

// phase A: t_i = b_i - (entries of row i outside the block) . x
__global__ void __launch_bounds__(kBlock)
    trs_block_a_kernel(int32_t pos0, int32_t nrows, const int32_t *__restrict__ order,
                       const int32_t *__restrict__ crp, const int32_t *__restrict__ cci,
                       const double *__restrict__ cv, const double *__restrict__ b,
                       const double *__restrict__ x, double *__restrict__ t_scratch,
                       const int32_t *stop, int cta_per_row)
{
    cudaTriggerProgrammaticLaunchCompletion();
    const bool off = stop != nullptr && *stop != 0;
    const int lane = threadIdx.x & 31;
    if (cta_per_row) {
        // long rows (the dense separator triangles: hundreds to thousands of outside entries
        // per row): the whole CTA strides over one row, 4 independent gathers per thread in
        // flight; with a warp per row only a few warps would carry the block's megabytes
        __shared__ double s_red[kBlock / 32];
        int32_t k0 = 0, k1 = 0, c0 = -1;
        double rhs = 0.0, v0 = 0.0;
        if (!off && (int32_t)blockIdx.x < nrows) {
            const int32_t p = pos0 + blockIdx.x;
            k0 = crp[p];
            k1 = crp[p + 1];
            rhs = b[order[p]];
            if (k0 + (int32_t)threadIdx.x < k1) {
                c0 = cci[k0 + threadIdx.x];
                v0 = cv[k0 + threadIdx.x];
            }
        }
        cudaGridDependencySynchronize();
        if (off) return;
        bool first = true;
        for (int32_t i = blockIdx.x; i < nrows; i += gridDim.x) {
            if (!first) {
                const int32_t p = pos0 + i;
                k0 = crp[p];
                k1 = crp[p + 1];
                rhs = b[order[p]];
            }
            double s = 0.0;
            for (int32_t k = k0 + threadIdx.x; k < k1; k += 4 * kBlock) {
                double t0 = (first && k == k0 + (int32_t)threadIdx.x) ? v0 * x[c0] : cv[k] * x[cci[k]];
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                if (k + kBlock < k1) t1 = cv[k + kBlock] * x[cci[k + kBlock]];
                if (k + 2 * kBlock < k1) t2 = cv[k + 2 * kBlock] * x[cci[k + 2 * kBlock]];
                if (k + 3 * kBlock < k1) t3 = cv[k + 3 * kBlock] * x[cci[k + 3 * kBlock]];
                s += (t0 + t1) + (t2 + t3);
            }
            s = block_sum(s, s_red);
            if (threadIdx.x == 0) t_scratch[i] = rhs - s;
            first = false;
        }
        return;
    }
    const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * kBlock) >> 5;
    int32_t k0 = 0, k1 = 0, c0 = -1;
    double rhs = 0.0, v0 = 0.0;
    if (!off && warp < nrows) {
        const int32_t p = pos0 + warp;
        k0 = crp[p];
        k1 = crp[p + 1];
        rhs = b[order[p]];
        if (k0 + lane < k1) {
            c0 = cci[k0 + lane];
            v0 = cv[k0 + lane];
        }
    }
    cudaGridDependencySynchronize();
    if (off) return;
    bool first = true;
    for (int32_t i = warp; i < nrows; i += nwarps) {
        if (!first) {
            const int32_t p = pos0 + i;
            k0 = crp[p];
            k1 = crp[p + 1];
            rhs = b[order[p]];
        }
        double s = 0.0;
        for (int32_t k = k0 + lane; k < k1; k += 32)
            s += (first && k == k0 + lane) ? v0 * x[c0] : cv[k] * x[cci[k]];
        s = warp_sum(s);
        if (lane == 0) t_scratch[i] = rhs - s;
        first = false;
    }
}

// phase B, its own launch (the kernel boundary replaces a fence + ticket + reload in the last
// CTA): x_K = Dinv t.  Dinv row-major (lower triangle); a warp per row with the lanes across
// the columns (coalesced), four rows of a warp in flight at once.  Before it waits for phase A
// the kernel pulls its rows of Dinv towards L2.
__global__ void __launch_bounds__(kBlock)
    trs_block_b_kernel(int32_t pos0, int32_t nrows, const int32_t *__restrict__ order,
                       const double *__restrict__ dinv, const double *__restrict__ t_scratch,
                       double *__restrict__ x, const int32_t *stop)
{
    __shared__ double s_t[kTrsBlock];
    cudaTriggerProgrammaticLaunchCompletion();
    const bool off = stop != nullptr && *stop != 0;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const int r0 = 4 * w;
    if (!off && r0 < nrows) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u;
            if (r < nrows)
                for (int j = 16 * lane; j <= r; j += 16 * 32)   // one 128-byte line per lane
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(dinv + (size_t)r * nrows + j));
        }
    }
    cudaGridDependencySynchronize();
    if (off) return;
    for (int i = threadIdx.x; i < nrows; i += kBlock) s_t[i] = t_scratch[i];
    __syncthreads();
    if (r0 >= nrows) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int jj = 0; jj < kTrsBlock / 32; ++jj) {
        const int j = lane + 32 * jj;
        if (j > r0 + 3) break;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u;
            if (r < nrows && j <= r) acc[u] += dinv[(size_t)r * nrows + j] * s_t[j];
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const double a = warp_sum(acc[u]);
        if (lane == 0 && r0 + u < nrows) x[order[pos0 + r0 + u]] = a;
    }
}

TrsPlan::TrsPlan(const Ctx &ctx, int32_t n, const int32_t *rp, const int32_t *ci,
                 const double *v, bool upper)
    : ctx_(ctx), n_(n), upper_(upper)
{
    nnz_ = rp[n];
    std::vector<int32_t> level(n, 0);
    std::vector<double> inv_diag(n, 1.0);
    int32_t maxl = 0;
    auto visit = [&](int32_t i) {
        int32_t l = 0;
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
            const int32_t c = ci[k];
            if (c == i) inv_diag[i] = 1.0 / v[k];
            else if (upper ? c > i : c < i) l = std::max(l, level[c] + 1);
        }
        level[i] = l;
        maxl = std::max(maxl, l);
    };
    if (upper) for (int32_t i = n - 1; i >= 0; --i) visit(i);
    else for (int32_t i = 0; i < n; ++i) visit(i);
    num_levels_ = n > 0 ? maxl + 1 : 0;
    level_ptr_.assign((size_t)num_levels_ + 1, 0);
    for (int32_t i = 0; i < n; ++i) level_ptr_[level[i] + 1]++;
    for (int32_t l = 0; l < num_levels_; ++l) level_ptr_[l + 1] += level_ptr_[l];
    std::vector<int32_t> order(n), cur(level_ptr_.begin(), level_ptr_.end() - 1);
    for (int32_t i = 0; i < n; ++i) order[cur[level[i]]++] = i;
    {
        // level-ordered copy of the factor, off-diagonal entries only
        std::vector<int32_t> prp((size_t)n + 1, 0), pci;
        std::vector<double> pv, pinv((size_t)n);
        pci.reserve((size_t)nnz_);
        pv.reserve((size_t)nnz_);
        for (int32_t i = 0; i < n; ++i) {
            const int32_t row = order[i];
            for (int32_t k = rp[row]; k < rp[row + 1]; ++k)
                if (ci[k] != row && (upper ? ci[k] > row : ci[k] < row)) {
                    pci.push_back(ci[k]);
                    pv.push_back(v[k]);
                }
            prp[i + 1] = (int32_t)pci.size();
            pinv[i] = inv_diag[row];
        }
        rp_ = ctx.upload(prp.data(), prp.size());
        ci_ = ctx.upload(pci.data(), pci.size());
        v_ = ctx.upload(pv.data(), pv.size());
        inv_diag_ = ctx.upload(pinv.data(), pinv.size());
    }
    order_ = ctx.upload(order.data(), (size_t)n);

    // ---- segments: wide levels one by one, runs of small levels cut into blocks ----------
    static_assert(kTrsBlock % 32 == 0, "the inverse blocks are walked 32 columns at a time");
    std::vector<int32_t> crp(1, 0), cci, pos_in_block((size_t)n, -1);
    std::vector<double> cv, dinv;
    std::vector<int32_t> chain_first;   // position of every chain row -> index into crp
    std::vector<int32_t> crp_of_pos((size_t)n + 1, 0);
    int32_t l = 0;
    std::vector<double> D, Dinv;
    while (l < num_levels_) {
        const int32_t rows = level_ptr_[l + 1] - level_ptr_[l];
        if (rows > kTrsSmallLevel) {
            // mode 1: more rows than the grid has warps, and short ones -> one thread per row
            int64_t lnnz = 0;
            for (int32_t i = level_ptr_[l]; i < level_ptr_[l + 1]; ++i)
                lnnz += rp[order[i] + 1] - rp[order[i]];
            segments_.push_back({0, l, 0, 0, (rows > 8192 && lnnz <= 8 * (int64_t)rows) ? 1 : 0});
            ++l;
            continue;
        }
        int32_t l1 = l + 1;
        while (l1 < num_levels_ && level_ptr_[l1 + 1] - level_ptr_[l1] <= kTrsSmallLevel) ++l1;
        const int32_t p_begin = level_ptr_[l], p_end = level_ptr_[l1];
        for (int32_t p0 = p_begin; p0 < p_end; p0 += kTrsBlock) {
            const int32_t nb = std::min(kTrsBlock, p_end - p0);
            for (int32_t i = 0; i < nb; ++i) pos_in_block[order[p0 + i]] = i;
            D.assign((size_t)nb * nb, 0.0);
            for (int32_t i = 0; i < nb; ++i) {
                const int32_t row = order[p0 + i];
                for (int32_t k = rp[row]; k < rp[row + 1]; ++k) {
                    const int32_t c = ci[k];
                    const int32_t q = pos_in_block[c];
                    if (c == row) {
                        D[(size_t)i * nb + i] = v[k];
                    } else if (q >= 0) {
                        // in-block dependency: always an earlier position (earlier level)
                        D[(size_t)i * nb + q] = v[k];
                    } else if (upper ? c > row : c < row) {
                        cci.push_back(c);
                        cv.push_back(v[k]);
                    }
                }
                crp.push_back((int32_t)cci.size());
            }
            // explicit inverse of the lower-triangular block (row-major D, forward substitution
            // on the identity), kept row-major with ld = nb for the device
            Dinv.assign((size_t)nb * nb, 0.0);
            for (int32_t j = 0; j < nb; ++j) {
                for (int32_t i = j; i < nb; ++i) {
                    double sacc = (i == j) ? 1.0 : 0.0;
                    for (int32_t q = j; q < i; ++q) sacc -= D[(size_t)i * nb + q] * Dinv[(size_t)j * nb + q];
                    Dinv[(size_t)j * nb + i] = sacc / D[(size_t)i * nb + i];
                }
            }
            // mode 2: long rows -> a CTA per row for the outside part
            const int64_t outside = (int64_t)cci.size() - crp[crp.size() - 1 - nb];
            segments_.push_back({1, p0, nb, (int64_t)dinv.size(), outside >= 128 * (int64_t)nb ? 2 : 0});
            for (int32_t i = 0; i < nb; ++i)      // Dinv above is column-major: transpose
                for (int32_t j = 0; j < nb; ++j) dinv.push_back(Dinv[(size_t)j * nb + i]);
            for (int32_t i = 0; i < nb; ++i) pos_in_block[order[p0 + i]] = -1;
            ++num_blocks_;
        }
        l = l1;
    }
    if (num_blocks_ > 0) {
        // crp is indexed by chain-row ordinal; the kernel wants it indexed by position, so
        // spread it out (positions outside chain runs keep empty ranges)
        std::vector<int32_t> by_pos((size_t)n + 1, 0);
        size_t ord = 0;
        int32_t last = 0;
        std::vector<char> is_chain((size_t)n, 0);
        for (const Segment &sg : segments_)
            if (sg.kind == 1)
                for (int32_t i = 0; i < sg.b; ++i) is_chain[sg.a + i] = 1;
        for (int32_t p = 0; p < n; ++p) {
            by_pos[p] = last;
            if (is_chain[p]) {
                last = crp[ord + 1];
                ++ord;
            }
        }
        by_pos[n] = last;
        chain_rp_ = ctx.upload(by_pos.data(), by_pos.size());
        chain_ci_ = ctx.upload(cci.data(), cci.size());
        chain_v_ = ctx.upload(cv.data(), cv.size());
        dinv_ = ctx.upload(dinv.data(), dinv.size());
        block_t_ = ctx.alloc_zero<double>(kTrsBlock);
    }

    // ---- work items of the dependency-driven solve, in an order in which every item only
    // waits for items before it: the wide levels in level order (a chunk of <= 32 short rows
    // for the lanes of a warp, or one row per warp), then per block its right-hand-side rows
    // followed by the rows of the explicit inverse ------------------------------------------
    // Every item also names ONE entry to watch before it looks at anything else: the
    // dependency that sits latest in the item order (entries become ready roughly in that
    // order).  A single lane polls that entry; only then do the lanes gather, re-checking each
    // value - thousands of waiting warps otherwise poll with every lane and saturate L2.
    std::vector<int32_t> pos_of((size_t)n, 0);
    for (int32_t i = 0; i < n; ++i) pos_of[order[i]] = i;
    auto latest_dep = [&](int32_t pos) {   // row id of the latest dependency of position pos
        const int32_t row = order[pos];
        int32_t best = -1, best_pos = -1;
        for (int32_t k = rp[row]; k < rp[row + 1]; ++k) {
            const int32_t c = ci[k];
            if (c == row || !(upper ? c > row : c < row)) continue;
            if (pos_of[c] > best_pos) {
                best_pos = pos_of[c];
                best = c;
            }
        }
        return std::make_pair(best, best_pos);
    };
    std::vector<Item> items;
    std::vector<int32_t> bpos0, bnb, bord0;
    std::vector<int64_t> boff;
    int32_t ord = 0;
    for (const Segment &sg : segments_) {
        if (sg.kind == 0) {
            const int32_t p0 = level_ptr_[sg.a], p1 = level_ptr_[sg.a + 1];
            int64_t lnnz = 0;
            for (int32_t i = p0; i < p1; ++i) lnnz += rp[order[i] + 1] - rp[order[i]] - 1;
            if (lnnz <= 8 * (int64_t)(p1 - p0)) {
                for (int32_t q = p0; q < p1; q += 32) {
                    const int32_t cnt = std::min(32, p1 - q);
                    int32_t gate = -1, gate_pos = -1;
                    for (int32_t i = 0; i < cnt; ++i) {
                        auto d = latest_dep(q + i);
                        if (d.second > gate_pos) {
                            gate_pos = d.second;
                            gate = d.first;
                        }
                    }
                    items.push_back({0, q, cnt, gate});
                }
            } else {
                for (int32_t q = p0; q < p1; ++q) items.push_back({1, q, 0, latest_dep(q).first});
            }
        } else {
            const int32_t blk = (int32_t)bpos0.size();
            bpos0.push_back(sg.a);
            bnb.push_back(sg.b);
            bord0.push_back(ord);
            boff.push_back(sg.dinv_off);
            for (int32_t i = 0; i < sg.b; ++i) {
                // latest dependency OUTSIDE the block (inside ones are folded into Dinv)
                const int32_t row = order[sg.a + i];
                int32_t gate = -1, gate_pos = -1;
                for (int32_t k = rp[row]; k < rp[row + 1]; ++k) {
                    const int32_t c = ci[k];
                    if (c == row || !(upper ? c > row : c < row)) continue;
                    const int32_t pc = pos_of[c];
                    if (pc >= sg.a && pc < sg.a + sg.b) continue;
                    if (pc > gate_pos) {
                        gate_pos = pc;
                        gate = c;
                    }
                }
                items.push_back({2, sg.a + i, ord + i, gate});
            }
            for (int32_t i = 0; i < sg.b; ++i) items.push_back({3, blk, i, 0});
            ord += sg.b;
        }
    }
    num_items_ = (int32_t)items.size();
    num_chain_rows_ = ord;
    items_ = ctx.upload(items.data(), items.size());
    counter_ = ctx.alloc_zero<int32_t>(4);
    t_ = ctx.alloc_zero<double>(std::max(ord, 1));
    blk_pos0_ = ctx.upload(bpos0.data(), bpos0.size());
    blk_nb_ = ctx.upload(bnb.data(), bnb.size());
    blk_ord0_ = ctx.upload(bord0.data(), bord0.size());
    blk_dinv_off_ = ctx.upload(boff.data(), boff.size());
}

TrsPlan::~TrsPlan()
{
    for (const Captured &c : graphs_) cudaGraphExecDestroy(c.exec);
    ctx_.release(rp_);
    ctx_.release(ci_);
    ctx_.release(v_);
    ctx_.release(order_);
    ctx_.release(inv_diag_);
    ctx_.release(chain_rp_);
    ctx_.release(chain_ci_);
    ctx_.release(chain_v_);
    ctx_.release(dinv_);
    ctx_.release(block_t_);
    ctx_.release(items_);
    ctx_.release(counter_);
    ctx_.release(t_);
    ctx_.release(blk_pos0_);
    ctx_.release(blk_nb_);
    ctx_.release(blk_ord0_);
    ctx_.release(blk_dinv_off_);
}

void TrsPlan::solve_levels(const double *b, double *x, const int32_t *stop)
{
    ctx_.use();
    if (n_ == 0) return;
    for (const Captured &c : graphs_)
        if (c.b == b && c.x == x && c.stop == stop) {
            SCHWZ_CUDA(cudaGraphLaunch(c.exec, ctx_.stream));
            count_launch(num_launches());
            return;
        }
    // capture the level launches once per (b, x, stop) triple; a few triples are kept (a
    // preconditioner is applied to different vectors by the same solver)
    if (graphs_.size() >= 4) {
        cudaGraphExecDestroy(graphs_.front().exec);
        graphs_.erase(graphs_.begin());
    }
    cudaGraph_t graph = nullptr;
    SCHWZ_CUDA(cudaStreamBeginCapture(ctx_.stream, cudaStreamCaptureModeThreadLocal));
    // every launch but the first may overlap its predecessor's tail (see trs_row)
    bool first_launch = true;
    auto launch = [&](auto kernel, int grid, auto... args) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kBlock);
        cfg.stream = ctx_.stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (first_launch || !g_trs_pdl) ? 0 : 1;
        first_launch = false;
        SCHWZ_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
    };
    for (const Segment &sg : segments_) {
        if (sg.kind == 0) {
            const int32_t l = sg.a;
            const int32_t rows = level_ptr_[l + 1] - level_ptr_[l];
            if (sg.mode == 1) {
                const int grid = std::min((rows + kBlock - 1) / kBlock, ctx_.vec_grid());
                launch(trs_level_thread_kernel, grid, level_ptr_[l], level_ptr_[l + 1],
                       (const int32_t *)order_, (const int32_t *)rp_, (const int32_t *)ci_,
                       (const double *)v_, (const double *)inv_diag_, b, x, stop);
                continue;
            }
            const int grid = std::min((rows + kTrsWarps - 1) / kTrsWarps, ctx_.vec_grid());
            launch(trs_levels_kernel, grid, level_ptr_[l], level_ptr_[l + 1],
                   (const int32_t *)order_, (const int32_t *)rp_, (const int32_t *)ci_,
                   (const double *)v_, (const double *)inv_diag_, b, x, stop);
        } else {
            const int grid = sg.mode == 2 ? sg.b : std::max(1, (sg.b + kTrsWarps - 1) / kTrsWarps);
            launch(trs_block_a_kernel, grid, sg.a, sg.b, (const int32_t *)order_,
                   (const int32_t *)chain_rp_, (const int32_t *)chain_ci_, (const double *)chain_v_,
                   b, (const double *)x, block_t_, stop, sg.mode == 2 ? 1 : 0);
            launch(trs_block_b_kernel, (sg.b + 4 * kTrsWarps - 1) / (4 * kTrsWarps), sg.a, sg.b,
                   (const int32_t *)order_, (const double *)(dinv_ + sg.dinv_off),
                   (const double *)block_t_, x, stop);
        }
    }
    SCHWZ_CUDA(cudaStreamEndCapture(ctx_.stream, &graph));
    Captured c{b, x, stop, nullptr};
    SCHWZ_CUDA(cudaGraphInstantiate(&c.exec, graph, 0));
    cudaGraphDestroy(graph);
    graphs_.push_back(c);
    SCHWZ_CUDA(cudaGraphLaunch(c.exec, ctx_.stream));
    count_launch(num_launches());
}

// =============================================================================
// Dependency-driven solve: ONE persistent kernel per triangular solve instead of one launch
// per level (331 dependent launches for the cfg5 factor).  The work items above are claimed in
// order from a counter; a warp that holds an item first loads everything that does not depend
// on the solution (row bounds, indices, values, right-hand side, reciprocal diagonal) and only
// then looks at the x entries it needs.  x (and the blocks' right-hand sides t) are pre-filled
// with an "unset" bit pattern (all ones, a NaN no computation produces), so an entry IS its own
// ready flag: a consumer re-reads it from L2 until it is set - one L2 round trip per dependency
// level (~0.15 us) instead of a kernel boundary plus a chain of dependent HBM loads (~7.5 us).
// Items are claimed in dependency order and only by running warps, so every wait is for an
// item that a running warp already holds: no deadlock, whatever part of the grid is resident
// (several subdomains' solves share the GPU).  Waits are bounded all the same: on expiry the
// abort word is raised, every wait falls through, and the host reports the failure.
// Summation order per row is the one of the level kernels above (same results bit for bit).
// =============================================================================
constexpr unsigned long long kTrsUnset = 0xffffffffffffffffull;

__device__ __forceinline__ double ld_l2(const double *p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_l2(double *p, double v)
{
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool trs_unset(double v)
{
    return (unsigned long long)__double_as_longlong(v) == kTrsUnset;
}
// the value at p once it has been produced (0 after an abort)
__device__ __forceinline__ double trs_wait(const double *p, volatile int32_t *abort_word)
{
    double v = ld_l2(p);
    unsigned int spins = 0;
    while (trs_unset(v)) {
        if ((++spins & 4095u) == 0) {
            if (*abort_word != 0) return 0.0;
            if (spins > (1u << 26)) {   // seconds: something is wrong - stop everybody
                atomicExch((int32_t *)abort_word, 1);
                return 0.0;
            }
        }
        v = ld_l2(p);
    }
    return v;
}
// one lane watches the item's latest dependency, the warp goes on when it is there
__device__ __forceinline__ void trs_gate(const double *p, int lane, volatile int32_t *abort_word)
{
    if (lane == 0) (void)trs_wait(p, abort_word);
    __syncwarp();
}

__global__ void __launch_bounds__(kBlock)
    trs_prepare_kernel(int32_t n, double *__restrict__ x, int32_t nt, double *__restrict__ t,
                       int32_t *counter, const int32_t *stop)
{
    if (stop != nullptr && *stop != 0) return;
    const double unset = __longlong_as_double((long long)kTrsUnset);
    const int64_t stride = (int64_t)gridDim.x * kBlock;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) x[i] = unset;
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < nt; i += stride) t[i] = unset;
    if (blockIdx.x == 0 && threadIdx.x == 0) counter[0] = 0;
}

struct TrsFlowArgs {
    int32_t num_items;
    const int4 *items;
    int32_t *counter;
    const int32_t *order, *prp, *ci;
    const double *v, *pinv;
    const int32_t *crp, *cci;
    const double *cv, *dinv;
    const int32_t *blk_pos0, *blk_nb, *blk_ord0;
    const int64_t *blk_off;
    double *t;
};

// sum over entries [k0, k1) of val[k] * X[col[k]] with the lanes of a warp striding over the
// entries, 4 gathers per lane in flight - the association of trs_row above.  The indices and
// values of the next round are loaded before the x entries of this one are looked at, and a
// lane only ever waits for an entry that is really still missing: whatever part of a long row
// refers to rows solved long ago is summed up while the latest dependencies are still in work.
__device__ __forceinline__ double trs_gather_row(int32_t k0, int32_t k1,
                                                 const int32_t *__restrict__ col,
                                                 const double *__restrict__ val, const double *X,
                                                 int lane, volatile int32_t *abort_word)
{
    double s = 0.0;
    int32_t c[4], cn[4];
    double vv[4], vn[4], xv[4];
    auto load_round = [&](int32_t k, int32_t *cc, double *vc) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int32_t kk = k + 32 * u;
            cc[u] = kk < k1 ? __ldg(col + kk) : -1;
            vc[u] = kk < k1 ? __ldg(val + kk) : 0.0;
        }
    };
    int32_t k = k0 + lane;
    if (k < k1) load_round(k, c, vv);
    for (; k < k1; k += 128) {
        const bool more = k + 128 < k1;
        if (more) load_round(k + 128, cn, vn);
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = c[u] >= 0 ? ld_l2(X + c[u]) : 0.0;
        double t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c[u] >= 0 && trs_unset(xv[u])) xv[u] = trs_wait(X + c[u], abort_word);
            t[u] = c[u] >= 0 ? vv[u] * xv[u] : 0.0;
        }
        s += (t[0] + t[1]) + (t[2] + t[3]);
        if (more) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                c[u] = cn[u];
                vv[u] = vn[u];
            }
        }
    }
    return warp_sum(s);
}

__global__ void __launch_bounds__(kBlock)
    trs_flow_kernel(TrsFlowArgs A, const double *__restrict__ b, double *x, const int32_t *stop)
{
    if (stop != nullptr && *stop != 0) return;
    const int lane = threadIdx.x & 31;
    volatile int32_t *abort_word = A.counter + 1;
    // a warp always holds its current item and has the next one claimed: the claim and the
    // item's descriptor travel while the current item is being worked on
    int32_t idx = 0;
    if (lane == 0) idx = atomicAdd(A.counter, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    while (idx < A.num_items) {
        int32_t next = 0;
        if (lane == 0) next = atomicAdd(A.counter, 1);
        const int4 it = __ldg(A.items + idx);
        if (it.x == 0) {
            // <= 32 short rows of one level, a lane each; everything that does not depend on
            // x is fetched before the warp looks at its latest dependency
            constexpr int kPre = 4;
            int32_t row = 0, k0 = 0, k1 = 0, pc[kPre];
            double rhs = 0.0, d = 0.0, pv[kPre];
            if (lane < it.z) {
                const int32_t pos = it.y + lane;
                row = A.order[pos];
                k0 = A.prp[pos];
                k1 = A.prp[pos + 1];
                rhs = b[row];
                d = A.pinv[pos];
#pragma unroll
                for (int u = 0; u < kPre; ++u) {
                    pc[u] = k0 + u < k1 ? __ldg(A.ci + k0 + u) : -1;
                    pv[u] = k0 + u < k1 ? __ldg(A.v + k0 + u) : 0.0;
                }
            }
            if (it.w >= 0) trs_gate(x + it.w, lane, abort_word);
            if (lane < it.z) {
                double s = 0.0;
#pragma unroll
                for (int u = 0; u < kPre; ++u)
                    if (pc[u] >= 0) s += pv[u] * trs_wait(x + pc[u], abort_word);
                for (int32_t k = k0 + kPre; k < k1; ++k) {
                    const int32_t c = A.ci[k];
                    const double vv = A.v[k];
                    s += vv * trs_wait(x + c, abort_word);
                }
                st_l2(x + row, (rhs - s) * d);
            }
        } else if (it.x == 1) {
            const int32_t pos = it.y;
            const int32_t row = A.order[pos];
            const int32_t k0 = A.prp[pos], k1 = A.prp[pos + 1];
            const double rhs = b[row], d = A.pinv[pos];
            const double s = trs_gather_row(k0, k1, A.ci, A.v, x, lane, abort_word);
            if (lane == 0) st_l2(x + row, (rhs - s) * d);
        } else if (it.x == 2) {
            // right-hand side of a block row: t = b - (entries outside the block) . x
            const int32_t pos = it.y;
            const int32_t k0 = A.crp[pos], k1 = A.crp[pos + 1];
            const double rhs = b[A.order[pos]];
            const double s = trs_gather_row(k0, k1, A.cci, A.cv, x, lane, abort_word);
            if (lane == 0) st_l2(A.t + it.z, rhs - s);
        } else {
            // row r of x_K = Dinv_K t_K (Dinv row-major, lower triangle)
            const int32_t blk = it.y, r = it.z;
            const int32_t nb = A.blk_nb[blk];
            const double *drow = A.dinv + A.blk_off[blk] + (size_t)r * nb;
            const double *tk = A.t + A.blk_ord0[blk];
            const int32_t row = A.order[A.blk_pos0[blk] + r];
            double dv[kTrsBlock / 32];
#pragma unroll
            for (int jj = 0; jj < kTrsBlock / 32; ++jj) {
                const int j = lane + 32 * jj;
                dv[jj] = j <= r ? __ldg(drow + j) : 0.0;
            }
            double acc = 0.0;
#pragma unroll
            for (int jj = 0; jj < kTrsBlock / 32; ++jj) {
                const int j = lane + 32 * jj;
                if (j <= r) acc += dv[jj] * trs_wait(tk + j, abort_word);
            }
            acc = warp_sum(acc);
            if (lane == 0) st_l2(x + row, acc);
        }
        idx = __shfl_sync(0xffffffffu, next, 0);
    }
}

void TrsPlan::solve_flow(const double *b, double *x, const int32_t *stop)
{
    ctx_.use();
    if (n_ == 0) return;
    SCHWZ_REQUIRE(b != x, "triangular solve: right-hand side and solution must not alias");
    static const int ctas_per_sm = [] {
        const char *e = std::getenv("SCHWZ_B200_TRS_CTAS_PER_SM");
        const int v = e ? std::atoi(e) : 0;
        return v > 0 ? v : 1;
    }();
    const int fill_grid = (int)std::max<int64_t>(
        1, std::min<int64_t>(((int64_t)n_ + kBlock - 1) / kBlock, ctx_.vec_grid()));
    trs_prepare_kernel<<<fill_grid, kBlock, 0, ctx_.stream>>>(n_, x, num_chain_rows_, t_, counter_,
                                                             stop);
    TrsFlowArgs A;
    A.num_items = num_items_;
    A.items = reinterpret_cast<const int4 *>(items_);
    A.counter = counter_;
    A.order = order_;
    A.prp = rp_;
    A.ci = ci_;
    A.v = v_;
    A.pinv = inv_diag_;
    A.crp = chain_rp_;
    A.cci = chain_ci_;
    A.cv = chain_v_;
    A.dinv = dinv_;
    A.blk_pos0 = blk_pos0_;
    A.blk_nb = blk_nb_;
    A.blk_ord0 = blk_ord0_;
    A.blk_off = blk_dinv_off_;
    A.t = t_;
    const int grid = std::max(1, std::min((num_items_ + kTrsWarps - 1) / kTrsWarps,
                                          ctx_.num_sms * ctas_per_sm));
    trs_flow_kernel<<<grid, kBlock, 0, ctx_.stream>>>(A, b, x, stop);
    SCHWZ_CUDA(cudaGetLastError());
    count_launch(2);
}

int32_t TrsPlan::error()
{
    int32_t e = 0;
    ctx_.use();
    SCHWZ_CUDA(cudaMemcpyAsync(&e, counter_ + 1, sizeof(e), cudaMemcpyDeviceToHost, ctx_.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx_.stream));
    return e;
}

// The level-per-launch graph (with programmatic dependent launch) is the default; the one-kernel
// solve stays selectable with SCHWZ_B200_TRS_LEVELS=0 (measurements: profiles/r2_sptrsv.md).
bool TrsPlan::uses_level_graph() const
{
    const char *e = std::getenv("SCHWZ_B200_TRS_LEVELS");
    return e ? e[0] != '0' : true;
}

void TrsPlan::solve(const double *b, double *x, const int32_t *stop)
{
    // Measured (profiles/r2_sptrsv.md), L + U pair of the cfg5 factor (1 462 levels) alone / 8
    // side by side, and of the ILU(0) wavefronts of a cfg2 strip (9 216 levels) alone / 2 side
    // by side:
    //   round 1 level graph (128-row blocks)          2489 / 372 us      51.0 / 30.0 ms
    //   one-kernel solve (trs_flow_kernel)             1811 / 740 us      30.9 / 16.6 ms
    //   level graph, 512-row blocks                    1850 / 325 us
    //   level graph, 512-row blocks + PDL prologues    1267 / 256 us      30.6 / 15.7 ms
    // The one-kernel solve removes the launch chain but pays ~1.7 us per dependency hop (several
    // L2 round trips) and its polling warps contend when solves share the GPU; the launch graph
    // with dependency-free prologues is as fast on deep chains and faster everywhere else.
    if (uses_level_graph()) solve_levels(b, x, stop);
    else solve_flow(b, x, stop);
}

}  // namespace schwz_b200
