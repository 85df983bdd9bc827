// Per-subdomain state of the RAS iteration, the stages of the outer loop and
// the loop itself (replaces SchwarzBase::run's loop body,
// source/schwarz_base.cpp:387-452, and everything it calls).
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "engine.hpp"

namespace schwz_b200 {

// =============================================================================
// NCCL through dlopen: the only collective on the path is the allgather of P
// residual norms (MPI_Allgather, source/solve.cpp:890-891).  Loading lazily
// keeps the library loadable where NCCL is absent and binds to whichever
// libnccl.so.2 the process already carries (torch's, under torchrun).
// =============================================================================
namespace {
struct Id128 {   // ncclUniqueId is passed by value: 128 bytes
    char b[128];
};
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi &nccl()
{
    static NcclApi api;
    if (!api.h) {
        api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        SCHWZ_REQUIRE(api.h != nullptr, "cannot load libnccl.so.2");
        api.GetUniqueId = (int (*)(void *))dlsym(api.h, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void **, int, Id128, int))dlsym(api.h, "ncclCommInitRank");
        api.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(
            api.h, "ncclAllGather");
        api.CommDestroy = (int (*)(void *))dlsym(api.h, "ncclCommDestroy");
        api.GetErrorString = (const char *(*)(int))dlsym(api.h, "ncclGetErrorString");
        SCHWZ_REQUIRE(api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy,
                      "libnccl lacks the expected symbols");
    }
    return api;
}
void nccl_check(int rc, const char *what)
{
    if (rc != 0) {
        const char *s = nccl().GetErrorString ? nccl().GetErrorString(rc) : "?";
        throw std::runtime_error(std::string("NCCL ") + what + ": " + s);
    }
}
}  // namespace

void comm_unique_id(void *id128) { nccl_check(nccl().GetUniqueId(id128), "ncclGetUniqueId"); }

Comm *comm_create(const Ctx &ctx, const void *id128, int nranks, int rank)
{
    ctx.use();
    auto *c = new Comm();
    c->ctx = &ctx;
    c->nranks = nranks;
    c->rank = rank;
    Id128 id;
    std::memcpy(&id, id128, sizeof(id));
    nccl_check(nccl().CommInitRank(&c->nccl, nranks, id, rank), "ncclCommInitRank");
    return c;
}

Comm::~Comm()
{
    if (nccl) ::schwz_b200::nccl().CommDestroy(nccl);
    if (ctx) {
        ctx->release(dev_in);
        ctx->release(dev_out);
    }
}

void comm_allgather_f64(Comm &c, const double *dev_in, int count, double *dev_out)
{
    c.ctx->use();
    // ncclDouble == 8
    nccl_check(nccl().AllGather(dev_in, dev_out, (size_t)count, 8, c.nccl, c.ctx->stream),
               "ncclAllGather");
}

// =============================================================================
// Mailbox
// =============================================================================
static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

MailboxLayout MailboxLayout::make(int64_t in_total, int32_t n_in, int32_t P, int64_t out_total,
                                  int64_t x_len)
{
    MailboxLayout m;
    m.recv_stride = align_up(std::max<int64_t>(in_total, 1) * 8, 256);
    m.flags_off = 2 * m.recv_stride;
    m.conv_off = m.flags_off + align_up(std::max(n_in, 1) * 8, 256);
    m.err_off = m.conv_off + align_up((int64_t)std::max(P, 3) * 4, 256);
    m.send_off = m.err_off + 256;
    m.slots_off = m.send_off + align_up(std::max<int64_t>(out_total, 1) * 8, 256);
    m.x_off = m.slots_off + align_up(std::max<int64_t>(in_total, 1) * 4, 256);
    m.bytes = m.x_off + align_up(std::max<int64_t>(x_len, 1) * 8, 256);
    return m;
}

// =============================================================================
// Ras
// =============================================================================
Ras::Ras(const Ctx &ctx_, Setup &setup, int32_t rank_, const double *host_rhs_global,
         const RasOptions &opt_)
    : ctx(ctx_), rank(rank_), P(setup.P()), opt(opt_)
{
    setup.build_matrices(rank);
    RankLayout &R = setup.rank(rank);
    local_size = R.local_size;
    local_size_x = R.local_size_x;
    overlap_size = R.overlap_size;
    n_halo = R.n_halo;
    first_row = setup.first_row()[rank];
    nbr_in = R.nbr_in;
    nbr_out = R.nbr_out;
    local_nnz = R.local.nnz();
    ctx.use();

    A.reset(csr_upload(ctx, R.local.nrows, R.local.ncols, R.local.rp.data(), R.local.ci.data(),
                       R.local.v.data()));
    HostCsr Ic;
    setup.compact_interface(rank, Ic);
    I.reset(csr_upload(ctx, Ic.nrows, Ic.ncols, Ic.rp.data(), Ic.ci.data(), Ic.v.data()));

    l2g_local_.assign(R.l2g.begin(), R.l2g.begin() + local_size_x);
    // A6: local_rhs = [rhs[own] ; rhs[overlap_row]]  (initialization.cpp:333-359)
    std::vector<double> lrhs((size_t)local_size_x, 1.0);
    if (host_rhs_global)
        for (int32_t k = 0; k < local_size_x; ++k) lrhs[k] = host_rhs_global[R.l2g[k]];
    local_rhs = ctx.upload(lrhs.data(), lrhs.size());
    // F8: the reference leaves these uninitialised and relies on zeros
    // (x itself lives in the peer-visible mailbox, allocated below)
    local_sol = ctx.alloc_zero<double>(local_size_x);
    init_guess = ctx.alloc_zero<double>(local_size_x);
    work = ctx.alloc_zero<double>(2 * (size_t)local_size_x);
    resnorm_dev = ctx.alloc_zero<double>(2);
    state = (OuterState *)ctx.alloc_zero<char>(sizeof(OuterState));
    {
        OuterState init{};
        init.resnorm = init.resnorm0 = init.gres0 = -1.0;
        init.finished_iter = -1;
        SCHWZ_CUDA(cudaMemcpyAsync(state, &init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    SCHWZ_CUDA(cudaEventCreateWithFlags(&ev_resid, cudaEventDisableTiming));
    num_converged_dev = ctx.alloc_zero<int32_t>(2);
    conv_sent = ctx.alloc_zero<int32_t>(std::max(P, 3));

    // halo tables.  in: scatter targets in compact numbering (g2l - 1);
    // out: gather sources inside my own block (global id - first_row).
    std::vector<int32_t> pos_of;   // global id -> compact slot for the non-own part
    {
        // ids beyond local_size appear once in l2g; binary search on a sorted copy
        std::vector<std::pair<int32_t, int32_t>> ext;
        ext.reserve(R.l2g.size() - local_size);
        for (int32_t k = local_size; k < (int32_t)R.l2g.size(); ++k) ext.emplace_back(R.l2g[k], k);
        std::sort(ext.begin(), ext.end());
        std::vector<int32_t> dst;
        for (size_t j = 0; j < R.get.size(); ++j) {
            in_count.push_back((int32_t)R.get[j].size());
            for (int32_t gid : R.get[j]) {
                auto it = std::lower_bound(ext.begin(), ext.end(), std::make_pair(gid, (int32_t)-1));
                SCHWZ_REQUIRE(it != ext.end() && it->first == gid, "halo id not in the index set");
                dst.push_back(it->second);
            }
        }
        in_total_ = (int32_t)dst.size();
        in_dst_ = ctx.upload(dst.data(), dst.size());
        // Get one-by-one: where each of those elements sits inside its owner's x (the own
        // block of a subdomain is stored first, in global order)
        std::vector<int32_t> rsrc;
        in_off_host_.assign(1, 0);
        for (size_t j = 0; j < R.get.size(); ++j) {
            const int32_t owner_first = setup.first_row()[nbr_in[j]];
            for (int32_t gid : R.get[j]) rsrc.push_back(gid - owner_first);
            in_off_host_.push_back((int32_t)rsrc.size());
        }
        in_remote_idx_ = ctx.upload(rsrc.data(), rsrc.size());
        in_off_ = ctx.upload(in_off_host_.data(), in_off_host_.size());
    }
    {
        std::vector<int32_t> src, off(1, 0);
        for (size_t j = 0; j < R.put.size(); ++j) {
            out_count.push_back((int32_t)R.put[j].size());
            for (int32_t gid : R.put[j]) src.push_back(gid - first_row);
            off.push_back((int32_t)src.size());
        }
        out_total_ = (int32_t)src.size();
        out_src_ = ctx.upload(src.data(), src.size());
        out_off_ = ctx.upload(off.data(), off.size());
        out_off_host_ = off;
    }
    mbox = MailboxLayout::make(in_total_, (int32_t)nbr_in.size(), P, out_total_,
                               (int64_t)local_size_x + n_halo);
    mailbox = ctx.alloc_zero<char>((size_t)mbox.bytes);
    x = (double *)(mailbox + mbox.x_off);
    if (in_total_ > 0)   // the slot table is what a one-by-one Put peer needs to know
        SCHWZ_CUDA(cudaMemcpyAsync(mailbox + mbox.slots_off, in_dst_, sizeof(int32_t) * (size_t)in_total_,
                                   cudaMemcpyDeviceToDevice, ctx.stream));
    const size_t no = nbr_out.size();
    const size_t ni = nbr_in.size();
    out_remote_slot_ = ctx.alloc_zero<int32_t>(std::max<size_t>(out_total_, 1));
    in_send_host_.assign(ni, nullptr);
    in_x_host_.assign(ni, nullptr);
    out_x_host_.assign(no, nullptr);
    send_seg_host_.assign(no, nullptr);
    for (size_t j = 0; j < no; ++j)
        send_seg_host_[j] = mailbox + mbox.send_off + wire_size() * (size_t)out_off_host_[j];
    in_send_dev_ = ctx.alloc_zero<const void *>(std::max<size_t>(ni, 1));
    in_x_dev_ = ctx.alloc_zero<const void *>(std::max<size_t>(ni, 1));
    out_x_dev_ = ctx.alloc_zero<double *>(std::max<size_t>(no, 1));
    send_seg_dev_ = ctx.alloc_zero<void *>(std::max<size_t>(no, 1));
    conv_peer_host_.assign((size_t)P, nullptr);
    conv_peer_host_[rank] = conv();
    conv_peer_dev_ = ctx.alloc_zero<int32_t *>((size_t)P);
    for (int b = 0; b < 2; ++b) out_dst_host_[b].assign(no, nullptr);
    out_flag_host_.assign(no, nullptr);
    out_conv_host_.assign(no, nullptr);
    out_same_process_.assign(no, 1);
    for (int b = 0; b < 2; ++b) out_dst_dev_[b] = ctx.alloc_zero<void *>(std::max<size_t>(no, 1));
    out_flag_dev_ = ctx.alloc_zero<unsigned long long *>(std::max<size_t>(no, 1));
    out_conv_dev_ = ctx.alloc_zero<int32_t *>(std::max<size_t>(no, 1));
    SCHWZ_CUDA(cudaEventCreateWithFlags(&ev_pushed, cudaEventDisableTiming));

    if (opt.local_solver == 2) {
        if (opt.non_symmetric) gmres.reset(new GmresSolver(ctx, *A, opt.restart_iter));
        else cg.reset(new CgSolver(ctx, *A));
        if (opt.local_precond != PRECOND_NONE) {
            // solve.cpp:486-652: generated once from the local matrix, before the loop
            precond.reset(new Preconditioner(ctx, R.local.nrows, R.local.rp.data(),
                                             R.local.ci.data(), R.local.v.data(),
                                             opt.local_precond, opt.precond_max_block_size));
            precond->release_host();
            if (gmres) gmres->set_precond(precond.get());
            else cg->set_precond(precond.get());
        }
    }
    ctx.sync();
}

Ras::~Ras()
{
    cudaSetDevice(ctx.device);
    cudaStreamSynchronize(ctx.stream);
    for (void *p : {(void *)in_off_, (void *)in_remote_idx_, (void *)out_remote_slot_,
                    (void *)in_send_dev_, (void *)in_x_dev_, (void *)out_x_dev_,
                    (void *)send_seg_dev_, (void *)conv_peer_dev_})
        ctx.release(p);
    hub.reset();
    if (ev_resid) cudaEventDestroy(ev_resid);
    for (void *p : {(void *)local_rhs, (void *)local_sol, (void *)init_guess, (void *)state,
                    (void *)work, (void *)resnorm_dev, (void *)num_converged_dev,
                    (void *)conv_sent, (void *)mailbox, (void *)in_dst_, (void *)out_src_,
                    (void *)out_off_, (void *)out_dst_dev_[0], (void *)out_dst_dev_[1],
                    (void *)out_flag_dev_, (void *)out_conv_dev_, (void *)fperm,
                    (void *)fperm_col})
        ctx.release(p);
    if (ev_pushed) cudaEventDestroy(ev_pushed);
    if (pinned_rhs_) cudaFreeHost(pinned_rhs_);
}

void Ras::upload_rhs(const double *host_rhs_global)
{
    ctx.use();
    // Only the overlap rows need a gather (small pinned staging area); the own
    // block is contiguous in the global numbering and is copied straight from
    // the caller's buffer (truly asynchronous when that buffer is pinned).
    if (overlap_size > 0) {
        if (!pinned_rhs_)
            SCHWZ_CUDA(cudaMallocHost((void **)&pinned_rhs_, sizeof(double) * (size_t)overlap_size));
        else
            SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));   // staging area still in use?
        for (int32_t k = 0; k < overlap_size; ++k)
            pinned_rhs_[k] = host_rhs_global[l2g_local_[local_size + k]];
        SCHWZ_CUDA(cudaMemcpyAsync(local_rhs + local_size, pinned_rhs_,
                                   sizeof(double) * (size_t)overlap_size, cudaMemcpyHostToDevice,
                                   ctx.stream));
    }
    SCHWZ_CUDA(cudaMemcpyAsync(local_rhs, host_rhs_global + first_row,
                               sizeof(double) * (size_t)local_size, cudaMemcpyHostToDevice,
                               ctx.stream));
}

void Ras::download_solution(double *host_solution_global)
{
    ctx.use();
    SCHWZ_CUDA(cudaMemcpyAsync(host_solution_global + first_row, x, sizeof(double) * (size_t)local_size,
                               cudaMemcpyDeviceToHost, ctx.stream));
}

void Ras::reset_state()
{
    ctx.use();
    SCHWZ_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * ((size_t)local_size_x + n_halo), ctx.stream));
    SCHWZ_CUDA(cudaMemsetAsync(init_guess, 0, sizeof(double) * (size_t)local_size_x, ctx.stream));
    SCHWZ_CUDA(cudaMemsetAsync(local_sol, 0, sizeof(double) * (size_t)local_size_x, ctx.stream));
    SCHWZ_CUDA(cudaMemsetAsync(conv(), 0, sizeof(int32_t) * (size_t)std::max(P, 3), ctx.stream));
    SCHWZ_CUDA(cudaMemsetAsync(err_word(), 0, sizeof(int32_t), ctx.stream));
    SCHWZ_CUDA(cudaMemsetAsync(conv_sent, 0, sizeof(int32_t) * (size_t)std::max(P, 3), ctx.stream));
    OuterState init{};
    init.resnorm = init.resnorm0 = init.gres0 = -1.0;
    init.finished_iter = -1;
    SCHWZ_CUDA(cudaMemcpyAsync(state, &init, sizeof(init), cudaMemcpyHostToDevice, ctx.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));   // `init` lives on this stack frame
    resnorm = resnorm0 = -1.0;
    gres = 0.0;
    gres0 = -1.0;
    num_converged = 0;
    finished = false;
    finished_iter = -1;
    // The halo values of the previous run still sit in the receive / send buffers.  A
    // synchronous run overwrites them before it reads them; a one-sided run would scatter them
    // at its first exchange if the peers' fresh values have not landed yet, so it clears them
    // on entry (not here: a peer process that has already started its next run may be writing
    // into them by now - reset is not a collective).  The epoch counters keep counting.
    recv_dirty_ = true;
}

void Ras::set_onesided(bool o)
{
    onesided_ = o;
    if (o && recv_dirty_) {
        ctx.use();
        SCHWZ_CUDA(cudaMemsetAsync(mailbox, 0, (size_t)(2 * mbox.recv_stride), ctx.stream));
        SCHWZ_CUDA(cudaMemsetAsync(mailbox + mbox.send_off, 0,
                                   (size_t)(mbox.slots_off - mbox.send_off), ctx.stream));
        recv_dirty_ = false;
    }
}

void Ras::fetch_state(OuterState &out)
{
    ctx.use();
    SCHWZ_CUDA(cudaMemcpyAsync(&out, state, sizeof(out), cudaMemcpyDeviceToHost, ctx.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));
    resnorm = out.resnorm;
    resnorm0 = out.resnorm0;
    gres = out.gres;
    gres0 = out.gres0;
    num_converged = out.num_converged;
    finished = out.stop != 0 && out.error == OUTER_OK;
    finished_iter = out.finished_iter;
}

float Ras::kernel_time_ms(int kind, int reps)
{
    ctx.use();
    double *p = work, *q = work + local_size_x;
    // warm-up launch, then `reps` timed launches back to back
    auto one = [&]() {
        switch (kind) {
        case 0:
            launch_spmv(ctx, *A, 1.0, p, 0.0, nullptr, q, EPI_DOT, p, resnorm_dev + 1, local_size_x,
                        nullptr);
            break;
        case 3:
            launch_spmv(ctx, *A, -1.0, x, 1.0, local_sol, q, EPI_NRM2, nullptr, resnorm_dev + 1,
                        local_size_x, nullptr);
            break;
        // 4: push + unpack, 5 / 6: the two halves on their own.  Timing launches rewrite / reread
        // the buffers of the last epoch: the counters do not move.
        case 4:
            exchange_push(0, true);
            exchange_unpack(0, false, false);
            break;
        case 5: exchange_push(0, true); break;
        case 6: exchange_unpack(0, false, false); break;
        default: SCHWZ_REQUIRE(cg != nullptr, "no CG solver on this subdomain"); cg->bench_step(kind, work);
        }
    };
    if ((kind == 1 || kind == 2) && cg) cg->bench_step(-1, work);
    one();
    SCHWZ_CUDA(cudaEventRecord(ctx.ev_start, ctx.stream));
    for (int i = 0; i < reps; ++i) one();
    SCHWZ_CUDA(cudaEventRecord(ctx.ev_stop, ctx.stream));
    SCHWZ_CUDA(cudaEventSynchronize(ctx.ev_stop));
    float ms = 0.f;
    SCHWZ_CUDA(cudaEventElapsedTime(&ms, ctx.ev_start, ctx.ev_stop));
    return ms / (float)reps;
}

void Ras::set_factors(const int32_t *Lrp, const int32_t *Lci, const double *Lv,
                      const int32_t *perm)
{
    // U = L^T (source/solve.cpp:286-304)
    HostCsr L;
    L.nrows = L.ncols = local_size_x;
    L.rp.assign(Lrp, Lrp + local_size_x + 1);
    L.ci.assign(Lci, Lci + Lrp[local_size_x]);
    L.v.assign(Lv, Lv + Lrp[local_size_x]);
    HostCsr U = transpose(L);
    Ltrs.reset(new TrsPlan(ctx, local_size_x, L.rp.data(), L.ci.data(), L.v.data(), false));
    Utrs.reset(new TrsPlan(ctx, local_size_x, U.rp.data(), U.ci.data(), U.v.data(), true));
    std::vector<int32_t> pv(local_size_x);
    for (int32_t i = 0; i < local_size_x; ++i) pv[i] = perm ? perm[i] : i;
    ctx.release(fperm);
    ctx.release(fperm_col);
    fperm_col = nullptr;
    fperm = ctx.upload(pv.data(), pv.size());
}

// LU variant (UMFPACK branch of the reference, source/solve.cpp:322-385): P A Q = L U with
// local_perm = P (row_permute) and local_inv_perm = Q (inverse row_permute)
void Ras::set_lu_factors(const int32_t *Lrp, const int32_t *Lci, const double *Lv,
                         const int32_t *Urp, const int32_t *Uci, const double *Uv,
                         const int32_t *row_perm, const int32_t *col_perm)
{
    Ltrs.reset(new TrsPlan(ctx, local_size_x, Lrp, Lci, Lv, false));
    Utrs.reset(new TrsPlan(ctx, local_size_x, Urp, Uci, Uv, true));
    std::vector<int32_t> pv(local_size_x), qv(local_size_x);
    for (int32_t i = 0; i < local_size_x; ++i) {
        pv[i] = row_perm ? row_perm[i] : i;
        qv[i] = col_perm ? col_perm[i] : i;
    }
    ctx.release(fperm);
    ctx.release(fperm_col);
    fperm = ctx.upload(pv.data(), pv.size());
    fperm_col = ctx.upload(qv.data(), qv.size());
}

void Ras::connect(int32_t j, void *peer_base, const MailboxLayout &pl, int32_t peer_recv_offset,
                  int32_t peer_flag_slot, bool same_process)
{
    SCHWZ_REQUIRE(j >= 0 && j < (int32_t)nbr_out.size(), "out-neighbour index out of range");
    char *base = (char *)peer_base;
    for (int b = 0; b < 2; ++b)
        out_dst_host_[b][j] = base + b * pl.recv_stride + wire_size() * (size_t)peer_recv_offset;
    out_flag_host_[j] = (unsigned long long *)(base + pl.flags_off) + peer_flag_slot;
    out_conv_host_[j] = (int32_t *)(base + pl.conv_off);
    out_same_process_[j] = same_process ? 1 : 0;
    out_x_host_[j] = (double *)(base + pl.x_off);
    conv_peer_host_[nbr_out[j]] = out_conv_host_[j];
    // the slots the neighbour keeps my elements in (its scatter table for my block)
    const int32_t cnt = out_off_host_[j + 1] - out_off_host_[j];
    if (cnt > 0) {
        ctx.use();
        SCHWZ_CUDA(cudaMemcpyAsync(out_remote_slot_ + out_off_host_[j],
                                   (const int32_t *)(base + pl.slots_off) + peer_recv_offset,
                                   sizeof(int32_t) * (size_t)cnt, cudaMemcpyDefault, ctx.stream));
    }
    peer_tables_dirty_ = true;
}

void Ras::connect_in(int32_t j, void *peer_base, const MailboxLayout &pl, int32_t peer_send_offset)
{
    SCHWZ_REQUIRE(j >= 0 && j < (int32_t)nbr_in.size(), "in-neighbour index out of range");
    const char *base = (const char *)peer_base;
    in_send_host_[j] = base + pl.send_off + wire_size() * (size_t)peer_send_offset;
    in_x_host_[j] = base + pl.x_off;
    conv_peer_host_[nbr_in[j]] = (int32_t *)((char *)peer_base + pl.conv_off);
    peer_tables_dirty_ = true;
}

void Ras::connect_conv(int32_t peer_rank, void *peer_base, const MailboxLayout &pl)
{
    SCHWZ_REQUIRE(peer_rank >= 0 && peer_rank < P, "subdomain id out of range");
    conv_peer_host_[peer_rank] = (int32_t *)((char *)peer_base + pl.conv_off);
    peer_tables_dirty_ = true;
}

void Ras::set_exchange_mode(int32_t mode)
{
    SCHWZ_REQUIRE(mode >= EXCHANGE_PUT_GATHERED && mode <= EXCHANGE_GET_ONE_BY_ONE,
                  "unknown exchange mode");
    exchange_mode = mode;
}

void Ras::upload_peer_tables()
{
    if (!peer_tables_dirty_) return;
    const size_t no = nbr_out.size();
    any_remote_ = false;
    for (size_t j = 0; j < no; ++j) {
        SCHWZ_REQUIRE(out_dst_host_[0][j] != nullptr, "out-neighbour not connected");
        if (!out_same_process_[j]) any_remote_ = true;
    }
    if (no) {
        ctx.use();
        for (int b = 0; b < 2; ++b)
            SCHWZ_CUDA(cudaMemcpyAsync(out_dst_dev_[b], out_dst_host_[b].data(), no * sizeof(void *),
                                       cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(out_flag_dev_, out_flag_host_.data(), no * sizeof(void *),
                                   cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(out_conv_dev_, out_conv_host_.data(), no * sizeof(void *),
                                   cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(out_x_dev_, out_x_host_.data(), no * sizeof(void *),
                                   cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(send_seg_dev_, send_seg_host_.data(), no * sizeof(void *),
                                   cudaMemcpyHostToDevice, ctx.stream));
    }
    const size_t ni = nbr_in.size();
    ctx.use();
    if (ni) {
        SCHWZ_CUDA(cudaMemcpyAsync(in_send_dev_, in_send_host_.data(), ni * sizeof(void *),
                                   cudaMemcpyHostToDevice, ctx.stream));
        SCHWZ_CUDA(cudaMemcpyAsync(in_x_dev_, in_x_host_.data(), ni * sizeof(void *),
                                   cudaMemcpyHostToDevice, ctx.stream));
    }
    SCHWZ_CUDA(cudaMemcpyAsync(conv_peer_dev_, conv_peer_host_.data(), (size_t)P * sizeof(void *),
                               cudaMemcpyHostToDevice, ctx.stream));
    SCHWZ_CUDA(cudaStreamSynchronize(ctx.stream));
    peer_tables_dirty_ = false;
}

// A8 send side: pack x[own] for every out-neighbour and store it into the
// neighbours' receive buffers (epoch parity selects the buffer), publish the
// epoch.  The flag store is always made: a same-process neighbour ignores it.
void Ras::exchange_push(int32_t iter, bool repush)
{
    upload_peer_tables();
    const int32_t no = (int32_t)nbr_out.size();
    const int32_t *stop = stop_ptr();
    // what travels is x[own]; right after a local solve that is solution_vector()[0..local_size)
    // as well, but x is what every variant (and every caller of the stages) agrees on
    if (no > 0 && exchange_mode == EXCHANGE_PUT_GATHERED) {
        if (onesided_) {
            // one receive buffer, no flags: the receiver takes whatever is there (:789-799)
            launch_halo_pack_push(ctx, no, out_off_, out_total_, out_src_, x, out_dst_dev_[0],
                                  nullptr, 0, stop, opt.use_mixed_precision != 0);
        } else {
            // Epochs count exchanges over the lifetime of the subdomain, so the loop may be
            // entered repeatedly (warm-up + timed runs).  Invariant: push_epoch_ is
            // unpack_epoch_ or unpack_epoch_ + 1, and tail_pending <=> a push nobody has unpacked
            // yet is out.  A real push REWRITES such a pending push (same epoch, same buffer,
            // current x) instead of following it with another; a timing / refresh push
            // (repush) rewrites the last epoch without changing what is pending - and becomes
            // the pending push when none is out.
            if (repush) {
                if (push_epoch_ == unpack_epoch_) {
                    ++push_epoch_;
                    tail_pending = true;
                }
            } else if (tail_pending) {
                tail_pending = false;
            } else {
                ++push_epoch_;
            }
            launch_halo_pack_push(ctx, no, out_off_, out_total_, out_src_, x,
                                  out_dst_dev_[push_epoch_ & 1], out_flag_dev_,
                                  (unsigned long long)push_epoch_, stop,
                                  opt.use_mixed_precision != 0);
        }
    } else if (no > 0 && exchange_mode == EXCHANGE_GET_GATHERED) {
        // pack_buffer into my own send buffer; the receivers come and get it (:807-818)
        launch_halo_pack_push(ctx, no, out_off_, out_total_, out_src_, x, send_seg_dev_, nullptr, 0,
                              stop, opt.use_mixed_precision != 0);
    } else if (no > 0 && exchange_mode == EXCHANGE_PUT_ONE_BY_ONE) {
        for (size_t j = 0; j < nbr_out.size(); ++j)
            SCHWZ_REQUIRE(out_x_host_[j] != nullptr, "out-neighbour not connected");
        launch_halo_put_elements(ctx, no, out_off_, out_total_, out_src_, out_remote_slot_, x,
                                 out_x_dev_, stop);
    }
    SCHWZ_CUDA(cudaEventRecord(ev_pushed, ctx.stream));
    last_push_iter = iter;
}

void Ras::wait_push_of(const Ras &nbr)
{
    ctx.use();
    SCHWZ_CUDA(cudaStreamWaitEvent(ctx.stream, nbr.ev_pushed, 0));
}

// A8 receive side: scatter the receive buffer of this epoch into the
// overlap + halo slots of x.
void Ras::exchange_unpack(int32_t iter, bool wait_flags, bool advance)
{
    const int32_t ni = (int32_t)nbr_in.size();
    const int32_t *stop = stop_ptr();
    if (exchange_mode != EXCHANGE_PUT_GATHERED) {
        SCHWZ_REQUIRE(!wait_flags, "only the Put-gathered exchange has a synchronous mode");
        if (ni == 0) return;
        upload_peer_tables();
        if (exchange_mode == EXCHANGE_PUT_ONE_BY_ONE) return;   // the Put was the update
        const bool gathered = exchange_mode == EXCHANGE_GET_GATHERED;
        for (int32_t j = 0; j < ni; ++j)
            SCHWZ_REQUIRE((gathered ? in_send_host_[j] : in_x_host_[j]) != nullptr,
                          "in-neighbour not connected (Get variants need connect_in)");
        // one-by-one moves ValueType elements of x itself: no float mirror (comm_helpers.hpp:58-89)
        launch_halo_pull(ctx, ni, in_off_, in_total_, in_dst_, gathered ? nullptr : in_remote_idx_,
                         gathered ? in_send_dev_ : in_x_dev_, x,
                         gathered && opt.use_mixed_precision != 0, stop);
        return;
    }
    if (onesided_) {
        if (ni == 0) return;
        launch_halo_unpack(ctx, ni, in_total_, in_dst_, mailbox, x, nullptr, 0, nullptr,
                           opt.use_mixed_precision != 0, stop);
        return;
    }
    const int64_t e = advance ? ++unpack_epoch_ : std::max<int64_t>(unpack_epoch_, 1);
    if (ni == 0) return;
    const void *recv = mailbox + (e & 1) * mbox.recv_stride;
    const unsigned long long *flags =
        wait_flags ? (const unsigned long long *)(mailbox + mbox.flags_off) : nullptr;
    // a timed-out wait raises the error word of the loop state (read by the host at its next
    // poll) and the mailbox's own error word (stage-by-stage callers: schwz_b200_ras_halo_error)
    launch_halo_unpack(ctx, ni, in_total_, in_dst_, recv, x, flags, (unsigned long long)e,
                       guarded_ ? &state->error : err_word(), opt.use_mixed_precision != 0, stop);
}

// A9: local_solution = local_rhs - I * x  (source/restricted_schwarz.cpp:992-1017)
void Ras::update_boundary()
{
    const int32_t *stop = stop_ptr();
    sol_in_guess_ = false;
    launch_copy_guarded(ctx, local_size, local_rhs, local_sol, stop);
    if (overlap_size > 0) {
        if (P > 1 && opt.overlap > 0 && I->nnz > 0)
            launch_spmv(ctx, *I, -1.0, x, 1.0, local_rhs + local_size, local_sol + local_size,
                        EPI_NONE, nullptr, nullptr, 0, stop);
        else
            launch_copy_guarded(ctx, overlap_size, local_rhs + local_size, local_sol + local_size,
                                stop);
    }
}

// A10: r = local_solution - A_loc [x_own ; x_overlap], ||r||_2
// (source/solve.cpp:828-843).  extract_local_vector is the identity in the
// compact layout; the norm is fused into the SpMV.
void Ras::local_residual()
{
    launch_spmv(ctx, *A, -1.0, x, 1.0, local_sol, work, EPI_NRM2, nullptr, resnorm_dev,
                local_size_x, stop_ptr());
}

// A12 / A13 (source/solve.cpp:709-781)
void Ras::local_solve()
{
    const int32_t *stop = stop_ptr();
    if (opt.local_solver == 2) {
        const int32_t cap = opt.local_max_iters == -1 ? local_size_x : opt.local_max_iters;
        if (opt.non_symmetric) gmres->solve(local_sol, init_guess, cap, opt.local_tol, stop);
        else cg->solve(local_sol, init_guess, cap, opt.local_tol, stop);
        sol_in_guess_ = true;
        // local_solution <- init_guess (:781) is not materialised: restrict_to_x and the
        // accessors read init_guess (solution_vector())
    } else {
        SCHWZ_REQUIRE(Ltrs && Utrs && fperm, "direct local solve without factors");
        double *perm_sol = work, *tmp = work + local_size_x;
        launch_permute(ctx, local_size_x, fperm, 0, local_sol, perm_sol, stop);
        Ltrs->solve(perm_sol, tmp, stop);
        Utrs->solve(tmp, perm_sol, stop);
        launch_permute(ctx, local_size_x, fperm_col ? fperm_col : fperm, 1, perm_sol, local_sol,
                       stop);
    }
}

const double *Ras::solution_vector() const { return sol_in_guess_ ? init_guess : local_sol; }

// A14: x[own] = local_solution[0 .. local_size)  (source/communicate.cpp:65-94)
void Ras::restrict_to_x() { launch_copy_guarded(ctx, local_size, solution_vector(), x, stop_ptr()); }

double Ras::true_residual_sq()
{
    // ||b_own - (A x)_own||^2 : own rows of the local matrix are complete rows
    // of the global matrix (no entry is dropped when overlap >= 2)
    launch_spmv(ctx, *A, -1.0, x, 1.0, local_rhs, work, EPI_NRM2SQ, nullptr, resnorm_dev + 1,
                local_size, nullptr);
    double out = 0.0;
    SCHWZ_CUDA(cudaMemcpyAsync(&out, resnorm_dev + 1, sizeof(double), cudaMemcpyDeviceToHost,
                               ctx.stream));
    ctx.sync();
    return out;
}

void Ras::conv_forward(int32_t converged_all_local)
{
    upload_peer_tables();
    launch_conv_forward(ctx, P, rank, converged_all_local, conv(), conv_sent,
                        (int32_t)nbr_out.size(), out_conv_dev_, num_converged_dev);
}

void Ras::conv_accumulate(int32_t converged_all_local)
{
    for (int32_t q = 0; q < P; ++q)
        SCHWZ_REQUIRE(conv_peer_host_[q] != nullptr,
                      "accumulate convergence check: every subdomain's flags must be connected "
                      "(schwz_b200_ras_connect_conv)");
    upload_peer_tables();
    launch_conv_accumulate(ctx, P, rank, converged_all_local, conv(), conv_peer_dev_,
                           num_converged_dev);
}

// one-sided decision on the device: ratio test on the residual norm the SpMV left in
// resnorm_dev, flag protocol, break test (source/solve.cpp:913-943)
void Ras::conv_decide(int32_t protocol, double tol, int32_t check, int32_t iter,
                      double *history_slot)
{
    upload_peer_tables();
    if (protocol == 1) {
        if (rank > 0) SCHWZ_REQUIRE(conv_peer_host_[(rank - 1) / 2] != nullptr, "tree parent not connected");
        for (int32_t p = 2 * rank + 1; p <= 2 * rank + 2; ++p)
            if (p < P) SCHWZ_REQUIRE(conv_peer_host_[p] != nullptr, "tree child not connected");
    } else if (protocol == 2) {
        for (int32_t q = 0; q < P; ++q)
            SCHWZ_REQUIRE(conv_peer_host_[q] != nullptr,
                          "accumulate convergence check: every subdomain's flags must be "
                          "connected (schwz_b200_ras_connect_conv)");
    }
    launch_ras_conv_decide(ctx, protocol, P, rank, state, resnorm_dev, tol, check, iter,
                           history_slot, conv(), conv_sent, (int32_t)nbr_out.size(), out_conv_dev_,
                           conv_peer_dev_, num_converged_dev);
}

void Ras::conv_tree(int32_t converged_all_local)
{
    upload_peer_tables();
    if (rank > 0) SCHWZ_REQUIRE(conv_peer_host_[(rank - 1) / 2] != nullptr, "tree parent not connected");
    for (int32_t p = 2 * rank + 1; p <= 2 * rank + 2; ++p)
        if (p < P) SCHWZ_REQUIRE(conv_peer_host_[p] != nullptr, "tree child not connected");
    launch_conv_tree(ctx, P, rank, converged_all_local, conv(), conv_peer_dev_, num_converged_dev);
}

// =============================================================================
// The outer loop over the subdomains of this process.  One host thread ENQUEUES the stages of
// all local subdomains, a few outer iterations ahead of the device; nothing is read back inside
// an iteration.  The convergence decision is taken on the device (ras_decide_kernel /
// ras_conv_decide_kernel) and raises a per-subdomain stop word that every later launch honours,
// so whatever was enqueued past the break point of source/schwarz_base.cpp:432-433 runs as
// no-ops.  The host looks at a snapshot of the loop state once per chunk of iterations (the
// snapshot taken at the end of chunk c is examined before chunk c + 2 is enqueued: the same
// deterministic point on every process, so the NCCL call counts stay equal).
//
// Synchronous mode, per iteration and subdomain:
//   unpack (waits for the neighbours' epoch) -> boundary update -> residual + norm ->
//   [hub: gather the local norms, ncclAllGather across processes, ordered sum, decision] ->
//   local solve -> restriction -> PUSH of the new boundary values for the next iteration.
// The push sits at the tail of the iteration that produced the values (the reference sends at
// the head of the next one and has a hook to overlap only its MPI_Wait,
// source/restricted_schwarz.cpp:887-892, 965): the transfer and its flag latency run while the
// neighbours are still in their own local solves, and the next iteration's unpack usually finds
// the epoch already published.
// =============================================================================
struct Ras::Hub {
    const Ctx *ctx = nullptr;
    std::vector<Ras *> members;
    int32_t P = 0;
    const double **norm_ptrs = nullptr;
    OuterState **state_ptrs = nullptr;
    int32_t *slot = nullptr;
    double *mine = nullptr, *all = nullptr, *history = nullptr;
    size_t history_cap = 0;
    OuterState *pinned = nullptr;            // 2 snapshots x nl
    std::vector<cudaEvent_t> ev_chunk[2];
    cudaEvent_t ev_decided = nullptr, ev_comm = nullptr;
    ~Hub()
    {
        if (!ctx) return;
        cudaSetDevice(ctx->device);
        for (void *p : {(void *)norm_ptrs, (void *)state_ptrs, (void *)slot, (void *)mine,
                        (void *)all, (void *)history})
            if (p) cudaFree(p);
        if (pinned) cudaFreeHost(pinned);
        for (auto &v : ev_chunk)
            for (cudaEvent_t e : v) cudaEventDestroy(e);
        if (ev_decided) cudaEventDestroy(ev_decided);
        if (ev_comm) cudaEventDestroy(ev_comm);
    }
};

static Ras::Hub &ensure_hub(std::vector<Ras *> &subs, int32_t P)
{
    Ras *owner = subs[0];
    if (owner->hub && owner->hub->members == subs && owner->hub->P == P) return *owner->hub;
    owner->hub.reset(new Ras::Hub());
    Ras::Hub &H = *owner->hub;
    const Ctx &ctx = owner->ctx;
    H.ctx = &ctx;
    H.members = subs;
    H.P = P;
    const size_t nl = subs.size();
    // the hub's kernels run on the first subdomain's device and touch the other subdomains'
    // words: peers must see each other when the subdomains of a process span devices
    for (Ras *r : subs)
        if (r->ctx.device != ctx.device) {
            int can = 0;
            SCHWZ_CUDA(cudaDeviceCanAccessPeer(&can, ctx.device, r->ctx.device));
            SCHWZ_REQUIRE(can, "subdomains of one process on devices without peer access");
            ctx.use();
            cudaError_t e = cudaDeviceEnablePeerAccess(r->ctx.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SCHWZ_CUDA(e);
            cudaGetLastError();
        }
    // norms in rank order (the order the allgather / the ordered sum need)
    std::vector<Ras *> by_rank(subs);
    std::sort(by_rank.begin(), by_rank.end(), [](Ras *a, Ras *b) { return a->rank < b->rank; });
    std::vector<const double *> np;
    for (Ras *r : by_rank) np.push_back(r->resnorm_dev);
    std::vector<OuterState *> sp;
    std::vector<int32_t> sl;
    for (Ras *r : subs) {
        sp.push_back(r->state);
        sl.push_back(r->rank);
    }
    H.norm_ptrs = ctx.upload(np.data(), nl);
    H.state_ptrs = ctx.upload(sp.data(), nl);
    H.slot = ctx.upload(sl.data(), nl);
    H.mine = ctx.alloc_zero<double>(nl);
    H.all = ctx.alloc_zero<double>((size_t)P);
    SCHWZ_CUDA(cudaMallocHost((void **)&H.pinned, 2 * nl * sizeof(OuterState)));
    for (auto &v : H.ev_chunk) {
        v.resize(nl);
        for (size_t i = 0; i < nl; ++i) {
            subs[i]->ctx.use();
            SCHWZ_CUDA(cudaEventCreateWithFlags(&v[i], cudaEventDisableTiming));
        }
    }
    ctx.use();
    SCHWZ_CUDA(cudaEventCreateWithFlags(&H.ev_decided, cudaEventDisableTiming));
    SCHWZ_CUDA(cudaEventCreateWithFlags(&H.ev_comm, cudaEventDisableTiming));
    ctx.sync();
    return H;
}

// One synchronous exchange outside the loop: every subdomain's overlap / halo entries of x
// become its neighbours' current values (what the final residual check needs after a run that
// ended on its iteration budget, source/solve.cpp:1025-1085).  Collective over all processes.
void ras_refresh_halo(std::vector<Ras *> &subs, int32_t P)
{
    if (P < 2) return;
    std::vector<Ras *> by_rank(P, nullptr);
    for (Ras *r : subs) by_rank[r->rank] = r;
    for (Ras *r : subs) {
        r->set_exchange_mode(EXCHANGE_PUT_GATHERED);
        r->set_onesided(false);
        r->exchange_push(0);
    }
    for (Ras *r : subs) {
        bool remote = false;
        for (int32_t p : r->nbr_in) {
            Ras *s = by_rank[p];
            if (s) r->wait_push_of(*s);
            else remote = true;
        }
        r->exchange_unpack(0, remote);
    }
    for (Ras *r : subs) r->ctx.sync();
}

void ras_run(std::vector<Ras *> &subs, const LoopOptions &o, LoopResult &res,
             double *history)
{
    const int nl = (int)subs.size();
    const int P = o.num_subdomains;
    SCHWZ_REQUIRE(nl > 0, "no subdomains");
    const bool multi_process = o.comm != nullptr && o.comm->nranks > 1;
    if (!multi_process) SCHWZ_REQUIRE(nl == P, "all subdomains must be local without a communicator");
    if (multi_process) {
        SCHWZ_REQUIRE(nl * o.comm->nranks == P, "subdomains must be spread evenly over the processes");
        for (int i = 0; i < nl; ++i)
            SCHWZ_REQUIRE(subs[i]->rank == o.comm->rank * nl + i,
                          "process p must hold the subdomains [p*nl, (p+1)*nl) in order");
    }
    SCHWZ_REQUIRE(o.enable_onesided || o.exchange_mode == EXCHANGE_PUT_GATHERED,
                  "Get / one-by-one exchanges exist in one-sided mode only");
    // map rank -> local subdomain (for same-process event waits)
    std::vector<Ras *> by_rank(P, nullptr);
    for (Ras *r : subs) by_rank[r->rank] = r;
    Ras::Hub &H = ensure_hub(subs, P);
    if (multi_process && !o.enable_onesided && P != nl) {
        // slot of local subdomain i inside the allgathered array = its rank; the local norms are
        // contiguous in rank order by the requirement above
        SCHWZ_REQUIRE(o.comm->ctx != nullptr, "communicator without a context");
    }
    for (Ras *r : subs) {
        r->set_exchange_mode(o.enable_onesided ? o.exchange_mode : EXCHANGE_PUT_GATHERED);
        r->set_onesided(o.enable_onesided != 0);
        r->set_guarded(true);
    }
    struct Unguard {
        std::vector<Ras *> &s;
        ~Unguard()
        {
            for (Ras *r : s) r->set_guarded(false);
        }
    } unguard{subs};

    // entry: a subdomain that converged in an earlier call stays finished until it is reset
    std::vector<OuterState> st(nl);
    int alive = 0;
    res.host_stream_syncs = res.host_event_waits = 0;
    for (int i = 0; i < nl; ++i) {
        subs[i]->fetch_state(st[i]);
        ++res.host_stream_syncs;
        if (!st[i].stop) ++alive;
    }
    auto fill_result = [&](int iters_if_running) {
        int fin = 0, last = -1;
        for (int i = 0; i < nl; ++i) {
            if (st[i].stop && st[i].error == OUTER_OK) {
                ++fin;
                last = std::max(last, st[i].finished_iter);
            }
        }
        res.converged = fin == nl ? 1 : 0;
        res.iters = fin == nl ? std::max(last, 0) : iters_if_running;
        res.global_resnorm = st[0].gres;
        res.global_resnorm0 = st[0].gres0;
    };
    if (alive == 0) {
        fill_result(0);
        res.iters = 0;
        res.elapsed_s = 0.0;
        return;
    }
    if (history) {
        const size_t need = (size_t)o.max_iters * nl;
        if (H.history_cap < need) {
            H.ctx->release(H.history);
            H.history = H.ctx->alloc<double>(need);
            H.history_cap = need;
        }
        H.ctx->use();
        SCHWZ_CUDA(cudaMemsetAsync(H.history, 0, need * sizeof(double), H.ctx->stream));
    }
    for (Ras *r : subs) {
        r->ctx.sync();
        ++res.host_stream_syncs;
    }

    const cudaStream_t hub_stream = H.ctx->stream;
    const double tol = o.tolerance;
    const int32_t protocol = o.conv_tree ? 1 : (o.conv_accumulate ? 2 : 0);
    auto enqueue_iteration = [&](int iter) {
        const int32_t check =
            (tol > 0.0 &&
             (o.iter_offset ? ((iter > (o.max_iters * 0.05)) || o.max_iters < 1000) : true)) ? 1 : 0;
        if (o.enable_onesided) {
            // every subdomain runs its whole loop body on its own stream: no ordering between
            // subdomains at all (one-sided semantics, source/restricted_schwarz.cpp:715-852)
            for (int i = 0; i < nl; ++i) {
                Ras *r = subs[i];
                if (P > 1 && iter > 0) {   // one-sided skips iteration 0 (:725)
                    r->exchange_push(iter);
                    r->exchange_unpack(iter, false);
                }
                r->update_boundary();
                r->local_residual();
                r->conv_decide(protocol, tol, check, iter,
                               history ? H.history + (size_t)iter * nl + i : nullptr);
                r->local_solve();
                r->restrict_to_x();
            }
            return;
        }
        // ---- 0 boundary exchange: the values were pushed at the tail of the previous pass ----
        if (P > 1) {
            for (Ras *r : subs) {
                bool remote = false;
                for (int32_t p : r->nbr_in) {
                    Ras *s = by_rank[p];
                    if (s) r->wait_push_of(*s);
                    else remote = true;
                }
                r->exchange_unpack(iter, remote);
            }
        }
        // ---- 1 boundary update, 2 convergence check ----------------------------------------
        for (Ras *r : subs) {
            r->update_boundary();
            r->local_residual();
            if (r != subs[0]) {
                r->ctx.use();
                SCHWZ_CUDA(cudaEventRecord(r->ev_resid, r->ctx.stream));
            }
        }
        H.ctx->use();
        for (Ras *r : subs)
            if (r != subs[0]) SCHWZ_CUDA(cudaStreamWaitEvent(hub_stream, r->ev_resid, 0));
        const double *all = H.mine;
        launch_gather_norms(*H.ctx, nl, H.norm_ptrs, H.mine);
        if (multi_process) {
            // MPI_Allgather of the local norms (source/solve.cpp:890-891) as an on-stream
            // ncclAllGather; the communicator's stream is ordered behind the hub's and back
            Comm &c = *o.comm;
            if (c.ctx->stream != hub_stream) {
                SCHWZ_CUDA(cudaEventRecord(H.ev_comm, hub_stream));
                c.ctx->use();
                SCHWZ_CUDA(cudaStreamWaitEvent(c.ctx->stream, H.ev_comm, 0));
            }
            comm_allgather_f64(c, H.mine, nl, H.all);
            if (c.ctx->stream != hub_stream) {
                SCHWZ_CUDA(cudaEventRecord(H.ev_comm, c.ctx->stream));
                H.ctx->use();
                SCHWZ_CUDA(cudaStreamWaitEvent(hub_stream, H.ev_comm, 0));
            }
            all = H.all;
        }
        launch_ras_decide(*H.ctx, P, nl, all, H.slot, H.state_ptrs, tol, check,
                          o.enable_global_check, iter, history ? H.history : nullptr);
        SCHWZ_CUDA(cudaEventRecord(H.ev_decided, hub_stream));
        for (Ras *r : subs)
            if (r != subs[0]) {
                r->ctx.use();
                SCHWZ_CUDA(cudaStreamWaitEvent(r->ctx.stream, H.ev_decided, 0));
            }
        // ---- 3 local solve, 4 restriction, then the push for the next iteration -------------
        for (Ras *r : subs) {
            r->local_solve();
            r->restrict_to_x();
            if (P > 1) r->exchange_push(iter + 1);
        }
    };

    auto t0 = std::chrono::steady_clock::now();
    if (P > 1 && !o.enable_onesided)
        for (Ras *r : subs) {
            // the exchange of iteration 0: x as it stands now (a push left over from the
            // previous call is rewritten with the same epoch)
            r->exchange_push(0);
        }
    // iterations enqueued between two looks at the loop state (SCHWZ_B200_OUTER_CHUNK)
    const char *env_chunk_s = std::getenv("SCHWZ_B200_OUTER_CHUNK");
    const int env_chunk = env_chunk_s ? std::atoi(env_chunk_s) : 0;
    const int K = std::max(1, o.chunk > 0 ? o.chunk : (env_chunk > 0 ? env_chunk : 4));
    int enq = 0, chunk_idx = 0;
    bool done = false, failed = false;
    auto examine = [&](int slot) {
        int stopped = 0;
        for (int i = 0; i < nl; ++i) {
            subs[i]->ctx.use();
            SCHWZ_CUDA(cudaEventSynchronize(H.ev_chunk[slot][i]));
            ++res.host_event_waits;
            const OuterState &S = H.pinned[(size_t)slot * nl + i];
            if (S.error != OUTER_OK) failed = true;
            if (S.stop) ++stopped;
        }
        if (stopped == nl || failed) done = true;
    };
    while (enq < o.max_iters && !done) {
        const int end = std::min(o.max_iters, enq + K);
        const int slot = chunk_idx & 1;
        for (; enq < end; ++enq) {
            enqueue_iteration(enq);
            // the loop state of this iteration travels to the host behind it (no wait here)
            for (int i = 0; i < nl; ++i) {
                subs[i]->ctx.use();
                SCHWZ_CUDA(cudaMemcpyAsync(&H.pinned[(size_t)slot * nl + i], subs[i]->state,
                                           sizeof(OuterState), cudaMemcpyDeviceToHost,
                                           subs[i]->ctx.stream));
            }
        }
        for (int i = 0; i < nl; ++i) {
            subs[i]->ctx.use();
            SCHWZ_CUDA(cudaEventRecord(H.ev_chunk[slot][i], subs[i]->ctx.stream));
        }
        if (chunk_idx >= 1) examine(slot ^ 1);
        ++chunk_idx;
    }
    if (P > 1 && !o.enable_onesided)
        for (Ras *r : subs) r->tail_pending = true;
    for (Ras *r : subs) r->ctx.sync();
    H.ctx->sync();
    if (multi_process) o.comm->ctx->sync();
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < nl; ++i) subs[i]->fetch_state(st[i]);
    res.host_stream_syncs += 2 * nl + 1 + (multi_process ? 1 : 0);
    if (history) {
        H.ctx->use();
        SCHWZ_CUDA(cudaMemcpyAsync(history, H.history, (size_t)o.max_iters * nl * sizeof(double),
                                   cudaMemcpyDeviceToHost, H.ctx->stream));
        H.ctx->sync();
    }
    for (int i = 0; i < nl; ++i) {
        switch (st[i].error) {
        case OUTER_HALO_TIMEOUT:
            throw std::runtime_error("halo exchange timed out: subdomain " +
                                     std::to_string(subs[i]->rank) +
                                     " waited for a neighbour's boundary values longer than "
                                     "SCHWZ_B200_HALO_TIMEOUT_MS");
        case OUTER_NAN: throw std::runtime_error("residual norm is NaN");
        case OUTER_DIVERGED: throw std::runtime_error("diverged");   // schwarz_base.cpp:424-428
        default: break;
        }
    }
    fill_result(enq);
    res.elapsed_s = std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace schwz_b200
