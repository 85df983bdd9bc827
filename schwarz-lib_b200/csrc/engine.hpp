// Device-resident local solvers and the per-subdomain state of the RAS
// iteration.  Internal; the C ABI in include/schwz_b200.h wraps these.
#pragma once
#include <memory>
#include <vector>

#include "device.hpp"
#include "setup.hpp"

namespace schwz_b200 {

// ---- NCCL (loaded lazily with dlopen; only the residual-norm allgather) ------
struct Comm {
    void *nccl = nullptr;   // ncclComm_t
    const Ctx *ctx = nullptr;
    int nranks = 1, rank = 0;
    double *dev_in = nullptr, *dev_out = nullptr;   // scratch for small gathers
    int cap = 0;
    ~Comm();
};
void comm_unique_id(void *id128);
Comm *comm_create(const Ctx &ctx, const void *id128, int nranks, int rank);
void comm_allgather_f64(Comm &c, const double *dev_in, int count, double *dev_out);

// one-CTA solvers for small subdomains (small_solvers.cu)
extern bool g_use_cg_graph;        // SCHWZ_B200_NO_CG_GRAPH=1 turns the CG graph replay off
extern bool g_trs_pdl;             // SCHWZ_B200_TRS_NO_PDL=1 turns programmatic dependent launch off
extern bool g_gmres_cgs2;          // SCHWZ_B200_GMRES_MGS=1 turns CGS2 off in the one-CTA GMRES
extern bool g_use_small_solvers;   // SCHWZ_B200_NO_SMALL=1 turns them off (A/B measurements)
bool cg_small_fits(int64_t n);
void launch_cg_small(const Ctx &ctx, const DeviceCsr &A, const double *b, double *x,
                     int32_t max_iters, double tol, CgScalars *out, const int32_t *outer_stop);
bool gmres_small_fits(int64_t n, int m);
void launch_gmres_small(const Ctx &ctx, const DeviceCsr &A, const double *b, double *x, double *V,
                        int32_t m, int32_t max_iters, double tol, double *resnorm_out,
                        double *r0_out, int32_t *total_out, const int32_t *outer_stop = nullptr);

// ---- local preconditioners (source/solve.cpp:486-652; precond.cu) --------------
class TrsPlan;
enum PrecondKind { PRECOND_NONE = 0, PRECOND_BLOCK_JACOBI = 1, PRECOND_ILU = 2, PRECOND_ISAI = 3 };

// what generation produces (host; setup time)
struct PrecondData {
    std::vector<int32_t> block_ptrs;   // block-Jacobi: block b = rows [block_ptrs[b], block_ptrs[b+1])
    std::vector<int64_t> block_off;    //   offset of its inverse in `blocks`
    std::vector<double> blocks;        //   inverse diagonal blocks, column-major, ld = block size
    HostCsr L, U, Li, Ui;              // ILU(0) factors and their sparse approximate inverses
    void generate(int32_t n, const int32_t *rp, const int32_t *ci, const double *v, int32_t kind,
                  int32_t max_block_size);
};

class Preconditioner {
public:
    // host CSR of the local matrix; generation happens here (setup time)
    Preconditioner(const Ctx &ctx, int32_t n, const int32_t *rp, const int32_t *ci,
                   const double *v, int32_t kind, int32_t max_block_size);
    ~Preconditioner();
    // z = M^-1 r on ctx.stream; dot_result (device, optional) <- r.z; the launches are
    // no-ops when *stop != 0
    void apply(const double *r, double *z, double *dot_result, const int32_t *stop);
    int32_t kind() const { return kind_; }
    int32_t size() const { return n_; }
    int64_t bytes_per_apply() const;
    bool uses_level_graphs() const;   // ILU: the triangular solves are CUDA graphs of their own
    void release_host();   // drop the host copy below (large subdomains)
    PrecondData host;      // what was generated (parity tests read it through the C ABI)

private:
    const Ctx &ctx_;
    int32_t n_, kind_;
    int64_t inv_bytes_ = 0;
    int32_t *dev_block_ptrs_ = nullptr, *dev_row_block_ = nullptr;
    int64_t *dev_block_off_ = nullptr;
    double *dev_blocks_ = nullptr, *tmp_ = nullptr, *in_ = nullptr;
    std::unique_ptr<TrsPlan> Ltrs_, Utrs_;
    std::unique_ptr<DeviceCsr> dev_Li_, dev_Ui_;
};

// ---- CG (Ginkgo Cg semantics; source/solve.cpp:469-478, 572-652, 746-754) ----
class CgSolver {
public:
    CgSolver(const Ctx &ctx, const DeviceCsr &A);
    ~CgSolver();
    // z = M^-1 r ; rho = r.z ; p = z + (rho/prev_rho) p  (with_preconditioner, :581-638)
    void set_precond(Preconditioner *M);
    // x holds the warm start; asynchronous on ctx.stream.  outer_stop: optional
    // device flag that turns the whole solve into a no-op.
    void solve(const double *b, double *x, int32_t max_iters, double tol,
               const int32_t *outer_stop = nullptr);
    void result(int32_t *iters, double *resnorm, double *resnorm0);   // synchronises
    int64_t bytes_per_iteration() const;
    void bench_step(int kind, double *scratch_x);   // 1: r update, 2: x/p update (timing only)

private:
    void iteration(double *x, cudaGraphConditionalHandle loop = 0);
    const Ctx &ctx_;
    const DeviceCsr &A_;
    int64_t n_;
    double *r_ = nullptr, *p_ = nullptr, *q_ = nullptr, *z_ = nullptr;
    Preconditioner *M_ = nullptr;
    CgScalars *s_ = nullptr;
    int32_t *pinned_stop_ = nullptr;   // 2 slots
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    // the captured launch sequence of a whole solve (fixed buffers, fixed budget)
    struct Captured {
        const double *b;
        double *x;
        int32_t max_iters;
        double tol;
        const int32_t *outer_stop;
        cudaGraphExec_t exec;
        int launches;
    };
    std::vector<Captured> graphs_;
    int plain_solves_ = 0;
    bool not_graphable_ = false;   // a capture failed once: plain enqueue from then on
};

// ---- GMRES(m) (Ginkgo Gmres semantics; source/solve.cpp:486-567) -------------
class GmresSolver {
public:
    GmresSolver(const Ctx &ctx, const DeviceCsr &A, int32_t restart);
    ~GmresSolver();
    // right preconditioning: w = A M^-1 v_k ; x += M^-1 (V y)  (:496-556)
    void set_precond(Preconditioner *M);
    // outer_stop: optional device flag that turns the whole solve into a no-op
    void solve(const double *b, double *x, int32_t max_iters, double tol,
               const int32_t *outer_stop = nullptr);
    void result(int32_t *iters, double *resnorm, double *resnorm0);

private:
    const Ctx &ctx_;
    const DeviceCsr &A_;
    int64_t n_;
    int32_t m_;
    double *V_ = nullptr;    // (m+1) x n basis, row-major by vector
    double *w_ = nullptr;
    double *pv_ = nullptr, *upd_ = nullptr;   // preconditioned vector, V y (only with M_)
    Preconditioner *M_ = nullptr;
    double *small_ = nullptr;   // Hessenberg, rotations, g, y and scalars
    int32_t *pinned_stop_ = nullptr;
    cudaEvent_t ev_[2] = {nullptr, nullptr};
};

// ---- level-scheduled sparse triangular solve ---------------------------------
class TrsPlan {
public:
    TrsPlan(const Ctx &ctx, int32_t n, const int32_t *rp, const int32_t *ci, const double *v,
            bool upper);
    ~TrsPlan();
    // stop: optional device flag, the launches are no-ops when it is set
    void solve(const double *b, double *x, const int32_t *stop = nullptr);
    int32_t num_levels() const { return num_levels_; }
    int32_t num_launches() const { return (int32_t)segments_.size() + num_blocks_; }
    int64_t nnz() const { return nnz_; }

private:
    const Ctx &ctx_;
    int32_t n_, num_levels_ = 0;
    int64_t nnz_ = 0;
    bool upper_;
    // the factor in level order, off-diagonal entries only: position i = row order_[i], entries
    // [rp_[i], rp_[i+1]), reciprocal diagonal inv_diag_[i]
    int32_t *rp_ = nullptr, *ci_ = nullptr, *order_ = nullptr;
    double *v_ = nullptr, *inv_diag_ = nullptr;
    std::vector<int32_t> level_ptr_;   // host: level l = positions [level_ptr[l], level_ptr[l+1])
    // Runs of consecutive small levels (the dense top of a nested-dissection factor: 3 % of
    // the rows, 60 % of the non-zeros, 90 % of the levels) are solved block-wise: a block =
    // up to kTrsBlock consecutive rows of the level order, x_K = Dinv_K (b_K - Lout_K x),
    // Dinv_K the explicit inverse of the block's own triangular part.
    struct Segment {
        int32_t kind;      // 0: one wide level, 1: one block
        int32_t a, b;      // kind 0: level index, unused; kind 1: first position, rows
        int64_t dinv_off;  // kind 1: offset of the dense inverse (column-major, ld = rows)
        int32_t mode;      // kind 0: 1 = one thread per row (short rows); kind 1: 2 = a CTA per
                           // row for the part outside the block (long rows); 0 = a warp per row
    };
    std::vector<Segment> segments_;
    int32_t *chain_rp_ = nullptr, *chain_ci_ = nullptr;   // outside entries, indexed by position
    double *chain_v_ = nullptr, *dinv_ = nullptr, *block_t_ = nullptr;
    int32_t num_blocks_ = 0;
    struct Captured {
        const double *b;
        double *x;
        const int32_t *stop;
        cudaGraphExec_t exec;
    };
    std::vector<Captured> graphs_;
    // dependency-driven solve (one persistent kernel): work items in dependency order,
    // claimed dynamically; a value that is still the "unset" bit pattern is not ready yet
    struct Item {
        int32_t kind, a, b, c;
    };
    Item *items_ = nullptr;
    int32_t num_items_ = 0, num_chain_rows_ = 0;
    int32_t *counter_ = nullptr;        // [0] next item, [1] error / abort word
    double *t_ = nullptr;               // block right-hand sides, one slot per chain row
    int32_t *blk_pos0_ = nullptr, *blk_nb_ = nullptr, *blk_ord0_ = nullptr;
    int64_t *blk_dinv_off_ = nullptr;
    void solve_levels(const double *b, double *x, const int32_t *stop);
    void solve_flow(const double *b, double *x, const int32_t *stop);

public:
    int32_t num_items() const { return num_items_; }
    bool uses_level_graph() const;      // which of the two solve() takes (see solvers.cu)
    int32_t error();                    // synchronises; non-zero: a wait timed out
};

struct RasOptions {
    double tolerance = 1e-6, local_tol = 1e-12;
    int32_t local_max_iters = -1, local_solver = 2, non_symmetric = 0, restart_iter = 1,
            overlap = 2;
    // settings.use_mixed_precision with MixedValueType = float: halo values travel as floats
    // in the gathered exchanges (restricted_schwarz.cpp:483-603, 769-787, 898-903)
    int32_t use_mixed_precision = 0;
    // metadata.local_precond / precond_max_block_size (PrecondKind; iterative local solve only)
    int32_t local_precond = 0, precond_max_block_size = 16;
};

// Layout of the peer-visible mailbox of a subdomain (byte offsets from base).
struct MailboxLayout {
    int64_t recv_stride = 0;   // bytes between the two receive buffers (epoch parity)
    int64_t flags_off = 0;     // n_in epoch words (u64)
    int64_t conv_off = 0;      // P convergence flags (i32)
    int64_t err_off = 0;       // 1 i32 error word
    int64_t send_off = 0;      // send buffer (out_total doubles): what Get-gathered peers read
    int64_t slots_off = 0;     // in_total i32: compact slot of every received element in x
    int64_t x_off = 0;         // x itself (local_size_x + n_halo doubles): one-by-one Put / Get
    int64_t bytes = 0;
    static MailboxLayout make(int64_t in_total, int32_t n_in, int32_t P, int64_t out_total,
                              int64_t x_len);
};

// How the halo values travel (Settings::comm_settings of the reference:
// enable_put / enable_get x enable_one_by_one, restricted_schwarz.cpp:753-851).
enum ExchangeMode {
    EXCHANGE_PUT_GATHERED = 0,     // pack + peer stores into the neighbour's receive buffer, unpack
    EXCHANGE_GET_GATHERED = 1,     // pack into the own send buffer; the receiver pulls + scatters
    EXCHANGE_PUT_ONE_BY_ONE = 2,   // element-wise peer stores straight into the neighbour's x
    EXCHANGE_GET_ONE_BY_ONE = 3    // element-wise peer loads straight from the neighbour's x
};

// One subdomain.  Vector layout in HBM (compact numbering = g2l - 1 of the
// reference, source/restricted_schwarz.cpp:155-180, 285-295):
//   x        [ own (local_size) | overlap (overlap_size) | halo (n_halo) ]
//   local_*  [ own | overlap ]                              (local_size_x)
class Ras {
public:
    Ras(const Ctx &ctx, Setup &setup, int32_t rank, const double *host_rhs_global,
        const RasOptions &opt);
    ~Ras();

    void set_factors(const int32_t *Lrp, const int32_t *Lci, const double *Lv,
                     const int32_t *perm);
    void set_lu_factors(const int32_t *Lrp, const int32_t *Lci, const double *Lv,
                        const int32_t *Urp, const int32_t *Uci, const double *Uv,
                        const int32_t *row_perm, const int32_t *col_perm);
    void connect(int32_t j_out, void *peer_base, const MailboxLayout &peer_layout,
                 int32_t peer_recv_offset, int32_t peer_flag_slot, bool same_process);
    // in-neighbour j_in (Get variants): its mailbox and the offset of my block inside its
    // send buffer (= get_displacements[neighbour], restricted_schwarz.cpp:642-658)
    void connect_in(int32_t j_in, void *peer_base, const MailboxLayout &peer_layout,
                    int32_t peer_send_offset);
    // any subdomain's convergence words (the tree's parent / children need not be halo
    // neighbours)
    void connect_conv(int32_t peer_rank, void *peer_base, const MailboxLayout &peer_layout);
    void set_exchange_mode(int32_t mode);

    // stages of SchwarzBase::run's loop body (source/schwarz_base.cpp:387-452); every launch
    // honours state->stop when `guarded` (the device-side outer loop), none does otherwise
    // (stage-by-stage callers decide on the host).
    // Exchange epochs (synchronous Put-gathered mode): every push carries the next epoch into the
    // neighbours' receive buffer of that parity and publishes it; every unpack consumes the
    // next epoch.  A push that follows one nobody has unpacked yet (the tail push of the previous
    // ras_run) rewrites it: same epoch, same buffer, current x.  repush = a timing / refresh
    // launch: writes the last epoch again without changing what is pending.
    // One-sided mode uses receive buffer 0 only and no flags: the receiver always sees the
    // freshest values, as with the reference's single recv_buffer.
    void exchange_push(int32_t iter, bool repush = false);
    void exchange_unpack(int32_t iter, bool wait_flags, bool advance = true);
    void update_boundary();
    void local_residual();      // -> *resnorm_dev
    void local_solve();
    void restrict_to_x();
    void wait_push_of(const Ras &nbr);
    double true_residual_sq();
    const double *solution_vector() const;   // what local_solve left: init_guess / local_sol
    void set_guarded(bool g) { guarded_ = g; }
    void set_onesided(bool o);
    bool tail_pending = false;   // the last push of the previous ras_run has not been unpacked
    void fetch_state(OuterState &out);       // synchronises
    // host-buffer entry points of the plugin path (what SolverRAS::initialize /
    // run move across PCIe: rhs in, solution out)
    void upload_rhs(const double *host_rhs_global);          // async H2D via pinned staging
    void download_solution(double *host_solution_global);    // async D2H of the own block
    void reset_state();                                      // x, init_guess <- 0; norms unlatched
    // average duration (ms) of one launch of a hot kernel on this subdomain's
    // data, CUDA events on the subdomain's stream: 0 SpMV+dot (CG q = A p),
    // 1 CG r update, 2 CG x/p update, 3 residual SpMV+norm, 4 halo push+unpack, 5 push, 6 unpack
    float kernel_time_ms(int kind, int reps);

    const Ctx &ctx;
    int32_t rank, P;
    RasOptions opt;
    int32_t local_size, local_size_x, overlap_size, n_halo, first_row;
    std::vector<int32_t> nbr_in, nbr_out, in_count, out_count;
    MailboxLayout mbox;
    char *mailbox = nullptr;
    double *x = nullptr, *local_rhs = nullptr, *local_sol = nullptr, *init_guess = nullptr,
           *work = nullptr;
    double *resnorm_dev = nullptr;
    OuterState *state = nullptr;           // device-resident loop bookkeeping
    cudaEvent_t ev_resid = nullptr;        // residual norm of this iteration is in resnorm_dev
    int32_t *num_converged_dev = nullptr, *conv_sent = nullptr;
    std::unique_ptr<DeviceCsr> A, I;
    std::unique_ptr<Preconditioner> precond;
    std::unique_ptr<CgSolver> cg;
    std::unique_ptr<GmresSolver> gmres;
    std::unique_ptr<TrsPlan> Ltrs, Utrs;
    int32_t *fperm = nullptr, *fperm_col = nullptr;   // row order P; column order Q (LU only)
    cudaEvent_t ev_pushed = nullptr;
    int32_t last_push_iter = -1;
    int64_t local_nnz = 0;

    // host mirror of `state` (refreshed when ras_run polls / returns)
    double resnorm = -1.0, resnorm0 = -1.0, gres = 0.0, gres0 = -1.0;
    int32_t num_converged = 0, finished_iter = -1;
    bool finished = false;
    void conv_decide(int32_t protocol, double tol, int32_t check, int32_t iter, double *history_slot);
    struct Hub;                            // per-process decision hub (owned by the first subdomain)
    std::unique_ptr<Hub> hub;

    int32_t *conv() const { return (int32_t *)(mailbox + mbox.conv_off); }
    int32_t *err_word() const { return (int32_t *)(mailbox + mbox.err_off); }
    void conv_forward(int32_t converged_all_local);
    void conv_tree(int32_t converged_all_local);
    void conv_accumulate(int32_t converged_all_local);
    int32_t exchange_mode = EXCHANGE_PUT_GATHERED;

private:
    int32_t in_total_ = 0, out_total_ = 0;
    int32_t *in_dst_ = nullptr, *out_src_ = nullptr, *out_off_ = nullptr;
    // Get / one-by-one variants
    int32_t *in_off_ = nullptr;            // prefix sums of the in-list lengths
    int32_t *in_remote_idx_ = nullptr;     // position of every in-element inside its owner's x
    int32_t *out_remote_slot_ = nullptr;   // slot of every out-element inside the receiver's x
    std::vector<int32_t> in_off_host_, out_off_host_;
    std::vector<const void *> in_send_host_, in_x_host_;     // per in-neighbour
    std::vector<double *> out_x_host_;                       // per out-neighbour
    std::vector<void *> send_seg_host_;
    const void **in_send_dev_ = nullptr, **in_x_dev_ = nullptr;
    double **out_x_dev_ = nullptr;
    void **send_seg_dev_ = nullptr;
    size_t wire_size() const { return opt.use_mixed_precision ? sizeof(float) : sizeof(double); }
    std::vector<int32_t *> conv_peer_host_;                  // per subdomain id
    int32_t **conv_peer_dev_ = nullptr;
    std::vector<void *> out_dst_host_[2];
    std::vector<unsigned long long *> out_flag_host_;
    std::vector<int32_t *> out_conv_host_;
    std::vector<char> out_same_process_;
    void **out_dst_dev_[2] = {nullptr, nullptr};
    unsigned long long **out_flag_dev_ = nullptr;
    int32_t **out_conv_dev_ = nullptr;
    bool any_remote_ = false, guarded_ = false, onesided_ = false;
    bool sol_in_guess_ = false;   // the last local solve left its result in init_guess (:781)
    bool recv_dirty_ = false;   // reset_state ran: stale halo values may sit in the receive buffers
    int64_t push_epoch_ = 0, unpack_epoch_ = 0;
    const int32_t *stop_ptr() const { return guarded_ ? &state->stop : nullptr; }
    std::vector<int32_t> l2g_local_;     // global ids of [own | overlap]
    double *pinned_rhs_ = nullptr;
    void upload_peer_tables();
    bool peer_tables_dirty_ = true;
};

struct LoopOptions {
    int32_t num_subdomains = 1, max_iters = 100;
    int32_t chunk = 0;   // outer iterations enqueued between two host polls (0: default)
    double tolerance = 1e-6;
    int32_t enable_onesided = 0, enable_global_check = 0, conv_decentralized = 0, iter_offset = 0;
    // one-sided only: ExchangeMode, and 1 = centralised tree instead of flag flooding
    int32_t exchange_mode = 0, conv_tree = 0, conv_accumulate = 0;
    Comm *comm = nullptr;
};
struct LoopResult {
    int32_t iters = 0, converged = 0;
    double global_resnorm = 0.0, global_resnorm0 = -1.0, elapsed_s = 0.0;
    int32_t host_stream_syncs = 0, host_event_waits = 0;
};
void ras_run(std::vector<Ras *> &subs, const LoopOptions &opt, LoopResult &res,
             double *resnorm_history);
void ras_refresh_halo(std::vector<Ras *> &subs, int32_t P);

}  // namespace schwz_b200
