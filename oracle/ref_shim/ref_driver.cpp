// Driver of oracle/_ref: runs the reference's OWN SolverRAS (compiled unmodified from
// /root/reference/source, see oracle/Makefile) with every MPI rank as a thread of this
// process (ref_shim/mock_mpi.cpp) and the Ginkgo stand-in (ref_shim/ginkgo/ginkgo.hpp),
// and hands the index sets, matrices, exchange lists, residual histories and iterates
// back through a small C ABI for tests/ to compare with oracle/schwz_oracle.cpp and the
// CUDA path. It mirrors what benchmarking/bench_ras.cpp:47-169 does with its flags.
// TEST INFRASTRUCTURE; never part of the product.
#include <cstdint>
#include <chrono>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include <mpi.h>
#include <restricted_schwarz.hpp>

extern "C" {
struct ref_config {
    int32_t num_subdomains;
    int32_t laplacian_n;        // >0: --explicit_laplacian --set_1d_laplacian_size
    const char *matrix_file;    // used when laplacian_n == 0
    int32_t partition;          // Settings::partition_settings value (0, 1, 4)
    int32_t overlap;
    int32_t max_iters;          // --num_iters
    double tolerance;           // --set_tol
    double local_tol;           // --local_tol
    int32_t local_max_iters;    // --local_max_iters (-1 = local size)
    int32_t non_symmetric;      // GMRES instead of CG
    int32_t restart_iter;
    int32_t enable_onesided;
    int32_t remote_put;         // 1 = put, 0 = get
    int32_t one_by_one;
    int32_t conv_tree;          // 1 = centralized-tree, 0 = decentralized
    int32_t enable_global_check;
    int32_t decentralized_accumulate;
    int32_t put_all_local_residual_norms;
    int32_t use_mixed_precision;
    const char *local_precond;  // "null", "block-jacobi", ...
    int32_t precond_max_block_size;
    int32_t record_iterates;    // keep x (length N) after every exchange, per rank
    int32_t run;                // 0 = initialize() only
    const char *metis_objtype;
    int32_t enable_overlap;     // --enable_comm_overlap
    int32_t flush_local;
    int32_t lock_local;
};
}

namespace {

using VT = double;
using IT = gko::int32;

struct RankOut {
    int64_t sizes[8] = {0};   // global_size, local_size, local_size_x, overlap_size, nnz_local, nnz_interface, n_in, n_out
    std::vector<IT> first_row, permutation, i_permutation, l2g, g2l, overlap_row;
    std::vector<uint32_t> partition_indices;
    std::vector<IT> lrp, lci, irp, ici, grp, gci;
    std::vector<VT> lv, iv, gv;
    std::vector<IT> nbr_in, nbr_out, put_disp, get_disp;
    std::vector<std::vector<IT>> get_lists, put_lists;
    std::vector<VT> local_rhs;
    int32_t iter_count = 0;
    double run_seconds = 0.0;   // wall time of SolverRAS::run on this rank
    std::vector<VT> local_res, local_conv_res;
    std::vector<std::vector<VT>> global_res;
    std::vector<std::vector<VT>> iterates;
    std::vector<VT> solution;   // rank 0 only
    std::string error;
};

template <typename M>
void take_csr(const M &m, std::vector<IT> &rp, std::vector<IT> &ci, std::vector<VT> &v)
{
    const auto nr = m->get_size()[0];
    rp.assign(m->get_const_row_ptrs(), m->get_const_row_ptrs() + nr + 1);
    const auto nnz = static_cast<size_t>(rp.empty() ? 0 : rp[nr]);
    ci.assign(m->get_const_col_idxs(), m->get_const_col_idxs() + nnz);
    v.assign(m->get_const_values(), m->get_const_values() + nnz);
}

template <typename Mixed>
class Probe : public schwz::SolverRAS<VT, IT, Mixed> {
public:
    using Base = schwz::SolverRAS<VT, IT, Mixed>;
    Probe(schwz::Settings &s, schwz::Metadata<VT, IT> &m, RankOut &out, bool record)
        : Base(s, m), out_(out), record_(record)
    {}
    void update_boundary(const schwz::Settings &settings, const schwz::Metadata<VT, IT> &metadata,
                         std::shared_ptr<gko::matrix::Dense<VT>> &local_solution,
                         const std::shared_ptr<gko::matrix::Dense<VT>> &local_rhs,
                         const std::shared_ptr<gko::matrix::Dense<VT>> &global_solution,
                         const std::shared_ptr<gko::matrix::Csr<VT, IT>> &interface_matrix) override
    {
        if (record_)
            out_.iterates.emplace_back(global_solution->get_const_values(),
                                       global_solution->get_const_values() + metadata.global_size);
        Base::update_boundary(settings, metadata, local_solution, local_rhs, global_solution,
                              interface_matrix);
    }
    void harvest_setup()
    {
        auto &md = this->metadata;
        const auto N = md.global_size;
        const auto P = md.num_subdomains;
        out_.sizes[0] = N;
        out_.sizes[1] = md.local_size;
        out_.sizes[2] = md.local_size_x;
        out_.sizes[3] = md.overlap_size;
        out_.first_row.assign(md.first_row->get_data(), md.first_row->get_data() + P + 1);
        out_.permutation.assign(md.permutation->get_data(), md.permutation->get_data() + N);
        out_.i_permutation.assign(md.i_permutation->get_data(), md.i_permutation->get_data() + N);
        out_.l2g.assign(md.local_to_global->get_data(), md.local_to_global->get_data() + N);
        out_.g2l.assign(md.global_to_local->get_data(), md.global_to_local->get_data() + N);
        out_.overlap_row.assign(md.overlap_row.get_const_data(),
                                md.overlap_row.get_const_data() + md.overlap_size);
        out_.partition_indices = this->partition_indices;
        take_csr(this->local_matrix, out_.lrp, out_.lci, out_.lv);
        if (this->interface_matrix->get_size()[0] > 0)
            take_csr(this->interface_matrix, out_.irp, out_.ici, out_.iv);
        take_csr(this->global_matrix, out_.grp, out_.gci, out_.gv);
        out_.sizes[4] = out_.lci.size();
        out_.sizes[5] = out_.ici.size();
        auto &cs = this->comm_struct;
        out_.sizes[6] = cs.num_neighbors_in;
        out_.sizes[7] = cs.num_neighbors_out;
        for (int j = 0; j < cs.num_neighbors_in; ++j) {
            out_.nbr_in.push_back(cs.neighbors_in->get_data()[j]);
            IT *l = cs.global_get->get_data()[j];
            out_.get_lists.emplace_back(l + 1, l + 1 + l[0]);
        }
        for (int j = 0; j < cs.num_neighbors_out; ++j) {
            out_.nbr_out.push_back(cs.neighbors_out->get_data()[j]);
            IT *l = cs.global_put->get_data()[j];
            out_.put_lists.emplace_back(l + 1, l + 1 + l[0]);
        }
        out_.local_rhs.assign(this->local_rhs->get_const_values(),
                              this->local_rhs->get_const_values() + md.local_size_x);
    }
    void harvest_run()
    {
        auto &md = this->metadata;
        auto &cs = this->comm_struct;
        const auto P = md.num_subdomains;
        out_.put_disp.assign(cs.put_displacements->get_data(), cs.put_displacements->get_data() + P + 1);
        out_.get_disp.assign(cs.get_displacements->get_data(), cs.get_displacements->get_data() + P + 1);
        out_.iter_count = md.iter_count;
        out_.local_res = md.post_process_data.local_residual_vector_out;
        out_.local_conv_res = md.post_process_data.local_converged_resnorm;
        out_.global_res = md.post_process_data.global_residual_vector_out;
    }

private:
    RankOut &out_;
    bool record_;
};

struct Run {
    ref_config cfg;
    std::string matrix_file, precond, metis_objtype;
    std::vector<RankOut> ranks;
    std::string log;
};

template <typename Mixed>
void rank_body(int rank, Run &run)
{
    const ref_config &c = run.cfg;
    RankOut &out = run.ranks[rank];
    // benchmarking/bench_ras.cpp:49-148
    schwz::Metadata<VT, IT> metadata;
    schwz::Settings settings("reference");
    metadata.mpi_communicator = MPI_COMM_WORLD;
    MPI_Comm_rank(metadata.mpi_communicator, &metadata.my_rank);
    MPI_Comm_size(metadata.mpi_communicator, &metadata.comm_size);
    metadata.tolerance = c.tolerance;
    metadata.max_iters = c.max_iters;
    metadata.num_subdomains = metadata.comm_size;
    metadata.num_threads = 1;
    metadata.oned_laplacian_size = c.laplacian_n;
    settings.shifted_iter = 1;
    settings.comm_settings.enable_onesided = c.enable_onesided != 0;
    settings.comm_settings.enable_put = c.remote_put != 0;
    settings.comm_settings.enable_get = c.remote_put == 0;
    settings.comm_settings.enable_one_by_one = c.one_by_one != 0;
    settings.comm_settings.enable_overlap = c.enable_overlap != 0;
    if (c.flush_local) {
        settings.comm_settings.enable_flush_all = false;
        settings.comm_settings.enable_flush_local = true;
    }
    if (c.lock_local) {
        settings.comm_settings.enable_lock_all = false;
        settings.comm_settings.enable_lock_local = true;
    }
    settings.convergence_settings.put_all_local_residual_norms = c.put_all_local_residual_norms != 0;
    settings.convergence_settings.enable_global_check_iter_offset = false;
    settings.convergence_settings.enable_global_check = c.enable_global_check != 0;
    if (c.conv_tree) {
        settings.convergence_settings.enable_global_simple_tree = true;
    } else {
        settings.convergence_settings.enable_decentralized_leader_election = true;
        settings.convergence_settings.enable_accumulate = c.decentralized_accumulate != 0;
    }
    metadata.local_solver_tolerance = c.local_tol;
    metadata.local_precond = run.precond;
    metadata.local_max_iters = c.local_max_iters;
    metadata.updated_max_iters = -1;
    settings.non_symmetric_matrix = c.non_symmetric != 0;
    settings.restart_iter = c.restart_iter;
    settings.use_mixed_precision = c.use_mixed_precision != 0;
    metadata.precond_max_block_size = c.precond_max_block_size;
    settings.matrix_filename = c.laplacian_n > 0 ? std::string("null") : run.matrix_file;
    settings.explicit_laplacian = c.laplacian_n > 0;
    settings.enable_random_rhs = false;
    settings.overlap = c.overlap;
    settings.metis_objtype = run.metis_objtype;
    settings.partition = static_cast<schwz::Settings::partition_settings>(c.partition);
    settings.local_solver = schwz::Settings::local_solver_settings::iterative_solver_ginkgo;
    metadata.init_mpi_wtime = MPI_Wtime();

    std::shared_ptr<gko::matrix::Dense<VT>> solution;
    Probe<Mixed> solver(settings, metadata, out, c.record_iterates != 0);
    solver.initialize();
    solver.harvest_setup();
    if (c.run) {
        const auto t_run = std::chrono::steady_clock::now();
        solver.run(solution);
        out.run_seconds =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run).count();
        solver.harvest_run();
        if (rank == 0 && solution)
            out.solution.assign(solution->get_const_values(),
                                solution->get_const_values() + metadata.global_size);
    }
}

struct Thunk {
    Run *run;
};

void rank_entry(int rank, void *arg)
{
    Run &run = *static_cast<Thunk *>(arg)->run;
    if (run.cfg.use_mixed_precision)
        rank_body<float>(rank, run);
    else
        rank_body<double>(rank, run);
}

std::mutex g_run_mutex;   // the MPI mock holds process-wide state: one run at a time

// std::cout of the reference is captured; all rank threads print into it concurrently, so
// the buffer serialises them (a plain ostringstream would be a data race).
class LockedBuf : public std::streambuf {
public:
    std::string str()
    {
        std::lock_guard<std::mutex> lk(m_);
        return text_;
    }

protected:
    std::streamsize xsputn(const char *s, std::streamsize n) override
    {
        std::lock_guard<std::mutex> lk(m_);
        text_.append(s, static_cast<size_t>(n));
        return n;
    }
    int overflow(int c) override
    {
        if (c != EOF) {
            std::lock_guard<std::mutex> lk(m_);
            text_.push_back(static_cast<char>(c));
        }
        return c;
    }

private:
    std::mutex m_;
    std::string text_;
};

template <typename T>
int copy_out(const std::vector<T> &v, T *out, int64_t cap)
{
    if (out && cap >= static_cast<int64_t>(v.size()) && !v.empty())
        std::memcpy(out, v.data(), v.size() * sizeof(T));
    return static_cast<int>(v.size());
}

}  // namespace

extern "C" {

// Runs the reference on cfg->num_subdomains threads. Returns a handle (never null); check
// ref_error(). stdout of the reference is captured into ref_log().
void *ref_run(const ref_config *cfg)
{
    std::lock_guard<std::mutex> lk(g_run_mutex);
    auto *run = new Run();
    run->cfg = *cfg;
    run->matrix_file = cfg->matrix_file ? cfg->matrix_file : "null";
    run->precond = cfg->local_precond ? cfg->local_precond : "null";
    run->metis_objtype = cfg->metis_objtype ? cfg->metis_objtype : "null";
    run->ranks.resize(cfg->num_subdomains);
    LockedBuf sink;
    auto *old = std::cout.rdbuf(&sink);
    Thunk t{run};
    try {
        mockmpi::run(cfg->num_subdomains, rank_entry, &t);
    } catch (const std::exception &e) {
        run->ranks[0].error = e.what();
    } catch (...) {
        run->ranks[0].error = "unknown exception";
    }
    std::cout.rdbuf(old);
    run->log = sink.str();
    return run;
}
void ref_free(void *h) { delete static_cast<Run *>(h); }
const char *ref_error(void *h) { return static_cast<Run *>(h)->ranks[0].error.c_str(); }
const char *ref_log(void *h) { return static_cast<Run *>(h)->log.c_str(); }

void ref_sizes(void *h, int rank, int64_t *out8)
{
    std::memcpy(out8, static_cast<Run *>(h)->ranks[rank].sizes, 8 * sizeof(int64_t));
}
#define REF_VEC(name, field, T)                                         \
    int ref_##name(void *h, int rank, T *out, int64_t cap)              \
    {                                                                   \
        return copy_out(static_cast<Run *>(h)->ranks[rank].field, out, cap); \
    }
REF_VEC(first_row, first_row, int32_t)
REF_VEC(permutation, permutation, int32_t)
REF_VEC(i_permutation, i_permutation, int32_t)
REF_VEC(l2g, l2g, int32_t)
REF_VEC(g2l, g2l, int32_t)
REF_VEC(overlap_row, overlap_row, int32_t)
REF_VEC(partition_indices, partition_indices, uint32_t)
REF_VEC(local_rp, lrp, int32_t)
REF_VEC(local_ci, lci, int32_t)
REF_VEC(local_v, lv, double)
REF_VEC(interface_rp, irp, int32_t)
REF_VEC(interface_ci, ici, int32_t)
REF_VEC(interface_v, iv, double)
REF_VEC(global_rp, grp, int32_t)
REF_VEC(global_ci, gci, int32_t)
REF_VEC(global_v, gv, double)
REF_VEC(neighbors_in, nbr_in, int32_t)
REF_VEC(neighbors_out, nbr_out, int32_t)
REF_VEC(put_displacements, put_disp, int32_t)
REF_VEC(get_displacements, get_disp, int32_t)
REF_VEC(local_rhs, local_rhs, double)
REF_VEC(local_residuals, local_res, double)
REF_VEC(local_converged_resnorm, local_conv_res, double)
REF_VEC(solution, solution, double)

int ref_get_list(void *h, int rank, int j, int32_t *out, int64_t cap)
{
    return copy_out(static_cast<Run *>(h)->ranks[rank].get_lists[j], out, cap);
}
int ref_put_list(void *h, int rank, int j, int32_t *out, int64_t cap)
{
    return copy_out(static_cast<Run *>(h)->ranks[rank].put_lists[j], out, cap);
}
int ref_global_residuals(void *h, int rank, int j, double *out, int64_t cap)
{
    auto &g = static_cast<Run *>(h)->ranks[rank].global_res;
    if (j >= static_cast<int>(g.size())) return 0;
    return copy_out(g[j], out, cap);
}
double ref_run_seconds(void *h, int rank) { return static_cast<Run *>(h)->ranks[rank].run_seconds; }
int ref_iter_count(void *h, int rank) { return static_cast<Run *>(h)->ranks[rank].iter_count; }
int ref_num_iterates(void *h, int rank)
{
    return static_cast<int>(static_cast<Run *>(h)->ranks[rank].iterates.size());
}
int ref_iterate(void *h, int rank, int k, double *out, int64_t cap)
{
    return copy_out(static_cast<Run *>(h)->ranks[rank].iterates[k], out, cap);
}

// ---- kernel-level probes: the Ginkgo stand-in's local solvers and preconditioners, built
// with exactly the builder calls of source/solve.cpp:469-652, on a caller-supplied CSR ----
struct RefPrecond {
    std::shared_ptr<gko::Executor> exec;
    std::shared_ptr<gko::matrix::Csr<VT, IT>> A;
    std::shared_ptr<const gko::LinOp> op;
    std::string kind;
};

static std::shared_ptr<gko::matrix::Csr<VT, IT>> make_csr(std::shared_ptr<gko::Executor> exec,
                                                          int n, const int32_t *rp,
                                                          const int32_t *ci, const double *v)
{
    auto A = gko::matrix::Csr<VT, IT>::create(exec, gko::dim<2>(n, n), rp[n]);
    std::copy(rp, rp + n + 1, A->get_row_ptrs());
    std::copy(ci, ci + rp[n], A->get_col_idxs());
    std::copy(v, v + rp[n], A->get_values());
    return std::shared_ptr<gko::matrix::Csr<VT, IT>>(std::move(A));
}

void *ref_precond_create(int n, const int32_t *rp, const int32_t *ci, const double *v,
                         const char *kind, int max_block_size)
{
    auto *p = new RefPrecond();
    p->exec = gko::ReferenceExecutor::create();
    p->A = make_csr(p->exec, n, rp, ci, v);
    p->kind = kind;
    auto exec = p->exec;
    if (p->kind == "block-jacobi") {   // solve.cpp:496-505
        using bj = gko::preconditioner::Jacobi<VT, IT>;
        p->op = gko::share(bj::build().with_max_block_size(max_block_size).on(exec)->generate(p->A));
    } else if (p->kind == "ilu") {     // solve.cpp:513-526
        auto par_ilu = gko::factorization::ParIlu<VT, IT>::build().on(exec)->generate(p->A);
        auto f = gko::preconditioner::Ilu<gko::solver::LowerTrs<VT, IT>,
                                          gko::solver::UpperTrs<VT, IT>, false>::build()
                     .on(exec);
        p->op = gko::share(f->generate(gko::share(par_ilu)));
    } else if (p->kind == "isai") {    // solve.cpp:540-549
        using LowerIsai = gko::preconditioner::LowerIsai<VT, IT>;
        using UpperIsai = gko::preconditioner::UpperIsai<VT, IT>;
        auto f = gko::preconditioner::Ilu<LowerIsai, UpperIsai, false, IT>::build().on(exec);
        p->op = gko::share(f->generate(p->A));
    }
    return p;
}
void ref_precond_free(void *h) { delete static_cast<RefPrecond *>(h); }
void ref_precond_apply(void *h, const double *b, double *x)
{
    auto *p = static_cast<RefPrecond *>(h);
    const auto n = p->A->get_size()[0];
    auto bb = gko::matrix::Dense<VT>::create(p->exec, gko::dim<2>(n, 1));
    auto xx = gko::matrix::Dense<VT>::create(p->exec, gko::dim<2>(n, 1));
    std::copy(b, b + n, bb->get_values());
    p->op->apply(bb.get(), xx.get());
    std::copy(xx->get_const_values(), xx->get_const_values() + n, x);
}
int ref_precond_block_ptrs(void *h, int32_t *out, int64_t cap)
{
    auto *p = static_cast<RefPrecond *>(h);
    auto j = dynamic_cast<const gko::preconditioner::Jacobi<VT, IT> *>(p->op.get());
    if (!j) return -1;
    return copy_out(j->get_block_pointers(), out, cap);
}
int ref_precond_blocks(void *h, double *out, int64_t cap)
{
    auto *p = static_cast<RefPrecond *>(h);
    auto j = dynamic_cast<const gko::preconditioner::Jacobi<VT, IT> *>(p->op.get());
    if (!j) return -1;
    return copy_out(j->get_blocks(), out, cap);
}
// which: 0 L, 1 U (ILU factors); 2 approximate inverse of L, 3 of U (ISAI).  rp == NULL: nnz
int64_t ref_precond_csr(void *h, int which, int32_t *rp, int32_t *ci, double *v)
{
    auto *p = static_cast<RefPrecond *>(h);
    std::shared_ptr<const gko::matrix::Csr<VT, IT>> m;
    using TrsIlu = gko::preconditioner::Ilu<gko::solver::LowerTrs<VT, IT>,
                                            gko::solver::UpperTrs<VT, IT>, false>;
    using IsaiIlu = gko::preconditioner::Ilu<gko::preconditioner::LowerIsai<VT, IT>,
                                             gko::preconditioner::UpperIsai<VT, IT>, false, IT>;
    if (auto t = dynamic_cast<const TrsIlu *>(p->op.get())) {
        if (which == 0) m = t->get_l_solver()->get_system_matrix();
        if (which == 1) m = t->get_u_solver()->get_system_matrix();
    } else if (auto s = dynamic_cast<const IsaiIlu *>(p->op.get())) {
        if (which == 2) m = s->get_l_solver()->get_approximate_inverse();
        if (which == 3) m = s->get_u_solver()->get_approximate_inverse();
    }
    if (!m) return -1;
    const auto n = m->get_size()[0];
    const int64_t nnz = m->get_const_row_ptrs()[n];
    if (rp) {
        std::copy(m->get_const_row_ptrs(), m->get_const_row_ptrs() + n + 1, rp);
        std::copy(m->get_const_col_idxs(), m->get_const_col_idxs() + nnz, ci);
        std::copy(m->get_const_values(), m->get_const_values() + nnz, v);
    }
    return nnz;
}
// Cg / Gmres(restart) with Combined(Iteration(max_iters), ResidualNormReduction(tol)) and an
// optional preconditioner handle, x = warm start in, solution out (solve.cpp:469-478, 572-652)
void ref_krylov_solve(int n, const int32_t *rp, const int32_t *ci, const double *v,
                      const double *b, double *x, int gmres, int restart, int max_iters,
                      double tol, void *precond)
{
    auto exec = gko::ReferenceExecutor::create();
    auto A = make_csr(exec, n, rp, ci, v);
    auto crit = gko::share(
        gko::stop::Combined::build()
            .with_criteria(gko::stop::Iteration::build().with_max_iters(max_iters).on(exec),
                           gko::stop::ResidualNormReduction<VT>::build()
                               .with_reduction_factor(tol)
                               .on(exec))
            .on(exec));
    std::shared_ptr<const gko::LinOp> pre;
    if (precond) pre = static_cast<RefPrecond *>(precond)->op;
    std::shared_ptr<gko::LinOp> solver;
    if (gmres) {
        auto b_ = gko::solver::Gmres<VT>::build().with_criteria(crit).with_krylov_dim(restart);
        if (pre) b_.with_generated_preconditioner(pre);
        solver = gko::share(b_.on(exec)->generate(A));
    } else {
        auto b_ = gko::solver::Cg<VT>::build().with_criteria(crit);
        if (pre) b_.with_generated_preconditioner(pre);
        solver = gko::share(b_.on(exec)->generate(A));
    }
    auto bb = gko::matrix::Dense<VT>::create(exec, gko::dim<2>(n, 1));
    auto xx = gko::matrix::Dense<VT>::create(exec, gko::dim<2>(n, 1));
    std::copy(b, b + n, bb->get_values());
    std::copy(x, x + n, xx->get_values());
    solver->apply(bb.get(), xx.get());
    std::copy(xx->get_const_values(), xx->get_const_values() + n, x);
}
}
