// Bodies of the solver façade (schwz_classes.hpp).  Every stage of the outer
// loop is a call into the C ABI of libschwz_b200.so; ranks are host threads of
// this process (mpi_compat.hpp), subdomain s runs on GPU s % num_devices.
//
// Reference being mirrored: source/schwarz_base.cpp (ctor :74-124, initialize
// :128-271, run :323-506), source/restricted_schwarz.cpp, source/
// initialization.cpp, source/solve.cpp, source/communicate.cpp.
#include <type_traits>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <random>
#include <stdexcept>
#include <mutex>
#include <sstream>
#include <typeinfo>

#include "schwz_classes.hpp"

namespace {
// The C ABI speaks int32.  With IndexType = int32 these views are the caller's own pointers;
// with IndexType = int64 (the reference instantiates both, include/settings.hpp:533-537) values
// are narrowed on the way in (range-checked) and widened on the way out.
template <typename I>
struct Out32 {   // a buffer the library fills
    Out32(I *p, size_t n) : dst_(p), tmp_(std::is_same<I, int32_t>::value ? 0 : n) {}
    operator int32_t *() { return std::is_same<I, int32_t>::value ? (int32_t *)(void *)dst_ : tmp_.data(); }
    ~Out32()
    {
        for (size_t k = 0; k < tmp_.size(); ++k) dst_[k] = (I)tmp_[k];
    }
    I *dst_;
    std::vector<int32_t> tmp_;
};
template <typename I>
struct In32 {    // a buffer the library reads
    In32(const I *p, size_t n) : src_(p), tmp_(std::is_same<I, int32_t>::value ? 0 : n)
    {
        for (size_t k = 0; k < tmp_.size(); ++k) {
            if (p[k] > (I)2147483647 || p[k] < (I)-2147483647 - 1)
                throw std::runtime_error("index does not fit the 32-bit device index type");
            tmp_[k] = (int32_t)p[k];
        }
    }
    operator const int32_t *() const
    {
        return std::is_same<I, int32_t>::value ? (const int32_t *)(const void *)src_ : tmp_.data();
    }
    const I *src_;
    std::vector<int32_t> tmp_;
};
}  // namespace

#define B200_CHECK(expr) ::schwz::b200::check((expr), __FILE__, __LINE__)

namespace schwz {

namespace b200 {
// error convention of the reference: exceptions carrying "file:line: message"
// (include/exception.hpp:42-62); main() prints the banner and returns 1
void check(int rc, const char *file, int line)
{
    if (rc != 0)
        throw std::runtime_error(std::string(file) + ":" + std::to_string(line) + ": " +
                                 schwz_b200_last_error());
}
}  // namespace b200

using schwz_mpi::RankGroup;

// ranks are threads: print whole lines atomically
#define SAY(stream_expr)                                  \
    do {                                                  \
        std::ostringstream _os;                           \
        _os << stream_expr << "\n";                       \
        std::lock_guard<std::mutex> _lk(g_print_mutex);   \
        std::cout << _os.str() << std::flush;             \
    } while (0)
static std::mutex g_print_mutex;
enum Slot { kSlotSetup = 0, kSlotRas = 1, kSlotCtx = 2, kSlotSolution = 3 };

// =============================================================================
// Initialize
// =============================================================================
template <typename V, typename I>
Initialize<V, I>::Initialize(Settings &settings_, Metadata<V, I> &metadata_)
    : settings(settings_), metadata(metadata_)
{
    MPI_Comm_rank(metadata.mpi_communicator, &metadata.my_rank);
    MPI_Comm_size(metadata.mpi_communicator, &metadata.comm_size);
    metadata.num_subdomains = metadata.comm_size;   // source/initialization.cpp:72-74
}

template <typename V, typename I>
void Initialize<V, I>::generate_rhs(std::vector<V> &rhs)
{
    // source/initialization.cpp:89-96 (default-seeded engine, U(0,1))
    std::uniform_real_distribution<double> unif(0.0, 1.0);
    std::default_random_engine engine;
    for (auto &v : rhs) v = unif(engine);
}

// Generated problems keep only the shape on the host unless they are small:
// rows are produced on demand inside the index-set builder (no 4 GB replica
// of the global matrix per rank).
constexpr gko::size_type kMaterializeLimit = gko::size_type(1) << 22;

template <typename V, typename I>
void Initialize<V, I>::setup_global_matrix(const std::string &filename,
                                           const gko::size_type &n,
                                           std::shared_ptr<gko::matrix::Csr<V, I>> &global_matrix)
{
    using mtx = gko::matrix::Csr<V, I>;
    auto host = settings.executor->get_master();
    if (settings.matrix_filename != "null") {
        int32_t nrows = 0;
        int64_t nnz = 0;
        int32_t *rp = nullptr, *ci = nullptr;
        double *v = nullptr;
        B200_CHECK(schwz_b200_read_mtx(filename.c_str(), &nrows, &nnz, &rp, &ci, &v));
        global_matrix = mtx::create(host, gko::dim<2>(nrows), nnz);
        global_matrix->allocate();
        std::copy(rp, rp + nrows + 1, global_matrix->get_row_ptrs());
        std::copy(ci, ci + nnz, global_matrix->get_col_idxs());
        std::copy(v, v + nnz, global_matrix->get_values());
        schwz_b200_host_free(rp);
        schwz_b200_host_free(ci);
        schwz_b200_host_free(v);
        SAY("Matrix from file " << filename);
    } else if (settings.explicit_laplacian) {
        const bool three_d = settings.laplacian_dim == 3;
        SAY((three_d ? "Laplacian 3D Matrix (generated in house) "
                              : "Laplacian 2D Matrix (generated in house) "));
        const gko::size_type N = three_d ? n * n * n : n * n;
        const gko::size_type nnz = three_d ? 7 * N - 6 * n * n : 5 * N - 4 * n;
        global_matrix = mtx::create(host, gko::dim<2>(N), nnz);
        if (N <= kMaterializeLimit) {
            global_matrix->allocate();
            if (three_d)
                schwz_b200_laplacian3d((int32_t)n, Out32<I>(global_matrix->get_row_ptrs(), N + 1),
                                       Out32<I>(global_matrix->get_col_idxs(), nnz),
                                       global_matrix->get_values());
            else
                schwz_b200_laplacian2d((int32_t)n, Out32<I>(global_matrix->get_row_ptrs(), N + 1),
                                       Out32<I>(global_matrix->get_col_idxs(), nnz),
                                       global_matrix->get_values());
        }
    } else {
        std::cerr << " Need to provide a matrix or enable the default laplacian matrix." << std::endl;
        std::exit(-1);
    }
}

template <typename V, typename I>
void Initialize<V, I>::partition(const Settings &settings, const Metadata<V, I> &metadata,
                                 const std::shared_ptr<gko::matrix::Csr<V, I>> &global_matrix,
                                 std::vector<unsigned int> &partition_indices)
{
    // source/initialization.cpp:279-329: rank 0 computes, everybody gets a copy
    // (the copy is the shared index-set object here)
    if (metadata.my_rank != 0) return;
    partition_indices.assign(metadata.global_size, 0u);
    const auto kind = settings.partition;
    if (kind == Settings::partition_metis) {
        SAY(" METIS partition");
        if (!global_matrix->has_arrays())
            throw std::runtime_error("METIS partitioning needs the stored global matrix");
        B200_CHECK(schwz_b200_partition_metis(
            (int32_t)metadata.global_size,
            In32<I>(global_matrix->get_const_row_ptrs(), metadata.global_size + 1),
            In32<I>(global_matrix->get_const_col_idxs(), global_matrix->get_num_stored_elements()),
            (int32_t)metadata.num_subdomains, settings.metis_objtype.c_str(),
            partition_indices.data()));
    } else if (kind == Settings::partition_regular) {
        SAY(" Regular 1D partition");
    } else if (kind == Settings::partition_regular2d) {
        SAY(" Regular 2D partition");
        // perfect-square subdomain counts: the reference's rule, literally
        // (include/partition_tools.hpp:70-94).  Otherwise that rule leaves subdomains without
        // rows (its sqrt truncates), so the px x py extension takes over (8 -> 2 x 4).
        const int32_t Pn = (int32_t)metadata.num_subdomains;
        int32_t sq = (int32_t)std::sqrt((double)Pn);
        while (sq * sq > Pn) --sq;
        while ((sq + 1) * (sq + 1) <= Pn) ++sq;
        if (sq * sq == Pn) {
            B200_CHECK(schwz_b200_partition_regular2d((int64_t)metadata.global_size, Pn,
                                                      partition_indices.data()));
        } else {
            SAY(" (subdomain count is not a perfect square: px x py extension)");
            B200_CHECK(schwz_b200_partition_regular2d_rect((int64_t)metadata.global_size, Pn, 0,
                                                           0, partition_indices.data()));
        }
        if (settings.write_debug_out) {
            std::ofstream file("part_indices.csv");
            file << "idx,subd\n";
            for (size_t i = 0; i < partition_indices.size(); ++i)
                file << i << "," << partition_indices[i] << "\n";
        }
    } else {
        throw std::runtime_error("partition type not implemented");   // SCHWARZ_NOT_IMPLEMENTED
    }
}

template <typename V, typename I>
void Initialize<V, I>::setup_vectors(const Settings &settings, const Metadata<V, I> &metadata,
                                     std::vector<V> &rhs,
                                     std::shared_ptr<gko::matrix::Dense<V>> &local_rhs,
                                     std::shared_ptr<gko::matrix::Dense<V>> &global_rhs,
                                     std::shared_ptr<gko::matrix::Dense<V>> &local_solution)
{
    // source/initialization.cpp:333-359.  The device copies are made when the
    // subdomain object is created (setup_comm_buffers); here only the host-side
    // views the public members promise.
    using vec = gko::matrix::Dense<V>;
    auto host = settings.executor->get_master();
    local_rhs = vec::create(host, gko::dim<2>(metadata.local_size_x, 1));
    const I *l2g = metadata.local_to_global->get_const_data();
    for (gko::size_type k = 0; k < metadata.local_size_x; ++k) local_rhs->at(k) = rhs[l2g[k]];
    if (metadata.global_size <= kMaterializeLimit) {
        global_rhs = vec::create(host, gko::dim<2>(metadata.global_size, 1));
        std::copy(rhs.begin(), rhs.end(), global_rhs->get_values());
    }
    local_solution = vec::create(host, gko::dim<2>(metadata.local_size_x, 1));
}

// =============================================================================
// Communicate
// =============================================================================
template <typename V, typename I, typename M>
void Communicate<V, I, M>::local_to_global_vector(const Settings &settings, const Metadata<V, I> &,
                                                  const std::shared_ptr<gko::matrix::Dense<V>> &,
                                                  std::shared_ptr<gko::matrix::Dense<V>> &)
{
    // source/communicate.cpp:65-94.  Only the solution_based update exists here.  The
    // residual_based branch (:86-90) is unreachable from bench_ras and cannot run upstream
    // either (it add_scales a local_size_x vector with a local_size one); selecting it is an
    // error rather than a silent fall-back to solution_based.
    if (settings.convergence_settings.convergence_crit ==
        Settings::convergence_settings::local_convergence_crit::residual_based)
        throw std::runtime_error("convergence_settings.convergence_crit = residual_based is not "
                                 "implemented (source/communicate.cpp:86-90); use solution_based");
    B200_CHECK(schwz_b200_ras_restrict(comm_dev_->ras));
}

template <typename V, typename I, typename M>
void Communicate<V, I, M>::clear(Settings &)
{}

// =============================================================================
// Solve
// =============================================================================
// --write_perm_data (source/solve.cpp:434-453): one index per line, a blank line at the end
static void write_perm_files(int rank, const std::vector<int32_t> &perm,
                             const std::vector<int32_t> &inv_perm)
{
    for (int which = 0; which < 2; ++which) {
        std::ofstream file((which == 0 ? "perm_" : "inv_perm_") + std::to_string(rank) + ".csv");
        for (int32_t i : (which == 0 ? perm : inv_perm)) file << i << "\n";
        file << std::endl;
    }
}

template <typename V, typename I, typename M>
void Solve<V, I, M>::setup_local_solver(
    const Settings &settings, Metadata<V, I> &metadata,
    const std::shared_ptr<gko::matrix::Csr<V, I>> &local_matrix,
    std::shared_ptr<gko::matrix::Csr<V, I>> &triangular_factor_l,
    std::shared_ptr<gko::matrix::Csr<V, I>> &triangular_factor_u,
    std::shared_ptr<gko::matrix::Permutation<I>> &local_perm,
    std::shared_ptr<gko::matrix::Permutation<I>> &local_inv_perm,
    std::shared_ptr<gko::matrix::Dense<V>> &)
{
    // source/solve.cpp:197-663
    local_residual_vector.assign(std::max<size_t>(metadata.num_subdomains, metadata.max_iters + 1),
                                 DBL_MAX);
    metadata.post_process_data.global_residual_vector_out =
        std::vector<std::vector<V>>(metadata.num_subdomains);
    const auto solver = settings.local_solver;
    const bool direct = solver == Settings::direct_solver_ginkgo ||
                        solver == Settings::direct_solver_cholmod ||
                        solver == Settings::direct_solver_umfpack;
    // the unsymmetric factorisation: --local_solver=direct-umfpack, or direct-ginkgo with
    // --local_factorization=umfpack (source/solve.cpp:145-171, 322-385)
    const bool want_lu = solver == Settings::direct_solver_umfpack ||
                         (solver == Settings::direct_solver_ginkgo &&
                          settings.factorization == "umfpack");
    if (direct && want_lu) {
        b200::State &D = *solve_dev_;
        const int32_t n = (int32_t)metadata.local_size_x;
        std::vector<int32_t> rp(n + 1), ci(local_matrix->get_num_stored_elements());
        std::vector<double> v(ci.size());
        B200_CHECK(schwz_b200_setup_local_matrix(D.setup, metadata.my_rank, rp.data(), ci.data(),
                                                 v.data()));
        D.factor_perm.resize(n);
        if (settings.naturally_ordered_factor)
            std::iota(D.factor_perm.begin(), D.factor_perm.end(), 0);
        else
            B200_CHECK(schwz_b200_host_nd_ordering(n, rp.data(), ci.data(), D.factor_perm.data()));
        B200_CHECK(schwz_b200_host_lu_create(n, rp.data(), ci.data(), v.data(),
                                             D.factor_perm.data(), 1e-3, &D.lu));
        int64_t lnnz = 0, unnz = 0;
        B200_CHECK(schwz_b200_host_lu_nnz(D.lu, &lnnz, &unnz));
        if (metadata.my_rank == 0)
            SAY(" Local direct factorization (sparse LU, threshold partial pivoting, "
                << (settings.naturally_ordered_factor ? "natural" : "nested dissection")
                << " column ordering)");
        SAY(" Process " << metadata.my_rank << " has factors with " << n << " rows and " << lnnz
                        << " + " << unnz << " non-zeros ");
        auto host = settings.executor->get_master();
        triangular_factor_l = gko::matrix::Csr<V, I>::create(host, gko::dim<2>(n), lnnz);
        triangular_factor_u = gko::matrix::Csr<V, I>::create(host, gko::dim<2>(n), unnz);
        std::vector<int32_t> rowp(n), tmp_rp(n + 1), tmp_ci(std::max(lnnz, unnz));
        std::vector<int32_t> tmp_rp2(n + 1), tmp_ci2(std::max(lnnz, unnz));
        std::vector<double> tmp_v(std::max(lnnz, unnz)), tmp_v2(std::max(lnnz, unnz));
        B200_CHECK(schwz_b200_host_lu_get(D.lu, tmp_rp.data(), tmp_ci.data(), tmp_v.data(),
                                          tmp_rp2.data(), tmp_ci2.data(), tmp_v2.data(),
                                          rowp.data()));
        local_perm = gko::matrix::Permutation<I>::create(host, std::vector<I>(rowp.begin(), rowp.end()));
        local_inv_perm = gko::matrix::Permutation<I>::create(
            host, std::vector<I>(D.factor_perm.begin(), D.factor_perm.end()));
        if (settings.write_perm_data) write_perm_files(metadata.my_rank, rowp, D.factor_perm);
        if (metadata.my_rank == 0)
            SAY(" Local direct solve with level-scheduled TRS");
    } else if (direct) {
        // factorisation on the host (replaces cholmod_analyze / cholmod_factorize,
        // solve.cpp:94-142); fill-reducing order unless --factor_ordering_natural
        b200::State &D = *solve_dev_;
        const int32_t n = (int32_t)metadata.local_size_x;
        std::vector<int32_t> rp(n + 1), ci(local_matrix->get_num_stored_elements());
        std::vector<double> v(ci.size());
        B200_CHECK(schwz_b200_setup_local_matrix(D.setup, metadata.my_rank, rp.data(), ci.data(),
                                                 v.data()));
        D.factor_perm.resize(n);
        if (settings.naturally_ordered_factor)
            std::iota(D.factor_perm.begin(), D.factor_perm.end(), 0);
        else
            B200_CHECK(schwz_b200_host_nd_ordering(n, rp.data(), ci.data(), D.factor_perm.data()));
        const int64_t lnnz = schwz_b200_host_cholesky(n, rp.data(), ci.data(), v.data(),
                                                      D.factor_perm.data(), nullptr, nullptr, nullptr);
        if (lnnz < 0) throw std::runtime_error(schwz_b200_last_error());
        D.L_rowptr.resize(n + 1);
        D.L_col.resize(lnnz);
        D.L_val.resize(lnnz);
        schwz_b200_host_cholesky(n, rp.data(), ci.data(), v.data(), D.factor_perm.data(),
                                 D.L_rowptr.data(), D.L_col.data(), D.L_val.data());
        D.have_factors = true;
        if (metadata.my_rank == 0)
            SAY(" Local direct factorization (simplicial LL^T, "
                      << (settings.naturally_ordered_factor ? "natural" : "nested dissection")
                      << " ordering)");
        SAY(" Process " << metadata.my_rank << " has factor with " << n << " rows and "
                  << lnnz << " non-zeros ");
        auto host = settings.executor->get_master();
        triangular_factor_l = gko::matrix::Csr<V, I>::create(host, gko::dim<2>(n), lnnz);
        triangular_factor_u = gko::matrix::Csr<V, I>::create(host, gko::dim<2>(n), lnnz);
        local_perm = gko::matrix::Permutation<I>::create(
            host, std::vector<I>(D.factor_perm.begin(), D.factor_perm.end()));
        local_inv_perm = local_perm;
        if (settings.write_perm_data)
            write_perm_files(metadata.my_rank, D.factor_perm, D.factor_perm);
        if (metadata.my_rank == 0)
            SAY(" Local direct solve with level-scheduled TRS");
    } else if (solver == Settings::iterative_solver_ginkgo) {
        const int l_max_iters = metadata.local_max_iters == -1 ? (int)local_matrix->get_size()[0]
                                                                : (int)metadata.local_max_iters;
        if (metadata.my_rank == 0) {
            SAY(" Local max iters " << l_max_iters << " with restart iter "
                      << settings.restart_iter);
            // the banners of source/solve.cpp:489-651, verbatim
            const std::string kry = settings.non_symmetric_matrix ? "GMRES" : "CG";
            const std::string &pc = metadata.local_precond;
            if (pc == "block-jacobi")
                SAY(" Local Ginkgo iterative solve(" << kry << ") with Block-Jacobi preconditioning ");
            else if (pc == "ilu")
                SAY(" Local Ginkgo iterative solve(" << kry << ") with ParILU preconditioning ");
            else if (pc == "isai")
                SAY(" Local Ginkgo iterative solve(" << kry << ") with ISAIpreconditioning ");
            else if (pc == "null")
                SAY(" Local Ginkgo iterative solve(" << kry << ") with no preconditioning ");
            else
                std::cerr << "Unsupported preconditioner." << std::endl;
        }
    } else {
        throw std::runtime_error("local solver not implemented");
    }
}

template <typename V, typename I, typename M>
void Solve<V, I, M>::local_solve(const Settings &settings, Metadata<V, I> &metadata,
                                 const std::shared_ptr<gko::matrix::Csr<V, I>> &,
                                 const std::shared_ptr<gko::matrix::Csr<V, I>> &,
                                 const std::shared_ptr<gko::matrix::Csr<V, I>> &,
                                 std::shared_ptr<gko::matrix::Permutation<I>> &,
                                 std::shared_ptr<gko::matrix::Permutation<I>> &,
                                 std::shared_ptr<gko::matrix::Dense<V>> &,
                                 std::shared_ptr<gko::matrix::Dense<V>> &,
                                 std::shared_ptr<gko::matrix::Dense<V>> &)
{
    // source/solve.cpp:667-792 — asynchronous on the subdomain's stream
    if (settings.local_solver == Settings::iterative_solver_ginkgo &&
        settings.reset_local_crit_iter != -1 &&
        (int)metadata.iter_count > settings.reset_local_crit_iter) {
        // :721-741: past that outer iteration the local iteration cap becomes
        // metadata.updated_max_iters (-1: the local size)
        B200_CHECK(schwz_b200_ras_set_local_max_iters(solve_dev_->ras,
                                                      (int32_t)metadata.updated_max_iters));
    }
    B200_CHECK(schwz_b200_ras_local_solve(solve_dev_->ras));
    metadata.post_process_data.local_converged_iter_count.push_back(0);
    metadata.post_process_data.local_timestamp.push_back(MPI_Wtime() - metadata.init_mpi_wtime);
}

template <typename V, typename I, typename M>
bool Solve<V, I, M>::check_local_convergence(const Settings &, Metadata<V, I> &metadata,
                                             V &local_resnorm, V &local_resnorm0)
{
    // source/solve.cpp:796-856
    bool locally_converged = false;
    local_resnorm = -1.0;
    const V tolerance = metadata.tolerance;
    if (tolerance >= 0.0) {
        B200_CHECK(schwz_b200_ras_local_residual(solve_dev_->ras));
        double nrm = 0.0;
        B200_CHECK(schwz_b200_ras_residual_norm(solve_dev_->ras, &nrm));   // the D2H of :841-843
        local_resnorm = nrm;
        if (local_resnorm0 < 0.0) local_resnorm0 = local_resnorm;
        locally_converged = (local_resnorm * local_resnorm) / (local_resnorm0 * local_resnorm0) <
                            (tolerance * tolerance);
    }
    metadata.post_process_data.local_converged_resnorm.push_back(local_resnorm / local_resnorm0);
    return locally_converged;
}

template <typename V, typename I, typename M>
void Solve<V, I, M>::check_global_convergence(
    const Settings &settings, Metadata<V, I> &metadata,
    struct Communicate<V, I, M>::comm_struct &, V &local_resnorm, V &local_resnorm0,
    V &global_resnorm, V &global_resnorm0, int &converged_all_local, int &num_converged_procs)
{
    // source/solve.cpp:860-955
    const int P = (int)metadata.num_subdomains;
    const int me = metadata.my_rank;
    const V tolerance = metadata.tolerance;
    auto &l_res = local_residual_vector;
    RankGroup &G = RankGroup::instance();
    if (settings.convergence_settings.enable_global_check && !settings.comm_settings.enable_onesided) {
        // MPI_Allgather of P doubles == a shared table + two barriers
        G.doubles()[me] = local_resnorm;
        G.barrier();
        for (int j = 0; j < P; ++j) l_res[j] = G.doubles()[j];
        G.barrier();
        global_resnorm = 0.0;
        for (int j = 0; j < P; ++j) {
            metadata.post_process_data.global_residual_vector_out[j].push_back(l_res[j]);
            if (l_res[j] != DBL_MAX) {
                global_resnorm += l_res[j];
            } else {
                global_resnorm = -1.0;
                break;
            }
        }
        if (global_resnorm >= 0.0) {
            if (global_resnorm0 < 0.0) global_resnorm0 = global_resnorm;
            if (global_resnorm / global_resnorm0 <= tolerance) converged_all_local++;
        }
    } else if (settings.comm_settings.enable_onesided) {
        if (local_resnorm / local_resnorm0 <= tolerance) converged_all_local++;
        l_res[me] = std::min(l_res[me], local_resnorm);
        for (int j = 0; j < P; ++j)
            metadata.post_process_data.global_residual_vector_out[j].push_back(l_res[j]);
    }
    if (settings.comm_settings.enable_onesided) {
        if (settings.convergence_settings.enable_decentralized_leader_election ||
            settings.convergence_settings.enable_global_simple_tree) {
            // flags live in the peer-visible mailboxes: flooding along the halo graph
            // (conv_tools.hpp:248-274) or the binary tree (conv_tools.hpp:147-209)
            int32_t n = 0;
            if (settings.convergence_settings.enable_global_simple_tree)
                B200_CHECK(schwz_b200_ras_conv_tree(solve_dev_->ras, converged_all_local));
            else if (settings.convergence_settings.enable_accumulate)   // conv_tools.hpp:230-247
                B200_CHECK(schwz_b200_ras_conv_accumulate(solve_dev_->ras, converged_all_local));
            else
                B200_CHECK(schwz_b200_ras_conv_set_local(solve_dev_->ras, converged_all_local));
            B200_CHECK(schwz_b200_ras_conv_count(solve_dev_->ras, &n));
            num_converged_procs = n;
        } else {
            SAY("Global Convergence check type unspecified");
            std::exit(-1);
        }
    } else {
        if (settings.convergence_settings.enable_global_check) {
            if (converged_all_local == 1) num_converged_procs = P;
        } else {
            num_converged_procs = 0;   // Allreduce of a count that is never incremented (F9)
        }
    }
}

template <typename V, typename I, typename M>
void Solve<V, I, M>::check_convergence(
    const Settings &settings, Metadata<V, I> &metadata,
    struct Communicate<V, I, M>::comm_struct &comm_struct, V &local_residual_norm,
    V &local_residual_norm0, V &global_residual_norm, V &global_residual_norm0,
    int &num_converged_procs)
{
    // source/solve.cpp:959-1005
    int num_converged_p =
        check_local_convergence(settings, metadata, local_residual_norm, local_residual_norm0) ? 1 : 0;
    if (std::isnan(local_residual_norm)) std::exit(-1);
    metadata.post_process_data.local_residual_vector_out.push_back(local_residual_norm);
    metadata.current_residual_norm = local_residual_norm;
    const auto iter = metadata.iter_count;
    metadata.min_residual_norm =
        (iter == 0 ? local_residual_norm : std::min(local_residual_norm, metadata.min_residual_norm));
    const bool iter_cond = settings.convergence_settings.enable_global_check_iter_offset
                               ? ((iter > (metadata.max_iters * 0.05)) || metadata.max_iters < 1000)
                               : true;
    if (metadata.tolerance > 0.0 && iter_cond) {
        int converged_all_local = 0;
        check_global_convergence(settings, metadata, comm_struct, local_residual_norm,
                                 local_residual_norm0, global_residual_norm, global_residual_norm0,
                                 converged_all_local, num_converged_p);
        num_converged_procs = num_converged_p;
    }
}

template <typename V, typename I, typename M>
void Solve<V, I, M>::compute_residual_norm(const Settings &, const Metadata<V, I> &metadata,
                                           V &mat_norm, V &rhs_norm, V &sol_norm, V &residual_norm)
{
    // source/solve.cpp:1025-1085 does an N-double Allreduce and a replicated
    // global SpMV; here every subdomain computes ||b_own - (A x)_own||^2 on its
    // own rows after one more exchange and the squares are summed.
    RankGroup &G = RankGroup::instance();
    const int me = metadata.my_rank, P = (int)metadata.num_subdomains;
    schwz_ras *ras = solve_dev_->ras;
    B200_CHECK(schwz_b200_ras_exchange_push(ras, 0));
    B200_CHECK(schwz_b200_ras_sync(ras));
    G.barrier();
    B200_CHECK(schwz_b200_ras_exchange_unpack(ras, 0, 0));
    double rsq = 0.0;
    B200_CHECK(schwz_b200_ras_true_residual_sq(ras, &rsq));
    G.doubles()[me] = rsq;
    G.barrier();
    double tot = 0.0;
    for (int j = 0; j < P; ++j) tot += G.doubles()[j];
    G.barrier();
    residual_norm = std::sqrt(tot);
    mat_norm = -1.0;
    sol_norm = -1.0;
    (void)rhs_norm;
}

template <typename V, typename I, typename M>
void Solve<V, I, M>::clear(Settings &)
{}

// =============================================================================
// SchwarzBase
// =============================================================================
template <typename V, typename I, typename M>
SchwarzBase<V, I, M>::SchwarzBase(Settings &settings_, Metadata<V, I> &metadata_)
    : Initialize<V, I>(settings_, metadata_), settings(settings_), metadata(metadata_)
{
    // source/schwarz_base.cpp:74-124.  The node-local rank of the reference is
    // the rank itself (one process); the device is rank % num_devices, which
    // lifts the one-rank-per-GPU limit.
    this->comm_dev_ = &dev_;
    this->solve_dev_ = &dev_;
    metadata.my_local_rank = metadata.my_rank;
    metadata.local_num_procs = metadata.comm_size;
    if (settings.executor_string == "cuda") {
        int num_devices = 0;
        B200_CHECK(schwz_b200_device_count(&num_devices));
        if (num_devices < 1) {
            std::cerr << " No CUDA devices available for rank " << metadata.my_rank << std::endl;
            std::exit(-1);   // source/utils.cpp:164-168
        }
        const int use = settings.num_devices > 0 ? std::min(settings.num_devices, num_devices)
                                                  : num_devices;
        dev_.device = metadata.my_rank % use;
        settings.executor =
            gko::CudaExecutor::create(dev_.device, gko::OmpExecutor::create(), false);
        SAY(" Rank " << metadata.my_rank << " with local rank " << metadata.my_local_rank
                  << " has " << dev_.device << " id of gpu");
        MPI_Barrier(metadata.mpi_communicator);
    } else {
        // --executor=omp|reference select the CPU executors of Ginkgo upstream;
        // this library has no CPU compute path
        throw std::runtime_error("executor '" + settings.executor_string +
                                 "' is not available: schwz-b200 runs the RAS path on CUDA "
                                 "devices only (use --executor=cuda)");
    }
    B200_CHECK(schwz_b200_ctx_create(dev_.device, &dev_.ctx));
}

template <typename V, typename I, typename M>
SchwarzBase<V, I, M>::~SchwarzBase()
{
    if (dev_.ras) schwz_b200_ras_destroy(dev_.ras);
    if (dev_.ctx) schwz_b200_ctx_destroy(dev_.ctx);
    if (dev_.setup && metadata.my_rank == 0) schwz_b200_setup_destroy(dev_.setup);
}

template <typename V, typename I, typename M>
void SchwarzBase<V, I, M>::initialize()
{
    // source/schwarz_base.cpp:128-271
    using vec_itype = gko::Array<I>;
    auto host = settings.executor->get_master();
    if (settings.explicit_laplacian || settings.matrix_filename != "null") {
        Initialize<V, I>::setup_global_matrix(settings.matrix_filename, metadata.oned_laplacian_size,
                                              this->global_matrix);
    } else {
        std::cerr << " Explicit laplacian needs to be enabled with the --explicit_laplacian flag or "
                     "deal.ii support needs to be enabled to generate the matrices"
                  << std::endl;
        std::exit(-1);
    }
    metadata.global_size = this->global_matrix->get_size()[0];
    const auto P = metadata.num_subdomains;

    rhs_host_.assign(metadata.global_size, 1.0);
    if (settings.enable_random_rhs && settings.explicit_laplacian)
        Initialize<V, I>::generate_rhs(rhs_host_);   // same default-seeded stream on every rank

    metadata.first_row = std::make_shared<vec_itype>(host, P + 1);
    auto &cs = this->comm_struct;
    cs.neighbors_in = std::make_shared<vec_itype>(host, P + 1);
    cs.neighbors_out = std::make_shared<vec_itype>(host, P + 1);
    cs.local_neighbors_in = std::make_shared<vec_itype>(host, P + 1);
    cs.local_neighbors_out = std::make_shared<vec_itype>(host, P + 1);
    cs.is_local_neighbor = std::vector<bool>(P + 1, false);
    cs.global_put = std::make_shared<gko::Array<I *>>(host, P + 1);
    cs.local_put = std::make_shared<gko::Array<I *>>(host, P + 1);
    cs.global_get = std::make_shared<gko::Array<I *>>(host, P + 1);
    cs.local_get = std::make_shared<gko::Array<I *>>(host, P + 1);
    std::vector<I> zeros(P + 1, 0);
    cs.get_displacements = std::make_shared<vec_itype>(host, zeros.begin(), zeros.end());
    cs.put_displacements = std::make_shared<vec_itype>(host, zeros.begin(), zeros.end());

    Initialize<V, I>::partition(settings, metadata, this->global_matrix, this->partition_indices);
    this->setup_local_matrices(settings, metadata, this->partition_indices, this->global_matrix,
                               this->local_matrix, this->interface_matrix);
    SAY("Subdomain " << metadata.my_rank << " has local problem size "
              << this->local_matrix->get_size()[0] << " with "
              << this->local_matrix->get_num_stored_elements() << " non-zeros ");
    Initialize<V, I>::setup_vectors(settings, metadata, rhs_host_, this->local_rhs, this->global_rhs,
                                    this->local_solution);
    Solve<V, I, M>::setup_local_solver(settings, metadata, this->local_matrix,
                                       this->triangular_factor_l, this->triangular_factor_u,
                                       this->local_perm, this->local_inv_perm, this->local_rhs);
    this->setup_comm_buffers();
}

template <typename V, typename I>
static void write_iters_and_residuals(int iter_count, std::vector<V> &res,
                                      std::vector<V> &local_iters, std::vector<V> &local_res,
                                      std::vector<V> &stamp, const std::string &filename)
{
    // source/schwarz_base.cpp:51-70
    std::ofstream file(filename);
    file << "iter,resnorm,localiter,localresnorm,timestamp\n";
    for (int i = 0; i < iter_count; ++i)
        file << i << "," << res[i] << "," << (i < (int)local_iters.size() ? local_iters[i] : 0)
             << "," << (i < (int)local_res.size() ? local_res[i] : 0) << ","
             << (i < (int)stamp.size() ? stamp[i] : 0) << "\n";
}

template <typename V, typename I, typename M>
void SchwarzBase<V, I, M>::run(std::shared_ptr<gko::matrix::Dense<V>> &solution)
{
    // source/schwarz_base.cpp:323-506
    using vec_vtype = gko::matrix::Dense<V>;
    using ValueType = V;   // MEASURE_ELAPSED_FUNC_TIME names it
    RankGroup &G = RankGroup::instance();
    auto host = settings.executor->get_master();
    if (!solution.get()) solution = vec_vtype::create(host, gko::dim<2>(metadata.global_size, 1));
    if (metadata.my_rank == 0) {
        M dummy1 = 0.0;
        V dummy2 = 1.0;
        SAY(" MixedValueType: " << typeid(dummy1).name()
                  << " ValueType: " << typeid(dummy2).name());
    }
    // the device holds x (zero-initialised, F8), work vectors and init_guess;
    // these host handles exist for signature compatibility
    std::shared_ptr<vec_vtype> global_solution = vec_vtype::create(host, gko::dim<2>(0, 1));
    std::shared_ptr<vec_vtype> work_vector = vec_vtype::create(host, gko::dim<2>(0, 1));
    std::shared_ptr<vec_vtype> init_guess = vec_vtype::create(host, gko::dim<2>(0, 1));

    this->setup_windows(settings, metadata, global_solution);

    V local_residual_norm = -1.0, local_residual_norm0 = -1.0, global_residual_norm = 0.0,
      global_residual_norm0 = -1.0;
    metadata.iter_count = 0;
    B200_CHECK(schwz_b200_ras_sync(dev_.ras));
    MPI_Barrier(MPI_COMM_WORLD);
    auto start_time = std::chrono::steady_clock::now();
    int num_converged_procs = 0;

    for (; metadata.iter_count < metadata.max_iters; ++(metadata.iter_count)) {
        MEASURE_ELAPSED_FUNC_TIME(this->exchange_boundary(settings, metadata, global_solution), 0,
                                  metadata.my_rank, boundary_exchange, metadata.iter_count);
        MEASURE_ELAPSED_FUNC_TIME(
            this->update_boundary(settings, metadata, this->local_solution, this->local_rhs,
                                  global_solution, this->interface_matrix),
            1, metadata.my_rank, boundary_update, metadata.iter_count);
        MEASURE_ELAPSED_FUNC_TIME(
            (Solve<V, I, M>::check_convergence(settings, metadata, this->comm_struct,
                                               local_residual_norm, local_residual_norm0,
                                               global_residual_norm, global_residual_norm0,
                                               num_converged_procs)),
            2, metadata.my_rank, convergence_check, metadata.iter_count);
        if (std::isnan(global_residual_norm) || global_residual_norm > 1e12) {
            SAY(" Rank " << metadata.my_rank << " diverged in " << metadata.iter_count
                      << " iters ");
            std::exit(-1);
        }
        if (num_converged_procs == (int)metadata.num_subdomains) {
            break;
        } else {
            MEASURE_ELAPSED_FUNC_TIME(
                (Solve<V, I, M>::local_solve(settings, metadata, this->local_matrix,
                                             this->triangular_factor_l, this->triangular_factor_u,
                                             this->local_perm, this->local_inv_perm, work_vector,
                                             init_guess, this->local_solution)),
                3, metadata.my_rank, local_solve, metadata.iter_count);
            MEASURE_ELAPSED_FUNC_TIME(
                (Communicate<V, I, M>::local_to_global_vector(settings, metadata,
                                                              this->local_solution, global_solution)),
                4, metadata.my_rank, expand_local_vec, metadata.iter_count);
        }
    }
    B200_CHECK(schwz_b200_ras_sync(dev_.ras));
    MPI_Barrier(MPI_COMM_WORLD);
    auto elapsed_time = std::chrono::duration<V>(std::chrono::steady_clock::now() - start_time);

    if (settings.write_iters_and_residuals &&
        settings.local_solver == Settings::iterative_solver_ginkgo) {
        std::string rank_string = std::to_string(metadata.my_rank);
        if (metadata.my_rank < 10) rank_string = "0" + rank_string;
        auto &pp = metadata.post_process_data;
        write_iters_and_residuals<V, I>((int)pp.local_residual_vector_out.size(),
                                        pp.local_residual_vector_out, pp.local_converged_iter_count,
                                        pp.local_converged_resnorm, pp.local_timestamp,
                                        "iter_res_" + rank_string + ".csv");
    }
    // every rank must take the same branch below (it contains barriers): decide
    // on the minimum over the ranks
    G.doubles()[metadata.my_rank] = num_converged_procs;
    G.barrier();
    bool all_converged = true;
    for (gko::size_type j = 0; j < metadata.num_subdomains; ++j)
        all_converged &= ((int)G.doubles()[j] == (int)metadata.num_subdomains);
    G.barrier();
    if (num_converged_procs < (int)metadata.num_subdomains) {
        SAY("Rank " << metadata.my_rank << " did not converge in " << metadata.iter_count
                  << " iterations.");
    } else {
        SAY(" Rank " << metadata.my_rank << " converged in " << metadata.iter_count
                  << " iterations ");
    }
    if (all_converged) {
        V mat_norm = -1.0, rhs_norm = -1.0, sol_norm = -1.0, residual_norm = -1.0;
        rhs_norm = std::sqrt(std::inner_product(rhs_host_.begin(), rhs_host_.end(),
                                                rhs_host_.begin(), V(0)));
        Solve<V, I, M>::compute_residual_norm(settings, metadata, mat_norm, rhs_norm, sol_norm,
                                              residual_norm);
        // gather_comm_data, source/schwarz_base.cpp:275-319
        auto &cs = this->comm_struct;
        for (gko::size_type i = 0; i < metadata.num_subdomains; ++i) {
            std::vector<int> cout_(metadata.num_subdomains, 0), cin_(metadata.num_subdomains, 0);
            std::vector<std::tuple<int, int>> send_tuple, recv_tuple;
            for (int j = 0; j < cs.num_neighbors_out; ++j) {
                send_tuple.emplace_back(cs.neighbors_out->get_data()[j], cs.global_put->get_data()[j][0]);
                cout_[cs.neighbors_out->get_data()[j]] = 1;
            }
            for (int j = 0; j < cs.num_neighbors_in; ++j) {
                recv_tuple.emplace_back(cs.neighbors_in->get_data()[j], cs.global_get->get_data()[j][0]);
                cin_[cs.neighbors_in->get_data()[j]] = 1;
            }
            for (gko::size_type j = 0; j < metadata.num_subdomains; ++j) {
                if (cout_[j] == 0) send_tuple.emplace_back((int)j, 0);
                if (cin_[j] == 0) recv_tuple.emplace_back((int)j, 0);
            }
            metadata.comm_data_struct.emplace_back((int)i, recv_tuple, send_tuple,
                                                   cs.num_neighbors_in, cs.num_neighbors_out);
        }
        if (metadata.my_rank == 0) {
            SAY(" residual norm " << residual_norm << "\n"
                      << " relative residual norm of solution " << residual_norm / rhs_norm << "\n"
                      << " Time taken for solve " << elapsed_time.count());
        }
    }
    // the reference copies its (allreduced) vector on rank 0; here every rank
    // writes its own block into rank 0's vector
    if (metadata.my_rank == 0) G.slots(kSlotSolution)[0] = solution->get_values();
    G.barrier();
    double *dst = (double *)G.slots(kSlotSolution)[0];
    B200_CHECK(schwz_b200_ras_download_solution(dev_.ras, dst));
    B200_CHECK(schwz_b200_ras_sync(dev_.ras));
    G.barrier();
}

// =============================================================================
// SolverRAS
// =============================================================================
template <typename V, typename I, typename M>
SolverRAS<V, I, M>::SolverRAS(Settings &settings, Metadata<V, I> &metadata)
    : SchwarzBase<V, I, M>(settings, metadata)
{}

template <typename V, typename I, typename M>
void SolverRAS<V, I, M>::setup_local_matrices(
    Settings &settings, Metadata<V, I> &metadata, std::vector<unsigned int> &partition_indices,
    std::shared_ptr<gko::matrix::Csr<V, I>> &global_matrix,
    std::shared_ptr<gko::matrix::Csr<V, I>> &local_matrix,
    std::shared_ptr<gko::matrix::Csr<V, I>> &interface_matrix)
{
    // source/restricted_schwarz.cpp:56-304.  Rank 0 builds the index sets of all
    // subdomains once (the MPI_Bcast of the partition vector at :73 becomes a
    // shared object); every rank then reads its own part.
    RankGroup &G = RankGroup::instance();
    b200::State &D = this->dev_;
    const int me = metadata.my_rank;
    const int P = (int)metadata.num_subdomains;
    const bool permute = settings.partition == Settings::partition_metis ||
                         settings.partition == Settings::partition_regular2d;
    if (me == 0) {
        schwz_setup *s = nullptr;
        const uint32_t *part = permute ? partition_indices.data() : nullptr;
        if (global_matrix->has_arrays()) {
            B200_CHECK(schwz_b200_setup_create(
                0, 0, (int32_t)metadata.global_size,
                In32<I>(global_matrix->get_const_row_ptrs(), metadata.global_size + 1),
                In32<I>(global_matrix->get_const_col_idxs(), global_matrix->get_num_stored_elements()),
                global_matrix->get_const_values(), P, permute ? 1 : 0, part, settings.overlap, &s));
        } else {
            B200_CHECK(schwz_b200_setup_create(settings.laplacian_dim == 3 ? 2 : 1,
                                               (int32_t)metadata.oned_laplacian_size,
                                               (int32_t)metadata.global_size, nullptr, nullptr,
                                               nullptr, P, permute ? 1 : 0, part, settings.overlap,
                                               &s));
        }
        G.slots(kSlotSetup)[0] = s;
    }
    G.barrier();
    D.setup = (schwz_setup *)G.slots(kSlotSetup)[0];
    // the index sets were built once by rank 0; reading them and building the local matrices is
    // re-entrant, so the ranks work side by side
    auto host = settings.executor->get_master();
    for (int turn = me; turn == me; ++turn) {
        {
            B200_CHECK(schwz_b200_setup_first_row(D.setup,
                                                  Out32<I>(metadata.first_row->get_data(), P + 1)));
            int64_t sz[8];
            B200_CHECK(schwz_b200_setup_sizes(D.setup, me, sz));
            metadata.local_size = sz[0];
            metadata.local_size_x = sz[1];
            metadata.local_size_o = metadata.global_size;
            metadata.overlap_size = sz[2];
            metadata.local_to_global = std::make_shared<gko::Array<I>>(host, sz[1] + sz[5]);
            B200_CHECK(schwz_b200_setup_l2g(
                D.setup, me, Out32<I>(metadata.local_to_global->get_data(), sz[1] + sz[5])));
            metadata.overlap_row = gko::Array<I>(
                host, metadata.local_to_global->get_data() + sz[0],
                metadata.local_to_global->get_data() + sz[1]);
            if (permute && metadata.global_size <= kMaterializeLimit) {
                metadata.permutation = std::make_shared<gko::Array<I>>(host, metadata.global_size);
                metadata.i_permutation = std::make_shared<gko::Array<I>>(host, metadata.global_size);
                B200_CHECK(schwz_b200_setup_permutation(
                    D.setup, Out32<I>(metadata.permutation->get_data(), metadata.global_size),
                    Out32<I>(metadata.i_permutation->get_data(), metadata.global_size)));
            }
            local_matrix = gko::matrix::Csr<V, I>::create(host, gko::dim<2>(sz[1]), sz[3]);
            interface_matrix = gko::matrix::Csr<V, I>::create(
                host, sz[4] > 0 ? gko::dim<2>(sz[1]) : gko::dim<2>(0), sz[4]);
            if ((gko::size_type)sz[1] <= kMaterializeLimit / 8 || settings.print_matrices) {
                local_matrix->allocate();
                B200_CHECK(schwz_b200_setup_local_matrix(
                    D.setup, me, Out32<I>(local_matrix->get_row_ptrs(), sz[1] + 1),
                    Out32<I>(local_matrix->get_col_idxs(), sz[3]), local_matrix->get_values()));
                interface_matrix->allocate();
                B200_CHECK(schwz_b200_setup_interface_matrix(
                    D.setup, me,
                    Out32<I>(interface_matrix->get_row_ptrs(), sz[4] > 0 ? sz[1] + 1 : 1),
                    Out32<I>(interface_matrix->get_col_idxs(), sz[4]),
                    interface_matrix->get_values()));
            }
        }
    }
    G.barrier();
}

template <typename V, typename I, typename M>
void SolverRAS<V, I, M>::setup_comm_buffers()
{
    // source/restricted_schwarz.cpp:308-604: neighbour lists and [count, ids...]
    // index lists in the reference's layout; send/recv buffers become the
    // peer-visible mailbox owned by the device object created here.
    RankGroup &G = RankGroup::instance();
    b200::State &D = this->dev_;
    auto &metadata = this->metadata;
    auto &settings = this->settings;
    auto &cs = this->comm_struct;
    const int me = metadata.my_rank;
    const int P = (int)metadata.num_subdomains;
    for (int turn = me; turn == me; ++turn) {   // every rank at once (the setup object is re-entrant)
        {
            int64_t sz[8];
            B200_CHECK(schwz_b200_setup_sizes(D.setup, me, sz));
            cs.num_neighbors_in = (int)sz[6];
            cs.num_neighbors_out = (int)sz[7];
            B200_CHECK(schwz_b200_setup_neighbors(
                D.setup, me, Out32<I>(cs.neighbors_in->get_data(), std::max<int64_t>(sz[6], 1)),
                Out32<I>(cs.neighbors_out->get_data(), std::max<int64_t>(sz[7], 1))));
            cs.recv.assign(P, 0);
            cs.send.assign(P, 0);
            for (int j = 0; j < cs.num_neighbors_in; ++j) {
                int32_t cnt = 0;
                B200_CHECK(schwz_b200_setup_get_count(D.setup, me, j, &cnt));
                cs.list_storage.emplace_back(cnt + 1);
                auto &l = cs.list_storage.back();
                l[0] = cnt;
                B200_CHECK(schwz_b200_setup_get_list(D.setup, me, j, Out32<I>(l.data() + 1, cnt)));
                cs.recv[cs.neighbors_in->get_data()[j]] = cnt;
            }
            for (int j = 0; j < cs.num_neighbors_out; ++j) {
                int32_t cnt = 0;
                B200_CHECK(schwz_b200_setup_put_count(D.setup, me, j, &cnt));
                cs.list_storage.emplace_back(cnt + 1);
                auto &l = cs.list_storage.back();
                l[0] = cnt;
                B200_CHECK(schwz_b200_setup_put_list(D.setup, me, j, Out32<I>(l.data() + 1, cnt)));
                cs.send[cs.neighbors_out->get_data()[j]] = cnt;
            }
            for (int j = 0; j < cs.num_neighbors_in; ++j)
                cs.global_get->get_data()[j] = cs.local_get->get_data()[j] = cs.list_storage[j].data();
            for (int j = 0; j < cs.num_neighbors_out; ++j)
                cs.global_put->get_data()[j] = cs.local_put->get_data()[j] =
                    cs.list_storage[cs.num_neighbors_in + j].data();
            // A5 displacement tables (source/restricted_schwarz.cpp:624-658)
            B200_CHECK(schwz_b200_setup_displacements(
                D.setup, me, Out32<I>(cs.put_displacements->get_data(), P + 1),
                Out32<I>(cs.get_displacements->get_data(), P + 1)));
            // the device object: local + interface matrices, vectors, mailbox
            schwz_ras_options o{};
            o.tolerance = metadata.tolerance;
            o.local_tol = metadata.local_solver_tolerance;
            o.local_max_iters = metadata.local_max_iters;
            o.local_solver = settings.local_solver == Settings::iterative_solver_ginkgo ? 2 : 1;
            o.non_symmetric = settings.non_symmetric_matrix ? 1 : 0;
            o.restart_iter = (int32_t)settings.restart_iter;
            o.overlap = settings.overlap;
            // float mirrors of the halo buffers exist only when MixedValueType is float
            // (restricted_schwarz.cpp:483-492); with M = double the conversion is the identity
            o.use_mixed_precision =
                (settings.use_mixed_precision && std::is_same<M, float>::value) ? 1 : 0;
            // metadata.local_precond (solve.cpp:486-652); an unknown name leaves the solve
            // unpreconditioned after the "Unsupported preconditioner." message, as upstream
            if (o.local_solver == 2) {
                const std::string &pc = metadata.local_precond;
                o.local_precond = pc == "block-jacobi" ? SCHWZ_PRECOND_BLOCK_JACOBI
                                  : pc == "ilu"        ? SCHWZ_PRECOND_ILU
                                  : pc == "isai"       ? SCHWZ_PRECOND_ISAI
                                                       : SCHWZ_PRECOND_NONE;
                o.precond_max_block_size = (int32_t)metadata.precond_max_block_size;
            }
            B200_CHECK(schwz_b200_ras_create(D.ctx, D.setup, me, this->rhs_host_.data(), &o, &D.ras));
            if (D.have_factors)
                B200_CHECK(schwz_b200_ras_set_factors(D.ras, D.L_rowptr.data(), D.L_col.data(),
                                                      D.L_val.data(), D.factor_perm.data()));
            if (D.lu) {
                B200_CHECK(schwz_b200_ras_set_lu_factors(D.ras, D.lu, D.factor_perm.data()));
                B200_CHECK(schwz_b200_host_lu_destroy(D.lu));
                D.lu = nullptr;
            }
            B200_CHECK(schwz_b200_setup_release_rank(D.setup, me));
            G.slots(kSlotRas)[me] = D.ras;
            G.slots(kSlotCtx)[me] = D.ctx;
        }
    }
    G.barrier();
}

template <typename V, typename I, typename M>
void SolverRAS<V, I, M>::setup_windows(const Settings &, const Metadata<V, I> &metadata,
                                       std::shared_ptr<gko::matrix::Dense<V>> &)
{
    // source/restricted_schwarz.cpp:608-711: MPI_Win_create / lock_all become
    // peer access + the mailbox pointers of the neighbours
    RankGroup &G = RankGroup::instance();
    const int P = (int)metadata.num_subdomains;
    if (metadata.my_rank == 0) {
        std::vector<schwz_ctx *> ctxs;
        for (int r = 0; r < P; ++r) ctxs.push_back((schwz_ctx *)G.slots(kSlotCtx)[r]);
        B200_CHECK(schwz_b200_enable_peers(ctxs.data(), P));
        std::vector<schwz_ras *> all;
        for (int r = 0; r < P; ++r) all.push_back((schwz_ras *)G.slots(kSlotRas)[r]);
        B200_CHECK(schwz_b200_ras_connect_local(all.data(), P, this->dev_.setup));
    }
    G.barrier();
}

template <typename V, typename I, typename M>
void SolverRAS<V, I, M>::exchange_boundary(const Settings &settings, const Metadata<V, I> &metadata,
                                           std::shared_ptr<gko::matrix::Dense<V>> &)
{
    // source/restricted_schwarz.cpp:715-988
    RankGroup &G = RankGroup::instance();
    schwz_ras *ras = this->dev_.ras;
    if (metadata.num_subdomains < 2) return;
    if (settings.comm_settings.enable_onesided) {
        if (metadata.iter_count == 0) return;   // :725
        // push into the neighbours' buffers (or pack for them to pull), then unpack
        // whatever mine holds (or pull) — no synchronisation with the neighbours
        // (one-sided semantics).  enable_put/enable_get x enable_one_by_one (:753-851)
        // select the data path.
        const auto &c = settings.comm_settings;
        const int32_t mode = (c.enable_put ? 0 : 1) + (c.enable_one_by_one ? 2 : 0);
        B200_CHECK(schwz_b200_ras_set_exchange_mode(ras, mode));
        B200_CHECK(schwz_b200_ras_set_onesided(ras, 1));
        B200_CHECK(schwz_b200_ras_exchange_push(ras, metadata.iter_count));
        B200_CHECK(schwz_b200_ras_exchange_unpack(ras, metadata.iter_count, 0));
        return;
    }
    // two-sided: receive completes before unpack.  Ranks are host
    // threads: after everybody has enqueued its push, each rank makes its
    // stream wait on the push events of its in-neighbours, then unpacks.
    B200_CHECK(schwz_b200_ras_exchange_push(ras, metadata.iter_count));
    G.barrier();
    auto &cs = this->comm_struct;
    for (int j = 0; j < cs.num_neighbors_in; ++j) {
        auto *nbr = (schwz_ras *)G.slots(kSlotRas)[cs.neighbors_in->get_data()[j]];
        B200_CHECK(schwz_b200_ras_wait_push_of(ras, nbr));
    }
    B200_CHECK(schwz_b200_ras_exchange_unpack(ras, metadata.iter_count, 0));
    G.barrier();   // nobody re-records its push event before all waits are enqueued
}

template <typename V, typename I, typename M>
void SolverRAS<V, I, M>::update_boundary(const Settings &, const Metadata<V, I> &,
                                         std::shared_ptr<gko::matrix::Dense<V>> &,
                                         const std::shared_ptr<gko::matrix::Dense<V>> &,
                                         const std::shared_ptr<gko::matrix::Dense<V>> &,
                                         const std::shared_ptr<gko::matrix::Csr<V, I>> &)
{
    // source/restricted_schwarz.cpp:992-1017: local_solution = local_rhs - I x
    B200_CHECK(schwz_b200_ras_update_boundary(this->dev_.ras));
}

// explicit instantiations (the C ABI is fp64 / int32: the types of
// benchmarking/bench_ras.cpp:204)
template class Initialize<double, gko::int32>;
template class Communicate<double, gko::int32, double>;
template class Communicate<double, gko::int32, float>;
template class Solve<double, gko::int32, double>;
template class Solve<double, gko::int32, float>;
template class SchwarzBase<double, gko::int32, double>;
template class SchwarzBase<double, gko::int32, float>;
template class SolverRAS<double, gko::int32, double>;
template class SolverRAS<double, gko::int32, float>;
// IndexType = int64 (include/settings.hpp:533-537 of the reference instantiates it too): the
// host-side index sets are handed out as int64, the device side stays int32 (N < 2^31)
template class Initialize<double, gko::int64>;
template class Communicate<double, gko::int64, double>;
template class Communicate<double, gko::int64, float>;
template class Solve<double, gko::int64, double>;
template class Solve<double, gko::int64, float>;
template class SchwarzBase<double, gko::int64, double>;
template class SchwarzBase<double, gko::int64, float>;
template class SolverRAS<double, gko::int64, double>;
template class SolverRAS<double, gko::int64, float>;

}  // namespace schwz
