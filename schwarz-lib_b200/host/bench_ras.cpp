// bench_ras — the reference's benchmark driver (benchmarking/bench_ras.cpp) on
// the B200 implementation.  BenchRas::solve wires flags into Settings /
// Metadata exactly as the reference does (:47-190) and calls
// SolverRAS::initialize() / run(); main() starts one host thread per subdomain
// where the reference is started as `mpirun -n P`.
#include <cmath>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <string>
#include <utility>

#include "bench_base.hpp"

template <typename ValueType, typename IndexType>
class BenchRas : public BenchBase<ValueType, IndexType> {
public:
    void run() { solve(MPI_COMM_WORLD); }

private:
    void solve(MPI_Comm mpi_communicator);
};

namespace {

// A string-valued flag selects one entry of a small table; an unknown value leaves the settings
// at their defaults, which is what the if / else-if chains of the reference amount to
// (benchmarking/bench_ras.cpp:72-149).
template <typename Action>
void select(const std::string &value, std::initializer_list<std::pair<const char *, Action>> table)
{
    for (const auto &entry : table)
        if (value == entry.first) {
            entry.second();
            return;
        }
}
using Apply = std::function<void()>;

void wire_communication(schwz::Settings &st)
{
    auto &c = st.comm_settings;
    c.enable_onesided = FLAGS_enable_onesided;
    c.enable_one_by_one = FLAGS_enable_one_by_one;
    c.enable_overlap = FLAGS_enable_comm_overlap;
    select<Apply>(FLAGS_remote_comm_type, {{"put", [&] { c.enable_put = true, c.enable_get = false; }},
                                          {"get", [&] { c.enable_put = false, c.enable_get = true; }}});
    select<Apply>(FLAGS_flush_type,
                  {{"flush-all", [&] { c.enable_flush_all = true; }},
                   {"flush-local", [&] { c.enable_flush_all = false, c.enable_flush_local = true; }}});
    select<Apply>(FLAGS_lock_type,
                  {{"lock-all", [&] { c.enable_lock_all = true; }},
                   {"lock-local", [&] { c.enable_lock_all = false, c.enable_lock_local = true; }}});
}

void wire_convergence(schwz::Settings &st)
{
    auto &c = st.convergence_settings;
    c.enable_global_check = FLAGS_enable_global_check;
    c.enable_global_check_iter_offset = FLAGS_enable_global_check_iter_offset;
    c.put_all_local_residual_norms = FLAGS_enable_put_all_local_residual_norms;
    select<Apply>(FLAGS_global_convergence_type,
                  {{"centralized-tree", [&] { c.enable_global_simple_tree = true; }},
                   {"decentralized", [&] {
                        c.enable_decentralized_leader_election = true;
                        c.enable_accumulate = FLAGS_enable_decentralized_accumulate;
                    }}});
}

void wire_problem_and_local_solver(schwz::Settings &st)
{
    using S = schwz::Settings;
    st.matrix_filename = FLAGS_matrix_filename;
    st.explicit_laplacian = FLAGS_explicit_laplacian;
    st.enable_random_rhs = FLAGS_enable_random_rhs;
    st.overlap = FLAGS_overlap;
    st.non_symmetric_matrix = FLAGS_non_symmetric_matrix;
    st.restart_iter = FLAGS_restart_iter;
    st.naturally_ordered_factor = FLAGS_factor_ordering_natural;
    st.reorder = FLAGS_local_reordering;
    st.factorization = FLAGS_local_factorization;
    select<Apply>(FLAGS_partition,
                  {{"regular", [&] { st.partition = S::partition_settings::partition_regular; }},
                   {"regular2d", [&] { st.partition = S::partition_settings::partition_regular2d; }},
                   {"metis", [&] {
                        st.partition = S::partition_settings::partition_metis;
                        st.metis_objtype = FLAGS_metis_objtype;
                    }}});
    using L = S::local_solver_settings;
    select<Apply>(FLAGS_local_solver,
                  {{"iterative-ginkgo", [&] { st.local_solver = L::iterative_solver_ginkgo; }},
                   {"direct-ginkgo", [&] { st.local_solver = L::direct_solver_ginkgo; }},
                   {"direct-cholmod", [&] { st.local_solver = L::direct_solver_cholmod; }},
                   {"direct-umfpack", [&] { st.local_solver = L::direct_solver_umfpack; }}});
    // output switches
    st.write_debug_out = FLAGS_enable_debug_write;
    st.write_perm_data = FLAGS_write_perm_data;
    st.write_iters_and_residuals = FLAGS_write_iters_and_residuals;
    st.print_matrices = FLAGS_print_matrices;
    st.shifted_iter = FLAGS_shifted_iter;
    st.debug_print = FLAGS_debug;
    // launcher additions (not in the reference)
    st.num_devices = FLAGS_num_devices;
    st.laplacian_dim = FLAGS_laplacian_dim;
}

}  // namespace

template <typename ValueType, typename IndexType>
void BenchRas<ValueType, IndexType>::solve(MPI_Comm mpi_communicator)
{
    // flags -> Settings / Metadata, field for field what benchmarking/bench_ras.cpp:47-150 sets
    schwz::Settings settings(FLAGS_executor);
    wire_communication(settings);
    wire_convergence(settings);
    wire_problem_and_local_solver(settings);

    schwz::Metadata<ValueType, IndexType> metadata;
    metadata.mpi_communicator = mpi_communicator;
    MPI_Comm_rank(mpi_communicator, &metadata.my_rank);
    MPI_Comm_size(mpi_communicator, &metadata.comm_size);
    metadata.num_subdomains = metadata.comm_size;
    metadata.num_threads = FLAGS_num_threads;
    metadata.oned_laplacian_size = FLAGS_set_1d_laplacian_size;
    metadata.tolerance = FLAGS_set_tol;
    metadata.max_iters = FLAGS_num_iters;
    metadata.local_solver_tolerance = FLAGS_local_tol;
    metadata.local_max_iters = FLAGS_local_max_iters;
    metadata.local_precond = FLAGS_local_precond;
    metadata.precond_max_block_size = FLAGS_precond_max_block_size;

    std::shared_ptr<gko::matrix::Dense<ValueType>> explicit_laplacian_solution;

    if (metadata.my_rank == 0) {
        std::cout << " Running on the " << FLAGS_executor << " executor on " << metadata.num_subdomains
                  << " ranks with " << FLAGS_num_threads << " threads" << std::endl;
        std::cout << " Problem Size: " << metadata.global_size << std::endl;
    }
    if (FLAGS_print_config && metadata.my_rank == 0) this->print_config();

    schwz::SolverRAS<ValueType, IndexType> solver(settings, metadata);
    solver.initialize();
    solver.run(explicit_laplacian_solution);
    std::string rank_string = std::to_string(metadata.my_rank);
    if (metadata.my_rank < 10) rank_string = "0" + rank_string;
    if (FLAGS_timings_file != "null")
        this->write_timings(metadata.time_struct, FLAGS_timings_file + "_" + rank_string + ".csv",
                            settings.comm_settings.enable_onesided);
    if (FLAGS_write_comm_data && !metadata.comm_data_struct.empty())
        this->write_comm_data(metadata.num_subdomains, metadata.my_rank, metadata.comm_data_struct,
                              "num_send_" + rank_string + ".csv", "num_recv_" + rank_string + ".csv");
    if (metadata.my_rank == 0 && FLAGS_debug && explicit_laplacian_solution) {
        double s = 0.0;
        for (gko::size_type i = 0; i < metadata.global_size; ++i)
            s += explicit_laplacian_solution->at(i) * explicit_laplacian_solution->at(i);
        std::cout << " solution norm " << std::sqrt(s) << std::endl;
    }
}

int main(int argc, char *argv[])
{
    try {
        initialize_argument_parsing(&argc, &argv);
        MPI_Init(&argc, &argv);
        int P = (int)FLAGS_num_subdomains;
        if (P == 0) {
            int n = 0;
            schwz_b200_device_count(&n);
            P = n > 0 ? n : 1;
        }
        schwz_mpi::RankGroup::instance().run(P, [](int) {
            if (FLAGS_index_bits == 64) {   // the reference instantiates (double, int64) too
                BenchRas<double, gko::int64> laplace_problem_2d;
                laplace_problem_2d.run();
            } else {
                BenchRas<double, int> laplace_problem_2d;
                laplace_problem_2d.run();
            }
        });
        MPI_Finalize();
    } catch (std::exception &exc) {
        std::cerr << std::endl
                  << std::endl
                  << "----------------------------------------------------" << std::endl;
        std::cerr << "Exception on processing: " << std::endl
                  << exc.what() << std::endl
                  << "Aborting!" << std::endl
                  << "----------------------------------------------------" << std::endl;
        return 1;
    } catch (...) {
        std::cerr << std::endl
                  << std::endl
                  << "----------------------------------------------------" << std::endl;
        std::cerr << "Unknown exception!" << std::endl
                  << "Aborting!" << std::endl
                  << "----------------------------------------------------" << std::endl;
        return 1;
    }
    return 0;
}
