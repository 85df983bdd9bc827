// bench_ras — the reference's benchmark driver (benchmarking/bench_ras.cpp) on
// the B200 implementation.  BenchRas::solve wires flags into Settings /
// Metadata exactly as the reference does (:47-190) and calls
// SolverRAS::initialize() / run(); main() starts one host thread per subdomain
// where the reference is started as `mpirun -n P`.
#include <cmath>
#include <iostream>

#include "bench_base.hpp"

template <typename ValueType, typename IndexType>
class BenchRas : public BenchBase<ValueType, IndexType> {
public:
    void run() { solve(MPI_COMM_WORLD); }

private:
    void solve(MPI_Comm mpi_communicator);
};

template <typename ValueType, typename IndexType>
void BenchRas<ValueType, IndexType>::solve(MPI_Comm mpi_communicator)
{
    schwz::Metadata<ValueType, IndexType> metadata;
    schwz::Settings settings(FLAGS_executor);

    metadata.mpi_communicator = mpi_communicator;
    MPI_Comm_rank(metadata.mpi_communicator, &metadata.my_rank);
    MPI_Comm_size(metadata.mpi_communicator, &metadata.comm_size);
    metadata.tolerance = FLAGS_set_tol;
    metadata.max_iters = FLAGS_num_iters;
    metadata.num_subdomains = metadata.comm_size;
    metadata.num_threads = FLAGS_num_threads;
    metadata.oned_laplacian_size = FLAGS_set_1d_laplacian_size;

    settings.write_debug_out = FLAGS_enable_debug_write;
    settings.write_perm_data = FLAGS_write_perm_data;
    settings.write_iters_and_residuals = FLAGS_write_iters_and_residuals;
    settings.print_matrices = FLAGS_print_matrices;
    settings.shifted_iter = FLAGS_shifted_iter;

    settings.comm_settings.enable_onesided = FLAGS_enable_onesided;
    if (FLAGS_remote_comm_type == "put") {
        settings.comm_settings.enable_put = true;
        settings.comm_settings.enable_get = false;
    } else if (FLAGS_remote_comm_type == "get") {
        settings.comm_settings.enable_put = false;
        settings.comm_settings.enable_get = true;
    }
    settings.comm_settings.enable_one_by_one = FLAGS_enable_one_by_one;
    settings.comm_settings.enable_overlap = FLAGS_enable_comm_overlap;
    if (FLAGS_flush_type == "flush-all") {
        settings.comm_settings.enable_flush_all = true;
    } else if (FLAGS_flush_type == "flush-local") {
        settings.comm_settings.enable_flush_all = false;
        settings.comm_settings.enable_flush_local = true;
    }
    if (FLAGS_lock_type == "lock-all") {
        settings.comm_settings.enable_lock_all = true;
    } else if (FLAGS_lock_type == "lock-local") {
        settings.comm_settings.enable_lock_all = false;
        settings.comm_settings.enable_lock_local = true;
    }

    settings.convergence_settings.put_all_local_residual_norms = FLAGS_enable_put_all_local_residual_norms;
    settings.convergence_settings.enable_global_check_iter_offset = FLAGS_enable_global_check_iter_offset;
    settings.convergence_settings.enable_global_check = FLAGS_enable_global_check;
    if (FLAGS_global_convergence_type == "centralized-tree") {
        settings.convergence_settings.enable_global_simple_tree = true;
    } else if (FLAGS_global_convergence_type == "decentralized") {
        settings.convergence_settings.enable_decentralized_leader_election = true;
        settings.convergence_settings.enable_accumulate = FLAGS_enable_decentralized_accumulate;
    }

    metadata.local_solver_tolerance = FLAGS_local_tol;
    metadata.local_precond = FLAGS_local_precond;
    metadata.local_max_iters = FLAGS_local_max_iters;
    settings.non_symmetric_matrix = FLAGS_non_symmetric_matrix;
    settings.restart_iter = FLAGS_restart_iter;
    metadata.precond_max_block_size = FLAGS_precond_max_block_size;
    settings.matrix_filename = FLAGS_matrix_filename;
    settings.explicit_laplacian = FLAGS_explicit_laplacian;
    settings.enable_random_rhs = FLAGS_enable_random_rhs;
    settings.overlap = FLAGS_overlap;
    settings.naturally_ordered_factor = FLAGS_factor_ordering_natural;
    settings.reorder = FLAGS_local_reordering;
    settings.factorization = FLAGS_local_factorization;
    if (FLAGS_partition == "metis") {
        settings.partition = schwz::Settings::partition_settings::partition_metis;
        settings.metis_objtype = FLAGS_metis_objtype;
    } else if (FLAGS_partition == "regular") {
        settings.partition = schwz::Settings::partition_settings::partition_regular;
    } else if (FLAGS_partition == "regular2d") {
        settings.partition = schwz::Settings::partition_settings::partition_regular2d;
    }
    if (FLAGS_local_solver == "iterative-ginkgo") {
        settings.local_solver = schwz::Settings::local_solver_settings::iterative_solver_ginkgo;
    } else if (FLAGS_local_solver == "direct-cholmod") {
        settings.local_solver = schwz::Settings::local_solver_settings::direct_solver_cholmod;
    } else if (FLAGS_local_solver == "direct-umfpack") {
        settings.local_solver = schwz::Settings::local_solver_settings::direct_solver_umfpack;
    } else if (FLAGS_local_solver == "direct-ginkgo") {
        settings.local_solver = schwz::Settings::local_solver_settings::direct_solver_ginkgo;
    }
    settings.debug_print = FLAGS_debug;
    // launcher additions
    settings.num_devices = FLAGS_num_devices;
    settings.laplacian_dim = FLAGS_laplacian_dim;

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
