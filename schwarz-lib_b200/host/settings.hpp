// schwz::Settings / schwz::Metadata — the user-visible configuration structs of
// schwarz-lib (include/settings.hpp:77-305 and :318-496 of the reference), kept
// field-for-field so that driver code written against the reference compiles
// unchanged.  Only the types behind gko:: / MPI_ differ (gko_compat.hpp,
// mpi_compat.hpp).
#pragma once
#include <chrono>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "gko_compat.hpp"
#include "mpi_compat.hpp"

#define MINIMAL_OVERLAP 2

namespace schwz {

struct Settings {
    // -- execution -----------------------------------------------------------
    std::string executor_string;                       // "reference" | "omp" | "cuda"
    std::shared_ptr<gko::Executor> executor = gko::ReferenceExecutor::create();
    std::shared_ptr<void> cuda_device_guard;           // RAII device guard in the reference

    // -- partitioning / problem ------------------------------------------------
    enum partition_settings {
        partition_regular = 0x0,
        partition_regular2d = 0x4,
        partition_metis = 0x1,
        partition_zoltan = 0x2,
        partition_custom = 0x3
    };
    partition_settings partition = partition_settings::partition_regular;
    gko::int32 overlap = MINIMAL_OVERLAP;
    std::string matrix_filename = "null";
    bool explicit_laplacian = true;
    bool use_mixed_precision = false;
    bool enable_random_rhs = false;
    bool print_matrices = false;
    bool debug_print = false;

    // -- local solver ----------------------------------------------------------
    enum local_solver_settings {
        direct_solver_cholmod = 0x0,
        direct_solver_umfpack = 0x5,
        direct_solver_ginkgo = 0x1,
        iterative_solver_ginkgo = 0x2,
        iterative_solver_dealii = 0x3,
        solver_custom = 0x4
    };
    local_solver_settings local_solver = local_solver_settings::iterative_solver_ginkgo;
    bool non_symmetric_matrix = false;
    unsigned int restart_iter = 1u;
    int reset_local_crit_iter = -1;
    bool naturally_ordered_factor = false;
    std::string metis_objtype;
    bool use_precond = false;

    // -- output ----------------------------------------------------------------
    bool write_debug_out = false;
    bool write_iters_and_residuals = false;
    bool enable_logging = false;
    bool write_perm_data = false;
    int shifted_iter = 1;

    struct comm_settings {
        bool enable_onesided = false;
        bool enable_overlap = false;
        bool enable_put = false;
        bool enable_get = true;
        bool stage_through_host = false;
        bool enable_one_by_one = false;
        bool enable_flush_local = false;
        bool enable_flush_all = true;
        bool enable_lock_local = false;
        bool enable_lock_all = true;
    };
    comm_settings comm_settings;

    struct convergence_settings {
        bool put_all_local_residual_norms = true;
        bool enable_global_simple_tree = false;
        bool enable_decentralized_leader_election = false;
        bool enable_global_check = true;
        bool enable_accumulate = false;
        bool enable_global_check_iter_offset = false;
        enum local_convergence_crit { residual_based = 0x0, solution_based = 0x1 };
        local_convergence_crit convergence_crit = local_convergence_crit::solution_based;
    };
    convergence_settings convergence_settings;

    std::string factorization = "cholmod";
    std::string reorder;

    // -- additions of this implementation (not in the reference) ---------------
    int num_devices = 0;       // 0 = all visible GPUs; subdomain s runs on GPU s % num_devices
    int laplacian_dim = 2;     // 3 = generated 3-D 7-pt Laplacian (the reference has no 3-D generator)

    Settings(std::string executor_string = "reference") : executor_string(executor_string) {}
};

template <typename ValueType, typename IndexType>
struct Metadata {
    MPI_Comm mpi_communicator = MPI_COMM_WORLD;

    gko::size_type global_size = 0;
    gko::size_type oned_laplacian_size = 0;
    gko::size_type local_size = 0;
    gko::size_type local_size_x = 0;
    gko::size_type local_size_o = 0;
    gko::size_type overlap_size = 0;
    gko::size_type num_subdomains = 1;

    int my_rank = 0;
    int my_local_rank = 0;
    int local_num_procs = 1;
    int comm_size = 1;
    int num_threads = 1;

    IndexType iter_count = 0;
    ValueType tolerance = 1e-6;
    ValueType local_solver_tolerance = 1e-12;
    IndexType max_iters = 100;
    IndexType local_max_iters = -1;
    IndexType updated_max_iters = -1;
    std::string local_precond = "null";
    unsigned int precond_max_block_size = 16;
    ValueType current_residual_norm = -1.0;
    ValueType min_residual_norm = -1.0;

    // (id, rank, last iteration, name, samples) per timed stage
    std::vector<std::tuple<int, int, int, std::string, std::vector<ValueType>>> time_struct;
    // (subdomain, [(from, count)], [(to, count)], #in, #out)
    std::vector<std::tuple<int, std::vector<std::tuple<int, int>>, std::vector<std::tuple<int, int>>,
                           int, int>>
        comm_data_struct;

    struct post_process_data {
        std::vector<std::vector<ValueType>> global_residual_vector_out;
        std::vector<ValueType> local_residual_vector_out;
        std::vector<ValueType> local_converged_iter_count;
        std::vector<ValueType> local_converged_resnorm;
        std::vector<ValueType> local_timestamp;
    };
    post_process_data post_process_data;
    double init_mpi_wtime = 0.0;

    std::shared_ptr<gko::Array<IndexType>> global_to_local;
    std::shared_ptr<gko::Array<IndexType>> local_to_global;
    gko::Array<IndexType> overlap_row;
    std::shared_ptr<gko::Array<IndexType>> first_row;
    std::shared_ptr<gko::Array<IndexType>> permutation;
    std::shared_ptr<gko::Array<IndexType>> i_permutation;
};

// Stage timer with the reference's bookkeeping (include/settings.hpp:508-523):
// the entry is created at iteration 0 and appended to afterwards.
#define MEASURE_ELAPSED_FUNC_TIME(_func, _id, _rank, _name, _iter)                          \
    {                                                                                       \
        auto _t0 = std::chrono::steady_clock::now();                                        \
        _func;                                                                              \
        auto _dt = std::chrono::duration<ValueType>(std::chrono::steady_clock::now() - _t0); \
        if (_iter == 0) {                                                                   \
            metadata.time_struct.push_back(std::make_tuple(                                 \
                _id, _rank, _iter, #_name, std::vector<ValueType>(1, _dt.count())));        \
        } else {                                                                            \
            std::get<2>(metadata.time_struct[_id]) = _iter;                                 \
            std::get<4>(metadata.time_struct[_id]).push_back(_dt.count());                  \
        }                                                                                   \
    }

}  // namespace schwz

namespace schwarz = schwz;   // BASELINE.json's north_star spells it schwarz::
