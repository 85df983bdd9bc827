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

    // -- problem and partitioning (values of the enumerators as in the reference) ----
    enum partition_settings { partition_regular = 0, partition_metis = 1, partition_zoltan = 2,
                              partition_custom = 3, partition_regular2d = 4 };
    partition_settings partition = partition_regular;
    std::string matrix_filename = "null", metis_objtype;
    gko::int32 overlap = MINIMAL_OVERLAP;
    bool explicit_laplacian = true, enable_random_rhs = false, use_mixed_precision = false;

    // -- local solver ----------------------------------------------------------
    enum local_solver_settings { direct_solver_cholmod = 0, direct_solver_ginkgo = 1,
                                 iterative_solver_ginkgo = 2, iterative_solver_dealii = 3,
                                 solver_custom = 4, direct_solver_umfpack = 5 };
    local_solver_settings local_solver = iterative_solver_ginkgo;
    std::string factorization = "cholmod", reorder;    // "cholmod" | "umfpack"; reorder is unused
    bool non_symmetric_matrix = false, naturally_ordered_factor = false, use_precond = false;
    unsigned int restart_iter = 1u;
    int reset_local_crit_iter = -1;   // past this outer iteration: metadata.updated_max_iters

    // -- output ----------------------------------------------------------------
    bool print_matrices = false, debug_print = false, write_debug_out = false,
         write_iters_and_residuals = false, write_perm_data = false, enable_logging = false;
    int shifted_iter = 1;

    // halo exchange (all defaults as upstream: one-sided off, Get, flush-all, lock-all)
    struct comm_settings {
        bool enable_onesided = false, enable_overlap = false, stage_through_host = false;
        bool enable_put = false, enable_get = true, enable_one_by_one = false;
        bool enable_flush_all = true, enable_flush_local = false;
        bool enable_lock_all = true, enable_lock_local = false;
    };
    comm_settings comm_settings;

    struct convergence_settings {
        bool enable_global_check = true, enable_global_check_iter_offset = false;
        bool enable_global_simple_tree = false, enable_decentralized_leader_election = false,
             enable_accumulate = false;
        bool put_all_local_residual_norms = true;
        enum local_convergence_crit { residual_based = 0, solution_based = 1 };
        local_convergence_crit convergence_crit = solution_based;
    };
    convergence_settings convergence_settings;

    // -- additions of this implementation (not in the reference) ---------------
    int num_devices = 0;       // 0 = all visible GPUs; subdomain s runs on GPU s % num_devices
    int laplacian_dim = 2;     // 3 = generated 3-D 7-pt Laplacian (the reference has no 3-D generator)

    Settings(std::string executor_string = "reference") : executor_string(executor_string) {}
};

template <typename ValueType, typename IndexType>
struct Metadata {
    MPI_Comm mpi_communicator = MPI_COMM_WORLD;
    int my_rank = 0, my_local_rank = 0, local_num_procs = 1, comm_size = 1, num_threads = 1;

    // sizes: global problem, generated-Laplacian edge, and this subdomain's own / own+overlap /
    // "locally global" / overlap counts
    gko::size_type global_size = 0, oned_laplacian_size = 0, num_subdomains = 1;
    gko::size_type local_size = 0, local_size_x = 0, local_size_o = 0, overlap_size = 0;

    // outer iteration
    IndexType iter_count = 0, max_iters = 100;
    ValueType tolerance = 1e-6, current_residual_norm = -1.0, min_residual_norm = -1.0;
    // local solve
    ValueType local_solver_tolerance = 1e-12;
    IndexType local_max_iters = -1, updated_max_iters = -1;
    std::string local_precond = "null";
    unsigned int precond_max_block_size = 16;

    // (id, rank, last iteration, name, samples) per timed stage
    std::vector<std::tuple<int, int, int, std::string, std::vector<ValueType>>> time_struct;
    // (subdomain, [(from, count)], [(to, count)], #in, #out)
    using neighbour_counts = std::vector<std::tuple<int, int>>;
    std::vector<std::tuple<int, neighbour_counts, neighbour_counts, int, int>> comm_data_struct;

    struct post_process_data {
        std::vector<std::vector<ValueType>> global_residual_vector_out;
        std::vector<ValueType> local_residual_vector_out, local_converged_iter_count,
            local_converged_resnorm, local_timestamp;
    };
    post_process_data post_process_data;
    double init_mpi_wtime = 0.0;

    // index sets (host): global <-> local numbering, overlap rows, block starts, partition
    // permutation and its inverse
    using index_array = gko::Array<IndexType>;
    std::shared_ptr<index_array> global_to_local, local_to_global, first_row, permutation,
        i_permutation;
    index_array overlap_row;
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
