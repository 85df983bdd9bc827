// Compile-only check (tests/test_host_api.py): user code written against the reference's public
// surface - include <restricted_schwarz.hpp>, subclass schwz::SolverRAS and override its five
// virtual extension points with the reference's signatures (include/restricted_schwarz.hpp:76-103),
// construct / initialize / run, read the public data members (include/schwarz_base.hpp:137-197) -
// compiles unchanged against schwarz-lib_b200/host/.
#include <restricted_schwarz.hpp>

#include <memory>
#include <vector>

template <typename V, typename I, typename M>
class MyRas : public schwz::SolverRAS<V, I, M> {
public:
    using Base = schwz::SolverRAS<V, I, M>;
    MyRas(schwz::Settings &s, schwz::Metadata<V, I> &m) : Base(s, m) {}

    void setup_local_matrices(schwz::Settings &settings, schwz::Metadata<V, I> &metadata,
                              std::vector<unsigned int> &partition_indices,
                              std::shared_ptr<gko::matrix::Csr<V, I>> &global_matrix,
                              std::shared_ptr<gko::matrix::Csr<V, I>> &local_matrix,
                              std::shared_ptr<gko::matrix::Csr<V, I>> &interface_matrix) override
    {
        Base::setup_local_matrices(settings, metadata, partition_indices, global_matrix, local_matrix,
                                   interface_matrix);
    }
    void setup_comm_buffers() override { Base::setup_comm_buffers(); }
    void setup_windows(const schwz::Settings &settings, const schwz::Metadata<V, I> &metadata,
                       std::shared_ptr<gko::matrix::Dense<V>> &main_buffer) override
    {
        Base::setup_windows(settings, metadata, main_buffer);
    }
    void exchange_boundary(const schwz::Settings &settings, const schwz::Metadata<V, I> &metadata,
                           std::shared_ptr<gko::matrix::Dense<V>> &global_solution) override
    {
        Base::exchange_boundary(settings, metadata, global_solution);
    }
    void update_boundary(const schwz::Settings &settings, const schwz::Metadata<V, I> &metadata,
                         std::shared_ptr<gko::matrix::Dense<V>> &local_solution,
                         const std::shared_ptr<gko::matrix::Dense<V>> &local_rhs,
                         const std::shared_ptr<gko::matrix::Dense<V>> &global_solution,
                         const std::shared_ptr<gko::matrix::Csr<V, I>> &interface_matrix) override
    {
        Base::update_boundary(settings, metadata, local_solution, local_rhs, global_solution,
                              interface_matrix);
    }
};

template <typename V, typename I, typename M>
void drive()
{
    schwarz::Settings settings("cuda");            // the north star spells the namespace schwarz::
    schwz::Metadata<V, I> metadata;
    settings.explicit_laplacian = true;
    settings.partition = schwz::Settings::partition_settings::partition_regular2d;
    settings.local_solver = schwz::Settings::local_solver_settings::iterative_solver_ginkgo;
    settings.comm_settings.enable_onesided = false;
    settings.convergence_settings.enable_global_check = true;
    metadata.oned_laplacian_size = 16;
    metadata.tolerance = 1e-6;
    metadata.max_iters = 10;
    MyRas<V, I, M> solver(settings, metadata);
    solver.initialize();
    std::shared_ptr<gko::matrix::Dense<V>> solution;
    solver.run(solution);
    // public data members of SchwarzBase
    (void)solver.local_matrix;
    (void)solver.interface_matrix;
    (void)solver.global_matrix;
    (void)solver.local_rhs;
    (void)solver.global_rhs;
    (void)solver.local_solution;
    (void)solver.global_solution;
    (void)solver.triangular_factor_l;
    (void)solver.triangular_factor_u;
    (void)solver.local_perm;
    (void)solver.local_inv_perm;
    (void)solver.local_residual_vector_out;
    (void)solver.global_residual_vector_out;
    (void)metadata.time_struct;
    (void)metadata.comm_data_struct;
}

template void drive<double, gko::int32, double>();
template void drive<double, gko::int32, float>();
template void drive<double, gko::int64, double>();
template void drive<double, gko::int64, float>();
