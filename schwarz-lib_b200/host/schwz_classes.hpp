// Class skeleton of schwarz-lib's solver façade, same names / inheritance /
// virtual override points as the reference:
//   Initialize   include/initialization.hpp:62-170
//   Communicate  include/communicate.hpp:62-300   (comm_struct :67-225)
//   Solve        include/solve.hpp
//   SchwarzBase  include/schwarz_base.hpp:76-216
//   SolverRAS    include/restricted_schwarz.hpp:61-104
// The bodies (schwz_host.cpp) drive the sm_100a kernels through the C ABI
// (include/schwz_b200.h); no Ginkgo, no MPI.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "../../include/schwz_b200.h"
#include "settings.hpp"

namespace schwz {

namespace b200 {
// Device-side state of one rank (subdomain) + what the ranks of a process share.
struct State {
    schwz_setup *setup = nullptr;   // shared by all ranks of the process (owned by rank 0)
    schwz_ctx *ctx = nullptr;
    schwz_ras *ras = nullptr;
    int device = 0;
    std::vector<int32_t> factor_perm, L_rowptr, L_col;
    std::vector<double> L_val;
    bool have_factors = false;
    schwz_lu *lu = nullptr;   // --local_factorization=umfpack: P A Q = L U, Q = factor_perm
    double resnorm = -1.0;
};
void check(int rc, const char *file, int line);
}  // namespace b200

template <typename ValueType = gko::default_precision, typename IndexType = gko::int32>
class Initialize : public Settings, public Metadata<ValueType, IndexType> {
public:
    Initialize(Settings &settings, Metadata<ValueType, IndexType> &metadata);
    virtual ~Initialize() = default;

    std::vector<unsigned int> partition_indices;
    std::vector<unsigned int> cell_weights;

    void generate_rhs(std::vector<ValueType> &rhs);
    void setup_global_matrix(const std::string &filename, const gko::size_type &oned_laplacian_size,
                             std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &global_matrix);
    void partition(const Settings &settings, const Metadata<ValueType, IndexType> &metadata,
                   const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &global_matrix,
                   std::vector<unsigned int> &partition_indices);
    void setup_vectors(const Settings &settings, const Metadata<ValueType, IndexType> &metadata,
                       std::vector<ValueType> &rhs,
                       std::shared_ptr<gko::matrix::Dense<ValueType>> &local_rhs,
                       std::shared_ptr<gko::matrix::Dense<ValueType>> &global_rhs,
                       std::shared_ptr<gko::matrix::Dense<ValueType>> &local_solution);
    virtual void setup_local_matrices(
        Settings &settings, Metadata<ValueType, IndexType> &metadata,
        std::vector<unsigned int> &partition_indices,
        std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &global_matrix,
        std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &local_matrix,
        std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &interface_matrix) = 0;

private:
    Settings &settings;
    Metadata<ValueType, IndexType> &metadata;
};

template <typename ValueType, typename IndexType, typename MixedValueType>
class Communicate {
public:
    virtual ~Communicate() = default;
    struct comm_struct {
        int num_neighbors_in = 0, num_neighbors_out = 0;
        std::shared_ptr<gko::Array<IndexType>> neighbors_in, neighbors_out;
        std::vector<bool> is_local_neighbor;
        int local_num_neighbors_in = 0, local_num_neighbors_out = 0;
        std::shared_ptr<gko::Array<IndexType>> local_neighbors_in, local_neighbors_out;
        // [count, i0, i1, ...] per neighbour, global indices (reference layout)
        std::shared_ptr<gko::Array<IndexType *>> global_put, local_put, global_get, local_get;
        std::vector<IndexType> send, recv;
        std::shared_ptr<gko::matrix::Dense<ValueType>> send_buffer, recv_buffer;
        std::shared_ptr<gko::matrix::Dense<MixedValueType>> mixedt_send_buffer, mixedt_recv_buffer;
        std::shared_ptr<gko::Array<IndexType>> get_displacements, put_displacements;
        std::vector<std::vector<IndexType>> list_storage;   // owns the [count, ...] lists
    };
    comm_struct comm_struct;

    virtual void setup_comm_buffers() = 0;
    virtual void setup_windows(const Settings &settings,
                               const Metadata<ValueType, IndexType> &metadata,
                               std::shared_ptr<gko::matrix::Dense<ValueType>> &main_buffer) = 0;
    virtual void exchange_boundary(const Settings &settings,
                                   const Metadata<ValueType, IndexType> &metadata,
                                   std::shared_ptr<gko::matrix::Dense<ValueType>> &global_solution) = 0;
    void local_to_global_vector(const Settings &settings,
                                const Metadata<ValueType, IndexType> &metadata,
                                const std::shared_ptr<gko::matrix::Dense<ValueType>> &local_vector,
                                std::shared_ptr<gko::matrix::Dense<ValueType>> &global_vector);
    virtual void update_boundary(
        const Settings &settings, const Metadata<ValueType, IndexType> &metadata,
        std::shared_ptr<gko::matrix::Dense<ValueType>> &local_solution,
        const std::shared_ptr<gko::matrix::Dense<ValueType>> &local_rhs,
        const std::shared_ptr<gko::matrix::Dense<ValueType>> &global_solution,
        const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &interface_matrix) = 0;
    void clear(Settings &settings);

protected:
    b200::State *comm_dev_ = nullptr;
};

template <typename ValueType, typename IndexType, typename MixedValueType>
class Solve {
public:
    virtual ~Solve() = default;

protected:
    void setup_local_solver(const Settings &settings, Metadata<ValueType, IndexType> &metadata,
                            const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &local_matrix,
                            std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &triangular_factor_l,
                            std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &triangular_factor_u,
                            std::shared_ptr<gko::matrix::Permutation<IndexType>> &local_perm,
                            std::shared_ptr<gko::matrix::Permutation<IndexType>> &local_inv_perm,
                            std::shared_ptr<gko::matrix::Dense<ValueType>> &local_rhs);
    void local_solve(const Settings &settings, Metadata<ValueType, IndexType> &metadata,
                     const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &local_matrix,
                     const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &triangular_factor_l,
                     const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &triangular_factor_u,
                     std::shared_ptr<gko::matrix::Permutation<IndexType>> &local_perm,
                     std::shared_ptr<gko::matrix::Permutation<IndexType>> &local_inv_perm,
                     std::shared_ptr<gko::matrix::Dense<ValueType>> &work_vector,
                     std::shared_ptr<gko::matrix::Dense<ValueType>> &init_guess,
                     std::shared_ptr<gko::matrix::Dense<ValueType>> &local_solution);
    bool check_local_convergence(const Settings &settings, Metadata<ValueType, IndexType> &metadata,
                                 ValueType &local_resnorm, ValueType &local_resnorm0);
    void check_global_convergence(
        const Settings &settings, Metadata<ValueType, IndexType> &metadata,
        struct Communicate<ValueType, IndexType, MixedValueType>::comm_struct &comm_struct,
        ValueType &local_resnorm, ValueType &local_resnorm0, ValueType &global_resnorm,
        ValueType &global_resnorm0, int &converged_all_local, int &num_converged_procs);
    void check_convergence(
        const Settings &settings, Metadata<ValueType, IndexType> &metadata,
        struct Communicate<ValueType, IndexType, MixedValueType>::comm_struct &comm_struct,
        ValueType &local_residual_norm, ValueType &local_residual_norm0,
        ValueType &global_residual_norm, ValueType &global_residual_norm0, int &num_converged_procs);
    void compute_residual_norm(const Settings &settings,
                               const Metadata<ValueType, IndexType> &metadata, ValueType &mat_norm,
                               ValueType &rhs_norm, ValueType &sol_norm, ValueType &residual_norm);
    void clear(Settings &settings);

    std::vector<ValueType> local_residual_vector;   // l_res of the reference (P doubles)
    b200::State *solve_dev_ = nullptr;
};

template <typename ValueType = gko::default_precision, typename IndexType = gko::int32,
          typename MixedValueType = gko::default_precision>
class SchwarzBase : public Initialize<ValueType, IndexType>,
                    public Communicate<ValueType, IndexType, MixedValueType>,
                    public Solve<ValueType, IndexType, MixedValueType> {
public:
    SchwarzBase(Settings &settings, Metadata<ValueType, IndexType> &metadata);
    ~SchwarzBase() override;

    void initialize();
    void run(std::shared_ptr<gko::matrix::Dense<ValueType>> &solution);

    std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> local_matrix;
    std::shared_ptr<gko::matrix::Permutation<IndexType>> local_perm;
    std::shared_ptr<gko::matrix::Permutation<IndexType>> local_inv_perm;
    std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> triangular_factor_l;
    std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> triangular_factor_u;
    std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> interface_matrix;
    std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> global_matrix;
    std::shared_ptr<gko::matrix::Dense<ValueType>> local_rhs;
    std::shared_ptr<gko::matrix::Dense<ValueType>> global_rhs;
    std::shared_ptr<gko::matrix::Dense<ValueType>> local_solution;
    std::shared_ptr<gko::matrix::Dense<ValueType>> global_solution;
    std::vector<ValueType> local_residual_vector_out;
    std::vector<std::vector<ValueType>> global_residual_vector_out;

protected:
    Settings &settings;
    Metadata<ValueType, IndexType> &metadata;
    b200::State dev_;
    std::vector<ValueType> rhs_host_;   // global rhs (permuted numbering)
};

template <typename ValueType = gko::default_precision, typename IndexType = gko::int32,
          typename MixedValueType = gko::default_precision>
class SolverRAS : public SchwarzBase<ValueType, IndexType, MixedValueType> {
public:
    SolverRAS(Settings &settings, Metadata<ValueType, IndexType> &metadata);

    void setup_local_matrices(
        Settings &settings, Metadata<ValueType, IndexType> &metadata,
        std::vector<unsigned int> &partition_indices,
        std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &global_matrix,
        std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &local_matrix,
        std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &interface_matrix) override;
    void setup_comm_buffers() override;
    void setup_windows(const Settings &settings, const Metadata<ValueType, IndexType> &metadata,
                       std::shared_ptr<gko::matrix::Dense<ValueType>> &main_buffer) override;
    void exchange_boundary(const Settings &settings, const Metadata<ValueType, IndexType> &metadata,
                           std::shared_ptr<gko::matrix::Dense<ValueType>> &global_solution) override;
    void update_boundary(
        const Settings &settings, const Metadata<ValueType, IndexType> &metadata,
        std::shared_ptr<gko::matrix::Dense<ValueType>> &local_solution,
        const std::shared_ptr<gko::matrix::Dense<ValueType>> &local_rhs,
        const std::shared_ptr<gko::matrix::Dense<ValueType>> &global_solution,
        const std::shared_ptr<gko::matrix::Csr<ValueType, IndexType>> &interface_matrix) override;
};

}  // namespace schwz
