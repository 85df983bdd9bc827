// Host-side index sets of one RAS problem: partition -> permutation -> overlap
// growth -> local / interface CSR -> halo lists -> displacements.
//
// Produces arrays bit-identical to SolverRAS::setup_local_matrices /
// setup_comm_buffers / setup_windows (source/restricted_schwarz.cpp:56-711)
// without replicating the global matrix per subdomain: rows come from a
// RowSource (generated stencil, stored CSR, or either one seen through the
// partition permutation); global->local lookups go through a small per-call index (own range +
// hash of the overlap / halo ids), so the subdomains of a process share no scratch.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace schwz_b200 {

struct HostCsr {
    int32_t nrows = 0, ncols = 0;
    std::vector<int32_t> rp, ci;
    std::vector<double> v;
    int64_t nnz() const { return rp.empty() ? 0 : rp[nrows]; }
};

// Row access to the (permuted) global matrix, entries in the order the
// reference's loops see them.
struct RowSource {
    int32_t N = 0;
    int max_row = 0;
    virtual ~RowSource() = default;
    // writes the entries of row i, returns their number (<= max_row)
    virtual int row(int32_t i, int32_t *cols, double *vals) const = 0;
};

std::unique_ptr<RowSource> make_laplacian2d(int32_t n);
std::unique_ptr<RowSource> make_laplacian3d(int32_t n);
std::unique_ptr<RowSource> make_stored(int32_t N, const int32_t *rp, const int32_t *ci,
                                       const double *v, bool copy);
HostCsr materialize(const RowSource &src);
HostCsr read_mtx(const std::string &path);   // gko::read + sort_by_column_index

void partition_regular2d(int64_t N, int32_t P, uint32_t *part);
bool partition_regular2d_rect(int64_t N, int32_t P, int32_t px, int32_t py, uint32_t *part);
int partition_metis(int32_t N, const int32_t *rp, const int32_t *ci, int32_t P,
                    const char *objtype, uint32_t *part);
int nd_ordering(int32_t n, const int32_t *rp, const int32_t *ci, int32_t *perm);
// simplicial LL^T of P A P^T (up-looking, elimination-tree reach); returns
// false when a pivot is not positive
bool host_cholesky(const HostCsr &A, const int32_t *perm, HostCsr &L);
HostCsr transpose(const HostCsr &A);
// sparse LU with threshold partial pivoting: P A Q = L U, q given (may be NULL), p returned
bool host_sparse_lu(const HostCsr &A, const int32_t *q, double diag_tol, HostCsr &L, HostCsr &U,
                    std::vector<int32_t> &p);
// fill-reducing ordering of the pattern of A + A^T (METIS_NodeND)
int nd_ordering_symmetrized(int32_t n, const int32_t *rp, const int32_t *ci, int32_t *perm);

struct RankLayout {
    bool have_index = false, have_matrix = false;
    int32_t local_size = 0, local_size_x = 0, overlap_size = 0, n_halo = 0;
    std::vector<int32_t> l2g;                 // local_size_x + n_halo global ids
    HostCsr local;                            // local_size_x^2, local columns
    HostCsr iface;                            // reference layout: global columns
    std::vector<int32_t> nbr_in, nbr_out;
    std::vector<std::vector<int32_t>> get, put;   // global ids, ascending
    std::vector<int32_t> put_disp, get_disp;      // P+1
};

class Setup {
public:
    Setup(std::unique_ptr<RowSource> base, int32_t P, int32_t partition_kind,
          const uint32_t *part, int32_t overlap);

    int32_t N() const { return N_; }
    int32_t P() const { return P_; }
    int32_t overlap() const { return overlap_; }
    bool permuted() const { return permuted_; }
    const std::vector<int32_t> &first_row() const { return first_row_; }
    const std::vector<int32_t> &perm() const { return perm_; }
    const std::vector<int32_t> &iperm() const { return iperm_; }
    const RowSource &rows() const { return *view_; }

    // index sets of every subdomain (cheap: O(nnz) once); needed because the
    // put-lists of a subdomain are its neighbours' get-lists
    void build_index_sets();
    // local + interface matrices of one subdomain
    void build_matrices(int32_t rank);
    void release(int32_t rank);
    RankLayout &rank(int32_t r) { return ranks_[r]; }

    // interface matrix in compact numbering (rows = overlap rows only,
    // columns = g2l-1, entries in the reference's sorted-by-global-column order)
    void compact_interface(int32_t rank, HostCsr &out) const;

private:
    void index_set(int32_t me);
    std::unique_ptr<RowSource> base_, view_;
    int32_t N_ = 0, P_ = 1, overlap_ = 2;
    bool permuted_ = false;
    std::vector<int32_t> first_row_, perm_, iperm_, local_p_size_;
    std::vector<RankLayout> ranks_;
    bool have_index_ = false;
};

}  // namespace schwz_b200
