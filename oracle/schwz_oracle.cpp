// =============================================================================
// schwz_oracle.cpp — CPU restatement of schwarz-lib's restricted additive
// Schwarz (RAS) path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.  The product (schwarz-lib_b200/) never links or calls it.
//
// PARITY STATUS: pinned.  oracle/_ref is the reference's own source/*.cpp compiled unmodified
// (oracle/Makefile) against stand-ins for MPI, Ginkgo and the generated config header
// (oracle/ref_shim/); tests/test_ref_pinning.py, test_ref_pinning_random.py and
// test_precond_pinning.py require this restatement to reproduce that reference run bit for
// bit: partition, permutation, permuted matrix, overlap / halo numbering, local and interface
// matrices, halo lists, displacement tables, the iterate after every exchange, the residual
// histories and the stopping iteration - on the Laplacian configs, ani4_crop with the
// reference's own METIS call, seeded random matrices, and with the local preconditioners.
// What stays unpinned, by construction: the bit patterns of upstream Ginkgo's own kernels
// (Ginkgo is not under /root/reference; its arithmetic is restated from the call sites and from
// the upstream reference-executor kernels), and CHOLMOD / UMFPACK (absent:
// the factorised variants are pinned mathematically, L L^T = P A P^T and P A Q = L U).
// Earlier anchors stay: SURVEY.md Appendix E known answers, scipy cross-checks of the local
// solves, the fixed point of the iteration.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).
// =============================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using idx = int32_t;

struct Csr {
    idx nrows = 0, ncols = 0;
    std::vector<idx> rp, ci;
    std::vector<double> v;
    idx nnz() const { return rp.empty() ? 0 : rp[nrows]; }
};

int g_threads = 1;

// -----------------------------------------------------------------------------
// Ginkgo-semantics primitives (restated; Ginkgo is not in /root/reference).
// Call sites: Csr::apply  restricted_schwarz.cpp:1014, solve.cpp:834,1019,1079;
//             Dense::compute_norm2 solve.cpp:841,1069-1082.
// -----------------------------------------------------------------------------

// c = alpha*A*b + beta*c, row-wise sequential in stored column order
// (Ginkgo reference/omp executor advanced_spmv).
void spmv_adv(const Csr &A, double alpha, const double *b, double beta,
              double *c)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (idx row = 0; row < A.nrows; ++row) {
        double acc = c[row] * beta;
        for (idx k = A.rp[row]; k < A.rp[row + 1]; ++k) {
            acc += alpha * A.v[k] * b[A.ci[k]];
        }
        c[row] = acc;
    }
}

// c = A*b
void spmv(const Csr &A, const double *b, double *c)
{
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (idx row = 0; row < A.nrows; ++row) {
        double acc = 0.0;
        for (idx k = A.rp[row]; k < A.rp[row + 1]; ++k) {
            acc += A.v[k] * b[A.ci[k]];
        }
        c[row] = acc;
    }
}

double dot(const double *a, const double *b, int64_t n)
{
    double s = 0.0;
#pragma omp parallel for num_threads(g_threads) schedule(static) reduction(+ : s)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

double norm2(const double *a, int64_t n) { return std::sqrt(dot(a, a, n)); }

// Csr::sort_by_column_index (restricted_schwarz.cpp:297-298,
// initialization.cpp:212): sort every row by column index, values follow.
void sort_by_column_index(Csr &A)
{
    std::vector<std::pair<idx, double>> tmp;
    for (idx r = 0; r < A.nrows; ++r) {
        tmp.clear();
        for (idx k = A.rp[r]; k < A.rp[r + 1]; ++k)
            tmp.emplace_back(A.ci[k], A.v[k]);
        std::stable_sort(tmp.begin(), tmp.end(),
                         [](const auto &x, const auto &y) {
                             return x.first < y.first;
                         });
        idx k = A.rp[r];
        for (auto &e : tmp) {
            A.ci[k] = e.first;
            A.v[k] = e.second;
            ++k;
        }
    }
}

// -----------------------------------------------------------------------------
// A1. 2-D 5-pt Laplacian, source/initialization.cpp:214-265 (literal,
// including the exclusion-set walk; the one-past-the-end read of
// exclusion_set at :250-252 is bounds-guarded, result identical).
// -----------------------------------------------------------------------------
Csr laplacian2d(int64_t n)
{
    Csr A;
    const int64_t N = n * n;
    A.nrows = A.ncols = (idx)N;
    A.rp.assign(N + 1, 0);
    A.ci.reserve(5 * N);
    A.v.reserve(5 * N);
    // :225-242 exclusion set: linearised (row*N + col) of the wrap-around
    // couplings (k*n, k*n-1) and (k*n-1, k*n), k = 1..n-1.
    std::vector<uint64_t> excl;
    for (int64_t i = 2; i < N; ++i) {
        uint64_t index = (uint64_t)(i - 1) * (uint64_t)n;
        if (index < (uint64_t)N) {  // index*index < N*N  (:233)
            excl.push_back(index * (uint64_t)N + (index - 1));
            excl.push_back((index - 1) * (uint64_t)N + index);
        } else {
            break;  // monotone in i: nothing further qualifies
        }
    }
    std::sort(excl.begin(), excl.end());
    // :227-230 std::map iterates offsets ascending: -n, -1, 0, +1, +n
    const int64_t ofs[5] = {-n, -1, 0, 1, n};
    const double val[5] = {-1, -1, 4, -1, -1};
    size_t cur = 0;
    for (int64_t i = 0; i < N; ++i) {
        for (int s = 0; s < 5; ++s) {
            // size_type (unsigned 64-bit) wrap-around arithmetic as at :252
            uint64_t lin = (uint64_t)i * (uint64_t)N + (uint64_t)(i + ofs[s]);
            bool in_excl = cur < excl.size() && excl[cur] == lin;
            if (0 <= i + ofs[s] && i + ofs[s] < N && !in_excl) {
                A.v.push_back(val[s]);
                A.ci.push_back((idx)(i + ofs[s]));
            }
            if (in_excl) ++cur;
        }
        A.rp[i + 1] = (idx)A.ci.size();
    }
    return A;
}

// Extension (the reference has no 3-D generator): 7-pt Laplacian on
// an n^3 grid, natural ordering, values {-1 x6, 6}, columns ascending.
Csr laplacian3d(int64_t n)
{
    Csr A;
    const int64_t N = n * n * n;
    A.nrows = A.ncols = (idx)N;
    A.rp.assign(N + 1, 0);
    for (int64_t z = 0; z < n; ++z)
        for (int64_t y = 0; y < n; ++y)
            for (int64_t x = 0; x < n; ++x) {
                int64_t i = (z * n + y) * n + x;
                if (z > 0) { A.ci.push_back((idx)(i - n * n)); A.v.push_back(-1); }
                if (y > 0) { A.ci.push_back((idx)(i - n)); A.v.push_back(-1); }
                if (x > 0) { A.ci.push_back((idx)(i - 1)); A.v.push_back(-1); }
                A.ci.push_back((idx)i); A.v.push_back(6);
                if (x < n - 1) { A.ci.push_back((idx)(i + 1)); A.v.push_back(-1); }
                if (y < n - 1) { A.ci.push_back((idx)(i + n)); A.v.push_back(-1); }
                if (z < n - 1) { A.ci.push_back((idx)(i + n * n)); A.v.push_back(-1); }
                A.rp[i + 1] = (idx)A.ci.size();
            }
    return A;
}

// -----------------------------------------------------------------------------
// A2. include/partition_tools.hpp:70-94 PartitionRegular2D (literal, incl. the
// truncating sqrt that leaves ranks >= sq_p^2 empty).
// -----------------------------------------------------------------------------
void partition_regular2d(int64_t N, int P, uint32_t *pi)
{
    int sq_n = static_cast<int>(std::sqrt((double)N));
    int sq_partn = static_cast<int>(std::sqrt((double)P));
    for (int j1 = 0; j1 < sq_partn; ++j1) {
        int offset2 = (int)(j1 * sq_partn * std::pow(sq_n / sq_partn, 2));
        for (int j2 = 0; j2 < sq_partn; ++j2) {
            int my_id = sq_partn * j1 + j2;
            int offset1 = (j2)*sq_n / sq_partn;
            for (int i1 = 0; i1 < sq_n / sq_partn; ++i1)
                for (int i2 = 0; i2 < sq_n / sq_partn; ++i2)
                    pi[offset2 + offset1 + (i1 * sq_n) + i2] = my_id;
        }
    }
}

// Rectangular extension of the rule above (NOT in the reference, whose sqrt truncation leaves
// subdomains without rows unless P is a perfect square): the
// n x n grid is cut into px block-rows x py block-columns, px * py == P, subdomain id =
// py * j1 + j2 exactly as include/partition_tools.hpp:84 numbers them, block-row j1 = grid rows
// [j1*n/px, (j1+1)*n/px), block-column j2 = grid columns [j2*n/py, (j2+1)*n/py) (integer
// division: every row gets an owner when n is not divisible).  px = py = 0 picks the most
// square factorisation with px <= py (8 -> 2 x 4, 4 -> 2 x 2, 64 -> 8 x 8).  For a perfect
// square P and n divisible by sqrt(P) it reproduces partition_regular2d bit for bit.
void regular2d_factors(int P, int &px, int &py)
{
    if (px > 0 && py > 0) return;
    px = 1;
    for (int d = 1; (int64_t)d * d <= P; ++d)
        if (P % d == 0) px = d;
    py = P / px;
}

int partition_regular2d_rect(int64_t N, int P, int px, int py, uint32_t *pi)
{
    const int64_t n = (int64_t)std::llround(std::sqrt((double)N));
    if (n * n != N) return -1;
    regular2d_factors(P, px, py);
    if ((int64_t)px * py != P || px > n || py > n) return -1;
    for (int j1 = 0; j1 < px; ++j1)
        for (int64_t r = j1 * n / px; r < (j1 + 1) * n / px; ++r)
            for (int j2 = 0; j2 < py; ++j2)
                for (int64_t c = j2 * n / py; c < (j2 + 1) * n / py; ++c)
                    pi[r * n + c] = (uint32_t)(py * j1 + j2);
    return 0;
}

// -----------------------------------------------------------------------------
// Per-subdomain ("rank") state.
// -----------------------------------------------------------------------------
struct Options {
    double tolerance = 1e-6;        // metadata.tolerance (--set_tol)
    double local_tol = 1e-12;       // metadata.local_solver_tolerance
    int local_max_iters = -1;       // metadata.local_max_iters
    int max_iters = 100;            // metadata.max_iters (--num_iters)
    int non_symmetric = 0;          // settings.non_symmetric_matrix -> GMRES
    int restart_iter = 1;           // settings.restart_iter
    int local_solver = 2;           // 2 iterative-ginkgo, 1 direct-ginkgo
    int enable_onesided = 0;        // comm_settings.enable_onesided
    int enable_put = 0;             // remote_comm_type == put
    int enable_one_by_one = 0;
    int enable_global_check = 0;    // convergence_settings.enable_global_check
    int conv_tree = 1;              // enable_global_simple_tree
    int conv_decentralized = 0;     // enable_decentralized_leader_election
    int enable_accumulate = 0;
    int iter_offset = 0;            // enable_global_check_iter_offset
    int overlap = 2;                // settings.overlap (for update_boundary)
    int use_mixed_precision = 0;    // settings.use_mixed_precision with MixedValueType = float
    int local_precond = 0;          // metadata.local_precond: 0 null, 1 block-jacobi, 2 ilu, 3 isai
    int precond_max_block_size = 16;  // metadata.precond_max_block_size
    int local_factorization = 0;    // settings.factorization: 0 "cholmod" (LL^T), 1 "umfpack" (LU)
};

struct Precond;
struct Rank {
    std::shared_ptr<Precond> precond;   // local preconditioner (solve.cpp:486-652)
    std::vector<idx> g2l, l2g;
    idx local_size = 0, local_size_x = 0, overlap_size = 0, n_halo = 0;
    std::vector<idx> overlap_row;
    Csr local, iface;
    std::vector<idx> nbr_in, nbr_out;
    std::vector<std::vector<idx>> get, put;
    std::vector<idx> put_disp, get_disp;
    // run state
    std::vector<double> x, local_rhs, local_sol, init_guess, work, recv_buf,
        send_buf;
    double resnorm = -1.0, resnorm0 = -1.0, gres = 0.0, gres0 = -1.0;
    int num_converged = 0;
    bool finished = false;
    int finished_iter = -1;
    std::vector<idx> conv, conv_sent, conv_local;
    std::vector<double> l_res;
    std::vector<double> res_hist;     // local_residual_vector_out
    std::vector<double> gres_hist;    // global residual norm per iteration
    std::vector<int> local_iter_hist; // inner iterations per outer iteration
    // direct solver
    Csr L, U;
    std::vector<idx> fperm, fperm_col;   // local_perm (P) and local_inv_perm (Q; empty = P)
    int last_local_iters = 0;
};

struct Problem {
    idx N = 0;
    int P = 1;
    int overlap = 2;
    int permuted = 0;
    Csr g;  // (permuted) global matrix
    std::vector<idx> first_row, perm, iperm;
    std::vector<Rank> ranks;
    std::vector<double> rhs;
    Options opt;
    int iter_count = 0;
    bool solver_ready = false;
};

// -----------------------------------------------------------------------------
// A3. SolverRAS::setup_local_matrices, source/restricted_schwarz.cpp:56-304.
// -----------------------------------------------------------------------------
void setup_partition(Problem &pb, const Csr &A, int partition_kind,
                     const uint32_t *part)
{
    const idx N = pb.N;
    const int P = pb.P;
    // :84, :97-102 default 1-D split
    const idx nb = (N + P - 1) / P;
    pb.first_row.assign(P + 1, 0);
    std::vector<idx> local_p_size(P);
    for (int p = 0; p < P; ++p) {
        local_p_size[p] = std::min<idx>(N - pb.first_row[p], nb);
        pb.first_row[p + 1] = pb.first_row[p] + local_p_size[p];
    }
    pb.permuted = 0;
    pb.g = A;
    // :105-152 metis / regular2d: stable counting-sort permutation + symmetric
    // permutation of the global matrix (rows in new order, columns renamed,
    // NOT re-sorted).
    if (partition_kind == 1) {
        pb.perm.assign(N, 0);
        pb.iperm.assign(N, 0);
        if (P > 1) {
            std::fill(local_p_size.begin(), local_p_size.end(), 0);
            for (idx i = 0; i < N; ++i) local_p_size[part[i]]++;
            pb.first_row[0] = 0;
            for (int p = 0; p < P; ++p)
                pb.first_row[p + 1] = pb.first_row[p] + local_p_size[p];
            for (idx i = 0; i < N; ++i) {
                pb.perm[pb.first_row[part[i]]] = i;
                pb.first_row[part[i]]++;
            }
            for (int p = P; p > 0; --p) pb.first_row[p] = pb.first_row[p - 1];
            pb.first_row[0] = 0;
            for (idx i = 0; i < N; ++i) pb.iperm[pb.perm[i]] = i;
        } else {
            // P == 1: the reference reads uninitialised permutation arrays
            // (SURVEY Appendix A.3); the oracle uses the identity.
            std::iota(pb.perm.begin(), pb.perm.end(), 0);
            std::iota(pb.iperm.begin(), pb.iperm.end(), 0);
        }
        Csr t;
        t.nrows = t.ncols = N;
        t.rp.assign(N + 1, 0);
        t.ci.resize(A.nnz());
        t.v.resize(A.nnz());
        idx nnz = 0;
        for (idx row = 0; row < N; ++row) {
            for (idx c = A.rp[pb.perm[row]]; c < A.rp[pb.perm[row] + 1]; ++c) {
                t.ci[nnz] = pb.iperm[A.ci[c]];
                t.v[nnz] = A.v[c];
                ++nnz;
            }
            t.rp[row + 1] = nnz;
        }
        pb.g = std::move(t);
        pb.permuted = 1;
    }
    for (int p = 0; p < P; ++p) {
        // metadata.local_size = local_p_size[my_rank]  (:181)
        pb.ranks[p].local_size = local_p_size[p];
    }
}

void setup_local_matrices(Problem &pb, int me)
{
    Rank &R = pb.ranks[me];
    const Csr &g = pb.g;
    const idx N = pb.N;
    const auto &first_row = pb.first_row;
    // :155-164
    R.g2l.assign(N, 0);
    R.l2g.assign(N, 0);
    idx num = 0;
    for (idx i = first_row[me]; i < first_row[me + 1]; ++i) {
        R.g2l[i] = 1 + num;
        R.l2g[num] = i;
        ++num;
    }
    // :166-180 overlap BFS, (overlap-1) layers
    idx old = 0;
    for (int k = 1; k < pb.overlap; ++k) {
        idx now = num;
        for (idx i = old; i < now; ++i) {
            for (idx j = g.rp[R.l2g[i]]; j < g.rp[R.l2g[i] + 1]; ++j) {
                if (R.g2l[g.ci[j]] == 0) {
                    R.l2g[num] = g.ci[j];
                    R.g2l[g.ci[j]] = 1 + num;
                    ++num;
                }
            }
        }
        old = now;
    }
    // :181-192
    R.local_size_x = num;
    R.overlap_size = num - R.local_size;
    R.overlap_row.assign(R.l2g.begin() + R.local_size,
                         R.l2g.begin() + R.local_size + R.overlap_size);
    // :194-216 count
    idx nnz_local = 0, nnz_interface = 0;
    for (idx i = first_row[me]; i < first_row[me + 1]; ++i)
        for (idx j = g.rp[i]; j < g.rp[i + 1]; ++j)
            if (R.g2l[g.ci[j]] != 0) ++nnz_local;
    for (idx k = 0; k < R.overlap_size; ++k) {
        idx t = R.overlap_row[k];
        for (idx j = g.rp[t]; j < g.rp[t + 1]; ++j) {
            if (R.g2l[g.ci[j]] != 0)
                ++nnz_local;
            else
                ++nnz_interface;
        }
    }
    // :218-236
    Csr &Lm = R.local;
    Lm.nrows = Lm.ncols = R.local_size_x;
    Lm.rp.assign(R.local_size_x + 1, 0);
    Lm.ci.assign(nnz_local, 0);
    Lm.v.assign(nnz_local, 0.0);
    Csr &Im = R.iface;
    const bool have_iface = nnz_interface > 0;
    if (have_iface) {
        Im.nrows = Im.ncols = R.local_size_x;  // declared lsx x lsx (:227-229)
        Im.rp.assign(R.local_size_x + 1, 0);
        Im.ci.assign(nnz_interface, 0);
        Im.v.assign(nnz_interface, 0.0);
    } else {
        Im = Csr();  // empty 0x0 (:231)
        Im.rp.assign(1, 0);
    }
    // :238-260 own rows
    num = 0;
    nnz_local = 0;
    for (idx i = first_row[me]; i < first_row[me + 1]; ++i) {
        for (idx j = g.rp[i]; j < g.rp[i + 1]; ++j) {
            if (R.g2l[g.ci[j]] != 0) {
                Lm.ci[nnz_local] = R.g2l[g.ci[j]] - 1;
                Lm.v[nnz_local] = g.v[j];
                ++nnz_local;
            }
        }
        if (have_iface) Im.rp[num + 1] = 0;
        Lm.rp[num + 1] = nnz_local;
        ++num;
    }
    // :262-284 overlap rows (only when nnz_interface > 0 — quirk kept)
    if (have_iface) {
        nnz_interface = 0;
        for (idx k = 0; k < R.overlap_size; ++k) {
            idx t = R.overlap_row[k];
            for (idx j = g.rp[t]; j < g.rp[t + 1]; ++j) {
                if (R.g2l[g.ci[j]] != 0) {
                    Lm.ci[nnz_local] = R.g2l[g.ci[j]] - 1;
                    Lm.v[nnz_local] = g.v[j];
                    ++nnz_local;
                } else {
                    Im.ci[nnz_interface] = g.ci[j];
                    Im.v[nnz_interface] = g.v[j];
                    ++nnz_interface;
                }
            }
            Lm.rp[num + 1] = nnz_local;
            Im.rp[num + 1] = nnz_interface;
            ++num;
        }
    } else {
        // overlap rows were counted into nnz_local at :207-216 but never
        // filled; only reachable when the overlap swallows everything / P == 1
        // (SURVEY Appendix D).  Fill them so the matrix is well formed.
        for (idx k = 0; k < R.overlap_size; ++k) {
            idx t = R.overlap_row[k];
            for (idx j = g.rp[t]; j < g.rp[t + 1]; ++j) {
                if (R.g2l[g.ci[j]] != 0) {
                    Lm.ci[nnz_local] = R.g2l[g.ci[j]] - 1;
                    Lm.v[nnz_local] = g.v[j];
                    ++nnz_local;
                }
            }
            Lm.rp[R.local_size + k + 1] = nnz_local;
        }
    }
    // :285-295 halo sweep: one more BFS layer over i in old..now-1 where `old`
    // is the start of the last overlap layer and `now` = rows filled above.
    idx now = num;
    idx hnum = R.local_size_x;  // appended beyond local_size_x
    // NB: in the reference `num` is reused as the append cursor and equals
    // `now` (= local_size_x when the interface loop ran).
    hnum = num;
    for (idx i = old; i < now; ++i) {
        for (idx j = g.rp[R.l2g[i]]; j < g.rp[R.l2g[i] + 1]; ++j) {
            if (R.g2l[g.ci[j]] == 0) {
                R.l2g[hnum] = g.ci[j];
                R.g2l[g.ci[j]] = 1 + hnum;
                ++hnum;
            }
        }
    }
    R.n_halo = hnum - num;
    // :297-298
    sort_by_column_index(Lm);
    if (have_iface) sort_by_column_index(Im);
}

// -----------------------------------------------------------------------------
// A4. SolverRAS::setup_comm_buffers, source/restricted_schwarz.cpp:308-604.
// The MPI handshake (:400-472) hands the owner the requester's list verbatim.
// -----------------------------------------------------------------------------
void setup_comm_buffers(Problem &pb)
{
    const int P = pb.P;
    for (int me = 0; me < P; ++me) {
        Rank &R = pb.ranks[me];
        R.nbr_in.clear();
        R.get.clear();
        for (int p = 0; p < P; ++p) {
            if (p == me) continue;
            std::vector<idx> lst;
            for (idx i = pb.first_row[p]; i < pb.first_row[p + 1]; ++i)
                if (R.g2l[i] != 0) lst.push_back(i);  // :362-368
            if (!lst.empty()) {
                R.nbr_in.push_back(p);
                R.get.push_back(std::move(lst));
            }
        }
    }
    for (int me = 0; me < P; ++me) {
        Rank &R = pb.ranks[me];
        R.nbr_out.clear();
        R.put.clear();
        for (int p = 0; p < P; ++p) {  // :426-472 ascending p
            if (p == me) continue;
            const Rank &Q = pb.ranks[p];
            for (size_t j = 0; j < Q.nbr_in.size(); ++j) {
                if (Q.nbr_in[j] == me) {
                    R.nbr_out.push_back(p);
                    R.put.push_back(Q.get[j]);
                }
            }
        }
        size_t num_recv = 0, num_send = 0;
        for (auto &l : R.get) num_recv += l.size();
        for (auto &l : R.put) num_send += l.size();
        R.recv_buf.assign(std::max<size_t>(num_recv, 1), 0.0);  // :481-603
        R.send_buf.assign(std::max<size_t>(num_send, 1), 0.0);
    }
}

// -----------------------------------------------------------------------------
// A5. SolverRAS::setup_windows displacement tables,
// source/restricted_schwarz.cpp:624-658 (MPI_Alltoall transposes the tables).
// -----------------------------------------------------------------------------
void setup_windows(Problem &pb)
{
    const int P = pb.P;
    std::vector<std::vector<idx>> in_pref(P), out_pref(P);
    for (int me = 0; me < P; ++me) {
        Rank &R = pb.ranks[me];
        std::vector<idx> t(P + 1, 0);
        for (size_t j = 0; j < R.nbr_in.size(); ++j)
            t[R.nbr_in[j] + 1] = (idx)R.get[j].size();
        for (int j = 0; j < P; ++j) t[j + 1] += t[j];
        in_pref[me] = t;
        std::vector<idx> u(P + 1, 0);
        for (size_t j = 0; j < R.nbr_out.size(); ++j)
            u[R.nbr_out[j] + 1] = (idx)R.put[j].size();
        for (int j = 0; j < P; ++j) u[j + 1] += u[j];
        out_pref[me] = u;
    }
    for (int me = 0; me < P; ++me) {
        Rank &R = pb.ranks[me];
        R.put_disp.assign(P + 1, 0);
        R.get_disp.assign(P + 1, 0);
        for (int q = 0; q < P; ++q) {
            // Alltoall: recv[q] = send_of_q[me]
            R.put_disp[q] = in_pref[q][me];
            R.get_disp[q] = out_pref[q][me];
        }
        // :676-691 convergence arrays
        R.conv.assign(P, 0);
        R.conv_sent.assign(P, 0);
        R.conv_local.assign(P, 0);
        // solve.cpp:220-221 window over local_residual_vector (uninitialised in
        // the reference; DBL_MAX here, logging only — SURVEY Appendix D)
        R.l_res.assign(std::max(P, pb.opt.max_iters + 1), DBL_MAX);
    }
}

// -----------------------------------------------------------------------------
// A6. Initialize::setup_vectors + SolverTools::extract_local_vector,
// source/initialization.cpp:333-359, include/solver_tools.hpp:101-116.
// -----------------------------------------------------------------------------
void extract_local_vector(const Problem &pb, int me, double *sub,
                          const double *vec)
{
    const Rank &R = pb.ranks[me];
    const idx first = pb.first_row[me];
    for (idx i = 0; i < R.local_size; ++i) sub[i] = vec[first + i];
    // Gather(copy): into[i] = from[idx[i]]  (include/gather.hpp:86-92)
    for (idx k = 0; k < R.overlap_size; ++k)
        sub[R.local_size + k] = vec[R.overlap_row[k]];
}

void setup_vectors(Problem &pb)
{
    for (int me = 0; me < pb.P; ++me) {
        Rank &R = pb.ranks[me];
        R.local_rhs.assign(R.local_size_x, 0.0);
        extract_local_vector(pb, me, R.local_rhs.data(), pb.rhs.data());
        R.local_sol.assign(R.local_size_x, 0.0);  // F8: zero-filled
        R.x.assign(pb.N, 0.0);
        R.init_guess.assign(R.local_size_x, 0.0);
        R.work.assign(2 * (size_t)R.local_size_x, 0.0);
        R.resnorm = R.resnorm0 = -1.0;
        R.gres = 0.0;
        R.gres0 = -1.0;
        R.num_converged = 0;
        R.finished = false;
        R.finished_iter = -1;
        R.res_hist.clear();
        R.gres_hist.clear();
        R.local_iter_hist.clear();
    }
    pb.iter_count = 0;
}

// -----------------------------------------------------------------------------
// Local preconditioners of the iterative local solve (call sites
// source/solve.cpp:486-652).  The arithmetic lives in Ginkgo, which is not in
// the tree: restated from the upstream reference-executor kernels
// [upstream-memory] and pinned bit-for-bit to the stand-in that oracle/_ref
// links (tests/test_precond_pinning.py).
//   kind 1  block-Jacobi : preconditioner::Jacobi(max_block_size), :496-505, :581-589
//   kind 2  ILU          : factorization::ParIlu + preconditioner::Ilu<LowerTrs,
//                          UpperTrs>, :513-532, :598-617
//   kind 3  ISAI         : preconditioner::Ilu<LowerIsai, UpperIsai>, :540-556, :625-638
// -----------------------------------------------------------------------------
struct Precond {
    int kind = 0;
    // block-Jacobi
    std::vector<idx> block_ptrs;
    std::vector<size_t> block_off;
    std::vector<double> blocks;   // inverse blocks, column-major, ld = block size
    // ILU factors and their sparse approximate inverses
    Csr L, U, Li, Ui;
    mutable std::vector<double> tmp;
};

void lower_trs(const Csr &L, const double *b, double *x);
void upper_trs(const Csr &U, const double *b, double *x);
Csr transpose(const Csr &A);

// jacobi::find_blocks: natural blocks (equal column patterns of neighbouring
// rows, at most max_bs rows), then greedy agglomeration up to max_bs.
static void jacobi_find_blocks(const Csr &A, idx max_bs, std::vector<idx> &bp)
{
    const idx n = A.nrows;
    bp.assign(1, 0);
    if (n == 0) return;
    std::vector<idx> nat(1, 0);
    idx cur = 1;
    for (idx i = 1; i < n; ++i) {
        const idx la = A.rp[i] - A.rp[i - 1], lb = A.rp[i + 1] - A.rp[i];
        bool same = (la == lb);
        for (idx k = 0; same && k < lb; ++k)
            same = A.ci[A.rp[i - 1] + k] == A.ci[A.rp[i] + k];
        if (cur < max_bs && same) {
            ++cur;
        } else {
            nat.push_back(nat.back() + cur);
            cur = 1;
        }
    }
    nat.push_back(nat.back() + cur);
    const size_t nn = nat.size() - 1;
    cur = nat[1] - nat[0];
    for (size_t i = 1; i < nn; ++i) {
        const idx bs = nat[i + 1] - nat[i];
        if (cur + bs <= max_bs) {
            cur += bs;
        } else {
            bp.push_back(bp.back() + cur);
            cur = bs;
        }
    }
    bp.push_back(bp.back() + cur);
}

// jacobi::generate: dense diagonal block, in-place Gauss-Jordan with implicit
// row pivoting (first row of largest magnitude), columns un-permuted on store.
static void jacobi_invert_block(std::vector<double> &B, idx bs, double *out)
{
    std::vector<idx> perm(bs);
    std::iota(perm.begin(), perm.end(), 0);
    auto at = [&](idx i, idx j) -> double & { return B[(size_t)i * bs + j]; };
    for (idx k = 0; k < bs; ++k) {
        idx piv = k;
        for (idx i = k + 1; i < bs; ++i)
            if (std::fabs(at(piv, k)) < std::fabs(at(i, k))) piv = i;
        if (piv != k) {
            for (idx j = 0; j < bs; ++j) std::swap(at(k, j), at(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        const double d = at(k, k);
        for (idx i = 0; i < bs; ++i) at(i, k) /= -d;
        at(k, k) = 0.0;
        for (idx i = 0; i < bs; ++i) {
            const double f = at(i, k);
            for (idx j = 0; j < bs; ++j) at(i, j) += f * at(k, j);
        }
        for (idx j = 0; j < bs; ++j) at(k, j) /= d;
        at(k, k) = 1.0 / d;
    }
    for (idx i = 0; i < bs; ++i)
        for (idx j = 0; j < bs; ++j) out[(size_t)perm[j] * bs + i] = at(i, j);
}

// factorization::ParIlu on the reference executor: L = strict lower part + unit
// diagonal, U = diagonal + strict upper part (zero / missing diagonal -> 1),
// then one sequential row-major sweep of the fixed-point update, which is
// ILU(0).  Per entry: s = a_rc - sum_k l_rk u_kc over the matching k in
// ascending order INCLUDING the term that holds the unknown, which is then
// added back (upstream's "last_operation" idiom - kept, it is visible in the
// last bit).
static void par_ilu(const Csr &A0, Csr &L, Csr &U)
{
    Csr A = A0;
    sort_by_column_index(A);
    const idx n = A.nrows;
    // explicit diagonal
    Csr D;
    D.nrows = D.ncols = n;
    D.rp.assign(n + 1, 0);
    for (idx r = 0; r < n; ++r) {
        bool placed = false;
        for (idx k = A.rp[r]; k < A.rp[r + 1]; ++k) {
            if (!placed && A.ci[k] > r) {
                D.ci.push_back(r);
                D.v.push_back(0.0);
                placed = true;
            }
            if (A.ci[k] == r) placed = true;
            D.ci.push_back(A.ci[k]);
            D.v.push_back(A.v[k]);
        }
        if (!placed) {
            D.ci.push_back(r);
            D.v.push_back(0.0);
        }
        D.rp[r + 1] = (idx)D.ci.size();
    }
    L = Csr();
    Csr Uc;   // U by columns: row c of Uc = column c of U, row indices ascending
    L.nrows = L.ncols = Uc.nrows = Uc.ncols = n;
    L.rp.assign(n + 1, 0);
    std::vector<idx> ucount(n, 0);
    for (idx r = 0; r < n; ++r)
        for (idx k = D.rp[r]; k < D.rp[r + 1]; ++k)
            if (D.ci[k] >= r) ucount[D.ci[k]]++;
    Uc.rp.assign(n + 1, 0);
    for (idx c = 0; c < n; ++c) Uc.rp[c + 1] = Uc.rp[c] + ucount[c];
    Uc.ci.resize(Uc.rp[n]);
    Uc.v.resize(Uc.rp[n]);
    std::vector<idx> ucur(Uc.rp.begin(), Uc.rp.end() - 1);
    for (idx r = 0; r < n; ++r) {
        for (idx k = D.rp[r]; k < D.rp[r + 1]; ++k) {
            const idx c = D.ci[k];
            if (c < r) {
                L.ci.push_back(c);
                L.v.push_back(D.v[k]);
            } else {
                const double val = (c == r && D.v[k] == 0.0) ? 1.0 : D.v[k];
                Uc.ci[ucur[c]] = r;
                Uc.v[ucur[c]] = val;
                ucur[c]++;
            }
        }
        L.ci.push_back(r);
        L.v.push_back(1.0);
        L.rp[r + 1] = (idx)L.ci.size();
    }
    for (idx r = 0; r < n; ++r)
        for (idx e = D.rp[r]; e < D.rp[r + 1]; ++e) {
            const idx c = D.ci[e];
            idx a = L.rp[r], b = Uc.rp[c];
            double s = D.v[e], last = 0.0;
            while (a < L.rp[r + 1] && b < Uc.rp[c + 1]) {
                const idx ka = L.ci[a], kb = Uc.ci[b];
                if (ka == kb) {
                    last = L.v[a] * Uc.v[b];
                    s -= last;
                } else {
                    last = 0.0;
                }
                if (ka <= kb) ++a;
                if (kb <= ka) ++b;
            }
            s += last;
            if (r > c) {
                const double w = s / Uc.v[Uc.rp[c + 1] - 1];
                if (std::isfinite(w)) L.v[a - 1] = w;
            } else if (std::isfinite(s)) {
                Uc.v[b - 1] = s;
            }
        }
    U = transpose(Uc);
}

// isai::generate_tri_inverse, sparsity power 1: row i of M ~ T^-1 on the
// pattern J of row i of T, from the dense system T(J,J)^T m = e.
static void isai_tri(const Csr &T0, bool lower, Csr &M)
{
    Csr T = T0;
    sort_by_column_index(T);
    M = T;
    std::vector<double> tri, m;
    for (idx row = 0; row < T.nrows; ++row) {
        const idx b = T.rp[row];
        const int sz = (int)(T.rp[row + 1] - b);
        if (sz == 0) continue;
        tri.assign((size_t)sz * sz, 0.0);
        for (int i = 0; i < sz; ++i) {
            const idx r2 = T.ci[b + i];
            idx ka = T.rp[r2], kb = b;
            while (ka < T.rp[r2 + 1] && kb < T.rp[row + 1]) {
                if (T.ci[ka] == T.ci[kb]) {
                    tri[(size_t)i * sz + (kb - b)] = T.v[ka];
                    ++ka;
                    ++kb;
                } else if (T.ci[ka] < T.ci[kb]) {
                    ++ka;
                } else {
                    ++kb;
                }
            }
        }
        m.assign(sz, 0.0);
        if (lower) {
            m[sz - 1] = 1.0;
            for (int c = sz - 1; c >= 0; --c) {
                const double t = m[c] / tri[(size_t)c * sz + c];
                m[c] = t;
                for (int r = c - 1; r >= 0; --r) m[r] -= t * tri[(size_t)c * sz + r];
            }
        } else {
            m[0] = 1.0;
            for (int c = 0; c < sz; ++c) {
                const double t = m[c] / tri[(size_t)c * sz + c];
                m[c] = t;
                for (int r = c + 1; r < sz; ++r) m[r] -= t * tri[(size_t)c * sz + r];
            }
        }
        bool finite = true;
        for (int i = 0; i < sz; ++i) finite = finite && std::isfinite(m[i]);
        for (int i = 0; i < sz; ++i)
            M.v[b + i] = finite ? m[i] : (T.ci[b + i] == row ? 1.0 : 0.0);
    }
}

void precond_generate(const Csr &A, int kind, int max_block_size, Precond &M)
{
    M = Precond();
    M.kind = kind;
    if (kind == 1) {
        jacobi_find_blocks(A, (idx)max_block_size, M.block_ptrs);
        const size_t nb = M.block_ptrs.size() - 1;
        M.block_off.assign(nb + 1, 0);
        for (size_t b = 0; b < nb; ++b) {
            const size_t bs = (size_t)(M.block_ptrs[b + 1] - M.block_ptrs[b]);
            M.block_off[b + 1] = M.block_off[b] + bs * bs;
        }
        M.blocks.assign(M.block_off[nb], 0.0);
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 64)
        for (int64_t b = 0; b < (int64_t)nb; ++b) {
            const idx r0 = M.block_ptrs[b], bs = M.block_ptrs[b + 1] - r0;
            std::vector<double> B((size_t)bs * bs, 0.0);
            for (idx i = 0; i < bs; ++i)
                for (idx k = A.rp[r0 + i]; k < A.rp[r0 + i + 1]; ++k)
                    if (A.ci[k] >= r0 && A.ci[k] < r0 + bs)
                        B[(size_t)i * bs + (A.ci[k] - r0)] = A.v[k];
            jacobi_invert_block(B, bs, M.blocks.data() + M.block_off[b]);
        }
    } else if (kind == 2 || kind == 3) {
        par_ilu(A, M.L, M.U);
        if (kind == 3) {
            isai_tri(M.L, true, M.Li);
            isai_tri(M.U, false, M.Ui);
        }
    }
    M.tmp.assign(A.nrows, 0.0);
}

// z = M^-1 r
void precond_apply(const Precond &M, const double *r, double *z, idx n)
{
    switch (M.kind) {
    case 1: {
        const int64_t nb = (int64_t)M.block_ptrs.size() - 1;
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int64_t b = 0; b < nb; ++b) {
            const idx r0 = M.block_ptrs[b], bs = M.block_ptrs[b + 1] - r0;
            const double *inv = M.blocks.data() + M.block_off[b];
            for (idx i = 0; i < bs; ++i) z[r0 + i] = 0.0;
            for (idx inner = 0; inner < bs; ++inner)
                for (idx i = 0; i < bs; ++i)
                    z[r0 + i] += inv[(size_t)inner * bs + i] * r[r0 + inner];
        }
        break;
    }
    case 2:
        lower_trs(M.L, r, M.tmp.data());
        upper_trs(M.U, M.tmp.data(), z);
        break;
    case 3:
        spmv(M.Li, r, M.tmp.data());
        spmv(M.Ui, M.tmp.data(), z);
        break;
    default: std::copy(r, r + n, z);
    }
}

// -----------------------------------------------------------------------------
// A12. Local iterative solve: Ginkgo Cg / Gmres semantics (SURVEY Appendix F;
// call sites source/solve.cpp:469-478, 486-652, 746-754;
// include/solver_tools.hpp:91-98).  Stop = Combined(Iteration(max),
// ResidualNormReduction(local_tol)) with the reduction measured against the
// residual of the initial guess.
// -----------------------------------------------------------------------------
int cg_solve(const Csr &A, const double *b, double *x, int max_iters,
             double factor, const Precond *M = nullptr)
{
    const idx n = A.nrows;
    std::vector<double> r(b, b + n), z(n, 0.0), p(n, 0.0), q(n, 0.0);
    spmv_adv(A, -1.0, x, 1.0, r.data());  // r = b - A x
    const double r0 = norm2(r.data(), n);
    double rho = 0.0, prev_rho = 1.0;
    int iter = -1;
    while (true) {
        if (M && M->kind)
            precond_apply(*M, r.data(), z.data(), n);  // z = M^-1 r
        else
            std::copy(r.begin(), r.end(), z.begin());
        rho = dot(r.data(), z.data(), n);
        ++iter;
        const double tau = norm2(r.data(), n);
        if (iter >= max_iters || tau < factor * r0) break;
        // step_1: p = z + (rho/prev_rho) p   (prev_rho == 0 -> p = z)
        if (prev_rho == 0.0) {
            std::copy(z.begin(), z.end(), p.begin());
        } else {
            const double t = rho / prev_rho;
#pragma omp parallel for num_threads(g_threads) schedule(static)
            for (idx i = 0; i < n; ++i) p[i] = z[i] + t * p[i];
        }
        spmv(A, p.data(), q.data());
        const double beta = dot(p.data(), q.data(), n);
        // step_2: x += (rho/beta) p ; r -= (rho/beta) q   (beta == 0 -> skip)
        if (beta != 0.0) {
            const double t = rho / beta;
#pragma omp parallel for num_threads(g_threads) schedule(static)
            for (idx i = 0; i < n; ++i) {
                x[i] += t * p[i];
                r[i] -= t * q[i];
            }
        }
        std::swap(prev_rho, rho);
    }
    return iter;
}

// Restarted GMRES(m), modified Gram-Schmidt, Givens rotations, implicit
// residual norm in the stopping test, right preconditioning: w = A M^-1 v_k and
// x += M^-1 (V y) at restart / termination.
int gmres_solve(const Csr &A, const double *b, double *x, int max_iters,
                double factor, int m, const Precond *M = nullptr)
{
    const bool pc = M && M->kind;
    const idx n = A.nrows;
    if (m < 1) m = 1;
    std::vector<std::vector<double>> V(m + 1, std::vector<double>(n, 0.0));
    std::vector<double> H((size_t)(m + 1) * m, 0.0);  // column-major (m+1) x m
    std::vector<double> cs(m, 0.0), sn(m, 0.0), g(m + 1, 0.0), y(m, 0.0);
    std::vector<double> r(n), w(n);
    auto Hc = [&](int i, int j) -> double & { return H[(size_t)j * (m + 1) + i]; };

    auto restart = [&]() -> double {
        std::copy(b, b + n, r.begin());
        spmv_adv(A, -1.0, x, 1.0, r.data());
        double rn = norm2(r.data(), n);
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = rn;
        for (idx i = 0; i < n; ++i) V[0][i] = rn != 0.0 ? r[i] / rn : 0.0;
        return rn;
    };
    auto update_x = [&](int k) {
        // back substitution H(0:k,0:k) y = g(0:k); x += V y
        for (int i = k - 1; i >= 0; --i) {
            double s = g[i];
            for (int j = i + 1; j < k; ++j) s -= Hc(i, j) * y[j];
            y[i] = s / Hc(i, i);
        }
        if (pc) {
            std::vector<double> upd(n, 0.0), pv(n, 0.0);
            for (int j = 0; j < k; ++j)
                for (idx i = 0; i < n; ++i) upd[i] += y[j] * V[j][i];
            precond_apply(*M, upd.data(), pv.data(), n);
            for (idx i = 0; i < n; ++i) x[i] += pv[i];
            return;
        }
        for (int j = 0; j < k; ++j)
            for (idx i = 0; i < n; ++i) x[i] += y[j] * V[j][i];
    };

    double resnorm = restart();
    const double r0 = resnorm;
    int total = -1, k = 0;
    while (true) {
        ++total;
        if (total >= max_iters || resnorm < factor * r0) break;
        if (k == m) {
            update_x(k);
            resnorm = restart();
            k = 0;
        }
        if (pc) {
            precond_apply(*M, V[k].data(), r.data(), n);  // r doubles as scratch
            spmv(A, r.data(), w.data());
        } else {
            spmv(A, V[k].data(), w.data());
        }
        for (int i = 0; i <= k; ++i) {
            double h = dot(w.data(), V[i].data(), n);
            Hc(i, k) = h;
            for (idx t = 0; t < n; ++t) w[t] -= h * V[i][t];
        }
        double hn = norm2(w.data(), n);
        Hc(k + 1, k) = hn;
        for (idx t = 0; t < n; ++t) V[k + 1][t] = hn != 0.0 ? w[t] / hn : 0.0;
        // apply previous rotations
        for (int i = 0; i < k; ++i) {
            double t = cs[i] * Hc(i, k) + sn[i] * Hc(i + 1, k);
            Hc(i + 1, k) = -sn[i] * Hc(i, k) + cs[i] * Hc(i + 1, k);
            Hc(i, k) = t;
        }
        // new rotation
        {
            double a = Hc(k, k), c = Hc(k + 1, k);
            if (a == 0.0) {
                cs[k] = 0.0;
                sn[k] = 1.0;
            } else {
                double sc = std::fabs(a) + std::fabs(c);
                double hyp = sc * std::sqrt((a / sc) * (a / sc) + (c / sc) * (c / sc));
                cs[k] = a / hyp;
                sn[k] = c / hyp;
            }
            Hc(k, k) = cs[k] * a + sn[k] * c;
            Hc(k + 1, k) = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            resnorm = std::fabs(g[k + 1]);
        }
        ++k;
    }
    update_x(k);
    return total;
}


// -----------------------------------------------------------------------------
// A13. Factorised direct variant (source/solve.cpp:75-174, 281-399, 709-720;
// include/solver_tools.hpp:69-87).  CHOLMOD is not available; the oracle does
// its own simplicial LL^T of P A P^T for a caller-supplied ordering `perm`
// (perm[k] = original index of the k-th row of the permuted matrix, as
// cholmod's L_factor->Perm; the ordering vector is shared between oracle and
// product, like the METIS partition vector).  U = L^T as CSR (solve.cpp:286-304).
// Row-by-row up-looking factorisation with column lists; small cases only.
// -----------------------------------------------------------------------------
bool cholesky(const Csr &A, const std::vector<idx> &perm, Csr &L)
{
    const idx n = A.nrows;
    std::vector<idx> inv(n);
    for (idx i = 0; i < n; ++i) inv[perm[i]] = i;
    std::vector<std::vector<std::pair<idx, double>>> cols(n);  // L(:,j), rows asc
    std::vector<double> diag(n, 0.0), w(n, 0.0);
    std::vector<char> inpat(n, 0);
    L = Csr();
    L.nrows = L.ncols = n;
    L.rp.assign(n + 1, 0);
    std::vector<idx> pat;
    for (idx i = 0; i < n; ++i) {
        pat.clear();
        double d = 0.0;
        const idx oi = perm[i];
        for (idx k = A.rp[oi]; k < A.rp[oi + 1]; ++k) {
            idx j = inv[A.ci[k]];
            if (j < i) {
                if (!inpat[j]) { inpat[j] = 1; pat.push_back(j); w[j] = 0.0; }
                w[j] += A.v[k];
            } else if (j == i) {
                d += A.v[k];
            }
        }
        // ascending sparse forward solve; fill discovered on the fly
        std::make_heap(pat.begin(), pat.end(), std::greater<idx>());
        std::vector<std::pair<idx, double>> row;
        while (!pat.empty()) {
            std::pop_heap(pat.begin(), pat.end(), std::greater<idx>());
            idx j = pat.back();
            pat.pop_back();
            double yj = w[j] / diag[j];
            inpat[j] = 0;
            w[j] = 0.0;
            row.emplace_back(j, yj);
            d -= yj * yj;
            for (auto &e : cols[j]) {  // rows k in (j, i)
                idx k = e.first;
                if (!inpat[k]) {
                    inpat[k] = 1;
                    w[k] = 0.0;
                    pat.push_back(k);
                    std::push_heap(pat.begin(), pat.end(), std::greater<idx>());
                }
                w[k] -= e.second * yj;
            }
        }
        if (!(d > 0.0)) return false;
        diag[i] = std::sqrt(d);
        for (auto &e : row) {
            cols[e.first].emplace_back(i, e.second);
            L.ci.push_back(e.first);
            L.v.push_back(e.second);
        }
        L.ci.push_back(i);
        L.v.push_back(diag[i]);
        L.rp[i + 1] = (idx)L.ci.size();
    }
    return true;
}

// UMFPACK branch of the factorised variant (source/solve.cpp:145-171, 322-385): the reference
// takes P, Q, L, U with P A Q = L U from UMFPACK (its row scaling is fetched but never applied,
// SURVEY Appendix D) and solves x = Q U^-1 L^-1 P b with Ginkgo's triangular solvers.  UMFPACK
// is not available; the oracle does its own dense LU with partial pivoting (Q = identity, P
// from the pivot search, first row of largest magnitude) and keeps the factors as CSR.  Small
// cases only (O(n^3)).
bool dense_lu(const Csr &A, Csr &L, Csr &U, std::vector<idx> &p)
{
    const idx n = A.nrows;
    std::vector<double> M((size_t)n * n, 0.0);
    for (idx i = 0; i < n; ++i)
        for (idx k = A.rp[i]; k < A.rp[i + 1]; ++k) M[(size_t)i * n + A.ci[k]] += A.v[k];
    p.resize(n);
    std::iota(p.begin(), p.end(), 0);
    for (idx k = 0; k < n; ++k) {
        idx piv = k;
        for (idx i = k + 1; i < n; ++i)
            if (std::fabs(M[(size_t)i * n + k]) > std::fabs(M[(size_t)piv * n + k])) piv = i;
        if (M[(size_t)piv * n + k] == 0.0) return false;
        if (piv != k) {
            for (idx j = 0; j < n; ++j) std::swap(M[(size_t)k * n + j], M[(size_t)piv * n + j]);
            std::swap(p[k], p[piv]);
        }
        const double d = M[(size_t)k * n + k];
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (idx i = k + 1; i < n; ++i) {
            double &lik = M[(size_t)i * n + k];
            if (lik == 0.0) continue;
            lik /= d;
            for (idx j = k + 1; j < n; ++j) M[(size_t)i * n + j] -= lik * M[(size_t)k * n + j];
        }
    }
    L = Csr();
    U = Csr();
    L.nrows = L.ncols = U.nrows = U.ncols = n;
    L.rp.assign(n + 1, 0);
    U.rp.assign(n + 1, 0);
    for (idx i = 0; i < n; ++i) {
        for (idx j = 0; j < i; ++j)
            if (M[(size_t)i * n + j] != 0.0) {
                L.ci.push_back(j);
                L.v.push_back(M[(size_t)i * n + j]);
            }
        L.ci.push_back(i);
        L.v.push_back(1.0);
        for (idx j = i; j < n; ++j)
            if (j == i || M[(size_t)i * n + j] != 0.0) {
                U.ci.push_back(j);
                U.v.push_back(M[(size_t)i * n + j]);
            }
        L.rp[i + 1] = (idx)L.ci.size();
        U.rp[i + 1] = (idx)U.ci.size();
    }
    return true;
}

Csr transpose(const Csr &A)
{
    Csr T;
    T.nrows = A.ncols;
    T.ncols = A.nrows;
    T.rp.assign(T.nrows + 1, 0);
    for (idx k = 0; k < A.nnz(); ++k) T.rp[A.ci[k] + 1]++;
    for (idx i = 0; i < T.nrows; ++i) T.rp[i + 1] += T.rp[i];
    T.ci.resize(A.nnz());
    T.v.resize(A.nnz());
    std::vector<idx> cur(T.rp.begin(), T.rp.end() - 1);
    for (idx r = 0; r < A.nrows; ++r)
        for (idx k = A.rp[r]; k < A.rp[r + 1]; ++k) {
            idx p = cur[A.ci[k]]++;
            T.ci[p] = r;
            T.v[p] = A.v[k];
        }
    return T;
}

// gko::solver::LowerTrs / UpperTrs apply (reference executor: serial
// substitution, diagonal taken from the matrix); SURVEY Appendix F.
void lower_trs(const Csr &L, const double *b, double *x)
{
    for (idx i = 0; i < L.nrows; ++i) {
        double s = b[i], d = 1.0;
        for (idx k = L.rp[i]; k < L.rp[i + 1]; ++k) {
            idx c = L.ci[k];
            if (c < i) s -= L.v[k] * x[c];
            else if (c == i) d = L.v[k];
        }
        x[i] = s / d;
    }
}
void upper_trs(const Csr &U, const double *b, double *x)
{
    for (idx i = U.nrows - 1; i >= 0; --i) {
        double s = b[i], d = 1.0;
        for (idx k = U.rp[i]; k < U.rp[i + 1]; ++k) {
            idx c = U.ci[k];
            if (c > i) s -= U.v[k] * x[c];
            else if (c == i) d = U.v[k];
        }
        x[i] = s / d;
    }
}

// -----------------------------------------------------------------------------
// A8. Boundary exchange.
// -----------------------------------------------------------------------------
// Two-sided, source/restricted_schwarz.cpp:856-973.  The oracle defines
// "synchronous" as receive-completes-before-unpack: every rank
// first gathers into its send buffer (Gather copy, :893-897), then every rank
// scatters what its neighbours sent (Scatter copy, :955-959).
void exchange_twosided(Problem &pb)
{
    const int P = pb.P;
    for (int p = 0; p < P; ++p) {
        Rank &R = pb.ranks[p];
        if (R.finished) continue;
        size_t num_put = 0;
        for (size_t j = 0; j < R.nbr_out.size(); ++j) {
            for (size_t k = 0; k < R.put[j].size(); ++k)
                R.send_buf[num_put + k] = R.x[R.put[j][k]];
            num_put += R.put[j].size();
        }
    }
    for (int q = 0; q < P; ++q) {
        Rank &R = pb.ranks[q];
        if (R.finished) continue;
        size_t num_get = 0;
        for (size_t j = 0; j < R.nbr_in.size(); ++j) {
            const Rank &S = pb.ranks[R.nbr_in[j]];
            // locate my block inside the sender's send buffer (MPI matches the
            // message by source rank; the block order is neighbors_out order)
            size_t off = 0;
            for (size_t jj = 0; jj < S.nbr_out.size(); ++jj) {
                if (S.nbr_out[jj] == q) break;
                off += S.put[jj].size();
            }
            // use_mixed_precision (restricted_schwarz.cpp:898-903, 952-954): the values travel
            // as MixedValueType (float mirrors of the send / receive buffers)
            for (size_t k = 0; k < R.get[j].size(); ++k)
                R.recv_buf[num_get + k] = pb.opt.use_mixed_precision
                                              ? (double)(float)S.send_buf[off + k]
                                              : S.send_buf[off + k];
            for (size_t k = 0; k < R.get[j].size(); ++k)
                R.x[R.get[j][k]] = R.recv_buf[num_get + k];
            num_get += R.get[j].size();
        }
    }
}

// One-sided, source/restricted_schwarz.cpp:715-852 + include/comm_helpers.hpp.
// Deterministic emulation of one admissible asynchronous schedule: ranks act in
// ascending order within an outer iteration, each one Put-ting into (or
// Get-ting from) its neighbours' buffers as they are at that moment and then
// unpacking whatever its own receive buffer holds.  iter 0 returns (:725).
void exchange_onesided_rank(Problem &pb, int p)
{
    Rank &R = pb.ranks[p];
    if (pb.iter_count == 0) return;
    const Options &o = pb.opt;
    if (o.enable_put) {
        if (o.enable_one_by_one) {
            // comm_helpers.hpp:58-89: element idx of my x -> element idx of q's x
            for (size_t j = 0; j < R.nbr_out.size(); ++j) {
                Rank &Q = pb.ranks[R.nbr_out[j]];
                for (idx g : R.put[j]) Q.x[g] = R.x[g];
            }
        } else {
            size_t num_put = 0;
            for (size_t j = 0; j < R.nbr_out.size(); ++j) {
                int q = R.nbr_out[j];
                Rank &Q = pb.ranks[q];
                for (size_t k = 0; k < R.put[j].size(); ++k)
                    R.send_buf[num_put + k] = R.x[R.put[j][k]];  // pack_buffer
                // transfer_buffer: MPI_Put at put_displacements[q]
                for (size_t k = 0; k < R.put[j].size(); ++k)
                    Q.recv_buf[R.put_disp[q] + k] = o.use_mixed_precision
                                                        ? (double)(float)R.send_buf[num_put + k]
                                                        : R.send_buf[num_put + k];   // :769-787
                num_put += R.put[j].size();
            }
            size_t num_get = 0;
            for (size_t j = 0; j < R.nbr_in.size(); ++j) {  // unpack_buffer
                for (size_t k = 0; k < R.get[j].size(); ++k)
                    R.x[R.get[j][k]] = R.recv_buf[num_get + k];
                num_get += R.get[j].size();
            }
        }
    } else {  // enable_get
        if (o.enable_one_by_one) {
            for (size_t j = 0; j < R.nbr_in.size(); ++j) {
                const Rank &Q = pb.ranks[R.nbr_in[j]];
                for (idx g : R.get[j]) R.x[g] = Q.x[g];
            }
        } else {
            size_t num_put = 0;
            for (size_t j = 0; j < R.nbr_out.size(); ++j) {
                for (size_t k = 0; k < R.put[j].size(); ++k)
                    R.send_buf[num_put + k] = R.x[R.put[j][k]];
                num_put += R.put[j].size();
            }
            size_t num_get = 0;
            for (size_t j = 0; j < R.nbr_in.size(); ++j) {
                int q = R.nbr_in[j];
                const Rank &Q = pb.ranks[q];
                // MPI_Get from q's send buffer at get_displacements[q]; the
                // reference tests global_put[p][0] here (:823-824, SURVEY
                // Appendix D) — the oracle uses the in-list count.
                for (size_t k = 0; k < R.get[j].size(); ++k)
                    R.recv_buf[num_get + k] = o.use_mixed_precision
                                                  ? (double)(float)Q.send_buf[R.get_disp[q] + k]
                                                  : Q.send_buf[R.get_disp[q] + k];   // :819-835
                for (size_t k = 0; k < R.get[j].size(); ++k)
                    R.x[R.get[j][k]] = R.recv_buf[num_get + k];
                num_get += R.get[j].size();
            }
        }
    }
}

// -----------------------------------------------------------------------------
// A9. SolverRAS::update_boundary, source/restricted_schwarz.cpp:992-1017:
// local_solution = local_rhs - I * x   (I has GLOBAL column indices).
// -----------------------------------------------------------------------------
void update_boundary(Problem &pb, int me)
{
    Rank &R = pb.ranks[me];
    R.local_sol = R.local_rhs;
    if (pb.P > 1 && pb.opt.overlap > 0 && R.iface.nrows > 0)
        spmv_adv(R.iface, -1.0, R.x.data(), 1.0, R.local_sol.data());
}

// -----------------------------------------------------------------------------
// A10. Solve::check_local_convergence, source/solve.cpp:796-856.
// -----------------------------------------------------------------------------
bool check_local_convergence(Problem &pb, int me)
{
    Rank &R = pb.ranks[me];
    bool locally_converged = false;
    R.resnorm = -1.0;
    const double tol = pb.opt.tolerance;
    if (tol >= 0.0) {
        double *local_b = R.work.data();
        double *local_x = R.work.data() + R.local_size_x;
        std::copy(R.local_sol.begin(), R.local_sol.end(), local_b);  // :828
        extract_local_vector(pb, me, local_x, R.x.data());           // :829-831
        spmv_adv(R.local, -1.0, local_x, 1.0, local_b);              // :834
        R.resnorm = norm2(local_b, R.local_size_x);                  // :841-843
        if (R.resnorm0 < 0.0) R.resnorm0 = R.resnorm;                // :845
        locally_converged = (R.resnorm * R.resnorm) /
                                (R.resnorm0 * R.resnorm0) <
                            (tol * tol);                             // :847-849
    }
    return locally_converged;
}

// -----------------------------------------------------------------------------
// A11. conv_tools (include/conv_tools.hpp:147-275) on plain shared arrays:
// a remote MPI_Put of `1` is a store into the target rank's conv array.
// -----------------------------------------------------------------------------
void conv_tree(Problem &pb, int me, int converged_all_local, int &num_conv)
{
    Rank &R = pb.ranks[me];
    const int P = pb.P;
    auto &c = R.conv;
    if (((c[0] == 1 && c[1] == 1) || (c[0] == 1 && me == P / 2 - 1) ||
         (me >= P / 2 && c[0] != 2)) &&
        converged_all_local > 0) {
        if (me == 0) {
            c[2] = 1;
        } else {
            int p = (me - 1) / 2;
            int id = (me % 2 == 0 ? 1 : 0);
            pb.ranks[p].conv[id] = 1;
        }
        c[0] = 2;
    }
    if (c[2] == 1) {
        int p = 2 * me + 1;
        if (p < P) pb.ranks[p].conv[2] = 1;
        ++p;
        if (p < P) pb.ranks[p].conv[2] = 1;
        c[1]++;
        num_conv = P;
    } else {
        num_conv = 0;
    }
}

void conv_decentralized(Problem &pb, int me, int converged_all_local,
                        int &num_conv)
{
    Rank &R = pb.ranks[me];
    const int P = pb.P;
    if (pb.opt.enable_accumulate) {  // :230-247
        if (converged_all_local == 1) {
            for (int j = 0; j < P; ++j) {
                if (j != me) pb.ranks[j].conv[0] += 1;
                else R.conv[0]++;
            }
        }
        num_conv = R.conv[0];
    } else {  // :248-274
        if (converged_all_local == 1) R.conv[me] = 1;
        R.conv_local = R.conv;
        num_conv = std::accumulate(R.conv.begin(), R.conv.end(), 0);
        for (size_t i = 0; i < R.nbr_out.size(); ++i) {
            Rank &Q = pb.ranks[R.nbr_out[i]];
            for (int j = 0; j < P; ++j)
                if (R.conv_sent[j] == 0 && R.conv_local[j] == 1) Q.conv[j] = 1;
        }
        R.conv_sent = R.conv_local;
    }
}

// -----------------------------------------------------------------------------
// A11. Solve::check_global_convergence / check_convergence,
// source/solve.cpp:860-955, 959-1005.  `allgather` (two-sided) is resolved by
// the caller having computed every rank's local norm first.
// -----------------------------------------------------------------------------
void check_global_convergence(Problem &pb, int me, int &converged_all_local,
                              int &num_conv)
{
    Rank &R = pb.ranks[me];
    const Options &o = pb.opt;
    const int P = pb.P;
    if (o.enable_global_check && !o.enable_onesided) {
        // :890-905 MPI_Allgather + ordered sum
        R.gres = 0.0;
        for (int j = 0; j < P; ++j) {
            double lj = pb.ranks[j].resnorm;
            if (lj != DBL_MAX) {
                R.gres += lj;
            } else {
                R.gres = -1.0;
                break;
            }
        }
        if (R.gres >= 0.0) {  // :908-912
            if (R.gres0 < 0.0) R.gres0 = R.gres;
            if (R.gres / R.gres0 <= o.tolerance) converged_all_local++;
        }
    } else if (o.enable_onesided) {
        if (R.resnorm / R.resnorm0 <= o.tolerance) converged_all_local++;  // :914
        R.l_res[me] = std::min(R.l_res[me], R.resnorm);
    }
    if (o.enable_onesided) {  // :927-943
        if (o.conv_tree) conv_tree(pb, me, converged_all_local, num_conv);
        else if (o.conv_decentralized)
            conv_decentralized(pb, me, converged_all_local, num_conv);
    } else {  // :944-954
        if (o.enable_global_check) {
            if (converged_all_local == 1) num_conv = P;
        } else {
            // MPI_Allreduce of (converged_all_local != 0): never incremented on
            // this branch -> 0.
            num_conv = 0;
        }
    }
}

void check_convergence(Problem &pb, int me)
{
    Rank &R = pb.ranks[me];
    const Options &o = pb.opt;
    int num_converged_p = check_local_convergence(pb, me) ? 1 : 0;  // :975-981
    R.res_hist.push_back(R.resnorm);                                // :985
    const int iter = pb.iter_count;
    bool iter_cond = o.iter_offset
                         ? ((iter > (o.max_iters * 0.05)) || o.max_iters < 1000)
                         : true;  // :992-996
    if (o.tolerance > 0.0 && iter_cond) {
        int converged_all_local = 0;
        check_global_convergence(pb, me, converged_all_local, num_converged_p);
        R.num_converged = num_converged_p;  // :1003
    }
    R.gres_hist.push_back(R.gres);
}

// -----------------------------------------------------------------------------
// A12/A13. Solve::local_solve, source/solve.cpp:667-792.
// -----------------------------------------------------------------------------
void local_solve(Problem &pb, int me)
{
    Rank &R = pb.ranks[me];
    const Options &o = pb.opt;
    if (o.local_solver == 2) {
        // solve.cpp:458-463 cap; :753-754 apply(rhs = local_solution,
        // x = init_guess) warm start (F10); :781 local_solution <- init_guess
        int cap = o.local_max_iters == -1 ? R.local.nrows : o.local_max_iters;
        int it;
        if (o.non_symmetric)
            it = gmres_solve(R.local, R.local_sol.data(), R.init_guess.data(),
                             cap, o.local_tol, o.restart_iter, R.precond.get());
        else
            it = cg_solve(R.local, R.local_sol.data(), R.init_guess.data(), cap,
                          o.local_tol, R.precond.get());
        R.last_local_iters = it;
        R.local_iter_hist.push_back(it);
        R.local_sol = R.init_guess;
    } else {
        // direct_solver_ginkgo, solve.cpp:709-720 + solver_tools.hpp:69-87:
        // perm_sol = P b (out[i] = in[perm[i]]); L y = perm_sol; U z = y;
        // local_solution = P^-1 z (out[perm[i]] = in[i]).
        const idx n = R.local_size_x;
        double *perm_sol = R.work.data();
        double *tmp = R.work.data() + n;
        for (idx i = 0; i < n; ++i) perm_sol[i] = R.local_sol[R.fperm[i]];
        lower_trs(R.L, perm_sol, tmp);
        upper_trs(R.U, tmp, perm_sol);
        const std::vector<idx> &q = R.fperm_col.empty() ? R.fperm : R.fperm_col;
        for (idx i = 0; i < n; ++i) R.local_sol[q[i]] = perm_sol[i];
        R.local_iter_hist.push_back(0);
    }
}

// A14. Communicate::local_to_global_vector, source/communicate.cpp:65-94
// (solution_based branch; residual_based is unreachable from bench_ras).
void local_to_global_vector(Problem &pb, int me)
{
    Rank &R = pb.ranks[me];
    const idx first = pb.first_row[me];
    for (idx i = 0; i < R.local_size; ++i) R.x[first + i] = R.local_sol[i];
}

// -----------------------------------------------------------------------------
// A15. One pass of the loop body of SchwarzBase::run,
// source/schwarz_base.cpp:387-452, for all subdomains.  Returns the number of
// ranks that left the loop in this pass (break at :432-433).
// -----------------------------------------------------------------------------
// Rank-parallel stepping (bench.py's CPU arm): the reference runs its P MPI ranks side by side,
// each with its OpenMP team.  With g_rank_threads > 1 the per-rank stages of a synchronous
// step run on that many host threads at once (nested teams of g_threads inside); every rank's
// arithmetic is the same sequence as in the serial sweep, so the iterates do not change.
int g_rank_threads = 1;
template <class F>
void for_ranks(int P, F f)
{
    if (g_rank_threads > 1) {
#pragma omp parallel for num_threads(g_rank_threads) schedule(static, 1)
        for (int p = 0; p < P; ++p) f(p);
    } else {
        for (int p = 0; p < P; ++p) f(p);
    }
}

int ras_step(Problem &pb)
{
    const int P = pb.P;
    const Options &o = pb.opt;
    int newly_finished = 0;
    if (!o.enable_onesided) {
        // synchronous: every stage completes on all ranks before the next
        exchange_twosided(pb);
        // local norms first (the allgather needs all of them)
        std::vector<int> loc(P, 0);
        for_ranks(P, [&](int p) {
            if (pb.ranks[p].finished) return;
            update_boundary(pb, p);
            loc[p] = check_local_convergence(pb, p) ? 1 : 0;
        });
        for (int p = 0; p < P; ++p)
            if (!pb.ranks[p].finished) pb.ranks[p].res_hist.push_back(pb.ranks[p].resnorm);
        for (int p = 0; p < P; ++p) {
            if (pb.ranks[p].finished) continue;
            Rank &R = pb.ranks[p];
            int num_converged_p = loc[p];
            const int iter = pb.iter_count;
            bool iter_cond =
                o.iter_offset
                    ? ((iter > (o.max_iters * 0.05)) || o.max_iters < 1000)
                    : true;
            if (o.tolerance > 0.0 && iter_cond) {
                int cal = 0;
                check_global_convergence(pb, p, cal, num_converged_p);
                R.num_converged = num_converged_p;
            }
            R.gres_hist.push_back(R.gres);
        }
        for (int p = 0; p < P; ++p) {
            Rank &R = pb.ranks[p];
            if (R.finished) continue;
            if (R.num_converged == P) {
                R.finished = true;
                R.finished_iter = pb.iter_count;
                ++newly_finished;
            }
        }
        for_ranks(P, [&](int p) {
            if (pb.ranks[p].finished) return;
            local_solve(pb, p);
            local_to_global_vector(pb, p);
        });
    } else {
        // asynchronous emulation: ranks run their whole loop body in turn
        for (int p = 0; p < P; ++p) {
            Rank &R = pb.ranks[p];
            if (R.finished) continue;
            exchange_onesided_rank(pb, p);
            update_boundary(pb, p);
            check_convergence(pb, p);
            if (R.num_converged == P) {
                R.finished = true;
                R.finished_iter = pb.iter_count;
                ++newly_finished;
                continue;
            }
            local_solve(pb, p);
            local_to_global_vector(pb, p);
        }
    }
    pb.iter_count++;
    return newly_finished;
}

// Solve::compute_residual_norm, source/solve.cpp:1025-1085: the global
// solution is the sum over ranks of the own parts (MPI_Allreduce SUM of
// zero-padded vectors), then r = b - A x on the (permuted) global matrix.
void final_residual(Problem &pb, double *x_out, double out[4])
{
    std::vector<double> xs(pb.N, 0.0);
    for (int p = 0; p < pb.P; ++p) {
        const Rank &R = pb.ranks[p];
        for (idx i = pb.first_row[p]; i < pb.first_row[p + 1]; ++i)
            xs[i] += R.x[i];
    }
    std::vector<double> r(pb.rhs);
    out[1] = norm2(pb.rhs.data(), pb.N);   // rhs_norm
    out[2] = norm2(xs.data(), pb.N);       // sol_norm
    spmv_adv(pb.g, -1.0, xs.data(), 1.0, r.data());
    out[0] = norm2(r.data(), pb.N);        // residual_norm
    out[3] = out[0] / out[1];
    if (x_out) std::copy(xs.begin(), xs.end(), x_out);
}

}  // namespace

// =============================================================================
// C API (ctypes-friendly)
// =============================================================================
extern "C" {

void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
// subdomains stepped side by side (1 = one after the other, each with the whole team)
void orc_set_rank_threads(int n)
{
    g_rank_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
    omp_set_dynamic(0);
    omp_set_max_active_levels(g_rank_threads > 1 ? 2 : 1);
#endif
}
int orc_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- generators / partitioners ---------------------------------------------
int64_t orc_laplacian2d(int n, idx *rp, idx *ci, double *v)
{
    Csr A = laplacian2d(n);
    std::copy(A.rp.begin(), A.rp.end(), rp);
    std::copy(A.ci.begin(), A.ci.end(), ci);
    std::copy(A.v.begin(), A.v.end(), v);
    return A.nnz();
}
int64_t orc_laplacian3d(int n, idx *rp, idx *ci, double *v)
{
    Csr A = laplacian3d(n);
    std::copy(A.rp.begin(), A.rp.end(), rp);
    std::copy(A.ci.begin(), A.ci.end(), ci);
    std::copy(A.v.begin(), A.v.end(), v);
    return A.nnz();
}
int orc_partition_regular2d_rect(int64_t N, int P, int px, int py, uint32_t *pi)
{
    return partition_regular2d_rect(N, P, px, py, pi);
}

void orc_partition_regular2d(int64_t N, int P, uint32_t *pi)
{
    std::fill(pi, pi + N, 0u);  // resize()-zeroed, initialization.cpp:285
    partition_regular2d(N, P, pi);
}

// ---- kernels exposed for unit parity ---------------------------------------
void orc_spmv(idx nrows, const idx *rp, const idx *ci, const double *v,
              double alpha, const double *x, double beta, double *y)
{
    Csr A;
    A.nrows = nrows;
    A.rp.assign(rp, rp + nrows + 1);
    A.ci.assign(ci, ci + rp[nrows]);
    A.v.assign(v, v + rp[nrows]);
    spmv_adv(A, alpha, x, beta, y);
}
int orc_cg(idx nrows, const idx *rp, const idx *ci, const double *v,
           const double *b, double *x, int max_iters, double factor)
{
    Csr A;
    A.nrows = A.ncols = nrows;
    A.rp.assign(rp, rp + nrows + 1);
    A.ci.assign(ci, ci + rp[nrows]);
    A.v.assign(v, v + rp[nrows]);
    return cg_solve(A, b, x, max_iters, factor);
}
int orc_gmres(idx nrows, const idx *rp, const idx *ci, const double *v,
              const double *b, double *x, int max_iters, double factor, int m)
{
    Csr A;
    A.nrows = A.ncols = nrows;
    A.rp.assign(rp, rp + nrows + 1);
    A.ci.assign(ci, ci + rp[nrows]);
    A.v.assign(v, v + rp[nrows]);
    return gmres_solve(A, b, x, max_iters, factor, m);
}
// ---- preconditioners (kernel-level parity) -----------------------------------
static Csr csr_from(idx nrows, const idx *rp, const idx *ci, const double *v)
{
    Csr A;
    A.nrows = A.ncols = nrows;
    A.rp.assign(rp, rp + nrows + 1);
    A.ci.assign(ci, ci + rp[nrows]);
    A.v.assign(v, v + rp[nrows]);
    return A;
}
void *orc_precond_create(idx nrows, const idx *rp, const idx *ci, const double *v, int kind,
                         int max_block_size)
{
    Precond *M = new Precond();
    precond_generate(csr_from(nrows, rp, ci, v), kind, max_block_size, *M);
    return M;
}
void orc_precond_free(void *h) { delete (Precond *)h; }
void orc_precond_apply(void *h, idx n, const double *r, double *z)
{
    precond_apply(*(Precond *)h, r, z, n);
}
int64_t orc_precond_block_ptrs(void *h, idx *out)
{
    Precond *M = (Precond *)h;
    if (out) std::copy(M->block_ptrs.begin(), M->block_ptrs.end(), out);
    return (int64_t)M->block_ptrs.size();
}
int64_t orc_precond_blocks(void *h, double *out)
{
    Precond *M = (Precond *)h;
    if (out) std::copy(M->blocks.begin(), M->blocks.end(), out);
    return (int64_t)M->blocks.size();
}
// which: 0 L, 1 U, 2 approximate inverse of L, 3 of U
int64_t orc_precond_csr(void *h, int which, idx *rp, idx *ci, double *v)
{
    Precond *M = (Precond *)h;
    const Csr &T = which == 0 ? M->L : which == 1 ? M->U : which == 2 ? M->Li : M->Ui;
    if (rp) {
        std::copy(T.rp.begin(), T.rp.end(), rp);
        std::copy(T.ci.begin(), T.ci.end(), ci);
        std::copy(T.v.begin(), T.v.end(), v);
    }
    return (int64_t)T.ci.size();
}
int orc_cg_pc(idx nrows, const idx *rp, const idx *ci, const double *v, const double *b,
              double *x, int max_iters, double factor, void *precond)
{
    return cg_solve(csr_from(nrows, rp, ci, v), b, x, max_iters, factor, (Precond *)precond);
}
int orc_gmres_pc(idx nrows, const idx *rp, const idx *ci, const double *v, const double *b,
                 double *x, int max_iters, double factor, int m, void *precond)
{
    return gmres_solve(csr_from(nrows, rp, ci, v), b, x, max_iters, factor, m,
                       (Precond *)precond);
}

// L (CSR, lower incl. diagonal) of P A P^T; returns nnz(L) or -1.  Call with
// L arrays == NULL to query the size.
int64_t orc_cholesky(idx nrows, const idx *rp, const idx *ci, const double *v,
                     const idx *perm, idx *Lrp, idx *Lci, double *Lv)
{
    Csr A, L;
    A.nrows = A.ncols = nrows;
    A.rp.assign(rp, rp + nrows + 1);
    A.ci.assign(ci, ci + rp[nrows]);
    A.v.assign(v, v + rp[nrows]);
    std::vector<idx> pv(perm, perm + nrows);
    if (!cholesky(A, pv, L)) return -1;
    if (Lrp) {
        std::copy(L.rp.begin(), L.rp.end(), Lrp);
        std::copy(L.ci.begin(), L.ci.end(), Lci);
        std::copy(L.v.begin(), L.v.end(), Lv);
    }
    return L.nnz();
}
void orc_trs(idx nrows, const idx *rp, const idx *ci, const double *v,
             int upper, const double *b, double *x)
{
    Csr A;
    A.nrows = A.ncols = nrows;
    A.rp.assign(rp, rp + nrows + 1);
    A.ci.assign(ci, ci + rp[nrows]);
    A.v.assign(v, v + rp[nrows]);
    if (upper) upper_trs(A, b, x);
    else lower_trs(A, b, x);
}

// ---- problem handle ---------------------------------------------------------
// partition_kind: 0 = regular (1-D split, no permutation), 1 = metis/regular2d
// (permutation by `part`).
void *orc_create(idx N, const idx *rp, const idx *ci, const double *v, int P,
                 int partition_kind, const uint32_t *part, int overlap)
{
    Problem *pb = new Problem();
    pb->N = N;
    pb->P = P;
    pb->overlap = overlap;
    pb->opt.overlap = overlap;
    Csr A;
    A.nrows = A.ncols = N;
    A.rp.assign(rp, rp + N + 1);
    A.ci.assign(ci, ci + rp[N]);
    A.v.assign(v, v + rp[N]);
    pb->ranks.resize(P);
    setup_partition(*pb, A, partition_kind, part);
    for (int p = 0; p < P; ++p) setup_local_matrices(*pb, p);
    setup_comm_buffers(*pb);
    setup_windows(*pb);
    pb->rhs.assign(N, 1.0);  // schwarz_base.cpp:169
    setup_vectors(*pb);
    return pb;
}
void orc_destroy(void *h) { delete (Problem *)h; }

void orc_first_row(void *h, idx *out)
{
    Problem *pb = (Problem *)h;
    std::copy(pb->first_row.begin(), pb->first_row.end(), out);
}
int orc_permutation(void *h, idx *perm, idx *iperm)
{
    Problem *pb = (Problem *)h;
    if (!pb->permuted) return 0;
    std::copy(pb->perm.begin(), pb->perm.end(), perm);
    std::copy(pb->iperm.begin(), pb->iperm.end(), iperm);
    return 1;
}
int64_t orc_global_matrix(void *h, idx *rp, idx *ci, double *v)
{
    Problem *pb = (Problem *)h;
    if (rp) {
        std::copy(pb->g.rp.begin(), pb->g.rp.end(), rp);
        std::copy(pb->g.ci.begin(), pb->g.ci.end(), ci);
        std::copy(pb->g.v.begin(), pb->g.v.end(), v);
    }
    return pb->g.nnz();
}
// out: local_size, local_size_x, overlap_size, nnz_local, nnz_interface,
//      n_halo, num_neighbors_in, num_neighbors_out
void orc_sizes(void *h, int rank, int64_t *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    out[0] = R.local_size;
    out[1] = R.local_size_x;
    out[2] = R.overlap_size;
    out[3] = R.local.nnz();
    out[4] = R.iface.nnz();
    out[5] = R.n_halo;
    out[6] = (int64_t)R.nbr_in.size();
    out[7] = (int64_t)R.nbr_out.size();
}
void orc_l2g(void *h, int rank, idx *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.l2g.begin(), R.l2g.begin() + R.local_size_x + R.n_halo, out);
}
void orc_g2l(void *h, int rank, idx *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.g2l.begin(), R.g2l.end(), out);
}
void orc_local_matrix(void *h, int rank, idx *rp, idx *ci, double *v)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.local.rp.begin(), R.local.rp.end(), rp);
    std::copy(R.local.ci.begin(), R.local.ci.end(), ci);
    std::copy(R.local.v.begin(), R.local.v.end(), v);
}
int orc_interface_matrix(void *h, int rank, idx *rp, idx *ci, double *v)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    if (R.iface.nrows == 0) return 0;
    std::copy(R.iface.rp.begin(), R.iface.rp.end(), rp);
    std::copy(R.iface.ci.begin(), R.iface.ci.end(), ci);
    std::copy(R.iface.v.begin(), R.iface.v.end(), v);
    return R.iface.nrows;
}
void orc_neighbors(void *h, int rank, idx *nin, idx *nout)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.nbr_in.begin(), R.nbr_in.end(), nin);
    std::copy(R.nbr_out.begin(), R.nbr_out.end(), nout);
}
int orc_get_list(void *h, int rank, int j, idx *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    if (out) std::copy(R.get[j].begin(), R.get[j].end(), out);
    return (int)R.get[j].size();
}
int orc_put_list(void *h, int rank, int j, idx *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    if (out) std::copy(R.put[j].begin(), R.put[j].end(), out);
    return (int)R.put[j].size();
}
void orc_displacements(void *h, int rank, idx *put_disp, idx *get_disp)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.put_disp.begin(), R.put_disp.end(), put_disp);
    std::copy(R.get_disp.begin(), R.get_disp.end(), get_disp);
}

// ---- run ---------------------------------------------------------------------
struct orc_options {
    double tolerance, local_tol;
    int32_t max_iters, local_max_iters, non_symmetric, restart_iter,
        local_solver, enable_onesided, enable_put, enable_one_by_one,
        enable_global_check, conv_tree, conv_decentralized, enable_accumulate,
        iter_offset, use_mixed_precision, local_precond, precond_max_block_size,
        local_factorization;
};

void orc_set_rhs(void *h, const double *rhs)
{
    Problem *pb = (Problem *)h;
    pb->rhs.assign(rhs, rhs + pb->N);
    setup_vectors(*pb);
}

// Factor orderings for the direct variant: perm_all = concatenation over ranks
// of local_size_x entries each (may be NULL -> natural ordering).
int orc_configure(void *h, const orc_options *o, const idx *perm_all)
{
    Problem *pb = (Problem *)h;
    Options &d = pb->opt;
    d.tolerance = o->tolerance;
    d.local_tol = o->local_tol;
    d.max_iters = o->max_iters;
    d.local_max_iters = o->local_max_iters;
    d.non_symmetric = o->non_symmetric;
    d.restart_iter = o->restart_iter;
    d.local_solver = o->local_solver;
    d.enable_onesided = o->enable_onesided;
    d.enable_put = o->enable_put;
    d.enable_one_by_one = o->enable_one_by_one;
    d.enable_global_check = o->enable_global_check;
    d.conv_tree = o->conv_tree;
    d.conv_decentralized = o->conv_decentralized;
    d.enable_accumulate = o->enable_accumulate;
    d.iter_offset = o->iter_offset;
    d.use_mixed_precision = o->use_mixed_precision;
    d.local_precond = o->local_precond;
    d.precond_max_block_size = o->precond_max_block_size;
    d.local_factorization = o->local_factorization;
    d.overlap = pb->overlap;
    setup_windows(*pb);
    setup_vectors(*pb);
    if (d.local_solver == 1) {
        size_t off = 0;
        for (int p = 0; p < pb->P; ++p) {
            Rank &R = pb->ranks[p];
            R.fperm.resize(R.local_size_x);
            if (perm_all)
                std::copy(perm_all + off, perm_all + off + R.local_size_x,
                          R.fperm.begin());
            else
                std::iota(R.fperm.begin(), R.fperm.end(), 0);
            off += R.local_size_x;
            R.fperm_col.clear();
            if (d.local_factorization == 1) {
                if (!dense_lu(R.local, R.L, R.U, R.fperm)) return -1;
                R.fperm_col.resize(R.local_size_x);
                std::iota(R.fperm_col.begin(), R.fperm_col.end(), 0);
                continue;
            }
            if (!cholesky(R.local, R.fperm, R.L)) return -1;
            R.U = transpose(R.L);
        }
    }
    for (int p = 0; p < pb->P; ++p) {
        Rank &R = pb->ranks[p];
        R.precond.reset();
        if (d.local_solver == 2 && d.local_precond != 0) {
            R.precond = std::make_shared<Precond>();
            precond_generate(R.local, d.local_precond, d.precond_max_block_size, *R.precond);
        }
    }
    pb->solver_ready = true;
    return 0;
}

int orc_step(void *h) { return ras_step(*(Problem *)h); }

// Runs the outer loop to max_iters or until every rank has left it; returns
// the iteration count at which the last rank stopped (metadata.iter_count).
int orc_run(void *h)
{
    Problem *pb = (Problem *)h;
    int alive = pb->P;
    while (pb->iter_count < pb->opt.max_iters && alive > 0) {
        alive -= ras_step(*pb);
        if (alive <= 0) return pb->iter_count - 1;
    }
    return pb->iter_count;
}
int orc_iter_count(void *h) { return ((Problem *)h)->iter_count; }

// per-rank state getters
void orc_x(void *h, int rank, double *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.x.begin(), R.x.end(), out);
}
void orc_local_solution(void *h, int rank, double *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.local_sol.begin(), R.local_sol.end(), out);
}
void orc_local_rhs(void *h, int rank, double *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    std::copy(R.local_rhs.begin(), R.local_rhs.end(), out);
}
// out: resnorm, resnorm0, gres, gres0, num_converged, finished, finished_iter,
//      last_local_iters
void orc_rank_status(void *h, int rank, double *out)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    out[0] = R.resnorm;
    out[1] = R.resnorm0;
    out[2] = R.gres;
    out[3] = R.gres0;
    out[4] = R.num_converged;
    out[5] = R.finished ? 1 : 0;
    out[6] = R.finished_iter;
    out[7] = R.last_local_iters;
}
int orc_history(void *h, int rank, double *res, double *gres, int32_t *liters)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    if (res) std::copy(R.res_hist.begin(), R.res_hist.end(), res);
    if (gres) std::copy(R.gres_hist.begin(), R.gres_hist.end(), gres);
    if (liters)
        std::copy(R.local_iter_hist.begin(), R.local_iter_hist.end(), liters);
    return (int)R.res_hist.size();
}
// out[4] = residual_norm, rhs_norm, sol_norm, relative residual; x_out (N) may
// be NULL.
void orc_final_residual(void *h, double *x_out, double *out)
{
    final_residual(*(Problem *)h, x_out, out);
}
int64_t orc_factor(void *h, int rank, idx *Lrp, idx *Lci, double *Lv)
{
    Rank &R = ((Problem *)h)->ranks[rank];
    if (Lrp) {
        std::copy(R.L.rp.begin(), R.L.rp.end(), Lrp);
        std::copy(R.L.ci.begin(), R.L.ci.end(), Lci);
        std::copy(R.L.v.begin(), R.L.v.end(), Lv);
    }
    return R.L.nnz();
}

}  // extern "C"
