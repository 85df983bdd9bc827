// Host index sets of the RAS iteration (see setup.hpp).  Reference being
// replaced: source/initialization.cpp:197-329, include/partition_tools.hpp,
// source/restricted_schwarz.cpp:56-711.  The rules implemented are those of
// SURVEY.md Appendix A; parity with the CPU oracle is asserted bit-exactly in
// tests/test_setup.py.
#include "setup.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace schwz_b200 {

// -----------------------------------------------------------------------------
// Row sources
// -----------------------------------------------------------------------------
namespace {

// 2-D 5-pt Laplacian rows in closed form.  The reference walks a sorted
// exclusion list of the wrap-around couplings (k*n, k*n-1) / (k*n-1, k*n),
// k = 1..n-1 (source/initialization.cpp:225-262); those are exactly the
// "-1" neighbour of a row-start point and the "+1" neighbour of a row-end
// point.  Values {-1,-1,4,-1,-1}, offsets ascending.
struct Lap2D final : RowSource {
    int64_t n;
    explicit Lap2D(int32_t n_) : n(n_)
    {
        N = (int32_t)(n * n);
        max_row = 5;
    }
    int row(int32_t i, int32_t *c, double *v) const override
    {
        int k = 0;
        const int64_t x = i % n;
        if (i - n >= 0) { c[k] = (int32_t)(i - n); v[k++] = -1.0; }
        if (x != 0) { c[k] = i - 1; v[k++] = -1.0; }
        c[k] = i; v[k++] = 4.0;
        if (x != n - 1) { c[k] = i + 1; v[k++] = -1.0; }
        if ((int64_t)i + n < N) { c[k] = (int32_t)(i + n); v[k++] = -1.0; }
        return k;
    }
};

// Extension (no 3-D generator in the reference): 7-pt, natural order.
struct Lap3D final : RowSource {
    int64_t n;
    explicit Lap3D(int32_t n_) : n(n_)
    {
        N = (int32_t)(n * n * n);
        max_row = 7;
    }
    int row(int32_t i, int32_t *c, double *v) const override
    {
        int k = 0;
        const int64_t x = i % n, y = (i / n) % n, z = i / (n * n);
        if (z > 0) { c[k] = (int32_t)(i - n * n); v[k++] = -1.0; }
        if (y > 0) { c[k] = (int32_t)(i - n); v[k++] = -1.0; }
        if (x > 0) { c[k] = i - 1; v[k++] = -1.0; }
        c[k] = i; v[k++] = 6.0;
        if (x < n - 1) { c[k] = i + 1; v[k++] = -1.0; }
        if (y < n - 1) { c[k] = (int32_t)(i + n); v[k++] = -1.0; }
        if (z < n - 1) { c[k] = (int32_t)(i + n * n); v[k++] = -1.0; }
        return k;
    }
};

struct Stored final : RowSource {
    std::vector<int32_t> rp_own, ci_own;
    std::vector<double> v_own;
    const int32_t *rp, *ci;
    const double *v;
    Stored(int32_t N_, const int32_t *rp_, const int32_t *ci_, const double *v_, bool copy)
    {
        N = N_;
        if (copy) {
            rp_own.assign(rp_, rp_ + N + 1);
            ci_own.assign(ci_, ci_ + rp_[N]);
            v_own.assign(v_, v_ + rp_[N]);
            rp = rp_own.data();
            ci = ci_own.data();
            v = v_own.data();
        } else {
            rp = rp_;
            ci = ci_;
            v = v_;
        }
        max_row = 0;
        for (int32_t i = 0; i < N; ++i) max_row = std::max(max_row, rp[i + 1] - rp[i]);
    }
    int row(int32_t i, int32_t *c, double *vals) const override
    {
        const int32_t a = rp[i], b = rp[i + 1];
        for (int32_t k = a; k < b; ++k) {
            c[k - a] = ci[k];
            vals[k - a] = v[k];
        }
        return b - a;
    }
};

// Symmetric permutation view: new row r = old row perm[r], columns renamed by
// iperm, entry order of the old row kept — NOT re-sorted
// (source/restricted_schwarz.cpp:135-151).
struct Permuted final : RowSource {
    const RowSource &base;
    const std::vector<int32_t> &perm, &iperm;
    Permuted(const RowSource &b, const std::vector<int32_t> &p, const std::vector<int32_t> &ip)
        : base(b), perm(p), iperm(ip)
    {
        N = b.N;
        max_row = b.max_row;
    }
    int row(int32_t i, int32_t *c, double *v) const override
    {
        int k = base.row(perm[i], c, v);
        for (int j = 0; j < k; ++j) c[j] = iperm[c[j]];
        return k;
    }
};

struct Identity final : RowSource {
    const RowSource &base;
    explicit Identity(const RowSource &b) : base(b)
    {
        N = b.N;
        max_row = b.max_row;
    }
    int row(int32_t i, int32_t *c, double *v) const override { return base.row(i, c, v); }
};

inline void sort_row(int32_t *c, double *v, int k)
{
    for (int a = 1; a < k; ++a) {   // rows are short: insertion sort, stable
        int32_t cc = c[a];
        double vv = v[a];
        int b = a - 1;
        while (b >= 0 && c[b] > cc) {
            c[b + 1] = c[b];
            v[b + 1] = v[b];
            --b;
        }
        c[b + 1] = cc;
        v[b + 1] = vv;
    }
}

}  // namespace

std::unique_ptr<RowSource> make_laplacian2d(int32_t n) { return std::make_unique<Lap2D>(n); }
std::unique_ptr<RowSource> make_laplacian3d(int32_t n) { return std::make_unique<Lap3D>(n); }
std::unique_ptr<RowSource> make_stored(int32_t N, const int32_t *rp, const int32_t *ci,
                                       const double *v, bool copy)
{
    return std::make_unique<Stored>(N, rp, ci, v, copy);
}

HostCsr materialize(const RowSource &src)
{
    HostCsr A;
    A.nrows = A.ncols = src.N;
    A.rp.assign((size_t)src.N + 1, 0);
    std::vector<int32_t> c(src.max_row);
    std::vector<double> v(src.max_row);
    for (int32_t i = 0; i < src.N; ++i) {
        int k = src.row(i, c.data(), v.data());
        A.ci.insert(A.ci.end(), c.begin(), c.begin() + k);
        A.v.insert(A.v.end(), v.begin(), v.begin() + k);
        A.rp[i + 1] = (int32_t)A.ci.size();
    }
    return A;
}

// MatrixMarket coordinate real {general, symmetric}; result sorted by column
// within each row (gko::read + sort_by_column_index,
// source/initialization.cpp:210-212).
HostCsr read_mtx(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Could not find the file \"" + path + "\"");
    std::string line;
    std::getline(in, line);
    bool symmetric = line.find("symmetric") != std::string::npos;
    bool pattern = line.find("pattern") != std::string::npos;
    while (std::getline(in, line))
        if (!line.empty() && line[0] != '%') break;
    long long M = 0, Ncol = 0, nz = 0;
    {
        std::istringstream ss(line);
        ss >> M >> Ncol >> nz;
    }
    struct Ent { int32_t r, c; double v; };
    std::vector<Ent> ents;
    ents.reserve((size_t)nz * (symmetric ? 2 : 1));
    for (long long k = 0; k < nz; ++k) {
        long long r, c;
        double v = 1.0;
        in >> r >> c;
        if (!pattern) in >> v;
        ents.push_back({(int32_t)(r - 1), (int32_t)(c - 1), v});
        if (symmetric && r != c) ents.push_back({(int32_t)(c - 1), (int32_t)(r - 1), v});
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent &a, const Ent &b) {
        return a.r != b.r ? a.r < b.r : a.c < b.c;
    });
    HostCsr A;
    A.nrows = (int32_t)M;
    A.ncols = (int32_t)Ncol;
    A.rp.assign((size_t)M + 1, 0);
    A.ci.reserve(ents.size());
    A.v.reserve(ents.size());
    for (auto &e : ents) {
        A.rp[e.r + 1]++;
        A.ci.push_back(e.c);
        A.v.push_back(e.v);
    }
    for (long long i = 0; i < M; ++i) A.rp[i + 1] += A.rp[i];
    return A;
}

// include/partition_tools.hpp:70-94 (truncating square roots; subdomain ids
// beyond sq_p^2 receive no rows).
void partition_regular2d(int64_t N, int32_t P, uint32_t *part)
{
    std::fill(part, part + N, 0u);
    const int sq_n = (int)std::sqrt((double)N);
    const int sq_p = (int)std::sqrt((double)P);
    const int b = sq_n / sq_p;
    for (int j1 = 0; j1 < sq_p; ++j1)
        for (int j2 = 0; j2 < sq_p; ++j2) {
            const int64_t base = (int64_t)j1 * sq_p * b * b + (int64_t)j2 * sq_n / sq_p;
            for (int i1 = 0; i1 < b; ++i1)
                for (int i2 = 0; i2 < b; ++i2)
                    part[base + (int64_t)i1 * sq_n + i2] = (uint32_t)(sq_p * j1 + j2);
        }
}

// px x py blocks of an n x n grid: the rectangular extension of the rule above for subdomain
// counts that are not perfect squares (there the reference's truncating sqrt leaves subdomains
// without rows; BASELINE.json configs[1] asks for a 2-D partition of 8 subdomains).
// Same numbering as include/partition_tools.hpp:84 (id = py * block_row + block_column); cuts
// at k*n/px and k*n/py.  px = py = 0: most square factorisation with px <= py.
bool partition_regular2d_rect(int64_t N, int32_t P, int32_t px, int32_t py, uint32_t *part)
{
    int64_t n = (int64_t)std::sqrt((double)N);
    while (n * n < N) ++n;
    while (n * n > N) --n;
    if (n * n != N || P < 1) return false;
    if (px <= 0 || py <= 0) {
        px = 1;
        for (int32_t d = 2; (int64_t)d * d <= P; ++d)
            if (P % d == 0) px = d;
        py = P / px;
    }
    if ((int64_t)px * py != P || px > n || py > n) return false;
    std::vector<int32_t> col_owner((size_t)n);
    for (int32_t j2 = 0; j2 < py; ++j2)
        for (int64_t c = j2 * n / py; c < (j2 + 1) * n / py; ++c) col_owner[c] = j2;
    for (int32_t j1 = 0; j1 < px; ++j1)
        for (int64_t r = j1 * n / px; r < (j1 + 1) * n / px; ++r) {
            uint32_t *row = part + r * n;
            for (int64_t c = 0; c < n; ++c) row[c] = (uint32_t)(py * j1 + col_owner[c]);
        }
    return true;
}

// -----------------------------------------------------------------------------
// METIS (static library shipped inside the CUDA toolkit; idx_t = int64,
// real_t = float — SURVEY F4).  Same call sequence as
// include/partition_tools.hpp:124-196.
// -----------------------------------------------------------------------------
extern "C" {
int METIS_SetDefaultOptions(int64_t *options);
int METIS_PartGraphRecursive(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy,
                             int64_t *vwgt, int64_t *vsize, int64_t *adjwgt, int64_t *nparts,
                             float *tpwgts, float *ubvec, int64_t *options, int64_t *objval,
                             int64_t *part);
int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy,
                        int64_t *vwgt, int64_t *vsize, int64_t *adjwgt, int64_t *nparts,
                        float *tpwgts, float *ubvec, int64_t *options, int64_t *objval,
                        int64_t *part);
int METIS_NodeND(int64_t *nvtxs, int64_t *xadj, int64_t *adjncy, int64_t *vwgt,
                 int64_t *options, int64_t *perm, int64_t *iperm);
}

int partition_metis(int32_t N, const int32_t *rp, const int32_t *ci, int32_t P,
                    const char *objtype, uint32_t *part)
{
    int64_t n = N, ncon = 1, nparts = std::min<int64_t>(N, P), objval = 0;
    int64_t options[40];
    if (METIS_SetDefaultOptions(options) != 1) return -1;
    if (objtype && std::strcmp(objtype, "edgecut") == 0) options[1] = 0;        // METIS_OBJTYPE_CUT
    else if (objtype && std::strcmp(objtype, "totalvol") == 0) options[1] = 1;  // METIS_OBJTYPE_VOL
    // the reference hands METIS the CSR structure as is, diagonal included
    std::vector<int64_t> xadj(rp, rp + N + 1), adj(ci, ci + rp[N]), out(N);
    int rc;
    if (nparts <= 8)
        rc = METIS_PartGraphRecursive(&n, &ncon, xadj.data(), adj.data(), nullptr, nullptr,
                                      nullptr, &nparts, nullptr, nullptr, options, &objval,
                                      out.data());
    else
        rc = METIS_PartGraphKway(&n, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr,
                                 &nparts, nullptr, nullptr, options, &objval, out.data());
    if (rc != 1) return -2;
    for (int32_t i = 0; i < N; ++i) part[i] = (uint32_t)out[i];
    return 0;
}

int nd_ordering(int32_t n, const int32_t *rp, const int32_t *ci, int32_t *perm)
{
    // NodeND wants a loop-free symmetric structure
    std::vector<int64_t> xadj((size_t)n + 1, 0), adj;
    adj.reserve(rp[n]);
    for (int32_t i = 0; i < n; ++i) {
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
            if (ci[k] != i) adj.push_back(ci[k]);
        xadj[i + 1] = (int64_t)adj.size();
    }
    int64_t nv = n;
    int64_t options[40];
    if (METIS_SetDefaultOptions(options) != 1) return -1;
    std::vector<int64_t> p(n), ip(n);
    if (METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, options, p.data(), ip.data()) != 1)
        return -2;
    for (int32_t i = 0; i < n; ++i) perm[i] = (int32_t)p[i];
    return 0;
}

int nd_ordering_symmetrized(int32_t n, const int32_t *rp, const int32_t *ci, int32_t *perm)
{
    // pattern of A + A^T without the diagonal, rows sorted and duplicate-free
    std::vector<std::vector<int32_t>> adj((size_t)n);
    for (int32_t i = 0; i < n; ++i)
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
            if (ci[k] != i) {
                adj[i].push_back(ci[k]);
                adj[ci[k]].push_back(i);
            }
    std::vector<int32_t> srp((size_t)n + 1, 0), sci;
    for (int32_t i = 0; i < n; ++i) {
        std::sort(adj[i].begin(), adj[i].end());
        adj[i].erase(std::unique(adj[i].begin(), adj[i].end()), adj[i].end());
        sci.insert(sci.end(), adj[i].begin(), adj[i].end());
        srp[i + 1] = (int32_t)sci.size();
    }
    return nd_ordering(n, srp.data(), sci.data(), perm);
}

HostCsr transpose(const HostCsr &A)
{
    HostCsr T;
    T.nrows = A.ncols;
    T.ncols = A.nrows;
    T.rp.assign((size_t)T.nrows + 1, 0);
    const int64_t nnz = A.nnz();
    for (int64_t k = 0; k < nnz; ++k) T.rp[A.ci[k] + 1]++;
    for (int32_t i = 0; i < T.nrows; ++i) T.rp[i + 1] += T.rp[i];
    T.ci.resize(nnz);
    T.v.resize(nnz);
    std::vector<int32_t> cur(T.rp.begin(), T.rp.end() - 1);
    for (int32_t r = 0; r < A.nrows; ++r)
        for (int32_t k = A.rp[r]; k < A.rp[r + 1]; ++k) {
            const int32_t q = cur[A.ci[k]]++;
            T.ci[q] = r;
            T.v[q] = A.v[k];
        }
    return T;
}

// Left-looking sparse LU with threshold partial pivoting, P A Q = L U (replaces
// umfpack_di_symbolic / umfpack_di_numeric + umfpack_di_get_numeric, source/solve.cpp:145-171,
// 322-385; the reference never applies UMFPACK's row scaling, SURVEY Appendix D, so none is
// computed).  Q is given (q[j] = column of A eliminated j-th; a fill-reducing ordering of the
// pattern of A + A^T), P is found: column j is the sparse triangular solve x = L^-1 A(:, q[j])
// over the reach of its pattern in the graph of L; the pivot is the diagonal entry A-row q[j]
// when it is still free and |x| >= diag_tol * max (UMFPACK's "symmetric strategy" preference
// for the diagonal), else the entry of largest magnitude.  L has a unit diagonal.
// Returns false for a structurally or numerically singular matrix.
bool host_sparse_lu(const HostCsr &A, const int32_t *q, double diag_tol, HostCsr &L, HostCsr &U,
                    std::vector<int32_t> &p)
{
    const int32_t n = A.nrows;
    const HostCsr Ac = transpose(A);   // row c of Ac = column c of A
    std::vector<int32_t> pinv((size_t)n, -1), mark((size_t)n, -1), xi((size_t)n), rstack((size_t)n),
        pstack((size_t)n);
    std::vector<double> x((size_t)n, 0.0);
    // factors by columns; L with the original row ids until the end
    std::vector<int32_t> Lp(1, 0), Li, Up(1, 0), Ui;
    std::vector<double> Lx, Ux;
    Li.reserve((size_t)A.nnz() * 4);
    Lx.reserve((size_t)A.nnz() * 4);
    Ui.reserve((size_t)A.nnz() * 4);
    Ux.reserve((size_t)A.nnz() * 4);
    p.assign((size_t)n, -1);
    for (int32_t j = 0; j < n; ++j) {
        const int32_t col = q ? q[j] : j;
        int32_t top = n;
        for (int32_t e = Ac.rp[col]; e < Ac.rp[col + 1]; ++e) {
            const int32_t start = Ac.ci[e];
            if (mark[start] == j) continue;
            int32_t head = 0;
            rstack[0] = start;
            while (head >= 0) {   // depth-first search, post-order onto xi[top..n)
                const int32_t i = rstack[head];
                const int32_t k = pinv[i];
                if (mark[i] != j) {
                    mark[i] = j;
                    pstack[head] = k < 0 ? 0 : Lp[k];
                }
                const int32_t pend = k < 0 ? 0 : Lp[k + 1];
                bool done = true;
                for (int32_t pp = pstack[head]; pp < pend; ++pp) {
                    const int32_t r = Li[pp];
                    if (mark[r] == j) continue;
                    pstack[head] = pp + 1;
                    rstack[++head] = r;
                    done = false;
                    break;
                }
                if (done) {
                    --head;
                    xi[--top] = i;
                }
            }
        }
        for (int32_t e = Ac.rp[col]; e < Ac.rp[col + 1]; ++e) x[Ac.ci[e]] = Ac.v[e];
        // x = L^-1 A(:, col) in topological order; pivotal rows become U(:, j)
        double amax = -1.0;
        int32_t ipiv = -1;
        for (int32_t px = top; px < n; ++px) {
            const int32_t i = xi[px];
            const int32_t k = pinv[i];
            if (k < 0) {
                const double a = std::fabs(x[i]);
                if (a > amax) {
                    amax = a;
                    ipiv = i;
                }
                continue;
            }
            const double xk = x[i];
            Ui.push_back(k);
            Ux.push_back(xk);
            for (int32_t pp = Lp[k] + 1; pp < Lp[k + 1]; ++pp) x[Li[pp]] -= Lx[pp] * xk;
        }
        if (ipiv < 0 || !(amax > 0.0)) return false;
        if (pinv[col] < 0 && mark[col] == j && std::fabs(x[col]) >= diag_tol * amax) ipiv = col;
        const double pivot = x[ipiv];
        pinv[ipiv] = j;
        p[j] = ipiv;
        Ui.push_back(j);
        Ux.push_back(pivot);
        Up.push_back((int32_t)Ui.size());
        Li.push_back(ipiv);   // unit diagonal first
        Lx.push_back(1.0);
        for (int32_t px = top; px < n; ++px) {
            const int32_t i = xi[px];
            if (pinv[i] < 0) {
                Li.push_back(i);
                Lx.push_back(x[i] / pivot);
            }
            x[i] = 0.0;
        }
        if (Li.size() > 2000000000u || Ui.size() > 2000000000u)
            throw std::runtime_error("LU factor exceeds int32 indexing");
        Lp.push_back((int32_t)Li.size());
    }
    for (auto &r : Li) r = pinv[r];
    HostCsr Lc, Uc;   // the column-wise factors, as "CSR of the transpose"
    Lc.nrows = Lc.ncols = Uc.nrows = Uc.ncols = n;
    Lc.rp = std::move(Lp);
    Lc.ci = std::move(Li);
    Lc.v = std::move(Lx);
    Uc.rp = std::move(Up);
    Uc.ci = std::move(Ui);
    Uc.v = std::move(Ux);
    L = transpose(Lc);
    U = transpose(Uc);
    return true;
}

// Up-looking sparse Cholesky driven by the elimination tree (replaces
// cholmod_analyze + cholmod_factorize with supernodal = 0, final_ll = 1,
// source/solve.cpp:94-142).  Works column-wise on U = L^T (which is also the
// layout the reference extracts, solve.cpp:286-304) and returns L.
bool host_cholesky(const HostCsr &A, const int32_t *perm, HostCsr &L)
{
    const int32_t n = A.nrows;
    std::vector<int32_t> pv(n), inv(n);
    for (int32_t i = 0; i < n; ++i) pv[i] = perm ? perm[i] : i;
    for (int32_t i = 0; i < n; ++i) inv[pv[i]] = i;
    // C = upper triangle of P A P^T, by columns (== lower triangle by rows)
    std::vector<int32_t> cp((size_t)n + 1, 0), cidx;
    std::vector<double> cval;
    for (int32_t k = 0; k < n; ++k) {
        const int32_t o = pv[k];
        for (int32_t q = A.rp[o]; q < A.rp[o + 1]; ++q) {
            const int32_t i = inv[A.ci[q]];
            if (i <= k) {
                cidx.push_back(i);
                cval.push_back(A.v[q]);
            }
        }
        cp[k + 1] = (int32_t)cidx.size();
    }
    // elimination tree
    std::vector<int32_t> parent(n, -1), anc(n, -1);
    for (int32_t k = 0; k < n; ++k)
        for (int32_t q = cp[k]; q < cp[k + 1]; ++q) {
            int32_t i = cidx[q];
            while (i != -1 && i < k) {
                const int32_t nx = anc[i];
                anc[i] = k;
                if (nx == -1) parent[i] = k;
                i = nx;
            }
        }
    // symbolic: column counts of L through row-subtree walks
    std::vector<int32_t> cnt(n, 1), mark(n, -1);
    for (int32_t k = 0; k < n; ++k) {
        mark[k] = k;
        for (int32_t q = cp[k]; q < cp[k + 1]; ++q) {
            int32_t i = cidx[q];
            while (i < k && mark[i] != k) {
                mark[i] = k;
                cnt[i]++;
                i = parent[i];
            }
        }
    }
    std::vector<int64_t> lp((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; ++i) lp[i + 1] = lp[i] + cnt[i];
    if (lp[n] > 2147483647LL) throw std::runtime_error("factor exceeds int32 indexing");
    std::vector<int32_t> li((size_t)lp[n]);
    std::vector<double> lx((size_t)lp[n]);
    std::vector<int64_t> fill(lp.begin(), lp.end() - 1);
    std::vector<double> x(n, 0.0);
    std::vector<int32_t> stack(n), path(n);
    std::fill(mark.begin(), mark.end(), -1);
    for (int32_t k = 0; k < n; ++k) {
        // pattern of row k of L in topological order
        int32_t top = n;
        mark[k] = k;
        double d = 0.0;
        for (int32_t q = cp[k]; q < cp[k + 1]; ++q) {
            int32_t i = cidx[q];
            if (i == k) { d += cval[q]; continue; }
            x[i] += cval[q];
            int32_t len = 0;
            while (mark[i] != k) {
                path[len++] = i;
                mark[i] = k;
                i = parent[i];
            }
            while (len > 0) stack[--top] = path[--len];
        }
        for (; top < n; ++top) {
            const int32_t i = stack[top];
            const double lki = x[i] / lx[lp[i]];
            x[i] = 0.0;
            for (int64_t q = lp[i] + 1; q < fill[i]; ++q) x[li[q]] -= lx[q] * lki;
            d -= lki * lki;
            const int64_t q = fill[i]++;
            li[q] = k;
            lx[q] = lki;
        }
        if (!(d > 0.0)) return false;
        const int64_t q = fill[k]++;
        li[q] = k;
        lx[q] = std::sqrt(d);
    }
    // columns of L == rows of U; return L as CSR
    HostCsr U;
    U.nrows = U.ncols = n;
    U.rp.resize((size_t)n + 1);
    for (int32_t i = 0; i <= n; ++i) U.rp[i] = (int32_t)lp[i];
    U.ci = std::move(li);
    U.v = std::move(lx);
    L = transpose(U);
    return true;
}

// -----------------------------------------------------------------------------
// Setup
// -----------------------------------------------------------------------------
Setup::Setup(std::unique_ptr<RowSource> base, int32_t P, int32_t partition_kind,
             const uint32_t *part, int32_t overlap)
    : base_(std::move(base)), N_(base_->N), P_(P), overlap_(overlap)
{
    if (P < 1) throw std::runtime_error("need at least one subdomain");
    // default 1-D split (source/restricted_schwarz.cpp:84, 97-102)
    const int32_t nb = (int32_t)(((int64_t)N_ + P - 1) / P);
    first_row_.assign((size_t)P + 1, 0);
    local_p_size_.assign(P, 0);
    for (int32_t p = 0; p < P; ++p) {
        local_p_size_[p] = std::min<int32_t>(N_ - first_row_[p], nb);
        first_row_[p + 1] = first_row_[p] + local_p_size_[p];
    }
    if (partition_kind == 1) {
        // stable counting sort by part id (:108-133)
        perm_.resize(N_);
        iperm_.resize(N_);
        if (P > 1) {
            if (!part) throw std::runtime_error("partition vector required");
            std::fill(local_p_size_.begin(), local_p_size_.end(), 0);
            for (int32_t i = 0; i < N_; ++i) {
                if (part[i] >= (uint32_t)P) throw std::runtime_error("partition id out of range");
                local_p_size_[part[i]]++;
            }
            for (int32_t p = 0; p < P; ++p) first_row_[p + 1] = first_row_[p] + local_p_size_[p];
            std::vector<int32_t> cursor(first_row_.begin(), first_row_.end() - 1);
            for (int32_t i = 0; i < N_; ++i) perm_[cursor[part[i]]++] = i;
            for (int32_t i = 0; i < N_; ++i) iperm_[perm_[i]] = i;
        } else {
            std::iota(perm_.begin(), perm_.end(), 0);
            std::iota(iperm_.begin(), iperm_.end(), 0);
        }
        permuted_ = true;
        view_ = std::make_unique<Permuted>(*base_, perm_, iperm_);
    } else {
        view_ = std::make_unique<Identity>(*base_);
    }
    ranks_.resize(P);
    for (int32_t p = 0; p < P; ++p) ranks_[p].local_size = local_p_size_[p];
}

// global id -> 1 + local slot of one subdomain (0 = absent) without the shared dense scratch of
// the index-set pass: the own block is a contiguous id range, the few thousand overlap / halo ids
// sit in a small open-addressing table.  Makes index_set / build_matrices / compact_interface
// re-entrant, so the subdomains of one process are set up side by side.
namespace {
struct LocalIndex {
    int32_t first = 0, own = 0;
    uint32_t mask = 0;
    std::vector<int32_t> key, val;
    size_t used = 0;
    LocalIndex(int32_t first_, int32_t own_, size_t n_ext) : first(first_), own(own_)
    {
        size_t cap = 16;
        while (cap < 2 * n_ext + 1) cap <<= 1;
        mask = (uint32_t)cap - 1;
        key.assign(cap, -1);
        val.assign(cap, 0);
    }
    static uint32_t hash(int32_t g) { return (uint32_t)g * 2654435761u; }
    void grow()
    {
        std::vector<int32_t> k2, v2;
        k2.swap(key);
        v2.swap(val);
        const size_t cap = 2 * k2.size();
        mask = (uint32_t)cap - 1;
        key.assign(cap, -1);
        val.assign(cap, 0);
        used = 0;
        for (size_t i = 0; i < k2.size(); ++i)
            if (k2[i] != -1) put(k2[i], v2[i]);
    }
    void put(int32_t g, int32_t v)
    {
        if (2 * (used + 1) > key.size()) grow();
        uint32_t h = hash(g) & mask;
        while (key[h] != -1) h = (h + 1) & mask;
        key[h] = g;
        val[h] = v;
        ++used;
    }
    int32_t operator()(int32_t g) const
    {
        const uint32_t d = (uint32_t)(g - first);
        if (d < (uint32_t)own) return 1 + (int32_t)d;
        uint32_t h = hash(g) & mask;
        while (key[h] != -1) {
            if (key[h] == g) return val[h];
            h = (h + 1) & mask;
        }
        return 0;
    }
};
}  // namespace

// Index set of one subdomain.
void Setup::index_set(int32_t me)
{
    RankLayout &R = ranks_[me];
    const RowSource &g = *view_;
    std::vector<int32_t> c(g.max_row);
    std::vector<double> v(g.max_row);
    auto &l2g = R.l2g;
    l2g.clear();
    l2g.reserve((size_t)R.local_size + 1024);
    int32_t num = 0;
    for (int32_t i = first_row_[me]; i < first_row_[me + 1]; ++i) {
        l2g.push_back(i);
        ++num;
    }
    LocalIndex g2l(first_row_[me], R.local_size, 1024);
    int32_t old = 0;
    for (int k = 1; k < overlap_; ++k) {
        const int32_t now = num;
        for (int32_t i = old; i < now; ++i) {
            const int len = g.row(l2g[i], c.data(), v.data());
            for (int j = 0; j < len; ++j)
                if (g2l(c[j]) == 0) {
                    l2g.push_back(c[j]);
                    g2l.put(c[j], 1 + num);
                    ++num;
                }
        }
        old = now;
    }
    R.local_size_x = num;
    R.overlap_size = num - R.local_size;
    // does an interface exist?  (decides how far the halo sweep runs, see
    // the nnz_interface > 0 guard at source/restricted_schwarz.cpp:263-295)
    bool have_iface = false;
    for (int32_t k = R.local_size; k < R.local_size_x && !have_iface; ++k) {
        const int len = g.row(l2g[k], c.data(), v.data());
        for (int j = 0; j < len; ++j)
            if (g2l(c[j]) == 0) {
                have_iface = true;
                break;
            }
    }
    const int32_t now = have_iface ? R.local_size_x : R.local_size;
    if (now == num) {   // append position == sweep end (always, see above)
        for (int32_t i = old; i < now; ++i) {
            const int len = g.row(l2g[i], c.data(), v.data());
            for (int j = 0; j < len; ++j)
                if (g2l(c[j]) == 0) {
                    l2g.push_back(c[j]);
                    g2l.put(c[j], 1 + num);
                    ++num;
                }
        }
    }
    R.n_halo = num - R.local_size_x;
    // get-lists: for p ascending, every id of p's range present in g2l,
    // ascending (:336-371) == the non-own ids sorted and split by owner
    std::vector<int32_t> ext(l2g.begin() + R.local_size, l2g.end());
    std::sort(ext.begin(), ext.end());
    R.nbr_in.clear();
    R.get.clear();
    size_t pos = 0;
    for (int32_t p = 0; p < P_ && pos < ext.size(); ++p) {
        if (p == me) continue;
        size_t end = pos;
        while (end < ext.size() && ext[end] < first_row_[p + 1]) ++end;
        if (end > pos && ext[pos] >= first_row_[p]) {
            R.nbr_in.push_back(p);
            R.get.emplace_back(ext.begin() + pos, ext.begin() + end);
        }
        pos = end;
    }
    R.have_index = true;
}

void Setup::build_index_sets()
{
    if (have_index_) return;
    // every subdomain on its own (no shared scratch): side by side on the host cores
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t p = 0; p < P_; ++p) index_set(p);
    // owner side of the handshake (:400-472): put-list p -> q is q's get-list
    // from p; neighbours_out ascending
    for (int32_t me = 0; me < P_; ++me) {
        RankLayout &R = ranks_[me];
        R.nbr_out.clear();
        R.put.clear();
        for (int32_t q = 0; q < P_; ++q) {
            if (q == me) continue;
            const RankLayout &Q = ranks_[q];
            for (size_t j = 0; j < Q.nbr_in.size(); ++j)
                if (Q.nbr_in[j] == me) {
                    R.nbr_out.push_back(q);
                    R.put.push_back(Q.get[j]);
                }
        }
    }
    // displacement tables (:624-658): prefix sums in ascending-rank order,
    // exchanged with an all-to-all
    std::vector<std::vector<int32_t>> in_pref(P_), out_pref(P_);
    for (int32_t me = 0; me < P_; ++me) {
        const RankLayout &R = ranks_[me];
        std::vector<int32_t> a((size_t)P_ + 1, 0), b((size_t)P_ + 1, 0);
        for (size_t j = 0; j < R.nbr_in.size(); ++j) a[R.nbr_in[j] + 1] = (int32_t)R.get[j].size();
        for (size_t j = 0; j < R.nbr_out.size(); ++j) b[R.nbr_out[j] + 1] = (int32_t)R.put[j].size();
        for (int32_t j = 0; j < P_; ++j) {
            a[j + 1] += a[j];
            b[j + 1] += b[j];
        }
        in_pref[me] = std::move(a);
        out_pref[me] = std::move(b);
    }
    for (int32_t me = 0; me < P_; ++me) {
        RankLayout &R = ranks_[me];
        R.put_disp.assign((size_t)P_ + 1, 0);
        R.get_disp.assign((size_t)P_ + 1, 0);
        for (int32_t q = 0; q < P_; ++q) {
            R.put_disp[q] = in_pref[q][me];
            R.get_disp[q] = out_pref[q][me];
        }
    }
    have_index_ = true;
}

// Local and interface matrices of one subdomain.
void Setup::build_matrices(int32_t me)
{
    build_index_sets();
    RankLayout &R = ranks_[me];
    if (R.have_matrix) return;
    const RowSource &g = *view_;
    std::vector<int32_t> c(g.max_row), lc(g.max_row), ic(g.max_row);
    std::vector<double> v(g.max_row), lv(g.max_row), iv(g.max_row);
    // g2l as it is when the reference builds the matrices: own + overlap only
    LocalIndex g2l(first_row_[me], R.local_size, (size_t)R.overlap_size);
    for (int32_t k = R.local_size; k < R.local_size_x; ++k) g2l.put(R.l2g[k], 1 + k);
    HostCsr &Lm = R.local;
    HostCsr &Im = R.iface;
    Lm = HostCsr();
    Im = HostCsr();
    Lm.nrows = Lm.ncols = R.local_size_x;
    Lm.rp.assign((size_t)R.local_size_x + 1, 0);
    Lm.ci.reserve((size_t)R.local_size_x * (size_t)std::min(g.max_row, 8));
    Lm.v.reserve(Lm.ci.capacity());
    std::vector<int32_t> irp((size_t)R.local_size_x + 1, 0);
    for (int32_t k = 0; k < R.local_size_x; ++k) {
        const int len = g.row(R.l2g[k], c.data(), v.data());
        int nl = 0, ni = 0;
        for (int j = 0; j < len; ++j) {
            const int32_t loc = g2l(c[j]);
            if (loc != 0) {
                lc[nl] = loc - 1;
                lv[nl++] = v[j];
            } else if (k >= R.local_size) {   // own rows drop such entries (:199-203)
                ic[ni] = c[j];
                iv[ni++] = v[j];
            }
        }
        sort_row(lc.data(), lv.data(), nl);
        sort_row(ic.data(), iv.data(), ni);
        Lm.ci.insert(Lm.ci.end(), lc.begin(), lc.begin() + nl);
        Lm.v.insert(Lm.v.end(), lv.begin(), lv.begin() + nl);
        Lm.rp[k + 1] = (int32_t)Lm.ci.size();
        Im.ci.insert(Im.ci.end(), ic.begin(), ic.begin() + ni);
        Im.v.insert(Im.v.end(), iv.begin(), iv.begin() + ni);
        irp[k + 1] = (int32_t)Im.ci.size();
    }
    if (!Im.ci.empty()) {
        Im.nrows = Im.ncols = R.local_size_x;   // declared square (:227-229)
        Im.rp = std::move(irp);
    } else {
        Im.nrows = Im.ncols = 0;                // empty 0x0 matrix (:231)
        Im.rp.assign(1, 0);
    }
    R.have_matrix = true;
}

void Setup::release(int32_t r)
{
    RankLayout &R = ranks_[r];
    R.local = HostCsr();
    R.iface = HostCsr();
    R.have_matrix = false;
}

void Setup::compact_interface(int32_t me, HostCsr &out) const
{
    const RankLayout &R = ranks_[me];
    out = HostCsr();
    out.nrows = R.overlap_size;
    out.ncols = R.local_size_x + R.n_halo;
    out.rp.assign((size_t)R.overlap_size + 1, 0);
    if (R.iface.nrows == 0) return;
    LocalIndex g2l(0, 0, (size_t)R.n_halo);   // interface columns are halo ids only
    for (int32_t k = R.local_size_x; k < R.local_size_x + R.n_halo; ++k) g2l.put(R.l2g[k], 1 + k);
    out.ci.reserve(R.iface.ci.size());
    out.v = R.iface.v;
    for (int32_t k = 0; k < R.overlap_size; ++k) {
        const int32_t row = R.local_size + k;
        for (int32_t q = R.iface.rp[row]; q < R.iface.rp[row + 1]; ++q) {
            const int32_t loc = g2l(R.iface.ci[q]);
            if (loc == 0) throw std::runtime_error("interface column outside the halo layer");
            out.ci.push_back(loc - 1);
        }
        out.rp[k + 1] = (int32_t)out.ci.size();
    }
}

}  // namespace schwz_b200
