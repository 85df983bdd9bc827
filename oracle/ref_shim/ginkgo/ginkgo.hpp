// Stand-in for <ginkgo/ginkgo.hpp>, written from scratch for ONE purpose: to let the
// reference's own sources under /root/reference/source compile and run here, unmodified,
// as the second oracle (oracle/_ref). TEST INFRASTRUCTURE; never part of the product.
//
// It provides only the API surface schwarz-lib touches (SURVEY.md section 8c lists the call
// sites) with host-only, sequential, deterministic semantics:
//   * containers: Array, dim, matrix::Dense / Csr / Permutation, Executor + MemorySpace;
//   * Csr::apply row-wise in stored order, Dense::compute_norm2 = sqrt(sum v^2);
//   * solver::Cg / Gmres with Combined(Iteration, ResidualNormReduction) criteria,
//     following the Ginkgo recurrences restated in SURVEY.md Appendix F;
//   * solver::LowerTrs / UpperTrs (serial substitution);
//   * preconditioner::Jacobi (block detection by supervariable agglomeration, Gauss-Jordan
//     block inverses), factorization::ParIlu (one sequential sweep = ILU(0)),
//     preconditioner::Ilu over LowerTrs/UpperTrs or LowerIsai/UpperIsai, restated from the
//     upstream reference-executor kernels as remembered [upstream-memory].
// What it proves: every integer/index set, buffer layout, exchange and convergence
// protocol coming out of oracle/_ref is produced by the reference's own code. What it does
// not prove: bit-level agreement with upstream Ginkgo's kernels (not in the tree).
#ifndef GINKGO_SHIM_HPP_
#define GINKGO_SHIM_HPP_

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <fstream>
#include <initializer_list>
#include <iostream>
#include <istream>
#include <memory>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace gko {

using size_type = std::size_t;
using int32 = std::int32_t;
using int64 = std::int64_t;
using default_precision = double;

struct NotSupported : std::runtime_error {
    explicit NotSupported(const std::string &what)
        : std::runtime_error("ginkgo shim: not supported: " + what)
    {}
};

template <typename T>
inline T zero()
{
    return T{};
}
template <typename T>
inline T one()
{
    return T(1);
}
template <typename T>
inline T squared_norm(const T &x)
{
    return x * x;
}

// ----------------------------------------------------------------------------- dim
template <size_type N, typename D = size_type>
struct dim {
    D v[N];
    dim()
    {
        for (auto &e : v) e = 0;
    }
    template <typename A>
    explicit dim(A a)
    {
        for (auto &e : v) e = static_cast<D>(a);
    }
    template <typename A, typename B>
    dim(A a, B b)
    {
        static_assert(N == 2, "two-argument dim is 2-D");
        v[0] = static_cast<D>(a);
        v[1] = static_cast<D>(b);
    }
    const D &operator[](size_type i) const { return v[i]; }
    D &operator[](size_type i) { return v[i]; }
    explicit operator bool() const
    {
        for (auto &e : v)
            if (e == 0) return false;
        return true;
    }
    friend bool operator==(const dim &a, const dim &b)
    {
        for (size_type i = 0; i < N; ++i)
            if (a.v[i] != b.v[i]) return false;
        return true;
    }
    friend bool operator!=(const dim &a, const dim &b) { return !(a == b); }
};

inline dim<2> transpose(const dim<2> &d) { return dim<2>(d[1], d[0]); }

// ------------------------------------------------------------------ pointer helpers
template <typename T>
inline T *lend(const std::unique_ptr<T> &p)
{
    return p.get();
}
template <typename T>
inline T *lend(const std::shared_ptr<T> &p)
{
    return p.get();
}
template <typename T>
inline T *lend(T *p)
{
    return p;
}

template <typename P>
inline std::shared_ptr<typename std::remove_reference<P>::type::element_type> share(P &&p)
{
    // like upstream: takes ownership even from an lvalue (solve.cpp:525 shares an lvalue
    // unique_ptr); a const shared_ptr is copied
    return std::shared_ptr<typename std::remove_reference<P>::type::element_type>(std::move(p));
}

template <typename T, typename U>
inline T *as(U *obj)
{
    if (auto p = dynamic_cast<T *>(obj)) return p;
    throw NotSupported(std::string("gko::as: object is not of the requested type"));
}
template <typename T, typename U>
inline const T *as(const U *obj)
{
    if (auto p = dynamic_cast<const T *>(obj)) return p;
    throw NotSupported(std::string("gko::as: object is not of the requested type"));
}
template <typename T, typename U>
inline std::shared_ptr<T> as(const std::shared_ptr<U> &obj)
{
    if (auto p = std::dynamic_pointer_cast<T>(obj)) return p;
    throw NotSupported(std::string("gko::as: object is not of the requested type"));
}

// ------------------------------------------------------------------------ executors
class OmpExecutor;
class ReferenceExecutor;
class CudaExecutor;

class Operation {
public:
    virtual ~Operation() = default;
    virtual void run(std::shared_ptr<const OmpExecutor>) const
    {
        throw NotSupported("Operation without an OpenMP implementation");
    }
    virtual void run(std::shared_ptr<const CudaExecutor>) const
    {
        throw NotSupported("Operation without a CUDA implementation");
    }
};

class MemorySpace {
public:
    template <typename T>
    void copy_from(const MemorySpace *, size_type n, const T *src, T *dst) const
    {
        if (n > 0) std::memmove(dst, src, n * sizeof(T));
    }
};

struct ExecInfo {
    void bind_to_core(int) {}
    void bind_to_cores(const std::vector<int> &) {}
};

class Executor : public std::enable_shared_from_this<Executor> {
public:
    virtual ~Executor() = default;
    virtual void run(const Operation &op) const = 0;
    virtual std::shared_ptr<Executor> get_master() = 0;
    virtual std::shared_ptr<const Executor> get_master() const = 0;
    std::shared_ptr<MemorySpace> get_mem_space() const { return mem_space_; }
    ExecInfo *get_exec_info() const { return &exec_info_; }
    template <typename T>
    void copy(size_type n, const T *src, T *dst) const
    {
        if (n > 0) std::memmove(dst, src, n * sizeof(T));
    }
    template <typename T>
    void copy_from(const Executor *, size_type n, const T *src, T *dst) const
    {
        if (n > 0) std::memmove(dst, src, n * sizeof(T));
    }
    void synchronize() const {}

protected:
    std::shared_ptr<MemorySpace> mem_space_ = std::make_shared<MemorySpace>();
    mutable ExecInfo exec_info_;
};

class OmpExecutor : public Executor {
public:
    static std::shared_ptr<OmpExecutor> create()
    {
        return std::shared_ptr<OmpExecutor>(new OmpExecutor());
    }
    void run(const Operation &op) const override
    {
        op.run(std::static_pointer_cast<const OmpExecutor>(this->shared_from_this()));
    }
    std::shared_ptr<Executor> get_master() override { return this->shared_from_this(); }
    std::shared_ptr<const Executor> get_master() const override
    {
        return this->shared_from_this();
    }

protected:
    OmpExecutor() = default;
};

class ReferenceExecutor : public OmpExecutor {
public:
    static std::shared_ptr<ReferenceExecutor> create()
    {
        return std::shared_ptr<ReferenceExecutor>(new ReferenceExecutor());
    }

protected:
    ReferenceExecutor() = default;
};

// No device in the oracle: the type exists so that the reference's cuda branches compile.
class CudaExecutor : public Executor {
public:
    static std::shared_ptr<CudaExecutor> create(int, std::shared_ptr<Executor>, bool = false)
    {
        throw NotSupported("CudaExecutor (the oracle is host-only)");
    }
    int get_device_id() const { return 0; }
    void run(const Operation &op) const override
    {
        op.run(std::static_pointer_cast<const CudaExecutor>(this->shared_from_this()));
    }
    std::shared_ptr<Executor> get_master() override { return master_; }
    std::shared_ptr<const Executor> get_master() const override { return master_; }

private:
    std::shared_ptr<Executor> master_;
};

// ---------------------------------------------------------------------------- Array
template <typename T>
class Array {
public:
    using value_type = T;
    Array() = default;
    explicit Array(std::shared_ptr<const Executor> exec) : exec_(std::move(exec)) {}
    Array(std::shared_ptr<const Executor> exec, size_type n)
        : exec_(std::move(exec)), n_(n), own_(n ? new T[n]() : nullptr), data_(own_.get())
    {}
    template <typename It, typename = typename std::enable_if<
                               !std::is_integral<It>::value>::type>
    Array(std::shared_ptr<const Executor> exec, It b, It e)
        : Array(std::move(exec), static_cast<size_type>(std::distance(b, e)))
    {
        std::copy(b, e, data_);
    }
    Array(std::shared_ptr<const Executor> exec, const Array &o) : Array(std::move(exec))
    {
        *this = o;
    }
    Array(const Array &o) : Array(o.exec_) { *this = o; }
    Array(Array &&o) noexcept
        : exec_(std::move(o.exec_)), n_(o.n_), own_(std::move(o.own_)), data_(o.data_),
          view_(o.view_)
    {
        o.n_ = 0;
        o.data_ = nullptr;
        o.view_ = false;
    }
    static Array view(std::shared_ptr<const Executor> exec, size_type n, T *data)
    {
        Array a(std::move(exec));
        a.n_ = n;
        a.data_ = data;
        a.view_ = true;
        return a;
    }
    // Ginkgo semantics: keep the own executor if there is one; an owning array is resized
    // only when the size differs (so raw pointers stay valid across same-size copies, which
    // restricted_schwarz.cpp:93-151 relies on); a view must match in size.
    Array &operator=(const Array &o)
    {
        if (&o == this) return *this;
        if (!exec_) exec_ = o.exec_;
        if (view_) {
            if (o.n_ != n_) throw std::length_error("ginkgo shim: assignment into a view of another size");
        } else {
            resize_and_reset(o.n_);
        }
        if (n_) std::copy(o.data_, o.data_ + n_, data_);
        return *this;
    }
    Array &operator=(Array &&o) noexcept
    {
        if (&o == this) return *this;
        if (!exec_) exec_ = o.exec_;
        n_ = o.n_;
        own_ = std::move(o.own_);
        data_ = o.data_;
        view_ = o.view_;
        o.n_ = 0;
        o.data_ = nullptr;
        o.view_ = false;
        return *this;
    }
    void resize_and_reset(size_type n)
    {
        if (n == n_ && !view_) return;
        own_.reset(n ? new T[n]() : nullptr);
        data_ = own_.get();
        n_ = n;
        view_ = false;
    }
    T *get_data() { return data_; }
    const T *get_data() const { return data_; }
    const T *get_const_data() const { return data_; }
    size_type get_num_elems() const { return n_; }
    std::shared_ptr<const Executor> get_executor() const { return exec_; }
    void set_executor(std::shared_ptr<const Executor> e) { exec_ = std::move(e); }
    bool is_owning() const { return !view_; }

private:
    std::shared_ptr<const Executor> exec_;
    size_type n_ = 0;
    std::unique_ptr<T[]> own_;
    T *data_ = nullptr;
    bool view_ = false;
};

// ------------------------------------------------------------------- logging stubs
class LinOp;
namespace log {
class Logger {
public:
    using mask_type = std::uint64_t;
    static constexpr mask_type iteration_complete_mask = mask_type{1} << 0;
    static constexpr mask_type criterion_check_completed_mask = mask_type{1} << 1;
    virtual ~Logger() = default;
    virtual void on_criterion_check_completed(size_type num_iterations,
                                              const LinOp *residual) const = 0;
};

struct criterion_data {
    size_type num_iterations = 0;
    std::unique_ptr<const LinOp> residual;
    std::unique_ptr<const LinOp> residual_norm;
    std::unique_ptr<const LinOp> solution;
};

class Record : public Logger {
public:
    struct logged_data {
        std::deque<std::unique_ptr<criterion_data>> criterion_check_completed;
    };
    static std::shared_ptr<Record> create(std::shared_ptr<const Executor>,
                                          std::shared_ptr<MemorySpace>, mask_type = 0,
                                          size_type max_storage = 1)
    {
        auto r = std::shared_ptr<Record>(new Record());
        r->max_storage_ = max_storage;
        return r;
    }
    static std::shared_ptr<Record> create(std::shared_ptr<const Executor> e, mask_type m = 0,
                                          size_type max_storage = 1)
    {
        return create(std::move(e), nullptr, m, max_storage);
    }
    const logged_data &get() const { return data_; }
    logged_data &get() { return data_; }
    void on_criterion_check_completed(size_type num_iterations,
                                      const LinOp *residual) const override;

private:
    Record() = default;
    mutable logged_data data_;
    size_type max_storage_ = 1;
};
}  // namespace log

class Loggable {
public:
    void add_logger(std::shared_ptr<const log::Logger> l) { loggers_.push_back(std::move(l)); }
    const std::vector<std::shared_ptr<const log::Logger>> &get_loggers() const { return loggers_; }

protected:
    std::vector<std::shared_ptr<const log::Logger>> loggers_;
};

// ---------------------------------------------------------------------------- LinOp
class LinOp : public Loggable {
public:
    virtual ~LinOp() = default;
    const dim<2> &get_size() const { return size_; }
    std::shared_ptr<const Executor> get_executor() const { return exec_; }
    void apply(const LinOp *b, LinOp *x) const { this->apply_impl(b, x); }
    void apply(const LinOp *alpha, const LinOp *b, const LinOp *beta, LinOp *x) const
    {
        this->apply_impl(alpha, b, beta, x);
    }
    virtual std::unique_ptr<LinOp> clone() const
    {
        throw NotSupported("clone of this LinOp type");
    }

protected:
    LinOp(std::shared_ptr<const Executor> exec, dim<2> size = dim<2>())
        : size_(size), exec_(std::move(exec))
    {}
    virtual void apply_impl(const LinOp *b, LinOp *x) const = 0;
    virtual void apply_impl(const LinOp *alpha, const LinOp *b, const LinOp *beta,
                            LinOp *x) const = 0;
    dim<2> size_;
    std::shared_ptr<const Executor> exec_;
};

class LinOpFactory : public Loggable {
public:
    virtual ~LinOpFactory() = default;
    std::shared_ptr<const Executor> get_executor() const { return exec_; }
    std::unique_ptr<LinOp> generate(std::shared_ptr<const LinOp> op) const
    {
        return this->generate_impl(std::move(op));
    }

protected:
    explicit LinOpFactory(std::shared_ptr<const Executor> e) : exec_(std::move(e)) {}
    virtual std::unique_ptr<LinOp> generate_impl(std::shared_ptr<const LinOp>) const = 0;
    std::shared_ptr<const Executor> exec_;
};

namespace matrix {

// ---------------------------------------------------------------------------- Dense
template <typename V = default_precision>
class Dense : public LinOp {
public:
    using value_type = V;
    static std::unique_ptr<Dense> create(std::shared_ptr<const Executor> exec,
                                         const dim<2> &size = dim<2>())
    {
        return std::unique_ptr<Dense>(new Dense(std::move(exec), size));
    }
    static std::unique_ptr<Dense> create(std::shared_ptr<const Executor> exec,
                                         const dim<2> &size, size_type stride)
    {
        auto d = std::unique_ptr<Dense>(new Dense(exec, dim<2>()));
        d->size_ = size;
        d->stride_ = stride;
        d->values_ = Array<V>(exec, size[0] * stride);
        return d;
    }
    static std::unique_ptr<Dense> create(std::shared_ptr<const Executor> exec,
                                         const dim<2> &size, Array<V> values, size_type stride)
    {
        auto d = std::unique_ptr<Dense>(new Dense(exec, dim<2>()));
        d->size_ = size;
        d->stride_ = stride;
        d->values_ = std::move(values);
        d->values_.set_executor(exec);
        return d;
    }
    V *get_values() { return values_.get_data(); }
    const V *get_values() const { return values_.get_const_data(); }
    const V *get_const_values() const { return values_.get_const_data(); }
    size_type get_stride() const { return stride_; }
    size_type get_num_stored_elements() const { return values_.get_num_elems(); }
    V &at(size_type r, size_type c) { return values_.get_data()[r * stride_ + c]; }
    const V &at(size_type r, size_type c) const { return values_.get_const_data()[r * stride_ + c]; }
    V &at(size_type i) { return values_.get_data()[i / size_[1] * stride_ + i % size_[1]]; }
    const V &at(size_type i) const
    {
        return values_.get_const_data()[i / size_[1] * stride_ + i % size_[1]];
    }

    // Ginkgo's EnablePolymorphicAssignment: *this = *other, keeping the own executor; a
    // view of equal size is written through.
    Dense *copy_from(const LinOp *other)
    {
        if (auto o = dynamic_cast<const Dense<V> *>(other)) {
            values_ = o->values_;
            size_ = o->size_;
            stride_ = o->stride_;
            return this;
        }
        if (auto o = dynamic_cast<const Dense<typename std::conditional<
                         std::is_same<V, double>::value, float, double>::type> *>(other)) {
            o->convert_to(this);
            return this;
        }
        throw NotSupported("Dense::copy_from of a non-Dense operand");
    }
    Dense *copy_from(std::unique_ptr<LinOp> other) { return copy_from(other.get()); }
    template <typename O>
    Dense *copy_from(const std::unique_ptr<O> &other)
    {
        return copy_from(static_cast<const LinOp *>(other.get()));
    }
    template <typename W>
    void convert_to(Dense<W> *result) const
    {
        if (!(result->get_size() == size_) || result->get_stride() != stride_) {
            auto tmp = Dense<W>::create(result->get_executor() ? result->get_executor() : exec_,
                                        size_, stride_);
            result->move_from_same(tmp.get());
        }
        const size_type n = size_[0] * stride_;
        for (size_type i = 0; i < n; ++i) result->get_values()[i] = static_cast<W>(get_const_values()[i]);
    }
    void move_from_same(Dense *o)
    {
        values_ = std::move(o->values_);
        size_ = o->size_;
        stride_ = o->stride_;
    }
    std::unique_ptr<LinOp> clone() const override
    {
        auto d = Dense::create(exec_, size_, stride_);
        std::copy(get_const_values(), get_const_values() + size_[0] * stride_, d->get_values());
        return d;
    }
    void fill(V v)
    {
        for (size_type r = 0; r < size_[0]; ++r)
            for (size_type c = 0; c < size_[1]; ++c) at(r, c) = v;
    }
    // column-wise: result(0,c) = sqrt(sum_r at(r,c)^2), sequential in row order
    void compute_norm2(LinOp *result) const
    {
        auto res = as<Dense<V>>(result);
        for (size_type c = 0; c < size_[1]; ++c) {
            V s = zero<V>();
            for (size_type r = 0; r < size_[0]; ++r) s += at(r, c) * at(r, c);
            res->at(0, c) = std::sqrt(s);
        }
    }
    void compute_dot(const LinOp *b, LinOp *result) const
    {
        auto bb = as<Dense<V>>(b);
        auto res = as<Dense<V>>(result);
        for (size_type c = 0; c < size_[1]; ++c) {
            V s = zero<V>();
            for (size_type r = 0; r < size_[0]; ++r) s += at(r, c) * bb->at(r, c);
            res->at(0, c) = s;
        }
    }
    void scale(const LinOp *alpha)
    {
        auto a = as<Dense<V>>(alpha);
        for (size_type r = 0; r < size_[0]; ++r)
            for (size_type c = 0; c < size_[1]; ++c)
                at(r, c) *= a->at(0, a->get_size()[1] == 1 ? 0 : c);
    }
    void add_scaled(const LinOp *alpha, const LinOp *b)
    {
        auto a = as<Dense<V>>(alpha);
        auto bb = as<Dense<V>>(b);
        for (size_type r = 0; r < size_[0]; ++r)
            for (size_type c = 0; c < size_[1]; ++c)
                at(r, c) += a->at(0, a->get_size()[1] == 1 ? 0 : c) * bb->at(r, c);
    }

protected:
    Dense(std::shared_ptr<const Executor> exec, const dim<2> &size)
        : LinOp(exec, size), values_(exec, size[0] * size[1]), stride_(size[1])
    {}
    void apply_impl(const LinOp *, LinOp *) const override
    {
        throw NotSupported("Dense::apply");
    }
    void apply_impl(const LinOp *, const LinOp *, const LinOp *, LinOp *) const override
    {
        throw NotSupported("Dense::apply");
    }
    Array<V> values_;
    size_type stride_ = 1;
};

// ------------------------------------------------------------------------------ Csr
template <typename V = default_precision, typename I = int32>
class Csr : public LinOp {
public:
    using value_type = V;
    using index_type = I;
    static std::unique_ptr<Csr> create(std::shared_ptr<const Executor> exec,
                                       const dim<2> &size = dim<2>(), size_type nnz = 0)
    {
        return std::unique_ptr<Csr>(new Csr(std::move(exec), size, nnz));
    }
    static std::unique_ptr<Csr> create(std::shared_ptr<const Executor> exec, const dim<2> &size,
                                       Array<V> values, Array<I> col_idxs, Array<I> row_ptrs)
    {
        auto m = std::unique_ptr<Csr>(new Csr(exec, dim<2>(), 0));
        m->size_ = size;
        m->values_ = std::move(values);
        m->col_idxs_ = std::move(col_idxs);
        m->row_ptrs_ = std::move(row_ptrs);
        return m;
    }
    V *get_values() { return values_.get_data(); }
    const V *get_values() const { return values_.get_const_data(); }
    const V *get_const_values() const { return values_.get_const_data(); }
    I *get_col_idxs() { return col_idxs_.get_data(); }
    const I *get_col_idxs() const { return col_idxs_.get_const_data(); }
    const I *get_const_col_idxs() const { return col_idxs_.get_const_data(); }
    I *get_row_ptrs() { return row_ptrs_.get_data(); }
    const I *get_row_ptrs() const { return row_ptrs_.get_const_data(); }
    const I *get_const_row_ptrs() const { return row_ptrs_.get_const_data(); }
    size_type get_num_stored_elements() const { return values_.get_num_elems(); }

    Csr *copy_from(const LinOp *other)
    {
        auto o = dynamic_cast<const Csr<V, I> *>(other);
        if (!o) throw NotSupported("Csr::copy_from of a non-Csr operand");
        values_ = o->values_;
        col_idxs_ = o->col_idxs_;
        row_ptrs_ = o->row_ptrs_;
        size_ = o->size_;
        return this;
    }
    Csr *copy_from(std::unique_ptr<LinOp> other) { return copy_from(other.get()); }
    template <typename O>
    Csr *copy_from(const std::unique_ptr<O> &other)
    {
        return copy_from(static_cast<const LinOp *>(other.get()));
    }
    std::unique_ptr<LinOp> clone() const override
    {
        auto m = Csr::create(exec_);
        m->copy_from(this);
        return m;
    }

    // every row sorted by column index, values follow (reference kernel semantics)
    void sort_by_column_index()
    {
        const I *rp = get_const_row_ptrs();
        I *ci = get_col_idxs();
        V *va = get_values();
        std::vector<std::pair<I, V>> tmp;
        for (size_type r = 0; r < size_[0]; ++r) {
            tmp.clear();
            for (I k = rp[r]; k < rp[r + 1]; ++k) tmp.emplace_back(ci[k], va[k]);
            std::stable_sort(tmp.begin(), tmp.end(),
                             [](const std::pair<I, V> &a, const std::pair<I, V> &b) {
                                 return a.first < b.first;
                             });
            I k = rp[r];
            for (auto &e : tmp) {
                ci[k] = e.first;
                va[k] = e.second;
                ++k;
            }
        }
    }
    bool is_sorted_by_column_index() const
    {
        const I *rp = get_const_row_ptrs();
        const I *ci = get_const_col_idxs();
        for (size_type r = 0; r < size_[0]; ++r)
            for (I k = rp[r] + 1; k < rp[r + 1]; ++k)
                if (ci[k - 1] > ci[k]) return false;
        return true;
    }
    std::unique_ptr<LinOp> transpose() const
    {
        const size_type nr = size_[0], nc = size_[1];
        const I *rp = get_const_row_ptrs();
        const I *ci = get_const_col_idxs();
        const V *va = get_const_values();
        const size_type nnz = static_cast<size_type>(rp[nr]);
        auto t = Csr::create(exec_, dim<2>(nc, nr), nnz);
        I *trp = t->get_row_ptrs();
        for (size_type c = 0; c <= nc; ++c) trp[c] = 0;
        for (size_type k = 0; k < nnz; ++k) trp[ci[k] + 1]++;
        for (size_type c = 0; c < nc; ++c) trp[c + 1] += trp[c];
        std::vector<I> next(trp, trp + nc);
        for (size_type r = 0; r < nr; ++r)
            for (I k = rp[r]; k < rp[r + 1]; ++k) {
                const I dst = next[ci[k]]++;
                t->get_col_idxs()[dst] = static_cast<I>(r);
                t->get_values()[dst] = va[k];
            }
        return t;
    }

    // Matrix Market "coordinate" reader (what gko::read<Csr> needs here): entries sorted
    // row-major, symmetric storage mirrored, explicit zeros kept out as Csr::read does.
    static std::unique_ptr<Csr> read_mtx(std::istream &is, std::shared_ptr<const Executor> exec)
    {
        std::string line;
        if (!std::getline(is, line)) throw std::runtime_error("ginkgo shim: empty matrix file");
        std::string banner, object, format, field, symmetry;
        {
            std::istringstream hs(line);
            hs >> banner >> object >> format >> field >> symmetry;
        }
        auto lower = [](std::string s) {
            for (auto &c : s) c = static_cast<char>(std::tolower(c));
            return s;
        };
        format = lower(format);
        field = lower(field);
        symmetry = lower(symmetry);
        if (banner != "%%MatrixMarket" || format != "coordinate")
            throw NotSupported("matrix file that is not MatrixMarket coordinate");
        while (std::getline(is, line))
            if (!line.empty() && line[0] != '%') break;
        size_type nr = 0, nc = 0, ne = 0;
        {
            std::istringstream ss(line);
            ss >> nr >> nc >> ne;
        }
        std::vector<std::tuple<I, I, V>> ent;
        ent.reserve(symmetry == "general" ? ne : 2 * ne);
        for (size_type e = 0; e < ne; ++e) {
            long long r, c;
            double v = 1.0;
            is >> r >> c;
            if (field != "pattern") is >> v;
            ent.emplace_back(static_cast<I>(r - 1), static_cast<I>(c - 1), static_cast<V>(v));
            if (symmetry != "general" && r != c)
                ent.emplace_back(static_cast<I>(c - 1), static_cast<I>(r - 1),
                                 static_cast<V>(symmetry == "skew-symmetric" ? -v : v));
        }
        std::stable_sort(ent.begin(), ent.end(),
                         [](const std::tuple<I, I, V> &a, const std::tuple<I, I, V> &b) {
                             return std::make_pair(std::get<0>(a), std::get<1>(a)) <
                                    std::make_pair(std::get<0>(b), std::get<1>(b));
                         });
        size_type nnz = 0;
        for (auto &e : ent) nnz += (std::get<2>(e) != zero<V>());
        auto m = Csr::create(exec, dim<2>(nr, nc), nnz);
        I *rp = m->get_row_ptrs();
        for (size_type r = 0; r <= nr; ++r) rp[r] = 0;
        size_type k = 0;
        for (auto &e : ent) {
            if (std::get<2>(e) == zero<V>()) continue;
            rp[std::get<0>(e) + 1]++;
            m->get_col_idxs()[k] = std::get<1>(e);
            m->get_values()[k] = std::get<2>(e);
            ++k;
        }
        for (size_type r = 0; r < nr; ++r) rp[r + 1] += rp[r];
        return m;
    }

protected:
    Csr(std::shared_ptr<const Executor> exec, const dim<2> &size, size_type nnz)
        : LinOp(exec, size), values_(exec, nnz), col_idxs_(exec, nnz), row_ptrs_(exec, size[0] + 1)
    {}
    // x = A b : per row a sequential sum in stored order starting from zero
    void apply_impl(const LinOp *b, LinOp *x) const override
    {
        auto bb = as<Dense<V>>(b);
        auto xx = as<Dense<V>>(x);
        const I *rp = get_const_row_ptrs();
        const I *ci = get_const_col_idxs();
        const V *va = get_const_values();
        const size_type ncol = xx->get_size()[1];
        for (size_type r = 0; r < size_[0]; ++r)
            for (size_type c = 0; c < ncol; ++c) {
                V acc = zero<V>();
                for (I k = rp[r]; k < rp[r + 1]; ++k) acc += va[k] * bb->at(ci[k], c);
                xx->at(r, c) = acc;
            }
    }
    // x = alpha A b + beta x : x(r) = beta*x(r), then += alpha * a_rk * b_k in stored order.
    // NOTE: no bounds check on b on purpose - update_boundary (restricted_schwarz.cpp:1008-
    // 1015) applies an interface matrix with GLOBAL column ids to a short view of x.
    void apply_impl(const LinOp *alpha, const LinOp *b, const LinOp *beta, LinOp *x) const override
    {
        auto bb = as<Dense<V>>(b);
        auto xx = as<Dense<V>>(x);
        const V al = as<Dense<V>>(alpha)->at(0, 0);
        const V be = as<Dense<V>>(beta)->at(0, 0);
        const I *rp = get_const_row_ptrs();
        const I *ci = get_const_col_idxs();
        const V *va = get_const_values();
        const size_type ncol = xx->get_size()[1];
        const size_type bs = bb->get_stride();
        const V *bv = bb->get_const_values();
        for (size_type r = 0; r < size_[0]; ++r)
            for (size_type c = 0; c < ncol; ++c) {
                V acc = xx->at(r, c) * be;
                for (I k = rp[r]; k < rp[r + 1]; ++k)
                    acc += al * va[k] * bv[static_cast<size_type>(ci[k]) * bs + c];
                xx->at(r, c) = acc;
            }
    }
    Array<V> values_;
    Array<I> col_idxs_;
    Array<I> row_ptrs_;
};

// ---------------------------------------------------------------------- Permutation
using mask_type = std::uint64_t;
static constexpr mask_type row_permute = mask_type{1};
static constexpr mask_type column_permute = mask_type{1 << 2};
static constexpr mask_type inverse_permute = mask_type{1 << 3};

template <typename I = int32>
class Permutation : public LinOp {
public:
    using index_type = I;
    static std::unique_ptr<Permutation> create(std::shared_ptr<const Executor> exec,
                                               const dim<2> &size, Array<I> perm,
                                               mask_type mask = row_permute)
    {
        auto p = std::unique_ptr<Permutation>(new Permutation(exec, size));
        p->perm_ = std::move(perm);
        p->mask_ = mask;
        return p;
    }
    I *get_permutation() { return perm_.get_data(); }
    const I *get_permutation() const { return perm_.get_const_data(); }
    const I *get_const_permutation() const { return perm_.get_const_data(); }
    size_type get_permutation_size() const { return perm_.get_num_elems(); }
    mask_type get_permute_mask() const { return mask_; }

protected:
    Permutation(std::shared_ptr<const Executor> exec, const dim<2> &size) : LinOp(exec, size) {}
    // row_permute: out[i] = in[perm[i]]; with inverse_permute: out[perm[i]] = in[i]
    void apply_impl(const LinOp *in, LinOp *out) const override
    {
        const I *p = get_const_permutation();
        const size_type n = get_permutation_size();
        auto dispatch = [&](auto *src, auto *dst) {
            const size_type nc = src->get_size()[1];
            for (size_type i = 0; i < n; ++i)
                for (size_type c = 0; c < nc; ++c) {
                    if (mask_ & inverse_permute)
                        dst->at(p[i], c) = src->at(i, c);
                    else
                        dst->at(i, c) = src->at(p[i], c);
                }
        };
        if (auto s = dynamic_cast<const Dense<double> *>(in))
            dispatch(s, as<Dense<double>>(out));
        else if (auto s2 = dynamic_cast<const Dense<float> *>(in))
            dispatch(s2, as<Dense<float>>(out));
        else
            throw NotSupported("Permutation::apply on a non-Dense operand");
    }
    void apply_impl(const LinOp *, const LinOp *, const LinOp *, LinOp *) const override
    {
        throw NotSupported("advanced Permutation::apply");
    }
    Array<I> perm_;
    mask_type mask_ = row_permute;
};

}  // namespace matrix

inline void log::Record::on_criterion_check_completed(size_type num_iterations,
                                                      const LinOp *residual) const
{
    auto d = std::unique_ptr<criterion_data>(new criterion_data());
    d->num_iterations = num_iterations;
    if (residual) d->residual = residual->clone();
    data_.criterion_check_completed.push_back(std::move(d));
    while (max_storage_ && data_.criterion_check_completed.size() > max_storage_)
        data_.criterion_check_completed.pop_front();
}

template <typename M, typename... Args>
inline std::unique_ptr<M> initialize(std::initializer_list<typename M::value_type> vals,
                                     std::shared_ptr<const Executor> exec, Args &&...)
{
    auto m = M::create(std::move(exec), dim<2>(vals.size(), 1));
    size_type i = 0;
    for (auto v : vals) m->at(i++, 0) = v;
    return m;
}

template <typename M, typename Stream, typename... Args>
inline std::unique_ptr<M> read(Stream &&is, std::shared_ptr<const Executor> exec, Args &&...)
{
    return M::read_mtx(is, std::move(exec));
}

// ------------------------------------------------------------------- stop criteria
namespace stop {

class CriterionFactory : public Loggable {
public:
    virtual ~CriterionFactory() = default;
    // folds this factory into (max_iters, reduction_factor); "no limit" = max / 0
    virtual void collect(size_type &max_iters, double &reduction,
                         std::vector<std::shared_ptr<const log::Logger>> &loggers) const = 0;

protected:
    void collect_loggers(std::vector<std::shared_ptr<const log::Logger>> &out) const
    {
        out.insert(out.end(), loggers_.begin(), loggers_.end());
    }
};

// the builder idiom: Type::build().with_x(..).on(exec) -> unique_ptr<Type::Factory>
#define GKO_SHIM_ON(FactoryT)                                                      \
    std::unique_ptr<FactoryT> on(std::shared_ptr<const Executor> exec) const      \
    {                                                                              \
        return std::unique_ptr<FactoryT>(new FactoryT(std::move(exec), *this));    \
    }

class Iteration {
public:
    class Factory;
    struct parameters_type {
        size_type max_iters = 0;
        parameters_type &with_max_iters(size_type n)
        {
            max_iters = n;
            return *this;
        }
        GKO_SHIM_ON(Factory)
    };
    class Factory : public CriterionFactory {
    public:
        Factory(std::shared_ptr<const Executor>, const parameters_type &p) : params_(p) {}
        const parameters_type &get_parameters() const { return params_; }
        void collect(size_type &max_iters, double &,
                     std::vector<std::shared_ptr<const log::Logger>> &lg) const override
        {
            max_iters = std::min(max_iters, params_.max_iters);
            collect_loggers(lg);
        }

    private:
        parameters_type params_;
    };
    static parameters_type build() { return {}; }
};

template <typename V = default_precision>
class ResidualNormReduction {
public:
    class Factory;
    struct parameters_type {
        V reduction_factor = static_cast<V>(1e-15);
        parameters_type &with_reduction_factor(V f)
        {
            reduction_factor = f;
            return *this;
        }
        GKO_SHIM_ON(Factory)
    };
    class Factory : public CriterionFactory {
    public:
        Factory(std::shared_ptr<const Executor>, const parameters_type &p) : params_(p) {}
        const parameters_type &get_parameters() const { return params_; }
        void collect(size_type &, double &reduction,
                     std::vector<std::shared_ptr<const log::Logger>> &lg) const override
        {
            reduction = std::max(reduction, static_cast<double>(params_.reduction_factor));
            collect_loggers(lg);
        }

    private:
        parameters_type params_;
    };
    static parameters_type build() { return {}; }
};

class Combined {
public:
    class Factory;
    struct parameters_type {
        std::vector<std::shared_ptr<const CriterionFactory>> criteria;
        template <typename... Args>
        parameters_type &with_criteria(Args &&... args)
        {
            criteria = {std::shared_ptr<const CriterionFactory>(std::forward<Args>(args))...};
            return *this;
        }
        GKO_SHIM_ON(Factory)
    };
    class Factory : public CriterionFactory {
    public:
        Factory(std::shared_ptr<const Executor>, const parameters_type &p) : params_(p) {}
        const parameters_type &get_parameters() const { return params_; }
        void collect(size_type &max_iters, double &reduction,
                     std::vector<std::shared_ptr<const log::Logger>> &lg) const override
        {
            for (auto &c : params_.criteria) c->collect(max_iters, reduction, lg);
            collect_loggers(lg);
        }

    private:
        parameters_type params_;
    };
    static parameters_type build() { return {}; }
};

}  // namespace stop

// -------------------------------------------------------------------- factorization
// ParILU on the reference executor [upstream-memory, Ginkgo 1.2/1.3 reference kernels]:
// L takes the strictly lower part of A plus a unit diagonal, U the upper part incl. the
// diagonal (a missing diagonal entry is added as a zero and initialised to one in U); then
// `iterations` (default 0 -> 1) sequential sweeps of the Chow-Patel fixed-point update over
// the entries of A in row-major order.  One sweep in that order reproduces ILU(0) exactly.
namespace factorization {

template <typename V = default_precision, typename I = int32>
class ParIlu : public LinOp {   // stands for the Composition(L, U) upstream returns
public:
    class Factory;
    struct parameters_type {
        size_type iterations = 0;
        bool skip_sorting = false;
        parameters_type &with_iterations(size_type n)
        {
            iterations = n;
            return *this;
        }
        parameters_type &with_skip_sorting(bool b)
        {
            skip_sorting = b;
            return *this;
        }
        GKO_SHIM_ON(Factory)
    };
    class Factory : public LinOpFactory {
    public:
        Factory(std::shared_ptr<const Executor> e, const parameters_type &p)
            : LinOpFactory(std::move(e)), params_(p)
        {}
        std::unique_ptr<ParIlu> generate(std::shared_ptr<const LinOp> op) const
        {
            return std::unique_ptr<ParIlu>(new ParIlu(exec_, params_, std::move(op)));
        }

    protected:
        std::unique_ptr<LinOp> generate_impl(std::shared_ptr<const LinOp> op) const override
        {
            return generate(std::move(op));
        }
        parameters_type params_;
    };
    static parameters_type build() { return {}; }
    std::shared_ptr<const matrix::Csr<V, I>> get_l_factor() const { return l_; }
    std::shared_ptr<const matrix::Csr<V, I>> get_u_factor() const { return u_; }

protected:
    ParIlu(std::shared_ptr<const Executor> e, const parameters_type &p,
           std::shared_ptr<const LinOp> op)
        : LinOp(e, op->get_size())
    {
        auto A0 = as<matrix::Csr<V, I>>(op.get());
        auto As = matrix::Csr<V, I>::create(e);
        As->copy_from(A0);
        if (!p.skip_sorting) As->sort_by_column_index();
        const size_type n = As->get_size()[0];
        // add_diagonal_elements: rows without a stored diagonal get an explicit zero
        std::vector<I> rp(n + 1, 0), ci;
        std::vector<V> va;
        {
            const I *arp = As->get_const_row_ptrs();
            const I *aci = As->get_const_col_idxs();
            const V *ava = As->get_const_values();
            for (size_type r = 0; r < n; ++r) {
                bool placed = false;
                for (I k = arp[r]; k < arp[r + 1]; ++k) {
                    if (!placed && static_cast<size_type>(aci[k]) > r) {
                        ci.push_back(static_cast<I>(r));
                        va.push_back(zero<V>());
                        placed = true;
                    }
                    if (static_cast<size_type>(aci[k]) == r) placed = true;
                    ci.push_back(aci[k]);
                    va.push_back(ava[k]);
                }
                if (!placed) {
                    ci.push_back(static_cast<I>(r));
                    va.push_back(zero<V>());
                }
                rp[r + 1] = static_cast<I>(ci.size());
            }
        }
        // initialize_row_ptrs_l_u + initialize_l_u
        std::vector<I> lrp(n + 1, 0), urp(n + 1, 0);
        for (size_type r = 0; r < n; ++r) {
            I nl = 0, nu = 0;
            for (I k = rp[r]; k < rp[r + 1]; ++k) {
                if (static_cast<size_type>(ci[k]) < r) ++nl;
                if (static_cast<size_type>(ci[k]) > r) ++nu;
            }
            lrp[r + 1] = lrp[r] + nl + 1;
            urp[r + 1] = urp[r] + nu + 1;
        }
        auto L = matrix::Csr<V, I>::create(e, dim<2>(n, n), static_cast<size_type>(lrp[n]));
        auto U = matrix::Csr<V, I>::create(e, dim<2>(n, n), static_cast<size_type>(urp[n]));
        std::copy(lrp.begin(), lrp.end(), L->get_row_ptrs());
        std::copy(urp.begin(), urp.end(), U->get_row_ptrs());
        for (size_type r = 0; r < n; ++r) {
            I il = lrp[r], iu = urp[r] + 1;   // U: the diagonal goes first
            V diag = one<V>();
            for (I k = rp[r]; k < rp[r + 1]; ++k) {
                const size_type c = static_cast<size_type>(ci[k]);
                if (c < r) {
                    L->get_col_idxs()[il] = ci[k];
                    L->get_values()[il] = va[k];
                    ++il;
                } else if (c == r) {
                    diag = va[k];
                } else {
                    U->get_col_idxs()[iu] = ci[k];
                    U->get_values()[iu] = va[k];
                    ++iu;
                }
            }
            L->get_col_idxs()[lrp[r + 1] - 1] = static_cast<I>(r);
            L->get_values()[lrp[r + 1] - 1] = one<V>();
            U->get_col_idxs()[urp[r]] = static_cast<I>(r);
            U->get_values()[urp[r]] = diag != zero<V>() ? diag : one<V>();
        }
        // the sweep works on U stored by columns (CSR of U^T), exactly as upstream
        auto Ut = std::unique_ptr<matrix::Csr<V, I>>(
            static_cast<matrix::Csr<V, I> *>(U->transpose().release()));
        const I *lr = L->get_const_row_ptrs(), *lc = L->get_const_col_idxs();
        const I *ur = Ut->get_const_row_ptrs(), *uc = Ut->get_const_col_idxs();
        V *lv = L->get_values(), *uv = Ut->get_values();
        const size_type sweeps = p.iterations == 0 ? 1 : p.iterations;
        for (size_type it = 0; it < sweeps; ++it)
            for (size_type row = 0; row < n; ++row)
                for (I el = rp[row]; el < rp[row + 1]; ++el) {
                    const size_type col = static_cast<size_type>(ci[el]);
                    I row_l = lr[row], row_u = ur[col];
                    V sum = va[el], last_operation = zero<V>();
                    while (row_l < lr[row + 1] && row_u < ur[col + 1]) {
                        const I col_l = lc[row_l], col_u = uc[row_u];
                        if (col_l == col_u) {
                            last_operation = lv[row_l] * uv[row_u];
                            sum -= last_operation;
                        } else {
                            last_operation = zero<V>();
                        }
                        if (col_l <= col_u) ++row_l;
                        if (col_u <= col_l) ++row_u;
                    }
                    sum += last_operation;   // undo the term that contains the unknown itself
                    if (row > col) {
                        const V to_write = sum / uv[ur[col + 1] - 1];
                        if (std::isfinite(to_write)) lv[row_l - 1] = to_write;
                    } else {
                        if (std::isfinite(sum)) uv[row_u - 1] = sum;
                    }
                }
        l_ = std::move(L);
        u_ = std::shared_ptr<matrix::Csr<V, I>>(
            static_cast<matrix::Csr<V, I> *>(Ut->transpose().release()));
    }
    void apply_impl(const LinOp *, LinOp *) const override
    {
        throw NotSupported("ParIlu composition apply");
    }
    void apply_impl(const LinOp *, const LinOp *, const LinOp *, LinOp *) const override
    {
        throw NotSupported("ParIlu composition apply");
    }
    std::shared_ptr<matrix::Csr<V, I>> l_, u_;
};

}  // namespace factorization

// -------------------------------------------------------------------- preconditioners
namespace preconditioner {

// Block-Jacobi [upstream-memory, Ginkgo reference kernels jacobi::find_blocks / generate /
// simple_apply]: natural blocks = runs of consecutive rows with identical column pattern
// (capped at max_block_size), agglomerated greedily while the sum stays <= max_block_size;
// every diagonal block is inverted by in-place Gauss-Jordan with implicit row pivoting and
// applied as a dense block times vector, inner index ascending.
template <typename V = default_precision, typename I = int32>
class Jacobi : public LinOp {
public:
    class Factory;
    struct parameters_type {
        unsigned max_block_size = 32u;
        parameters_type &with_max_block_size(unsigned n)
        {
            max_block_size = n;
            return *this;
        }
        GKO_SHIM_ON(Factory)
    };
    class Factory : public LinOpFactory {
    public:
        Factory(std::shared_ptr<const Executor> e, const parameters_type &p)
            : LinOpFactory(std::move(e)), params_(p)
        {}
        std::unique_ptr<Jacobi> generate(std::shared_ptr<const LinOp> op) const
        {
            return std::unique_ptr<Jacobi>(new Jacobi(exec_, params_, std::move(op)));
        }

    protected:
        std::unique_ptr<LinOp> generate_impl(std::shared_ptr<const LinOp> op) const override
        {
            return generate(std::move(op));
        }
        parameters_type params_;
    };
    static parameters_type build() { return {}; }
    size_type get_num_blocks() const { return block_ptrs_.size() - 1; }
    const std::vector<I> &get_block_pointers() const { return block_ptrs_; }
    const std::vector<V> &get_blocks() const { return blocks_; }   // column-major, ld = block size

protected:
    Jacobi(std::shared_ptr<const Executor> e, const parameters_type &p,
           std::shared_ptr<const LinOp> op)
        : LinOp(e, op->get_size())
    {
        if (p.max_block_size < 1 || p.max_block_size > 32)
            throw NotSupported("Jacobi max_block_size outside [1, 32]");
        auto A = as<matrix::Csr<V, I>>(op.get());
        const size_type n = A->get_size()[0];
        const I *rp = A->get_const_row_ptrs();
        const I *ci = A->get_const_col_idxs();
        const V *va = A->get_const_values();
        const I mbs = static_cast<I>(p.max_block_size);
        // find_natural_blocks
        std::vector<I> bp(n + 1, 0);
        size_type nb = 0;
        if (n > 0) {
            nb = 1;
            I cur = 1;
            for (size_type i = 1; i < n; ++i) {
                const I *prev = ci + rp[i - 1], *curr = ci + rp[i], *next = ci + rp[i + 1];
                const bool same = (next - curr) == (curr - prev) && std::equal(curr, next, prev);
                if (cur < mbs && same) {
                    ++cur;
                } else {
                    bp[nb] = bp[nb - 1] + cur;
                    ++nb;
                    cur = 1;
                }
            }
            bp[nb] = bp[nb - 1] + cur;
        }
        // agglomerate_supervariables
        size_type na = 0;
        if (nb > 0) {
            na = 1;
            I cur = bp[1] - bp[0];
            for (size_type i = 1; i < nb; ++i) {
                const I bs = bp[i + 1] - bp[i];
                if (cur + bs <= mbs) {
                    cur += bs;
                } else {
                    bp[na] = bp[na - 1] + cur;
                    ++na;
                    cur = bs;
                }
            }
            bp[na] = bp[na - 1] + cur;
        }
        block_ptrs_.assign(bp.begin(), bp.begin() + na + 1);
        block_off_.assign(na + 1, 0);
        for (size_type b = 0; b < na; ++b) {
            const size_type bs = static_cast<size_type>(block_ptrs_[b + 1] - block_ptrs_[b]);
            block_off_[b + 1] = block_off_[b] + bs * bs;
        }
        blocks_.assign(block_off_[na], zero<V>());
        std::vector<V> blk;
        std::vector<I> perm;
        for (size_type b = 0; b < na; ++b) {
            const I r0 = block_ptrs_[b];
            const size_type bs = static_cast<size_type>(block_ptrs_[b + 1] - r0);
            blk.assign(bs * bs, zero<V>());   // row-major work copy
            for (size_type i = 0; i < bs; ++i)
                for (I k = rp[r0 + i]; k < rp[r0 + i + 1]; ++k)
                    if (ci[k] >= r0 && ci[k] < r0 + static_cast<I>(bs))
                        blk[i * bs + (ci[k] - r0)] = va[k];
            perm.resize(bs);
            for (size_type i = 0; i < bs; ++i) perm[i] = static_cast<I>(i);
            for (size_type k = 0; k < bs; ++k) {
                size_type cp = k;   // choose_pivot: first row of maximal magnitude in column k
                for (size_type i = k + 1; i < bs; ++i)
                    if (std::abs(blk[cp * bs + k]) < std::abs(blk[i * bs + k])) cp = i;
                for (size_type j = 0; j < bs; ++j) std::swap(blk[k * bs + j], blk[cp * bs + j]);
                std::swap(perm[k], perm[cp]);
                // apply_gauss_jordan_transform(k, k)
                const V d = blk[k * bs + k];
                for (size_type i = 0; i < bs; ++i) blk[i * bs + k] /= -d;
                blk[k * bs + k] = zero<V>();
                for (size_type i = 0; i < bs; ++i)
                    for (size_type j = 0; j < bs; ++j)
                        blk[i * bs + j] += blk[i * bs + k] * blk[k * bs + j];
                for (size_type j = 0; j < bs; ++j) blk[k * bs + j] /= d;
                blk[k * bs + k] = one<V>() / d;
            }
            // undo the row permutation on the columns and store column-major
            V *out = blocks_.data() + block_off_[b];
            for (size_type i = 0; i < bs; ++i)
                for (size_type j = 0; j < bs; ++j) out[perm[j] * bs + i] = blk[i * bs + j];
        }
    }
    void apply_impl(const LinOp *b, LinOp *x) const override
    {
        auto bb = as<matrix::Dense<V>>(b);
        auto xx = as<matrix::Dense<V>>(x);
        const size_type nb = block_ptrs_.size() - 1;
        for (size_type blk = 0; blk < nb; ++blk) {
            const size_type r0 = static_cast<size_type>(block_ptrs_[blk]);
            const size_type bs = static_cast<size_type>(block_ptrs_[blk + 1]) - r0;
            const V *inv = blocks_.data() + block_off_[blk];
            for (size_type c = 0; c < xx->get_size()[1]; ++c) {
                for (size_type i = 0; i < bs; ++i) xx->at(r0 + i, c) = zero<V>();
                for (size_type inner = 0; inner < bs; ++inner)
                    for (size_type i = 0; i < bs; ++i)
                        xx->at(r0 + i, c) += inv[inner * bs + i] * bb->at(r0 + inner, c);
            }
        }
    }
    void apply_impl(const LinOp *, const LinOp *, const LinOp *, LinOp *) const override
    {
        throw NotSupported("advanced Jacobi::apply");
    }
    std::vector<I> block_ptrs_;
    std::vector<size_type> block_off_;
    std::vector<V> blocks_;
};

// ISAI of a triangular factor [upstream-memory, isai::generate_tri_inverse on the reference
// executor, sparsity_power = 1]: row i of the approximate inverse M has the pattern J of row
// i of the factor T and solves (M T)(i, J) = e_i(J) through the dense |J| x |J| system
// T(J, J)^T m = e.  Applying it is one SpMV.
template <typename V, typename I, bool Lower>
class IsaiBase : public LinOp {
public:
    std::shared_ptr<const matrix::Csr<V, I>> get_approximate_inverse() const { return inv_; }

protected:
    IsaiBase(std::shared_ptr<const Executor> e, std::shared_ptr<const LinOp> op)
        : LinOp(e, op->get_size())
    {
        auto T0 = as<matrix::Csr<V, I>>(op.get());
        auto T = matrix::Csr<V, I>::create(e);
        T->copy_from(T0);
        T->sort_by_column_index();
        auto M = matrix::Csr<V, I>::create(e);
        M->copy_from(T.get());
        const size_type n = T->get_size()[0];
        const I *rp = T->get_const_row_ptrs();
        const I *ci = T->get_const_col_idxs();
        const V *tv = T->get_const_values();
        V *mv = M->get_values();
        std::vector<V> tri, rhs;
        for (size_type row = 0; row < n; ++row) {
            const I b = rp[row];
            const int sz = static_cast<int>(rp[row + 1] - b);
            tri.assign(static_cast<size_t>(sz) * sz, zero<V>());
            for (int i = 0; i < sz; ++i) {
                const I r2 = ci[b + i];
                I ka = rp[r2], kb = b;   // forall_matching(row r2 of T, pattern)
                while (ka < rp[r2 + 1] && kb < rp[row + 1]) {
                    if (ci[ka] == ci[kb]) {
                        tri[static_cast<size_t>(i) * sz + (kb - b)] = tv[ka];
                        ++ka;
                        ++kb;
                    } else if (ci[ka] < ci[kb]) {
                        ++ka;
                    } else {
                        ++kb;
                    }
                }
            }
            rhs.assign(sz, zero<V>());
            if (sz > 0) {
                if (Lower) {
                    rhs[sz - 1] = one<V>();
                    for (int col = sz - 1; col >= 0; --col) {
                        const V bot = rhs[col] / tri[static_cast<size_t>(col) * sz + col];
                        rhs[col] = bot;
                        for (int r = col - 1; r >= 0; --r)
                            rhs[r] -= bot * tri[static_cast<size_t>(col) * sz + r];
                    }
                } else {
                    rhs[0] = one<V>();
                    for (int col = 0; col < sz; ++col) {
                        const V top = rhs[col] / tri[static_cast<size_t>(col) * sz + col];
                        rhs[col] = top;
                        for (int r = col + 1; r < sz; ++r)
                            rhs[r] -= top * tri[static_cast<size_t>(col) * sz + r];
                    }
                }
            }
            bool finite = true;
            for (int i = 0; i < sz; ++i) finite = finite && std::isfinite(rhs[i]);
            for (int i = 0; i < sz; ++i)
                mv[b + i] = finite ? rhs[i]
                                   : (static_cast<size_type>(ci[b + i]) == row ? one<V>() : zero<V>());
        }
        inv_ = std::move(M);
    }
    void apply_impl(const LinOp *b, LinOp *x) const override { inv_->apply(b, x); }
    void apply_impl(const LinOp *alpha, const LinOp *b, const LinOp *beta, LinOp *x) const override
    {
        inv_->apply(alpha, b, beta, x);
    }
    std::shared_ptr<matrix::Csr<V, I>> inv_;
};

#define GKO_SHIM_ISAI(Name, lower)                                                          \
    template <typename V = default_precision, typename I = int32>                           \
    class Name : public IsaiBase<V, I, lower> {                                             \
    public:                                                                                 \
        class Factory;                                                                      \
        struct parameters_type {                                                            \
            GKO_SHIM_ON(Factory)                                                            \
        };                                                                                  \
        class Factory : public LinOpFactory {                                               \
        public:                                                                             \
            Factory(std::shared_ptr<const Executor> e, const parameters_type &)             \
                : LinOpFactory(std::move(e))                                                \
            {}                                                                              \
            std::unique_ptr<Name> generate(std::shared_ptr<const LinOp> op) const           \
            {                                                                               \
                return std::unique_ptr<Name>(new Name(exec_, std::move(op)));               \
            }                                                                               \
                                                                                            \
        protected:                                                                          \
            std::unique_ptr<LinOp> generate_impl(std::shared_ptr<const LinOp> op) const override \
            {                                                                               \
                return generate(std::move(op));                                             \
            }                                                                               \
        };                                                                                  \
        static parameters_type build() { return {}; }                                       \
                                                                                            \
    protected:                                                                              \
        Name(std::shared_ptr<const Executor> e, std::shared_ptr<const LinOp> op)            \
            : IsaiBase<V, I, lower>(std::move(e), std::move(op))                            \
        {}                                                                                  \
    }
GKO_SHIM_ISAI(LowerIsai, true);
GKO_SHIM_ISAI(UpperIsai, false);

// Ilu<LSolve, USolve>: x = USolve(LSolve(b)).  Generated from a ParIlu result, or from a plain
// matrix (then ParILU is run first, as upstream does for a non-composition operand).
template <typename L, typename U, bool ReverseApply = false, typename I = int32>
class Ilu : public LinOp {
public:
    class Factory;
    struct parameters_type {
        GKO_SHIM_ON(Factory)
    };
    class Factory : public LinOpFactory {
    public:
        Factory(std::shared_ptr<const Executor> e, const parameters_type &)
            : LinOpFactory(std::move(e))
        {}
        std::unique_ptr<Ilu> generate(std::shared_ptr<const LinOp> op) const
        {
            return std::unique_ptr<Ilu>(new Ilu(exec_, std::move(op)));
        }

    protected:
        std::unique_ptr<LinOp> generate_impl(std::shared_ptr<const LinOp> op) const override
        {
            return generate(std::move(op));
        }
    };
    static parameters_type build() { return {}; }
    std::shared_ptr<const L> get_l_solver() const { return l_solver_; }
    std::shared_ptr<const U> get_u_solver() const { return u_solver_; }

protected:
    using V = default_precision;
    Ilu(std::shared_ptr<const Executor> e, std::shared_ptr<const LinOp> op)
        : LinOp(e, op->get_size())
    {
        static_assert(!ReverseApply, "reverse apply is not used by the reference");
        auto fact = std::dynamic_pointer_cast<const factorization::ParIlu<V, I>>(op);
        if (!fact)
            fact = std::shared_ptr<const factorization::ParIlu<V, I>>(
                factorization::ParIlu<V, I>::build().on(e)->generate(op));
        l_solver_ = std::shared_ptr<const L>(L::build().on(e)->generate(fact->get_l_factor()));
        u_solver_ = std::shared_ptr<const U>(U::build().on(e)->generate(fact->get_u_factor()));
    }
    void apply_impl(const LinOp *b, LinOp *x) const override
    {
        auto bb = as<matrix::Dense<V>>(b);
        auto tmp = matrix::Dense<V>::create(exec_, bb->get_size());
        l_solver_->apply(b, tmp.get());
        u_solver_->apply(tmp.get(), x);
    }
    void apply_impl(const LinOp *, const LinOp *, const LinOp *, LinOp *) const override
    {
        throw NotSupported("advanced Ilu::apply");
    }
    std::shared_ptr<const L> l_solver_;
    std::shared_ptr<const U> u_solver_;
};

}  // namespace preconditioner

// ---------------------------------------------------------------------------- solvers
namespace solver {

// parameters shared by Cg and Gmres
template <typename Params>
struct iterative_params {
    std::vector<std::shared_ptr<const stop::CriterionFactory>> criteria;
    std::shared_ptr<const LinOpFactory> preconditioner;
    std::shared_ptr<const LinOp> generated_preconditioner;
    template <typename... Args>
    Params &with_criteria(Args &&... args)
    {
        criteria = {std::shared_ptr<const stop::CriterionFactory>(std::forward<Args>(args))...};
        return static_cast<Params &>(*this);
    }
    template <typename F>
    Params &with_preconditioner(F &&f)
    {
        preconditioner = std::shared_ptr<const LinOpFactory>(std::forward<F>(f));
        return static_cast<Params &>(*this);
    }
    Params &with_generated_preconditioner(std::shared_ptr<const LinOp> p)
    {
        generated_preconditioner = std::move(p);
        return static_cast<Params &>(*this);
    }
};

template <typename V>
class IterativeBase : public LinOp {
public:
    std::shared_ptr<const LinOp> get_system_matrix() const { return system_; }
    std::shared_ptr<const LinOp> get_preconditioner() const { return precond_; }
    // fork-only API used at solve.cpp:747-751
    void set_stop_criterion_factory(std::shared_ptr<const stop::CriterionFactory> f)
    {
        criteria_ = {std::move(f)};
    }

protected:
    template <typename P>
    IterativeBase(std::shared_ptr<const Executor> e, const P &p, std::shared_ptr<const LinOp> A)
        : LinOp(e, transpose(A->get_size())), system_(std::move(A)), criteria_(p.criteria)
    {
        if (p.generated_preconditioner)
            precond_ = p.generated_preconditioner;
        else if (p.preconditioner)
            precond_ = p.preconditioner->generate(system_);
    }
    struct stop_state {
        size_type max_iters = static_cast<size_type>(-1);
        double reduction = 0.0;
        std::vector<std::shared_ptr<const log::Logger>> loggers;
    };
    stop_state make_stop() const
    {
        stop_state s;
        for (auto &c : criteria_) c->collect(s.max_iters, s.reduction, s.loggers);
        return s;
    }
    // Combined(Iteration, ResidualNormReduction): iter >= max OR tau < factor * tau0
    static bool check(const stop_state &s, size_type iter, V tau, V tau0, const LinOp *residual)
    {
        for (auto &l : s.loggers) l->on_criterion_check_completed(iter, residual);
        return iter >= s.max_iters || (s.reduction > 0.0 && tau < static_cast<V>(s.reduction) * tau0);
    }
    void apply_impl(const LinOp *alpha, const LinOp *b, const LinOp *beta, LinOp *x) const override
    {
        auto xx = as<matrix::Dense<V>>(x);
        auto tmp = xx->clone();
        this->apply_impl(b, tmp.get());
        xx->scale(beta);
        xx->add_scaled(alpha, tmp.get());
    }
    using LinOp::apply_impl;
    std::shared_ptr<const LinOp> system_;
    std::shared_ptr<const LinOp> precond_;
    std::vector<std::shared_ptr<const stop::CriterionFactory>> criteria_;
};

#define GKO_SHIM_SOLVER_FACTORY(Solver)                                                       \
    class Factory : public LinOpFactory {                                                     \
    public:                                                                                   \
        Factory(std::shared_ptr<const Executor> e, const parameters_type &p)                  \
            : LinOpFactory(std::move(e)), params_(p)                                          \
        {}                                                                                    \
        std::unique_ptr<Solver> generate(std::shared_ptr<const LinOp> A) const                \
        {                                                                                     \
            return std::unique_ptr<Solver>(new Solver(exec_, params_, std::move(A)));         \
        }                                                                                     \
        const parameters_type &get_parameters() const { return params_; }                     \
                                                                                              \
    protected:                                                                                \
        std::unique_ptr<LinOp> generate_impl(std::shared_ptr<const LinOp> A) const override   \
        {                                                                                     \
            return generate(std::move(A));                                                    \
        }                                                                                     \
        parameters_type params_;                                                              \
    };                                                                                        \
    static parameters_type build() { return {}; }

// CG with Ginkgo's recurrences: r = b - A x, z = p = q = 0,
// prev_rho = 1; loop { z = M^-1 r; rho = r.z; ++iter; stop test on ||r||; p = z +
// (rho/prev_rho) p; q = A p; beta = p.q; x += (rho/beta) p; r -= (rho/beta) q; swap }.
template <typename V = default_precision>
class Cg : public IterativeBase<V> {
public:
    class Factory;
    struct parameters_type : iterative_params<parameters_type> {
        GKO_SHIM_ON(Factory)
    };
    GKO_SHIM_SOLVER_FACTORY(Cg)

protected:
    Cg(std::shared_ptr<const Executor> e, const parameters_type &p, std::shared_ptr<const LinOp> A)
        : IterativeBase<V>(e, p, std::move(A))
    {}
    using IterativeBase<V>::apply_impl;
    void apply_impl(const LinOp *b, LinOp *x) const override
    {
        using Vec = matrix::Dense<V>;
        auto bb = as<Vec>(b);
        auto xx = as<Vec>(x);
        const size_type n = bb->get_size()[0];
        auto exec = this->exec_;
        auto r = Vec::create(exec, dim<2>(n, 1));
        auto z = Vec::create(exec, dim<2>(n, 1));
        auto p = Vec::create(exec, dim<2>(n, 1));
        auto q = Vec::create(exec, dim<2>(n, 1));
        auto one_op = initialize<Vec>({one<V>()}, exec);
        auto neg_one_op = initialize<Vec>({-one<V>()}, exec);
        r->copy_from(bb);
        this->system_->apply(neg_one_op.get(), xx, one_op.get(), r.get());
        auto norm = [&](const Vec *v) {
            V s = zero<V>();
            for (size_type i = 0; i < n; ++i) s += v->at(i, 0) * v->at(i, 0);
            return std::sqrt(s);
        };
        auto dot = [&](const Vec *a, const Vec *c) {
            V s = zero<V>();
            for (size_type i = 0; i < n; ++i) s += a->at(i, 0) * c->at(i, 0);
            return s;
        };
        const V tau0 = norm(r.get());
        auto stop = this->make_stop();
        V rho = zero<V>(), prev_rho = one<V>();
        size_type iter = static_cast<size_type>(-1);
        while (true) {
            if (this->precond_)
                this->precond_->apply(r.get(), z.get());
            else
                z->copy_from(r.get());
            rho = dot(r.get(), z.get());
            ++iter;
            if (IterativeBase<V>::check(stop, iter, norm(r.get()), tau0, r.get())) break;
            if (prev_rho == zero<V>()) {
                p->copy_from(z.get());
            } else {
                const V t = rho / prev_rho;
                for (size_type i = 0; i < n; ++i) p->at(i, 0) = z->at(i, 0) + t * p->at(i, 0);
            }
            this->system_->apply(p.get(), q.get());
            const V beta = dot(p.get(), q.get());
            if (beta != zero<V>()) {
                const V t = rho / beta;
                for (size_type i = 0; i < n; ++i) {
                    xx->at(i, 0) += t * p->at(i, 0);
                    r->at(i, 0) -= t * q->at(i, 0);
                }
            }
            std::swap(prev_rho, rho);
        }
    }
};

// Restarted GMRES(krylov_dim): modified Gram-Schmidt, Givens rotations, the stop test on
// the implicit residual norm, right preconditioning, x updated at restart / termination.
template <typename V = default_precision>
class Gmres : public IterativeBase<V> {
public:
    class Factory;
    struct parameters_type : iterative_params<parameters_type> {
        size_type krylov_dim = 100u;
        parameters_type &with_krylov_dim(size_type m)
        {
            krylov_dim = m;
            return *this;
        }
        GKO_SHIM_ON(Factory)
    };
    GKO_SHIM_SOLVER_FACTORY(Gmres)
    size_type get_krylov_dim() const { return m_; }

protected:
    Gmres(std::shared_ptr<const Executor> e, const parameters_type &p, std::shared_ptr<const LinOp> A)
        : IterativeBase<V>(e, p, std::move(A)), m_(p.krylov_dim ? p.krylov_dim : 1)
    {}
    using IterativeBase<V>::apply_impl;
    void apply_impl(const LinOp *b, LinOp *x) const override
    {
        using Vec = matrix::Dense<V>;
        auto bb = as<Vec>(b);
        auto xx = as<Vec>(x);
        const size_type n = bb->get_size()[0];
        const size_type m = m_;
        auto exec = this->exec_;
        std::vector<std::unique_ptr<Vec>> Vb;
        for (size_type i = 0; i <= m; ++i) Vb.push_back(Vec::create(exec, dim<2>(n, 1)));
        std::vector<V> H((m + 1) * m, zero<V>()), cs(m), sn(m), g(m + 1), y(m);
        auto Hc = [&](size_type i, size_type j) -> V & { return H[j * (m + 1) + i]; };
        auto r = Vec::create(exec, dim<2>(n, 1));
        auto w = Vec::create(exec, dim<2>(n, 1));
        auto pv = Vec::create(exec, dim<2>(n, 1));
        auto upd = Vec::create(exec, dim<2>(n, 1));
        auto one_op = initialize<Vec>({one<V>()}, exec);
        auto neg_one_op = initialize<Vec>({-one<V>()}, exec);
        auto norm = [&](const Vec *v) {
            V s = zero<V>();
            for (size_type i = 0; i < n; ++i) s += v->at(i, 0) * v->at(i, 0);
            return std::sqrt(s);
        };
        auto dot = [&](const Vec *a, const Vec *c) {
            V s = zero<V>();
            for (size_type i = 0; i < n; ++i) s += a->at(i, 0) * c->at(i, 0);
            return s;
        };
        auto restart = [&]() {
            r->copy_from(bb);
            this->system_->apply(neg_one_op.get(), xx, one_op.get(), r.get());
            const V rn = norm(r.get());
            std::fill(g.begin(), g.end(), zero<V>());
            g[0] = rn;
            for (size_type i = 0; i < n; ++i) Vb[0]->at(i, 0) = rn != zero<V>() ? r->at(i, 0) / rn : zero<V>();
            return rn;
        };
        auto update_x = [&](size_type k) {
            for (size_type ii = k; ii-- > 0;) {
                V s = g[ii];
                for (size_type j = ii + 1; j < k; ++j) s -= Hc(ii, j) * y[j];
                y[ii] = s / Hc(ii, ii);
            }
            if (this->precond_) {
                upd->fill(zero<V>());
                for (size_type j = 0; j < k; ++j)
                    for (size_type i = 0; i < n; ++i) upd->at(i, 0) += y[j] * Vb[j]->at(i, 0);
                this->precond_->apply(upd.get(), pv.get());
                for (size_type i = 0; i < n; ++i) xx->at(i, 0) += pv->at(i, 0);
            } else {
                for (size_type j = 0; j < k; ++j)
                    for (size_type i = 0; i < n; ++i) xx->at(i, 0) += y[j] * Vb[j]->at(i, 0);
            }
        };
        auto stop = this->make_stop();
        V resnorm = restart();
        const V tau0 = resnorm;
        size_type total = static_cast<size_type>(-1), k = 0;
        while (true) {
            ++total;
            if (IterativeBase<V>::check(stop, total, resnorm, tau0, r.get())) break;
            if (k == m) {
                update_x(k);
                resnorm = restart();
                k = 0;
            }
            if (this->precond_) {
                this->precond_->apply(Vb[k].get(), pv.get());
                this->system_->apply(pv.get(), w.get());
            } else {
                this->system_->apply(Vb[k].get(), w.get());
            }
            for (size_type i = 0; i <= k; ++i) {
                const V h = dot(w.get(), Vb[i].get());
                Hc(i, k) = h;
                for (size_type t = 0; t < n; ++t) w->at(t, 0) -= h * Vb[i]->at(t, 0);
            }
            const V hn = norm(w.get());
            Hc(k + 1, k) = hn;
            for (size_type t = 0; t < n; ++t) Vb[k + 1]->at(t, 0) = hn != zero<V>() ? w->at(t, 0) / hn : zero<V>();
            for (size_type i = 0; i < k; ++i) {
                const V t = cs[i] * Hc(i, k) + sn[i] * Hc(i + 1, k);
                Hc(i + 1, k) = -sn[i] * Hc(i, k) + cs[i] * Hc(i + 1, k);
                Hc(i, k) = t;
            }
            {
                const V a = Hc(k, k), c = Hc(k + 1, k);
                if (a == zero<V>()) {
                    cs[k] = zero<V>();
                    sn[k] = one<V>();
                } else {
                    const V sc = std::fabs(a) + std::fabs(c);
                    const V hyp = sc * std::sqrt((a / sc) * (a / sc) + (c / sc) * (c / sc));
                    cs[k] = a / hyp;
                    sn[k] = c / hyp;
                }
                Hc(k, k) = cs[k] * a + sn[k] * c;
                Hc(k + 1, k) = zero<V>();
                g[k + 1] = -sn[k] * g[k];
                g[k] = cs[k] * g[k];
                resnorm = std::fabs(g[k + 1]);
            }
            ++k;
        }
        update_x(k);
    }
    size_type m_;
};

// serial forward / backward substitution, diagonal taken from the matrix
template <typename V, typename I, bool Lower>
class TrsBase : public LinOp {
public:
    std::shared_ptr<const matrix::Csr<V, I>> get_system_matrix() const { return system_; }

protected:
    TrsBase(std::shared_ptr<const Executor> e, std::shared_ptr<const LinOp> A)
        : LinOp(e, A->get_size()),
          system_(std::dynamic_pointer_cast<const matrix::Csr<V, I>>(A))
    {
        if (!system_) throw NotSupported("triangular solver on a non-Csr operand");
    }
    void apply_impl(const LinOp *b, LinOp *x) const override
    {
        auto bb = as<matrix::Dense<V>>(b);
        auto xx = as<matrix::Dense<V>>(x);
        const I *rp = system_->get_const_row_ptrs();
        const I *ci = system_->get_const_col_idxs();
        const V *va = system_->get_const_values();
        const size_type n = size_[0];
        for (size_type c = 0; c < xx->get_size()[1]; ++c)
            for (size_type s = 0; s < n; ++s) {
                const size_type r = Lower ? s : n - 1 - s;
                V acc = bb->at(r, c), d = one<V>();
                for (I k = rp[r]; k < rp[r + 1]; ++k) {
                    const size_type col = static_cast<size_type>(ci[k]);
                    if (col == r)
                        d = va[k];
                    else if (Lower ? col < r : col > r)
                        acc -= va[k] * xx->at(col, c);
                }
                xx->at(r, c) = acc / d;
            }
    }
    void apply_impl(const LinOp *, const LinOp *, const LinOp *, LinOp *) const override
    {
        throw NotSupported("advanced triangular apply");
    }
    std::shared_ptr<const matrix::Csr<V, I>> system_;
};

#define GKO_SHIM_TRS(Name, lower)                                                           \
    template <typename V = default_precision, typename I = int32>                           \
    class Name : public TrsBase<V, I, lower> {                                              \
    public:                                                                                 \
        class Factory;                                                                      \
        struct parameters_type {                                                            \
            size_type num_rhs = 1;                                                          \
            parameters_type &with_num_rhs(size_type n)                                      \
            {                                                                               \
                num_rhs = n;                                                                \
                return *this;                                                               \
            }                                                                               \
            GKO_SHIM_ON(Factory)                                                            \
        };                                                                                  \
        GKO_SHIM_SOLVER_FACTORY(Name)                                                       \
    protected:                                                                              \
        Name(std::shared_ptr<const Executor> e, const parameters_type &,                    \
             std::shared_ptr<const LinOp> A)                                                \
            : TrsBase<V, I, lower>(e, std::move(A))                                         \
        {}                                                                                  \
    }
GKO_SHIM_TRS(LowerTrs, true);
GKO_SHIM_TRS(UpperTrs, false);

}  // namespace solver
}  // namespace gko

#endif  // GINKGO_SHIM_HPP_
