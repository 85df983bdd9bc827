// Minimal stand-in for the Ginkgo types that appear in schwarz-lib's public
// signatures (Settings, Metadata, SchwarzBase, SolverRAS).  Ginkgo is not a
// dependency of this implementation: arithmetic lives behind the C ABI
// (include/schwz_b200.h); these classes are plain host containers so that code
// written against the reference's headers keeps compiling.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace gko {

using size_type = std::size_t;
using int32 = std::int32_t;
using int64 = std::int64_t;
using default_precision = double;

template <int N>
struct dim {
    std::array<size_type, N> v{};
    dim() = default;
    explicit dim(size_type n) { v.fill(n); }
    dim(size_type r, size_type c)
    {
        static_assert(N == 2, "two extents need dim<2>");
        v[0] = r;
        v[1] = c;
    }
    size_type operator[](int i) const { return v[i]; }
};

// ---- executors: only identity and device id matter here ------------------------
class Executor : public std::enable_shared_from_this<Executor> {
public:
    virtual ~Executor() = default;
    virtual std::shared_ptr<Executor> get_master() { return shared_from_this(); }
    virtual bool is_device() const { return false; }
    virtual int get_device_id() const { return -1; }
};
class ReferenceExecutor : public Executor {
public:
    static std::shared_ptr<ReferenceExecutor> create() { return std::make_shared<ReferenceExecutor>(); }
};
class OmpExecutor : public Executor {
public:
    static std::shared_ptr<OmpExecutor> create() { return std::make_shared<OmpExecutor>(); }
};
class CudaExecutor : public Executor {
public:
    CudaExecutor(int id, std::shared_ptr<Executor> master) : id_(id), master_(std::move(master)) {}
    static std::shared_ptr<CudaExecutor> create(int id, std::shared_ptr<Executor> master, bool = false)
    {
        return std::make_shared<CudaExecutor>(id, std::move(master));
    }
    std::shared_ptr<Executor> get_master() override { return master_; }
    bool is_device() const override { return true; }
    int get_device_id() const override { return id_; }

private:
    int id_;
    std::shared_ptr<Executor> master_;
};

// ---- Array -----------------------------------------------------------------------
template <typename T>
class Array {
public:
    Array() = default;
    Array(std::shared_ptr<Executor> exec, size_type n) : exec_(std::move(exec)), data_(n) {}
    template <typename It>
    Array(std::shared_ptr<Executor> exec, It b, It e) : exec_(std::move(exec)), data_(b, e)
    {}
    T *get_data() { return data_.data(); }
    const T *get_const_data() const { return data_.data(); }
    size_type get_num_elems() const { return data_.size(); }
    std::shared_ptr<Executor> get_executor() const { return exec_; }
    void resize(size_type n) { data_.resize(n); }
    std::vector<T> &vec() { return data_; }

private:
    std::shared_ptr<Executor> exec_;
    std::vector<T> data_;
};

namespace matrix {

template <typename V>
class Dense {
public:
    Dense(std::shared_ptr<Executor> exec, dim<2> size)
        : exec_(std::move(exec)), size_(size), v_(size[0] * size[1], V{})
    {}
    static std::shared_ptr<Dense> create(std::shared_ptr<Executor> exec, dim<2> size = dim<2>(0, 0))
    {
        return std::make_shared<Dense>(std::move(exec), size);
    }
    dim<2> get_size() const { return size_; }
    V *get_values() { return v_.data(); }
    const V *get_const_values() const { return v_.data(); }
    V &at(size_type i, size_type j = 0) { return v_[i * size_[1] + j]; }
    const V &at(size_type i, size_type j = 0) const { return v_[i * size_[1] + j]; }
    void copy_from(const Dense *o)
    {
        size_ = o->size_;
        v_ = o->v_;
    }
    std::shared_ptr<Executor> get_executor() const { return exec_; }

private:
    std::shared_ptr<Executor> exec_;
    dim<2> size_;
    std::vector<V> v_;
};

// CSR container.  For big problems the arrays may be left empty ("sizes only"):
// the matrix then lives on the device behind the C ABI and get_size() /
// get_num_stored_elements() still answer.
template <typename V, typename I>
class Csr {
public:
    explicit Csr(std::shared_ptr<Executor> exec, dim<2> size = dim<2>(0, 0), size_type nnz = 0)
        : exec_(std::move(exec)), size_(size), nnz_(nnz)
    {}
    static std::shared_ptr<Csr> create(std::shared_ptr<Executor> exec, dim<2> size = dim<2>(0, 0),
                                       size_type nnz = 0)
    {
        return std::make_shared<Csr>(std::move(exec), size, nnz);
    }
    dim<2> get_size() const { return size_; }
    size_type get_num_stored_elements() const { return nnz_; }
    void allocate()
    {
        rp_.assign(size_[0] + 1, 0);
        ci_.assign(nnz_, 0);
        v_.assign(nnz_, V{});
    }
    bool has_arrays() const { return !rp_.empty(); }
    I *get_row_ptrs() { return rp_.data(); }
    I *get_col_idxs() { return ci_.data(); }
    V *get_values() { return v_.data(); }
    const I *get_const_row_ptrs() const { return rp_.data(); }
    const I *get_const_col_idxs() const { return ci_.data(); }
    const V *get_const_values() const { return v_.data(); }
    void set_shape(dim<2> size, size_type nnz)
    {
        size_ = size;
        nnz_ = nnz;
    }

private:
    std::shared_ptr<Executor> exec_;
    dim<2> size_;
    size_type nnz_;
    std::vector<I> rp_, ci_;
    std::vector<V> v_;
};

template <typename I>
class Permutation {
public:
    Permutation(std::shared_ptr<Executor> exec, std::vector<I> p) : exec_(std::move(exec)), p_(std::move(p)) {}
    static std::shared_ptr<Permutation> create(std::shared_ptr<Executor> exec, std::vector<I> p)
    {
        return std::make_shared<Permutation>(std::move(exec), std::move(p));
    }
    const I *get_const_permutation() const { return p_.data(); }
    size_type get_permutation_size() const { return p_.size(); }

private:
    std::shared_ptr<Executor> exec_;
    std::vector<I> p_;
};

}  // namespace matrix
}  // namespace gko
