// In-process stand-in for the handful of MPI calls the reference's driver and
// facade make (benchmarking/bench_ras.cpp:55-56,206-221; source/
// initialization.cpp:72-73; source/schwarz_base.cpp:116,453).  A "rank" is a
// host thread of this process; there is no MPI on the data path — halo values
// move by peer stores, flags by peer-mapped words.
#pragma once
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace schwz_mpi {

class RankGroup {
public:
    static RankGroup &instance()
    {
        static RankGroup g;
        return g;
    }
    int size() const { return size_; }
    static int &my_rank()
    {
        static thread_local int r = 0;
        return r;
    }
    // runs fn(rank) on `n` threads and joins them; exceptions are rethrown
    void run(int n, const std::function<void(int)> &fn)
    {
        size_ = n;
        arrived_ = 0;
        generation_ = 0;
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> err(n);
        for (int r = 0; r < n; ++r)
            th.emplace_back([&, r] {
                my_rank() = r;
                try {
                    fn(r);
                } catch (...) {
                    err[r] = std::current_exception();
                    abort_barriers();
                }
            });
        for (auto &t : th) t.join();
        size_ = 1;
        for (auto &e : err)
            if (e) std::rethrow_exception(e);
    }
    void barrier()
    {
        std::unique_lock<std::mutex> lk(m_);
        if (aborted_) throw std::runtime_error("rank group aborted");
        const long gen = generation_;
        if (++arrived_ == size_) {
            arrived_ = 0;
            ++generation_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return generation_ != gen || aborted_; });
            if (aborted_) throw std::runtime_error("rank group aborted");
        }
    }
    // slot table shared by the ranks (solver handles, norms, ...)
    std::vector<void *> &slots(int which)
    {
        std::lock_guard<std::mutex> lk(m_);
        // sized once for every slot kind: a later resize would move the tables other ranks
        // hold references to
        if (tables_.empty()) tables_.resize(16);
        if ((int)tables_.size() <= which) throw std::runtime_error("slot table index out of range");
        if ((int)tables_[which].size() < size_) tables_[which].resize(size_, nullptr);
        return tables_[which];
    }
    std::vector<double> &doubles()
    {
        std::lock_guard<std::mutex> lk(m_);
        if ((int)dbl_.size() < size_) dbl_.resize(size_, 0.0);
        return dbl_;
    }

private:
    void abort_barriers()
    {
        std::lock_guard<std::mutex> lk(m_);
        aborted_ = true;
        cv_.notify_all();
    }
    int size_ = 1, arrived_ = 0;
    long generation_ = 0;
    bool aborted_ = false;
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<std::vector<void *>> tables_;
    std::vector<double> dbl_;
};

}  // namespace schwz_mpi

using MPI_Comm = int;
constexpr MPI_Comm MPI_COMM_WORLD = 0;
inline int MPI_Comm_rank(MPI_Comm, int *r)
{
    *r = schwz_mpi::RankGroup::my_rank();
    return 0;
}
inline int MPI_Comm_size(MPI_Comm, int *s)
{
    *s = schwz_mpi::RankGroup::instance().size();
    return 0;
}
inline int MPI_Barrier(MPI_Comm)
{
    schwz_mpi::RankGroup::instance().barrier();
    return 0;
}
inline double MPI_Wtime()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int MPI_Init(int *, char ***) { return 0; }
inline int MPI_Finalize() { return 0; }
