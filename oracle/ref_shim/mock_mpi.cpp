// Implementation of the in-process MPI mock declared in ref_shim/mpi.h.
#include <mpi.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <exception>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <tuple>
#include <vector>

struct MPI_Request_impl {
    bool is_recv = false, done = true;
    void *buf = nullptr;
    size_t bytes = 0;
    int src = 0, dst = 0, tag = 0;
};

namespace {
int g_n = 1;
thread_local int t_rank = 0;
thread_local int t_win_counter = 0;
std::mutex g_m;
std::condition_variable g_cv;
int g_arrived = 0;
long g_gen = 0;
bool g_abort = false;
std::map<std::tuple<int, int, int>, std::deque<std::vector<char>>> g_mail;   // (src,dst,tag)
std::vector<const void *> g_coll_ptr;
struct WinEntry { char *base = nullptr; int unit = 1; };
std::map<int, std::vector<WinEntry>> g_win;

size_t dt_size(MPI_Datatype t) { return (size_t)(t & 0xff); }
int dt_kind(MPI_Datatype t) { return t >> 8; }

void barrier()
{
    std::unique_lock<std::mutex> lk(g_m);
    if (g_abort) throw std::runtime_error("mock MPI aborted");
    long gen = g_gen;
    if (++g_arrived == g_n) {
        g_arrived = 0;
        ++g_gen;
        g_cv.notify_all();
    } else {
        g_cv.wait(lk, [&] { return g_gen != gen || g_abort; });
        if (g_abort) throw std::runtime_error("mock MPI aborted");
    }
}

template <typename T>
void combine(T *dst, const T *src, int n, MPI_Op op)
{
    for (int i = 0; i < n; ++i) {
        if (op == MPI_SUM) dst[i] += src[i];
        else if (op == MPI_MIN) dst[i] = src[i] < dst[i] ? src[i] : dst[i];
        else if (op == MPI_MAX) dst[i] = src[i] > dst[i] ? src[i] : dst[i];
    }
}
void combine_dt(void *dst, const void *src, int n, MPI_Datatype t, MPI_Op op)
{
    if (t == MPI_DOUBLE) combine((double *)dst, (const double *)src, n, op);
    else if (t == MPI_FLOAT) combine((float *)dst, (const float *)src, n, op);
    else if (t == MPI_INT) combine((int *)dst, (const int *)src, n, op);
    else if (t == MPI_LONG) combine((long *)dst, (const long *)src, n, op);
    else if (t == MPI_UNSIGNED) combine((unsigned *)dst, (const unsigned *)src, n, op);
    else throw std::runtime_error("mock MPI: reduction on an unsupported datatype");
}
}  // namespace

namespace mockmpi {
void run(int nranks, void (*fn)(int, void *), void *arg)
{
    g_n = nranks;
    g_arrived = 0;
    g_abort = false;
    g_mail.clear();
    g_win.clear();
    g_coll_ptr.assign(nranks, nullptr);
    std::vector<std::thread> th;
    std::vector<std::exception_ptr> err(nranks);
    for (int r = 0; r < nranks; ++r)
        th.emplace_back([&, r] {
            t_rank = r;
            t_win_counter = 0;
            try {
                fn(r, arg);
            } catch (...) {
                err[r] = std::current_exception();
                std::lock_guard<std::mutex> lk(g_m);
                g_abort = true;
                g_cv.notify_all();
            }
        });
    for (auto &t : th) t.join();
    for (auto &e : err)
        if (e) std::rethrow_exception(e);
}
}  // namespace mockmpi

int MPI_Init(int *, char ***) { return 0; }
int MPI_Init_thread(int *, char ***, int req, int *prov)
{
    *prov = req;
    return 0;
}
int MPI_Finalize() { return 0; }
int MPI_Comm_rank(MPI_Comm, int *r)
{
    *r = t_rank;
    return 0;
}
int MPI_Comm_size(MPI_Comm, int *s)
{
    *s = g_n;
    return 0;
}
int MPI_Comm_split_type(MPI_Comm c, int, int, MPI_Info, MPI_Comm *out)
{
    *out = c;   // one node
    return 0;
}
int MPI_Barrier(MPI_Comm)
{
    barrier();
    return 0;
}
double MPI_Wtime()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm)
{
    if (t_rank == root) g_coll_ptr[root] = buf;
    barrier();
    if (t_rank != root) std::memcpy(buf, g_coll_ptr[root], count * dt_size(t));
    barrier();
    return 0;
}
int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request *req)
{
    std::vector<char> msg((const char *)buf, (const char *)buf + count * dt_size(t));
    {
        std::lock_guard<std::mutex> lk(g_m);
        g_mail[std::make_tuple(t_rank, dest, tag)].push_back(std::move(msg));
    }
    g_cv.notify_all();
    *req = new MPI_Request_impl();   // eager: complete at once
    return 0;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Request *req)
{
    auto *r = new MPI_Request_impl();
    r->is_recv = true;
    r->done = false;
    r->buf = buf;
    r->bytes = count * dt_size(t);
    r->src = src;
    r->dst = t_rank;
    r->tag = tag;
    *req = r;
    // The reference scatters out of recv_buffer right after posting its receives and never
    // waits on them (restricted_schwarz.cpp:926-962): with a real MPI the data
    // is there only if the network won the race. The mock resolves the race the way the
    // authors intended - the receive is complete when MPI_Irecv returns (every matching
    // send of the reference is posted before the receive, so this cannot deadlock).
    return MPI_Wait(req, nullptr);
}
int MPI_Wait(MPI_Request *req, MPI_Status *st)
{
    MPI_Request_impl *r = *req;
    if (r == nullptr) return 0;
    if (r->is_recv && !r->done) {
        std::unique_lock<std::mutex> lk(g_m);
        auto key = std::make_tuple(r->src, r->dst, r->tag);
        g_cv.wait(lk, [&] { return g_abort || !g_mail[key].empty(); });
        if (g_abort) throw std::runtime_error("mock MPI aborted");
        auto &msg = g_mail[key].front();
        std::memcpy(r->buf, msg.data(), msg.size() < r->bytes ? msg.size() : r->bytes);
        g_mail[key].pop_front();
        r->done = true;
    }
    if (st) {
        st->source = r->src;
        st->tag = r->tag;
    }
    // requests are small; the reference waits through copies of the handle, so
    // they are intentionally not freed here
    return 0;
}
int MPI_Alltoall(const void *s, int sc, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm)
{
    g_coll_ptr[t_rank] = s;
    barrier();
    const size_t b = sc * dt_size(st);
    for (int q = 0; q < g_n; ++q)
        std::memcpy((char *)r + q * b, (const char *)g_coll_ptr[q] + t_rank * b, b);
    barrier();
    return 0;
}
int MPI_Allgather(const void *s, int sc, MPI_Datatype st, void *r, int, MPI_Datatype, MPI_Comm)
{
    g_coll_ptr[t_rank] = s;
    barrier();
    const size_t b = sc * dt_size(st);
    for (int q = 0; q < g_n; ++q) std::memcpy((char *)r + q * b, g_coll_ptr[q], b);
    barrier();
    return 0;
}
int MPI_Allreduce(const void *s, void *r, int count, MPI_Datatype t, MPI_Op op, MPI_Comm)
{
    g_coll_ptr[t_rank] = s;
    barrier();
    std::vector<char> acc((const char *)g_coll_ptr[0], (const char *)g_coll_ptr[0] + count * dt_size(t));
    for (int q = 1; q < g_n; ++q) combine_dt(acc.data(), g_coll_ptr[q], count, t, op);   // rank order
    barrier();
    std::memcpy(r, acc.data(), acc.size());
    return 0;
}
int MPI_Win_create(void *base, MPI_Aint, int disp_unit, MPI_Info, MPI_Comm, MPI_Win *win)
{
    const int id = t_win_counter++;
    {
        std::lock_guard<std::mutex> lk(g_m);
        auto &v = g_win[id];
        if ((int)v.size() < g_n) v.resize(g_n);
        v[t_rank].base = (char *)base;
        v[t_rank].unit = disp_unit;
    }
    barrier();
    *win = id;
    return 0;
}
int MPI_Win_lock_all(int, MPI_Win) { return 0; }
int MPI_Win_unlock_all(MPI_Win) { return 0; }
int MPI_Win_lock(int, int, int, MPI_Win) { return 0; }
int MPI_Win_unlock(int, MPI_Win) { return 0; }
int MPI_Win_flush(int, MPI_Win) { return 0; }
int MPI_Win_flush_local(int, MPI_Win) { return 0; }
int MPI_Win_free(MPI_Win *) { return 0; }
int MPI_Put(const void *o, int oc, MPI_Datatype ot, int target, MPI_Aint disp, int, MPI_Datatype, MPI_Win w)
{
    std::lock_guard<std::mutex> lk(g_m);
    WinEntry &e = g_win[w][target];
    std::memcpy(e.base + (size_t)disp * e.unit, o, oc * dt_size(ot));
    return 0;
}
int MPI_Get(void *o, int oc, MPI_Datatype ot, int target, MPI_Aint disp, int, MPI_Datatype, MPI_Win w)
{
    std::lock_guard<std::mutex> lk(g_m);
    WinEntry &e = g_win[w][target];
    std::memcpy(o, e.base + (size_t)disp * e.unit, oc * dt_size(ot));
    return 0;
}
int MPI_Accumulate(const void *o, int oc, MPI_Datatype ot, int target, MPI_Aint disp, int, MPI_Datatype,
                   MPI_Op op, MPI_Win w)
{
    std::lock_guard<std::mutex> lk(g_m);
    WinEntry &e = g_win[w][target];
    combine_dt(e.base + (size_t)disp * e.unit, o, oc, ot, op);
    return 0;
}
