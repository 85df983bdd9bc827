// Shared host/device helpers for the sm_100a kernels of the RAS hot path.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace schwz_b200 {

// ---- error convention --------------------------------------------------------
// The reference throws ::CudaError(file, line, func, code) through
// SCHWARZ_ASSERT_NO_CUDA_ERRORS (include/exception_helpers.hpp:82-91); the C
// ABI turns exceptions into status codes + schwz_b200_last_error().
struct CudaFailure : std::runtime_error {
    int code;
    CudaFailure(const char *file, int line, const char *what, int code_)
        : std::runtime_error(std::string(file) + ":" + std::to_string(line) +
                             ": " + what + ": " +
                             cudaGetErrorName((cudaError_t)code_) + ": " +
                             cudaGetErrorString((cudaError_t)code_)),
          code(code_)
    {}
};

#define SCHWZ_CUDA(expr)                                                      \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess)                                                \
            throw ::schwz_b200::CudaFailure(__FILE__, __LINE__, #expr, _e);   \
    } while (0)

#define SCHWZ_REQUIRE(cond, msg)                                             \
    do {                                                                     \
        if (!(cond))                                                         \
            throw std::runtime_error(std::string(__FILE__) + ":" +           \
                                     std::to_string(__LINE__) + ": " + msg); \
    } while (0)

extern std::atomic<int64_t> g_launches;   // kernels launched by this library
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

constexpr int kBlock = 256;       // threads per CTA for every kernel here
constexpr int kSpmvUnroll = 8;    // nnz per thread per CTA tile
constexpr int kSpmvTile = kBlock * kSpmvUnroll;   // nnz staged per CTA
constexpr int kVecCtasPerSM = 8;  // resident CTAs/SM targeted by vector kernels
// (grids are sized from the SM count the device reports: Ctx::num_sms, 148 on B200)
constexpr int kMaxPartials = 65536;

#ifdef __CUDACC__
// Deterministic CTA-wide sum (fixed shuffle tree, fixed warp order).
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v, double *warp_buf)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();   // warp_buf may still be read from a previous call
    if (lane == 0) warp_buf[w] = v;
    __syncthreads();
    double s = 0.0;
    if (w == 0) {
        s = lane < (kBlock / 32) ? warp_buf[lane] : 0.0;
        s = warp_sum(s);
    }
    return s;   // valid in warp 0
}

// "last CTA finishes the reduction" ticket.  Returns true in every thread of
// the CTA that arrived last; that CTA then sums partials[0..gridDim.x) in a
// fixed order, so the result does not depend on CTA scheduling.
__device__ __forceinline__ bool last_cta(unsigned int *ticket)
{
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0u;   // re-arm for the next launch
    }
    __syncthreads();
    return s_last != 0;
}

__device__ __forceinline__ double reduce_partials(const double *partials, int n,
                                                  double *warp_buf)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += kBlock) s += __ldcg(partials + i);
    return block_sum(s, warp_buf);
}

// system-scope release/acquire on peer-mapped words (epoch flags, convergence
// flags) — the replacement for MPI_Win_flush + MPI_Put of a flag.
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_i32(int *p, int v)
{
    asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_sys_i32(const int *p)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif  // __CUDACC__

}  // namespace schwz_b200
