// common.cuh -- shared helpers for the nbk CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nbk.h"

namespace nbk {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define NBK_CHECK(call)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            throw ::nbk::Error(NBK_ERR_CUDA, _b);                                                    \
        }                                                                                            \
    } while (0)

#define NBK_REQUIRE(cond, code, msg)                                  \
    do {                                                              \
        if (!(cond)) throw ::nbk::Error((code), std::string(msg));    \
    } while (0)

// Stream the calling API entry point works on: DevBuf allocations are stream-ordered, from the library's OWN memory pool
// (one per device, created on first use, release threshold unlimited: temporaries are recycled, not returned to the driver --
// cudaMalloc/cudaFree of multi-GB scratch costs more than the kernels using it).  The device's default pool, and with it
// every other user of cudaMallocAsync in the process, is left alone; nbk_release_cached_memory trims the private pool.
inline cudaStream_t& cur_stream() { static thread_local cudaStream_t s = nullptr; return s; }
inline cudaMemPool_t nbk_pool(int device = -1) {
    static cudaMemPool_t pools[64] = {nullptr};
    if (device < 0) cudaGetDevice(&device);
    if (device < 0 || device >= 64) return nullptr;
    if (!pools[device]) {
        static std::mutex m;
        std::lock_guard<std::mutex> g(m);
        if (!pools[device]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = device;
            cudaMemPool_t pool = nullptr;
            if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            pools[device] = pool;
        }
    }
    return pools[device];
}
// stream-ordered allocation from the library's pool (falls back to the default pool if a private one cannot be created)
inline cudaError_t nbk_malloc_async(void** ptr, size_t bytes, cudaStream_t st) {
    cudaMemPool_t pool = nbk_pool();
    return pool ? cudaMallocFromPoolAsync(ptr, bytes, pool, st) : cudaMallocAsync(ptr, bytes, st);
}

// RAII device buffer
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) {
            cudaError_t e = nbk_malloc_async((void**)&p, count * sizeof(T), cur_stream());
            if (e != cudaSuccess) {
                p = nullptr; n = 0;
                cudaGetLastError();
                throw Error(NBK_ERR_NOMEM, std::string("device allocation failed: ") + cudaGetErrorString(e));
            }
        }
    }
    void release() {
        if (p) cudaFreeAsync(p, cur_stream());
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
    operator T*() const { return p; }
};

// NBK_TRACE=1 : print host-side phase timings (each point synchronises the stream; debugging only)
struct Tracer {
    cudaStream_t st; bool on; double last;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit Tracer(cudaStream_t s) : st(s), on(getenv("NBK_TRACE") != nullptr), last(0) { if (on) { cudaStreamSynchronize(st); last = now(); } }
    void point(const char* label) {
        if (!on) return;
        cudaStreamSynchronize(st);
        double t = now();
        fprintf(stderr, "[nbk] %-32s %10.3f ms\n", label, t - last);
        last = t;
    }
};

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Size of the LEFT child of a node of `size` particles (the right child takes the rest).
//   reference shape (aligned = 0): ceil(size / 2)  (KDTree.cxx:1012: split index = start + (size - 1) / 2);
//   warp-aligned shape (NBK_WARP_ALIGNED trees): a node above 32 particles splits at a multiple of 32 -- the balanced split of its
//   32-particle units -- so that every run of 32 consecutive tree positions [32 g, 32 g + 32) is exactly one node: the warp's
//   query group of the density / FOF kernels coincides with a node for ANY particle count, not only for powers of two.
//   Still a median split (the cut plane moves by less than 32 particles' worth); left >= right, depth unchanged.
__host__ __device__ __forceinline__ int64_t split_left(int64_t size, int aligned) {
    if (aligned && size > 32) { const int64_t u = (size + 31) >> 5; return ((u + 1) >> 1) << 5; }
    return (size + 1) >> 1;
}

// 4-wide coordinate records: storage type S is float (16 B) or double (32 B)
template <class S> struct Vec4;
template <> struct __align__(16) Vec4<float> { float x, y, z, w; };
template <> struct __align__(32) Vec4<double> { double x, y, z, w; };

// node record, heap order (root 0, children 2i+1 / 2i+2): tight fp32 bounds rounded outward + range
struct __align__(16) NodeLo { float x, y, z; int start; };  // start < 0 : node absent
struct __align__(16) NodeHi { float x, y, z; int end; };

#ifdef __CUDACC__
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// order preserving bit keys
__device__ __forceinline__ uint32_t sort_key(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ uint64_t sort_key(double f) {
    uint64_t u = (uint64_t)__double_as_longlong(f);
    return u ^ ((u >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}
// outward rounding of storage coordinates to fp32 bounds
__device__ __forceinline__ float round_down(float v) { return v; }
__device__ __forceinline__ float round_up(float v) { return v; }
__device__ __forceinline__ float round_down(double v) { return __double2float_rd(v); }
__device__ __forceinline__ float round_up(double v) { return __double2float_ru(v); }
#endif

}  // namespace nbk
