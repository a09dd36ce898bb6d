// api.cu -- the C ABI of include/nbk.h: argument checking, host<->device staging, call sequencing.
// No computation happens on the host and nothing here falls back to a CPU path: every entry point either
// launches the CUDA kernels of build.cu / knn.cu / fof.cu / ball.cu or returns an error.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <thread>

#include "tree.h"

using namespace nbk;

static thread_local std::string g_err;
namespace nbk { void set_last_error(const char* msg) { g_err = msg ? msg : ""; } }

#define NBK_API_BEGIN try {
#define NBK_API_END                                                          \
    }                                                                        \
    catch (const nbk::Error& e) { g_err = e.what(); return e.code; }         \
    catch (const std::bad_alloc&) { g_err = "host allocation failed"; return NBK_ERR_NOMEM; } \
    catch (const std::exception& e) { g_err = e.what(); return NBK_ERR_ARG; } \
    return NBK_OK;

namespace {

struct DeviceGuard {
    int prev = -1;
    cudaStream_t prev_stream;
    explicit DeviceGuard(int dev, cudaStream_t st = nullptr) {
        cudaGetDevice(&prev);
        if (dev != prev) NBK_CHECK(cudaSetDevice(dev));
        prev_stream = cur_stream();
        cur_stream() = st;
    }
    ~DeviceGuard() { cur_stream() = prev_stream; if (prev >= 0) cudaSetDevice(prev); }
};

// device + stream of a tree for the duration of an entry point, and the tree's lock (see nbk_tree::mtx)
struct TreeGuard {
    std::unique_lock<std::mutex> lock;
    DeviceGuard dev;
    explicit TreeGuard(const nbk_tree* t) : lock(t->mtx), dev(t->device, t->stream) {}
};

// ---- staging kernels -----------------------------------------------------------------------------
// raw strided reals (device visible) -> packed doubles [n][3]
template <class R>
__global__ void gather3_kernel(const unsigned char* base, int64_t stride, int64_t n, double* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const R* p = reinterpret_cast<const R*>(base + i * stride);
    out[3 * i] = (double)p[0]; out[3 * i + 1] = (double)p[1]; out[3 * i + 2] = (double)p[2];
}
template <class R>
__global__ void gather1_kernel(const unsigned char* base, int64_t stride, int64_t n, double* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (double)*reinterpret_cast<const R*>(base + i * stride);
}
__global__ void count_inexact_kernel(const double* a, int64_t n3, unsigned long long* cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n3) { double v = a[i]; bad = ((double)(float)v != v); }
    unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(cnt, (unsigned long long)__popc(m));
}
template <class S>
__global__ void pack4_kernel(const double* a, int64_t n, Vec4<S>* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Vec4<S> v;
    v.x = (S)a[3 * i]; v.y = (S)a[3 * i + 1]; v.z = (S)a[3 * i + 2]; v.w = (S)0;
    out[i] = v;
}
__global__ void scatter_f64_kernel(int64_t n, const int32_t* order, const double* src, double* dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[order[i]] = src[i];
}
__global__ void scatter_i32_kernel(int64_t n, const int32_t* order, const int32_t* src, int32_t* dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[order[i]] = src[i];
}
__global__ void gather_u8_kernel(int64_t n, const int32_t* order, const uint8_t* src, uint8_t* dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[order[i]];
}
// per-particle rows of w doubles: caller order (by ID) <-> tree order
__global__ void gather_rows_f64_kernel(int64_t n, int w, const int32_t* order, const double* src_by_id, double* dst_tree) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    int64_t p = i / w; int c = (int)(i - p * w);
    dst_tree[i] = src_by_id[(int64_t)order[p] * w + c];
}
__global__ void scatter_rows_f64_kernel(int64_t n, int w, const int32_t* order, const double* src_tree, double* dst_by_id) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    int64_t p = i / w; int c = (int)(i - p * w);
    dst_by_id[(int64_t)order[p] * w + c] = src_tree[i];
}
__global__ void check_indices_kernel(int64_t m, const int32_t* q, int32_t n, int* bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m && (q[i] < 0 || q[i] >= n)) *bad = 1;
}
__global__ void gather_i32_kernel(int64_t n, const int32_t* order, const int32_t* src, int32_t* dst) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[order[i]];
}

// Bring n strided records of `comps` reals (host or device memory) into a packed device double[n][comps].
// Device input: converted in place by a gather kernel.  Host input: contiguous arrays are copied as they are
// (cudaMemcpyAsync straight from the caller's buffer -- pinned buffers run at PCIe speed); strided AoS records
// (the reference's 88-byte Particle) are packed by a few host threads into pinned chunks, double buffered
// against the copies.  Widening to fp64 always happens on the device.
void stage3(const void* src, int64_t stride, int real_bytes, bool on_device, int64_t n, int comps, double* d_out, cudaStream_t st) {
    const int tb = 256;
    auto convert = [&](const void* dsrc, int64_t dstride) {
        if (comps == 3) {
            if (real_bytes == 8) gather3_kernel<double><<<div_up(n, tb), tb, 0, st>>>((const unsigned char*)dsrc, dstride, n, d_out);
            else gather3_kernel<float><<<div_up(n, tb), tb, 0, st>>>((const unsigned char*)dsrc, dstride, n, d_out);
        } else {
            if (real_bytes == 8) gather1_kernel<double><<<div_up(n, tb), tb, 0, st>>>((const unsigned char*)dsrc, dstride, n, d_out);
            else gather1_kernel<float><<<div_up(n, tb), tb, 0, st>>>((const unsigned char*)dsrc, dstride, n, d_out);
        }
        NBK_CHECK(cudaGetLastError());
    };
    if (on_device) { convert(src, stride); return; }
    const int64_t rec = (int64_t)comps * real_bytes;
    if (real_bytes == 8 && stride == rec) {   // packed doubles: no conversion needed
        NBK_CHECK(cudaMemcpyAsync(d_out, src, (size_t)n * rec, cudaMemcpyHostToDevice, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        return;
    }
    DevBuf<unsigned char> raw((size_t)n * rec);
    if (stride == rec) {
        NBK_CHECK(cudaMemcpyAsync(raw.p, src, (size_t)n * rec, cudaMemcpyHostToDevice, st));
    } else {
        const int64_t chunk = 1 << 22;
        unsigned char* pin[2] = {nullptr, nullptr};
        cudaEvent_t done[2];
        for (int b = 0; b < 2; b++) {
            NBK_CHECK(cudaMallocHost((void**)&pin[b], (size_t)std::min(chunk, n) * rec));
            NBK_CHECK(cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming));
        }
        unsigned hw = std::thread::hardware_concurrency();
        int nth = (int)std::max(1u, std::min(hw ? hw : 4u, 16u));
        int b = 0;
        for (int64_t c0 = 0; c0 < n; c0 += chunk, b ^= 1) {
            int64_t cn = std::min(chunk, n - c0);
            NBK_CHECK(cudaEventSynchronize(done[b]));
            auto work = [&](int tid) {
                int64_t lo = cn * tid / nth, hi = cn * (tid + 1) / nth;
                const unsigned char* bp = (const unsigned char*)src + (c0 + lo) * stride;
                unsigned char* o = pin[b] + lo * rec;
                for (int64_t i = lo; i < hi; i++, bp += stride, o += rec) memcpy(o, bp, (size_t)rec);
            };
            if (cn < 65536 || nth == 1) { for (int tdx = 0; tdx < nth; tdx++) work(tdx); }
            else {
                std::vector<std::thread> th;
                for (int tdx = 0; tdx < nth; tdx++) th.emplace_back(work, tdx);
                for (auto& x : th) x.join();
            }
            NBK_CHECK(cudaMemcpyAsync(raw.p + c0 * rec, pin[b], (size_t)cn * rec, cudaMemcpyHostToDevice, st));
            NBK_CHECK(cudaEventRecord(done[b], st));
        }
        NBK_CHECK(cudaStreamSynchronize(st));
        for (int q = 0; q < 2; q++) { cudaFreeHost(pin[q]); cudaEventDestroy(done[q]); }
    }
    convert(raw.p, rec);
    NBK_CHECK(cudaStreamSynchronize(st));
}

// KernelConstruction, reference KDTree.cxx:1144-1183 + SmoothingKernels.h:32-56
double kern_w(int type, double r, double h) {
    if (type == NBK_KSPH) {
        double rh = r / h;
        if (rh <= 1.0) return (1.0 - 0.75 * (2.0 - rh) * (rh * rh));
        else if (rh > 1.0 && rh <= 2.0) return 0.25 * (2.0 - rh) * (2.0 - rh) * (2.0 - rh);
        return 0;
    }
    if (type == NBK_KGAUSS) { double rh = r / h * 2.0; return exp(-0.5 * pow(rh, 2.0)); }
    if (type == NBK_KEPAN) { double rh = r / h * 0.5; return rh <= 1.0 ? (1.0 - rh * rh) : 0.; }
    double rh = r / h * 0.5;
    return (rh <= 1.0);
}
void build_kernel_table(nbk_tree& t) {
    const int ND = t.nd;
    double kn = 1.0 / ND * pow(M_1_PI, ND / 2.) * tgamma(ND / 2. + 1.);
    int type = t.kerntype;
    if (type == NBK_KSPH) kn *= ND * (ND + 1.) * (ND + 2.) * (ND + 3.) / (6 * (pow(2., ND + 1) - 1.));
    else if (type == NBK_KGAUSS) kn *= pow(0.5 * M_1_PI, ND / 2.);
    else if (type == NBK_KEPAN) kn *= ND * (ND + 2.) * pow(0.5, ND + 1.);
    else if (type != NBK_KTH) { type = NBK_KSPH; kn *= ND * (ND + 1.) * (ND + 2.) * (ND + 3.) / (6 * (pow(2., ND + 1) - 1.)); }
    t.kernnorm = kn;
    t.h_kernel.resize(t.kernres);
    double delta = 2.0 / (double)(t.kernres - 1);
    for (int i = 0; i < t.kernres; i++) t.h_kernel[i] = kn * kern_w(type, i * delta, 1.0);
    NBK_CHECK(nbk_malloc_async((void**)&t.d_kernel, sizeof(double) * t.kernres, t.stream));
    NBK_CHECK(cudaStreamSynchronize(t.stream));
    NBK_CHECK(cudaMemcpy(t.d_kernel, t.h_kernel.data(), sizeof(double) * t.kernres, cudaMemcpyHostToDevice));
}

// root representatives (nbk_fof_roots): tree-order root indices -> IDs, delivered by ID
__global__ void roots_to_ids_kernel(int64_t n, const int32_t* __restrict__ order, const int32_t* __restrict__ root_tree, int32_t* __restrict__ out_by_id) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t r = root_tree[i];
    out_by_id[order[i]] = r < 0 ? -1 : order[r];
}

struct CallTimer {
    nbk_tree& t;
    explicit CallTimer(nbk_tree& tt) : t(tt) { cudaEventRecord(t.ev0, t.stream); }
    void stop() {
        cudaEventRecord(t.ev1, t.stream);
        cudaEventSynchronize(t.ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, t.ev0, t.ev1);
        t.last_call_ms = ms;
    }
};

}  // namespace

extern "C" {

const char* nbk_last_error(void) { return g_err.c_str(); }

int nbk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int nbk_create(const nbk_particles* p, int64_t n, int bucket, int treetype, int kerntype, int kernres, int split,
               const double* period, int flags, int device, nbk_tree** out) {
    NBK_API_BEGIN
    NBK_REQUIRE(out != nullptr && p != nullptr && p->pos != nullptr, NBK_ERR_ARG, "nbk_create: null argument");
    *out = nullptr;
    NBK_REQUIRE(n >= 1 && n < ((int64_t)1 << 31) - 64, NBK_ERR_ARG, "nbk_create: particle count must be in [1, 2^31)");
    NBK_REQUIRE(bucket >= 1 && bucket <= 1024, NBK_ERR_ARG, "nbk_create: bucket size must be in [1,1024]");
    NBK_REQUIRE(p->real_bytes == 4 || p->real_bytes == 8, NBK_ERR_ARG, "nbk_create: real_bytes must be 4 or 8");
    // reference TreeTypeCheck (KDTree.cxx:1095-1105) prints and leaves root=NULL; here it is an error code
    NBK_REQUIRE(treetype >= NBK_TPHYS && treetype <= NBK_TMETRIC, NBK_ERR_ARG, "Error in type of tree specified");
    NBK_REQUIRE(treetype == NBK_TPHYS || treetype == NBK_TPHS || treetype == NBK_TVEL, NBK_ERR_UNSUPPORTED,
                "tree types TPROJ / TMETRIC have no device implementation");
    NBK_REQUIRE(split == 0, NBK_ERR_UNSUPPORTED, "only KDTREE_SPLIT_SPREAD is implemented on the device");
    NBK_REQUIRE(!(treetype != NBK_TPHYS && p->vel == nullptr), NBK_ERR_ARG, "TVEL / TPHS trees need velocities");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw Error(NBK_ERR_CUDA, "no CUDA device: nbk has no CPU fallback");
    }
    if (device < 0) NBK_CHECK(cudaGetDevice(&device));
    DeviceGuard guard(device);

    std::unique_ptr<nbk_tree> t(new nbk_tree);
    t->device = device;
    t->n = n; t->bucket = bucket; t->treetype = treetype; t->kerntype = kerntype;
    t->aligned = (flags & NBK_WARP_ALIGNED) ? 1 : 0;
    if (kernres < 100) kernres = 100;      // KDTree.cxx:1145-1148
    t->kernres = kernres;
    t->nd = (treetype == NBK_TPHS) ? 6 : 3;
    t->periodic = period != nullptr;
    if (period) for (int d = 0; d < 3; d++) t->period[d] = period[d];
    NBK_CHECK(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    cur_stream() = t->stream;
    NBK_CHECK(cudaEventCreate(&t->ev0)); NBK_CHECK(cudaEventCreate(&t->ev1));
    NBK_CHECK(cudaEventCreate(&t->ev2)); NBK_CHECK(cudaEventCreate(&t->ev3));
    cudaStream_t st = t->stream;
    const int tb = 256;

    // ---- stage ------------------------------------------------------------------------------------
    // The tree is built on one phase-space half only (positions; velocities for TVEL).  With host input whose storage
    // width is known up front, the other half and the masses are staged by a side thread on its own stream while the main
    // stream sorts and partitions; the build joins it right before the final particle gather.
    NBK_CHECK(cudaEventRecord(t->ev0, st));
    const bool tvel = treetype == NBK_TVEL;
    const void* prim_src = tvel ? p->vel : p->pos;
    const int64_t prim_stride = tvel ? p->vel_stride : p->pos_stride;
    const void* sec_src = tvel ? p->pos : p->vel;
    const int64_t sec_stride = tvel ? p->pos_stride : p->vel_stride;
    const bool host_in = p->on_device == 0;
    const bool width_known = p->real_bytes == 4 || (flags & (NBK_STORE_F64 | NBK_STORE_F32));
    const bool overlap = host_in && width_known && (sec_src || p->mass);
    DevBuf<double> rprim((size_t)3 * n), rsec, rmass;
    stage3(prim_src, prim_stride, p->real_bytes, !host_in, n, 3, rprim.p, st);
    if (!overlap) {
        if (sec_src) { rsec.alloc((size_t)3 * n); stage3(sec_src, sec_stride, p->real_bytes, !host_in, n, 3, rsec.p, st); }
        if (p->mass) { rmass.alloc((size_t)n); stage3(p->mass, p->mass_stride, p->real_bytes, !host_in, n, 1, rmass.p, st); }
    }
    NBK_CHECK(cudaEventRecord(t->ev1, st));

    // ---- storage precision: fp32 when it is exact (or forced), fp64 otherwise ---------------------------
    int store = 4;
    if (p->real_bytes == 8 && !width_known) {
        DevBuf<unsigned long long> cnt(1);
        NBK_CHECK(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), st));
        count_inexact_kernel<<<div_up(3 * n, tb), tb, 0, st>>>(rprim.p, 3 * n, cnt.p);
        if (sec_src) count_inexact_kernel<<<div_up(3 * n, tb), tb, 0, st>>>(rsec.p, 3 * n, cnt.p);
        unsigned long long c = 0;
        NBK_CHECK(cudaMemcpyAsync(&c, cnt.p, sizeof(c), cudaMemcpyDeviceToHost, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        t->inexact = (int64_t)c;
        if (c > 0) store = 8;
    }
    if (flags & NBK_STORE_F64) store = 8;
    if (flags & NBK_STORE_F32) store = 4;
    if (store == 8) t->inexact = 0;
    t->store_bytes = store;

    NBK_CHECK(cudaEventRecord(t->ev2, st));
    auto run = [&](auto tag) {
        typedef decltype(tag) S;
        DevBuf<Vec4<S>> a(n), b(sec_src ? n : 0);
        if (p->mass && overlap) rmass.alloc((size_t)n);
        pack4_kernel<S><<<div_up(n, tb), tb, 0, st>>>(rprim.p, n, a.p);
        rprim.release();                                   // stream ordered: freed once the pack has run
        std::thread side;
        std::exception_ptr side_err;
        cudaStream_t s2 = nullptr;
        if (overlap) {
            NBK_CHECK(cudaStreamSynchronize(st));          // the buffers the side stream fills exist from here on
            NBK_CHECK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            Vec4<S>* bp = b.p; double* mp = rmass.p;
            const int dev = device;
            side = std::thread([&, bp, mp, dev, s2]() {
                try {
                    NBK_CHECK(cudaSetDevice(dev));
                    cur_stream() = s2;
                    if (sec_src) {
                        DevBuf<double> tmp((size_t)3 * n);
                        stage3(sec_src, sec_stride, p->real_bytes, false, n, 3, tmp.p, s2);
                        pack4_kernel<S><<<div_up(n, tb), tb, 0, s2>>>(tmp.p, n, bp);
                        NBK_CHECK(cudaStreamSynchronize(s2));
                    }
                    if (p->mass) stage3(p->mass, p->mass_stride, p->real_bytes, false, n, 1, mp, s2);
                    NBK_CHECK(cudaStreamSynchronize(s2));
                } catch (...) { side_err = std::current_exception(); }
            });
            t->before_gather = [&side]() { if (side.joinable()) side.join(); };
        } else if (sec_src) {
            pack4_kernel<S><<<div_up(n, tb), tb, 0, st>>>(rsec.p, n, b.p);
        }
        struct Joiner { std::thread& th; cudaStream_t& s; ~Joiner() { if (th.joinable()) th.join(); if (s) cudaStreamDestroy(s); } } joiner{side, s2};
        if (!overlap) { NBK_CHECK(cudaStreamSynchronize(st)); rsec.release(); }
        build_tree<S>(*t, a.p, sec_src ? b.p : nullptr, p->mass ? rmass.p : nullptr);
        t->before_gather = nullptr;
        if (side.joinable()) side.join();
        if (side_err) std::rethrow_exception(side_err);
    };
    if (store == 4) run(float());
    else run(double());
    NBK_CHECK(cudaEventRecord(t->ev3, st));
    NBK_CHECK(cudaEventSynchronize(t->ev3));
    float ms = 0;
    NBK_CHECK(cudaEventElapsedTime(&ms, t->ev0, t->ev1)); t->h2d_ms = ms;
    NBK_CHECK(cudaEventElapsedTime(&ms, t->ev2, t->ev3)); t->build_ms = ms;
    build_kernel_table(*t);
    *out = t.release();
    NBK_API_END
}

int nbk_destroy(nbk_tree* t) {
    NBK_API_BEGIN
    if (!t) return NBK_OK;
    { std::lock_guard<std::mutex> wait_for_running_calls(t->mtx); }
    delete t;                                  // ~nbk_tree frees the buffers, events and the stream
    NBK_API_END
}

namespace {
__global__ void offset_nodes_kernel(int64_t nslots, NodeLo* nlo, NodeHi* nhi, int off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nslots && nlo[i].start >= 0) { nlo[i].start += off; nhi[i].end += off; }
}
__global__ void offset_i32_kernel(int64_t n, int32_t* a, int off) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += off;
}
}  // namespace

int nbk_attach_halo(nbk_tree* t, nbk_tree* halo) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && halo && t != halo, NBK_ERR_ARG, "nbk_attach_halo: null argument");
    NBK_REQUIRE(t->n_main == 0 && halo->n_main == 0, NBK_ERR_ARG, "nbk_attach_halo: a halo is already attached");
    NBK_REQUIRE(t->device == halo->device && t->store_bytes == halo->store_bytes && t->treetype == NBK_TPHYS && halo->treetype == NBK_TPHYS &&
                    t->bucket == halo->bucket && (t->sec == nullptr) == (halo->sec == nullptr),
                NBK_ERR_ARG, "nbk_attach_halo: the two trees must be TPHYS trees on one device with the same storage width, bucket size and columns");
    NBK_REQUIRE(t->n + halo->n < ((int64_t)1 << 31) - 64, NBK_ERR_ARG, "nbk_attach_halo: too many particles");
    TreeGuard guard(t);
    cudaStream_t st = t->stream;
    NBK_CHECK(cudaStreamSynchronize(halo->stream));
    const int64_t n1 = t->n, n2 = halo->n, n = n1 + n2;
    const size_t vb = (size_t)t->store_bytes * 4;
    DevBuf<unsigned char> prim((size_t)n * vb), sec(t->sec ? (size_t)n * vb : 0);
    DevBuf<double> mass(n);
    DevBuf<int32_t> order(n);
    auto cat = [&](void* dst, const void* a, const void* b, size_t elem) {
        NBK_CHECK(cudaMemcpyAsync(dst, a, (size_t)n1 * elem, cudaMemcpyDeviceToDevice, st));
        NBK_CHECK(cudaMemcpyAsync((unsigned char*)dst + (size_t)n1 * elem, b, (size_t)n2 * elem, cudaMemcpyDeviceToDevice, st));
    };
    cat(prim.p, t->prim, halo->prim, vb);
    if (t->sec) cat(sec.p, t->sec, halo->sec, vb);
    cat(mass.p, t->mass, halo->mass, sizeof(double));
    cat(order.p, t->order, halo->order, sizeof(int32_t));
    offset_i32_kernel<<<div_up(n2, 256), 256, 0, st>>>(n2, order.p + n1, (int)n1);            // halo IDs follow the main IDs
    offset_nodes_kernel<<<div_up(halo->nslots, 256), 256, 0, st>>>(halo->nslots, halo->nlo, halo->nhi, (int)n1);
    NBK_CHECK(cudaGetLastError());
    void* old[] = {t->prim, t->sec, t->mass, t->order, halo->prim, halo->sec, halo->mass, halo->order, halo->cutdim, halo->d_kernel};
    for (void* b : old) if (b) cudaFreeAsync(b, st);
    NBK_CHECK(cudaStreamSynchronize(st));
    t->prim = prim.p; prim.p = nullptr;
    t->sec = sec.p; sec.p = nullptr;
    t->mass = mass.p; mass.p = nullptr;
    t->order = order.p; order.p = nullptr;
    t->nlo2 = halo->nlo; t->nhi2 = halo->nhi;
    t->aligned2 = halo->aligned;
    t->knn_fp32_ok = -1;
    t->n_main = n1; t->n = n;
    t->device_bytes += halo->device_bytes;
    // the halo handle is consumed: its particle arrays are freed above, its node arrays now belong to t
    halo->prim = halo->sec = nullptr; halo->mass = nullptr; halo->order = nullptr; halo->cutdim = nullptr; halo->d_kernel = nullptr;
    halo->nlo = nullptr; halo->nhi = nullptr;
    delete halo;
    NBK_API_END
}

int nbk_get_info(const nbk_tree* t, nbk_info* info) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && info, NBK_ERR_ARG, "nbk_get_info: null argument");
    memset(info, 0, sizeof(*info));
    info->n = t->n; info->bucket = t->bucket; info->treetype = t->treetype; info->kerntype = t->kerntype;
    info->kernres = t->kernres; info->nd = t->nd;
    info->num_nodes = (int32_t)t->num_nodes; info->num_leaves = (int32_t)t->num_leaves; info->depth = t->depth;
    info->store_bytes = t->store_bytes; info->periodic = t->periodic ? 1 : 0; info->inexact_coords = t->inexact;
    info->kernnorm = t->kernnorm;
    for (int d = 0; d < 3; d++) info->period[d] = t->period[d];
    info->build_ms = t->build_ms; info->h2d_ms = t->h2d_ms; info->last_kernel_ms = t->last_kernel_ms;
    info->last_call_ms = t->last_call_ms; info->last_launches = t->last_launches; info->device_bytes = t->device_bytes; info->last_flagged = t->last_flagged;
    info->warp_aligned = t->aligned;
    NBK_API_END
}

int nbk_get_order(const nbk_tree* t, int32_t* ids, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && ids, NBK_ERR_ARG, "nbk_get_order: null argument");
    TreeGuard guard(t);
    NBK_CHECK(cudaMemcpyAsync(ids, t->order, sizeof(int32_t) * t->n, (flags & NBK_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->stream));
    NBK_CHECK(cudaStreamSynchronize(t->stream));
    NBK_API_END
}

int nbk_get_kernel_table(const nbk_tree* t, double* table) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && table, NBK_ERR_ARG, "nbk_get_kernel_table: null argument");
    memcpy(table, t->h_kernel.data(), sizeof(double) * t->kernres);
    NBK_API_END
}

int nbk_get_nodes(const nbk_tree* t, int64_t* num_slots, int32_t* start, int32_t* end, int32_t* cutdim, float* bounds) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && num_slots, NBK_ERR_ARG, "nbk_get_nodes: null argument");
    *num_slots = t->nslots;
    if (!start && !end && !cutdim && !bounds) return NBK_OK;
    TreeGuard guard(t);
    std::vector<NodeLo> lo(t->nslots);
    std::vector<NodeHi> hi(t->nslots);
    std::vector<int8_t> cd(t->nslots);
    NBK_CHECK(cudaMemcpy(lo.data(), t->nlo, sizeof(NodeLo) * t->nslots, cudaMemcpyDeviceToHost));
    NBK_CHECK(cudaMemcpy(hi.data(), t->nhi, sizeof(NodeHi) * t->nslots, cudaMemcpyDeviceToHost));
    NBK_CHECK(cudaMemcpy(cd.data(), t->cutdim, t->nslots, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < t->nslots; i++) {
        if (start) start[i] = lo[i].start;
        if (end) end[i] = hi[i].end;
        if (cutdim) cutdim[i] = (lo[i].start >= 0 && hi[i].end - lo[i].start > t->bucket) ? cd[i] : -1;   // leaf iff size <= bucket
        if (bounds) {
            bounds[6 * i + 0] = lo[i].x; bounds[6 * i + 1] = hi[i].x; bounds[6 * i + 2] = lo[i].y;
            bounds[6 * i + 3] = hi[i].y; bounds[6 * i + 4] = lo[i].z; bounds[6 * i + 5] = hi[i].z;
        }
    }
    NBK_API_END
}

static void require_no_halo(const nbk_tree* t, const char* what) {
    if (t->n_main) throw Error(NBK_ERR_UNSUPPORTED, std::string(what) + ": a tree with an attached halo serves the Calc* family only");
}
static void require_knn_tree(const nbk_tree* t) {
    // Q4: TPHS trees with the constructor-default Aniso=0 take the metric path in the reference; no device version.
    NBK_REQUIRE(t->treetype == NBK_TPHYS || t->treetype == NBK_TVEL, NBK_ERR_UNSUPPORTED,
                "kNN / density on a TPHS tree (6D / metric search) has no device implementation");
}

int nbk_knn_particles(nbk_tree* t, int k, int64_t q0, int64_t q1, int32_t* nn, double* d2, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_knn_particles: null tree");
    require_knn_tree(t);
    require_no_halo(t, "nbk_knn_particles");
    NBK_REQUIRE(q0 >= 0 && q1 <= t->n && q0 <= q1, NBK_ERR_ARG, "nbk_knn_particles: bad query range");
    NBK_REQUIRE(k >= 1, NBK_ERR_ARG, "nbk_knn_particles: k must be >= 1");
    TreeGuard guard(t);
    const int64_t rows = q1 - q0;
    const bool dev = flags & NBK_DEVICE_PTRS;
    DevBuf<int32_t> dnn;
    DevBuf<double> dd2;
    KnnArgs a;
    a.k = k; a.mode = 0; a.q0 = q0; a.q1 = q1;
    a.periodic = t->periodic && t->treetype != NBK_TVEL;   // the reference never reflects velocity searches (KDSplitNode.cxx:1082-1085)
    a.strict = flags & NBK_STRICT_PERIODIC; a.tree_form = flags & NBK_KNN_TREE_FORM; a.out_ids = flags & NBK_OUT_IDS;
    if (dev) { a.nn = nn; a.d2 = d2; }
    else {
        if (nn) { dnn.alloc((size_t)rows * k); a.nn = dnn.p; }
        if (d2) { dd2.alloc((size_t)rows * k); a.d2 = dd2.p; }
    }
    CallTimer tm(*t);
    launch_knn(*t, a);
    tm.stop();
    t->last_kernel_ms = t->last_call_ms; t->last_launches = 1;
    if (!dev) {
        if (nn) NBK_CHECK(cudaMemcpyAsync(nn, dnn.p, dnn.bytes(), cudaMemcpyDeviceToHost, t->stream));
        if (d2) NBK_CHECK(cudaMemcpyAsync(d2, dd2.p, dd2.bytes(), cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
    NBK_API_END
}

int nbk_knn_points(nbk_tree* t, int k, int64_t m, const double* x, int32_t* nn, double* d2, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && x, NBK_ERR_ARG, "nbk_knn_points: null argument");
    require_knn_tree(t);
    require_no_halo(t, "nbk_knn_points");
    NBK_REQUIRE(k >= 1 && m >= 0, NBK_ERR_ARG, "nbk_knn_points: bad k or m");
    if (m == 0) return NBK_OK;
    TreeGuard guard(t);
    const bool dev = flags & NBK_DEVICE_PTRS;
    DevBuf<double> dx, dd2;
    DevBuf<int32_t> dnn;
    KnnArgs a;
    a.k = k; a.mode = 1; a.q0 = 0; a.q1 = m;
    a.periodic = t->periodic && t->treetype != NBK_TVEL;
    a.strict = flags & NBK_STRICT_PERIODIC; a.out_ids = flags & NBK_OUT_IDS;
    if (dev) { a.xq = x; a.nn = nn; a.d2 = d2; }
    else {
        dx.alloc((size_t)3 * m);
        NBK_CHECK(cudaMemcpyAsync(dx.p, x, dx.bytes(), cudaMemcpyHostToDevice, t->stream));
        a.xq = dx.p;
        if (nn) { dnn.alloc((size_t)m * k); a.nn = dnn.p; }
        if (d2) { dd2.alloc((size_t)m * k); a.d2 = dd2.p; }
    }
    CallTimer tm(*t);
    launch_knn(*t, a);
    tm.stop();
    t->last_kernel_ms = t->last_call_ms; t->last_launches = 1;
    if (!dev) {
        if (nn) NBK_CHECK(cudaMemcpyAsync(nn, dnn.p, dnn.bytes(), cudaMemcpyDeviceToHost, t->stream));
        if (d2) NBK_CHECK(cudaMemcpyAsync(d2, dd2.p, dd2.bytes(), cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
    NBK_API_END
}

// FindNearestCheck / FindNearestCriterion: the exact kernel with candidate filters
static void knn_filtered_call(nbk_tree* t, int k, int64_t q0, int64_t q1, int64_t m, const double* x, const double* v, int criterion,
                              const double* params, const int32_t* check, int32_t* nn, double* d2, int flags) {
    require_knn_tree(t);
    require_no_halo(t, "filtered kNN");
    NBK_REQUIRE(t->treetype == NBK_TPHYS, NBK_ERR_UNSUPPORTED, "FindNearestCheck / FindNearestCriterion need a physical tree");
    NBK_REQUIRE(k >= 1, NBK_ERR_ARG, "filtered kNN: k must be >= 1");
    NBK_REQUIRE(criterion >= 0 || check, NBK_ERR_ARG, "filtered kNN: neither a criterion nor check values given");
    TreeGuard guard(t);
    const bool dev = flags & NBK_DEVICE_PTRS;
    const int64_t n = t->n;
    KnnArgs a;
    a.k = k;
    a.periodic = t->periodic;
    a.strict = true;            // the reference walks all 8 images unconditionally (KDSplitNode.cxx:1255-1340): any exact schedule agrees
    a.tree_form = flags & NBK_KNN_TREE_FORM;
    a.out_ids = flags & NBK_OUT_IDS;
    if (criterion >= 0) {
        NBK_REQUIRE(params, NBK_ERR_ARG, "filtered kNN: criterion without params");
        if (!(criterion == NBK_FOF3D || criterion == NBK_FOF6D))
            throw Error(criterion == NBK_FOFVEL ? NBK_ERR_UNSUPPORTED : NBK_ERR_ARG, "FindNearestCriterion: only FOF3d and FOF6d have device implementations");
        a.crit_mode = criterion == NBK_FOF3D ? 2 : 4;
        a.cp0 = params[6]; a.cp1 = params[7];
    }
    DevBuf<int32_t> dchk, dchk_tree, dnn;
    DevBuf<double> dx, dv, dd2;
    if (check) {
        const int32_t* src = check;
        if (!dev) {
            dchk.alloc(n);
            NBK_CHECK(cudaMemcpyAsync(dchk.p, check, sizeof(int32_t) * n, cudaMemcpyHostToDevice, t->stream));
            src = dchk.p;
        }
        if (flags & NBK_TREE_ORDER) a.cand_excl = src;
        else {
            dchk_tree.alloc(n);
            gather_i32_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src, dchk_tree.p);
            a.cand_excl = dchk_tree.p;
        }
    }
    int64_t rows;
    if (x) {
        a.mode = 1; a.q0 = 0; a.q1 = m; rows = m;
        if (dev) { a.xq = x; a.vq = v; }
        else {
            dx.alloc((size_t)3 * m);
            NBK_CHECK(cudaMemcpyAsync(dx.p, x, dx.bytes(), cudaMemcpyHostToDevice, t->stream));
            a.xq = dx.p;
            if (v) { dv.alloc((size_t)3 * m); NBK_CHECK(cudaMemcpyAsync(dv.p, v, dv.bytes(), cudaMemcpyHostToDevice, t->stream)); a.vq = dv.p; }
        }
        NBK_REQUIRE(a.crit_mode != 4 || a.vq, NBK_ERR_ARG, "FindNearestCriterion(FOF6d) at a point needs the query velocity");
    } else {
        NBK_REQUIRE(q0 >= 0 && q1 <= n && q0 <= q1, NBK_ERR_ARG, "filtered kNN: bad query range");
        a.mode = 0; a.q0 = q0; a.q1 = q1; rows = q1 - q0;
    }
    if (rows == 0) return;
    if (dev) { a.nn = nn; a.d2 = d2; }
    else {
        if (nn) { dnn.alloc((size_t)rows * k); a.nn = dnn.p; }
        if (d2) { dd2.alloc((size_t)rows * k); a.d2 = dd2.p; }
    }
    NBK_REQUIRE(a.nn || a.d2, NBK_ERR_ARG, "filtered kNN: no output requested");
    CallTimer tm(*t);
    launch_knn(*t, a);
    tm.stop();
    t->last_kernel_ms = t->last_call_ms; t->last_launches = 1;
    if (!dev) {
        if (nn) NBK_CHECK(cudaMemcpyAsync(nn, dnn.p, dnn.bytes(), cudaMemcpyDeviceToHost, t->stream));
        if (d2) NBK_CHECK(cudaMemcpyAsync(d2, dd2.p, dd2.bytes(), cudaMemcpyDeviceToHost, t->stream));
    }
    NBK_CHECK(cudaStreamSynchronize(t->stream));
}

int nbk_knn_filtered_particles(nbk_tree* t, int k, int64_t q0, int64_t q1, int criterion, const double* params, const int32_t* check,
                               int32_t* nn, double* d2, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_knn_filtered_particles: null tree");
    knn_filtered_call(t, k, q0, q1, 0, nullptr, nullptr, criterion, params, check, nn, d2, flags);
    NBK_API_END
}
int nbk_knn_filtered_points(nbk_tree* t, int k, int64_t m, const double* x, const double* v, int criterion, const double* params,
                            const int32_t* check, int32_t* nn, double* d2, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && (x || m == 0) && m >= 0, NBK_ERR_ARG, "nbk_knn_filtered_points: null argument");
    if (m == 0) return NBK_OK;
    knn_filtered_call(t, k, 0, 0, m, x, v, criterion, params, check, nn, d2, flags);
    NBK_API_END
}

// FindNearestPhase: the exact kernel with 6D keys (knn.cu, PHASE instantiation)
static void knn_phase_call(nbk_tree* t, int k, int64_t q0, int64_t q1, int64_t m, const double* x, const double* v,
                           int32_t* nn, double* d2, int flags) {
    require_no_halo(t, "phase-space kNN");
    // the walk prunes on the position half of the distance: positions must be the tree coordinates (the reference's
    // FindNearestPhase walks whatever tree it is called on with GetPhase(cut_dim), KDSplitNode.cxx:69-95)
    NBK_REQUIRE(t->treetype == NBK_TPHYS || t->treetype == NBK_TPHS, NBK_ERR_UNSUPPORTED, "FindNearestPhase needs a TPHYS or TPHS tree");
    NBK_REQUIRE(t->vel4() != nullptr, NBK_ERR_ARG, "FindNearestPhase: the tree was built without velocities");
    NBK_REQUIRE(k >= 1, NBK_ERR_ARG, "phase-space kNN: k must be >= 1");
    TreeGuard guard(t);
    const bool dev = flags & NBK_DEVICE_PTRS;
    KnnArgs a;
    a.k = k; a.phase = true;
    a.periodic = t->periodic;
    a.strict = true;          // the reference walks all 8 position images unconditionally (KDSplitNode.cxx:1153-1187): any exact schedule agrees
    a.tree_form = true;       // periodic particle form: k + 1 slots, LoadNN(k) keeps the k farthest, i.e. drops the query itself (KDFindNearest.cxx:349,358-359)
    a.out_ids = flags & NBK_OUT_IDS;
    DevBuf<int32_t> dnn;
    DevBuf<double> dx, dv, dd2;
    int64_t rows;
    if (x) {
        NBK_REQUIRE(v, NBK_ERR_ARG, "FindNearestPhase at a point needs the query velocity");
        a.mode = 1; a.q0 = 0; a.q1 = m; rows = m;
        if (dev) { a.xq = x; a.vq = v; }
        else {
            dx.alloc((size_t)3 * m); dv.alloc((size_t)3 * m);
            NBK_CHECK(cudaMemcpyAsync(dx.p, x, dx.bytes(), cudaMemcpyHostToDevice, t->stream));
            NBK_CHECK(cudaMemcpyAsync(dv.p, v, dv.bytes(), cudaMemcpyHostToDevice, t->stream));
            a.xq = dx.p; a.vq = dv.p;
        }
    } else {
        NBK_REQUIRE(q0 >= 0 && q1 <= t->n && q0 <= q1, NBK_ERR_ARG, "phase-space kNN: bad query range");
        a.mode = 0; a.q0 = q0; a.q1 = q1; rows = q1 - q0;
    }
    if (rows == 0) return;
    if (dev) { a.nn = nn; a.d2 = d2; }
    else {
        if (nn) { dnn.alloc((size_t)rows * k); a.nn = dnn.p; }
        if (d2) { dd2.alloc((size_t)rows * k); a.d2 = dd2.p; }
    }
    NBK_REQUIRE(a.nn || a.d2, NBK_ERR_ARG, "phase-space kNN: no output requested");
    CallTimer tm(*t);
    launch_knn(*t, a);
    tm.stop();
    t->last_kernel_ms = t->last_call_ms; t->last_launches = 1;
    if (!dev) {
        if (nn) NBK_CHECK(cudaMemcpyAsync(nn, dnn.p, dnn.bytes(), cudaMemcpyDeviceToHost, t->stream));
        if (d2) NBK_CHECK(cudaMemcpyAsync(d2, dd2.p, dd2.bytes(), cudaMemcpyDeviceToHost, t->stream));
    }
    NBK_CHECK(cudaStreamSynchronize(t->stream));
}

int nbk_knn_phase_particles(nbk_tree* t, int k, int64_t q0, int64_t q1, int32_t* nn, double* d2, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_knn_phase_particles: null tree");
    knn_phase_call(t, k, q0, q1, 0, nullptr, nullptr, nn, d2, flags);
    NBK_API_END
}
int nbk_knn_phase_points(nbk_tree* t, int k, int64_t m, const double* x, const double* v, int32_t* nn, double* d2, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && ((x && v) || m == 0) && m >= 0, NBK_ERR_ARG, "nbk_knn_phase_points: null argument");
    if (m == 0) return NBK_OK;
    knn_phase_call(t, k, 0, 0, m, x, v, nn, d2, flags);
    NBK_API_END
}

// per-particle double output: tree-order device buffer -> caller (ID order unless NBK_TREE_ORDER)
static void deliver_f64(nbk_tree* t, const double* src_tree, double* dst, int flags) {
    const int64_t n = t->n;
    const bool dev = flags & NBK_DEVICE_PTRS;
    if (flags & NBK_TREE_ORDER) {
        NBK_CHECK(cudaMemcpyAsync(dst, src_tree, sizeof(double) * n, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->stream));
    } else if (dev) {
        scatter_f64_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src_tree, dst);
    } else {
        DevBuf<double> tmp(n);
        scatter_f64_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src_tree, tmp.p);
        NBK_CHECK(cudaMemcpyAsync(dst, tmp.p, sizeof(double) * n, cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
    NBK_CHECK(cudaStreamSynchronize(t->stream));
}
static void deliver_i32(nbk_tree* t, const int32_t* src_tree, int32_t* dst, int flags) {
    const int64_t n = t->n;
    const bool dev = flags & NBK_DEVICE_PTRS;
    if (flags & NBK_TREE_ORDER) {
        NBK_CHECK(cudaMemcpyAsync(dst, src_tree, sizeof(int32_t) * n, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->stream));
    } else if (dev) {
        scatter_i32_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src_tree, dst);
    } else {
        DevBuf<int32_t> tmp(n);
        scatter_i32_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src_tree, tmp.p);
        NBK_CHECK(cudaMemcpyAsync(dst, tmp.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
    NBK_CHECK(cudaStreamSynchronize(t->stream));
}

static void smooth_call(nbk_tree* t, int k, int veldens_k, double* rho, double* hsm, int flags, const uint8_t* active = nullptr) {
    require_knn_tree(t);
    NBK_REQUIRE(t->treetype == NBK_TPHYS || veldens_k == 0, NBK_ERR_UNSUPPORTED, "CalcVelDensity needs a physical tree");
    NBK_REQUIRE(k >= 1 && k < t->n, NBK_ERR_ARG, "smoothing needs 1 <= Nsmooth < numparts");
    TreeGuard guard(t);
    const int64_t n = t->n;
    DevBuf<double> drho(rho ? n : 0), dh(hsm ? n : 0);
    KnnArgs a;
    a.k = k; a.mode = 0; a.q0 = 0; a.q1 = t->n_main ? t->n_main : n;      // with a halo attached only the main particles are queries
    a.periodic = false;                       // quirk Q2: all Calc* searches ignore the period (KDCalcSmoothQuantities.cxx:227,338)
    a.rho = rho ? drho.p : nullptr; a.hsm = hsm ? dh.p : nullptr; a.veldens_k = veldens_k;
    DevBuf<uint8_t> dact, dact_tree;
    if (active) {
        const uint8_t* src = active;
        if (!(flags & NBK_DEVICE_PTRS)) {
            dact.alloc(n);
            NBK_CHECK(cudaMemcpyAsync(dact.p, active, (size_t)n, cudaMemcpyHostToDevice, t->stream));
            src = dact.p;
        }
        if (flags & NBK_TREE_ORDER) a.active = src;
        else {
            dact_tree.alloc(n);
            gather_u8_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src, dact_tree.p);
            a.active = dact_tree.p;
        }
    }
    if (hsm) NBK_CHECK(cudaMemsetAsync(dh.p, 0, dh.bytes(), t->stream));
    CallTimer tm(*t);
    if (rho) NBK_CHECK(cudaMemsetAsync(drho.p, 0, drho.bytes(), t->stream));
    NBK_CHECK(cudaEventRecord(t->ev2, t->stream));
    launch_knn(*t, a);
    NBK_CHECK(cudaEventRecord(t->ev3, t->stream));
    tm.stop();
    float ms = 0;
    NBK_CHECK(cudaEventElapsedTime(&ms, t->ev2, t->ev3));
    t->last_kernel_ms = ms; t->last_launches += (rho ? 1 : 0);
    if (rho) deliver_f64(t, drho.p, rho, flags);
    if (hsm) deliver_f64(t, dh.p, hsm, flags);
}

int nbk_calc_density(nbk_tree* t, int nsmooth, double* rho, double* hsm, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && rho, NBK_ERR_ARG, "nbk_calc_density: null argument");
    smooth_call(t, nsmooth, 0, rho, hsm, flags);
    NBK_API_END
}
int nbk_calc_density_subset(nbk_tree* t, int nsmooth, const uint8_t* active, double* rho, double* hsm, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && rho && active, NBK_ERR_ARG, "nbk_calc_density_subset: null argument");
    smooth_call(t, nsmooth, 0, rho, hsm, flags, active);
    NBK_API_END
}
int nbk_calc_veldensity(nbk_tree* t, int nsmooth, int nsearch, double* rho, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && rho, NBK_ERR_ARG, "nbk_calc_veldensity: null argument");
    if (nsmooth > nsearch) nsmooth = nsearch;   // KDCalcSmoothQuantities.cxx:319-322
    NBK_REQUIRE(nsmooth >= 1, NBK_ERR_ARG, "nbk_calc_veldensity: Nsmooth must be >= 1");
    smooth_call(t, nsearch, nsmooth, rho, nullptr, flags);
    NBK_API_END
}
int nbk_smoothing_scale(nbk_tree* t, int nsmooth, double* hsm, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && hsm, NBK_ERR_ARG, "nbk_smoothing_scale: null argument");
    smooth_call(t, nsmooth, 0, nullptr, hsm, flags);
    NBK_API_END
}

// CalcSmoothVel / CalcSmoothVelDisp: rows of `width` doubles per particle (3: mean velocity, 9: dispersion tensor)
// width 3: mean velocity (moment 1) or skewness / kurtosis (moment 3 / 4, which also read the dispersions); width 9: dispersion
static void smooth_moments_call(nbk_tree* t, int k, const double* rho, const double* smvel, double* out, int width, int flags, int moment = 1,
                                const double* smdisp = nullptr) {
    require_knn_tree(t);
    require_no_halo(t, "CalcSmoothVel*");
    NBK_REQUIRE(t->treetype == NBK_TPHYS, NBK_ERR_UNSUPPORTED, "CalcSmoothVel* need a physical tree");      // KDCalcSmoothQuantities.cxx:488-491
    NBK_REQUIRE(t->sec != nullptr, NBK_ERR_ARG, "CalcSmoothVel* need velocities");
    NBK_REQUIRE(k >= 1 && k < t->n, NBK_ERR_ARG, "smoothing needs 1 <= Nsmooth < numparts");
    NBK_REQUIRE(out && (width == 3 || smvel) && (moment == 1 || (smvel && smdisp)), NBK_ERR_ARG, "CalcSmoothVel*: null argument");
    TreeGuard guard(t);
    const int64_t n = t->n;
    const bool dev = flags & NBK_DEVICE_PTRS, tree_order = flags & NBK_TREE_ORDER;
    const int tb = 256;
    // inputs -> device, tree order
    auto stage_rows = [&](const double* src, int w, DevBuf<double>& hostcopy, DevBuf<double>& tree) -> const double* {
        const double* d = src;
        if (!dev) {
            hostcopy.alloc((size_t)n * w);
            NBK_CHECK(cudaMemcpyAsync(hostcopy.p, src, sizeof(double) * n * w, cudaMemcpyHostToDevice, t->stream));
            d = hostcopy.p;
        }
        if (tree_order) return d;
        tree.alloc((size_t)n * w);
        gather_rows_f64_kernel<<<div_up(n * w, tb), tb, 0, t->stream>>>(n, w, t->order, d, tree.p);
        return tree.p;
    };
    DevBuf<double> rho_h, rho_t, sv_h, sv_t, sd_h, sd_t, acc((size_t)n * width), out_id;
    KnnArgs a;
    a.k = k; a.mode = 0; a.q0 = 0; a.q1 = n; a.periodic = false;      // quirk Q2
    CallTimer tm(*t);
    t->last_launches = 0;
    if (rho) a.rho_in = stage_rows(rho, 1, rho_h, rho_t);
    else {
        // densityset != 1: the reference calls CalcDensity(Nsmooth) first (KDCalcSmoothQuantities.cxx:492)
        rho_t.alloc(n);
        NBK_CHECK(cudaMemsetAsync(rho_t.p, 0, rho_t.bytes(), t->stream));
        KnnArgs d = a;
        d.rho = rho_t.p;
        launch_knn(*t, d);
        a.rho_in = rho_t.p;
    }
    if (width == 9 || moment > 1) a.smvel_in = stage_rows(smvel, 3, sv_h, sv_t);
    if (moment > 1) { a.smdisp_in = stage_rows(smdisp, 9, sd_h, sd_t); a.moment = moment; }
    NBK_CHECK(cudaMemsetAsync(acc.p, 0, acc.bytes(), t->stream));
    if (moment > 1) a.smhigh_out = acc.p;
    else if (width == 3) a.smvel_out = acc.p;
    else a.smdisp_out = acc.p;
    const int64_t before = t->last_launches;
    launch_knn(*t, a);
    t->last_launches += before;
    tm.stop();
    t->last_kernel_ms = t->last_call_ms;
    const double* res = acc.p;
    if (!tree_order) {
        if (dev) { scatter_rows_f64_kernel<<<div_up(n * width, tb), tb, 0, t->stream>>>(n, width, t->order, acc.p, out); res = nullptr; }
        else {
            out_id.alloc((size_t)n * width);
            scatter_rows_f64_kernel<<<div_up(n * width, tb), tb, 0, t->stream>>>(n, width, t->order, acc.p, out_id.p);
            res = out_id.p;
        }
    }
    if (res) NBK_CHECK(cudaMemcpyAsync(out, res, sizeof(double) * n * width, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->stream));
    NBK_CHECK(cudaStreamSynchronize(t->stream));
    NBK_CHECK(cudaGetLastError());
}

int nbk_calc_smooth_vel(nbk_tree* t, int nsmooth, const double* rho, double* smvel, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && smvel, NBK_ERR_ARG, "nbk_calc_smooth_vel: null argument");
    smooth_moments_call(t, nsmooth, rho, nullptr, smvel, 3, flags);
    NBK_API_END
}
int nbk_calc_smooth_veldisp(nbk_tree* t, int nsmooth, const double* rho, const double* smvel, double* smveldisp, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && smvel && smveldisp, NBK_ERR_ARG, "nbk_calc_smooth_veldisp: null argument");
    smooth_moments_call(t, nsmooth, rho, smvel, smveldisp, 9, flags);
    NBK_API_END
}

int nbk_calc_smooth_velskew(nbk_tree* t, int nsmooth, const double* rho, const double* smvel, const double* smveldisp, double* smvelskew, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && smvel && smveldisp && smvelskew, NBK_ERR_ARG, "nbk_calc_smooth_velskew: null argument");
    smooth_moments_call(t, nsmooth, rho, smvel, smvelskew, 3, flags, 3, smveldisp);
    NBK_API_END
}
int nbk_calc_smooth_velkurtosis(nbk_tree* t, int nsmooth, const double* rho, const double* smvel, const double* smveldisp, double* smvelkurt, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && smvel && smveldisp && smvelkurt, NBK_ERR_ARG, "nbk_calc_smooth_velkurtosis: null argument");
    smooth_moments_call(t, nsmooth, rho, smvel, smvelkurt, 3, flags, 4, smveldisp);
    NBK_API_END
}

static void fof_call(nbk_tree* t, FofArgs& a, const int32_t* precheck, int32_t* group, int64_t* ngroups, nbk_fof_lists* lists, int flags) {
    NBK_REQUIRE(group && ngroups, NBK_ERR_ARG, "FOF: null output");
    require_no_halo(t, "FOF");
    NBK_REQUIRE(t->treetype == NBK_TPHYS || t->treetype == NBK_TPHS, NBK_ERR_UNSUPPORTED, "FOF needs a TPHYS or TPHS tree");
    TreeGuard guard(t);
    const int64_t n = t->n;
    const bool dev = flags & NBK_DEVICE_PTRS;
    DevBuf<int32_t> dpre, dpre_tree, dgroup(n), dlen, dhead, dnext, dtail;
    if (precheck) {
        const int32_t* src = precheck;
        if (!dev) {
            dpre.alloc(n);
            NBK_CHECK(cudaMemcpyAsync(dpre.p, precheck, sizeof(int32_t) * n, cudaMemcpyHostToDevice, t->stream));
            src = dpre.p;
        }
        if (flags & NBK_TREE_ORDER) a.precheck_tree = src;
        else {
            dpre_tree.alloc(n);
            gather_i32_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src, dpre_tree.p);
            a.precheck_tree = dpre_tree.p;
        }
    }
    a.group_tree = dgroup.p;
    if (lists && lists->len) { dlen.alloc(n + 1); a.len = dlen.p; }
    if (lists && lists->head) { if (dev) a.head = lists->head; else { dhead.alloc(n); a.head = dhead.p; } }
    if (lists && lists->next) { if (dev) a.next = lists->next; else { dnext.alloc(n); a.next = dnext.p; } }
    if (lists && lists->tail) { if (dev) a.tail = lists->tail; else { dtail.alloc(n); a.tail = dtail.p; } }
    CallTimer tm(*t);
    launch_fof(*t, a);
    tm.stop();
    *ngroups = a.ngroups;
    deliver_i32(t, dgroup.p, group, flags);
    if (lists && !dev) {
        if (lists->head) NBK_CHECK(cudaMemcpyAsync(lists->head, dhead.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, t->stream));
        if (lists->next) NBK_CHECK(cudaMemcpyAsync(lists->next, dnext.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, t->stream));
        if (lists->tail) NBK_CHECK(cudaMemcpyAsync(lists->tail, dtail.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
    if (lists && lists->len) {
        NBK_CHECK(cudaMemcpyAsync(lists->len, dlen.p, sizeof(int32_t) * (a.ngroups + 1), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
}

int nbk_fof(nbk_tree* t, double fdist, int minnum, int order, const int32_t* precheck, int32_t* group, int64_t* ngroups,
            nbk_fof_lists* lists, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_fof: null tree");
    FofArgs a;
    a.mode = t->treetype == NBK_TPHS ? 1 : 0;
    a.p0 = fdist * fdist;                       // KDFOF.cxx:33
    a.prune_x2 = a.p0;
    a.minnum = minnum; a.order = order;
    fof_call(t, a, precheck, group, ngroups, lists, flags);
    NBK_API_END
}

int nbk_fof_criterion(nbk_tree* t, int criterion, const double* params, int minnum, int order, const int32_t* precheck,
                      int32_t* group, int64_t* ngroups, nbk_fof_lists* lists, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && params, NBK_ERR_ARG, "nbk_fof_criterion: null argument");
    NBK_REQUIRE(criterion == NBK_FOF3D || criterion == NBK_FOF6D, criterion == NBK_FOFVEL ? NBK_ERR_UNSUPPORTED : NBK_ERR_ARG,
                "nbk_fof_criterion: only FOF3d and FOF6d have device implementations (host FOFcompfunc callbacks cannot run on the GPU)");
    FofArgs a;
    a.mode = criterion == NBK_FOF3D ? 2 : 4;
    a.p0 = params[6]; a.p1 = params[7];
    // any linked pair has sum(dx^2)/params[6] < 1; the tiny factor covers the rounding of the divided sum
    a.prune_x2 = params[6] * (1.0 + 1e-12);
    a.minnum = minnum; a.order = order;
    fof_call(t, a, precheck, group, ngroups, lists, flags);
    NBK_API_END
}

int nbk_fof_criterion_basis(nbk_tree* t, int criterion, const double* params, int minnum, int order, const int32_t* check,
                            int32_t* group, int64_t* ngroups, nbk_fof_lists* lists, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && params && check, NBK_ERR_ARG, "nbk_fof_criterion_basis: null argument");
    NBK_REQUIRE(criterion == NBK_FOF3D || criterion == NBK_FOF6D, criterion == NBK_FOFVEL ? NBK_ERR_UNSUPPORTED : NBK_ERR_ARG,
                "nbk_fof_criterion_basis: only FOF3d and FOF6d have device implementations");
    FofArgs a;
    a.mode = criterion == NBK_FOF3D ? 2 : 4;
    a.p0 = params[6]; a.p1 = params[7];
    a.prune_x2 = params[6] * (1.0 + 1e-12);
    a.minnum = minnum; a.order = order;
    a.attach = true;
    fof_call(t, a, check, group, ngroups, lists, flags);
    NBK_API_END
}

struct CritSpec { int mode = 0; double p0 = 0, p1 = 0, prune_x2 = 0; };

// criterion code + reference params[] -> device predicate (shared by FOFCriterion* and SearchCriterion*)
static CritSpec crit_from_params(int criterion, const double* params, const char* who) {
    if (!(criterion == NBK_FOF3D || criterion == NBK_FOF6D))
        throw Error(criterion == NBK_FOFVEL ? NBK_ERR_UNSUPPORTED : NBK_ERR_ARG,
                    std::string(who) + ": only FOF3d and FOF6d have device implementations (host FOFcompfunc callbacks cannot run on the GPU; "
                                       "FOFVel's result depends on the reference's traversal order, see DESIGN.md)");
    CritSpec c;
    c.mode = criterion == NBK_FOF3D ? 2 : 4;
    c.p0 = params[6]; c.p1 = params[7];
    // any accepted pair has sum(dx^2)/params[6] < 1; the tiny factor covers the rounding of the divided sum
    c.prune_x2 = params[6] * (1.0 + 1e-12);
    return c;
}

int nbk_fof_roots(nbk_tree* t, int criterion, double fdist, const double* params, const int32_t* precheck, int32_t* root, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && root, NBK_ERR_ARG, "nbk_fof_roots: null argument");
    require_no_halo(t, "FOF");
    NBK_REQUIRE(t->treetype == NBK_TPHYS || t->treetype == NBK_TPHS, NBK_ERR_UNSUPPORTED, "FOF needs a TPHYS or TPHS tree");
    FofArgs a;
    if (criterion < 0) {
        NBK_REQUIRE(fdist > 0, NBK_ERR_ARG, "nbk_fof_roots: linking length must be positive");
        a.mode = t->treetype == NBK_TPHS ? 1 : 0;
        a.p0 = fdist * fdist; a.prune_x2 = a.p0;
    } else {
        NBK_REQUIRE(params, NBK_ERR_ARG, "nbk_fof_roots: null params");
        CritSpec c = crit_from_params(criterion, params, "nbk_fof_roots");
        a.mode = c.mode; a.p0 = c.p0; a.p1 = c.p1; a.prune_x2 = c.prune_x2;
    }
    a.minnum = 1; a.order = 0;
    TreeGuard guard(t);
    const int64_t n = t->n;
    const bool dev = flags & NBK_DEVICE_PTRS;
    DevBuf<int32_t> dpre, dpre_tree, droot(n), dout;
    if (precheck) {
        const int32_t* src = precheck;
        if (!dev) { dpre.alloc(n); NBK_CHECK(cudaMemcpyAsync(dpre.p, precheck, sizeof(int32_t) * n, cudaMemcpyHostToDevice, t->stream)); src = dpre.p; }
        dpre_tree.alloc(n);
        gather_i32_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, src, dpre_tree.p);
        a.precheck_tree = dpre_tree.p;
    }
    a.roots_tree = droot.p;
    CallTimer tm(*t);
    launch_fof(*t, a);
    tm.stop();
    int32_t* out = root;
    if (!dev) { dout.alloc(n); out = dout.p; }
    roots_to_ids_kernel<<<div_up(n, 256), 256, 0, t->stream>>>(n, t->order, droot.p, out);
    if (!dev) NBK_CHECK(cudaMemcpyAsync(root, dout.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, t->stream));
    NBK_CHECK(cudaStreamSynchronize(t->stream));
    NBK_API_END
}

int nbk_union_pairs(int device, int64_t nnodes, int64_t npairs, const int32_t* a, const int32_t* b, int32_t* root) {
    NBK_API_BEGIN
    NBK_REQUIRE(nnodes >= 0 && npairs >= 0 && (npairs == 0 || (a && b)) && (nnodes == 0 || root), NBK_ERR_ARG, "nbk_union_pairs: bad argument");
    int dev = device;
    if (dev < 0) NBK_CHECK(cudaGetDevice(&dev));
    DeviceGuard guard(dev, nullptr);
    launch_union_pairs(nullptr, nnodes, npairs, a, b, root);
    NBK_API_END
}

static void ball_call(nbk_tree* t, double fdist2, const CritSpec* crit, int64_t m, const int32_t* qidx, const double* x, const double* vq,
                      int64_t* offsets, int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags) {
    NBK_REQUIRE(offsets && total, NBK_ERR_ARG, "ball search: null output");
    NBK_REQUIRE(m >= 0 && cap >= 0, NBK_ERR_ARG, "ball search: negative count");
    require_no_halo(t, "ball search");
    NBK_REQUIRE(t->treetype == NBK_TPHYS || t->treetype == NBK_TPHS, NBK_ERR_UNSUPPORTED, "SearchBallPos needs positions as tree coordinates");
    TreeGuard guard(t);
    const bool dev = flags & NBK_DEVICE_PTRS;
    if (m == 0) {                               // no queries: one row offset (0), nothing else
        *total = 0;
        if (dev) { NBK_CHECK(cudaMemsetAsync(offsets, 0, sizeof(int64_t), t->stream)); NBK_CHECK(cudaStreamSynchronize(t->stream)); }
        else offsets[0] = 0;
        return;
    }
    DevBuf<int32_t> dq, didx;
    DevBuf<double> dx, dv, dd2;
    DevBuf<int64_t> doff;
    BallArgs a;
    a.r2 = fdist2; a.m = m; a.cap = idx ? cap : 0; a.out_ids = flags & NBK_OUT_IDS;
    if (crit) { a.mode = crit->mode; a.p0 = crit->p0; a.p1 = crit->p1; a.prune_x2 = crit->prune_x2; }
    if (dev) { a.qidx = qidx; a.xq = x; a.vq = vq; a.offsets = offsets; a.idx = idx; a.d2 = idx ? d2 : nullptr; }
    else {
        if (qidx) { dq.alloc(m); NBK_CHECK(cudaMemcpyAsync(dq.p, qidx, sizeof(int32_t) * m, cudaMemcpyHostToDevice, t->stream)); a.qidx = dq.p; }
        if (x) { dx.alloc((size_t)3 * m); NBK_CHECK(cudaMemcpyAsync(dx.p, x, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, t->stream)); a.xq = dx.p; }
        if (vq) { dv.alloc((size_t)3 * m); NBK_CHECK(cudaMemcpyAsync(dv.p, vq, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, t->stream)); a.vq = dv.p; }
        doff.alloc(m + 1); a.offsets = doff.p;
        if (idx && cap > 0) { didx.alloc(cap); a.idx = didx.p; if (d2) { dd2.alloc(cap); a.d2 = dd2.p; } }
    }
    if (qidx && m > 0) {
        // tree indices come from the caller: check them on the device side of the boundary
        DevBuf<int> bad(1);
        NBK_CHECK(cudaMemsetAsync(bad.p, 0, sizeof(int), t->stream));
        check_indices_kernel<<<div_up(m, 256), 256, 0, t->stream>>>(m, a.qidx, (int32_t)t->n, bad.p);
        int hb = 0;
        NBK_CHECK(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
        NBK_REQUIRE(hb == 0, NBK_ERR_ARG, "ball search: particle index out of range");
    }
    CallTimer tm(*t);
    launch_ball(*t, a);
    tm.stop();
    *total = a.total;
    if (!dev) {
        NBK_CHECK(cudaMemcpyAsync(offsets, doff.p, sizeof(int64_t) * (m + 1), cudaMemcpyDeviceToHost, t->stream));
        if (idx && cap > 0) NBK_CHECK(cudaMemcpyAsync(idx, didx.p, sizeof(int32_t) * std::min(cap, a.total), cudaMemcpyDeviceToHost, t->stream));
        if (idx && d2 && cap > 0) NBK_CHECK(cudaMemcpyAsync(d2, dd2.p, sizeof(double) * std::min(cap, a.total), cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
    }
    if (idx && a.total > cap) throw Error(NBK_ERR_CAPACITY, "ball search: index buffer too small (see *total)");
}

int nbk_ball_particles(nbk_tree* t, double fdist2, int64_t m, const int32_t* qidx, int64_t* offsets, int32_t* idx, double* d2, int64_t cap,
                       int64_t* total, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && (qidx || m == 0), NBK_ERR_ARG, "nbk_ball_particles: null argument");
    ball_call(t, fdist2, nullptr, m, qidx, nullptr, nullptr, offsets, idx, d2, cap, total, flags);
    NBK_API_END
}
int nbk_ball_points(nbk_tree* t, double fdist2, int64_t m, const double* x, int64_t* offsets, int32_t* idx, double* d2, int64_t cap,
                    int64_t* total, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && (x || m == 0), NBK_ERR_ARG, "nbk_ball_points: null argument");
    ball_call(t, fdist2, nullptr, m, nullptr, x, nullptr, offsets, idx, d2, cap, total, flags);
    NBK_API_END
}
int nbk_search_criterion_particles(nbk_tree* t, int criterion, const double* params, int64_t m, const int32_t* qidx, int64_t* offsets,
                                   int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && params && (qidx || m == 0), NBK_ERR_ARG, "nbk_search_criterion_particles: null argument");
    CritSpec c = crit_from_params(criterion, params, "nbk_search_criterion_particles");
    ball_call(t, 0.0, &c, m, qidx, nullptr, nullptr, offsets, idx, d2, cap, total, flags);
    NBK_API_END
}
int nbk_search_criterion_points(nbk_tree* t, int criterion, const double* params, int64_t m, const double* x, const double* v,
                                int64_t* offsets, int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && params && (x || m == 0), NBK_ERR_ARG, "nbk_search_criterion_points: null argument");
    CritSpec c = crit_from_params(criterion, params, "nbk_search_criterion_points");
    NBK_REQUIRE(c.mode != 4 || v || m == 0, NBK_ERR_ARG, "nbk_search_criterion_points: FOF6d needs the query velocities");
    ball_call(t, 0.0, &c, m, nullptr, x, v, offsets, idx, d2, cap, total, flags);
    NBK_API_END
}

// ---- *Particle / *Position forms of the smoothed estimators ------------------------------------------------------
static void smooth_gather_call(nbk_tree* t, int k, int veldens_k, int64_t m, const int32_t* qidx, const double* x, const double* v,
                               double* out, int flags) {
    require_knn_tree(t);
    require_no_halo(t, "Calc*Particle / Calc*Position");
    NBK_REQUIRE(t->treetype == NBK_TPHYS || veldens_k == 0, NBK_ERR_UNSUPPORTED, "CalcVelDensity needs a physical tree");
    NBK_REQUIRE(k >= 1 && k < t->n, NBK_ERR_ARG, "smoothing needs 1 <= Nsmooth < numparts");
    NBK_REQUIRE(m >= 0 && out, NBK_ERR_ARG, "Calc*Particle / Calc*Position: bad count or null output");
    if (m == 0) return;
    TreeGuard guard(t);
    const bool dev = flags & NBK_DEVICE_PTRS;
    DevBuf<int32_t> dq;
    DevBuf<double> dx, dv, dout;
    KnnArgs a;
    a.k = k; a.veldens_k = veldens_k; a.gather = true;
    a.periodic = false;                       // quirk Q2: every Calc* search ignores the period (KDCalcSmoothQuantities.cxx:784,1108)
    if (x) {
        a.mode = 1; a.q0 = 0; a.q1 = m;
        if (dev) { a.xq = x; a.vq = v; }
        else {
            dx.alloc((size_t)3 * m);
            NBK_CHECK(cudaMemcpyAsync(dx.p, x, dx.bytes(), cudaMemcpyHostToDevice, t->stream));
            a.xq = dx.p;
            if (v) { dv.alloc((size_t)3 * m); NBK_CHECK(cudaMemcpyAsync(dv.p, v, dv.bytes(), cudaMemcpyHostToDevice, t->stream)); a.vq = dv.p; }
        }
    } else if (qidx) {
        a.mode = 0; a.nq = m;
        if (dev) a.qlist = qidx;
        else { dq.alloc(m); NBK_CHECK(cudaMemcpyAsync(dq.p, qidx, dq.bytes(), cudaMemcpyHostToDevice, t->stream)); a.qlist = dq.p; }
        DevBuf<int> bad(1);
        NBK_CHECK(cudaMemsetAsync(bad.p, 0, sizeof(int), t->stream));
        check_indices_kernel<<<div_up(m, 256), 256, 0, t->stream>>>(m, a.qlist, (int32_t)t->n, bad.p);
        int hb = 0;
        NBK_CHECK(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, t->stream));
        NBK_CHECK(cudaStreamSynchronize(t->stream));
        NBK_REQUIRE(hb == 0, NBK_ERR_ARG, "Calc*Particle: particle index out of range");
    } else {
        NBK_REQUIRE(m <= t->n, NBK_ERR_ARG, "Calc*Particle: more queries than particles");
        a.mode = 0; a.q0 = 0; a.q1 = m;       // no list: tree indices 0..m-1 (m = n: every particle, 32 neighbouring queries per warp)
    }
    if (dev) a.rho = out;
    else { dout.alloc(m); a.rho = dout.p; }
    CallTimer tm(*t);
    launch_knn(*t, a);
    tm.stop();
    t->last_kernel_ms = t->last_call_ms;
    if (!dev) NBK_CHECK(cudaMemcpyAsync(out, dout.p, sizeof(double) * m, cudaMemcpyDeviceToHost, t->stream));
    NBK_CHECK(cudaStreamSynchronize(t->stream));
}

int nbk_calc_density_particles(nbk_tree* t, int nsmooth, int64_t m, const int32_t* qidx, double* rho, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_calc_density_particles: null tree");
    smooth_gather_call(t, nsmooth, 0, m, qidx, nullptr, nullptr, rho, flags);
    NBK_API_END
}
int nbk_calc_veldensity_particles(nbk_tree* t, int nsmooth, int nsearch, int64_t m, const int32_t* qidx, double* rho, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_calc_veldensity_particles: null tree");
    if (nsmooth > nsearch) nsmooth = nsearch;   // KDCalcSmoothQuantities.cxx:855-858
    NBK_REQUIRE(nsmooth >= 1, NBK_ERR_ARG, "nbk_calc_veldensity_particles: Nsmooth must be >= 1");
    smooth_gather_call(t, nsearch, nsmooth, m, qidx, nullptr, nullptr, rho, flags);
    NBK_API_END
}
int nbk_calc_density_points(nbk_tree* t, int nsmooth, int64_t m, const double* x, double* rho, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && (x || m == 0), NBK_ERR_ARG, "nbk_calc_density_points: null argument");
    smooth_gather_call(t, nsmooth, 0, m, nullptr, x, nullptr, rho, flags);
    NBK_API_END
}
int nbk_calc_veldensity_points(nbk_tree* t, int nsmooth, int nsearch, int64_t m, const double* x, const double* v, double* rho, int flags) {
    NBK_API_BEGIN
    NBK_REQUIRE(t && ((x && v) || m == 0), NBK_ERR_ARG, "nbk_calc_veldensity_points: null argument");
    if (nsmooth > nsearch) nsmooth = nsearch;   // KDCalcSmoothQuantities.cxx:1160-1163
    NBK_REQUIRE(nsmooth >= 1, NBK_ERR_ARG, "nbk_calc_veldensity_points: Nsmooth must be >= 1");
    smooth_gather_call(t, nsearch, nsmooth, m, nullptr, x, v, rho, flags);
    NBK_API_END
}

int nbk_set_option(const char* name, int64_t value) {
    NBK_API_BEGIN
    NBK_REQUIRE(name != nullptr, NBK_ERR_ARG, "nbk_set_option: null name");
    NBK_REQUIRE(set_knn_option(name, value) || set_fof_option(name, value), NBK_ERR_ARG, std::string("nbk_set_option: unknown option ") + name);
    NBK_API_END
}

int nbk_release_cached_memory(int device) {
    NBK_API_BEGIN
    if (device < 0) NBK_CHECK(cudaGetDevice(&device));
    DeviceGuard guard(device);
    NBK_CHECK(cudaDeviceSynchronize());
    cudaMemPool_t pool = nbk_pool(device);
    if (pool) NBK_CHECK(cudaMemPoolTrimTo(pool, 0));
    NBK_API_END
}

int nbk_device_arrays(const nbk_tree* t, const void** pos4, const void** vel4, const void** mass, const int32_t** order) {
    NBK_API_BEGIN
    NBK_REQUIRE(t, NBK_ERR_ARG, "nbk_device_arrays: null tree");
    if (pos4) *pos4 = t->pos4();
    if (vel4) *vel4 = t->vel4();
    if (mass) *mass = t->mass;
    if (order) *order = t->order;
    NBK_API_END
}

}  // extern "C"
