// sharded.cu -- the slab-sharded tree behind the C ABI (nbk_comm_* / nbk_sharded_*): one process per GPU, NCCL over
// NVLink for the exchange steps, the single-GPU entry points of this library for all the work.
//
// The reference has no distributed code: VELOCIraptor decomposes the domain over MPI ranks outside the library
// and builds one local KDTree per rank.  This file is that outer layer for one node of GPUs, reachable from C / C++ (an MPI
// rank broadcasts the 128-byte communicator id with MPI_Bcast and calls nbk_comm_init_rank; the tests do the same with
// torch.distributed).  nbodylib_b200/sharded.py drives the same algorithm with torch collectives and a pluggable engine
// (its CPU tests pin the host logic); the two are compared on hardware by tests/test_gpu_sharded.py.
//
//   * decomposition: the global box is cut into `nranks` slabs along x; every rank passes the particles of its slab in
//     GLOBAL coordinates, which are never shifted (ghosts keep their owners' exact values);
//   * halo exchange: particles within h of a slab face go to the neighbour across it (grouped ncclSend / ncclRecv); h is one
//     number for the group (ncclAllReduce max) and, for three or more ranks, must stay below the slab width;
//   * CalcDensity: owned particles in the main tree, ghosts in an attached second tree (nbk_attach_halo); after the pass
//     every owned k-ball is checked against the halo (ncclAllReduce of the violation count; widen and repeat); the
//     symmetric scatter terms deposited on ghosts travel home and are added there;
//   * FOF / FOFCriterion: local components over owned + ghosts on a tree that is periodic with the global periods
//     (nbk_fof_roots), cross-slab edges = (my component of a particle I sent, the neighbour's component of its ghost copy),
//     ncclAllGather of the edges and of the boundary components' owned sizes, the same union on every rank
//     (launch_union_pairs), minnum filter, one global numbering.
#include <nccl.h>

#include <algorithm>
#include <memory>
#include <numeric>

#include "../../include/nbk_sharded.h"
#include "sort_scan.cuh"
#include "tree.h"

struct nbk_comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1, device = 0;
    cudaStream_t stream = nullptr;
};

namespace nbk {

#define NBK_NCCL(call)                                                                                     \
    do {                                                                                                   \
        ncclResult_t _r = (call);                                                                          \
        if (_r != ncclSuccess) {                                                                           \
            char _b[512];                                                                                  \
            snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(_r)); \
            throw ::nbk::Error(NBK_ERR_CUDA, _b);                                                          \
        }                                                                                                  \
    } while (0)

// ---- small kernels ------------------------------------------------------------------------------------------------------
template <class R>
__global__ void sh_face_flags_kernel(int64_t n, const R* __restrict__ pos, double lo, double hi, int want_lo, int want_hi, uint32_t* fl_lo, uint32_t* fl_hi) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = (double)pos[3 * i];
    fl_lo[i] = (want_lo && x < lo) ? 1u : 0u;          // within h of the left face: x < x0 + h
    fl_hi[i] = (want_hi && x >= hi) ? 1u : 0u;         // within h of the right face: x >= x1 - h
}
__global__ void sh_compact_kernel(int64_t n, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ scan, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[scan[i]] = (int32_t)i;
}
// rows of doubles: pos(3) [vel(3)] [mass(1)] of the listed particles
template <class R>
__global__ void sh_gather_rows_kernel(int64_t m, const int32_t* __restrict__ idx, const R* __restrict__ pos, const R* __restrict__ vel, const R* __restrict__ mass,
                                      int cols, double* __restrict__ rows) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t i = idx[j];
    double* r = rows + j * cols;
    r[0] = (double)pos[3 * i]; r[1] = (double)pos[3 * i + 1]; r[2] = (double)pos[3 * i + 2];
    int c = 3;
    if (vel) { r[c] = (double)vel[3 * i]; r[c + 1] = (double)vel[3 * i + 1]; r[c + 2] = (double)vel[3 * i + 2]; c += 3; }
    if (mass) r[c] = (double)mass[i];
}
__global__ void sh_gid_kernel(int64_t m, const int32_t* __restrict__ idx, int64_t gid0, int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[j] = gid0 + (idx ? (int64_t)idx[j] : j);
}
// local arrays of a slab tree: owned particles followed by the ghost rows
template <class R>
__global__ void sh_concat_kernel(int64_t n, int64_t g, const R* __restrict__ own, const double* __restrict__ rows, int cols, int col0, int width, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + g) return;
    for (int c = 0; c < width; c++) out[i * width + c] = i < n ? (double)own[i * width + c] : rows[(i - n) * cols + col0 + c];
}
template <class R>
__global__ void sh_widen_kernel(int64_t n, const R* __restrict__ in, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
__global__ void sh_col_kernel(int64_t g, const double* __restrict__ rows, int cols, int col0, int width, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g) return;
    for (int c = 0; c < width; c++) out[i * width + c] = rows[i * cols + col0 + c];
}
template <class R>
__global__ void sh_halo_check_kernel(int64_t n, const R* __restrict__ pos, const double* __restrict__ hsm, double x0, double x1, double h, int chk_lo, int chk_hi,
                                     unsigned long long* bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool b = false;
    if (i < n) {
        const double x = (double)pos[3 * i], rk = 2.0 * hsm[i];
        b = (chk_lo && rk > (x - x0) + h) || (chk_hi && rk > (x1 - x) + h);
    }
    unsigned m = __ballot_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(bad, (unsigned long long)__popc(m));
}
__global__ void sh_index_add_kernel(int64_t m, const int32_t* __restrict__ idx, const double* __restrict__ val, double* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) atomicAdd(&out[idx[j]], val[j]);
}
__global__ void sh_count_roots_kernel(int64_t n, const int32_t* __restrict__ root, int32_t* __restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && root[i] >= 0) atomicAdd(&cnt[root[i]], 1);
}
__global__ void sh_names_kernel(int64_t m, const int32_t* __restrict__ which, int64_t off, const int32_t* __restrict__ root, const int64_t* __restrict__ gid,
                                int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t i = which ? (int64_t)which[j] : off + j;
    out[j] = gid[root[i]];
}
__global__ void sh_touch_kernel(int64_t m, const int32_t* __restrict__ which, int64_t off, const int32_t* __restrict__ root, uint32_t* __restrict__ touch) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t i = which ? (int64_t)which[j] : off + j;
    touch[root[i]] = 1u;
}
__global__ void sh_edge_kernel(int64_t m, const int64_t* __restrict__ mine, const int64_t* __restrict__ peer, uint32_t* __restrict__ flag) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) flag[j] = mine[j] != peer[j] ? 1u : 0u;
}
__global__ void sh_edge_compact_kernel(int64_t m, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ scan, const int64_t* __restrict__ mine,
                                       const int64_t* __restrict__ peer, int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m && flag[j]) { out[2 * (int64_t)scan[j]] = mine[j]; out[2 * (int64_t)scan[j] + 1] = peer[j]; }
}
__global__ void sh_node_table_kernel(int64_t m, const int32_t* __restrict__ tidx, const int64_t* __restrict__ gid, const int32_t* __restrict__ cnt, int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) { out[2 * j] = gid[tidx[j]]; out[2 * j + 1] = (int64_t)cnt[tidx[j]]; }
}
__global__ void sh_unique_flag_kernel(int64_t m, const uint64_t* __restrict__ sorted, uint32_t* __restrict__ flag) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) flag[j] = (j == 0 || sorted[j] != sorted[j - 1]) ? 1u : 0u;
}
__global__ void sh_unique_compact_kernel(int64_t m, const uint64_t* __restrict__ sorted, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ scan,
                                         int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m && flag[j]) out[scan[j]] = (int64_t)sorted[j];
}
__device__ __forceinline__ int sh_search(const int64_t* __restrict__ names, int nn, int64_t key) {
    int lo = 0, hi = nn - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (names[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void sh_lookup_kernel(int64_t m, const int64_t* __restrict__ keys, int stride, const int64_t* __restrict__ names, int nn, int32_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[j] = sh_search(names, nn, keys[j * stride]);
}
__global__ void sh_comp_size_kernel(int64_t m, const int64_t* __restrict__ nodes, const int64_t* __restrict__ names, int nn, const int32_t* __restrict__ comp,
                                    unsigned long long* __restrict__ size) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) atomicAdd(&size[comp[sh_search(names, nn, nodes[2 * j])]], (unsigned long long)nodes[2 * j + 1]);
}
__global__ void sh_valid_comp_kernel(int nn, const int32_t* __restrict__ comp, const unsigned long long* __restrict__ size, int minnum, uint32_t* __restrict__ flag) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nn) flag[j] = (comp[j] == j && size[j] >= (unsigned long long)minnum) ? 1u : 0u;
}
__global__ void sh_interior_flag_kernel(int64_t n_all, const int32_t* __restrict__ root, const uint32_t* __restrict__ touch, const int32_t* __restrict__ cnt, int minnum,
                                        uint32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_all) flag[i] = (root[i] == (int32_t)i && !touch[i] && cnt[i] >= minnum) ? 1u : 0u;
}
__global__ void sh_sizes_kernel(int64_t m, const int32_t* __restrict__ idx, const int32_t* __restrict__ cnt, int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[j] = (int64_t)cnt[idx[j]];
}
__global__ void sh_csize_kernel(int64_t m, const int32_t* __restrict__ idx, const unsigned long long* __restrict__ size, int64_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[j] = (int64_t)size[idx[j]];
}
__global__ void sh_scatter_i32_kernel(int64_t m, const int32_t* __restrict__ idx, const int32_t* __restrict__ val, int32_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) out[idx[j]] = val[j];
}
__global__ void sh_touch_lut_kernel(int64_t m, const int32_t* __restrict__ tidx, const int64_t* __restrict__ gid, const int64_t* __restrict__ names, int nn,
                                    const int32_t* __restrict__ comp, const int32_t* __restrict__ comp_gid, int32_t* __restrict__ lut) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) lut[tidx[j]] = comp_gid[comp[sh_search(names, nn, gid[tidx[j]])]];
}
__global__ void sh_label_kernel(int64_t n, const int32_t* __restrict__ root, const int32_t* __restrict__ lut, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = root[i] >= 0 ? lut[root[i]] : 0;
}

// indices of the set flags (ascending); returns the count
static int64_t compact_flags(cudaStream_t st, int64_t n, uint32_t* flag /* n + 1, clobbered? no: kept */, DevBuf<int32_t>& out) {
    if (n <= 0) { out.alloc(0); return 0; }
    DevBuf<uint32_t> scan(n + 1), scratch(scan_scratch_elems(n + 1));
    NBK_CHECK(cudaMemsetAsync(flag + n, 0, sizeof(uint32_t), st));
    exclusive_scan_u32(flag, scan.p, n + 1, scratch.p, st);
    uint32_t tot = 0;
    NBK_CHECK(cudaMemcpyAsync(&tot, scan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NBK_CHECK(cudaStreamSynchronize(st));
    out.alloc(tot);
    if (tot) sh_compact_kernel<<<div_up(n, 256), 256, 0, st>>>(n, flag, scan.p, out.p);
    return (int64_t)tot;
}

}  // namespace nbk

using namespace nbk;

struct nbk_sharded {
    nbk_comm* comm = nullptr;
    int64_t n = 0, gid0 = 0, n_global = 0;
    int real_bytes = 4;
    bool periodic = false, has_vel = false;
    double box[3] = {1, 1, 1}, x0 = 0, x1 = 1, slab_width = 1, h_knn = 0;
    // this rank's particles, packed [n][3] / [n] of the input real type, device
    void *pos = nullptr, *vel = nullptr, *mass = nullptr;
    // density state: main tree (+ attached halo), the lists of particles sent across each face
    nbk_tree* dens_tree = nullptr;
    int64_t dens_all = 0, dens_nl = 0, dens_nr = 0;      // particles of the density tree, ghosts received from the left / right
    DevBuf<int32_t> dens_send_l, dens_send_r;
    double dens_h = 0;
    // FOF state: local tree over owned + ghosts
    nbk_tree* fof_tree = nullptr;
    double fof_h = -1; bool fof_vel = false;
    int64_t fof_all = 0, fof_nl = 0, fof_nr = 0;
    DevBuf<int32_t> fof_send_l, fof_send_r;
    DevBuf<int64_t> fof_gid;
    // statistics of the last calls
    int64_t ghosts_knn = 0, ghosts_fof = 0, density_setups = 0, fof_setups = 0;
    double last_kernel_ms = 0, last_call_ms = 0;
    int64_t last_launches = 0, last_flagged = 0;
};

namespace {

struct CommGuard {
    int prev = -1;
    cudaStream_t prev_stream;
    explicit CommGuard(nbk_comm* c) {
        cudaGetDevice(&prev);
        if (prev != c->device) cudaSetDevice(c->device);
        prev_stream = cur_stream();
        cur_stream() = c->stream;
    }
    ~CommGuard() { cur_stream() = prev_stream; if (prev >= 0) cudaSetDevice(prev); }
};

double group_max(nbk_comm* c, double v) {
    if (c->nranks == 1) return v;
    DevBuf<double> d(1);
    NBK_CHECK(cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NBK_NCCL(ncclAllReduce(d.p, d.p, 1, ncclDouble, ncclMax, c->nccl, c->stream));
    NBK_CHECK(cudaMemcpyAsync(&v, d.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NBK_CHECK(cudaStreamSynchronize(c->stream));
    return v;
}
// every rank's (a, b) pair
std::vector<int64_t> gather_pairs(nbk_comm* c, int64_t a, int64_t b) {
    std::vector<int64_t> out(2 * (size_t)c->nranks);
    if (c->nranks == 1) { out[0] = a; out[1] = b; return out; }
    DevBuf<int64_t> d(2 * (size_t)c->nranks);
    const int64_t mine[2] = {a, b};
    NBK_CHECK(cudaMemcpyAsync(d.p + 2 * c->rank, mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    NBK_NCCL(ncclAllGather(d.p + 2 * c->rank, d.p, 2, ncclInt64, c->nccl, c->stream));
    NBK_CHECK(cudaMemcpyAsync(out.data(), d.p, sizeof(int64_t) * out.size(), cudaMemcpyDeviceToHost, c->stream));
    NBK_CHECK(cudaStreamSynchronize(c->stream));
    return out;
}
// one buffer to each face neighbour, one from each: byte counts; returns the received byte counts through the DevBufs' sizes
void face_exchange(nbk_comm* c, const void* to_left, int64_t bytes_l, const void* to_right, int64_t bytes_r, DevBuf<unsigned char>& from_left,
                   DevBuf<unsigned char>& from_right) {
    const int W = c->nranks, left = (c->rank - 1 + W) % W, right = (c->rank + 1) % W;
    std::vector<int64_t> tab = gather_pairs(c, bytes_l, bytes_r);
    const int64_t n_from_left = tab[2 * (size_t)left + 1], n_from_right = tab[2 * (size_t)right];
    from_left.alloc((size_t)n_from_left);
    from_right.alloc((size_t)n_from_right);
    NBK_NCCL(ncclGroupStart());
    if (bytes_r) NBK_NCCL(ncclSend(to_right, (size_t)bytes_r, ncclUint8, right, c->nccl, c->stream));
    if (bytes_l) NBK_NCCL(ncclSend(to_left, (size_t)bytes_l, ncclUint8, left, c->nccl, c->stream));
    if (n_from_left) NBK_NCCL(ncclRecv(from_left.p, (size_t)n_from_left, ncclUint8, left, c->nccl, c->stream));
    if (n_from_right) NBK_NCCL(ncclRecv(from_right.p, (size_t)n_from_right, ncclUint8, right, c->nccl, c->stream));
    NBK_NCCL(ncclGroupEnd());
    NBK_CHECK(cudaStreamSynchronize(c->stream));
}
// variable-length rows of int64 from every rank, concatenated in rank order
void allgather_rows(nbk_comm* c, const int64_t* rows, int64_t m, int width, DevBuf<int64_t>& out, int64_t* total) {
    const int W = c->nranks;
    if (W == 1) {
        out.alloc((size_t)(m * width));
        if (m) NBK_CHECK(cudaMemcpyAsync(out.p, rows, sizeof(int64_t) * m * width, cudaMemcpyDeviceToDevice, c->stream));
        *total = m;
        return;
    }
    std::vector<int64_t> tab = gather_pairs(c, m, 0);
    int64_t mx = 1, tot = 0;
    for (int r = 0; r < W; r++) { mx = std::max(mx, tab[2 * (size_t)r]); tot += tab[2 * (size_t)r]; }
    DevBuf<int64_t> pad((size_t)(mx * width) * (size_t)W);
    int64_t* mine = pad.p + (size_t)c->rank * (size_t)(mx * width);
    if (m) NBK_CHECK(cudaMemcpyAsync(mine, rows, sizeof(int64_t) * m * width, cudaMemcpyDeviceToDevice, c->stream));
    NBK_NCCL(ncclAllGather(mine, pad.p, (size_t)(mx * width), ncclInt64, c->nccl, c->stream));
    out.alloc((size_t)(tot * width));
    int64_t at = 0;
    for (int r = 0; r < W; r++) {
        const int64_t mr = tab[2 * (size_t)r];
        if (mr) NBK_CHECK(cudaMemcpyAsync(out.p + at * width, pad.p + (size_t)r * (size_t)(mx * width), sizeof(int64_t) * mr * width, cudaMemcpyDeviceToDevice, c->stream));
        at += mr;
    }
    NBK_CHECK(cudaStreamSynchronize(c->stream));
    *total = tot;
}

struct Halo {
    DevBuf<int32_t> send_l, send_r;
    DevBuf<unsigned char> rows_l, rows_r, gid_l, gid_r;     // received: double rows [m][cols], int64 ids
    int64_t nl = 0, nr = 0;
    int cols = 0;
};

template <class R>
void halo_exchange(nbk_sharded* s, double h, bool wrap, bool with_vel, bool with_mass, bool with_gid, Halo& H) {
    nbk_comm* c = s->comm;
    cudaStream_t st = c->stream;
    const int W = c->nranks;
    H.cols = 3 + (with_vel ? 3 : 0) + (with_mass ? 1 : 0);
    if (W == 1) return;
    const int64_t n = s->n;
    const bool want_lo = wrap || c->rank > 0, want_hi = wrap || c->rank < W - 1;
    DevBuf<uint32_t> fl(n + 1), fh(n + 1);
    sh_face_flags_kernel<R><<<div_up(n, 256), 256, 0, st>>>(n, (const R*)s->pos, s->x0 + h, s->x1 - h, want_lo, want_hi, fl.p, fh.p);
    const int64_t ml = compact_flags(st, n, fl.p, H.send_l), mr = compact_flags(st, n, fh.p, H.send_r);
    DevBuf<double> rl((size_t)(ml * H.cols)), rr((size_t)(mr * H.cols));
    if (ml) sh_gather_rows_kernel<R><<<div_up(ml, 256), 256, 0, st>>>(ml, H.send_l.p, (const R*)s->pos, with_vel ? (const R*)s->vel : nullptr, with_mass ? (const R*)s->mass : nullptr, H.cols, rl.p);
    if (mr) sh_gather_rows_kernel<R><<<div_up(mr, 256), 256, 0, st>>>(mr, H.send_r.p, (const R*)s->pos, with_vel ? (const R*)s->vel : nullptr, with_mass ? (const R*)s->mass : nullptr, H.cols, rr.p);
    face_exchange(c, rl.p, ml * H.cols * 8, rr.p, mr * H.cols * 8, H.rows_l, H.rows_r);
    H.nl = (int64_t)H.rows_l.n / (H.cols * 8);
    H.nr = (int64_t)H.rows_r.n / (H.cols * 8);
    if (with_gid) {
        DevBuf<int64_t> gl((size_t)ml), gr((size_t)mr);
        if (ml) sh_gid_kernel<<<div_up(ml, 256), 256, 0, st>>>(ml, H.send_l.p, s->gid0, gl.p);
        if (mr) sh_gid_kernel<<<div_up(mr, 256), 256, 0, st>>>(mr, H.send_r.p, s->gid0, gr.p);
        face_exchange(c, gl.p, ml * 8, gr.p, mr * 8, H.gid_l, H.gid_r);
    }
    NBK_CHECK(cudaGetLastError());
}

void check_halo(nbk_sharded* s, double h, const char* what) {
    // two ranks: the one neighbour's whole slab is everything there is, so any width is complete
    if (s->comm->nranks > 2 && !(h < s->slab_width)) {
        char b[256];
        snprintf(b, sizeof(b), "%s needs a halo of %.6g but a slab is only %.6g wide: particles two slabs away would be missing", what, h, s->slab_width);
        throw Error(NBK_ERR_ARG, b);
    }
}

nbk_tree* make_tree(nbk_sharded* s, const void* pos, const void* vel, const void* mass, int64_t n, int real_bytes, const double* period, int flags = 0) {
    nbk_particles np;
    memset(&np, 0, sizeof(np));
    np.pos = pos; np.pos_stride = 3 * real_bytes;
    np.vel = vel; np.vel_stride = 3 * real_bytes;
    np.mass = mass; np.mass_stride = real_bytes;
    np.real_bytes = real_bytes; np.on_device = 1;
    nbk_tree* t = nullptr;
    NBK_CHECK(cudaStreamSynchronize(s->comm->stream));       // the library builds on the tree's own stream
    // warp-aligned shape: a slab's particle count is arbitrary, and nobody inspects the shape of these trees
    const int rc = nbk_create(&np, n, 16, NBK_TPHYS, NBK_KEPAN, 1000, 0, period, flags | NBK_WARP_ALIGNED, s->comm->device, &t);
    if (rc != NBK_OK) throw Error(rc, nbk_last_error());
    return t;
}

void close_density(nbk_sharded* s) { if (s->dens_tree) { nbk_destroy(s->dens_tree); s->dens_tree = nullptr; } }
void close_fof(nbk_sharded* s) { if (s->fof_tree) { nbk_destroy(s->fof_tree); s->fof_tree = nullptr; } s->fof_h = -1; }

template <class R>
void density_setup(nbk_sharded* s) {
    nbk_comm* c = s->comm;
    cudaStream_t st = c->stream;
    check_halo(s, s->h_knn, "CalcDensity");
    Halo H;
    halo_exchange<R>(s, s->h_knn, false, false, true, false, H);
    const int64_t g = H.nl + H.nr;
    // fp32 input is stored as it is; fp64 input is stored as fp32 when every coordinate is representable -- a property the
    // owned particles and the ghosts must share for the two trees to be attached
    const int store = s->real_bytes == 4 ? NBK_STORE_F32 : 0;
    nbk_tree* tree = make_tree(s, s->pos, nullptr, s->mass, s->n, s->real_bytes, nullptr, store);
    if (g > 0) {
        DevBuf<double> gpos((size_t)(3 * g)), gmass((size_t)g);
        if (H.nl) {
            sh_col_kernel<<<div_up(H.nl, 256), 256, 0, st>>>(H.nl, (const double*)H.rows_l.p, H.cols, 0, 3, gpos.p);
            sh_col_kernel<<<div_up(H.nl, 256), 256, 0, st>>>(H.nl, (const double*)H.rows_l.p, H.cols, 3, 1, gmass.p);
        }
        if (H.nr) {
            sh_col_kernel<<<div_up(H.nr, 256), 256, 0, st>>>(H.nr, (const double*)H.rows_r.p, H.cols, 0, 3, gpos.p + 3 * H.nl);
            sh_col_kernel<<<div_up(H.nr, 256), 256, 0, st>>>(H.nr, (const double*)H.rows_r.p, H.cols, 3, 1, gmass.p + H.nl);
        }
        nbk_tree* halo = nullptr;
        try {
            halo = make_tree(s, gpos.p, nullptr, gmass.p, g, 8, nullptr, store);
            nbk_info a, b;
            nbk_get_info(tree, &a); nbk_get_info(halo, &b);
            if (a.store_bytes != b.store_bytes) {                 // one side needs fp64: both get it
                if (a.store_bytes == 4) { nbk_destroy(tree); tree = nullptr; tree = make_tree(s, s->pos, nullptr, s->mass, s->n, s->real_bytes, nullptr, NBK_STORE_F64); }
                else { nbk_destroy(halo); halo = nullptr; halo = make_tree(s, gpos.p, nullptr, gmass.p, g, 8, nullptr, NBK_STORE_F64); }
            }
        } catch (...) { if (tree) nbk_destroy(tree); if (halo) nbk_destroy(halo); throw; }
        const int rc = nbk_attach_halo(tree, halo);
        if (rc != NBK_OK) { nbk_destroy(tree); nbk_destroy(halo); throw Error(rc, nbk_last_error()); }
    }
    s->dens_tree = tree;
    s->dens_all = s->n + g; s->dens_nl = H.nl; s->dens_nr = H.nr; s->dens_h = s->h_knn;
    s->dens_send_l = std::move(H.send_l); s->dens_send_r = std::move(H.send_r);
    s->ghosts_knn = g;
    s->density_setups++;
}

template <class R>
void calc_density(nbk_sharded* s, int k, double* rho_out, int flags) {
    nbk_comm* c = s->comm;
    cudaStream_t st = c->stream;
    const int W = c->nranks;
    const int64_t n = s->n;
    DevBuf<double> rho, hsm;
    for (int attempt = 0;; attempt++) {
        if (!s->dens_tree) density_setup<R>(s);
        rho.alloc((size_t)s->dens_all); hsm.alloc((size_t)s->dens_all);
        NBK_CHECK(cudaStreamSynchronize(st));
        const int rc = nbk_calc_density(s->dens_tree, k, rho.p, hsm.p, NBK_DEVICE_PTRS);
        if (rc != NBK_OK) throw Error(rc, nbk_last_error());
        nbk_info info;
        nbk_get_info(s->dens_tree, &info);
        s->last_kernel_ms = info.last_kernel_ms; s->last_call_ms = info.last_call_ms; s->last_launches = info.last_launches; s->last_flagged = info.last_flagged;
        if (W == 1) break;
        DevBuf<unsigned long long> bad(1);
        NBK_CHECK(cudaMemsetAsync(bad.p, 0, sizeof(unsigned long long), st));
        sh_halo_check_kernel<R><<<div_up(n, 256), 256, 0, st>>>(n, (const R*)s->pos, hsm.p, s->x0, s->x1, s->dens_h, c->rank > 0, c->rank < W - 1, bad.p);
        NBK_NCCL(ncclAllReduce(bad.p, bad.p, 1, ncclUint64, ncclSum, c->nccl, st));
        unsigned long long hb = 0;
        NBK_CHECK(cudaMemcpyAsync(&hb, bad.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        if (hb == 0) break;
        NBK_REQUIRE(attempt < 6, NBK_ERR_ARG, "nbk_sharded_calc_density: halo still too narrow after 6 widenings");
        s->h_knn = group_max(c, s->h_knn * 1.6);
        close_density(s);
    }
    // scatter terms deposited on ghosts go home: what comes back from my left neighbour concerns the particles I sent left
    if (W > 1) {
        DevBuf<unsigned char> back_l, back_r;
        face_exchange(c, rho.p + n, s->dens_nl * 8, rho.p + n + s->dens_nl, s->dens_nr * 8, back_l, back_r);
        const int64_t bl = (int64_t)back_l.n / 8, br = (int64_t)back_r.n / 8;
        NBK_REQUIRE(bl == (int64_t)s->dens_send_l.n && br == (int64_t)s->dens_send_r.n, NBK_ERR_CUDA, "internal: halo return sizes do not match the lists sent");
        if (bl) sh_index_add_kernel<<<div_up(bl, 256), 256, 0, st>>>(bl, s->dens_send_l.p, (const double*)back_l.p, rho.p);
        if (br) sh_index_add_kernel<<<div_up(br, 256), 256, 0, st>>>(br, s->dens_send_r.p, (const double*)back_r.p, rho.p);
        NBK_CHECK(cudaGetLastError());
    }
    NBK_CHECK(cudaMemcpyAsync(rho_out, rho.p, sizeof(double) * n, (flags & NBK_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    NBK_CHECK(cudaStreamSynchronize(st));
}

template <class R>
void fof_setup(nbk_sharded* s, double hw, bool with_vel) {
    if (s->fof_tree && s->fof_h == hw && s->fof_vel == with_vel) return;
    close_fof(s);
    nbk_comm* c = s->comm;
    cudaStream_t st = c->stream;
    check_halo(s, hw, "FOF");
    Halo H;
    halo_exchange<R>(s, hw, s->periodic, with_vel, false, true, H);
    const int64_t n = s->n, g = H.nl + H.nr, n_all = n + g;
    DevBuf<double> grows((size_t)(g * H.cols));
    if (H.nl) NBK_CHECK(cudaMemcpyAsync(grows.p, H.rows_l.p, H.rows_l.n, cudaMemcpyDeviceToDevice, st));
    if (H.nr) NBK_CHECK(cudaMemcpyAsync(grows.p + H.nl * H.cols, H.rows_r.p, H.rows_r.n, cudaMemcpyDeviceToDevice, st));
    // local arrays (doubles: exact for fp32 and fp64 input; the tree picks fp32 storage when every value is fp32-representable)
    DevBuf<double> pos((size_t)(3 * n_all)), vel(with_vel ? (size_t)(3 * n_all) : 0);
    sh_concat_kernel<R><<<div_up(n_all, 256), 256, 0, st>>>(n, g, (const R*)s->pos, grows.p, H.cols, 0, 3, pos.p);
    if (with_vel) sh_concat_kernel<R><<<div_up(n_all, 256), 256, 0, st>>>(n, g, (const R*)s->vel, grows.p, H.cols, 3, 3, vel.p);
    NBK_CHECK(cudaGetLastError());
    // the local tree wraps with the GLOBAL periods: ghosts keep their true coordinates
    s->fof_tree = make_tree(s, pos.p, with_vel ? vel.p : nullptr, nullptr, n_all, 8, s->periodic ? s->box : nullptr);
    s->fof_gid.alloc((size_t)n_all);
    sh_gid_kernel<<<div_up(n, 256), 256, 0, st>>>(n, nullptr, s->gid0, s->fof_gid.p);
    if (H.nl) NBK_CHECK(cudaMemcpyAsync(s->fof_gid.p + n, H.gid_l.p, H.gid_l.n, cudaMemcpyDeviceToDevice, st));
    if (H.nr) NBK_CHECK(cudaMemcpyAsync(s->fof_gid.p + n + H.nl, H.gid_r.p, H.gid_r.n, cudaMemcpyDeviceToDevice, st));
    NBK_CHECK(cudaStreamSynchronize(st));
    s->fof_h = hw; s->fof_vel = with_vel;
    s->fof_all = n_all; s->fof_nl = H.nl; s->fof_nr = H.nr;
    s->fof_send_l = std::move(H.send_l); s->fof_send_r = std::move(H.send_r);
    s->ghosts_fof = g;
    s->fof_setups++;
}

template <class R>
void run_fof(nbk_sharded* s, int criterion, double fdist, const double* params, int minnum, int order, int32_t* group_out, int64_t* ngroups, int flags) {
    nbk_comm* c = s->comm;
    cudaStream_t st = c->stream;
    const int W = c->nranks;
    const bool with_vel = criterion == NBK_FOF6D;
    NBK_REQUIRE(!with_vel || s->has_vel, NBK_ERR_ARG, "FOF6d needs velocities");
    const double reach = criterion < 0 ? fdist : std::sqrt(params[6]);
    fof_setup<R>(s, reach * (1.0 + 1e-9) + 1e-300, with_vel);
    const int64_t n = s->n, n_all = s->fof_all, nl = s->fof_nl, nr = s->fof_nr, g = nl + nr;
    const int64_t ml = (int64_t)s->fof_send_l.n, mr = (int64_t)s->fof_send_r.n, msent = ml + mr;
    DevBuf<int32_t> root((size_t)n_all);
    {
        NBK_CHECK(cudaStreamSynchronize(st));
        const int rc = nbk_fof_roots(s->fof_tree, criterion, fdist, params, nullptr, root.p, NBK_DEVICE_PTRS);
        if (rc != NBK_OK) throw Error(rc, nbk_last_error());
        nbk_info info;
        nbk_get_info(s->fof_tree, &info);
        s->last_kernel_ms = info.last_kernel_ms; s->last_call_ms = info.last_call_ms; s->last_launches = info.last_launches; s->last_flagged = 0;
    }
    DevBuf<int32_t> cnt((size_t)n_all);
    NBK_CHECK(cudaMemsetAsync(cnt.p, 0, cnt.bytes(), st));
    sh_count_roots_kernel<<<div_up(n, 256), 256, 0, st>>>(n, root.p, cnt.p);            // owned members only: ghosts count at home
    // ---- cross-slab edges ----------------------------------------------------------------------------------------------------
    DevBuf<int64_t> ghost_name((size_t)g), mine((size_t)msent), peer((size_t)msent);
    if (g) sh_names_kernel<<<div_up(g, 256), 256, 0, st>>>(g, nullptr, n, root.p, s->fof_gid.p, ghost_name.p);
    if (ml) sh_names_kernel<<<div_up(ml, 256), 256, 0, st>>>(ml, s->fof_send_l.p, 0, root.p, s->fof_gid.p, mine.p);
    if (mr) sh_names_kernel<<<div_up(mr, 256), 256, 0, st>>>(mr, s->fof_send_r.p, 0, root.p, s->fof_gid.p, mine.p + ml);
    if (W > 1) {
        DevBuf<unsigned char> back_l, back_r;
        face_exchange(c, ghost_name.p, nl * 8, ghost_name.p + nl, nr * 8, back_l, back_r);
        NBK_REQUIRE((int64_t)back_l.n == ml * 8 && (int64_t)back_r.n == mr * 8, NBK_ERR_CUDA, "internal: halo return sizes do not match the lists sent");
        if (ml) NBK_CHECK(cudaMemcpyAsync(peer.p, back_l.p, back_l.n, cudaMemcpyDeviceToDevice, st));
        if (mr) NBK_CHECK(cudaMemcpyAsync(peer.p + ml, back_r.p, back_r.n, cudaMemcpyDeviceToDevice, st));
    }
    DevBuf<int64_t> edges;
    int64_t nedge = 0;
    if (msent) {
        DevBuf<uint32_t> ef(msent + 1), es(msent + 1), scratch(scan_scratch_elems(msent + 1));
        sh_edge_kernel<<<div_up(msent, 256), 256, 0, st>>>(msent, mine.p, peer.p, ef.p);
        NBK_CHECK(cudaMemsetAsync(ef.p + msent, 0, sizeof(uint32_t), st));
        exclusive_scan_u32(ef.p, es.p, msent + 1, scratch.p, st);
        uint32_t tot = 0;
        NBK_CHECK(cudaMemcpyAsync(&tot, es.p + msent, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        nedge = tot;
        edges.alloc((size_t)(2 * nedge));
        if (nedge) sh_edge_compact_kernel<<<div_up(msent, 256), 256, 0, st>>>(msent, ef.p, es.p, mine.p, peer.p, edges.p);
    }
    // the local components that touch the boundary (hold a sent particle or a ghost), with their owned sizes
    DevBuf<uint32_t> touch((size_t)n_all + 1);
    NBK_CHECK(cudaMemsetAsync(touch.p, 0, touch.bytes(), st));
    if (ml) sh_touch_kernel<<<div_up(ml, 256), 256, 0, st>>>(ml, s->fof_send_l.p, 0, root.p, touch.p);
    if (mr) sh_touch_kernel<<<div_up(mr, 256), 256, 0, st>>>(mr, s->fof_send_r.p, 0, root.p, touch.p);
    if (g) sh_touch_kernel<<<div_up(g, 256), 256, 0, st>>>(g, nullptr, n, root.p, touch.p);
    DevBuf<int32_t> tidx;
    const int64_t nt = compact_flags(st, n_all, touch.p, tidx);
    DevBuf<int64_t> node_tab((size_t)(2 * nt));
    if (nt) sh_node_table_kernel<<<div_up(nt, 256), 256, 0, st>>>(nt, tidx.p, s->fof_gid.p, cnt.p, node_tab.p);
    DevBuf<int64_t> all_edges, all_nodes;
    int64_t ne_all = 0, nn_all = 0;
    allgather_rows(c, edges.p, nedge, 2, all_edges, &ne_all);
    allgather_rows(c, node_tab.p, nt, 2, all_nodes, &nn_all);
    // ---- the same union on every rank ---------------------------------------------------------------------------------------
    const int64_t nkeys = nn_all + 2 * ne_all;
    int nn = 0;
    DevBuf<int64_t> names;
    DevBuf<int32_t> comp;
    DevBuf<unsigned long long> csize;
    if (nkeys > 0) {
        DevBuf<uint64_t> ka((size_t)nkeys), kb((size_t)nkeys);
        DevBuf<uint32_t> va((size_t)nkeys), vb((size_t)nkeys);
        // node names (column 0 of the node table) and both columns of the edge table
        NBK_CHECK(cudaMemcpy2DAsync(ka.p, 8, all_nodes.p, 16, 8, (size_t)nn_all, cudaMemcpyDeviceToDevice, st));
        if (ne_all) NBK_CHECK(cudaMemcpyAsync(ka.p + nn_all, all_edges.p, sizeof(int64_t) * 2 * ne_all, cudaMemcpyDeviceToDevice, st));
        int key_bits = 8;
        while (key_bits < 64 && (s->n_global >> key_bits)) key_bits += 8;             // names are global particle ids
        RadixSortPlan<uint64_t> plan(nkeys);
        DevBuf<uint32_t> temp(plan.temp_u32());
        uint64_t* rk; uint32_t* rv;
        radix_sort_pairs<uint64_t>(ka.p, va.p, kb.p, vb.p, nkeys, key_bits, true, temp.p, st, &rk, &rv);
        DevBuf<uint32_t> uf((size_t)nkeys + 1), us((size_t)nkeys + 1), scratch(scan_scratch_elems(nkeys + 1));
        sh_unique_flag_kernel<<<div_up(nkeys, 256), 256, 0, st>>>(nkeys, rk, uf.p);
        NBK_CHECK(cudaMemsetAsync(uf.p + nkeys, 0, sizeof(uint32_t), st));
        exclusive_scan_u32(uf.p, us.p, nkeys + 1, scratch.p, st);
        uint32_t tot = 0;
        NBK_CHECK(cudaMemcpyAsync(&tot, us.p + nkeys, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        nn = (int)tot;
        names.alloc((size_t)nn);
        sh_unique_compact_kernel<<<div_up(nkeys, 256), 256, 0, st>>>(nkeys, rk, uf.p, us.p, names.p);
        DevBuf<int32_t> ea((size_t)ne_all), eb((size_t)ne_all);
        if (ne_all) {
            sh_lookup_kernel<<<div_up(ne_all, 256), 256, 0, st>>>(ne_all, all_edges.p, 2, names.p, nn, ea.p);
            sh_lookup_kernel<<<div_up(ne_all, 256), 256, 0, st>>>(ne_all, all_edges.p + 1, 2, names.p, nn, eb.p);
        }
        comp.alloc((size_t)nn);
        launch_union_pairs(st, nn, ne_all, ea.p, eb.p, comp.p);
        csize.alloc((size_t)nn);
        NBK_CHECK(cudaMemsetAsync(csize.p, 0, csize.bytes(), st));
        if (nn_all) sh_comp_size_kernel<<<div_up(nn_all, 256), 256, 0, st>>>(nn_all, all_nodes.p, names.p, nn, comp.p, csize.p);
    }
    // valid boundary components (replicated) and this rank's interior components
    DevBuf<int32_t> vcomp, iroots;
    int64_t nb = 0, ni = 0;
    if (nn) {
        DevBuf<uint32_t> vf((size_t)nn + 1);
        sh_valid_comp_kernel<<<div_up(nn, 256), 256, 0, st>>>(nn, comp.p, csize.p, minnum, vf.p);
        nb = compact_flags(st, nn, vf.p, vcomp);
    }
    {
        DevBuf<uint32_t> inf((size_t)n_all + 1);
        sh_interior_flag_kernel<<<div_up(n_all, 256), 256, 0, st>>>(n_all, root.p, touch.p, cnt.p, minnum, inf.p);
        ni = compact_flags(st, n_all, inf.p, iroots);
    }
    DevBuf<int64_t> isz((size_t)ni), bsz((size_t)nb), all_isz;
    if (ni) sh_sizes_kernel<<<div_up(ni, 256), 256, 0, st>>>(ni, iroots.p, cnt.p, isz.p);
    if (nb) sh_csize_kernel<<<div_up(nb, 256), 256, 0, st>>>(nb, vcomp.p, csize.p, bsz.p);
    int64_t ni_all = 0;
    allgather_rows(c, isz.p, ni, 1, all_isz, &ni_all);
    std::vector<int64_t> counts = gather_pairs(c, ni, 0);
    // ---- one global numbering (group-table sized: on the host) ---------------------------------------------------------------
    std::vector<int64_t> hsz((size_t)(nb + ni_all));
    if (nb) NBK_CHECK(cudaMemcpyAsync(hsz.data(), bsz.p, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st));
    if (ni_all) NBK_CHECK(cudaMemcpyAsync(hsz.data() + nb, all_isz.p, sizeof(int64_t) * ni_all, cudaMemcpyDeviceToHost, st));
    NBK_CHECK(cudaStreamSynchronize(st));
    const int64_t ng = nb + ni_all;
    std::vector<int64_t> perm((size_t)ng);
    std::iota(perm.begin(), perm.end(), (int64_t)0);             // natural order: boundary components first, then interior by (rank, representative)
    if (order) std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return hsz[(size_t)a] > hsz[(size_t)b]; });
    std::vector<int32_t> gid_of((size_t)ng);
    for (int64_t i = 0; i < ng; i++) gid_of[(size_t)perm[(size_t)i]] = (int32_t)(i + 1);
    int64_t my0 = nb;
    for (int r = 0; r < c->rank; r++) my0 += counts[2 * (size_t)r];
    // ---- labels of the owned particles ----------------------------------------------------------------------------------------
    DevBuf<int32_t> lut((size_t)n_all);
    NBK_CHECK(cudaMemsetAsync(lut.p, 0, lut.bytes(), st));
    if (ni) {
        DevBuf<int32_t> ids((size_t)ni);
        NBK_CHECK(cudaMemcpyAsync(ids.p, gid_of.data() + my0, sizeof(int32_t) * ni, cudaMemcpyHostToDevice, st));
        sh_scatter_i32_kernel<<<div_up(ni, 256), 256, 0, st>>>(ni, iroots.p, ids.p, lut.p);
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    if (nn && nt) {
        DevBuf<int32_t> comp_gid((size_t)nn);
        NBK_CHECK(cudaMemsetAsync(comp_gid.p, 0, comp_gid.bytes(), st));
        if (nb) {
            DevBuf<int32_t> ids((size_t)nb);
            NBK_CHECK(cudaMemcpyAsync(ids.p, gid_of.data(), sizeof(int32_t) * nb, cudaMemcpyHostToDevice, st));
            sh_scatter_i32_kernel<<<div_up(nb, 256), 256, 0, st>>>(nb, vcomp.p, ids.p, comp_gid.p);
            NBK_CHECK(cudaStreamSynchronize(st));
        }
        sh_touch_lut_kernel<<<div_up(nt, 256), 256, 0, st>>>(nt, tidx.p, s->fof_gid.p, names.p, nn, comp.p, comp_gid.p, lut.p);
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    DevBuf<int32_t> lab((size_t)n);
    sh_label_kernel<<<div_up(n, 256), 256, 0, st>>>(n, root.p, lut.p, lab.p);
    NBK_CHECK(cudaGetLastError());
    NBK_CHECK(cudaMemcpyAsync(group_out, lab.p, sizeof(int32_t) * n, (flags & NBK_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    NBK_CHECK(cudaStreamSynchronize(st));
    *ngroups = ng;
}

template <class R>
void stage_local(nbk_sharded* s, const nbk_particles* p) {
    // packed copies of this rank's columns on the device (input: host or device, any stride)
    cudaStream_t st = s->comm->stream;
    const int64_t n = s->n;
    auto stage = [&](const void* src, int64_t stride, int comps) -> void* {
        R* dst = nullptr;
        if (nbk_malloc_async((void**)&dst, sizeof(R) * (size_t)n * comps, st) != cudaSuccess) { cudaGetLastError(); throw Error(NBK_ERR_NOMEM, "device allocation failed"); }
        if (stride == (int64_t)sizeof(R) * comps) {
            NBK_CHECK(cudaMemcpyAsync(dst, src, sizeof(R) * (size_t)n * comps, p->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        } else {
            NBK_CHECK(cudaMemcpy2DAsync(dst, sizeof(R) * comps, src, (size_t)stride, sizeof(R) * comps, (size_t)n, p->on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        }
        return dst;
    };
    s->pos = stage(p->pos, p->pos_stride, 3);
    if (p->vel) { s->vel = stage(p->vel, p->vel_stride, 3); s->has_vel = true; }
    if (p->mass) s->mass = stage(p->mass, p->mass_stride, 1);
    else {
        std::vector<R> ones((size_t)n, (R)1);
        R* dst = nullptr;
        if (nbk_malloc_async((void**)&dst, sizeof(R) * (size_t)n, st) != cudaSuccess) { cudaGetLastError(); throw Error(NBK_ERR_NOMEM, "device allocation failed"); }
        NBK_CHECK(cudaMemcpyAsync(dst, ones.data(), sizeof(R) * (size_t)n, cudaMemcpyHostToDevice, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        s->mass = dst;
    }
    NBK_CHECK(cudaStreamSynchronize(st));
}

}  // namespace

#define NBK_SH_BEGIN try {
#define NBK_SH_END                                                                        \
    }                                                                                     \
    catch (const nbk::Error& e) { nbk::set_last_error(e.what()); return e.code; }          \
    catch (const std::bad_alloc&) { nbk::set_last_error("host allocation failed"); return NBK_ERR_NOMEM; } \
    catch (const std::exception& e) { nbk::set_last_error(e.what()); return NBK_ERR_ARG; } \
    return NBK_OK;

extern "C" {

int nbk_comm_unique_id(unsigned char id[128]) {
    NBK_SH_BEGIN
    NBK_REQUIRE(id != nullptr, NBK_ERR_ARG, "nbk_comm_unique_id: null argument");
    static_assert(sizeof(ncclUniqueId) <= 128, "ncclUniqueId must fit the 128-byte id of the C ABI");
    ncclUniqueId u;
    NBK_NCCL(ncclGetUniqueId(&u));
    memset(id, 0, 128);
    memcpy(id, &u, sizeof(u));
    NBK_SH_END
}

int nbk_comm_init_rank(int nranks, int rank, const unsigned char id[128], int device, nbk_comm** out) {
    NBK_SH_BEGIN
    NBK_REQUIRE(out && id && nranks >= 1 && rank >= 0 && rank < nranks, NBK_ERR_ARG, "nbk_comm_init_rank: bad argument");
    *out = nullptr;
    if (device < 0) NBK_CHECK(cudaGetDevice(&device));
    std::unique_ptr<nbk_comm> c(new nbk_comm);
    c->rank = rank; c->nranks = nranks; c->device = device;
    int prev = -1;
    cudaGetDevice(&prev);
    NBK_CHECK(cudaSetDevice(device));
    NBK_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (nranks > 1) {
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        NBK_NCCL(ncclCommInitRank(&c->nccl, nranks, u, rank));
    }
    if (prev >= 0) cudaSetDevice(prev);
    *out = c.release();
    NBK_SH_END
}

int nbk_comm_destroy(nbk_comm* c) {
    NBK_SH_BEGIN
    if (!c) return NBK_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->nccl) ncclCommDestroy(c->nccl);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
    NBK_SH_END
}

int nbk_sharded_create(nbk_comm* c, const nbk_particles* p, int64_t n_local, const double box[3], const double* slab_edges, int periodic, int knn_k, double halo,
                       nbk_sharded** out) {
    NBK_SH_BEGIN
    NBK_REQUIRE(c && p && p->pos && box && out && n_local >= 1, NBK_ERR_ARG, "nbk_sharded_create: bad argument");
    NBK_REQUIRE(p->real_bytes == 4 || p->real_bytes == 8, NBK_ERR_ARG, "nbk_sharded_create: real_bytes must be 4 or 8");
    *out = nullptr;
    CommGuard guard(c);
    std::unique_ptr<nbk_sharded> s(new nbk_sharded);
    s->comm = c; s->n = n_local; s->real_bytes = p->real_bytes; s->periodic = periodic != 0;
    for (int d = 0; d < 3; d++) s->box[d] = box[d];
    if (slab_edges) {
        s->slab_width = box[0];
        for (int r = 0; r < c->nranks; r++) {
            NBK_REQUIRE(slab_edges[r + 1] > slab_edges[r], NBK_ERR_ARG, "nbk_sharded_create: slab_edges must ascend strictly");
            s->slab_width = std::min(s->slab_width, slab_edges[r + 1] - slab_edges[r]);      // the narrowest slab bounds the halo
        }
        NBK_REQUIRE(slab_edges[0] == 0.0 && slab_edges[c->nranks] == box[0], NBK_ERR_ARG, "nbk_sharded_create: slab_edges must run from 0 to box[0]");
        s->x0 = slab_edges[c->rank]; s->x1 = slab_edges[c->rank + 1];
    } else {
        s->slab_width = box[0] / c->nranks;
        s->x0 = box[0] * c->rank / c->nranks;
        s->x1 = box[0] * (c->rank + 1) / c->nranks;
    }
    if (p->real_bytes == 4) stage_local<float>(s.get(), p); else stage_local<double>(s.get(), p);
    std::vector<int64_t> counts = gather_pairs(c, n_local, 0);
    for (int r = 0; r < c->nranks; r++) { if (r < c->rank) s->gid0 += counts[2 * (size_t)r]; s->n_global += counts[2 * (size_t)r]; }
    // halo for the k-NN ball: a few times the radius that holds k particles at the slab's mean density -- the LARGEST such
    // radius over the ranks, so that what a rank receives is what its own completeness test assumes
    const double vol = (s->x1 - s->x0) * box[1] * box[2];
    const double h = halo > 0 ? halo : 2.5 * std::cbrt((double)(knn_k > 0 ? knn_k : 64) * vol / (double)n_local / (4.0 * 3.14159265358979323846 / 3.0));
    s->h_knn = group_max(c, h);
    *out = s.release();
    NBK_SH_END
}

int nbk_sharded_destroy(nbk_sharded* s) {
    NBK_SH_BEGIN
    if (!s) return NBK_OK;
    CommGuard guard(s->comm);
    close_density(s);
    close_fof(s);
    cudaStreamSynchronize(s->comm->stream);
    for (void* b : {s->pos, s->vel, s->mass}) if (b) cudaFreeAsync(b, s->comm->stream);
    cudaStreamSynchronize(s->comm->stream);
    delete s;
    NBK_SH_END
}

int nbk_sharded_calc_density(nbk_sharded* s, int nsmooth, double* rho, int flags) {
    NBK_SH_BEGIN
    NBK_REQUIRE(s && rho && nsmooth >= 1, NBK_ERR_ARG, "nbk_sharded_calc_density: bad argument");
    CommGuard guard(s->comm);
    if (s->real_bytes == 4) calc_density<float>(s, nsmooth, rho, flags); else calc_density<double>(s, nsmooth, rho, flags);
    NBK_SH_END
}

int nbk_sharded_fof(nbk_sharded* s, int criterion, double fdist, const double* params, int minnum, int order, int32_t* group, int64_t* ngroups, int flags) {
    NBK_SH_BEGIN
    NBK_REQUIRE(s && group && ngroups, NBK_ERR_ARG, "nbk_sharded_fof: null argument");
    NBK_REQUIRE(criterion < 0 ? fdist > 0 : (params != nullptr && (criterion == NBK_FOF3D || criterion == NBK_FOF6D)), NBK_ERR_ARG,
                "nbk_sharded_fof: FOF(fdist) needs a positive linking length, FOFCriterion one of FOF3d / FOF6d with its params");
    CommGuard guard(s->comm);
    if (s->real_bytes == 4) run_fof<float>(s, criterion, fdist, params, minnum, order, group, ngroups, flags);
    else run_fof<double>(s, criterion, fdist, params, minnum, order, group, ngroups, flags);
    NBK_SH_END
}

int nbk_sharded_get_info(const nbk_sharded* s, nbk_sharded_info* info) {
    NBK_SH_BEGIN
    NBK_REQUIRE(s && info, NBK_ERR_ARG, "nbk_sharded_get_info: null argument");
    memset(info, 0, sizeof(*info));
    info->n_local = s->n; info->n_global = s->n_global; info->first_global_id = s->gid0;
    info->rank = s->comm->rank; info->nranks = s->comm->nranks;
    info->h_knn = s->h_knn; info->ghosts_knn = s->ghosts_knn; info->ghosts_fof = s->ghosts_fof;
    info->density_setups = s->density_setups; info->fof_setups = s->fof_setups; info->last_kernel_ms = s->last_kernel_ms;
    info->last_call_ms = s->last_call_ms; info->last_launches = s->last_launches; info->last_flagged = s->last_flagged;
    NBK_SH_END
}

int nbk_sharded_release(nbk_sharded* s) {
    NBK_SH_BEGIN
    NBK_REQUIRE(s != nullptr, NBK_ERR_ARG, "nbk_sharded_release: null argument");
    CommGuard guard(s->comm);
    close_density(s);
    close_fof(s);
    NBK_SH_END
}

}  // extern "C"
