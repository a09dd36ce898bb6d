// fof.cu -- friends-of-friends as lock-free union-find over ball-search links.
//
// Replaces the serial breadth-first search of the reference: KDTree::FOF KDFOF.cxx:29-153,
// KDTree::FOFCriterion :157-265, LeafNode::FOFSearchBall KDLeafNode.cxx:527-590, FOFSearchCriterion :591-619,
// SplitNode::FOFSearchBall[Periodic] KDSplitNode.cxx:921-988,1602-1681, criteria FOFFunc.h:30-55.
//
// The reference's groups are the connected components of the strict '<' link relation;
// here every particle searches its ball with the shared warp traversal and each link found is merged
// with atomicCAS hooking (larger root under smaller root, so a component's root is its smallest tree
// index) and path halving; a second pass flattens, counts, filters by minnum and numbers the groups.
// The link predicate is evaluated in fp64 with the reference's operation order, so the relation --
// and therefore the partition -- is bit-identical.
#include "sort_scan.cuh"
#include "traverse.cuh"
#include "tree.h"

namespace nbk {

// Warps per CTA.  A CTA stays resident until its slowest warp is done, and the warps' costs differ by orders of magnitude (a group
// inside a halo core against one in a void): small CTAs hand their slots back sooner.  Measured at 512^3: 3D link 51.1 ms with 8
// warps, 39.6 ms with 4; 6D link 116.0 / 109.9 / 107.3 ms with 8 / 4 / 2.
constexpr int FOF_WARPS = 4;
constexpr int FOF6_WARPS = 2;      // the general (6D / fp64) link kernel

__device__ __forceinline__ int uf_load(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__device__ __forceinline__ int uf_find(int* parent, int a) {
    while (true) {
        int p = uf_load(parent + a);
        if (p == a) return a;
        int gp = uf_load(parent + p);
        if (gp == p) return p;
        parent[a] = gp;          // path halving: any ancestor is a valid parent (parents only ever decrease)
        a = gp;
    }
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }
        if (atomicCAS(parent + b, b, a) == b) return;
    }
}

struct FofParams {
    const NodeLo* nlo; const NodeHi* nhi; int bucket;
    const void* P; const void* V;
    int64_t n;
    int mode; double p0, p1; float prune_f;
    int periodic; double period[3];
    const int32_t* excl;     // tree order, non-zero => particle takes no part
    int* parent;
};

template <class S>
struct FofVisitor {
    const Vec4<S>* P; const Vec4<S>* V;
    double* tile;          // [6][32]
    int* parent;
    const int32_t* excl;
    double qx, qy, qz, vx, vy, vz;
    double p0, p1;
    float prune_f;
    int mode, self;
    bool on, shifted;
    unsigned lane;

    __device__ __forceinline__ bool need(float lb) const { return lb < prune_f; }

    __device__ __forceinline__ bool linked(int j) const { return crit_linked(mode, p0, p1, qx, qy, qz, vx, vy, vz, tile, j); }

    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[start + base + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
                if (mode == 1 || mode == 4) {
                    Vec4<S> u = V[start + base + lane];
                    tile[96 + lane] = (double)u.x; tile[128 + lane] = (double)u.y; tile[160 + lane] = (double)u.z;
                }
            }
            __syncwarp();
            if (!on) continue;
            for (int j = 0; j < m; j++) {
                int c = start + base + j;
                // the unshifted relation is exactly symmetric, so each pair is merged once (from its lower index)
                if (shifted ? (c == self) : (c <= self)) continue;
                if (linked(j)) {
                    if (excl && excl[c]) continue;
                    uf_union(parent, self, c);
                }
            }
        }
    }
};

template <class S>
__global__ void __launch_bounds__(FOF6_WARPS * 32) fof_link_kernel(FofParams prm) {
    __shared__ double s_tile[FOF6_WARPS][192];
    __shared__ int s_stack[FOF6_WARPS][TRAV_STACK];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    int64_t group = (int64_t)blockIdx.x * FOF6_WARPS + w;
    int64_t qi = group * 32 + lane;
    if (group * 32 >= prm.n) return;
    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
    bool valid = qi < prm.n;
    if (valid && prm.excl && prm.excl[qi]) valid = false;
    FofVisitor<S> v;
    v.P = P; v.V = V; v.tile = s_tile[w]; v.parent = prm.parent; v.excl = prm.excl;
    v.p0 = prm.p0; v.p1 = prm.p1; v.prune_f = prm.prune_f; v.mode = prm.mode; v.lane = lane;
    v.self = valid ? (int)qi : -1;
    v.on = valid; v.shifted = false;
    double x0 = 0, y0 = 0, z0 = 0;
    v.vx = v.vy = v.vz = 0;
    if (valid) {
        Vec4<S> c = P[qi];
        x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z;
        if (prm.mode == 1 || prm.mode == 4) { Vec4<S> u = V[qi]; v.vx = (double)u.x; v.vy = (double)u.y; v.vz = (double)u.z; }
    }
    const int nimg = prm.periodic ? 8 : 1;
    for (int img = 0; img < nimg; img++) {
        // reflected target positions (DistFunc.h:343-355): +p if x < p/2 else -p.  An image whose ball misses the
        // root box dies at the first node test, which subsumes the reference's fdist2 > sval*sval tests.
        v.qx = (img & 1) ? ((x0 < prm.period[0] / 2.0) ? x0 + prm.period[0] : x0 - prm.period[0]) : x0;
        v.qy = (img & 2) ? ((y0 < prm.period[1] / 2.0) ? y0 + prm.period[1] : y0 - prm.period[1]) : y0;
        v.qz = (img & 4) ? ((z0 < prm.period[2] / 2.0) ? z0 + prm.period[2] : z0 - prm.period[2]) : z0;
        v.shifted = img != 0;
        QueryBox qb = make_qbox(v.qx, v.qy, v.qz);
        traverse(prm.nlo, prm.nhi, prm.bucket, s_stack[w], v, qb, valid);
    }
}

// ---- 3D ball on fp32 storage: fp32 screen, exact fp64 test for the survivors ------------------------------------------------
// Stored coordinates are exact fp32 values, so for the unshifted query the fp32 evaluation of d2 (3 subtractions, one
// multiplication, two FMAs) has a relative error below 4 * 2^-24: a candidate whose fp32 d2 exceeds fdist2 * (1 + 2^-20)
// cannot pass the exact test.  Only the few candidates near or inside the ball get the reference's fp64 expression, so the
// link relation -- and the partition -- stay bit-identical.  A periodic image of the query (x +- period) is not an fp32
// value in general: shifted images skip the screen (almost all of them die at the root box test anyway).
struct FofVisitor3F {
    const Vec4<float>* P;
    float4* tile;          // [32]
    int* parent;
    const int32_t* excl;
    double qx, qy, qz, p0;
    float qxf, qyf, qzf, limf, prune_f;
    int self;
    bool on, shifted;
    unsigned lane;

    __device__ __forceinline__ bool need(float lb) const { return lb < prune_f; }
    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        for (int base = 0; base < cnt; base += 32) {
            const int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) { Vec4<float> c = P[start + base + lane]; tile[lane] = make_float4(c.x, c.y, c.z, 0.f); }
            __syncwarp();
            if (!on) continue;
            for (int j = 0; j < m; j++) {
                const int c = start + base + j;
                // the unshifted relation is exactly symmetric, so each pair is merged once (from its lower index)
                if (shifted ? (c == self) : (c <= self)) continue;
                const float4 t = tile[j];
                if (!shifted) {
                    const float dx = qxf - t.x, dy = qyf - t.y, dz = qzf - t.z;
                    if (!(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))) <= limf)) continue;
                }
                if (dist2_ref(qx, qy, qz, (double)t.x, (double)t.y, (double)t.z) < p0) {
                    if (excl && excl[c]) continue;
                    uf_union(parent, self, c);
                }
            }
        }
    }
};

__global__ void __launch_bounds__(FOF_WARPS * 32) fof_link3f_kernel(FofParams prm) {
    __shared__ float4 s_tile[FOF_WARPS][32];
    __shared__ int s_stack[FOF_WARPS][TRAV_STACK];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int64_t group = (int64_t)blockIdx.x * FOF_WARPS + w;
    const int64_t qi = group * 32 + lane;
    if (group * 32 >= prm.n) return;
    const Vec4<float>* P = reinterpret_cast<const Vec4<float>*>(prm.P);
    bool valid = qi < prm.n;
    if (valid && prm.excl && prm.excl[qi]) valid = false;
    FofVisitor3F v;
    v.P = P; v.tile = s_tile[w]; v.parent = prm.parent; v.excl = prm.excl;
    v.p0 = prm.p0; v.prune_f = prm.prune_f; v.lane = lane;
    v.limf = __fmul_ru(__double2float_ru(prm.p0), 1.00000095367431640625f);
    v.self = valid ? (int)qi : -1;
    v.on = valid; v.shifted = false;
    double x0 = 0, y0 = 0, z0 = 0;
    v.qxf = v.qyf = v.qzf = 0.f;
    if (valid) {
        Vec4<float> c = P[qi];
        x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z;
        v.qxf = c.x; v.qyf = c.y; v.qzf = c.z;
    }
    const int nimg = prm.periodic ? 8 : 1;
    for (int img = 0; img < nimg; img++) {
        v.qx = (img & 1) ? ((x0 < prm.period[0] / 2.0) ? x0 + prm.period[0] : x0 - prm.period[0]) : x0;
        v.qy = (img & 2) ? ((y0 < prm.period[1] / 2.0) ? y0 + prm.period[1] : y0 - prm.period[1]) : y0;
        v.qz = (img & 4) ? ((z0 < prm.period[2] / 2.0) ? z0 + prm.period[2] : z0 - prm.period[2]) : z0;
        v.shifted = img != 0;
        QueryBox qb = make_qbox(v.qx, v.qy, v.qz);
        traverse(prm.nlo, prm.nhi, prm.bucket, s_stack[w], v, qb, valid);
    }
}

// ---- FOFCriterionSetBasisForLinks (KDFOF.cxx:268-378, KDLeafNode.cxx:620-652) -------------------------------------
// Only particles whose check value is 0 may start or extend a group ("basis" particles); the others can be linked INTO
// a group by a basis particle but never link further.  The basis particles' groups are therefore the connected components
// of the basis-basis links (the ordinary link pass with the others excluded), and every other particle joins one of the
// groups that reach it.  The reference hands it to the group its serial search discovers first, i.e. the one whose first
// basis member comes first in tree order; here that is the smallest root among the linked basis neighbours (a root is its
// component's smallest tree index).
template <class S>
struct AttachVisitor : FofVisitor<S> {
    int best;
    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)this->lane < m) {
                Vec4<S> c = this->P[start + base + this->lane];
                this->tile[this->lane] = (double)c.x; this->tile[32 + this->lane] = (double)c.y; this->tile[64 + this->lane] = (double)c.z;
                if (this->mode == 1 || this->mode == 4) {
                    Vec4<S> u = this->V[start + base + this->lane];
                    this->tile[96 + this->lane] = (double)u.x; this->tile[128 + this->lane] = (double)u.y; this->tile[160 + this->lane] = (double)u.z;
                }
            }
            __syncwarp();
            if (!this->on) continue;
            for (int j = 0; j < m; j++) {
                int c = start + base + j;
                if (c == this->self || this->excl[c]) continue;           // only basis particles hand out membership
                if (this->linked(j)) {
                    int r = c;
                    while (true) { int p = this->parent[r]; if (p == r) break; r = p; }
                    best = min(best, r);
                }
            }
        }
    }
};

template <class S>
__global__ void __launch_bounds__(FOF_WARPS * 32) fof_attach_kernel(FofParams prm, int* __restrict__ attach_to) {
    __shared__ double s_tile[FOF_WARPS][192];
    __shared__ int s_stack[FOF_WARPS][TRAV_STACK];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    int64_t group = (int64_t)blockIdx.x * FOF_WARPS + w;
    int64_t qi = group * 32 + lane;
    if (group * 32 >= prm.n) return;
    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
    const bool valid = qi < prm.n && prm.excl[qi] != 0;                  // queries: the particles that are NOT a basis
    if (!__any_sync(0xffffffffu, valid)) return;
    AttachVisitor<S> v;
    v.P = P; v.V = V; v.tile = s_tile[w]; v.parent = prm.parent; v.excl = prm.excl;
    v.p0 = prm.p0; v.p1 = prm.p1; v.prune_f = prm.prune_f; v.mode = prm.mode; v.lane = lane;
    v.self = valid ? (int)qi : -1;
    v.on = valid; v.shifted = false; v.best = 0x7fffffff;
    double x0 = 0, y0 = 0, z0 = 0;
    v.vx = v.vy = v.vz = 0;
    if (valid) {
        Vec4<S> c = P[qi];
        x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z;
        if (prm.mode == 1 || prm.mode == 4) { Vec4<S> u = V[qi]; v.vx = (double)u.x; v.vy = (double)u.y; v.vz = (double)u.z; }
    }
    const int nimg = prm.periodic ? 8 : 1;
    for (int img = 0; img < nimg; img++) {
        v.qx = (img & 1) ? ((x0 < prm.period[0] / 2.0) ? x0 + prm.period[0] : x0 - prm.period[0]) : x0;
        v.qy = (img & 2) ? ((y0 < prm.period[1] / 2.0) ? y0 + prm.period[1] : y0 - prm.period[1]) : y0;
        v.qz = (img & 4) ? ((z0 < prm.period[2] / 2.0) ? z0 + prm.period[2] : z0 - prm.period[2]) : z0;
        QueryBox qb = make_qbox(v.qx, v.qy, v.qz);
        traverse(prm.nlo, prm.nhi, prm.bucket, s_stack[w], v, qb, valid);
    }
    if (valid) attach_to[qi] = v.best == 0x7fffffff ? -1 : v.best;
}
// attached particles become children of the root they joined and stop being excluded
__global__ void fof_attach_apply_kernel(int64_t n, const int32_t* excl, const int* attach_to, int* parent, int32_t* excl_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t x = excl[i];
    if (x != 0 && attach_to[i] >= 0) { parent[i] = attach_to[i]; x = 0; }
    excl_out[i] = x;
}

// root (tree index) of every particle, with path halving; excluded particles get -1
__global__ void fof_roots_kernel(int64_t n, int* parent, const int32_t* excl, int32_t* root) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    root[i] = (excl && excl[i]) ? -1 : uf_find(parent, (int)i);
}
// connected components of an explicit edge list (nbk_union_pairs): the cross-slab merge of the sharded FOF
__global__ void uf_init_kernel(int64_t n, int* parent) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (int)i;
}
__global__ void uf_pairs_kernel(int64_t m, const int32_t* __restrict__ a, const int32_t* __restrict__ b, int* parent) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) uf_union(parent, a[i], b[i]);
}
__global__ void uf_flatten_kernel(int64_t n, int* parent, int32_t* root) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) root[i] = uf_find(parent, (int)i);
}
void launch_union_pairs(cudaStream_t st, int64_t nnodes, int64_t npairs, const int32_t* a, const int32_t* b, int32_t* root) {
    if (nnodes <= 0) return;
    DevBuf<int> parent(nnodes);
    uf_init_kernel<<<div_up(nnodes, 256), 256, 0, st>>>(nnodes, parent.p);
    if (npairs > 0) uf_pairs_kernel<<<div_up(npairs, 256), 256, 0, st>>>(npairs, a, b, parent.p);
    uf_flatten_kernel<<<div_up(nnodes, 256), 256, 0, st>>>(nnodes, parent.p, root);
    NBK_CHECK(cudaGetLastError());
    NBK_CHECK(cudaStreamSynchronize(st));
}

__global__ void fof_init_kernel(int64_t n, int* parent, uint32_t* size) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = (int)i; size[i] = 0; }
}
__global__ void fof_flatten_kernel(int64_t n, int* parent, uint32_t* size) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = uf_find(parent, (int)i);
    atomicAdd(&size[r], 1u);
}
// after flatten every chain is short; make parent[i] the root itself
__global__ void fof_root_kernel(int64_t n, int* parent, const uint32_t* size, int minnum, const int32_t* excl, uint32_t* flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = (int)i;
    while (true) { int p = parent[r]; if (p == r) break; r = p; }
    parent[i] = r;   // races only write the same final root
    bool isroot = (r == (int)i);
    flag[i] = (isroot && size[i] >= (uint32_t)minnum && !(excl && excl[i])) ? 1u : 0u;
}
// flagscan = exclusive scan of flag.  order==0: id = rank in tree order + 1
__global__ void fof_label_kernel(int64_t n, const int* parent, const uint32_t* flag, const uint32_t* flagscan, const uint32_t* newid, int32_t* group) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = parent[i];
    int g = 0;
    if (flag[r]) g = newid ? (int)newid[flagscan[r]] : (int)flagscan[r] + 1;
    group[i] = g;
}
// compact valid roots: keys = ~size (so ascending key == descending size), vals = rank in tree order
__global__ void fof_compact_kernel(int64_t n, const uint32_t* flag, const uint32_t* flagscan, const uint32_t* size, uint32_t* keys, uint32_t* vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    uint32_t r = flagscan[i];
    keys[r] = ~size[i];
    vals[r] = r;
}
__global__ void fof_newid_kernel(int64_t ng, const uint32_t* sorted_vals, uint32_t* newid) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ng) newid[sorted_vals[i]] = (uint32_t)i + 1;
}
// group lengths pLen[gid] (KDFOF.cxx:74,98-104)
__global__ void fof_len_kernel(int64_t n, const uint32_t* flag, const uint32_t* flagscan, const uint32_t* newid, const uint32_t* size, int32_t* len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    int g = newid ? (int)newid[flagscan[i]] : (int)flagscan[i] + 1;
    len[g] = (int)size[i];
}

// ---- pHead / pNext / pTail (KDFOF.cxx:52-67): members of a group chained in ascending tree index -------------------
__global__ void fof_list_keys_kernel(int64_t n, const int32_t* group, uint32_t* keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (uint32_t)group[i];
}
// sorted by group id (stable, so ascending tree index inside a group): chain neighbours, remember each run's ends
__global__ void fof_list_next_kernel(int64_t n, const uint32_t* sk, const uint32_t* sv, int32_t* next, uint32_t* run_first, uint32_t* run_last) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t g = sk[p], i = sv[p];
    if (g == 0) { next[i] = -1; return; }
    bool first = (p == 0) || sk[p - 1] != g;
    bool last = (p == n - 1) || sk[p + 1] != g;
    next[i] = last ? -1 : (int32_t)sv[p + 1];
    if (first) run_first[g] = i;
    if (last) run_last[g] = i;
}
__global__ void fof_list_ends_kernel(int64_t n, const int32_t* group, const uint32_t* run_first, const uint32_t* run_last, int32_t* head, int32_t* tail) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int g = group[i];
    if (head) head[i] = g > 0 ? (int32_t)run_first[g] : (int32_t)i;
    if (tail) tail[i] = g > 0 ? (int32_t)run_last[g] : (int32_t)i;
}

static void build_fof_lists(nbk_tree& t, FofArgs& a, int64_t* launches) {
    const int64_t n = t.n;
    cudaStream_t st = t.stream;
    const int tb = 256;
    DevBuf<uint32_t> ka(n), kb(n), va(n), vb(n), rf((size_t)a.ngroups + 2), rl((size_t)a.ngroups + 2);
    DevBuf<int32_t> next_tmp(a.next ? 0 : n);
    int32_t* next = a.next ? a.next : next_tmp.p;
    RadixSortPlan<uint32_t> plan(n);
    DevBuf<uint32_t> temp(plan.temp_u32());
    fof_list_keys_kernel<<<div_up(n, tb), tb, 0, st>>>(n, a.group_tree, ka.p);
    int bits = 8;
    while (bits < 32 && ((uint64_t)1 << bits) <= (uint64_t)a.ngroups) bits += 8;
    uint32_t *rk, *rv;
    radix_sort_pairs<uint32_t>(ka.p, va.p, kb.p, vb.p, n, bits, true, temp.p, st, &rk, &rv, launches);
    fof_list_next_kernel<<<div_up(n, tb), tb, 0, st>>>(n, rk, rv, next, rf.p, rl.p);
    fof_list_ends_kernel<<<div_up(n, tb), tb, 0, st>>>(n, a.group_tree, rf.p, rl.p, a.head, a.tail);
    *launches += 3;
    NBK_CHECK(cudaStreamSynchronize(st));
}

static int g_fof_screen = 1;
bool set_fof_option(const char* name, int64_t value) {
    if (std::string(name) == "fof_screen") { g_fof_screen = (int)value; return true; }
    return false;
}

void launch_fof(nbk_tree& t, FofArgs& a) {
    const int64_t n = t.n;
    cudaStream_t st = t.stream;
    DevBuf<int> parent(n);
    DevBuf<uint32_t> size(n), flag(n + 1), flagscan(n + 1), scratch(scan_scratch_elems(n + 1));
    const int tb = 256;
    int64_t launches = 0;
    Tracer tr(st);
    fof_init_kernel<<<div_up(n, tb), tb, 0, st>>>(n, parent.p, size.p);
    FofParams p;
    p.nlo = t.nlo; p.nhi = t.nhi; p.bucket = t.bucket;
    p.P = t.pos4(); p.V = t.vel4(); p.n = n;
    p.mode = a.mode; p.p0 = a.p0; p.p1 = a.p1;
    p.prune_f = __builtin_nextafterf((float)a.prune_x2, INFINITY);
    if ((double)p.prune_f < a.prune_x2) p.prune_f = __builtin_nextafterf(p.prune_f, INFINITY);
    p.periodic = t.periodic ? 1 : 0;
    for (int d = 0; d < 3; d++) p.period[d] = t.period[d];
    p.excl = a.precheck_tree; p.parent = parent.p;
    if (a.mode == 1 || a.mode == 4) NBK_REQUIRE(p.V != nullptr, NBK_ERR_ARG, "6D FOF needs velocities");
    int64_t groups = (n + 31) / 32;
    NBK_CHECK(cudaEventRecord(t.ev2, st));
    if (t.store_bytes == 4 && a.mode == 0 && g_fof_screen) fof_link3f_kernel<<<div_up(groups, FOF_WARPS), FOF_WARPS * 32, 0, st>>>(p);
    else if (t.store_bytes == 4) fof_link_kernel<float><<<div_up(groups, FOF6_WARPS), FOF6_WARPS * 32, 0, st>>>(p);
    else fof_link_kernel<double><<<div_up(groups, FOF6_WARPS), FOF6_WARPS * 32, 0, st>>>(p);
    DevBuf<int32_t> excl2;
    const int32_t* excl_final = a.precheck_tree;
    if (a.attach) {
        NBK_REQUIRE(a.precheck_tree != nullptr, NBK_ERR_ARG, "FOFCriterionSetBasisForLinks needs the check values");
        DevBuf<int> attach_to(n);
        excl2.alloc(n);
        if (t.store_bytes == 4) fof_attach_kernel<float><<<div_up(groups, FOF_WARPS), FOF_WARPS * 32, 0, st>>>(p, attach_to.p);
        else fof_attach_kernel<double><<<div_up(groups, FOF_WARPS), FOF_WARPS * 32, 0, st>>>(p, attach_to.p);
        fof_attach_apply_kernel<<<div_up(n, tb), tb, 0, st>>>(n, a.precheck_tree, attach_to.p, parent.p, excl2.p);
        NBK_CHECK(cudaStreamSynchronize(st));
        excl_final = excl2.p;
        launches += 2;
    }
    NBK_CHECK(cudaEventRecord(t.ev3, st));
    tr.point("fof link");
    if (a.roots_tree) {
        // representatives only (nbk_fof_roots): every particle's root = the smallest tree index of its component, -1 for
        // particles that take no part; no minnum filter, no numbering
        fof_roots_kernel<<<div_up(n, tb), tb, 0, st>>>(n, parent.p, excl_final, a.roots_tree);
        NBK_CHECK(cudaStreamSynchronize(st));
        NBK_CHECK(cudaGetLastError());
        float ms = 0;
        NBK_CHECK(cudaEventElapsedTime(&ms, t.ev2, t.ev3));
        t.last_kernel_ms = ms;
        t.last_launches = launches + 3;
        a.ngroups = 0;
        return;
    }
    fof_flatten_kernel<<<div_up(n, tb), tb, 0, st>>>(n, parent.p, size.p);
    fof_root_kernel<<<div_up(n, tb), tb, 0, st>>>(n, parent.p, size.p, a.minnum, excl_final, flag.p);
    NBK_CHECK(cudaMemsetAsync(flag.p + n, 0, sizeof(uint32_t), st));
    exclusive_scan_u32(flag.p, flagscan.p, n + 1, scratch.p, st, &launches);
    uint32_t ng32 = 0;
    NBK_CHECK(cudaMemcpyAsync(&ng32, flagscan.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NBK_CHECK(cudaStreamSynchronize(st));
    a.ngroups = ng32;
    tr.point("fof flatten+root+scan");
    launches += 5;
    DevBuf<uint32_t> newid;
    if (a.order && ng32 > 1) {
        DevBuf<uint32_t> ka(ng32), kb(ng32), va(ng32), vb(ng32);
        RadixSortPlan<uint32_t> plan(ng32);
        DevBuf<uint32_t> temp(plan.temp_u32());
        newid.alloc(ng32);
        fof_compact_kernel<<<div_up(n, tb), tb, 0, st>>>(n, flag.p, flagscan.p, size.p, ka.p, va.p);
        uint32_t *rk, *rv;
        radix_sort_pairs<uint32_t>(ka.p, va.p, kb.p, vb.p, ng32, 32, false, temp.p, st, &rk, &rv, &launches);
        fof_newid_kernel<<<div_up(ng32, tb), tb, 0, st>>>(ng32, rv, newid.p);
        launches += 2;
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    tr.point("fof order");
    fof_label_kernel<<<div_up(n, tb), tb, 0, st>>>(n, parent.p, flag.p, flagscan.p, newid.p, a.group_tree);
    launches++;
    if (a.len) {
        NBK_CHECK(cudaMemsetAsync(a.len, 0, sizeof(int32_t) * (ng32 + 1), st));
        fof_len_kernel<<<div_up(n, tb), tb, 0, st>>>(n, flag.p, flagscan.p, newid.p, size.p, a.len);
        launches++;
    }
    if (a.head || a.next || a.tail) build_fof_lists(t, a, &launches);
    NBK_CHECK(cudaStreamSynchronize(st));
    NBK_CHECK(cudaGetLastError());
    tr.point("fof label");
    float ms = 0;
    NBK_CHECK(cudaEventElapsedTime(&ms, t.ev2, t.ev3));
    t.last_kernel_ms = ms;
    t.last_launches = launches;
}

}  // namespace nbk
