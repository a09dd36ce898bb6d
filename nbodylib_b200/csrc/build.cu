// build.cu -- level-parallel exact-median kd-tree build (replaces the recursive quickselect build of the
// reference: KDTree::BuildNodes KDTree.cxx:970-1055, DetermineSplitDim :459-504, MedianPos :328-370,
// SpreadestPos/BoundaryandMeanPos :27-167).
//
// Same splitting rule as the reference, evaluated for every node of a level at once:
//   * split dimension = largest (max-min) extent, ties -> lowest dimension  (KDTree.cxx:489-499)
//   * split index     = start + (size-1)/2 ; left = [start, split+1), right = [split+1, end)  (:1012-1053)
//   * leaf            = size <= bucket                                                      (:994)
// so the tree SHAPE (node count, leaf count, every node's [start,end)) equals the reference's; only the
// placement of particles with exactly equal split coordinates and the order inside a leaf may differ.
//
// Method (no recursion, no per-node launches):
//   1. one hand-written radix sort per dimension gives three index arrays ord[d], each sorted by coordinate d;
//   2. for every level: a node's extent in d is coord[ord[d][end-1]] - coord[ord[d][start]] (free), the median
//      particle is ord[cd][m]; every particle gets a side bit, and the two other arrays are stable-partitioned
//      per node by that bit with one device-wide prefix sum (the partition target of a node's first element is
//      data independent, so no segmented scan is needed);
//   3. the final ord[0] is the tree order; particles are gathered into tree-order Vec4 arrays.
// All kernels stream HBM with coalesced accesses; the only random accesses are the 1-byte side-bit gathers.
#include "sort_scan.cuh"
#include "tree.h"

namespace nbk {

template <class S>
__device__ __forceinline__ S comp(const Vec4<S>& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

template <class S, class K>
__global__ void make_keys_kernel(const Vec4<S>* __restrict__ P, int64_t n, int d, K* __restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = sort_key(comp(P[i], d));
}

template <class S>
__global__ void finalize_kernel(int64_t n, const uint32_t* __restrict__ ord, const Vec4<S>* __restrict__ prim_in, const Vec4<S>* __restrict__ sec_in,
                                const double* __restrict__ mass_in, Vec4<S>* __restrict__ prim, Vec4<S>* __restrict__ sec, double* __restrict__ mass,
                                int32_t* __restrict__ order) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t pid = ord[i];
    prim[i] = prim_in[pid];
    if (sec_in) sec[i] = sec_in[pid];
    mass[i] = mass_in ? mass_in[pid] : 1.0;
    order[i] = (int32_t)pid;
}

// ================================================================================================ build v2
// Rank-space build.  After the three radix sorts every particle is described by its three RANKS (position in the
// array sorted by x, by y, by z; ties already broken by the stable sort, i.e. by (coordinate, id)).  The level loop works
// on ranks only -- no coordinates, no ties, no gathers:
//   * three arrays A[d] (one per sort order) carry the rank triplet of the particle at each position as SoA columns;
//   * a node's extent in d comes from the sorted keys at the ranks of its first / last element of A[d];
//   * side of an element = (its rank in the cut dimension > the median's rank); A[cd] itself splits by position;
//   * every array is stable-partitioned with one device-wide scan per level (partition offsets are data independent).
// One level streams 12 columns once for the flags (8 B/particle) and 9 columns in + 9 out for the move (72 B/particle),
// all coalesced.  Levels whose nodes hold at most V2_CAP particles are finished by ONE kernel: a CTA loads its node's
// rank triplets, converts them to local ranks (binary search in the node's own sorted columns) and runs the remaining
// levels of the same algorithm in shared memory.
constexpr int V2_CAP = PRIM_TILE;   // 4096: node size handled by one CTA; also >= the tile size of the level kernels
constexpr int V2_T = 512;
constexpr int V2_PER = V2_CAP / V2_T;   // 8 positions per thread

struct Cols { uint32_t* c[3][3]; };   // c[d][e][i]: rank in dimension e of the particle at position i of the array sorted by d

template <class K> struct KeyCoord;
template <> struct KeyCoord<uint32_t> {
    typedef float type;
    static __device__ __forceinline__ float get(uint32_t k) { return __uint_as_float((k >> 31) ? (k ^ 0x80000000u) : ~k); }
};
template <> struct KeyCoord<uint64_t> {
    typedef double type;
    static __device__ __forceinline__ double get(uint64_t k) { return __longlong_as_double((long long)((k >> 63) ? (k ^ 0x8000000000000000ull) : ~k)); }
};

__global__ void v2_rank_scatter_kernel(int64_t n, const uint32_t* __restrict__ o0, const uint32_t* __restrict__ o1, const uint32_t* __restrict__ o2,
                                       uint32_t* __restrict__ R /* [n][4] */) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    R[(size_t)o0[i] * 4 + 0] = (uint32_t)i;
    R[(size_t)o1[i] * 4 + 1] = (uint32_t)i;
    R[(size_t)o2[i] * 4 + 2] = (uint32_t)i;
}
__global__ void v2_rank_gather_kernel(int64_t n, const uint32_t* __restrict__ od, const uint4* __restrict__ R, uint32_t* __restrict__ c0,
                                      uint32_t* __restrict__ c1, uint32_t* __restrict__ c2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = R[od[i]];
    c0[i] = r.x; c1[i] = r.y; c2[i] = r.z;
}

// one thread per node of a GLOBAL level (every node of such a level is split)
template <class K>
__global__ void v2_level_nodes_kernel(int level, int64_t n, int al, Cols cur, const K* __restrict__ sk0, const K* __restrict__ sk1, const K* __restrict__ sk2,
                                      NodeLo* __restrict__ nlo, NodeHi* __restrict__ nhi, int8_t* __restrict__ cutdim, int32_t* __restrict__ lv_cd,
                                      uint32_t* __restrict__ lv_mr, uint32_t* __restrict__ rcount) {
    typedef typename KeyCoord<K>::type S;
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t C = (int64_t)1 << level;
    if (j >= C) return;
    int64_t idx = C - 1 + j;
    int s, e;
    if (level == 0) { s = 0; e = (int)n; }
    else {
        int64_t p = (idx - 1) >> 1;
        int ps = nlo[p].start, pe = nhi[p].end;
        int pm = ps + (int)split_left(pe - ps, al) - 1;
        if (idx & 1) { s = ps; e = pm + 1; } else { s = pm + 1; e = pe; }
    }
    S lo0 = KeyCoord<K>::get(sk0[cur.c[0][0][s]]), hi0 = KeyCoord<K>::get(sk0[cur.c[0][0][e - 1]]);
    S lo1 = KeyCoord<K>::get(sk1[cur.c[1][1][s]]), hi1 = KeyCoord<K>::get(sk1[cur.c[1][1][e - 1]]);
    S lo2 = KeyCoord<K>::get(sk2[cur.c[2][2][s]]), hi2 = KeyCoord<K>::get(sk2[cur.c[2][2][e - 1]]);
    double e0 = (double)hi0 - (double)lo0, e1 = (double)hi1 - (double)lo1, e2 = (double)hi2 - (double)lo2;
    int cd = 0; double best = e0;
    if (e1 > best) { cd = 1; best = e1; }
    if (e2 > best) { cd = 2; }
    NodeLo a; a.x = round_down(lo0); a.y = round_down(lo1); a.z = round_down(lo2); a.start = s;
    NodeHi b; b.x = round_up(hi0); b.y = round_up(hi1); b.z = round_up(hi2); b.end = e;
    nlo[idx] = a; nhi[idx] = b;
    cutdim[idx] = (int8_t)cd;
    const int m = s + (int)split_left(e - s, al) - 1;
    lv_cd[j] = cd;
    lv_mr[j] = cur.c[cd][cd][m];
    rcount[j] = (uint32_t)(e - (m + 1));
}

// the (at most two) nodes a tile of a global level overlaps.  The node holding position `base` is found by walking the
// split rule from the root (pure arithmetic: the tree shape depends only on n), no search through memory.
struct TileNodes { int jA, sB, mA, mB, cdA, cdB; uint32_t mrA, mrB; };
__device__ __forceinline__ TileNodes v2_tile_nodes(int64_t base, int level, int64_t n, int al, const int32_t* __restrict__ lv_cd, const uint32_t* __restrict__ lv_mr) {
    int s = 0, e = (int)n, j = 0;
    for (int l = 0; l < level; l++) {
        const int m = s + (int)split_left(e - s, al) - 1;
        if (base > m) { s = m + 1; j = 2 * j + 1; } else { e = m + 1; j = 2 * j; }
    }
    TileNodes t;
    t.jA = j; t.mA = s + (int)split_left(e - s, al) - 1;
    const int C = 1 << level;
    const bool hasB = j + 1 < C;
    t.cdA = lv_cd[j]; t.mrA = lv_mr[j];
    t.cdB = hasB ? lv_cd[j + 1] : 0; t.mrB = hasB ? lv_mr[j + 1] : 0u;
    t.sB = 0x7fffffff; t.mB = 0;
    if (hasB) {
        // the next node at this level: climb to the first ancestor entered through its left child, step to the sibling,
        // then keep left; equivalently walk the split rule for position e (the first position after node A)
        int s2 = 0, e2 = (int)n;
        for (int l = 0; l < level; l++) {
            const int m2 = s2 + (int)split_left(e2 - s2, al) - 1;
            if (e > m2) s2 = m2 + 1; else e2 = m2 + 1;
        }
        t.sB = s2; t.mB = s2 + (int)split_left(e2 - s2, al) - 1;
    }
    return t;
}

__global__ void __launch_bounds__(PRIM_THREADS) v2_count_kernel(int level, int64_t n, int al, Cols cur, const int32_t* __restrict__ lv_cd,
                                                                const uint32_t* __restrict__ lv_mr, uint32_t* __restrict__ tsum, int ntiles) {
    __shared__ uint32_t swarp[8];
    __shared__ TileNodes tn_s;
    const int d = blockIdx.y;
    const int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
    if (threadIdx.x == 0) tn_s = v2_tile_nodes(base, level, n, al, lv_cd, lv_mr);
    __syncthreads();
    const TileNodes tn = tn_s;
    uint32_t cnt = 0;
#pragma unroll 4
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int64_t i = base + r * PRIM_THREADS + threadIdx.x;
        if (i < n) {
            const bool inB = i >= tn.sB;
            const int cd = inB ? tn.cdB : tn.cdA;
            const int m = inB ? tn.mB : tn.mA;
            bool f;
            if (cd == d) f = i > m;
            else f = cur.c[d][cd][i] > (inB ? tn.mrB : tn.mrA);
            cnt += f ? 1u : 0u;
        }
    }
    uint32_t tot;
    block_excl_scan_256(cnt, swarp, &tot);
    if (threadIdx.x == 0) tsum[(size_t)d * ntiles + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(PRIM_THREADS, 3) v2_scatter_kernel(int level, int64_t n, int al, Cols cur, Cols nxt, const int32_t* __restrict__ lv_cd,
                                                                     const uint32_t* __restrict__ lv_mr, const uint32_t* __restrict__ tscan, int ntiles,
                                                                     const uint32_t* __restrict__ Rb) {
    __shared__ uint32_t cnt[PRIM_ITEMS * 8];
    __shared__ uint32_t swarp[8];
    __shared__ TileNodes tn_s;
    const int d = blockIdx.y;
    const unsigned lane = lane_id(), w = threadIdx.x >> 5, lt = lanemask_lt();
    const int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
    if (threadIdx.x == 0) tn_s = v2_tile_nodes(base, level, n, al, lv_cd, lv_mr);
    __syncthreads();
    const TileNodes tn = tn_s;
    const uint32_t RbA = Rb[tn.jA], RbB = (tn.sB != 0x7fffffff) ? Rb[tn.jA + 1] : 0u;
    uint32_t bal[PRIM_ITEMS];
    uint32_t fbits = 0;
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int64_t i = base + r * PRIM_THREADS + threadIdx.x;
        bool f = false;
        if (i < n) {
            const bool inB = i >= tn.sB;
            const int cd = inB ? tn.cdB : tn.cdA;
            const int m = inB ? tn.mB : tn.mA;
            if (cd == d) f = i > m;
            else f = cur.c[d][cd][i] > (inB ? tn.mrB : tn.mrA);
        }
        uint32_t b = __ballot_sync(0xffffffffu, f);
        bal[r] = b;
        fbits |= (f ? 1u : 0u) << r;
        if (lane == 0) cnt[r * 8 + w] = __popc(b);
    }
    __syncthreads();
    uint32_t v = threadIdx.x < PRIM_ITEMS * 8 ? cnt[threadIdx.x] : 0u;
    uint32_t tot;
    uint32_t ex = block_excl_scan_256(v, swarp, &tot);
    if (threadIdx.x < PRIM_ITEMS * 8) cnt[threadIdx.x] = ex;
    __syncthreads();
    const uint32_t tile_off = tscan[(size_t)d * ntiles + blockIdx.x] - tscan[(size_t)d * ntiles];
    const uint32_t* __restrict__ s0 = cur.c[d][0];
    const uint32_t* __restrict__ s1 = cur.c[d][1];
    const uint32_t* __restrict__ s2 = cur.c[d][2];
    uint32_t* __restrict__ t0 = nxt.c[d][0];
    uint32_t* __restrict__ t1 = nxt.c[d][1];
    uint32_t* __restrict__ t2 = nxt.c[d][2];
    // batches of 8 elements: all 24 loads of a batch are in flight before the first scattered store (the source and target
    // columns never alias, but a store between two loads would otherwise order them)
#pragma unroll
    for (int r0 = 0; r0 < PRIM_ITEMS; r0 += 8) {
        uint32_t v0[8], v1[8], v2[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int64_t i = base + (r0 + q) * PRIM_THREADS + threadIdx.x;
            if (i < n) { v0[q] = s0[i]; v1[q] = s1[i]; v2[q] = s2[i]; }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int r = r0 + q;
            int64_t i = base + r * PRIM_THREADS + threadIdx.x;
            if (i < n) {
                const bool inB = i >= tn.sB;
                const int m = inB ? tn.mB : tn.mA;
                const uint32_t P = tile_off + cnt[r * 8 + w] + __popc(bal[r] & lt);
                const uint32_t R = P - (inB ? RbB : RbA);
                const int64_t dest = ((fbits >> r) & 1u) ? (int64_t)m + 1 + R : i - (int64_t)R;
                t0[dest] = v0[q]; t1[dest] = v1[q]; t2[dest] = v2[q];
            }
        }
    }
}

// shared-memory bytes of the small-node kernel
static inline size_t v2_small_smem(int coord_bytes, int ntab) {
    return (size_t)3 * V2_CAP * 2 + 2 * V2_CAP * 2 + (size_t)3 * V2_CAP * coord_bytes + (size_t)V2_CAP * 4 + (size_t)ntab * 2 * 9 + (size_t)ntab + 16 + 16 * 8;
}

// One CTA finishes one node of at most V2_CAP particles: all remaining levels in shared memory.
//   set-up : element = local x-rank; local ranks in y and z by binary search of the element's global rank in the node's own
//            sorted column; the three local sorted orders; coordinates by element (one gather of the particle record).
//   levels : node records (extent from the first / last element of each order), then ONE pass partitions the three orders:
//            thread t owns positions 8t..8t+7 of every order, the three flag counts are scanned together (packed in 64 bits),
//            prefix values at node starts go through a per-node table.
template <class S>
__global__ void __launch_bounds__(V2_T) v2_small_kernel(int L, int nsub, int64_t n, int bucket, int al, Cols cur, const Vec4<S>* __restrict__ P_in,
                                                        const uint32_t* __restrict__ ordx, NodeLo* __restrict__ nlo, NodeHi* __restrict__ nhi,
                                                        int8_t* __restrict__ cutdim, uint32_t* __restrict__ tree_ord, int ntab) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint16_t* ord0 = reinterpret_cast<uint16_t*>(sm);
    uint16_t* ord1 = ord0 + V2_CAP;
    uint16_t* ord2 = ord1 + V2_CAP;
    uint16_t* rk1 = ord2 + V2_CAP;
    uint16_t* rk2 = rk1 + V2_CAP;
    S* cx = reinterpret_cast<S*>(rk2 + V2_CAP);                        // coordinates by element
    S* cy = cx + V2_CAP;
    S* cz = cy + V2_CAP;
    uint32_t* tmp = reinterpret_cast<uint32_t*>(cz + V2_CAP);           // set-up: one sorted column of global ranks
    uint16_t* pnode = reinterpret_cast<uint16_t*>(tmp);                 // level loop: node of each position (0xffff: in a finished leaf)
    uint16_t* n_s = reinterpret_cast<uint16_t*>(tmp + V2_CAP);          // [2][ntab] ping-pong by sublevel parity
    uint16_t* n_e = n_s + 2 * ntab;                                     // [2][ntab]; 0 = node absent
    uint16_t* n_m = n_e + 2 * ntab;
    uint16_t* n_mr = n_m + ntab;
    uint16_t* n_pre = n_mr + ntab;                                      // [3][ntab] exclusive flag prefix at the node's first position
    int8_t* n_cd = reinterpret_cast<int8_t*>(n_pre + 3 * ntab);         // 0..2 split dimension, 3 = not split (leaf or absent)
    unsigned long long* wsum = reinterpret_cast<unsigned long long*>(sm + ((reinterpret_cast<unsigned char*>(n_cd + ntab) - sm + 15) & ~(size_t)15));
    __shared__ int se_s[2];

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, w = tid >> 5;
    const int64_t jL = blockIdx.x;
    if (tid == 0) {
        int64_t idx = ((int64_t)1 << L) - 1 + jL;
        int s, e;
        if (L == 0) { s = 0; e = (int)n; }
        else {
            int64_t p = (idx - 1) >> 1;
            int ps = nlo[p].start, pe = nhi[p].end;
            int pm = ps + (int)split_left(pe - ps, al) - 1;
            if (idx & 1) { s = ps; e = pm + 1; } else { s = pm + 1; e = pe; }
        }
        se_s[0] = s; se_s[1] = e;
    }
    __syncthreads();
    const int s = se_s[0], e = se_s[1];
    const int S_ = e - s;

    // ---- set-up ----------------------------------------------------------------------------------------------
    for (int i = tid; i < S_; i += V2_T) {
        ord0[i] = (uint16_t)i;
        const Vec4<S> p = P_in[ordx[cur.c[0][0][s + i]]];
        cx[i] = p.x; cy[i] = p.y; cz[i] = p.z;
    }
    for (int dd = 1; dd <= 2; dd++) {
        const uint32_t* colsorted = cur.c[dd][dd] + s;       // global ranks in dimension dd of this node, ascending
        const uint32_t* mine = cur.c[0][dd] + s;             // global rank in dd of element i
        uint16_t* rk = dd == 1 ? rk1 : rk2;
        uint16_t* od = dd == 1 ? ord1 : ord2;
        __syncthreads();
        for (int i = tid; i < S_; i += V2_T) tmp[i] = colsorted[i];
        __syncthreads();
        for (int i = tid; i < S_; i += V2_T) {
            const uint32_t g = mine[i];
            int lo = 0, hi = S_ - 1;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (tmp[mid] < g) lo = mid + 1; else hi = mid; }
            rk[i] = (uint16_t)lo;
            od[lo] = (uint16_t)i;
        }
    }
    __syncthreads();
    for (int i = tid; i < V2_CAP; i += V2_T) pnode[i] = 0;
    if (tid == 0) { n_s[0] = 0; n_e[0] = (uint16_t)S_; }
    __syncthreads();

    for (int t = 0; t <= nsub; t++) {
        const int nn = 1 << t, tb = t & 1;
        uint16_t* cs_ = n_s + tb * ntab; uint16_t* ce_ = n_e + tb * ntab;
        uint16_t* ns_ = n_s + (tb ^ 1) * ntab; uint16_t* ne_ = n_e + (tb ^ 1) * ntab;
        // ---- nodes of this sublevel: record, split decision, children ranges ------------------------------------
        for (int k = tid; k < nn; k += V2_T) {
            const int s1 = cs_[k], e1 = ce_[k];
            bool split = false;
            if (e1 != 0) {
                const int64_t hidx = (((int64_t)1 << (L + t)) - 1) + (jL << t) + k;
                S lo0 = cx[ord0[s1]], hi0 = cx[ord0[e1 - 1]];
                S lo1 = cy[ord1[s1]], hi1 = cy[ord1[e1 - 1]];
                S lo2 = cz[ord2[s1]], hi2 = cz[ord2[e1 - 1]];
                double x0 = (double)hi0 - (double)lo0, x1 = (double)hi1 - (double)lo1, x2 = (double)hi2 - (double)lo2;
                int cd = 0; double best = x0;
                if (x1 > best) { cd = 1; best = x1; }
                if (x2 > best) { cd = 2; }
                split = (e1 - s1) > bucket;
                NodeLo a; a.x = round_down(lo0); a.y = round_down(lo1); a.z = round_down(lo2); a.start = s + s1;
                NodeHi b; b.x = round_up(hi0); b.y = round_up(hi1); b.z = round_up(hi2); b.end = s + e1;
                nlo[hidx] = a; nhi[hidx] = b;
                cutdim[hidx] = split ? (int8_t)cd : (int8_t)-1;
                if (split) {
                    const int m1 = s1 + (int)split_left(e1 - s1, al) - 1;
                    const int em = cd == 0 ? ord0[m1] : (cd == 1 ? ord1[m1] : ord2[m1]);       // the median element
                    n_m[k] = (uint16_t)m1; n_cd[k] = (int8_t)cd;
                    n_mr[k] = (uint16_t)(cd == 0 ? em : (cd == 1 ? (int)rk1[em] : (int)rk2[em]));
                    if (t < nsub) { ns_[2 * k] = (uint16_t)s1; ne_[2 * k] = (uint16_t)(m1 + 1); ns_[2 * k + 1] = (uint16_t)(m1 + 1); ne_[2 * k + 1] = (uint16_t)e1; }
                }
            }
            if (!split) {
                n_cd[k] = 3;
                if (t < nsub) { ns_[2 * k] = 0; ne_[2 * k] = 0; ns_[2 * k + 1] = 0; ne_[2 * k + 1] = 0; }
            }
        }
        __syncthreads();
        if (t == nsub) break;
        // ---- stable partition of the three local orders, one pass ---------------------------------------------------
        uint16_t el[3][V2_PER];
        unsigned fl[3] = {0u, 0u, 0u};
        unsigned act = 0, isstart = 0;
        int kk[V2_PER];
        unsigned long long c = 0;           // three 16-bit-safe counts packed: bits 0.., 20.., 40..
#pragma unroll
        for (int r = 0; r < V2_PER; r++) {
            const int i = tid * V2_PER + r;
            kk[r] = 0xffff;
            el[0][r] = el[1][r] = el[2][r] = 0;
            if (i < S_) {
                el[0][r] = ord0[i]; el[1][r] = ord1[i]; el[2][r] = ord2[i];
                const int k = pnode[i];
                kk[r] = k;
                if (k != 0xffff) {
                    const int cd = n_cd[k];
                    if (cd != 3) {
                        act |= 1u << r;
                        if (i == (int)cs_[k]) isstart |= 1u << r;
                        const int m1 = n_m[k], mr = n_mr[k];
                        const bool pos_side = i > m1;
#pragma unroll
                        for (int d = 0; d < 3; d++) {
                            const int ee = el[d][r];
                            const int rnk = cd == 0 ? ee : (cd == 1 ? (int)rk1[ee] : (int)rk2[ee]);
                            const bool f = (cd == d) ? pos_side : (rnk > mr);
                            if (f) { fl[d] |= 1u << r; c += 1ull << (20 * d); }
                        }
                    }
                }
            }
        }
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += v; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        unsigned long long run = inc - c;
        for (int q = 0; q < (int)w; q++) run += wsum[q];
        int pre[3][V2_PER];
#pragma unroll
        for (int r = 0; r < V2_PER; r++) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                pre[d][r] = (int)((run >> (20 * d)) & 0xfffffull);
                if ((isstart >> r) & 1u) n_pre[d * ntab + kk[r]] = (uint16_t)pre[d][r];
                if ((fl[d] >> r) & 1u) run += 1ull << (20 * d);
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < V2_PER; r++) {
            const int i = tid * V2_PER + r;
            if (i < S_) {
                int dst[3] = {i, i, i};
                if ((act >> r) & 1u) {
                    const int k = kk[r];
                    const int m1 = n_m[k];
#pragma unroll
                    for (int d = 0; d < 3; d++) {
                        const int R = pre[d][r] - (int)n_pre[d * ntab + k];
                        dst[d] = ((fl[d] >> r) & 1u) ? m1 + 1 + R : i - R;
                    }
                    pnode[i] = (uint16_t)(2 * k + (i > m1 ? 1 : 0));
                } else if (kk[r] != 0xffff) {
                    pnode[i] = (uint16_t)0xffff;
                }
                ord0[dst[0]] = el[0][r]; ord1[dst[1]] = el[1][r]; ord2[dst[2]] = el[2][r];
            }
        }
        __syncthreads();
    }
    // ---- tree order of this node's range ------------------------------------------------------------------------------------
    for (int i = tid; i < S_; i += V2_T) tree_ord[s + i] = ordx[cur.c[0][0][s + ord0[i]]];
}

template <class S> struct KeyOf;
template <> struct KeyOf<float> { typedef uint32_t type; };
template <> struct KeyOf<double> { typedef uint64_t type; };

// Tree shape: depends only on n, bucket and the split rule (split_left; reference: left = ceil(size/2), leaf iff size <= bucket,
// KDTree.cxx:994,1012)
static void tree_shape(nbk_tree& t) {
    const int64_t n = t.n;
    const int bucket = t.bucket, al = t.aligned;
    int depth = 0;
    {
        int64_t smax = n;                       // the left child is never the smaller one
        while (smax > bucket) { smax = split_left(smax, al); depth++; }
    }
    NBK_REQUIRE(depth <= 30, NBK_ERR_ARG, "tree too deep for 32-bit node indices");
    t.depth = depth;
    t.nslots = ((int64_t)1 << (depth + 1)) - 1;
    // node sizes of a level with their multiplicities (two sizes for the reference rule, a few more with a ragged last unit)
    std::vector<std::pair<int64_t, int64_t>> lv(1, std::make_pair(n, (int64_t)1)), nx;
    int64_t nodes = 0, leaves = 0;
    for (int l = 0; l <= depth && !lv.empty(); l++) {
        nx.clear();
        for (const auto& sc : lv) {
            nodes += sc.second;
            if (sc.first <= bucket) { leaves += sc.second; continue; }
            const int64_t left = split_left(sc.first, al);
            for (int64_t ch : {left, sc.first - left}) {
                bool found = false;
                for (auto& e : nx) if (e.first == ch) { e.second += sc.second; found = true; break; }
                if (!found) nx.push_back(std::make_pair(ch, sc.second));
            }
        }
        NBK_REQUIRE(nx.size() <= 64, NBK_ERR_ARG, "internal: too many node sizes on a level");
        lv.swap(nx);
    }
    t.num_nodes = nodes; t.num_leaves = leaves;
}

template <class S>
static void build_tree_v2(nbk_tree& t, const Vec4<S>* prim_in, const Vec4<S>* sec_in, const double* mass_in) {
    typedef typename KeyOf<S>::type K;
    const int64_t n = t.n;
    cudaStream_t st = t.stream;
    const int bucket = t.bucket;
    int64_t launches = 0;
    Tracer tr(st);
    tree_shape(t);
    const int depth = t.depth;
    // global levels: all levels whose nodes are larger than V2_CAP (every node of such a level is split: V2_CAP >= bucket)
    const int al = t.aligned;
    int L = 0;
    for (int64_t smax = n; smax > V2_CAP; smax = split_left(smax, al)) L++;      // largest node of level L <= V2_CAP
    NBK_REQUIRE(L <= depth, NBK_ERR_ARG, "internal: bucket larger than the shared-memory node capacity");
    const int nsub = depth - L;
    const int ntab = 1 << nsub;

    DevBuf<Vec4<S>> prim(n), sec(sec_in ? n : 0);
    DevBuf<double> mass(n);
    DevBuf<int32_t> order(n);
    DevBuf<NodeLo> nlo(t.nslots);
    DevBuf<NodeHi> nhi(t.nslots);
    DevBuf<int8_t> cutdim(t.nslots);
    NBK_CHECK(cudaMemsetAsync(nlo.p, 0xff, nlo.bytes(), st));      // absent nodes: start = end = -1, cutdim = -1
    NBK_CHECK(cudaMemsetAsync(nhi.p, 0xff, nhi.bytes(), st));
    NBK_CHECK(cudaMemsetAsync(cutdim.p, 0xff, cutdim.bytes(), st));

    DevBuf<K> skeys[3];
    DevBuf<uint32_t> ordx(n);                // pid at x-rank
    DevBuf<uint32_t> colbuf((size_t)18 * n); // two sets of 9 rank columns
    Cols cols[2];
    for (int b = 0; b < 2; b++)
        for (int d = 0; d < 3; d++)
            for (int e = 0; e < 3; e++) cols[b].c[d][e] = colbuf.p + ((size_t)(b * 9 + d * 3 + e)) * n;
    {
        DevBuf<uint32_t> ordA1(n), ordA2(n), ordB[3];
        uint32_t* bufA[3] = {ordx.p, ordA1.p, ordA2.p};      // the x sort works in ordx (its result buffer)
        DevBuf<K> keys_a(n), keys_b(n);
        RadixSortPlan<K> plan(n);
        DevBuf<uint32_t> temp(plan.temp_u32());
        const int kb = (int)sizeof(K) * 8;
        uint32_t* od[3];
        for (int d = 0; d < 3; d++) {
            skeys[d].alloc(n);
            ordB[d].alloc(n);
            make_keys_kernel<S, K><<<div_up(n, 256), 256, 0, st>>>(prim_in, n, d, keys_a.p);
            launches++;
            K* rk; uint32_t* rv;
            radix_sort_pairs<K>(keys_a.p, bufA[d], keys_b.p, ordB[d].p, n, kb, true, temp.p, st, &rk, &rv, &launches);
            NBK_CHECK(cudaMemcpyAsync(skeys[d].p, rk, sizeof(K) * n, cudaMemcpyDeviceToDevice, st));
            od[d] = rv;
        }
        if (od[0] != ordx.p) NBK_CHECK(cudaMemcpyAsync(ordx.p, od[0], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
        tr.point("build: 3 radix sorts");
        {
            DevBuf<uint32_t> R((size_t)4 * n);
            v2_rank_scatter_kernel<<<div_up(n, 256), 256, 0, st>>>(n, od[0], od[1], od[2], R.p);
            for (int d = 0; d < 3; d++)
                v2_rank_gather_kernel<<<div_up(n, 256), 256, 0, st>>>(n, od[d], reinterpret_cast<const uint4*>(R.p), cols[0].c[d][0], cols[0].c[d][1], cols[0].c[d][2]);
            launches += 4;
            NBK_CHECK(cudaStreamSynchronize(st));
        }
        tr.point("build: rank columns");
    }
    int cur = 0;
    if (L > 0) {
        const int ntiles = div_up(n, PRIM_TILE);
        const size_t maxnodes = (size_t)1 << (L - 1);
        DevBuf<int32_t> lv_cd(maxnodes);
        DevBuf<uint32_t> lv_mr(maxnodes), rcount(maxnodes);
        DevBuf<uint32_t> tsum((size_t)3 * ntiles);
        size_t sc = scan_scratch_elems((int64_t)3 * ntiles), sc2 = scan_scratch_elems((int64_t)maxnodes);
        DevBuf<uint32_t> scratch(sc > sc2 ? sc : sc2);
        for (int l = 0; l < L; l++) {
            const int64_t C = (int64_t)1 << l;
            v2_level_nodes_kernel<K><<<div_up(C, 128), 128, 0, st>>>(l, n, al, cols[cur], skeys[0].p, skeys[1].p, skeys[2].p, nlo.p, nhi.p, cutdim.p,
                                                                     lv_cd.p, lv_mr.p, rcount.p);
            exclusive_scan_u32(rcount.p, rcount.p, C, scratch.p, st, &launches);
            v2_count_kernel<<<dim3(ntiles, 3), PRIM_THREADS, 0, st>>>(l, n, al, cols[cur], lv_cd.p, lv_mr.p, tsum.p, ntiles);
            exclusive_scan_u32(tsum.p, tsum.p, (int64_t)3 * ntiles, scratch.p, st, &launches);
            v2_scatter_kernel<<<dim3(ntiles, 3), PRIM_THREADS, 0, st>>>(l, n, al, cols[cur], cols[cur ^ 1], lv_cd.p, lv_mr.p, tsum.p, ntiles, rcount.p);
            launches += 3;
            cur ^= 1;
            if (tr.on) { char lb[64]; snprintf(lb, sizeof(lb), "build: level %d", l); tr.point(lb); }
        }
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    {
        DevBuf<uint32_t> tree_ord(n);
        const size_t smem = v2_small_smem((int)sizeof(S), ntab);
        NBK_REQUIRE(smem <= 227 * 1024, NBK_ERR_ARG, "internal: small-node kernel does not fit shared memory");
        NBK_CHECK(cudaFuncSetAttribute(v2_small_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        v2_small_kernel<S><<<(unsigned)((int64_t)1 << L), V2_T, smem, st>>>(L, nsub, n, bucket, al, cols[cur], prim_in, ordx.p, nlo.p, nhi.p, cutdim.p,
                                                                           tree_ord.p, ntab);
        NBK_CHECK(cudaGetLastError());
        tr.point("build: small nodes");
        if (t.before_gather) { t.before_gather(); t.before_gather = nullptr; }
        finalize_kernel<S><<<div_up(n, 256), 256, 0, st>>>(n, tree_ord.p, prim_in, sec_in, mass_in, prim.p, sec.p, mass.p, order.p);
        launches += 2;
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    NBK_CHECK(cudaGetLastError());
    tr.point("build: gather");

    t.device_bytes = (int64_t)(prim.bytes() + sec.bytes() + mass.bytes() + order.bytes() + nlo.bytes() + nhi.bytes() + cutdim.bytes());
    t.prim = prim.p; prim.p = nullptr;
    t.sec = sec.p; sec.p = nullptr;
    t.mass = mass.p; mass.p = nullptr;
    t.order = order.p; order.p = nullptr;
    t.nlo = nlo.p; nlo.p = nullptr;
    t.nhi = nhi.p; nhi.p = nullptr;
    t.cutdim = cutdim.p; cutdim.p = nullptr;
    t.last_launches = launches;
}

template <class S>
void build_tree(nbk_tree& t, const Vec4<S>* prim_in, const Vec4<S>* sec_in, const double* mass_in) {
    build_tree_v2<S>(t, prim_in, sec_in, mass_in);
}

template void build_tree<float>(nbk_tree&, const Vec4<float>*, const Vec4<float>*, const double*);
template void build_tree<double>(nbk_tree&, const Vec4<double>*, const Vec4<double>*, const double*);

}  // namespace nbk
