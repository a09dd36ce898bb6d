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
__global__ void build_level_nodes_kernel(int level, int64_t n, int bucket, const Vec4<S>* __restrict__ P, const uint32_t* __restrict__ o0,
                                         const uint32_t* __restrict__ o1, const uint32_t* __restrict__ o2, NodeLo* __restrict__ nlo,
                                         NodeHi* __restrict__ nhi, int8_t* __restrict__ cutdim, uint32_t* __restrict__ rcount) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t C = (int64_t)1 << level;
    if (j >= C) return;
    int64_t idx = C - 1 + j;
    int s, e;
    if (level == 0) { s = 0; e = (int)n; }
    else {
        int64_t p = (idx - 1) >> 1;
        int ps = nlo[p].start, pe = nhi[p].end;
        if (ps < 0 || pe - ps <= bucket) {
            NodeLo a; a.x = a.y = a.z = 0.f; a.start = -1;
            NodeHi b; b.x = b.y = b.z = 0.f; b.end = -1;
            nlo[idx] = a; nhi[idx] = b; cutdim[idx] = -1; rcount[j] = 0;
            return;
        }
        int pm = ps + (pe - ps - 1) / 2;
        if (idx & 1) { s = ps; e = pm + 1; } else { s = pm + 1; e = pe; }
    }
    S lo0 = P[o0[s]].x, hi0 = P[o0[e - 1]].x;
    S lo1 = P[o1[s]].y, hi1 = P[o1[e - 1]].y;
    S lo2 = P[o2[s]].z, hi2 = P[o2[e - 1]].z;
    double e0 = (double)hi0 - (double)lo0, e1 = (double)hi1 - (double)lo1, e2 = (double)hi2 - (double)lo2;
    int cd = 0; double best = e0;
    if (e1 > best) { cd = 1; best = e1; }
    if (e2 > best) { cd = 2; }
    bool split = (e - s) > bucket;
    NodeLo a; a.x = round_down(lo0); a.y = round_down(lo1); a.z = round_down(lo2); a.start = s;
    NodeHi b; b.x = round_up(hi0); b.y = round_up(hi1); b.z = round_up(hi2); b.end = e;
    nlo[idx] = a; nhi[idx] = b;
    cutdim[idx] = split ? (int8_t)cd : (int8_t)-1;
    rcount[j] = split ? (uint32_t)(e - (s + (e - s - 1) / 2 + 1)) : 0u;
}

// advance pos_node to this level's node and publish each particle's side bit (by particle id).
__global__ void mark_side_kernel(int level, int64_t n, int bucket, const uint32_t* __restrict__ o0, const uint32_t* __restrict__ o1,
                                 const uint32_t* __restrict__ o2, const NodeLo* __restrict__ nlo, const NodeHi* __restrict__ nhi,
                                 const int8_t* __restrict__ cutdim, int32_t* __restrict__ pos_node, uint8_t* __restrict__ side) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int pn = pos_node[i];
    if (level > 0) {
        int Fprev = (1 << (level - 1)) - 1;
        if (pn >= Fprev) {
            int ps = nlo[pn].start, pe = nhi[pn].end;
            if (pe - ps > bucket) {
                int pm = ps + (pe - ps - 1) / 2;
                pn = ((int)i > pm) ? 2 * pn + 2 : 2 * pn + 1;
                pos_node[i] = pn;
            }
        }
    }
    int F = (1 << level) - 1;
    if (pn >= F) {
        int s = nlo[pn].start, e = nhi[pn].end;
        if (e - s > bucket) {
            int m = s + (e - s - 1) / 2;
            int cd = cutdim[pn];
            const uint32_t* o = cd == 0 ? o0 : (cd == 1 ? o1 : o2);
            side[o[i]] = ((int)i > m) ? 1 : 0;
        }
    }
}

struct PartNode { int active, s, m, cd; };
__device__ __forceinline__ PartNode part_node(int pn, int F, int bucket, const NodeLo* nlo, const NodeHi* nhi, const int8_t* cutdim) {
    PartNode r; r.active = 0; r.s = 0; r.m = 0; r.cd = -1;
    if (pn >= F) {
        int s = nlo[pn].start, e = nhi[pn].end;
        if (e - s > bucket) { r.active = 1; r.s = s; r.m = s + (e - s - 1) / 2; r.cd = cutdim[pn]; }
    }
    return r;
}

__global__ void __launch_bounds__(PRIM_THREADS) part_count_kernel(int level, int64_t n, int bucket, const uint32_t* __restrict__ o0,
                                                                  const uint32_t* __restrict__ o1, const uint32_t* __restrict__ o2,
                                                                  const NodeLo* __restrict__ nlo, const NodeHi* __restrict__ nhi,
                                                                  const int8_t* __restrict__ cutdim, const int32_t* __restrict__ pos_node,
                                                                  const uint8_t* __restrict__ side, uint32_t* __restrict__ tsum, int ntiles,
                                                                  uint32_t* __restrict__ fbits) {
    // also publishes every element's side flag as one bit (one 32-bit ballot per warp and round) so that the scatter
    // kernel does not have to repeat the random side[] gather
    __shared__ uint32_t swarp[8];
    const int d = blockIdx.y;
    const uint32_t* o = d == 0 ? o0 : (d == 1 ? o1 : o2);
    const int F = (1 << level) - 1;
    int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
    const size_t words_per_dim = (size_t)ntiles * (PRIM_TILE / 32);
    uint32_t cnt = 0;
#pragma unroll 4
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int64_t i = base + r * PRIM_THREADS + threadIdx.x;
        uint32_t f = 0;
        if (i < n) {
            PartNode pn = part_node(pos_node[i], F, bucket, nlo, nhi, cutdim);
            if (pn.active) f = (pn.cd == d) ? ((int)i > pn.m ? 1u : 0u) : (uint32_t)side[o[i]];
        }
        uint32_t b = __ballot_sync(0xffffffffu, f != 0);
        if ((threadIdx.x & 31) == 0) fbits[(size_t)d * words_per_dim + (size_t)((base + r * PRIM_THREADS + threadIdx.x) >> 5)] = b;
        cnt += f;
    }
    uint32_t tot;
    block_excl_scan_256(cnt, swarp, &tot);
    if (threadIdx.x == 0) tsum[(size_t)d * ntiles + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(PRIM_THREADS) part_scatter_kernel(int level, int64_t n, int bucket, const uint32_t* __restrict__ o0,
                                                                    const uint32_t* __restrict__ o1, const uint32_t* __restrict__ o2,
                                                                    uint32_t* __restrict__ n0, uint32_t* __restrict__ n1, uint32_t* __restrict__ n2,
                                                                    const NodeLo* __restrict__ nlo, const NodeHi* __restrict__ nhi,
                                                                    const int8_t* __restrict__ cutdim, const int32_t* __restrict__ pos_node,
                                                                    const uint32_t* __restrict__ fbits_all, const uint32_t* __restrict__ tscan, int ntiles,
                                                                    const uint32_t* __restrict__ Rb) {
    __shared__ uint32_t cnt[PRIM_ITEMS * 8];
    __shared__ uint32_t swarp[8];
    const int d = blockIdx.y;
    const uint32_t* o = d == 0 ? o0 : (d == 1 ? o1 : o2);
    uint32_t* out = d == 0 ? n0 : (d == 1 ? n1 : n2);
    const int F = (1 << level) - 1;
    const unsigned lane = lane_id(), w = threadIdx.x >> 5, lt = lanemask_lt();
    int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
    uint32_t bal[PRIM_ITEMS];
    uint32_t fbits = 0;
    const size_t words_per_dim = (size_t)ntiles * (PRIM_TILE / 32);
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        // the flag ballots were written by part_count_kernel: one broadcast word per warp and round
        uint32_t b = fbits_all[(size_t)d * words_per_dim + (size_t)((base + r * PRIM_THREADS + threadIdx.x) >> 5)];
        bal[r] = b;
        fbits |= ((b >> lane) & 1u) << r;
        if (lane == 0) cnt[r * 8 + w] = __popc(b);
    }
    __syncthreads();
    uint32_t v = threadIdx.x < PRIM_ITEMS * 8 ? cnt[threadIdx.x] : 0u;
    uint32_t tot;
    uint32_t ex = block_excl_scan_256(v, swarp, &tot);
    if (threadIdx.x < PRIM_ITEMS * 8) cnt[threadIdx.x] = ex;
    __syncthreads();
    const uint32_t tile_off = tscan[(size_t)d * ntiles + blockIdx.x] - tscan[(size_t)d * ntiles];
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int64_t i = base + r * PRIM_THREADS + threadIdx.x;
        if (i < n) {
            uint32_t pid = o[i];
            int node = pos_node[i];
            PartNode pn = part_node(node, F, bucket, nlo, nhi, cutdim);
            int64_t dest = i;
            if (pn.active) {
                uint32_t P = tile_off + cnt[r * 8 + w] + __popc(bal[r] & lt);
                uint32_t R = P - Rb[node - F];
                dest = ((fbits >> r) & 1u) ? (int64_t)pn.m + 1 + R : i - (int64_t)R;
            }
            out[dest] = pid;
        }
    }
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

template <class S> struct KeyOf;
template <> struct KeyOf<float> { typedef uint32_t type; };
template <> struct KeyOf<double> { typedef uint64_t type; };

template <class S>
void build_tree(nbk_tree& t, const Vec4<S>* prim_in, const Vec4<S>* sec_in, const double* mass_in) {
    typedef typename KeyOf<S>::type K;
    const int64_t n = t.n;
    cudaStream_t st = t.stream;
    const int bucket = t.bucket;
    int64_t launches = 0;
    Tracer tr(st);

    // ---- shape (depends only on n and bucket) ------------------------------------------------------
    int depth = 0;
    {
        int64_t smax = n;
        while (smax > bucket) { smax = (smax + 1) / 2; depth++; }
    }
    NBK_REQUIRE(depth <= 30, NBK_ERR_ARG, "tree too deep for 32-bit node indices");
    t.depth = depth;
    t.nslots = ((int64_t)1 << (depth + 1)) - 1;
    {   // node / leaf counts: at most two distinct sizes per level
        int64_t sz[2] = {n, -1}, ct[2] = {1, 0};
        int64_t nodes = 0, leaves = 0;
        for (int l = 0; l <= depth; l++) {
            int64_t nsz[2] = {-1, -1}, nct[2] = {0, 0};
            for (int q = 0; q < 2; q++) {
                if (ct[q] == 0) continue;
                nodes += ct[q];
                if (sz[q] <= bucket) { leaves += ct[q]; continue; }
                int64_t ch[2] = {(sz[q] + 1) / 2, sz[q] / 2};
                for (int c = 0; c < 2; c++) {
                    int slot = (nsz[0] == ch[c] || nsz[0] < 0) ? 0 : 1;
                    if (nsz[slot] >= 0 && nsz[slot] != ch[c]) throw Error(NBK_ERR_ARG, "internal: >2 node sizes on a level");
                    nsz[slot] = ch[c]; nct[slot] += ct[q];
                }
            }
            sz[0] = nsz[0]; sz[1] = nsz[1]; ct[0] = nct[0]; ct[1] = nct[1];
        }
        t.num_nodes = nodes; t.num_leaves = leaves;
    }

    // ---- persistent outputs ------------------------------------------------------------------------
    DevBuf<Vec4<S>> prim(n), sec(sec_in ? n : 0);
    DevBuf<double> mass(n);
    DevBuf<int32_t> order(n);
    DevBuf<NodeLo> nlo(t.nslots);
    DevBuf<NodeHi> nhi(t.nslots);
    DevBuf<int8_t> cutdim(t.nslots);

    // ---- temporaries -------------------------------------------------------------------------------
    DevBuf<uint32_t> ordA[3], ordB[3];
    for (int d = 0; d < 3; d++) { ordA[d].alloc(n); ordB[d].alloc(n); }
    {
        DevBuf<K> keys_a(n), keys_b(n);
        RadixSortPlan<K> plan(n);
        DevBuf<uint32_t> temp(plan.temp_u32());
        const int kb = (int)sizeof(K) * 8;
        for (int d = 0; d < 3; d++) {
            make_keys_kernel<S, K><<<div_up(n, 256), 256, 0, st>>>(prim_in, n, d, keys_a.p);
            launches++;
            K* rk; uint32_t* rv;
            radix_sort_pairs<K>(keys_a.p, ordA[d].p, keys_b.p, ordB[d].p, n, kb, true, temp.p, st, &rk, &rv, &launches);
            if (rv != ordA[d].p) {  // odd number of passes never happens for 32/64-bit keys, but stay safe
                DevBuf<uint32_t> tmp = std::move(ordA[d]); ordA[d] = std::move(ordB[d]); ordB[d] = std::move(tmp);
            }
        }
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    tr.point("build: 3 radix sorts");
    if (depth > 0) {
        const int ntiles = div_up(n, PRIM_TILE);
        DevBuf<int32_t> pos_node(n);
        DevBuf<uint8_t> side(n);
        DevBuf<uint32_t> rcount((size_t)1 << (depth > 0 ? depth - 1 : 0));
        DevBuf<uint32_t> tsum((size_t)3 * ntiles);
        DevBuf<uint32_t> fbits((size_t)3 * ntiles * (PRIM_TILE / 32));
        size_t sc = scan_scratch_elems((int64_t)3 * ntiles);
        size_t sc2 = scan_scratch_elems((int64_t)rcount.n);
        DevBuf<uint32_t> scratch(sc > sc2 ? sc : sc2);
        NBK_CHECK(cudaMemsetAsync(pos_node.p, 0, pos_node.bytes(), st));
        uint32_t *o[3] = {ordA[0].p, ordA[1].p, ordA[2].p}, *nw[3] = {ordB[0].p, ordB[1].p, ordB[2].p};
        for (int l = 0; l < depth; l++) {
            int64_t C = (int64_t)1 << l;
            build_level_nodes_kernel<S><<<div_up(C, 128), 128, 0, st>>>(l, n, bucket, prim_in, o[0], o[1], o[2], nlo.p, nhi.p, cutdim.p, rcount.p);
            exclusive_scan_u32(rcount.p, rcount.p, C, scratch.p, st, &launches);
            mark_side_kernel<<<div_up(n, 256), 256, 0, st>>>(l, n, bucket, o[0], o[1], o[2], nlo.p, nhi.p, cutdim.p, pos_node.p, side.p);
            part_count_kernel<<<dim3(ntiles, 3), PRIM_THREADS, 0, st>>>(l, n, bucket, o[0], o[1], o[2], nlo.p, nhi.p, cutdim.p, pos_node.p, side.p, tsum.p, ntiles, fbits.p);
            exclusive_scan_u32(tsum.p, tsum.p, (int64_t)3 * ntiles, scratch.p, st, &launches);
            part_scatter_kernel<<<dim3(ntiles, 3), PRIM_THREADS, 0, st>>>(l, n, bucket, o[0], o[1], o[2], nw[0], nw[1], nw[2], nlo.p, nhi.p, cutdim.p,
                                                                           pos_node.p, fbits.p, tsum.p, ntiles, rcount.p);
            launches += 4;
            for (int d = 0; d < 3; d++) { uint32_t* tp = o[d]; o[d] = nw[d]; nw[d] = tp; }
            if (tr.on) { char lb[64]; snprintf(lb, sizeof(lb), "build: level %d", l); tr.point(lb); }
        }
        // deepest level: leaves only
        {
            int64_t C = (int64_t)1 << depth;
            DevBuf<uint32_t> rc2((size_t)C);
            build_level_nodes_kernel<S><<<div_up(C, 128), 128, 0, st>>>(depth, n, bucket, prim_in, o[0], o[1], o[2], nlo.p, nhi.p, cutdim.p, rc2.p);
            launches++;
            finalize_kernel<S><<<div_up(n, 256), 256, 0, st>>>(n, o[0], prim_in, sec_in, mass_in, prim.p, sec.p, mass.p, order.p);
            launches++;
            NBK_CHECK(cudaStreamSynchronize(st));
        }
    } else {
        DevBuf<uint32_t> rc2(1);
        build_level_nodes_kernel<S><<<1, 128, 0, st>>>(0, n, bucket, prim_in, ordA[0].p, ordA[1].p, ordA[2].p, nlo.p, nhi.p, cutdim.p, rc2.p);
        finalize_kernel<S><<<div_up(n, 256), 256, 0, st>>>(n, ordA[0].p, prim_in, sec_in, mass_in, prim.p, sec.p, mass.p, order.p);
        launches += 2;
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    NBK_CHECK(cudaGetLastError());
    tr.point("build: leaves + gather");

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

template void build_tree<float>(nbk_tree&, const Vec4<float>*, const Vec4<float>*, const double*);
template void build_tree<double>(nbk_tree&, const Vec4<double>*, const Vec4<double>*, const double*);

}  // namespace nbk
