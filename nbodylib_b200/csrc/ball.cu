// ball.cu -- batched fixed-radius search (CSR output), two passes over the shared warp traversal.
//
// Replaces KDTree::SearchBallPosTagged (reference KDFindNearest.cxx:618-688) and the node code behind it
// (KDSplitNode.cxx:455-690, KDLeafNode.cxx:247-413, periodic forms KDSplitNode.cxx:1341-1426): every
// particle with d2 < fdist2 (strict, fp64, reference operation order).  Pass 1 counts per query, a prefix
// sum gives the row offsets, pass 2 writes the indices.  Leaves are visited left to right, so each image's
// contribution to a row is ascending in tree index.
//
// The same two passes serve KDTree::SearchCriterionTagged / SearchCriterion (KDFindNearest.cxx:590-603,643-706; leaf code
// KDLeafNode.cxx:414-492): the predicate becomes one of the in-tree FOFcompfunc criteria (crit_linked in traverse.cuh) and
// pruning uses the position radius the criterion implies.  The dense forms (SearchBallPos / SearchCriterion with nn[] and
// dist2[] arrays of N entries, KDFindNearest.cxx:567-603) are served from the CSR rows plus the optional d2 output.
#include "sort_scan.cuh"
#include "traverse.cuh"
#include "tree.h"

namespace nbk {

constexpr int BALL_WARPS = 4;      // small CTAs: a CTA waits for its slowest warp (see fof.cu)

struct BallParams {
    const NodeLo* nlo; const NodeHi* nhi; int bucket;
    const void* P; const void* V; const int32_t* order;
    int64_t m; const int32_t* qidx; const double* xq; const double* vq;
    int mode; double p0, p1;   // predicate (crit_linked); mode 0: p0 = fdist2
    float r2f;                 // position pruning radius^2, rounded up
    int periodic; double period[3];
    uint32_t* counts;          // pass 1 out / pass 2: exclusive offsets in
    int32_t* idx; double* d2; int64_t cap; int out_ids;
};

template <class S, bool FILL>
struct BallVisitor {
    const Vec4<S>* P; const Vec4<S>* V; const int32_t* order;
    double* tile;          // [6][32]
    double qx, qy, qz, vx, vy, vz, p0, p1;
    float r2f;
    int mode;
    int self;              // excluded tree index or -1
    bool on;
    unsigned lane;
    uint32_t count;        // running count (pass 1) / write cursor (pass 2)
    int32_t* idx; double* d2; int64_t cap; int out_ids;

    __device__ __forceinline__ bool need(float lb) const { return lb < r2f; }
    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[start + base + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
                if (crit_needs_vel(mode)) {
                    Vec4<S> u = V[start + base + lane];
                    tile[96 + lane] = (double)u.x; tile[128 + lane] = (double)u.y; tile[160 + lane] = (double)u.z;
                }
            }
            __syncwarp();
            if (!on) continue;
            for (int j = 0; j < m; j++) {
                int c = start + base + j;
                if (c == self) continue;
                if (crit_linked(mode, p0, p1, qx, qy, qz, vx, vy, vz, tile, j)) {
                    if (FILL) {
                        if ((int64_t)count < cap) {
                            idx[count] = out_ids ? order[c] : c;
                            // dense forms report the position distance^2 (KDLeafNode.cxx:255-275, 414-428)
                            if (d2) d2[count] = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
                        }
                    }
                    count++;
                }
            }
        }
    }
};

template <class S, bool FILL>
__global__ void __launch_bounds__(BALL_WARPS * 32) ball_kernel(BallParams prm) {
    __shared__ double s_tile[BALL_WARPS][192];
    __shared__ int s_stack[BALL_WARPS][TRAV_STACK];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    int64_t group = (int64_t)blockIdx.x * BALL_WARPS + w;
    int64_t qi = group * 32 + lane;
    if (group * 32 >= prm.m) return;
    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
    const bool valid = qi < prm.m;
    BallVisitor<S, FILL> v;
    v.P = P; v.V = V; v.order = prm.order; v.tile = s_tile[w]; v.p0 = prm.p0; v.p1 = prm.p1; v.mode = prm.mode; v.r2f = prm.r2f; v.lane = lane;
    v.on = valid; v.self = -1; v.idx = prm.idx; v.d2 = prm.d2; v.cap = prm.cap; v.out_ids = prm.out_ids;
    v.count = (FILL && valid) ? prm.counts[qi] : 0u;
    double x0 = 0, y0 = 0, z0 = 0;
    v.vx = v.vy = v.vz = 0;
    if (valid) {
        if (prm.qidx) {
            int t = prm.qidx[qi];
            Vec4<S> c = P[t];
            x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z;
            if (crit_needs_vel(prm.mode)) { Vec4<S> u = V[t]; v.vx = (double)u.x; v.vy = (double)u.y; v.vz = (double)u.z; }
            // ball search, quirk Q5: periodic target forms go through coordinates and keep the target.  Criterion search:
            // the target is always left out (i != target, KDLeafNode.cxx:445-456; the periodic form compares whole
            // particles and its unshifted copy equals the target, :805-810)
            if (!prm.periodic || prm.mode >= 2) v.self = t;
        } else {
            x0 = prm.xq[3 * qi]; y0 = prm.xq[3 * qi + 1]; z0 = prm.xq[3 * qi + 2];
            if (crit_needs_vel(prm.mode) && prm.vq) { v.vx = prm.vq[3 * qi]; v.vy = prm.vq[3 * qi + 1]; v.vz = prm.vq[3 * qi + 2]; }
        }
    }
    const int nimg = prm.periodic ? 8 : 1;
    for (int img = 0; img < nimg; img++) {
        v.qx = (img & 1) ? ((x0 < prm.period[0] / 2.0) ? x0 + prm.period[0] : x0 - prm.period[0]) : x0;
        v.qy = (img & 2) ? ((y0 < prm.period[1] / 2.0) ? y0 + prm.period[1] : y0 - prm.period[1]) : y0;
        v.qz = (img & 4) ? ((z0 < prm.period[2] / 2.0) ? z0 + prm.period[2] : z0 - prm.period[2]) : z0;
        QueryBox qb = make_qbox(v.qx, v.qy, v.qz);
        traverse<BallVisitor<S, FILL>, true>(prm.nlo, prm.nhi, prm.bucket, s_stack[w], v, qb, valid);
    }
    if (!FILL && valid) prm.counts[qi] = v.count;
}

// 64-bit total of the per-query counts: the offsets are scanned in 32 bits, so a batch whose total does not fit is refused
__global__ void ball_total_kernel(int64_t m, const uint32_t* __restrict__ counts, unsigned long long* total) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = i < m ? counts[i] : 0ull;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}
__global__ void ball_offsets_kernel(int64_t m, const uint32_t* scan, int64_t* offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= m) offsets[i] = (int64_t)scan[i];
}

void launch_ball(nbk_tree& t, BallArgs& a) {
    cudaStream_t st = t.stream;
    const int64_t m = a.m;
    if (m <= 0) { a.total = 0; return; }
    DevBuf<uint32_t> counts(m + 1), scratch(scan_scratch_elems(m + 1));
    NBK_CHECK(cudaMemsetAsync(counts.p, 0, counts.bytes(), st));
    BallParams p;
    p.nlo = t.nlo; p.nhi = t.nhi; p.bucket = t.bucket; p.P = t.pos4(); p.V = t.vel4(); p.order = t.order;
    p.m = m; p.qidx = a.qidx; p.xq = a.xq; p.vq = a.vq;
    p.mode = a.mode; p.p0 = a.mode == 0 ? a.r2 : a.p0; p.p1 = a.p1;
    const double prune = a.mode == 0 ? a.r2 : a.prune_x2;
    p.r2f = __builtin_nextafterf((float)prune, INFINITY);
    if ((double)p.r2f < prune) p.r2f = __builtin_nextafterf(p.r2f, INFINITY);
    if (a.mode == 1 || a.mode == 4) NBK_REQUIRE(p.V != nullptr, NBK_ERR_ARG, "6D search needs velocities");
    p.periodic = t.periodic ? 1 : 0;
    for (int d = 0; d < 3; d++) p.period[d] = t.period[d];
    p.counts = counts.p; p.idx = a.idx; p.d2 = a.d2; p.cap = a.cap; p.out_ids = a.out_ids ? 1 : 0;
    int blocks = div_up((m + 31) / 32, BALL_WARPS);
    int64_t launches = 0;
    NBK_CHECK(cudaEventRecord(t.ev2, st));
    if (t.store_bytes == 4) ball_kernel<float, false><<<blocks, BALL_WARPS * 32, 0, st>>>(p);
    else ball_kernel<double, false><<<blocks, BALL_WARPS * 32, 0, st>>>(p);
    NBK_CHECK(cudaEventRecord(t.ev3, st));
    {
        DevBuf<unsigned long long> tot64(1);
        NBK_CHECK(cudaMemsetAsync(tot64.p, 0, sizeof(unsigned long long), st));
        ball_total_kernel<<<div_up(m, 256), 256, 0, st>>>(m, counts.p, tot64.p);
        unsigned long long h64 = 0;
        NBK_CHECK(cudaMemcpyAsync(&h64, tot64.p, sizeof(h64), cudaMemcpyDeviceToHost, st));
        NBK_CHECK(cudaStreamSynchronize(st));
        NBK_REQUIRE(h64 < 0xffffffffull, NBK_ERR_ARG, "ball / criterion search: the batch returns 2^32 or more entries in all; split the query batch");
    }
    exclusive_scan_u32(counts.p, counts.p, m + 1, scratch.p, st, &launches);
    ball_offsets_kernel<<<div_up(m + 1, 256), 256, 0, st>>>(m, counts.p, a.offsets);
    uint32_t tot = 0;
    NBK_CHECK(cudaMemcpyAsync(&tot, counts.p + m, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NBK_CHECK(cudaStreamSynchronize(st));
    a.total = tot;
    launches += 2;
    if (a.idx && a.cap > 0) {
        if (t.store_bytes == 4) ball_kernel<float, true><<<blocks, BALL_WARPS * 32, 0, st>>>(p);
        else ball_kernel<double, true><<<blocks, BALL_WARPS * 32, 0, st>>>(p);
        launches++;
        NBK_CHECK(cudaStreamSynchronize(st));
    }
    NBK_CHECK(cudaGetLastError());
    float ms = 0;
    NBK_CHECK(cudaEventElapsedTime(&ms, t.ev2, t.ev3));
    t.last_kernel_ms = ms;
    t.last_launches = launches;
}

}  // namespace nbk
