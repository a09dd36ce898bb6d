// traverse.cuh -- warp-per-query-bucket traversal shared by kNN, ball search and FOF.
//
// One warp owns 32 queries that are adjacent in tree order (spatially compact).  The warp walks the
// heap-ordered node arrays with ONE shared stack (uniform control flow, one broadcast load per node);
// every lane tests the node against its own query with a rigorous fp32 LOWER bound of the squared
// distance to the node box (directed rounding), and the node is opened if any lane needs it.  Leaf
// particles are staged once per warp into a shared-memory tile (coalesced 16/32-byte loads, widened
// to fp64) and every lane evaluates the reference's fp64 distance against the broadcast tile.
//
// Replaces SplitNode::FindNearestPos / SearchBallPos / FOFSearchBall recursion (reference
// KDSplitNode.cxx:15-41, 455-690, 921-988) -- pruning there is Arya-Mount incremental offsets; any
// conservative pruning yields the same result set, so box bounds are used here.
#pragma once
#include "common.cuh"

namespace nbk {

constexpr int TRAV_STACK = 64;

// lower bound (rounded toward -inf at every step) of the squared distance from a query known to lie in
// [qlo,qhi] (component-wise fp32 enclosure of the fp64 query) to the box [lo,hi].
__device__ __forceinline__ float box_lb(float qlx, float qly, float qlz, float qhx, float qhy, float qhz, const NodeLo& lo, const NodeHi& hi) {
    float dx = fmaxf(fmaxf(__fsub_rd(lo.x, qhx), __fsub_rd(qlx, hi.x)), 0.f);
    float dy = fmaxf(fmaxf(__fsub_rd(lo.y, qhy), __fsub_rd(qly, hi.y)), 0.f);
    float dz = fmaxf(fmaxf(__fsub_rd(lo.z, qhz), __fsub_rd(qlz, hi.z)), 0.f);
    return __fadd_rd(__fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy)), __fmul_rd(dz, dz));
}
// upper bound (rounded toward +inf) of the squared distance from the query to the farthest box corner
__device__ __forceinline__ float box_ub(float qlx, float qly, float qlz, float qhx, float qhy, float qhz, const NodeLo& lo, const NodeHi& hi) {
    float dx = fmaxf(__fsub_ru(qhx, lo.x), __fsub_ru(hi.x, qlx));
    float dy = fmaxf(__fsub_ru(qhy, lo.y), __fsub_ru(hi.y, qly));
    float dz = fmaxf(__fsub_ru(qhz, lo.z), __fsub_ru(hi.z, qlz));
    return __fadd_ru(__fadd_ru(__fmul_ru(dx, dx), __fmul_ru(dy, dy)), __fmul_ru(dz, dz));
}

// the reference distance: ((dx*dx)+dy*dy)+dz*dz, one rounding per operation (DistFunc.h:14-19)
__device__ __forceinline__ double dist2_ref(double ax, double ay, double az, double bx, double by, double bz) {
    double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Link / search predicates shared by FOF, FOFCriterion, SearchBall* and SearchCriterion* (strict '<', fp64, the
// reference's operation order):
//   mode 0  3D ball            DistanceSqd(pos) < p0                               DistFunc.h:14-19
//   mode 1  6D ball (TPHS)     DistanceSqd(pos) + DistanceSqd(vel) < p0            KDLeafNode.cxx:570-572
//   mode 2  FOF3d              sum dx*dx/p0 < 1                                    FOFFunc.h:30-35
//   mode 4  FOF6d              sum (dx*dx/p0 + dv*dv/p1), interleaved, < 1         FOFFunc.h:48-55
// (mode 3, FOFVel, is not implemented: see DESIGN.md)
__device__ __forceinline__ bool crit_needs_vel(int mode) { return mode == 1 || mode == 4; }
// candidate j of a staged tile ([6][32] doubles: x, y, z, vx, vy, vz rows); velocities are read only by the 6D predicates
__device__ __forceinline__ bool crit_linked(int mode, double p0, double p1, double qx, double qy, double qz, double vx, double vy, double vz,
                                            const double* __restrict__ tile, int j) {
    const double cx = tile[j], cy = tile[32 + j], cz = tile[64 + j];
    if (mode == 0) return dist2_ref(qx, qy, qz, cx, cy, cz) < p0;
    const double ux = tile[96 + j], uy = tile[128 + j], uz = tile[160 + j];
    if (mode == 1) {
        double d = dist2_ref(qx, qy, qz, cx, cy, cz);
        d = __dadd_rn(d, dist2_ref(vx, vy, vz, ux, uy, uz));
        return d < p0;
    }
    double dx = __dsub_rn(qx, cx), dy = __dsub_rn(qy, cy), dz = __dsub_rn(qz, cz);
    if (mode == 2) {
        double t = __ddiv_rn(__dmul_rn(dx, dx), p0);
        t = __dadd_rn(t, __ddiv_rn(__dmul_rn(dy, dy), p0));
        t = __dadd_rn(t, __ddiv_rn(__dmul_rn(dz, dz), p0));
        return t < 1.0;
    }
    double wx = __dsub_rn(vx, ux), wy = __dsub_rn(vy, uy), wz = __dsub_rn(vz, uz);
    double t = __ddiv_rn(__dmul_rn(dx, dx), p0);
    t = __dadd_rn(t, __ddiv_rn(__dmul_rn(wx, wx), p1));
    t = __dadd_rn(t, __ddiv_rn(__dmul_rn(dy, dy), p0));
    t = __dadd_rn(t, __ddiv_rn(__dmul_rn(wy, wy), p1));
    t = __dadd_rn(t, __ddiv_rn(__dmul_rn(dz, dz), p0));
    t = __dadd_rn(t, __ddiv_rn(__dmul_rn(wz, wz), p1));
    return t < 1.0;
}

struct QueryBox { float lx, ly, lz, hx, hy, hz; };
__device__ __forceinline__ QueryBox make_qbox(double x, double y, double z) {
    QueryBox q;
    q.lx = __double2float_rd(x); q.hx = __double2float_ru(x);
    q.ly = __double2float_rd(y); q.hy = __double2float_ru(y);
    q.lz = __double2float_rd(z); q.hz = __double2float_ru(z);
    return q;
}

// Visitor interface (all methods called by the full warp):
//   bool need(float lb)             per-lane: could this lane accept something at distance^2 >= lb ?
//   void leaf(int start, int cnt, int node, unsigned nmask)   process particles [start, start+cnt) of node `node`;
//                                   nmask = ballot of the lanes whose bound still reaches the node's box
// ORDERED: always descend the left child first so leaves are met in ascending tree-index order.
//
// Control flow: a node is tested when it is reached as a child; the nearer child (by vote of the lanes that
// need it) is descended into directly and only the other child goes on the stack, to be re-tested against the
// lanes' current bounds when popped.
template <class V, bool ORDERED = false>
__device__ __forceinline__ void traverse_from(const NodeLo* __restrict__ nlo, const NodeHi* __restrict__ nhi, int bucket, int* stack, V& v,
                                              const QueryBox& q, bool on, int root, NodeLo lo, NodeHi hi) {
    const unsigned lane = lane_id();
    int sp = 0;
    int node = root;
    unsigned nmask;     // lanes that need `node`
    {
        float lb = box_lb(q.lx, q.ly, q.lz, q.hx, q.hy, q.hz, lo, hi);
        nmask = __ballot_sync(0xffffffffu, on && v.need(lb));
        if (!nmask) return;
    }
    while (true) {
        // invariant: `node` (bounds lo/hi) is needed by at least one lane
        bool descend = false;
        if (hi.end - lo.start <= bucket) {
            v.leaf(lo.start, hi.end - lo.start, node, nmask);
        } else {
            const int c1 = 2 * node + 1, c2 = c1 + 1;
            NodeLo l1 = nlo[c1]; NodeHi h1 = nhi[c1];
            NodeLo l2 = nlo[c2]; NodeHi h2 = nhi[c2];
            float b1 = box_lb(q.lx, q.ly, q.lz, q.hx, q.hy, q.hz, l1, h1);
            float b2 = box_lb(q.lx, q.ly, q.lz, q.hx, q.hy, q.hz, l2, h2);
            bool n1 = on && v.need(b1), n2 = on && v.need(b2);
            unsigned m1 = __ballot_sync(0xffffffffu, n1), m2 = __ballot_sync(0xffffffffu, n2);
            if (m1 | m2) {
                bool first1;
                if (ORDERED) first1 = true;
                else {
                    unsigned p1 = __ballot_sync(0xffffffffu, (n1 || n2) && (b1 <= b2));
                    first1 = 2 * __popc(p1) >= __popc(m1 | m2);
                }
                if (m1 && m2) {
                    if (lane == 0) stack[sp] = first1 ? c2 : c1;       // the stack is lane 0's alone (see the pop): no warp-level ordering needed
                    sp++;
                }
                bool take1 = m1 && (first1 || !m2);
                node = take1 ? c1 : c2;
                lo = take1 ? l1 : l2;
                hi = take1 ? h1 : h2;
                nmask = take1 ? m1 : m2;
                descend = true;
            }
        }
        if (descend) continue;
        // pop until a still-needed node is found
        bool found = false;
        while (sp > 0) {
            int top = 0;
            --sp;
            if (lane == 0) top = stack[sp];
            node = __shfl_sync(0xffffffffu, top, 0);     // only lane 0 ever touches the stack: no shared-memory hazard between lanes
            lo = nlo[node]; hi = nhi[node];
            float lb = box_lb(q.lx, q.ly, q.lz, q.hx, q.hy, q.hz, lo, hi);
            nmask = __ballot_sync(0xffffffffu, on && v.need(lb));
            if (nmask) { found = true; break; }
        }
        if (!found) return;
    }
}

template <class V, bool ORDERED = false>
__device__ __forceinline__ void traverse(const NodeLo* __restrict__ nlo, const NodeHi* __restrict__ nhi, int bucket, int* stack, V& v,
                                         const QueryBox& q, bool on) {
    traverse_from<V, ORDERED>(nlo, nhi, bucket, stack, v, q, on, 0, nlo[0], nhi[0]);
}

// Bottom-up walk for queries that are particles of the tree: `first` .. `last` (tree positions, inclusive range end
// exclusive) all lie in one node A0, the deepest node that holds the whole group (found by arithmetic: the tree's shape is a
// function of (n, bucket) and the split rule only, split_left).  A0's subtree is walked first, then the sibling subtree of A0 and of
// every ancestor up to the root: the near field comes first whatever the lanes' bounds are (they may still be infinite), and
// the upper levels cost one independent box test each instead of a chain of dependent node loads from the root.  The next
// sibling's record is loaded before the current subtree is walked.
template <class V>
__device__ __forceinline__ void traverse_bottom_up(const NodeLo* __restrict__ nlo, const NodeHi* __restrict__ nhi, int bucket, int* stack, V& v,
                                                   const QueryBox& q, bool on, int64_t ntree, int aligned, int64_t first, int64_t last) {
    int node = 0;
    {
        int64_t s = 0, e = ntree;
        while (e - s > bucket) {
            const int64_t mid = s + split_left(e - s, aligned);
            if (last <= mid) { node = 2 * node + 1; e = mid; }
            else if (first >= mid) { node = 2 * node + 2; s = mid; }
            else break;
        }
    }
    int visit = node, cur = node;
    NodeLo lo = nlo[node];
    NodeHi hi = nhi[node];
    while (true) {
        const bool more = cur != 0;
        int nxt = 0;
        NodeLo lon = lo; NodeHi hin = hi;
        if (more) { nxt = (cur & 1) ? cur + 1 : cur - 1; lon = nlo[nxt]; hin = nhi[nxt]; }
        traverse_from(nlo, nhi, bucket, stack, v, q, on, visit, lo, hi);
        if (!more) break;
        visit = nxt; lo = lon; hi = hin;
        cur = (cur - 1) >> 1;
    }
}

}  // namespace nbk
