// knn.cu -- k-nearest-neighbour search with the SPH density / velocity-density / smoothing-scale
// epilogues fused into the same kernel.
//
// Replaces (reference): LeafNode::FindNearestPos KDLeafNode.cxx:15-28,119-130; SplitNode::FindNearestPos
// KDSplitNode.cxx:15-41; FindNearestPosPeriodic :1119-1148; PriorityQueue.h:14-85; drivers
// KDFindNearest.cxx:247-334,462-554; KDTree::CalcDensity / CalcVelDensity KDCalcSmoothQuantities.cxx:203-389;
// the phase-space forms LeafNode::FindNearestPhase KDLeafNode.cxx:43-57,143-154, SplitNode::FindNearestPhase(Periodic)
// KDSplitNode.cxx:69-95,260-289,1153-1187 and their drivers KDFindNearest.cxx:347-361,543-555 (PHASE instantiation).
//
// Layout: one warp = 32 queries adjacent in tree order; lane l owns query l and a bounded max-heap in shared
// memory, slot-major / lane-minor ([slot][32]) so that lanes touching different slots never bank-conflict.
// Candidates come from the shared traversal (traverse.cuh); the distance is the reference's fp64 expression on
// exactly widened coordinates, so neighbour sets and every d2 are bit-identical to the reference whatever the
// traversal order.
//
// Two kernels:
//   knn_exact_kernel  heap of (fp64 d2, int32 index), 12 B/entry.  Used when neighbour lists are materialised
//                     (FindNearest API), for the periodic image schedule, and as the fallback below.
//   knn_hp_kernel     the density family (CalcDensity / CalcVelDensity / smoothing scale), non periodic target
//                     form: heap of packed (fp32 key | candidate id) words with a rigorous error bound (described
//                     at the kernel).  Queries whose k-th / (k+1)-th keys are closer than the error band are
//                     re-run by the exact kernel.
#include <string.h>

#include "traverse.cuh"
#include "tree.h"

namespace nbk {

constexpr int KNN_WARPS = 2;      // warps per CTA of the fp64-heap kernel: small CTAs hand their slots back sooner (176.8 -> 168.7 ms at 256^3, k = 64)
constexpr double KNN_SENTINEL = 1e32;   // reference MAXVALUE (Precision.h:49)

// smoothing kernel interpolation, KDCalcSmoothQuantities.cxx:12-15
__device__ __forceinline__ double wsm(double r, int i, int size, double delta, const double* __restrict__ x) {
    if (i < size - 1) { double a = x[i], b = x[i + 1]; return (a + (b - a) * (r - delta * i) / delta); }
    return x[i];
}
// the same with 1/delta precomputed: one multiplication instead of an fp64 division per neighbour (differs from the
// reference expression by one rounding, ~1e-16 relative, inside the density tolerance)
__device__ __forceinline__ double wsm_fast(double r, int i, int size, double delta, double inv_delta, const double* __restrict__ x) {
    if (i < size - 1) { double a = x[i], b = x[i + 1]; return (a + (b - a) * ((r - delta * i) * inv_delta)); }
    return x[i];
}

struct KnnParams {
    const NodeLo* nlo; const NodeHi* nhi; int bucket;
    const NodeLo* nlo2; const NodeHi* nhi2; int bucket2;   // attached halo tree (or null)
    const void* P; const void* V; const double* mass; const int32_t* order;
    int64_t n;
    int64_t q0, q1; const double* xq; int mode;      // mode 0: particles [q0,q1) or qlist; 1: points
    const int32_t* qlist; int64_t nq;                 // optional explicit particle list (exact kernel)
    int k, kcap;
    int periodic, strict, tree_form;
    double period[3];
    int32_t* nn; double* d2out; int out_ids;
    double* rho; double* hsm; int veldens_k;
    const double* kern; int kernres;
    int* flag_count; int32_t* flag_list;              // fast kernel: queries that need the exact kernel
    const uint8_t* active;                            // optional query mask (tree order)
    int gather;                                       // *Particle / *Position forms: gather-only sum, result per ROW in rho
    const double* vq;                                 // point forms that need a query velocity (velocity density, FOF6d filter), m x 3
    // filtered search (FindNearestCheck / FindNearestCriterion): candidates must have cand_excl[c] == 0 and / or meet the
    // criterion crit_mode (0: none, 2: FOF3d, 4: FOF6d; crit_linked in traverse.cuh) relative to the query
    const int32_t* cand_excl; int crit_mode; double cp0, cp1;
    // smoothed velocity moments (CalcSmoothVel / CalcSmoothVelDisp): densities in, accumulators out, all tree order
    const double* rho_in; const double* smvel_in; double* smvel_out; double* smdisp_out;
    const double* smdisp_in; double* smhigh_out; int moment;      // skewness (3) / kurtosis (4) about the receiver's mean, in units of its dispersion
    int64_t n_tree;                                   // particles of the main tree (= n unless a halo is attached)
    int aligned;                                      // the main tree's split rule (split_left)
    int tr_max;                                       // density kernel: transposed screening threshold (lanes needing a tile)
    int phase;                                        // FindNearestPhase: the key is the 6D distance (exact kernel, PHASE instantiation)
};

// ================================================================================================ exact
struct WarpHeap {
    double* H;   // [kcap][32]
    int* I;      // [kcap][32]
    int k;
    unsigned lane;
    __device__ __forceinline__ double& h(int s) { return H[s * 32 + lane]; }
    __device__ __forceinline__ int& i(int s) { return I[s * 32 + lane]; }
    __device__ __forceinline__ void init() {
        for (int s = 0; s < k; s++) { h(s) = KNN_SENTINEL; i(s) = -1; }
    }
    // sift (d,id) down from slot p in a heap of size n
    __device__ __forceinline__ void sift_down(int p, int n, double d, int id) {
        while (true) {
            int c = 2 * p + 1;
            if (c >= n) break;
            double dc = h(c);
            if (c + 1 < n) {
                double dr = h(c + 1);
                if (dr > dc) { c = c + 1; dc = dr; }
            }
            if (d >= dc) break;
            h(p) = dc; i(p) = i(c);
            p = c;
        }
        h(p) = d; i(p) = id;
    }
    __device__ __forceinline__ void replace_top(double d, int id) { sift_down(0, k, d, id); }
    // in-place ascending sort of the first n slots (heap must be valid on [0,n))
    __device__ __forceinline__ void sort_ascending(int n) {
        for (int e = n - 1; e > 0; e--) {
            double dl = h(e); int il = i(e);
            h(e) = h(0); i(e) = i(0);
            sift_down(0, e, dl, il);
        }
    }
    __device__ __forceinline__ void heapify(int n) {
        for (int p = n / 2 - 1; p >= 0; p--) { double d = h(p); int id = i(p); sift_down(p, n, d, id); }
    }
};

// PHASE: the heap key is the reference's 6D PhaseDistSqd (DistFunc.h:41-49: one left-to-right sum of the six squares,
// positions first).  The tree stays the position tree: the position part of the distance is a lower bound of the 6D
// distance, so the box test prunes conservatively and the k-nearest set is the one any exact search finds.
__device__ __forceinline__ double phase_dist2_ref(double qx, double qy, double qz, double vx, double vy, double vz,
                                                  const double* __restrict__ tile, int j) {
    double d = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
    const double wx = __dsub_rn(vx, tile[96 + j]), wy = __dsub_rn(vy, tile[128 + j]), wz = __dsub_rn(vz, tile[160 + j]);
    d = __dadd_rn(d, __dmul_rn(wx, wx));
    d = __dadd_rn(d, __dmul_rn(wy, wy));
    return __dadd_rn(d, __dmul_rn(wz, wz));
}

template <class S, bool FILTER = false, bool PHASE = false>
struct KnnVisitor {
    const Vec4<S>* P;
    const Vec4<S>* V;   // FILTER with a 6D criterion, PHASE
    double* tile;       // [6][32]
    WarpHeap hp;
    double qx, qy, qz;
    double vx, vy, vz;  // query velocity (FILTER with FOF6d, PHASE)
    const int32_t* excl; int crit_mode; double cp0, cp1;
    double top;         // heap top (current k-th distance^2)
    float topf;         // top rounded up
    int self;           // tree index of the query particle, or -1 (coordinate form)
    bool target_form;   // skip self and d2==0 (KDLeafNode.cxx:15-28)
    bool on;
    unsigned lane;

    __device__ __forceinline__ bool need(float lb) const { return lb < topf; }
    __device__ __forceinline__ void settop() { top = hp.h(0); topf = __double2float_ru(top); }
    __device__ __forceinline__ double dist(int j) const {
        if (PHASE) return phase_dist2_ref(qx, qy, qz, vx, vy, vz, tile, j);
        return dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
    }

    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[start + base + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
                if (PHASE || (FILTER && crit_mode == 4)) {
                    Vec4<S> u = V[start + base + lane];
                    tile[96 + lane] = (double)u.x; tile[128 + lane] = (double)u.y; tile[160 + lane] = (double)u.z;
                }
            }
            __syncwarp();
            unsigned acc = 0;
#pragma unroll 4
            for (int j = 0; j < m; j++) {
                double d2 = dist(j);
                bool ok = on && d2 < top;
                if (target_form) ok = ok && (start + base + j != self) && d2 > 0.0;
                if (FILTER) {
                    // KDLeafNode.cxx:88-118,202-246: i != target, 0 < d2 < top, check(bucket[i]) == 0 / cmp(target, bucket[i]) == 1
                    ok = ok && (start + base + j != self) && d2 > 0.0;
                    if (ok && excl) ok = excl[start + base + j] == 0;
                    if (ok && crit_mode) ok = crit_linked(crit_mode, cp0, cp1, qx, qy, qz, vx, vy, vz, tile, j);
                }
                acc |= (ok ? 1u : 0u) << j;
            }
            while (__any_sync(0xffffffffu, acc != 0)) {
                if (acc) {
                    int j = __ffs(acc) - 1;
                    acc &= acc - 1;
                    double d2 = dist(j);
                    if (d2 < top) { hp.replace_top(d2, start + base + j); settop(); }
                }
            }
        }
    }
};

constexpr int EXACT_TILE_DOUBLES = 192;   // [6][32]: positions, and velocities for the FOF6d-filtered search
static __host__ __device__ inline size_t exact_warp_bytes(int kcap) { return (size_t)kcap * 32 * 12 + EXACT_TILE_DOUBLES * 8 + TRAV_STACK * 4; }

template <class S, bool FILTER, bool PHASE = false>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_exact_kernel(KnnParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const size_t warp_bytes = exact_warp_bytes(prm.kcap);
    unsigned char* base = smem_raw + w * warp_bytes;
    WarpHeap hp;
    hp.H = reinterpret_cast<double*>(base);
    hp.I = reinterpret_cast<int*>(base + (size_t)prm.kcap * 32 * 8);
    double* tile = reinterpret_cast<double*>(base + (size_t)prm.kcap * 32 * 12);
    int* stack = reinterpret_cast<int*>(base + (size_t)prm.kcap * 32 * 12 + EXACT_TILE_DOUBLES * 8);
    hp.k = prm.kcap; hp.lane = lane;

    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    // explicit query lists (the fast kernels' fallback) hold scattered particles: one query per warp (lane 0) keeps each
    // warp's traversal short instead of serialising 32 unrelated searches
    const int64_t nrows = prm.qlist ? prm.nq : (prm.q1 - prm.q0);
    const int64_t group = (int64_t)blockIdx.x * KNN_WARPS + w;
    const int64_t row = prm.qlist ? group : group * 32 + lane;
    if ((prm.qlist ? group : group * 32) >= nrows) return;                 // whole warp out of range
    bool valid = prm.qlist ? (lane == 0) : (row < nrows);
    const int64_t qi = valid ? (prm.qlist ? (int64_t)prm.qlist[row] : prm.q0 + row) : 0;
    if (valid && prm.active && prm.mode == 0 && !prm.qlist && !prm.active[qi]) valid = false;

    KnnVisitor<S, FILTER, PHASE> v;
    v.P = P; v.V = reinterpret_cast<const Vec4<S>*>(prm.V); v.tile = tile; v.hp = hp; v.lane = lane;
    v.excl = prm.cand_excl; v.crit_mode = prm.crit_mode; v.cp0 = prm.cp0; v.cp1 = prm.cp1;
    v.vx = v.vy = v.vz = 0;
    v.on = valid;
    v.self = -1; v.target_form = false;
    double x0 = 0, y0 = 0, z0 = 0;
    if (valid) {
        if (prm.mode == 0) {
            Vec4<S> c = P[qi];
            x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z;
            v.self = (int)qi;
            v.target_form = !prm.periodic;      // periodic particle searches use the coordinate form (KDSplitNode.cxx:1075-1080)
            if (PHASE || (FILTER && prm.crit_mode == 4)) { Vec4<S> u = v.V[qi]; v.vx = (double)u.x; v.vy = (double)u.y; v.vz = (double)u.z; }
        } else {
            x0 = prm.xq[3 * qi]; y0 = prm.xq[3 * qi + 1]; z0 = prm.xq[3 * qi + 2];
            if (PHASE || (FILTER && prm.crit_mode == 4)) { v.vx = prm.vq[3 * qi]; v.vy = prm.vq[3 * qi + 1]; v.vz = prm.vq[3 * qi + 2]; }
        }
    }
    v.qx = x0; v.qy = y0; v.qz = z0;
    v.hp.init();
    v.settop();
    {
        QueryBox qb = make_qbox(x0, y0, z0);
        traverse(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid);
        if (prm.nlo2) traverse(prm.nlo2, prm.nhi2, prm.bucket2, stack, v, qb, valid);
    }
    if (prm.periodic) {
        // reference image schedule: 3 faces, 3 edges, corner; each tested against the CURRENT top
        // (KDSplitNode.cxx:1125-1147, DistFunc.h:326-355)
        const double px = prm.period[0], py = prm.period[1], pz = prm.period[2];
        const double sx = (x0 < px / 2.0) ? x0 + px : x0 - px, ax = (x0 < px / 2.0) ? x0 : sx;   // ax: value squared in the 2D/ND tests
        const double sy = (y0 < py / 2.0) ? y0 + py : y0 - py, ay = (y0 < py / 2.0) ? y0 : sy;
        const double sz = (z0 < pz / 2.0) ? z0 + pz : z0 - pz, az = (z0 < pz / 2.0) ? z0 : sz;
        for (int img = 1; img <= 7; img++) {
            // order: x, y, z, xy, xz, yz, xyz
            const int mx = (img == 1 || img == 4 || img == 5 || img == 7);
            const int my = (img == 2 || img == 4 || img == 6 || img == 7);
            const int mz = (img == 3 || img == 5 || img == 6 || img == 7);
            bool go;
            if (img <= 3) {
                double sval = img == 1 ? ((x0 < px / 2.0) ? x0 : -sx) : (img == 2 ? ((y0 < py / 2.0) ? y0 : -sy) : ((z0 < pz / 2.0) ? z0 : -sz));
                go = sqrt(v.top) > sval;
            } else {
                double s2 = 0;
                if (mx) s2 = __dadd_rn(s2, __dmul_rn(ax, ax));
                if (my) s2 = __dadd_rn(s2, __dmul_rn(ay, ay));
                if (mz) s2 = __dadd_rn(s2, __dmul_rn(az, az));
                double sval = sqrt(s2);
                go = (prm.strict ? sqrt(v.top) : v.top) > sval;     // quirk Q1 unless strict
            }
            go = go && valid;
            if (!__any_sync(0xffffffffu, go)) continue;
            v.qx = mx ? sx : x0; v.qy = my ? sy : y0; v.qz = mz ? sz : z0;
            v.on = go;
            QueryBox qb = make_qbox(v.qx, v.qy, v.qz);
            traverse(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, go);
        }
        v.on = valid;
        v.qx = x0; v.qy = y0; v.qz = z0;
    }
    if (!valid) return;   // no warp-collective operations below this line

    // ------------------------------------------------------------------------------------ epilogues
    const int kc = prm.kcap;
    if (prm.hsm) prm.hsm[prm.gather ? row : qi] = 0.5 * sqrt(v.top);
    bool sorted = false;
    if (prm.rho && prm.veldens_k == 0 && prm.gather) {
        // CalcDensityParticle / CalcDensityPosition (KDCalcSmoothQuantities.cxx:768-844, 1092-1148): gather only, weight
        // 1.0 * W, no scatter.  The reference pops its heap, i.e. sums from the farthest neighbour inwards: same order here.
        const double hi = 0.5 * sqrt(v.top);
        const double norm = 1.0 / pow(hi, 3.0);
        const double delta = 2.0 / (double)(prm.kernres - 1);
        v.hp.sort_ascending(kc);
        sorted = true;
        double acc = 0;
        for (int s = kc - 1; s >= 0; s--) {
            int id = v.hp.i(s);
            if (id < 0) continue;
            double rij = sqrt(v.hp.h(s));
            double r = rij / hi;
            double Wij = wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
            acc += Wij * prm.mass[id];
        }
        prm.rho[row] = acc;
    }
    if (prm.rho && prm.veldens_k == 0 && !prm.gather) {
        // R1: CalcDensity (KDCalcSmoothQuantities.cxx:260-300), symmetric gather + scatter
        const double hi = 0.5 * sqrt(v.top);
        const double norm = 1.0 / pow(hi, 3.0);
        const double delta = 2.0 / (double)(prm.kernres - 1);
        const double mi = prm.mass[qi];
        double acc = 0;
        for (int s = 0; s < kc; s++) {
            int id = v.hp.i(s);
            if (id < 0) continue;
            double rij = sqrt(v.hp.h(s));
            double r = rij / hi;
            double Wij = 0.5 * wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
            acc += Wij * prm.mass[id];
            atomicAdd(&prm.rho[id], Wij * mi);
        }
        atomicAdd(&prm.rho[qi], acc);
    }
    if (prm.smvel_out || prm.smdisp_out || prm.smhigh_out) {
        // CalcSmoothVel / CalcSmoothVelDisp (KDCalcSmoothQuantities.cxx:480-614): symmetric gather + scatter with weights
        // 0.5 * W(r_ij, h_i) * m / rho of the CONTRIBUTING particle; the dispersion is taken about the smoothed mean velocity
        // of the RECEIVING particle (:594-611)
        const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
        const double hi = 0.5 * sqrt(v.top);
        const double norm = 1.0 / pow(hi, 3.0);
        const double delta = 2.0 / (double)(prm.kernres - 1);
        const Vec4<S> vq4 = V[qi];
        const double vi[3] = {(double)vq4.x, (double)vq4.y, (double)vq4.z};
        const double wi = prm.mass[qi] / prm.rho_in[qi];        // temp = Wij / rho * m, evaluated as Wij * (m / rho): one extra rounding
        double mi[3] = {0, 0, 0};
        if (prm.smdisp_out || prm.smhigh_out) { mi[0] = prm.smvel_in[3 * qi]; mi[1] = prm.smvel_in[3 * qi + 1]; mi[2] = prm.smvel_in[3 * qi + 2]; }
        // CalcSmoothVelSkew / CalcSmoothVelKurtosis (KDCalcSmoothQuantities.cxx:617-765): per component k the third / fourth power of
        // (v - smoothed mean of the RECEIVER) over the receiver's dispersion sigma_kk^1.5 / sigma_kk^2; the kurtosis form subtracts 3
        // from EVERY contribution (:745,751), which is reproduced
        double si[3] = {1, 1, 1};
        if (prm.smhigh_out) {
            for (int a = 0; a < 3; a++) { const double d = prm.smdisp_in[9 * qi + 4 * a]; si[a] = prm.moment == 3 ? pow(d, 1.5) : d * d; }
        }
        double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < kc; s++) {
            int id = v.hp.i(s);
            if (id < 0) continue;
            double rij = sqrt(v.hp.h(s));
            double r = rij / hi;
            double Wij = 0.5 * wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
            const Vec4<S> vj4 = V[id];
            const double vj[3] = {(double)vj4.x, (double)vj4.y, (double)vj4.z};
            const double tj = Wij * (prm.mass[id] / prm.rho_in[id]);
            const double ti = Wij * wi;
            if (prm.smvel_out) {
                for (int a = 0; a < 3; a++) { acc[a] += tj * vj[a]; atomicAdd(&prm.smvel_out[3 * (int64_t)id + a], ti * vi[a]); }
            } else if (prm.smhigh_out) {
                for (int a = 0; a < 3; a++) {
                    const double dj = vj[a] - mi[a], di = vi[a] - prm.smvel_in[3 * (int64_t)id + a];
                    const double dd = prm.smdisp_in[9 * (int64_t)id + 4 * a];
                    if (prm.moment == 3) {
                        acc[a] += tj * dj * dj * dj / si[a];
                        atomicAdd(&prm.smhigh_out[3 * (int64_t)id + a], ti * di * di * di / pow(dd, 1.5));
                    } else {
                        acc[a] += tj * dj * dj * dj * dj / si[a] - 3.;
                        atomicAdd(&prm.smhigh_out[3 * (int64_t)id + a], ti * di * di * di * di / (dd * dd) - 3.);
                    }
                }
            } else {
                const double mj[3] = {prm.smvel_in[3 * (int64_t)id], prm.smvel_in[3 * (int64_t)id + 1], prm.smvel_in[3 * (int64_t)id + 2]};
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++) {
                        acc[3 * a + b] += tj * (vj[a] - mi[a]) * (vj[b] - mi[b]);
                        atomicAdd(&prm.smdisp_out[9 * (int64_t)id + 3 * a + b], ti * (vi[a] - mj[a]) * (vi[b] - mj[b]));
                    }
            }
        }
        if (prm.smvel_out) { for (int a = 0; a < 3; a++) atomicAdd(&prm.smvel_out[3 * qi + a], acc[a]); }
        else if (prm.smhigh_out) { for (int a = 0; a < 3; a++) atomicAdd(&prm.smhigh_out[3 * qi + a], acc[a]); }
        else { for (int a = 0; a < 9; a++) atomicAdd(&prm.smdisp_out[9 * qi + a], acc[a]); }
    }
    if (prm.rho && prm.veldens_k > 0) {
        // R2: CalcVelDensity (KDCalcSmoothQuantities.cxx:335-383); the *Particle / *Position forms (:845-921, :1150-1207)
        // compute the same number for one target (gather: result per row; point form: query velocity from vq)
        const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
        double vix, viy, viz;
        if (prm.mode == 1) { vix = prm.vq[3 * qi]; viy = prm.vq[3 * qi + 1]; viz = prm.vq[3 * qi + 2]; }
        else { Vec4<S> vi = V[qi]; vix = (double)vi.x; viy = (double)vi.y; viz = (double)vi.z; }
        int kx = 0;
        for (int s = 0; s < kc; s++) {
            int id = v.hp.i(s);
            if (id < 0) continue;
            Vec4<S> vj = V[id];
            double vd = sqrt(dist2_ref(vix, viy, viz, (double)vj.x, (double)vj.y, (double)vj.z));
            v.hp.h(kx) = vd; v.hp.i(kx) = id; kx++;
        }
        int kv = min(prm.veldens_k, kx);
        double rho = 0;
        if (kv > 0) {
            v.hp.heapify(kv);
            for (int s = kv; s < kx; s++) {
                double vd = v.hp.h(s);
                if (vd < v.hp.h(0)) v.hp.sift_down(0, kv, vd, v.hp.i(s));
            }
            const double hi = 0.5 * v.hp.h(0);
            const double norm = 1.0 / pow(hi, 3.0);
            const double delta = 2.0 / (double)(prm.kernres - 1);
            // pop in descending order like the reference so the sum is accumulated in the same order
            for (int e = kv; e > 0; e--) {
                double rij = v.hp.h(0);
                double r = rij / hi;
                rho = rho + wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
                double dl = v.hp.h(e - 1); int il = v.hp.i(e - 1);
                v.hp.sift_down(0, e - 1, dl, il);
            }
        }
        prm.rho[prm.gather ? row : qi] = rho;
    }
    if (prm.nn || prm.d2out) {
        if (!sorted) v.hp.sort_ascending(kc);
        // periodic particle searches carry k+1 slots: FindNearestPos(tt) drops the farthest, FindNearest(tt) the nearest (Q3)
        const int off = (prm.kcap > prm.k && prm.tree_form) ? 1 : 0;
        const int64_t orow = row * (int64_t)prm.k;
        for (int j = 0; j < prm.k; j++) {
            int id = v.hp.i(j + off);
            if (prm.nn) prm.nn[orow + j] = (prm.out_ids && id >= 0) ? prm.order[id] : id;
            if (prm.d2out) prm.d2out[orow + j] = v.hp.h(j + off);
        }
    }
}

// ========================================================================= density kernel: heap of packed words
// The density family needs (a) the exact set of the k nearest, (b) the exact k-th distance and (c) a sum over the set.
// Per lane (= query) the kernel keeps the k+2 smallest candidates seen so far as 32-bit WORDS in a 4-ary max-heap in shared
// memory:
//     word = (fp32 key & ~0x1fff) | candidate id,     candidate id = (tile sequence number << 6) | slot in the tile
// (the warp's tile list maps a sequence number back to a tree index; words order like their keys, so the heap needs no
// separate index array and the kernel no global scratch).  A tile is screened with ONE fp32 distance and one unsigned
// comparison per candidate into a bit mask; the set bits are inserted in SIMT-serial rounds.  After the traversal the full
// keys of the k+2 survivors are recomputed from their indices.
//
// Exactness.  Full keys are a = fl32(d2) evaluated with 3 subtractions, 1 multiplication, 2 FMAs on exact fp32
// coordinates: |a - d2| <= 5.01 * 2^-24 * d2 =: eps * d2 (fp64 storage: a = RN_fp32 of the reference's fp64 d2, eps = 2^-24).
//   * A candidate screened out (a > edge * (1 + 2^-20), edge = upper edge of the root word's 2^-10-wide bin) or lying in a
//     subtree whose box lower bound reaches that limit is farther, in exact arithmetic, than every word the heap holds then
//     and later (2^-20 > 2 eps).
//   * A candidate that reaches a round and is rejected, or a root that is evicted, lies in the root's bin or above; the lane
//     remembers the smallest such word (`mindrop`).
//   * Among the k+2 survivors let m1 be the (k+1)-th smallest full key and m2 the k-th.  If m1 > m2 * (1 + 2^-20) -- and, when
//     some dropped word shares the final root's bin, if that bin's lower edge exceeds m2 * (1 + 2^-20) as well (the root is
//     the (k+2)-th word, so this fails only when three consecutive keys fall into one 2^-10-wide bin) -- every dropped
//     candidate and the two extra survivors are farther in exact arithmetic than each of the k kept ones: the k smallest
//     full keys ARE the k nearest.
//   * The exact k-th distance is the reference fp64 d2 of the entry with key m2 provided the next smaller key is below
//     m2 * (1 - 2^-20); the SPH sums use fp64 d2 recomputed from the indices.
//   * Queries failing a gap test (~4e-4 of them), meeting an underflowing key or more than 128 tiles are appended to a list
//     and re-run by knn_exact_kernel (fp64 heap).
// Traversal: bottom-up (traverse_bottom_up): the group's own node first, then the sibling subtrees of its ancestors.
// Persistent grid: warps draw 32-query groups from a global counter, so the grid is one wave whatever the particle count.
constexpr float AP_TINY = 7.888609052210118e-31f;           // 2^-100: below it the relative error bound of a key is not guaranteed
constexpr float AP_HUGE = 1.0e37f;                           // launch_knn sends boxes that could exceed it to the exact kernel
constexpr float AP_WIDEN = 1.00000095367431640625f;          // 1 + 2^-20
constexpr float AP_NARROW = 0.99999904632568359375f;         // 1 - 2^-20
// A staged tile is a whole node of the scanned tree level: 32 slots when that level's nodes hold at most 32 particles (always
// so for 2^m particles), else 40 (the level holds 21..40).  Candidate id: 7 bits tile sequence number + 5 / 6 bits slot.
template <int TILE> struct ApWord {
    static constexpr int SLOT_BITS = TILE <= 32 ? 5 : 6;
    static constexpr unsigned CIDMASK = (1u << (6 + SLOT_BITS)) - 1u;      // candidate id = (tile sequence number, 6 bits: AP_MAXTILES) , slot
    static constexpr unsigned KEYMASK = ~CIDMASK;
};
constexpr int AP_MAXTILES = 64;
constexpr float AP_INF = __builtin_huge_valf();

#ifdef NBK_STATS
__device__ unsigned long long g_stats[8];   // 0 tiles, 2 insertion rounds, 3 candidates past the screen (lane-level), 4 flagged, 6 tile list overflows
#define STAT(i, v) do { if (lane_id() == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#define STAT_LANE(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define STAT(i, v)
#define STAT_LANE(i, v)
#endif

// the key of a candidate: the same expression in the tile scan and in the final pass over the survivors
__device__ __forceinline__ float ap_key(float qx, float qy, float qz, float cx, float cy, float cz) {
    const float dx = __fsub_rn(qx, cx), dy = __fsub_rn(qy, cy), dz = __fsub_rn(qz, cz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}
__device__ __forceinline__ float ap_key(double qx, double qy, double qz, double cx, double cy, double cz) {
    return __double2float_rn(dist2_ref(qx, qy, qz, cx, cy, cz));
}

// Leaf tile staged in shared memory.
template <class S, int AP_TILE> struct ApTile;
template <int AP_TILE> struct ApTile<float, AP_TILE> {
    static constexpr int TILE_BYTES = AP_TILE * 16;
    float4* t;
    float qx, qy, qz;
    __device__ __forceinline__ void init(void* mem, double x, double y, double z) {
        t = reinterpret_cast<float4*>(mem);
        qx = (float)x; qy = (float)y; qz = (float)z;     // exact: fp32 storage
    }
    __device__ __forceinline__ void load(const Vec4<float>* __restrict__ P, int first, int m, unsigned lane) {
        // slots past the end of the leaf hold NaN: every comparison on them is false, so the scan runs over whole groups of 8
        const float nanf_ = __int_as_float(0x7fc00000);
        float4 mine = make_float4(nanf_, nanf_, nanf_, 0.f), more = mine;
        if ((int)lane < m) { Vec4<float> c = P[first + lane]; mine = make_float4(c.x, c.y, c.z, 0.f); }
        if (AP_TILE > 32 && 32 + (int)lane < m) { Vec4<float> c = P[first + 32 + lane]; more = make_float4(c.x, c.y, c.z, 0.f); }
        t[lane] = mine;
        if (AP_TILE > 32 && (int)lane < AP_TILE - 32) t[32 + lane] = more;
    }
    static constexpr bool TRANSPOSABLE = true;
    __device__ __forceinline__ float key(int j) const { const float4 c = t[j]; return ap_key(qx, qy, qz, c.x, c.y, c.z); }
    __device__ __forceinline__ float key_of(const Vec4<float>& c) const { return ap_key(qx, qy, qz, c.x, c.y, c.z); }
    // key of lane q's query against candidate j of the staged tile (all lanes call it with the same q)
    __device__ __forceinline__ unsigned key_bits_for(int q, int j) const {
        const float4 c = t[j];
        const float x = __shfl_sync(0xffffffffu, qx, q), y = __shfl_sync(0xffffffffu, qy, q), z = __shfl_sync(0xffffffffu, qz, q);
        return __float_as_uint(ap_key(x, y, z, c.x, c.y, c.z));
    }
    __device__ __forceinline__ bool coincident(int j) const { const float4 c = t[j]; return qx == c.x && qy == c.y && qz == c.z; }
};
template <int AP_TILE> struct ApTile<double, AP_TILE> {
    static constexpr int TILE_BYTES = 3 * AP_TILE * 8;
    double* t;
    double qx, qy, qz;
    __device__ __forceinline__ void init(void* mem, double x, double y, double z) { t = reinterpret_cast<double*>(mem); qx = x; qy = y; qz = z; }
    __device__ __forceinline__ void load(const Vec4<double>* __restrict__ P, int first, int m, unsigned lane) {
        const double nan_ = __longlong_as_double(0x7ff8000000000000ll);
        double cx = nan_, cy = nan_, cz = nan_;
        if ((int)lane < m) { Vec4<double> c = P[first + lane]; cx = c.x; cy = c.y; cz = c.z; }
        t[lane] = cx; t[AP_TILE + lane] = cy; t[2 * AP_TILE + lane] = cz;
        if (AP_TILE > 32 && (int)lane < AP_TILE - 32) {
            cx = cy = cz = nan_;
            if (32 + (int)lane < m) { Vec4<double> c = P[first + 32 + lane]; cx = c.x; cy = c.y; cz = c.z; }
            t[32 + lane] = cx; t[AP_TILE + 32 + lane] = cy; t[2 * AP_TILE + 32 + lane] = cz;
        }
    }
    static constexpr bool TRANSPOSABLE = false;
    __device__ __forceinline__ float key(int j) const { return ap_key(qx, qy, qz, t[j], t[AP_TILE + j], t[2 * AP_TILE + j]); }
    __device__ __forceinline__ float key_of(const Vec4<double>& c) const { return ap_key(qx, qy, qz, c.x, c.y, c.z); }
    __device__ __forceinline__ unsigned key_bits_for(int, int) const { return 0u; }
    __device__ __forceinline__ bool coincident(int j) const { return qx == t[j] && qy == t[AP_TILE + j] && qz == t[2 * AP_TILE + j]; }
};

// SPH epilogues over a lane's neighbour list: entry s (0 <= s < cnt) is tree index idx_at(s).  All 32 lanes call it (`on`:
// this lane has a result to compute); D: per-lane doubles [k][32] in shared memory for the kv < kx velocity-density
// selection -- it may overlap the storage idx_at reads, which is consumed from the last entry down, row by row.
template <class S, class IdxAt>
__device__ __forceinline__ void sc_epilogue(const KnnParams& prm, const Vec4<S>* __restrict__ P, IdxAt idx_at, double* Dbase, unsigned lane, bool on,
                                            int cnt, double d2max, double x0, double y0, double z0, int64_t qi) {
    if (on && prm.hsm) prm.hsm[qi] = 0.5 * sqrt(d2max);
    if (on && prm.rho && prm.veldens_k == 0) {
        const double hi = 0.5 * sqrt(d2max);
        const double norm = 1.0 / pow(hi, 3.0);
        const double delta = 2.0 / (double)(prm.kernres - 1);
        const double mi = prm.mass[qi];
        const double inv_hi = 1.0 / hi, inv_delta = 1.0 / delta, half_res = 0.5 * (prm.kernres - 1), half_norm = 0.5 * norm;
        double acc = 0;
        for (int s0 = 0; s0 < cnt; s0 += 4) {
            // four neighbours per trip: the gathers of P and mass are issued together
            int id[4]; Vec4<S> c[4]; double mj[4];
#pragma unroll
            for (int u = 0; u < 4; u++) id[u] = (s0 + u < cnt) ? idx_at(s0 + u) : -1;
#pragma unroll
            for (int u = 0; u < 4; u++) { const int ci = id[u] >= 0 ? id[u] : 0; c[u] = P[ci]; mj[u] = prm.mass[ci]; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (id[u] >= 0) {
                    const double d2 = dist2_ref(x0, y0, z0, (double)c[u].x, (double)c[u].y, (double)c[u].z);
                    // r = rij / hi with one rounding more than the reference's division; the k-th neighbour sits at exactly 2
                    const double r = d2 == d2max ? 2.0 : sqrt(d2) * inv_hi;
                    const double Wij = wsm_fast(r, (int)(r * half_res), prm.kernres, delta, inv_delta, prm.kern) * half_norm;
                    acc += Wij * mj[u];
                    atomicAdd(&prm.rho[id[u]], Wij * mi);
                }
            }
        }
        atomicAdd(&prm.rho[qi], acc);
    }
    if (prm.rho && prm.veldens_k > 0) {
        const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
        Vec4<S> vi = on ? V[qi] : Vec4<S>();
        const double delta = 2.0 / (double)(prm.kernres - 1);
        const int kv = on ? min(prm.veldens_k, cnt) : 0;
        double rho = 0;
        if (prm.veldens_k >= prm.k) {
            // every spatial neighbour is used: h from the largest velocity distance, then the sum (two passes)
            if (on && kv > 0) {
                double vmax = 0;
                for (int s = 0; s < cnt; s++) {
                    Vec4<S> vj = V[idx_at(s)];
                    vmax = fmax(vmax, sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z)));
                }
                const double hi = 0.5 * vmax;
                const double norm = 1.0 / pow(hi, 3.0);
                for (int s = 0; s < cnt; s++) {
                    Vec4<S> vj = V[idx_at(s)];
                    double r = sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z)) / hi;
                    rho = rho + wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
                }
            }
        } else {
            // kv < kx: exact fp64 selection on a per-lane array of doubles in shared memory.  The array overlaps the list's own
            // storage: row s of the doubles covers rows 2s, 2s+1 of the list, so the list is consumed from its last row down,
            // all lanes in step.
            double* D = Dbase + lane;
            const int nmax = __reduce_max_sync(0xffffffffu, on ? cnt : 0);
            for (int s = nmax - 1; s >= 0; s--) {
                const bool in = on && s < cnt;
                double vd = 0;
                if (in) {
                    Vec4<S> vj = V[idx_at(s)];
                    vd = sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z));
                }
                __syncwarp();
                if (in) D[s * 32] = vd;
                __syncwarp();
            }
            auto dsift = [&](int p, int n, double d) {
                while (true) {
                    int c = 2 * p + 1;
                    if (c >= n) break;
                    double dc = D[c * 32];
                    if (c + 1 < n) { double dr = D[(c + 1) * 32]; if (dr > dc) { c = c + 1; dc = dr; } }
                    if (d >= dc) break;
                    D[p * 32] = dc;
                    p = c;
                }
                D[p * 32] = d;
            };
            if (kv > 0) {
                for (int p = kv / 2 - 1; p >= 0; p--) dsift(p, kv, D[p * 32]);
                for (int s = kv; s < cnt; s++) {
                    double vd = D[s * 32];
                    if (vd < D[0]) dsift(0, kv, vd);
                }
                const double hi = 0.5 * D[0];
                const double norm = 1.0 / pow(hi, 3.0);
                for (int e = kv; e > 0; e--) {
                    double r = D[0] / hi;
                    rho = rho + wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
                    dsift(0, e - 1, D[(e - 1) * 32]);
                }
            }
        }
        if (on) prm.rho[qi] = rho;
    }
}

// Per-lane 4-ary max-heap of words.  The ROOT lives in a register (it is the lane's bound anyway); nodes 1 .. 4G live in shared
// memory: the four children of node p are the components of ONE 16-byte group at kb + p*512 (one LDS.128 per level, 3 levels for
// 66 entries), node p >= 1 at kb + ((p-1)>>2)*512 + ((p-1)&3)*4.  Gs*512 bytes per warp, Gs = ceil((k+2)/4).
struct WordHeap4 {
    unsigned char* kb;   // this lane's byte base inside the [Gs][32] float4 groups
    int G;               // nodes 0 .. G-1 have children
    __device__ __forceinline__ float* nodep(int p) const { return reinterpret_cast<float*>(kb + ((p - 1) >> 2) * 512 + ((p - 1) & 3) * 4); }
    // place xk at node p >= 1 (overwriting it) and sift it down
    __device__ __forceinline__ void sift(int p, float xk) {
        unsigned char* pa = reinterpret_cast<unsigned char*>(nodep(p));
        while (p < G) {
            unsigned char* ga = kb + p * 512;
            const float4 ck = *reinterpret_cast<const float4*>(ga);
            const bool a = ck.x >= ck.y, b = ck.z >= ck.w;
            const float m01 = a ? ck.x : ck.y, m23 = b ? ck.z : ck.w;
            const bool c = m01 >= m23;
            const float m = c ? m01 : m23;
            if (xk >= m) break;
            const int j = c ? (a ? 0 : 1) : (b ? 2 : 3);
            *reinterpret_cast<float*>(pa) = m;
            pa = ga + 4 * j;
            p = 4 * p + 1 + j;
        }
        *reinterpret_cast<float*>(pa) = xk;
    }
    // xk takes the place of the root (which is dropped); returns the new root
    __device__ __forceinline__ float replace_root(float xk) {
        const float4 ck = *reinterpret_cast<const float4*>(kb);
        const bool a = ck.x >= ck.y, b = ck.z >= ck.w;
        const float m01 = a ? ck.x : ck.y, m23 = b ? ck.z : ck.w;
        const bool c = m01 >= m23;
        const float m = c ? m01 : m23;
        if (xk >= m) return xk;
        sift(1 + (c ? (a ? 0 : 1) : (b ? 2 : 3)), xk);       // the largest child moves up into the register, xk sinks from its slot
        return m;
    }
};

template <class S, int AP_TILE>
struct HpVisitor {
    static constexpr unsigned AP_CIDMASK = ApWord<AP_TILE>::CIDMASK, AP_KEYMASK = ApWord<AP_TILE>::KEYMASK;
    const Vec4<S>* P;
    ApTile<S, AP_TILE> tile;
    WordHeap4 hp;
    int* tl;              // shared: the warp's tile list (tile sequence number -> first tree index)
    int nt;               // tiles scanned so far (warp-uniform)
    float topw;           // heap root (node 0 lives here, not in shared memory): the lane's (k+2)-th smallest word so far (+inf while it has fewer)
    float limf;           // screen / node-test limit: upper edge of the root's bin * (1 + 2^-20); -1: the lane accepts nothing
    unsigned limu;        // the same limit for the unsigned screen test (bits(a) - bits(AP_TINY) < limu); 0: nothing passes
    unsigned mindrop;     // smallest word rejected or evicted in a round once the heap was full
    int filled, kcap;     // heap slots taken so far (the first k+2 candidates are stored without sifting), k + 2
    int tr_max;           // tiles needed by at most this many lanes are screened in the transposed form
    bool failed;
    unsigned lane;
    static constexpr unsigned TINY_BITS = 0x0d800000u;      // bits of AP_TINY = 2^-100
    __device__ __forceinline__ bool need(float lb) const { return lb < limf; }
    __device__ __forceinline__ void set_limit(float lim) { limf = lim; limu = __float_as_uint(lim) - TINY_BITS + 1u; }
    __device__ __forceinline__ void settop(float w) {
        topw = w;
        set_limit(w < AP_HUGE ? fminf(__fmul_ru(__uint_as_float(__float_as_uint(w) | AP_CIDMASK), AP_WIDEN), AP_HUGE) : AP_HUGE);
    }
    __device__ __forceinline__ void fail() { failed = true; limf = -1.f; limu = 0u; }
    __device__ __forceinline__ void leaf(int start, int n, int, unsigned nmask) {
        for (int base = 0; base < n; base += AP_TILE) {
            const int m = min(AP_TILE, n - base);
            if (nt >= AP_MAXTILES) { if (limu != 0u) { STAT(6, 1); fail(); } return; }      // tile list full: the whole group goes to the exact kernel
            __syncwarp();
            tile.load(P, start + base, m, lane);
            if (lane == 0) tl[nt] = start + base;
            __syncwarp();
            const unsigned cid0 = (unsigned)nt << ApWord<AP_TILE>::SLOT_BITS;
            nt++;
            STAT(0, 1);
            // pass 1: one unsigned comparison per candidate (NaN padding, keys below 2^-100 -- the query itself, coincident
            // particles, underflow -- and keys above the limit all fail it).  acc: slots 0..31, acch: slots 32..39
            unsigned acc = 0u, acch = 0u, tiny = 0u;
            if (ApTile<S, AP_TILE>::TRANSPOSABLE && __popc(nmask) <= tr_max) {
                // few lanes need this tile: lane j takes candidate j (and 32 + j) and the needing queries are broadcast one at a
                // time, so a query costs one key per lane instead of 32 (same expression, same bits as the direct form)
                unsigned rem = nmask;
                while (rem) {
                    const int q = __ffs(rem) - 1;
                    rem &= rem - 1u;
                    const unsigned lq = __shfl_sync(0xffffffffu, limu, q);
                    const unsigned u = tile.key_bits_for(q, lane) - TINY_BITS;
                    unsigned mk = __ballot_sync(0xffffffffu, u < lq), tn = __ballot_sync(0xffffffffu, (int)u < 0), mkh = 0u;
                    if (AP_TILE > 32 && m > 32) {
                        const unsigned u2 = tile.key_bits_for(q, 32 + (lane & 7)) - TINY_BITS;
                        mkh = __ballot_sync(0xffffffffu, (int)lane < AP_TILE - 32 && u2 < lq);
                        tn |= __ballot_sync(0xffffffffu, (int)lane < AP_TILE - 32 && (int)u2 < 0);
                    }
                    if ((int)lane == q) { acc = mk; acch = mkh; tiny = tn ? 0x80000000u : 0u; }
                }
            } else {
                for (int j0 = 0; j0 < m; j0 += 8) {
                    float a[8];
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) a[jj] = tile.key(j0 + jj);
                    unsigned a8 = 0u;
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) {
                        const unsigned u = __float_as_uint(a[jj]) - TINY_BITS;
                        tiny |= u;
                        if (u < limu) a8 |= 1u << jj;
                    }
                    if (j0 < 32) acc |= a8 << j0; else acch = a8;
                }
            }
            if ((int)tiny < 0 && limu != 0u) {
                // rare: the query itself and coincident particles (never neighbours), or a distance that underflows fp32
#pragma unroll 1
                for (int j = 0; j < m; j++) {
                    const float a = tile.key(j);
                    if (a < AP_TINY && !tile.coincident(j)) fail();
                }
                if (failed) { acc = 0u; acch = 0u; }
            }
            STAT_LANE(3, __popc(acc) + __popc(acch));
            // pass 2: SIMT-serial insertion rounds over the set bits.  The first k+1 candidates of a lane fill the heap's slots
            // directly (the bound is still infinite) and are heapified together when the last slot is taken.
            while (__any_sync(0xffffffffu, (acc | acch) != 0u)) {
                STAT(2, 1);
                if (acc | acch) {
                    int j;
                    if (acc) { j = __ffs(acc) - 1; acc &= acc - 1u; }
                    else { j = 31 + __ffs(acch); acch &= acch - 1u; }
                    const unsigned wb = (__float_as_uint(tile.key(j)) & AP_KEYMASK) | (cid0 + (unsigned)j);
                    const float w = __uint_as_float(wb);
                    if (filled < kcap) {
                        // candidates 1 .. k+1 go to nodes 1 .. k+1 as they come; the (k+2)-th completes the heap: the inner nodes
                        // are heapified and the newcomer is sifted in from the root
                        filled++;
                        if (filled < kcap) *hp.nodep(filled) = w;
                        else {
                            for (int p = hp.G - 1; p >= 1; p--) hp.sift(p, *hp.nodep(p));
                            settop(hp.replace_root(w));
                        }
                    } else {
                        // the word that does not stay in the heap: the candidate itself, or the root it evicts
                        const unsigned out = w < topw ? __float_as_uint(topw) : wb;
                        mindrop = min(mindrop, out);
                        if (w < topw) settop(hp.replace_root(w));
                    }
                }
            }
        }
    }
};

// One CTA per SM with as many warps as registers (96 per thread: 21 warps) and shared memory (the per-warp heaps) allow: the
// warps never synchronise with each other, so the CTA size is free, and one large CTA wastes no per-CTA reserved shared memory
// (5 CTAs x 4 warps left 13 KB per SM unused -- one warp's worth).
constexpr int HP_MAX_WARPS = 24;
constexpr int HP_STACK = 32;       // traversal stack entries of this kernel: at most one per tree level (depth <= 30)
static inline int hp_store_groups(int k) { return (k + 2 + 3) / 4; }   // 16-byte groups per lane: nodes 1 .. k+2 (the last one is the root's slot of the final pass)
static inline size_t hp_warp_bytes(int k, bool want_doubles, int tile_bytes) {
    size_t region = (size_t)hp_store_groups(k) * 512;
    if (want_doubles) region += (size_t)k * 32 * 8;
    return region + tile_bytes + HP_STACK * 4 + AP_MAXTILES * 4;
}

template <class S, bool HALO, int AP_TILE>
__global__ void __launch_bounds__(HP_MAX_WARPS * 32, 1) knn_hp_kernel(KnnParams prm, int want_doubles, int* __restrict__ work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned full = 0xffffffffu;
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    // the heap keeps k + 2 words: the k nearest, the (k+1)-th that certifies them, and one more so that a candidate dropped
    // from the root's bin almost never sits next to the k-th (see the certificate below)
    const int k = prm.k, kcap = k + 2;
    const int G = (kcap - 1 + 3) / 4;                      // nodes 0 .. G-1 have children (the last group may be partly padding)
    const int Gs = (kcap + 3) / 4;                         // stored groups: nodes 1 .. 4 Gs, 4 Gs >= kcap
    const size_t heap_bytes = (size_t)Gs * 512;
    const size_t region = heap_bytes + (want_doubles ? (size_t)k * 32 * 8 : 0);
    const size_t warp_bytes = region + ApTile<S, AP_TILE>::TILE_BYTES + HP_STACK * 4 + AP_MAXTILES * 4;
    unsigned char* base = smem_raw + w * warp_bytes;
    void* tile_mem = base + region;
    int* stack = reinterpret_cast<int*>(base + region + ApTile<S, AP_TILE>::TILE_BYTES);
    int* tl = stack + HP_STACK;
    unsigned char* kb = base + lane * 16;

    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const int64_t ngroups = (prm.q1 - prm.q0 + 31) >> 5;
    while (true) {
        int64_t group = 0;
        if (lane == 0) group = (int64_t)atomicAdd(work_counter, 1);
        group = __shfl_sync(full, group, 0);
        if (group >= ngroups) break;
        const int64_t g0 = prm.q0 + group * 32;
        const int64_t g1 = g0 + 32 < prm.q1 ? g0 + 32 : prm.q1;
        const int64_t qi = g0 + lane;
        const bool valid = qi < prm.q1 && (!prm.active || prm.active[qi]);
        if (!__any_sync(full, valid)) continue;
        double x0 = 0, y0 = 0, z0 = 0;
        if (valid) { Vec4<S> c = P[qi]; x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z; }
        const QueryBox qb = make_qbox(x0, y0, z0);

        // ------------------------------------------------------------------------------------------ select
        constexpr unsigned AP_KEYMASK = ApWord<AP_TILE>::KEYMASK;
        HpVisitor<S, AP_TILE> v;
        v.P = P; v.lane = lane;
        v.tile.init(tile_mem, x0, y0, z0);
        v.hp.kb = kb; v.hp.G = G;
        v.tl = tl; v.nt = 0;
        v.mindrop = 0xffffffffu;
        v.filled = 0; v.kcap = kcap; v.tr_max = prm.tr_max;
        v.failed = false;
        // the padding nodes of the last child group hold 0 (never the largest child); real nodes are written before they are read
        for (int p = kcap; p <= 4 * Gs; p++) *v.hp.nodep(p) = 0.f;
        v.topw = AP_INF;
        if (valid) v.set_limit(AP_HUGE); else { v.limf = -1.f; v.limu = 0u; }
        traverse_bottom_up(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid, prm.n_tree, prm.aligned, g0, g1);
        if (HALO) traverse(prm.nlo2, prm.nhi2, prm.bucket2, stack, v, qb, valid);

        // ------------------------------------------------------------------------ the k nearest and the k-th
        // the candidates a lane holds: slot s = node s + 1.  Heap complete: nodes 1 .. k+1 and the root, which is written to the
        // slot behind them; otherwise (fewer than k + 2 candidates exist) nodes 1 .. filled, never heapified
        auto word_at = [kb](int s) -> unsigned* { return reinterpret_cast<unsigned*>(kb + (s >> 2) * 512 + (s & 3) * 4); };
        const unsigned rootw = __float_as_uint(v.topw);
        int cnt = 0;
        if (valid && !v.failed) {
            cnt = v.filled;
            if (cnt == kcap) *word_at(kcap - 1) = rootw;
        }
        auto idx_at = [kb, tl](int s) -> int {
            const unsigned wd = *reinterpret_cast<const unsigned*>(kb + (s >> 2) * 512 + (s & 3) * 4);
            return tl[(wd >> ApWord<AP_TILE>::SLOT_BITS) & (AP_MAXTILES - 1)] + (int)(wd & ((1u << ApWord<AP_TILE>::SLOT_BITS) - 1u));
        };
        // full keys of the survivors: the four largest f0 >= f1 >= f2 >= f3 and the slots of the first three
        float f0 = -1.f, f1 = -1.f, f2 = -1.f, f3 = -1.f;
        int t0 = -1, t1 = -1, t2 = -1;
        {
            const int nmax = __reduce_max_sync(full, cnt);
            for (int s0 = 0; s0 < nmax; s0 += 4) {
                Vec4<S> c[4];
#pragma unroll
                for (int u = 0; u < 4; u++) c[u] = P[s0 + u < cnt ? idx_at(s0 + u) : 0];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (s0 + u < cnt) {
                        const float a = v.tile.key_of(c[u]);
                        const int s = s0 + u;
                        const bool g0_ = a > f0, g1_ = a > f1, g2_ = a > f2, g3_ = a > f3;
                        f3 = g2_ ? f2 : (g3_ ? a : f3);
                        f2 = g1_ ? f1 : (g2_ ? a : f2);  t2 = g1_ ? t1 : (g2_ ? s : t2);
                        f1 = g0_ ? f0 : (g1_ ? a : f1);  t1 = g0_ ? t0 : (g1_ ? s : t1);
                        f0 = g0_ ? a : f0;               t0 = g0_ ? s : t0;
                    }
                }
            }
        }
        bool flagged = v.failed;
        const int drop = cnt > k ? cnt - k : 0;             // survivors beyond the k nearest: 2 when the heap is full
        float kth = f0, above = AP_INF, below = f1;
        int skth = t0;
        if (drop == 1) { above = f0; kth = f1; below = f2; skth = t1; }
        if (drop == 2) { above = f1; kth = f2; below = f3; skth = t2; }
        const int nk = cnt - drop;
        const bool short_of_k = nk < k;                     // fewer than k candidates exist: the reference's heap keeps sentinels
        if (valid && !flagged) {
            const float kth_band = __fmul_ru(kth, AP_WIDEN);
            // the (k+1)-th full key certifies the set if it lies above the k-th by more than the error band ...
            if (drop >= 1 && !(above > kth_band)) flagged = true;
            // ... and so must every dropped candidate: those dropped in a round lie in the root's bin or above, and if one shares
            // the final root's bin it may lie anywhere in it, so the k-th has to stay below that bin's lower edge
            if (drop == 2 && (v.mindrop & AP_KEYMASK) <= (rootw & AP_KEYMASK) && !(__uint_as_float(rootw & AP_KEYMASK) > kth_band)) flagged = true;
            if (!short_of_k && nk >= 2 && !(below < __fmul_rd(kth, AP_NARROW))) flagged = true;
        }
        if (valid && !flagged && drop >= 1) {
            // close the slots of the dropped survivors with the last entries
            int last = cnt - 1;
            int da = t0, db = drop == 2 ? t1 : -1;           // slots to drop
            if (db > da) { const int t = da; da = db; db = t; }      // larger slot first
            if (da != last) { *word_at(da) = *word_at(last); if (skth == last) skth = da; }
            last--;
            if (db >= 0) {
                if (db != last) { *word_at(db) = *word_at(last); if (skth == last) skth = db; }
                last--;
            }
        }
        if (valid && flagged) {
            STAT_LANE(4, 1);
            int slot = atomicAdd(prm.flag_count, 1);
            prm.flag_list[slot] = (int)qi;
        }
        // -------------------------------------------------------------------------------------- epilogues
        const bool on = valid && !flagged;
        double d2max = KNN_SENTINEL;
        if (on && !short_of_k) {
            const Vec4<S> c = P[idx_at(skth)];
            d2max = dist2_ref(x0, y0, z0, (double)c.x, (double)c.y, (double)c.z);
        }
        sc_epilogue<S>(prm, P, idx_at, reinterpret_cast<double*>(base + heap_bytes), lane, on, on ? nk : 0, d2max, x0, y0, z0, qi);
        __syncwarp();
    }
}


// tuning overrides of the density kernel (nbk_set_option); the defaults are what ships
static int g_knn_leaf = 0, g_knn_exact = 0, g_knn_transpose = -1;
bool set_knn_option(const char* name, int64_t value) {
    const std::string s(name);
    if (s == "knn_leaf") { g_knn_leaf = (int)value; return true; }
    if (s == "knn_exact") { g_knn_exact = (int)value; return true; }
    if (s == "knn_transpose") { g_knn_transpose = (int)value; return true; }
    return false;
}

static void fill_common(KnnParams& p, nbk_tree& t, const KnnArgs& a) {
    p.nlo = t.nlo; p.nhi = t.nhi; p.bucket = t.bucket;
    p.nlo2 = t.nlo2; p.nhi2 = t.nhi2; p.bucket2 = t.bucket;
    p.P = t.prim; p.V = t.vel4(); p.mass = t.mass; p.order = t.order;
    p.n = t.n;
    p.n_tree = t.n_main ? t.n_main : t.n;
    p.aligned = t.aligned;
    p.tr_max = g_knn_transpose >= 0 ? g_knn_transpose : 12;
    p.q0 = a.q0; p.q1 = a.q1; p.xq = a.xq; p.mode = a.mode;
    p.qlist = a.qlist; p.nq = a.nq;
    p.gather = a.gather ? 1 : 0; p.vq = a.vq;
    p.cand_excl = a.cand_excl; p.crit_mode = a.crit_mode; p.cp0 = a.cp0; p.cp1 = a.cp1;
    p.rho_in = a.rho_in; p.smvel_in = a.smvel_in; p.smvel_out = a.smvel_out; p.smdisp_out = a.smdisp_out;
    p.smdisp_in = a.smdisp_in; p.smhigh_out = a.smhigh_out; p.moment = a.moment;
    p.k = a.k;
    p.periodic = a.periodic; p.strict = a.strict; p.tree_form = a.tree_form;
    for (int d = 0; d < 3; d++) p.period[d] = t.period[d];
    p.nn = a.nn; p.d2out = a.d2; p.out_ids = a.out_ids;
    p.rho = a.rho; p.hsm = a.hsm; p.veldens_k = a.veldens_k;
    p.kern = t.d_kernel; p.kernres = t.kernres;
    p.flag_count = nullptr; p.flag_list = nullptr;
    p.active = a.active;
    p.phase = a.phase ? 1 : 0;
}

static void run_exact(nbk_tree& t, KnnParams& p, int64_t rows) {
    size_t smem = exact_warp_bytes(p.kcap) * KNN_WARPS;
    NBK_REQUIRE(smem <= 227 * 1024, NBK_ERR_ARG, "k too large for the shared-memory heaps (max ~145 at 4 warps/CTA)");
    int blocks = p.qlist ? div_up(rows, KNN_WARPS) : div_up((rows + 31) / 32, KNN_WARPS);
    const bool filter = p.cand_excl != nullptr || p.crit_mode != 0;
    auto go = [&](auto kern) {
        NBK_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p);
    };
    if (p.phase) {
        NBK_REQUIRE(!filter, NBK_ERR_UNSUPPORTED, "phase-space kNN takes no candidate filters");
        if (t.store_bytes == 4) go(knn_exact_kernel<float, false, true>); else go(knn_exact_kernel<double, false, true>);
    }
    else if (t.store_bytes == 4) { if (filter) go(knn_exact_kernel<float, true>); else go(knn_exact_kernel<float, false>); }
    else { if (filter) go(knn_exact_kernel<double, true>); else go(knn_exact_kernel<double, false>); }
    NBK_CHECK(cudaGetLastError());
}

// The fp32 keys of the density kernel must not overflow: squared distances inside the union of the root boxes stay below
// AP_HUGE (decided once per tree; anything larger, or non finite, runs on the fp64-heap kernel).
static bool ap_range_ok(nbk_tree& t) {
    if (t.knn_fp32_ok < 0) {
        NodeLo lo[2]; NodeHi hi[2];
        int nb = 1;
        NBK_CHECK(cudaMemcpyAsync(&lo[0], t.nlo, sizeof(NodeLo), cudaMemcpyDeviceToHost, t.stream));
        NBK_CHECK(cudaMemcpyAsync(&hi[0], t.nhi, sizeof(NodeHi), cudaMemcpyDeviceToHost, t.stream));
        if (t.nlo2) {
            NBK_CHECK(cudaMemcpyAsync(&lo[1], t.nlo2, sizeof(NodeLo), cudaMemcpyDeviceToHost, t.stream));
            NBK_CHECK(cudaMemcpyAsync(&hi[1], t.nhi2, sizeof(NodeHi), cudaMemcpyDeviceToHost, t.stream));
            nb = 2;
        }
        NBK_CHECK(cudaStreamSynchronize(t.stream));
        double mn[3] = {lo[0].x, lo[0].y, lo[0].z}, mx[3] = {hi[0].x, hi[0].y, hi[0].z};
        if (nb == 2) {
            const double l2[3] = {lo[1].x, lo[1].y, lo[1].z}, h2[3] = {hi[1].x, hi[1].y, hi[1].z};
            for (int d = 0; d < 3; d++) { mn[d] = l2[d] < mn[d] ? l2[d] : mn[d]; mx[d] = h2[d] > mx[d] ? h2[d] : mx[d]; }
        }
        double ext2 = 0;
        for (int d = 0; d < 3; d++) ext2 += (mx[d] - mn[d]) * (mx[d] - mn[d]);
        t.knn_fp32_ok = (ext2 < 1.0e36) ? 1 : 0;             // false for NaN / inf as well
    }
    return t.knn_fp32_ok == 1;
}

void launch_knn(nbk_tree& t, const KnnArgs& a) {
    NBK_REQUIRE(a.k >= 1, NBK_ERR_ARG, "k must be >= 1");
    KnnParams p;
    fill_common(p, t, a);
    if (a.veldens_k > 0) NBK_REQUIRE(p.V != nullptr, NBK_ERR_ARG, "velocity density needs velocities");
    if (a.qlist) NBK_REQUIRE(a.mode == 0, NBK_ERR_ARG, "explicit query lists hold particle indices");
    if (a.veldens_k > 0 && a.mode == 1) NBK_REQUIRE(a.vq != nullptr && a.gather, NBK_ERR_ARG, "point form of the velocity density needs query velocities");
    const int64_t rows = a.qlist ? a.nq : a.q1 - a.q0;
    if (rows <= 0) return;
    t.last_launches = 0;
    t.last_flagged = 0;
    if (a.crit_mode == 4) NBK_REQUIRE(p.V != nullptr && (a.mode == 0 || a.vq != nullptr), NBK_ERR_ARG, "FOF6d-filtered search needs velocities");
    if (a.phase) {
        NBK_REQUIRE(p.V != nullptr && (a.mode == 0 || a.vq != nullptr), NBK_ERR_ARG, "phase-space search needs velocities");
        NBK_REQUIRE(!a.rho && !a.hsm && !a.qlist && !a.smvel_out && !a.smdisp_out && !a.smhigh_out, NBK_ERR_UNSUPPORTED,
                    "phase-space search returns neighbour lists only");
    }
    if (a.smvel_out || a.smdisp_out || a.smhigh_out) {
        NBK_REQUIRE(a.mode == 0 && !a.qlist && a.rho_in && p.V, NBK_ERR_ARG, "smoothed velocity moments need particle queries, densities and velocities");
        NBK_REQUIRE(!a.smdisp_out || a.smvel_in, NBK_ERR_ARG, "CalcSmoothVelDisp needs the smoothed mean velocities");
        NBK_REQUIRE(!a.smhigh_out || (a.smvel_in && a.smdisp_in && (a.moment == 3 || a.moment == 4)), NBK_ERR_ARG, "CalcSmoothVelSkew / Kurtosis need the smoothed mean velocities and dispersions");
    }
    const bool smooth_only = a.mode == 0 && !a.periodic && !a.nn && !a.d2 && (a.rho || a.hsm) && !a.gather && !a.qlist && !a.cand_excl && !a.crit_mode &&
                             !a.smvel_out && !a.smdisp_out && !a.smhigh_out && !a.phase;
    if (smooth_only && !g_knn_exact && ap_range_ok(t)) {
        // ---- packed-word heap kernel (knn_hp_kernel), exact kernel for the flagged queries --------------------------
        p.kcap = a.k;
        // Nodes of up to `leaf` particles are scanned as one tile: the level whose nodes hold 21..40 particles (exactly one
        // level does: sizes halve) -- a tile and a bit; with a fixed threshold of 32 a particle count just above a power of
        // two would be scanned as half-empty 16/17-particle tiles.
        {
            int64_t sz = t.n_main ? t.n_main : t.n;
            while (sz > 40) sz = split_left(sz, t.aligned);
            int leaf = g_knn_leaf > 0 ? g_knn_leaf : (int)sz;
            if (leaf > p.bucket) p.bucket = leaf;
            if (t.nlo2) {
                sz = t.n - t.n_main;
                while (sz > 40) sz = split_left(sz, t.aligned2);
                leaf = g_knn_leaf > 0 ? g_knn_leaf : (int)sz;
                if (leaf > p.bucket2) p.bucket2 = leaf;
            }
        }
        const int want_doubles = (a.veldens_k > 0 && a.veldens_k < a.k) ? 1 : 0;
        const int tile_slots = (p.bucket <= 32 && (!t.nlo2 || p.bucket2 <= 32)) ? 32 : 40;
        const int tile_bytes = t.store_bytes == 4 ? tile_slots * 16 : 3 * tile_slots * 8;
        const size_t warp_bytes = hp_warp_bytes(a.k, want_doubles, tile_bytes);
        int nsm = 0, smem_max = 0;
        NBK_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, t.device));
        NBK_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, t.device));
        const int64_t ngroups = (rows + 31) / 32;
        int warps = (int)std::min<int64_t>(HP_MAX_WARPS, (int64_t)smem_max / (int64_t)warp_bytes);
        NBK_REQUIRE(warps >= 1, NBK_ERR_ARG, "k too large for the shared-memory heaps");
        if (ngroups < (int64_t)nsm * warps) warps = (int)std::max<int64_t>(1, (ngroups + nsm - 1) / nsm);      // small inputs: spread over the SMs
        const size_t smem = warp_bytes * warps;
        DevBuf<int> counters(2);
        DevBuf<int32_t> flist(rows);
        NBK_CHECK(cudaMemsetAsync(counters.p, 0, 2 * sizeof(int), t.stream));
        p.flag_count = counters.p; p.flag_list = flist.p;
        auto go = [&](auto kern) {
            NBK_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t blocks = std::min<int64_t>(nsm, (ngroups + warps - 1) / warps);
            kern<<<(int)blocks, warps * 32, smem, t.stream>>>(p, want_doubles, counters.p + 1);
        };
        if (tile_slots == 32) {
            if (t.store_bytes == 4) { if (t.nlo2) go(knn_hp_kernel<float, true, 32>); else go(knn_hp_kernel<float, false, 32>); }
            else { if (t.nlo2) go(knn_hp_kernel<double, true, 32>); else go(knn_hp_kernel<double, false, 32>); }
        } else {
            if (t.store_bytes == 4) { if (t.nlo2) go(knn_hp_kernel<float, true, 40>); else go(knn_hp_kernel<float, false, 40>); }
            else { if (t.nlo2) go(knn_hp_kernel<double, true, 40>); else go(knn_hp_kernel<double, false, 40>); }
        }
        NBK_CHECK(cudaGetLastError());
#ifdef NBK_STATS
        {
            unsigned long long h[8];
            NBK_CHECK(cudaStreamSynchronize(t.stream));
            NBK_CHECK(cudaMemcpyFromSymbol(h, g_stats, sizeof(h)));
            double g = (double)((rows + 31) / 32);
            fprintf(stderr, "[nbk stats] per warp: tiles %.1f rounds %.1f ; per lane: past the screen %.1f ; flagged %llu tile-list overflows %llu\n",
                    h[0] / g, h[2] / g, h[3] / (double)rows, h[4], h[6]);
            unsigned long long z[8] = {0};
            NBK_CHECK(cudaMemcpyToSymbol(g_stats, z, sizeof(z)));
        }
#endif
        t.last_launches += 2;
        int nflag = 0;
        NBK_CHECK(cudaMemcpyAsync(&nflag, counters.p, sizeof(int), cudaMemcpyDeviceToHost, t.stream));
        NBK_CHECK(cudaStreamSynchronize(t.stream));
        t.last_flagged = nflag;
        if (nflag > 0) {
            KnnParams pe = p;
            pe.kcap = a.k;
            pe.bucket = t.bucket; pe.bucket2 = t.bucket;
            pe.qlist = flist.p; pe.nq = nflag;
            pe.flag_count = nullptr; pe.flag_list = nullptr;
            run_exact(t, pe, nflag);
            t.last_launches += 1;
        }
        return;
    }
    p.kcap = a.k + ((a.periodic && a.mode == 0) ? 1 : 0);
    // the reference's periodic FindNearestCheck / FindNearestCriterion search k+1 and drop the nearest in every form
    if ((a.cand_excl || a.crit_mode) && a.periodic && a.tree_form) p.kcap = a.k + 1;
    // FindNearestPhase(x, v) holds k slots; the particle form k + 1 with the nearest dropped (tree_form set by the caller)
    run_exact(t, p, rows);
    t.last_launches += 1;
}

}  // namespace nbk
