// knn.cu -- k-nearest-neighbour search with the SPH density / velocity-density / smoothing-scale
// epilogues fused into the same kernel.
//
// Replaces (reference): LeafNode::FindNearestPos KDLeafNode.cxx:15-28,119-130; SplitNode::FindNearestPos
// KDSplitNode.cxx:15-41; FindNearestPosPeriodic :1119-1148; PriorityQueue.h:14-85; drivers
// KDFindNearest.cxx:247-334,462-554; KDTree::CalcDensity / CalcVelDensity KDCalcSmoothQuantities.cxx:203-389.
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
//   knn_ap_kernel     the density family (CalcDensity / CalcVelDensity / smoothing scale), non periodic target
//                     form: append + prune selection on fp32 keys with a rigorous error bound, no heap and no
//                     serial insertion rounds (described at the kernel).  Queries whose k-th / (k+1)-th keys are
//                     closer than the error band are re-run by the exact kernel.
#include <string.h>

#include "traverse.cuh"
#include "tree.h"

namespace nbk {

constexpr int KNN_WARPS = 4;
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
    int64_t n_tree;                                   // particles of the main tree (= n unless a halo is attached)
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

template <class S, bool FILTER = false>
struct KnnVisitor {
    const Vec4<S>* P;
    const Vec4<S>* V;   // FILTER with a 6D criterion only
    double* tile;       // [6][32]
    WarpHeap hp;
    double qx, qy, qz;
    double vx, vy, vz;  // query velocity (FILTER, FOF6d)
    const int32_t* excl; int crit_mode; double cp0, cp1;
    double top;         // heap top (current k-th distance^2)
    float topf;         // top rounded up
    int self;           // tree index of the query particle, or -1 (coordinate form)
    bool target_form;   // skip self and d2==0 (KDLeafNode.cxx:15-28)
    bool on;
    unsigned lane;

    __device__ __forceinline__ bool need(float lb) const { return lb < topf; }
    __device__ __forceinline__ void settop() { top = hp.h(0); topf = __double2float_ru(top); }

    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[start + base + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
                if (FILTER && crit_mode == 4) {
                    Vec4<S> u = V[start + base + lane];
                    tile[96 + lane] = (double)u.x; tile[128 + lane] = (double)u.y; tile[160 + lane] = (double)u.z;
                }
            }
            __syncwarp();
            unsigned acc = 0;
#pragma unroll 4
            for (int j = 0; j < m; j++) {
                double d2 = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
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
                    double d2 = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
                    if (d2 < top) { hp.replace_top(d2, start + base + j); settop(); }
                }
            }
        }
    }
};

constexpr int EXACT_TILE_DOUBLES = 192;   // [6][32]: positions, and velocities for the FOF6d-filtered search
static __host__ __device__ inline size_t exact_warp_bytes(int kcap) { return (size_t)kcap * 32 * 12 + EXACT_TILE_DOUBLES * 8 + TRAV_STACK * 4; }

template <class S, bool FILTER>
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

    KnnVisitor<S, FILTER> v;
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
            if (FILTER && prm.crit_mode == 4) { Vec4<S> u = v.V[qi]; v.vx = (double)u.x; v.vy = (double)u.y; v.vz = (double)u.z; }
        } else {
            x0 = prm.xq[3 * qi]; y0 = prm.xq[3 * qi + 1]; z0 = prm.xq[3 * qi + 2];
            if (FILTER && prm.crit_mode == 4) { v.vx = prm.vq[3 * qi]; v.vy = prm.vq[3 * qi + 1]; v.vz = prm.vq[3 * qi + 2]; }
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
    if (prm.smvel_out || prm.smdisp_out) {
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
        if (prm.smdisp_out) { mi[0] = prm.smvel_in[3 * qi]; mi[1] = prm.smvel_in[3 * qi + 1]; mi[2] = prm.smvel_in[3 * qi + 2]; }
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

// ====================================================================================== append + prune select
// The density family needs (a) the exact set of the k nearest, (b) the exact k-th distance and (c) a sum over the set.
// Per lane (= query) the kernel keeps an APPEND BUFFER of fp32 keys in shared memory ([cap][32], slot-major) and, entry
// for entry, the candidates' tree indices in a global scratch block ([cap][32] per resident warp, L2 resident).  A
// candidate is screened with ONE fp32 distance and, if it is below the lane's bound, appended by two predicated stores:
// there is no heap and no SIMT-serial insertion round -- every lane works on every instruction of the tile scan.
// When some lane's buffer is nearly full, the whole warp PRUNES in lock step: a 32-bin histogram of the lane's keys
// finds the bin holding rank R = k+1, one compaction pass keeps the bins below it, and the (<= 8) members of that bin
// are ranked by a sorting network in registers; the lane's bound becomes its exact R-th smallest key.  The last prune
// leaves exactly the k+1 smallest keys.
//
// Exactness.  Keys are a = fl32(d2) evaluated with 3 subtractions, 1 multiplication, 2 FMAs on exact fp32 coordinates:
// |a - d2| <= 5.01 * 2^-24 * d2 =: eps * d2 (fp64 storage: a = RN_fp32 of the reference's fp64 d2, eps = 2^-24).
//   * Every candidate that is dropped -- screened out (a > bound * (1 + 2^-20)), in a subtree whose box lower bound
//     reaches bound * (1 + 2^-20), or pruned -- has a >= the lane's final (k+1)-th smallest key m1.
//   * If m1 > m2 * (1 + 2^-20) (m2 = k-th smallest key), every dropped candidate and the (k+1)-th itself are farther in
//     exact arithmetic than each of the k kept ones, because 2^-20 > 2 eps: the k smallest keys ARE the k nearest.
//   * The exact k-th distance is the reference fp64 d2 of the entry with key m2 provided the third largest key m3 is
//     below m2 * (1 - 2^-20); the SPH sums use fp64 d2 recomputed from the indices.
//   * Queries failing either gap test (~2e-4 of them), meeting an under/overflowing key or a degenerate histogram are
//     appended to a list and re-run by knn_exact_kernel (fp64 heap).
// Traversal: bottom-up (traverse_bottom_up): the group's own node first, then the sibling subtrees of its ancestors.
// Persistent grid: warps draw 32-query groups from a global counter, so the grid is one wave whatever the particle count.
constexpr float AP_TINY = 7.888609052210118e-31f;           // 2^-100: below it the relative error bound of a key is not guaranteed
constexpr float AP_HUGE = 1.0e37f;
constexpr float AP_WIDEN = 1.00000095367431640625f;          // 1 + 2^-20
constexpr float AP_NARROW = 0.99999904632568359375f;         // 1 - 2^-20
constexpr int AP_STASH = 8;
constexpr int AP_AUX_BYTES = 2048;                           // histogram u16 [32][32]; then the stash: float [8][32] + int [8][32]
constexpr int AP_MAX_LEVELS = 6;

#ifdef NBK_STATS
__device__ unsigned long long g_stats[8];   // 0 tiles, 1 prunes, 2 prune levels, 3 appended (lane-level), 4 flagged, 5 failed
#define STAT(i, v) do { if (lane_id() == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#define STAT_LANE(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define STAT(i, v)
#define STAT_LANE(i, v)
#endif

// Leaf tile staged in shared memory and the per-candidate key.
template <class S> struct ApTile;
template <> struct ApTile<float> {
    static constexpr int TILE_BYTES = 32 * 16;
    float4* t;
    float qx, qy, qz;
    __device__ __forceinline__ void init(void* mem, double x, double y, double z) {
        t = reinterpret_cast<float4*>(mem);
        qx = (float)x; qy = (float)y; qz = (float)z;     // exact: fp32 storage
    }
    __device__ __forceinline__ void load(const Vec4<float>* __restrict__ P, int first, int m, unsigned lane) {
        // slots past the end of the leaf hold NaN: every comparison on them is false, so the scan runs over whole groups of 8
        const float nanf_ = __int_as_float(0x7fc00000);
        float4 mine = make_float4(nanf_, nanf_, nanf_, 0.f);
        if ((int)lane < m) { Vec4<float> c = P[first + lane]; mine = make_float4(c.x, c.y, c.z, 0.f); }
        t[lane] = mine;
    }
    __device__ __forceinline__ float key(int j) const {
        const float4 c = t[j];
        const float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
        return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
    }
    __device__ __forceinline__ bool coincident(int j) const { const float4 c = t[j]; return qx == c.x && qy == c.y && qz == c.z; }
};
template <> struct ApTile<double> {
    static constexpr int TILE_BYTES = 96 * 8;
    double* t;
    double qx, qy, qz;
    __device__ __forceinline__ void init(void* mem, double x, double y, double z) { t = reinterpret_cast<double*>(mem); qx = x; qy = y; qz = z; }
    __device__ __forceinline__ void load(const Vec4<double>* __restrict__ P, int first, int m, unsigned lane) {
        const double nan_ = __longlong_as_double(0x7ff8000000000000ll);
        double cx = nan_, cy = nan_, cz = nan_;
        if ((int)lane < m) { Vec4<double> c = P[first + lane]; cx = c.x; cy = c.y; cz = c.z; }
        t[lane] = cx; t[32 + lane] = cy; t[64 + lane] = cz;
    }
    __device__ __forceinline__ float key(int j) const { return __double2float_rn(dist2_ref(qx, qy, qz, t[j], t[32 + j], t[64 + j])); }
    __device__ __forceinline__ bool coincident(int j) const { return qx == t[j] && qy == t[32 + j] && qz == t[64 + j]; }
};

__device__ __forceinline__ void ap_cex(float& ka, int& ia, float& kb, int& ib) {
    const bool sw = ka > kb;
    const float k0 = sw ? kb : ka, k1 = sw ? ka : kb;
    const int i0 = sw ? ib : ia, i1 = sw ? ia : ib;
    ka = k0; kb = k1; ia = i0; ib = i1;
}

// Lock-step prune of the lanes' append buffers (all 32 lanes call it).  Lane state: cnt entries (keys kb[s*32], indices
// lg[s*32]), all keys <= bound * (1 + 2^-20).  Lanes with cnt > R keep their R smallest keys (FINAL: exactly; otherwise
// possibly a few more when the rank-R bin holds more than 8 keys) and get bound = largest kept key.  Returns
// (bound bits << 32 | cnt); cnt = -1: the lane could not be resolved (degenerate keys) and goes to the exact kernel.
template <bool FINAL>
__device__ __noinline__ unsigned long long ap_prune(float* kb, int* lg, unsigned char* aux, int cnt, int R, float bound) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = lane_id();
    bool act = cnt > R;
    if (!__any_sync(full, act)) return ((unsigned long long)__float_as_uint(bound) << 32) | (unsigned)cnt;
    STAT(1, 1);
    float tlo = 0.f, thi = __fmul_ru(bound, AP_WIDEN);
    int cbelow = 0, level = 0;
    if (__any_sync(full, act && !(thi <= 3.0e38f))) {
        // lanes still on the infinite bound: the range is their largest key
        const int nmax = __reduce_max_sync(full, act ? cnt : 0);
        float mx = 0.f;
#pragma unroll 4
        for (int s = 0; s < nmax; s++) if (s < cnt) mx = fmaxf(mx, kb[s * 32]);
        if (!(thi <= 3.0e38f)) thi = mx;
    }
    unsigned short* myh = reinterpret_cast<unsigned short*>(aux) + lane;      // histogram column: bin b at myh[b*32]
    float* sk = reinterpret_cast<float*>(aux) + lane;                         // stash keys [8][32]
    int* si = reinterpret_cast<int*>(aux) + 256 + lane;                       // stash indices [8][32]
    while (__any_sync(full, act)) {
        STAT(2, 1);
        const int nmax = __reduce_max_sync(full, act ? cnt : 0);
        __syncwarp();
        {
            uint4* h4 = reinterpret_cast<uint4*>(aux);
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int i = 0; i < AP_AUX_BYTES / 512; i++) h4[i * 32 + lane] = z;
        }
        __syncwarp();
        const float scale = (thi > tlo) ? 31.99f / (thi - tlo) : 0.f;
        // ---- histogram of the keys inside [tlo, thi] ------------------------------------------------------
#pragma unroll 4
        for (int s = 0; s < nmax; s++) {
            if (act && s < cnt) {
                const float key = kb[s * 32];
                if (key >= tlo && key <= thi) {
                    const int b = min(__float2int_rz((key - tlo) * scale), 31);
                    myh[b * 32] += 1;
                }
            }
        }
        // ---- the bin that holds rank R --------------------------------------------------------------------
        int bstar = -1, cbefore = 0, p = 0;
        {
            int cum = cbelow;
#pragma unroll 8
            for (int b = 0; b < 32; b++) {
                const int h = myh[b * 32];
                if (bstar < 0 && cum + h >= R) { bstar = b; cbefore = cum; p = h; }
                cum += h;
            }
        }
        if (act && bstar < 0) { cnt = -1; act = false; }            // cannot happen for consistent state; never loop on it
        const int need = R - cbefore;                                // members of bin bstar that complete the R smallest
        const bool small = p <= AP_STASH;
        __syncwarp();                                                // histogram columns are dead: the region becomes the stash
        // ---- compaction: bins below bstar stay, bins above go, members of bstar go to the stash (or stay) -----
        int wp = 0, ns = 0;
        float mn = __int_as_float(0x7f800000), mx = 0.f;
        for (int s0 = 0; s0 < nmax; s0 += 8) {
            int id[8];
            float ky[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const bool in = act && s0 + u < cnt;
                id[u] = in ? lg[(s0 + u) * 32] : 0;
                ky[u] = in ? kb[(s0 + u) * 32] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (act && s0 + u < cnt) {
                    const float key = ky[u];
                    bool keep = key < tlo, member = false;
                    if (!keep && key <= thi) {
                        const int b = min(__float2int_rz((key - tlo) * scale), 31);
                        keep = b < bstar; member = b == bstar;
                    }
                    if (member) {
                        if (small) { sk[ns * 32] = key; si[ns * 32] = id[u]; ns++; }
                        else { keep = true; mn = fminf(mn, key); mx = fmaxf(mx, key); }
                    }
                    if (keep) { kb[wp * 32] = key; lg[wp * 32] = id[u]; wp++; }
                }
            }
        }
        if (act) {
            if (small) {
                float k8[8]; int i8[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { const bool in = u < ns; k8[u] = in ? sk[u * 32] : __int_as_float(0x7f800000); i8[u] = in ? si[u * 32] : -1; }
                // 19-comparator network for 8 keys (Batcher odd-even merge sort)
                ap_cex(k8[0], i8[0], k8[1], i8[1]); ap_cex(k8[2], i8[2], k8[3], i8[3]); ap_cex(k8[4], i8[4], k8[5], i8[5]); ap_cex(k8[6], i8[6], k8[7], i8[7]);
                ap_cex(k8[0], i8[0], k8[2], i8[2]); ap_cex(k8[1], i8[1], k8[3], i8[3]); ap_cex(k8[4], i8[4], k8[6], i8[6]); ap_cex(k8[5], i8[5], k8[7], i8[7]);
                ap_cex(k8[1], i8[1], k8[2], i8[2]); ap_cex(k8[5], i8[5], k8[6], i8[6]);
                ap_cex(k8[0], i8[0], k8[4], i8[4]); ap_cex(k8[1], i8[1], k8[5], i8[5]); ap_cex(k8[2], i8[2], k8[6], i8[6]); ap_cex(k8[3], i8[3], k8[7], i8[7]);
                ap_cex(k8[2], i8[2], k8[4], i8[4]); ap_cex(k8[3], i8[3], k8[5], i8[5]);
                ap_cex(k8[1], i8[1], k8[2], i8[2]); ap_cex(k8[3], i8[3], k8[4], i8[4]); ap_cex(k8[5], i8[5], k8[6], i8[6]);
                float b = bound;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (u < need) { kb[wp * 32] = k8[u]; lg[wp * 32] = i8[u]; wp++; b = k8[u]; }
                }
                bound = b; cnt = wp; act = false;
            } else {
                // more than 8 keys in the rank-R bin: keep the whole bin; the bound is its largest key
                cnt = wp;
                if (FINAL) {
                    cbelow = cbefore; tlo = mn; thi = mx;
                    if (!(mx > mn) || ++level >= AP_MAX_LEVELS) { cnt = -1; act = false; }      // all keys of the bin equal: exact kernel
                } else { bound = mx; act = false; }
            }
        }
    }
    __syncwarp();
    return ((unsigned long long)__float_as_uint(bound) << 32) | (unsigned)cnt;
}

template <class S>
struct ApVisitor {
    const Vec4<S>* P;
    ApTile<S> tile;
    float* kb;            // shared: this lane's key column, entry s at kb[s*32]
    int* lg;              // global: this lane's index column, entry s at lg[s*32]
    unsigned char* aux;   // shared: the warp's histogram / stash region
    int cnt, cap, R;
    float bound, limf;    // bound: largest kept key after the last prune (or +inf); limf = bound * (1 + 2^-20), -1: lane accepts nothing
    bool failed;
    unsigned lane;
    __device__ __forceinline__ bool need(float lb) const { return lb < limf; }
    __device__ __forceinline__ void fail() { failed = true; limf = -1.f; cnt = 0; }
    template <bool FINAL>
    __device__ __forceinline__ void prune() {
        const unsigned long long r = ap_prune<FINAL>(kb, lg, aux, failed ? 0 : cnt, R, bound);
        const int c = (int)(unsigned)(r & 0xffffffffull);
        if (!failed) {
            if (c < 0 || (!FINAL && c > cap - 8)) fail();
            else { cnt = c; bound = __uint_as_float((unsigned)(r >> 32)); limf = __fmul_ru(bound, AP_WIDEN); }
        }
    }
    __device__ __forceinline__ void leaf(int start, int n, int, unsigned) {
        for (int base = 0; base < n; base += 32) {
            const int m = min(32, n - base);
            __syncwarp();
            tile.load(P, start + base, m, lane);
            __syncwarp();
            STAT(0, 1);
            for (int j0 = 0; j0 < m; j0 += 8) {
                if (__any_sync(0xffffffffu, cnt > cap - 8)) prune<false>();
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const float a = tile.key(j0 + jj);
                    if (a <= limf) {
                        if (a >= AP_TINY && a <= AP_HUGE) { kb[cnt * 32] = a; lg[cnt * 32] = start + base + j0 + jj; cnt++; STAT_LANE(3, 1); }
                        else if (!tile.coincident(j0 + jj)) fail();     // the query itself and coincident particles are never neighbours
                    }
                }
            }
        }
    }
};

// SPH epilogues over a lane's neighbour list L (entry s at L[s*32], global or shared); D: per-lane doubles [k][32] in shared
// memory for the kv < kx velocity-density selection.
template <class S>
__device__ __forceinline__ void sc_epilogue(const KnnParams& prm, const Vec4<S>* __restrict__ P, const int* L, double* Dbase, unsigned lane,
                                            int cnt, double d2max, double x0, double y0, double z0, int64_t qi) {
    if (prm.hsm) prm.hsm[qi] = 0.5 * sqrt(d2max);
    if (prm.rho && prm.veldens_k == 0) {
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
            for (int u = 0; u < 4; u++) id[u] = (s0 + u < cnt) ? L[(s0 + u) * 32] : -1;
#pragma unroll
            for (int u = 0; u < 4; u++) if (id[u] >= 0) { c[u] = P[id[u]]; mj[u] = prm.mass[id[u]]; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (id[u] >= 0) {
                    const double rij = sqrt(dist2_ref(x0, y0, z0, (double)c[u].x, (double)c[u].y, (double)c[u].z));
                    const double r = rij * inv_hi;
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
        const Vec4<S> vi = V[qi];
        const double delta = 2.0 / (double)(prm.kernres - 1);
        const int kv = min(prm.veldens_k, cnt);
        double rho = 0;
        if (kv == cnt) {
            // every spatial neighbour is used: h from the largest velocity distance, then the sum (two passes)
            double vmax = 0;
            for (int s = 0; s < cnt; s++) {
                Vec4<S> vj = V[L[s * 32]];
                vmax = fmax(vmax, sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z)));
            }
            const double hi = 0.5 * vmax;
            const double norm = 1.0 / pow(hi, 3.0);
            for (int s = 0; s < cnt; s++) {
                Vec4<S> vj = V[L[s * 32]];
                double r = sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z)) / hi;
                rho = rho + wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
            }
        } else if (kv > 0) {
            // kv < kx: exact fp64 selection on a per-lane array of doubles in shared memory
            double* D = Dbase + lane;
            for (int s = 0; s < cnt; s++) {
                Vec4<S> vj = V[L[s * 32]];
                D[s * 32] = sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z));
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
        prm.rho[qi] = rho;
    }
}

static inline int ap_capacity(int k) {
    int c = 2 * k > k + 48 ? 2 * k : k + 48;
    return (c + 7) & ~7;
}
static inline size_t ap_warp_bytes(int k, int cap, bool want_doubles, int tile_bytes) {
    size_t region = (size_t)cap * 32 * 4;
    if (want_doubles && (size_t)k * 32 * 8 > region) region = (size_t)k * 32 * 8;
    return region + AP_AUX_BYTES + tile_bytes + TRAV_STACK * 4;
}

template <class S, bool HALO>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_ap_kernel(KnnParams prm, int want_doubles, int cap, int* __restrict__ work_counter,
                                                                int32_t* __restrict__ logbuf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned full = 0xffffffffu;
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int k = prm.k, R = k + 1;
    size_t region = (size_t)cap * 32 * 4;
    if (want_doubles && (size_t)k * 32 * 8 > region) region = (size_t)k * 32 * 8;
    const size_t warp_bytes = region + AP_AUX_BYTES + ApTile<S>::TILE_BYTES + TRAV_STACK * 4;
    unsigned char* base = smem_raw + w * warp_bytes;
    unsigned char* aux = base + region;
    void* tile_mem = base + region + AP_AUX_BYTES;
    int* stack = reinterpret_cast<int*>(base + region + AP_AUX_BYTES + ApTile<S>::TILE_BYTES);
    int* lg = logbuf + ((size_t)blockIdx.x * KNN_WARPS + w) * (size_t)cap * 32 + lane;
    float* kb = reinterpret_cast<float*>(base) + lane;

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
        ApVisitor<S> v;
        v.P = P; v.lane = lane;
        v.tile.init(tile_mem, x0, y0, z0);
        v.kb = kb; v.lg = lg; v.aux = aux;
        v.cnt = 0; v.cap = cap; v.R = R;
        v.bound = __int_as_float(0x7f800000);
        v.limf = valid ? v.bound : -1.f;
        v.failed = false;
        traverse_bottom_up(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid, prm.n_tree, g0, g1);
        if (HALO) traverse(prm.nlo2, prm.nhi2, prm.bucket2, stack, v, qb, valid);
        v.template prune<true>();

        // ------------------------------------------------------------------------ the k nearest and the k-th
        // three largest keys (m1 >= m2 >= m3) and the slots of the first two
        int cnt = v.cnt;
        float m1 = -1.f, m2 = -1.f, m3 = -1.f;
        int s1 = -1, s2 = -1;
        {
            const int nmax = __reduce_max_sync(full, valid ? cnt : 0);
#pragma unroll 4
            for (int s = 0; s < nmax; s++) {
                if (s < cnt) {
                    const float key = kb[s * 32];
                    if (key > m1) { m3 = m2; m2 = m1; s2 = s1; m1 = key; s1 = s; }
                    else if (key > m2) { m3 = m2; m2 = key; s2 = s; }
                    else if (key > m3) m3 = key;
                }
            }
        }
        bool flagged = v.failed;
        int nk = cnt, skth = s1;
        float kth = m1, below = m2;
        if (valid && !flagged && cnt == R) {
            // k+1 entries: the largest is the (k+1)-th neighbour; it leaves, and certifies the set if the gap is wide enough
            if (!(m1 > __fmul_ru(m2, AP_WIDEN))) flagged = true;
            const int last = cnt - 1;
            if (s1 != last) lg[s1 * 32] = lg[last * 32];
            if (s2 == last) s2 = s1;
            nk = cnt - 1; skth = s2; kth = m2; below = m3;
        }
        const bool short_of_k = nk < k;                     // fewer than k candidates exist: the reference's heap keeps sentinels
        if (valid && !flagged && !short_of_k && nk >= 2 && !(below < __fmul_rd(kth, AP_NARROW))) flagged = true;
        if (valid && flagged) {
            STAT_LANE(4, 1);
            int slot = atomicAdd(prm.flag_count, 1);
            prm.flag_list[slot] = (int)qi;
        }
        __syncwarp();   // the key buffer is dead from here on (the velocity-density selection reuses it)

        // -------------------------------------------------------------------------------------- epilogues
        if (valid && !flagged) {
            double d2max = KNN_SENTINEL;
            if (!short_of_k) {
                const Vec4<S> c = P[lg[skth * 32]];
                d2max = dist2_ref(x0, y0, z0, (double)c.x, (double)c.y, (double)c.z);
            }
            sc_epilogue<S>(prm, P, lg, reinterpret_cast<double*>(base), lane, nk, d2max, x0, y0, z0, qi);
        }
        __syncwarp();
    }
}


static void fill_common(KnnParams& p, nbk_tree& t, const KnnArgs& a) {
    p.nlo = t.nlo; p.nhi = t.nhi; p.bucket = t.bucket;
    p.nlo2 = t.nlo2; p.nhi2 = t.nhi2; p.bucket2 = t.bucket;
    p.P = t.prim; p.V = t.vel4(); p.mass = t.mass; p.order = t.order;
    p.n = t.n;
    p.n_tree = t.n_main ? t.n_main : t.n;
    p.q0 = a.q0; p.q1 = a.q1; p.xq = a.xq; p.mode = a.mode;
    p.qlist = a.qlist; p.nq = a.nq;
    p.gather = a.gather ? 1 : 0; p.vq = a.vq;
    p.cand_excl = a.cand_excl; p.crit_mode = a.crit_mode; p.cp0 = a.cp0; p.cp1 = a.cp1;
    p.rho_in = a.rho_in; p.smvel_in = a.smvel_in; p.smvel_out = a.smvel_out; p.smdisp_out = a.smdisp_out;
    p.k = a.k;
    p.periodic = a.periodic; p.strict = a.strict; p.tree_form = a.tree_form;
    for (int d = 0; d < 3; d++) p.period[d] = t.period[d];
    p.nn = a.nn; p.d2out = a.d2; p.out_ids = a.out_ids;
    p.rho = a.rho; p.hsm = a.hsm; p.veldens_k = a.veldens_k;
    p.kern = t.d_kernel; p.kernres = t.kernres;
    p.flag_count = nullptr; p.flag_list = nullptr;
    p.active = a.active;
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
    if (t.store_bytes == 4) { if (filter) go(knn_exact_kernel<float, true>); else go(knn_exact_kernel<float, false>); }
    else { if (filter) go(knn_exact_kernel<double, true>); else go(knn_exact_kernel<double, false>); }
    NBK_CHECK(cudaGetLastError());
}

// tuning overrides of the density kernel (nbk_set_option); the defaults are what ships
static int g_knn_cap = 0, g_knn_leaf = 0, g_knn_exact = 0;
bool set_knn_option(const char* name, int64_t value) {
    const std::string s(name);
    if (s == "knn_cap") { g_knn_cap = (int)value; return true; }
    if (s == "knn_leaf") { g_knn_leaf = (int)value; return true; }
    if (s == "knn_exact") { g_knn_exact = (int)value; return true; }
    return false;
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
    if (a.smvel_out || a.smdisp_out) {
        NBK_REQUIRE(a.mode == 0 && !a.qlist && a.rho_in && p.V, NBK_ERR_ARG, "smoothed velocity moments need particle queries, densities and velocities");
        NBK_REQUIRE(!a.smdisp_out || a.smvel_in, NBK_ERR_ARG, "CalcSmoothVelDisp needs the smoothed mean velocities");
    }
    const bool smooth_only = a.mode == 0 && !a.periodic && !a.nn && !a.d2 && (a.rho || a.hsm) && !a.gather && !a.qlist && !a.cand_excl && !a.crit_mode &&
                             !a.smvel_out && !a.smdisp_out;
    if (smooth_only && !g_knn_exact) {
        // ---- append + prune kernel, exact kernel for the flagged queries -----------------------------------------
        p.kcap = a.k;
        // Nodes of up to `leaf` particles are scanned as one tile: the level whose nodes hold 21..40 particles (exactly one
        // level does: sizes halve) -- a tile and a bit; with a fixed threshold of 32 a particle count just above a power of
        // two would be scanned as half-empty 16/17-particle tiles.
        {
            int64_t sz = t.n_main ? t.n_main : t.n;
            while (sz > 40) sz = (sz + 1) / 2;
            int leaf = g_knn_leaf > 0 ? g_knn_leaf : (int)sz;
            if (leaf > p.bucket) p.bucket = leaf;
            if (t.nlo2) {
                sz = t.n - t.n_main;
                while (sz > 40) sz = (sz + 1) / 2;
                leaf = g_knn_leaf > 0 ? g_knn_leaf : (int)sz;
                if (leaf > p.bucket2) p.bucket2 = leaf;
            }
        }
        const int want_doubles = (a.veldens_k > 0 && a.veldens_k < a.k) ? 1 : 0;
        int cap = ap_capacity(a.k);
        if (g_knn_cap > a.k + 16) cap = (g_knn_cap + 7) & ~7;
        NBK_REQUIRE(cap < 65536, NBK_ERR_ARG, "k too large");
        const size_t smem = ap_warp_bytes(a.k, cap, want_doubles, t.store_bytes == 4 ? 32 * 16 : 96 * 8) * KNN_WARPS;
        NBK_REQUIRE(smem <= 227 * 1024, NBK_ERR_ARG, "k too large for the shared-memory candidate buffers");
        int nsm = 0, per_sm = 0;
        NBK_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, t.device));
        const int64_t ngroups = (rows + 31) / 32;
        DevBuf<int> counters(2);
        DevBuf<int32_t> flist(rows);
        DevBuf<int32_t> logbuf;
        NBK_CHECK(cudaMemsetAsync(counters.p, 0, 2 * sizeof(int), t.stream));
        p.flag_count = counters.p; p.flag_list = flist.p;
        auto go = [&](auto kern) {
            NBK_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NBK_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KNN_WARPS * 32, smem));
            NBK_REQUIRE(per_sm >= 1, NBK_ERR_ARG, "k too large for the shared-memory candidate buffers");
            int64_t blocks = (int64_t)nsm * per_sm;
            if (blocks > (ngroups + KNN_WARPS - 1) / KNN_WARPS) blocks = (ngroups + KNN_WARPS - 1) / KNN_WARPS;
            // index log: one [cap][32] block per resident warp, rewritten in place group after group (stays in L2)
            logbuf.alloc((size_t)blocks * KNN_WARPS * cap * 32);
            kern<<<(int)blocks, KNN_WARPS * 32, smem, t.stream>>>(p, want_doubles, cap, counters.p + 1, logbuf.p);
        };
        if (t.store_bytes == 4) { if (t.nlo2) go(knn_ap_kernel<float, true>); else go(knn_ap_kernel<float, false>); }
        else { if (t.nlo2) go(knn_ap_kernel<double, true>); else go(knn_ap_kernel<double, false>); }
        NBK_CHECK(cudaGetLastError());
#ifdef NBK_STATS
        {
            unsigned long long h[8];
            NBK_CHECK(cudaStreamSynchronize(t.stream));
            NBK_CHECK(cudaMemcpyFromSymbol(h, g_stats, sizeof(h)));
            double g = (double)((rows + 31) / 32);
            fprintf(stderr, "[nbk stats] per warp: tiles %.1f prunes %.2f levels %.2f ; per lane: appended %.1f ; flagged %llu\n",
                    h[0] / g, h[1] / g, h[2] / g, h[3] / (double)rows, h[4]);
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
    run_exact(t, p, rows);
    t.last_launches += 1;
}

}  // namespace nbk
