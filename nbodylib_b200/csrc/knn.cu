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
//   knn_fast_kernel   the density family (CalcDensity / CalcVelDensity / smoothing scale), non periodic target
//                     form.  Heap entries are ONE 64-bit word: (fp32 key << 32 | index), key = RN_fp32(d2) of the
//                     exact fp64 d2.  fp64->fp32 rounding is monotone, so the k smallest keys are the exact k
//                     nearest unless the k-th and (k+1)-th keys are EQUAL; the heap carries k+1 entries to see that
//                     case, and such queries (a ~1e-5 fraction on random data) are appended to a list that the
//                     exact kernel re-runs.  Exact d2 for the SPH weights is recomputed from the indices.
//                     8 B/entry instead of 12 and half the shared-memory traffic per sift step; the heap is
//                     bulk-loaded from the k+1 tree-order neighbours of the bucket (heapify) instead of k
//                     full-depth insertions.
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
    int bulk_align;                                   // select + log kernel: the bulk-loaded range starts at a multiple of this (0: centred window)
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

// ================================================================================================= fast
constexpr unsigned FKEY_INF = 0x7f800000u;   // +inf: empty slot

#ifdef NBK_STATS
__device__ unsigned long long g_stats[8];   // 0 tiles, 1 candidate evals (warp-level), 2 rounds, 3 sifts (lane-level), 4 accepted bits, 5 leaves skipped
#define STAT(i, v) do { if (lane_id() == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#define STAT_LANE(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define STAT(i, v)
#define STAT_LANE(i, v)
#endif

// Per-lane 4-ary max-heap in shared memory.  Node p's four children are nodes 4p+1..4p+4 and their fp32 keys sit
// in ONE 16-byte group, so a sift step costs one LDS.128 instead of two dependent 8-byte loads, and a 65-entry
// heap is 3 levels deep instead of 6.  Key of node p: group (p+3)>>2, component (p+3)&3 (node 0 = group 0, comp 3),
// groups are [group][lane] float4 (conflict-free 16-byte lane stride); indices are [node][lane] int32.
struct Heap4 {
    float4* K4;   // [G+1][32]
    int* I;       // [NN][32]
    int G;        // internal nodes: 0..G-1 ; NN = 4G+1 nodes
    unsigned lane;
    __device__ __forceinline__ float* keyp(int p) const { return reinterpret_cast<float*>(K4 + (((p + 3) >> 2) * 32 + lane)) + ((p + 3) & 3); }
    __device__ __forceinline__ int& idx(int p) const { return I[p * 32 + lane]; }
    __device__ __forceinline__ float rootkey() const { return *keyp(0); }
    // place (xk, xi) at node p and sift it down
    __device__ __forceinline__ void sift(int p, float xk, int xi) {
        while (p < G) {
            const float4 ck = K4[(p + 1) * 32 + lane];
            const float m = fmaxf(fmaxf(ck.x, ck.y), fmaxf(ck.z, ck.w));
            if (xk >= m) break;
            const int j = (ck.x == m) ? 0 : ((ck.y == m) ? 1 : ((ck.z == m) ? 2 : 3));
            const int c = 4 * p + 1 + j;
            *keyp(p) = m;
            idx(p) = idx(c);
            p = c;
        }
        *keyp(p) = xk;
        idx(p) = xi;
    }
};

template <class S>
struct FastVisitor {
    const Vec4<S>* P;
    double* tile;
    Heap4 hp;
    double qx, qy, qz;
    double topd;        // (double) of the heap-top key; 0 for lanes without a query (nothing is ever accepted)
    float topf;
    int self;
    int r0, r1;         // tree-index range already loaded into the heap (skipped during the traversal)
    unsigned lane;

    __device__ __forceinline__ bool need(float lb) const { return lb < topf; }
    __device__ __forceinline__ void settop() { topf = hp.rootkey(); topd = (double)topf; }

    template <bool OVERLAP>
    __device__ __forceinline__ void scan_tile(int first, int m) {
        unsigned acc = 0;
        STAT(0, 1); STAT(1, m);
#pragma unroll 4
        for (int j = 0; j < m; j++) {
            const int c = first + j;
            double d2 = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
            bool ok = d2 < topd && d2 > 0.0 && c != self;
            if (OVERLAP) ok = ok && (c < r0 || c >= r1);
            acc |= (ok ? 1u : 0u) << j;
        }
        STAT_LANE(4, __popc(acc));
        while (__any_sync(0xffffffffu, acc != 0)) {
            STAT(2, 1);
            if (acc) {
                int j = __ffs(acc) - 1;
                acc &= acc - 1;
                double d2 = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
                if (d2 < topd) { STAT_LANE(3, 1); hp.sift(0, __double2float_rn(d2), first + j); settop(); }
            }
        }
    }

    __device__ __forceinline__ void leaf(int start, int cnt, int = 0, unsigned = 0) {
        if (start >= r0 && start + cnt <= r1) return;          // warp-uniform: leaf entirely preloaded
        const bool overlap = start < r1 && start + cnt > r0;   // warp-uniform
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[start + base + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
            }
            __syncwarp();
            if (overlap) scan_tile<true>(start + base, m);
            else scan_tile<false>(start + base, m);
        }
    }
};

static inline int heap4_groups(int kcap) { return (kcap - 1 + 3) / 4; }
static inline size_t fast_warp_bytes(int kcap) {
    int G = heap4_groups(kcap);
    return (size_t)(G + 1) * 32 * 16 + (size_t)(4 * G + 1) * 32 * 4 + 96 * 8 + TRAV_STACK * 4;
}

template <class S>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_fast_kernel(KnnParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int kcap = prm.kcap;                        // k + 1 real slots
    const int G = (kcap - 1 + 3) / 4, NN = 4 * G + 1;
    const size_t warp_bytes = (size_t)(G + 1) * 32 * 16 + (size_t)NN * 32 * 4 + 96 * 8 + TRAV_STACK * 4;
    unsigned char* base = smem_raw + w * warp_bytes;
    Heap4 hp;
    hp.K4 = reinterpret_cast<float4*>(base);
    hp.I = reinterpret_cast<int*>(base + (size_t)(G + 1) * 32 * 16);
    hp.G = G; hp.lane = lane;
    double* tile = reinterpret_cast<double*>(base + (size_t)(G + 1) * 32 * 16 + (size_t)NN * 32 * 4);
    int* stack = reinterpret_cast<int*>(base + (size_t)(G + 1) * 32 * 16 + (size_t)NN * 32 * 4 + 96 * 8);

    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const int64_t group = (int64_t)blockIdx.x * KNN_WARPS + w;
    const int64_t g0 = prm.q0 + group * 32;
    if (g0 >= prm.q1) return;
    const int64_t qi = g0 + lane;
    const bool valid = qi < prm.q1 && (!prm.active || prm.active[qi]);

    FastVisitor<S> v;
    v.P = P; v.tile = tile; v.hp = hp; v.lane = lane;
    v.self = valid ? (int)qi : -1;
    double x0 = 0, y0 = 0, z0 = 0;
    if (valid) { Vec4<S> c = P[qi]; x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z; }
    v.qx = x0; v.qy = y0; v.qz = z0;

    // ---- bulk load: the kcap particles around the bucket in tree order, then heapify -----------------------
    {
        // exactly kcap tree positions: every one of them is either skipped for good (self, coincident) or stored,
        // so marking the range as "already seen" for the traversal never loses a candidate
        int64_t want = (int64_t)kcap;
        int64_t r0 = g0 + 16 - want / 2;
        if (r0 + want > prm.n) r0 = prm.n - want;
        if (r0 < 0) r0 = 0;
        int64_t r1 = r0 + want;
        if (r1 > prm.n) r1 = prm.n;
        v.r0 = (int)r0; v.r1 = (int)r1;
        int filled = 0;
        for (int64_t b0 = r0; b0 < r1; b0 += 32) {
            int m = (int)min((int64_t)32, r1 - b0);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[b0 + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
            }
            __syncwarp();
            for (int j = 0; j < m; j++) {
                double d2 = dist2_ref(x0, y0, z0, tile[j], tile[32 + j], tile[64 + j]);
                if (valid && (int)(b0 + j) != v.self && d2 > 0.0) {
                    *v.hp.keyp(filled) = __double2float_rn(d2);
                    v.hp.idx(filled) = (int)(b0 + j);
                    filled++;
                }
            }
        }
        // empty real slots wait for candidates (+inf); the padding up to 4G+1 nodes, and every slot of a lane without
        // a query, holds key 0 / index -1: never evicted, never accepted against
        for (; filled < NN; filled++) {
            *v.hp.keyp(filled) = (valid && filled < kcap) ? __uint_as_float(FKEY_INF) : 0.f;
            v.hp.idx(filled) = -1;
        }
        for (int p = G - 1; p >= 0; p--) v.hp.sift(p, *v.hp.keyp(p), v.hp.idx(p));
        v.settop();
    }
    {
        QueryBox qb = make_qbox(x0, y0, z0);
        traverse(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid);
    }
    if (!valid) return;

    // ---- exactness test: drop the (k+1)-th; the k smallest keys are the exact kNN iff key_k < key_{k+1} -------
    const float key_kp1 = v.hp.rootkey();
    v.hp.sift(0, 0.f, -1);                                       // the root becomes padding
    const float key_k = v.hp.rootkey();
    const bool short_of_k = __float_as_uint(key_k) == FKEY_INF;  // fewer than k candidates exist (n <= k)
    if (key_k == key_kp1 && !short_of_k) {
        int slot = atomicAdd(prm.flag_count, 1);
        prm.flag_list[slot] = (int)qi;
        return;                                                  // the exact kernel redoes this query entirely
    }
    // exact k-th distance: the largest exact d2 among the entries sharing the top key
    double d2max = 0;
    for (int s = 0; s < NN; s++) {
        int id = v.hp.idx(s);
        if (id >= 0 && *v.hp.keyp(s) == key_k) {
            Vec4<S> c = P[id];
            d2max = fmax(d2max, dist2_ref(x0, y0, z0, (double)c.x, (double)c.y, (double)c.z));
        }
    }
    if (short_of_k) d2max = KNN_SENTINEL;
    if (prm.hsm) prm.hsm[qi] = 0.5 * sqrt(d2max);
    if (prm.rho && prm.veldens_k == 0) {
        const double hi = 0.5 * sqrt(d2max);
        const double norm = 1.0 / pow(hi, 3.0);
        const double delta = 2.0 / (double)(prm.kernres - 1);
        const double mi = prm.mass[qi];
        double acc = 0;
        for (int s = 0; s < NN; s++) {
            int id = v.hp.idx(s);
            if (id < 0) continue;
            Vec4<S> c = P[id];
            double rij = sqrt(dist2_ref(x0, y0, z0, (double)c.x, (double)c.y, (double)c.z));
            double r = rij / hi;
            double Wij = 0.5 * wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
            acc += Wij * prm.mass[id];
            atomicAdd(&prm.rho[id], Wij * mi);
        }
        atomicAdd(&prm.rho[qi], acc);
    }
    if (prm.rho && prm.veldens_k > 0) {
        // R2.  The fp64 velocity distances (up to k per lane) are written over the lane's own heap storage, which is dead
        // by now: doubles 0..2G+1 over the lane's key groups, the rest over pairs of the lane's index slots that have
        // already been consumed (double t >= 2G+2 uses index slots 2u, 2u+1 with u = t-2G-2 <= s-2G-2, and 2u+1 <= s for
        // every read position s <= 4G).  Everything stays lane-private, so no cross-lane ordering is needed.
        const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
        const int nA = 2 * (G + 1);
        auto dget = [&](int t) -> double {
            if (t < nA) return reinterpret_cast<const double*>(v.hp.K4 + ((t >> 1) * 32 + lane))[t & 1];
            const int u = t - nA;
            return __hiloint2double(v.hp.I[(2 * u + 1) * 32 + lane], v.hp.I[(2 * u) * 32 + lane]);
        };
        auto dset = [&](int t, double d) {
            if (t < nA) { reinterpret_cast<double*>(v.hp.K4 + ((t >> 1) * 32 + lane))[t & 1] = d; return; }
            const int u = t - nA;
            v.hp.I[(2 * u) * 32 + lane] = __double2loint(d);
            v.hp.I[(2 * u + 1) * 32 + lane] = __double2hiint(d);
        };
        Vec4<S> vi = V[qi];
        int kx = 0;
        for (int s = 0; s < NN; s++) {
            int id = v.hp.idx(s);
            if (id < 0) continue;
            Vec4<S> vj = V[id];
            dset(kx, sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z)));
            kx++;
        }
        auto dsift = [&](int p, int n, double d) {      // binary max-heap on the doubles
            while (true) {
                int c = 2 * p + 1;
                if (c >= n) break;
                double dc = dget(c);
                if (c + 1 < n) { double dr = dget(c + 1); if (dr > dc) { c = c + 1; dc = dr; } }
                if (d >= dc) break;
                dset(p, dc);
                p = c;
            }
            dset(p, d);
        };
        const int kv = min(prm.veldens_k, kx);
        double rho = 0;
        if (kv > 0) {
            for (int p = kv / 2 - 1; p >= 0; p--) dsift(p, kv, dget(p));
            for (int s = kv; s < kx; s++) {
                double vd = dget(s);
                if (vd < dget(0)) dsift(0, kv, vd);
            }
            const double hi = 0.5 * dget(0);
            const double norm = 1.0 / pow(hi, 3.0);
            const double delta = 2.0 / (double)(prm.kernres - 1);
            // pop in descending order like the reference so the sum is accumulated in the same order
            for (int e = kv; e > 0; e--) {
                double rij = dget(0);
                double r = rij / hi;
                rho = rho + wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
                dsift(0, e - 1, dget(e - 1));
            }
        }
        prm.rho[qi] = rho;
    }
}

// ==================================================================================== select-then-collect
// The density family needs only (a) the exact k-th distance and (b) a sum over the k nearest neighbours, so the
// search is split in two traversals sharing one small shared-memory region per warp:
//   select  : per-lane 4-ary max-heap of fp32 KEYS ONLY (k+1 of them) -> key_k, key_{k+1}.  No indices are carried,
//             so a sift step moves 4 bytes, and the region is half the size of a (key,index) heap: twice the resident
//             warps for this latency-bound kernel.
//   collect : fixed-radius pass, every candidate with RN_fp32(d2) <= key_k is appended to the lane's index list
//             (exactly the k nearest when key_k < key_{k+1}; otherwise the query goes to the exact kernel); appends
//             are predicated stores, there are no serial insertion rounds.  The exact fp64 k-th distance is the max
//             over the appended candidates.
//   epilogue: SPH sums over the list, as before.
struct KeyHeap4 {
    unsigned char* kb;   // this lane's byte base inside the [G+1][32] float4 groups (group g of the lane at kb + g*512)
    int G;
    // key of node p: group (p+3)>>2, component (p+3)&3 ; children of p = the four components of group p+1
    __device__ __forceinline__ float* keyp(int p) const { return reinterpret_cast<float*>(kb + ((p + 3) >> 2) * 512 + ((p + 3) & 3) * 4); }
    __device__ __forceinline__ float rootkey() const { return *reinterpret_cast<const float*>(kb + 12); }
    // place xk at node p and sift it down; returns the key that ends up at node p
    __device__ __forceinline__ float sift(int p, float xk) {
        unsigned char* pa = reinterpret_cast<unsigned char*>(keyp(p));
        unsigned char* ga = kb + (p + 1) * 512;
        float at_p = xk;
        bool moved = false;
        while (p < G) {
            const float4 ck = *reinterpret_cast<const float4*>(ga);
            const bool a = ck.x >= ck.y, b = ck.z >= ck.w;
            const float m01 = a ? ck.x : ck.y, m23 = b ? ck.z : ck.w;
            const bool c = m01 >= m23;
            const float m = c ? m01 : m23;
            if (xk >= m) break;
            const int j = c ? (a ? 0 : 1) : (b ? 2 : 3);
            *reinterpret_cast<float*>(pa) = m;
            if (!moved) { at_p = m; moved = true; }
            pa = ga + 4 * j;
            p = 4 * p + 1 + j;
            ga = kb + (p + 1) * 512;
        }
        *reinterpret_cast<float*>(pa) = xk;
        return at_p;
    }
};

// Leaf tile staged in shared memory + the per-candidate tests of the two passes.
// fp32 storage: the tile keeps the float4 records (one LDS.128 per candidate) and candidates are first screened with
// an fp32 distance: inputs are exact, the fp32 evaluation has a relative error below 5 * 2^-24, so a candidate whose
// fp32 d2 exceeds limit * (1 + 2^-20) cannot pass the exact fp64 test, which is then evaluated only for the few
// survivors.  fp64 storage: no screen (fp32 rounding of the coordinates would not bound the error), exact test only.
template <class S> struct LeafTile;
template <> struct LeafTile<float> {
    static constexpr int TILE_BYTES = 32 * 16;
    float4* t;
    float qxf, qyf, qzf;
    __device__ __forceinline__ void init(void* mem, double qx, double qy, double qz) {
        t = reinterpret_cast<float4*>(mem);
        qxf = (float)qx; qyf = (float)qy; qzf = (float)qz;
    }
    __device__ __forceinline__ void load(const Vec4<float>* P, int first, int m, unsigned lane) {
        // slots past the end of the leaf hold NaN: every screen / comparison on them is false, so the scan loops can run
        // over whole groups of 8 without a bound check
        const float nanf_ = __int_as_float(0x7fc00000);
        float4 mine = make_float4(nanf_, nanf_, nanf_, 0.f);
        if ((int)lane < m) { Vec4<float> c = P[first + lane]; mine = make_float4(c.x, c.y, c.z, 0.f); }
        t[lane] = mine;
    }
    static __device__ __forceinline__ float screen_limit(float lim) { return __fmul_ru(lim, 1.00000095367431640625f); }
    static __device__ __forceinline__ bool screen_one(float qx, float qy, float qz, const float4& c, float limf) {
        const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
        return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))) <= limf;     // 3 sub + mul + 2 fma; error < 4 * 2^-24
    }
    __device__ __forceinline__ bool screen(int j, float limf, double) const { return screen_one(qxf, qyf, qzf, t[j], limf); }
    __device__ __forceinline__ double exact(int j, double qx, double qy, double qz) const {
        const float4 c = t[j];
        return dist2_ref(qx, qy, qz, (double)c.x, (double)c.y, (double)c.z);
    }
};
template <> struct LeafTile<double> {
    static constexpr int TILE_BYTES = 96 * 8;
    double* t;
    double qx_, qy_, qz_;
    __device__ __forceinline__ void init(void* mem, double qx, double qy, double qz) { t = reinterpret_cast<double*>(mem); qx_ = qx; qy_ = qy; qz_ = qz; }
    __device__ __forceinline__ void load(const Vec4<double>* P, int first, int m, unsigned lane) {
        const double nan_ = __longlong_as_double(0x7ff8000000000000ll);
        double cx = nan_, cy = nan_, cz = nan_;
        if ((int)lane < m) { Vec4<double> c = P[first + lane]; cx = c.x; cy = c.y; cz = c.z; }
        t[lane] = cx; t[32 + lane] = cy; t[64 + lane] = cz;
    }
    static __device__ __forceinline__ float screen_limit(float lim) { return lim; }
    __device__ __forceinline__ bool screen(int j, float, double limd) const { return dist2_ref(qx_, qy_, qz_, t[j], t[32 + j], t[64 + j]) < limd; }
    __device__ __forceinline__ double exact(int j, double qx, double qy, double qz) const { return dist2_ref(qx, qy, qz, t[j], t[32 + j], t[64 + j]); }
};

constexpr int SC_LEAFCAP = 240;   // leaves remembered by the select pass for the collect pass (per warp)

template <class S, bool LOG = false>
struct SelectVisitor {
    int* log;           // LOG: this lane's column of the warp's [logcap][32] insertion log (global scratch)
    int nlog, logcap;
    const Vec4<S>* P;
    LeafTile<S> tile;
    KeyHeap4 hp;
    double qx, qy, qz;
    double topd;        // (double) of the heap-top key; 0 for lanes without a query (nothing is ever accepted)
    float topf, limf;   // limf: fp32 screening limit derived from topf
    int r0, r1;         // tree-index range bulk-loaded into the heap (skipped during the traversal)
    int* leaflist;      // [SC_LEAFCAP] node index of every leaf scanned, warp-uniform
    int nleaf;
    unsigned lane;
    __device__ __forceinline__ bool need(float lb) const { return lb < topf; }
    __device__ __forceinline__ void settop(float k) { topf = k; topd = (double)k; limf = LeafTile<S>::screen_limit(k); }
    template <bool OVERLAP>
    __device__ __forceinline__ void scan_tile(int first, int m, unsigned nmask) {
        // pass 1: one cheap test per candidate (the query itself and coincident particles have d2 == 0 and are weeded out
        // in pass 2, like bulk-loaded candidates); pass 2: serial insertion rounds over the set bits
        unsigned acc = 0;
        STAT(0, 1); STAT(1, m);
        for (int j0 = 0; j0 < m; j0 += 8) {
            unsigned a8 = 0;
#pragma unroll
            for (int jj = 0; jj < 8; jj++) if (tile.screen(j0 + jj, limf, topd)) a8 |= 1u << jj;
            acc |= a8 << j0;
        }
        STAT_LANE(4, __popc(acc));
        while (__any_sync(0xffffffffu, acc != 0)) {
            STAT(2, 1);
            if (acc) {
                const int j = __ffs(acc) - 1;
                acc &= acc - 1;
                const int c = first + j;
                const double d2 = tile.exact(j, qx, qy, qz);
                if (d2 < topd && d2 > 0.0 && (!OVERLAP || c < r0 || c >= r1)) {
                    settop(hp.sift(0, __double2float_rn(d2)));
                    STAT_LANE(3, 1);
                    if (LOG) { if (nlog < logcap) { *log = c; log += 32; } nlog++; }
                }
            }
        }
    }
    __device__ __forceinline__ void leaf(int start, int cnt, int node, unsigned nmask) {
        if (start >= r0 && start + cnt <= r1) return;
        if (!LOG) {
            if (nleaf < SC_LEAFCAP && lane == 0) leaflist[nleaf] = node;
            nleaf++;
        }
        const bool overlap = start < r1 && start + cnt > r0;
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            tile.load(P, start + base, m, lane);
            __syncwarp();
            if (overlap) scan_tile<true>(start + base, m, nmask);
            else scan_tile<false>(start + base, m, nmask);
        }
    }
};

template <class S>
struct CollectVisitor {
    const Vec4<S>* P;
    LeafTile<S> tile;
    int* L;            // this lane's column of the [k][32] index list (entry s at L[s*32])
    double qx, qy, qz;
    double d2max;
    double thr_d;      // candidates with d2 >= thr_d cannot qualify (fp64 pre-test)
    float thr, limf;   // qualify iff RN_fp32(d2) <= thr ; -1 for lanes that collect nothing
    int cnt, cap;
    int r0, r1;        // with skip_range: candidates in [r0,r1) are left to the separate scan of the bulk range
    bool skip_bulk;    // traversal form: leave [r0,r1) out (it was scanned separately)
    unsigned lane;
    __device__ __forceinline__ bool need(float lb) const { return lb <= thr; }
    template <bool SKIP>
    __device__ __forceinline__ void scan(int start, int n, unsigned nmask) {
        for (int base = 0; base < n; base += 32) {
            int m = min(32, n - base);
            __syncwarp();
            tile.load(P, start + base, m, lane);
            __syncwarp();
            unsigned acc = 0;
            for (int j0 = 0; j0 < m; j0 += 8) {
                unsigned a8 = 0;
#pragma unroll
                for (int jj = 0; jj < 8; jj++) if (tile.screen(j0 + jj, limf, thr_d)) a8 |= 1u << jj;
                acc |= a8 << j0;
            }
            while (__any_sync(0xffffffffu, acc != 0)) {
                if (acc) {
                    const int j = __ffs(acc) - 1;
                    acc &= acc - 1;
                    const int c = start + base + j;
                    const double d2 = tile.exact(j, qx, qy, qz);
                    if (d2 < thr_d && d2 > 0.0 && __double2float_rn(d2) <= thr && cnt < cap && (!SKIP || c < r0 || c >= r1)) {
                        L[cnt * 32] = c;
                        cnt++;
                        d2max = fmax(d2max, d2);
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void leaf(int start, int n, int, unsigned nmask) {
        if (skip_bulk) scan<true>(start, n, nmask);
        else scan<false>(start, n, nmask);
    }
};

// SPH epilogues over a lane's neighbour list L (entry s at L[s*32]; `base` = the warp's shared region, whose part after
// the list holds the per-lane doubles of the kv < kx velocity-density selection).
template <class S>
__device__ __forceinline__ void sc_epilogue(const KnnParams& prm, const Vec4<S>* __restrict__ P, const int* L, unsigned char* base, unsigned lane,
                                            int k, int cnt, double d2max, double x0, double y0, double z0, int64_t qi) {
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
            // kv < kx: exact fp64 selection on a per-lane array of doubles placed after the list (want_doubles layout)
            double* D = reinterpret_cast<double*>(base + (size_t)k * 32 * 4);
            for (int s = 0; s < cnt; s++) {
                Vec4<S> vj = V[L[s * 32]];
                D[s * 32 + lane] = sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z));
            }
            auto dsift = [&](int p, int n, double d) {
                while (true) {
                    int c = 2 * p + 1;
                    if (c >= n) break;
                    double dc = D[c * 32 + lane];
                    if (c + 1 < n) { double dr = D[(c + 1) * 32 + lane]; if (dr > dc) { c = c + 1; dc = dr; } }
                    if (d >= dc) break;
                    D[p * 32 + lane] = dc;
                    p = c;
                }
                D[p * 32 + lane] = d;
            };
            for (int p = kv / 2 - 1; p >= 0; p--) dsift(p, kv, D[p * 32 + lane]);
            for (int s = kv; s < cnt; s++) {
                double vd = D[s * 32 + lane];
                if (vd < D[lane]) dsift(0, kv, vd);
            }
            const double hi = 0.5 * D[lane];
            const double norm = 1.0 / pow(hi, 3.0);
            for (int e = kv; e > 0; e--) {
                double r = D[lane] / hi;
                rho = rho + wsm(r, (int)(r * 0.5 * (prm.kernres - 1)), prm.kernres, delta, prm.kern) * norm;
                dsift(0, e - 1, D[(e - 1) * 32 + lane]);
            }
        }
        prm.rho[qi] = rho;
    }
}

static inline size_t sc_warp_bytes(int k, bool want_doubles) {
    int G = heap4_groups(k + 1);
    size_t keys = (size_t)(G + 1) * 32 * 16, list = (size_t)k * 32 * 4;
    size_t region = keys > list ? keys : list;
    if (want_doubles) region = list + (size_t)k * 32 * 8 > region ? list + (size_t)k * 32 * 8 : region;
    return region + 96 * 8 + TRAV_STACK * 4 + SC_LEAFCAP * 4;
}

template <class S>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_sc_kernel(KnnParams prm, int want_doubles) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int k = prm.k, kcap = k + 1;
    const int G = (kcap - 1 + 3) / 4, NN = 4 * G + 1;
    size_t region = (size_t)(G + 1) * 32 * 16;
    if ((size_t)k * 32 * 4 > region) region = (size_t)k * 32 * 4;
    if (want_doubles && (size_t)k * 32 * 12 > region) region = (size_t)k * 32 * 12;
    const size_t warp_bytes = region + 96 * 8 + TRAV_STACK * 4 + SC_LEAFCAP * 4;
    unsigned char* base = smem_raw + w * warp_bytes;
    void* tile_mem = base + region;
    int* stack = reinterpret_cast<int*>(base + region + 96 * 8);
    int* leaflist = reinterpret_cast<int*>(base + region + 96 * 8 + TRAV_STACK * 4);

    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const int64_t group = (int64_t)blockIdx.x * KNN_WARPS + w;
    const int64_t g0 = prm.q0 + group * 32;
    if (g0 >= prm.q1) return;
    const int64_t qi = g0 + lane;
    const bool valid = qi < prm.q1 && (!prm.active || prm.active[qi]);
    double x0 = 0, y0 = 0, z0 = 0;
    if (valid) { Vec4<S> c = P[qi]; x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z; }
    const QueryBox qb = make_qbox(x0, y0, z0);
    int nleaf, r0i, r1i;

    // ---------------------------------------------------------------------------------------------- select
    float key_k, key_kp1;
    {
        SelectVisitor<S> v;
        v.P = P; v.lane = lane;
        v.tile.init(tile_mem, x0, y0, z0);
        v.leaflist = leaflist; v.nleaf = 0;
        v.hp.kb = base + lane * 16; v.hp.G = G;
        const int self = valid ? (int)qi : -1;
        v.qx = x0; v.qy = y0; v.qz = z0;
        int64_t want = (int64_t)kcap;
        int64_t r0 = g0 + 16 - want / 2;
        if (r0 + want > prm.n) r0 = prm.n - want;
        if (r0 < 0) r0 = 0;
        int64_t r1 = r0 + want;
        if (r1 > prm.n) r1 = prm.n;
        v.r0 = (int)r0; v.r1 = (int)r1;
        int filled = 0;
        for (int64_t b0 = r0; b0 < r1; b0 += 32) {
            int m = (int)min((int64_t)32, r1 - b0);
            __syncwarp();
            v.tile.load(P, (int)b0, m, lane);
            __syncwarp();
            for (int j = 0; j < m; j++) {
                double d2 = v.tile.exact(j, x0, y0, z0);
                if (valid && (int)(b0 + j) != self && d2 > 0.0) { *v.hp.keyp(filled) = __double2float_rn(d2); filled++; }
            }
        }
        for (; filled < NN; filled++) *v.hp.keyp(filled) = (valid && filled < kcap) ? __uint_as_float(FKEY_INF) : 0.f;
        for (int p = G - 1; p >= 0; p--) v.hp.sift(p, *v.hp.keyp(p));
        v.settop(v.hp.rootkey());
        traverse(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid);
        key_kp1 = v.hp.rootkey();
        v.hp.sift(0, 0.f);
        key_k = v.hp.rootkey();
        nleaf = v.nleaf; r0i = v.r0; r1i = v.r1;
    }
    const bool short_of_k = __float_as_uint(key_k) == FKEY_INF;
    const bool flagged = valid && key_k == key_kp1 && !short_of_k;
    if (flagged) {
        int slot = atomicAdd(prm.flag_count, 1);
        prm.flag_list[slot] = (int)qi;
    }
    __syncwarp();   // the key groups are dead from here on: the region is reused for the index list

    // --------------------------------------------------------------------------------------------- collect
    CollectVisitor<S> c2;
    c2.P = P; c2.L = reinterpret_cast<int*>(base) + lane; c2.lane = lane; c2.skip_bulk = false;
    c2.tile.init(tile_mem, x0, y0, z0);
    c2.qx = x0; c2.qy = y0; c2.qz = z0;
    c2.r0 = r0i; c2.r1 = r1i;
    c2.cnt = 0; c2.cap = k; c2.d2max = 0;
    c2.thr = (valid && !flagged) ? key_k : -1.f;
    c2.thr_d = (valid && !flagged) ? (double)__uint_as_float(__float_as_uint(key_k) + (short_of_k ? 0u : 1u)) : -1.0;   // next float above key_k
    if (short_of_k && valid) c2.thr_d = 3.0e38 * 10.0;
    c2.limf = LeafTile<S>::screen_limit(c2.thr);
    const bool collecting = valid && !flagged;
    if (nleaf <= SC_LEAFCAP) {
        // every leaf that can hold one of the k nearest was scanned by the select pass (a lane's final neighbours were
        // below its bound at all times): re-scan exactly those tiles, plus the bulk-loaded range, without walking the tree
        c2.template scan<false>(r0i, r1i - r0i, 0xffffffffu);
        for (int t = 0; t < nleaf; t++) {
            const int node = leaflist[t];
            const NodeLo lo = prm.nlo[node];
            const NodeHi hi = prm.nhi[node];
            const float lb = box_lb(qb.lx, qb.ly, qb.lz, qb.hx, qb.hy, qb.hz, lo, hi);
            const unsigned nmask = __ballot_sync(0xffffffffu, collecting && c2.need(lb));
            if (!nmask) continue;                                                  // the bounds have tightened since
            c2.template scan<true>(lo.start, hi.end - lo.start, nmask);
        }
    } else {
        traverse(prm.nlo, prm.nhi, prm.bucket, stack, c2, qb, collecting);
    }
    if (!collecting) return;

    // -------------------------------------------------------------------------------------------- epilogues
    sc_epilogue<S>(prm, P, c2.L, base, lane, k, c2.cnt, short_of_k ? KNN_SENTINEL : c2.d2max, x0, y0, z0, qi);
}

// =========================================================================================== select + log
// Same select pass, but every candidate that enters a lane's heap is also appended to the lane's insertion LOG in a
// global scratch (one [logcap][32] block per resident warp, written once and read once, so it lives in L2).  Every one of
// the final k nearest was inserted at some point (its d2 was below the lane's bound at all times), so the collect pass
// is a filter over the log (~1.4 k entries per lane) plus the bulk-loaded range instead of a second scan of ~65 leaf tiles.
// Persistent grid: warps draw 32-query groups from a global counter, so a warp's scratch block is reused and the
// grid is exactly one wave whatever the particle count.
static inline size_t sl_warp_bytes(int k, bool want_doubles, int tile_bytes) {
    int G = heap4_groups(k + 1);
    size_t keys = (size_t)(G + 1) * 32 * 16, list = (size_t)k * 32 * 4;
    size_t region = keys > list ? keys : list;
    if (want_doubles) region = list + (size_t)k * 32 * 8 > region ? list + (size_t)k * 32 * 8 : region;
    return region + tile_bytes + TRAV_STACK * 4;
}

template <class S, int MB, bool HALO>
__global__ void __launch_bounds__(KNN_WARPS * 32, MB) knn_sl_kernel(KnnParams prm, int want_doubles, int* __restrict__ work_counter,
                                                                int32_t* __restrict__ logbuf, int logcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const int k = prm.k, kcap = k + 1;
    const int G = (kcap - 1 + 3) / 4, NN = 4 * G + 1;
    size_t region = (size_t)(G + 1) * 32 * 16;
    if ((size_t)k * 32 * 4 > region) region = (size_t)k * 32 * 4;
    if (want_doubles && (size_t)k * 32 * 12 > region) region = (size_t)k * 32 * 12;
    const size_t warp_bytes = region + LeafTile<S>::TILE_BYTES + TRAV_STACK * 4;
    unsigned char* base = smem_raw + w * warp_bytes;
    void* tile_mem = base + region;
    int* stack = reinterpret_cast<int*>(base + region + LeafTile<S>::TILE_BYTES);
    int* log = logbuf + ((size_t)blockIdx.x * KNN_WARPS + w) * (size_t)logcap * 32 + lane;

    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    const int64_t ngroups = (prm.q1 - prm.q0 + 31) >> 5;
    while (true) {
        int64_t group = 0;
        if (lane == 0) group = (int64_t)atomicAdd(work_counter, 1);
        group = __shfl_sync(0xffffffffu, group, 0);
        if (group >= ngroups) break;
        const int64_t g0 = prm.q0 + group * 32;
        const int64_t qi = g0 + lane;
        const bool valid = qi < prm.q1 && (!prm.active || prm.active[qi]);
        if (!__any_sync(0xffffffffu, valid)) continue;
        double x0 = 0, y0 = 0, z0 = 0;
        if (valid) { Vec4<S> c = P[qi]; x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z; }
        const QueryBox qb = make_qbox(x0, y0, z0);
        int r0i, r1i, nlog;

        // ------------------------------------------------------------------------------------------ select
        float key_k, key_kp1;
        {
            SelectVisitor<S, true> v;
            v.P = P; v.lane = lane;
            v.tile.init(tile_mem, x0, y0, z0);
            v.leaflist = nullptr; v.nleaf = 0;
            v.log = log; v.nlog = 0; v.logcap = logcap;
            v.hp.kb = base + lane * 16; v.hp.G = G;
            const int self = valid ? (int)qi : -1;
            v.qx = x0; v.qy = y0; v.qz = z0;
            int64_t want = (int64_t)kcap;
            int64_t r0 = g0 + 16 - want / 2;
            // ... or, better, the node-aligned block that holds the group: with 2^m particles the tree positions [j*64, (j+1)*64)
            // are one node, a compact set, whereas a window centred on the group straddles three nodes whose tree-order
            // neighbours can lie across a high-level cut plane -- a looser first bound and more insertions (k = 64: 428 -> 413 ms)
            if (prm.bulk_align > 0) r0 = g0 - (g0 % prm.bulk_align);
            if (r0 + want > prm.n) r0 = prm.n - want;
            if (r0 < 0) r0 = 0;
            int64_t r1 = r0 + want;
            if (r1 > prm.n) r1 = prm.n;
            v.r0 = (int)r0; v.r1 = (int)r1;
            int filled = 0;
            for (int64_t b0 = r0; b0 < r1; b0 += 32) {
                int m = (int)min((int64_t)32, r1 - b0);
                __syncwarp();
                v.tile.load(P, (int)b0, m, lane);
                __syncwarp();
                for (int j = 0; j < m; j++) {
                    double d2 = v.tile.exact(j, x0, y0, z0);
                    if (valid && (int)(b0 + j) != self && d2 > 0.0) { *v.hp.keyp(filled) = __double2float_rn(d2); filled++; }
                }
            }
            for (; filled < NN; filled++) *v.hp.keyp(filled) = (valid && filled < kcap) ? __uint_as_float(FKEY_INF) : 0.f;
            for (int p = G - 1; p >= 0; p--) v.hp.sift(p, *v.hp.keyp(p));
            v.settop(v.hp.rootkey());
            traverse(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid);
            if (HALO) traverse(prm.nlo2, prm.nhi2, prm.bucket2, stack, v, qb, valid);     // attached halo tree: its own instantiation,
                                                                                          // the second inlined walk costs the plain kernel 3 %
            key_kp1 = v.hp.rootkey();
            v.hp.sift(0, 0.f);
            key_k = v.hp.rootkey();
            r0i = v.r0; r1i = v.r1; nlog = v.nlog;
        }
        const bool short_of_k = __float_as_uint(key_k) == FKEY_INF;
        // inexact key order: the exact kernel redoes the query
        const bool flagged = valid && key_k == key_kp1 && !short_of_k;
        if (flagged) {
            int slot = atomicAdd(prm.flag_count, 1);
            prm.flag_list[slot] = (int)qi;
        }
        __syncwarp();   // the key groups are dead from here on: the region is reused for the index list

        // ----------------------------------------------------------------------------------------- collect
        const bool collecting = valid && !flagged;
        CollectVisitor<S> c2;
        c2.P = P; c2.L = reinterpret_cast<int*>(base) + lane; c2.lane = lane; c2.skip_bulk = true;
        c2.tile.init(tile_mem, x0, y0, z0);
        c2.qx = x0; c2.qy = y0; c2.qz = z0;
        c2.r0 = r0i; c2.r1 = r1i;
        c2.cnt = 0; c2.cap = k; c2.d2max = 0;
        c2.thr = collecting ? key_k : -1.f;
        c2.thr_d = collecting ? (double)__uint_as_float(__float_as_uint(key_k) + (short_of_k ? 0u : 1u)) : -1.0;   // next float above key_k
        if (short_of_k && collecting) c2.thr_d = 3.0e38 * 10.0;
        c2.limf = LeafTile<S>::screen_limit(c2.thr);
        c2.template scan<false>(r0i, r1i - r0i, 0xffffffffu);      // the bulk-loaded range
        const bool overflowed = collecting && nlog > logcap;      // log incomplete: this lane collects by a second traversal
        {
            const int mynl = (collecting && !overflowed) ? nlog : 0;
            const int nmax = __reduce_max_sync(0xffffffffu, mynl);
            int cnt = c2.cnt;
            double d2max = c2.d2max;
            for (int s0 = 0; s0 < nmax; s0 += 4) {
                int cidx[4];
                Vec4<S> pc[4];
#pragma unroll
                for (int u = 0; u < 4; u++) cidx[u] = (s0 + u < mynl) ? log[(s0 + u) * 32] : -1;
#pragma unroll
                for (int u = 0; u < 4; u++) if (cidx[u] >= 0) pc[u] = P[cidx[u]];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (cidx[u] >= 0) {
                        const double d2 = dist2_ref(x0, y0, z0, (double)pc[u].x, (double)pc[u].y, (double)pc[u].z);
                        if (d2 < c2.thr_d && __double2float_rn(d2) <= c2.thr && cnt < k) {
                            c2.L[cnt * 32] = cidx[u];
                            cnt++;
                            d2max = fmax(d2max, d2);
                        }
                    }
                }
            }
            c2.cnt = cnt; c2.d2max = d2max;
        }
        if (__any_sync(0xffffffffu, overflowed)) {
            const float thr_keep = c2.thr, limf_keep = c2.limf;
            const double thrd_keep = c2.thr_d;
            if (!overflowed) { c2.thr = -1.f; c2.thr_d = -1.0; c2.limf = -1.f; }      // the other lanes are complete
            traverse(prm.nlo, prm.nhi, prm.bucket, stack, c2, qb, overflowed);
            if (HALO) traverse(prm.nlo2, prm.nhi2, prm.bucket2, stack, c2, qb, overflowed);
            c2.thr = thr_keep; c2.limf = limf_keep; c2.thr_d = thrd_keep;
        }
        if (collecting) sc_epilogue<S>(prm, P, c2.L, base, lane, k, c2.cnt, short_of_k ? KNN_SENTINEL : c2.d2max, x0, y0, z0, qi);
        __syncwarp();
    }
}


static void fill_common(KnnParams& p, nbk_tree& t, const KnnArgs& a) {
    p.nlo = t.nlo; p.nhi = t.nhi; p.bucket = t.bucket;
    p.nlo2 = t.nlo2; p.nhi2 = t.nhi2; p.bucket2 = t.bucket;
    p.P = t.prim; p.V = t.vel4(); p.mass = t.mass; p.order = t.order;
    p.n = t.n;
    p.q0 = a.q0; p.q1 = a.q1; p.xq = a.xq; p.mode = a.mode;
    p.qlist = a.qlist; p.nq = a.nq;
    p.gather = a.gather ? 1 : 0; p.vq = a.vq;
    p.cand_excl = a.cand_excl; p.crit_mode = a.crit_mode; p.cp0 = a.cp0; p.cp1 = a.cp1;
    p.rho_in = a.rho_in; p.smvel_in = a.smvel_in; p.smvel_out = a.smvel_out; p.smdisp_out = a.smdisp_out;
    {
        int al = 32;
        while (al * 2 <= a.k) al *= 2;                     // largest power of two <= k, at least the 32-query group
        p.bulk_align = getenv("NBK_KNN_BULK_ALIGN") ? atoi(getenv("NBK_KNN_BULK_ALIGN")) : al;
    }
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
    if (smooth_only && getenv("NBK_KNN_EXACT_ONLY") == nullptr) {
        // ---- fast path + exact fallback for the flagged queries ----------------------------------------------
        p.kcap = a.k + 1;
        const char* em = getenv("NBK_KNN_MODE");
        int mode = em ? atoi(em) : 2;           // 2: select + insertion log (default), 1: select-then-collect, 0: (key,index) heap
        if (t.nlo2) mode = 2;                   // only the default kernel (and the exact one) walk an attached halo tree
        // nodes of up to `leaf` particles are scanned as one tile: fewer node tests and better balanced insertion rounds
        // The level whose nodes hold 21..40 particles (exactly one level does: sizes halve) -- a tile and a bit; with a fixed
        // threshold of 32 a particle count just above a power of two would be scanned as half-empty 16/17-particle tiles.
        {
            const char* e = getenv("NBK_KNN_LEAF");
            int64_t sz = t.n_main ? t.n_main : t.n;
            while (sz > 40) sz = (sz + 1) / 2;
            int leaf = e ? atoi(e) : (int)sz;
            if (leaf > p.bucket) p.bucket = leaf;
            if (t.nlo2) {
                sz = t.n - t.n_main;
                while (sz > 40) sz = (sz + 1) / 2;
                leaf = e ? atoi(e) : (int)sz;
                if (leaf > p.bucket2) p.bucket2 = leaf;
            }
        }
        const int want_doubles = (a.veldens_k > 0 && a.veldens_k < a.k) ? 1 : 0;
        if (mode == 2) {
            size_t smem = sl_warp_bytes(a.k, want_doubles, t.store_bytes == 4 ? 32 * 16 : 96 * 8) * KNN_WARPS;
            NBK_REQUIRE(smem <= 227 * 1024, NBK_ERR_ARG, "k too large for the shared-memory heaps");
            const char* e = getenv("NBK_KNN_LOGCAP");
            int logcap = e ? atoi(e) : 4 * a.k;
            if (logcap < 64) logcap = 64;
            int nsm = 0, per_sm = 0;
            NBK_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, t.device));
            // register budget follows the shared-memory budget: MB = CTAs per SM the kernel is compiled for (5: 96, 6: 80, 8: 64 regs)
            int fit = (int)(233472 / (smem + 1024));
            const char* emb = getenv("NBK_KNN_MB");
            if (emb) fit = atoi(emb);
            const int mb = fit >= 8 ? 8 : (fit >= 6 ? 6 : 5);
            const int64_t ngroups = (rows + 31) / 32;
            DevBuf<int> counters(2);
            DevBuf<int32_t> flist(rows);
            DevBuf<int32_t> logbuf;
            NBK_CHECK(cudaMemsetAsync(counters.p, 0, 2 * sizeof(int), t.stream));
            p.flag_count = counters.p; p.flag_list = flist.p;
            bool window = false;
            size_t old_persist_limit = 0;
            auto go = [&](auto kern) {
                NBK_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                NBK_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, KNN_WARPS * 32, smem));
                NBK_REQUIRE(per_sm >= 1, NBK_ERR_ARG, "k too large for the shared-memory heaps");
                int64_t blocks = (int64_t)nsm * per_sm;
                if (blocks > (ngroups + KNN_WARPS - 1) / KNN_WARPS) blocks = (ngroups + KNN_WARPS - 1) / KNN_WARPS;
                logbuf.alloc((size_t)blocks * KNN_WARPS * logcap * 32);
                // The log is written once and read once by the same warp, then overwritten by the warp's next query group:
                // pinned in L2 (persisting access window) its lines are rewritten in place and never travel to HBM.
                if (getenv("NBK_KNN_NO_L2PIN") == nullptr) {
                    int max_persist = 0, max_window = 0;
                    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, t.device);
                    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, t.device);
                    if (max_persist > 0 && max_window > 0) {
                        size_t bytes = logbuf.bytes();
                        size_t win = bytes < (size_t)max_window ? bytes : (size_t)max_window;
                        cudaDeviceGetLimit(&old_persist_limit, cudaLimitPersistingL2CacheSize);
                        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
                        cudaStreamAttrValue av;
                        memset(&av, 0, sizeof(av));
                        av.accessPolicyWindow.base_ptr = logbuf.p;
                        av.accessPolicyWindow.num_bytes = win;
                        av.accessPolicyWindow.hitRatio = win <= (size_t)max_persist ? 1.0f : (float)((double)max_persist / (double)win);
                        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                        window = cudaStreamSetAttribute(t.stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
                        if (!window) cudaGetLastError();
                    }
                }
                kern<<<(int)blocks, KNN_WARPS * 32, smem, t.stream>>>(p, want_doubles, counters.p + 1, logbuf.p, logcap);
                if (window) {
                    cudaStreamAttrValue av;
                    memset(&av, 0, sizeof(av));
                    av.accessPolicyWindow.num_bytes = 0;
                    cudaStreamSetAttribute(t.stream, cudaStreamAttributeAccessPolicyWindow, &av);
                }
            };
            if (t.store_bytes == 4) {
                if (t.nlo2) { if (mb == 8) go(knn_sl_kernel<float, 8, true>); else if (mb == 6) go(knn_sl_kernel<float, 6, true>); else go(knn_sl_kernel<float, 5, true>); }
                else if (mb == 8) go(knn_sl_kernel<float, 8, false>); else if (mb == 6) go(knn_sl_kernel<float, 6, false>); else go(knn_sl_kernel<float, 5, false>);
            } else {
                if (t.nlo2) { if (mb == 8) go(knn_sl_kernel<double, 8, true>); else if (mb == 6) go(knn_sl_kernel<double, 6, true>); else go(knn_sl_kernel<double, 5, true>); }
                else if (mb == 8) go(knn_sl_kernel<double, 8, false>); else if (mb == 6) go(knn_sl_kernel<double, 6, false>); else go(knn_sl_kernel<double, 5, false>);
            }
            NBK_CHECK(cudaGetLastError());
#ifdef NBK_STATS
            {
                unsigned long long h[8];
                NBK_CHECK(cudaStreamSynchronize(t.stream));
                NBK_CHECK(cudaMemcpyFromSymbol(h, g_stats, sizeof(h)));
                double g = (double)((rows + 31) / 32);
                fprintf(stderr, "[nbk stats] per warp: tiles %.1f cand %.1f rounds %.1f ; per lane: screened-in %.1f inserted %.1f\n",
                        h[0] / g, h[1] / g, h[2] / g, h[4] / (double)rows, h[3] / (double)rows);
                unsigned long long z[8] = {0};
                NBK_CHECK(cudaMemcpyToSymbol(g_stats, z, sizeof(z)));
            }
#endif
            t.last_launches += 2;
            int nflag = 0;
            NBK_CHECK(cudaMemcpyAsync(&nflag, counters.p, sizeof(int), cudaMemcpyDeviceToHost, t.stream));
            NBK_CHECK(cudaStreamSynchronize(t.stream));
            if (window) {
                // hand the set-aside lines back to the normal L2: the carve-out would otherwise stay in force for every later
                // kernel of the process (builds and FOF lose ~2/3 of the L2 they stream their node / rank arrays through)
                cudaCtxResetPersistingL2Cache();
                if (getenv("NBK_KNN_KEEP_L2_LIMIT") == nullptr) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, old_persist_limit);
            }
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
        size_t warp_bytes = mode == 1 ? sc_warp_bytes(a.k, want_doubles) : fast_warp_bytes(p.kcap);
        size_t smem = warp_bytes * KNN_WARPS;
        NBK_REQUIRE(smem <= 227 * 1024, NBK_ERR_ARG, "k too large for the shared-memory heaps");
        DevBuf<int> fcount(1);
        DevBuf<int32_t> flist(rows);
        NBK_CHECK(cudaMemsetAsync(fcount.p, 0, sizeof(int), t.stream));
        p.flag_count = fcount.p; p.flag_list = flist.p;
        int blocks = div_up((rows + 31) / 32, KNN_WARPS);
        if (mode == 1) {
            if (t.store_bytes == 4) {
                NBK_CHECK(cudaFuncSetAttribute(knn_sc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                knn_sc_kernel<float><<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p, want_doubles);
            } else {
                NBK_CHECK(cudaFuncSetAttribute(knn_sc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                knn_sc_kernel<double><<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p, want_doubles);
            }
        } else if (t.store_bytes == 4) {
            NBK_CHECK(cudaFuncSetAttribute(knn_fast_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            knn_fast_kernel<float><<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p);
        } else {
            NBK_CHECK(cudaFuncSetAttribute(knn_fast_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            knn_fast_kernel<double><<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p);
        }
        NBK_CHECK(cudaGetLastError());
#ifdef NBK_STATS
        {
            unsigned long long h[8];
            NBK_CHECK(cudaStreamSynchronize(t.stream));
            NBK_CHECK(cudaMemcpyFromSymbol(h, g_stats, sizeof(h)));
            double g = (double)((rows + 31) / 32);
            fprintf(stderr, "[nbk stats] per warp: tiles %.1f cand %.1f rounds %.1f ; per lane: accepted-bits %.1f sifts %.1f\n",
                    h[0] / g, h[1] / g, h[2] / g, h[4] / (double)rows, h[3] / (double)rows);
            unsigned long long z[8] = {0};
            NBK_CHECK(cudaMemcpyToSymbol(g_stats, z, sizeof(z)));
        }
#endif
        t.last_launches += 2;
        int nflag = 0;
        NBK_CHECK(cudaMemcpyAsync(&nflag, fcount.p, sizeof(int), cudaMemcpyDeviceToHost, t.stream));
        NBK_CHECK(cudaStreamSynchronize(t.stream));
        t.last_flagged = nflag;
        if (nflag > 0) {
            KnnParams pe = p;
            pe.kcap = a.k;
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
