// knn.cu -- k-nearest-neighbour search with the SPH density / velocity-density / smoothing-scale
// epilogues fused into the same kernel.
//
// Replaces (reference): LeafNode::FindNearestPos KDLeafNode.cxx:15-28,119-130; SplitNode::FindNearestPos
// KDSplitNode.cxx:15-41; FindNearestPosPeriodic :1119-1148; PriorityQueue.h:14-85; drivers
// KDFindNearest.cxx:247-334,462-554; KDTree::CalcDensity / CalcVelDensity KDCalcSmoothQuantities.cxx:203-389.
//
// Layout: one warp = 32 queries adjacent in tree order; lane l owns query l and a bounded max-heap of
// (fp64 d2, int32 index) in shared memory, slot-major / lane-minor ([slot][32]) so that lanes touching
// different slots never bank-conflict.  Candidates come from the shared traversal (traverse.cuh); the
// distance is the reference's fp64 expression on exactly widened coordinates, so the neighbour set and
// every d2 are bit-identical to the reference whatever the traversal order.
#include "traverse.cuh"
#include "tree.h"

namespace nbk {

constexpr int KNN_WARPS = 4;
constexpr double KNN_SENTINEL = 1e32;   // reference MAXVALUE (Precision.h:49)

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

template <class S>
struct KnnVisitor {
    const Vec4<S>* P;
    double* tile;       // [3][32]
    WarpHeap hp;
    double qx, qy, qz;
    double top;         // heap top (current k-th distance^2)
    float topf;         // top rounded up
    int self;           // tree index of the query particle, or -1 (coordinate form)
    bool target_form;   // skip self and d2==0 (KDLeafNode.cxx:15-28)
    bool on;
    unsigned lane;

    __device__ __forceinline__ bool need(float lb) const { return lb < topf; }
    __device__ __forceinline__ bool whole(const QueryBox&, const NodeLo&, const NodeHi&, bool) const { return false; }
    __device__ __forceinline__ void settop() { top = hp.h(0); topf = __double2float_ru(top); }

    __device__ __forceinline__ void leaf(int start, int cnt) {
        for (int base = 0; base < cnt; base += 32) {
            int m = min(32, cnt - base);
            __syncwarp();
            if ((int)lane < m) {
                Vec4<S> c = P[start + base + lane];
                tile[lane] = (double)c.x; tile[32 + lane] = (double)c.y; tile[64 + lane] = (double)c.z;
            }
            __syncwarp();
            unsigned acc = 0;
#pragma unroll 4
            for (int j = 0; j < m; j++) {
                double d2 = dist2_ref(qx, qy, qz, tile[j], tile[32 + j], tile[64 + j]);
                bool ok = on && d2 < top;
                if (target_form) ok = ok && (start + base + j != self) && d2 > 0.0;
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

// smoothing kernel interpolation, KDCalcSmoothQuantities.cxx:12-15
__device__ __forceinline__ double wsm(double r, int i, int size, double delta, const double* __restrict__ x) {
    if (i < size - 1) { double a = x[i], b = x[i + 1]; return (a + (b - a) * (r - delta * i) / delta); }
    return x[i];
}

struct KnnParams {
    const NodeLo* nlo; const NodeHi* nhi; int bucket;
    const void* P; const void* V; const double* mass; const int32_t* order;
    int64_t q0, q1; const double* xq; int mode;
    int k, kcap;
    int periodic, strict, tree_form;
    double period[3];
    int32_t* nn; double* d2out; int out_ids;
    double* rho; double* hsm; int veldens_k;
    const double* kern; int kernres;
};

template <class S>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_kernel(KnnParams prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const size_t warp_bytes = (size_t)prm.kcap * 32 * 12 + 96 * 8 + TRAV_STACK * 4;
    unsigned char* base = smem_raw + w * warp_bytes;
    WarpHeap hp;
    hp.H = reinterpret_cast<double*>(base);
    hp.I = reinterpret_cast<int*>(base + (size_t)prm.kcap * 32 * 8);
    double* tile = reinterpret_cast<double*>(base + (size_t)prm.kcap * 32 * 12);
    int* stack = reinterpret_cast<int*>(base + (size_t)prm.kcap * 32 * 12 + 96 * 8);
    hp.k = prm.kcap; hp.lane = lane;

    const Vec4<S>* P = reinterpret_cast<const Vec4<S>*>(prm.P);
    int64_t group = (int64_t)blockIdx.x * KNN_WARPS + w;
    int64_t qi = prm.q0 + group * 32 + lane;
    if (prm.q0 + group * 32 >= prm.q1) return;       // whole warp out of range
    const bool valid = qi < prm.q1;

    KnnVisitor<S> v;
    v.P = P; v.tile = tile; v.hp = hp; v.lane = lane;
    v.on = valid;
    v.self = -1; v.target_form = false;
    double x0 = 0, y0 = 0, z0 = 0;
    if (valid) {
        if (prm.mode == 0) {
            Vec4<S> c = P[qi];
            x0 = (double)c.x; y0 = (double)c.y; z0 = (double)c.z;
            v.self = (int)qi;
            v.target_form = !prm.periodic;      // periodic particle searches use the coordinate form (KDSplitNode.cxx:1075-1080)
        } else {
            x0 = prm.xq[3 * qi]; y0 = prm.xq[3 * qi + 1]; z0 = prm.xq[3 * qi + 2];
        }
    }
    v.qx = x0; v.qy = y0; v.qz = z0;
    v.hp.init();
    v.settop();
    {
        QueryBox qb = make_qbox(x0, y0, z0);
        traverse(prm.nlo, prm.nhi, prm.bucket, stack, v, qb, valid);
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
    if (prm.hsm) prm.hsm[qi] = 0.5 * sqrt(v.top);
    if (prm.rho && prm.veldens_k == 0) {
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
    if (prm.rho && prm.veldens_k > 0) {
        // R2: CalcVelDensity (KDCalcSmoothQuantities.cxx:335-383)
        const Vec4<S>* V = reinterpret_cast<const Vec4<S>*>(prm.V);
        Vec4<S> vi = V[qi];
        int kx = 0;
        for (int s = 0; s < kc; s++) {
            int id = v.hp.i(s);
            if (id < 0) continue;
            Vec4<S> vj = V[id];
            double vd = sqrt(dist2_ref((double)vi.x, (double)vi.y, (double)vi.z, (double)vj.x, (double)vj.y, (double)vj.z));
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
        prm.rho[qi] = rho;
    }
    if (prm.nn || prm.d2out) {
        v.hp.sort_ascending(kc);
        // periodic particle searches carry k+1 slots: FindNearestPos(tt) drops the farthest, FindNearest(tt) the nearest (Q3)
        const int off = (prm.kcap > prm.k && prm.tree_form) ? 1 : 0;
        const int64_t row = (qi - prm.q0) * (int64_t)prm.k;
        for (int j = 0; j < prm.k; j++) {
            int id = v.hp.i(j + off);
            if (prm.nn) prm.nn[row + j] = (prm.out_ids && id >= 0) ? prm.order[id] : id;
            if (prm.d2out) prm.d2out[row + j] = v.hp.h(j + off);
        }
    }
}

void launch_knn(nbk_tree& t, const KnnArgs& a) {
    NBK_REQUIRE(a.k >= 1, NBK_ERR_ARG, "k must be >= 1");
    KnnParams p;
    p.nlo = t.nlo; p.nhi = t.nhi; p.bucket = t.bucket;
    p.P = t.prim; p.V = t.vel4(); p.mass = t.mass; p.order = t.order;
    p.q0 = a.q0; p.q1 = a.q1; p.xq = a.xq; p.mode = a.mode;
    p.k = a.k;
    p.kcap = a.k + ((a.periodic && a.mode == 0) ? 1 : 0);
    p.periodic = a.periodic; p.strict = a.strict; p.tree_form = a.tree_form;
    for (int d = 0; d < 3; d++) p.period[d] = t.period[d];
    p.nn = a.nn; p.d2out = a.d2; p.out_ids = a.out_ids;
    p.rho = a.rho; p.hsm = a.hsm; p.veldens_k = a.veldens_k;
    p.kern = t.d_kernel; p.kernres = t.kernres;
    if (a.veldens_k > 0) NBK_REQUIRE(p.V != nullptr, NBK_ERR_ARG, "velocity density needs velocities");
    int64_t rows = a.q1 - a.q0;
    if (rows <= 0) return;
    size_t warp_bytes = (size_t)p.kcap * 32 * 12 + 96 * 8 + TRAV_STACK * 4;
    size_t smem = warp_bytes * KNN_WARPS;
    NBK_REQUIRE(smem <= 227 * 1024, NBK_ERR_ARG, "k too large for the shared-memory heaps (max ~145 at 4 warps/CTA)");
    int64_t groups = (rows + 31) / 32;
    int blocks = div_up(groups, KNN_WARPS);
    if (t.store_bytes == 4) {
        NBK_CHECK(cudaFuncSetAttribute(knn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_kernel<float><<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p);
    } else {
        NBK_CHECK(cudaFuncSetAttribute(knn_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_kernel<double><<<blocks, KNN_WARPS * 32, smem, t.stream>>>(p);
    }
    NBK_CHECK(cudaGetLastError());
}

}  // namespace nbk
