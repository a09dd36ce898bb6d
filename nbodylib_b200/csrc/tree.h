// tree.h -- the device-resident tree object behind nbk_tree, and the launch entry points each .cu exports.
#pragma once
#include <functional>
#include <mutex>

#include "common.cuh"

struct nbk_tree {
    // entry points that work on a tree hold this lock: the tree's stream, events and last_* fields are per tree, so two host
    // threads calling into ONE tree are serialised here (different trees run concurrently)
    mutable std::mutex mtx;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;

    int64_t n = 0;
    int bucket = 16, treetype = 0, kerntype = 2, kernres = 1000, nd = 3;
    bool periodic = false;
    double period[3] = {0, 0, 0};
    int store_bytes = 4;            // 4: Vec4<float>, 8: Vec4<double>
    int64_t inexact = 0;
    int knn_fp32_ok = -1;           // density kernel: squared distances inside the root box fit fp32 (-1: not decided yet)

    // tree-order particle arrays (HBM).  prim = coordinates the tree is built on (pos; vel for TVEL),
    // sec = the other phase-space half (may be null).
    void* prim = nullptr;
    void* sec = nullptr;
    double* mass = nullptr;
    int32_t* order = nullptr;       // ID at tree index

    // nodes, heap order
    nbk::NodeLo* nlo = nullptr;
    nbk::NodeHi* nhi = nullptr;
    int8_t* cutdim = nullptr;
    int64_t nslots = 0;
    int depth = 0;                  // deepest level holding nodes
    int aligned2 = 0;               // split rule of the attached halo tree
    int aligned = 0;                // 1: warp-aligned shape (NBK_WARP_ALIGNED, split_left in common.cuh), 0: the reference's
    int64_t num_nodes = 0, num_leaves = 0;

    // optional second tree over "halo" particles (nbk_attach_halo): its particles are appended to the arrays above at
    // positions [n_main, n), its nodes live in nlo2/nhi2 with ranges already offset by n_main.  Queries of the Calc* family
    // then run for the main particles only and traverse both trees with the same per-query state.
    int64_t n_main = 0;             // 0: no halo attached
    nbk::NodeLo* nlo2 = nullptr;
    nbk::NodeHi* nhi2 = nullptr;

    // smoothing kernel table
    double* d_kernel = nullptr;
    std::vector<double> h_kernel;
    double kernnorm = 0;

    // timings
    double build_ms = 0, h2d_ms = 0, last_kernel_ms = 0, last_call_ms = 0;
    int64_t last_launches = 0;
    int64_t last_flagged = 0;       // queries the fp32-key kNN kernel handed to the exact kernel
    int64_t device_bytes = 0;

    // nbk_create: called by the build right before the particle gather; joins the side thread that stages the secondary
    // phase-space half and the masses (their host->device copies overlap the sorts and the level loop)
    std::function<void()> before_gather;

    nbk_tree() {}
    nbk_tree(const nbk_tree&) = delete;
    nbk_tree& operator=(const nbk_tree&) = delete;
    // frees everything the tree owns (also on the error paths of nbk_create, where the tree is held by a unique_ptr)
    ~nbk_tree() {
        int prev = -1;
        cudaGetDevice(&prev);
        if (prev != device) cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        void* bufs[] = {prim, sec, mass, order, nlo, nhi, cutdim, d_kernel, nlo2, nhi2};
        for (void* b : bufs) if (b) cudaFreeAsync(b, stream);
        if (stream) cudaStreamSynchronize(stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev2) cudaEventDestroy(ev2);
        if (ev3) cudaEventDestroy(ev3);
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
        if (prev >= 0 && prev != device) cudaSetDevice(prev);
    }

    const void* pos4() const { return treetype == NBK_TVEL ? sec : prim; }
    const void* vel4() const { return treetype == NBK_TVEL ? prim : sec; }
};

namespace nbk {

// build.cu : consumes input-order arrays (device), fills t.prim/sec/mass/order/nodes.
template <class S>
void build_tree(nbk_tree& t, const Vec4<S>* prim_in, const Vec4<S>* sec_in, const double* mass_in);

// knn.cu
struct KnnArgs {
    int k = 0;                 // neighbours wanted
    int mode = 0;              // 0 particle targets (tree index range), 1 arbitrary points
    int64_t q0 = 0, q1 = 0;    // particle range (mode 0) or [0,m) (mode 1)
    const double* xq = nullptr;  // device, m x 3 (mode 1)
    bool periodic = false;     // run the reference's periodic image schedule
    bool strict = false;       // NBK_STRICT_PERIODIC
    bool tree_form = false;    // NBK_KNN_TREE_FORM
    // outputs (device); any may be null
    int32_t* nn = nullptr;     // rows x k
    double* d2 = nullptr;      // rows x k
    bool out_ids = false;
    double* rho = nullptr;     // density accumulators, tree order (atomic adds), pre-zeroed
    double* hsm = nullptr;     // smoothing scale, tree order
    const uint8_t* active = nullptr;  // device, tree order, optional: only these particles are queries
    int veldens_k = 0;         // >0: CalcVelDensity with Nsmooth=veldens_k, Nsearch=k; rho gets the value (no atomics)
    // *Particle / *Position forms (KDCalcSmoothQuantities.cxx:768-921, 1092-1207): gather-only sums, one value per query ROW
    bool gather = false;       // rho (and hsm) are indexed by row, weight 1.0 * W, no scatter
    const int32_t* qlist = nullptr;  // device, explicit tree indices (mode 0) instead of the range [q0,q1)
    int64_t nq = 0;
    const double* vq = nullptr;      // device, m x 3 query velocities (mode 1 velocity density / FOF6d filter)
    // FindNearestCheck / FindNearestCriterion (KDFindNearest.cxx:363-441): candidate filters of the exact kernel
    const int32_t* cand_excl = nullptr;  // device, tree order: non-zero => never a neighbour
    int crit_mode = 0;                   // 0 none, 2 FOF3d, 4 FOF6d (predicate codes of FofArgs::mode)
    double cp0 = 0, cp1 = 0;
    // CalcSmoothVel / CalcSmoothVelDisp (KDCalcSmoothQuantities.cxx:480-614); device, tree order, accumulators pre-zeroed
    const double* rho_in = nullptr;      // densities (n)
    const double* smvel_in = nullptr;    // smoothed mean velocities (n x 3), input of the dispersion
    double* smvel_out = nullptr;         // n x 3
    double* smdisp_out = nullptr;        // n x 9 (row-major 3x3)
    const double* smdisp_in = nullptr;   // dispersions (n x 9), input of the skewness / kurtosis
    double* smhigh_out = nullptr;        // n x 3: skewness (moment 3) or kurtosis (moment 4)
    int moment = 0;
    bool phase = false;                  // FindNearestPhase: 6D distance keys (needs velocities; mode 1: vq)
};
void launch_knn(nbk_tree& t, const KnnArgs& a);
bool set_knn_option(const char* name, int64_t value);   // nbk_set_option names starting with "knn_"

// fof.cu
struct FofArgs {
    int mode = 0;              // 0: 3D ball  1: 6D ball (TPHS)  2: FOF3d  3: FOFVel  4: FOF6d  (same codes as oracle)
    double p0 = 0, p1 = 0;     // mode 0/1: p0 = fdist^2; mode 2/3: p0 = params[6]; mode 4: p0 = params[6], p1 = params[7]
    double prune_x2 = 0;       // spatial pruning radius^2 (>= any linked pair's position distance^2)
    int minnum = 8, order = 0;
    const int32_t* precheck_tree = nullptr;  // device, tree order, may be null
    bool attach = false;                     // FOFCriterionSetBasisForLinks: precheck != 0 particles cannot link but can be linked
    int32_t* group_tree = nullptr;           // device out, tree order
    int32_t* roots_tree = nullptr;           // device out, tree order: when set, only each particle's root (tree index) is produced
    int64_t ngroups = 0;
    int32_t *head = nullptr, *next = nullptr, *tail = nullptr, *len = nullptr;  // device, optional
};
void launch_fof(nbk_tree& t, FofArgs& a);
void launch_union_pairs(cudaStream_t st, int64_t nnodes, int64_t npairs, const int32_t* a, const int32_t* b, int32_t* root);
bool set_fof_option(const char* name, int64_t value);   // nbk_set_option names starting with "fof_"

// ball.cu
struct BallArgs {
    double r2 = 0;
    int64_t m = 0;
    const int32_t* qidx = nullptr;  // device, particle form (tree indices) or null
    const double* xq = nullptr;     // device, point form
    int64_t* offsets = nullptr;     // device m+1
    int32_t* idx = nullptr;         // device cap
    double* d2 = nullptr;           // device cap, optional: position distance^2 of every entry (dense forms)
    int64_t cap = 0, total = 0;
    bool out_ids = false;
    // criterion search (SearchCriterion*): predicate codes of FofArgs::mode (2: FOF3d, 4: FOF6d); mode 0 = ball of radius^2 r2
    int mode = 0;
    double p0 = 0, p1 = 0, prune_x2 = 0;
    const double* vq = nullptr;     // device, point form with velocities (SearchCriterionTagged(Particle&))
};
void launch_ball(nbk_tree& t, BallArgs& a);

// api.cu: what nbk_last_error() reports for the calling thread (libnbk_sharded.so reports through the same call)
void set_last_error(const char* msg);

}  // namespace nbk
