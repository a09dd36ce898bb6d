/*! \file KDTree.h (nbodylib_b200 shim)
 *  Header-only NBody::KDTree with the reference's public interface (reference src/KDTree/KDTree.h:81-657) whose
 *  methods marshal into the C ABI of include/nbk.h (libnbk.so, CUDA sm_100a).  A program written against the
 *  reference compiles against this header unchanged for the calls on the hot path:
 *
 *    ctor (Particle*, numparts, bucket_size, TreeType, KernType, KernRes, SplittingCriterion, Aniso, ScaleSpace, Period)
 *    GetNumNodes / GetNumLeafNodes / GetBucketSize / GetTreeType / GetKernType / GetKernNorm / GetPeriod
 *    FindNearest / FindNearestPos (Int_t tt | Double_t* x | Coordinate | whole system)
 *    FindNearestPhase (Int_t tt | Double_t* x, v | Coordinate x, v | whole system); FindNearest on a TPHS tree with Aniso = -1
 *    FindNearestVel (Int_t tt | Double_t* v | Coordinate v | whole system) on TVEL trees
 *    SearchBallPosTagged (Int_t tt | Double_t* x | Coordinate; array and vector forms), dense SearchBall / SearchBallPos
 *    SearchCriterionTagged (Int_t tt | Particle&; array and vector forms), dense SearchCriterion (FOF3d / FOF6d)
 *    CalcDensity, CalcVelDensity, CalcSmoothingScale (new: north star), CalcDensityParticle, CalcVelDensityParticle,
 *    CalcDensityPosition, CalcVelDensityPosition, CalcSmoothLocalValue, CalcSmoothVel, CalcSmoothVelDisp, CalcSmoothVelSkew,
 *    CalcSmoothVelKurtosis
 *    FindNearestCheck, FindNearestCriterion (Int_t tt | Particle | Coordinate)
 *    FOF, FOFCriterion, FOFCriterionSetBasisForLinks, FOFCriterionParticle (FOF3d / FOF6d), GetRoot / FindLeafNode (host mirror of the node arrays)
 *    OverWriteInputOrder, SetResetOrder, ~KDTree (restores the caller's particle order)
 *
 *  Semantics kept from the reference: the caller's Particle array is permuted IN PLACE into tree order and
 *  Particle::id is overwritten with the input index (KDTree.cxx:1291); nn[] / tt / tagged[] are tree-order indices;
 *  FOF results are new[]-allocated arrays indexed by ID that the caller delete[]s; the destructor sorts the array
 *  back by ID unless OverWriteInputOrder() was called (KDTree.cxx:1340-1362).
 *  Differences: errors throw std::runtime_error instead of printf+exit; calls without a device implementation
 *  (TPROJ/TMETRIC trees, host FOFcompfunc callbacks other than FOF3d/FOF6d, the metric searches of TPHS trees with Aniso >= 0, FOFNN*) are absent or throw -- there is
 *  no CPU fallback.  Per-particle calls launch one small kernel each; loops over all particles should use the
 *  whole-system forms.
 */
#ifndef NBK_SHIM_KDTREE_H
#define NBK_SHIM_KDTREE_H

#ifdef NBK_USE_REFERENCE_PARTICLE
// the consumer has the NBodylib headers (src/NBody, src/Math and src/KDTree on the include path AFTER this directory):
// Particle / System / Coordinate / Matrix, the priority queue and the FOF criteria are the reference's own
#include <NBody.h>
#include <NBodyMath.h>
#include <PriorityQueue.h>
#include <FOFFunc.h>
namespace NBody {
// the reference defines these next to its node classes (KDNode.h:29-35), which this header replaces
#ifdef LARGETREE
typedef long int Int_tree_t;
typedef unsigned long int UInt_tree_t;
#else
typedef int Int_tree_t;
typedef unsigned int UInt_tree_t;
#endif
}
#else
#include "nbk_standalone_types.h"
#endif
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <new>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nbk.h"
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <type_traits>

// the marshalling loops over the caller's particle array run on all host threads when the consumer is built with OpenMP
#if defined(_OPENMP)
#define NBK_SHIM_PARALLEL_FOR _Pragma("omp parallel for schedule(static)")
#else
#define NBK_SHIM_PARALLEL_FOR
#endif

namespace NBody {

/// uninitialised per-particle scratch of the shim (a std::vector would zero a GB-sized array on one thread before the parallel
/// loop that fills it touches it)
template <class T>
struct NbkRawBuf {
    T* p;
    explicit NbkRawBuf(size_t n) : p(static_cast<T*>(::operator new(sizeof(T) * (n ? n : 1)))) {}
    ~NbkRawBuf() { ::operator delete(p); }
    NbkRawBuf(const NbkRawBuf&) = delete;
    NbkRawBuf& operator=(const NbkRawBuf&) = delete;
    T* data() { return p; }
    T& operator[](size_t i) { return p[i]; }
};

/// Page-aligned scratch for the permutation of the caller's particle array, with transparent huge pages requested where the
/// platform has them: at 512^3 the scratch is 11.8 GB, i.e. 2.9 M first-touch faults with 4 KiB pages.
struct NbkPermScratch {
    void* p = nullptr;
    size_t bytes = 0;
    bool mapped = false;
    explicit NbkPermScratch(size_t n) {
#if defined(__linux__)
        const size_t huge = (size_t)2 << 20;
        bytes = (n + huge - 1) / huge * huge;
        void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m != MAP_FAILED) {
#ifdef MADV_HUGEPAGE
            (void)madvise(m, bytes, MADV_HUGEPAGE);
#endif
            p = m; mapped = true;
            return;
        }
#endif
        p = ::operator new(n ? n : 1);
    }
    ~NbkPermScratch() {
#if defined(__linux__)
        if (mapped) { munmap(p, bytes); return; }
#endif
        ::operator delete(p);
    }
    NbkPermScratch(const NbkPermScratch&) = delete;
    NbkPermScratch& operator=(const NbkPermScratch&) = delete;
};

/// a[i] <- a[src(i)] for a permutation src of 0..n-1, on all host threads (what the reference does with in-place quickselect
/// swaps while it builds, KDTree.cxx:328-370, and with std::sort by ID in its destructor, :1347).  Two streaming passes
/// through a scratch copy: a gather (random reads of whole records, software prefetched) and a block copy back.  Trivially
/// copyable records move as bytes; any other Particle type goes through its move operations.  The reference's own Particle
/// has a user-written copy constructor, so the compiler cannot call it trivially copyable even in the default build, where
/// every member is a scalar: a consumer who knows its Particle build can be relocated byte-wise (no member pointing into the
/// object itself: true for every NBodylib configuration, whose optional members are unique_ptrs) defines
/// NBK_SHIM_RELOCATE_BYTES to take the byte path -- each record is moved exactly twice and never copied or destroyed.
template <class P, class F>
inline void nbk_permute_impl(P* a, int64_t n, F src, void* scratch, std::true_type /* trivially copyable */) {
    unsigned char* tmp = static_cast<unsigned char*>(scratch);
    const int64_t ahead = 16;
    NBK_SHIM_PARALLEL_FOR
    for (int64_t i = 0; i < n; i++) {
        if (i + ahead < n) {
            const char* nx = reinterpret_cast<const char*>(a + src(i + ahead));
            __builtin_prefetch(nx); __builtin_prefetch(nx + 64);
        }
        std::memcpy(tmp + sizeof(P) * (size_t)i, static_cast<const void*>(a + src(i)), sizeof(P));
    }
    const int64_t blk = 1 << 16;          // records per block of the copy back
    NBK_SHIM_PARALLEL_FOR
    for (int64_t b = 0; b < (n + blk - 1) / blk; b++) {
        const int64_t i0 = b * blk, i1 = std::min(n, i0 + blk);
        std::memcpy(static_cast<void*>(a + i0), tmp + sizeof(P) * (size_t)i0, sizeof(P) * (size_t)(i1 - i0));
    }
}
template <class P, class F>
inline void nbk_permute_impl(P* a, int64_t n, F src, void* scratch, std::false_type) {
    P* tmp = static_cast<P*>(scratch);
    NBK_SHIM_PARALLEL_FOR
    for (int64_t i = 0; i < n; i++) new (tmp + i) P(std::move(a[src(i)]));
    NBK_SHIM_PARALLEL_FOR
    for (int64_t i = 0; i < n; i++) { a[i] = std::move(tmp[i]); tmp[i].~P(); }
}
template <class P, class F>
inline void nbk_permute_records(P* a, int64_t n, F src) {
    if (n <= 0) return;
    NbkPermScratch scratch(sizeof(P) * (size_t)n);
#ifdef NBK_SHIM_RELOCATE_BYTES
    nbk_permute_impl(a, n, src, scratch.p, std::true_type());
#else
    nbk_permute_impl(a, n, src, scratch.p, typename std::is_trivially_copyable<P>::type());
#endif
}

/// Host mirror of one tree node (reference KDNode.h:45-334 Node, :343-484 SplitNode, :492-610 LeafNode): what callers of
/// KDTree::GetRoot() / FindLeafNode() read.  IDs number the nodes depth first, left before right, like the reference's
/// BuildNodes.  Boundaries are the node's particle bounding box (fp32, rounded outward when the tree stores fp64
/// coordinates); the cut value of a split node is the largest cut-dimension coordinate of its left child, i.e. the median
/// particle's coordinate (KDTree.cxx:1012-1013).
class Node {
    friend class KDTree;
protected:
    Int_t nid = 0, bucket_start = 0, bucket_end = 0, numdim = 3;
    Double_t xbnd[6][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
    int cut_dim = -1;
    Double_t cut_val = 0;
    Node *left = nullptr, *right = nullptr;
public:
    virtual ~Node() {}
    Int_t GetID() const { return nid; }
    Int_t GetCount() const { return bucket_end - bucket_start; }
    Int_t GetStart() const { return bucket_start; }
    Int_t GetEnd() const { return bucket_end; }
    Double_t GetBoundary(int i, int j) const { return xbnd[i][j]; }
    bool GetLeaf() const { return left == nullptr; }
};
class SplitNode : public Node {
public:
    int GetCutDim() const { return cut_dim; }
    Double_t GetCutValue() const { return cut_val; }
    Node* GetLeft() const { return left; }
    Node* GetRight() const { return right; }
};
class LeafNode : public Node {};

class KDTree {
public:
    const static int TPHYS = 0, TPROJ = 1, TVEL = 2, TPHS = 3, TMETRIC = 4;
    const static int KSPH = 0, KGAUSS = 1, KEPAN = 2, KTH = 3;
    const static int KDTREE_SPLIT_ENTROPY = 1, KDTREE_SPLIT_DISPERSION = 2, KDTREE_SPLIT_MAXINTERPARTICLESPACING = 3, KDTREE_SPLIT_SPREAD = 0;

private:
    nbk_tree* h = nullptr;
    Particle* bucket = nullptr;
    Int_t numparts = 0;
    int anisotropic = 0;      // the constructor's Aniso: -1 = plain phase-space search on a TPHS tree (KDTree.h:157-158)
    Int_t bucket_size_caller = 16;
    bool iresetorder = true;
    nbk_info info{};
    Double_t* period = nullptr;

    std::mutex dev_mutex;          // device calls on one tree are serialised (callers may sit inside an OpenMP loop)
    unsigned long long serial = 0;  // distinguishes trees that reuse an address (per-thread caches key on it)

    static void check(int rc) {
        if (rc != NBK_OK) throw std::runtime_error(std::string("nbk: ") + nbk_last_error());
    }
    static unsigned long long next_serial() { static std::mutex m; static unsigned long long c = 0; std::lock_guard<std::mutex> g(m); return ++c; }

    // Per-thread block cache behind the per-particle FindNearest*(Int_t tt) calls.  The reference's callers loop over
    // every particle (often inside `omp parallel for`, e.g. reference tests/test_kdtree.cxx:279-301) and call the
    // per-particle form; one kernel launch per call would waste the device.  The first call of a thread computes the
    // neighbours of the whole block of consecutive tree indices around tt in ONE batched device query; the following
    // calls of that thread (static / dynamic / guided chunks are runs of consecutive indices) are served from host memory.
    struct KnnBlock {
        unsigned long long serial = 0; Int_t b0 = 0, b1 = 0, k = 0; int flags = -1;
        int crit = -2; FOFcheckfunc checkfn = nullptr; double params[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // filtered searches
        std::vector<int32_t> nn; std::vector<double> d2;
    };
    static KnnBlock& tls_block() { static thread_local KnnBlock b; return b; }
    Int_t block_size(Int_t k) const {
        Int_t b = numparts / 64;
        if (b < 4096) b = 4096;
        if (b > 65536) b = 65536;
        while ((size_t)b * (size_t)k * 12 > ((size_t)64 << 20) && b > 1024) b >>= 1;     // <= 64 MiB of host cache per thread
        return b;
    }
    void refresh() { check(nbk_get_info(h, &info)); }

    // host mirror of the node arrays, built on the first GetRoot() / FindLeafNode() call
    std::vector<std::unique_ptr<Node>> nodes;
    Node* root = nullptr;
    std::vector<Double_t> kernel;   // KernelConstruction table (nbk_get_kernel_table), fetched on first use
    Node* mirror(int64_t slot, const std::vector<int32_t>& st, const std::vector<int32_t>& en, const std::vector<int32_t>& cd,
                 const std::vector<float>& bd, Int_t& next_id) {
        Node* nd;
        const bool leaf = cd[slot] < 0;
        if (leaf) nd = new LeafNode(); else nd = new SplitNode();
        nodes.emplace_back(nd);
        nd->nid = next_id++;
        nd->bucket_start = st[slot]; nd->bucket_end = en[slot];
        for (int j = 0; j < 3; j++) { nd->xbnd[j][0] = bd[6 * slot + 2 * j]; nd->xbnd[j][1] = bd[6 * slot + 2 * j + 1]; }
        if (!leaf) {
            nd->cut_dim = cd[slot];
            nd->left = mirror(2 * slot + 1, st, en, cd, bd, next_id);
            nd->right = mirror(2 * slot + 2, st, en, cd, bd, next_id);
            nd->cut_val = nd->left->xbnd[nd->cut_dim][1];
        }
        return nd;
    }
    void build_mirror() {
        std::lock_guard<std::mutex> g(dev_mutex);
        if (root) return;
        int64_t ns = 0;
        check(nbk_get_nodes(h, &ns, NULL, NULL, NULL, NULL));
        std::vector<int32_t> st(ns), en(ns), cd(ns);
        std::vector<float> bd((size_t)6 * ns);
        check(nbk_get_nodes(h, &ns, st.data(), en.data(), cd.data(), bd.data()));
        Int_t next_id = 0;
        nodes.reserve((size_t)info.num_nodes);
        root = mirror(0, st, en, cd, bd, next_id);
    }
    const std::vector<Double_t>& kernel_table() {
        std::lock_guard<std::mutex> g(dev_mutex);
        if (kernel.empty()) {
            std::vector<double> kt(info.kernres);
            check(nbk_get_kernel_table(h, kt.data()));
            kernel.assign(kt.begin(), kt.end());
        }
        return kernel;
    }
    int crit_code(FOFcompfunc cmp, const char* who) {
        // inline criteria are recognised by address inside the caller's translation unit
        if (cmp == (FOFcompfunc)&FOF3d) return NBK_FOF3D;
        if (cmp == (FOFcompfunc)&FOF6d) return NBK_FOF6D;
        throw std::runtime_error(std::string("nbk shim: ") + who + " supports FOF3d and FOF6d; host callbacks cannot run on the device");
    }

public:
    KDTree(Particle* p, Int_t nparts, Int_t bucket_size = 16, int TreeType = TPHYS, int KernType = KEPAN, int KernRes = 1000,
           int SplittingCriterion = KDTREE_SPLIT_SPREAD, int Aniso = 0, int ScaleSpace = 0, Double_t* Period = NULL,
           Double_t** metric = NULL, bool iBuildInParallel = true, bool iKeepInputOrder = false, Double_t Rdistadapt = -1,
           Double_t AdaptiveMedianFac = 0.0, Int_t min_bucket_size = 16)
        : bucket(p), numparts(nparts), anisotropic(Aniso), bucket_size_caller(bucket_size) {
        (void)metric; (void)iBuildInParallel;
        // ScaleSpace: the reference accumulates into xmean[] without ever initialising it and starts the variance sums from 1.0
        // (KDTree.h:140, KDTree.cxx:1063-1083,1293): its result is not defined, so there is nothing to be faithful to
        if (ScaleSpace) throw std::runtime_error("nbk shim: ScaleSpace has no device implementation");
        if (iKeepInputOrder) throw std::runtime_error("nbk shim: keep-order builds have no device implementation");
        // Rdistadapt / AdaptiveMedianFac / SplittingCriterion shape the REFERENCE's tree (where leaves stop: size <= b and radius <
        // Rdistadapt, or size <= min_bucket_size, KDTree.cxx:988-991; where the cut goes, :785-960; which dimension is cut,
        // :459-504).  Every search, density and FOF result is a function of the particle set alone, so they are served from the
        // device tree, whose shape follows the median / largest-spread rule with leaves of up to min_bucket_size particles in the
        // adaptive case; GetNumNodes / GetNumLeafNodes / GetRoot describe the device tree.
        if (Rdistadapt > 0) bucket_size = std::max<Int_t>(1, std::min(bucket_size, min_bucket_size));
        (void)AdaptiveMedianFac;
        if (SplittingCriterion < KDTREE_SPLIT_SPREAD || SplittingCriterion > KDTREE_SPLIT_MAXINTERPARTICLESPACING) throw std::runtime_error("nbk shim: unknown splitting criterion");
        SplittingCriterion = KDTREE_SPLIT_SPREAD;
        NbkRawBuf<Double_t> mass((size_t)numparts);
        NBK_SHIM_PARALLEL_FOR
        for (Int_t i = 0; i < numparts; i++) { bucket[i].SetID(i); mass[i] = bucket[i].GetMass(); }      // KDTree.cxx:1291
        nbk_particles np;
        np.pos = bucket[0].GetPosition(); np.pos_stride = (int64_t)sizeof(Particle);
        np.vel = bucket[0].GetVelocity(); np.vel_stride = (int64_t)sizeof(Particle);
        np.mass = mass.data(); np.mass_stride = (int64_t)sizeof(Double_t);
        np.real_bytes = (int32_t)sizeof(Double_t); np.on_device = 0;
        if (Period != NULL) { period = new Double_t[3]; for (int k = 0; k < 3; k++) period[k] = Period[k]; }
        double per[3] = {0, 0, 0};
        if (period) for (int k = 0; k < 3; k++) per[k] = (double)period[k];
        check(nbk_create(&np, numparts, (int)bucket_size, TreeType, KernType, KernRes, SplittingCriterion, period ? per : NULL, 0, -1, &h));
        serial = next_serial();
        refresh();
        // bring the caller's array into tree order (the reference does this with in-place quickselect swaps)
        NbkRawBuf<int32_t> order((size_t)numparts);
        check(nbk_get_order(h, order.data(), 0));
        permute([&](Int_t i) { return (Int_t)order[i]; });
    }
    KDTree(System& s, Int_t bucket_size = 16, int TreeType = TPHYS, int KernType = KEPAN, int KernRes = 1000, int SplittingCriterion = 0,
           int Aniso = 0, int ScaleSpace = 0)
        : KDTree(s.Parts(), s.GetNumParts(), bucket_size, TreeType, KernType, KernRes, SplittingCriterion, Aniso, ScaleSpace,
                 (s.GetPeriod()[0] == 0 && s.GetPeriod()[1] == 0 && s.GetPeriod()[2] == 0) ? (Double_t*)NULL : s.GetPeriod().GetCoord()) {}
    KDTree(const KDTree&) = delete;
    KDTree& operator=(const KDTree&) = delete;

    ~KDTree() {
        if (h) nbk_destroy(h);
        if (period) delete[] period;
        if (iresetorder && bucket) {
            // reference: std::sort(bucket, bucket+numparts, IDCompareVec); ids are a permutation of 0..N-1 -> O(N) placement
            NbkRawBuf<Int_t> src((size_t)numparts);
            NBK_SHIM_PARALLEL_FOR
            for (Int_t i = 0; i < numparts; i++) src[bucket[i].GetID()] = i;
            permute([&](Int_t i) { return src[i]; });
        }
    }

    Int_t GetNumNodes() { return info.num_nodes; }
    Int_t GetNumLeafNodes() { return info.num_leaves; }
    Int_t GetBucketSize() { return bucket_size_caller; }
    Int_t GetTreeType() { return info.treetype; }
    Int_t GetKernType() { return info.kerntype; }
    Double_t GetKernNorm() { return info.kernnorm; }
    Double_t GetPeriod(int j) { return period[j]; }
    nbk_tree* GetHandle() { return h; }

    // ---- nearest neighbours (KDFindNearest.cxx:247-334, 444-554) ----------------------------------------------
    void FindNearestPos(Int_t tt, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { knn_cached(tt, nn, dist2, Nsearch, 0); }
    /// On a TPHS tree FindNearest is the phase-space search when the tree was built with Aniso = -1 (KDFindNearest.cxx:260-262,
    /// 300-301); with the constructor default Aniso = 0 the reference takes its metric path (quirk Q4), which is not built.
    void FindNearest(Int_t tt, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        if (phase_tree("FindNearest")) knn_cached(tt, nn, dist2, Nsearch, 0, -3);
        else knn_cached(tt, nn, dist2, Nsearch, NBK_KNN_TREE_FORM);
    }
    // ---- velocity-space nearest neighbours (KDFindNearest.cxx:335-346, 451-452, 530-540, 559-561) ---------------
    /// The reference's FindNearestVel walks whatever tree it is called on with velocity coordinates against the tree's cut
    /// planes, so it is only meaningful on a TVEL tree: k slots, target form, never reflected -- what nbk_knn_* do on TVEL trees.
    void FindNearestVel(Int_t tt, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { require_vel_tree("FindNearestVel"); knn_cached(tt, nn, dist2, Nsearch, 0); }
    void FindNearestVel(Double_t* v, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { require_vel_tree("FindNearestVel"); FindNearestPos(v, nn, dist2, Nsearch); }
    void FindNearestVel(Coordinate v, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { FindNearestVel(v.GetCoord(), nn, dist2, Nsearch); }
    void FindNearestVel(Int_t** nn, Double_t** dist2, Int_t Nsearch = 64) { require_vel_tree("FindNearestVel"); knn_all(nn, dist2, Nsearch, 0); }
    // ---- phase-space nearest neighbours (KDFindNearest.cxx:347-361, 543-555; PhaseDistSqd, DistFunc.h:41-49) ------
    void FindNearestPhase(Int_t tt, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { knn_cached(tt, nn, dist2, Nsearch, 0, -3); }
    void FindNearestPhase(Double_t* x, Double_t* v, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        std::vector<double> d2(Nsearch);
        std::vector<int32_t> n32(Nsearch);
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]}, vv[3] = {(double)v[0], (double)v[1], (double)v[2]};
        {
            std::lock_guard<std::mutex> g(dev_mutex);
            check(nbk_knn_phase_points(h, (int)Nsearch, 1, xx, vv, n32.data(), d2.data(), 0));
        }
        for (Int_t j = 0; j < Nsearch; j++) { nn[j] = n32[j]; dist2[j] = d2[j]; }
    }
    void FindNearestPhase(Coordinate x, Coordinate v, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        FindNearestPhase(x.GetCoord(), v.GetCoord(), nn, dist2, Nsearch);
    }
    void FindNearestPhase(Int_t** nn, Double_t** dist2, Int_t Nsearch = 64) { knn_all(nn, dist2, Nsearch, 0, true); }
    void FindNearestPos(Double_t* x, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        std::vector<double> d2(Nsearch);
        std::vector<int32_t> n32(Nsearch);
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]};
        {
            std::lock_guard<std::mutex> g(dev_mutex);
            check(nbk_knn_points(h, (int)Nsearch, 1, xx, n32.data(), d2.data(), 0));
        }
        for (Int_t j = 0; j < Nsearch; j++) { nn[j] = n32[j]; dist2[j] = d2[j]; }
    }
    /// on a TPHS tree x holds six numbers, position then velocity (KDFindNearest.cxx:474-477)
    void FindNearest(Double_t* x, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        if (phase_tree("FindNearest")) FindNearestPhase(x, x + 3, nn, dist2, Nsearch);
        else FindNearestPos(x, nn, dist2, Nsearch);
    }
    void FindNearestPos(Coordinate x, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { FindNearestPos(x.GetCoord(), nn, dist2, Nsearch); }
    /// whole-system forms: nn[i][j], dist2[i][j] for every tree index i
    void FindNearestPos(Int_t** nn, Double_t** dist2, Int_t Nsearch = 64) { knn_all(nn, dist2, Nsearch, 0); }
    void FindNearest(Int_t** nn, Double_t** dist2, Int_t Nsearch = 64) {
        if (phase_tree("FindNearest")) knn_all(nn, dist2, Nsearch, 0, true);
        else knn_all(nn, dist2, Nsearch, NBK_KNN_TREE_FORM);
    }

    // ---- filtered nearest neighbours (KDFindNearest.cxx:363-441) ------------------------------------------------
    /// Per-particle forms are served from the per-thread block cache like FindNearest(tt).  The caller's FOFcheckfunc runs on
    /// the host for every particle when a block is computed (its values are what the device receives); params are read when
    /// the block is computed and compared by value afterwards.
    void FindNearestCheck(Int_t tt, FOFcheckfunc check_, Double_t* params, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        knn_cached(tt, nn, dist2, Nsearch, period ? NBK_KNN_TREE_FORM : 0, -1, check_, params);
    }
    void FindNearestCriterion(Int_t tt, FOFcompfunc cmp, Double_t* params, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        knn_cached(tt, nn, dist2, Nsearch, period ? NBK_KNN_TREE_FORM : 0, crit_code(cmp, "FindNearestCriterion"), nullptr, params);
    }
    void FindNearestCheck(Coordinate x, FOFcheckfunc check_, Double_t* params, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        knn_filtered_point(x.GetCoord(), NULL, -1, check_, params, nn, dist2, Nsearch);
    }
    void FindNearestCheck(Particle p, FOFcheckfunc check_, Double_t* params, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        knn_filtered_point(p.GetPosition(), NULL, -1, check_, params, nn, dist2, Nsearch);
    }
    void FindNearestCriterion(Particle p, FOFcompfunc cmp, Double_t* params, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        knn_filtered_point(p.GetPosition(), p.GetVelocity(), crit_code(cmp, "FindNearestCriterion"), nullptr, params, nn, dist2, Nsearch);
    }

    // ---- fixed radius (KDFindNearest.cxx:567-688) ------------------------------------------------------------
    Int_t SearchBallPosTagged(Int_t tt, Double_t fdist2, Int_t* tagged) {
        std::vector<Int_t> v = SearchBallPosTagged(tt, fdist2);
        std::copy(v.begin(), v.end(), tagged);
        return (Int_t)v.size();
    }
    Int_t SearchBallPosTagged(Double_t* x, Double_t fdist2, Int_t* tagged) {
        std::vector<Int_t> v = SearchBallPosTagged(x, fdist2);
        std::copy(v.begin(), v.end(), tagged);
        return (Int_t)v.size();
    }
    Int_t SearchBallPosTagged(Coordinate x, Double_t fdist2, Int_t* tagged) { return SearchBallPosTagged(x.GetCoord(), fdist2, tagged); }
    std::vector<Int_t> SearchBallPosTagged(Int_t tt, Double_t fdist2) { return ball_cached(tt, (double)fdist2, -1, NULL); }
    std::vector<Int_t> SearchBallPosTagged(Double_t* x, Double_t fdist2) {
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]};
        std::vector<int32_t> idx;
        csr_row([&](int64_t* off, int32_t* ix, double* d2, int64_t cap, int64_t* tot, int fl) {
            return nbk_ball_points(h, (double)fdist2, 1, xx, off, ix, d2, cap, tot, fl); }, 0, idx, NULL);
        return std::vector<Int_t>(idx.begin(), idx.end());
    }
    std::vector<Int_t> SearchBallPosTagged(Coordinate x, Double_t fdist2) { return SearchBallPosTagged(x.GetCoord(), fdist2); }

    /// dense forms (KDFindNearest.cxx:567-587): nn[ID] = imark, dist2[ID] = d2 for every particle inside the ball; nn and
    /// dist2 are the caller's numparts-sized arrays, indexed by particle ID.  Non periodic target form: the target itself is
    /// not marked (the reference marks it only when a whole node containing it is swallowed, quirk Q5).
    void SearchBallPos(Int_t tt, Double_t fdist2, Int_t imark, Int_t* nn, Double_t* dist2) {
        int32_t q = (int32_t)tt;
        std::vector<int32_t> idx; std::vector<double> d2;
        csr_row([&](int64_t* off, int32_t* ix, double* dd, int64_t cap, int64_t* tot, int fl) {
            return nbk_ball_particles(h, (double)fdist2, 1, &q, off, ix, dd, cap, tot, fl); }, NBK_OUT_IDS, idx, &d2);
        for (size_t j = 0; j < idx.size(); j++) { nn[idx[j]] = imark; dist2[idx[j]] = d2[j]; }
    }
    void SearchBallPos(Double_t* x, Double_t fdist2, Int_t imark, Int_t* nn, Double_t* dist2) {
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]};
        std::vector<int32_t> idx; std::vector<double> d2;
        csr_row([&](int64_t* off, int32_t* ix, double* dd, int64_t cap, int64_t* tot, int fl) {
            return nbk_ball_points(h, (double)fdist2, 1, xx, off, ix, dd, cap, tot, fl); }, NBK_OUT_IDS, idx, &d2);
        for (size_t j = 0; j < idx.size(); j++) { nn[idx[j]] = imark; dist2[idx[j]] = d2[j]; }
    }
    void SearchBallPos(Coordinate x, Double_t fdist2, Int_t imark, Int_t* nn, Double_t* dist2) { SearchBallPos(x.GetCoord(), fdist2, imark, nn, dist2); }
    /// SearchBall dispatches on the tree type (KDFindNearest.cxx:557-565); only position trees have a device ball search
    void SearchBall(Int_t tt, Double_t fdist2, Int_t imark, Int_t* nn, Double_t* dist2) { require_pos_tree("SearchBall"); SearchBallPos(tt, fdist2, imark, nn, dist2); }
    void SearchBall(Double_t* x, Double_t fdist2, Int_t imark, Int_t* nn, Double_t* dist2) { require_pos_tree("SearchBall"); SearchBallPos(x, fdist2, imark, nn, dist2); }
    void SearchBall(Coordinate x, Double_t fdist2, Int_t imark, Int_t* nn, Double_t* dist2) { require_pos_tree("SearchBall"); SearchBallPos(x.GetCoord(), fdist2, imark, nn, dist2); }

    // ---- criterion search (KDFindNearest.cxx:590-603, 660-706; FOF3d / FOF6d) ----------------------------------
    std::vector<Int_t> SearchCriterionTagged(Int_t tt, FOFcompfunc cmp, Double_t* params) {
        return ball_cached(tt, 0.0, crit_code(cmp, "SearchCriterionTagged"), params);
    }
    std::vector<Int_t> SearchCriterionTagged(Particle& p, FOFcompfunc cmp, Double_t* params) {
        const int crit = crit_code(cmp, "SearchCriterionTagged");
        double pr[16]; for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        double xx[3] = {(double)p.GetPosition(0), (double)p.GetPosition(1), (double)p.GetPosition(2)};
        double vv[3] = {(double)p.GetVelocity(0), (double)p.GetVelocity(1), (double)p.GetVelocity(2)};
        std::vector<int32_t> idx;
        csr_row([&](int64_t* off, int32_t* ix, double* d2, int64_t cap, int64_t* tot, int fl) {
            return nbk_search_criterion_points(h, crit, pr, 1, xx, vv, off, ix, d2, cap, tot, fl); }, 0, idx, NULL);
        return std::vector<Int_t>(idx.begin(), idx.end());
    }
    Int_t SearchCriterionTagged(Int_t tt, FOFcompfunc cmp, Double_t* params, Int_t* tagged) {
        std::vector<Int_t> v = SearchCriterionTagged(tt, cmp, params);
        std::copy(v.begin(), v.end(), tagged);
        return (Int_t)v.size();
    }
    Int_t SearchCriterionTagged(Particle& p, FOFcompfunc cmp, Double_t* params, Int_t* tagged) {
        std::vector<Int_t> v = SearchCriterionTagged(p, cmp, params);
        std::copy(v.begin(), v.end(), tagged);
        return (Int_t)v.size();
    }
    /// dense forms: nn[ID] = imark (and dist2[ID] = position distance^2) for every particle meeting the criterion.  The
    /// reference additionally skips particles whose nn[ID] is non-zero and <= imark (KDLeafNode.cxx:418): same here.
    void SearchCriterion(Int_t tt, FOFcompfunc cmp, Double_t* params, Int_t imark, Int_t* nn, Double_t* dist2) {
        const int crit = crit_code(cmp, "SearchCriterion");
        double pr[16]; for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        int32_t q = (int32_t)tt;
        std::vector<int32_t> idx; std::vector<double> d2;
        csr_row([&](int64_t* off, int32_t* ix, double* dd, int64_t cap, int64_t* tot, int fl) {
            return nbk_search_criterion_particles(h, crit, pr, 1, &q, off, ix, dd, cap, tot, fl); }, NBK_OUT_IDS, idx, &d2);
        for (size_t j = 0; j < idx.size(); j++)
            if (nn[idx[j]] > imark || nn[idx[j]] == 0) { nn[idx[j]] = imark; if (dist2) dist2[idx[j]] = d2[j]; }
    }
    void SearchCriterion(Int_t tt, FOFcompfunc cmp, Double_t* params, Int_t imark, Int_t* nn) { SearchCriterion(tt, cmp, params, imark, nn, NULL); }

    // ---- node mirrors (KDTree.h:288, KDFindNearest.cxx:709-749) -------------------------------------------------
    Node* GetRoot() { if (!root) build_mirror(); return root; }
    Node* FindLeafNode(Int_t tt) {
        Node* np = GetRoot();
        while (!np->GetLeaf()) np = (tt < np->left->GetEnd()) ? np->left : np->right;
        return np;
    }
    Node* FindLeafNode(Double_t* x) {
        Node* np = GetRoot();
        while (!np->GetLeaf()) { const int k = np->cut_dim; np = (x[k] < np->left->GetBoundary(k, 1)) ? np->left : np->right; }
        return np;
    }

    // ---- smoothed estimators (KDCalcSmoothQuantities.cxx:203-389) ----------------------------------------------
    void CalcDensity(Int_t Nsmooth = 64) {
        NbkRawBuf<double> rho((size_t)numparts);
        check(nbk_calc_density(h, (int)Nsmooth, rho.data(), NULL, NBK_TREE_ORDER));
        NBK_SHIM_PARALLEL_FOR
        for (Int_t i = 0; i < numparts; i++) bucket[i].SetDensity(rho[i]);
    }
    void CalcVelDensity(Int_t Nsmooth = 64, Int_t Nsearch = 64) {
        NbkRawBuf<double> rho((size_t)numparts);
        check(nbk_calc_veldensity(h, (int)Nsmooth, (int)Nsearch, rho.data(), NBK_TREE_ORDER));
        NBK_SHIM_PARALLEL_FOR
        for (Int_t i = 0; i < numparts; i++) bucket[i].SetDensity(rho[i]);
    }
    /// hi = 0.5*sqrt(d2 of the Nsmooth-th neighbour) for every particle, indexed by ID (new[]: caller delete[]s)
    Double_t* CalcSmoothingScale(Int_t Nsmooth = 64) {
        std::vector<double> hs(numparts);
        check(nbk_smoothing_scale(h, (int)Nsmooth, hs.data(), 0));
        Double_t* out = new Double_t[numparts];
        for (Int_t i = 0; i < numparts; i++) out[i] = hs[i];
        return out;
    }

    /// KDCalcSmoothQuantities.cxx:480-614: smoothed mean velocity / velocity dispersion tensor of every particle, new[] arrays
    /// indexed by particle ID that the caller delete[]s.  densityset != 1 recomputes the densities first (and stores them in the
    /// particles like the reference's CalcDensity call does); meanvelset != 1 recomputes smvel (the argument is then ignored).
    Coordinate* CalcSmoothVel(Int_t Nsmooth = 64, int densityset = 1) {
        if (densityset != 1) CalcDensity(Nsmooth);
        std::vector<double> rho(numparts), sv((size_t)3 * numparts);
        for (Int_t i = 0; i < numparts; i++) rho[i] = bucket[i].GetDensity();                 // tree order
        check(nbk_calc_smooth_vel(h, (int)Nsmooth, rho.data(), sv.data(), NBK_TREE_ORDER));
        Coordinate* out = new Coordinate[numparts];
        for (Int_t i = 0; i < numparts; i++) { Coordinate& c = out[bucket[i].GetID()]; for (int j = 0; j < 3; j++) c[j] = sv[(size_t)3 * i + j]; }
        return out;
    }
    Matrix* CalcSmoothVelDisp(Coordinate* smvel, Int_t Nsmooth = 64, int densityset = 1, int meanvelset = 1) {
        if (densityset != 1) CalcDensity(Nsmooth);
        Coordinate* own = NULL;
        if (meanvelset != 1 || smvel == NULL) smvel = own = CalcSmoothVel(Nsmooth);
        std::vector<double> rho(numparts), sv((size_t)3 * numparts), sd((size_t)9 * numparts);
        for (Int_t i = 0; i < numparts; i++) {
            rho[i] = bucket[i].GetDensity();
            const Coordinate& c = smvel[bucket[i].GetID()];
            for (int j = 0; j < 3; j++) sv[(size_t)3 * i + j] = c[j];
        }
        if (own) delete[] own;
        check(nbk_calc_smooth_veldisp(h, (int)Nsmooth, rho.data(), sv.data(), sd.data(), NBK_TREE_ORDER));
        Matrix* out = new Matrix[numparts];
        for (Int_t i = 0; i < numparts; i++) { Matrix& mm = out[bucket[i].GetID()]; for (int j = 0; j < 3; j++) for (int l = 0; l < 3; l++) mm(j, l) = sd[(size_t)9 * i + 3 * j + l]; }
        return out;
    }

    Coordinate* higher_moment(int moment, Coordinate* smvel, Matrix* smveldisp, Int_t Nsmooth, int densityset, int meanvelset, int veldispset) {
        if (densityset != 1) CalcDensity(Nsmooth);
        Coordinate* own_v = NULL;
        Matrix* own_d = NULL;
        if (meanvelset != 1 || smvel == NULL) smvel = own_v = CalcSmoothVel(Nsmooth);
        if (veldispset != 1 || smveldisp == NULL) smveldisp = own_d = CalcSmoothVelDisp(smvel, Nsmooth);
        std::vector<double> rho(numparts), sv((size_t)3 * numparts), sd((size_t)9 * numparts), hm((size_t)3 * numparts);
        for (Int_t i = 0; i < numparts; i++) {
            rho[i] = bucket[i].GetDensity();
            const Coordinate& c = smvel[bucket[i].GetID()];
            const Matrix& mm = smveldisp[bucket[i].GetID()];
            for (int j = 0; j < 3; j++) { sv[(size_t)3 * i + j] = c[j]; for (int l = 0; l < 3; l++) sd[(size_t)9 * i + 3 * j + l] = mm(j, l); }
        }
        if (own_v) delete[] own_v;
        if (own_d) delete[] own_d;
        check(moment == 3 ? nbk_calc_smooth_velskew(h, (int)Nsmooth, rho.data(), sv.data(), sd.data(), hm.data(), NBK_TREE_ORDER)
                          : nbk_calc_smooth_velkurtosis(h, (int)Nsmooth, rho.data(), sv.data(), sd.data(), hm.data(), NBK_TREE_ORDER));
        Coordinate* out = new Coordinate[numparts];
        for (Int_t i = 0; i < numparts; i++) { Coordinate& c = out[bucket[i].GetID()]; for (int j = 0; j < 3; j++) c[j] = hm[(size_t)3 * i + j]; }
        return out;
    }
    /// KDCalcSmoothQuantities.cxx:617-765: smoothed velocity skewness / kurtosis per component, new[] arrays indexed by particle ID.
    /// *set != 1 recomputes the corresponding input like the reference does (the passed array is then ignored, not deleted).
    Coordinate* CalcSmoothVelSkew(Coordinate* smvel, Matrix* smveldisp, Int_t Nsmooth = 64, int densityset = 1, int meanvelset = 1, int veldispset = 1) {
        return higher_moment(3, smvel, smveldisp, Nsmooth, densityset, meanvelset, veldispset);
    }
    Coordinate* CalcSmoothVelKurtosis(Coordinate* smvel, Matrix* smveldisp, Int_t Nsmooth = 64, int densityset = 1, int meanvelset = 1, int veldispset = 1) {
        return higher_moment(4, smvel, smveldisp, Nsmooth, densityset, meanvelset, veldispset);
    }

    /// single-target forms (KDCalcSmoothQuantities.cxx:768-921, 1092-1207): gather-only, value returned.  One small device
    /// query per call; loops over many targets should use nbk_calc_density_particles / nbk_calc_veldensity_particles.
    Double_t CalcDensityParticle(Int_t target, Int_t Nsmooth = 64) {
        int32_t q = (int32_t)target; double out = 0;
        std::lock_guard<std::mutex> g(dev_mutex);
        check(nbk_calc_density_particles(h, (int)Nsmooth, 1, &q, &out, 0));
        return out;
    }
    Double_t CalcVelDensityParticle(Int_t target, Int_t Nsmooth = 64, Int_t Nsearch = 64, int iflag = 0, PriorityQueue* pq = NULL,
                                    PriorityQueue* pq2 = NULL, Int_t* nnIDs = NULL, Double_t* vdist = NULL) {
        (void)iflag; (void)pq; (void)pq2; (void)nnIDs; (void)vdist;     // caller-provided scratch of the reference: not needed
        int32_t q = (int32_t)target; double out = 0;
        std::lock_guard<std::mutex> g(dev_mutex);
        check(nbk_calc_veldensity_particles(h, (int)Nsmooth, (int)Nsearch, 1, &q, &out, 0));
        return out;
    }
    Double_t CalcDensityPosition(Double_t* x, Int_t Nsmooth = 64, Double_t* v = NULL) {
        (void)v;                                                          // only read by the reference's phase-space trees
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]}, out = 0;
        std::lock_guard<std::mutex> g(dev_mutex);
        check(nbk_calc_density_points(h, (int)Nsmooth, 1, xx, &out, 0));
        return out;
    }
    Double_t CalcVelDensityPosition(Double_t* x, Double_t* v, Int_t Nsmooth = 64, Int_t Nsearch = 64) {
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]}, vv[3] = {(double)v[0], (double)v[1], (double)v[2]}, out = 0;
        std::lock_guard<std::mutex> g(dev_mutex);
        check(nbk_calc_veldensity_points(h, (int)Nsmooth, (int)Nsearch, 1, xx, vv, &out, 0));
        return out;
    }
    Double_t CalcDensityPosition(Coordinate x, Int_t Nsmooth = 64, Coordinate v = Coordinate(0.)) { (void)v; return CalcDensityPosition(x.GetCoord(), Nsmooth); }
    Double_t CalcVelDensityPosition(Coordinate x, Coordinate v, Int_t Nsmooth = 64, Int_t Nsearch = 64) { return CalcVelDensityPosition(x.GetCoord(), v.GetCoord(), Nsmooth, Nsearch); }

    /// KDCalcSmoothQuantities.cxx:1704-1735: kernel-weighted sum over a caller-supplied neighbour list (host data in, one
    /// number out), evaluated with the tree's kernel table exactly as the reference's Wsm does (:12-15).  The queue form
    /// holds squared distances and is emptied; the array form holds distances in descending order.
    Double_t CalcSmoothLocalValue(Int_t Nsmooth, PriorityQueue* pq, Double_t* weight) {
        const Double_t hi = 0.5 * std::sqrt(pq->TopPriority());
        const Double_t norm = 1.0 / std::pow(hi, (Double_t)(info.nd * 1.));
        Double_t value = 0;
        for (Int_t j = 0; j < Nsmooth; j++) { value += wsm(std::sqrt(pq->TopPriority()) / hi) * norm * weight[j]; pq->Pop(); }
        return value;
    }
    Double_t CalcSmoothLocalValue(Int_t Nsmooth, Double_t* dist, Double_t* weight) {
        const Double_t hi = 0.5 * dist[0];
        const Double_t norm = 1.0 / std::pow(hi, (Double_t)(info.nd * 1.));
        Double_t value = 0;
        for (Int_t j = 0; j < Nsmooth; j++) value += wsm(dist[j] / hi) * norm * weight[j];
        return value;
    }

    // ---- FOF (KDFOF.cxx:29-265) --------------------------------------------------------------------------------
    Int_t* FOF(Double_t fdist, Int_t& numgroup, Int_t minnum = 8, int order = 0, Int_tree_t* pHead = NULL, Int_tree_t* pNext = NULL,
               Int_tree_t* pTail = NULL, Int_tree_t* pLen = NULL, int ipcheckflag = 0, FOFcheckfunc check_ = Pnocheck, Double_t* params = NULL) {
        std::vector<int32_t> pre;
        fill_precheck(pre, ipcheckflag, check_, params);
        return run_fof(pre, pHead, pNext, pTail, pLen, numgroup, [&](const int32_t* pc, int32_t* g, int64_t* ng, nbk_fof_lists* l) {
            return nbk_fof(h, (double)fdist, (int)minnum, order, pc, g, ng, l, 0);
        });
    }
    Int_t* FOFCriterion(FOFcompfunc cmp, Double_t* params, Int_t& numgroups, Int_t minnum = 8, int order = 0, int ipcheckflag = 0,
                        FOFcheckfunc check_ = Pnocheck, Int_tree_t* pHead = NULL, Int_tree_t* pNext = NULL, Int_tree_t* pTail = NULL,
                        Int_tree_t* pLen = NULL) {
        const int crit = crit_code(cmp, "FOFCriterion");
        std::vector<int32_t> pre;
        fill_precheck(pre, ipcheckflag, check_, params);
        double pr[16];
        for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        return run_fof(pre, pHead, pNext, pTail, pLen, numgroups, [&](const int32_t* pc, int32_t* g, int64_t* ng, nbk_fof_lists* l) {
            return nbk_fof_criterion(h, crit, pr, (int)minnum, order, pc, g, ng, l, 0);
        });
    }

    /// KDFOF.cxx:268-378.  `check_` is evaluated here on the host for every particle (it is the caller's code); the
    /// device receives the values.  (ipcheckflag is not read by the reference's implementation either.)
    Int_t* FOFCriterionSetBasisForLinks(FOFcompfunc cmp, Double_t* params, Int_t& numgroup, Int_t minnum = 8, int order = 0, int ipcheckflag = 0,
                                        FOFcheckfunc check_ = Pnocheck, Int_tree_t* pHead = NULL, Int_tree_t* pNext = NULL, Int_tree_t* pTail = NULL,
                                        Int_tree_t* pLen = NULL) {
        (void)ipcheckflag;
        const int crit = crit_code(cmp, "FOFCriterionSetBasisForLinks");
        std::vector<int32_t> pre;
        fill_precheck(pre, 1, check_, params);
        double pr[16];
        for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        return run_fof(pre, pHead, pNext, pTail, pLen, numgroup, [&](const int32_t* pc, int32_t* g, int64_t* ng, nbk_fof_lists* l) {
            return nbk_fof_criterion_basis(h, crit, pr, (int)minnum, order, pc, g, ng, l, 0);
        });
    }

    /// KDFOF.cxx:686-734 FOFCriterionParticle: grows group iGroup from particle `target` (a tree index).  pfof is the caller's
    /// group array indexed by ID; every particle reachable from the target through chains of cmp-linked particles whose pfof is
    /// >= 0 and != iGroup joins the group (members of other groups are taken over, KDLeafNode.cxx:603-616; particles with a
    /// negative tag and the group's existing members are not walked through).  Returns pLen[iGroup] = the group's size.  The
    /// scratch arrays of the reference's breadth-first search (pGroupHead, Fifo) are not needed; pHead / pNext / pTail (tree
    /// index space, optional) are rebuilt for the group: the target first, then its members in ascending tree index.
    /// Device work: one component search over the whole tree (nbk_fof_roots) per call.
    Int_t FOFCriterionParticle(FOFcompfunc cmp, Int_t* pfof, Int_t target, Int_t iGroup, Double_t* params, Int_tree_t* pGroupHead = NULL,
                               Int_tree_t* Fifo = NULL, Int_tree_t* pHead = NULL, Int_tree_t* pTail = NULL, Int_tree_t* pNext = NULL, Int_tree_t* pLen = NULL) {
        (void)Fifo;
        const int crit = crit_code(cmp, "FOFCriterionParticle");
        if (target < 0 || target >= numparts) throw std::runtime_error("nbk shim: particle index out of range");
        double pr[16]; for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        const Int_t tid = bucket[target].GetID();
        std::vector<int32_t> excl((size_t)numparts), root((size_t)numparts);
        NBK_SHIM_PARALLEL_FOR
        for (Int_t i = 0; i < numparts; i++) excl[(size_t)i] = (pfof[i] < 0 || (pfof[i] == iGroup && i != tid)) ? 1 : 0;      // by ID
        {
            std::lock_guard<std::mutex> g(dev_mutex);
            check(nbk_fof_roots(h, crit, 0.0, pr, excl.data(), root.data(), 0));
        }
        const int32_t rt = root[(size_t)tid];
        Int_t len = 0;
        for (Int_t i = 0; i < numparts; i++) { if (root[(size_t)i] == rt && rt >= 0) pfof[i] = iGroup; if (pfof[i] == iGroup) len++; }
        if (pGroupHead) pGroupHead[iGroup] = target;
        if (pLen) pLen[iGroup] = len;
        if (pHead || pNext || pTail) {
            Int_t prev = target, tail = target;
            for (Int_t i = 0; i < numparts; i++) if (i != target && pfof[bucket[i].GetID()] == iGroup) tail = i;
            for (Int_t i = 0; i < numparts; i++) {
                const bool in = pfof[bucket[i].GetID()] == iGroup;
                if (pHead) pHead[i] = in ? target : i;
                if (pTail) pTail[i] = in ? tail : i;
                if (pNext) pNext[i] = -1;
            }
            if (pNext) for (Int_t i = 0; i < numparts; i++) if (i != target && pfof[bucket[i].GetID()] == iGroup) { pNext[prev] = i; prev = i; }
        }
        return len;
    }
    /// forget the cached FOFcheckfunc values (call after changing the particle fields the function reads)
    void InvalidateCheckCache() { std::lock_guard<std::mutex> g(chk_mutex); chk_vals.clear(); chk_fn = nullptr; }

    // ---- ordering (KDTree.cxx:1358-1362) -----------------------------------------------------------------------
    void OverWriteInputOrder() {
        iresetorder = false;
        for (Int_t i = 0; i < numparts; i++) bucket[i].SetID(i);
    }
    void SetResetOrder(bool a) { iresetorder = a; }

private:
    /// bucket[i] <- bucket[src(i)] for a permutation src, moving every particle twice (into a scratch array and back); both
    /// passes are parallel and the scratch array is raw storage: no element is default-constructed or copied
    /// (KDTree.cxx:328-370 does it with in-place quickselect swaps; ~KDTree with std::sort, :1347)
    template <class F>
    void permute(F src) { nbk_permute_records(bucket, (int64_t)numparts, [&](int64_t i) { return (int64_t)src((Int_t)i); }); }
    void require_vel_tree(const char* who) {
        if (info.treetype != TVEL) throw std::runtime_error(std::string("nbk shim: ") + who + " needs a TVEL tree (on any other tree the reference prunes velocity queries with position cut planes)");
    }
    void require_pos_tree(const char* who) {
        if (info.treetype != TPHYS && info.treetype != TPHS) throw std::runtime_error(std::string("nbk shim: ") + who + " has a device implementation on position trees only");
    }
    /// Wsm of the reference (KDCalcSmoothQuantities.cxx:12-15) on the tree's kernel table, r = rij/hi
    Double_t wsm(Double_t r) {
        const std::vector<Double_t>& K = kernel_table();
        const int size = info.kernres;
        const Double_t delta = 2.0 / (Double_t)(size - 1);
        const int i = (int)(r * 0.5 * (size - 1));
        if (i < size - 1) return K[i] + (K[i + 1] - K[i]) * (r - delta * i) / delta;
        return K[i];
    }
    /// CSR protocol for a single query row in ONE device round trip: the call is made with the capacity the thread's last
    /// row needed (at least 256 entries); only when the row is larger does NBK_ERR_CAPACITY report the size for a second call
    template <class F>
    void csr_row(F call, int flags, std::vector<int32_t>& idx, std::vector<double>* d2) {
        static thread_local size_t hint = 256;
        int64_t off[2], tot = 0;
        std::lock_guard<std::mutex> g(dev_mutex);
        idx.resize(hint);
        if (d2) d2->resize(hint);
        int rc = call(off, idx.data(), d2 ? d2->data() : (double*)NULL, (int64_t)idx.size(), &tot, flags);
        if (rc == NBK_ERR_CAPACITY) {
            hint = (size_t)tot + (size_t)tot / 4 + 16;
            idx.resize(hint);
            if (d2) d2->resize(hint);
            rc = call(off, idx.data(), d2 ? d2->data() : (double*)NULL, (int64_t)idx.size(), &tot, flags);
        }
        check(rc);
        idx.resize((size_t)tot);
        if (d2) d2->resize((size_t)tot);
    }
    /// Per-thread block cache behind SearchBallPosTagged(Int_t tt) / SearchCriterionTagged(Int_t tt): like knn_cached, the
    /// first call of a thread answers the whole block of consecutive tree indices around tt with one batched device query
    /// (CSR), the following calls are copies from host memory (reference callers loop over every particle:
    /// tests/test_kdtree.cxx:325-341).  crit < 0: ball of radius^2 r2; otherwise NBK_FOF3D / NBK_FOF6D with params.
    struct BallBlock {
        unsigned long long serial = 0; Int_t b0 = 0, b1 = 0; int crit = -2; double r2 = -1, params[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<int64_t> off; std::vector<int32_t> idx;
    };
    static BallBlock& tls_ball() { static thread_local BallBlock b; return b; }
    std::vector<Int_t> ball_cached(Int_t tt, double r2, int crit, const Double_t* params) {
        if (tt < 0 || tt >= numparts) throw std::runtime_error("nbk shim: particle index out of range");
        BallBlock& c = tls_ball();
        double pr[16] = {0};
        if (crit >= 0 && params) for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        bool same = c.serial == serial && c.crit == crit && c.r2 == r2 && tt >= c.b0 && tt < c.b1;
        for (int j = 0; same && j < 8; j++) same = c.params[j] == pr[j];
        if (!same) {
            const Int_t B = 4096;
            const Int_t b0 = (tt / B) * B, b1 = std::min(numparts, b0 + B);
            const int64_t m = b1 - b0;
            std::vector<int32_t> q((size_t)m);
            for (int64_t i = 0; i < m; i++) q[(size_t)i] = (int32_t)(b0 + i);
            c.off.resize((size_t)m + 1);
            if (c.idx.size() < (size_t)64 * (size_t)m) c.idx.resize((size_t)64 * (size_t)m);
            int64_t tot = 0;
            std::lock_guard<std::mutex> g(dev_mutex);
            auto call = [&]() {
                return crit < 0 ? nbk_ball_particles(h, r2, m, q.data(), c.off.data(), c.idx.data(), NULL, (int64_t)c.idx.size(), &tot, 0)
                                : nbk_search_criterion_particles(h, crit, pr, m, q.data(), c.off.data(), c.idx.data(), NULL, (int64_t)c.idx.size(), &tot, 0);
            };
            int rc = call();
            if (rc == NBK_ERR_CAPACITY) { c.idx.resize((size_t)tot + (size_t)tot / 4 + 16); rc = call(); }
            c.serial = 0;                       // the block is invalid until the call has succeeded
            check(rc);
            c.serial = serial; c.b0 = b0; c.b1 = b1; c.crit = crit; c.r2 = r2;
            for (int j = 0; j < 8; j++) c.params[j] = pr[j];
        }
        const int64_t r0 = c.off[(size_t)(tt - c.b0)], r1 = c.off[(size_t)(tt - c.b0) + 1];
        return std::vector<Int_t>(c.idx.begin() + r0, c.idx.begin() + r1);
    }
    /// FOFcheckfunc values of every particle in tree order, computed once per (function, params) and kept: the caller's
    /// function runs on the host, the device receives the values.  InvalidateCheckCache() after changing the particle fields
    /// the function reads.
    std::mutex chk_mutex;
    std::vector<int32_t> chk_vals;
    FOFcheckfunc chk_fn = nullptr;
    double chk_params[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int32_t* check_values(FOFcheckfunc fn, Double_t* params) {
        double pr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (params) for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        std::lock_guard<std::mutex> g(chk_mutex);
        bool same = chk_fn == fn && !chk_vals.empty();
        for (int j = 0; same && j < 8; j++) same = chk_params[j] == pr[j];
        if (!same) {
            chk_vals.resize((size_t)numparts);
            NBK_SHIM_PARALLEL_FOR
            for (Int_t i = 0; i < numparts; i++) chk_vals[(size_t)i] = fn(bucket[i], params);
            chk_fn = fn;
            for (int j = 0; j < 8; j++) chk_params[j] = pr[j];
        }
        return chk_vals.data();
    }
    /// true on a TPHS tree built with Aniso = -1; throws on a TPHS tree with a metric (Aniso >= 0)
    bool phase_tree(const char* who) const {
        if (info.treetype != TPHS) return false;
        if (anisotropic != -1) throw std::runtime_error(std::string("nbk shim: ") + who + " on a TPHS tree with Aniso >= 0 is the reference's metric search, which has no device implementation (build the tree with Aniso = -1 for the plain 6D search)");
        return true;
    }
    /// crit == -2: plain search; -3: phase-space search; otherwise filtered (crit -1: check function only, >= 0: NBK_FOF3D / NBK_FOF6D)
    void knn_cached(Int_t tt, Int_t* nn, Double_t* dist2, Int_t k, int flags, int crit = -2, FOFcheckfunc checkfn = nullptr, Double_t* params = NULL) {
        if (tt < 0 || tt >= numparts) throw std::runtime_error("nbk shim: particle index out of range");
        KnnBlock& c = tls_block();
        double pr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (crit > -2 && params) for (int j = 0; j < 8; j++) pr[j] = (double)params[j];       // criterion AND check forms: compared by value
        bool same = c.serial == serial && c.k == k && c.flags == flags && c.crit == crit && c.checkfn == checkfn && tt >= c.b0 && tt < c.b1;
        for (int j = 0; same && j < 8; j++) same = c.params[j] == pr[j];
        if (!same) {
            const Int_t B = block_size(k);
            const Int_t b0 = (tt / B) * B, b1 = std::min(numparts, b0 + B);
            c.nn.resize((size_t)(b1 - b0) * k);
            c.d2.resize((size_t)(b1 - b0) * k);
            if (crit == -2) {
                std::lock_guard<std::mutex> g(dev_mutex);
                check(nbk_knn_particles(h, (int)k, b0, b1, c.nn.data(), c.d2.data(), flags));
            } else if (crit == -3) {
                std::lock_guard<std::mutex> g(dev_mutex);
                check(nbk_knn_phase_particles(h, (int)k, b0, b1, c.nn.data(), c.d2.data(), flags));
            } else {
                const int32_t* chk = checkfn ? check_values(checkfn, params) : NULL;                     // tree order, once per (function, params)
                std::lock_guard<std::mutex> g(dev_mutex);
                check(nbk_knn_filtered_particles(h, (int)k, b0, b1, crit, pr, chk, c.nn.data(), c.d2.data(), flags | NBK_TREE_ORDER));
            }
            c.serial = serial; c.b0 = b0; c.b1 = b1; c.k = k; c.flags = flags; c.crit = crit; c.checkfn = checkfn;
            for (int j = 0; j < 8; j++) c.params[j] = pr[j];
        }
        const size_t row = (size_t)(tt - c.b0) * k;
        for (Int_t j = 0; j < k; j++) { nn[j] = c.nn[row + j]; dist2[j] = c.d2[row + j]; }
    }
    void knn_filtered_point(const Double_t* x, const Double_t* v, int crit, FOFcheckfunc checkfn, Double_t* params, Int_t* nn, Double_t* dist2, Int_t k) {
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]}, vv[3] = {0, 0, 0}, pr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (v) for (int j = 0; j < 3; j++) vv[j] = (double)v[j];
        if (crit >= 0 && params) for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        std::vector<int32_t> n32(k);
        std::vector<double> d2(k);
        const int32_t* chk = checkfn ? check_values(checkfn, params) : NULL;
        {
            std::lock_guard<std::mutex> g(dev_mutex);
            check(nbk_knn_filtered_points(h, (int)k, 1, xx, v ? vv : NULL, crit, pr, chk, n32.data(), d2.data(),
                                          NBK_TREE_ORDER | (period ? NBK_KNN_TREE_FORM : 0)));
        }
        for (Int_t j = 0; j < k; j++) { nn[j] = n32[j]; dist2[j] = d2[j]; }
    }
    void knn_range(Int_t q0, Int_t q1, Int_t* nn, Double_t* dist2, Int_t k, int flags) {
        std::vector<int32_t> n32((size_t)(q1 - q0) * k);
        std::vector<double> d2((size_t)(q1 - q0) * k);
        {
            std::lock_guard<std::mutex> g(dev_mutex);
            check(nbk_knn_particles(h, (int)k, q0, q1, n32.data(), d2.data(), flags));
        }
        for (size_t j = 0; j < n32.size(); j++) { nn[j] = n32[j]; dist2[j] = d2[j]; }
    }
    void knn_all(Int_t** nn, Double_t** dist2, Int_t k, int flags, bool phase = false) {
        const Int_t chunk = 1 << 20;
        std::vector<int32_t> n32((size_t)std::min(chunk, numparts) * k);
        std::vector<double> d2(n32.size());
        for (Int_t q0 = 0; q0 < numparts; q0 += chunk) {
            Int_t q1 = std::min(numparts, q0 + chunk);
            {
                std::lock_guard<std::mutex> g(dev_mutex);
                check(phase ? nbk_knn_phase_particles(h, (int)k, q0, q1, n32.data(), d2.data(), flags)
                            : nbk_knn_particles(h, (int)k, q0, q1, n32.data(), d2.data(), flags));
            }
            for (Int_t i = q0; i < q1; i++)
                for (Int_t j = 0; j < k; j++) { nn[i][j] = n32[(size_t)(i - q0) * k + j]; dist2[i][j] = d2[(size_t)(i - q0) * k + j]; }
        }
    }
    void fill_precheck(std::vector<int32_t>& pre, int ipcheckflag, FOFcheckfunc check_, Double_t* params) {
        if (!ipcheckflag) return;
        pre.resize(numparts);
        for (Int_t i = 0; i < numparts; i++) pre[bucket[i].GetID()] = check_(bucket[i], params);   // KDFOF.cxx:65
    }
    template <class F>
    Int_t* run_fof(const std::vector<int32_t>& pre, Int_tree_t* pHead, Int_tree_t* pNext, Int_tree_t* pTail, Int_tree_t* pLen, Int_t& numgroup, F call) {
        const bool any = pHead || pNext || pTail || pLen;
        std::vector<int32_t> g(numparts), len(pLen ? (size_t)numparts + 1 : 0), hd(pHead ? numparts : 0), nx(pNext ? numparts : 0), tl(pTail ? numparts : 0);
        nbk_fof_lists lists = {pHead ? hd.data() : NULL, pNext ? nx.data() : NULL, pTail ? tl.data() : NULL, pLen ? len.data() : NULL};
        int64_t ng = 0;
        check(call(pre.empty() ? NULL : pre.data(), g.data(), &ng, any ? &lists : NULL));
        numgroup = (Int_t)ng;
        if (pLen) for (int64_t i = 0; i <= ng && i < numparts; i++) pLen[i] = len[i];
        for (Int_t i = 0; i < numparts; i++) {
            if (pHead) pHead[i] = hd[i];
            if (pNext) pNext[i] = nx[i];
            if (pTail) pTail[i] = tl[i];
        }
        Int_t* out = new Int_t[numparts];
        for (Int_t i = 0; i < numparts; i++) out[i] = g[i];
        return out;
    }
};

}  // namespace NBody
#endif
