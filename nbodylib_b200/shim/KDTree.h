/*! \file KDTree.h (nbodylib_b200 shim)
 *  Header-only NBody::KDTree with the reference's public interface (reference src/KDTree/KDTree.h:81-657) whose
 *  methods marshal into the C ABI of include/nbk.h (libnbk.so, CUDA sm_100a).  A program written against the
 *  reference compiles against this header unchanged for the calls on the hot path:
 *
 *    ctor (Particle*, numparts, bucket_size, TreeType, KernType, KernRes, SplittingCriterion, Aniso, ScaleSpace, Period)
 *    GetNumNodes / GetNumLeafNodes / GetBucketSize / GetTreeType / GetKernType / GetKernNorm / GetPeriod
 *    FindNearest / FindNearestPos (Int_t tt | Double_t* x | Coordinate | whole system)
 *    SearchBallPosTagged (Int_t tt | Double_t* x | Coordinate; array and vector forms)
 *    CalcDensity, CalcVelDensity, CalcSmoothingScale (new: north star), FOF, FOFCriterion (FOF3d / FOF6d)
 *    OverWriteInputOrder, SetResetOrder, ~KDTree (restores the caller's particle order)
 *
 *  Semantics kept from the reference: the caller's Particle array is permuted IN PLACE into tree order and
 *  Particle::id is overwritten with the input index (KDTree.cxx:1291); nn[] / tt / tagged[] are tree-order indices;
 *  FOF results are new[]-allocated arrays indexed by ID that the caller delete[]s; the destructor sorts the array
 *  back by ID unless OverWriteInputOrder() was called (KDTree.cxx:1340-1362).
 *  Differences: errors throw std::runtime_error instead of printf+exit; calls without a device implementation
 *  (TPROJ/TMETRIC trees, host FOFcompfunc callbacks other than FOF3d/FOF6d, CalcSmoothVel*, FOFNN*) throw -- there is
 *  no CPU fallback.  Per-particle calls launch one small kernel each; loops over all particles should use the
 *  whole-system forms.
 */
#ifndef NBK_SHIM_KDTREE_H
#define NBK_SHIM_KDTREE_H

#ifndef NBK_USE_REFERENCE_PARTICLE
#include "Particle.h"
#endif
#include <algorithm>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/nbk.h"

namespace NBody {

class KDTree {
public:
    const static int TPHYS = 0, TPROJ = 1, TVEL = 2, TPHS = 3, TMETRIC = 4;
    const static int KSPH = 0, KGAUSS = 1, KEPAN = 2, KTH = 3;
    const static int KDTREE_SPLIT_ENTROPY = 1, KDTREE_SPLIT_DISPERSION = 2, KDTREE_SPLIT_MAXINTERPARTICLESPACING = 3, KDTREE_SPLIT_SPREAD = 0;

private:
    nbk_tree* h = nullptr;
    Particle* bucket = nullptr;
    Int_t numparts = 0;
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

public:
    KDTree(Particle* p, Int_t nparts, Int_t bucket_size = 16, int TreeType = TPHYS, int KernType = KEPAN, int KernRes = 1000,
           int SplittingCriterion = KDTREE_SPLIT_SPREAD, int Aniso = 0, int ScaleSpace = 0, Double_t* Period = NULL,
           Double_t** metric = NULL, bool iBuildInParallel = true, bool iKeepInputOrder = false, Double_t Rdistadapt = -1,
           Double_t AdaptiveMedianFac = 0.0, Int_t min_bucket_size = 16)
        : bucket(p), numparts(nparts) {
        (void)Aniso; (void)metric; (void)iBuildInParallel; (void)min_bucket_size;
        if (ScaleSpace) throw std::runtime_error("nbk shim: ScaleSpace has no device implementation");
        if (iKeepInputOrder || Rdistadapt > 0 || AdaptiveMedianFac > 0) throw std::runtime_error("nbk shim: adaptive / keep-order builds have no device implementation");
        for (Int_t i = 0; i < numparts; i++) bucket[i].SetID(i);                    // KDTree.cxx:1291
        std::vector<Double_t> mass(numparts);
        for (Int_t i = 0; i < numparts; i++) mass[i] = bucket[i].GetMass();
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
        std::vector<int32_t> order(numparts);
        check(nbk_get_order(h, order.data(), 0));
        std::vector<Particle> tmp(bucket, bucket + numparts);
        for (Int_t i = 0; i < numparts; i++) bucket[i] = tmp[order[i]];
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
            std::vector<Particle> tmp(bucket, bucket + numparts);
            for (Int_t i = 0; i < numparts; i++) bucket[tmp[i].GetID()] = tmp[i];
        }
    }

    Int_t GetNumNodes() { return info.num_nodes; }
    Int_t GetNumLeafNodes() { return info.num_leaves; }
    Int_t GetBucketSize() { return info.bucket; }
    Int_t GetTreeType() { return info.treetype; }
    Int_t GetKernType() { return info.kerntype; }
    Double_t GetKernNorm() { return info.kernnorm; }
    Double_t GetPeriod(int j) { return period[j]; }
    nbk_tree* GetHandle() { return h; }

    // ---- nearest neighbours (KDFindNearest.cxx:247-334, 444-554) ----------------------------------------------
    void FindNearestPos(Int_t tt, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { knn_cached(tt, nn, dist2, Nsearch, 0); }
    void FindNearest(Int_t tt, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { knn_cached(tt, nn, dist2, Nsearch, NBK_KNN_TREE_FORM); }
    void FindNearestPos(Double_t* x, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) {
        std::vector<double> d2(Nsearch);
        std::vector<int32_t> n32(Nsearch);
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]};
        check(nbk_knn_points(h, (int)Nsearch, 1, xx, n32.data(), d2.data(), 0));
        for (Int_t j = 0; j < Nsearch; j++) { nn[j] = n32[j]; dist2[j] = d2[j]; }
    }
    void FindNearest(Double_t* x, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { FindNearestPos(x, nn, dist2, Nsearch); }
    void FindNearestPos(Coordinate x, Int_t* nn, Double_t* dist2, Int_t Nsearch = 64) { FindNearestPos(x.GetCoord(), nn, dist2, Nsearch); }
    /// whole-system forms: nn[i][j], dist2[i][j] for every tree index i
    void FindNearestPos(Int_t** nn, Double_t** dist2, Int_t Nsearch = 64) { knn_all(nn, dist2, Nsearch, 0); }
    void FindNearest(Int_t** nn, Double_t** dist2, Int_t Nsearch = 64) { knn_all(nn, dist2, Nsearch, NBK_KNN_TREE_FORM); }

    // ---- fixed radius (KDFindNearest.cxx:618-688) ------------------------------------------------------------
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
    std::vector<Int_t> SearchBallPosTagged(Int_t tt, Double_t fdist2) {
        int32_t q = (int32_t)tt;
        int64_t off[2], tot = 0;
        check(nbk_ball_particles(h, (double)fdist2, 1, &q, off, NULL, 0, &tot, 0));
        std::vector<int32_t> idx((size_t)std::max<int64_t>(tot, 1));
        check(nbk_ball_particles(h, (double)fdist2, 1, &q, off, idx.data(), (int64_t)idx.size(), &tot, 0));
        return std::vector<Int_t>(idx.begin(), idx.begin() + tot);
    }
    std::vector<Int_t> SearchBallPosTagged(Double_t* x, Double_t fdist2) {
        double xx[3] = {(double)x[0], (double)x[1], (double)x[2]};
        int64_t off[2], tot = 0;
        check(nbk_ball_points(h, (double)fdist2, 1, xx, off, NULL, 0, &tot, 0));
        std::vector<int32_t> idx((size_t)std::max<int64_t>(tot, 1));
        check(nbk_ball_points(h, (double)fdist2, 1, xx, off, idx.data(), (int64_t)idx.size(), &tot, 0));
        return std::vector<Int_t>(idx.begin(), idx.begin() + tot);
    }
    std::vector<Int_t> SearchBallPosTagged(Coordinate x, Double_t fdist2) { return SearchBallPosTagged(x.GetCoord(), fdist2); }

    // ---- smoothed estimators (KDCalcSmoothQuantities.cxx:203-389) ----------------------------------------------
    void CalcDensity(Int_t Nsmooth = 64) {
        std::vector<double> rho(numparts);
        check(nbk_calc_density(h, (int)Nsmooth, rho.data(), NULL, NBK_TREE_ORDER));
        for (Int_t i = 0; i < numparts; i++) bucket[i].SetDensity(rho[i]);
    }
    void CalcVelDensity(Int_t Nsmooth = 64, Int_t Nsearch = 64) {
        std::vector<double> rho(numparts);
        check(nbk_calc_veldensity(h, (int)Nsmooth, (int)Nsearch, rho.data(), NBK_TREE_ORDER));
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
        // inline criteria are recognised by address inside the caller's translation unit
        int crit;
        if (cmp == (FOFcompfunc)&FOF3d) crit = NBK_FOF3D;
        else if (cmp == (FOFcompfunc)&FOF6d) crit = NBK_FOF6D;
        else throw std::runtime_error("nbk shim: FOFCriterion supports FOF3d and FOF6d; host callbacks cannot run on the device");
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
        int crit;
        if (cmp == (FOFcompfunc)&FOF3d) crit = NBK_FOF3D;
        else if (cmp == (FOFcompfunc)&FOF6d) crit = NBK_FOF6D;
        else throw std::runtime_error("nbk shim: FOFCriterionSetBasisForLinks supports FOF3d and FOF6d; host callbacks cannot run on the device");
        std::vector<int32_t> pre;
        fill_precheck(pre, 1, check_, params);
        double pr[16];
        for (int j = 0; j < 8; j++) pr[j] = (double)params[j];
        return run_fof(pre, pHead, pNext, pTail, pLen, numgroup, [&](const int32_t* pc, int32_t* g, int64_t* ng, nbk_fof_lists* l) {
            return nbk_fof_criterion_basis(h, crit, pr, (int)minnum, order, pc, g, ng, l, 0);
        });
    }

    // ---- ordering (KDTree.cxx:1358-1362) -----------------------------------------------------------------------
    void OverWriteInputOrder() {
        iresetorder = false;
        for (Int_t i = 0; i < numparts; i++) bucket[i].SetID(i);
    }
    void SetResetOrder(bool a) { iresetorder = a; }

private:
    void knn_cached(Int_t tt, Int_t* nn, Double_t* dist2, Int_t k, int flags) {
        if (tt < 0 || tt >= numparts) throw std::runtime_error("nbk shim: particle index out of range");
        KnnBlock& c = tls_block();
        if (c.serial != serial || c.k != k || c.flags != flags || tt < c.b0 || tt >= c.b1) {
            const Int_t B = block_size(k);
            const Int_t b0 = (tt / B) * B, b1 = std::min(numparts, b0 + B);
            c.nn.resize((size_t)(b1 - b0) * k);
            c.d2.resize((size_t)(b1 - b0) * k);
            {
                std::lock_guard<std::mutex> g(dev_mutex);
                check(nbk_knn_particles(h, (int)k, b0, b1, c.nn.data(), c.d2.data(), flags));
            }
            c.serial = serial; c.b0 = b0; c.b1 = b1; c.k = k; c.flags = flags;
        }
        const size_t row = (size_t)(tt - c.b0) * k;
        for (Int_t j = 0; j < k; j++) { nn[j] = c.nn[row + j]; dist2[j] = c.d2[row + j]; }
    }
    void knn_range(Int_t q0, Int_t q1, Int_t* nn, Double_t* dist2, Int_t k, int flags) {
        std::vector<int32_t> n32((size_t)(q1 - q0) * k);
        std::vector<double> d2((size_t)(q1 - q0) * k);
        check(nbk_knn_particles(h, (int)k, q0, q1, n32.data(), d2.data(), flags));
        for (size_t j = 0; j < n32.size(); j++) { nn[j] = n32[j]; dist2[j] = d2[j]; }
    }
    void knn_all(Int_t** nn, Double_t** dist2, Int_t k, int flags) {
        const Int_t chunk = 1 << 20;
        std::vector<int32_t> n32((size_t)std::min(chunk, numparts) * k);
        std::vector<double> d2(n32.size());
        for (Int_t q0 = 0; q0 < numparts; q0 += chunk) {
            Int_t q1 = std::min(numparts, q0 + chunk);
            check(nbk_knn_particles(h, (int)k, q0, q1, n32.data(), d2.data(), flags));
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
