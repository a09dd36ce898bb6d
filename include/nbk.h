/* include/nbk.h -- the drop-in boundary: a plain C ABI over the B200 kd-tree hot path.
 *
 * The reference (pelahi/NBodylib) has no FFI layer: consumers include <KDTree.h> and link the C++
 * class NBody::KDTree (reference src/KDTree/KDTree.h:81-657).  This header is what a binding for that
 * class binds instead: every entry point below names the reference interface it replaces.  The
 * header-only C++ class in nbodylib_b200/shim/ (same class / method names as the reference) is a thin
 * marshalling layer over these calls; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - all functions return NBK_OK (0) or a negative nbk_status; nbk_last_error() gives the text
 *     (the reference printf()s and exit()s instead: KDCalcSmoothQuantities.cxx:205-208).
 *   - "tree index" = position of a particle in tree order (the reference permutes the caller's array
 *     into that order, KDTree.cxx:328-370); "ID" = position in the caller's input order
 *     (KDTree.cxx:1291).  nbk_get_order() returns ID-at-tree-index so a caller can permute its array.
 *   - pointers are HOST memory unless the call is given NBK_DEVICE_PTRS, in which case every in/out
 *     array of that call is device memory on the tree's GPU (no copies; used for the resident-data
 *     benchmark and by callers that already hold particles on the GPU).
 *   - there is no CPU fallback anywhere: without a CUDA device nbk_create fails with NBK_ERR_CUDA.
 */
#ifndef NBK_H
#define NBK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbk_tree nbk_tree;

typedef enum {
    NBK_OK = 0,
    NBK_ERR_ARG = -1,          /* bad argument (tree type, k, NULL pointer ...)                        */
    NBK_ERR_CUDA = -2,         /* CUDA runtime error / no device                                       */
    NBK_ERR_UNSUPPORTED = -3,  /* valid reference call with no device implementation (no CPU fallback) */
    NBK_ERR_NOMEM = -4,
    NBK_ERR_CAPACITY = -5      /* caller-provided output buffer too small (required size reported)     */
} nbk_status;

/* reference KDTree.h:90  TPHYS=0,TPROJ=1,TVEL=2,TPHS=3,TMETRIC=4 */
enum { NBK_TPHYS = 0, NBK_TPROJ = 1, NBK_TVEL = 2, NBK_TPHS = 3, NBK_TMETRIC = 4 };
/* reference KDTree.h:97  KSPH=0,KGAUSS=1,KEPAN=2,KTH=3 */
enum { NBK_KSPH = 0, NBK_KGAUSS = 1, NBK_KEPAN = 2, NBK_KTH = 3 };
/* reference FOFFunc.h:30-55: the FOFcompfunc criteria that exist in-tree */
enum { NBK_FOF3D = 0, NBK_FOFVEL = 1, NBK_FOF6D = 2 };

/* flags (bitwise or) */
enum {
    NBK_DEVICE_PTRS = 1 << 0,   /* in/out arrays of this call are device pointers                              */
    NBK_TREE_ORDER = 1 << 1,    /* per-particle outputs indexed by tree index instead of by ID                  */
    NBK_STRICT_PERIODIC = 1 << 2, /* periodic kNN: dimensionally consistent edge/corner image tests instead of
                                     the reference's (KDSplitNode.cxx:1134-1146)            */
    NBK_KNN_TREE_FORM = 1 << 3, /* periodic particle kNN: FindNearest(tt) self rule instead of FindNearestPos(tt)
                                   (KDFindNearest.cxx:247-334, quirk Q3)                                        */
    NBK_STORE_F64 = 1 << 4,     /* nbk_create: keep coordinates as fp64 in HBM even if fp32 would be exact      */
    NBK_STORE_F32 = 1 << 5,     /* nbk_create: force fp32 storage (coordinates rounded if not representable)    */
    NBK_OUT_IDS = 1 << 6,       /* neighbour / ball outputs hold particle IDs instead of tree indices           */
    NBK_WARP_ALIGNED = 1 << 7   /* nbk_create: warp-aligned tree shape.  Nodes above 32 particles split at a multiple of
                                   32 (the balanced split of their 32-particle units) instead of at ceil(size/2)
                                   (KDTree.cxx:1012), so every 32 consecutive tree positions are one node and the
                                   kernels' query groups coincide with nodes for ANY particle count (with the reference
                                   rule only for powers of two; elsewhere CalcDensity is ~20 % slower).  Still a median
                                   split in the dimension of largest spread; depth, node count of the upper levels and
                                   all search / density / FOF RESULTS are unchanged -- what changes is the tree order
                                   (nbk_get_order) and the node arrays (nbk_get_nodes), which then no longer equal the
                                   reference's.  Meant for trees whose shape the caller never inspects (the slab trees
                                   of nbk_sharded.h use it).                                                          */
};

/* Strided, layout-agnostic description of the caller's particles (reference Particle.h:264-354; in the
 * default build sizeof(Particle)=88 with mass@0, position@8, velocity@32, id@60, rho@72).
 * pos/vel point at the first particle's x / vx: three consecutive reals; *_stride in bytes. */
typedef struct {
    const void* pos;  int64_t pos_stride;
    const void* vel;  int64_t vel_stride;    /* may be NULL: velocities 0 (TVEL/TPHS/veldensity/6D FOF need it)  */
    const void* mass; int64_t mass_stride;   /* may be NULL: unit masses (reference NOMASS build)                */
    int32_t real_bytes;                      /* 8 = double (reference default), 4 = float                        */
    int32_t on_device;                       /* 0: host memory, 1: device memory                                  */
} nbk_particles;

typedef struct {
    int64_t n;
    int32_t bucket, treetype, kerntype, kernres, nd;
    int32_t num_nodes, num_leaves, depth;    /* KDTree::GetNumNodes/GetNumLeafNodes (KDTree.h:282-283)             */
    int32_t store_bytes;                     /* 4 or 8: coordinate storage chosen                                  */
    int32_t periodic;
    int64_t inexact_coords;                  /* # input coordinates that fp32 storage had to round (0 => exact)   */
    double  kernnorm;                        /* KDTree::GetKernNorm (KDTree.h:287)                                 */
    double  period[3];
    double  build_ms;                        /* device time of the last build (CUDA events)                        */
    double  h2d_ms;                          /* host->device staging time of the build                             */
    double  last_kernel_ms;                  /* device time of the dominant kernel of the last query call          */
    double  last_call_ms;                    /* device time of the whole last query call (all kernels, no copies)   */
    int64_t last_launches;                   /* kernels launched by the last call                                   */
    int64_t device_bytes;                    /* HBM held by the tree                                                */
    int64_t last_flagged;                    /* density calls: queries re-run by the exact-heap kernel (fp32 key ties) */
    int32_t warp_aligned;                    /* 1: built with NBK_WARP_ALIGNED                                     */
    int32_t reserved;
} nbk_info;

const char* nbk_last_error(void);
int nbk_device_count(void);

/* Replaces NBody::KDTree::KDTree(Particle*, numparts, bucket_size, TreeType, KernType, KernRes,
 * SplittingCriterion, Aniso, ScaleSpace, Period, ...) (reference KDTree.h:229-245, KDTree.cxx:1238-1306).
 * split must be 0 (KDTREE_SPLIT_SPREAD); period NULL => non periodic.  device < 0 => current device. */
int nbk_create(const nbk_particles* p, int64_t n, int bucket, int treetype, int kerntype, int kernres,
               int split, const double* period, int flags, int device, nbk_tree** out);
/* Replaces KDTree::~KDTree (KDTree.cxx:1340-1356); order restoration is the shim's job. */
int nbk_destroy(nbk_tree* t);
/* KDTree::GetNumNodes/GetNumLeafNodes/GetKernNorm/GetPeriod ... (KDTree.h:282-289) */
int nbk_get_info(const nbk_tree* t, nbk_info* info);
/* ids[i] = ID of the particle at tree index i (what the reference leaves in Particle::id after the
 * in-place reorder, KDTree.cxx:1291 + :328-370) */
int nbk_get_order(const nbk_tree* t, int32_t* ids, int flags);
/* Kernel table of KernelConstruction (KDTree.cxx:1144-1183): table[kernres] */
int nbk_get_kernel_table(const nbk_tree* t, double* table);
/* host mirror of the node arrays (KDTree::GetRoot / FindLeafNode consumers, KDTree.h:288,384-386):
 * per node (heap order, root 0, children 2i+1/2i+2) start,end (tree indices; start<0 => absent),
 * cut dimension (-1 leaf), bounds[2*nd] (lo0,hi0,lo1,...).  Pass NULL arrays to query *num_slots. */
int nbk_get_nodes(const nbk_tree* t, int64_t* num_slots, int32_t* start, int32_t* end, int32_t* cutdim, float* bounds);

/* Replaces the per-particle loops over KDTree::FindNearestPos(Int_t tt,...) / FindNearest(Int_t tt,...)
 * (KDFindNearest.cxx:247-334, whole-system forms :444-459) for tree indices [q0,q1).
 * nn: (q1-q0) x k tree indices (or IDs with NBK_OUT_IDS), d2: (q1-q0) x k, rows ascending.
 * Non periodic tree: self and coincident particles excluded (KDLeafNode.cxx:15-28).
 * Periodic tree: reference image schedule; self at slot 0 unless NBK_KNN_TREE_FORM (quirk Q3). */
int nbk_knn_particles(nbk_tree* t, int k, int64_t q0, int64_t q1, int32_t* nn, double* d2, int flags);
/* Replaces KDTree::FindNearestPos(Double_t *x,...) / (Coordinate x,...) (KDFindNearest.cxx:462-554)
 * for m query points x[m][3]. */
int nbk_knn_points(nbk_tree* t, int k, int64_t m, const double* x, int32_t* nn, double* d2, int flags);

/* Replace KDTree::FindNearestPhase(Int_t tt, nn, dist2, Nsearch) and FindNearestPhase(Double_t* x, Double_t* v, ...) /
 * (Coordinate x, Coordinate v, ...) (KDFindNearest.cxx:347-361,543-555; leaf code KDLeafNode.cxx:43-57,143-154) and with
 * them KDTree::FindNearest(tt | x, ...) on a TPHS tree built with Aniso = -1 (KDFindNearest.cxx:260-262,300-301: the same
 * search): the k nearest in the plain 6D phase-space distance PhaseDistSqd (DistFunc.h:41-49), positions and velocities in
 * the caller's units.  Tree indices [q0,q1) / m points x[m][3], v[m][3]; outputs like nbk_knn_particles.  TPHYS or TPHS
 * tree with velocities.  Non periodic particle form: the particle itself and 6D-coincident particles are not neighbours.
 * Periodic tree: the position is searched in all 8 images (KDSplitNode.cxx:1153-1187; velocities are never reflected); the
 * particle form returns the k nearest after the query itself.  The metric forms (Aniso >= 0, quirk Q4) are not built. */
int nbk_knn_phase_particles(nbk_tree* t, int k, int64_t q0, int64_t q1, int32_t* nn, double* d2, int flags);
int nbk_knn_phase_points(nbk_tree* t, int k, int64_t m, const double* x, const double* v, int32_t* nn, double* d2, int flags);

/* Replaces KDTree::FindNearestCheck(Int_t tt | Particle p | Coordinate x, check, params, nn, dist2, Nsearch) and
 * KDTree::FindNearestCriterion(Int_t tt | Particle p, cmp, params, nn, dist2, Nsearch) (KDFindNearest.cxx:363-441; leaf code
 * KDLeafNode.cxx:88-118,202-246): the k nearest among the particles i != target with 0 < d2 that pass the filters --
 *   check  (optional, n entries by ID, or tree order with NBK_TREE_ORDER): the caller's FOFcheckfunc values; only
 *          check == 0 particles can be neighbours;
 *   criterion (NBK_FOF3D / NBK_FOF6D, or -1 for none) with the reference's params[] (criterion parameters at [6], [7]):
 *          only particles with cmp(target, i, params) == 1 can be neighbours.  Point form: v = query velocities (FOF6d).
 * Rows ascending in d2; missing neighbours are (-1, 1e32) like the reference's sentinels.  Periodic trees: minimum image
 * over the reference's reflections.  The reference's periodic forms search Nsearch+1 and then drop the NEAREST entry
 * (KDFindNearest.cxx:365,376: LoadNN leaves the smallest of k+1 in the queue, although the target is never queued);
 * NBK_KNN_TREE_FORM reproduces that, without it the k nearest are returned. */
int nbk_knn_filtered_particles(nbk_tree* t, int k, int64_t q0, int64_t q1, int criterion, const double* params,
                               const int32_t* check, int32_t* nn, double* d2, int flags);
int nbk_knn_filtered_points(nbk_tree* t, int k, int64_t m, const double* x, const double* v, int criterion, const double* params,
                            const int32_t* check, int32_t* nn, double* d2, int flags);

/* Replaces KDTree::SearchBallPosTagged(Int_t tt / Double_t* x, fdist2, tagged) (KDFindNearest.cxx:618-688)
 * for a batch: CSR rows; offsets[m+1]; idx capacity cap; *total = entries required.
 * Rows hold every particle with d2 < fdist2 (strict; minimum image over the reference's reflections
 * when periodic).  Particle form (qidx = tree indices): the query itself is excluded when non periodic,
 * included when periodic (quirk Q5).  Rows are sorted ascending (periodic trees: ascending within each of the up to 8 image
 * passes, which are concatenated).  A batch whose rows hold 2^32 entries or more in all is refused (NBK_ERR_ARG): split it.
 * d2 (optional, cap entries, same layout as idx): squared position distance of every entry -- what the dense forms
 * KDTree::SearchBallPos(tt / x, fdist2, imark, nn, dist2) (KDFindNearest.cxx:567-587) write into dist2[ID]; the shim
 * builds the dense nn[] / dist2[] arrays from a row.  idx == NULL (or cap == 0): count pass only. */
int nbk_ball_particles(nbk_tree* t, double fdist2, int64_t m, const int32_t* qidx, int64_t* offsets,
                       int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags);
int nbk_ball_points(nbk_tree* t, double fdist2, int64_t m, const double* x, int64_t* offsets,
                    int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags);

/* Replaces KDTree::SearchCriterionTagged(Int_t tt | Particle& p, cmp, params, tagged) and the dense
 * KDTree::SearchCriterion(tt, cmp, params, imark, nn[, dist2]) (KDFindNearest.cxx:590-603,643-706; leaf code
 * KDLeafNode.cxx:414-492) for a batch and the in-tree criteria NBK_FOF3D / NBK_FOF6D: every particle i != target with
 * cmp(target, i, params) true, CSR like nbk_ball_*.  The reference prunes the tree walk with params[1] (position
 * distance^2) and evaluates cmp on every particle of a visited leaf; here pruning uses the radius the criterion itself
 * implies (params[6]), which gives the same set whenever params[1] >= params[6] (the way callers set it) and the full
 * criterion set otherwise.  Point form: v = query velocities (m x 3, needed by FOF6d, may be NULL for FOF3d);
 * nothing is excluded (a Particle that is not in the tree never compares equal to a tree particle). */
int nbk_search_criterion_particles(nbk_tree* t, int criterion, const double* params, int64_t m, const int32_t* qidx,
                                   int64_t* offsets, int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags);
int nbk_search_criterion_points(nbk_tree* t, int criterion, const double* params, int64_t m, const double* x, const double* v,
                                int64_t* offsets, int32_t* idx, double* d2, int64_t cap, int64_t* total, int flags);

/* Replaces KDTree::CalcDensity(Nsmooth) (KDCalcSmoothQuantities.cxx:203-305).  rho[n] by ID
 * (or tree index with NBK_TREE_ORDER); hsm (optional) = 0.5*sqrt(d2_k), the smoothing scale. */
int nbk_calc_density(nbk_tree* t, int nsmooth, double* rho, double* hsm, int flags);
/* Same, restricted to the query particles whose `active` byte (n entries, by ID or tree index like rho) is non-zero.
 * Inactive particles still act as neighbours and still receive the scatter term of the active ones; their own
 * gather+scatter contribution is left out.  This is what a slab-sharded caller needs: queries for owned particles
 * only, ghost particles as pure neighbours (nbodylib_b200/sharded.py). */
int nbk_calc_density_subset(nbk_tree* t, int nsmooth, const uint8_t* active, double* rho, double* hsm, int flags);
/* Replaces KDTree::CalcVelDensity(Nsmooth, Nsearch) (KDCalcSmoothQuantities.cxx:309-389). */
int nbk_calc_veldensity(nbk_tree* t, int nsmooth, int nsearch, double* rho, int flags);
/* "CalcSmoothingScale" (named by the north star; = hi of KDCalcSmoothQuantities.cxx:260). */
int nbk_smoothing_scale(nbk_tree* t, int nsmooth, double* hsm, int flags);

/* Replace the single-target forms KDTree::CalcDensityParticle(target, Nsmooth) / CalcVelDensityParticle(target, Nsmooth,
 * Nsearch) (KDCalcSmoothQuantities.cxx:768-921) and CalcDensityPosition(x, Nsmooth) / CalcVelDensityPosition(x, v, Nsmooth,
 * Nsearch) (:1092-1207) for a batch of m queries: gather-only sums (weight 1.0 * W, no scatter term, no 0.5 factor), one
 * value per query in rho[m].  Particle forms: qidx = tree indices (target search: the particle itself and coincident
 * particles are not neighbours); qidx == NULL means tree indices 0..m-1.  Point forms: coordinate search (d2 = 0 counts).
 * Like every Calc* call they ignore the period (quirk Q2). */
int nbk_calc_density_particles(nbk_tree* t, int nsmooth, int64_t m, const int32_t* qidx, double* rho, int flags);
int nbk_calc_veldensity_particles(nbk_tree* t, int nsmooth, int nsearch, int64_t m, const int32_t* qidx, double* rho, int flags);
int nbk_calc_density_points(nbk_tree* t, int nsmooth, int64_t m, const double* x, double* rho, int flags);
int nbk_calc_veldensity_points(nbk_tree* t, int nsmooth, int nsearch, int64_t m, const double* x, const double* v, double* rho, int flags);

/* Replace KDTree::CalcSmoothVel(Nsmooth, densityset) and KDTree::CalcSmoothVelDisp(smvel, Nsmooth, densityset, meanvelset)
 * (KDCalcSmoothQuantities.cxx:480-614): kernel-smoothed mean velocity (n x 3) and velocity dispersion tensor (n x 9, row-major
 * 3x3) of every particle, symmetric gather + scatter like CalcDensity with weights 0.5 W(r_ij, h_i) m / rho of the
 * contributing particle; the dispersion is taken about the receiving particle's smoothed mean velocity.
 * rho: the particles' densities (n, what CalcDensity left in Particle::rho); NULL = compute CalcDensity(nsmooth) first (the
 * reference's densityset != 1).  smvel: the result of nbk_calc_smooth_vel.  All arrays by ID (tree order with NBK_TREE_ORDER). */
int nbk_calc_smooth_vel(nbk_tree* t, int nsmooth, const double* rho, double* smvel, int flags);
int nbk_calc_smooth_veldisp(nbk_tree* t, int nsmooth, const double* rho, const double* smvel, double* smveldisp, int flags);
/* Replace KDTree::CalcSmoothVelSkew(smvel, smveldisp, Nsmooth, ...) and KDTree::CalcSmoothVelKurtosis(...)
 * (KDCalcSmoothQuantities.cxx:617-765): per velocity component k, the kernel-smoothed third / fourth power of (v_k - smoothed mean
 * of the receiving particle) in units of the receiver's dispersion, sigma_kk^1.5 / sigma_kk^2 (n x 3); same symmetric gather +
 * scatter and weights as CalcSmoothVelDisp.  The reference subtracts 3 from EVERY kurtosis contribution (2 Nsmooth per particle
 * on average) rather than once; that is reproduced.  smvel / smveldisp: results of the two calls above. */
int nbk_calc_smooth_velskew(nbk_tree* t, int nsmooth, const double* rho, const double* smvel, const double* smveldisp, double* smvelskew, int flags);
int nbk_calc_smooth_velkurtosis(nbk_tree* t, int nsmooth, const double* rho, const double* smvel, const double* smveldisp, double* smvelkurt, int flags);

/* Optional FOF by-products in tree-index space (reference KDFOF.cxx:52-55: pHead,pNext,pTail,pLen).
 * Any pointer may be NULL.  head/next/tail have n entries: the members of a group are chained in ascending tree
 * index (head[i] = first member, next[i] = following member or -1, tail[i] = last member; a particle outside any
 * group is its own one-element list).  len needs n+1 entries of room; entries 0..ngroups are written (index = group
 * id, the reference's pLen[iGroup]). */
typedef struct { int32_t* head; int32_t* next; int32_t* tail; int32_t* len; } nbk_fof_lists;

/* Replaces KDTree::FOF(fdist, numgroup, minnum, order, pHead,pNext,pTail,pLen, ipcheckflag, check, params)
 * (KDFOF.cxx:29-153).  On a TPHS tree the link distance is 6D (KDLeafNode.cxx:570-572).
 * precheck (optional, n entries by ID): non-zero entries are excluded from linking exactly like a
 * FOFcheckfunc returning non-zero (KDFOF.cxx:65,72); group[n] by ID (or tree index). */
int nbk_fof(nbk_tree* t, double fdist, int minnum, int order, const int32_t* precheck, int32_t* group,
            int64_t* ngroups, nbk_fof_lists* lists, int flags);
/* Replaces KDTree::FOFCriterion(cmp, params, numgroups, minnum, order, ...) (KDFOF.cxx:157-265) for the
 * in-tree criteria; params laid out as the reference expects (FOFFunc.h:8-15: [0] tree type, [1],[2]
 * pos/vel pruning length^2, [6],[7] criterion parameters). */
int nbk_fof_criterion(nbk_tree* t, int criterion, const double* params, int minnum, int order,
                      const int32_t* precheck, int32_t* group, int64_t* ngroups, nbk_fof_lists* lists, int flags);

/* Replaces KDTree::FOFCriterionSetBasisForLinks(cmp, params, numgroup, minnum, order, ipcheckflag, check, ...)
 * (KDFOF.cxx:268-378, KDLeafNode.cxx:620-652).  check (n entries by ID, the values of the caller's FOFcheckfunc): only
 * particles with check == 0 start or extend groups; the others can be linked into a group but never link further.
 * The groups of the check == 0 particles are the connected components of their mutual links (identical to the
 * reference).  A particle with check != 0 that is linked by members of several groups joins the group whose first
 * member comes first in tree order -- the group the reference's serial search reaches it from -- so the assignment
 * can differ from the reference only where the two tree orders differ inside a leaf. */
int nbk_fof_criterion_basis(nbk_tree* t, int criterion, const double* params, int minnum, int order,
                            const int32_t* check, int32_t* group, int64_t* ngroups, nbk_fof_lists* lists, int flags);

/* Domain decomposition support (no counterpart in the reference, which has no distributed code: VELOCIraptor builds one
 * KDTree per MPI rank over local + imported particles).  Appends the particles of `halo` (a second TPHYS tree on the same
 * device, e.g. the ghost layer received from the neighbouring ranks) to `t` as a SECOND tree: IDs of the halo particles
 * follow t's (n_main .. n_main + n_halo - 1), `halo` is consumed.  nbk_calc_density / nbk_calc_veldensity /
 * nbk_smoothing_scale then run their queries for t's own particles only and search both trees; output arrays have
 * n_main + n_halo entries (the symmetric scatter term of CalcDensity is deposited on halo particles too, for the caller
 * to send home).  Keeping the halo out of the main tree leaves the main tree's shape -- and the alignment of the warp
 * query groups with its nodes -- independent of the halo size.  Other queries return NBK_ERR_UNSUPPORTED on such a tree. */
int nbk_attach_halo(nbk_tree* t, nbk_tree* halo);

/* Scratch buffers are recycled through the device's stream-ordered memory pool and kept across calls and
 * trees (allocating and freeing GBs through the driver costs more than the kernels).  This hands the cached
 * memory back to the driver (e.g. before another library needs the HBM). */
int nbk_release_cached_memory(int device);

/* Building blocks of a domain-decomposed FOF (VELOCIraptor stitches its MPI domains outside the library; here
 * the slab-sharded driver does): the components of FOF (criterion < 0: KDTree::FOF(fdist), KDFOF.cxx:29-153) or FOFCriterion
 * (NBK_FOF3D / NBK_FOF6D with the reference's params[], KDFOF.cxx:157-265) WITHOUT the minnum filter and the numbering:
 * root[ID] = ID of one fixed member of the particle's component (the same for all its members), -1 for particles excluded
 * by precheck.  NBK_DEVICE_PTRS: precheck / root are device pointers. */
int nbk_fof_roots(nbk_tree* t, int criterion, double fdist, const double* params, const int32_t* precheck, int32_t* root, int flags);
/* Connected components of an explicit edge list on the device (the cross-domain merge step): nodes 0..nnodes-1, edges
 * (a[i], b[i]); root[v] = smallest node of v's component.  a, b, root are DEVICE pointers on `device` (-1: the current one). */
int nbk_union_pairs(int device, int64_t nnodes, int64_t npairs, const int32_t* a, const int32_t* b, int32_t* root);

/* Process-wide tuning overrides (none is needed for correct results; the defaults are what the benchmarks run).  No
 * reference counterpart: the reference's only knobs are constructor arguments.  Names:
 *   "knn_leaf"    particles per scanned tile of the density kernel (0 = the tree level holding 21..40 particles)
 *   "knn_exact"   1: the Calc* family runs on the fp64-heap kernel only
 *   "knn_transpose" density kernel: tiles needed by at most this many of a warp's 32 queries are screened query-by-query
 *                 (-1 = default 12, 0 = never)
 *   "fof_screen"  0: the 3D link kernel skips its fp32 screen
 * Unknown names return NBK_ERR_ARG. */
int nbk_set_option(const char* name, int64_t value);

/* Device-resident views for callers that stay on the GPU (sharded driver, benchmarks). */
int nbk_device_arrays(const nbk_tree* t, const void** pos4, const void** vel4, const void** mass, const int32_t** order);

#ifdef __cplusplus
}
#endif
#endif /* NBK_H */
