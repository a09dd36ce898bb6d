/* nbk_sharded.h -- C ABI of the slab-sharded tree: one process per GPU, NCCL over NVLink for the exchange steps.
 *
 * Library: libnbk_sharded.so (links libnbk.so and NCCL; libnbk.so itself has no NCCL dependency).
 *
 * The reference (pelahi/NBodylib) has no distributed code: its user VELOCIraptor decomposes the domain over MPI ranks
 * outside the library, imports a ghost layer and builds one local NBody::KDTree per rank.  These entry points
 * are that outer layer for the GPUs of one node, for the two calls of the hot path that need an exchange step:
 *   KDTree::CalcDensity(Nsmooth)                          (KDCalcSmoothQuantities.cxx:203-305)
 *   KDTree::FOF(fdist, ...) / KDTree::FOFCriterion(...)   (KDFOF.cxx:29-153, :157-265)
 * Results are those of ONE tree over all ranks' particles: densities to rounding (the order of the fp64 sums differs), FOF
 * partitions identical, group ids one numbering for the whole job.
 *
 * Decomposition: the global box [0,box[0]) x [0,box[1]) x [0,box[2]) is cut into `nranks` slabs along x with faces at
 * slab_edges[0] = 0 < slab_edges[1] < ... < slab_edges[nranks] = box[0] (NULL: equal widths); rank r passes the particles with
 * slab_edges[r] <= x < slab_edges[r+1], in GLOBAL coordinates.  The caller owns the decomposition, like VELOCIraptor owns its MPI
 * domains: equal widths, or faces at the x quantiles for equal particle counts.  Global particle id = position in the
 * concatenation of the ranks' arrays in rank order.
 *
 * Every call is COLLECTIVE: all ranks of the communicator make it, with the same arguments apart from the particle data.
 * Errors: nbk_status codes of nbk.h, text from nbk_last_error().  A rank that fails leaves its peers waiting in NCCL, like
 * any collective; argument errors that every rank detects identically (halo wider than a slab, bad criterion) are raised
 * before any exchange.
 */
#ifndef NBK_SHARDED_H
#define NBK_SHARDED_H

#include "nbk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbk_comm nbk_comm;
typedef struct nbk_sharded nbk_sharded;

/* Rendezvous, NCCL style: ONE rank calls nbk_comm_unique_id and hands the 128 bytes to the others by whatever channel the
 * host program has (MPI_Bcast in an MPI code, a file, torch.distributed in the tests); then every rank calls
 * nbk_comm_init_rank(nranks, its rank, the id, its device [-1: the current one]).  nranks == 1 needs no id. */
int nbk_comm_unique_id(unsigned char id[128]);
int nbk_comm_init_rank(int nranks, int rank, const unsigned char id[128], int device, nbk_comm** out);
int nbk_comm_destroy(nbk_comm* c);

/* This rank's slab.  p: its particles (host or device, any stride; copied -- the caller's arrays are not kept), n_local >= 1.
 * periodic != 0: FOF wraps with the periods box[] (CalcDensity never does: the reference's Calc* family ignores the period,
 * SURVEY.md quirk Q2).  knn_k: the Nsmooth the density halo is sized for (<= 0: 64); halo > 0 overrides the initial halo
 * width (it is widened automatically until every owned k-ball is complete).  With three or more ranks a halo must stay below the
 * narrowest slab's width (particles two slabs away would be missing): such calls fail with NBK_ERR_ARG on every rank.  Mirrors NBody::KDTree's constructor for
 * TPHYS / KEPAN / bucket 16, the configuration of the headline workload. */
int nbk_sharded_create(nbk_comm* c, const nbk_particles* p, int64_t n_local, const double box[3], const double* slab_edges,
                       int periodic, int knn_k, double halo, nbk_sharded** out);
int nbk_sharded_destroy(nbk_sharded* s);

/* KDTree::CalcDensity(nsmooth) over the global particle set; rho[n_local] for this rank's particles in input order
 * (device pointer with NBK_DEVICE_PTRS). */
int nbk_sharded_calc_density(nbk_sharded* s, int nsmooth, double* rho, int flags);

/* criterion < 0: KDTree::FOF(fdist, numgroups, minnum, order); criterion NBK_FOF3D / NBK_FOF6D: KDTree::FOFCriterion(cmp,
 * params, numgroups, minnum, order) with the reference's params[] (FOFFunc.h:8-15).  group[n_local]: global group id of
 * this rank's particles (0 = none); *ngroups = number of groups of the whole job (the same on every rank).  order != 0:
 * ids descend in group size (stable); otherwise groups that cross a slab face come first, then each rank's interior groups.
 * The slab's local tree (owned + ghost layer of the linking length) stays resident between calls with the same reach. */
int nbk_sharded_fof(nbk_sharded* s, int criterion, double fdist, const double* params, int minnum, int order, int32_t* group,
                    int64_t* ngroups, int flags);

typedef struct {
    int64_t n_local, n_global, first_global_id;
    int32_t rank, nranks;
    double  h_knn;                 /* density halo width in use (one number for the group)                  */
    int64_t ghosts_knn, ghosts_fof;/* ghost particles held for the last density / FOF call                   */
    int64_t density_setups, fof_setups; /* halo exchanges + tree builds so far (1 = the state stayed resident) */
    double  last_kernel_ms;        /* device time of the dominant kernel of the last call on this rank       */
    double  last_call_ms;          /* device time of the library call (density / FOF roots) inside the last call */
    int64_t last_launches;         /* kernels that library call launched (the exchange layer's own small kernels not counted) */
    int64_t last_flagged;          /* density: queries re-run by the exact-heap kernel                       */
} nbk_sharded_info;
int nbk_sharded_get_info(const nbk_sharded* s, nbk_sharded_info* info);
/* Drops the resident slab trees (density tree + halo, FOF tree); the next call rebuilds what it needs.  Not collective. */
int nbk_sharded_release(nbk_sharded* s);

#ifdef __cplusplus
}
#endif
#endif
