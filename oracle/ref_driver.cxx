/* oracle/ref_driver.cxx -- TEST INFRASTRUCTURE ONLY.
 *
 * A C-ABI "dumper" around the UNMODIFIED reference classes (NBody::KDTree, NBody::Particle),
 * compiled together with the reference sources where they lie under /root/reference by
 * oracle/Makefile into oracle/_ref/libnbref.so.  Python tests / bench.py (--impl reference and the
 * cpu_baseline leg) drive it through ctypes (nbodylib_b200/_oracle.py is NOT used by the product
 * path; see tests/refdriver.py).
 *
 * Every entry point converts the reference's tree-order indices into particle IDs (= input order,
 * reference KDTree.cxx:1291) so results can be compared with any other tree layout.
 *
 * Reference entry points driven here:
 *   KDTree ctor               src/KDTree/KDTree.cxx:1238-1306
 *   FindNearest / Pos (tt)    src/KDTree/KDFindNearest.cxx:247-334   (looped as tests/test_kdtree.cxx:279-301)
 *   FindNearestPos (x)        src/KDTree/KDFindNearest.cxx:462-554
 *   SearchBallPosTagged       src/KDTree/KDFindNearest.cxx:618-688
 *   CalcDensity/CalcVelDensity src/KDTree/KDCalcSmoothQuantities.cxx:203-389
 *   FOF / FOFCriterion        src/KDTree/KDFOF.cxx:29-265
 *   CalcDensityParticle / CalcVelDensityParticle / Calc*Position / CalcSmoothLocalValue
 *                             src/KDTree/KDCalcSmoothQuantities.cxx:768-921, 1092-1207, 1704-1735
 *   SearchCriterionTagged, dense SearchBallPos / SearchCriterion   src/KDTree/KDFindNearest.cxx:567-603, 660-706
 *   FindLeafNode, node getters   src/KDTree/KDFindNearest.cxx:709-736, KDNode.h
 */
#include <KDTree.h>
#include <omp.h>
#include <chrono>
#include <cstring>
#include <vector>

using namespace NBody;
using namespace std;

struct RefTree {
    Particle* parts;
    Int_t n;
    KDTree* tree;
    double period[3];
    bool periodic;
    double build_seconds;
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" {

int ref_sizeof_particle() { return (int)sizeof(Particle); }
int ref_max_threads() { return omp_get_max_threads(); }

/* pos, vel: n x 3 doubles (row-major); mass: n doubles or NULL (=1). period NULL => non periodic. */
void* ref_create(long n, const double* pos, const double* vel, const double* mass, int bucket, int treetype,
                 int kerntype, int kernres, const double* period, int aniso) {
    RefTree* h = new RefTree;
    h->n = (Int_t)n;
    h->parts = new Particle[n];
    for (long i = 0; i < n; i++) {
        h->parts[i].SetMass(mass ? mass[i] : 1.0);
        h->parts[i].SetPosition(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        if (vel) h->parts[i].SetVelocity(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        else h->parts[i].SetVelocity(0, 0, 0);
        h->parts[i].SetID(i);
        h->parts[i].SetPID(i);
        h->parts[i].SetType(0);
    }
    h->periodic = (period != NULL);
    if (period) for (int j = 0; j < 3; j++) h->period[j] = period[j];
    double t0 = now_s();
    h->tree = new KDTree(h->parts, h->n, bucket, treetype, kerntype, kernres, 0, aniso, 0,
                         period ? h->period : NULL);
    h->build_seconds = now_s() - t0;
    return h;
}

void ref_destroy(void* hv) {
    RefTree* h = (RefTree*)hv;
    delete h->tree;
    delete[] h->parts;
    delete h;
}

double ref_build_seconds(void* hv) { return ((RefTree*)hv)->build_seconds; }
long ref_num_nodes(void* hv) { return ((RefTree*)hv)->tree->GetNumNodes(); }
long ref_num_leaves(void* hv) { return ((RefTree*)hv)->tree->GetNumLeafNodes(); }
double ref_kernnorm(void* hv) { return ((RefTree*)hv)->tree->GetKernNorm(); }

/* ids[i] = particle ID sitting at tree index i */
void ref_order(void* hv, int* ids) {
    RefTree* h = (RefTree*)hv;
    for (Int_t i = 0; i < h->n; i++) ids[i] = (int)h->parts[i].GetID();
}

/* kNN for queries q0..q1 (particle IDs, input order).  which: 0 = FindNearestPos(tt), 1 = FindNearest(tt), 2 = FindNearestVel(tt).
 * out_ids / out_d2 : (q1-q0) x k, neighbour particle IDs.  Returns wall seconds of the query loop. */
double ref_knn_particles(void* hv, int which, int k, long q0, long q1, int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
    double t0 = now_s();
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long q = q0; q < q1; q++) {
            Int_t tt = where[q];
            if (which == 0) h->tree->FindNearestPos(tt, nn.data(), d2.data(), k);
            else if (which == 2) h->tree->FindNearestVel(tt, nn.data(), d2.data(), k);
            else h->tree->FindNearest(tt, nn.data(), d2.data(), k);
            if (out_ids) {
                for (int j = 0; j < k; j++) {
                    out_ids[(q - q0) * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                    out_d2[(q - q0) * k + j] = d2[j];
                }
            }
        }
    }
    return now_s() - t0;
}

/* the same for an explicit list of query IDs (sampled parity checks at sizes where the whole system is too much) */
double ref_knn_particle_list(void* hv, int which, int k, long m, const int* qids, int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
    double t0 = now_s();
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long q = 0; q < m; q++) {
            Int_t tt = where[qids[q]];
            if (which == 0) h->tree->FindNearestPos(tt, nn.data(), d2.data(), k);
            else if (which == 2) h->tree->FindNearestVel(tt, nn.data(), d2.data(), k);
            else h->tree->FindNearest(tt, nn.data(), d2.data(), k);
            for (int j = 0; j < k; j++) {
                out_ids[q * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                out_d2[q * k + j] = d2[j];
            }
        }
    }
    return now_s() - t0;
}

void ref_set_threads(int n) { omp_set_num_threads(n); }

/* kNN around arbitrary positions (m x 3 doubles), FindNearestPos(Double_t*) */
double ref_knn_points(void* hv, int k, long m, const double* x, int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
    double t0 = now_s();
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long q = 0; q < m; q++) {
            Double_t xx[3] = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
            h->tree->FindNearestPos(xx, nn.data(), d2.data(), k);
            for (int j = 0; j < k; j++) {
                out_ids[q * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                out_d2[q * k + j] = d2[j];
            }
        }
    }
    return now_s() - t0;
}

/* SearchBallPosTagged(tt, r2) for the given query IDs; CSR output (offsets has m+1 entries);
 * `cap` is the capacity of out_ids; returns total count (may exceed cap => truncated). */
long ref_ball_particles(void* hv, double r2, long m, const int* qids, long* offsets, int* out_ids, long cap) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
    long tot = 0;
    vector<Int_t> tagged(h->n);
    for (long q = 0; q < m; q++) {
        offsets[q] = tot;
        Int_t nt = h->tree->SearchBallPosTagged(where[qids[q]], r2, tagged.data());
        for (Int_t j = 0; j < nt; j++) {
            if (tot < cap) out_ids[tot] = (int)h->parts[tagged[j]].GetID();
            tot++;
        }
    }
    offsets[m] = tot;
    return tot;
}

long ref_ball_points(void* hv, double r2, long m, const double* x, long* offsets, int* out_ids, long cap) {
    RefTree* h = (RefTree*)hv;
    long tot = 0;
    vector<Int_t> tagged(h->n);
    for (long q = 0; q < m; q++) {
        offsets[q] = tot;
        Double_t xx[3] = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
        Int_t nt = h->tree->SearchBallPosTagged(xx, r2, tagged.data());
        for (Int_t j = 0; j < nt; j++) {
            if (tot < cap) out_ids[tot] = (int)h->parts[tagged[j]].GetID();
            tot++;
        }
    }
    offsets[m] = tot;
    return tot;
}

/* library CalcDensity (serial); rho_by_id[n] */
double ref_calc_density(void* hv, int k, double* rho_by_id) {
    RefTree* h = (RefTree*)hv;
    double t0 = now_s();
    h->tree->CalcDensity(k);
    double dt = now_s() - t0;
    for (Int_t i = 0; i < h->n; i++) rho_by_id[h->parts[i].GetID()] = h->parts[i].GetDensity();
    return dt;
}

double ref_calc_veldensity(void* hv, int kv, int kx, double* rho_by_id) {
    RefTree* h = (RefTree*)hv;
    double t0 = now_s();
    h->tree->CalcVelDensity(kv, kx);
    double dt = now_s() - t0;
    for (Int_t i = 0; i < h->n; i++) rho_by_id[h->parts[i].GetID()] = h->parts[i].GetDensity();
    return dt;
}

/* "full-host OpenMP kNN-density" (variant ii): the caller-side OpenMP loop over
 * FindNearestPos-equivalent searches of tests/test_kdtree.cxx:279-301 plus the R1 accumulation of
 * KDCalcSmoothQuantities.cxx:260-300 with atomics on the scatter term.  Uses only public reference
 * methods; the kernel table is rebuilt here exactly as KDTree.cxx:1144-1183 (KEPAN / KSPH only).
 * Queries the tree-order range [i0,i1) so a bounded sample can be timed.  Non-periodic search (quirk Q2). */
double ref_calc_density_omp(void* hv, int k, long i0, long i1, int kernres, int kerntype, double* rho_by_id) {
    RefTree* h = (RefTree*)hv;
    const int ND = 3;
    vector<Double_t> Kernel(kernres);
    double kernnorm = h->tree->GetKernNorm();
    double delta = 2.0 / (Double_t)(kernres - 1);
    for (int i = 0; i < kernres; i++) {
        double r = i * delta;
        Kernel[i] = kernnorm * (kerntype == KDTree::KSPH ? WSPH(r, 1.0) : (kerntype == KDTree::KEPAN ? WEpan(r, 1.0)
                              : (kerntype == KDTree::KGAUSS ? WGauss(r, 1.0) : WTH(r, 1.0))));
    }
    vector<double> rho(h->n, 0.0);
    /* the periodic tree's FindNearestPos(tt) would do a periodic search; CalcDensity never does (Q2).
     * A non-periodic per-particle kNN through the public API needs a tree without a period, so the
     * caller must have created this handle with period=NULL. */
    double t0 = now_s();
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long i = i0; i < i1; i++) {
            h->tree->FindNearestPos((Int_t)i, nn.data(), d2.data(), k);
            Double_t hi = 0.5 * sqrt(d2[k - 1]);
            Double_t norm = 1.0 / pow(hi, (Double_t)(ND * 1.));
            Double_t mi = h->parts[i].GetMass(), acc = 0;
            for (int j = k - 1; j >= 0; j--) {
                Double_t rij = sqrt(d2[j]);
                Double_t r = rij / hi;
                int idx = (int)(r * 0.5 * (kernres - 1));
                Double_t W = (idx < kernres - 1) ? (Kernel[idx] + (Kernel[idx + 1] - Kernel[idx]) * (r - delta * idx) / delta)
                                                 : Kernel[idx];
                Double_t Wij = 0.5 * W * norm;
                acc += Wij * h->parts[nn[j]].GetMass();
                Double_t add = Wij * mi;
#pragma omp atomic
                rho[nn[j]] += add;
            }
#pragma omp atomic
            rho[i] += acc;
        }
    }
    double dt = now_s() - t0;
    if (rho_by_id) for (Int_t i = 0; i < h->n; i++) rho_by_id[h->parts[i].GetID()] = rho[i];
    return dt;
}

/* FOF(fdist) -> group_by_id[n]; returns seconds; *ngroups set */
double ref_fof(void* hv, double fdist, int minnum, int order, int* group_by_id, long* ngroups) {
    RefTree* h = (RefTree*)hv;
    Int_t ng = 0;
    double t0 = now_s();
    Int_t* g = h->tree->FOF(fdist, ng, minnum, order);
    double dt = now_s() - t0;
    for (Int_t i = 0; i < h->n; i++) group_by_id[i] = (int)g[i];
    delete[] g;
    *ngroups = ng;
    return dt;
}

/* FOFCriterion with one of the in-tree criteria: 0 = FOF3d, 1 = FOFVel, 2 = FOF6d (FOFFunc.h:30-55).
 * params must hold >= 8 doubles laid out as the reference expects. */
double ref_fof_criterion(void* hv, int crit, double* params, int minnum, int order, int* group_by_id, long* ngroups) {
    RefTree* h = (RefTree*)hv;
    Int_t ng = 0;
    FOFcompfunc cmp = crit == 0 ? FOF3d : (crit == 1 ? FOFVel : FOF6d);
    double t0 = now_s();
    Int_t* g = h->tree->FOFCriterion(cmp, params, ng, minnum, order);
    double dt = now_s() - t0;
    for (Int_t i = 0; i < h->n; i++) group_by_id[i] = (int)g[i];
    delete[] g;
    *ngroups = ng;
    return dt;
}

/* check function for the ipcheckflag / SetBasisForLinks entry points: particles whose type is non-zero are "checked out" */
static int check_by_type(Particle& p, Double_t*) { return p.GetType() != 0 ? -1 : 0; }

/* set Particle::type from an array indexed by ID (the tree has permuted the array) */
void ref_set_types(void* hv, const int* type_by_id) {
    RefTree* h = (RefTree*)hv;
    for (Int_t i = 0; i < h->n; i++) h->parts[i].SetType(type_by_id[h->parts[i].GetID()]);
}

/* which: 0 = FOF(fdist = params[0]) with ipcheckflag, 1 = FOFCriterion(cmp) with ipcheckflag,
 *        2 = FOFCriterionSetBasisForLinks(cmp); crit as in ref_fof_criterion */
double ref_fof_checked(void* hv, int which, int crit, double* params, int minnum, int order, int* group_by_id, long* ngroups) {
    RefTree* h = (RefTree*)hv;
    Int_t ng = 0;
    FOFcompfunc cmp = crit == 0 ? FOF3d : (crit == 1 ? FOFVel : FOF6d);
    double t0 = now_s();
    Int_t* g;
    if (which == 0) g = h->tree->FOF(params[0], ng, minnum, order, NULL, NULL, NULL, NULL, 1, check_by_type, params);
    else if (which == 1) g = h->tree->FOFCriterion(cmp, params, ng, minnum, order, 1, check_by_type);
    else g = h->tree->FOFCriterionSetBasisForLinks(cmp, params, ng, minnum, order, 1, check_by_type);
    double dt = now_s() - t0;
    for (Int_t i = 0; i < h->n; i++) group_by_id[i] = (int)g[i];
    delete[] g;
    *ngroups = ng;
    return dt;
}

/* Particle::ScalePhase on the whole array (Particle.h:666) -- used for the 6D FOF form (A) */
void ref_scale_phase(long n, double* pos, double* vel, double xs, double vs) {
    Particle p;
    for (long i = 0; i < n; i++) {
        p.SetPosition(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        p.SetVelocity(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        Double_t a = xs, b = vs;
        p.ScalePhase(a, b);
        for (int j = 0; j < 3; j++) { pos[3 * i + j] = p.GetPosition(j); vel[3 * i + j] = p.GetVelocity(j); }
    }
}

/* per-node dump for structural comparisons: walks from the root (KDNode.h getters). */
static void walk(Node* nd, vector<int>& starts, vector<int>& ends, vector<int>& leaf, vector<int>& cutdim) {
    starts.push_back((int)nd->GetStart());
    ends.push_back((int)nd->GetEnd());
    leaf.push_back(nd->GetLeaf() ? 1 : 0);
    if (!nd->GetLeaf()) {
        cutdim.push_back(((SplitNode*)nd)->GetCutDim());
        walk(((SplitNode*)nd)->GetLeft(), starts, ends, leaf, cutdim);
        walk(((SplitNode*)nd)->GetRight(), starts, ends, leaf, cutdim);
    } else cutdim.push_back(-1);
}
long ref_dump_nodes(void* hv, int* start, int* end, int* isleaf, int* cutdim, long cap) {
    RefTree* h = (RefTree*)hv;
    vector<int> s, e, l, c;
    walk(h->tree->GetRoot(), s, e, l, c);
    long m = (long)s.size();
    for (long i = 0; i < m && i < cap; i++) { start[i] = s[i]; end[i] = e[i]; isleaf[i] = l[i]; cutdim[i] = c[i]; }
    return m;
}

/* ---- single-target estimators (KDCalcSmoothQuantities.cxx:768-921, 1092-1207), looped over query IDs / points ---- */
void ref_calc_density_particles(void* hv, int k, long m, const int* qids, double* out) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
#pragma omp parallel for schedule(guided)
    for (long q = 0; q < m; q++) out[q] = h->tree->CalcDensityParticle(where[qids[q]], k);
}
void ref_calc_veldensity_particles(void* hv, int kv, int kx, long m, const int* qids, double* out) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
#pragma omp parallel for schedule(guided)
    for (long q = 0; q < m; q++) out[q] = h->tree->CalcVelDensityParticle(where[qids[q]], kv, kx);
}
void ref_calc_density_points(void* hv, int k, long m, const double* x, double* out) {
    RefTree* h = (RefTree*)hv;
#pragma omp parallel for schedule(guided)
    for (long q = 0; q < m; q++) {
        Double_t xx[3] = {x[3 * q], x[3 * q + 1], x[3 * q + 2]};
        out[q] = h->tree->CalcDensityPosition(xx, k);
    }
}
void ref_calc_veldensity_points(void* hv, int kv, int kx, long m, const double* x, const double* v, double* out) {
    RefTree* h = (RefTree*)hv;
#pragma omp parallel for schedule(guided)
    for (long q = 0; q < m; q++) {
        Double_t xx[3] = {x[3 * q], x[3 * q + 1], x[3 * q + 2]}, vv[3] = {v[3 * q], v[3 * q + 1], v[3 * q + 2]};
        out[q] = h->tree->CalcVelDensityPosition(xx, vv, kv, kx);
    }
}
/* CalcSmoothLocalValue(Nsmooth, Double_t* dist, Double_t* weight): dist descending (KDCalcSmoothQuantities.cxx:1721-1735) */
double ref_smooth_local_value(void* hv, int k, double* dist, double* weight) {
    RefTree* h = (RefTree*)hv;
    return h->tree->CalcSmoothLocalValue(k, dist, weight);
}

/* SearchCriterionTagged(tt, cmp, params, tagged) for the given query IDs; CSR like ref_ball_particles */
long ref_search_criterion_particles(void* hv, int crit, double* params, long m, const int* qids, long* offsets, int* out_ids, long cap) {
    RefTree* h = (RefTree*)hv;
    FOFcompfunc cmp = crit == 0 ? FOF3d : (crit == 1 ? FOFVel : FOF6d);
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
    long tot = 0;
    vector<Int_t> tagged(8 * (size_t)h->n + 8);       // the periodic form can report a particle once per image
    for (long q = 0; q < m; q++) {
        offsets[q] = tot;
        Int_t nt = h->tree->SearchCriterionTagged(where[qids[q]], cmp, params, tagged.data());
        for (Int_t j = 0; j < nt; j++) {
            if (tot < cap) out_ids[tot] = (int)h->parts[tagged[j]].GetID();
            tot++;
        }
    }
    offsets[m] = tot;
    return tot;
}
/* SearchCriterionTagged(Particle& p, ...) for particles that are not in the tree (positions x, velocities v) */
long ref_search_criterion_points(void* hv, int crit, double* params, long m, const double* x, const double* v, long* offsets, int* out_ids, long cap) {
    RefTree* h = (RefTree*)hv;
    FOFcompfunc cmp = crit == 0 ? FOF3d : (crit == 1 ? FOFVel : FOF6d);
    long tot = 0;
    vector<Int_t> tagged(8 * (size_t)h->n + 8);
    for (long q = 0; q < m; q++) {
        offsets[q] = tot;
        Particle p;
        p.SetPosition(x[3 * q], x[3 * q + 1], x[3 * q + 2]);
        if (v) p.SetVelocity(v[3 * q], v[3 * q + 1], v[3 * q + 2]);
        p.SetID(-1); p.SetPID(-1);
        Int_t nt = h->tree->SearchCriterionTagged(p, cmp, params, tagged.data());
        for (Int_t j = 0; j < nt; j++) {
            if (tot < cap) out_ids[tot] = (int)h->parts[tagged[j]].GetID();
            tot++;
        }
    }
    offsets[m] = tot;
    return tot;
}
/* dense SearchBallPos(tt | x, fdist2, imark, nn, dist2) (KDFindNearest.cxx:567-587): nn / dist2 have n entries (by ID)
 * and are updated in place; qid < 0 selects the position form. */
void ref_search_ball_dense(void* hv, long qid, const double* x, double r2, int imark, int* nn_by_id, double* d2_by_id) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> nn(nn_by_id, nn_by_id + h->n);
    vector<Double_t> d2(d2_by_id, d2_by_id + h->n);
    if (qid >= 0) {
        Int_t tt = -1;
        for (Int_t i = 0; i < h->n; i++) if (h->parts[i].GetID() == qid) { tt = i; break; }
        h->tree->SearchBallPos(tt, r2, imark, nn.data(), d2.data());
    } else {
        Double_t xx[3] = {x[0], x[1], x[2]};
        h->tree->SearchBallPos(xx, r2, imark, nn.data(), d2.data());
    }
    for (Int_t i = 0; i < h->n; i++) { nn_by_id[i] = (int)nn[i]; d2_by_id[i] = d2[i]; }
}
/* dense SearchCriterion(tt, cmp, params, imark, nn, dist2) (KDFindNearest.cxx:590-596) */
void ref_search_criterion_dense(void* hv, int crit, double* params, long qid, int imark, int* nn_by_id, double* d2_by_id) {
    RefTree* h = (RefTree*)hv;
    FOFcompfunc cmp = crit == 0 ? FOF3d : (crit == 1 ? FOFVel : FOF6d);
    vector<Int_t> nn(nn_by_id, nn_by_id + h->n);
    vector<Double_t> d2(d2_by_id, d2_by_id + h->n);
    Int_t tt = -1;
    for (Int_t i = 0; i < h->n; i++) if (h->parts[i].GetID() == qid) { tt = i; break; }
    h->tree->SearchCriterion(tt, cmp, params, imark, nn.data(), d2.data());
    for (Int_t i = 0; i < h->n; i++) { nn_by_id[i] = (int)nn[i]; d2_by_id[i] = d2[i]; }
}
/* FindLeafNode(tt) / FindLeafNode(x): the leaf's particle IDs (sorted by the caller), count returned */
long ref_find_leaf(void* hv, long qid, const double* x, int* member_ids, long cap) {
    RefTree* h = (RefTree*)hv;
    Node* nd;
    if (qid >= 0) {
        Int_t tt = -1;
        for (Int_t i = 0; i < h->n; i++) if (h->parts[i].GetID() == qid) { tt = i; break; }
        nd = h->tree->FindLeafNode(tt);
    } else {
        Double_t xx[3] = {x[0], x[1], x[2]};
        nd = h->tree->FindLeafNode(xx);
    }
    long c = 0;
    for (Int_t i = nd->GetStart(); i < nd->GetEnd(); i++, c++) if (c < cap) member_ids[c] = (int)h->parts[i].GetID();
    return c;
}
/* split nodes in depth-first order (left before right): node ID, cut dimension, cut value, left child's upper boundary in
 * the cut dimension */
static void walk_cuts(Node* nd, vector<int>& ids, vector<int>& dims, vector<double>& vals, vector<double>& leftmax) {
    if (nd->GetLeaf()) return;
    SplitNode* sp = (SplitNode*)nd;
    ids.push_back((int)sp->GetID()); dims.push_back(sp->GetCutDim()); vals.push_back(sp->GetCutValue());
    leftmax.push_back(sp->GetLeft()->GetBoundary(sp->GetCutDim(), 1));
    walk_cuts(sp->GetLeft(), ids, dims, vals, leftmax);
    walk_cuts(sp->GetRight(), ids, dims, vals, leftmax);
}
long ref_dump_cuts(void* hv, int* ids, int* dims, double* vals, double* leftmax, long cap) {
    RefTree* h = (RefTree*)hv;
    vector<int> a, b; vector<double> c, d;
    walk_cuts(h->tree->GetRoot(), a, b, c, d);
    long m = (long)a.size();
    for (long i = 0; i < m && i < cap; i++) { ids[i] = a[i]; dims[i] = b[i]; vals[i] = c[i]; leftmax[i] = d[i]; }
    return m;
}

/* FindNearestCheck(tt | Coordinate x, check_by_type, ...) and FindNearestCriterion(tt | Particle p, cmp, params, ...)
 * (KDFindNearest.cxx:363-441).  Queries are particle IDs q0..q1 (x == NULL) or m points x (with velocities v for the
 * Particle form); crit < 0 selects the check form (Particle::type != 0 => excluded, set with ref_set_types). */
void ref_knn_filtered(void* hv, int crit, double* params, int k, long q0, long q1, long m, const double* x, const double* v,
                      int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
    FOFcompfunc cmp = crit == 0 ? FOF3d : (crit == 1 ? FOFVel : FOF6d);
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
    const long rows = x ? m : q1 - q0;
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long r = 0; r < rows; r++) {
            if (x) {
                if (crit < 0) {
                    Coordinate c(x[3 * r], x[3 * r + 1], x[3 * r + 2]);
                    h->tree->FindNearestCheck(c, check_by_type, params, nn.data(), d2.data(), k);
                } else {
                    Particle p;
                    p.SetPosition(x[3 * r], x[3 * r + 1], x[3 * r + 2]);
                    if (v) p.SetVelocity(v[3 * r], v[3 * r + 1], v[3 * r + 2]);
                    p.SetID(-1); p.SetPID(-1);
                    h->tree->FindNearestCriterion(p, cmp, params, nn.data(), d2.data(), k);
                }
            } else {
                Int_t tt = where[q0 + r];
                if (crit < 0) h->tree->FindNearestCheck(tt, check_by_type, params, nn.data(), d2.data(), k);
                else h->tree->FindNearestCriterion(tt, cmp, params, nn.data(), d2.data(), k);
            }
            for (int j = 0; j < k; j++) {
                out_ids[r * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                out_d2[r * k + j] = d2[j];
            }
        }
    }
}

/* CalcDensity(k) followed by CalcSmoothVel(k) and CalcSmoothVelDisp(smvel, k) (KDCalcSmoothQuantities.cxx:480-614); outputs by ID:
 * rho[n], smvel[n][3], smdisp[n][9] (row-major) */
void ref_calc_smooth_vel(void* hv, int k, double* rho_by_id, double* smvel_by_id, double* smdisp_by_id) {
    RefTree* h = (RefTree*)hv;
    h->tree->CalcDensity(k);
    for (Int_t i = 0; i < h->n; i++) rho_by_id[h->parts[i].GetID()] = h->parts[i].GetDensity();
    Coordinate* sv = h->tree->CalcSmoothVel(k, 1);
    for (Int_t i = 0; i < h->n; i++) for (int j = 0; j < 3; j++) smvel_by_id[3 * i + j] = sv[i][j];
    if (smdisp_by_id) {
        Matrix* sd = h->tree->CalcSmoothVelDisp(sv, k, 1, 1);
        for (Int_t i = 0; i < h->n; i++) for (int j = 0; j < 3; j++) for (int l = 0; l < 3; l++) smdisp_by_id[9 * i + 3 * j + l] = sd[i](j, l);
        delete[] sd;
    }
    delete[] sv;
}

/* CalcSmoothVelSkew / CalcSmoothVelKurtosis (KDCalcSmoothQuantities.cxx:617-765) after CalcDensity, CalcSmoothVel, CalcSmoothVelDisp;
 * outputs by ID, [n][3] each */
void ref_calc_smooth_higher(void* hv, int k, double* skew_by_id, double* kurt_by_id) {
    RefTree* h = (RefTree*)hv;
    h->tree->CalcDensity(k);
    Coordinate* sv = h->tree->CalcSmoothVel(k, 1);
    Matrix* sd = h->tree->CalcSmoothVelDisp(sv, k, 1, 1);
    Coordinate* sk = h->tree->CalcSmoothVelSkew(sv, sd, k, 1, 1, 1);
    Coordinate* ku = h->tree->CalcSmoothVelKurtosis(sv, sd, k, 1, 1, 1);
    for (Int_t i = 0; i < h->n; i++) for (int j = 0; j < 3; j++) { skew_by_id[3 * i + j] = sk[i][j]; kurt_by_id[3 * i + j] = ku[i][j]; }
    delete[] sv; delete[] sd; delete[] sk; delete[] ku;
}

/* FindNearestPhase(tt) for the given particle IDs (which = 0) or FindNearest(tt) (which = 1: on a TPHS tree built with
 * anisotropic = -1 the same 6D search, KDFindNearest.cxx:260-262,300-301).  Neighbour particle IDs, rows ascending. */
void ref_knn_phase_particles(void* hv, int which, int k, long m, const int* qids, int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
    vector<Int_t> where(h->n);
    for (Int_t i = 0; i < h->n; i++) where[h->parts[i].GetID()] = i;
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long q = 0; q < m; q++) {
            Int_t tt = where[qids[q]];
            if (which == 0) h->tree->FindNearestPhase(tt, nn.data(), d2.data(), k);
            else h->tree->FindNearest(tt, nn.data(), d2.data(), k);
            for (int j = 0; j < k; j++) {
                out_ids[q * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                out_d2[q * k + j] = d2[j];
            }
        }
    }
}
/* FindNearestPhase(Double_t* x, Double_t* v, ...) about arbitrary phase-space points */
void ref_knn_phase_points(void* hv, int k, long m, const double* x, const double* v, int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long q = 0; q < m; q++) {
            Double_t xx[3] = {x[3 * q], x[3 * q + 1], x[3 * q + 2]}, vv[3] = {v[3 * q], v[3 * q + 1], v[3 * q + 2]};
            h->tree->FindNearestPhase(xx, vv, nn.data(), d2.data(), k);
            for (int j = 0; j < k; j++) {
                out_ids[q * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                out_d2[q * k + j] = d2[j];
            }
        }
    }
}

/* FindNearestVel(Double_t* v, ...) about arbitrary velocities */
void ref_knn_vel_points(void* hv, int k, long m, const double* v, int* out_ids, double* out_d2) {
    RefTree* h = (RefTree*)hv;
#pragma omp parallel
    {
        vector<Int_t> nn(k);
        vector<Double_t> d2(k);
#pragma omp for schedule(guided)
        for (long q = 0; q < m; q++) {
            Double_t vv[3] = {v[3 * q], v[3 * q + 1], v[3 * q + 2]};
            h->tree->FindNearestVel(vv, nn.data(), d2.data(), k);
            for (int j = 0; j < k; j++) {
                out_ids[q * k + j] = nn[j] >= 0 ? (int)h->parts[nn[j]].GetID() : -1;
                out_d2[q * k + j] = d2[j];
            }
        }
    }
}

}  // extern "C"
