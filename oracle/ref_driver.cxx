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

/* kNN for queries q0..q1 (particle IDs, input order).  which: 0 = FindNearestPos(tt), 1 = FindNearest(tt).
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

}  // extern "C"
