"""Host-side mirror of the reference's ``NBody::KDTree`` interface (reference src/KDTree/KDTree.h:81-657)
over the C ABI of include/nbk.h.  Method names, argument meaning and index conventions follow the
reference so that the parity tests read like the reference's own driver (src/tests/test_kdtree.cxx):

* ``FindNearestPos`` returns tree-order indices (KDTree.h:295) unless ``ids=True``;
* ``CalcDensity`` / ``CalcVelDensity`` / ``FOF`` / ``FOFCriterion`` return arrays indexed by particle ID
  (= input order, KDTree.h:399,474);
* periodic trees follow the reference's image schedule including quirk Q1 unless ``strict=True``.

Inputs may be numpy arrays (host, copied) or torch CUDA tensors (device pointers, no copy).
All computation happens in CUDA kernels; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _lib as L

TPHYS, TPROJ, TVEL, TPHS, TMETRIC = 0, 1, 2, 3, 4
KSPH, KGAUSS, KEPAN, KTH = 0, 1, 2, 3
FOF3D, FOFVEL, FOF6D = 0, 1, 2


def _is_torch(a):
    return a is not None and type(a).__module__.startswith("torch")


def _ptr(a):
    """raw pointer of a numpy array or torch tensor.  The library works on its own CUDA stream: a torch tensor handed to it
    must be complete, so torch's current stream is drained first (results are complete when a call returns: every entry
    point synchronises the library's stream before returning)."""
    if a is None:
        return None
    if _is_torch(a):
        if a.is_cuda:
            import torch
            torch.cuda.current_stream(a.device).synchronize()
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class KDTree:
    TPHYS, TPROJ, TVEL, TPHS, TMETRIC = 0, 1, 2, 3, 4
    KSPH, KGAUSS, KEPAN, KTH = 0, 1, 2, 3

    def __init__(self, pos, vel=None, mass=None, bucket_size=16, TreeType=TPHYS, KernType=KEPAN, KernRes=1000,
                 SplittingCriterion=0, Period=None, device=-1, flags=0, Aniso=0):
        """Mirror of KDTree(Particle*, numparts, bucket_size, TreeType, KernType, KernRes, SplittingCriterion,
        Aniso, ScaleSpace, Period) (KDTree.h:229-245) with the particle array given as pos/vel/mass columns."""
        self._lib = L.load()
        self._h = C.c_void_p()
        self.anisotropic = int(Aniso)     # -1: FindNearest on a TPHS tree is the plain 6D search (KDTree.h:157-158)
        dev_in = _is_torch(pos)
        keep = []

        def prep(a, cols):
            if a is None:
                return None, 0, 0
            if _is_torch(a):
                assert a.is_cuda and a.is_contiguous()
                import torch
                torch.cuda.current_stream(a.device).synchronize()     # the library reads it on its own stream
                rb = a.element_size()
                keep.append(a)
                return C.c_void_p(a.data_ptr()), cols * rb, rb
            a = np.ascontiguousarray(a)
            if a.dtype not in (np.float32, np.float64):
                a = a.astype(np.float64)
            keep.append(a)
            return C.c_void_p(a.ctypes.data), cols * a.dtype.itemsize, a.dtype.itemsize

        p = L.NbkParticles()
        p.pos, p.pos_stride, rb = prep(pos, 3)
        p.vel, p.vel_stride, rbv = prep(vel, 3)
        p.mass, p.mass_stride, rbm = prep(mass, 1)
        for x in (rbv, rbm):
            if x and x != rb:
                raise ValueError("pos / vel / mass must share one dtype")
        p.real_bytes = rb
        p.on_device = 1 if dev_in else 0
        n = int(pos.shape[0])
        per = None
        if Period is not None:
            per = np.ascontiguousarray(Period, dtype=np.float64)
        L.check(self._lib.nbk_create(C.byref(p), n, int(bucket_size), int(TreeType), int(KernType), int(KernRes),
                                     int(SplittingCriterion), _ptr(per), int(flags), int(device), C.byref(self._h)))
        self.n = n
        self.period = per
        del keep

    # ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.nbk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def attach_halo(self, halo):
        """nbk_attach_halo: `halo` (another KDTree on the same device) becomes this tree's second tree and is consumed.
        IDs of its particles follow this tree's; Calc* then query this tree's own particles against both."""
        L.check(self._lib.nbk_attach_halo(self._h, halo._h))
        halo._h = C.c_void_p()
        self.n_main = self.n
        self.n = int(self.info.n)

    @property
    def info(self):
        i = L.NbkInfo()
        L.check(self._lib.nbk_get_info(self._h, C.byref(i)))
        return i

    def GetNumNodes(self):
        return self.info.num_nodes

    def GetNumLeafNodes(self):
        return self.info.num_leaves

    def GetKernNorm(self):
        return self.info.kernnorm

    def GetBucketSize(self):
        return self.info.bucket

    def GetTreeType(self):
        return self.info.treetype

    def GetPeriod(self, j):
        return self.info.period[j]

    def order(self):
        """ID of the particle at each tree index (what Particle::GetID returns after the reference's reorder)."""
        ids = np.empty(self.n, dtype=np.int32)
        L.check(self._lib.nbk_get_order(self._h, _ptr(ids), 0))
        return ids

    def kernel_table(self):
        t = np.empty(self.info.kernres)
        L.check(self._lib.nbk_get_kernel_table(self._h, _ptr(t)))
        return t

    def nodes(self):
        ns = C.c_int64()
        L.check(self._lib.nbk_get_nodes(self._h, C.byref(ns), None, None, None, None))
        m = ns.value
        s, e, c = (np.empty(m, dtype=np.int32) for _ in range(3))
        b = np.empty((m, 6), dtype=np.float32)
        L.check(self._lib.nbk_get_nodes(self._h, C.byref(ns), _ptr(s), _ptr(e), _ptr(c), _ptr(b)))
        return s, e, c, b

    def FindLeafNode(self, q):
        """KDTree::FindLeafNode(Int_t tt) / FindLeafNode(Double_t *x) (KDFindNearest.cxx:709-736) on the host mirror of the
        node arrays: returns (slot, start, end) of the leaf holding tree index q, or the one a position descends into
        (left while x[cut] < the left child's upper boundary)."""
        if getattr(self, "_nodes", None) is None:
            self._nodes = self.nodes()
        s, e, c, b = self._nodes
        slot = 0
        while c[slot] >= 0:
            left = 2 * slot + 1
            if np.ndim(q) == 0:
                slot = left if q < e[left] else left + 1
            else:
                k = c[slot]
                slot = left if q[k] < b[left, 2 * k + 1] else left + 1
        return slot, int(s[slot]), int(e[slot])

    # ---- nearest neighbours -------------------------------------------------------------------
    def FindNearestPos(self, Nsearch=64, q0=0, q1=None, ids=False, tree_form=False, strict=False, out=None):
        """Whole-system / range form of FindNearestPos(Int_t tt, ...) (KDFindNearest.cxx:320-334,444-459).
        Returns (nn, dist2), rows = tree indices q0..q1.  out=(nn, d2) torch CUDA tensors => no copies."""
        q1 = self.n if q1 is None else q1
        flags = (L.OUT_IDS if ids else 0) | (L.KNN_TREE_FORM if tree_form else 0) | (L.STRICT_PERIODIC if strict else 0)
        if out is not None:
            nn, d2 = out
            flags |= L.DEVICE_PTRS
        else:
            nn = np.empty((q1 - q0, Nsearch), dtype=np.int32)
            d2 = np.empty((q1 - q0, Nsearch), dtype=np.float64)
        L.check(self._lib.nbk_knn_particles(self._h, int(Nsearch), int(q0), int(q1), _ptr(nn), _ptr(d2), flags))
        return nn, d2

    def FindNearest(self, Nsearch=64, **kw):
        """FindNearest(Int_t tt, ...) (KDFindNearest.cxx:247-318): on a periodic tree the target itself is dropped."""
        if self.info.treetype == TPHS:
            # KDFindNearest.cxx:260-262,300-301: phase-space search when the tree was built with Aniso = -1; the metric
            # searches of Aniso >= 0 (quirk Q4) are not built
            if self.anisotropic != -1:
                raise L.NbkError(-3, "FindNearest on a TPHS tree with Aniso >= 0 is the reference's metric search (no device implementation)")
            return self.FindNearestPhase(Nsearch, **kw)
        kw.setdefault("tree_form", True)
        return self.FindNearestPos(Nsearch, **kw)

    def FindNearestPosPoints(self, x, Nsearch=64, ids=False, strict=False):
        """Batched FindNearestPos(Double_t *x, ...) (KDFindNearest.cxx:462-554)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        m = len(x)
        nn = np.empty((m, Nsearch), dtype=np.int32)
        d2 = np.empty((m, Nsearch), dtype=np.float64)
        flags = (L.OUT_IDS if ids else 0) | (L.STRICT_PERIODIC if strict else 0)
        L.check(self._lib.nbk_knn_points(self._h, int(Nsearch), m, _ptr(x), _ptr(nn), _ptr(d2), flags))
        return nn, d2

    def FindNearestVel(self, Nsearch=64, q0=0, q1=None, v=None, ids=False):
        """Range / batched form of KDTree::FindNearestVel(Int_t tt | Double_t *v, ...) (KDFindNearest.cxx:335-346,530-540): the
        Nsearch nearest in velocity space on a TVEL tree (never reflected: KDSplitNode.cxx:1082-1085)."""
        if self.info.treetype != TVEL:
            raise L.NbkError(-3, "FindNearestVel needs a TVEL tree (the reference prunes with the tree's own cut planes)")
        if v is not None:
            return self.FindNearestPosPoints(v, Nsearch, ids=ids)
        return self.FindNearestPos(Nsearch, q0=q0, q1=q1, ids=ids)

    def FindNearestPhase(self, Nsearch=64, q0=0, q1=None, x=None, v=None, ids=False):
        """Range / batched form of KDTree::FindNearestPhase(Int_t tt, ...) and FindNearestPhase(Double_t *x, Double_t *v, ...)
        (KDFindNearest.cxx:347-361,543-555): the Nsearch nearest in the plain 6D distance PhaseDistSqd (DistFunc.h:41-49).  Also
        what FindNearest does on a TPHS tree built with Aniso = -1 (KDFindNearest.cxx:260-262).  Rows = tree indices q0..q1,
        or the points (x, v)."""
        flags = L.OUT_IDS if ids else 0
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
            v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
            assert x.shape == v.shape
            rows = len(x)
        else:
            q1 = self.n if q1 is None else q1
            rows = q1 - q0
        nn = np.empty((rows, Nsearch), dtype=np.int32)
        d2 = np.empty((rows, Nsearch), dtype=np.float64)
        if x is not None:
            L.check(self._lib.nbk_knn_phase_points(self._h, int(Nsearch), rows, _ptr(x), _ptr(v), _ptr(nn), _ptr(d2), flags))
        else:
            L.check(self._lib.nbk_knn_phase_particles(self._h, int(Nsearch), int(q0), int(q1), _ptr(nn), _ptr(d2), flags))
        return nn, d2

    def _knn_filtered(self, Nsearch, cmp, params, check, q0, q1, x, v, ids, tree_form):
        flags = (L.OUT_IDS if ids else 0) | (L.KNN_TREE_FORM if tree_form else 0)
        pr = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        chk = None if check is None else np.ascontiguousarray(check, dtype=np.int32)
        if chk is not None:
            assert chk.shape == (self.n,)
        crit = -1 if cmp is None else int(cmp)
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
            v = None if v is None else np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
            rows = len(x)
        else:
            q1 = self.n if q1 is None else q1
            rows = q1 - q0
        nn = np.empty((rows, Nsearch), dtype=np.int32)
        d2 = np.empty((rows, Nsearch), dtype=np.float64)
        if x is not None:
            L.check(self._lib.nbk_knn_filtered_points(self._h, int(Nsearch), rows, _ptr(x), _ptr(v), crit, _ptr(pr), _ptr(chk), _ptr(nn), _ptr(d2), flags))
        else:
            L.check(self._lib.nbk_knn_filtered_particles(self._h, int(Nsearch), int(q0), int(q1), crit, _ptr(pr), _ptr(chk), _ptr(nn), _ptr(d2), flags))
        return nn, d2

    def FindNearestCheck(self, check, Nsearch=64, q0=0, q1=None, x=None, ids=False, tree_form=None):
        """Range / batched form of KDTree::FindNearestCheck(Int_t tt | Coordinate x, check, params, nn, dist2, Nsearch)
        (KDFindNearest.cxx:396-441): the Nsearch nearest among the particles whose check value (array by ID, the caller's
        FOFcheckfunc evaluated per particle) is 0.  tree_form (default: on for periodic trees) reproduces the reference's
        periodic forms, which search Nsearch+1 and drop the nearest."""
        tree_form = (self.period is not None) if tree_form is None else tree_form
        return self._knn_filtered(Nsearch, None, None, check, q0, q1, x, None, ids, tree_form)

    def FindNearestCriterion(self, cmp, params, Nsearch=64, q0=0, q1=None, x=None, v=None, ids=False, tree_form=None):
        """Range / batched form of KDTree::FindNearestCriterion(Int_t tt | Particle p, cmp, params, nn, dist2, Nsearch)
        (KDFindNearest.cxx:363-394) for cmp in {FOF3D, FOF6D}: the Nsearch nearest among the particles meeting the
        criterion relative to the target; rows are padded with (-1, 1e32) when fewer qualify."""
        tree_form = (self.period is not None) if tree_form is None else tree_form
        return self._knn_filtered(Nsearch, cmp, params, None, q0, q1, x, v, ids, tree_form)

    # ---- fixed radius -------------------------------------------------------------------------
    def _csr(self, call, m, ids, want_d2):
        """two-pass CSR protocol of nbk_ball_* / nbk_search_criterion_*: count, allocate, fill"""
        off = np.empty(m + 1, dtype=np.int64)
        tot = C.c_int64()
        flags = L.OUT_IDS if ids else 0
        L.check(call(_ptr(off), None, None, 0, C.byref(tot), flags))
        idx = np.empty(max(tot.value, 1), dtype=np.int32)
        d2 = np.empty(max(tot.value, 1), dtype=np.float64) if want_d2 else None
        L.check(call(_ptr(off), _ptr(idx), _ptr(d2), len(idx), C.byref(tot), flags))
        if want_d2:
            return off, idx[:tot.value], d2[:tot.value]
        return off, idx[:tot.value]

    def SearchBallPosTagged(self, tt, fdist2, ids=False, want_d2=False):
        """Batched SearchBallPosTagged(Int_t tt, fdist2, tagged) (KDFindNearest.cxx:618-626); tt = tree indices.
        Returns CSR (offsets, tagged[, dist2])."""
        q = np.ascontiguousarray(tt, dtype=np.int32)
        return self._csr(lambda *a: self._lib.nbk_ball_particles(self._h, float(fdist2), len(q), _ptr(q), *a), len(q), ids, want_d2)

    def SearchBallPosTaggedPoints(self, x, fdist2, ids=False, want_d2=False):
        """Batched SearchBallPosTagged(Double_t *x, fdist2, tagged) (KDFindNearest.cxx:628-636)."""
        q = np.ascontiguousarray(x, dtype=np.float64)
        return self._csr(lambda *a: self._lib.nbk_ball_points(self._h, float(fdist2), len(q), _ptr(q), *a), len(q), ids, want_d2)

    def SearchBallPos(self, tt, fdist2, imark, nn, dist2):
        """Dense KDTree::SearchBallPos(Int_t tt | Double_t *x, fdist2, imark, nn, dist2) (KDFindNearest.cxx:567-587): marks
        nn[ID] = imark and dist2[ID] = d2 for every particle within the ball; nn / dist2 are the caller's N-entry arrays
        (indexed by ID) and are updated in place.  tt: a tree index, or a position (3 floats)."""
        if np.ndim(tt) == 0:
            off, idx, d2 = self.SearchBallPosTagged([int(tt)], fdist2, ids=True, want_d2=True)
        else:
            off, idx, d2 = self.SearchBallPosTaggedPoints(np.asarray(tt, dtype=np.float64).reshape(1, 3), fdist2, ids=True, want_d2=True)
        nn[idx] = imark
        dist2[idx] = d2
        return len(idx)

    def SearchCriterionTagged(self, tt, cmp, params, ids=False, want_d2=False):
        """Batched KDTree::SearchCriterionTagged(Int_t tt, cmp, params, tagged) (KDFindNearest.cxx:660-668) for cmp in
        {FOF3D, FOF6D}; tt = tree indices.  Returns CSR (offsets, tagged[, dist2])."""
        q = np.ascontiguousarray(tt, dtype=np.int32)
        pr = np.ascontiguousarray(params, dtype=np.float64)
        return self._csr(lambda *a: self._lib.nbk_search_criterion_particles(self._h, int(cmp), _ptr(pr), len(q), _ptr(q), *a), len(q), ids, want_d2)

    def SearchCriterionTaggedPoints(self, x, v, cmp, params, ids=False, want_d2=False):
        """Batched KDTree::SearchCriterionTagged(Particle &p, cmp, params, tagged) (KDFindNearest.cxx:669-677) for particles
        that are not in the tree, given as positions x and velocities v (v may be None for FOF3D)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        v = None if v is None else np.ascontiguousarray(v, dtype=np.float64)
        pr = np.ascontiguousarray(params, dtype=np.float64)
        return self._csr(lambda *a: self._lib.nbk_search_criterion_points(self._h, int(cmp), _ptr(pr), len(x), _ptr(x), _ptr(v), *a), len(x), ids, want_d2)

    def SearchCriterion(self, tt, cmp, params, imark, nn, dist2=None):
        """Dense KDTree::SearchCriterion(Int_t tt, cmp, params, imark, nn[, dist2]) (KDFindNearest.cxx:590-603)."""
        off, idx, d2 = self.SearchCriterionTagged([int(tt)], cmp, params, ids=True, want_d2=True)
        nn[idx] = imark
        if dist2 is not None:
            dist2[idx] = d2
        return len(idx)

    # ---- smoothed estimators ------------------------------------------------------------------
    def CalcDensity(self, Nsmooth=64, want_h=False, out=None):
        """KDTree::CalcDensity (KDCalcSmoothQuantities.cxx:203-305); returns rho indexed by ID."""
        if out is not None:     # torch CUDA tensor (no copy) or a host buffer (numpy / pinned torch CPU tensor)
            dev = _is_torch(out) and out.is_cuda
            L.check(self._lib.nbk_calc_density(self._h, int(Nsmooth), _ptr(out), None, L.DEVICE_PTRS if dev else 0))
            return out
        rho = np.empty(self.n)
        h = np.empty(self.n) if want_h else None
        L.check(self._lib.nbk_calc_density(self._h, int(Nsmooth), _ptr(rho), _ptr(h), 0))
        return (rho, h) if want_h else rho

    def CalcDensityInto(self, Nsmooth, rho_out, hsm_out=None):
        """CalcDensity into caller arrays indexed by ID (torch CUDA tensors => device pointers, no copies).  On a tree with
        an attached halo the arrays have n_main + n_halo entries and only the main particles are queries."""
        flags = L.DEVICE_PTRS if _is_torch(rho_out) else 0
        L.check(self._lib.nbk_calc_density(self._h, int(Nsmooth), _ptr(rho_out), _ptr(hsm_out), flags))
        return rho_out

    def CalcDensitySubset(self, Nsmooth, active, rho_out, hsm_out=None):
        """CalcDensity restricted to the query particles with active[id] != 0 (uint8); ghosts act as neighbours only.
        active / rho_out / hsm_out: torch CUDA tensors (device pointers) or numpy arrays, all indexed by ID."""
        flags = L.DEVICE_PTRS if _is_torch(rho_out) else 0
        L.check(self._lib.nbk_calc_density_subset(self._h, int(Nsmooth), _ptr(active), _ptr(rho_out), _ptr(hsm_out), flags))
        return rho_out

    def CalcVelDensity(self, Nsmooth=64, Nsearch=64, out=None):
        """KDTree::CalcVelDensity (KDCalcSmoothQuantities.cxx:309-389)."""
        if out is not None:
            dev = _is_torch(out) and out.is_cuda
            L.check(self._lib.nbk_calc_veldensity(self._h, int(Nsmooth), int(Nsearch), _ptr(out), L.DEVICE_PTRS if dev else 0))
            return out
        rho = np.empty(self.n)
        L.check(self._lib.nbk_calc_veldensity(self._h, int(Nsmooth), int(Nsearch), _ptr(rho), 0))
        return rho

    def CalcDensityParticle(self, target, Nsmooth=64):
        """KDTree::CalcDensityParticle(target, Nsmooth) (KDCalcSmoothQuantities.cxx:768-844) for a batch of tree indices
        (None: every particle, in tree order): gather-only density, one value per query."""
        if target is None:
            q, m = None, self.n
        else:
            q = np.ascontiguousarray(np.atleast_1d(target), dtype=np.int32)
            m = len(q)
        out = np.empty(m)
        L.check(self._lib.nbk_calc_density_particles(self._h, int(Nsmooth), m, _ptr(q), _ptr(out), 0))
        return out if target is None or np.ndim(target) else float(out[0])

    def CalcVelDensityParticle(self, target, Nsmooth=64, Nsearch=64):
        """KDTree::CalcVelDensityParticle(target, Nsmooth, Nsearch) (KDCalcSmoothQuantities.cxx:845-921), batched."""
        if target is None:
            q, m = None, self.n
        else:
            q = np.ascontiguousarray(np.atleast_1d(target), dtype=np.int32)
            m = len(q)
        out = np.empty(m)
        L.check(self._lib.nbk_calc_veldensity_particles(self._h, int(Nsmooth), int(Nsearch), m, _ptr(q), _ptr(out), 0))
        return out if target is None or np.ndim(target) else float(out[0])

    def CalcDensityPosition(self, x, Nsmooth=64):
        """KDTree::CalcDensityPosition(Double_t *x, Nsmooth) (KDCalcSmoothQuantities.cxx:1092-1148), batched: x is (m, 3)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(x))
        L.check(self._lib.nbk_calc_density_points(self._h, int(Nsmooth), len(x), _ptr(x), _ptr(out), 0))
        return out

    def CalcVelDensityPosition(self, x, v, Nsmooth=64, Nsearch=64):
        """KDTree::CalcVelDensityPosition(Double_t *x, Double_t *v, Nsmooth, Nsearch) (KDCalcSmoothQuantities.cxx:1150-1207)."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        assert x.shape == v.shape
        out = np.empty(len(x))
        L.check(self._lib.nbk_calc_veldensity_points(self._h, int(Nsmooth), int(Nsearch), len(x), _ptr(x), _ptr(v), _ptr(out), 0))
        return out

    def CalcSmoothLocalValue(self, Nsmooth, dist, weight):
        """KDTree::CalcSmoothLocalValue(Nsmooth, Double_t *dist, Double_t *weight) (KDCalcSmoothQuantities.cxx:1721-1735):
        kernel-weighted sum over a caller-supplied list of Nsmooth distances (descending, dist[0] the largest) and
        weights.  Host data in, one number out: evaluated here with the tree's kernel table (nbk_get_kernel_table), the
        same few lines the C++ shim carries."""
        info = self.info
        kern, kernres, nd = self.kernel_table(), info.kernres, info.nd
        hi = 0.5 * dist[0]
        norm = 1.0 / hi ** float(nd)
        delta = 2.0 / (kernres - 1)
        value = 0.0
        for j in range(Nsmooth):
            r = dist[j] / hi
            i = int(r * 0.5 * (kernres - 1))
            w = kern[i] + (kern[i + 1] - kern[i]) * (r - delta * i) / delta if i < kernres - 1 else kern[i]
            value += w * norm * weight[j]
        return value

    def CalcSmoothVel(self, Nsmooth=64, rho=None):
        """KDTree::CalcSmoothVel(Nsmooth, densityset) (KDCalcSmoothQuantities.cxx:480-541): smoothed mean velocity of every
        particle, (n, 3) by ID.  rho = the particles' densities by ID (None: CalcDensity(Nsmooth) is computed first)."""
        r = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64)
        out = np.empty((self.n, 3))
        L.check(self._lib.nbk_calc_smooth_vel(self._h, int(Nsmooth), _ptr(r), _ptr(out), 0))
        return out

    def CalcSmoothVelDisp(self, smvel, Nsmooth=64, rho=None):
        """KDTree::CalcSmoothVelDisp(smvel, Nsmooth, densityset) (KDCalcSmoothQuantities.cxx:542-614): (n, 3, 3) by ID."""
        r = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64)
        sv = np.ascontiguousarray(smvel, dtype=np.float64)
        assert sv.shape == (self.n, 3)
        out = np.empty((self.n, 3, 3))
        L.check(self._lib.nbk_calc_smooth_veldisp(self._h, int(Nsmooth), _ptr(r), _ptr(sv), _ptr(out), 0))
        return out

    def _higher(self, call, smvel, smveldisp, Nsmooth, rho):
        r = None if rho is None else np.ascontiguousarray(rho, dtype=np.float64)
        sv = np.ascontiguousarray(smvel, dtype=np.float64)
        sd = np.ascontiguousarray(smveldisp, dtype=np.float64)
        assert sv.shape == (self.n, 3) and sd.shape == (self.n, 3, 3)
        out = np.empty((self.n, 3))
        L.check(call(self._h, int(Nsmooth), _ptr(r), _ptr(sv), _ptr(sd), _ptr(out), 0))
        return out

    def CalcSmoothVelSkew(self, smvel, smveldisp, Nsmooth=64, rho=None):
        """KDTree::CalcSmoothVelSkew(smvel, smveldisp, Nsmooth, ...) (KDCalcSmoothQuantities.cxx:617-689): (n, 3) by ID."""
        return self._higher(self._lib.nbk_calc_smooth_velskew, smvel, smveldisp, Nsmooth, rho)

    def CalcSmoothVelKurtosis(self, smvel, smveldisp, Nsmooth=64, rho=None):
        """KDTree::CalcSmoothVelKurtosis(smvel, smveldisp, Nsmooth, ...) (KDCalcSmoothQuantities.cxx:692-765): (n, 3) by ID; like the
        reference, 3 is subtracted from every neighbour contribution."""
        return self._higher(self._lib.nbk_calc_smooth_velkurtosis, smvel, smveldisp, Nsmooth, rho)

    def CalcSmoothingScale(self, Nsmooth=64):
        """hi = 0.5*sqrt(d2 of the Nsmooth-th neighbour) (KDCalcSmoothQuantities.cxx:260)."""
        h = np.empty(self.n)
        L.check(self._lib.nbk_smoothing_scale(self._h, int(Nsmooth), _ptr(h), 0))
        return h

    # ---- friends of friends -------------------------------------------------------------------
    def FOF(self, fdist, minnum=8, order=0, precheck=None, want_len=False, out=None, want_lists=False):
        """KDTree::FOF(fdist, numgroup, minnum, order, ...) (KDFOF.cxx:29-153).  Returns (pfof by ID, numgroup)."""
        ng = C.c_int64()
        if out is not None:
            L.check(self._lib.nbk_fof(self._h, float(fdist), int(minnum), int(order), None, _ptr(out), C.byref(ng), None, L.DEVICE_PTRS))
            return out, ng.value
        g = np.empty(self.n, dtype=np.int32)
        pre = None if precheck is None else np.ascontiguousarray(precheck, dtype=np.int32)
        lists, plen = None, None
        if want_lists:
            plen = np.zeros(self.n + 1, dtype=np.int32)
            head, nxt, tail = (np.empty(self.n, dtype=np.int32) for _ in range(3))
            lists = L.NbkFofLists(head.ctypes.data, nxt.ctypes.data, tail.ctypes.data, plen.ctypes.data)
        elif want_len:
            plen = np.zeros(self.n + 1, dtype=np.int32)
            lists = L.NbkFofLists(None, None, None, plen.ctypes.data)
        L.check(self._lib.nbk_fof(self._h, float(fdist), int(minnum), int(order), _ptr(pre), _ptr(g), C.byref(ng),
                                  C.byref(lists) if lists else None, 0))
        if want_lists:
            return g, ng.value, {"pHead": head, "pNext": nxt, "pTail": tail, "pLen": plen[:ng.value + 1]}
        if want_len:
            return g, ng.value, plen[:ng.value + 1]
        return g, ng.value

    def FOFRoots(self, fdist=0.0, cmp=-1, params=None, precheck=None, out=None):
        """nbk_fof_roots: the components of FOF(fdist) (cmp < 0) or FOFCriterion(cmp, params) without the minnum filter and the
        numbering: root[ID] = ID of one fixed member of the particle's component (-1: excluded by precheck).  The building
        block of the slab-sharded FOF.  out: int32 torch CUDA tensor => the result stays on the device."""
        pr = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        if out is not None:
            L.check(self._lib.nbk_fof_roots(self._h, int(cmp), float(fdist), _ptr(pr), None, _ptr(out), L.DEVICE_PTRS))
            return out
        pre = None if precheck is None else np.ascontiguousarray(precheck, dtype=np.int32)
        root = np.empty(self.n, dtype=np.int32)
        L.check(self._lib.nbk_fof_roots(self._h, int(cmp), float(fdist), _ptr(pr), _ptr(pre), _ptr(root), 0))
        return root

    def FOFCriterionSetBasisForLinks(self, cmp, params, check, minnum=8, order=0):
        """KDTree::FOFCriterionSetBasisForLinks(cmp, params, numgroup, minnum, order, ipcheckflag, check) (KDFOF.cxx:268-378).
        check: the FOFcheckfunc values by ID; only check == 0 particles start / extend groups, the others can only be
        linked into one (include/nbk.h states the tie rule)."""
        ng = C.c_int64()
        g = np.empty(self.n, dtype=np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64)
        chk = np.ascontiguousarray(check, dtype=np.int32)
        assert chk.shape == (self.n,)
        L.check(self._lib.nbk_fof_criterion_basis(self._h, int(cmp), _ptr(params), int(minnum), int(order), _ptr(chk), _ptr(g),
                                                  C.byref(ng), None, 0))
        return g, ng.value

    def FOFCriterion(self, cmp, params, minnum=8, order=0, precheck=None, out=None):
        """KDTree::FOFCriterion(cmp, params, numgroups, minnum, order) (KDFOF.cxx:157-265) for cmp in
        {FOF3D, FOF6D} (FOFFunc.h:30-55).  out: int32 torch CUDA tensor (by ID) => the group ids stay on the device."""
        ng = C.c_int64()
        params = np.ascontiguousarray(params, dtype=np.float64)
        if out is not None:
            L.check(self._lib.nbk_fof_criterion(self._h, int(cmp), _ptr(params), int(minnum), int(order), None, _ptr(out),
                                                C.byref(ng), None, L.DEVICE_PTRS))
            return out, ng.value
        g = np.empty(self.n, dtype=np.int32)
        pre = None if precheck is None else np.ascontiguousarray(precheck, dtype=np.int32)
        L.check(self._lib.nbk_fof_criterion(self._h, int(cmp), _ptr(params), int(minnum), int(order), _ptr(pre), _ptr(g),
                                            C.byref(ng), None, 0))
        return g, ng.value
