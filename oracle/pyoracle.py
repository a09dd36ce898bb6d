"""ctypes front-ends for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Port``  -- oracle/liboracle.so, the plain-C restatement (oracle/nbk_oracle.c).  Always available
  (built by ``__graft_entry__.build()`` / ``make -C oracle port``).
* ``Ref``   -- oracle/_ref/libnbref.so, the unmodified reference sources + oracle/ref_driver.cxx.
  Built only where /root/reference exists; the prebuilt .so travels to the GPU box.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product package (nbodylib_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libnbref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _l(a):
    return None if a is None else a.ctypes.data_as(_lp)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def build_port():
    if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "nbk_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return PORT_SO


def build_ref(reference="/root/reference"):
    """Compile the reference where it lies; no-op (returns None) if the sources are absent."""
    if os.path.isdir(os.path.join(reference, "src", "KDTree")):
        src = os.path.join(HERE, "ref_driver.cxx")
        if not os.path.exists(REF_SO) or os.path.getmtime(REF_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", HERE, "ref", "REF=" + reference], stdout=subprocess.DEVNULL)
    return REF_SO if os.path.exists(REF_SO) else None


def have_ref():
    return os.path.exists(REF_SO)


HARNESS = os.path.join(HERE, "_ref", "test_kdtree_shim")


def build_harness(reference="/root/reference"):
    """The reference's own test program (src/tests/test_kdtree.cxx, unchanged) compiled against nbodylib_b200/shim/KDTree.h with
    the reference's NBody / Math headers and linked with libnbk.so (oracle/Makefile target `harness`).  No-op without the
    reference sources; returns the binary's path or None."""
    if os.path.isdir(os.path.join(reference, "src", "tests")):
        deps = [os.path.join(HERE, "harness_main.cxx"), os.path.join(HERE, "..", "nbodylib_b200", "shim", "KDTree.h"),
                os.path.join(HERE, "..", "nbodylib_b200", "libnbk.so")]
        if not os.path.exists(HARNESS) or any(os.path.getmtime(HARNESS) < os.path.getmtime(d) for d in deps if os.path.exists(d)):
            subprocess.check_call(["make", "-B", "-C", HERE, "harness", "REF=" + reference], stdout=subprocess.DEVNULL)
    return HARNESS if os.path.exists(HARNESS) else None


class Port:
    """Brute-force restatement; all indices are particle IDs (input order)."""

    def __init__(self):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.orc_kernel_table.restype = C.c_double
        L.orc_kernel_table.argtypes = [C.c_int, C.c_int, C.c_int, _dp]
        L.orc_knn_particles.argtypes = [C.c_long, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_long, C.c_long, _ip, _dp]
        L.orc_knn_points.argtypes = [C.c_long, _dp, C.c_int, _dp, C.c_int, C.c_long, _dp, _ip, _dp]
        L.orc_density.argtypes = [C.c_long, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.orc_veldensity.argtypes = [C.c_long, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_fof.restype = C.c_long
        L.orc_fof.argtypes = [C.c_long, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int, _ip]
        L.orc_ball_points.restype = C.c_long
        L.orc_ball_points.argtypes = [C.c_long, _dp, _dp, C.c_double, C.c_long, _dp, _lp, _ip, C.c_long]
        L.orc_knn_phase_particles.argtypes = [C.c_long, _dp, _dp, C.c_int, _dp, C.c_long, _ip, _ip, _dp]
        L.orc_knn_phase_points.argtypes = [C.c_long, _dp, _dp, C.c_int, _dp, C.c_long, _dp, _dp, _ip, _dp]

    def kernel_table(self, nd, kerntype, kernres):
        t = np.zeros(kernres)
        norm = self.lib.orc_kernel_table(nd, kerntype, kernres, _d(t))
        return norm, t

    def knn_particles(self, pos, k, period=None, which=0, strict=0, q0=0, q1=None):
        pos = _f64(pos)
        n = len(pos)
        q1 = n if q1 is None else q1
        ids = np.zeros((q1 - q0, k), dtype=np.int32)
        d2 = np.zeros((q1 - q0, k))
        self.lib.orc_knn_particles(n, _d(pos), k, _d(_f64(period)), which, strict, q0, q1, _i(ids), _d(d2))
        return ids, d2

    def knn_points(self, pos, x, k, period=None, strict=0):
        pos, x = _f64(pos), _f64(x)
        ids = np.zeros((len(x), k), dtype=np.int32)
        d2 = np.zeros((len(x), k))
        self.lib.orc_knn_points(len(pos), _d(pos), k, _d(_f64(period)), strict, len(x), _d(x), _i(ids), _d(d2))
        return ids, d2

    def knn_phase_particles(self, pos, vel, qids, k, period=None):
        """FindNearestPhase(tt): 6D neighbours of the particles `qids` (IDs)"""
        pos, vel = _f64(pos), _f64(vel)
        q = np.ascontiguousarray(qids, dtype=np.int32)
        ids = np.zeros((len(q), k), dtype=np.int32)
        d2 = np.zeros((len(q), k))
        self.lib.orc_knn_phase_particles(len(pos), _d(pos), _d(vel), k, _d(_f64(period)), len(q), _i(q), _i(ids), _d(d2))
        return ids, d2

    def knn_phase_points(self, pos, vel, x, v, k, period=None):
        """FindNearestPhase(x, v): 6D neighbours of arbitrary phase-space points"""
        pos, vel, x, v = _f64(pos), _f64(vel), _f64(x), _f64(v)
        ids = np.zeros((len(x), k), dtype=np.int32)
        d2 = np.zeros((len(x), k))
        self.lib.orc_knn_phase_points(len(pos), _d(pos), _d(vel), k, _d(_f64(period)), len(x), _d(x), _d(v), _i(ids), _d(d2))
        return ids, d2

    def density(self, pos, mass, k, kerntype=2, kernres=1000):
        pos, mass = _f64(pos), _f64(mass)
        rho = np.zeros(len(pos))
        h = np.zeros(len(pos))
        self.lib.orc_density(len(pos), _d(pos), _d(mass), k, kerntype, kernres, _d(rho), _d(h))
        return rho, h

    def veldensity(self, pos, vel, kv, kx, kerntype=2, kernres=1000):
        pos, vel = _f64(pos), _f64(vel)
        rho = np.zeros(len(pos))
        self.lib.orc_veldensity(len(pos), _d(pos), _d(vel), kv, kx, kerntype, kernres, _d(rho))
        return rho

    def fof(self, pos, vel, mode, params, period=None, minnum=8, order=0):
        pos, vel, params = _f64(pos), _f64(vel), _f64(params)
        g = np.zeros(len(pos), dtype=np.int32)
        ng = self.lib.orc_fof(len(pos), _d(pos), _d(vel), mode, _d(params), _d(_f64(period)), minnum, order, _i(g))
        return g, ng

    def ball_points(self, pos, x, r2, period=None, cap=None):
        pos, x = _f64(pos), _f64(x)
        cap = cap or 64 * len(x) + 1024
        while True:
            off = np.zeros(len(x) + 1, dtype=np.int64)
            ids = np.zeros(cap, dtype=np.int32)
            tot = self.lib.orc_ball_points(len(pos), _d(pos), _d(_f64(period)), r2, len(x), _d(x), _l(off), _i(ids), cap)
            if tot <= cap:
                return off, ids[:tot]
            cap = tot


class Ref:
    """The reference itself (NBody::KDTree) behind oracle/ref_driver.cxx."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise RuntimeError("oracle/_ref/libnbref.so not built (needs /root/reference; run make -C oracle ref)")
            L = C.CDLL(REF_SO)
            L.ref_create.restype = C.c_void_p
            L.ref_create.argtypes = [C.c_long, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
            L.ref_destroy.argtypes = [C.c_void_p]
            for f in ("ref_build_seconds", "ref_kernnorm"):
                getattr(L, f).restype = C.c_double
                getattr(L, f).argtypes = [C.c_void_p]
            for f in ("ref_num_nodes", "ref_num_leaves"):
                getattr(L, f).restype = C.c_long
                getattr(L, f).argtypes = [C.c_void_p]
            L.ref_order.argtypes = [C.c_void_p, _ip]
            L.ref_knn_particles.restype = C.c_double
            L.ref_knn_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_long, _ip, _dp]
            L.ref_knn_particle_list.restype = C.c_double
            L.ref_knn_particle_list.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _ip, _ip, _dp]
            L.ref_set_threads.argtypes = [C.c_int]
            L.ref_knn_points.restype = C.c_double
            L.ref_knn_points.argtypes = [C.c_void_p, C.c_int, C.c_long, _dp, _ip, _dp]
            L.ref_ball_particles.restype = C.c_long
            L.ref_ball_particles.argtypes = [C.c_void_p, C.c_double, C.c_long, _ip, _lp, _ip, C.c_long]
            L.ref_ball_points.restype = C.c_long
            L.ref_ball_points.argtypes = [C.c_void_p, C.c_double, C.c_long, _dp, _lp, _ip, C.c_long]
            L.ref_calc_density.restype = C.c_double
            L.ref_calc_density.argtypes = [C.c_void_p, C.c_int, _dp]
            L.ref_calc_veldensity.restype = C.c_double
            L.ref_calc_veldensity.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
            L.ref_calc_density_omp.restype = C.c_double
            L.ref_calc_density_omp.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, _dp]
            L.ref_fof.restype = C.c_double
            L.ref_fof.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, _ip, _lp]
            L.ref_fof_criterion.restype = C.c_double
            L.ref_fof_criterion.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, _ip, _lp]
            L.ref_set_types.argtypes = [C.c_void_p, _ip]
            L.ref_fof_checked.restype = C.c_double
            L.ref_fof_checked.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_int, C.c_int, _ip, _lp]
            L.ref_scale_phase.argtypes = [C.c_long, _dp, _dp, C.c_double, C.c_double]
            L.ref_dump_nodes.restype = C.c_long
            L.ref_dump_nodes.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, C.c_long]
            L.ref_calc_density_particles.argtypes = [C.c_void_p, C.c_int, C.c_long, _ip, _dp]
            L.ref_calc_veldensity_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _ip, _dp]
            L.ref_calc_density_points.argtypes = [C.c_void_p, C.c_int, C.c_long, _dp, _dp]
            L.ref_calc_veldensity_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _dp, _dp, _dp]
            L.ref_smooth_local_value.restype = C.c_double
            L.ref_smooth_local_value.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
            L.ref_search_criterion_particles.restype = C.c_long
            L.ref_search_criterion_particles.argtypes = [C.c_void_p, C.c_int, _dp, C.c_long, _ip, _lp, _ip, C.c_long]
            L.ref_search_criterion_points.restype = C.c_long
            L.ref_search_criterion_points.argtypes = [C.c_void_p, C.c_int, _dp, C.c_long, _dp, _dp, _lp, _ip, C.c_long]
            L.ref_search_ball_dense.argtypes = [C.c_void_p, C.c_long, _dp, C.c_double, C.c_int, _ip, _dp]
            L.ref_search_criterion_dense.argtypes = [C.c_void_p, C.c_int, _dp, C.c_long, C.c_int, _ip, _dp]
            L.ref_find_leaf.restype = C.c_long
            L.ref_find_leaf.argtypes = [C.c_void_p, C.c_long, _dp, _ip, C.c_long]
            L.ref_dump_cuts.restype = C.c_long
            L.ref_dump_cuts.argtypes = [C.c_void_p, _ip, _ip, _dp, _dp, C.c_long]
            L.ref_knn_filtered.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_long, C.c_long, C.c_long, _dp, _dp, _ip, _dp]
            L.ref_calc_smooth_vel.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
            L.ref_calc_smooth_higher.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
            L.ref_knn_phase_particles.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _ip, _ip, _dp]
            L.ref_knn_phase_points.argtypes = [C.c_void_p, C.c_int, C.c_long, _dp, _dp, _ip, _dp]
            L.ref_knn_vel_points.argtypes = [C.c_void_p, C.c_int, C.c_long, _dp, _ip, _dp]
            L.ref_max_threads.restype = C.c_int
            L.ref_sizeof_particle.restype = C.c_int
            cls._lib = L
        return cls._lib

    TPHYS, TPROJ, TVEL, TPHS = 0, 1, 2, 3
    KSPH, KGAUSS, KEPAN, KTH = 0, 1, 2, 3

    def __init__(self, pos, vel=None, mass=None, bucket=16, treetype=0, kerntype=2, kernres=1000, period=None, aniso=-1):
        L = self.lib()
        self.pos, self.vel, self.mass = _f64(pos), _f64(vel), _f64(mass)
        self.n = len(self.pos)
        self.kerntype, self.kernres = kerntype, kernres
        self.period = _f64(period)
        self.h = L.ref_create(self.n, _d(self.pos), _d(self.vel), _d(self.mass), bucket, treetype, kerntype, kernres,
                              _d(self.period), aniso)

    def close(self):
        if self.h:
            self.lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def build_seconds(self):
        return self.lib().ref_build_seconds(self.h)

    @property
    def num_nodes(self):
        return self.lib().ref_num_nodes(self.h)

    @property
    def num_leaves(self):
        return self.lib().ref_num_leaves(self.h)

    @property
    def kernnorm(self):
        return self.lib().ref_kernnorm(self.h)

    @classmethod
    def max_threads(cls):
        return cls.lib().ref_max_threads()

    def order(self):
        ids = np.zeros(self.n, dtype=np.int32)
        self.lib().ref_order(self.h, _i(ids))
        return ids

    def knn_particles(self, k, which=0, q0=0, q1=None, want=True):
        q1 = self.n if q1 is None else q1
        ids = np.zeros((q1 - q0, k), dtype=np.int32) if want else None
        d2 = np.zeros((q1 - q0, k)) if want else None
        sec = self.lib().ref_knn_particles(self.h, which, k, q0, q1, _i(ids), _d(d2))
        self.last_seconds = sec
        return ids, d2

    def knn_particle_list(self, qids, k, which=0):
        """FindNearestPos(tt) (which 0) / FindNearest(tt) (1) / FindNearestVel(tt) (2) for an explicit list of particle IDs"""
        q = np.ascontiguousarray(qids, dtype=np.int32)
        ids = np.zeros((len(q), k), dtype=np.int32)
        d2 = np.zeros((len(q), k))
        self.last_seconds = self.lib().ref_knn_particle_list(self.h, which, k, len(q), _i(q), _i(ids), _d(d2))
        return ids, d2

    @classmethod
    def set_threads(cls, n):
        """OpenMP thread count of the reference's loops (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
        cls.lib().ref_set_threads(int(n))

    def knn_points(self, x, k):
        x = _f64(x)
        ids = np.zeros((len(x), k), dtype=np.int32)
        d2 = np.zeros((len(x), k))
        self.last_seconds = self.lib().ref_knn_points(self.h, k, len(x), _d(x), _i(ids), _d(d2))
        return ids, d2

    def knn_phase_particles(self, qids, k, which=0):
        """FindNearestPhase(tt) (which = 0) or FindNearest(tt) (which = 1) for a list of particle IDs"""
        q = np.ascontiguousarray(qids, dtype=np.int32)
        ids = np.zeros((len(q), k), dtype=np.int32)
        d2 = np.zeros((len(q), k))
        self.lib().ref_knn_phase_particles(self.h, which, k, len(q), _i(q), _i(ids), _d(d2))
        return ids, d2

    def knn_vel_points(self, v, k):
        """FindNearestVel(Double_t *v, ...)"""
        v = _f64(v)
        ids = np.zeros((len(v), k), dtype=np.int32)
        d2 = np.zeros((len(v), k))
        self.lib().ref_knn_vel_points(self.h, k, len(v), _d(v), _i(ids), _d(d2))
        return ids, d2

    def knn_phase_points(self, x, v, k):
        x, v = _f64(x), _f64(v)
        ids = np.zeros((len(x), k), dtype=np.int32)
        d2 = np.zeros((len(x), k))
        self.lib().ref_knn_phase_points(self.h, k, len(x), _d(x), _d(v), _i(ids), _d(d2))
        return ids, d2

    def ball_particles(self, qids, r2):
        qids = np.ascontiguousarray(qids, dtype=np.int32)
        cap = 64 * len(qids) + 1024
        while True:
            off = np.zeros(len(qids) + 1, dtype=np.int64)
            ids = np.zeros(cap, dtype=np.int32)
            tot = self.lib().ref_ball_particles(self.h, r2, len(qids), _i(qids), _l(off), _i(ids), cap)
            if tot <= cap:
                return off, ids[:tot]
            cap = tot

    def ball_points(self, x, r2):
        x = _f64(x)
        cap = 64 * len(x) + 1024
        while True:
            off = np.zeros(len(x) + 1, dtype=np.int64)
            ids = np.zeros(cap, dtype=np.int32)
            tot = self.lib().ref_ball_points(self.h, r2, len(x), _d(x), _l(off), _i(ids), cap)
            if tot <= cap:
                return off, ids[:tot]
            cap = tot

    def calc_density(self, k):
        rho = np.zeros(self.n)
        self.last_seconds = self.lib().ref_calc_density(self.h, k, _d(rho))
        return rho

    def calc_veldensity(self, kv, kx):
        rho = np.zeros(self.n)
        self.last_seconds = self.lib().ref_calc_veldensity(self.h, kv, kx, _d(rho))
        return rho

    def calc_density_omp(self, k, i0=0, i1=None, want=True):
        """Full-host OpenMP kNN-density (BASELINE.md section 3 variant ii); tree must be non periodic."""
        i1 = self.n if i1 is None else i1
        rho = np.zeros(self.n) if want else None
        self.last_seconds = self.lib().ref_calc_density_omp(self.h, k, i0, i1, self.kernres, self.kerntype, _d(rho))
        return rho

    # ---- single-target estimators, criterion search, dense forms, node getters -------------------------------
    def calc_density_particles(self, qids, k):
        q = np.ascontiguousarray(qids, dtype=np.int32)
        out = np.zeros(len(q))
        self.lib().ref_calc_density_particles(self.h, k, len(q), _i(q), _d(out))
        return out

    def calc_veldensity_particles(self, qids, kv, kx):
        q = np.ascontiguousarray(qids, dtype=np.int32)
        out = np.zeros(len(q))
        self.lib().ref_calc_veldensity_particles(self.h, kv, kx, len(q), _i(q), _d(out))
        return out

    def calc_density_points(self, x, k):
        x = _f64(x)
        out = np.zeros(len(x))
        self.lib().ref_calc_density_points(self.h, k, len(x), _d(x), _d(out))
        return out

    def calc_veldensity_points(self, x, v, kv, kx):
        x, v = _f64(x), _f64(v)
        out = np.zeros(len(x))
        self.lib().ref_calc_veldensity_points(self.h, kv, kx, len(x), _d(x), _d(v), _d(out))
        return out

    def smooth_local_value(self, dist, weight):
        dist, weight = _f64(dist).copy(), _f64(weight).copy()
        return self.lib().ref_smooth_local_value(self.h, len(dist), _d(dist), _d(weight))

    def _csr(self, call, m):
        cap = 64 * m + 1024
        while True:
            off = np.zeros(m + 1, dtype=np.int64)
            ids = np.zeros(cap, dtype=np.int32)
            tot = call(_l(off), _i(ids), cap)
            if tot <= cap:
                return off, ids[:tot]
            cap = tot

    def search_criterion_particles(self, qids, crit, params):
        q = np.ascontiguousarray(qids, dtype=np.int32)
        pr = _f64(params).copy()
        return self._csr(lambda off, ids, cap: self.lib().ref_search_criterion_particles(self.h, crit, _d(pr), len(q), _i(q), off, ids, cap), len(q))

    def search_criterion_points(self, x, v, crit, params):
        x = _f64(x)
        v = _f64(v)
        pr = _f64(params).copy()
        return self._csr(lambda off, ids, cap: self.lib().ref_search_criterion_points(self.h, crit, _d(pr), len(x), _d(x), _d(v), off, ids, cap), len(x))

    def search_ball_dense(self, q, r2, imark, nn, dist2):
        """q: particle ID (int) or a position; nn (int32) / dist2 (float64) by ID, updated in place"""
        if np.ndim(q) == 0:
            self.lib().ref_search_ball_dense(self.h, int(q), None, r2, imark, _i(nn), _d(dist2))
        else:
            x = _f64(q)
            self.lib().ref_search_ball_dense(self.h, -1, _d(x), r2, imark, _i(nn), _d(dist2))

    def search_criterion_dense(self, qid, crit, params, imark, nn, dist2):
        pr = _f64(params).copy()
        self.lib().ref_search_criterion_dense(self.h, crit, _d(pr), int(qid), imark, _i(nn), _d(dist2))

    def find_leaf(self, q):
        ids = np.zeros(4096, dtype=np.int32)
        if np.ndim(q) == 0:
            c = self.lib().ref_find_leaf(self.h, int(q), None, _i(ids), len(ids))
        else:
            x = _f64(q)
            c = self.lib().ref_find_leaf(self.h, -1, _d(x), _i(ids), len(ids))
        return np.sort(ids[:c])

    def knn_filtered(self, k, crit=-1, params=None, q0=0, q1=None, x=None, v=None):
        """FindNearestCheck (crit < 0; Particle::type != 0 excluded, see set_types) / FindNearestCriterion (crit 0 | 2) for the
        particle IDs q0..q1, or for the points x (velocities v)"""
        pr = _f64(np.zeros(10) if params is None else params).copy()
        x, v = _f64(x), _f64(v)
        q1 = self.n if q1 is None else q1
        rows = len(x) if x is not None else q1 - q0
        ids = np.zeros((rows, k), dtype=np.int32)
        d2 = np.zeros((rows, k))
        self.lib().ref_knn_filtered(self.h, crit, _d(pr), k, q0, q1, rows, _d(x), _d(v), _i(ids), _d(d2))
        return ids, d2

    def calc_smooth_vel(self, k):
        """CalcDensity(k), CalcSmoothVel(k), CalcSmoothVelDisp(smvel, k): (rho, smvel (n,3), smdisp (n,3,3)) by ID"""
        rho, sv, sd = np.zeros(self.n), np.zeros((self.n, 3)), np.zeros((self.n, 3, 3))
        self.lib().ref_calc_smooth_vel(self.h, k, _d(rho), _d(sv), _d(sd))
        return rho, sv, sd

    def calc_smooth_higher(self, k):
        """CalcSmoothVelSkew / CalcSmoothVelKurtosis after density, mean velocity and dispersion: (skew (n,3), kurtosis (n,3)) by ID"""
        sk, ku = np.zeros((self.n, 3)), np.zeros((self.n, 3))
        self.lib().ref_calc_smooth_higher(self.h, k, _d(sk), _d(ku))
        return sk, ku

    def dump_cuts(self):
        cap = self.n + 64
        ids, dims = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32)
        vals, lmax = np.zeros(cap), np.zeros(cap)
        m = self.lib().ref_dump_cuts(self.h, _i(ids), _i(dims), _d(vals), _d(lmax), cap)
        return ids[:m], dims[:m], vals[:m], lmax[:m]

    def fof(self, fdist, minnum=8, order=0):
        g = np.zeros(self.n, dtype=np.int32)
        ng = C.c_long(0)
        self.last_seconds = self.lib().ref_fof(self.h, fdist, minnum, order, _i(g), C.byref(ng))
        return g, ng.value

    def fof_criterion(self, crit, params, minnum=8, order=0):
        params = _f64(params).copy()
        g = np.zeros(self.n, dtype=np.int32)
        ng = C.c_long(0)
        self.last_seconds = self.lib().ref_fof_criterion(self.h, crit, _d(params), minnum, order, _i(g), C.byref(ng))
        return g, ng.value

    def set_types(self, type_by_id):
        """Particle::type by ID; the checked FOF entry points treat type != 0 as FOFcheckfunc() != 0"""
        t = np.ascontiguousarray(type_by_id, dtype=np.int32)
        self.lib().ref_set_types(self.h, _i(t))

    def fof_checked(self, which, crit, params, minnum=8, order=0):
        """which: 0 FOF(params[0]) with ipcheckflag, 1 FOFCriterion with ipcheckflag, 2 FOFCriterionSetBasisForLinks"""
        params = _f64(params).copy()
        g = np.zeros(self.n, dtype=np.int32)
        ng = C.c_long(0)
        self.last_seconds = self.lib().ref_fof_checked(self.h, which, crit, _d(params), minnum, order, _i(g), C.byref(ng))
        return g, ng.value

    def dump_nodes(self):
        cap = 4 * self.n // 8 + 64
        a = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
        m = self.lib().ref_dump_nodes(self.h, _i(a[0]), _i(a[1]), _i(a[2]), _i(a[3]), cap)
        return [x[:m] for x in a]

    @classmethod
    def scale_phase(cls, pos, vel, xs, vs):
        pos, vel = _f64(pos).copy(), _f64(vel).copy()
        cls.lib().ref_scale_phase(len(pos), _d(pos), _d(vel), xs, vs)
        return pos, vel


def canonical_groups(g):
    """Relabel a FOF group array so label = smallest member ID + 1 (0 stays 0)."""
    g = np.asarray(g)
    out = np.zeros_like(g)
    idx = np.nonzero(g > 0)[0]
    if len(idx) == 0:
        return out
    first = {}
    labs = g[idx]
    order = np.argsort(labs, kind="stable")
    sl = labs[order]
    starts = np.r_[0, np.nonzero(np.diff(sl))[0] + 1]
    mins = np.minimum.reduceat(idx[order], starts)
    lut = dict(zip(sl[starts].tolist(), mins.tolist()))
    out[idx] = np.array([lut[v] for v in labs.tolist()], dtype=g.dtype) + 1
    return out
