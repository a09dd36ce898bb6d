#!/usr/bin/env python
"""bench.py -- headline benchmark of the kd-tree hot path (BASELINE.json: kNN-density particles/s and FOF
particles/s on the clustered periodic box).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--ng 512] [--impl reference]

One "step" = one pass of the fused kNN(k=64)+SPH-density kernel (KDTree::CalcDensity(64)) over every particle
of the resident tree; `value` = particles/s with the tree and particles already in HBM.  `e2e` = the same
metric through the C ABI with HOST buffers: host->device copy of pos/vel/mass, tree build, CalcDensity, and the
device->host read of rho, all inside the timed region.  FOF, build and velocity-density throughputs of the same
tree are reported in `extra` (each timed the same way, K steps).  N>1: one process per GPU (torchrun), every
rank owns one slab of an N-times larger periodic box (weak scaling) -- see nbodylib_b200/sharded.py.

`--impl reference` times the reference's own CPU implementation (oracle/_ref, the unmodified NBodylib sources)
with every host thread on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_NN = 64
# global lattice (in units of the per-GPU cube) for N ranks: N * ng^3 particles in all, slabs along x
SHARD_DIMS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
ALG_BYTES = {"knn_density": 24, "veldensity": 36, "fof3d": 24, "fof6d": 36, "build": 36}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def shim_e2e(pos, vel, mass, k, period, reps):
    """KDTree(Particle*) + CalcDensity + ~KDTree on an array of 88-byte reference-layout particles through the C++ shim
    (nbodylib_b200/libnbk_shimbench.so, built from examples/shim_bench.cxx); the first repetition is a warm-up."""
    import ctypes as C
    so = os.path.join(ROOT, "nbodylib_b200", "libnbk_shimbench.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    n = len(pos)
    sec = np.zeros((reps + 1, 4))
    rho_sum = C.c_double(0)
    err = C.create_string_buffer(512)
    per = np.ascontiguousarray(period, dtype=np.float64)
    L.nbk_shim_e2e.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_char_p, C.c_int]
    rc = L.nbk_shim_e2e(pos.ctypes.data, vel.ctypes.data, mass.ctypes.data, n, k, per.ctypes.data, reps + 1, sec.ctypes.data, C.byref(rho_sum), err, 512)
    if rc != 0:
        return {"error": err.value.decode()[:300]}
    t = sec[1:].mean(0)
    return {"unit": "particles/s", "ms_per_step": float(t[3]) * 1e3, "constructor_ms": float(t[0]) * 1e3, "calc_density_ms": float(t[1]) * 1e3,
            "destructor_ms": float(t[2]) * 1e3, "particle_bytes": 88, "host_threads": host_cores(), "rho_sum": rho_sum.value,
            "includes": "NBody::KDTree(Particle*, N, 16, TPHYS, KEPAN, 1000, 0, 0, 0, period) [SetID, strided H2D of the fp64 fields, build, "
                        "array permuted into tree order], CalcDensity(%d) [rho into the particles], ~KDTree [order restored]" % k}


def workload_name(ng, nh, world=1):
    """config.workload of both arms (the reference arm times a bounded sample of the same box on the host cores)"""
    if world == 1:
        return "clustered periodic box %d^3 per GPU (ZA lattice + %d Plummer halos), KDTree bucket=16, CalcDensity(%d): kNN + SPH density" % (ng, nh, K_NN)
    dims = SHARD_DIMS[world]
    return ("one clustered periodic box of %d x %d x %d lattice cells (%d particles, ZA + %d Plummer halos) cut into %d slabs along x, one per GPU, "
            "KDTree bucket=16, CalcDensity(%d): kNN + SPH density" % (dims[0] * ng, dims[1] * ng, dims[2] * ng, world * ng ** 3, nh * world, world, K_NN))


def host_cores():
    """cores this process may use (the affinity mask of the container, not the machine's core count)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def sample_fraction(ng):
    """side of the sub-cube the CPU legs work on: ~16.8 M particles (about 20 s of host work: build + kNN-density + FOF)"""
    f = 1.0
    while (ng * f) ** 3 > 1.7e7 and f > 1.0 / 64:
        f *= 0.5
    return f


def sample_subvolume(pos, vel, mass, frac_side=0.25):
    """bounded CPU sample of the same workload: every particle inside the sub-cube [0, frac_side)^3"""
    sel = (pos[:, 0] < frac_side) & (pos[:, 1] < frac_side) & (pos[:, 2] < frac_side)
    return pos[sel].double().cpu().numpy(), vel[sel].double().cpu().numpy(), mass[sel].double().cpu().numpy()


def cpu_reference_leg(pos, vel, mass, k, steps=1, warmup=0, fof_ll=None, fof6d_params=None, check_queries=0):
    """The reference's CPU path on the host cores: full-host OpenMP kNN-density (BASELINE.md section 3 variant ii:
    omp-parallel loop over FindNearestPos + the CalcDensity accumulation) through oracle/_ref when it is present
    ('reference'), else the brute-force port ('port').  Optionally the library's (serial) FOF / FOFCriterion(FOF6d), and the
    reference's neighbour distances of `check_queries` sample particles + its FOF labels for bench.py's self-check."""
    from oracle import pyoracle
    n = len(pos)
    if pyoracle.have_ref():
        # every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the reference on one thread
        pyoracle.Ref.set_threads(host_cores())
        R = pyoracle.Ref(pos, vel, mass, period=None)     # Calc* never use the period (quirk Q2)
        for _ in range(warmup):
            R.calc_density_omp(k, 0, min(n, 100000), want=False)
        ts = []
        for _ in range(steps):
            R.calc_density_omp(k, want=False)
            ts.append(R.last_seconds)
        cores = pyoracle.Ref.max_threads()
        build_s = R.build_seconds
        out = {"kind": "reference", "cores": cores, "seconds": ts, "n": n, "build_seconds": build_s, "fof_seconds": None, "check": None}
        if check_queries:
            q = np.sort(np.random.default_rng(7).choice(n, min(check_queries, n), replace=False)).astype(np.int32)
            _, d2 = R.knn_particle_list(q, k)
            out["check"] = {"qids": q, "d2": d2}
        if fof_ll is not None:
            g, _ = R.fof(fof_ll, 20, 1)               # the library's FOF is serial (KDFOF.cxx:70-107): "full host" == 1 core
            out["fof_seconds"] = R.last_seconds
            if out["check"] is not None:
                out["check"]["fof"] = g
        R.close()
        if fof6d_params is not None:
            # FOFCriterion walks every leaf its pruning radius touches with the criterion: ~5x the cost of FOF per particle, so a
            # smaller sample (1/8 of the sub-cube) keeps the leg bounded
            half = pos.max() / 2.0
            s6 = (pos[:, 0] < half) & (pos[:, 1] < half) & (pos[:, 2] < half)
            R6 = pyoracle.Ref(pos[s6], vel[s6], mass[s6], period=None)
            R6.fof_criterion(2, fof6d_params, 20, 1)
            out["fof6d_seconds"], out["fof6d_n"] = R6.last_seconds, int(s6.sum())
            R6.close()
        return out
    P = pyoracle.Port()
    m = min(n, 20000)
    t0 = time.time()
    P.density(pos[:m], mass[:m], k)
    return {"kind": "port", "cores": os.cpu_count(), "seconds": [time.time() - t0], "n": m, "build_seconds": 0.0, "fof_seconds": None, "check": None}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    import torch
    from nbodylib_b200.synth import clustered_box
    ng = args.ng
    pos, vel, mass = clustered_box(ng, seed=2025, nhalo=max(8, min(8192, ng ** 3 // 16384)), device="cuda" if torch.cuda.is_available() else "cpu")
    frac = sample_fraction(ng)
    sp, sv, sm = sample_subvolume(pos, vel, mass, frac)
    del pos, vel, mass
    # bounded: the rate is per particle, so a few timed passes are enough (each is ~10 s of full-host work); ~60 s in all
    leg = cpu_reference_leg(sp, sv, sm, K_NN, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    dt = float(np.mean(leg["seconds"]))
    val = leg["n"] / dt
    sample = "all %d particles of the sub-cube [0,%.2f)^3 of the %d^3 clustered box (tree built over the sample only; per-query cost grows ~log N, so this flatters the CPU by ~10%% at 512^3)" % (leg["n"], frac, ng)
    line = {
        "impl": "reference", "metric": "knn_density_particles_per_s", "value": val, "unit": "particles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "timed_passes": len(leg["seconds"]), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(ng, max(8, min(8192, ng ** 3 // 16384)), world), "particles_per_gpu": ng ** 3, "particles_total": world * ng ** 3, "k": K_NN},
        "cpu_baseline": {"value": val, "unit": "particles/s", "cores": leg["cores"], "kind": leg["kind"], "sample": sample,
                         "build_seconds_sample": leg["build_seconds"]},
        "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ng", type=int, default=512, help="particles per dimension PER GPU (512 = BASELINE config 3)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: native libraries that print there (NCCL's version banner, for one) are
    # sent to stderr by pointing file descriptor 1 at it; the line itself goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    # fail fast instead of hanging the caller: a healthy run ends within ~2 minutes at any N
    import signal
    signal.alarm(int(os.environ.get("BENCH_WATCHDOG_S", "900")))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nbodylib_b200 import KDTree
    from nbodylib_b200.synth import clustered_box

    t_start = time.perf_counter()
    ng = args.ng
    nh = max(8, min(8192, ng ** 3 // 16384))
    # N GPUs: ONE periodic box of N * ng^3 particles (8 GPUs at ng = 512: the 1024^3 cube of BASELINE config 5), cut into N
    # slabs along x; every rank generates the same global field and keeps its slab
    dims = SHARD_DIMS[world]
    box = np.array(dims, dtype=np.float64)
    if world > 1:
        # slab faces at the x quantiles (equal particle counts, the load-balanced decomposition a production code uses); NBK_SLABS=width
        # cuts equal-width slabs instead (then the slab holding the largest halo has ~20 % more particles at 8 ranks)
        slabs = "equal_width" if os.environ.get("NBK_SLABS", "count") == "width" else "equal_count"
        pos, vel, mass, edges = clustered_box(ng, seed=2025, nhalo=nh * world, device="cuda", dims=dims, slab=(rank, world, slabs), return_edges=True)
        torch.cuda.empty_cache()
    else:
        pos, vel, mass = clustered_box(ng, seed=2025, nhalo=nh, device="cuda")
    n = int(pos.shape[0])                       # this rank's particles
    n_total = n
    if world > 1:
        tcount = torch.tensor([n], dtype=torch.int64, device="cuda")
        dist.all_reduce(tcount)
        n_total = int(tcount.item())
    period = box.copy()
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def progress(msg):
        if os.environ.get("BENCH_VERBOSE"):
            print("[bench rank %d %.1fs] %s" % (rank, time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    if world > 1:
        # the slab-sharded tree: through the C ABI (include/nbk_sharded.h: exchange + merge in C++ over the library's own NCCL
        # communicator) unless NBK_SHARDED_DRIVER=torch selects the torch.distributed driver of the same algorithm
        from nbodylib_b200 import sharded as _sh
        ShardedTree = _sh.ShardedTree if os.environ.get("NBK_SHARDED_DRIVER", "native") == "torch" else _sh.NativeShardedTree
        tree = ShardedTree(pos, vel, mass, period=period, rank=rank, world=world, box=box, edges=edges)
    else:
        tree = KDTree(pos, vel, mass, Period=period, device=local)
    info = tree.info
    rho = torch.empty(tree.n_owned if world > 1 else n, dtype=torch.float64, device="cuda")

    def step():
        tree.CalcDensity(K_NN, out=rho)

    progress("tree ready")
    for _ in range(args.warmup):
        step()
        progress("warm-up step done")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    kernel_ms, call_ms, launches = [], [], 0
    t0 = time.perf_counter()
    # CUDA events on torch's current stream bracket the K steps.  The library works on its own stream, drains torch's
    # stream before every call and synchronises its own before returning, so everything a step launches (library kernels,
    # the sharded driver's torch ops and NCCL transfers) completes between the two records.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
        i = tree.info
        kernel_ms.append(i.last_kernel_ms)
        call_ms.append(i.last_call_ms)
        launches += int(i.last_launches)
    ev1.record()
    progress("timed steps done")
    barrier()
    wall = time.perf_counter() - t0
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    tmax = torch.tensor([dev_s, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_s, wall = float(tmax[0].item()), float(tmax[1].item())
    clocks = sampler.stop() if sampler else None
    if world > 1:
        # one more (untimed) step with device-synchronised section timers of the sharded driver
        tree.profile = True
        step()
        tree.profile = False
    ms_step = dev_s * 1e3 / args.steps
    value = n_total / (dev_s / args.steps)

    # ---- untimed self-check of the timed step (N = 1): the same CalcDensity through the INDEPENDENT fp64-heap kernel ------
    checked = None
    if world == 1:
        from nbodylib_b200 import set_option
        h1 = torch.empty(n, dtype=torch.float64, device="cuda")
        rho2 = torch.empty(n, dtype=torch.float64, device="cuda")
        h2 = torch.empty(n, dtype=torch.float64, device="cuda")
        tree.CalcDensityInto(K_NN, rho, h1)
        flagged = int(tree.info.last_flagged)
        set_option("knn_exact", 1)
        try:
            tree.CalcDensityInto(K_NN, rho2, h2)
            heap_ms = float(tree.info.last_kernel_ms)
        finally:
            set_option("knn_exact", 0)
        checked = {"against": "knn_exact_kernel (fp64 (d2, index) heap per query, the kernel behind FindNearest; bit-exact against the reference at 1 M in tests/)",
                   "particles": n, "h_mismatches": int((h1 != h2).sum().item()),
                   "rho_max_rel_diff": float(((rho - rho2).abs() / rho2.abs().clamp_min(1e-300)).max().item()),
                   "queries_rerun_by_exact_kernel": flagged, "exact_kernel_ms": heap_ms}
        checked["ok"] = checked["h_mismatches"] == 0 and checked["rho_max_rel_diff"] < 1e-10
        del rho2, h2
        progress("self-check done")

    # ---- other stages of the same resident tree (rank-local), each K steps ------------------------------------
    extra = {}
    rows = {}
    if world == 1:
        g = torch.empty(n, dtype=torch.int32, device="cuda")
        ts, ks = [], []
        for it in range(2 + args.steps):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            _, ngroups = tree.FOF(0.2 / ng, 20, 1, out=g)
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(time.perf_counter() - t1); ks.append(tree.info.last_kernel_ms)
        fof_ms, fof_kms = float(np.mean(ts)) * 1e3, float(np.mean(ks))
        rows["fof3d"] = {"metric": "fof3d_particles_per_s", "value": n / (fof_ms * 1e-3), "unit": "particles/s", "ms_per_step": fof_ms,
                         "call": "KDTree::FOF(0.2 mean spacings, minnum 20, order 1), periodic; group ids stay on the device", "groups": int(ngroups),
                         "roofline": {"bound": "hbm", "kernel": "fof_link3f_kernel", "kernel_ms": fof_kms, "algorithmic_bytes_per_particle": ALG_BYTES["fof3d"],
                                      "achieved": n * ALG_BYTES["fof3d"] / (fof_kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": n * ALG_BYTES["fof3d"] / (fof_kms * 1e-3) / 1e9 / peak, "traffic": None}}
        g_fof3d = g
        ts = []
        for _ in range(max(1, args.steps // 2)):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            tree.CalcVelDensity(K_NN, K_NN, out=rho)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t1)
        extra["veldensity_particles_per_s"] = n / float(np.mean(ts))
        # BASELINE config 4: 6D phase-space FOF with the in-tree criterion (FOFFunc.h:48-55)
        sv2 = float(((vel - vel.mean(0)) ** 2).sum(1).mean().item() / 3.0)
        params = np.zeros(10)
        params[1] = params[6] = (0.2 / ng) ** 2
        params[2] = params[7] = (1.25 ** 2) * sv2
        g6 = torch.empty(n, dtype=torch.int32, device="cuda")
        ts, ks = [], []
        for it in range(1 + max(2, args.steps // 2)):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            _, ng6 = tree.FOFCriterion(2, params, 20, 1, out=g6)
            torch.cuda.synchronize()
            if it >= 1:
                ts.append(time.perf_counter() - t1); ks.append(tree.info.last_kernel_ms)
        f6_ms, f6_kms = float(np.mean(ts)) * 1e3, float(np.mean(ks))
        rows["fof6d"] = {"metric": "fof6d_particles_per_s", "value": n / (f6_ms * 1e-3), "unit": "particles/s", "ms_per_step": f6_ms,
                         "call": "KDTree::FOFCriterion(FOF6d, ll_x = 0.2 spacings, ll_v = 1.25 sigma_v, minnum 20, order 1), periodic; group ids stay on the device",
                         "groups": int(ng6),
                         "roofline": {"bound": "hbm", "kernel": "fof_link_kernel<float> (6D predicate)", "kernel_ms": f6_kms, "algorithmic_bytes_per_particle": ALG_BYTES["fof6d"],
                                      "achieved": n * ALG_BYTES["fof6d"] / (f6_kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": n * ALG_BYTES["fof6d"] / (f6_kms * 1e-3) / 1e9 / peak, "traffic": None}}
        del g6
        # tree build from device-resident arrays, K times (the first build of the process also pays for growing the memory pool)
        bms = []
        for _ in range(max(2, args.steps)):
            with KDTree(pos, vel, mass, Period=period, device=local) as tb:
                bms.append(tb.info.build_ms)
        build_ms = float(np.mean(bms[1:]))
        rows["build"] = {"metric": "build_particles_per_s", "value": n / (build_ms * 1e-3), "unit": "particles/s", "ms_per_step": build_ms,
                         "call": "KDTree(Particle*, N, bucket 16): radix sorts + rank-space level loop + shared-memory small-node kernel, device-resident input",
                         "first_build_ms": float(info.build_ms),
                         "roofline": {"bound": "hbm", "kernel": "whole build (28 launches; v2_scatter_kernel dominates)", "kernel_ms": build_ms,
                                      "algorithmic_bytes_per_particle": ALG_BYTES["build"], "achieved": n * ALG_BYTES["build"] / (build_ms * 1e-3) / 1e9,
                                      "peak": peak, "unit": "GB/s", "frac": n * ALG_BYTES["build"] / (build_ms * 1e-3) / 1e9 / peak, "traffic": None}}
    if world > 1:
        # BASELINE configs 4 / 5 on the slab-sharded box: 3D FOF and 6D FOF (in-tree criterion), K steps each; a step is the
        # whole call on a resident slab tree (local union-find over owned + ghosts, cross-slab merge, global numbering)
        def timed(fn, reps):
            ts = []
            for it in range(1 + reps):
                barrier(); t1 = time.perf_counter()
                out = fn()
                barrier()
                tt = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                if it > 0:
                    ts.append(float(tt.item()))
            return float(np.mean(ts)), out

        tree.close_density()
        torch.cuda.empty_cache()
        try:
            sums = torch.cat([vel.double().sum(0), (vel.double() ** 2).sum(0)])
            dist.all_reduce(sums)
            mean = sums[:3] / n_total
            sv2 = float(((sums[3:] / n_total) - mean ** 2).sum().item() / 3.0)
            params = np.zeros(10)
            params[1] = params[6] = (0.2 / ng) ** 2
            params[2] = params[7] = (1.25 ** 2) * sv2
            dt, (g3, ng3) = timed(lambda: tree.FOF(0.2 / ng, 20, 1), max(2, args.steps // 2))
            rows["fof3d"] = {"metric": "fof3d_particles_per_s", "value": n_total / dt, "unit": "particles/s", "ms_per_step": dt * 1e3, "groups": int(ng3),
                             "call": "ShardedTree.FOF(0.2 mean spacings, minnum 20, order 1), periodic global box: local union-find over owned + ghosts, "
                                     "all-gathered cross-slab edges, device-side union, global numbering; max over ranks",
                             "ghosts_rank0": int(tree.stats.get("ghosts_fof", 0)), "link_kernel_ms_rank0": float(tree.info.last_kernel_ms)}
            del g3
            dt, (g6, ng6) = timed(lambda: tree.FOFCriterion(2, params, 20, 1), max(2, args.steps // 2))
            rows["fof6d"] = {"metric": "fof6d_particles_per_s", "value": n_total / dt, "unit": "particles/s", "ms_per_step": dt * 1e3, "groups": int(ng6),
                             "call": "ShardedTree.FOFCriterion(FOF6d, ll_x = 0.2 spacings, ll_v = 1.25 sigma_v, minnum 20, order 1), periodic global box; "
                                     "velocities travel with the ghosts; max over ranks",
                             "ghosts_rank0": int(tree.stats.get("ghosts_fof", 0)), "link_kernel_ms_rank0": float(tree.info.last_kernel_ms)}
            del g6
        except Exception as ex:  # the headline line must survive a failure of these rows
            extra["fof_error"] = repr(ex)[:300]
        extra["sharded_rank0"] = dict(tree.stats)
        extra["sharded_driver"] = "C ABI (libnbk_sharded.so, NCCL from C++)" if ShardedTree is _sh.NativeShardedTree else "torch.distributed driver (nbodylib_b200/sharded.py)"
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"particles": int(tree.n_owned), "library_ms_per_step": float(np.mean(call_ms))})
        extra["per_rank"] = per_rank
        extra["box"] = list(map(float, box))
        extra["slab_faces_x"] = [float(e) for e in edges]
        extra["decomposition"] = "%d x-slabs, faces at the x quantiles (equal particle counts)" % world if slabs == "equal_count" else "%d x-slabs of equal width" % world
        extra["particles_total"] = n_total
    tree.close()

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------
    e2e = None
    e2e_aos = None
    if not args.no_e2e and world > 1:
        # every rank: pinned host arrays -> its GPU -> slab-sharded tree (halo exchange, local build) -> CalcDensity -> rho on the host
        hp, hm = (x.cpu().pin_memory() for x in (pos, mass))
        out = torch.empty(n, dtype=torch.float64).pin_memory()
        del pos, vel, mass
        ts = []
        for it in range(1 + max(1, args.steps // 2)):
            barrier(); t1 = time.perf_counter()
            if ShardedTree is _sh.NativeShardedTree:            # host pointers straight into the C ABI
                dp = dm = None
                st = ShardedTree(hp, None, hm, period=period, rank=rank, world=world, box=box, device=local, edges=edges)
            else:
                dp, dm = hp.to("cuda", non_blocking=True), hm.to("cuda", non_blocking=True)
                st = ShardedTree(dp, None, dm, period=period, rank=rank, world=world, box=box, edges=edges)
            r = st.CalcDensity(K_NN)
            out.copy_(r)
            barrier()
            tt = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            st.close()
            del st, dp, dm, r
            if it > 0:
                ts.append(float(tt.item()))
        e2e = {"value": n_total / float(np.mean(ts)), "unit": "particles/s", "h2d_bytes_per_step": int(16 * n_total),
               "d2h_bytes_per_step": int(8 * n_total), "ms_per_step": float(np.mean(ts)) * 1e3,
               "includes": "per rank: H2D of pos/mass (fp32, pinned host arrays), halo exchange, tree builds (owned + halo), CalcDensity(64) with scatter return, D2H of rho (fp64, pinned); max over ranks"}
    if not args.no_e2e and world == 1:
        hp, hv, hm = (x.cpu().pin_memory().numpy() for x in (pos, vel, mass))
        out = torch.empty(n, dtype=torch.float64).pin_memory().numpy()      # the caller's (pinned) result buffer
        outg = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
        reps = 1 + max(1, args.steps // 2)

        def e2e_loop(body):
            ts = []
            for it in range(reps):
                t1 = time.perf_counter()
                body()
                if it > 0:
                    ts.append(time.perf_counter() - t1)
            return float(np.mean(ts))

        def e2e_density():
            with KDTree(hp, hv, hm, Period=period, device=local) as t2:
                t2.CalcDensity(K_NN, out=out)

        def e2e_fof3d():
            with KDTree(hp, None, None, Period=period, device=local) as t2:
                t2.FOF(0.2 / ng, 20, 1, out=outg)

        def e2e_fof6d():
            with KDTree(hp, hv, None, Period=period, device=local) as t2:
                t2.FOFCriterion(2, params, 20, 1, out=outg)

        def e2e_build():
            with KDTree(hp, hv, hm, Period=period, device=local):
                pass

        dt = e2e_loop(e2e_density)
        e2e = {"value": n / dt, "unit": "particles/s", "h2d_bytes_per_step": int(hp.nbytes + hv.nbytes + hm.nbytes),
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": dt * 1e3,
               "includes": "H2D of pos/vel/mass (fp32, pinned host arrays), tree build, CalcDensity(64), D2H of rho (fp64, pinned host array)"}
        dt = e2e_loop(e2e_fof3d)
        rows["fof3d"]["e2e"] = {"value": n / dt, "unit": "particles/s", "h2d_bytes_per_step": int(hp.nbytes), "d2h_bytes_per_step": int(outg.nbytes),
                                "ms_per_step": dt * 1e3, "includes": "H2D of pos (fp32, pinned), tree build, FOF, D2H of the group ids (int32, pinned)"}
        dt = e2e_loop(e2e_fof6d)
        rows["fof6d"]["e2e"] = {"value": n / dt, "unit": "particles/s", "h2d_bytes_per_step": int(hp.nbytes + hv.nbytes), "d2h_bytes_per_step": int(outg.nbytes),
                                "ms_per_step": dt * 1e3, "includes": "H2D of pos/vel (fp32, pinned), tree build, FOFCriterion(FOF6d), D2H of the group ids (int32, pinned)"}
        dt = e2e_loop(e2e_build)
        rows["build"]["e2e"] = {"value": n / dt, "unit": "particles/s", "h2d_bytes_per_step": int(hp.nbytes + hv.nbytes + hm.nbytes), "d2h_bytes_per_step": 0,
                                "ms_per_step": dt * 1e3, "includes": "H2D of pos/vel/mass (fp32, pinned), tree build (nothing is read back: the tree stays on the device)"}
        # the drop-in path: 88-byte fp64 Particle[] through the header-only C++ shim (examples/shim_bench.cxx)
        e2e_aos = shim_e2e(hp, hv, hm, K_NN, period, 2)
        if e2e_aos is not None:
            e2e_aos["value"] = n / (e2e_aos["ms_per_step"] * 1e-3)
        del hp, hv, hm, out, outg

    # ---- CPU baseline on the host cores (rank 0, N=1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        frac = sample_fraction(ng)
        sel = (pos[:, 0] < frac) & (pos[:, 1] < frac) & (pos[:, 2] < frac)
        sp, sv, sm = pos[sel].double().cpu().numpy(), vel[sel].double().cpu().numpy(), mass[sel].double().cpu().numpy()
        leg = cpu_reference_leg(sp, sv, sm, K_NN, fof_ll=0.2 / ng, fof6d_params=params, check_queries=200000)
        dt = float(np.mean(leg["seconds"]))
        sample_txt = "all %d particles of the sub-cube [0,%.2f)^3 of the same box, tree built over the sample only" % (leg["n"], frac)
        cpu = {"value": leg["n"] / dt, "unit": "particles/s", "cores": leg["cores"], "kind": leg["kind"],
               "sample": sample_txt + "; full-host OpenMP kNN(k=%d)+density accumulation (BASELINE.md 3, variant ii)" % K_NN, "seconds": dt}
        if leg.get("fof_seconds"):
            rows["fof3d"]["cpu_baseline"] = {"value": leg["n"] / leg["fof_seconds"], "unit": "particles/s", "cores": 1, "kind": leg["kind"],
                                             "sample": sample_txt + "; KDTree::FOF is serial in the library (KDFOF.cxx:70-107), non periodic", "seconds": leg["fof_seconds"]}
        if leg.get("fof6d_seconds"):
            rows["fof6d"]["cpu_baseline"] = {"value": leg["fof6d_n"] / leg["fof6d_seconds"], "unit": "particles/s", "cores": 1, "kind": leg["kind"],
                                             "sample": "the first %d particles of that sample's tree order region [0,%.3f)^3; KDTree::FOFCriterion(FOF6d) is serial in the library, non periodic" % (leg["fof6d_n"], frac / 2),
                                             "seconds": leg["fof6d_seconds"]}
        if leg.get("build_seconds"):
            rows["build"]["cpu_baseline"] = {"value": leg["n"] / leg["build_seconds"], "unit": "particles/s", "cores": leg["cores"], "kind": leg["kind"],
                                             "sample": sample_txt + "; KDTree constructor (OpenMP tasks over subtrees, KDTree.cxx:1015-1053)", "seconds": leg["build_seconds"]}
        # ---- the device results against the reference on the sample, where the sample's answer is the whole box's answer --------
        if checked is not None and leg.get("check") is not None:
            ck = leg["check"]
            sid = torch.nonzero(sel).flatten()
            face = np.minimum(sp, frac - sp).min(1)                     # distance of every sample particle to the sub-cube's faces
            q = ck["qids"]
            rk = np.sqrt(ck["d2"][:, -1])
            inner = rk < face[q]                                         # the reference's k-ball lies inside the sub-cube: its k nearest in the sample are its k nearest in the box
            h_dev = h1[sid[torch.from_numpy(q[inner].astype(np.int64)).cuda()]].cpu().numpy()
            checked["reference_on_sample"] = {
                "queries": int(len(q)), "with_k_ball_inside_the_sample": int(inner.sum()),
                "h_mismatches": int((h_dev != 0.5 * rk[inner]).sum())}
            checked["ok"] = checked["ok"] and checked["reference_on_sample"]["h_mismatches"] == 0 and int(inner.sum()) > 0
            # FOF: a sample group none of whose members lies within one linking length of a sub-cube face is a group of the box
            gs = ck["fof"]
            gd = g_fof3d[sid].cpu().numpy()
            ll = 0.2 / ng
            ngs = int(gs.max())
            near = np.zeros(ngs + 1, dtype=bool)
            near[np.unique(gs[face < ll])] = True
            near[0] = True
            safe = ~near[gs]                                              # members of interior sample groups
            lab_s, lab_d = gs[safe], gd[safe]
            first = np.full(ngs + 1, -1, dtype=np.int64)
            first[lab_s] = lab_d
            same_label = bool(np.array_equal(first[lab_s], lab_d)) and bool((lab_d > 0).all())
            size_s = np.bincount(lab_s, minlength=ngs + 1)
            size_d_all = torch.bincount(g_fof3d.long()).cpu().numpy()
            grp = np.nonzero(size_s)[0]
            same_size = bool(np.array_equal(size_s[grp], size_d_all[first[grp]])) if same_label else False
            checked["reference_fof_on_sample"] = {"interior_groups": int(len(grp)), "members": int(safe.sum()), "one_device_group_each": same_label,
                                                  "sizes_equal": same_size}
            checked["ok"] = checked["ok"] and same_label and same_size and len(grp) > 0

    if rank == 0:
        kms = float(np.mean(kernel_ms))
        achieved = n * ALG_BYTES["knn_density"] / (kms * 1e-3) / 1e9
        traffic, ncu_facts = None, {}
        tp = os.path.join(ROOT, "profiles", "knn_density_traffic.json")
        if os.path.exists(tp):
            ncu_facts = json.load(open(tp))
            traffic = ncu_facts.get("dram_bytes_per_particle", 0) * n
        line = {
            "metric": "knn_density_particles_per_s", "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(ng, nh, world),
                       "particles_per_gpu": n_total // world, "particles_total": n_total, "k": K_NN,
                       "storage": "fp32 coordinates (exact), fp32 screening keys with a certified error band, fp64 distances and sums for the results",
                       "l2": "inputs (%.1f GB) exceed L2, no flush needed" % (n * 16 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "knn_hp_kernel<float,false,32>", "kernel_ms": kms,
                         "algorithmic_bytes_per_particle": ALG_BYTES["knn_density"],
                         "note": "issue-slot bound tree traversal + selection, not HBM bound: see DESIGN.md section 4; traffic = measured DRAM bytes of one launch (ncu)",
                         **ncu_facts},
            "timer": "CUDA events around the K steps (barrier + device synchronize on both sides), max over ranks; wall_ms_per_step = host clock around the same region; library_ms_per_step = the library's own CUDA events around each call on its stream",
            "wall_ms_per_step": wall * 1e3 / args.steps, "library_ms_per_step": float(np.mean(call_ms)),
            "cpu_baseline": cpu, "e2e": e2e, "e2e_aos": e2e_aos, "gpu_launches": launches, "clocks": clocks, "checked": checked, "rows": rows, "extra": extra,
        }
        rt = os.path.join(ROOT, "profiles", "row_traffic.json")
        row_facts = json.load(open(rt)) if os.path.exists(rt) else {}
        for name, r in rows.items():
            if "roofline" in r:
                r["roofline"]["peak_source"] = peak_src
                f = row_facts.get(name)
                if f:       # measured DRAM bytes per particle of the row's kernel(s) (ncu, 256^3 capture) scaled to this run's particle count
                    r["roofline"]["traffic"] = f["dram_bytes_per_particle"] * n
                    r["roofline"]["ncu"] = f
        emit(line)
    if world > 1:
        _sh.NativeShardedTree.shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
