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
ALG_BYTES = {"knn_density": 24, "veldensity": 36, "fof3d": 24, "build": 36}


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


def workload_name(ng, nh):
    return "clustered periodic box %d^3 per GPU (ZA lattice + %d Plummer halos), KDTree bucket=16, CalcDensity(%d): kNN + SPH density" % (ng, nh, K_NN)


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


def cpu_reference_leg(pos, vel, mass, k, steps=1, warmup=0, fof_ll=None):
    """The reference's CPU path on the host cores: full-host OpenMP kNN-density (BASELINE.md section 3 variant ii:
    omp-parallel loop over FindNearestPos + the CalcDensity accumulation) through oracle/_ref when it is present
    ('reference'), else the brute-force port ('port')."""
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
        fof_s = None
        if fof_ll is not None:
            R.fof(fof_ll, 20, 1)               # the library's FOF is serial (KDFOF.cxx:70-107): "full host" == 1 core
            fof_s = R.last_seconds
        R.close()
        return {"kind": "reference", "cores": cores, "seconds": ts, "n": n, "build_seconds": build_s, "fof_seconds": fof_s}
    P = pyoracle.Port()
    m = min(n, 20000)
    t0 = time.time()
    P.density(pos[:m], mass[:m], k)
    return {"kind": "port", "cores": os.cpu_count(), "seconds": [time.time() - t0], "n": m, "build_seconds": 0.0, "fof_seconds": None}


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
        "config": {"workload": workload_name(ng, max(8, min(8192, ng ** 3 // 16384))), "particles_per_gpu": ng ** 3, "k": K_NN},
        "cpu_baseline": {"value": val, "unit": "particles/s", "cores": leg["cores"], "kind": leg["kind"], "sample": sample,
                         "build_seconds_sample": leg["build_seconds"]},
        "e2e": {"value": val, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    n = ng ** 3
    nh = max(8, min(8192, n // 16384))
    pos, vel, mass = clustered_box(ng, seed=2025 + 10 * rank, nhalo=nh, device="cuda")
    period = np.ones(3)
    peak, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def progress(msg):
        if os.environ.get("BENCH_VERBOSE"):
            print("[bench rank %d %.1fs] %s" % (rank, time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    if world > 1:
        from nbodylib_b200.sharded import ShardedTree
        tree = ShardedTree(pos, vel, mass, period=period, rank=rank, world=world)
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
    ms_step = dev_s * 1e3 / args.steps
    value = n * world / (dev_s / args.steps)

    # ---- other stages of the same resident tree (rank-local), each K steps ------------------------------------
    extra = {}
    if world == 1:
        g = torch.empty(n, dtype=torch.int32, device="cuda")
        ts = []
        for _ in range(args.steps):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            _, ngroups = tree.FOF(0.2 / ng, 20, 1, out=g)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t1)
        extra["fof3d_particles_per_s"] = n / float(np.mean(ts))
        extra["fof3d_ms"] = float(np.mean(ts)) * 1e3
        extra["fof3d_link_kernel_ms"] = tree.info.last_kernel_ms
        extra["fof3d_groups"] = int(ngroups)
        extra["fof3d_hbm_frac"] = n * ALG_BYTES["fof3d"] / (tree.info.last_kernel_ms * 1e-3) / 1e9 / peak
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
        torch.cuda.synchronize(); t1 = time.perf_counter()
        g6, ng6 = tree.FOFCriterion(2, params, 20, 1)
        torch.cuda.synchronize()
        extra["fof6d_particles_per_s"] = n / (time.perf_counter() - t1)
        extra["fof6d_link_kernel_ms"] = tree.info.last_kernel_ms
        extra["fof6d_groups"] = int(ng6)
        extra["fof6d_note"] = "FOFCriterion(FOF6d), host group array returned (includes 0.5 GB D2H)"
        del g6
        del g
        # tree build from device-resident arrays, K times (the first build of the process also pays for growing the memory pool)
        bms = []
        for _ in range(max(2, args.steps)):
            with KDTree(pos, vel, mass, Period=period, device=local) as tb:
                bms.append(tb.info.build_ms)
        extra["build_ms"] = float(np.mean(bms[1:]))
        extra["build_ms_first"] = float(info.build_ms)
        extra["build_particles_per_s"] = n / (extra["build_ms"] * 1e-3)
        extra["build_hbm_frac"] = n * ALG_BYTES["build"] / (extra["build_ms"] * 1e-3) / 1e9 / peak
    if world > 1:
        extra["sharded_rank0"] = dict(tree.stats)
    if world > 1 and os.environ.get("BENCH_SHARDED_FOF"):
        # BASELINE config 5: 3D FOF of the whole slab-sharded periodic box (halo exchange, local union-find, cross-slab merge).
        # Off by default: a rank-local failure between two collectives would hang the other ranks.
        try:
            tree.close_density()
            progress("sharded FOF starts")
            barrier(); t1 = time.perf_counter()
            gfof, ngl = tree.FOF(0.2 / ng, 20, 1)
            barrier()
            tf = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            extra["fof3d_particles_per_s"] = n * world / float(tf.item())
            extra["fof3d_ms"] = float(tf.item()) * 1e3
            extra["fof3d_groups"] = int(ngl)
            extra["fof3d_note"] = "ShardedTree.FOF: halo exchange + local tree build + union-find + cross-slab merge, all inside the timed region"
            del gfof
            progress("sharded FOF done")
        except Exception as ex:  # the headline line must survive a failure of this extra
            extra["fof3d_error"] = repr(ex)[:300]
    tree.close()

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------
    e2e = None
    if not args.no_e2e and world > 1:
        # every rank: pinned host arrays -> its GPU -> slab-sharded tree (halo exchange, local build) -> CalcDensity -> rho on the host
        hp, hm = (x.cpu().pin_memory() for x in (pos, mass))
        out = torch.empty(n, dtype=torch.float64).pin_memory()
        del pos, vel, mass
        ts = []
        for it in range(1 + max(1, args.steps // 2)):
            barrier(); t1 = time.perf_counter()
            dp, dm = hp.to("cuda", non_blocking=True), hm.to("cuda", non_blocking=True)
            st = ShardedTree(dp, None, dm, period=period, rank=rank, world=world)
            r = st.CalcDensity(K_NN)
            out.copy_(r)
            barrier()
            tt = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            st.close()
            del st, dp, dm, r
            if it > 0:
                ts.append(float(tt.item()))
        e2e = {"value": n * world / float(np.mean(ts)), "unit": "particles/s", "h2d_bytes_per_step": int((hp.numel() * hp.element_size() + hm.numel() * hm.element_size()) * world),
               "d2h_bytes_per_step": int(out.numel() * 8 * world), "ms_per_step": float(np.mean(ts)) * 1e3,
               "includes": "per rank: H2D of pos/mass (fp32, pinned host arrays), halo exchange, tree builds (owned + halo), CalcDensity(64) with scatter return, D2H of rho (fp64, pinned); max over ranks"}
    if not args.no_e2e and world == 1:
        hp, hv, hm = (x.cpu().pin_memory().numpy() for x in (pos, vel, mass))
        out = torch.empty(n, dtype=torch.float64).pin_memory().numpy()      # the caller's (pinned) result buffer
        ts = []
        for it in range(1 + max(1, args.steps // 2)):
            t1 = time.perf_counter()
            with KDTree(hp, hv, hm, Period=period, device=local) as t2:
                t2.CalcDensity(K_NN, out=out)
            if it > 0:
                ts.append(time.perf_counter() - t1)
        e2e = {"value": n / float(np.mean(ts)), "unit": "particles/s", "h2d_bytes_per_step": int(hp.nbytes + hv.nbytes + hm.nbytes),
               "d2h_bytes_per_step": int(out.nbytes), "ms_per_step": float(np.mean(ts)) * 1e3,
               "includes": "H2D of pos/vel/mass (fp32, pinned host arrays), tree build, CalcDensity(64), D2H of rho (fp64, pinned host array)"}

    # ---- CPU baseline on the host cores (rank 0, N=1 only) -------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        frac = sample_fraction(ng)
        sp, sv, sm = sample_subvolume(pos, vel, mass, frac)
        leg = cpu_reference_leg(sp, sv, sm, K_NN, fof_ll=0.2 / ng)
        dt = float(np.mean(leg["seconds"]))
        if leg.get("fof_seconds"):
            extra["cpu_fof3d_particles_per_s"] = leg["n"] / leg["fof_seconds"]
            extra["cpu_fof3d_note"] = "reference KDTree::FOF (serial in the library) on the same sub-cube sample, non periodic"
            extra["cpu_build_particles_per_s"] = leg["n"] / leg["build_seconds"] if leg["build_seconds"] else None
        cpu = {"value": leg["n"] / dt, "unit": "particles/s", "cores": leg["cores"], "kind": leg["kind"],
               "sample": "all %d particles of the sub-cube [0,%.2f)^3 of the same box, full-host OpenMP kNN(k=%d)+density accumulation (BASELINE.md 3, variant ii), tree built over the sample only" % (leg["n"], frac, K_NN),
               "seconds": dt}

    if rank == 0:
        kms = float(np.mean(kernel_ms))
        achieved = n * ALG_BYTES["knn_density"] / (kms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "knn_density_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_particle", 0) * n
        line = {
            "metric": "knn_density_particles_per_s", "value": value, "unit": "particles/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(ng, nh),
                       "particles_per_gpu": n, "k": K_NN, "storage": "fp32 coordinates (exact), fp64 distance arithmetic",
                       "l2": "inputs (%.1f GB) exceed L2, no flush needed" % (n * 16 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "knn_sl_kernel<float,6,false>", "kernel_ms": kms,
                         "algorithmic_bytes_per_particle": ALG_BYTES["knn_density"],
                         "note": "issue-slot bound tree traversal, not HBM bound: see DESIGN.md section 4; traffic = measured DRAM bytes of one launch (ncu), dominated by the insertion log",
                         "issue_active_pct_ncu": 65.7, "warp_instructions_per_particle_ncu": 2404,
                         "ncu_source": "profiles/r1_07_knn_select_log_512cube_k64.txt (full capture of this kernel body), profiles/r1_09_launches_bench_512cube_final_summary.txt (launch list of this command)"},
            "timer": "CUDA events around the K steps (barrier + device synchronize on both sides), max over ranks; wall_ms_per_step = host clock around the same region; library_ms_per_step = the library's own CUDA events around each call on its stream",
            "wall_ms_per_step": wall * 1e3 / args.steps, "library_ms_per_step": float(np.mean(call_ms)),
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
