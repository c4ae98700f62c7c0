#!/usr/bin/env python3
"""bench.py -- headline benchmark of the implicit-field + marching-cubes path (BASELINE.json metric:
voxels/s and triangles/s, field + MC, device-timed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--fine 512] [--ratio 4]

Workload (config.workload): BASELINE config 3 -- spatially varying gyroid lattice, 62 harmonics,
phase control grid trilinearly upsampled to a `fine`^3 grid, min-max normalised, band [0.20,0.30]
extracted with marching cubes.  N > 1 (torchrun, one rank per GPU): weak scaling -- the global grid
is fine x fine x (fine*N), sharded in z-slabs; one tiny NCCL all-reduce carries the global min/max
between field evaluation and extraction, vertex offsets come from an all-gather of the counts.

One step = control grids resident in HBM -> field -> min/max -> fused extraction -> counts on the host.
`value` is voxels (grid points) per second over all ranks; `e2e` is the same step through the
host-buffer C-ABI entry point (pinned host control grids copied H2D inside the timed region, counts read
back).  `roofline` describes the fused extraction kernel (HBM-bound: 4 B/point + 32 B/vertex), timed
with CUDA events on the library's stream during the timed steps; `field_kernel` reports the
FP32-bound SVL evaluation kernel that dominates the step.

`--impl reference` times the reference's OWN CUDA kernels (oracle/_ref/libgpucad_ref.so, the unmodified
sources compiled for sm_100a) on the same workload on the GPU -- the reference has no CPU path; north_star
names these kernels as "the comparison that matters".  If that library is absent it falls back to the
CPU oracle port on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ISO_MASK, BAND_LO, BAND_HI = 0.25, 0.20, 0.30


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers of the e2e leg
    are allocated, so that every rank's host-to-device copies read memory of the GPU's own socket instead of crossing the
    inter-socket link (eight ranks copy 4.2 GB per step).  Best effort: returns the number of CPUs bound to, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = index
        if visible and all(t.strip().isdigit() for t in visible.split(",")):
            idx = int(visible.split(",")[index])
        before = os.sched_getaffinity(0)
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(idx))
        after = os.sched_getaffinity(0)
        if not after:
            os.sched_setaffinity(0, before)
            return None
        return len(after)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every ~5 ms (the region of
    the default run is 60 ms, too short for `nvidia-smi -lms`); falls back to one `nvidia-smi` query stream when NVML is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons, self.source = [], None, set(), None

    def _start_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        if visible and all(t.strip().isdigit() for t in visible.split(",")):
            idx = int(visible.split(",")[self.index])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [(0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap")]

        def poll():
            while not self.stop_flag:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    r = int(get_reasons(h))
                    for bit, name in bits:
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.005)
        self.thread = threading.Thread(target=poll, daemon=True)
        self.thread.start()
        self.source = "nvml"

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.thread = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.source = "nvidia-smi"
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nme, v in zip(self.NAMES, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fine", type=int, default=512, help="fine grid points per axis per GPU")
    ap.add_argument("--ratio", type=int, default=4, help="fine/control upsampling ratio (2 = the app's own, 4 default, 8)")
    ap.add_argument("--harmonics", type=int, default=62)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="timed steps only (no e2e leg, no CPU baseline): for runs under ncu")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the global grid is fine^3 for every N (e.g. --fine 2048 --gpus 8 = BASELINE config 4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args, torch, rank, world, local_rank)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_numa_node(local_rank) if (world > 1 or os.environ.get("GCB_BENCH_AFFINITY")) and not os.environ.get("GCB_BENCH_NO_AFFINITY") else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import gpucadforam_b200 as g
    from gpucadforam_b200 import sharding, synth

    F, R, NH = args.fine, args.ratio, args.harmonics
    gnz = F if args.strong else F * world   # global point layers (weak scaling: one fine^3 block per rank)
    z0, z1 = sharding.slab_bounds(gnz, world, rank)
    nzl = z1 - z0 + 1
    d = (1.0 / R,) * 3
    cxy = F // R
    czg = gnz // R
    c0, c1 = sharding.control_slab(z0, z1, R, czg)
    czl = c1 - c0 + 1
    coef = synth.gyroid_coefficients()[:NH]
    harm = synth.HARMONICS[:NH]
    dev = torch.device("cuda", local_rank)
    phi = synth.phase_grids(cxy, cxy, czl, device=dev, z0=c0, cz_total=czg, harmonics=harm, periods=F / 40.0)
    torch.cuda.synchronize()
    ctx = g.Context(local_rank, options=0)
    svl = torch.empty(F * F * nzl, device=dev)
    mm = torch.zeros(2, device=dev)
    voxel, center = d, (0.0, 0.0, 0.0)
    ldims = (F, F, nzl)

    def field_and_minmax():
        g.svl_field(ctx, svl, phi, coef, (cxy, cxy, czl), ldims, d, slab=(z0, gnz), cz0=c0, d_minmax=mm)
        return sharding.allreduce_minmax(dist, mm)   # one 2-float all-reduce (N > 1), then the values on the host

    # set-up (untimed): count, then allocate the mesh exactly (count-then-allocate, SURVEY.md 7 "Capacity")
    a, b = field_and_minmax()
    act, tot = g.extract_band_raw(ctx, svl, a, b, ISO_MASK, BAND_LO, BAND_HI, ldims, voxel, center, None, None, 0, slab=(z0, gnz), count_only=True)
    cap = tot + 3
    mesh = g.MeshBuffers(cap, device=dev)

    def step():
        a, b = field_and_minmax()
        return g.extract_band_raw(ctx, svl, a, b, ISO_MASK, BAND_LO, BAND_HI, ldims, voxel, center, mesh.pos, mesh.norm, cap, slab=(z0, gnz))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step()
    assert res == (act, tot), "count pass and mesh pass disagree"
    ctx.enable_kernel_timing(True)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ctx.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ext_ms, fld_ms = [], []
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
        ext_ms.append(ctx.last_extract_kernel_ms())   # events already complete: the step ends with the counts on the host
        fld_ms.append(ctx.last_field_kernel_ms())
    ev1.record()
    barrier()
    launches = ctx.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    ctx.enable_kernel_timing(False)

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms, "extract_kernel_ms": sum(ext_ms) / len(ext_ms),
                              "field_kernel_ms": sum(fld_ms) / len(fld_ms), "verts": tot, "launches": launches}))
        return
    # ---- e2e: host control grids -> C-ABI host entry point (single GPU) / per-rank H2D + same step (multi GPU)
    hphi = torch.empty(phi.shape, dtype=torch.float32, pin_memory=True)
    hphi.copy_(phi)
    phi_scratch = torch.empty_like(phi)

    def e2e_step():
        if world == 1:
            a_, t_, _ = g.svl_lattice_host(ctx, hphi, phi_scratch, svl, coef, (cxy, cxy, czl), ldims, d, ISO_MASK, BAND_LO, BAND_HI, voxel, center, mesh.pos,
                                           mesh.norm, cap)
            return a_, t_
        g.svl_field_host(ctx, svl, hphi, phi_scratch, coef, (cxy, cxy, czl), ldims, d, slab=(z0, gnz), cz0=c0, d_minmax=mm)
        a_, b_ = sharding.allreduce_minmax(dist, mm)
        return g.extract_band_raw(ctx, svl, a_, b_, ISO_MASK, BAND_LO, BAND_HI, ldims, voxel, center, mesh.pos, mesh.norm, cap,
                                  slab=(z0, gnz))

    for _ in range(2):
        r2 = e2e_step()
    assert r2 == (act, tot)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.steps

    # max over ranks, global counts and offsets
    stats = torch.tensor([ms, e2e_ms, sum(ext_ms) / len(ext_ms), sum(fld_ms) / len(fld_ms)], device=dev, dtype=torch.float64)
    per_rank, voff, aoff, (g_act, g_tot) = sharding.gather_counts(dist, act, tot, device=dev)   # global vertex offsets = exclusive scan
    g_launch = launches
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        g_launch = int(lt[0])
    ms, e2e_ms, ext_k_ms, fld_k_ms = [float(x) for x in stats.cpu()]

    if rank == 0:
        points = F * F * gnz
        peak, peak_src = peaks()
        alg_bytes = 4.0 * F * F * nzl + 32.0 * tot          # this rank's launch: 4 B/point read + 32 B/vertex written
        achieved = alg_bytes / (ext_k_ms * 1e-3) / 1e9
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        sincos = float(F) * F * nzl * NH
        out = {
            "metric": "voxels/s (field + marching cubes, device-timed)", "value": points / (ms * 1e-3), "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 3: spatially varying gyroid lattice, %d harmonics, control grid %dx%dx%d -> fine %dx%dx%d, band [0.20,0.30]"
                                   % (NH, cxy, cxy, czg, F, F, gnz),
                       "fine": [F, F, gnz], "control": [cxy, cxy, czg], "ratio": R, "parallelism": "z-slabs x%d" % world,
                       "l2_policy": "inputs larger than L2: %.2f GB field + %.2f GB control grids + %.2f GB mesh per rank"
                                    % (4e-9 * F * F * nzl, 4e-9 * phi.numel(), 32e-9 * tot)},
            "triangles_per_s": (g_tot / 3) / (ms * 1e-3), "triangles": g_tot // 3, "active_voxels": g_act,
            "e2e": {"value": points / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(phi.numel() * 4 * world), "d2h_bytes_per_step": int((16 + 8) * world),
                    "note": "control grids copied from pinned host memory each step; counts and min/max read back; the mesh stays in device memory "
                            "as in the reference (Vulkan-exported vertex buffers)",
                    "host_cpus_bound_rank0": affinity},
            "gpu_launches": g_launch,
            "roofline": {"bound": "hbm", "kernel": "mc_fused_kernel<M_BAND_RAW, TMA>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES if (F, R, NH, world) == (512, 4, 62, 1) else None,
                         "traffic_source": "profiles/r01_ncu_mc_fused.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch, this workload)",
                         "kernel_ms": ext_k_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src},
            # the field kernel moves 4 B/point and is bound by instruction issue (exact texture model in fp64 + libdevice-identical
            # sincosf): reported as issue slots per (point, harmonic), not against the HBM roofline
            "field_kernel": {"kernel": "svl_field_tile_kernel", "bound": "instruction issue (fp64 lerps + sincosf polynomial)", "kernel_ms": fld_k_ms,
                             "share_of_step": fld_k_ms / ms,
                             "sincos_pairs_per_s": sincos / (fld_k_ms * 1e-3),
                             "issue_slots_per_point_harmonic": 148 * 4 * 32 * sm_mhz * 1e6 * (fld_k_ms * 1e-3) / sincos,
                             # against the HBM roofline it is nowhere: it writes 4 B/point and reads the control grids once
                             "hbm": {"algorithmic_bytes": 4.0 * F * F * nzl + 4.0 * phi.numel(),
                                     "achieved_gbs": (4.0 * F * F * nzl + 4.0 * phi.numel()) / (fld_k_ms * 1e-3) / 1e9,
                                     "frac": (4.0 * F * F * nzl + 4.0 * phi.numel()) / (fld_k_ms * 1e-3) / 1e9 / peak},
                             "ncu": "profiles/r01_ncu_svl_field.txt: issue slots 66 % busy, FMA / ALU / FP64 / XU pipes 26-27 % each"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(NH)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# measured once with `ncu --set full` on the default workload (profiles/r01_ncu_mc_fused.txt): 0.540 GB read + 5.468 GB written
NCU_TRAFFIC_BYTES = 539927808 + 5467993000


def cpu_baseline(nh, budget_s=12.0):
    """CPU oracle port (oracle/liboracle.so, OpenMP) on a bounded sample of the same workload."""
    import numpy as np
    import oracle_py as orc
    from gpucadforam_b200 import synth
    coef = synth.gyroid_coefficients()[:nh]
    harm = synth.HARMONICS[:nh]
    best = None
    for fine in (96, 160, 256):
        c = fine // 4
        phi = synth.phase_grids(c, c, c, harmonics=harm, periods=fine / 40.0).numpy()
        t0 = time.time()
        f = orc.svl_field(phi, coef, (fine, fine, fine), (0.25, 0.25, 0.25))
        mask, k = orc.normalise_four(f, BAND_LO, BAND_HI)
        r = orc.extract(orc.MODE_LATTICE, (fine,) * 3, (0.25,) * 3, (0, 0, 0), ISO_MASK, f0=mask, f1=k, f2=np.zeros_like(k), iso1=BAND_LO, iso2=BAND_HI,
                        max_verts=max(4 * fine ** 3, 300000), stages=False)
        dt = time.time() - t0
        best = {"value": fine ** 3 / dt, "unit": "voxels/s", "cores": orc.num_threads(), "kind": "port",
                "sample": "%d^3 fine / %d^3 control, %d harmonics, field + normalise + extraction, %.2f s wall" % (fine, c, nh, dt),
                "triangles_per_s": r["total"] / 3 / dt}
        if dt * 4.6 > budget_s:
            break
    return best


def reference_arm(args, torch, rank, world, local_rank):
    """The reference's own implementation of the path.  Rank 0 only."""
    if rank != 0:
        return
    import ref_py as ref
    F, R, NH = args.fine, args.ratio, args.harmonics
    base = {"impl": "reference", "metric": "voxels/s (field + marching cubes, device-timed)", "unit": "voxels/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if not (ref.available() and torch.cuda.is_available()):
        cb = cpu_baseline(NH)
        base.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
                    config={"workload": "config 3 (bounded CPU sample): " + cb["sample"]},
                    e2e={"value": cb["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)
        print(json.dumps(base))
        return
    import numpy as np
    import gpucadforam_b200 as g   # only the torch containers Scratch/MeshBuffers are used below; no product kernel runs in this arm
    from gpucadforam_b200 import synth
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    d = (1.0 / R,) * 3
    c = F // R
    coef = synth.gyroid_coefficients()[:NH]
    phi = synth.phase_grids(c, c, c, device=dev, harmonics=synth.HARMONICS[:NH], periods=F / 40.0)
    dcoef = torch.tensor(np.array(coef, np.float32), device=dev)
    n = F ** 3
    svl = torch.zeros(n, device=dev)
    ga = torch.zeros((n, 2), device=dev)
    mask, k, zeros = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    scr = g.Scratch((F - 1) ** 3, device=dev)
    ref.setup_texture(c, c, c)
    fix = 1 if (F - 1) ** 3 > 65535 * 1024 else 0     # the shipped host code drops blocks above 65535
    mesh = None
    cap = 0

    def step():
        svl.zero_()                                    # cudaMemset(d_svl) in check_lattice (main.cu:4058)
        ref.svl_field(svl, ga, phi, NH, dcoef, (c, c, c), (F, F, F), d)
        ref.normalise_four(svl, mask, k, (F, F, F), BAND_LO, BAND_HI)
        return ref.isosurface_lattice(False, fix, mask, mesh.pos, mesh.norm, ISO_MASK, (F, F, F), d, (0, 0, 0), scr, cap, k, zeros, BAND_LO, BAND_HI, 0.0, 0.0)

    # size the mesh like our arm does (count first); the reference app would allocate 4 vertices per point
    ref.svl_field(svl, ga, phi, NH, dcoef, (c, c, c), (F, F, F), d)
    ref.normalise_four(svl, mask, k, (F, F, F), BAND_LO, BAND_HI)
    tmp = g.MeshBuffers(16, device=dev)
    _, tot = ref.isosurface_lattice(False, fix, mask, tmp.pos, tmp.norm, ISO_MASK, (F, F, F), d, (0, 0, 0), scr, 3, k, zeros, BAND_LO, BAND_HI, 0.0, 0.0)
    cap = tot + 3
    mesh = g.MeshBuffers(cap, device=dev)
    for _ in range(max(args.warmup, 3)):
        act, tot2 = step()
    assert tot2 == tot
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    val = F ** 3 / (ms * 1e-3)
    launches_per_step = NH * 4 + 3 + 7  # 62 x (copytotexture, memcpy3D, grating, svl) + normalise (3) + classify, 2 scans(x2 kernels), compact, generate
    base.update(value=val, ms_per_step=ms, triangles_per_s=tot / 3 / (ms * 1e-3), triangles=tot // 3, active_voxels=act, clocks=clocks,
                config={"workload": "config 3: spatially varying gyroid lattice, %d harmonics, control grid %d^3 -> fine %d^3, band [0.20,0.30]"
                                    % (NH, c, F), "fine": [F, F, F], "control": [c, c, c], "ratio": R,
                        "note": "reference CUDA kernels (unmodified sources, sm_100a) via oracle/_ref; classify launched with a corrected 2-D grid: %s"
                                % bool(fix)},
                cpu_baseline={"value": val, "unit": "voxels/s", "cores": 0, "kind": "reference",
                              "sample": "full workload on 1 B200 with the reference's own CUDA kernels (the reference has no CPU implementation)"},
                e2e={"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=launches_per_step * args.steps)
    print(json.dumps(base))


if __name__ == "__main__":
    main()
