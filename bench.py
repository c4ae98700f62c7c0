#!/usr/bin/env python3
"""bench.py -- headline benchmark of the implicit-field + marching-cubes path (BASELINE.json metric:
voxels/s and triangles/s, field + MC, device-timed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--fine 512] [--ratio 4]

Workload (config.workload): BASELINE config 3 -- spatially varying gyroid lattice, 62 harmonics,
phase control grid trilinearly upsampled to a `fine`^3 grid, min-max normalised, band [0.20,0.30]
extracted with marching cubes.  N > 1 (torchrun, one rank per GPU): weak scaling -- the global grid
is fine x fine x (fine*N), sharded in z-slabs; one tiny NCCL all-reduce carries the global min/max
between field evaluation and extraction, vertex offsets come from an all-gather of the counts.

One step = control grids resident in HBM -> field -> min/max (stays in device memory) -> fused extraction -> counts on the host.
`value` is voxels (grid points) per second over all ranks, DEFAULT library mode (field bit-identical to the reference kernels).
`e2e` is the same workload with HOST control grids, every step copying them H2D inside the timed region and reading its counts back:
on one GPU through the two-deep job pipeline (gcb_svl_lattice_host_submit / _wait; `e2e.blocking_call` = one blocking call per
step), on N > 1 through the same pipeline in two halves per job (gcb_svl_slab_host_submit_field / _submit_extract around the
stream-ordered NCCL all-reduce of the range), with `e2e.h2d_probe` naming the host-side limiter by measurement.
`workflow` (N = 1, both arms) is the whole lattice job with the producer on the device: period field -> phase solve -> field -> mesh.
`roofline` describes the kernel that dominates the step, `roofline_other_kernel` the other one: the SVL field kernel (writes 4 B /
point: nothing to stream, bound by instruction issue -- reported with the HBM figure the contract asks for AND its actual limiter) and
the fused extraction kernel (HBM-bound: 4 B/point + 32 B/vertex); both timed with CUDA events on the library's stream in the timed steps.
Secondary legs in the same line: `fast_field` (GCB_OPT_FAST_FIELD: hardware cosine, stated tolerance), `ratio2` (the app's own upsampling
ratio), `configs` (BASELINE configs 1, 2, 5 at full size; our side -- the reference kernels' side is in the --impl reference line) and
`config4` (the 2048-wide grid of BASELINE config 4, z-slab sharded: the full 2048^3 at N >= 4).

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


def workload_config(F, R, NH, gnz, world, strong):
    """The `config` object of the JSON line: identical in both arms (the driver compares them)."""
    c = F // R
    return {"workload": "config 3: spatially varying gyroid lattice, %d harmonics, control grid %dx%dx%d -> fine %dx%dx%d, band [0.20,0.30]"
                        % (NH, c, c, gnz // R, F, F, gnz),
            "fine": [F, F, gnz], "control": [c, c, gnz // R], "ratio": R, "harmonics": NH, "parallelism": "z-slabs x%d" % world,
            "scaling_mode": "strong" if strong else "weak",
            "l2_policy": "inputs larger than L2: field, control grids and mesh are each far above 126 MB per rank and are streamed once per step"}


class SvlLeg:
    """One SVL-lattice workload (configs 3 / 4) on this rank's z-slab: set-up, device-timed steps, end-to-end steps."""

    def __init__(self, env, F, R, NH, gnz, fast=False, spectrum="gyroid", cuts=None):
        torch, g, sharding, synth = env["torch"], env["g"], env["sharding"], env["synth"]
        self.env, self.F, self.R, self.NH, self.gnz, self.fast = env, F, R, NH, gnz, fast
        rank, world, dev = env["rank"], env["world"], env["dev"]
        sharding.validate_slabs(gnz, world)
        self.cuts = list(cuts) if cuts else [sharding.slab_bounds(gnz, world, r)[0] for r in range(world)] + [gnz - 1]
        self.z0, self.z1 = self.cuts[rank], self.cuts[rank + 1]
        self.nzl = self.z1 - self.z0 + 1
        self.d = (1.0 / R,) * 3
        self.cxy, self.czg = F // R, gnz // R
        self.c0, c1 = sharding.control_slab(self.z0, self.z1, R, self.czg)
        self.czl = c1 - self.c0 + 1
        self.coef = (synth.gyroid_coefficients() if spectrum == "gyroid" else synth.schwarz_p_coefficients())[:NH]
        self.phi = synth.phase_grids(self.cxy, self.cxy, self.czl, device=dev, z0=self.c0, cz_total=self.czg, harmonics=synth.HARMONICS[:NH], periods=F / 40.0)
        torch.cuda.synchronize()
        self.ctx = env["ctx"]
        self.opts = g._capi.GCB_OPT_FAST_FIELD if fast else 0
        self.svl = torch.empty(F * F * self.nzl, device=dev)
        self.mm = torch.zeros(2, device=dev)
        self.ldims = (F, F, self.nzl)
        self.ctx.set_options(self.opts)
        a, b = self._field_and_minmax()
        self.act, self.tot = g.extract_band_raw(self.ctx, self.svl, a, b, ISO_MASK, BAND_LO, BAND_HI, self.ldims, self.d, (0.0, 0.0, 0.0), None, None, 0,
                                                slab=(self.z0, gnz), count_only=True)
        self.cap = self.tot + 3   # count-then-allocate (SURVEY.md 7 "Capacity"); +3: the reference's `index < maxVerts - 3` guard
        self.mesh = g.MeshBuffers(self.cap, device=dev)
        self.hphi = None

    def _field_and_minmax(self):
        g, sharding = self.env["g"], self.env["sharding"]
        g.svl_field(self.ctx, self.svl, self.phi, self.coef, (self.cxy, self.cxy, self.czl), self.ldims, self.d, slab=(self.z0, self.gnz), cz0=self.c0, d_minmax=self.mm)
        return sharding.allreduce_minmax(self.env["dist"], self.mm)   # set-up only: the values on the host

    def step(self):
        """field -> min/max (stays on the device; N > 1: one 2-float NCCL all-reduce) -> extraction reading the range in place -> counts"""
        g, sharding = self.env["g"], self.env["sharding"]
        g.svl_field(self.ctx, self.svl, self.phi, self.coef, (self.cxy, self.cxy, self.czl), self.ldims, self.d, slab=(self.z0, self.gnz), cz0=self.c0, d_minmax=self.mm)
        ab = sharding.allreduce_minmax_device(self.env["dist"], self.mm)
        return g.extract_band_raw_dev(self.ctx, self.svl, ab, ISO_MASK, BAND_LO, BAND_HI, self.ldims, self.d, (0.0, 0.0, 0.0), self.mesh.pos, self.mesh.norm, self.cap,
                                      slab=(self.z0, self.gnz))

    def _host_buffers(self):
        torch = self.env["torch"]
        if self.hphi is None:
            self.hphi = torch.empty(self.phi.shape, dtype=torch.float32, pin_memory=True)
            self.hphi.copy_(self.phi)
            self.phi_scratch = torch.empty_like(self.phi)
        return (self.cxy, self.cxy, self.czl)

    def e2e_step(self):
        """one blocking call with HOST control grids (H2D inside), counts read back"""
        g, sharding = self.env["g"], self.env["sharding"]
        cd = self._host_buffers()
        if self.env["world"] == 1:
            a_, t_, _ = g.svl_lattice_host(self.ctx, self.hphi, self.phi_scratch, self.svl, self.coef, cd, self.ldims, self.d, ISO_MASK, BAND_LO, BAND_HI, self.d,
                                           (0.0, 0.0, 0.0), self.mesh.pos, self.mesh.norm, self.cap)
            return a_, t_
        g.svl_field_host(self.ctx, self.svl, self.hphi, self.phi_scratch, self.coef, cd, self.ldims, self.d, slab=(self.z0, self.gnz), cz0=self.c0, d_minmax=self.mm)
        ab = sharding.allreduce_minmax_device(self.env["dist"], self.mm)
        return g.extract_band_raw_dev(self.ctx, self.svl, ab, ISO_MASK, BAND_LO, BAND_HI, self.ldims, self.d, (0.0, 0.0, 0.0), self.mesh.pos, self.mesh.norm, self.cap,
                                      slab=(self.z0, self.gnz))

    def ctx_on_torch_stream(self):
        """The sharded pipeline orders the library's kernels and NCCL's all-reduce on ONE stream: the context was created on the legacy
        default stream, which is torch's current stream unless somebody switched it."""
        return self.env["torch"].cuda.current_stream().cuda_stream == 0

    def e2e_pipelined(self, steps):
        """`steps` jobs through the two-deep job pipeline: every job copies its control grids from pinned host memory and has its counts read
        back; job i+1's copies overlap job i's kernels.  One rank: gcb_svl_lattice_host_submit / _wait.  Several ranks: the job is enqueued in
        two halves (gcb_svl_slab_host_submit_field / _submit_extract) with the 2-float NCCL all-reduce of the range between them, all
        stream-ordered -- the host only waits for the counts of the job before last.  Returns the results."""
        g, torch, sharding = self.env["g"], self.env["torch"], self.env["sharding"]
        cd = self._host_buffers()
        if not hasattr(self, "phi_scratch2") or self.phi_scratch2 is None:
            self.phi_scratch2 = torch.empty_like(self.phi)
            self.mm2 = [torch.zeros(2, device=self.env["dev"]) for _ in range(2)]
        scr = [self.phi_scratch, self.phi_scratch2]
        ab = [None, None]   # keeps the reduced range of each slot's job alive until that job is done

        def submit(i):
            if self.env["world"] == 1:
                g.svl_lattice_host_submit(self.ctx, i % 2, self.hphi, scr[i % 2], self.svl, self.coef, cd, self.ldims, self.d, ISO_MASK, BAND_LO, BAND_HI, self.d,
                                          (0.0, 0.0, 0.0), self.mesh.pos, self.mesh.norm, self.cap)
                return
            sl = i % 2
            g.svl_slab_host_submit_field(self.ctx, sl, self.hphi, scr[sl], self.svl, self.coef, cd, self.ldims, self.d, (self.z0, self.gnz), self.c0, self.mm2[sl])
            ab[sl] = sharding.allreduce_minmax_device(self.env["dist"], self.mm2[sl])
            g.svl_slab_host_submit_extract(self.ctx, sl, self.svl, ab[sl], ISO_MASK, BAND_LO, BAND_HI, self.ldims, self.d, (0.0, 0.0, 0.0), self.mesh.pos,
                                           self.mesh.norm, self.cap, slab=(self.z0, self.gnz))
        res = []
        submit(0)
        for i in range(1, steps):
            submit(i)
            res.append(g.svl_lattice_host_wait(self.ctx, (i - 1) % 2)[:2])
        res.append(g.svl_lattice_host_wait(self.ctx, (steps - 1) % 2)[:2])
        return res

    def run(self, steps, warmup, e2e=True, sampler=None):
        """Device-timed steps (CUDA events on the current stream, barrier + synchronize on both sides), then the e2e steps.
        Returns this rank's numbers; reduce() turns them into the whole-job line."""
        torch, env = self.env["torch"], self.env
        self.ctx.set_options(self.opts)
        for _ in range(warmup):
            res = self.step()
        assert res == (self.act, self.tot), "count pass and mesh pass disagree"
        self.ctx.enable_kernel_timing(True)
        env["barrier"]()
        if sampler is not None:
            sampler.start()
        self.ctx.reset_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ext_ms, fld_ms = [], []
        env["barrier"]()
        ev0.record()
        for _ in range(steps):
            self.step()
            ext_ms.append(self.ctx.last_extract_kernel_ms())   # events already complete: the step ends with the counts on the host
            fld_ms.append(self.ctx.last_field_kernel_ms())
        ev1.record()
        env["barrier"]()
        launches = self.ctx.launch_count()
        clocks = sampler.stop() if sampler is not None else None
        ms = ev0.elapsed_time(ev1) / steps
        self.ctx.enable_kernel_timing(False)
        out = {"ms": ms, "ext_ms": sum(ext_ms) / len(ext_ms), "fld_ms": sum(fld_ms) / len(fld_ms), "launches": launches, "clocks": clocks, "e2e_ms": None}
        if e2e:
            for _ in range(2):
                r2 = self.e2e_step()
            assert r2 == (self.act, self.tot)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            env["barrier"]()
            e0.record()
            for _ in range(steps):
                self.e2e_step()
            e1.record()
            env["barrier"]()
            out["e2e_ms"] = out["e2e_blocking_ms"] = e0.elapsed_time(e1) / steps
            if env["world"] == 1 or self.ctx_on_torch_stream():
                assert all(r == (self.act, self.tot) for r in self.e2e_pipelined(3))
                env["barrier"]()
                e0.record()
                self.e2e_pipelined(steps)
                e1.record()
                env["barrier"]()
                out["e2e_ms"] = e0.elapsed_time(e1) / steps
                out["e2e_pipelined"] = True
        return out

    def reduce(self, r):
        """max over ranks of the times, global counts (all-gather + exclusive scan = global vertex offsets)."""
        torch, env, sharding = self.env["torch"], self.env, self.env["sharding"]
        dist, dev = env["dist"], env["dev"]
        stats = torch.tensor([r["ms"], r["e2e_ms"] or 0.0, r["ext_ms"], r["fld_ms"]], device=dev, dtype=torch.float64)
        per_rank, voff, aoff, (g_act, g_tot) = sharding.gather_counts(dist, self.act, self.tot, device=dev)
        g_launch = r["launches"]
        if env["world"] > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
            lt = torch.tensor([r["launches"]], device=dev, dtype=torch.int64)
            dist.all_reduce(lt)
            g_launch = int(lt[0])
        ms, e2e_ms, ext_k, fld_k = [float(x) for x in stats.cpu()]
        return {"ms": ms, "e2e_ms": e2e_ms if r["e2e_ms"] is not None else None, "e2e_blocking_ms": r.get("e2e_blocking_ms"), "ext_ms": ext_k, "fld_ms": fld_k,
                "launches": g_launch,
                "g_act": g_act, "g_tot": g_tot, "per_rank_verts": [v for (_, v) in per_rank]}

    def summary(self, red, peak):
        """Compact record of a secondary leg (fast mode, ratio 2, config 4)."""
        F, gnz = self.F, self.gnz
        points = F * F * gnz
        alg_ext = 4.0 * F * F * self.nzl + 32.0 * self.tot
        out = {"grid": [F, F, gnz], "control": [self.cxy, self.cxy, self.czg], "ratio": self.R, "field_mode": "fast" if self.fast else "exact",
               "ms_per_step": red["ms"], "value": points / (red["ms"] * 1e-3), "unit": "voxels/s", "triangles": red["g_tot"] // 3,
               "triangles_per_s": (red["g_tot"] / 3) / (red["ms"] * 1e-3), "active_voxels": red["g_act"],
               "field_kernel_ms": red["fld_ms"], "extract_kernel_ms": red["ext_ms"],
               "extraction_hbm_frac": alg_ext / (red["ext_ms"] * 1e-3) / 1e9 / peak}
        if self.fast:
            # the fast field kernel's own roof: one MUFU.COS per (point, harmonic) on the XU pipe, 16 lanes per SM and clock
            ph = float(F) * F * self.nzl * self.NH
            xu_peak = 16.0 * 148 * 1965e6
            out["field_xu_roofline"] = {"bound": "XU (MUFU) pipe", "achieved_cos_per_s": ph / (red["fld_ms"] * 1e-3), "peak_cos_per_s": xu_peak,
                                        "frac": ph / (red["fld_ms"] * 1e-3) / xu_peak, "note": "peak = 16 lanes x 148 SMs x 1965 MHz; ncu: profiles/r02_ncu_svl_field_fast.txt"}
        if red["e2e_ms"] is not None:
            out["e2e"] = {"value": points / (red["e2e_ms"] * 1e-3), "unit": "voxels/s", "ms_per_step": red["e2e_ms"], "blocking_call_ms": red["e2e_blocking_ms"],
                          "h2d_bytes_per_step": int(self.phi.numel() * 4 * self.env["world"]), "d2h_bytes_per_step": int((16 + 8) * self.env["world"])}
        return out

    def free(self):
        torch = self.env["torch"]
        for name in ("phi", "svl", "mesh", "hphi", "phi_scratch", "phi_scratch2", "mm"):
            if hasattr(self, name):
                setattr(self, name, None)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()


def workflow_leg(env, F, R, NH, reps=2):
    """The lattice workflow with the producer ON THE DEVICE (Multitopo::spatial_lattice_run, main.cu:3904-4037): period field -> normalise_three ->
    phase solve of all harmonics on the control grid -> SVL field on the fine grid -> band extraction.  Nothing but the coefficients crosses
    PCIe; this is where the batched phase solve shows in a whole-job number.  Single rank."""
    torch, g, synth, ctx = env["torch"], env["g"], env["synth"], env["ctx"]
    dev = env["dev"]
    c = F // R
    nc, d = c ** 3, (1.0 / R,) * 3
    harm, coef = synth.HARMONICS[:NH], synth.gyroid_coefficients()[:NH]
    per, phi, svl = torch.zeros(nc, device=dev), torch.zeros(NH, nc, device=dev), torch.empty(F ** 3, device=dev)
    lat = g.Gratings(ctx)
    ctx.set_options(0)
    state = {}

    def job(pos, norm, cap):
        lat.period_data(per, c, c, c, 1.0, 1.0, 1.0, c / 2.0, c / 2.0, c / 2.0, "z")                    # main.cu:3927-3931
        lat.GPU_buffer_normalise_three(per, per, nc, float(c // 10), float(c // 4))
        state["fi"], _ = g.svl_phase_solve(ctx, phi, per, harm, (c, c, c), (1.0, 1.0, 1.0), latticetype="r", uniform_type=2, iters=500, end_res=0.01)
        return g.svl_lattice(ctx, svl, phi, coef, (c, c, c), (F, F, F), d, ISO_MASK, BAND_LO, BAND_HI, d, (0.0, 0.0, 0.0), pos, norm, cap)

    probe = g.MeshBuffers(3, device=dev)
    _, tot, _ = job(probe.pos, probe.norm, 3)            # count pass (also the warm-up)
    mesh = g.MeshBuffers(tot + 3, device=dev)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ms, ps = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        e[0].record()
        lat.period_data(per, c, c, c, 1.0, 1.0, 1.0, c / 2.0, c / 2.0, c / 2.0, "z")
        lat.GPU_buffer_normalise_three(per, per, nc, float(c // 10), float(c // 4))
        e[1].record()
        fi, _ = g.svl_phase_solve(ctx, phi, per, harm, (c, c, c), (1.0, 1.0, 1.0), latticetype="r", uniform_type=2, iters=500, end_res=0.01)
        e[2].record()
        act, tot2, _ = g.svl_lattice(ctx, svl, phi, coef, (c, c, c), (F, F, F), d, ISO_MASK, BAND_LO, BAND_HI, d, (0.0, 0.0, 0.0), mesh.pos, mesh.norm, tot + 3)
        e[3].record()
        torch.cuda.synchronize()
        assert tot2 == tot
        ms.append(e[0].elapsed_time(e[3]))
        ps.append(e[1].elapsed_time(e[2]))
    out = {"what": "period field -> normalise_three -> phase solve (%d harmonics, control %d^3, CG <= 500 it, 0.01) -> SVL field + band extraction %d^3; "
                   "producer on the device, only coefficients cross PCIe" % (NH, c, F),
           "ms_per_job": min(ms), "phase_solve_ms": min(ps), "field_and_extraction_ms": min(ms) - min(ps), "value": F ** 3 / (min(ms) * 1e-3), "unit": "voxels/s",
           "cg_iterations_total": int(sum(fi) - len(fi)), "triangles": tot // 3, "active_voxels": act,
           "note": "the reference's own sequence for the same job is timed in the --impl reference line (`workflow`)"}
    del per, phi, svl, mesh
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


# cost model of the slab balancing (ms per point / per vertex at 512-wide rows, from the kernel times of the N = 1 line): the field
# kernel and the stage + classify half of the extraction scale with points, the emission with vertices
COST_PER_POINT = {False: (9.8 + 0.67) / 134.2e6, True: (2.6 + 0.67) / 134.2e6}
COST_PER_VERTEX = 1.57 / 172.6e6


def balanced_leg(env, F, R, NH, gnz, fast):
    """N > 1: count on the even z-cut, all-gather the counts, re-cut the slabs so that the ESTIMATED cost per rank is even (the lattice's
    triangle density varies 1.5x along z), and set the leg up again on the new cuts.  Any partition concatenates to the same global
    mesh; the even cut stays if the re-cut would not change it."""
    sharding = env["sharding"]
    leg = SvlLeg(env, F, R, NH, gnz, fast=fast)
    if env["world"] == 1 or os.environ.get("GCB_BENCH_EVEN_SLABS"):
        return leg, None
    per_rank, _, _, _ = sharding.gather_counts(env["dist"], leg.act, leg.tot, device=env["dev"])
    cuts = sharding.balanced_cuts(gnz, env["world"], leg.cuts, [v for (_, v) in per_rank], COST_PER_POINT[bool(fast)], COST_PER_VERTEX, F * F,
                                  align=R if R % 2 == 0 else 2)   # cuts on control-cell boundaries: the field kernels' tiles then span one cell in z
    verts = [v for (_, v) in per_rank]
    a, b = COST_PER_POINT[bool(fast)], COST_PER_VERTEX

    def worst(c):   # the model's slowest rank on partition c (vertices spread evenly inside each measured slab)
        w = 0.0
        for r in range(env["world"]):
            t = 0.0
            for q in range(env["world"]):
                ov = max(0, min(c[r + 1], leg.cuts[q + 1]) - max(c[r], leg.cuts[q]))
                t += ov * (a * F * F + b * verts[q] / max(leg.cuts[q + 1] - leg.cuts[q], 1))
            w = max(w, t)
        return w
    gain = 1.0 - worst(cuts) / worst(leg.cuts)
    info = {"even_cut_vertices_per_rank": verts, "cell_layer_cuts": cuts, "modelled_gain": round(gain, 4), "applied": bool(cuts != leg.cuts and gain >= 0.05)}
    if not info["applied"]:   # below a modelled 5 % the re-cut measured inside the run-to-run noise (N = 8, exact field: 12.41 ms even vs 12.50 ms re-cut): keep the even cut
        info["cell_layer_cuts"] = leg.cuts
        return leg, info
    leg.free()
    return SvlLeg(env, F, R, NH, gnz, fast=fast, cuts=cuts), info


def h2d_probe(env, nbytes=512 << 20, reps=4):
    """Multi-rank runs: host-to-device bandwidth of one pinned buffer per rank, every rank copying at the same time and rank 0
    alone -- names the limiter of the e2e leg (PCIe links shared behind a switch / host memory) by measurement."""
    torch, dist, dev = env["torch"], env["dist"], env["dev"]
    h = torch.empty(nbytes // 4, dtype=torch.float32, pin_memory=True)
    h.zero_()
    dbuf = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)

    def gbs():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dbuf.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    dbuf.copy_(h)
    env["barrier"]()
    together = gbs()
    env["barrier"]()
    alone = gbs() if env["rank"] == 0 else 0.0
    env["barrier"]()
    t = torch.tensor([together, alone], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(env["world"])]
    dist.all_gather(allt, t)
    rows = [[float(x) for x in a.cpu()] for a in allt]
    return {"bytes": nbytes, "per_rank_gbs_all_ranks_copying": [round(r[0], 2) for r in rows], "rank0_gbs_alone": round(rows[0][1], 2),
            "aggregate_gbs": round(sum(r[0] for r in rows), 1)}


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
    ap.add_argument("--profile", action="store_true", help="timed steps only (no e2e leg, no extra legs, no CPU baseline): for runs under ncu")
    ap.add_argument("--fast-field", action="store_true", help="headline leg in GCB_OPT_FAST_FIELD mode (default: exact, with the fast mode reported beside it)")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the global grid is fine^3 for every N (e.g. --fine 2048 --gpus 8 = BASELINE config 4)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary legs (fast mode, ratio 2, configs 1/2/5, config 4)")
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

    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    env = {"torch": torch, "g": g, "sharding": sharding, "synth": synth, "dist": dist, "dev": dev, "rank": rank, "world": world, "barrier": barrier,
           "ctx": g.Context(local_rank, options=0)}
    ctx = env["ctx"]
    F, R, NH = args.fine, args.ratio, args.harmonics
    gnz = F if args.strong else F * world   # global point layers (weak scaling: one fine^3 block per rank)
    peak, peak_src = peaks()

    # ---- headline leg: BASELINE config 3, exact field (bit-identical to the reference kernels) unless --fast-field
    leg, balance = balanced_leg(env, F, R, NH, gnz, args.fast_field)
    raw = leg.run(args.steps, args.warmup, e2e=not args.profile, sampler=ClockSampler(local_rank) if rank == 0 else None)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": raw["ms"], "extract_kernel_ms": raw["ext_ms"], "field_kernel_ms": raw["fld_ms"],
                              "verts": leg.tot, "launches": raw["launches"], "field_mode": "fast" if args.fast_field else "exact"}))
        return
    red = leg.reduce(raw)
    clocks = raw["clocks"]
    nzl, tot, phi_elems = leg.nzl, leg.tot, leg.phi.numel()
    probe = h2d_probe(env) if world > 1 else None
    leg.free()

    extra = {}
    if not args.no_extra:
        # ---- fast field mode on the same workload (GCB_OPT_FAST_FIELD; tolerance stated in include/gpucad_b200.h)
        if not args.fast_field:
            fl, fbal = balanced_leg(env, F, R, NH, gnz, True)
            extra["fast_field"] = fl.summary(fl.reduce(fl.run(args.steps, args.warmup, e2e=True)), peak)
            if fbal:
                extra["fast_field"]["slab_balance"] = fbal
            extra["fast_field"]["note"] = ("GCB_OPT_FAST_FIELD: |c_h| cos(phi_h + arg c_h) with MUFU.COS + packed-fp32 lerps; field within sum|c_h| (ulp(phi)/2 + 4e-6) "
                                           "of the exact field, extraction bit-exact on that field (tests/test_gpu_parity.py::test_svl_field_fast_mode)")
            fl.free()
        ctx.set_options(0)
        if world == 1 and (F, R) == (512, 4):
            # ---- the reference app's own upsampling ratio 2 (control 256^3, 4.2 GB of control grids)
            r2 = SvlLeg(env, F, 2, NH, gnz)
            extra["ratio2"] = r2.summary(r2.reduce(r2.run(3, 3, e2e=True)), peak)
            r2.free()
            # ---- BASELINE configs 1, 2, 5 at full size through the legacy call sequences (and the fused entry points)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import config_bench
            extra["configs"] = config_bench.run_ours(g, ctx, args.steps, args.warmup)
            extra["configs"]["note"] = "this library only; the reference kernels' times for the same configs are in the --impl reference line"
            # ---- the whole lattice workflow with the producer on the device (phase solve -> field -> mesh)
            try:
                extra["workflow"] = workflow_leg(env, F, R, NH)
            except Exception as e:  # noqa: BLE001
                extra["workflow"] = {"error": str(e)[:200]}
        if not args.strong and F == 512:
            # ---- BASELINE config 4: 2048-wide grid, z-slab sharded.  The 2048^3 mesh alone is ~354 GB (3.7 G triangles x 96 B), so the
            # full grid runs at N >= 4 (<= 106 GB per rank); N = 1 / 2 run the tallest 2048-wide stack of 513 point layers per rank
            g4 = min(2048, 512 * world + 1)
            try:
                c4 = SvlLeg(env, 2048, 4, NH, g4)
                s4 = c4.summary(c4.reduce(c4.run(2, 3, e2e=(world >= 4))), peak)
                s4["complete_2048_cubed"] = bool(g4 == 2048)
                s4["scaling"] = "strong"
                s4["note"] = "BASELINE config 4 (2048^3 SVL lattice, z-slabs)" if g4 == 2048 else \
                    "2048x2048x%d stack (%d point layers per rank): the full 2048^3 mesh (~354 GB) needs the HBM of >= 4 GPUs" % (g4, (g4 - 1) // world + 1)
                extra["config4"] = s4
                c4.free()
            except Exception as e:  # noqa: BLE001  (an allocation failure here must not lose the headline line)
                extra["config4"] = {"error": str(e)[:200]}

    if rank == 0:
        points = F * F * gnz
        ms, e2e_ms, ext_k_ms, fld_k_ms = red["ms"], red["e2e_ms"], red["ext_ms"], red["fld_ms"]
        alg_ext = 4.0 * F * F * nzl + 32.0 * tot          # this rank's launch: 4 B/point read + 32 B/vertex written
        alg_fld = 4.0 * F * F * nzl + 4.0 * phi_elems       # 4 B/point written + the control grids read once
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        sincos = float(F) * F * nzl * NH
        ext_roof = {"bound": "hbm", "kernel": "mc_fused_kernel<M_BAND_RAW, TMA>", "achieved": alg_ext / (ext_k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": alg_ext / (ext_k_ms * 1e-3) / 1e9 / peak,
                    "traffic": NCU_TRAFFIC_BYTES if (F, R, NH, world) == (512, 4, 62, 1) else None,
                    "traffic_source": "profiles/r01_ncu_mc_fused.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch, this workload)",
                    "kernel_ms": ext_k_ms, "share_of_step": ext_k_ms / ms, "algorithmic_bytes": alg_ext, "peak_source": peak_src}
        # the field kernel has (almost) nothing to read: against the HBM roofline it is nowhere by construction; what bounds it is
        # instruction issue / the FP32 pipes, reported as issue slots per (point, harmonic) next to the HBM figure the contract asks for
        fld_name = "svl_field_fast_kernel" if args.fast_field else "svl_field_tile_kernel"
        fld_roof = {"bound": "hbm", "kernel": fld_name, "achieved": alg_fld / (fld_k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": alg_fld / (fld_k_ms * 1e-3) / 1e9 / peak, "traffic": NCU_FIELD_TRAFFIC_BYTES if (F, R, NH, world) == (512, 4, 62, 1) else None,
                    "traffic_source": "profiles/r01_ncu_svl_field.txt", "kernel_ms": fld_k_ms, "share_of_step": fld_k_ms / ms, "algorithmic_bytes": alg_fld,
                    "peak_source": peak_src,
                    "actual_limiter": {"bound": "instruction issue (FP32 FMA / FP64 / XU pipes; nothing to stream)",
                                       "point_harmonics_per_s": sincos / (fld_k_ms * 1e-3),
                                       "issue_slots_per_point_harmonic": 148 * 4 * 32 * sm_mhz * 1e6 * (fld_k_ms * 1e-3) / sincos,
                                       "ncu": "profiles/r02_ncu_svl_field*.txt (issue-slot and pipe utilisation of this kernel)"}}
        dominant, other = (fld_roof, ext_roof) if fld_k_ms >= ext_k_ms else (ext_roof, fld_roof)
        out = {
            "metric": "voxels/s (field + marching cubes, device-timed)", "value": points / (ms * 1e-3), "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(F, R, NH, gnz, world, args.strong),
            "field_mode": "fast (GCB_OPT_FAST_FIELD)" if args.fast_field else "exact (field bit-identical to the reference kernels)",
            "triangles_per_s": (red["g_tot"] / 3) / (ms * 1e-3), "triangles": red["g_tot"] // 3, "active_voxels": red["g_act"],
            "per_rank_vertices": red["per_rank_verts"],
            "slab_balance": balance,
            "e2e": {"value": points / (e2e_ms * 1e-3), "unit": "voxels/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(phi_elems * 4 * world), "d2h_bytes_per_step": int((16 + 8) * world),
                    "mode": ("two-deep job pipeline (gcb_svl_lattice_host_submit / _wait): every step copies its control grids from pinned host memory and "
                             "reads its counts back; step i+1's copies overlap step i's kernels") if world == 1 else
                            ("two-deep job pipeline per rank (gcb_svl_slab_host_submit_field -> NCCL all-reduce of the range, stream-ordered -> "
                             "gcb_svl_slab_host_submit_extract; _wait one job later): step i+1's copies overlap step i's kernels") if raw.get("e2e_pipelined") else
                            "one blocking sequence per step and rank (host control grids -> field -> NCCL min/max on the device -> extraction -> counts)",
                    "blocking_call": {"ms_per_step": red["e2e_blocking_ms"], "value": points / (red["e2e_blocking_ms"] * 1e-3),
                                      "note": "one blocking call sequence per step, nothing overlapped across steps"},
                    "note": "control grids copied from pinned host memory each step; counts and min/max read back; the mesh stays in device memory "
                            "as in the reference (Vulkan-exported vertex buffers)",
                    "host_cpus_bound_rank0": affinity},
            "gpu_launches": red["launches"],
            "roofline": dominant, "roofline_other_kernel": other,
            "clocks": clocks,
        }
        if probe:
            out["e2e"]["h2d_probe"] = probe
        out.update(extra)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(NH)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# measured once with `ncu --set full` on the default workload (profiles/r01_ncu_mc_fused.txt): 0.540 GB read + 5.468 GB written
NCU_TRAFFIC_BYTES = 539927808 + 5467993000
# profiles/r01_ncu_svl_field.txt: 0.52 GB read (control grids once) + 0.50 GB written (the field)
NCU_FIELD_TRAFFIC_BYTES = 521009664 + 500735744


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
    """The reference's own implementation of the path: its unmodified CUDA kernels (oracle/_ref) on one GPU.  Rank 0 only; at
    N > 1 it times ONE fine^3 block -- one rank's share of the weak-scaling workload -- and reports that block's voxels/s."""
    if rank != 0:
        return
    import ref_py as ref
    F, R, NH = args.fine, args.ratio, args.harmonics
    gnz = F if args.strong else F * world
    base = {"impl": "reference", "metric": "voxels/s (field + marching cubes, device-timed)", "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(F, R, NH, gnz, world, args.strong)}
    if not (ref.available() and torch.cuda.is_available()):
        cb = cpu_baseline(NH)
        cb["sample"] = "bounded CPU sample of the workload (no GPU / reference library here): " + cb["sample"]
        base.update(value=cb["value"], ms_per_step=None, cpu_baseline=cb,
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
                reference_note="reference CUDA kernels (unmodified sources, sm_100a) via oracle/_ref, device-resident inputs; classify launched with a "
                               "corrected 2-D grid: %s" % bool(fix),
                cpu_baseline={"value": val, "unit": "voxels/s", "cores": 0, "kind": "reference",
                              "sample": "one %d^3 block (%s) on 1 B200 with the reference's own CUDA kernels -- the reference has no CPU implementation"
                                        % (F, "the whole workload" if gnz == F else "1/%d of the workload: one rank's share" % world)},
                e2e={"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=launches_per_step * args.steps)
    if not args.no_extra and world == 1 and (F, R) == (512, 4):
        # the whole lattice workflow as the reference runs it (main.cu:3904-4037): period_data, GPU_buffer_normalise_three, per harmonic
        # finding_phi + host-driven GPUCG_lattice, then the field loop and the extraction timed above.  One job (it takes seconds).
        try:
            nc = c ** 3
            harm = synth.HARMONICS[:NH]
            per = torch.zeros(nc, device=dev)
            phi_w = torch.zeros(NH, nc, device=dev)
            wcap = [0]

            def wjob(count_only):
                ref.period_data(per, (c, c, c), (1.0, 1.0, 1.0), (c / 2.0,) * 3, "z")
                ref.normalise_three(per, per, nc, float(c // 10), float(c // 4))
                its = 0
                for h in range(NH):
                    ref.finding_phi(phi_w[h], per, (c, c, c), harm[h], (1.0, 1.0, 1.0), latticetype="r", uniform_type=2)
                    its += ref.cg(phi_w[h], (c, c, c), 500, 0.01)[0]
                svl.zero_()
                ref.svl_field(svl, ga, phi_w, NH, dcoef, (c, c, c), (F, F, F), d)
                ref.normalise_four(svl, mask, k, (F, F, F), BAND_LO, BAND_HI)
                if count_only:
                    return ref.isosurface_lattice(False, fix, mask, tmp.pos, tmp.norm, ISO_MASK, (F, F, F), d, (0, 0, 0), scr, 3, k, zeros, BAND_LO, BAND_HI, 0.0, 0.0), its
                return ref.isosurface_lattice(False, fix, mask, wmesh.pos, wmesh.norm, ISO_MASK, (F, F, F), d, (0, 0, 0), scr, wcap[0], k, zeros, BAND_LO, BAND_HI,
                                              0.0, 0.0), its
            (_, wtot), _ = wjob(True)
            mesh = None
            torch.cuda.empty_cache()
            wcap[0] = wtot + 3
            wmesh = g.MeshBuffers(wcap[0], device=dev)
            torch.cuda.synchronize()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            (wact, wtot2), wits = wjob(False)
            w1.record()
            torch.cuda.synchronize()
            wms = w0.elapsed_time(w1)
            base["workflow"] = {"what": "period_data -> GPU_buffer_normalise_three -> %d x (finding_phi + GPUCG_lattice) on control %d^3 -> %d x (texture upload, "
                                        "grating, svl) -> GPU_buffer_normalise_four -> computeIsosurface_lattice %d^3: the reference's kernels and host loops" % (NH, c, NH, F),
                                "ms_per_job": wms, "value": F ** 3 / (wms * 1e-3), "unit": "voxels/s", "cg_iterations_total": int(wits - NH), "triangles": wtot2 // 3,
                                "active_voxels": wact}
            del per, phi_w, wmesh
        except Exception as e:  # noqa: BLE001
            base["workflow"] = {"error": str(e)[:200]}
        # the reference kernels on BASELINE configs 1, 2, 5 (the product arm reports its own times for the same configs)
        del svl, ga, mask, k, zeros, scr, mesh, phi
        ref.delete_texture()
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config_bench
        base["configs"] = config_bench.run_reference(g, ref, args.steps, max(args.warmup, 3))
    print(json.dumps(base))


if __name__ == "__main__":
    main()
