#!/usr/bin/env python3
"""BASELINE config 5 on N GPUs: the topology-optimised density is a STORED field, so every rank holds the point layers of its
z-slab (density, grid_points, d_result) and receives the one +z halo layer from the rank above over NCCL point-to-point
(NVLink P2P on the NVSwitch box) -- SURVEY.md 8e.  Per step and rank: refine (2x upsample of the rank's coarse planes) ->
halo exchange (24 B per point of one layer) -> computeIsosurface_2 through the legacy entry point with the slab's gridcenter
-> all-gather of {active, verts}.  Timed on the device, max over ranks.  With --check rank 0 also runs the whole grid alone
and every rank compares its mesh with its span of that single-GPU mesh, byte for byte.

    python tools/config5_multi.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/config5_multi.py --check > profiles/rNN_config5_n2.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gpucadforam_b200 as g  # noqa: E402
from gpucadforam_b200 import sharding, synth  # noqa: E402

ISO = 0.4  # VolumeFraction (ImguiApp.cpp:273 neighbourhood)


def bytes_equal(a, b):
    return bool(torch.equal(a.contiguous().view(torch.uint8).reshape(-1), b.contiguous().view(torch.uint8).reshape(-1)))


class SlabPipeline:
    """refine + halo + extraction for point layers z0..z1 of the global fine grid (owned: z0..z1-1, the top rank also z1)."""

    def __init__(self, ctx, coarse, cdims, fdims, d, z0, z1, top, max_verts=None):
        cx, cy, cz = cdims
        fx, fy, fz = fdims
        self.ctx, self.fdims, self.d, self.z0 = ctx, fdims, d, z0
        self.nzl = z1 - z0 + 1
        self.owned = self.nzl if top else self.nzl - 1
        self.plane = fx * fy
        c0, c1 = sharding.control_slab(z0, z0 + self.owned - 1, 2, cz)
        self.cdl = (cx, cy, c1 - c0 + 1)
        self.coarse = coarse.view(cz, cy, cx)[c0:c1 + 1].contiguous().view(-1)
        npl = self.plane * self.nzl
        self.dens = torch.zeros(npl, device="cuda")
        self.vol_topo = torch.zeros((npl, 4), dtype=torch.int32, device="cuda")  # grid_points: val = 0, t = 0
        self.result = torch.zeros(npl, device="cuda")
        self.ldims = (fx, fy, self.nzl)
        self.center = sharding.slab_gridcenter((0.0, 0.0, 0.0), z0)
        self.lat, self.iso = g.Gratings(ctx), g.Isosurface(ctx)
        self.lat.setupTexture(*self.cdl)
        self.pbuf = torch.zeros(self.coarse.numel(), device="cuda")
        self.pp = self.lat.pitched(self.pbuf, cx, cy)
        self.scr = g.Scratch((fx - 1) * (fy - 1) * (self.nzl - 1))
        self.mesh = None
        self.max_verts = max_verts
        self.halo_bytes = 0

    def step(self, use_dist):
        cx, cy, czl = self.cdl
        fx, fy, _ = self.fdims
        self.lat.copytotexture(self.coarse, self.pp, cx, cy, czl)
        self.lat.updateTexture(self.pp)
        self.lat.refine(self.dens, fx, fy, self.owned, *self.d)
        self.halo_bytes = sharding.exchange_halo_planes(dist if use_dist else None, [(self.dens, self.plane), (self.vol_topo.view(-1), self.plane * 4),
                                                                                     (self.result, self.plane)], self.nzl)
        if self.mesh is None:  # count first, then allocate the mesh exactly (the reference preallocates 4 vertices per point)
            probe = g.MeshBuffers(3)
            _, tot = self.iso.computeIsosurface_2(probe.pos, probe.norm, ISO, self.scr, self.ldims, self.d, self.center, 3, self.vol_topo, self.dens, 0.0,
                                                  self.result)
            self.max_verts = self.max_verts or tot + 3
            self.mesh = g.MeshBuffers(self.max_verts)
        return self.iso.computeIsosurface_2(self.mesh.pos, self.mesh.norm, ISO, self.scr, self.ldims, self.d, self.center, self.max_verts, self.vol_topo,
                                            self.dens, 0.0, self.result)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--coarse", default="384,192,192")
    ap.add_argument("--blocking", action="store_true", help="every legacy call ends in a synchronise, as the reference wrappers do")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    use_dist = world > 1
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cdims = tuple(int(v) for v in args.coarse.split(","))
    cx, cy, cz = cdims
    fdims, d = (2 * cx, 2 * cy, 2 * cz), (0.5, 0.5, 0.5)
    fx, fy, fz = fdims
    # the stored input of the path; made once on rank 0 and broadcast so that every rank slices the same array (set-up, not timed)
    coarse = synth.cantilever_density(cx, cy, cz, struts=40, sigma=1.5, device="cuda").contiguous().view(-1) if rank == 0 else torch.zeros(cx * cy * cz,
                                                                                                                                           device="cuda")
    if use_dist:
        dist.broadcast(coarse, 0)
    ctx = g.Context(local, options=0 if args.blocking else g._capi.GCB_OPT_ASYNC_FIELDS)
    z0, z1 = sharding.slab_bounds(fz, world, rank)
    pipe = SlabPipeline(ctx, coarse, cdims, fdims, d, z0, z1, top=(rank == world - 1))

    def step():
        act, tot = pipe.step(use_dist)
        return sharding.gather_counts(dist if use_dist else None, act, tot, device="cuda")

    for _ in range(max(args.warmup, 1)):
        out = step()
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
    if use_dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    per_rank, voff, aoff, (ta, tv) = out
    line = {"config": 5, "workload": "cantilever density %dx%dx%d (coarse %dx%dx%d, 40 struts, sigma 1.5): refine + +z halo layer + computeIsosurface_2, iso 0.4"
            % (fx, fy, fz, cx, cy, cz), "n_gpus": world, "scaling": "strong", "legacy_calls": "blocking" if args.blocking else "enqueue only (GCB_OPT_ASYNC_FIELDS)", "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms[0]),
            "voxels_per_s": fx * fy * fz / (float(ms[0]) * 1e-3), "triangles_per_s": tv / 3 / (float(ms[0]) * 1e-3), "active_voxels": ta,
            "triangles": tv // 3, "per_rank_active_verts": per_rank, "vertex_offsets": voff,
            "halo": {"bytes_per_rank_per_step": fx * fy * 24 if use_dist else 0, "transport": "NCCL send/recv (P2P over NVLink)" if use_dist else "none",
                     "layers": "density 4 B + grid_points 16 B + d_result 4 B per point of one layer"}}
    if args.check:
        ok = torch.ones(1, device="cuda")
        n_single = torch.zeros(2, dtype=torch.int64, device="cuda")
        single = None
        if rank == 0:  # pipe's texture is replaced here; pipe is not stepped again, only its mesh is read
            whole = SlabPipeline(ctx, coarse, cdims, fdims, d, 0, fz - 1, top=True)
            a1, t1 = whole.step(False)
            n_single[0], n_single[1] = a1, t1
            single = whole.mesh
        if use_dist:
            dist.broadcast(n_single, 0)
        a1, t1 = int(n_single[0]), int(n_single[1])
        counts_ok = (a1, t1) == (ta, tv)
        mesh_ok = False
        if counts_ok:
            pos1 = single.pos[:t1].contiguous() if rank == 0 else torch.zeros((t1, 4), device="cuda")
            norm1 = single.norm[:t1].contiguous() if rank == 0 else torch.zeros((t1, 4), device="cuda")
            if use_dist:
                dist.broadcast(pos1, 0)
                dist.broadcast(norm1, 0)
            mine_t = per_rank[rank][1]
            same = bytes_equal(pipe.mesh.pos[:mine_t], pos1[voff[rank]:voff[rank] + mine_t]) and bytes_equal(pipe.mesh.norm[:mine_t],
                                                                                                              norm1[voff[rank]:voff[rank] + mine_t])
            ok[0] = 1.0 if same else 0.0
            if use_dist:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            mesh_ok = bool(ok[0] > 0)
        line["parity_vs_single_gpu"] = {"counts": counts_ok, "mesh_bytes_every_rank": mesh_ok, "single_gpu_active_verts": [a1, t1]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
