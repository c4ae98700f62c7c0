"""Host-side logic of z-slab sharding.  Pure Python / torch.distributed, no CUDA calls, so the
N > 1 path is covered by world_size-2 gloo tests on the CPU (tests/test_sharding_gloo.py) and shared by bench.py.

The path shards into independent units plus ONE tiny exchange step:
  * rank r owns cell layers [z0, z1) of the global grid and holds point layers z0..z1 (its +z halo plane);
  * analytic / control-grid fields are evaluated locally for exactly those planes -- no halo traffic;
  * global min/max of the field (normalisation) : one all-reduce of 2 floats between field and extraction;
  * global vertex offsets                       : one all-gather of {active, verts} per rank, exclusive scan.
Concatenating the rank meshes in rank order reproduces the single-GPU buffers byte for byte, because the
reference orders vertices by ascending linear cell id with z slowest (MarchingCubes_kernel.cu:120-136, :2160).
"""
import torch


def slab_bounds(gnz, world, rank, align=2):
    """Cell layers [z0, z1) owned by `rank` of a grid with gnz point layers (gnz-1 cell layers).  Interior boundaries are
    multiples of `align` (2: the fused SVL kernel evaluates 2x2x2 point blocks and wants slabs to start on an even layer)."""
    cells = gnz - 1

    def cut(r):
        if r <= 0:
            return 0
        if r >= world:
            return cells
        return min(cells, max(0, int(round(r * cells / world / align)) * align))
    return cut(rank), cut(rank + 1)


def control_slab(z0, z1, ratio, czg):
    """Control-grid planes [c0, c1] a fine slab holding point layers z0..z1 samples (trilinear: floor(z/ratio) and +1)."""
    c0 = z0 // ratio
    c1 = min(z1 // ratio + 1, czg - 1)
    return c0, c1


def allreduce_minmax(dist, mm):
    """Global {min, max} from per-rank {min, max} (2-element tensor on the backend's device) with ONE collective:
    max-reduce of {-min, max}.  Returns python floats."""
    t = torch.stack([-mm[0], mm[1]])
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return -float(t[0]), float(t[1])


def gather_counts(dist, active, verts, device="cpu"):
    """All-gather of per-rank {active, verts}; returns (per-rank list, exclusive vertex offsets, exclusive active offsets, totals)."""
    mine = torch.tensor([int(active), int(verts)], dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        allc = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(allc, mine)
    else:
        allc = [mine]
    per_rank = [(int(c[0]), int(c[1])) for c in allc]
    voff, aoff, v, a = [], [], 0, 0
    for (ac, vc) in per_rank:
        aoff.append(a)
        voff.append(v)
        a += ac
        v += vc
    return per_rank, voff, aoff, (a, v)
