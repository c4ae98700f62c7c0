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


def validate_slabs(gnz, world, align=2):
    """Raises (identically on every rank: pure function of its arguments) if the cut leaves some rank without a cell layer."""
    empty = [r for r in range(world) if slab_bounds(gnz, world, r, align)[1] <= slab_bounds(gnz, world, r, align)[0]]
    if empty:
        raise ValueError("z-slab sharding: %d point layers over %d ranks (alignment %d) leaves ranks %s without a cell layer" % (gnz, world, align, empty))


def balanced_cuts(gnz, world, cuts, verts_per_rank, cost_per_point, cost_per_vertex, points_per_layer, align=2):
    """Slab cuts that equalise the estimated cost per rank instead of the number of layers.  `cuts` (world + 1 cell-layer boundaries)
    is the partition the counts were taken on and `verts_per_rank` its vertex counts (one count-only pass + the all-gather of the
    counts); inside a measured slab the vertices are taken as spread evenly over its layers.  cost(layer) = cost_per_point *
    points_per_layer + cost_per_vertex * vertices(layer).  Pure function of its arguments: every rank computes the same cuts.
    Interior cuts are multiples of `align`, every slab keeps at least `align` cell layers."""
    cells = gnz - 1
    if world == 1:
        return [0, cells]
    layer_cost = []
    for r in range(world):
        n = cuts[r + 1] - cuts[r]
        per_layer = cost_per_point * points_per_layer + cost_per_vertex * (verts_per_rank[r] / max(n, 1))
        layer_cost += [per_layer] * n
    total = sum(layer_cost)
    out, acc, z = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while z < cells and acc + layer_cost[z] <= target:
            acc += layer_cost[z]
            z += 1
        # nearest multiple of align to the (fractional) crossing point, kept monotone with room for the remaining ranks
        frac = (target - acc) / layer_cost[z] if z < cells else 0.0
        c = int(round((z + frac) / align)) * align
        c = max(c, out[-1] + align)
        c = min(c, cells - align * (world - r))
        out.append(c)
    out.append(cells)
    if any(b - a < align for a, b in zip(out[:-1], out[1:])):
        return list(cuts)   # degenerate (fewer layers than ranks * align): keep the partition that was given
    return out


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


def allreduce_minmax_device(dist, mm):
    """Same reduction, result left on the device: a 2-element tensor {min, max} the extraction kernel reads in place
    (gcb_extract_band_raw_dev), so that no host synchronisation sits between the field and the extraction."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return mm
    t = torch.stack([-mm[0], mm[1]])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return torch.stack([-t[0], t[1]])


def gather_counts(dist, active, verts, device="cpu"):
    """All-gather of per-rank {active, verts}; returns (per-rank list, exclusive vertex offsets, exclusive active offsets, totals)."""
    mine = torch.tensor([int(active), int(verts)], dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        allc = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(allc, mine)
    else:
        allc = [mine]
    per_rank = [(int(a), int(v)) for (a, v) in torch.stack(allc).tolist()]  # one device-to-host read for all ranks
    voff, aoff, v, a = [], [], 0, 0
    for (ac, vc) in per_rank:
        aoff.append(a)
        voff.append(v)
        a += ac
        v += vc
    return per_rank, voff, aoff, (a, v)


def slab_gridcenter(gridcenter, z0):
    """gridcenter a rank passes to the legacy extraction entry points for a slab whose first point layer is global layer z0:
    the reference computes positions as (gridPos - gridcenter) * voxelSize (MarchingCubes_kernel.cu:1888-1890), and
    (z_local - (gc_z - z0)) equals (z_global - gc_z) exactly in fp32 (small integers / half-integers)."""
    return (float(gridcenter[0]), float(gridcenter[1]), float(gridcenter[2]) - float(z0))


def exchange_halo_planes(dist, fields, nzl):
    """+z halo of STORED fields (density, grid_points, d_result ...): every local buffer holds `nzl` point layers,
    layers 0..nzl-2 owned, layer nzl-1 = the first owned layer of the rank above (the top rank owns its last layer too).
    `fields` is a list of (flat tensor, elements per point layer).  One batched send/recv per neighbour pair: NCCL point-to-point
    (NVLink P2P on an NVSwitch box) for CUDA tensors, gloo for the CPU tests.  MC cells only look at +1
    (MarchingCubes_kernel.cu:889-896), so nothing travels upwards."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    rank, world = dist.get_rank(), dist.get_world_size()
    # An empty slab (more ranks than aligned cell layers) is a configuration error.  It is detected COLLECTIVELY -- one MIN
    # all-reduce of the local layer count -- so that every rank raises before any point-to-point operation is posted; a rank that
    # raised on its own would leave its neighbours waiting in batch_isend_irecv forever.
    least = torch.tensor([int(nzl)], dtype=torch.int64, device=fields[0][0].device if fields else "cpu")
    dist.all_reduce(least, op=dist.ReduceOp.MIN)
    if int(least[0]) < 2:
        raise ValueError("exchange_halo_planes: a rank owns no point layer (smallest slab holds %d layers, rank %d holds %d): "
                         "more ranks than aligned cell layers" % (int(least[0]), rank, nzl))
    ops, nbytes = [], 0
    for (buf, plane) in fields:
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, buf[:plane], rank - 1))
        if rank < world - 1:
            halo = buf[(nzl - 1) * plane: nzl * plane]
            ops.append(dist.P2POp(dist.irecv, halo, rank + 1))
            nbytes += halo.numel() * halo.element_size()
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return nbytes
