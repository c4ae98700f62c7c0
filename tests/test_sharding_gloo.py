"""N > 1 host logic on the CPU: world_size-2 gloo processes, the CPU oracle standing in for the per-rank GPU work.
Checks that z-slab sharding + one min/max all-reduce + one count all-gather reproduce the single-rank mesh byte for byte."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import oracle_py as orc  # noqa: E402
from gpucadforam_b200 import sharding  # noqa: E402

N = 40
VOX = (0.5, 0.5, 0.5)


def _field():
    return orc.create_lattice(N, N, N, 0)


def _band_raw_slab(slab, z0, gnz, lo, hi):
    """oracle version of gcb_extract_band_raw on a slab: normalise with the GLOBAL min/max, global boundary faces, global z."""
    nzl = slab.shape[0]
    k = ((slab - np.float32(lo)) / (np.float32(hi) - np.float32(lo))).astype(np.float32)
    mask = ((k >= np.float32(cases.BAND_LO)) & (k <= np.float32(cases.BAND_HI))).astype(np.float32)
    gz = np.arange(z0, z0 + nzl)
    face = np.zeros(slab.shape, bool)
    face[:, 0, :] = face[:, -1, :] = face[:, :, 0] = face[:, :, -1] = True
    face[gz == 0] = True
    face[gz == gnz - 1] = True
    k[face] = 0.0
    mask[face] = 0.0
    r = orc.extract(orc.MODE_LATTICE_ONE, (N, N, nzl), VOX, (0, 0, -float(z0)), cases.ISO_MASK, f0=mask, f1=k, iso1=cases.BAND_LO, iso2=cases.BAND_HI)
    return r


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = _field()
    z0, z1 = sharding.slab_bounds(N, world, rank)
    slab = np.ascontiguousarray(f[z0:z1 + 1])
    lo_l, hi_l = orc.minmax(slab)
    lo, hi = sharding.allreduce_minmax(dist, torch.tensor([lo_l, hi_l], dtype=torch.float32))
    dev_pair = sharding.allreduce_minmax_device(dist, torch.tensor([lo_l, hi_l], dtype=torch.float32))   # the sync-free variant bench.py uses
    assert (float(dev_pair[0]), float(dev_pair[1])) == (lo, hi)
    r = _band_raw_slab(slab, z0, N, lo, hi)
    per_rank, voff, aoff, totals = sharding.gather_counts(dist, r["active"], r["total"])
    np.savez(os.path.join(out, "rank%d.npz" % rank), pos=r["pos"][:r["total"]], norm=r["norm"][:r["total"]],
             comp=r["compVoxelArray"].astype(np.int64) + z0 * (N - 1) * (N - 1), voff=voff[rank], aoff=aoff[rank], totals=np.array(totals), lohi=np.array([lo, hi]))
    dist.destroy_process_group()


def test_two_rank_slabs_reproduce_single_rank(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    f = _field()
    lo, hi = orc.minmax(f)
    single = _band_raw_slab(f, 0, N, lo, hi)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert tuple(parts[0]["lohi"]) == (np.float32(lo), np.float32(hi)) or np.allclose(parts[0]["lohi"], [lo, hi], rtol=0, atol=0)
    assert tuple(parts[0]["totals"]) == (single["active"], single["total"])
    assert int(parts[0]["voff"]) == 0 and int(parts[1]["voff"]) == len(parts[0]["pos"])
    pos = np.concatenate([p["pos"] for p in parts])
    norm = np.concatenate([p["norm"] for p in parts])
    comp = np.concatenate([p["comp"] for p in parts])
    t = single["total"]
    assert np.array_equal(pos.view(np.uint32), single["pos"][:t].view(np.uint32))
    assert np.array_equal(norm.view(np.uint32), single["norm"][:t].view(np.uint32))
    assert np.array_equal(comp, single["compVoxelArray"].astype(np.int64))


def test_slab_bounds_cover_every_cell_layer_once():
    for gnz in (2, 17, 512, 2048):
        for world in (1, 2, 3, 4, 8):
            b = [sharding.slab_bounds(gnz, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == gnz - 1
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_control_slab_covers_sampled_planes():
    for ratio in (2, 4, 8):
        czg = 64
        gnz = czg * ratio
        for world in (2, 4, 8):
            for r in range(world):
                z0, z1 = sharding.slab_bounds(gnz, world, r)
                c0, c1 = sharding.control_slab(z0, z1, ratio, czg)
                for z in (z0, z1):
                    i = z // ratio
                    assert c0 <= i <= c1 and min(i + 1, czg - 1) <= c1


# ---------------------------------------------------------------- stored fields: one-plane +z halo exchange (config 5 sharded)
def _topo_inputs():
    """config 5 in miniature with non-trivial stored state: density (refined), grid_points with crossing parameters, d_result."""
    T = cases.TOPO
    fx, fy, fz = T["fdims"]
    npts = fx * fy * fz
    dens = orc.refine(cases.topo_coarse(T), T["fdims"], T["d"]).reshape(-1)
    rng = np.random.RandomState(5)
    gp = np.zeros(npts, orc.GP_DTYPE)
    gp["t_x"] = np.where(rng.rand(npts) < 0.2, rng.rand(npts), 0).astype(np.float32)
    gp["t_z"] = np.where(rng.rand(npts) < 0.2, rng.rand(npts), 0).astype(np.float32)
    gp["val"][:150] = -1
    result = rng.rand(npts).astype(np.float32)
    return T, dens, gp, result


def _topo_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, dens, gp, result = _topo_inputs()
    fx, fy, fz = T["fdims"]
    plane = fx * fy
    z0, z1 = sharding.slab_bounds(fz, world, rank)
    nzl = z1 - z0 + 1
    owned = nzl if rank == world - 1 else nzl - 1            # the top rank owns its last point layer too

    def local(a, width, np_dtype, poison):
        """local buffer of nzl point layers: owned layers filled, the halo layer poisoned until the exchange fills it"""
        flat = np.ascontiguousarray(a[z0 * plane:(z0 + owned) * plane]).view(np_dtype).reshape(-1)
        buf = torch.from_numpy(np.full(nzl * plane * width, poison, np_dtype))
        buf[:flat.size] = torch.from_numpy(flat)
        return buf
    l_dens = local(dens, 1, np.float32, np.nan)
    l_res = local(result, 1, np.float32, np.nan)
    l_gp = local(gp, 4, np.int32, 0x7fc00000)
    nbytes = sharding.exchange_halo_planes(dist, [(l_dens, plane), (l_gp, plane * 4), (l_res, plane)], nzl)
    assert nbytes == (0 if rank == world - 1 else plane * 24)
    gc = sharding.slab_gridcenter((0.0, 0.0, 0.0), z0)
    r = orc.extract(orc.MODE_TOPO, (fx, fy, nzl), T["d"], gc, T["iso"], f0=l_dens.numpy(), f1=l_res.numpy(), gp=l_gp.numpy().view(orc.GP_DTYPE).reshape(-1),
                    iso1=0.0)
    per_rank, voff, aoff, totals = sharding.gather_counts(dist, r["active"], r["total"])
    np.savez(os.path.join(out, "topo%d.npz" % rank), pos=r["pos"][:r["total"]], norm=r["norm"][:r["total"]],
             comp=r["compVoxelArray"].astype(np.int64) + z0 * (fx - 1) * (fy - 1), voff=voff[rank], totals=np.array(totals))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stored_field_halo_exchange_reproduces_single_rank(tmp_path, world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_topo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    T, dens, gp, result = _topo_inputs()
    single = orc.extract(orc.MODE_TOPO, T["fdims"], T["d"], (0, 0, 0), T["iso"], f0=dens, f1=result, gp=gp, iso1=0.0)
    parts = [np.load(os.path.join(str(tmp_path), "topo%d.npz" % r)) for r in range(world)]
    t = single["total"]
    assert t > 0 and all(len(p["pos"]) > 0 for p in parts)
    assert tuple(parts[0]["totals"]) == (single["active"], t)
    assert [int(p["voff"]) for p in parts] == list(np.cumsum([0] + [len(p["pos"]) for p in parts[:-1]]))
    assert np.array_equal(np.concatenate([p["pos"] for p in parts]).view(np.uint32), single["pos"][:t].view(np.uint32))
    assert np.array_equal(np.concatenate([p["norm"] for p in parts]).view(np.uint32), single["norm"][:t].view(np.uint32))
    assert np.array_equal(np.concatenate([p["comp"] for p in parts]), single["compVoxelArray"].astype(np.int64))


def test_slab_gridcenter_shift_is_exact_in_fp32():
    """(z_local - (gc_z - z0)) == (z_global - gc_z) bit for bit for the centres the reference uses: 0 (initMC_two, main.cu:2125-2139)
    and half-integer grid centres, for every layer of a 2048-high grid and every even slab start."""
    z = np.arange(0, 2048, dtype=np.float32)
    for gc in (0.0, 1023.5, 383.5, -12.25, 191.0):
        for z0 in (0, 2, 256, 1024, 1790, 2046):
            shifted = np.float32(sharding.slab_gridcenter((0.0, 0.0, gc), z0)[2])
            zl = z[z0:] - np.float32(z0)
            assert np.array_equal((zl - shifted).view(np.uint32), (z[z0:] - np.float32(gc)).view(np.uint32))


def test_halo_exchange_single_rank_is_a_no_op():
    buf = torch.arange(12, dtype=torch.float32)
    assert sharding.exchange_halo_planes(None, [(buf, 4)], 3) == 0
    assert torch.equal(buf, torch.arange(12, dtype=torch.float32))


def _empty_slab_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gnz = 5
    z0, z1 = sharding.slab_bounds(gnz, world, rank)
    nzl = (z1 - z0 + 1) if z1 > z0 else 1
    buf = torch.zeros(max(nzl, 1) * 4)
    try:
        sharding.exchange_halo_planes(dist, [(buf, 4)], nzl)
        verdict = "no error"
    except ValueError as e:
        verdict = "ValueError" if "owns no point layer" in str(e) else "other: %s" % e
    open(os.path.join(out, "empty%d.txt" % rank), "w").write(verdict)
    dist.destroy_process_group()


def test_empty_slab_fails_on_every_rank_instead_of_hanging(tmp_path):
    """5 point layers over 4 ranks leaves ranks without a cell layer: every rank must raise BEFORE any send/recv is posted
    (a rank that raised alone would leave its neighbours blocked in batch_isend_irecv)."""
    world = 4
    assert sharding.slab_bounds(5, world, 0) == (0, 0)
    with pytest.raises(ValueError):
        sharding.validate_slabs(5, world)
    sharding.validate_slabs(40, 3)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_empty_slab_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert [open(os.path.join(str(tmp_path), "empty%d.txt" % r)).read() for r in range(world)] == ["ValueError"] * world


def test_cpp_slab_bounds_equal_the_python_ones():
    """gcb_slab_bounds / gcb_control_slab (csrc/multi.cu, the C++ multi-GPU host) cut exactly like sharding.slab_bounds / control_slab
    (ties of the even split included: Python's round() and nearbyint() both round half to even)."""
    import ctypes as C
    from gpucadforam_b200 import _capi
    lib = _capi.load()
    for gnz in (5, 17, 40, 41, 129, 385, 513, 2048, 2049):
        for world in (1, 2, 3, 4, 5, 7, 8, 16):
            for align in (1, 2, 4):
                for rank in range(world):
                    z0, z1 = C.c_uint(0), C.c_uint(0)
                    assert lib.gcb_slab_bounds(gnz, world, rank, align, C.byref(z0), C.byref(z1)) == 0
                    assert (z0.value, z1.value) == sharding.slab_bounds(gnz, world, rank, align), (gnz, world, rank, align)
    for (z0, z1, ratio, czg) in ((0, 64, 4, 128), (64, 127, 4, 32), (10, 33, 2, 17), (0, 511, 4, 128)):
        c0, c1 = C.c_int(0), C.c_int(0)
        assert lib.gcb_control_slab(z0, z1, ratio, czg, C.byref(c0), C.byref(c1)) == 0
        assert (c0.value, c1.value) == sharding.control_slab(z0, z1, ratio, czg)


def test_balanced_cuts_equalise_cost_and_stay_valid():
    """sharding.balanced_cuts: with vertex counts that vary 1.5x between the outer and inner slabs (the bench lattice at N = 8), the
    re-cut partition's estimated cost per rank is within a few percent of the mean, cuts are aligned, monotone and cover the grid."""
    gnz, world = 4096, 8
    eq = [sharding.slab_bounds(gnz, world, r)[0] for r in range(world)] + [gnz - 1]
    verts = [266272662, 222891732, 184530948, 173026026, 172751532, 184811874, 222420294, 265232904]
    ppl = 512 * 512
    for (a, b) in ((7.8e-8, 0.9e-8), (2.4e-8, 0.9e-8), (0.0, 1.0)):
        cuts = sharding.balanced_cuts(gnz, world, eq, verts, a, b, ppl)
        assert cuts[0] == 0 and cuts[-1] == gnz - 1 and len(cuts) == world + 1
        assert all(c % 2 == 0 for c in cuts[1:-1]) and all(y - x >= 2 for x, y in zip(cuts[:-1], cuts[1:]))

        def cost(lo, hi):   # the model the function itself uses: vertices spread evenly inside each measured slab
            t = 0.0
            for r in range(world):
                ov = max(0, min(hi, eq[r + 1]) - max(lo, eq[r]))
                t += ov * (a * ppl + b * verts[r] / (eq[r + 1] - eq[r]))
            return t
        costs = [cost(cuts[r], cuts[r + 1]) for r in range(world)]
        before = [cost(eq[r], eq[r + 1]) for r in range(world)]
        assert max(costs) / (sum(costs) / world) < 1.02
        assert max(costs) <= max(before) + 1e-9
    assert sharding.balanced_cuts(100, 1, [0, 99], [5], 1, 1, 10) == [0, 99]
    assert sharding.balanced_cuts(9, 4, [0, 2, 4, 6, 8], [1, 100, 1, 1], 0.0, 1.0, 4) == [0, 2, 4, 6, 8]   # no room to move: unchanged


def _balanced_worker(rank, world, port, out):
    """bench.py's balanced_leg in miniature: count on the even cut, all-gather, re-cut by estimated cost, extract on the new cut."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = _field()
    lo, hi = orc.minmax(f)   # the global range (the all-reduce itself is covered by the test above)
    even = [sharding.slab_bounds(N, world, r)[0] for r in range(world)] + [N - 1]
    r0 = _band_raw_slab(np.ascontiguousarray(f[even[rank]:even[rank + 1] + 1]), even[rank], N, lo, hi)
    per_rank, _, _, _ = sharding.gather_counts(dist, r0["active"], r0["total"])
    cuts = sharding.balanced_cuts(N, world, even, [v for (_, v) in per_rank], 1e-3, 1.0, N * N)
    z0, z1 = cuts[rank], cuts[rank + 1]
    r = _band_raw_slab(np.ascontiguousarray(f[z0:z1 + 1]), z0, N, lo, hi)
    per_rank2, voff, aoff, totals = sharding.gather_counts(dist, r["active"], r["total"])
    np.savez(os.path.join(out, "bal%d.npz" % rank), pos=r["pos"][:r["total"]], norm=r["norm"][:r["total"]], voff=voff[rank], totals=np.array(totals),
             cuts=np.array(cuts), even=np.array(even), before=np.array([v for (_, v) in per_rank]), after=np.array([v for (_, v) in per_rank2]))
    dist.destroy_process_group()


def test_cost_balanced_slabs_still_concatenate_to_the_single_rank_mesh(tmp_path):
    world = 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_balanced_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    f = _field()
    lo, hi = orc.minmax(f)
    single = _band_raw_slab(f, 0, N, lo, hi)
    parts = [np.load(os.path.join(str(tmp_path), "bal%d.npz" % r)) for r in range(world)]
    t = single["total"]
    assert all(np.array_equal(p["cuts"], parts[0]["cuts"]) for p in parts)          # every rank derived the same cuts
    assert tuple(parts[0]["totals"]) == (single["active"], t)
    assert np.array_equal(np.concatenate([p["pos"] for p in parts]).view(np.uint32), single["pos"][:t].view(np.uint32))
    assert np.array_equal(np.concatenate([p["norm"] for p in parts]).view(np.uint32), single["norm"][:t].view(np.uint32))
    before, after = parts[0]["before"], parts[0]["after"]
    assert after.max() <= before.max()                                                  # the heaviest rank did not get heavier
