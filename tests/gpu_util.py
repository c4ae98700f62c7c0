"""Helpers shared by the GPU parity tests and the golden generator."""
import numpy as np
import torch

import gpucadforam_b200 as g
from gpucadforam_b200 import _capi

import cases
import oracle_py as orc
import ref_py as ref

GP_T = torch.int32  # grid_points viewed as 4 x int32 on the device


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda", dtype=dtype).contiguous()


def gp_zeros(n):
    return torch.zeros((n, 4), dtype=torch.int32, device="cuda")


def gp_to_numpy(t):
    return t.cpu().numpy().view(orc.GP_DTYPE).reshape(-1)


def gp_from_numpy(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32).reshape(-1, 4)).cuda()


def ref_field_buffer(n):
    """Output buffer of n floats for a REFERENCE field kernel: several of them (e.g. implicit_pyramid_frustum_kernel, Modelling.cu:555)
    write without a `tx < size` guard, i.e. up to the end of their last 1024-thread block -- the tail is allocated so that the
    reference's overrun stays inside this tensor (compute-sanitizer memcheck flags it otherwise)."""
    return torch.zeros((n + 1023) // 1024 * 1024 + 1024, device="cuda")[:n]


def max_verts_for(dims):
    return max(4 * dims[0] * dims[1] * dims[2], 300000)  # main.cu:2850


def bits(t):
    return t.contiguous().view(torch.int32)


def assert_bits_equal(a, b, what):
    a, b = bits(a), bits(b)
    nbad = int((a != b).sum())
    assert nbad == 0, "%s: %d of %d words differ" % (what, nbad, a.numel())


def ulp_diff(a, b):
    """max |a-b| in units of float32 ulps (sign-magnitude ordered ints)."""
    ia = a.contiguous().view(torch.int32).to(torch.int64)
    ib = b.contiguous().view(torch.int32).to(torch.int64)
    ia = torch.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = torch.where(ib < 0, -(ib & 0x7fffffff), ib)
    return int((ia - ib).abs().max())


def stage_dict(scr, ncell, active):
    return dict(voxelVerts=scr.voxelVerts[:ncell].cpu().numpy().astype(np.uint32),
                voxelOccupied=scr.voxelOccupied[:ncell].cpu().numpy().astype(np.uint32),
                voxelVertsScan=scr.voxelVertsScan[:ncell].cpu().numpy().astype(np.uint32),
                voxelOccupiedScan=scr.voxelOccupiedScan[:ncell].cpu().numpy().astype(np.uint32),
                compVoxelArray=scr.compVoxelArray[:active].cpu().numpy().astype(np.uint32))


def compare_extractions(a, b, what, exact_mesh=True, rtol=1e-5):
    """a, b: dicts {stage arrays, pos, norm, active, total}; stage arrays and counts must be bit-exact."""
    assert a["active"] == b["active"], "%s: activeVoxels %d vs %d" % (what, a["active"], b["active"])
    assert a["total"] == b["total"], "%s: totalVerts %d vs %d" % (what, a["total"], b["total"])
    for k in ("voxelVerts", "voxelOccupied", "voxelVertsScan", "voxelOccupiedScan", "compVoxelArray"):
        if a.get(k) is not None and b.get(k) is not None:
            assert np.array_equal(a[k], b[k]), "%s: stage array %s differs" % (what, k)
    t = a["total"]
    pa, pb = np.asarray(a["pos"][:t]), np.asarray(b["pos"][:t])
    na, nb = np.asarray(a["norm"][:t]), np.asarray(b["norm"][:t])
    if exact_mesh:
        assert np.array_equal(pa.view(np.uint32), pb.view(np.uint32)), "%s: positions not bit-identical" % what
        assert np.array_equal(na.view(np.uint32), nb.view(np.uint32)), "%s: normals not bit-identical" % what
    else:
        # north_star tolerance: 1e-5 relative (positions relative to the grid extent, normals to their length)
        scale = max(1.0, float(np.abs(pa[:, :3]).max())) if t else 1.0
        assert np.allclose(pa, pb, rtol=0, atol=rtol * scale), "%s: positions differ by %g" % (what, np.abs(pa - pb).max())
        nscale = np.maximum(np.linalg.norm(na[:, :3].astype(np.float64), axis=1, keepdims=True), 1e-3) if t else 1.0
        assert np.all(np.abs(na[:, :3].astype(np.float64) - nb[:, :3]) <= rtol * 10 * nscale + 1e-7), "%s: normals differ" % what
        assert np.array_equal(na[:, 3], nb[:, 3]), "%s: norm.w differs" % what


def mine_result(scr, mesh, dims, active, total):
    ncell = (dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1)
    r = stage_dict(scr, ncell, active)
    r.update(pos=mesh.pos.cpu().numpy(), norm=mesh.norm.cpu().numpy(), active=active, total=total)
    return r
