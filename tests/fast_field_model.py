"""numpy restatement of the FAST SVL field mode (GCB_OPT_FAST_FIELD, svl_field_fast_kernel in csrc/fields.cu): same operation
order in float32 up to the cosine itself (the GPU uses MUFU.COS, numpy a correctly rounded cos).  TEST INFRASTRUCTURE ONLY:
it checks the algorithm (amplitude / phase form, per-cell phase reduction, packed lerp order) on the CPU and gives the GPU test
a second reference next to the exact field."""
import math

import numpy as np

F = np.float32


def fast_field(phi, coef, fdims, ratio):
    """phi [nh, cz, cy, cx] float32 control grids, coef [(re, im)], fine dims (nx, ny, nz), power-of-two ratio."""
    nh, cz, cy, cx = phi.shape
    nx, ny, nz = fdims
    d = F(1.0 / ratio)
    fx, fy, fz = np.arange(nx), np.arange(ny), np.arange(nz)
    ix, iy, iz = np.minimum(fx // ratio, cx - 1), np.minimum(fy // ratio, cy - 1), fz // ratio
    ix1, iy1 = np.minimum(ix + 1, cx - 1), np.minimum(iy + 1, cy - 1)
    izc, iz1 = np.clip(iz, 0, cz - 1), np.clip(iz + 1, 0, cz - 1)
    # weights exactly as the kernel forms them: even point (f & (r-1)) * d, odd point that + d
    def w(f):
        even = (f // 2) * 2
        w0 = (even % ratio).astype(F) * d
        return np.where(f % 2 == 0, w0, w0 + d).astype(F)
    wx, wy, wz = w(fx), w(fy), w(fz)
    inv2pi, magic = F(0.15915494309189535), F(12582912.0)
    hi, lo = np.array([0x40C90FDB], np.uint32).view(F)[0], F(-1.7484555e-7)
    out = np.zeros((nz, ny, nx), F)
    Z, Y, X = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    for h in range(nh):
        g = phi[h]
        re, im = float(coef[h][0]), float(coef[h][1])
        th, am = F(math.atan2(im, re)), F(math.hypot(re, im))
        corner = g[izc][:, iy][:, :, ix]                                  # tap (0,0,0) of every point's cell
        fq = ((corner * inv2pi).astype(F) + magic).astype(F) - magic
        def red(zi, yi):
            p = g[zi][:, yi][:, :, ix]
            pn = g[zi][:, yi][:, :, ix1]
            pr = (fq.astype(np.float64) * (-float(hi)) + p.astype(np.float64)).astype(F)   # fma: exact product, one rounding
            pr = (fq.astype(np.float64) * (-float(lo)) + pr.astype(np.float64)).astype(F)
            pr = (pr + th).astype(F)                                      # theta is folded into the staged value
            dxv = (pn - p).astype(F)
            return (wx[None, None, :].astype(np.float64) * dxv.astype(np.float64) + pr.astype(np.float64)).astype(F)   # x-lerp, fma
        L00, L01, L10, L11 = red(izc, iy), red(izc, iy1), red(iz1, iy), red(iz1, iy1)
        wyb, wzb = wy[None, :, None], wz[:, None, None]
        D0, D1 = (L01 - L00).astype(F), (L11 - L10).astype(F)
        m0 = (wyb.astype(np.float64) * D0.astype(np.float64) + L00.astype(np.float64)).astype(F)
        m1 = (wyb.astype(np.float64) * D1.astype(np.float64) + L10.astype(np.float64)).astype(F)
        E = (m1 - m0).astype(F)
        v = (wzb.astype(np.float64) * E.astype(np.float64) + m0.astype(np.float64)).astype(F)
        c = np.cos(v.astype(np.float64)).astype(F)
        out = (c.astype(np.float64) * float(am) + out.astype(np.float64)).astype(F)
    return out


def bound(phi, coef):
    """Stated tolerance of the fast mode against the exact field: sum_h |c_h| (ulp(max |phi_h|) / 2 + 4e-6)."""
    b = 0.0
    for h in range(phi.shape[0]):
        m = float(np.abs(phi[h]).max())
        ulp = float(np.spacing(F(m))) if m > 0 else 0.0
        b += math.hypot(float(coef[h][0]), float(coef[h][1])) * (0.5 * ulp + 4e-6)
    return b
