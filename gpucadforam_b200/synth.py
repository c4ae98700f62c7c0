"""Deterministic synthetic INPUTS for the BASELINE.json configs.

Nothing here is on the measured path: these functions only manufacture the arrays a caller of
the reference would already hold (control-grid phases phi_h, Fourier coefficients c_h, a
topology-optimised density).  torch is used as an array library so the same code runs on the
CPU (tests) and on the GPU (bench set-up at 512^3 and above).
"""
import math

import numpy as np
import torch

# spatial_lattice_run (main.cu:3949-3956): k, j, i in [-2, 2], first 62 of 125 (everything before DC)
HARMONICS = [(i, j, k) for k in range(-2, 3) for j in range(-2, 3) for i in range(-2, 3)][:62]


def _unit_cell_coefficients(expr, n):
    ax = ((np.arange(n, dtype=np.float64) / (n - 1)) - 0.5) / 0.5
    a = (3.14 * ax).astype(np.float32).astype(np.float64)
    zz, yy, xx = np.meshgrid(a, a, a, indexing="ij")
    f = expr(xx, yy, zz)
    spec = np.fft.fftn(f) / f.size
    out = []
    for (i, j, k) in HARMONICS:
        c = spec[k % n, j % n, i % n]
        out.append((float(np.float32(c.real)), float(np.float32(c.imag))))
    return out


def gyroid_coefficients(n=61):
    """c_h of the gyroid unit cell truncated to 5x5x5 (main.cu:3578-3711 in spirit): FFT of the
    reference's unit-cell expression (Fft_lattice.cu:28-34) sampled on n^3 points, divided by n^3."""
    return _unit_cell_coefficients(lambda x, y, z: np.cos(x) * np.sin(y) + np.cos(y) * np.sin(z) + np.cos(z) * np.sin(x), n)


def schwarz_p_coefficients(n=61):
    """c_h of the Schwarz-P unit cell, lattice type 1 of create_lattice (Fft_lattice.cu:37-40), same recipe as the gyroid set."""
    return _unit_cell_coefficients(lambda x, y, z: np.cos(x) + np.cos(y) + np.cos(z), n)


def phase_grids(cx, cy, cz, device="cpu", z0=0, cz_total=None, harmonics=None, periods=6.0, dtype=torch.float32):
    """phi_h on a control grid (cx, cy, cz) -> tensor [nh, cz, cy, cx].

    Smooth spatially varying period and rotation about z (round-lattice / variable-period flavour of
    finding_phi, Gratings.cu:100-417): K(r) = 2 pi / P(r) * Rz(theta(r)) (i, j, k), phi = K . r.
    z0 / cz_total describe a z-slab of a taller global control grid so multi-rank runs generate
    exactly the planes a single rank would."""
    harmonics = HARMONICS if harmonics is None else harmonics
    cz_total = cz if cz_total is None else cz_total
    n = float(max(cx, cy))  # independent of the z extent so z-slab / weak-scaling runs see the same lattice period
    x = torch.arange(cx, device=device, dtype=torch.float64) - (cx - 1) / 2.0
    y = torch.arange(cy, device=device, dtype=torch.float64) - (cy - 1) / 2.0
    z = torch.arange(z0, z0 + cz, device=device, dtype=torch.float64) - (cz_total - 1) / 2.0
    Z, Y, X = torch.meshgrid(z, y, x, indexing="ij")
    period = (n / periods) * (1.0 + 0.35 * (X / n) + 0.2 * (Y / n) * (Z / n))
    theta = 0.6 * (Z / n) + 0.3 * (X / n) * (Y / n)
    ct, st = torch.cos(theta), torch.sin(theta)
    u = (ct * X + st * Y) / period
    v = (-st * X + ct * Y) / period
    w = Z / period
    out = torch.empty((len(harmonics), cz, cy, cx), device=device, dtype=dtype)
    for h, (i, j, k) in enumerate(harmonics):
        out[h] = (2.0 * math.pi * (i * u + j * v + k * w)).to(dtype)
    return out


def cantilever_density(nx, ny, nz, seed=1234, struts=40, sigma=1.5, device="cpu"):
    """Synthetic topology-optimised cantilever on a COARSE grid (nx, ny, nz): union of capsule struts
    between the clamped face x=0 and a tip load point, 1 inside / 0.07 outside (MinDens,
    ImguiApp.cpp:273), Gaussian-blurred.  Returned as [nz, ny, nx] float32."""
    rng = np.random.RandomState(seed)
    x = torch.arange(nx, device=device, dtype=torch.float32)
    y = torch.arange(ny, device=device, dtype=torch.float32)
    z = torch.arange(nz, device=device, dtype=torch.float32)
    Z, Y, X = torch.meshgrid(z, y, x, indexing="ij")
    tip = np.array([nx - 3.0, ny / 2.0, nz / 2.0])
    nodes = [np.array([1.0, rng.uniform(2, ny - 3), rng.uniform(2, nz - 3)]) for _ in range(struts // 2)]
    nodes += [np.array([rng.uniform(nx * 0.2, nx * 0.8), rng.uniform(2, ny - 3), rng.uniform(2, nz - 3)]) for _ in range(struts // 2)]
    dens = torch.full((nz, ny, nx), 0.07, device=device, dtype=torch.float32)
    r = max(1.5, min(ny, nz) / 28.0)
    for s in range(struts):
        a = nodes[s]
        b = tip if s % 3 == 0 else nodes[(s * 7 + 3) % len(nodes)]
        ab = b - a
        l2 = float(ab @ ab) + 1e-9
        t = ((X - a[0]) * ab[0] + (Y - a[1]) * ab[1] + (Z - a[2]) * ab[2]) / l2
        t = t.clamp(0, 1)
        d2 = (X - (a[0] + t * ab[0])) ** 2 + (Y - (a[1] + t * ab[1])) ** 2 + (Z - (a[2] + t * ab[2])) ** 2
        dens = torch.where(d2 <= r * r, torch.ones_like(dens), dens)
    # separable Gaussian blur, sigma in coarse voxels
    rad = int(math.ceil(3 * sigma))
    k = torch.exp(-0.5 * (torch.arange(-rad, rad + 1, device=device, dtype=torch.float32) / sigma) ** 2)
    k = k / k.sum()
    d = dens[None, None]
    for dim in range(3):
        shape = [1, 1, 1, 1, 1]
        shape[2 + dim] = -1
        pad = [0, 0, 0, 0, 0, 0]
        pad[2 * (2 - dim)] = rad
        pad[2 * (2 - dim) + 1] = rad
        d = torch.nn.functional.conv3d(torch.nn.functional.pad(d, pad, mode="replicate"), k.view(shape))
    return d[0, 0].contiguous()
