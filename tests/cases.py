"""Small, deterministic test cases shared by the golden generator (tests/golden/make_golden.py, runs the
reference CUDA kernels on the GPU box), the CPU oracle tests and the GPU parity tests.  Sizes are
chosen so the CPU oracle finishes in well under a second."""
import numpy as np

# reference UI defaults: bound_isoVal / bound_isoValone / bound_isoValtwo (ImguiApp.cpp:103-105)
ISO_MASK, BAND_LO, BAND_HI = 0.25, 0.20, 0.30

GYROID = dict(n=32, type=0)                     # config 1 in miniature (create_lattice -> normalise -> band -> latticeone)
TPMS_TYPES = [0, 1, 2, 3, 4, 5]

# config 2 in miniature: fine grid 48^3, dx2 = 0.5 (coarse 24^3)
CSG = dict(dims=(48, 48, 48), d=(0.5, 0.5, 0.5),
           sphere=dict(center=(0.0, 0.0, 0.0), radius=7.5, thickness=2.0),
           cuboid=dict(center=(1.0, 0.5, -0.5), angles=(0.3, 0.2, 0.1), xw=17.0, yw=9.0, zw=11.0),
           cylinder=dict(center=(0.0, 0.0, 0.0), axis=(0.0, 0.0, 1.0), radius=3.4, tr=2.0, ta=40.0))

# region / domain display (f-4): vol_topo <- retained sphere, primitive_fixed <- retained rotated cuboid, dynamic <- sphere field
# (32*32*24 points is a multiple of 1024: the reference's cuboid kernel has no bounds guard)
REGION = dict(dims=(32, 32, 24), d=(0.5, 0.5, 0.5),
              topo_sphere=dict(center=(-2.5, 1.0, 0.5), radius=3.0, thickness=1.0),
              cuboid=dict(center=(0.5, 0.0, 0.0), angles=(0.1, 0.2, 0.3), xw=9.0, yw=7.0, zw=6.0),
              dyn_sphere=dict(center=(2.0, 0.0, 0.0), radius=4.0, thickness=1.0))
REGION_MODES = ["make_region", "show_region", "show_domain"]

PRIMS = dict(dims=(40, 36, 44), d=(0.5, 0.5, 0.5), center=(0.7, -0.4, 0.3), angles=(0.3, 0.2, 0.1))

# configs 3/4 in miniature: control 16x16x8 -> fine 32x32x16 (ratio 2, the app's own), and control 8^3 -> fine 32^3 (ratio 4)
# (point counts are multiples of 1024: the reference's min/max reduction reads uninitialised shared memory otherwise)
SVL = dict(cdims=(16, 16, 8), fdims=(32, 32, 16), d=(0.5, 0.5, 0.5), nh=10)
SVL4 = dict(cdims=(8, 8, 8), fdims=(32, 32, 32), d=(0.25, 0.25, 0.25), nh=6)

# config 5 in miniature: coarse 24x12x12 -> fine 48x24x24, iso = VolumeFraction 0.4
TOPO = dict(cdims=(24, 12, 12), fdims=(48, 24, 24), d=(0.5, 0.5, 0.5), iso=0.4)


def svl_inputs(cfg):
    from gpucadforam_b200 import synth
    cx, cy, cz = cfg["cdims"]
    phi = synth.phase_grids(cx, cy, cz, periods=3.0, harmonics=synth.HARMONICS[20:20 + cfg["nh"]]).numpy()
    coef = synth.gyroid_coefficients()[20:20 + cfg["nh"]]
    # make sure no coefficient is exactly zero so every harmonic contributes
    coef = [(c[0] + 0.05 * ((i % 3) - 1), c[1] + 0.03 * ((i % 2) * 2 - 1)) for i, c in enumerate(coef)]
    coef = [(float(np.float32(a)), float(np.float32(b))) for a, b in coef]
    return phi, coef


def topo_coarse(cfg):
    from gpucadforam_b200 import synth
    cx, cy, cz = cfg["cdims"]
    return synth.cantilever_density(cx, cy, cz, struts=12, sigma=1.0).numpy()


# SVL phase solve (f-2) in miniature: control grid 16x12x8, reference defaults latticetype 'r' / period_type 2 (main.cu:3959-3962)
PHASE = dict(dims=(16, 12, 8), d=(1.0, 1.0, 1.0), harmonics=[(1, 0, 0), (0, 1, 0), (1, -2, 1), (-2, 1, 2)], iters=500, end_res=0.01)


def phase_period(cfg):
    nx, ny, nz = cfg["dims"]
    rng = np.random.RandomState(41)
    return rng.uniform(3.0, 7.0, nx * ny * nz).astype(np.float32)
