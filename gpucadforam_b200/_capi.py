"""ctypes binding of libgpucad_b200.so (C ABI declared in include/gpucad_b200.h).

The library is the product; this module only loads it and declares signatures.  There is no
fallback: if the shared object is missing the import raises, and if no CUDA device is present
`gcb_create` fails (the C library has no CPU path).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCB_LIB_PATH") or os.path.join(_HERE, "libgpucad_b200.so")  # GCB_LIB_PATH: A/B builds of the library


class Uint3(C.Structure):
    _fields_ = [("x", C.c_uint), ("y", C.c_uint), ("z", C.c_uint)]


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class PitchedPtr(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pitch", C.c_size_t), ("xsize", C.c_size_t), ("ysize", C.c_size_t)]


class Slab(C.Structure):
    _fields_ = [("z0", C.c_uint), ("gnz", C.c_uint)]


P = C.c_void_p
U = C.c_uint
I = C.c_int
F = C.c_float
ULL = C.c_ulonglong
PU = C.POINTER(C.c_uint)
PULL = C.POINTER(C.c_ulonglong)
PF = C.POINTER(C.c_float)

# name -> (restype, argtypes).  Every symbol include/gpucad_b200.h declares is listed here;
# tests/test_capi_symbols.py checks the two stay in sync.
SIGNATURES = {
    "gcb_create": (I, [C.POINTER(P), I, P]),
    "gcb_destroy": (I, [P]),
    "gcb_last_error": (C.c_char_p, [P]),
    "gcb_set_stream": (I, [P, P]),
    "gcb_set_options": (I, [P, U]),
    "gcb_launch_count": (ULL, [P]),
    "gcb_reset_launch_count": (None, [P]),
    "gcb_enable_kernel_timing": (I, [P, I]),
    "gcb_last_extract_kernel_ms": (F, [P]),
    "gcb_last_field_kernel_ms": (F, [P]),
    "gcb_allocateTextures_s": (I, [P, C.POINTER(P), C.POINTER(P)]),
    "gcb_destroyAllTextureObjects": (I, [P]),
    "gcb_tables": (None, [PU, PU]),
    "gcb_computeIsosurface": (I, [P, P, Uint3, P, P, F, U, P, P, P, P, Uint3, Uint3, Uint3, Float3, Float3, PU, PU, P, U,
                                  P, P, P, P, F, F, I, I, I, I, I, I, I, I, I, C.POINTER(C.c_size_t)]),
    "gcb_computeIsosurface_region": (I, [P, P, P, F, U, P, P, P, P, Uint3, Uint3, Uint3, Float3, Float3, PU, PU, P, U,
                                         P, P, P, P, P, F, F, I, I, I, I, I, I, I, I, I, I, I, P]),
    "gcb_computeIsosurface_lattice": (I, [P, P, P, P, F, U, P, P, P, P, Uint3, Uint3, Uint3, Float3, Float3, PU, PU, P, U,
                                          P, P, F, F, F, F]),
    "gcb_computeIsosurface_latticeone": (I, [P, P, P, P, F, U, P, P, P, P, Uint3, Uint3, Uint3, Float3, Float3, PU, PU, P, U,
                                             P, F, F]),
    "gcb_computeIsosurface_2": (I, [P, P, P, F, U, P, P, P, P, Uint3, Uint3, Uint3, Float3, Float3, PU, PU, P, U,
                                    P, P, P, P, F, P, P]),
    "gcb_computeIsosurface_topo": (I, [P, P, P, F, U, P, P, P, P, Uint3, Uint3, Uint3, Float3, Float3, PU, PU, P, U,
                                       P, P, P, P, F, P, P, I, P]),
    "gcb_copy_parameter": (I, [P, P, F, Uint3, Uint3, Uint3, Float3, U, P, P, P, I, I, F, F, I, I, I]),
    "gcb_patch_topo_field": (I, [P, P, I, I, I, P]),
    "gcb_distance_from_line": (I, [P, P, Float3, Float3, F, F, F, I, I, I, F, F, F, I]),
    "gcb_sphere_with_center": (I, [P, P, Float3, F, F, I, I, I, F, F, F, I]),
    "gcb_cuboid": (I, [P, P, Float3, Float3, F, F, F, I, I, I, F, F, F]),
    "gcb_cuboid_shell": (I, [P, P, Float3, Float3, F, F, F, F, I, I, I, F, F, F]),
    "gcb_torus_with_center": (I, [P, P, Float3, Float3, F, F, I, I, I, F, F, F]),
    "gcb_cone_with_base_radius_height": (I, [P, P, Float3, Float3, F, F, I, I, I, F, F, F]),
    "gcb_cone_frustum": (I, [P, P, Float3, Float3, F, F, F, I, I, I, F, F, F]),
    "gcb_pyramid_frustum": (I, [P, P, Float3, Float3, F, F, F, F, F, I, I, I, F, F, F]),
    "gcb_create_lattice": (I, [P, P, U, U, U, U, U]),
    "gcb_GPU_buffer_normalise_buffer": (I, [P, P, P, I]),
    "gcb_GPU_buffer_normalise_four": (I, [P, P, P, P, C.c_size_t, I, I, I, F, F]),
    "gcb_GPU_buffer_normalise_three": (I, [P, P, P, C.c_size_t, F, F]),
    "gcb_period_data": (I, [P, P, I, I, I, F, F, F, F, F, F, C.c_char]),
    "gcb_angle_data": (I, [P, P, I, I, I, F, F, F, F, F, F, C.c_char]),
    "gcb_grating": (I, [P, P, I, I, I, F, F, F]),
    "gcb_refine": (I, [P, P, I, I, I, F, F, F]),
    "gcb_svl": (I, [P, P, P, I, I, I, I, P]),
    "gcb_topo_field": (I, [P, P, P, F, I, I, I]),
    "gcb_primitive_field": (I, [P, P, P, P, F, I, I, I, I, I]),
    "gcb_setupTexture": (I, [P, I, I, I]),
    "gcb_copytotexture": (I, [P, P, PitchedPtr, I, I, I]),
    "gcb_updateTexture": (I, [P, PitchedPtr]),
    "gcb_deleteTexture": (I, [P]),
    "gcb_file_write_obj": (I, [P, P, U, C.c_char_p]),
    "gcb_unit_lattice_spectrum": (I, [P, P, I, I, I, I, P]),
    "gcb_finding_phi": (I, [P, P, P, I, I, I, I, I, I, F, F, F, C.c_char, I, F, F, F, F, F, F, I]),
    "gcb_GPUCG_lattice": (I, [P, P, I, I, I, I, I, F, C.POINTER(I), PF]),
    "gcb_svl_phase_solve": (I, [P, P, P, I, C.POINTER(I), I, I, I, F, F, F, C.c_char, I, F, F, F, F, F, F, I, I, F, C.POINTER(I), PF]),
    "gcb_svl_field": (I, [P, P, P, I, PF, I, I, I, I, I, I, I, Slab, F, F, F, I, P]),
    "gcb_svl_field_host": (I, [P, P, P, P, I, PF, I, I, I, I, I, I, I, Slab, F, F, F, P]),
    "gcb_minmax": (I, [P, P, C.c_size_t, PF, PF]),
    "gcb_extract_band_raw": (I, [P, P, F, F, F, F, F, Uint3, Slab, Float3, Float3, P, P, ULL, P, I, PULL, PULL]),
    "gcb_extract_band_raw_dev": (I, [P, P, P, F, F, F, Uint3, Slab, Float3, Float3, P, P, ULL, P, I, PULL, PULL]),
    "gcb_band_lattice_from_raw": (I, [P, P, Uint3, F, F, F, Float3, Float3, P, P, ULL, P, PULL, PULL, PF]),
    "gcb_tpms_lattice": (I, [P, P, U, Uint3, F, F, F, Float3, Float3, P, P, ULL, P, PULL, PULL, PF]),
    "gcb_density_surface": (I, [P, P, I, I, I, P, I, I, I, F, F, F, F, Float3, Float3, P, P, ULL, P, PULL, PULL]),
    "gcb_csg_retain_primitive": (I, [P, I, Float3, Float3, PF, I, I, P, P, I, I, I, F, F, F, F, I, I, I]),
    "gcb_svl_lattice_host_submit": (I, [P, I, P, P, P, I, PF, I, I, I, I, I, I, F, F, F, F, F, F, Float3, Float3, P, P, ULL]),
    "gcb_svl_lattice_host_wait": (I, [P, I, PULL, PULL, PF]),
    "gcb_svl_slab_host_submit_field": (I, [P, I, P, P, P, I, PF, I, I, I, I, I, I, I, Slab, F, F, F, P]),
    "gcb_svl_slab_host_submit_extract": (I, [P, I, P, P, F, F, F, Uint3, Slab, Float3, Float3, P, P, ULL]),
    "gcb_svl_lattice": (I, [P, P, P, I, PF, I, I, I, I, I, I, F, F, F, F, F, F, Float3, Float3, P, P, ULL, PULL, PULL, PF]),
    "gcb_svl_lattice_host": (I, [P, P, P, P, I, PF, I, I, I, I, I, I, F, F, F, F, F, F, Float3, Float3, P, P, ULL, PULL, PULL, PF]),
    "gcb_multi_create": (I, [C.POINTER(P), I, C.POINTER(I)]),
    "gcb_multi_destroy": (I, [P]),
    "gcb_multi_size": (I, [P]),
    "gcb_multi_ctx": (P, [P, I]),
    "gcb_multi_device": (I, [P, I]),
    "gcb_multi_last_error": (C.c_char_p, [P]),
    "gcb_multi_last_ms": (F, [P]),
    "gcb_slab_bounds": (I, [U, I, I, U, PU, PU]),
    "gcb_control_slab": (I, [U, U, I, I, C.POINTER(I), C.POINTER(I)]),
    "gcb_multi_svl_lattice": (I, [P, C.POINTER(P), C.POINTER(P), I, PF, I, I, C.POINTER(I), C.POINTER(I), I, I, U, F, F, F, F, F, F, Float3, Float3,
                                  C.POINTER(P), C.POINTER(P), PULL, I, PULL, PULL, PULL, PF]),
    "gcb_multi_computeIsosurface_2": (I, [P, C.POINTER(P), C.POINTER(P), C.POINTER(P), Uint3, Float3, Float3, F, F, C.POINTER(P), C.POINTER(P), PULL, C.POINTER(P),
                                          PULL, PULL, PULL, PULL]),
}

GCB_OPT_FILL_STAGE_ARRAYS = 1
GCB_OPT_LEGACY_MEMSET = 2
GCB_OPT_NO_TMA = 4
GCB_OPT_OBJ_HOST = 8
GCB_OPT_ASYNC_FIELDS = 16
GCB_OPT_FAST_FIELD = 32


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            "libgpucad_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C gpucadforam_b200/csrc). There is no CPU or PyTorch fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib
