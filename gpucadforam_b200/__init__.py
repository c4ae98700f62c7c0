"""gpucadforam_b200 -- B200-native implicit-field + marching-cubes path of GPUCADforAM.

Python is only the test/bench harness language here: the product is the C-ABI shared library
`libgpucad_b200.so` (hand-written CUDA for sm_100a, see csrc/ and include/gpucad_b200.h) and
the C++ host mirror in host/.  This package loads the library through ctypes and mirrors the
reference's host classes (same method names and argument meaning) on top of torch CUDA tensors,
which are used purely as device-memory handles.

Importing the package without the built library raises ImportError -- there is no fallback.
"""
from . import _capi
from .api import (Context, Isosurface, Modelling, Fft_lattice, Gratings, File_output, MeshBuffers, Scratch,
                  svl_lattice, svl_lattice_host, svl_lattice_host_submit, svl_lattice_host_wait, svl_slab_host_submit_field, svl_slab_host_submit_extract, extract_band_raw, extract_band_raw_dev, band_lattice_from_raw, tpms_lattice, density_surface, csg_retain_primitive, svl_field, svl_field_host, minmax, unit_lattice_spectrum, finding_phi, GPUCG_lattice, svl_phase_solve)

__all__ = ["Context", "Isosurface", "Modelling", "Fft_lattice", "Gratings", "File_output", "MeshBuffers", "Scratch",
           "svl_lattice", "svl_lattice_host", "svl_lattice_host_submit", "svl_lattice_host_wait", "svl_slab_host_submit_field", "svl_slab_host_submit_extract", "extract_band_raw", "extract_band_raw_dev", "band_lattice_from_raw", "tpms_lattice", "density_surface", "csg_retain_primitive", "svl_field", "svl_field_host", "minmax", "unit_lattice_spectrum", "finding_phi", "GPUCG_lattice", "svl_phase_solve", "_capi"]
