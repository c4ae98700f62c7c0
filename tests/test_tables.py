"""Marching-cubes table invariants and identity with the reference's tables.h
whenever /root/reference is present (it is not on the GPU box)."""
import os
import re
import subprocess
import sys

import numpy as np

import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TABLES = "/root/reference/src/tables.h"


def test_numverts_matches_tritable():
    tri, nv = orc.tables()
    for c in range(256):
        used = [e for e in tri[c] if e != 255]
        assert len(used) == nv[c]
        assert nv[c] % 3 == 0 and nv[c] <= 15
        assert all(e < 12 for e in used)
        # terminators only at the tail
        assert list(tri[c][:len(used)]) == used
    assert nv[0] == 0 and nv[255] == 0
    assert int(nv.sum()) == 2460
    # the table is NOT complement-symmetric, so inside/outside polarity matters
    assert any(nv[c] != nv[255 - c] for c in range(256))


def test_every_triangle_uses_crossing_edges():
    """An edge referenced by case c must join corners of different sign in c."""
    tri, nv = orc.tables()
    ends = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    for c in range(256):
        for e in tri[c][:nv[c]]:
            a, b = ends[e]
            assert ((c >> a) & 1) != ((c >> b) & 1)


def test_product_library_tables_equal_oracle():
    import ctypes as C
    from gpucadforam_b200 import _capi
    lib = _capi.load()
    tri = (C.c_uint * 4096)()
    nv = (C.c_uint * 256)()
    lib.gcb_tables(tri, nv)
    otri, onv = orc.tables()
    assert np.array_equal(np.array(tri[:]).reshape(256, 16), otri)
    assert np.array_equal(np.array(nv[:]), onv)


def test_packed_table_equals_reference_header():
    if not os.path.exists(REF_TABLES):
        import pytest
        pytest.skip("reference tree not present")
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "tools", "pack_mc_tables.py"), "--check", REF_TABLES,
                          os.path.join(ROOT, "gpucadforam_b200", "csrc", "mc_tables_packed.inc")])
    assert rc == 0
    src = open(REF_TABLES).read()
    m = re.search(r"numVertsTable\[256\]\s*=\s*\{(.*?)\};", src, re.S)
    ref_nv = [int(x) for x in re.findall(r"\d+", m.group(1))]
    assert list(orc.tables()[1]) == ref_nv
