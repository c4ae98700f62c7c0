"""CPU check of the fast SVL field algorithm (GCB_OPT_FAST_FIELD): its numpy restatement stays within the stated bound of the
oracle's exact field, at both upsampling ratios the tests use and for large phases."""
import numpy as np
import pytest

import cases
import oracle_py as orc
from fast_field_model import bound, fast_field


@pytest.mark.parametrize("name", ["SVL", "SVL4"])
def test_fast_field_model_within_stated_bound_of_the_oracle_field(name):
    cfg = getattr(cases, name)
    phi, coef = cases.svl_inputs(cfg)
    ratio = int(round(1.0 / cfg["d"][0]))
    exact = orc.svl_field(phi, coef, cfg["fdims"], cfg["d"])
    fast = fast_field(phi, coef, cfg["fdims"], ratio)
    err = float(np.abs(fast.astype(np.float64) - exact).max())
    b = bound(phi, coef)
    print("%s: max |fast - exact| = %.3g, bound %.3g, field range [%.3f, %.3f]" % (name, err, b, exact.min(), exact.max()))
    assert err <= b + 2e-5  # 2e-5: the oracle's own glibc-vs-libdevice tolerance (the exact GPU field is compared with atol 2e-5 too)


def test_fast_field_model_large_phases():
    """|phi| ~ 250 rad (the 512^3 bench workload): the per-cell reduction keeps the error at the 1e-6 level, far below ulp(phi)."""
    from gpucadforam_b200 import synth
    nh = 8
    phi = synth.phase_grids(8, 8, 8, periods=2.0, harmonics=synth.HARMONICS[:nh]).numpy()
    phi = (phi + np.float32(231.7) * np.array([1, -1, 1, 1, -1, 1, -1, 1], np.float32)[:, None, None, None]).astype(np.float32)
    assert float(np.abs(phi).max()) > 200.0
    coef = synth.gyroid_coefficients()[:nh]
    coef = [(c[0] + 0.1, c[1] - 0.05) for c in coef]
    fd = (32, 32, 32)
    exact = orc.svl_field(phi, coef, fd, (0.25, 0.25, 0.25))
    fast = fast_field(phi, coef, fd, 4)
    err = float(np.abs(fast.astype(np.float64) - exact).max())
    print("large phases: max |fast - exact| = %.3g, bound %.3g" % (err, bound(phi, coef)))
    assert err <= bound(phi, coef) + 2e-5
