"""Host logic: the tabulated named configs equal what the reference's .prm files say (when the tree is present)."""
import os

import pytest

REF = "/root/reference/data"
FILES = {"shell_3d_classic": "aqua_planet_shell_test_3d-classic.prm", "shell_3d_feec": "aqua_planet_shell_test_3d-feec.prm",
         "cube_3d": "aqua_planet_cube_test_3d.prm", "annulus_2d": "aqua_planet_test_2d.prm"}


def test_derived_numbers_of_named_configs():
    import dycore_b200  # noqa: F401
    from dycore_b200 import params
    p = params.NAMED["shell_3d_classic"]
    assert (p.inv_re, p.inv_pe, p.R0_scaled, p.R1_scaled) == (1e-2, 1e-3, 1.0, 3.0)
    assert p.g_scale * p.gravity_constant == 1.0
    q = params.NAMED["annulus_2d"]
    assert abs(q.inv_re - 0.1) < 1e-15 and abs(q.R0_scaled - 10.0) < 1e-12 and abs(q.R1_scaled - 30.0) < 1e-12


@pytest.mark.parametrize("name", sorted(FILES))
def test_named_configs_equal_prm_files(name):
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present (GPU box)")
    from dycore_b200 import params
    assert params.read_prm(os.path.join(REF, FILES[name])) == params.NAMED[name]
