"""Cross-validation container (include/dcp_dump.h).

CPU: the container written by the plain-C writer is read back bit-exactly; a harness problem survives the round
trip and the oracle gives identical results on the reloaded arrays (so a dump carries everything the path needs).
GPU: the device model built from a dump equals the one built from the live problem; and, when the environment names
a dump made by a real deal.II build of the reference (DCP_REFERENCE_DUMP), the CUDA path is compared against the
matrices and right-hand sides stored in it at the tolerances of BASELINE.json."""
import os
import subprocess
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_writer_and_python_reader_agree(tmp_path):
    import dycore_b200  # noqa: F401
    from dycore_b200 import dump
    src = tmp_path / "w.c"
    src.write_text(textwrap.dedent("""
        #include <dcp_dump.h>
        int main(int argc, char** argv) {
          double a[3] = {1.5, -2.25, 1e-300};
          int32_t b[5] = {1, -2, 3, -4, 2147483647};
          int64_t c[2] = {1ll << 40, -7};
          FILE* f = dcp_dump_open(argv[1]);
          if (!f) return 1;
          int rc = dcp_dump_array(f, "values", DCP_DUMP_F64, 3, a) | dcp_dump_array(f, "idx", DCP_DUMP_I32, 5, b) |
                   dcp_dump_array(f, "rowptr", DCP_DUMP_I64, 2, c) | dcp_dump_array(f, "empty", DCP_DUMP_F64, 0, 0) |
                   dcp_dump_scalar(f, "n_cells", 384) | dcp_dump_array(f, "spec", DCP_DUMP_I8, 23, "geometry=shell,refine=2");
          return rc | dcp_dump_close(f);
        }"""))
    exe = tmp_path / "w"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = tmp_path / "t.dcpd"
    subprocess.run([str(exe), str(out)], check=True)
    arrays, scalars, spec = dump.read_dump(str(out))
    assert np.array_equal(arrays["values"], [1.5, -2.25, 1e-300])
    assert np.array_equal(arrays["idx"], [1, -2, 3, -4, 2147483647]) and arrays["idx"].dtype == np.int32
    assert np.array_equal(arrays["rowptr"], [1 << 40, -7]) and arrays["empty"].size == 0
    assert scalars == {"n_cells": 384} and spec == {"geometry": "shell", "refine": "2"}


def test_problem_round_trip_feeds_the_oracle(problem_factory, tmp_path):
    import dycore_b200  # noqa: F401
    from dycore_b200 import dump, params
    from oracle import oracle as orc
    from util import synthetic_fields
    P = problem_factory(geometry="shell", refine=1)
    path = str(tmp_path / "p.dcpd")
    dump.dump_problem(P, path)
    Q = dump.DumpProblem(path)
    assert sorted(Q.names()) == sorted(P.names())
    for n in P.names():
        assert np.array_equal(P[n], Q[n]) and P[n].dtype == Q[n].dtype, n
    prm = orc.params_from(params.NAMED["shell_3d_classic"])
    u, T = synthetic_fields(P)
    a, ra = orc.assemble_nse_system(P, prm, u, T)
    b, rb = orc.assemble_nse_system(Q, prm, u, T)
    assert np.array_equal(a, b) and np.array_equal(ra, rb)


@pytest.mark.gpu
def test_model_from_dump_equals_model_from_problem(problem_factory, tmp_path):
    import dycore_b200  # noqa: F401
    from dycore_b200 import device, dump, params
    from util import synthetic_fields
    P = problem_factory(geometry="shell", refine=1)
    path = str(tmp_path / "p.dcpd")
    dump.dump_problem(P, path)
    Q = dump.DumpProblem(path)
    mp = params.NAMED["shell_3d_classic"]
    u, T = synthetic_fields(P)
    ctx = device.Context(0)
    out = []
    for prob in (P, Q):
        m = device.BoussinesqModel.from_problem(ctx, prob, mp)
        m.assemble_nse_system(u, T)
        out.append((m.nse_matrix.block(0, 0).values(), m.nse_rhs))
        m.close()
    ctx.close()
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-12 * np.abs(out[0][0]).max()   # atomics: summation order only
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-12 * np.abs(out[0][1]).max()


@pytest.mark.gpu
def test_against_a_dump_of_the_real_reference():
    """Arrays expected besides the inputs: ref.nse.bIJ.val, ref.pre.bIJ.val, ref.temp.mass.val, ref.temp.stiff.val,
    ref.nse_rhs, ref.temp_rhs, in.old_nse, in.old_temp, in.nse_solution."""
    path = os.environ.get("DCP_REFERENCE_DUMP")
    if not path:
        pytest.skip("DCP_REFERENCE_DUMP not set (needs a deal.II build of the reference, absent in this image)")
    import dycore_b200  # noqa: F401
    from dycore_b200 import device, dump, params
    from util import rel_err_max
    Q = dump.DumpProblem(path)
    mp = params.NAMED[Q.spec.get("parameters", "shell_3d_classic")]
    ctx = device.Context(0)
    m = device.BoussinesqModel.from_problem(ctx, Q, mp)
    m.assemble_nse_system(Q["in.old_nse"], Q["in.old_temp"])
    m.assemble_nse_preconditioner()
    m.assemble_temperature_matrix()
    m.assemble_temperature_rhs(Q["in.old_temp"], Q["in.nse_solution"])
    nb = 3 if "feec" in str(Q.spec.get("family", "classic")) else 2
    for i in range(nb):
        for j in range(nb):
            for which, mat in (("nse", m.nse_matrix), ("pre", m.nse_preconditioner_matrix)):
                name = f"ref.{which}.b{i}{j}.val"
                if name in Q:
                    assert rel_err_max(mat.block(i, j).values(), Q[name]) <= 1e-12, name
    assert rel_err_max(m.temperature_mass_matrix.values(), Q["ref.temp.mass.val"]) <= 1e-12
    assert rel_err_max(m.temperature_stiffness_matrix.values(), Q["ref.temp.stiff.val"]) <= 1e-12
    assert rel_err_max(m.nse_rhs, Q["ref.nse_rhs"]) <= 1e-12
    assert rel_err_max(m.temperature_rhs, Q["ref.temp_rhs"]) <= 1e-12
    m.close()
    ctx.close()
