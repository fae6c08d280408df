"""The C++ host mirror (include/dcp.hpp) above the C ABI, driven by tests/cpp/host_mirror_test.cpp.

The reference is C++; this is the shape its own code would use: dcp::BoussinesqModel with the reference's member
names, matrices that satisfy the deal.II vmult concept, dcp::Error derived from std::runtime_error."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-C", ROOT, "tests/cpp/host_mirror_test"], check=True, env=env, capture_output=True)
    assert os.path.exists(EXE)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_host_mirror_compiles_and_fails_loudly_without_a_device():
    _build()
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("spec", ["geometry=shell,refine=1", "geometry=shell,refine=2"])
def test_host_mirror_matches_oracle_on_gpu(spec):
    if not os.path.exists(EXE):
        _build()
    r = subprocess.run([EXE, spec], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK"), r.stdout
