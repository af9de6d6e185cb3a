"""GPU acceptance test of the drop-in boundary: the reference's OWN tests/*.cpp and examples/*.cpp, compiled unmodified
against cuddhelmholtz_b200/cxx + libcuddh_b200.so (cuddhelmholtz_b200/cxx/Makefile, built in the build container into
oracle/_ref/compat/), run here. Skipped when the prebuilt binaries are absent."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import REF_DRIVER, ROOT

COMPAT = os.path.join(ROOT, "oracle", "_ref", "compat")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(os.path.join(COMPAT, "reftest")), reason="compat binaries not built")]


def run(exe, cwd, timeout=900):
    os.makedirs(os.path.join(cwd, "solution"), exist_ok=True)
    r = subprocess.run([os.path.join(COMPAT, exe)], cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_reference_test_suite_passes_against_the_dropin():
    # tests/test.cpp: 92 checks (quadrature 29, basis 26, gmres 1, linalg 6, mass 24, stiffness 6); cwd = repo root so that
    # the mesh fixture path baked into the binary (oracle/_ref/compat/mesh) resolves
    out = run("reftest", ROOT)
    m = re.search(r"(\d+)\s*/\s*(\d+) tests passed", out)
    assert m, out[-1500:]
    assert "[ - ]" not in out, [l for l in out.splitlines() if "[ - ]" in l]
    assert int(m.group(1)) == int(m.group(2)) == 92


def test_poisson_example_runs(tmp_path):
    out = run("Poisson", str(tmp_path))
    assert "GMRES successfully converged" in out
    u = np.fromfile(tmp_path / "solution" / "poisson.0000")
    assert u.size == (15 * 3 + 1) ** 2 and np.all(np.isfinite(u))


@pytest.mark.skipif(not os.path.exists(REF_DRIVER), reason="oracle/_ref/ref_driver not built")
def test_ddh_example_matches_reference_build(tmp_path):
    # examples/DDH.cpp (uniform_rect(128), degree 3, omega = 2 pi 12.8, FP32 GMRES(20), maxit 100, tol 1e-4) linked against
    # this library vs. the same flow through the unmodified reference library (ref_driver ddh)
    from oracle.rdmp import read_rdmp
    out = run("DDH", str(tmp_path))
    assert "GMRES successfully converged" in out
    U = np.fromfile(tmp_path / "solution" / "ddh.0000")
    omega = 2 * np.pi * 128 / 10
    dump = tmp_path / "ref.bin"
    subprocess.check_call([REF_DRIVER, "ddh", "128", "4", repr(float(omega)), "20", "100", "1e-4", "2024", str(dump)], timeout=1500)
    r = read_rdmp(str(dump))
    assert bool(r["success"][0])
    it = int(re.search(r"After (\d+) iterations", out).group(1))
    assert abs(it - int(r["num_iter"][0])) <= 1, (it, int(r["num_iter"][0]))
    assert np.linalg.norm(U - r["U"]) / np.linalg.norm(r["U"]) < 1e-3
