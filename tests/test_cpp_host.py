"""C++ host side (aspect_b200/cpp/): the header-only mirror of StokesMatrixFreeHandler over the C ABI and its synthetic box
driver.  CPU: the C++ driver's index maps are bit-exact those of the Python driver; the example links against the
library and fails loudly without a CUDA device.  GPU: the example reproduces the reference's nsinker_gmg_gc screen output."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from aspect_b200 import synthetic
from util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "aspect_b200", "cpp")]
LINK = ["-L", os.path.join(ROOT, "aspect_b200"), "-lsmf_b200", "-Wl,-rpath," + os.path.join(ROOT, "aspect_b200")]

pytestmark = pytest.mark.skipif(not shutil.which("g++"), reason="g++ not available")


def _read(f, dtype):
    n, = struct.unpack("<q", f.read(8))
    return np.frombuffer(f.read(n * np.dtype(dtype).itemsize), dtype=dtype)


@pytest.mark.parametrize("dim,R,bc", [(2, 3, "tangential"), (3, 2, "zero"), (3, 1, "tangential")])
def test_cpp_box_driver_matches_python_driver(tmp_path, dim, R, bc):
    exe = tmp_path / "dump"
    subprocess.check_call(["g++", "-std=c++17", "-O1", *INC, os.path.join(ROOT, "examples", "dump_box_maps.cc"), "-o", str(exe)])
    out = tmp_path / "maps.bin"
    subprocess.check_call([str(exe), str(dim), str(R), bc, str(out)])
    hier = synthetic.BoxHierarchy(dim, R)
    with open(out, "rb") as f:
        for l, L in enumerate(hier.levels):
            n_cells, n_u, n_p = struct.unpack("<3q", f.read(24))
            assert (n_cells, n_u, n_p) == (L.n_cells, L.n_u, L.n_p)
            assert np.array_equal(_read(f, np.uint32).reshape(L.cell_dofs_u.shape), L.cell_dofs_u)
            assert np.array_equal(_read(f, np.uint32).reshape(L.cell_dofs_p.shape), L.cell_dofs_p)
            assert np.array_equal(_read(f, np.uint32), L.constrained_u(bc))
            ch = _read(f, np.int32)
            if l > 0:
                assert np.array_equal(ch.reshape(-1, 1 << dim), hier.children(l))


def _build_example(tmp_path):
    exe = tmp_path / "nsinker_box"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", *INC, os.path.join(ROOT, "examples", "nsinker_box.cc"),
                           *LINK, "-o", str(exe)])
    return exe


def test_cpp_example_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(_build_example(tmp_path)), "2", "4"], capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA device is required" in r.stderr
    assert "Number of degrees of freedom: 2312 (2187+125)" in r.stdout


@pytest.mark.gpu
def test_cpp_example_reproduces_the_reference_screen_output(tmp_path):
    g = golden("nsinker_gmg_gc")
    r = subprocess.run([str(_build_example(tmp_path)), "2", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    its = g["stokes_iterations_cheap_expensive"][0][0]
    assert f"Number of active cells: {g['active_cells'][0]} (on {g['levels'][0]} levels)" in r.stdout
    assert f"GMG coarse size A: {g['coarse_sizes_A_S'][0][0]}, coarse size S: {g['coarse_sizes_A_S'][0][1]}" in r.stdout
    assert f"Viscosity range: {g['viscosity_range'][0][0]:g} - {g['viscosity_range'][0][1]:g}" in r.stdout
    assert f"Stokes solver (GMRES): {its}+0 iterations." in r.stdout
    assert f"A block preconditioner: {g['A_preconditioner_applications'][0]}+0 iterations." in r.stdout
