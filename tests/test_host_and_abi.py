"""CPU-only checks of the product's host logic and of the C-ABI library (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import coeff, mesh as omesh
from aspect_b200 import _lib, benchmarks, cell_data, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dim,R", [(2, 0), (2, 1), (2, 3), (3, 0), (3, 1), (3, 2), (3, 3)])
def test_index_maps_bit_exact(dim, R):
    """Vectorised synthetic driver vs the oracle's literal first-touch restatement of App. A.4: integer-exact."""
    ho, hp = omesh.Hierarchy(dim, R), synthetic.BoxHierarchy(dim, R)
    for lo, lp in zip(ho.levels, hp.levels):
        assert np.array_equal(lo.cells, lp.cells)
        assert np.array_equal(lo.cell_dofs_u, lp.cell_dofs_u) and np.array_equal(lo.cell_dofs_p, lp.cell_dofs_p)
        assert (lo.n_u, lo.n_p) == (lp.n_u, lp.n_p)
        for bc in ("zero", "tangential", "none"):
            assert np.array_equal(lo.constrained_u(bc), lp.constrained_u(bc))
    for l in range(1, R + 1):
        assert np.array_equal(ho.children(l), hp.children(l))


def test_dof_counts_of_the_configs():
    """SURVEY.md section 8 size table (N_u = dim (2n+1)^dim, N_p = (n+1)^dim) and the goldens' DoF counts."""
    for dim, R, nu, npp in [(2, 2, 162, 25), (2, 4, 2178, 289), (3, 2, 2187, 125), (3, 4, 107811, 4913)]:
        L = synthetic.BoxLevel(dim, R)
        assert (L.n_u, L.n_p) == (nu, npp)
    L0 = synthetic.BoxLevel(3, 0)
    assert (L0.n_u, L0.n_p) == (81, 8)


@pytest.mark.parametrize("averaging,op", [("none", "none"), ("harmonic average", "harmonic"), ("arithmetic average", "arithmetic"),
                                          ("geometric average", "geometric"), ("harmonic average only viscosity", "harmonic")])
def test_viscosity_tables_match_oracle(averaging, op):
    hp, ho = synthetic.BoxHierarchy(3, 2), omesh.Hierarchy(3, 2)
    eta_q, _ = benchmarks.nsinker_material(hp.active.quadrature_points(), 4)
    eta_o, _ = coeff.nsinker(ho.active.quadrature_points(), 4)
    assert np.array_equal(eta_q, eta_o)
    children = [None] + [hp.children(l) for l in range(1, 3)]
    a, lv, mn, mx = cell_data.fill_cell_data(3, children, eta_q, averaging)
    ao, lvo, mno, mxo = coeff.viscosity_tables(ho, eta_o, op)
    assert np.array_equal(a, ao) and (mn, mx) == (mno, mxo)
    for x, y in zip(lv, lvo):
        assert np.array_equal(x, y)
    if op == "harmonic":
        assert f"{mn:.6g} - {mx:.6g}" == "0.01 - 0.198921"      # golden "Viscosity range" of tests/nsinker_gmg_gc


def test_rhs_and_pressure_scaling_match_oracle():
    from oracle import assembled
    hp, ho = synthetic.BoxHierarchy(2, 3), omesh.Hierarchy(2, 3)
    eta_q, rhs, ps = benchmarks.solcx_problem(hp)
    _, rho_o = coeff.solcx(ho.active.quadrature_points())
    f = assembled.rhs_gravity(ho.active, rho_o, np.array([0, -1.0]))
    f[ho.active.constrained_u("tangential")] = 0
    assert np.allclose(rhs[:hp.active.n_u], f, rtol=0, atol=1e-17)
    assert abs(ps - coeff.pressure_scaling_box(ho, lambda x: coeff.solcx(x)[0])) < 1e-9 and abs(ps - 1e5) < 1e-6


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "smf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smf_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()          # raises if the CUDA library has not been built
    syms = _header_symbols()
    assert len(syms) >= 30 and set(syms) == set(_lib.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/smf_b200.h but not exported"


def test_no_cpu_fallback_without_device():
    """Without a CUDA device smf_create must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    cfg = _lib.Config(3, 2, 2, 0)
    ctx = ctypes.c_void_p()
    assert lib.smf_create(ctypes.byref(cfg), ctypes.byref(ctx)) == _lib.SMF_ERR_CUDA and not ctx.value
    from aspect_b200 import StokesMatrixFreeHandler, SmfError
    with pytest.raises(SmfError):
        StokesMatrixFreeHandler(3).setup_dofs(synthetic.BoxHierarchy(3, 1))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "aspect_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("oracle-free", ""), f


def test_ctypes_structs_match_the_header(tmp_path):
    """sizeof / offsetof of every struct of include/smf_b200.h, measured by a C program compiled with gcc, against the ctypes
    mirrors in aspect_b200/_lib.py (a silent mismatch would corrupt parameters on the GPU box)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    pairs = {"smf_config": _lib.Config, "smf_level_mesh": _lib.LevelMesh, "smf_exchange": _lib.Exchange, "smf_cell_data": _lib.CellData,
             "smf_newton_data": _lib.NewtonData, "smf_solve_params": _lib.SolveParams, "smf_solve_result": _lib.SolveResult}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "smf_b200.h"', 'int main(void){']
    header = open(os.path.join(ROOT, "include", "smf_b200.h")).read()
    for cname, cls in pairs.items():
        body = re.search(r"typedef struct \{([^}]*)\}\s*" + cname + r"\s*;", header).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = [n for decl in body.split(";") if decl.strip() for n in re.findall(r"[\*\s](\w+)\s*(?:,|$)", decl.strip())]
        assert names == [f[0] for f in cls._fields_], f"{cname}: field names / order differ: {names} vs {[f[0] for f in cls._fields_]}"
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for n in names:
            lines.append(f'printf("{cname}.{n} %zu\\n", offsetof({cname}, {n}));')
    lines.append("return 0;}")
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for n, _ in cls._fields_:
            assert int(got[f"{cname}.{n}"]) == getattr(cls, n).offset, f"{cname}.{n}"


def test_stokes_gmg_type_parameter():
    """`Stokes GMG type` (parameters.cc:448-453): both values are accepted on uniformly refined meshes (same hierarchy, name()
    follows the reference's "GMG" / "GMG-GC"); global coarsening on an adaptive mesh and unknown values are refused before any
    device work."""
    from aspect_b200 import SmfError, StokesMatrixFreeHandler, _lib
    from aspect_b200 import adaptive as padapt
    h = StokesMatrixFreeHandler(3)
    assert h.name() == "GMG"
    h.parameters["Stokes GMG type"] = "global coarsening"
    assert h.name() == "GMG-GC"
    hier = padapt.AdaptiveBoxHierarchy.refined(3, 1, lambda x, l: x[2] < 0.5, 1)
    with pytest.raises(SmfError) as e:
        h.setup_dofs(hier, "zero")
    assert e.value.args[0] == _lib.SMF_ERR_UNSUPPORTED or "global coarsening" in str(e.value)
    h.parameters["Stokes GMG type"] = "algebraic"
    with pytest.raises(SmfError):
        h.setup_dofs(hier, "zero")
