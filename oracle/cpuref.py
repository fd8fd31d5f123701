"""ctypes front-end of oracle/csrc/smf_cpu_solver.cc, the C++/OpenMP/SIMD implementation of the reference's block-GMG
Stokes solve on the host cores (TEST INFRASTRUCTURE ONLY: the checker of the GPU parity tests at BASELINE cfg 1 / cfg 2
sizes and the `cpu_baseline` / `--impl reference` arm of bench.py).

`CppBlockGMG` takes the same inputs as oracle/cpu.py's CpuBackend (levels with cell->DoF maps, child tables, viscosity
tables, constrained velocity DoFs) and offers
  * solve(): the complete cheap-phase solve in C++ (smoother set-up, eigenvalue estimation, tolerance, GMRES), no Python
    in the iteration;
  * the backend interface of oracle/solver.py (apply_A, apply_M, dinv_*, prolongate_*, restrict_*, stokes, BT) so that
    the Python restatements of IDR(s) / FGMRES can drive the same C++ operators, and the C++ pieces can be compared one
    by one with the numpy/assembled oracle (tests/test_cpuref.py).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class _Result(ctypes.Structure):
    _fields_ = [("iterations", ctypes.c_int64), ("n_prec", ctypes.c_int64), ("residual", ctypes.c_double),
                ("tolerance", ctypes.c_double), ("initial_residual", ctypes.c_double), ("seconds", ctypes.c_double),
                ("setup_seconds", ctypes.c_double), ("seconds_vcycle_A", ctypes.c_double), ("seconds_vcycle_S", ctypes.c_double),
                ("seconds_stokes", ctypes.c_double), ("seconds_ortho", ctypes.c_double), ("seconds_loop", ctypes.c_double),
                ("converged", ctypes.c_int)]


def _has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return all(k in fl for k in ("avx512f", "avx512dq", "avx512vl", "avx512bw", "avx512cd"))
    except OSError:
        pass
    return False


def lib():
    """The widest build the host supports (x86-64-v4 = AVX-512, 8 cells per SIMD batch; else x86-64-v3 = AVX2, 4 cells)."""
    global _lib
    if _lib is None:
        names = (["libsmf_cpuref_w8.so"] if _has_avx512() and os.environ.get("SMFC_FORCE_W4") != "1" else []) + ["libsmf_cpuref_w4.so"]
        for name in names:
            path = os.path.join(_HERE, name)
            if os.path.exists(path):
                _lib = ctypes.CDLL(path)
                break
        if _lib is None:
            raise RuntimeError("oracle/libsmf_cpuref_w*.so missing: run `make -C oracle`")
        _lib.smfc_create.restype = ctypes.c_void_p
        _lib.smfc_width.restype = ctypes.c_int
        _lib.smfc_num_threads.restype = ctypes.c_int
        _lib.smfc_time_ortho.restype = ctypes.c_double
    return _lib


def set_num_threads(n=None):
    """One thread per core the process may run on (torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm)."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().smfc_set_num_threads(ctypes.c_int(int(n)))
    return int(lib().smfc_num_threads())


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CppBlockGMG:
    def __init__(self, levels, children, active_eta, level_eta, pressure_scaling, constrained_u, compressible=False):
        L = lib()
        self.n_levels = len(levels)
        fin = levels[-1]
        self.dim = dim = fin.dim
        self.n_u, self.n_p = int(fin.n_u), int(fin.n_p)
        self.h = ctypes.c_void_p(L.smfc_create(dim, self.n_levels))
        self.sizes = [(int(lv.n_u), int(lv.n_p)) for lv in levels]

        def set_level(idx, lv, eta, con):
            cu, cp = _i64(lv.cell_dofs_u), _i64(lv.cell_dofs_p)
            eta = _f64(eta).reshape(len(cu), -1)
            con = _i64(con if con is not None else [])
            L.smfc_set_level(self.h, ctypes.c_int(idx), ctypes.c_int64(len(cu)), ctypes.c_int64(lv.n_u), ctypes.c_int64(lv.n_p),
                             _p(cu), _p(cp), _p(_f64(lv.h)), _p(eta), ctypes.c_int(eta.shape[1]), ctypes.c_double(pressure_scaling),
                             ctypes.c_int(int(compressible)), _p(con), ctypes.c_int64(len(con)))
        for l, lv in enumerate(levels):
            set_level(l, lv, level_eta[l], constrained_u[l])
        set_level(-1, fin, active_eta, constrained_u[-1])
        for l in range(1, self.n_levels):
            ch = _i64(children[l])
            c, f = levels[l - 1], levels[l]
            con = _i64(constrained_u[l - 1] if constrained_u[l - 1] is not None else [])
            L.smfc_set_transfer(self.h, l, 0, ctypes.c_int64(len(ch)), _p(ch), _p(_i64(c.cell_dofs_u)), _p(_i64(f.cell_dofs_u)),
                                ctypes.c_int64(c.n_u), ctypes.c_int64(f.n_u), _p(con), ctypes.c_int64(len(con)))
            L.smfc_set_transfer(self.h, l, 1, ctypes.c_int64(len(ch)), _p(ch), _p(_i64(c.cell_dofs_p)), _p(_i64(f.cell_dofs_p)),
                                ctypes.c_int64(c.n_p), ctypes.c_int64(f.n_p), None, ctypes.c_int64(0))
        L.smfc_finalize(self.h)

    def close(self):
        if self.h:
            lib().smfc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the complete solve in C++ ---------------------------------------------------------------------------
    def solve(self, rhs, x0=None, tol_rel=1e-6, restart=100, max_its=200):
        x = np.zeros(self.n_u + self.n_p) if x0 is None else _f64(x0).copy()
        res = _Result()
        lib().smfc_solve(self.h, _p(_f64(rhs)), _p(x), ctypes.c_double(tol_rel), ctypes.c_int(restart), ctypes.c_int(max_its),
                         ctypes.byref(res))
        return dict(x=x, iterations=int(res.iterations), n_prec=int(res.n_prec), residual=res.residual, tolerance=res.tolerance,
                    initial_residual=res.initial_residual, converged=bool(res.converged), seconds=res.seconds,
                    setup_seconds=res.setup_seconds, split=dict(vcycle_A=res.seconds_vcycle_A, vcycle_S=res.seconds_vcycle_S,
                                                                 stokes_vmult=res.seconds_stokes, gmres_ortho=res.seconds_ortho,
                                                                 iteration_loop=res.seconds_loop))

    def time_ortho(self, k, reps=2):
        """Seconds of one CGS2 orthogonalisation against a basis of k block vectors (the solve's own routine)."""
        return float(lib().smfc_time_ortho(self.h, ctypes.c_int(int(k)), ctypes.c_int(int(reps))))

    # ---- backend interface of oracle/solver.py -----------------------------------------------------------------
    def _out(self, n):
        return np.empty(n)

    def apply_A(self, l, x):
        y = self._out(self.sizes[l][0]); lib().smfc_vmult_A(self.h, ctypes.c_int(l), _p(_f64(x)), _p(y)); return y

    def apply_M(self, l, x):
        y = self._out(self.sizes[l][1]); lib().smfc_vmult_mass(self.h, ctypes.c_int(l), _p(_f64(x)), _p(y)); return y

    def dinv_A(self, l):
        y = self._out(self.sizes[l][0]); lib().smfc_dinv(self.h, ctypes.c_int(l), 0, _p(y)); return y

    def dinv_M(self, l):
        y = self._out(self.sizes[l][1]); lib().smfc_dinv(self.h, ctypes.c_int(l), 1, _p(y)); return y

    def prolongate_u(self, l, x):
        y = self._out(self.sizes[l][0]); lib().smfc_transfer(self.h, ctypes.c_int(l), 0, 0, _p(_f64(x)), _p(y)); return y

    def prolongate_p(self, l, x):
        y = self._out(self.sizes[l][1]); lib().smfc_transfer(self.h, ctypes.c_int(l), 1, 0, _p(_f64(x)), _p(y)); return y

    def restrict_u(self, l, x):
        y = self._out(self.sizes[l - 1][0]); lib().smfc_transfer(self.h, ctypes.c_int(l), 0, 1, _p(_f64(x)), _p(y)); return y

    def restrict_p(self, l, x):
        y = self._out(self.sizes[l - 1][1]); lib().smfc_transfer(self.h, ctypes.c_int(l), 1, 1, _p(_f64(x)), _p(y)); return y

    def stokes(self, xu, xp):
        yu, yp = self._out(self.n_u), self._out(self.n_p)
        lib().smfc_vmult_stokes(self.h, _p(_f64(xu)), _p(_f64(xp)), _p(yu), _p(yp)); return yu, yp

    def BT(self, xp):
        y = self._out(self.n_u); lib().smfc_vmult_BT(self.h, _p(_f64(xp)), _p(y)); return y

    def A_active(self, xu):
        y = self._out(self.n_u); lib().smfc_vmult_A(self.h, ctypes.c_int(-1), _p(_f64(xu)), _p(y)); return y

    def M_active(self, xp):
        y = self._out(self.n_p); lib().smfc_vmult_mass(self.h, ctypes.c_int(-1), _p(_f64(xp)), _p(y)); return y

    # ---- the C++ V-cycles / preconditioner (after setup_smoothers) ----------------------------------------------
    def setup_smoothers(self):
        lib().smfc_setup_smoothers(self.h)

    def smoother_info(self, level, family):
        out = np.empty(4); lib().smfc_smoother_info(self.h, ctypes.c_int(level), ctypes.c_int(family), _p(out))
        return dict(lambda_min=out[0], lambda_max=out[1], cg_iterations=int(out[2]), theta=out[3])

    def vcycle(self, family, src):
        y = self._out(self.n_u if family == 0 else self.n_p)
        lib().smfc_vcycle(self.h, ctypes.c_int(family), _p(_f64(src)), _p(y)); return y

    def precondition(self, r):
        z = self._out(self.n_u + self.n_p); lib().smfc_precondition(self.h, _p(_f64(r)), _p(z)); return z


def from_hierarchy(hier, active_eta, level_eta, pressure_scaling, bc, compressible=False):
    """`hier`: oracle.mesh.Hierarchy or the product's synthetic BoxHierarchy (maps are data, not algorithm)."""
    cu = [lev.constrained_u(bc) for lev in hier.levels]
    children = [None] + [hier.children(l) for l in range(1, hier.R + 1)]
    return CppBlockGMG(hier.levels, children, active_eta, level_eta, pressure_scaling, cu, compressible)
