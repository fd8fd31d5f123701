"""ctypes front-end of the C matrix-free restatement oracle/csrc/smf_cpu.c (TEST INFRASTRUCTURE ONLY).

`CpuBackend` exposes the backend interface of oracle/solver.py on top of the C cell kernels, so the
whole block-GMG solve can run matrix-free on the host cores: this is the checker the CUDA path is
compared with at sizes where assembled matrices are too large, and the `cpu_baseline` /
`--impl reference` arm of bench.py.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class _Level(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_int), ("n_cells", ctypes.c_int64),
                ("cell_u", ctypes.c_void_p), ("cell_p", ctypes.c_void_p),
                ("h", ctypes.c_void_p), ("eta", ctypes.c_void_p), ("n_eta", ctypes.c_int),
                ("ps", ctypes.c_double), ("compressible", ctypes.c_int),
                ("mask_u", ctypes.c_void_p),
                ("order", ctypes.c_void_p), ("color_ptr", ctypes.c_void_p), ("n_colors", ctypes.c_int)]


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libsmf_cpu.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libsmf_cpu.so missing: run `make -C oracle`")
        _lib = ctypes.CDLL(path)
        _lib.smf_cpu_num_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CpuLevel:
    """Arrays of one level in the layout smf_cpu.c expects (keeps them alive)."""

    def __init__(self, dim, cells_ijk, cell_u, cell_p, h, eta, ps, constrained_u, n_u, n_p, compressible=False):
        self.dim, self.n_u, self.n_p = dim, int(n_u), int(n_p)
        self.cell_u = np.ascontiguousarray(cell_u, dtype=np.int64)
        self.cell_p = np.ascontiguousarray(cell_p, dtype=np.int64)
        self.h = np.ascontiguousarray(h, dtype=np.float64)
        self.eta = np.ascontiguousarray(eta, dtype=np.float64).reshape(len(self.cell_u), -1)
        self.mask = np.zeros(self.n_u, dtype=np.uint8)
        if constrained_u is not None and len(constrained_u):
            self.mask[np.asarray(constrained_u)] = 1
        color = np.zeros(len(self.cell_u), dtype=np.int64)
        for d in range(dim):
            color += (np.asarray(cells_ijk)[:, d] & 1) << d
        self.order = np.ascontiguousarray(np.argsort(color, kind="stable"), dtype=np.int64)
        self.color_ptr = np.ascontiguousarray(np.concatenate([[0], np.cumsum(np.bincount(color, minlength=1 << dim))]), dtype=np.int64)
        self.c = _Level(dim, len(self.cell_u), _p(self.cell_u), _p(self.cell_p), _p(self.h), _p(self.eta),
                        self.eta.shape[1], float(ps), int(compressible), _p(self.mask), _p(self.order),
                        _p(self.color_ptr), 1 << dim)

    def vmult_A(self, x):
        y = np.empty(self.n_u)
        lib().smf_cpu_vmult_A(ctypes.byref(self.c), ctypes.c_int64(self.n_u), _p(np.ascontiguousarray(x)), _p(y))
        return y

    def vmult_mass(self, x):
        y = np.empty(self.n_p)
        lib().smf_cpu_vmult_mass(ctypes.byref(self.c), ctypes.c_int64(self.n_p), _p(np.ascontiguousarray(x)), _p(y))
        return y

    def vmult_stokes(self, xu, xp):
        yu, yp = np.empty(self.n_u), np.empty(self.n_p)
        lib().smf_cpu_vmult_stokes(ctypes.byref(self.c), ctypes.c_int64(self.n_u), ctypes.c_int64(self.n_p),
                                   _p(np.ascontiguousarray(xu)), _p(np.ascontiguousarray(xp)), _p(yu), _p(yp))
        return yu, yp

    def vmult_BT(self, xp):
        yu = np.empty(self.n_u)
        lib().smf_cpu_vmult_BT(ctypes.byref(self.c), ctypes.c_int64(self.n_u), _p(np.ascontiguousarray(xp)), _p(yu))
        return yu

    def diag_A(self):
        d = np.empty(self.n_u)
        lib().smf_cpu_diag_A(ctypes.byref(self.c), ctypes.c_int64(self.n_u), _p(d))
        return d

    def diag_mass(self):
        d = np.empty(self.n_p)
        lib().smf_cpu_diag_mass(ctypes.byref(self.c), ctypes.c_int64(self.n_p), _p(d))
        return d


class CpuTransfer:
    def __init__(self, dim, npd, ncomp, children, cell_c, cell_f, n_c, n_f, constrained_c):
        self.dim, self.npd, self.ncomp = dim, npd, ncomp
        self.children = np.ascontiguousarray(children, dtype=np.int64)
        self.cell_c = np.ascontiguousarray(cell_c, dtype=np.int64)
        self.cell_f = np.ascontiguousarray(cell_f, dtype=np.int64)
        self.n_c, self.n_f = int(n_c), int(n_f)
        self.mask = np.zeros(self.n_c, dtype=np.uint8)
        if constrained_c is not None and len(constrained_c):
            self.mask[np.asarray(constrained_c)] = 1
        # valence = number of parent patches touching a fine DoF
        nn = npd ** dim
        npatch = (2 * npd - 1) ** dim
        val = np.zeros(self.n_f)
        # count via children maps: a fine dof is in parent's patch if in any child; count unique per parent
        for par_chunk in np.array_split(np.arange(len(self.children)), max(1, len(self.children) // 65536)):
            dofs = self.cell_f[self.children[par_chunk]].reshape(len(par_chunk), -1)
            dofs = np.sort(dofs, axis=1)
            first = np.ones_like(dofs, dtype=bool)
            first[:, 1:] = dofs[:, 1:] != dofs[:, :-1]
            np.add.at(val, dofs[first], 1.0)
        self.weights = np.ascontiguousarray(1.0 / np.maximum(val, 1.0))

    def prolongate(self, xc):
        y = np.empty(self.n_f)
        lib().smf_cpu_prolongate(self.dim, self.npd, self.ncomp, ctypes.c_int64(len(self.children)), _p(self.children),
                                 _p(self.cell_c), _p(self.cell_f), _p(self.weights), _p(self.mask),
                                 ctypes.c_int64(self.n_f), _p(np.ascontiguousarray(xc)), _p(y))
        return y

    def restrict(self, xf):
        y = np.zeros(self.n_c)
        lib().smf_cpu_restrict_add(self.dim, self.npd, self.ncomp, ctypes.c_int64(len(self.children)), _p(self.children),
                                   _p(self.cell_c), _p(self.cell_f), _p(self.weights), _p(self.mask),
                                   _p(np.ascontiguousarray(xf)), _p(y))
        return y


class CpuBackend:
    """Matrix-free backend for oracle.solver.BlockGMGStokes over the C kernels.

    `levels`: list (coarse -> fine) of objects with attributes dim, cells (ijk), cell_dofs_u, cell_dofs_p, h,
    n_u, n_p (oracle.mesh.Level or the product's synthetic levels -- maps are data, not algorithm);
    `children[l]`: (n_cells(l-1), 2^dim) child table."""

    def __init__(self, levels, children, active_eta, level_eta, pressure_scaling, constrained_u, compressible=False):
        self.n_levels = len(levels)
        fin = levels[-1]
        dim = fin.dim
        self.active = CpuLevel(dim, fin.cells, fin.cell_dofs_u, fin.cell_dofs_p, fin.h, active_eta, pressure_scaling,
                               constrained_u[-1], fin.n_u, fin.n_p, compressible)
        self.lv = [CpuLevel(dim, L.cells, L.cell_dofs_u, L.cell_dofs_p, L.h, level_eta[l], pressure_scaling,
                            constrained_u[l], L.n_u, L.n_p, compressible) for l, L in enumerate(levels)]
        self.tu = [None] + [CpuTransfer(dim, 3, dim, children[l], levels[l - 1].cell_dofs_u, levels[l].cell_dofs_u,
                                        levels[l - 1].n_u, levels[l].n_u, constrained_u[l - 1]) for l in range(1, self.n_levels)]
        self.tp = [None] + [CpuTransfer(dim, 2, 1, children[l], levels[l - 1].cell_dofs_p, levels[l].cell_dofs_p,
                                        levels[l - 1].n_p, levels[l].n_p, None) for l in range(1, self.n_levels)]
        self.n_u, self.n_p = fin.n_u, fin.n_p
        self._dA = [None] * self.n_levels
        self._dM = [None] * self.n_levels

    def apply_A(self, l, x): return self.lv[l].vmult_A(x)
    def apply_M(self, l, x): return self.lv[l].vmult_mass(x)

    def dinv_A(self, l):
        if self._dA[l] is None:
            self._dA[l] = 1.0 / self.lv[l].diag_A()
        return self._dA[l]

    def dinv_M(self, l):
        if self._dM[l] is None:
            self._dM[l] = 1.0 / self.lv[l].diag_mass()
        return self._dM[l]

    def prolongate_u(self, l, x): return self.tu[l].prolongate(x)
    def prolongate_p(self, l, x): return self.tp[l].prolongate(x)
    def restrict_u(self, l, x): return self.tu[l].restrict(x)
    def restrict_p(self, l, x): return self.tp[l].restrict(x)
    def stokes(self, xu, xp): return self.active.vmult_stokes(xu, xp)
    def BT(self, xp): return self.active.vmult_BT(xp)
    def A_active(self, xu): return self.active.vmult_A(xu)
    def M_active(self, xp): return self.active.vmult_mass(xp)


def backend_from_hierarchy(hier, active_eta, level_eta, pressure_scaling, bc, compressible=False):
    cu = [lev.constrained_u(bc) for lev in hier.levels]
    children = [None] + [hier.children(l) for l in range(1, hier.R + 1)]
    return CpuBackend(hier.levels, children, active_eta, level_eta, pressure_scaling, cu, compressible)
