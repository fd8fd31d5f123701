"""Host-side mirror of ASPECT's StokesMatrixFreeHandler over the C ABI.

Same life cycle and names as the reference
(/root/reference/include/aspect/simulator/solver/stokes_matrix_free.h:46-98 and
 stokes_matrix_free_local_smoothing.h:211-236):
    setup_dofs() -> assemble() [evaluate_material_model] -> build_preconditioner() -> solve()
and the same operator objects with `vmult(dst, src)`:
    stokes_matrix, A_block_matrix, BT_block, Schur_complement_block_matrix,
    mg_matrices_A_block[l], mg_matrices_Schur_complement[l], prec_A, prec_Schur.
All compute goes through aspect_b200/libsmf_b200.so; this file only marshals arrays.
"""
import ctypes
import os

import numpy as np

from . import _lib, cell_data
from ._lib import SMF_ACTIVE, SmfError, ptr


class DeviceVector:
    """LinearAlgebra::distributed::Vector<double> analogue: n doubles in HBM, owned by the handler's context."""

    def __init__(self, handler, n):
        self.h, self.n = handler, int(n)
        p = ctypes.c_void_p()
        handler._check(handler.lib.smf_vec_alloc(handler.ctx, ctypes.c_int64(self.n), ctypes.byref(p)))
        self.ptr = p.value

    def from_host(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.n
        self.h._check(self.h.lib.smf_vec_h2d(self.h.ctx, ptr(self.ptr), ptr(a), ctypes.c_int64(self.n)))
        self.h.synchronize()
        return self

    def zero(self):
        self.h._check(self.h.lib.smf_vec_set(self.h.ctx, ptr(self.ptr), ctypes.c_double(0.0), ctypes.c_int64(self.n)))
        return self

    def to_host(self):
        out = np.empty(self.n)
        self.h._check(self.h.lib.smf_vec_d2h(self.h.ctx, ptr(out), ptr(self.ptr), ctypes.c_int64(self.n)))
        return out

    def free(self):
        if self.ptr:
            self.h.lib.smf_vec_free(self.h.ctx, ptr(self.ptr))
            self.ptr = None


class _Op:
    def __init__(self, handler, fn):
        self.h, self.fn = handler, fn

    def vmult(self, dst, src):
        self.fn(dst, src)


class StokesMatrixFreeHandler:
    """`Stokes solver type = block GMG`, `Stokes GMG type = local smoothing`, Q2-Q1."""

    def __init__(self, dim, device=0, stokes_velocity_degree=2):
        self.lib = _lib.load()
        self.dim = dim
        self.degree = stokes_velocity_degree
        self.device = device
        self.ctx = None
        self.hier = None
        self.minimum_viscosity = self.maximum_viscosity = None
        self.allow_structured_bricks = True      # implementation switches (smf_set_option); tests exercise both paths
        self.use_cuda_graphs = True
        self.device_cell_data = True             # assemble(): averaging / projection / level tables on the device (False: host numpy)
        self.use_warp_brick_kernel = os.environ.get("SMF_WARP_BRICK", "0") == "1"    # structured levels: one warp per brick, 3 components per thread
        self.use_coarse_tail = True              # Schur V-cycle: coarse levels in one single-block kernel
        self.use_lattice_velocity = os.environ.get("SMF_LATTICE_VELOCITY", "0") == "1"   # experimental: TMA-staged A-block kernel on lattice-ordered vectors
        self.replicate_coarse_A = os.environ.get("SMF_REPLICATE_A", "1") == "1"   # >1 rank: coarse levels of prec_A on the global mesh, no exchanges
        self.use_lattice_schur = os.environ.get("SMF_LATTICE", "1") == "1"    # Schur V-cycle on box levels in lattice order (node-centred fused kernels)
        self.use_pipelined_kernel = True         # structured levels: persistent cp.async-pipelined velocity kernel
        self.max_grid_blocks = 0                 # > 0 caps the persistent grids (tests)
        self.p2p_exchange = True                 # >1 rank: halo exchange as one kernel over NVLink peer memory (False: NCCL send/recv)
        self.gram_reorthogonalization = True     # GMRES: CGS2 in one sweep over the basis (False: two sweeps)
        self.curved_sum_factorization = True     # curved cells: sum-factorised kernel (False: the general per-cell restatement kernel)
        self.curved_tma_geometry = os.environ.get("SMF_CURVED_TMA", "0") == "1"    # geometry stream staged by cp.async.bulk + mbarrier
        # Chebyshev eigenvalue start vector from partition-independent global DoF numbers (always on for >1 rank)
        self.global_start_vector = False
        # "Solver parameters/Stokes solver parameters" (names verbatim from source/simulator/parameters.cc:439-584)
        self.parameters = {
            "Linear solver tolerance": 1e-7,
            "Number of cheap Stokes solver steps": 1000,
            "Maximum number of expensive Stokes solver steps": 1000,
            "GMRES solver restart length": 100,
            "Stokes GMG type": "local smoothing",                 # or "global coarsening": uniformly refined meshes only (setup_dofs)
            "Krylov method for cheap solver steps": "GMRES",      # or "IDR(s)"
            "IDR(s) parameter": 2,
            "Linear solver A block tolerance": 1e-2,
            "Linear solver S block tolerance": 1e-6,
            # global parameters the solve's post-processing reads (parameters.cc: "Pressure normalization", "Surface pressure")
            "Pressure normalization": "no",                        # "surface" | "volume" | "no" (reference default: surface)
            "Surface pressure": 0.0,
        }

    def name(self):
        """`StokesSolver::Interface::name()`: "GMG" (local smoothing) or "GMG-GC" (stokes_matrix_free_global_coarsening.cc:142-145)."""
        return "GMG-GC" if self.parameters["Stokes GMG type"] == "global coarsening" else "GMG"

    def _check_gmg_type(self, hierarchy):
        """`Stokes GMG type` (parameters.cc:448-453).  On a UNIFORMLY refined mesh the coarsening sequence of the global-coarsening
        handler (create_geometric_coarsening_sequence, global_coarsening.cc:1466) is the sequence of refinement levels, its level
        operators, smoothers and coarse solver are the local-smoothing ones (:796-913) and its two-level transfers are the same
        polynomial embeddings, so this hierarchy serves both types -- the reference's own global-coarsening goldens nsinker_gmg_gc
        (21), nsinker_gmg_gc_no_avg, and gmg_test_project_q1-style runs are what tests/ reproduces on it.  On an ADAPTIVE mesh
        global coarsening needs levels that are themselves meshes with hanging nodes and transfers between them: not built."""
        kind = self.parameters["Stokes GMG type"]
        if kind not in ("local smoothing", "global coarsening"):
            raise SmfError(_lib.SMF_ERR_INVALID, "Stokes GMG type: local smoothing | global coarsening")
        if kind == "global coarsening" and getattr(hierarchy, "adaptive", False):
            raise SmfError(_lib.SMF_ERR_UNSUPPORTED,
                           "Stokes GMG type = global coarsening on an adaptively refined mesh (levels with hanging nodes) is not built; "
                           "use local smoothing")

    # ------------------------------------------------------------------
    def _check(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            msg = self.lib.smf_last_error(self.ctx).decode() if self.ctx else "context creation failed (no CUDA device?)"
            raise SmfError(rc, msg)
        return rc

    def synchronize(self):
        self._check(self.lib.smf_synchronize(self.ctx))

    def set_stream(self, cuda_stream):
        self._check(self.lib.smf_set_stream(self.ctx, ctypes.c_void_p(int(cuda_stream))))

    def close(self):
        if self.ctx:
            self.lib.smf_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def setup_dofs(self, hierarchy, boundary="zero"):
        """hierarchy: levels[l] with n_cells, n_u, n_p, cell_dofs_u, cell_dofs_p, h, constrained_u(bc); children(l).
        boundary: 'zero' (Zero velocity boundary indicators = all), 'tangential' (Tangential velocity ... = all)."""
        self._check_gmg_type(hierarchy)
        self.close()
        self.hier = hierarchy
        self.boundary = boundary
        self._pnorm_kind = None            # the weights live in the context that was just destroyed
        n_levels = len(hierarchy.levels)
        cfg = _lib.Config(self.dim, self.degree, n_levels, self.device)
        ctx = ctypes.c_void_p()
        rc = self.lib.smf_create(ctypes.byref(cfg), ctypes.byref(ctx))
        if rc != 0:
            raise SmfError(rc, "smf_create failed: a CUDA device is required (no CPU fallback)")
        self.ctx = ctx
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_STRUCTURED_BRICKS, int(self.allow_structured_bricks)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_CUDA_GRAPHS, int(self.use_cuda_graphs)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_ORTHO_SINGLE_PASS, int(self.gram_reorthogonalization)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_PIPELINED_KERNEL, int(self.use_pipelined_kernel)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_WARP_BRICK_KERNEL, int(self.use_warp_brick_kernel)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_COARSE_TAIL, int(self.use_coarse_tail)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_LATTICE_SCHUR, int(self.use_lattice_schur)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_LATTICE_VELOCITY, int(self.use_lattice_velocity)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_REPLICATE_COARSE_A, int(self.replicate_coarse_A)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_MAX_GRID_BLOCKS, int(self.max_grid_blocks)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_P2P_EXCHANGE, int(self.p2p_exchange)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_CURVED_SUMFAC, int(self.curved_sum_factorization)))
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_CURVED_TMA, int(self.curved_tma_geometry)))
        self._keep = []
        self.rank, self.world = getattr(hierarchy, "rank", 0), getattr(hierarchy, "world", 1)
        self._comm_init()
        for l in list(range(n_levels)) + [SMF_ACTIVE]:
            L = hierarchy.levels[l] if l >= 0 else hierarchy.active
            cu = np.ascontiguousarray(L.cell_dofs_u, dtype=np.uint32)
            cp = np.ascontiguousarray(L.cell_dofs_p, dtype=np.uint32)
            cons = np.ascontiguousarray(L.constrained_u(boundary), dtype=np.uint32)
            h = np.ascontiguousarray(L.h, dtype=np.float64)
            if getattr(L, "jinv", None) is not None:
                # curved cells: the mapping data MatrixFree holds per quadrature point (geometry = 1)
                ji = np.ascontiguousarray(L.jinv, dtype=np.float64)
                jw = np.ascontiguousarray(L.jxw, dtype=np.float64)
                assert ji.shape == (L.n_cells, 3 ** self.dim, self.dim, self.dim) and jw.shape == (L.n_cells, 3 ** self.dim)
                m = _lib.LevelMesh(L.n_cells, L.n_u, L.n_p, ptr(cu), ptr(cp), len(cons), ptr(cons), 0, ptr(None), 1, ptr(h), ptr(ji), ptr(jw))
            else:
                m = _lib.LevelMesh(L.n_cells, L.n_u, L.n_p, ptr(cu), ptr(cp), len(cons), ptr(cons), 0, ptr(None), 0, ptr(h), ptr(None), ptr(None))
            self._check(self.lib.smf_set_level(self.ctx, l, ctypes.byref(m)))
        if self.world > 1 or self.global_start_vector:
            for l in list(range(n_levels)) + [SMF_ACTIVE]:
                L = hierarchy.levels[l] if l >= 0 else hierarchy.active
                for family in (0, 1):
                    self._set_exchange(l, family, L)
        for l in range(1, n_levels):
            ch = np.ascontiguousarray(hierarchy.children(l), dtype=np.int32)
            self._check(self.lib.smf_set_transfer(self.ctx, l, ptr(ch)))
        self._check(self.lib.smf_set_copy_indices(self.ctx, ptr(None), ptr(None)))
        self.adaptive = bool(getattr(hierarchy, "adaptive", False))
        self._lines = {}
        if self.adaptive:
            # hanging-node lines of the active mesh, refinement edges and copy indices of the levels
            u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
            for family, name in ((0, "u"), (1, "p")):
                dof, start, master, weight = hierarchy.active.constraint_lines(name, boundary)
                self._lines[name] = (dof.astype(np.int64), start.astype(np.int64), master.astype(np.int64), weight)
                self._check(self.lib.smf_set_constraints(self.ctx, SMF_ACTIVE, family, ctypes.c_int64(len(dof)), ptr(u32(dof)), ptr(u32(start)),
                                                         ptr(u32(master)), ptr(np.ascontiguousarray(weight, dtype=np.float64))))
            for l in range(n_levels):
                L = hierarchy.levels[l]
                eu, ep = u32(L.edge_u), u32(L.edge_p)
                self._check(self.lib.smf_set_refinement_edges(self.ctx, l, ctypes.c_int64(len(eu)), ptr(eu), ctypes.c_int64(len(ep)), ptr(ep)))
                au, lu, ap, lp = hierarchy.copy_indices(l)
                self._check(self.lib.smf_set_level_copy_indices(self.ctx, l, 0, ctypes.c_int64(len(au)), ptr(u32(au)), ptr(u32(lu))))
                self._check(self.lib.smf_set_level_copy_indices(self.ctx, l, 1, ctypes.c_int64(len(ap)), ptr(u32(ap)), ptr(u32(lp))))
        # no-normal-flux lines on a curved boundary (`Tangential velocity boundary indicators` on the shell): constraint lines on
        # the active level AND on every multigrid level (user constraints of MGConstrainedDoFs, local_smoothing.cc:1719-1789)
        self._level_lines = {}
        if not self.adaptive and hasattr(hierarchy.active, "constraint_lines"):
            u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
            for l in list(range(n_levels)) + [SMF_ACTIVE]:
                L = hierarchy.levels[l] if l >= 0 else hierarchy.active
                dof, start, master, weight = L.constraint_lines("u", boundary)
                if len(dof) == 0:
                    continue
                self._check(self.lib.smf_set_constraints(self.ctx, l, 0, ctypes.c_int64(len(dof)), ptr(u32(dof)), ptr(u32(start)),
                                                         ptr(u32(master)), ptr(np.ascontiguousarray(weight, dtype=np.float64))))
                if l >= 0:
                    self._level_lines[l] = (dof.astype(np.int64), start.astype(np.int64), master.astype(np.int64), weight)
                else:
                    self._lines["u"] = (dof.astype(np.int64), start.astype(np.int64), master.astype(np.int64), weight)
        self.n_u, self.n_p = hierarchy.active.n_u, hierarchy.active.n_p
        self.n_levels = n_levels
        # operator objects (duck-typed vmult like the reference's)
        lib, c = self.lib, self._check
        self.stokes_matrix = _Op(self, lambda d, s: c(lib.smf_vmult_stokes(self.ctx, ptr(d[0].ptr), ptr(d[1].ptr), ptr(s[0].ptr), ptr(s[1].ptr))))
        self.A_block_matrix = _Op(self, lambda d, s: c(lib.smf_vmult_A(self.ctx, SMF_ACTIVE, ptr(d.ptr), ptr(s.ptr))))
        self.BT_block = _Op(self, lambda d, s: c(lib.smf_vmult_BT(self.ctx, ptr(d[0].ptr), ptr(s[1].ptr))))
        self.Schur_complement_block_matrix = _Op(self, lambda d, s: c(lib.smf_vmult_mass(self.ctx, SMF_ACTIVE, ptr(d.ptr), ptr(s.ptr))))
        self.mg_matrices_A_block = [_Op(self, lambda d, s, l=l: c(lib.smf_vmult_A(self.ctx, l, ptr(d.ptr), ptr(s.ptr)))) for l in range(n_levels)]
        self.mg_matrices_Schur_complement = [_Op(self, lambda d, s, l=l: c(lib.smf_vmult_mass(self.ctx, l, ptr(d.ptr), ptr(s.ptr)))) for l in range(n_levels)]
        self.prec_A = _Op(self, lambda d, s: c(lib.smf_vcycle_A(self.ctx, ptr(d.ptr), ptr(s.ptr))))
        self.prec_Schur = _Op(self, lambda d, s: c(lib.smf_vcycle_S(self.ctx, ptr(d.ptr), ptr(s.ptr))))
        self.preconditioner_cheap = _Op(self, lambda d, s: c(lib.smf_precondition(self.ctx, ptr(d[0].ptr), ptr(d[1].ptr), ptr(s[0].ptr), ptr(s[1].ptr))))

    def _comm_init(self):
        """One context per rank: rank 0 creates the NCCL id, torch.distributed carries it to the others."""
        idbuf = ctypes.create_string_buffer(_lib.SMF_UNIQUE_ID_BYTES)
        if self.world > 1:
            import torch.distributed as dist
            if self.rank == 0:
                self._check(self.lib.smf_comm_get_unique_id(idbuf))
            box = [idbuf.raw]
            dist.broadcast_object_list(box, src=0)
            idbuf = ctypes.create_string_buffer(box[0], _lib.SMF_UNIQUE_ID_BYTES)
        self._check(self.lib.smf_comm_init(self.ctx, self.rank, self.world, idbuf))

    def _set_exchange(self, level, family, L):
        fam = "u" if family == 0 else "p"
        ncomp = self.dim if family == 0 else 1
        lists = L.exchange_lists(fam) if self.world > 1 else {}
        ranks = sorted(lists)
        arrs = [np.ascontiguousarray(lists[q].astype(np.uint32) * ncomp) for q in ranks]     # first DoF of every shared node
        gid = L.global_node_ids(fam)
        if ncomp > 1:
            gid = (gid[:, None] * ncomp + np.arange(ncomp)[None, :]).ravel()
        gid = np.ascontiguousarray(gid, dtype=np.int64)
        rk = np.ascontiguousarray(ranks, dtype=np.int32)
        cnt = np.ascontiguousarray([len(a) for a in arrs], dtype=np.int64)
        ptrs = (ctypes.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        x = _lib.Exchange(len(ranks), ptr(rk), ptr(cnt), ctypes.cast(ptrs, ctypes.c_void_p), ptr(gid))
        self._check(self.lib.smf_set_exchange(self.ctx, level, family, ctypes.byref(x)))

    def exchange_add(self, vec, family, level=SMF_ACTIVE):
        """compress(add) + update_ghost_values of a device vector (e.g. a right-hand side assembled per rank)."""
        self._check(self.lib.smf_exchange_add(self.ctx, level, family, ptr(vec.ptr)))

    def initialize_dof_vector(self, block=None, level=SMF_ACTIVE):
        """block None -> (velocity, pressure) pair for the active level; 0 / 1 -> a single vector of that level."""
        L = self.hier.active if level == SMF_ACTIVE else self.hier.levels[level]
        if block is None:
            return (DeviceVector(self, L.n_u), DeviceVector(self, L.n_p))
        return DeviceVector(self, L.n_u if block == 0 else L.n_p)

    # ------------------------------------------------------------------
    def assemble(self, viscosity_q, material_averaging="none", pressure_scaling=1.0, is_compressible=False):
        """evaluate_material_model(): viscosity_q = plugin output at the velocity quadrature points of the
        active cells, (n_cells, 3^dim)."""
        self.pressure_scaling = float(pressure_scaling)
        if getattr(self, "adaptive", False):
            # adaptive mesh: cell-wise averaging as always; the level tables follow interpolate_to_mg on the partial levels
            # (a level cell carries its own value where it is active, else the mean of its children; local_smoothing.cc:220-226)
            q = cell_data.average_viscosity(np.asarray(viscosity_q, dtype=np.float64), material_averaging)
            self.minimum_viscosity, self.maximum_viscosity = float(q.min()), float(q.max())
            w = cell_data.quadrature_weights(self.dim)
            dg0 = (q * w[None, :]).sum(axis=1) / w.sum()
            active = q if material_averaging == "none" else dg0[:, None]
            levels = [np.ascontiguousarray(v[:, None]) for v in self.hier.level_viscosity(dg0)]
            self.set_cell_data(SMF_ACTIVE, np.ascontiguousarray(active), pressure_scaling, is_compressible)
            for l in range(self.n_levels):
                self.set_cell_data(l, levels[l], pressure_scaling, is_compressible)
            self._tables = (active, levels)
            return
        if self.device_cell_data:
            # averaging, DGQ0 projection, min / max and the level tables on the device (smf_fill_cell_data)
            op = material_averaging.replace(" only viscosity", "")
            codes = {"none": 0, "arithmetic average": 1, "harmonic average": 2, "geometric average": 3, "project to Q1": 4}
            if op not in codes:
                raise SmfError(_lib.SMF_ERR_UNSUPPORTED, f"unsupported Material averaging for the GMG path: {material_averaging}")
            q = np.ascontiguousarray(viscosity_q, dtype=np.float64)
            assert q.shape == (self.hier.active.n_cells, 3 ** self.dim)
            mm = np.zeros(2)
            self._check(self.lib.smf_fill_cell_data(self.ctx, ptr(q), codes[op], ctypes.c_double(pressure_scaling), int(is_compressible), ptr(mm)))
            self.minimum_viscosity, self.maximum_viscosity = float(mm[0]), float(mm[1])
            self._tables = None
            return
        children = [None] + [self.hier.children(l) for l in range(1, self.n_levels)]
        active, levels, self.minimum_viscosity, self.maximum_viscosity = cell_data.fill_cell_data(
            self.dim, children, viscosity_q, material_averaging, getattr(self.hier.active, "jxw", None))
        self.set_cell_data(SMF_ACTIVE, active, pressure_scaling, is_compressible)
        for l in range(self.n_levels):
            self.set_cell_data(l, levels[l], pressure_scaling, is_compressible)
        self._tables = (active, levels)

    def set_free_surface_faces(self, cells, face_numbers, stabilization_term_table):
        """`apply_stabilization_free_surface_faces` + `free_surface_stabilization_term_table` of the active OperatorCellData
        (local_smoothing.cc:528-612): boundary faces (active cell, local face number) and g = pressure perturbation * g_hat at the
        3^(dim-1) face Gauss points, (n_faces, 3^(dim-1), dim).  Pass empty arrays to clear."""
        c = np.ascontiguousarray(cells, dtype=np.uint32)
        f = np.ascontiguousarray(face_numbers, dtype=np.int32)
        t = np.ascontiguousarray(stabilization_term_table, dtype=np.float64)
        assert len(c) == len(f) and (len(c) == 0 or t.shape == (len(c), 3 ** (self.dim - 1), self.dim))
        self._check(self.lib.smf_set_free_surface_faces(self.ctx, ctypes.c_int64(len(c)), ptr(c), ptr(f), ptr(t)))

    def force_nonsymmetric_A_block_solver(self, on=True):
        """`Force nonsymmetric A block solver`: BiCGStab for the A block in the expensive phase."""
        self._check(self.lib.smf_set_option(self.ctx, _lib.SMF_OPT_NONSYMMETRIC_A_SOLVER, int(bool(on))))

    def get_cell_data(self, level):
        """The viscosity table of a level as the kernels see it, (n_cells, 1 or 3^dim)."""
        L = self.hier.active if level == SMF_ACTIVE else self.hier.levels[level]
        buf = np.zeros(L.n_cells * 3 ** self.dim)
        n = ctypes.c_int(0)
        self._check(self.lib.smf_get_cell_data(self.ctx, level, ptr(buf), ctypes.byref(n)))
        return buf[:L.n_cells * n.value].reshape(L.n_cells, n.value).copy()

    @property
    def active_table(self):
        return self._tables[0] if self._tables else self.get_cell_data(SMF_ACTIVE)

    @property
    def level_tables(self):
        return self._tables[1] if self._tables else [self.get_cell_data(l) for l in range(self.n_levels)]

    def set_cell_data(self, level, table, pressure_scaling, is_compressible=False):
        t = np.ascontiguousarray(table, dtype=np.float64)
        if t.ndim == 1:
            t = t[:, None]
        d = _lib.CellData(ptr(t), t.shape[1], float(pressure_scaling), int(is_compressible))
        self._check(self.lib.smf_set_cell_data(self.ctx, level, ctypes.byref(d)))

    def set_newton_data(self, enable_newton_derivatives=False, symmetrize_newton_system=False, average_newton_factors=False,
                        enable_prescribed_dilation=False, strain_rate_table=None, newton_factor_wrt_strain_rate_table=None,
                        newton_factor_wrt_pressure_table=None, dilation_lhs_term_table=None,
                        dilation_derivative_wrt_strain_rate_table=None, dilation_derivative_wrt_pressure_table=None):
        """The Newton / prescribed-dilation fields of the active `OperatorCellData` (matrix_free_operators.h:128-196; filled by
        local_smoothing.cc:296-612 in the reference): tables [cells][3^dim][n_sym] / [cells][3^dim], symmetric tensors in
        deal.II's unrolled order.  Call with no arguments to clear."""
        keep = [None if t is None else np.ascontiguousarray(t, dtype=np.float64) for t in (
            strain_rate_table, newton_factor_wrt_strain_rate_table, newton_factor_wrt_pressure_table, dilation_lhs_term_table,
            dilation_derivative_wrt_strain_rate_table, dilation_derivative_wrt_pressure_table)]
        if not (enable_newton_derivatives or enable_prescribed_dilation):
            self._check(self.lib.smf_set_newton_data(self.ctx, None))
            return
        d = _lib.NewtonData(int(enable_newton_derivatives), int(symmetrize_newton_system), int(average_newton_factors),
                            int(enable_prescribed_dilation), 0, *[None if t is None else t.ctypes.data for t in keep])
        self._check(self.lib.smf_set_newton_data(self.ctx, ctypes.byref(d)))

    def build_preconditioner(self):
        self._check(self.lib.smf_compute_diagonals(self.ctx))
        for l in sorted(getattr(self, "_level_lines", {})):
            self._constrained_level_diagonal(l)

    def _constrained_level_diagonal(self, l):
        """Masters' entries of diag(C^T A C) on a level with constraint lines (what MatrixFreeTools::compute_diagonal returns in
        the reference, matrix_free_operators.cc:1068-1098): smf_compute_diagonals knows the operator without the lines, so the
        level operator WITH its lines is probed with unit vectors.  Masters of nodes that share no cell do not interact (a line
        couples the components of one node only), so one probe per (colour, component) gives all their entries at once;
        nodes are coloured greedily over the cell cliques.  3 x (number of colours <= ~30) operator applications per level."""
        dof, start, master, weight = self._level_lines[l]
        L = self.hier.levels[l]
        nodes = np.unique(dof // self.dim)
        is_line_node = np.zeros(L.n_u // self.dim, bool); is_line_node[nodes] = True
        nn = 3 ** self.dim
        cell_nodes = (np.asarray(L.cell_dofs_u[:, :nn]) // self.dim).astype(np.int64)
        touched = np.flatnonzero(is_line_node[cell_nodes].any(axis=1))
        colour = np.full(L.n_u // self.dim, -1, dtype=np.int64)
        neigh = {int(n): set() for n in nodes}
        for c in touched:
            group = [int(n) for n in cell_nodes[c] if is_line_node[n]]
            for n in group:
                neigh[n].update(group)
        for n in nodes:                                   # greedy colouring in ascending node order
            used = {colour[m] for m in neigh[int(n)] if colour[m] >= 0}
            k = 0
            while k in used:
                k += 1
            colour[n] = k
        is_line_dof = np.zeros(L.n_u, bool); is_line_dof[dof] = True
        src, dst = self.initialize_dof_vector(0, l), self.initialize_dof_vector(0, l)
        idx, val = [], []
        for k in range(int(colour.max()) + 1):
            group = nodes[colour[nodes] == k]
            for c in range(self.dim):
                probe = self.dim * group + c
                probe = probe[~is_line_dof[probe]]           # masters only: the eliminated component has a unit diagonal
                if len(probe) == 0:
                    continue
                v = np.zeros(L.n_u); v[probe] = 1.0
                src.from_host(v)
                self.mg_matrices_A_block[l].vmult(dst, src)
                idx.append(probe); val.append(dst.to_host()[probe])
        idx = np.concatenate(idx); val = np.concatenate(val)
        self._check(self.lib.smf_set_inverse_diagonal_entries(self.ctx, l, 0, ctypes.c_int64(len(idx)),
                                                              ptr(np.ascontiguousarray(idx, dtype=np.uint32)),
                                                              ptr(np.ascontiguousarray(1.0 / val, dtype=np.float64))))

    def get_matrix_diagonal_inverse(self, level, which):
        L = self.hier.active if level == SMF_ACTIVE else self.hier.levels[level]
        out = np.empty(L.n_u if which == 0 else L.n_p)
        self._check(self.lib.smf_get_inverse_diagonal(self.ctx, level, which, ptr(out)))
        return out

    def setup_smoothers(self):
        self._check(self.lib.smf_setup_smoothers(self.ctx))

    def eigenvalue_estimates(self, level, which):
        out = np.zeros(3)
        self._check(self.lib.smf_get_eigenvalue_estimates(self.ctx, level, which, ptr(out)))
        return out

    def prolongate(self, which, to_level, dst, src):
        self._check(self.lib.smf_prolongate(self.ctx, which, to_level, ptr(dst.ptr), ptr(src.ptr)))

    def restrict_and_add(self, which, from_level, dst, src):
        self._check(self.lib.smf_restrict_add(self.ctx, which, from_level, ptr(dst.ptr), ptr(src.ptr)))

    # ------------------------------------------------------------------
    def _params(self, host_call=False):
        p = self.parameters
        methods = {"GMRES": 0, "IDR(s)": 1}              # parameters.cc:469-481
        if p["Krylov method for cheap solver steps"] not in methods:
            raise SmfError(_lib.SMF_ERR_INVALID, "Krylov method for cheap solver steps: GMRES | IDR(s)")
        return _lib.SolveParams(p["Linear solver tolerance"], p["Number of cheap Stokes solver steps"],
                                p["GMRES solver restart length"], methods[p["Krylov method for cheap solver steps"]],
                                p["IDR(s) parameter"], p["Maximum number of expensive Stokes solver steps"],
                                p["Linear solver A block tolerance"], p["Linear solver S block tolerance"],
                                int(host_call and self.pressure_scaling != 1.0), int(host_call and self._pressure_normalization_set()),
                                self._surface_pressure())

    def _surface_pressure(self):
        """`Surface pressure` is added only by `Pressure normalization = surface`; with `volume` the adjustment is
        -integral / volume alone (helper_functions.cc:858-861)."""
        p = self.parameters
        return float(p["Surface pressure"]) if p["Pressure normalization"] == "surface" else 0.0

    def _pressure_normalization_set(self):
        """Uploads the normalisation weights when "Pressure normalization" asks for them (once per mesh and kind)."""
        kind = self.parameters["Pressure normalization"]
        if kind == "no":
            return False
        if getattr(self, "_pnorm_kind", None) != kind:
            self.set_pressure_normalization(kind)
        return True

    def set_pressure_normalization(self, kind):
        """Weights of `Simulator::normalize_pressure` (helper_functions.cc:765-870) for the box meshes of this package:
        integral of every Q1 pressure shape function over this rank's cells ("volume") or over the top boundary faces
        ("surface", the faces with the last coordinate at its maximum)."""
        act, dim = self.hier.active, self.dim
        w = np.zeros(act.n_p)
        cp = np.asarray(act.cell_dofs_p, dtype=np.int64)
        if getattr(act, "jxw", None) is not None:
            # curved cells: integral of every Q1 shape function with the cell's JxW values, or over the outer boundary faces
            if kind == "volume":
                from .benchmarks import _phi1_at_quad
                np.add.at(w, cp, np.asarray(act.jxw) @ _phi1_at_quad(dim))
            elif kind == "surface":
                w, _ = act.surface_weights()
            else:
                raise SmfError(_lib.SMF_ERR_INVALID, "Pressure normalization: surface | volume | no")
        elif kind == "volume":
            np.add.at(w, cp.ravel(), float(np.prod(act.h)) / 2 ** dim)
        elif kind == "surface":
            top = np.asarray(act.cells)[:, dim - 1] == np.asarray(act.cells)[:, dim - 1].max()
            if getattr(self.hier, "world", 1) > 1:       # only the ranks whose subdomain touches the global top boundary
                top &= np.asarray(act.cells)[:, dim - 1] == act.grid_p - 2
            upper = [a for a in range(2 ** dim) if (a >> (dim - 1)) & 1]
            area = float(np.prod(act.h[:dim - 1])) / 2 ** (dim - 1)
            np.add.at(w, cp[top][:, upper].ravel(), area)
        else:
            raise SmfError(_lib.SMF_ERR_INVALID, "Pressure normalization: surface | volume | no")
        self._check(self.lib.smf_set_pressure_normalization(self.ctx, ptr(w)))
        self._pnorm_kind = kind

    def postprocess_pressure(self, sol_p, pressure_scaling=1.0, normalize=True):
        """Rescale and normalise a device pressure vector in place; returns the pressure normalisation adjustment."""
        adj = ctypes.c_double(0.0)
        self._check(self.lib.smf_postprocess_pressure(self.ctx, ptr(sol_p.ptr), ctypes.c_double(pressure_scaling), int(normalize),
                                                      ctypes.c_double(self._surface_pressure()), ctypes.byref(adj)))
        return adj.value

    def assemble_stokes_rhs(self, force_q, rhs_u):
        """Velocity block of the right-hand side on the device (assemblers/stokes.cc:387-396): force_q = density * gravity at
        the quadrature points, (n_cells, 3^dim, dim); rhs_u = DeviceVector of the velocity block."""
        f = np.ascontiguousarray(force_q, dtype=np.float64)
        assert f.shape == (self.hier.active.n_cells, 3 ** self.dim, self.dim)
        self._check(self.lib.smf_assemble_rhs(self.ctx, ptr(f), ptr(rhs_u.ptr)))

    def solve(self, system_rhs, solution_vector=None, out=None):
        """Host-buffer solve (the reference-facing call): system_rhs = concat(f_u, f_p); solution_vector = initial
        guess in physical units (or None).  Returns (solution, SolverOutputs-like dict).  Raises SmfError with
        code SMF_NO_CONVERGENCE if the cheap phase fails (SolverControl::NoConvergence).
        `out`: optional float64 array of n_u + n_p (e.g. pinned memory) that receives the solution in place."""
        nu, np_ = self.n_u, self.n_p
        rhs = np.ascontiguousarray(system_rhs, dtype=np.float64)
        if out is not None:
            sol = out
            sol[:] = 0.0 if solution_vector is None else solution_vector
        else:
            sol = np.zeros(nu + np_) if solution_vector is None else np.array(solution_vector, dtype=np.float64)
        if self.pressure_scaling != 1.0:
            sol[nu:] /= self.pressure_scaling                  # local_smoothing.cc:975
        su, sp = np.ascontiguousarray(sol[:nu]), np.ascontiguousarray(sol[nu:])
        ru, rp = np.ascontiguousarray(rhs[:nu]), np.ascontiguousarray(rhs[nu:])
        prm, res = self._params(host_call=True), _lib.SolveResult()      # (pressure rescaled / normalised on the device, :1421, :1441)
        rc = self.lib.smf_solve_host(self.ctx, ctypes.byref(prm), ptr(su), ptr(sp), ptr(ru), ptr(rp), ctypes.byref(res))
        out = self._outputs(res)
        self._check(rc, allow=(_lib.SMF_NO_CONVERGENCE,))
        out["converged"] = rc == 0
        if getattr(self, "adaptive", False) or getattr(self, "_lines", None):
            self.distribute_constraints(sol)
        return sol, out

    def distribute_constraints(self, sol):
        """`current_constraints.distribute(solution)` for the hanging-node lines (local_smoothing.cc:1424): the solve leaves
        zeros in the constrained entries; they take the value of their masters.  sol = host [u | p], changed in place."""
        for name, off in (("u", 0), ("p", self.n_u)):
            dof, start, master, weight = self._lines.get(name, (np.zeros(0, np.int64),) * 4)
            if len(dof):
                vals = np.add.reduceat(weight * sol[off + master], start[:-1]) if len(master) else np.zeros(len(dof))
                vals[start[:-1] == start[1:]] = 0.0
                sol[off + dof] = vals
        return sol

    def condense_rhs(self, rhs):
        """`constraints.distribute_local_to_global` on an assembled right-hand side: rhs += C^T-part of the hanging rows, then
        the constrained rows are zeroed.  rhs = host [u | p], changed in place."""
        for name, off in (("u", 0), ("p", self.n_u)):
            dof, start, master, weight = self._lines.get(name, (np.zeros(0, np.int64),) * 4)
            if len(dof):
                np.add.at(rhs, off + master, weight * np.repeat(rhs[off + dof], np.diff(start)))
                rhs[off + dof] = 0.0
        return rhs

    def correct_stokes_rhs(self, rhs, u0):
        """`StokesMatrixFreeHandler::correct_stokes_rhs` (local_smoothing.cc:617-754): rhs -= Stokes(u0) for inhomogeneous
        boundary values; rhs and u0 are (velocity, pressure) DeviceVector pairs, u0 = boundary values, zero elsewhere."""
        self._check(self.lib.smf_correct_stokes_rhs(self.ctx, ptr(rhs[0].ptr), ptr(rhs[1].ptr), ptr(u0[0].ptr), ptr(u0[1].ptr)))

    def solve_device(self, sol, rhs):
        """Device-resident solve: sol, rhs are (velocity, pressure) DeviceVector pairs (pressure already scaled)."""
        prm, res = self._params(), _lib.SolveResult()
        rc = self.lib.smf_solve(self.ctx, ctypes.byref(prm), ptr(sol[0].ptr), ptr(sol[1].ptr), ptr(rhs[0].ptr), ptr(rhs[1].ptr),
                                ctypes.byref(res))
        out = self._outputs(res)
        self._check(rc, allow=(_lib.SMF_NO_CONVERGENCE,))
        out["converged"] = rc == 0
        return out

    def execute_solver_timings(self, rhs, n_timings=10, out=print):
        """`Stokes solver parameters/Execute solver timings` (local_smoothing.cc:60-65, :1119-1260): warm-up + n_timings
        wall-clock timings (synchronised) of stokes_vmult (10 repeats), stokes_preconditioner, A_preconditioner,
        S_preconditioner (5 repeats), Stokes_solve_cheap_idr and Stokes_solve_cheap_gmres on device vectors, with the same
        names and the same "prepare" steps (src = dst between timings; the solves restart from the zero guess).
        rhs = host array [f_u | f_p].  Returns {name: average seconds per call}."""
        import time
        self.setup_smoothers()
        nu = self.n_u
        r = self.initialize_dof_vector(); d = self.initialize_dof_vector(); s = self.initialize_dof_vector()
        results = {}
        lib, c = self.lib, self._check

        def copy(dst, src):
            for b in (0, 1):
                c(lib.smf_vec_copy(self.ctx, ptr(dst[b].ptr), ptr(src[b].ptr), ctypes.c_int64(src[b].n)))

        def reset():
            r[0].from_host(rhs[:nu]); r[1].from_host(rhs[nu:])
            d[0].zero(); d[1].zero()
            copy(s, r)

        def time_this(name, repeats, body, prepare):
            out(f"Timing {name} {n_timings} time(s) and repeat {repeats} time(s) within each timing:")
            body(); self.synchronize()
            total = 0.0
            for _ in range(n_timings):
                prepare(); self.synchronize()
                t0 = time.perf_counter()
                for _ in range(repeats):
                    body()
                self.synchronize()
                dt = (time.perf_counter() - t0) / repeats
                out(f"\t...{dt:g}")
                total += dt
            results[name] = total / n_timings
            out(f"\taverage wall time of all: {results[name]:g} seconds")

        src_is_dst = lambda: copy(s, d)
        reset(); time_this("stokes_vmult", 10, lambda: self.stokes_matrix.vmult(d, s), src_is_dst)
        reset(); time_this("stokes_preconditioner", 1, lambda: self.preconditioner_cheap.vmult(d, s), src_is_dst)
        reset(); time_this("A_preconditioner", 1, lambda: self.prec_A.vmult(d[0], s[0]), src_is_dst)
        reset(); time_this("S_preconditioner", 5, lambda: self.prec_Schur.vmult(d[1], s[1]), src_is_dst)
        saved = dict(self.parameters)
        self.parameters["Number of cheap Stokes solver steps"] = 1000          # (:1211: always cheap iterations)
        self.parameters["Maximum number of expensive Stokes solver steps"] = 0
        zero = lambda: (d[0].zero(), d[1].zero())
        try:
            for name, method in (("Stokes_solve_cheap_idr", "IDR(s)"), ("Stokes_solve_cheap_gmres", "GMRES")):
                self.parameters["Krylov method for cheap solver steps"] = method
                reset(); time_this(name, 1, lambda: self.solve_device(d, r), zero)
        finally:
            self.parameters = saved
        for v in r + d + s:
            v.free()
        return results

    @staticmethod
    def _outputs(res):
        return {k: getattr(res, k) for k, _ in _lib.SolveResult._fields_}

    PROFILE_BUCKETS = ("stokes_vmult", "A_fine", "A_coarse", "mass", "BT", "transfer", "chebyshev_vec", "gmres_ortho", "other_vec",
                       "halo_exchange")

    def profile_enable(self, mode):
        """'Execute solver timings' analogue: 0 off, 1 finest-level operator kernels, 2 every launch."""
        self._check(self.lib.smf_profile_enable(self.ctx, int(mode)))

    def profile_read(self):
        ms = np.zeros(len(self.PROFILE_BUCKETS))
        cnt = np.zeros(len(self.PROFILE_BUCKETS), dtype=np.int64)
        self._check(self.lib.smf_profile_read(self.ctx, ptr(ms), ptr(cnt)))
        return {k: (float(m), int(c)) for k, m, c in zip(self.PROFILE_BUCKETS, ms, cnt)}

    def exchange_transport(self):
        return {0: "single rank", 1: "nccl send/recv", 2: "fused peer-memory kernel (CUDA IPC over NVLink)"}[int(self.lib.smf_exchange_transport(self.ctx))]

    def measure_fp64_peak(self):
        """Measured FP64 pipe peak of the device in TFLOP/s (DFMA micro-kernel inside the library)."""
        out = ctypes.c_double()
        self._check(self.lib.smf_measure_fp64_peak(self.ctx, ctypes.byref(out)))
        return float(out.value)

    def lattice_schur_active(self):
        """True when prec_Schur runs on the lattice path (box levels; decided by setup_smoothers)."""
        return bool(self.lib.smf_lattice_schur_active(self.ctx))

    def replicated_A_levels(self):
        """Several ranks: number of coarse prec_A levels every rank computes redundantly on the global mesh."""
        return int(self.lib.smf_replicated_A_levels(self.ctx))

    def lattice_replicated_levels(self):
        """Several ranks: number of coarse Schur levels every rank computes redundantly on the global lattice."""
        return int(self.lib.smf_lattice_replicated_levels(self.ctx))

    def exchange_count(self):
        return int(self.lib.smf_exchange_count(self.ctx))

    def kernel_launches(self):
        return int(self.lib.smf_kernel_launches(self.ctx))

    def get_cell_data_memory_consumption(self):
        return int(self.lib.smf_device_bytes(self.ctx))
