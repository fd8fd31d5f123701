"""ctypes binding of the C ABI in include/smf_b200.h (aspect_b200/libsmf_b200.so).

There is deliberately no fallback: if the CUDA library is missing or no CUDA device is visible,
everything here raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmf_b200.so")

SMF_OK, SMF_ERR_INVALID, SMF_ERR_CUDA, SMF_ERR_UNSUPPORTED, SMF_NO_CONVERGENCE = 0, 1, 2, 3, 4
SMF_ACTIVE = -1
SMF_OPT_STRUCTURED_BRICKS, SMF_OPT_CUDA_GRAPHS, SMF_OPT_ORTHO_SINGLE_PASS, SMF_OPT_PIPELINED_KERNEL, SMF_OPT_MAX_GRID_BLOCKS = 0, 1, 2, 3, 4
SMF_OPT_P2P_EXCHANGE = 5
SMF_OPT_CURVED_SUMFAC = 6
SMF_OPT_CURVED_TMA = 7
SMF_OPT_WARP_BRICK_KERNEL = 8
SMF_OPT_COARSE_TAIL = 9
SMF_OPT_NONSYMMETRIC_A_SOLVER = 10
SMF_OPT_LATTICE_SCHUR = 11
SMF_OPT_LATTICE_VELOCITY = 12
SMF_OPT_REPLICATE_COARSE_A = 13
SMF_UNIQUE_ID_BYTES = 128

c_i64, c_int, c_dbl, c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p


class Config(ctypes.Structure):
    _fields_ = [("dim", c_int), ("degree", c_int), ("n_levels", c_int), ("device", c_int)]


class LevelMesh(ctypes.Structure):
    _fields_ = [("n_cells", c_i64), ("n_dofs_u", c_i64), ("n_dofs_p", c_i64),
                ("cell_dofs_u", c_vp), ("cell_dofs_p", c_vp),
                ("n_constrained_u", c_i64), ("constrained_u", c_vp),
                ("n_constrained_p", c_i64), ("constrained_p", c_vp),
                ("geometry", c_int), ("cell_size", c_vp), ("inverse_jacobians", c_vp), ("jxw", c_vp)]


class Exchange(ctypes.Structure):
    _fields_ = [("n_neighbors", c_int), ("neighbor_ranks", c_vp), ("counts", c_vp), ("nodes", c_vp), ("global_dof_ids", c_vp)]


class CellData(ctypes.Structure):
    _fields_ = [("viscosity", c_vp), ("n_per_cell", c_int), ("pressure_scaling", c_dbl), ("is_compressible", c_int)]


class NewtonData(ctypes.Structure):
    _fields_ = [("enable_newton_derivatives", c_int), ("symmetrize_newton_system", c_int), ("average_newton_factors", c_int),
                ("enable_prescribed_dilation", c_int), ("apply_stabilization_free_surface_faces", c_int),
                ("strain_rate_table", c_vp), ("newton_factor_wrt_strain_rate_table", c_vp),
                ("newton_factor_wrt_pressure_table", c_vp), ("dilation_lhs_term_table", c_vp),
                ("dilation_derivative_wrt_strain_rate_table", c_vp), ("dilation_derivative_wrt_pressure_table", c_vp)]


class SolveParams(ctypes.Structure):
    _fields_ = [("linear_solver_tolerance", c_dbl), ("n_cheap_stokes_solver_steps", c_int),
                ("gmres_restart_length", c_int), ("krylov_method", c_int), ("idr_s_parameter", c_int),
                ("n_expensive_stokes_solver_steps", c_int), ("linear_solver_A_block_tolerance", c_dbl),
                ("linear_solver_S_block_tolerance", c_dbl), ("rescale_pressure", c_int), ("normalize_pressure", c_int),
                ("surface_pressure", c_dbl)]


class SolveResult(ctypes.Structure):
    _fields_ = [("iterations", c_int), ("n_A_preconditioner", c_int), ("n_S_preconditioner", c_int),
                ("initial_residual", c_dbl), ("final_residual", c_dbl), ("tolerance", c_dbl),
                ("setup_seconds", c_dbl), ("solve_seconds", c_dbl), ("iterations_expensive", c_int),
                ("n_A_expensive", c_int), ("n_S_expensive", c_int), ("pressure_normalization_adjustment", c_dbl)]


# every symbol include/smf_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "smf_create", "smf_destroy", "smf_last_error", "smf_set_option", "smf_set_stream", "smf_set_level", "smf_set_transfer",
    "smf_set_copy_indices", "smf_set_constraints", "smf_set_refinement_edges", "smf_set_level_copy_indices", "smf_set_cell_data", "smf_fill_cell_data", "smf_get_cell_data", "smf_set_newton_data", "smf_set_free_surface_faces", "smf_vec_alloc", "smf_vec_free", "smf_vec_h2d", "smf_vec_d2h", "smf_vec_set", "smf_vec_copy",
    "smf_synchronize", "smf_vmult_stokes", "smf_vmult_A", "smf_vmult_BT", "smf_vmult_mass", "smf_vmult_stokes_host", "smf_correct_stokes_rhs",
    "smf_compute_diagonals", "smf_get_inverse_diagonal", "smf_set_inverse_diagonal_entries", "smf_setup_smoothers", "smf_get_eigenvalue_estimates",
    "smf_prolongate", "smf_restrict_add", "smf_vcycle_A", "smf_vcycle_S", "smf_precondition", "smf_solve",
    "smf_solve_host", "smf_kernel_launches", "smf_device_bytes", "smf_profile_enable", "smf_profile_read",
    "smf_exchange_count", "smf_exchange_transport", "smf_lattice_schur_active", "smf_lattice_replicated_levels", "smf_detect_box_lattice", "smf_replicated_A_levels", "smf_assemble_rhs", "smf_set_pressure_normalization",
    "smf_postprocess_pressure", "smf_measure_fp64_peak", "smf_comm_get_unique_id", "smf_comm_init", "smf_set_exchange", "smf_exchange_add",
]

_lib = None


class SmfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"smf error {code}: {msg}")
        self.code = code


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m aspect_b200.build` "
                               "(there is no CPU fallback for the block-GMG path)")
        lib = ctypes.CDLL(LIB_PATH)
        lib.smf_last_error.restype = ctypes.c_char_p
        lib.smf_last_error.argtypes = [c_vp]
        lib.smf_kernel_launches.restype = c_i64
        lib.smf_kernel_launches.argtypes = [c_vp]
        lib.smf_device_bytes.restype = c_i64
        lib.smf_device_bytes.argtypes = [c_vp]
        lib.smf_exchange_count.restype = c_i64
        lib.smf_exchange_count.argtypes = [c_vp]
        lib.smf_destroy.restype = None
        lib.smf_destroy.argtypes = [c_vp]
        _lib = lib
    return _lib


def ptr(a):
    """Host numpy array, raw device address (int) or None -> void*."""
    if a is None:
        return c_vp(0)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(c_vp)
    return c_vp(int(a))
