"""B200-native matrix-free GMG Stokes path (drop-in for ASPECT's `Stokes solver type = block GMG`).

Package contents (only what the path needs):
  csrc/            CUDA kernels + C ABI implementation  -> libsmf_b200.so (include/smf_b200.h)
  _lib.py          ctypes binding of the C ABI
  handler.py       StokesMatrixFreeHandler mirror (setup_dofs / assemble / build_preconditioner / solve)
  cell_data.py     host-side viscosity-table fill (stays in the plugin layer, as in the reference)
  synthetic.py     synthetic box-mesh driver producing the deal.II-style index maps
  benchmarks.py    nsinker / SolCx coefficient + right-hand-side generators (material "plugins")
"""
from .handler import StokesMatrixFreeHandler, DeviceVector  # noqa: F401
from ._lib import SmfError, SMF_ACTIVE  # noqa: F401
