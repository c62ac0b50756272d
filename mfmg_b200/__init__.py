"""mfmg_b200 -- B200-native (sm_100a) multigrid V-cycle apply behind mfmg's operator API.

  mfmg_b200.csrc       hand-written CUDA kernels + the C ABI (include/mfmg_b200.h)
  mfmg_b200.device     host-side mirror of the reference's device operator interface
                       (SparseMatrixDevice, CudaMatrixOperator, CudaSmoother, CudaSolver, Hierarchy)
  mfmg_b200.hostsetup  setup path that stays on the host (assembly, AMGe restrictor, R A R^T)
"""
__version__ = "0.1.0"
