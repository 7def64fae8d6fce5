"""densepoints_b200 -- B200 (sm_100a) implementation of the photometric hot path of
DensePoints' PMVS method (scoring, refinement, expansion) behind a C ABI.

  capi      ctypes binding of include/densepoints_cuda.h (no CPU fallback)
  scenes    synthetic multi-view scenes + seed patches (numpy)
  build     nvcc build of csrc/ -> _build/libdensepoints_cuda.so
"""
__all__ = ["capi", "scenes", "build"]
