"""cubecl_b200: B200-native (sm_100a) implementation of CubeCL's dense linear-algebra hot path.

  matmul.launch / reduce.launch over ComputeClient + TensorHandle  ->  C ABI (include/cubecl_b200.h)
  ->  prebuilt sm_100a cubins: tcgen05/TMA GEMM (csrc/gemm_tcgen05.cu), HBM-bound reductions (csrc/reduce.cu).

There is no CPU implementation in this package; the CPU oracle lives in /oracle and is test infrastructure only.
"""
from . import matmul, reduce, synth  # noqa: F401
from ._ffi import B200Error  # noqa: F401
from .client import ComputeClient, Handle, ServerError, TensorHandle  # noqa: F401

__all__ = ["ComputeClient", "Handle", "TensorHandle", "ServerError", "B200Error", "matmul", "reduce", "synth"]
