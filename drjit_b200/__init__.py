"""drjit_b200 -- B200-native (sm_100a) data-parallel primitives behind Dr.Jit-Core's seam.

Layers (bottom-up):
  csrc/           hand-written CUDA kernels + the C ABI (``include/drjit_b200.h``)
  _lib.py         ctypes binding of the shared library (no fallback: import fails without it)
  ops.py          Python host mirror of the reference's operator surface for this path
                  (dr.sum / dr.block_reduce / dr.prefix_sum / dr.compress / dr.scatter_reduce /
                  dr.detail.block_mkperm ...) on torch CUDA tensors -- torch is only used for
                  device memory and streams
  dist.py         one-process-per-GPU sharding (NCCL only for the tiny combine messages)
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)
from .ops import (ReduceOp, ReduceMode, VarType, all, any, block_mkperm, block_prefix_reduce,  # noqa: F401,A004
                  block_prefix_sum, block_reduce, block_sum, compress, cumsum, dot, max, min,
                  prefix_sum, prod, scatter_add, scatter_reduce, sum, launch_count, version,
                  JitFlag, KernelType, set_flag, flag, kernel_history, kernel_history_clear,
                  reserve_scratch, sort, argsort, sort_with_indices, call_reduce,
                  scatter_reduce_packet, scatter_inc)

__all__ = ["ReduceOp", "ReduceMode", "VarType", "all", "any", "block_mkperm", "block_prefix_reduce",
           "block_prefix_sum", "block_reduce", "block_sum", "compress", "cumsum", "dot", "max",
           "min", "prefix_sum", "prod", "scatter_add", "scatter_reduce", "sum", "launch_count",
           "version", "JitFlag", "KernelType", "set_flag", "flag", "kernel_history",
           "kernel_history_clear", "reserve_scratch", "sort", "argsort", "sort_with_indices", "call_reduce",
           "scatter_reduce_packet", "scatter_inc"]
