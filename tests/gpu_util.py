"""Helpers shared by the GPU parity tests (numpy <-> torch views that keep the bit pattern)."""
import numpy as np
import torch

import drjit_b200 as dr
from drjit_b200 import VarType

NP_SIGNED_VIEW = {"u8": np.uint8, "u32": np.int32, "i32": np.int32, "u64": np.int64, "i64": np.int64,
                  "f16": np.float16, "f32": np.float32, "f64": np.float64}
VT = {"u8": VarType.UInt8, "u32": VarType.UInt32, "i32": VarType.Int32, "u64": VarType.UInt64,
      "i64": VarType.Int64, "f16": VarType.Float16, "f32": VarType.Float32, "f64": VarType.Float64}
NP = {"u8": np.uint8, "u32": np.uint32, "i32": np.int32, "u64": np.uint64, "i64": np.int64,
      "f16": np.float16, "f32": np.float32, "f64": np.float64}
OPS = {"add": dr.ReduceOp.Add, "mul": dr.ReduceOp.Mul, "min": dr.ReduceOp.Min, "max": dr.ReduceOp.Max,
       "and": dr.ReduceOp.And, "or": dr.ReduceOp.Or}


def to_dev(a, vt, misalign=0):
    """numpy array -> CUDA tensor holding the same bytes (unsigned types travel as signed views).
    misalign > 0 returns a view starting `misalign` elements into a larger allocation."""
    a = np.ascontiguousarray(a, NP[vt])
    host = torch.from_numpy(a.view(NP_SIGNED_VIEW[vt]))
    if misalign == 0:
        return host.cuda()
    buf = torch.empty(a.size + misalign, dtype=host.dtype, device="cuda")
    buf[misalign:].copy_(host)
    return buf[misalign:]


def to_np(t, vt):
    return t.cpu().numpy().view(NP[vt])
