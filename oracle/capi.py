"""numpy front-end of oracle.c (see its header). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# VarType / ReduceOp values: ext/drjit-core/include/drjit-core/jit.h:597-611, :990-1014
VT = {"bool": 1, "i8": 3, "u8": 4, "i16": 5, "u16": 6, "i32": 7, "u32": 8,
      "i64": 9, "u64": 10, "f16": 13, "f32": 14, "f64": 15}
OP = {"add": 1, "mul": 2, "min": 3, "max": 4, "and": 5, "or": 6}
NP = {"bool": np.uint8, "u8": np.uint8, "i32": np.int32, "u32": np.uint32, "i64": np.int64,
      "u64": np.uint64, "f16": np.float16, "f32": np.float32, "f64": np.float64}


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, os.path.join(_HERE, "liboracle.so")],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        vp, u32, u64, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
        L.oracle_fill_fmix32_u32.argtypes = [vp, u64, u64, u32, u32]
        L.oracle_fill_fmix32_u64.argtypes = [vp, u64, u64]
        L.oracle_fill_unit_f32.argtypes = [vp, u64, u64, u32]
        L.oracle_fill_mask_u8.argtypes = [vp, u64, u64, u32]
        L.oracle_reduce_identity.argtypes = [i32, i32]
        L.oracle_reduce_identity.restype = u64
        L.oracle_block_reduce.argtypes = [i32, i32, u32, u32, vp, vp, i32]
        L.oracle_block_prefix_reduce.argtypes = [i32, i32, u32, u32, i32, i32, vp, vp, i32]
        L.oracle_reduce_dot.argtypes = [i32, vp, vp, u32, vp, i32]
        L.oracle_compress.argtypes = [vp, u32, vp]
        L.oracle_compress.restype = u32
        L.oracle_block_mkperm.argtypes = [vp, u32, u32, u32, vp, vp]
        L.oracle_block_mkperm.restype = ctypes.c_int64
        L.oracle_all.argtypes = [vp, u32]
        L.oracle_any.argtypes = [vp, u32]
        L.oracle_scatter_reduce.argtypes = [i32, i32, vp, u32, vp, vp, vp, u32, i32]
        L.oracle_scatter_inc.argtypes = [vp, u32, vp, vp, u32, vp]
        L.oracle_memset.argtypes = [vp, u32, u32, vp]
        L.oracle_scatter_add_expand_f32.argtypes = [vp, u32, vp, vp, u32, u32, vp]
        L.oracle_aggregate.argtypes = [vp, vp, u32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class OracleError(RuntimeError):
    pass


def _check(rv, what):
    if rv == -1:
        raise OracleError(f"{what}: invalid block size / argument")
    if rv == -2:
        raise OracleError(f"{what}: unsupported type/op")
    if rv == -3:
        raise OracleError(f"{what}: index out of range")


# ----------------------------------------------------------------- generators
def fmix32(n, start=0, xor=0, mask=0xFFFFFFFF):
    out = np.empty(n, np.uint32)
    lib().oracle_fill_fmix32_u32(_p(out), start, n, xor, mask)
    return out


def fmix32_u64(n, start=0):
    out = np.empty(n, np.uint64)
    lib().oracle_fill_fmix32_u64(_p(out), start, n)
    return out


def unit_f32(n, start=0, xor=0):
    out = np.empty(n, np.float32)
    lib().oracle_fill_unit_f32(_p(out), start, n, xor)
    return out


def mask_u8(n, threshold, start=0):
    out = np.empty(n, np.uint8)
    lib().oracle_fill_mask_u8(_p(out), start, n, threshold)
    return out


# ----------------------------------------------------------------- primitives
def block_reduce(vt, op, x, block_size, acc64=False):
    x = np.ascontiguousarray(x, NP[vt])
    n = x.size
    if n == 0:
        return np.empty(0, NP[vt])
    if block_size == 0 or block_size > n:
        raise OracleError("block_reduce: invalid block size")
    out = np.empty((n + block_size - 1) // block_size, NP[vt])
    _check(lib().oracle_block_reduce(VT[vt], OP[op], n, block_size, _p(x), _p(out), int(acc64)), "block_reduce")
    return out


def block_prefix_reduce(vt, op, x, block_size, exclusive, reverse, acc64=False):
    x = np.ascontiguousarray(x, NP[vt])
    out = np.empty_like(x)
    _check(lib().oracle_block_prefix_reduce(VT[vt], OP[op], x.size, block_size, int(exclusive),
                                            int(reverse), _p(x), _p(out), int(acc64)), "block_prefix_reduce")
    return out


def reduce_dot(vt, a, b, acc64=False):
    a = np.ascontiguousarray(a, NP[vt]); b = np.ascontiguousarray(b, NP[vt])
    out = np.zeros(1, NP[vt])
    _check(lib().oracle_reduce_dot(VT[vt], _p(a), _p(b), a.size, _p(out), int(acc64)), "reduce_dot")
    return out[0]


def compress(mask):
    mask = np.ascontiguousarray(mask, np.uint8)
    out = np.empty(mask.size, np.uint32)
    c = lib().oracle_compress(_p(mask), mask.size, _p(out))
    return out[:c].copy()


def block_mkperm(keys, block_size, bucket_count, want_offsets=True):
    keys = np.ascontiguousarray(keys, np.uint32)
    perm = np.empty(keys.size, np.uint32)
    offsets = np.zeros(4 * bucket_count + 1, np.uint32) if want_offsets else None
    rv = lib().oracle_block_mkperm(_p(keys), keys.size, block_size, bucket_count, _p(perm), _p(offsets))
    if rv == -1:
        raise OracleError("block_mkperm: bucket_count cannot be zero")
    _check(rv, "block_mkperm")
    return perm, offsets, int(rv)


def all_(mask):
    mask = np.ascontiguousarray(mask, np.uint8)
    return bool(lib().oracle_all(_p(mask), mask.size))


def any_(mask):
    mask = np.ascontiguousarray(mask, np.uint8)
    return bool(lib().oracle_any(_p(mask), mask.size))


def scatter_reduce(vt, op, target, value, index, mask=None, acc64=False):
    target = np.array(target, NP[vt], copy=True)
    value = np.ascontiguousarray(value, NP[vt]); index = np.ascontiguousarray(index, np.uint32)
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    _check(lib().oracle_scatter_reduce(VT[vt], OP[op], _p(target), target.size, _p(value), _p(index),
                                       _p(m), value.size, int(acc64)), "scatter_reduce")
    return target


def scatter_reduce_packet(vt, op, target, values, index, mask=None, acc64=False):
    """Packet scatter-reduce, target[index[i] * n + k] op= values[k][i] (jit_var_scatter_packet,
    jit.h:1107-1120: "analogous to n separate scatters from indices index*n + [0, 1, .., n-1]";
    CUDA template src/cuda_packet.cpp:168-327). Restated exactly as that sentence: the n component
    scatters run through the scalar restatement above, serial in element order."""
    n = len(values)
    index = np.ascontiguousarray(index, np.uint32)
    flat_index = (index.astype(np.uint64)[:, None] * n + np.arange(n, dtype=np.uint64)[None, :]).reshape(-1)
    assert flat_index.size == 0 or int(flat_index.max()) < 2 ** 32
    flat_value = np.ascontiguousarray(np.stack([np.asarray(v, NP[vt]) for v in values], axis=1)).reshape(-1)
    flat_mask = np.repeat(np.ascontiguousarray(mask, np.uint8), n) if mask is not None else None
    return scatter_reduce(vt, op, target, flat_value, flat_index.astype(np.uint32), flat_mask, acc64)


def scatter_inc(target, index, mask=None, size=None):
    """dr.scatter_inc in serial element order: returns (target_after, out). index=None: counter 0."""
    target = np.array(target, np.uint32, copy=True)
    idx = np.ascontiguousarray(index, np.uint32) if index is not None else None
    n = idx.size if idx is not None else (int(size) if size is not None else np.asarray(mask).size)
    m = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    out = np.empty(n, np.uint32)
    _check(lib().oracle_scatter_inc(_p(target), target.size, _p(idx), _p(m), n, _p(out)), "scatter_inc")
    return target, out


def scatter_add_expand_f32(target, value, index, workers, scratch=None):
    """In-place dr.scatter_add in the reference's CPU ReduceMode.Expand form (see oracle.c)."""
    assert target.dtype == np.float32 and value.dtype == np.float32 and index.dtype == np.uint32
    if scratch is None:
        scratch = np.empty(workers * target.size, np.float32)
    _check(lib().oracle_scatter_add_expand_f32(_p(target), target.size, _p(value), _p(index), value.size,
                                               workers, _p(scratch)), "scatter_add_expand_f32")
    return target


def reduce_identity(vt, op):
    return int(lib().oracle_reduce_identity(VT[vt], OP[op]))
