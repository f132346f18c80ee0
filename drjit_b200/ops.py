"""Python host mirror of the reference's operator surface for the primitive path.

Names, argument meaning and error behaviour follow the Dr.Jit Python API so that the parity
tests read like the reference's own (citations relative to the reference tree):
  dr.sum/prod/min/max/all/any/dot ........ src/python/reduce.cpp:214-450
  dr.block_reduce / dr.block_sum ......... src/python/reduce.cpp:704-740
  dr.block_prefix_reduce / prefix_sum .... src/python/reduce.cpp:742-769, drjit/_reduce.py:269-312
  dr.compress ............................ src/python/reduce.cpp:659-674
  dr.scatter_reduce / scatter_add ........ src/python/memop.cpp:396-404 (ArrayNf values: packet form)
  dr.scatter_inc ......................... src/python/memop.cpp:408-445 (jit_var_scatter_inc, jit.h:1126-1143)
  dr.detail.block_mkperm ................. src/python/detail.cpp:481-520
Arrays are 1-D contiguous torch CUDA tensors; torch supplies device memory and the current
stream only. Every function goes through the C ABI (``_lib``); there is no torch fallback.
"""
import builtins
import ctypes
import enum

import torch

from . import _lib
from ._lib import check, lib


class VarType(enum.IntEnum):     # include/drjit-core/jit.h:597-611
    Void = 0; Bool = 1; Int8 = 3; UInt8 = 4; Int16 = 5; UInt16 = 6; Int32 = 7; UInt32 = 8
    Int64 = 9; UInt64 = 10; Float16 = 13; Float32 = 14; Float64 = 15


class ReduceOp(enum.IntEnum):    # include/drjit-core/jit.h:990-1014
    Identity = 0; Add = 1; Mul = 2; Min = 3; Max = 4; And = 5; Or = 6


class ReduceMode(enum.IntEnum):  # include/drjit-core/jit.h:1017-1066
    Auto = 0; Direct = 1; Local = 2; NoConflicts = 3; Expand = 4; Permute = 5


_DTYPE_TO_VT = {
    torch.bool: VarType.Bool, torch.uint8: VarType.UInt8, torch.int32: VarType.Int32,
    torch.int64: VarType.Int64, torch.float16: VarType.Float16, torch.float32: VarType.Float32,
    torch.float64: VarType.Float64,
}
for _name, _vt in (("uint32", VarType.UInt32), ("uint64", VarType.UInt64)):
    if hasattr(torch, _name):
        _DTYPE_TO_VT[getattr(torch, _name)] = _vt


def version():
    return lib.drjit_b200_version().decode()


def launch_count(reset=False):
    """Kernels launched by the library on this thread since the last reset."""
    return int(lib.drjit_b200_launch_count(int(reset)))


class KernelType(enum.IntEnum):  # include/drjit-core/jit.h:2597-2632 (+ this library's extensions >= 256)
    JIT = 0; BlockReduce = 1; BlockPrefixReduce = 2; Dot = 3; BatchedGemm = 4; Compress = 5
    MkPerm = 6; Memcpy = 7; Memset = 8; Poke = 9; Aggregate = 10; LLVMHostFunc = 11
    ScatterReduce = 256; Sort = 257; PeerExchange = 258


class JitFlag(enum.IntFlag):     # the two flags of jit.h:1680-1783 that act at the seam (cuda_ts.cpp:19-46)
    KernelHistory = 1; LaunchBlocking = 2


class _HistoryEntry(ctypes.Structure):
    _fields_ = [("type", ctypes.c_uint32), ("size", ctypes.c_uint32), ("launches", ctypes.c_uint32),
                ("execution_time", ctypes.c_float)]


def set_flag(flag, value=True):
    """dr.set_flag(dr.JitFlag.KernelHistory, True) -- thread-local, like the reference."""
    cur = int(lib.drjit_b200_flags())
    lib.drjit_b200_set_flags(cur | int(flag) if value else cur & ~int(flag))


def flag(flag):   # noqa: A002
    return bool(int(lib.drjit_b200_flags()) & int(flag))


def kernel_history(max_entries=4096):
    """dr.kernel_history(): list of dicts {backend, type, size, launches, execution_time (ms)} for the
    primitive calls recorded since the last call (src/python/history.cpp); clears the history."""
    buf = (_HistoryEntry * max_entries)()
    n = int(lib.drjit_b200_kernel_history(ctypes.cast(buf, ctypes.c_void_p), max_entries))
    return [{"backend": "cuda", "type": KernelType(buf[i].type), "size": int(buf[i].size),
             "launches": int(buf[i].launches), "execution_time": float(buf[i].execution_time)} for i in range(n)]


def kernel_history_clear():
    lib.drjit_b200_kernel_history_clear()


def reserve_scratch(nbytes, device=None):
    """Pre-size the library's scratch arena of the current stream (needed before CUDA-graph capture)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with _on(dev):
        check(lib.drjit_b200_reserve_scratch(ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream), int(nbytes)))


class _Noop:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOOP = _Noop()


def _on(device):
    """Context in which ``device`` is current: a no-op when it already is (the usual case, one
    process per GPU) -- entering torch.cuda.device() costs a few microseconds per call, which is
    visible next to primitives that take 20-100 us per shard."""
    idx = device.index if isinstance(device, torch.device) else device
    if idx is None or idx == torch.cuda.current_device():
        return _NOOP
    return torch.cuda.device(device)


def _vt(x, vt=None):
    if vt is not None:
        return VarType(vt)
    try:
        return _DTYPE_TO_VT[x.dtype]
    except KeyError:
        raise RuntimeError(f"drjit_b200: unsupported dtype {x.dtype}") from None


def _check_array(x, name="array"):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"drjit_b200: {name} must be a CUDA tensor (there is no CPU fallback)")
    if x.dim() != 1 or not x.is_contiguous():
        raise RuntimeError(f"drjit_b200: {name} must be a contiguous 1-D array")
    if x.numel() > 0xFFFFFFFF:
        raise RuntimeError("drjit_b200: arrays are limited to 2^32-1 entries (jitc_check_size)")
    return x


def _stream(x):
    return ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def _ptr(x):
    return ctypes.c_void_p(x.data_ptr()) if x is not None else None


# --------------------------------------------------------------------------- reductions
def block_reduce(op, value, block_size, vt=None, out=None):
    """dr.block_reduce(op, value, block_size): reduce contiguous blocks (last one may be short)."""
    x = _check_array(value)
    n = x.numel()
    with _on(x.device):
        if n == 0:
            return x.new_empty(0)
        blocks = (n + block_size - 1) // block_size if block_size else 1
        if out is None:
            out = torch.empty(blocks, dtype=x.dtype, device=x.device)
        check(lib.drjit_b200_block_reduce(_stream(x), _vt(x, vt), int(op), n, block_size, _ptr(x), _ptr(out)))
    return out


def block_sum(value, block_size, vt=None):
    return block_reduce(ReduceOp.Add, value, block_size, vt)


def _reduce(op, value, vt=None):
    x = _check_array(value)
    if x.numel() == 0:
        raise RuntimeError("drjit_b200: reduction of an empty array")
    return block_reduce(op, x, x.numel(), vt)


def sum(value, vt=None):   # noqa: A001
    """dr.sum(value): 1-element array (asynchronous, like the reference)."""
    return _reduce(ReduceOp.Add, value, vt)


def prod(value, vt=None):
    return _reduce(ReduceOp.Mul, value, vt)


def min(value, vt=None):   # noqa: A001
    return _reduce(ReduceOp.Min, value, vt)


def max(value, vt=None):   # noqa: A001
    return _reduce(ReduceOp.Max, value, vt)


def _all_any(fn, mask):
    m = _check_array(mask, "mask")
    if m.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError("drjit_b200: all()/any() expect a boolean mask")
    res = ctypes.c_int(0)
    with _on(m.device):
        check(fn(_stream(m), _ptr(m), m.numel(), ctypes.byref(res)))
    return bool(res.value)


def all(mask):  # noqa: A001
    """dr.all(mask) -> bool (synchronous, jitc_all)."""
    return _all_any(lib.drjit_b200_all, mask)


def any(mask):  # noqa: A001
    """dr.any(mask) -> bool (synchronous, jitc_any)."""
    return _all_any(lib.drjit_b200_any, mask)


def dot(a, b):
    """dr.dot(a, b) for floating point arrays: 1-element array."""
    a = _check_array(a, "a"); b = _check_array(b, "b")
    if a.dtype != b.dtype or a.numel() != b.numel():
        raise RuntimeError("drjit_b200: dot(): incompatible operands")
    out = torch.empty(1, dtype=a.dtype, device=a.device)
    with _on(a.device):
        check(lib.drjit_b200_reduce_dot(_stream(a), _vt(a), _ptr(a), _ptr(b), a.numel(), _ptr(out)))
    return out


# --------------------------------------------------------------------------- prefix reductions
def block_prefix_reduce(op, value, block_size, exclusive=True, reverse=False, vt=None, out=None):
    """dr.block_prefix_reduce(op, value, block_size, exclusive, reverse); ``out=value`` = in place."""
    x = _check_array(value)
    n = x.numel()
    if out is None:
        out = torch.empty_like(x)
    if n == 0:
        return out
    with _on(x.device):
        check(lib.drjit_b200_block_prefix_reduce(_stream(x), _vt(x, vt), int(op), n, block_size,
                                                 int(exclusive), int(reverse), _ptr(x), _ptr(out)))
    return out


def block_prefix_sum(value, block_size, exclusive=True, reverse=False, vt=None):
    return block_prefix_reduce(ReduceOp.Add, value, block_size, exclusive, reverse, vt)


def prefix_sum(value, exclusive=True, reverse=False, vt=None):
    """dr.prefix_sum(value): exclusive by default (drjit/_reduce.py)."""
    x = _check_array(value)
    if x.numel() == 0:
        return torch.empty_like(x)
    return block_prefix_reduce(ReduceOp.Add, x, x.numel(), exclusive, reverse, vt)


def cumsum(value, reverse=False, vt=None):
    """dr.cumsum(value): inclusive prefix sum."""
    return prefix_sum(value, exclusive=False, reverse=reverse, vt=vt)


# --------------------------------------------------------------------------- compress
def compress(mask):
    """dr.compress(mask): ascending indices of the true entries (synchronous: the result is
    shrunk to the count, like jitc_var_compress, src/var.cpp:2382-2415)."""
    m = _check_array(mask, "mask")
    if m.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError("drjit_b200: compress() expects a boolean mask")
    n = m.numel()
    idx_dtype = getattr(torch, "uint32", torch.int32)
    out = torch.empty(n, dtype=torch.int32, device=m.device)
    count = ctypes.c_uint32(0)
    with _on(m.device):
        check(lib.drjit_b200_compress(_stream(m), _ptr(m), n, _ptr(out), ctypes.byref(count)))
    res = out[:count.value]
    return res.view(idx_dtype) if idx_dtype is not torch.int32 else res


def compress_into(mask, out):
    """jit_compress() as the seam has it (cuda_ts.cpp:683-763): indices into the caller's buffer
    ``out`` (at least mask.numel() int32 entries), returns the count after the synchronisation."""
    m = _check_array(mask, "mask")
    if m.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError("drjit_b200: compress() expects a boolean mask")
    if out.numel() < m.numel() or out.element_size() != 4:
        raise RuntimeError("drjit_b200: compress_into() needs an output buffer of mask.numel() 32-bit entries")
    count = ctypes.c_uint32(0)
    with _on(m.device):
        check(lib.drjit_b200_compress(_stream(m), _ptr(m), m.numel(), _ptr(out), ctypes.byref(count)))
    return count.value


# --------------------------------------------------------------------------- mkperm
_pinned_cache = {}


def _pinned_offsets(bucket_count):
    """Host-pinned (device-mapped) table, jit_malloc(..., shared=true) in the reference."""
    need = 4 * bucket_count + 1
    buf = _pinned_cache.get("offsets")
    if buf is None or buf.numel() < need:
        buf = torch.empty(builtins.max(need, 1 << 16), dtype=torch.int32).pin_memory()
        _pinned_cache["offsets"] = buf
    return buf


def block_mkperm(values, block_size, bucket_count, want_offsets=True, perm=None, raw_table=False):
    """dr.detail.block_mkperm(values, block_size, bucket_count) -> (perm, offsets | None).

    ``perm`` is a device array; it is complete in stream order (the call itself only waits for
    the bucket table, cuda_ts.cpp:953-967). ``offsets`` is a host int64 numpy-like tensor of
    shape (unique, 4) with rows {bucket id, start, size, 0}, only for block_size == size.
    ``perm``: optional preallocated result. ``raw_table=True`` returns (perm, pinned offsets
    buffer, unique count) exactly as the seam function leaves them (what jit_var_call_reduce reads,
    call.cpp:1324-1378) -- the buffer is reused by the next call."""
    v = _check_array(values, "values")
    if _vt(v) not in (VarType.UInt32, VarType.Int32):
        raise RuntimeError("drjit_b200: block_mkperm() expects 32-bit integer keys")
    n = v.numel()
    if perm is None:
        perm = torch.empty(n, dtype=torch.int32, device=v.device)
    unique = ctypes.c_uint32(0)
    offsets = _pinned_offsets(bucket_count) if (want_offsets and block_size == n and n > 0) else None
    with _on(v.device):
        check(lib.drjit_b200_block_mkperm(_stream(v), _ptr(v), n, block_size, bucket_count, _ptr(perm),
                                          _ptr(offsets), ctypes.byref(unique)))
    if raw_table:
        return perm, offsets, unique.value
    table = None
    if offsets is not None:
        table = offsets[:4 * unique.value].clone().view(-1, 4).to(torch.int64) & 0xFFFFFFFF
    return perm, table


def call_reduce(ids, bucket_count, payloads=()):
    """jit_var_call_reduce (src/call.cpp:1268-1389) for evaluated callable IDs: returns
    ``(perm, table, permuted)`` where ``table`` is a host int64 tensor (unique, 4) with rows
    {bucket id, start, size, 0} ordered by decreasing size, and ``permuted[k] = payloads[k][perm]``
    (32-bit element arrays, at most 4) produced by the same scatter pass."""
    v = _check_array(ids, "ids")
    if _vt(v) not in (VarType.UInt32, VarType.Int32):
        raise RuntimeError("drjit_b200: call_reduce() expects 32-bit integer callable IDs")
    n = v.numel()
    pays = [_check_array(p, "payload") for p in payloads]
    for p in pays:
        if p.element_size() != 4 or p.numel() != n:
            raise RuntimeError("drjit_b200: call_reduce(): payloads must be 32-bit arrays of the size of `ids`")
    outs = [torch.empty_like(p) for p in pays]
    perm = torch.empty(n, dtype=torch.int32, device=v.device)
    offsets = _pinned_offsets(bucket_count)
    unique = ctypes.c_uint32(0)
    k = len(pays)
    arr_in = (ctypes.c_void_p * builtins.max(k, 1))(*[p.data_ptr() for p in pays])
    arr_out = (ctypes.c_void_p * builtins.max(k, 1))(*[o.data_ptr() for o in outs])
    with _on(v.device):
        check(lib.drjit_b200_call_reduce(_stream(v), _ptr(v), n, bucket_count, _ptr(perm), _ptr(offsets), k,
                                         arr_in, arr_out, ctypes.byref(unique)))
    table = offsets[:4 * unique.value].clone().view(-1, 4).to(torch.int64) & 0xFFFFFFFF
    return perm, table, outs


# --------------------------------------------------------------------------- sort
def _sort(value, descending, want_values, want_indices, vt=None):
    x = _check_array(value)
    t = _vt(x, vt)
    if t not in (VarType.Int32, VarType.UInt32, VarType.Float32, VarType.Int64, VarType.UInt64, VarType.Float64):
        raise RuntimeError(f"drjit_b200: sort(): unsupported type {t.name}")
    n = x.numel()
    values = torch.empty_like(x) if want_values else None
    index = torch.empty(n, dtype=torch.int32, device=x.device) if want_indices else None
    if n:
        with _on(x.device):
            check(lib.drjit_b200_sort(_stream(x), int(t), n, int(bool(descending)), _ptr(x), _ptr(values), _ptr(index)))
    return values, index


def sort(value, descending=False, vt=None):
    """dr.sort(value, descending=False) for flat arrays (drjit/__init__.py:1784-1850)."""
    return _sort(value, descending, True, False, vt)[0]


def argsort(value, descending=False, vt=None):
    """dr.argsort(value, descending=False): stable sorting permutation (UInt32 bits in an int32 tensor)."""
    return _sort(value, descending, False, True, vt)[1]


def sort_with_indices(value, descending=False, vt=None):
    """(dr.sort(value), dr.argsort(value)) from one run of the passes."""
    return _sort(value, descending, True, True, vt)


# --------------------------------------------------------------------------- scatter-reduce
def scatter_reduce(op, target, value, index, active=None, mode=ReduceMode.Auto, vt=None):
    """dr.scatter_reduce(op, target, value, index, active, mode): in-place on ``target``. A list / tuple
    of component arrays as ``value`` is the packet form (``dr.scatter_reduce`` of an ArrayNf,
    src/python/memop.cpp:396-404 -> jit_var_scatter_packet): see scatter_reduce_packet."""
    if isinstance(value, (list, tuple)):
        return scatter_reduce_packet(op, target, value, index, active, mode, vt)
    t = _check_array(target, "target"); val = _check_array(value, "value"); idx = _check_array(index, "index")
    if t.dtype != val.dtype:
        raise RuntimeError("drjit_b200: scatter_reduce(): target/value type mismatch")
    if _vt(idx) not in (VarType.UInt32, VarType.Int32) or idx.numel() != val.numel():
        raise RuntimeError("drjit_b200: scatter_reduce(): index must be a 32-bit integer array of matching size")
    m = None
    if active is not None:
        m = _check_array(active, "active")
        if m.dtype not in (torch.bool, torch.uint8) or m.numel() != val.numel():
            raise RuntimeError("drjit_b200: scatter_reduce(): invalid mask")
    with _on(t.device):
        check(lib.drjit_b200_scatter_reduce(_stream(t), _vt(t, vt), int(op), int(mode), _ptr(t), t.numel(),
                                            _ptr(val), _ptr(idx), _ptr(m), val.numel()))
    return target


def scatter_add(target, value, index, active=None, mode=ReduceMode.Auto):
    return scatter_reduce(ReduceOp.Add, target, value, index, active, mode)


def _check_mask(active, n, what):
    if active is None:
        return None
    m = _check_array(active, "active")
    if m.dtype not in (torch.bool, torch.uint8) or m.numel() != n:
        raise RuntimeError("drjit_b200: %s(): invalid mask" % what)
    return m


def scatter_reduce_packet(op, target, values, index, active=None, mode=ReduceMode.Auto, vt=None):
    """Packet scatter-reduce (jit_var_scatter_packet, jit.h:1107-1120; PTX template
    src/cuda_packet.cpp:168-327): ``target[index[i] * n + k] op= values[k][i]`` for the ``n`` component
    arrays in ``values``; ``target`` is the flat array of packets. In-place on ``target``."""
    t = _check_array(target, "target"); idx = _check_array(index, "index")
    vals = [_check_array(v, "value") for v in values]
    n = len(vals)
    if n == 0 or builtins.any(v.dtype != t.dtype or v.numel() != idx.numel() for v in vals):
        raise RuntimeError("drjit_b200: scatter_reduce_packet(): components must match the target's type "
                           "and the index array's size")
    if _vt(idx) not in (VarType.UInt32, VarType.Int32):
        raise RuntimeError("drjit_b200: scatter_reduce_packet(): index must be a 32-bit integer array")
    if t.numel() % n:
        raise RuntimeError("drjit_b200: scatter_reduce_packet(): target size is not a multiple of the packet size")
    m = _check_mask(active, idx.numel(), "scatter_reduce_packet")
    ptrs = (ctypes.c_void_p * n)(*[_ptr(v) for v in vals])
    with _on(t.device):
        check(lib.drjit_b200_scatter_reduce_packet(_stream(t), _vt(t, vt), int(op), int(mode), _ptr(t),
                                                   t.numel() // n, ptrs, n, _ptr(idx), _ptr(m), idx.numel()))
    return target


def scatter_inc(target, index=None, active=None, size=None, out=None):
    """dr.scatter_inc(target, index, active) (src/python/memop.cpp:408-445 -> jit_var_scatter_inc, jit.h:1126-1143;
    PTX template src/cuda_scatter.cpp:356-393): atomically ``out[i] = target[index[i]]++``; masked
    elements receive 0. ``index=None`` with ``size`` (or a mask) is the queue form
    ``dr.scatter_inc(counter, 0)``: every element increments ``target[0]``."""
    t = _check_array(target, "target")
    if _vt(t) not in (VarType.UInt32, VarType.Int32):
        raise RuntimeError("drjit_b200: scatter_inc(): target must be a 32-bit integer array")
    idx = None
    if index is not None:
        idx = _check_array(index, "index")
        if _vt(idx) not in (VarType.UInt32, VarType.Int32):
            raise RuntimeError("drjit_b200: scatter_inc(): index must be a 32-bit integer array")
        n = idx.numel()
    elif size is not None:
        n = int(size)
    elif active is not None:
        n = active.numel()
    else:
        raise RuntimeError("drjit_b200: scatter_inc(): without an index array, pass size= or a mask")
    m = _check_mask(active, n, "scatter_inc")
    if out is None:
        out = torch.empty(n, dtype=t.dtype, device=t.device)
    elif out.numel() != n or out.dtype != t.dtype:
        raise RuntimeError("drjit_b200: scatter_inc(): invalid output array")
    with _on(t.device):
        check(lib.drjit_b200_scatter_inc(_stream(t), _ptr(t), t.numel(), _ptr(idx), _ptr(m), n, _ptr(out)))
    return out


# --------------------------------------------------------------------------- misc
def memset(tensor, value_bytes):
    """jit_memset_async: fill ``tensor`` (any dtype) with the repeated byte pattern."""
    x = _check_array(tensor)
    isize = len(value_bytes)
    buf = ctypes.create_string_buffer(bytes(value_bytes), isize)
    nbytes = x.numel() * x.element_size()
    if nbytes % isize:
        raise RuntimeError("drjit_b200: memset(): size is not a multiple of the pattern")
    with _on(x.device):
        check(lib.drjit_b200_memset_async(_stream(x), _ptr(x), nbytes // isize, isize, buf))
    return tensor


def fill_fmix32(out, kind, start=0, xor=0, and_=0xFFFFFFFF):
    """Synthetic benchmark input generated on the device (tests/reductions.cpp:5-13)."""
    x = _check_array(out)
    with _on(x.device):
        check(lib.drjit_b200_fill_fmix32(_stream(x), kind, _ptr(x), start, x.numel(), xor, and_))
    return out


# --------------------------------------------------------------------------- shard-local forms
def prefix_reduce_carry(op, value, exclusive=True, reverse=False, carry_in=None, total_out=None,
                        vt=None, out=None):
    """Full-array prefix reduction of one shard whose running value starts at ``carry_in``
    (1-element device tensor, None = identity); ``total_out`` (1-element device tensor)
    receives op(carry_in, reduce(shard)). Asynchronous."""
    x = _check_array(value)
    if out is None:
        out = torch.empty_like(x)
    with _on(x.device):
        check(lib.drjit_b200_prefix_reduce_carry(_stream(x), _vt(x, vt), int(op), x.numel(), int(exclusive),
                                                 int(reverse), _ptr(x), _ptr(out), _ptr(carry_in),
                                                 _ptr(total_out)))
    return out


def compress_async(mask, index_base=0, out=None, count=None):
    """Asynchronous compress of one shard: indices are offset by ``index_base``; returns
    (out, count) where ``count`` is a 1-element device tensor (no host synchronisation)."""
    m = _check_array(mask, "mask")
    if out is None:
        out = torch.empty(m.numel(), dtype=torch.int32, device=m.device)
    if count is None:
        count = torch.zeros(1, dtype=torch.int32, device=m.device)
    with _on(m.device):
        check(lib.drjit_b200_compress_async(_stream(m), _ptr(m), m.numel(), index_base, _ptr(out), _ptr(count)))
    return out, count


def mkperm_sharded(values, bucket_count, index_base=0, perm=None, hist=None):
    """Shard-local mkperm: (perm with entries index_base + i, per-bucket histogram), asynchronous."""
    v = _check_array(values, "values")
    if perm is None:
        perm = torch.empty(v.numel(), dtype=torch.int32, device=v.device)
    if hist is None:
        hist = torch.empty(bucket_count, dtype=torch.int32, device=v.device)
    with _on(v.device):
        check(lib.drjit_b200_mkperm_sharded(_stream(v), _ptr(v), v.numel(), bucket_count, index_base,
                                            _ptr(perm), _ptr(hist)))
    return perm, hist
