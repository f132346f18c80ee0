"""ctypes front-end of the UNMODIFIED reference drjit-core built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (oracle pinning, incumbent GPU kernels, CPU baseline).
Backend ids follow JitBackend (include/drjit-core/jit.h:46-60): CUDA = 1, LLVM = 2.
With backend=LLVM all pointers are host (numpy) pointers; with backend=CUDA they
are device pointers (pass integers, e.g. torch ``tensor.data_ptr()``).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ORACLE_REF_DIR=oracle/_ref_b200 selects the build whose CUDA primitives are libdrjit_b200.so's
# (oracle/ref_build `make b200`; tests/test_insitu_gpu.py runs in a subprocess with it)
REF_DIR = os.environ.get("ORACLE_REF_DIR") or os.path.join(_HERE, "_ref")
CUDA, LLVM = 1, 2

_lib = None
_backends = 0


def available():
    return os.path.exists(os.path.join(REF_DIR, "libref_shim.so"))


def lib(cuda=False, llvm=True):
    """Load the shim and initialise the requested reference backends (idempotent)."""
    global _lib, _backends
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle`)")
        # The reference dlopens libLLVM; point it at the stub unless the user has a real one
        os.environ.setdefault("DRJIT_LIBLLVM_PATH", os.path.join(_HERE, "_ref", "libLLVM.so"))
        L = ctypes.CDLL(os.path.join(REF_DIR, "libref_shim.so"))
        vp, u32, i32, i64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int64
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_init.argtypes = [i32, i32]
        L.ref_malloc.argtypes = [i32, ctypes.c_size_t, i32]; L.ref_malloc.restype = vp
        L.ref_free.argtypes = [vp]
        L.ref_memcpy.argtypes = [i32, vp, vp, ctypes.c_size_t]
        L.ref_cuda_stream.restype = vp
        L.ref_llvm_set_thread_count.argtypes = [u32]
        L.ref_block_reduce.argtypes = [i32, i32, i32, u32, u32, vp, vp]
        L.ref_block_prefix_reduce.argtypes = [i32, i32, i32, u32, u32, i32, i32, vp, vp]
        L.ref_compress.argtypes = [i32, vp, u32, vp]; L.ref_compress.restype = i64
        L.ref_block_mkperm.argtypes = [i32, vp, u32, u32, u32, vp, vp]; L.ref_block_mkperm.restype = i64
        L.ref_reduce_dot.argtypes = [i32, i32, vp, vp, u32, vp]
        L.ref_scatter_reduce.argtypes = [i32, i32, i32, i32, vp, u32, vp, vp, u32]
        if hasattr(L, "ref_kernel_history"):
            L.ref_set_flag.argtypes = [u32, i32]
            L.ref_kernel_history.argtypes = [vp, vp, vp, vp, u32]; L.ref_kernel_history.restype = u32
        if hasattr(L, "ref_scatter_reduce_masked"):
            L.ref_scatter_reduce_masked.argtypes = [i32, i32, i32, i32, vp, u32, vp, vp, vp, u32]
            L.ref_kernel_history_ir_count.argtypes = [ctypes.c_char_p]; L.ref_kernel_history_ir_count.restype = u32
        if hasattr(L, "ref_scatter_packet"):
            L.ref_scatter_packet.argtypes = [i32, i32, i32, i32, vp, u32, ctypes.POINTER(vp), u32, vp, vp, u32]
            L.ref_scatter_inc.argtypes = [i32, vp, u32, vp, vp, u32, vp]
        _lib = L
    want = (2 if cuda else 0) | (4 if llvm else 0)
    if want & ~_backends:
        _backends |= _lib.ref_init(int(cuda), int(llvm))
    return _lib


def has_backend(b):
    return bool(_backends & (1 << b))


class RefError(RuntimeError):
    pass


def _chk(rv):
    if rv != 0:
        raise RefError(_lib.ref_last_error().decode())


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return a.ctypes.data_as(ctypes.c_void_p)


from .capi import VT, OP, NP  # noqa: E402  (shared enum tables)


# ------------------------------------------------ LLVM-backend (host) helpers
def block_reduce(vt, op, x, block_size):
    L = lib()
    x = np.ascontiguousarray(x, NP[vt]); n = x.size
    out = np.empty(max(1, (n + max(block_size, 1) - 1) // max(block_size, 1)), NP[vt])
    _chk(L.ref_block_reduce(LLVM, VT[vt], OP[op], n, block_size, _p(x), _p(out))); L.ref_sync()
    return out if n else out[:0]


def block_prefix_reduce(vt, op, x, block_size, exclusive, reverse):
    L = lib()
    x = np.ascontiguousarray(x, NP[vt]); out = np.empty_like(x)
    _chk(L.ref_block_prefix_reduce(LLVM, VT[vt], OP[op], x.size, block_size, int(exclusive), int(reverse),
                                   _p(x), _p(out))); L.ref_sync()
    return out


def reduce_dot(vt, a, b):
    L = lib()
    a = np.ascontiguousarray(a, NP[vt]); b = np.ascontiguousarray(b, NP[vt]); out = np.zeros(1, NP[vt])
    _chk(L.ref_reduce_dot(LLVM, VT[vt], _p(a), _p(b), a.size, _p(out)))
    return out[0]


def compress(mask):
    L = lib()
    mask = np.ascontiguousarray(mask, np.uint8); out = np.empty(mask.size, np.uint32)
    c = L.ref_compress(LLVM, _p(mask), mask.size, _p(out))
    if c < 0:
        raise RefError(L.ref_last_error().decode())
    return out[:c].copy()


def block_mkperm(keys, block_size, bucket_count, want_offsets=True):
    L = lib()
    keys = np.ascontiguousarray(keys, np.uint32); perm = np.empty(keys.size, np.uint32)
    offsets = np.zeros(4 * bucket_count + 1, np.uint32) if want_offsets else None
    rv = L.ref_block_mkperm(LLVM, _p(keys), keys.size, block_size, bucket_count, _p(perm), _p(offsets))
    if rv < 0:
        raise RefError(L.ref_last_error().decode())
    L.ref_sync()
    return perm, offsets, int(rv)
