"""ctypes binding of ``lib/libdrjit_b200.so`` (the C ABI declared in ``include/drjit_b200.h``).

The product path has no fallback: if the shared library is missing this module raises
at import time, and every call fails loudly when the CUDA device is unusable.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRJIT_B200_LIB selects another build of the same library (developer tooling: the
# -DDRJIT_B200_EXPERIMENTS variant that scripts/ use for A/B timing). Never a fallback.
LIB_PATH = os.environ.get("DRJIT_B200_LIB") or os.path.join(_HERE, "lib", "libdrjit_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first "
        "(`python -c 'import __graft_entry__ as g; g.build()'` or `make -C drjit_b200/csrc`). "
        "drjit_b200 has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

vp, u32, u64, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
pu32, pint = ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); mirrors include/drjit_b200.h one to one
SIGNATURES = {
    "drjit_b200_last_error": (ctypes.c_char_p, []),
    "drjit_b200_version": (ctypes.c_char_p, []),
    "drjit_b200_init": (i32, []),
    "drjit_b200_shutdown": (i32, []),
    "drjit_b200_set_allocator": (i32, [vp, vp, vp]),
    "drjit_b200_launch_count": (u64, [i32]),
    "drjit_b200_reserve_scratch": (i32, [vp, ctypes.c_size_t]),
    "drjit_b200_set_launch_hook": (i32, [vp, vp]),
    "drjit_b200_set_flags": (i32, [u32]),
    "drjit_b200_flags": (u32, []),
    "drjit_b200_kernel_history": (u32, [vp, u32]),
    "drjit_b200_kernel_history_clear": (None, []),
    "drjit_b200_memset_async": (i32, [vp, vp, u32, u32, vp]),
    "drjit_b200_block_reduce": (i32, [vp, i32, i32, u32, u32, vp, vp]),
    "drjit_b200_block_reduce_bool": (i32, [vp, vp, u32, vp, i32]),
    "drjit_b200_all": (i32, [vp, vp, u32, pint]),
    "drjit_b200_any": (i32, [vp, vp, u32, pint]),
    "drjit_b200_reduce_dot": (i32, [vp, i32, vp, vp, u32, vp]),
    "drjit_b200_block_prefix_reduce": (i32, [vp, i32, i32, u32, u32, i32, i32, vp, vp]),
    "drjit_b200_compress": (i32, [vp, vp, u32, vp, pu32]),
    "drjit_b200_block_mkperm": (i32, [vp, vp, u32, u32, u32, vp, vp, pu32]),
    "drjit_b200_call_reduce": (i32, [vp, vp, u32, u32, vp, vp, u32, ctypes.POINTER(vp), ctypes.POINTER(vp), pu32]),
    "drjit_b200_sort": (i32, [vp, i32, u32, i32, vp, vp, vp]),
    "drjit_b200_poke": (i32, [vp, vp, vp, u32]),
    "drjit_b200_aggregate": (i32, [vp, vp, vp, u32]),
    "drjit_b200_scatter_reduce": (i32, [vp, i32, i32, i32, vp, u32, vp, vp, vp, u32]),
    "drjit_b200_scatter_reduce_packet": (i32, [vp, i32, i32, i32, vp, u32, vp, u32, vp, vp, u32]),
    "drjit_b200_scatter_inc": (i32, [vp, vp, u32, vp, vp, u32, vp]),
    "drjit_b200_prefix_reduce_carry": (i32, [vp, i32, i32, u32, i32, i32, vp, vp, vp, vp]),
    "drjit_b200_compress_async": (i32, [vp, vp, u32, u32, vp, vp]),
    "drjit_b200_mkperm_sharded": (i32, [vp, vp, u32, u32, u32, vp, vp]),
    "drjit_b200_fill_fmix32": (i32, [vp, i32, vp, u64, u64, u32, u32]),
    # multi-GPU forms (peer-memory communicator)
    "drjit_b200_comm_create": (i32, [u32, u32, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "drjit_b200_comm_handle": (i32, [vp, vp]),
    "drjit_b200_comm_connect": (i32, [vp, vp]),
    "drjit_b200_comm_connect_local": (i32, [ctypes.POINTER(vp), u32]),
    "drjit_b200_comm_destroy": (i32, [vp]),
    "drjit_b200_comm_reduce": (i32, [vp, vp, i32, i32, i32, u32, vp, vp]),
    "drjit_b200_comm_reduce_dot": (i32, [vp, vp, i32, vp, vp, u32, vp]),
    "drjit_b200_comm_all": (i32, [vp, vp, vp, u32, pint]),
    "drjit_b200_comm_any": (i32, [vp, vp, vp, u32, pint]),
    "drjit_b200_comm_prefix_reduce": (i32, [vp, vp, i32, i32, u32, i32, i32, vp, vp, vp, i32]),
    "drjit_b200_comm_compress": (i32, [vp, vp, vp, u32, u32, vp, pu32]),
    "drjit_b200_comm_mkperm": (i32, [vp, vp, vp, u32, u32, u32, vp, vp, vp, vp, pu32]),
    "drjit_b200_comm_allreduce": (i32, [vp, vp, i32, i32, vp, u32]),
    "drjit_b200_comm_allgather": (i32, [vp, vp, vp, u32, vp]),
    "drjit_b200_comm_fold": (i32, [vp, vp, i32, i32, i32, vp, vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export the header's symbol
    _fn.restype = _res
    _fn.argtypes = _args

OK, EINVAL, EUNSUPPORTED, ECUDA, EFATAL = 0, -1, -2, -3, -4


class FatalError(RuntimeError):
    """The reference would call jitc_fail() -> abort() here (CUDA error, bucket_count == 0 ...)."""


def check(status):
    """Map a status code to the exception the reference surfaces in Python:
    jitc_raise() -> std::runtime_error -> RuntimeError; jitc_fail() -> abort()."""
    if status == OK:
        return
    msg = lib.drjit_b200_last_error().decode()
    if status in (EINVAL, EUNSUPPORTED):
        raise RuntimeError(msg)
    raise FatalError(msg)
