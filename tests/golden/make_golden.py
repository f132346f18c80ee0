"""Generates tests/golden/ref_llvm.npz from the UNMODIFIED reference library.

Run in the build container (needs /root/reference + `make -C oracle`):

    python tests/golden/make_golden.py

Every output below is produced by the reference's own LLVMThreadState CPU
primitives (ext/drjit-core/src/llvm_ts.cpp:265-933) through oracle/_ref/libref_shim.so;
inputs are regenerated at test time from the fmix32 generator
(ext/drjit-core/tests/reductions.cpp:5-13), so only outputs are stored: integer
outputs (bit-exact contract) as a 64-bit SHA-256 digest of the raw bytes, float
outputs (tolerance contract) in full at sizes <= 333.
The fixture lets the oracle (and through it the CUDA path) be pinned on machines
where oracle/_ref is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import capi, ref  # noqa: E402

# (size, block_size) pairs: a subset of the reference's red_sizes grid (reductions.cpp:73-76)
SIZES = [1, 2, 3, 7, 10, 32, 60, 250, 333, 1024, 5000, 16384 + 4]
INT_TYPES = ["u8", "i32", "u32", "i64", "u64"]
FLT_TYPES = ["f16", "f32", "f64"]
INT_OPS = ["add", "mul", "min", "max", "and", "or"]
FLT_OPS = ["add", "mul", "min", "max"]


def make_input(vt, n):
    """Deterministic input for type `vt` (shared with tests/test_oracle.py)."""
    raw = capi.fmix32(n)
    if vt == "u8":
        return (raw & 0xFF).astype(np.uint8)
    if vt == "u32":
        return raw
    if vt == "i32":
        return raw.view(np.int32)
    if vt == "u64":
        return capi.fmix32_u64(n)
    if vt == "i64":
        return (capi.fmix32_u64(n) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)
    u = capi.unit_f32(n)
    if vt == "f32":
        return (u * 2 - 0.5).astype(np.float32)          # mixed sign, |x| < 1.5
    if vt == "f64":
        return (u.astype(np.float64) * 2 - 0.5)
    if vt == "f16":
        return (u * 2 - 0.5).astype(np.float16)
    raise ValueError(vt)


def digest(a):
    """First 8 bytes of SHA-256 over the raw little-endian bytes (+ length)."""
    import hashlib
    a = np.ascontiguousarray(a)
    h = hashlib.sha256(a.tobytes() + str(a.size).encode()).digest()
    return np.frombuffer(h[:8], np.uint64).copy()


def main():
    out = {}
    ref.lib()
    ref._lib.ref_llvm_set_thread_count(1)  # serial order == oracle order (float bit-exactness)
    for vt in INT_TYPES + FLT_TYPES:
        ops = INT_OPS if vt in INT_TYPES else FLT_OPS
        for n in SIZES:
            x = make_input(vt, n)
            for bs in SIZES:
                if bs > n:
                    continue
                for op in ops:
                    if op == "mul" and vt in FLT_TYPES and bs > 60:
                        continue  # products of many |x|<1.5 values under/overflow: uninformative
                    is_int = vt in INT_TYPES
                    if not is_int and n > 333:
                        continue
                    pack = digest if is_int else (lambda a: a)
                    out[f"br/{vt}/{op}/{n}/{bs}"] = pack(ref.block_reduce(vt, op, x, bs))
                    for ex in (0, 1):
                        for rev in (0, 1):
                            out[f"bp/{vt}/{op}/{n}/{bs}/{ex}{rev}"] = \
                                pack(ref.block_prefix_reduce(vt, op, x, bs, ex, rev))
    for vt in FLT_TYPES:
        for n in [1, 5, 100, 5000]:
            a, b = make_input(vt, n), make_input(vt, n)[::-1].copy()
            out[f"dot/{vt}/{n}"] = np.array([ref.reduce_dot(vt, a, b)])
    for n in [1, 9, 4095, 4096, 4097, 8192, 20001]:
        for thr in [0, 3, 128, 253, 256]:
            out[f"compress/{n}/{thr}"] = digest(ref.compress(capi.mask_u8(n, thr)))
    for n, buckets in [(1, 1), (24, 1), (185, 24), (622, 185), (1473, 622), (20001, 4096), (20001, 37)]:
        keys = capi.fmix32(n) % np.uint32(buckets)
        perm, offsets, unique = ref.block_mkperm(keys, n, buckets)
        out[f"mkperm/{n}/{buckets}/perm"] = digest(perm)
        out[f"mkperm/{n}/{buckets}/offsets"] = digest(offsets[:4 * unique])
        for bs in (7, 256):
            if bs < n:
                perm, _, _ = ref.block_mkperm(keys, bs, buckets, want_offsets=False)
                out[f"mkperm_block/{n}/{buckets}/{bs}"] = digest(perm)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_llvm.npz")
    # pack the 1-element digests into two parallel arrays (npz per-entry overhead)
    dkeys = sorted(k for k, v in out.items() if v.dtype == np.uint64 and v.size == 1 and not k.startswith("b") or
                   (k.startswith("b") and k.split("/")[1] in INT_TYPES))
    packed = {k: v for k, v in out.items() if k not in set(dkeys)}
    packed["digest_keys"] = np.array(dkeys)
    packed["digest_vals"] = np.array([out[k][0] for k in dkeys], np.uint64)
    np.savez_compressed(path, **packed)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
