"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol
that include/drjit_b200.h declares (and nothing else), the C++ adapter compiles against it,
and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drjit_b200", "lib", "libdrjit_b200.so")
HEADER = os.path.join(ROOT, "include", "drjit_b200.h")


def header_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"DRJIT_B200_API\s+[\w\s\*]+?\b(drjit_b200_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for name in ["drjit_b200_block_reduce", "drjit_b200_block_prefix_reduce", "drjit_b200_reduce_dot",
                 "drjit_b200_compress", "drjit_b200_block_mkperm", "drjit_b200_scatter_reduce",
                 "drjit_b200_memset_async", "drjit_b200_poke", "drjit_b200_aggregate",
                 "drjit_b200_block_reduce_bool", "drjit_b200_all", "drjit_b200_any"]:
        assert name in syms
    assert len(syms) >= 20


def test_library_exports_exactly_the_header():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == header_symbols()
    lib = ctypes.CDLL(LIB)
    for name in header_symbols():
        getattr(lib, name)


def test_shipped_library_reads_no_environment():
    """Developer switches (phase toggles, tile sizes, the experimental kernels) exist only in the
    -DDRJIT_B200_EXPERIMENTS build that scripts/ use: the shipped library carries no DRJIT_B200_*
    variable name and its sources call getenv() only inside experiments-only blocks."""
    blob = open(LIB, "rb").read()
    # (getenv itself is still imported: the statically linked CUDA runtime reads CUDA_* variables)
    assert not re.search(rb"DRJIT_B200_[A-Z0-9_]+", blob), "an environment variable name is compiled in"
    assert b"scatter16" not in blob, "the experimental 16-bit staging kernel is in the shipped library"
    src = "".join(open(os.path.join(ROOT, "drjit_b200", "csrc", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "drjit_b200", "csrc")) if f.endswith((".cu", ".cuh", ".h")))
    outside = re.sub(r"#if defined\(DRJIT_B200_EXPERIMENTS\).*?#e(?:lse|ndif)", "", src, flags=re.S)
    assert "getenv" not in outside, "getenv() outside an experiments-only block"


def test_python_binding_covers_every_symbol():
    from drjit_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_no_torch_types_in_the_abi():
    src = open(HEADER).read()
    assert "torch" not in src and "at::" not in src and "#include <cuda" not in src


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import drjit_b200 as dr
    from drjit_b200._lib import lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dr.sum(torch.ones(8))
    # straight through the C ABI: CUDA error code, never a silent CPU result
    buf = (ctypes.c_uint32 * 8)()
    rv = lib.drjit_b200_block_reduce(None, 8, 1, 8, 8, buf, buf)
    assert rv in (-3, -4)
    assert lib.drjit_b200_last_error().decode() != ""
    # argument errors are still reported with the reference's wording
    rv = lib.drjit_b200_block_reduce(None, 8, 1, 8, 0, buf, buf)
    assert rv == -1 and b"invalid block size" in lib.drjit_b200_last_error()


def test_product_never_touches_the_oracle():
    """A product path that routes through oracle/ would void every parity claim."""
    pkg = os.path.join(ROOT, "drjit_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text.replace("oracle-backed", "").replace("oracle/", "ORACLE_DOC/") or \
                       f == "dist.py", (f, "mentions the oracle")
                assert "import oracle" not in text and "from oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read()


def test_cxx_adapter_compiles_and_links(tmp_path):
    """include/drjit_b200_thread_state.h: CUDAThreadState-shaped adapter over the C ABI"""
    src = tmp_path / "adapter.cpp"
    src.write_text(r'''
#include "drjit_b200_thread_state.h"
#include <cstring>
int main() {
    drjit_b200::ThreadState ts(nullptr);
    uint32_t buf[8] = {};
    // argument validation happens before any CUDA call: must throw like jitc_raise()
    try { ts.block_reduce(VarType::UInt32, ReduceOp::Add, 8, 0, buf, buf); return 1; }
    catch (const std::runtime_error &e) { if (!strstr(e.what(), "invalid block size")) return 2; }
    try { ts.block_prefix_reduce(VarType::UInt32, ReduceOp::Add, 8, 9, true, false, buf, buf); return 3; }
    catch (const std::runtime_error &e) { if (!strstr(e.what(), "invalid block size")) return 4; }
    try { ts.memset_async(buf, 8, 3, buf); return 5; }
    catch (const std::runtime_error &e) { if (!strstr(e.what(), "invalid element size")) return 6; }
    // size == 0 is a silent no-op / returns 0 (cuda_ts.cpp:200,685,790)
    ts.block_reduce(VarType::Float32, ReduceOp::Add, 0, 1, buf, buf);
    if (ts.compress((const uint8_t *) buf, 0, buf) != 0) return 7;
    if (ts.block_mkperm(buf, 0, 1, 4, buf, nullptr) != 0) return 8;
    // scatter forms: an odd packet is rejected like cuda_packet.cpp:184-186; empty inputs are no-ops
    const void *comps[3] = { buf, buf, buf };
    try { ts.scatter_reduce_packet(VarType::Float32, ReduceOp::Add, ReduceMode::Auto, buf, 1, comps, 3, buf, nullptr, 1); return 9; }
    catch (const std::runtime_error &e) { if (!strstr(e.what(), "not supported by reduction")) return 10; }
    ts.scatter_reduce_packet(VarType::Float32, ReduceOp::Add, ReduceMode::Auto, buf, 1, comps, 2, buf, nullptr, 0);
    ts.scatter_inc(buf, 1, nullptr, nullptr, 0, buf);
    const void *hole[2] = { buf, nullptr };
    try { ts.scatter_reduce_packet(VarType::Float32, ReduceOp::Add, ReduceMode::Auto, buf, 1, hole, 2, buf, nullptr, 4); return 11; }
    catch (const std::runtime_error &e) { if (!strstr(e.what(), "has no data")) return 12; }
    try { ts.scatter_inc(buf, 1, nullptr, nullptr, 4, nullptr); return 13; }
    catch (const std::runtime_error &e) { if (!strstr(e.what(), "null target")) return 14; }
    static_assert(sizeof(AggregationEntry) == 16, "layout");
    static_assert(sizeof(drjit_b200_call_bucket) == 16, "a table row is overwritten in place by a CallBucket");
    return 0;
}
''')
    exe = tmp_path / "adapter"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           LIB, f"-Wl,-rpath,{os.path.dirname(LIB)}"])
    rc = subprocess.call([str(exe)])
    assert rc == 0, f"adapter self-test failed with code {rc}"
