"""CPU-only static check of the shipped sm_100a code: the hot kernels contain the instruction
classes DESIGN.md section 4 claims (TMA bulk copies + mbarriers in scan / compress, redux.sync,
128-bit accesses, shared atomics + L2 bulk prefetch in mkperm, native global reductions incl. the
two-wide f16 forms in scatter_reduce). Full listing: scripts/sass_evidence.py -> profiles/sass_evidence.txt."""
import functools
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drjit_b200", "lib", "libdrjit_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")

SCAN = "_ZN3djb20prefix_reduce_kernelIjNS_5OpAddELb0ELb1ELj8ELj3ELj2EEEvNS_12PrefixParamsE"
COMPRESS = "_ZN3djb15compress_kernelILj8ELj1ELj3ELj3ELb1ELb0EEEvNS_14CompressParamsE"
SUM = "_ZN3djb25block_reduce_chunk_kernelIfNS_5OpAddELb0ELb1ELb0EEEvPKT_S4_PS2_PNS_3AccIS2_E4typeEPjjjjjNS_7PeerCtxEj"
SUM_PEER = SUM.replace("ELb0ELb1ELb0EEE", "ELb0ELb1ELb1EEE")
MKPERM_SCATTER = "_ZN3djb26mkperm_tile_scatter_kernelILj1024ELj%uELb0EEEvNS_16MkpermTileParamsE"
MKPERM_HIST = "_ZN3djb23mkperm_tile_hist_kernelILj1024ELj48EEEvNS_16MkpermTileParamsE"
MKPERM_STABLE = "_ZN3djb33mkperm_tile_scatter_stable_kernelILj1024ELj8EEEvNS_16MkpermTileParamsE"
SCATTER = "_ZN3djb21scatter_reduce_kernelI%sNS_5Op%sELb0EEEvNS_13ScatterParamsE"


@functools.lru_cache(maxsize=None)
def sass(mangled):
    """SASS of one kernel of the shipped library (cuobjdump warns on stderr for every other cubin)"""
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, LIB], capture_output=True, text=True).stdout
    assert "Function : " + mangled in out, f"kernel {mangled} is not in the library"
    assert "arch = sm_100a" in out
    return out


def test_only_sm_100a_code_is_shipped():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], check=True, capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("kernel", [SCAN, COMPRESS])
def test_scan_and_compress_are_tma_staged(kernel):
    b = sass(kernel)
    assert re.search(r"\bUBLKCP", b), "no cp.async.bulk"
    assert re.search(r"\bSYNCS", b), "no mbarrier"
    assert re.search(r"\bC?REDUX", b), "no redux.sync"
    assert re.search(r"\bLDS\.128", b), "no 128-bit shared loads"


def test_reductions_use_wide_loads_and_shuffles():
    b = sass(SUM)
    assert re.search(r"\bLDG\.E\.\S*128", b)
    assert re.search(r"\bSHFL", b)


@pytest.mark.parametrize("kernel", [MKPERM_SCATTER % 48, MKPERM_SCATTER % 40, MKPERM_SCATTER % 32, MKPERM_HIST])
def test_mkperm_tile_kernels(kernel):
    b = sass(kernel)
    assert re.search(r"\bATOMS", b), "no shared atomics"
    assert re.search(r"\bUBLKPF", b), "no L2 bulk prefetch"
    assert re.search(r"\bLDG\.E\.\S*128", b), "keys are not read with 128-bit loads"


def test_fused_reduction_exchanges_over_peer_memory():
    """The sharded reduction publishes its partial and spins on the flags inside the same kernel:
    system-scope release stores / acquire loads, and no such code in the single-GPU instantiation."""
    b = sass(SUM_PEER)
    assert re.search(r"\bSTG?\.E\.\S*STRONG\.SYS", b), "no system-scope release store"
    assert re.search(r"\bLDG?\.E\.\S*STRONG\.SYS", b), "no system-scope acquire load"
    assert re.search(r"\bMEMBAR\.\S*SYS", b), "no system-scope fence"
    assert not re.search(r"\bMEMBAR\.\S*SYS", sass(SUM))


def test_stable_mkperm_ranks_with_ballots():
    assert len(re.findall(r"\bVOTEU?\b", sass(MKPERM_STABLE))) >= 256


def test_scatter_reduce_uses_native_reductions():
    assert re.search(r"\bREDG?\.E\.ADD\.F32", sass(SCATTER % ("f", "Add")))
    for op in ("Add", "Min", "Max"):
        b = sass(SCATTER % ("6__half", op))
        assert re.search(rf"\bREDG?\.E\.{op.upper()}\.F16x2", b), f"f16 {op}: not a two-wide f16 reduction"
        assert not re.search(r"\bATOMG?\.E\.CAS", b), f"f16 {op}: compare-and-swap loop"


def test_fused_compress_exchanges_inside_the_kernel():
    """The PEER instantiation of the compaction publishes the shard's count to every rank from inside
    the kernel (system-scope stores and loads of the scalar cells); the single-GPU one contains none."""
    peer = COMPRESS.replace("ELb0EEEvNS_", "ELb1EEEvNS_")
    assert peer != COMPRESS
    b = sass(peer)
    assert re.search(r"\bSTG?\.E\.\S*STRONG\.SYS", b), "no system-scope store"
    assert re.search(r"\bLDG?\.E\.\S*STRONG\.SYS", b), "no system-scope load"
    assert not re.search(r"\bSTG?\.E\.\S*STRONG\.SYS", sass(COMPRESS))


PACKET = "_ZN3djb21scatter_packet_kernelI%sNS_5Op%sELj%uELb%uEEEvNS_12PacketParamsE"


def test_packet_scatter_uses_vector_reductions():
    """scatter_packet.cu: a whole packet (or 16 bytes of it) per reduction instruction
    (red.global.v4.f32.add / red.global.v8.f16.<op>.noftz, cuda_packet.cpp:224-259)"""
    assert re.search(r"\bREDG\.E\.ADD\.F32x4", sass(PACKET % ("f", "Add", 4, 0)))
    assert re.search(r"\bREDG\.E\.ADD\.F32x2", sass(PACKET % ("f", "Add", 2, 0)))
    assert re.search(r"\bREDG\.E\.MAX\.F16x8", sass(PACKET % ("6__half", "Max", 8, 0)))
    assert re.search(r"\bREDG\.E\.ADD\.F16x4", sass(PACKET % ("6__half", "Add", 4, 0)))
    local = sass(PACKET % ("f", "Add", 4, 1))
    assert re.search(r"\bMATCH\.ANY", local) and re.search(r"\bREDG\.E\.ADD\.F32x4", local)


def test_scatter_inc_aggregates_per_cta():
    """small counter arrays: shared-memory atomics with return + one global atomic per touched counter
    and tile; coherent warps are detected with one vote (match.all)"""
    b = sass("_ZN3djb26scatter_inc_private_kernelENS_9IncParamsE")
    assert re.search(r"\bATOMS\.ADD", b) and re.search(r"\bMATCH\.ALL", b) and re.search(r"\bATOMG\.E\.ADD", b)
    g = sass("_ZN3djb18scatter_inc_kernelENS_9IncParamsE")
    assert re.search(r"\bMATCH\.ANY", g) and re.search(r"\bATOMG\.E\.ADD", g)
