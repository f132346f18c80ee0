#!/usr/bin/env python3
"""Static SASS evidence: which instruction classes the hot kernels of the shipped library contain.

Runs `cuobjdump -sass` on drjit_b200/lib/libdrjit_b200.so (no GPU needed) and counts, per hot
kernel, the mnemonics that prove the design points DESIGN.md §4 claims: 128-bit loads (LDG.E.128 /
LDS.128), TMA bulk copies (UBLKCP) and L2 bulk prefetch (UBLKPF), mbarriers (SYNCS), redux.sync
(REDUX), warp shuffles / votes, shared atomics, global reductions (RED).

    python scripts/sass_evidence.py > profiles/sass_evidence.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drjit_b200", "lib", "libdrjit_b200.so")

# demangled-name prefixes of the kernels the bench and the BASELINE configs run
HOT = [
    "void djb::prefix_reduce_kernel<unsigned int, djb::OpAdd, false, true, 8u, 3u, 2u>",
    "void djb::compress_kernel<8u, 1u, 3u, 3u, true, false>",
    "void djb::compress_kernel<8u, 1u, 3u, 3u, true, true>",
    "void djb::mkperm_tile_scatter_kernel<1024u, 48u, false>",
    "void djb::mkperm_tile_scatter_kernel<1024u, 24u, true>",
    "void djb::mkperm_tile_scatter_stable_kernel<1024u, 8u>",
    "void djb::mkperm_tile_hist_kernel<1024u, 48u>",
    "void djb::block_reduce_chunk_kernel<float, djb::OpAdd, false, true, false>",
    "void djb::block_reduce_chunk_kernel<float, djb::OpAdd, false, true, true>",
    "void djb::prefix_group_blocks_kernel<float, djb::OpAdd, 32u, false>",
    "void djb::block_reduce_group_kernel<float, djb::OpAdd, true, false>",
    "void djb::scatter_reduce_kernel<float, djb::OpAdd, false>",
    "void djb::scatter_packet_vec_kernel<float, djb::OpAdd, 4u, 4u>",
    "void djb::scatter_packet_kernel<__half, djb::OpAdd, 8u, false>",
    "djb::scatter_inc_queue_kernel",
    "djb::scatter_inc_private_kernel",
]

CLASSES = [
    ("LDG.E.*STRONG.GPU (descriptors)", re.compile(r"\bLDG\.E\.\S*STRONG\.GPU")),
    ("LDG.E.*128 (128-bit global loads)", re.compile(r"\bLDG\.E\.(?:(?!STRONG)\S)*128")),
    ("LDS.128", re.compile(r"\bLDS\.128")),
    ("STG.E.*128 (128-bit global stores)", re.compile(r"\bSTG\.E\.\S*128")),
    ("ATOMS (shared atomics)", re.compile(r"\bATOMS")),
    ("RED/REDG (global reductions)", re.compile(r"\bREDG?\.E")),
    ("REDG vector forms (F32x2/x4, F16x4/x8)", re.compile(r"\bREDG\.E\.\w+\.F(32x[24]|16x[48])")),
    ("MATCH (match.any / match.all)", re.compile(r"\bMATCH\.")),
    ("REDUX (redux.sync)", re.compile(r"\bC?REDUX")),
    ("SHFL", re.compile(r"\bSHFL")),
    ("VOTE", re.compile(r"\bVOTEU?\b")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS")),
    ("UBLKCP (cp.async.bulk global->shared, TMA)", re.compile(r"\bUBLKCP")),
    ("UBLKPF (cp.async.bulk.prefetch.L2)", re.compile(r"\bUBLKPF")),
]


def main():
    if not os.path.exists(LIB):
        sys.exit(f"{LIB} missing: run `make -C drjit_b200/csrc` first")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    mangled = re.findall(r"Function : (\S+)", sass)
    demangled = subprocess.run(["c++filt"], input="\n".join(mangled), check=True, capture_output=True,
                               text=True).stdout.splitlines()
    names = dict(zip(mangled, demangled))
    bodies = re.split(r"\s*Function : (\S+)", sass)[1:]
    counts = {}
    for m, body in zip(bodies[0::2], bodies[1::2]):
        counts[names[m]] = {label: len(rx.findall(body)) for label, rx in CLASSES}

    print("# SASS evidence (cuobjdump -sass drjit_b200/lib/libdrjit_b200.so, sm_100a): instruction classes per hot kernel")
    print("# static instruction counts; produced by scripts/sass_evidence.py")
    print()
    for prefix in HOT:
        hits = [n for n in counts if n.startswith(prefix)]
        if not hits:
            print(f"{prefix}  -- NOT FOUND in the library")
            continue
        name = sorted(hits)[0]
        print(name[:150])
        for label, _ in CLASSES:
            c = counts[name][label]
            if c:
                print(f"    {label:<45} {c}")


if __name__ == "__main__":
    main()
