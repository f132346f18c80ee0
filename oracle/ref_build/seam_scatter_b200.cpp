/*
 * seam_scatter_b200.cpp -- SURVEY.md section 8, row f3: the scatter-reduce PTX template of the
 * reference's JIT (ext/drjit-core/src/cuda_scatter.cpp:246-354, warp pre-reduction :125-244) with a
 * Blackwell-era fast path for the case that dominates counters and integer histograms inside fused
 * kernels: ReduceMode::Local on 32-bit integers.
 *
 * The reference finds the lanes of a warp that hit the same address with match.any and then reduces
 * their values with five shuffle + op steps -- but only when ALL 32 lanes are active and agree
 * (`peers == -1`); a warp with masked-off lanes or a tail warp takes the general path, a
 * data-dependent loop of shuffles, ballots and bit tricks (~14 instructions per round, 5 rounds for
 * a warp-wide group), even when every active lane targets the same counter. That case -- counters,
 * loss accumulators, `scatter_reduce(op, target, value, 0, active)` -- is what this template makes
 * cheap: when all ACTIVE lanes agree, sm_80+ hardware does the whole reduction in one instruction:
 *
 *     match.all.sync.b32  peers|coherent, key, active   // do all active lanes hit one address?
 *     redux.sync.<op>.<t> acc, value, active            // one instruction, any set of active lanes
 *     @leader red.global.<op>.<t> [addr], acc           // lowest active lane: one atomic
 *
 * for op in {add, min, max, and, or}, t in {s32, u32, b32}. Warps whose lanes disagree fall through
 * to the reference's own generator, which is emitted behind the fast path. Two things measured on
 * the B200 shaped this (profiles/r4j_scatter_template_ab*.txt): redux.sync with k disjoint group
 * masks costs ~25 cycles per mask, so it only pays for ONE group (a first version that used it for
 * every group was 2.4x slower at 2^20 random bins); and match.any is the bottleneck of the Local mode
 * on this chip (~64 cycles per warp: 107 G elements/s at 2^20 bins against 189 G/s for plain REDs),
 * so the coherence test must not be a second match.any -- match.all is a single cheap vote. Everything else -- floats, 64-bit types, Direct /
 * NoConflicts modes, f16 pairs, the packet path of cuda_packet.cpp -- goes to the reference's
 * generator unchanged.
 *
 * Integration (oracle/ref_build/Makefile, `make b200`): the unmodified cuda_scatter.o is linked twice,
 * once with jitc_cuda_render_scatter_reduce weakened (so this definition wins for every caller) and
 * once with that symbol renamed to ref_... and all its other globals localized (so the original stays
 * callable from here). A maintainer would simply add the fast path at the top of the `mode ==
 * ReduceMode::Local` branch (cuda_scatter.cpp:279). Exercised on the B200 through the reference's
 * tracer by tests/test_insitu_gpu.py (jit_var_scatter with ReduceMode::Local against the oracle; the
 * kernel IR recorded by JitFlag::KernelHistory must contain redux.sync).
 */
#include "eval.h"
#include "var.h"
#include "op.h"
#include "log.h"
#include "cuda_eval.h"
#include "cuda_scatter.h"

extern void jitc_cuda_prepare_index(const Variable *ptr, const Variable *index, const Variable *value);

/// the reference's generator (second, renamed copy of cuda_scatter.o)
extern void ref_jitc_cuda_render_scatter_reduce(const Variable *v, const Variable *ptr, const Variable *value,
                                                const Variable *index, const Variable *mask)
    asm("ref__Z31jitc_cuda_render_scatter_reducePK8VariableS1_S1_S1_S1_");

void jitc_cuda_render_scatter_reduce(const Variable *v, const Variable *ptr, const Variable *value,
                                     const Variable *index, const Variable *mask) {
    const ReduceOp op = (ReduceOp) (uint32_t) v->literal;
    const ReduceMode mode = (ReduceMode) (uint32_t) (v->literal >> 32);
    const VarType vt = (VarType) value->type;
    const ThreadState *ts = thread_state_cuda;

    const bool int32 = vt == VarType::Int32 || vt == VarType::UInt32;
    const bool op_ok = op == ReduceOp::Add || op == ReduceOp::Min || op == ReduceOp::Max ||
                       op == ReduceOp::And || op == ReduceOp::Or;
    // redux.sync: PTX ISA 7.0, sm_80 and later
    if (!(mode == ReduceMode::Local && int32 && op_ok && ts->ptx_version >= 70 && ts->compute_capability >= 80)) {
        ref_jitc_cuda_render_scatter_reduce(v, ptr, value, index, mask);
        return;
    }

    const bool is_unmasked = mask->is_literal() && mask->literal == 1;
    const uint32_t uid = v->reg_index;
    if (!is_unmasked)
        fmt("    @!$v bra l_$u_b200_done;\n", mask, uid);

    jitc_cuda_prepare_index(ptr, index, value);      // address of the target element -> %rd3

    const char *name, *type;
    switch (op) {
        case ReduceOp::Add: name = "add"; type = "u32"; break;          // (wraps identically for s32)
        case ReduceOp::Min: name = "min"; type = vt == VarType::Int32 ? "s32" : "u32"; break;
        case ReduceOp::Max: name = "max"; type = vt == VarType::Int32 ? "s32" : "u32"; break;
        case ReduceOp::And: name = "and"; type = "b32"; break;
        default:            name = "or";  type = "b32"; break;
    }

    fmt("    {\n"
        "        .reg .b32 %b2_active, %b2_key, %b2_peers, %b2_below, %b2_acc;\n"
        "        .reg .b64 %b2_word;\n"
        "        .reg .pred %b2_leader, %b2_coherent;\n"
        "        activemask.b32 %b2_active;\n"
        "        shr.b64 %b2_word, %rd3, 2;\n"
        "        cvt.u32.u64 %b2_key, %b2_word;\n"
        "        match.all.sync.b32 %b2_peers|%b2_coherent, %b2_key, %b2_active;\n"
        "        @!%b2_coherent bra l_$u_b200_general;\n"
        "        redux.sync.$s.$s %b2_acc, $v, %b2_active;\n"
        "        mov.u32 %b2_below, %lanemask_lt;\n"
        "        and.b32 %b2_below, %b2_below, %b2_active;\n"
        "        setp.eq.u32 %b2_leader, %b2_below, 0;\n"
        "        @%b2_leader red.global.$s.$s [%rd3], %b2_acc;\n"
        "        bra l_$u_b200_done2;\n"
        "    }\n"
        "l_$u_b200_general:\n",
        uid, name, type, value, name, type, uid, uid);

    // lanes of this warp disagree: the reference's generator (its own mask test is a no-op here)
    ref_jitc_cuda_render_scatter_reduce(v, ptr, value, index, mask);

    fmt("l_$u_b200_done2:\n", uid);
    if (!is_unmasked)
        fmt("\nl_$u_b200_done:\n", uid);
}
