/*
 * Stub LLVM-C shared library (TEST INFRASTRUCTURE ONLY).
 *
 * The reference's LLVM backend dlopens libLLVM and resolves ~60 LLVM-C symbols
 * in jitc_llvm_api_init() (ext/drjit-core/src/llvm_api.cpp:60-165), then calls a
 * handful during jitc_llvm_init() (src/llvm_core.cpp:76-190) and
 * jitc_llvm_orcv2_init() (src/llvm_orcv2.cpp:28-72). The CPU primitives this
 * repository uses as oracle/baseline -- LLVMThreadState::{block_reduce,
 * block_prefix_reduce, reduce_dot, compress, block_mkperm}
 * (src/llvm_ts.cpp:265-933) -- are plain C++ on the nanothread pool and never
 * invoke the JIT. This stub therefore only has to make initialisation succeed
 * on hosts that have no libLLVM.so. Anything that would really compile IR aborts.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

static void stub_die(const char *name) {
    fprintf(stderr, "llvm_stub: %s() called -- the stub libLLVM cannot JIT-compile.\n", name);
    abort();
}

#define NOP(name)  void name(void) { }
#define DIE(name)  void *name(void) { stub_die(#name); return 0; }

/* --- called during init: must behave ---------------------------------- */
NOP(LLVMLinkInMCJIT)
NOP(LLVMInitializeX86AsmPrinter)      NOP(LLVMInitializeX86Disassembler)
NOP(LLVMInitializeX86Target)          NOP(LLVMInitializeX86TargetInfo)
NOP(LLVMInitializeX86TargetMC)
NOP(LLVMInitializeAArch64AsmPrinter)  NOP(LLVMInitializeAArch64Disassembler)
NOP(LLVMInitializeAArch64Target)      NOP(LLVMInitializeAArch64TargetInfo)
NOP(LLVMInitializeAArch64TargetMC)

char *LLVMCreateMessage(const char *s) { return strdup(s); }
void  LLVMDisposeMessage(char *s) { free(s); }
char *LLVMGetDefaultTargetTriple(void) {
#if defined(__aarch64__)
    return strdup("aarch64-unknown-linux-gnu");
#else
    return strdup("x86_64-unknown-linux-gnu");
#endif
}
char *LLVMGetHostCPUName(void) { return strdup("generic"); }
char *LLVMGetHostCPUFeatures(void) {
#if defined(__aarch64__)
    return strdup("+neon,+fma");
#else
    return strdup("+sse4.2,+avx,+avx2,+fma,+f16c");
#endif
}
void *LLVMGetGlobalContext(void) { static int ctx; return &ctx; }
void *LLVMCreateDisasm(const char *t, void *a, int b, void *c, void *d) {
    (void) t; (void) a; (void) b; (void) c; (void) d; return 0;
}
void  LLVMDisasmDispose(void *p) { (void) p; }
int   LLVMSetDisasmOptions(void *p, uint64_t o) { (void) p; (void) o; return 1; }
void  LLVMGetVersion(unsigned *major, unsigned *minor, unsigned *patch) {
    *major = 18; *minor = 0; *patch = 0;
}
int LLVMGetTargetFromTriple(const char *triple, void **target, char **err) {
    static int tgt; (void) triple; *target = &tgt; if (err) *err = 0; return 0;
}
void *LLVMCreateTargetMachine(void *t, const char *a, const char *b, const char *c,
                              int d, int e, int f) {
    static int tm; (void) t; (void) a; (void) b; (void) c; (void) d; (void) e; (void) f;
    return &tm;
}
void  LLVMDisposeTargetMachine(void *p) { (void) p; }
void *LLVMOrcJITTargetMachineBuilderCreateFromTargetMachine(void *tm) { return tm; }
void *LLVMOrcCreateLLJITBuilder(void) { static int b; return &b; }
void  LLVMOrcLLJITBuilderSetJITTargetMachineBuilder(void *a, void *b) { (void) a; (void) b; }
void  LLVMOrcLLJITBuilderSetObjectLinkingLayerCreator(void *a, void *b, void *c) {
    (void) a; (void) b; (void) c;
}
void *LLVMOrcCreateLLJIT(void **out, void *builder) {
    static int jit; (void) builder; *out = &jit; return 0;
}
void *LLVMOrcLLJITGetMainJITDylib(void *j) { return j; }
void *LLVMOrcDisposeLLJIT(void *j) { (void) j; return 0; }
char *LLVMGetErrorMessage(void *e) { (void) e; return strdup("llvm_stub"); }

/* --- only reachable when something tries to JIT: abort loudly ---------- */
DIE(LLVMAddModule)                    DIE(LLVMDisposeModule)
DIE(LLVMCreateMemoryBufferWithMemoryRange) DIE(LLVMParseIRInContext)
DIE(LLVMPrintModuleToString)          DIE(LLVMGetGlobalValueAddress)
DIE(LLVMRemoveModule)                 DIE(LLVMDisasmInstruction)
DIE(LLVMVerifyModule)
DIE(LLVMCreatePassManager)            DIE(LLVMRunPassManager)
DIE(LLVMDisposePassManager)           DIE(LLVMAddLICMPass)
DIE(LLVMCreatePassBuilderOptions)     DIE(LLVMPassBuilderOptionsSetLoopVectorization)
DIE(LLVMPassBuilderOptionsSetLoopUnrolling) DIE(LLVMPassBuilderOptionsSetSLPVectorization)
DIE(LLVMDisposePassBuilderOptions)    DIE(LLVMRunPasses)
DIE(LLVMModuleCreateWithName)         DIE(LLVMGetExecutionEngineTargetMachine)
DIE(LLVMCreateMCJITCompilerForModule) DIE(LLVMCreateSimpleMCJITMemoryManager)
DIE(LLVMDisposeExecutionEngine)       DIE(LLVMGetFunctionAddress)
DIE(LLVMOrcCreateNewThreadSafeContext) DIE(LLVMOrcDisposeThreadSafeContext)
DIE(LLVMOrcCreateNewThreadSafeModule) DIE(LLVMOrcLLJITAddLLVMIRModule)
DIE(LLVMOrcLLJITLookup)
DIE(LLVMOrcCreateRTDyldObjectLinkingLayerWithMCJITMemoryManagerLikeCallbacks)
DIE(LLVMOrcJITDylibClear)
/* version-probe symbols (llvm_api.cpp:205-221) */
NOP(LLVMDisposeErrorMessage) NOP(LLVMCreateBinary) NOP(LLVMBuildFreeze)
NOP(LLVMIsPoison) NOP(LLVMAddMetadataToInst) NOP(LLVMDeleteInstruction)
