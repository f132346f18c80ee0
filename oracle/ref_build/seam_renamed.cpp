/*
 * seam_renamed.cpp -- compiles the reference's UNMODIFIED src/cuda_ts.cpp (included from where it
 * lies, nothing copied) with the eight primitive methods of CUDAThreadState renamed out of the way,
 * so that seam_b200.cpp can define them on top of libdrjit_b200.so while everything else in that
 * file (launch, memcpy, batched_gemm, coop_vec_pack, enqueue_host_func, ...) stays the reference's.
 *
 * TEST INFRASTRUCTURE (oracle/ref_build: the in-situ proof of the drop-in boundary).
 * A maintainer would simply replace the eight bodies in src/cuda_ts.cpp by those of seam_b200.cpp;
 * the renaming exists only because this repository may not carry a modified copy of that file.
 *
 * How: the headers cuda_ts.cpp needs are included first (their include guards make the second
 * inclusion a no-op), then the method names are #defined to unused names while cuda_ts.h and
 * cuda_ts.cpp are read. In this one translation unit CUDAThreadState therefore declares
 * `ref_unused_block_reduce(...)` etc. as plain members (`override` is defined away, since they no
 * longer override anything); the vtable and every other translation unit use the real declaration
 * of cuda_ts.h, and the real symbols come from seam_b200.cpp.
 */
#include "internal.h"
#include "log.h"
#include "util.h"
#include "var.h"
#include "optix.h"
#include "eval.h"

#define override
#define memset_async        ref_unused_memset_async
#define block_reduce        ref_unused_block_reduce
#define reduce_dot          ref_unused_reduce_dot
#define block_prefix_reduce ref_unused_block_prefix_reduce
#define compress            ref_unused_compress
#define block_mkperm        ref_unused_block_mkperm
#define poke                ref_unused_poke
#define aggregate           ref_unused_aggregate

#include "cuda_ts.cpp"      /* found through -I$(REF)/src */
