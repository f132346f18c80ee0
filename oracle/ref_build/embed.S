/* Embeds the reference's LZ4-packed resources (kernel cache dictionary and the
   precompiled compute_75 PTX) under the symbol names that
   ext/drjit-core/resources/kernels.h:16-21 declares. Paths come from the Makefile. */
    .section .rodata
    .global kernels_dict
    .type kernels_dict, @object
    .balign 64
kernels_dict:
    .incbin KERNELS_DICT
    .global kernels_75
    .type kernels_75, @object
    .balign 64
kernels_75:
    .incbin KERNELS_75
    .section .note.GNU-stack,"",@progbits
