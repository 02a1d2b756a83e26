/* Embeds the default model (model19_check3, RDW container) in the shared library, the way the reference
 * compiles its weight tables in (src/rade_enc_data.c / src/rade_dec_data.c). */
    .section .rodata
    .global rade_b200_default_weights
    .global rade_b200_default_weights_end
    .balign 64
rade_b200_default_weights:
    .incbin RADE_WEIGHTS_FILE
rade_b200_default_weights_end:
    .byte 0
    .section .note.GNU-stack,"",@progbits
