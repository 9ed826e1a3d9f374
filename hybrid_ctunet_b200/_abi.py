"""argtypes/restype declarations for the non-GEMM entry points of include/ctunet_b200.h."""
import ctypes as C

P, I, L, F, D = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double

SIGS = {
    "ctu_ffn_fused": (P, L, P, P, P, P, P, L, P, L, L, I, I, P),
    "ctu_in_stats": (P, I, I, L, I, P, I, P),
    "ctu_in_apply": (P, I, P, I, P, I, P, I, P, I, I, L, I, F, I, F, P),
    "ctu_layernorm": (P, I, L, P, P, P, L, P, I, L, L, I, F, P),
    "ctu_patchify_ln": (P, I, I, I, I, I, P, P, P, F, P),
    "ctu_pwa_fuse": (P, P, P, L, I, I, P),
    "ctu_subsample": (P, I, I, I, I, P, I, I, I, I, I, I, P),
    "ctu_attention": (P, I, I, I, P, I, P, I, I, I, I, I, I, I, I, P, P),
    "ctu_conv_cin1": (P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, P),
    "ctu_blend_accumulate": (P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, P),
    "ctu_blend_count": (P, P, I, I, I, I, I, I, I, I, I, P),
    "ctu_blend_normalize": (P, P, P, I, L, P),
    # backward pass
    "ctu_in_bwd_stats": (P, I, P, I, P, I, P, I, P, I, P, I, I, L, I, F, I, F, P, P),
    "ctu_in_bwd_apply": (P, I, P, I, P, I, P, I, P, I, P, I, I, I, L, I, F, I, F, P, P, I, P, I, P),
    "ctu_layernorm_bwd": (P, I, L, P, P, L, P, I, L, P, L, P, L, P, P, L, I, F, P),
    "ctu_gelu": (P, P, L, P),
    "ctu_gelu_bwd": (P, P, P, L, P),
    "ctu_pwa_fuse_bwd": (P, P, P, P, P, L, I, I, P),
    "ctu_colsum": (P, I, L, L, L, P, P),
    "ctu_cf_to_cl": (P, P, I, I, L, I, I, P),
    "ctu_head_bwd": (P, P, L, P, P, L, I, P, I, P, I, L, I, I, P),
    "ctu_space_to_depth": (P, I, P, I, I, I, I, I, I, I, I, P),
    "ctu_subsample_bwd": (P, I, P, I, I, I, I, I, I, I, I, I, I, P),
    "ctu_im2col_cin1": (P, P, I, I, I, I, I, I, I, I, I, I, I, I, I, I, P),
    "ctu_accumulate": (P, I, L, P, I, L, L, I, P),
    "ctu_cast_f32_bf16": (P, L, P, L, L, I, P),
    "ctu_patchify_ln_bwd": (P, I, I, I, I, I, P, P, P, F, P),
    "ctu_ensemble_argmax": (P, P, I, L, P, P, P, P, P, P),
    "ctu_cin1_k1_stats": (P, P, I, L, I, P, P, I, P),
    "ctu_cc_filter_largest": (P, P, I, I, I, D, I, D, P, P, P, P),
    "ctu_invert_resample": (P, I, P, P, P),
    "ctu_invert_ensemble_argmax": (P, P, I, P, P, P, P, P, P, P),
    "ctu_adamw_step": (P, I, L, D, D, D, D, D, L, P),
    "ctu_pack_weights": (P, I, L, P),
    "ctu_unpack_grads": (P, I, L, P),
    "ctu_stats_fold": (P, I, I, I, I, D, P),
    "ctu_dice_ce_fwd": (P, P, I, I, L, P, P),
    "ctu_dice_ce_bwd": (P, P, I, I, L, P, P, P, P),
    "ctu_dice_ce_finalize": (P, P, P, P, P, P),
    "ctu_gather3d": (P, P, I, I, I, I, I, I, I, P, P, P, P),
    "ctu_attention_delta": (P, L, P, L, P, L, I, I, P),
    "ctu_attention_bwd": (P, I, I, I, P, I, P, P, P, P, I, P, P, I, I, I, I, I, I, I, I, P),
}


def declare(lib):
    for name, argtypes in SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = list(argtypes)
        fn.restype = C.c_int
