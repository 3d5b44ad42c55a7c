"""ctypes binding of libsdeo.so (include/sdeo.h). No fallback: a missing library is a hard error."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# SDEO_LIB: load another build of the same ABI (A/B comparisons of kernel changes on one GPU box)
LIB_PATH = os.environ.get("SDEO_LIB") or os.path.join(HERE, "_C", "libsdeo.so")

SDEO_EPI_NORMAL, SDEO_EPI_GEGLU, SDEO_EPI_QKV = 0, 1, 2
SDEO_ACT_NONE, SDEO_ACT_SILU, SDEO_ACT_QUICK_GELU = 0, 1, 2


class ConvArgs(Structure):
    """Mirror of `struct sdeo_conv_args` (include/sdeo.h)."""
    _fields_ = [
        ("x1", c_void_p), ("x2", c_void_p),
        ("n", c_int32), ("h", c_int32), ("w", c_int32),
        ("c1", c_int32), ("ld1", c_int32), ("c2", c_int32), ("ld2", c_int32),
        ("w_packed", c_void_p),
        ("cout", c_int32), ("ksize", c_int32), ("stride", c_int32), ("pad", c_int32),
        ("epi_mode", c_int32), ("act", c_int32),
        ("bias", c_void_p), ("emb", c_void_p), ("emb_step", c_void_p), ("residual", c_void_p), ("ldr", c_int32), ("residual_f32", c_int32),
        ("scale", c_float),
        ("y", c_void_p), ("ldy", c_int32), ("y_fp32", c_int32), ("y2", c_void_p), ("ldy2", c_int32),
        ("q", c_void_p), ("k", c_void_p), ("vt", c_void_p),
        ("heads", c_int32), ("dhead", c_int32), ("tokens", c_int32), ("ldv", c_int32), ("qkv_first", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
        ("gn_stats", c_void_p),
        ("row_stats", c_void_p), ("row_stats_ld", c_int32),
        ("ln_stats", c_void_p), ("ln_parts", c_int32), ("ln_ld", c_int32), ("ln_c", c_int32), ("ln_eps", ctypes.c_float),
        ("ln_csum", c_void_p),
        ("pad_hi", c_int32),
        ("gnf_stats1", c_void_p), ("gnf_stats2", c_void_p), ("gnf_parts1", c_int32), ("gnf_parts2", c_int32),
        ("gnf_gamma", c_void_p), ("gnf_beta", c_void_p), ("gnf_groups", c_int32), ("gnf_eps", ctypes.c_float),
        ("gnf_silu", c_int32),
        ("up2_phase", c_int32),
    ]


# name -> (restype, argtypes); every symbol declared in include/sdeo.h
SIGNATURES = {
    "sdeo_last_error": (c_char_p, []),
    "sdeo_version": (c_int, []),
    "sdeo_set_pdl": (c_int, [c_int]),
    "sdeo_set_trace": (c_int, [c_void_p]),
    "sdeo_conv_workspace_bytes": (c_size_t, [POINTER(ConvArgs)]),
    "sdeo_conv_counter_bytes": (c_size_t, []),
    "sdeo_conv2d": (c_int, [POINTER(ConvArgs), c_void_p]),
    "sdeo_conv_autotune": (c_int, [c_int]),
    "sdeo_conv_set_cta_budget": (c_int, [c_int]),
    "sdeo_conv_gn_stats_slots": (c_int, [POINTER(ConvArgs), POINTER(c_int32), POINTER(c_int32)]),
    "sdeo_conv_plan_describe": (c_int, [POINTER(ConvArgs), c_int32, POINTER(c_int32), c_int32]),
    "sdeo_conv_row_stats_parts": (c_int, [POINTER(ConvArgs), POINTER(c_int32), POINTER(c_int32)]),
    "sdeo_groupnorm_apply_stats": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_int32, c_void_p]),
    "sdeo_gn_stats_fold": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, POINTER(c_int32), c_void_p]),
    "sdeo_packed_rows": (c_int32, [c_int32]),
    "sdeo_packed_k": (c_int32, [c_int32, c_int32, c_int32]),
    "sdeo_pick_bn": (c_int32, [c_int32, c_int32, c_int32]),
    "sdeo_pack_conv_weight": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "sdeo_pack_geglu_bias": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "sdeo_groupnorm_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "sdeo_groupnorm_plan": (c_int, [c_int32, c_int32, c_int64, c_void_p]),
    "sdeo_groupnorm_nhwc": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                    c_int32, c_int32, c_float, c_int32, c_void_p, c_size_t, c_void_p]),
    "sdeo_groupnorm_f16_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "sdeo_groupnorm_f16_plan": (c_int, [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_groupnorm_f16_variant": (c_int, [c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_groupnorm_f16_slab_plan": (c_int, [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_groupnorm_f16_visits": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_int32]),
    "sdeo_groupnorm_nhwc_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float,
                                        c_int32, c_void_p, c_size_t, c_void_p]),
    "sdeo_layernorm": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "sdeo_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                               c_int32, c_float, c_void_p]),
    "sdeo_cfg_ddim_step": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                   c_void_p]),
    "sdeo_cfg_ddim_step_noise_table": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                               c_void_p]),
    "sdeo_counter_add": (c_int, [c_void_p, c_int32, c_void_p]),
    "sdeo_nchw_to_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "sdeo_nhwc_bf16_to_nchw": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_nhwc_f32_to_nchw": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_upsample_nearest2x": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_add_scaled": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "sdeo_timestep_embedding": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "sdeo_attention_causal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_float, c_void_p]),
    "sdeo_embedding_add": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                   c_void_p]),
    "sdeo_softmax_rows": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "sdeo_silu": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "sdeo_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "sdeo_bf16_to_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "sdeo_image_to_u8": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "sdeo_memset_async": (c_int, [c_void_p, c_int, c_size_t, c_void_p]),
    "sdeo_canny_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "sdeo_canny_u8": (c_int, [c_void_p, c_int32, c_int32, c_int32, ctypes.c_double, ctypes.c_double, c_void_p, c_void_p,
                              c_size_t, c_void_p]),
    "sdeo_edges_to_hint": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "sdeo_axpby_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "sdeo_mask_blend_table_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                          ctypes.c_int64, c_void_p]),
    "sdeo_mask_blend_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                    c_int32, c_int64, c_void_p]),
    "sdeo_split_terms": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int64, c_int32, c_int32, ctypes.c_uint32,
                                 c_void_p]),
    "sdeo_split_terms_weight": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                        ctypes.c_uint32, c_void_p]),
    "sdeo_groupnorm_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                   c_int32, c_float, c_int32, c_void_p]),
    "sdeo_layernorm_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
    "sdeo_attention_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                   c_int32, c_int32, c_int32, c_int32, c_float, c_void_p]),
    "sdeo_geglu_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "sdeo_silu_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "sdeo_timestep_embedding_f32": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p]),
}

_lib = None


class SdeoError(RuntimeError):
    pass


def load():
    """Returns the loaded library; raises if libsdeo.so has not been built (there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdeoError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. "
            "Run `python -m stablediffusioneo_b200.build` (needs nvcc). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sdeo_last_error()
        raise SdeoError(f"{what} failed with {rc}: {msg.decode() if msg else ''}")
