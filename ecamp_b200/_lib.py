"""ctypes binding of libecamp_b200.so (the C ABI declared in include/ecamp_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ECAMP_B200_LIB: load another build of the SAME library (A/B measurements of kernel variants); never a fallback
LIB_PATH = os.environ.get("ECAMP_B200_LIB") or os.path.join(_HERE, "lib", "libecamp_b200.so")

GEMM_GELU, GEMM_DGELU, GEMM_DROPOUT, GEMM_AUX_GRAD = 1, 2, 4, 8


class Epilogue(ctypes.Structure):
    _fields_ = [
        ("bias", ctypes.c_void_p),
        ("aux_in", ctypes.c_void_p),
        ("aux_out", ctypes.c_void_p),
        ("ld_aux", ctypes.c_int32),
        ("residual", ctypes.c_void_p),
        ("ld_res", ctypes.c_int32),
        ("out_f32", ctypes.c_void_p),
        ("ld_f32", ctypes.c_int32),
        ("out_bf16", ctypes.c_void_p),
        ("ld_bf16", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("drop_p", ctypes.c_float),
        ("seed", ctypes.c_uint64),
        ("site", ctypes.c_uint64),
        ("colsum_out", ctypes.c_void_p),
    ]


class Shape(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int32), ("T", ctypes.c_int32), ("len_keep", ctypes.c_int32), ("has_big", ctypes.c_int32),
                ("ce_rows", ctypes.c_int32)]


class Batch(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("image", "ids", "labels", "attention_mask", "type_ids", "weights",
                                               "column", "row", "noise")]


class ClsIO(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("image", "pos_embed", "fc_norm_w", "fc_norm_b", "head_w16", "head_b", "dp_scale",
                                               "logits", "d_logits", "g_pos_embed", "g_fc_norm_w", "g_fc_norm_b", "g_head_w",
                                               "g_head_b")]


class Attn(ctypes.Structure):
    _fields_ = [("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("ldq", ctypes.c_int32), ("ldk", ctypes.c_int32), ("ldv", ctypes.c_int32),
                ("o", ctypes.c_void_p), ("ldo", ctypes.c_int32), ("lse", ctypes.c_void_p),
                ("key_mask", ctypes.c_void_p),
                ("B", ctypes.c_int32), ("H", ctypes.c_int32), ("Sq", ctypes.c_int32), ("Sk", ctypes.c_int32),
                ("D", ctypes.c_int32), ("scale", ctypes.c_float), ("drop_p", ctypes.c_float),
                ("seed", ctypes.c_uint64), ("site", ctypes.c_uint64),
                ("d_o", ctypes.c_void_p), ("ld_do", ctypes.c_int32), ("delta", ctypes.c_void_p),
                ("dq", ctypes.c_void_p), ("dk", ctypes.c_void_p), ("dv", ctypes.c_void_p),
                ("lddq", ctypes.c_int32), ("lddk", ctypes.c_int32), ("lddv", ctypes.c_int32),
                ("cs_q", ctypes.c_void_p), ("cs_k", ctypes.c_void_p), ("cs_v", ctypes.c_void_p)]


class SgdTensor(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("buf", ctypes.c_void_p), ("numel", ctypes.c_int64)]


# every symbol include/ecamp_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ecamp_abi_version", "ecamp_last_error", "ecamp_launch_count", "ecamp_gemm_bf16", "ecamp_gemm_fp32", "ecamp_gemm_fp32_ws_bytes", "ecamp_gemm_set_cta_pair", "ecamp_gemm_set_tma_epilogue", "ecamp_gemm_set_direct_epilogue", "ecamp_random_masking", "ecamp_resize_patchify",
    "ecamp_layernorm_fwd", "ecamp_layernorm_bwd", "ecamp_layernorm_ws_floats", "ecamp_layernorm_set_bwd_slab", "ecamp_ce_rows_bias", "ecamp_ce_set_fused", "ecamp_sr_set_window_skip", "ecamp_set_side_stream", "ecamp_attention_set_tcgen05", "ecamp_attention_fwd",
    "ecamp_attention_bwd", "ecamp_mim_loss", "ecamp_sr_loss_fwd", "ecamp_sr_loss_bwd", "ecamp_sr_ws_floats",
    "ecamp_pred_grad", "ecamp_ce_rows", "ecamp_param_count", "ecamp_param_name", "ecamp_param_numel",
    "ecamp_param_decay", "ecamp_param_grad_offset", "ecamp_grad_floats", "ecamp_shadow_bytes",
    "ecamp_adam_table_bytes", "ecamp_adam_chunk_bytes", "ecamp_ctx_create", "ecamp_ctx_destroy", "ecamp_ctx_bind",
    "ecamp_workspace_bytes", "ecamp_ctx_set_precision", "ecamp_ctx_workspace_bytes", "ecamp_ctx_set_workspace", "ecamp_refresh_shadows", "ecamp_forward",
    "ecamp_backward_stage_count", "ecamp_backward_stage_range", "ecamp_backward", "ecamp_backward_stages", "ecamp_adamw_step", "ecamp_adamw_step_groups", "ecamp_adamw_step_range",
    "ecamp_debug_buffer", "ecamp_cls_workspace_bytes", "ecamp_cls_set_workspace", "ecamp_cls_forward", "ecamp_cls_backward",
    "ecamp_sgd_table_bytes", "ecamp_sgd_chunk_bytes", "ecamp_sgd_build_tables", "ecamp_grad_sumsq", "ecamp_sgd_momentum_step",
    "ecamp_attention_probs", "ecamp_cross_attention_probs", "ecamp_image_u8_normalize", "ecamp_image_resample_kmax", "ecamp_image_resized_crop_ws_bytes", "ecamp_image_resized_crop",
    "ecamp_image_resized_crop_host",
    "ecamp_text_mask_draw_count", "ecamp_text_context_mask", "ecamp_text_template_weights", "ecamp_text_mask_and_weights",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"ecamp_b200: {LIB_PATH} is missing - build it with ecamp_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no CPU or PyTorch fallback")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ecamp_last_error.restype = ctypes.c_char_p
        _lib.ecamp_abi_version.restype = ctypes.c_int
        _lib.ecamp_param_name.restype = ctypes.c_char_p
        for f in ("ecamp_param_numel", "ecamp_param_grad_offset", "ecamp_grad_floats", "ecamp_shadow_bytes",
                  "ecamp_adam_table_bytes", "ecamp_adam_chunk_bytes", "ecamp_workspace_bytes", "ecamp_ctx_workspace_bytes", "ecamp_launch_count", "ecamp_gemm_fp32_ws_bytes", "ecamp_image_resized_crop_ws_bytes",
                  "ecamp_cls_workspace_bytes", "ecamp_sgd_table_bytes", "ecamp_sgd_chunk_bytes"):
            getattr(_lib, f).restype = ctypes.c_int64
        for f in ("ecamp_layernorm_ws_floats", "ecamp_sr_ws_floats"):
            getattr(_lib, f).restype = ctypes.c_size_t
        _lib.ecamp_debug_buffer.restype = ctypes.c_void_p
        _lib.ecamp_ctx_destroy.restype = None
        _lib.ecamp_ctx_destroy.argtypes = [ctypes.c_void_p]
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().ecamp_last_error()
        raise RuntimeError(f"ecamp_b200: {what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def cur_stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a, b, *, a_mn=False, b_mn=False, M=None, N=None, K=None, bias=None, aux_in=None, aux_out=None,
         residual=None, out_f32=None, out_bf16=None, flags=0, drop_p=0.0, seed=0, site=0, tile_n=0, colsum_out=None):
    """D = epilogue(A . B^T) on the tcgen05 kernel.  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(-1) == 1 and b.stride(-1) == 1
    if M is None:
        M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    ep = Epilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.aux_in = aux_in.data_ptr() if aux_in is not None else None
    ep.aux_out = aux_out.data_ptr() if aux_out is not None else None
    aux = aux_in if aux_in is not None else aux_out
    ep.ld_aux = aux.stride(0) if aux is not None else 0
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.ld_res = residual.stride(0) if residual is not None else 0
    ep.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    ep.ld_f32 = out_f32.stride(0) if out_f32 is not None else 0
    ep.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
    ep.ld_bf16 = out_bf16.stride(0) if out_bf16 is not None else 0
    ep.flags, ep.drop_p, ep.seed, ep.site = flags, drop_p, seed, site
    ep.colsum_out = colsum_out.data_ptr() if colsum_out is not None else None
    rc = lib().ecamp_gemm_bf16(ptr(a), ctypes.c_int32(a.stride(0)), ctypes.c_int32(int(a_mn)), ptr(b),
                               ctypes.c_int32(b.stride(0)), ctypes.c_int32(int(b_mn)), ctypes.c_int32(M),
                               ctypes.c_int32(N), ctypes.c_int32(K), ctypes.byref(ep), ctypes.c_int32(tile_n),
                               cur_stream())
    check(rc, "ecamp_gemm_bf16")


def gemm_fp32(a, b, *, a_mn=False, b_mn=False, bias=None, aux_in=None, aux_out=None, residual=None, out_f32=None, out_act=None,
              flags=0, drop_p=0.0, seed=0, site=0, colsum_out=None):
    """fp32-accurate GEMM (bf16 x 3 split operands on the tcgen05 kernel).  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N])."""
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.stride(-1) == 1 and b.stride(-1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N = b.shape[1] if b_mn else b.shape[0]
    ep = Epilogue()
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.aux_in = aux_in.data_ptr() if aux_in is not None else None
    ep.aux_out = aux_out.data_ptr() if aux_out is not None else None
    aux = aux_in if aux_in is not None else aux_out
    ep.ld_aux = aux.stride(0) if aux is not None else 0
    ep.residual = residual.data_ptr() if residual is not None else None
    ep.ld_res = residual.stride(0) if residual is not None else 0
    ep.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    ep.ld_f32 = out_f32.stride(0) if out_f32 is not None else 0
    ep.out_bf16 = out_act.data_ptr() if out_act is not None else None
    ep.ld_bf16 = out_act.stride(0) if out_act is not None else 0
    ep.flags, ep.drop_p, ep.seed, ep.site = flags, drop_p, seed, site
    ep.colsum_out = colsum_out.data_ptr() if colsum_out is not None else None
    need = lib().ecamp_gemm_fp32_ws_bytes(M, N, K)
    ws = torch.empty(need, dtype=torch.uint8, device=a.device)
    rc = lib().ecamp_gemm_fp32(ptr(a), ctypes.c_int32(a.stride(0)), ctypes.c_int32(int(a_mn)), ptr(b), ctypes.c_int32(b.stride(0)),
                               ctypes.c_int32(int(b_mn)), ctypes.c_int32(M), ctypes.c_int32(N), ctypes.c_int32(K), ctypes.byref(ep),
                               ptr(ws), ctypes.c_int64(need), cur_stream())
    check(rc, "ecamp_gemm_fp32")
