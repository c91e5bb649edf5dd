"""ctypes binding of libtinyaudio_b200.so (include/tinyaudio_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used only for device memory and streams; every function takes raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# TA_LIB_VARIANT=pdl selects the A/B build `make PDL=1` leaves next to the default library (csrc/Makefile)
_VARIANT = os.environ.get("TA_LIB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, f"libtinyaudio_b200{'_' + _VARIANT if _VARIANT else ''}.so")

c_void_p, c_int, c_ll, c_float = C.c_void_p, C.c_int, C.c_longlong, C.c_float
P = c_void_p

# epilogue modes (enum in the header)
SKINNY_BF16, SKINNY_F32_RESID, SKINNY_SWIGLU, SKINNY_PARTIAL = range(4)
EPI_BF16, EPI_BF16_GELU, EPI_BF16_RESID, EPI_F32_RESID, EPI_F32, EPI_SWIGLU, EPI_SWIGLU_BWD, EPI_BF16_ROPE, EPI_BF16_ROWDOT = range(9)
ENC_PTRS_PER_LAYER = 12
LM_PTRS_PER_LAYER = 20
LM_LORA_GRADS_PER_LAYER = 8
LM_GRADS_PER_LAYER = 8
(LM_G_WQKV, LM_G_WO, LM_G_WGU, LM_G_WD, LM_G_LN1, LM_G_LN2, LM_G_QNORM, LM_G_KNORM) = range(8)
(ENC_LN1_W, ENC_LN1_B, ENC_WQKV, ENC_BQKV, ENC_WO, ENC_BO, ENC_LN2_W, ENC_LN2_B, ENC_W1, ENC_B1, ENC_W2, ENC_B2) = range(12)
(LM_LN1_W, LM_WQKV, LM_WQKV_T, LM_QNORM_W, LM_KNORM_W, LM_WO, LM_WO_T, LM_LN2_W, LM_WGU, LM_WGU_T, LM_WD, LM_WD_T,
 LM_LORA_A_QKV, LM_LORA_A_O, LM_LORA_A_GU, LM_LORA_A_D, LM_LORA_BT_QKV, LM_LORA_BT_O, LM_LORA_BT_GU, LM_LORA_BT_D) = range(20)


class GemmEpilogue(C.Structure):
    _fields_ = [("out", P), ("ldo", c_ll), ("bias", P), ("resid", P), ("ldr", c_ll), ("out2", P), ("ldo2", c_ll),
                ("aux", P), ("ldaux", c_ll), ("alpha", c_float), ("rope_cos", P), ("rope_sin", P), ("rope_seq", c_int),
                ("rope_cols", c_int), ("zero_f32", P), ("ld_zero", c_ll)]


class EncoderWeights(C.Structure):
    _fields_ = [("n_layers", c_int), ("dim", c_int), ("ffn", c_int), ("heads", c_int), ("head_dim", c_int),
                ("rot_dim", c_int), ("n_mels", c_int), ("max_pos", c_int), ("ln_eps", c_float),
                ("conv1_w", P), ("conv1_b", P), ("conv2_w", P), ("conv2_b", P), ("lnf_w", P), ("lnf_b", P),
                ("rope_cos", P), ("rope_sin", P), ("layers", C.POINTER(P))]


class MlpProjectorWeights(C.Structure):
    _fields_ = [("in_dim", c_int), ("hidden", c_int), ("out_dim", c_int), ("eps", c_float),
                ("w1", P), ("norm_w", P), ("w2", P), ("w2_t", P), ("norm2_w", P)]


class LmWeights(C.Structure):
    _fields_ = [("n_layers", c_int), ("dim", c_int), ("ffn", c_int), ("n_q_heads", c_int), ("n_kv_heads", c_int),
                ("head_dim", c_int), ("max_pos", c_int), ("vocab", c_ll), ("vocab_pad", c_ll), ("eps", c_float), ("lora_pad", c_int),
                ("embed_f32", P), ("embed_bf16", P), ("embed_bf16_t", P), ("final_norm_w", P),
                ("rope_cos", P), ("rope_sin", P), ("layers", C.POINTER(P))]


class LmStepArgs(C.Structure):
    _fields_ = [("B", c_int), ("S", c_int), ("n_labelled", c_int), ("with_backward", c_int),
                ("inputs_embeds", P), ("label_rows", P), ("label_targets", P), ("inv_num_items", c_float),
                ("loss", P), ("row_loss", P), ("d_inputs_embeds", P), ("workspace", P), ("workspace_bytes", c_ll),
                ("final_hidden", P), ("lora_grads", C.POINTER(P)), ("k_cache", P), ("v_cache", P), ("cache_max_seq", c_int),
                ("lm_grads", C.POINTER(P)), ("d_embed", P), ("d_final_norm", P), ("input_ids", P), ("audio_token_id", c_ll),
                ("position_ids", P), ("kv_start", P), ("lora_grads_zeroed", c_int)]


_SIGS = {
    "ta_version": ([], c_int),
    "ta_launch_count": ([], C.c_ulonglong),
    "ta_set_pdl": ([c_int], c_int),
    "ta_gemm_bf16": ([P, c_ll, P, c_ll, c_int, c_int, c_int, c_int, C.POINTER(GemmEpilogue), P], c_int),
    "ta_gemm_bf16_tn": ([P, c_ll, P, c_ll, c_int, c_int, c_int, P, c_ll, c_float, P], c_int),
    "ta_gemm_set_tile_n": ([c_int], c_int),
    "ta_gemm_set_resid_tma": ([c_int], c_int),
    "ta_gemm_set_swiglu_bwd_tma": ([c_int], c_int),
    "ta_gemm_set_cta_pair": ([c_int], c_int),
    "ta_gemm_set_tail_split": ([c_int], c_int),
    "ta_gemm_set_tn_splitk": ([c_int], c_int),
    "ta_logmel_workspace_floats": ([c_int, c_int, C.POINTER(c_ll)], c_int),
    "ta_logmel_fwd": ([P, c_ll, c_int, c_int, P, P, P, P], c_int),
    "ta_mel_to_conv1_im2col": ([P, c_int, c_int, P, P], c_int),
    "ta_attn_fwd": ([P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_int, c_float, P], c_int),
    "ta_attn_set_bwd_variant": ([c_int], c_int),
    "ta_lm_set_fused_attn_dsum": ([c_int], c_int),
    "ta_attn_set_tc_lm": ([c_int], c_int),
    "ta_attn_tc_lm_ring_slots": ([], c_int),
    "ta_attn_set_tc": ([c_int], c_int),
    "ta_attn_set_trace": ([P, c_int], c_int),
    "ta_debug_set": ([c_int, c_int], c_int),
    "ta_attn_bwd": ([P, P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int,
                     c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_ll, c_int, c_float, P], c_int),
    "ta_im2col_k3": ([P, P, c_int, c_int, c_int, c_int, P], c_int),
    "ta_layernorm_set_reverse": ([c_int], c_int),
    "ta_layernorm_bf16": ([P, P, P, P, c_ll, c_int, c_float, P], c_int),
    "ta_rmsnorm_f32": ([P, P, P, P, c_ll, c_int, c_float, P], c_int),
    "ta_rmsnorm_f32_bwd": ([P, P, P, P, P, c_ll, c_int, c_float, c_int, P], c_int),
    "ta_enc_rope": ([P, P, P, c_ll, c_int, c_int, c_int, c_int, P], c_int),
    "ta_lm_qknorm_rope_fwd": ([P, P, P, P, P, P, c_ll, c_int, c_int, c_int, c_float, P], c_int),
    "ta_lm_qknorm_rope_bwd": ([P, P, P, P, P, P, P, P, P, c_ll, c_int, c_int, c_int, c_float, P], c_int),
    "ta_proj_norm_fwd": ([P, P, P, c_ll, c_int, c_float, c_int, P], c_int),
    "ta_proj_norm_bwd": ([P, P, P, c_int, P, P, c_ll, c_int, c_float, c_int, P], c_int),
    "ta_audio_index": ([P, P, P, c_int, c_int, c_int, c_ll, P], c_int),
    "ta_embed_scatter": ([P, P, P, P, P, c_ll, c_int, c_ll, P], c_int),
    "ta_audio_grad_gather": ([P, P, P, c_ll, c_int, P], c_int),
    "ta_ce_fwd_bwd": ([P, c_ll, P, c_ll, c_int, c_int, c_float, P, P, c_int, P], c_int),
    "ta_transpose_bf16": ([P, P, c_int, c_int, c_ll, c_ll, P], c_int),
    "ta_cast_f32_bf16": ([P, P, c_ll, P], c_int),
    "ta_frame_stack": ([P, P, c_int, c_int, c_int, c_int, c_int, P], c_int),
    "ta_frame_keep_mask": ([P, P, c_ll, c_int, P], c_int),
    "ta_add_layernorm_fwd": ([P, P, P, c_ll, P, P, P, c_ll, P, P, P, c_ll, c_int, c_float, P], c_int),
    "ta_add_layernorm_bwd": ([P, P, P, P, P, c_ll, P, P, c_ll, P, P, P, P, P, P, c_ll, c_int, P], c_int),
    "ta_add_layernorm_bwd_partial_floats": ([c_int], c_ll),
    "ta_gelu_fwd_bf16": ([P, P, c_ll, P], c_int),
    "ta_gelu_bwd_bf16": ([P, P, P, c_ll, P], c_int),
    "ta_colsum_bf16": ([P, c_ll, P, c_ll, c_int, P], c_int),
    "ta_label_rows": ([P, c_int, c_int, P, P, P, P], c_int),
    "ta_assemble_prompts": ([P, P, P, P, c_int, P, c_int, P, c_int, c_ll, c_ll, c_int, c_int, P, P, P, P], c_int),
    "ta_window_attn_fwd": ([P, P, P, P, P, c_ll, c_int, c_int, c_int, c_int, c_float, P], c_int),
    "ta_window_attn_bwd": ([P, P, P, P, P, P, P, P, c_ll, c_int, c_int, c_int, c_int, c_float, P], c_int),
    "ta_window_attn_set_variant": ([c_int], c_int),
    "ta_grad_sumsq": ([P, c_ll, P, P], c_int),
    "ta_adamw_clip_step": ([P, P, P, P, c_ll, c_float, c_float, c_float, c_float, c_float, c_int, c_float, P, P], c_int),
    "ta_encoder_workspace_bytes": ([C.POINTER(EncoderWeights), c_int, c_int, C.POINTER(c_ll)], c_int),
    "ta_encoder_forward": ([C.POINTER(EncoderWeights), P, c_int, c_int, P, c_ll, P, P], c_int),
    "ta_mlp_projector_forward": ([C.POINTER(MlpProjectorWeights), P, c_ll, P, P, P, P, P], c_int),
    "ta_mlp_projector_backward_workspace_bytes": ([C.POINTER(MlpProjectorWeights), c_ll, C.POINTER(c_ll)], c_int),
    "ta_mlp_projector_backward": ([C.POINTER(MlpProjectorWeights), P, c_ll, P, P, P, P, P, c_ll, P, P, P, P, P], c_int),
    "ta_lm_workspace_bytes": ([C.POINTER(LmWeights), c_int, c_int, c_int, c_int, C.POINTER(c_ll)], c_int),
    "ta_lm_forward_backward": ([C.POINTER(LmWeights), C.POINTER(LmStepArgs), P], c_int),
    "ta_lm_hidden_to_logits": ([C.POINTER(LmWeights), P, P, c_int, P, P, P], c_int),
    "ta_rmsnorm_dw": ([P, P, P, c_ll, c_int, c_float, P, P], c_int),
    "ta_qknorm_dw": ([P, P, P, P, P, c_ll, c_int, c_int, c_int, c_float, P, P, P], c_int),
    "ta_embed_grad_scatter": ([P, P, P, c_ll, c_int, c_ll, c_ll, P], c_int),
    "ta_pack_weight": ([P, c_int, c_int, P, c_ll, P, c_ll, c_int, c_int, c_int, P], c_int),
    "ta_skinny_gemm_bf16": ([P, c_ll, P, c_ll, c_int, c_int, c_int, c_int, P, c_ll, P, c_int, P], c_int),
    "ta_decode_resid_rmsnorm": ([P, P, c_int, c_int, P, P, P, c_int, c_float, c_ll, P], c_int),
    "ta_decode_attn": ([P, P, P, P, c_ll, P, c_int, c_int, c_int, c_int, c_float, P], c_int),
    "ta_argmax_rows": ([P, c_ll, c_int, c_int, P, P], c_int),
    "ta_lm_decode_workspace_bytes": ([C.POINTER(LmWeights), c_int, C.POINTER(c_ll)], c_int),
    "ta_lm_decode_step": ([C.POINTER(LmWeights), P, P, c_int, P, P, c_int, c_int, P, c_ll, P, P, P, P], c_int),
}

EXPORTED_SYMBOLS = sorted(list(_SIGS) + ["ta_last_error_string"])

_lib = None


class TinyAudioB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TinyAudioB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C tiny_audio_b200/csrc`).  tiny_audio_b200 has no CPU / PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.ta_last_error_string.argtypes = []
    lib.ta_last_error_string.restype = C.c_char_p
    for name, (args, res) in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    # tuning / A-B switches (both kernels of each pair are parity-tested; defaults are the fast ones)
    if os.environ.get("TA_GEMM_CTA_PAIR") is not None:
        lib.ta_gemm_set_cta_pair(int(os.environ["TA_GEMM_CTA_PAIR"]))
    if os.environ.get("TA_GEMM_TN_SPLITK") is not None:
        lib.ta_gemm_set_tn_splitk(int(os.environ["TA_GEMM_TN_SPLITK"]))
    if os.environ.get("TA_GEMM_RESID_TMA") is not None:
        lib.ta_gemm_set_resid_tma(int(os.environ["TA_GEMM_RESID_TMA"]))
    if os.environ.get("TA_GEMM_SWIGLU_BWD_TMA") is not None:
        lib.ta_gemm_set_swiglu_bwd_tma(int(os.environ["TA_GEMM_SWIGLU_BWD_TMA"]))
    if os.environ.get("TA_LN_REVERSE") is not None:
        lib.ta_layernorm_set_reverse(int(os.environ["TA_LN_REVERSE"]))
    if os.environ.get("TA_GEMM_TAIL_SPLIT") is not None:
        lib.ta_gemm_set_tail_split(int(os.environ["TA_GEMM_TAIL_SPLIT"]))
    if os.environ.get("TA_PDL") is not None:          # effective only in a `make PDL=1` build of the library
        if lib.ta_set_pdl(int(os.environ["TA_PDL"])) == 0 and int(os.environ["TA_PDL"]):
            raise TinyAudioB200Error("TA_PDL=1 requested but libtinyaudio_b200.so was built without PDL (make -C tiny_audio_b200/csrc PDL=1)")
    if os.environ.get("TA_WINDOW_ATTN_VARIANT") is not None:
        lib.ta_window_attn_set_variant(int(os.environ["TA_WINDOW_ATTN_VARIANT"]))
    if os.environ.get("TA_LM_FUSED_ATTN_DSUM") is not None:
        lib.ta_lm_set_fused_attn_dsum(int(os.environ["TA_LM_FUSED_ATTN_DSUM"]))
    if os.environ.get("TA_ATTN_BWD_VARIANT") is not None:
        lib.ta_attn_set_bwd_variant(int(os.environ["TA_ATTN_BWD_VARIANT"]))
    if os.environ.get("TA_ATTN_TC_LM") is not None:
        lib.ta_attn_set_tc_lm(int(os.environ["TA_ATTN_TC_LM"]))
    if os.environ.get("TA_ATTN_TC") is not None:
        lib.ta_attn_set_tc(int(os.environ["TA_ATTN_TC"]))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ta_last_error_string().decode("utf-8", "replace")
        raise TinyAudioB200Error(f"libtinyaudio_b200 error {rc}: {msg}")


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TinyAudioB200Error("tiny_audio_b200 kernels need CUDA tensors (no CPU fallback)")


# --------------------------------------------------------------------------------------------------
# thin, typed convenience wrappers used by the host modules and the unit tests
# --------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, epi: int = EPI_BF16, out: Optional[torch.Tensor] = None,
         bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None,
         aux: Optional[torch.Tensor] = None, alpha: float = 1.0, k: Optional[int] = None,
         rope: Optional[tuple] = None, seq: int = 0, zero: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a @ b.T);  a [M,K] bf16, b [N,K] bf16 (both row-major, K contiguous)."""
    lib = load()
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    M, K = a.shape[0], (k if k is not None else a.shape[1])
    N = b.shape[0]
    if out is None:
        if epi == EPI_SWIGLU:
            out = torch.empty(M, N // 2, device=a.device, dtype=torch.bfloat16)
        elif epi == EPI_SWIGLU_BWD:
            out = torch.empty(M, 2 * N, device=a.device, dtype=torch.bfloat16)
        elif epi in (EPI_F32, EPI_F32_RESID):
            out = torch.empty(M, N, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    e = GemmEpilogue(ptr(out), out.stride(0), ptr(bias), ptr(resid), resid.stride(0) if resid is not None else 0,
                     ptr(out2), out2.stride(0) if (out2 is not None and out2.dim() == 2) else 0, ptr(aux),
                     aux.stride(0) if aux is not None else 0, alpha,
                     ptr(rope[0]) if rope else None, ptr(rope[1]) if rope else None, rope[2] if rope else seq, rope[3] if rope else 0,
                     ptr(zero), zero.stride(0) if zero is not None else 0)
    check(lib.ta_gemm_bf16(ptr(a), a.stride(0), ptr(b), b.stride(0), M, N, K, epi, C.byref(e), stream_ptr()))
    return out


def gemm_tn(at: torch.Tensor, bt: torch.Tensor, out: Optional[torch.Tensor] = None, alpha: float = 1.0) -> torch.Tensor:
    """out fp32 [M, N] = alpha * at.T @ bt;  at [K, M] bf16, bt [K, N] bf16 (row-major): dW = dY^T X without transposed copies."""
    lib = load()
    require_cuda(at, bt)
    assert at.dtype == torch.bfloat16 and bt.dtype == torch.bfloat16 and at.stride(-1) == 1 and bt.stride(-1) == 1
    K, M = at.shape
    N = bt.shape[1]
    assert bt.shape[0] == K
    if out is None:
        out = torch.empty(M, N, device=at.device, dtype=torch.float32)
    check(lib.ta_gemm_bf16_tn(ptr(at), at.stride(0), ptr(bt), bt.stride(0), M, N, K, ptr(out), out.stride(0), alpha, stream_ptr()))
    return out


def pointer_table(tensors) -> "C.Array":
    arr = (P * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr
