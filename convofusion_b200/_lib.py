"""ctypes binding of include/convofusion_b200.h.  There is no CPU fallback: if the shared library
cannot be loaded every call raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

ABI_VERSION = 2
N_STREAMS = 5
N_BRANCH = 7
F32, BF16 = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2
ACT = {"none": 0, "gelu": 1, "silu": 2, "relu": 3, "leaky01": 4}
SCHED_DDIM, SCHED_DDPM = 0, 1

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libconvofusion_b200.so"

DENOISER_LAYER_FIELDS = [
    "ln1_g", "ln1_b", "w_in", "b_in", "w_so", "b_so", "tb1_g", "tb1_b", "w_tb1", "b_tb1", "ln2_g", "ln2_b",
    "w_qx", "b_qx", "w_fu", "b_fu", "tb2_g", "tb2_b", "w_tb2", "b_tb2", "ln3_g", "ln3_b", "w_ff1", "b_ff1",
    "w_ff2", "b_ff2"]
VAE_LAYER_FIELDS = [
    "ln1_g", "ln1_b", "w_in", "b_in", "w_so", "b_so", "ln2_g", "ln2_b", "w_q", "b_q", "w_kv", "b_kv", "w_co",
    "b_co", "ln3_g", "ln3_b", "w_ff1", "b_ff1", "w_ff2", "b_ff2"]


class DenoiserLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in DENOISER_LAYER_FIELDS]


class DenoiserWeights(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("d_model", "latent_dim", "n_tokens", "n_layers", "n_heads", "ff_size",
                                          "precision", "pe_len")] +
                [(n, C.c_void_p) for n in ("w_embed", "tok_bias", "w_t1", "b_t1", "w_t2", "b_t2", "w_tbmod",
                                           "b_tbmod", "stream_emb", "pe_mem", "lnf_g", "lnf_b", "w_out", "b_out")] +
                [("w_zx", C.c_void_p * N_STREAMS), ("a_zx", C.c_void_p * N_STREAMS), ("w_yx", C.c_void_p * N_STREAMS),
                 ("layers", C.POINTER(DenoiserLayer))])


class Memory(C.Structure):
    _fields_ = [("cond", C.c_void_p * N_STREAMS), ("mask", C.c_void_p * N_STREAMS), ("slot", C.c_void_p * N_STREAMS),
                ("slot_host", C.c_void_p * N_STREAMS), ("n_slots", C.c_int32 * N_STREAMS), ("len", C.c_int32 * N_STREAMS)]


class Schedule(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_steps", C.c_int32), ("clip_sample", C.c_int32), ("guidance_scale", C.c_float),
                ("timesteps", C.POINTER(C.c_int64)), ("coef", C.POINTER(C.c_float))]


class VaeLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in VAE_LAYER_FIELDS]


class VaeDecoder(C.Structure):
    _fields_ = [("layers", C.POINTER(VaeLayer)), ("w_skip", C.c_void_p * 4), ("b_skip", C.c_void_p * 4),
                ("lnf_g", C.c_void_p), ("lnf_b", C.c_void_p), ("w_final", C.c_void_p), ("b_final", C.c_void_p),
                ("n_out", C.c_int32)]


VAE_ENC_LAYER_FIELDS = ["ln1_g", "ln1_b", "w_in", "b_in", "w_so", "b_so", "ln2_g", "ln2_b", "w_ff1", "b_ff1", "w_ff2", "b_ff2"]


class VaeEncLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in VAE_ENC_LAYER_FIELDS]


class VaeEncoder(C.Structure):
    _fields_ = [("layers", C.POINTER(VaeEncLayer)), ("w_skip", C.c_void_p * 4), ("b_skip", C.c_void_p * 4),
                ("lnf_g", C.c_void_p), ("lnf_b", C.c_void_p), ("tokens", C.c_void_p), ("w_emb", C.c_void_p),
                ("b_emb", C.c_void_p), ("n_in", C.c_int32), ("col0", C.c_int32)]


class VaeWeights(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("d_model", "n_layers", "n_heads", "ff_size", "precision", "pe_len")] +
                [("pe_query", C.c_void_p), ("pe_mem", C.c_void_p), ("part", VaeDecoder * 2),
                 ("pe_enc", C.c_void_p), ("enc", VaeEncoder * 2)])


# name -> (restype, argtypes); every symbol include/convofusion_b200.h declares.
_P, _I, _F, _LL = C.c_void_p, C.c_int, C.c_float, C.c_int64
PROTOTYPES = {
    "cfb_abi_version": (C.c_int, []),
    "cfb_last_error": (C.c_char_p, []),
    "cfb_set_gemm_backend": (C.c_int, [_I]),
    "cfb_set_shared_plan": (C.c_int, [_I]),
    "cfb_set_rowblock": (C.c_int, [_I]),
    "cfb_set_fp32_tensor_cores": (C.c_int, [_I]),
    "cfb_set_cross_tc": (C.c_int, [_I]),
    "cfb_set_bf16_activation_terms": (C.c_int, [_I]),
    "cfb_set_bf16_activation_sites": (C.c_int, [_I]),
    "cfb_set_bf16_activation_f16": (C.c_int, [_I]),
    "cfb_set_vae_f16": (C.c_int, [_I]),
    "cfb_get_vae_f16": (C.c_int, []),
    "cfb_launch_count": (C.c_ulonglong, []),
    "cfb_denoiser_create": (C.c_int, [C.POINTER(DenoiserWeights), C.POINTER(_P)]),
    "cfb_denoiser_destroy": (None, [_P]),
    "cfb_denoiser_attach_f16_weights": (C.c_int, [_P, C.POINTER(DenoiserWeights)]),
    "cfb_denoiser_set_chains": (C.c_int, [_P, C.c_int]),
    "cfb_denoiser_forward": (C.c_int, [_P, _P, _I, _LL, C.POINTER(Memory), _P, C.POINTER(_P), _P]),
    "cfb_denoiser_weg_forward": (C.c_int, [_P, _P, _I, _LL, C.POINTER(Memory), _I, _P, _P]),
    "cfb_denoiser_weg_backward": (C.c_int, [_P, _P, _P, _P]),
    "cfb_sample": (C.c_int, [_P, C.POINTER(Schedule), C.POINTER(Memory), _I, _I, _I, _P, _P, _P, _I, _P, C.POINTER(_P),
                             _I, _P]),
    "cfb_guidance_sched_step": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "cfb_vae_create": (C.c_int, [C.POINTER(VaeWeights), C.POINTER(_P)]),
    "cfb_vae_destroy": (None, [_P]),
    "cfb_vae_decode": (C.c_int, [_P, _P, _I, _I, _I, C.POINTER(C.c_int32), _P, _P]),
    "cfb_vae_encode": (C.c_int, [_P, _P, _I, _I, C.POINTER(C.c_int32), _P, _P, _P, _P]),
    "cfb_linear": (C.c_int, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "cfb_layernorm": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "cfb_mha": (C.c_int, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "cfb_keypoints3d": (C.c_int, [_P, C.c_longlong, _P, _P]),
    "cfb_audio_encoder": (C.c_int, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
}

_lib = None


class CfbError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) the in-tree CUDA library; raise loudly when it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CfbError(f"{LIB_PATH} is missing: run `python -m convofusion_b200.build` (needs nvcc). "
                           "convofusion_b200 has no CPU fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.cfb_abi_version() != ABI_VERSION:
            raise CfbError("libconvofusion_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().cfb_last_error().decode("utf-8", "replace")
        exc = ValueError if status == -1 else CfbError
        raise exc(f"convofusion_b200 [{status}]: {msg}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
