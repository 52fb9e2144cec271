"""ctypes binding of libsfb200.so (the C ABI declared in include/sfb200.h).

There is no fallback: if the library is missing or a call fails the caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsfb200.so")

P, LL, I, F = c_void_p, c_longlong, c_int, c_float
PP = ctypes.POINTER(c_void_p)
FP = ctypes.POINTER(c_float)   # host array of floats
ABI_VERSION = 11

# name -> argtypes, mirroring include/sfb200.h one to one
SIGNATURES = {
    "sfb_gemm_bf16": [P, LL, P, LL, P, I, I, I, I, P, LL, P, LL, P, LL, I, P, LL, P, LL, I, I, I, P, LL, P],
    "sfb_gemm_bf16_stats": [P, LL, P, LL, P, I, I, I, I, P, LL, P, LL, P, LL, I, P, LL, P, LL, I, I, I, P, P, P, F, P],
    "sfb_attention_fwd": [P, LL, LL, P, P, LL, LL, P, LL, LL, I, I, I, I, I, F, P, LL, P],
    "sfb_attention_fwd_qnorm": [P, LL, LL, P, P, LL, LL, P, LL, LL, I, I, I, I, I, F, P, I, F, P, LL, P],
    "sfb_modulation_table": [P, P, P, I, I, I, I, LL, LL, P],
    "sfb_ln_modulate": [P, LL, P, LL, I, I, F, P, P, LL, I, I, P],
    "sfb_ln_modulate_stats": [P, LL, P, LL, I, I, F, P, P, LL, I, I, P, I, P],
    "sfb_ln_affine": [P, LL, P, LL, I, I, F, P, P, P],
    "sfb_rmsnorm": [P, LL, P, LL, I, I, F, P, P],
    "sfb_qk_norm_rope": [P, LL, P, LL, P, LL, P, P, F, P, P, I, I, I, I, I, I, I, I, I, P, P, LL, LL, P, P, LL, LL, P],
    "sfb_qk_norm_rope_stats": [P, LL, P, LL, P, LL, P, P, F, P, I, I, I, P, P, I, I, I, I, I, I, I, I, I, P, P, LL, LL, P, P, LL, LL, P],
    "sfb_kv_roll": [P, I, I, LL, LL, LL, LL, LL, P],
    "sfb_patchify": [P, LL, LL, LL, LL, LL, P, I, I, I, I, I, P],
    "sfb_sinusoid": [P, I, P, I, I, P],
    "sfb_skinny_linear": [P, LL, P, LL, P, P, LL, I, I, I, I, P],
    "sfb_head_finish": [P, LL, P, LL, LL, LL, LL, LL, P, I, P, P, I, P, P, I, I, I, I, I, P],
    "sfb_add_noise": [P, P, P, I, P, P, I, P, I, I, P],
    "sfb_cfg_unipc_step": [P, P, P, P, P, P, P, P, P, LL, FP, I, I, P],
    # VAE decoder
    "sfb_vae_latent_in": [P, LL, P, P, P, P, P, I, P],
    "sfb_vae_norm_silu": [P, LL, P, P, LL, LL, I, I, P],
    "sfb_causal_conv3d_cl": [P, I, I, I, I, I, I, P, P, I, I, I, P, LL, P, P, LL, I, P, LL, P],
    "sfb_upsample2x_cl": [P, P, I, I, I, I, P],
    "sfb_softmax_rows": [P, LL, P, LL, I, I, F, P],
    "sfb_transpose_bf16": [P, LL, P, LL, I, I, P],
    "sfb_vae_pixel_out": [P, I, P, I, LL, P],
    # UMT5 text encoder
    "sfb_t5_rmsnorm": [P, LL, P, LL, I, I, F, P, P],
    "sfb_softmax_bias_rows": [P, LL, P, LL, P, P, LL, I, I, P],
    "sfb_t5_gated_gelu": [P, LL, P, LL, P, LL, I, I, P],
    # Ulysses head-parallel path: PP = host array of device pointers (ctypes c_void_p * n)
    "sfb_qk_norm_rope_sp": [P, LL, P, LL, P, LL, P, P, F, P, P, I, I, I, I, I, I, I, I, I, I, PP, LL, PP, PP, LL, P],
    "sfb_attention_fwd_sp": [P, LL, P, P, LL, PP, I, I, LL, I, I, I, I, F, P, LL, P],
    "sfb_peer_barrier": [PP, I, I, P],
}


def ptr_array(ptrs):
    """Host array of device pointers for the PP arguments."""
    return (c_void_p * len(ptrs))(*ptrs)


class SfbError(RuntimeError):
    pass


_lib = None


def load(path: str | None = None) -> ctypes.CDLL:
    """Load the shared library (once) and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("SFB200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise SfbError(
            f"{path} not found: build it with `python -m self_forcing_b200.build` "
            "(there is no CPU / PyTorch fallback for the B200 kernels)")
    lib = ctypes.CDLL(path)
    lib.sfb_last_error.restype = c_char_p
    lib.sfb_last_error.argtypes = []
    lib.sfb_abi_version.restype = c_int
    lib.sfb_abi_version.argtypes = []
    lib.sfb_attention_workspace_bytes.restype = c_longlong
    lib.sfb_attention_workspace_bytes.argtypes = []
    lib.sfb_gemm_workspace_bytes.restype = c_longlong
    lib.sfb_gemm_workspace_bytes.argtypes = []
    lib.sfb_causal_conv3d_workspace_bytes.restype = c_longlong
    lib.sfb_causal_conv3d_workspace_bytes.argtypes = [c_longlong, c_int, c_int, c_int]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    if lib.sfb_abi_version() != ABI_VERSION:
        raise SfbError(f"{path} has ABI version {lib.sfb_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().sfb_last_error().decode("utf-8", "replace")
        raise SfbError(f"{what} failed (status {status}): {msg}")
