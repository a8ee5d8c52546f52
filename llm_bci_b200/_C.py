"""ctypes binding of the C ABI in include/ndt1_b200.h.

This is the whole "extension": torch is used for device memory and streams
only; every call passes raw device pointers (``tensor.data_ptr()``), sizes and
the current ``cudaStream_t``.  There is no fallback: if the library is missing
or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libndt1_b200.so")
ABI_VERSION = 3
MAX_LAYERS = 32
MAX_DAYS = 64

ACT = {"identity": 0, "softsign": 1, "gelu": 2, "relu": 3}
METHOD = {"ctc": 0, "endtoend": 0, "mlm": 1, "autoregressive": 2}
LOSS_CTC, LOSS_POISSON_LOG, LOSS_POISSON_RATE, LOSS_MSE = 0, 1, 2, 3
PRECISION = {"fp32": 0, "bf16": 1}
MASK_MODE = {"temporal": 0, "neuron": 1, "region": 1, "random": 2, "co-smooth": 3}

_p = C.c_void_p


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "precision", "n_channels", "input_dim", "max_F", "embed_bias", "embed_act", "pos", "stack_active",
        "stack_size", "stack_stride", "block_token", "day_token", "n_blocks", "n_days", "adapt", "n_layers", "hidden",
        "n_heads", "inter", "attention_bias", "mlp_bias", "mlp_act", "use_rope")] + [("rope_theta", C.c_float)] + [
        (n, C.c_int32) for n in ("context_forward", "context_backward", "factors_active", "factors_size", "factors_act",
                                 "factors_bias", "method", "loss_kind", "n_outputs", "blank_id", "zero_infinity",
                                 "decoder_relu")] + [(n, C.c_float) for n in ("p_embed", "p_transformer", "p_factors")] + [
        (n, C.c_int32) for n in ("max_batch", "max_T", "max_targets")]


_LAYER_FIELDS = ("ln1_w", "ln1_b", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b", "ln2_w", "ln2_b", "up_w", "up_b",
                 "down_w", "down_b")


class LayerTensors(C.Structure):
    _fields_ = [(n, _p) for n in _LAYER_FIELDS]


class Tensors(C.Structure):
    _fields_ = [(n, _p) for n in ("embed_w", "embed_b", "proj_w", "proj_b", "pos_w", "block_emb", "day_emb")] + [
        ("embed_w_day", _p * MAX_DAYS), ("embed_b_day", _p * MAX_DAYS), ("layer", LayerTensors * MAX_LAYERS)] + [(n, _p) for n in ("out_norm_w", "out_norm_b", "factors_w", "factors_b",
                                                                   "dec_w", "dec_b")]


class Batch(C.Structure):
    _fields_ = [(n, _p) for n in ("spikes", "spikes_mask", "spikes_timestamp", "spikes_lengths", "block_idx", "day_idx",
                                  "targets", "targets_lengths", "recon_targets", "targets_mask")] + [
        ("B", C.c_int32), ("T", C.c_int32), ("S", C.c_int32), ("training", C.c_int32), ("need_backward", C.c_int32), ("encoder_only", C.c_int32),
        ("seed", C.c_uint64), ("spikes_bf16", _p), ("seed_ptr", _p)]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 160), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double)]


class Outputs(C.Structure):
    _fields_ = [(n, _p) for n in ("loss", "n_examples", "preds", "out_mask", "loss_mask", "out_lengths", "features")]


# name -> (restype, argtypes); must list every function include/ndt1_b200.h declares
_i, _i64, _u64, _f, _d, _sz = C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_size_t
PROTOTYPES = {
    "ndt1_last_error": (C.c_char_p, []),
    "ndt1_abi_version": (_i, []),
    "ndt1_smooth_noise": (_i, [_p, _p, _i, _i, _i, C.POINTER(C.c_float), _i, _f, _f, _p, _p, _i, _u64, _p, _p, _i, _p]),
    "ndt1_masker_apply": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ndt1_bernoulli_u8": (_i, [_p, _i64, _f, _u64, _u64, _p]),
    "ndt1_uniform_f32": (_i, [_p, _i64, _u64, _u64, _p]),
    "ndt1_pad_pack": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _d, _p]),
    "ndt1_ctc_workspace_bytes": (_sz, [_i, _i, _i]),
    "ndt1_ctc_loss": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "ndt1_ctc_greedy_decode": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "ndt1_edit_distance": (_i, [_p, _p, _i, _p, _p, _i, _i, _p, _p]),
    "ndt1_recon_loss": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ndt1_layernorm_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i64, _i, _f, _p]),
    "ndt1_linear_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "ndt1_linear_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    "ndt1_linear_drop_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _f, _u64, _u64, _p]),
    "ndt1_linear_drop_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _sz, _f, _u64, _u64, _p]),
    "ndt1_linear_workspace_bytes": (_sz, [_i, _i, _i]),
    "ndt1_splice_rows": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _d, _p]),
    "ndt1_unsplice_rows": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ndt1_stack_valid": (_i, [_p, _p, _i, _i, _i, _p]),
    "ndt1_attention_workspace_bytes": (_sz, [_i, _i, _i]),
    "ndt1_attention_bf16": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _u64, _u64, _u64, _p, _p, _p, _i, _p]),
    "ndt1_attention_f32": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _u64, _u64, _u64, _p, _p, _p, _p]),
    "ndt1_attention_mm_saved_bytes": (_sz, [_i, _i, _i, _i, _f]),
    "ndt1_attention_mm_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ndt1_attention_mm_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _u64, _u64, _p]),
    "ndt1_attention_mm_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _u64, _u64, _p]),
    "ndt1_xent_loss": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "ndt1_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _p]),
    "ndt1_dropout_inplace": (_i, [_p, _i64, _f, _u64, _u64, _p]),
    "ndt1_adamw_step": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i, _f, _p]),
    "ndt1_engine_create": (_i, [C.POINTER(Config), C.POINTER(_p)]),
    "ndt1_engine_destroy": (None, [_p]),
    "ndt1_engine_arena_bytes": (_sz, [_p]),
    "ndt1_engine_out_len": (_i, [_p, _i]),
    "ndt1_engine_forward": (_i, [_p, C.POINTER(Tensors), C.POINTER(Batch), C.POINTER(Outputs), _p]),
    "ndt1_engine_backward": (_i, [_p, C.POINTER(Tensors), C.POINTER(Tensors), _p, _p]),
    "ndt1_engine_backward_features": (_i, [_p, C.POINTER(Tensors), C.POINTER(Tensors), _p, _p]),
    "ndt1_engine_launch_count": (_i64, [_p]),
    "ndt1_engine_set_overlap": (_i, [_p, _i]),
    "ndt1_engine_set_weight_shadow": (_i, [_p, _p, _p, _i64]),
    "ndt1_adamw_step_fused": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i, _f, _p, _i, _p]),
    "ndt1_debug_attention_timeline": (_i, [_p]),
    "ndt1_debug_gemm_timeline": (_i, [_p]),
    "ndt1_engine_stage_count": (_i, [_p]),
    "ndt1_engine_wait_stage": (_i, [_p, _i, _p]),
    "ndt1_engine_set_rope_tables": (_i, [_p, _p, _p, _i]),
    "ndt1_profile_gemm_begin": (_i, []),
    "ndt1_profile_gemm_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ndt1_launch_counter": (_i64, []),
    "ndt1_event_create": (_i, [C.POINTER(_p)]),
    "ndt1_event_destroy": (_i, [_p]),
    "ndt1_event_record": (_i, [_p, _p]),
    "ndt1_stream_wait_event": (_i, [_p, _p]),
    "ndt1_profile_begin": (_i, []),
    "ndt1_profile_end": (_i, [C.POINTER(ProfileEntry), _i, C.POINTER(C.c_int)]),
    "ndt1_dropout_scales": (_i, [_p, _i64, _f, _u64, _u64, _p]),
}

_lib = None


def build_library(verbose: bool = False) -> str:
    """Compile csrc/ into libndt1_b200.so (nvcc, sm_100a).  Used by __graft_entry__.build()."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(max(1, os.cpu_count() or 1))]
    r = subprocess.run(cmd, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libndt1_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the library (once) and install the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the NDT1 kernels)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(L, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.ndt1_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libndt1_b200.so has ABI {L.ndt1_abi_version()}, binding expects {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ndt1_last_error()
        raise RuntimeError(f"{what or 'ndt1 call'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def profile_begin() -> None:
    check(lib().ndt1_profile_begin(), "ndt1_profile_begin")


def profile_end(capacity: int = 256):
    """[{name, launches, ms, flops, bytes}] per kernel of the library since profile_begin() (CUDA events around every launch)."""
    ent = (ProfileEntry * capacity)()
    n = C.c_int(0)
    check(lib().ndt1_profile_end(ent, capacity, C.byref(n)), "ndt1_profile_end")
    return [dict(name=ent[i].name.decode(), launches=int(ent[i].launches), ms=float(ent[i].ms), flops=float(ent[i].flops), bytes=float(ent[i].bytes))
            for i in range(n.value)]
