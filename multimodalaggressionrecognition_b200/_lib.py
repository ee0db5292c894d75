"""ctypes binding of libmar.so (include/mar.h).

The library is loaded lazily on first use and never stored on a module instance, so the drop-in
nn.Modules stay picklable (the reference checkpoints by pickling the whole trainer,
trainer.py:336).  There is NO fallback: if libmar.so is missing or the device is not sm_100 every
op raises RuntimeError.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAR_LIB", os.path.join(_HERE, "libmar.so"))   # MAR_LIB: debugging builds of the same ABI

MAR_F32, MAR_BF16 = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
EPI_RELU_PRE, EPI_DROPOUT, EPI_RELU_POST = 1, 2, 4

P = c_void_p
# name -> (restype, argtypes); mirrors include/mar.h one to one (tests/test_abi.py checks the header).
PROTOTYPES = {
    "mar_version": (c_int, []),
    "mar_last_error": (c_char_p, []),
    "mar_device_info": (c_int, [P, P, P]),
    "mar_launch_count": (c_int64, []),
    "mar_launch_count_reset": (None, []),
    "mar_last_engine": (c_int, []),
    "mar_rng_init": (c_int, [P, c_uint64, c_uint64, P]),
    "mar_rng_advance": (c_int, [P, P]),
    "mar_linear_fwd": (c_int, [P, c_int64, P, P, P, c_int64, P, c_int64, c_int64, c_int64, c_int64, c_int, c_int,
                               c_int, c_float, P, c_uint32, c_int, P]),
    "mar_linear_bwd_epilogue": (c_int, [P, P, P, P, c_int64, c_int64, c_int, c_int, c_int, c_float, P, c_uint32, c_int64, P]),
    "mar_linear_dgrad": (c_int, [P, P, P, P, P, c_float, P, c_int64, P, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "mar_linear_wgrad": (c_int, [P, P, c_int64, P, c_int64, c_int64, c_int64, c_int, c_int, c_int, P]),
    "mar_cast_weight": (c_int, [P, P, P, c_int64, c_int64, c_int, P]),
    "mar_cast": (c_int, [P, c_int, P, c_int, c_int64, P]),
    "mar_attention_fwd": (c_int, [P, P, P, P, c_int64, c_int64, c_int64, c_int64, c_int, c_float, P, c_uint32, P, c_int, P]),
    "mar_attention_dropbits_words": (c_int64, [c_int64, c_int64, c_int64]),
    "mar_attention_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int64, c_int, c_float, P, c_int, P]),
    "mar_attention_bwd_work_floats": (c_int64, [c_int64, c_int64, c_int64, c_int64]),
    "mar_layernorm_fwd": (c_int, [P, P, P, P, P, P, P, c_int64, c_int64, c_float, c_int, P]),
    "mar_layernorm_bwd": (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, P]),
    "mar_layernorm_bwd_dropout": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, c_float, P, c_uint32, P]),
    "mar_layernorm_fwd_mapped": (c_int, [P, P, P, P, P, P, P, c_int64, c_int64, c_float, c_int, c_int, c_int64, P, P, P, P, P, P]),
    "mar_layernorm_bwd_mapped": (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int64, c_int, c_int, c_int64, P, P, P, P, P, P]),
    "mar_meanpool_fwd": (c_int, [P, P, c_int64, c_int64, c_int64, c_int, P]),
    "mar_meanpool_bwd": (c_int, [P, P, c_int64, c_int64, c_int64, c_int, P]),
    "mar_rowzero_mask": (c_int, [P, P, c_int64, c_int64, c_int, P]),
    "mar_concat_rows": (c_int, [P, P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "mar_cross_entropy_fwd": (c_int, [P, P, P, P, P, P, c_int64, c_int64, P]),
    "mar_focal_loss_fwd": (c_int, [P, P, P, c_float, P, P, c_int64, c_int64, P]),
    "mar_gru_fwd": (c_int, [P, P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "mar_gru_bwd": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "mar_gru_work_floats": (c_int64, [c_int64, c_int64, c_int64]),
    "mar_lstm_fwd": (c_int, [P, P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "mar_lstm_bwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P]),
    "mar_adam_step": (c_int, [P, P, P, P, P, c_int64, c_float, c_float, c_float, c_float, P]),
    "mar_adam_tick": (c_int, [P, P]),
    "mar_adam_step_segments": (c_int, [P, P, P, P, P, P, P, P, c_int64, c_int, c_int, c_float, c_float, c_float, c_float, P, P]),
    "mar_label_weight_sum": (c_int, [P, P, P, c_int64, c_int64, P]),
    "mar_dp_block_bytes": (c_int64, [c_int64]),
    "mar_dp_ctrl_bytes": (c_int64, []),
    "mar_peer_alloc": (c_int, [P, c_int64]),
    "mar_peer_free": (c_int, [P]),
    "mar_peer_export": (c_int, [P, P]),
    "mar_peer_import": (c_int, [P, P]),
    "mar_peer_close": (c_int, [P]),
    "mar_dp_allreduce_adam": (c_int, [P, c_int, c_int64, c_int, c_int, P, P, P, P, P, P, P, c_int, c_int, c_float, c_float,
                                      c_float, c_float, P, P]),
    "mar_dp_check": (c_int, [P, P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libmar.so and bind every prototype; raises RuntimeError if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C multimodalaggressionrecognition_b200/csrc`). There is no CPU or PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)   # AttributeError => header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().mar_last_error().decode("utf-8", "replace")


def call(name: str, *args) -> None:
    """Call an int-returning entry point; raise RuntimeError carrying mar_last_error() on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")


_checked_devices = set()


def check_device(index: int) -> None:
    """The kernels are sm_100a only; fail loudly anywhere else."""
    if index in _checked_devices:
        return
    import torch
    with torch.cuda.device(index):
        call("mar_device_info", None, None, None)
    _checked_devices.add(index)
