"""autograd.Function wrappers over the libmar.so C ABI.

Every Function launches hand-written sm_100a kernels on torch's current CUDA stream; tensors
(inputs, outputs, saved-for-backward, workspaces) are allocated by PyTorch's caching allocator, so
the whole step is CUDA-graph capturable.  No op has a PyTorch or CPU fallback.
"""
from __future__ import annotations

import contextlib
import weakref
import threading
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (EPI_DROPOUT, EPI_RELU_POST, EPI_RELU_PRE, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, MAR_BF16,
                   MAR_F32, call)

# --------------------------------------------------------------------------------------
# precision mode
# --------------------------------------------------------------------------------------
_state = threading.local()


def _cfg():
    if not hasattr(_state, "dtype"):
        _state.dtype = torch.bfloat16
        _state.engine = ENGINE_AUTO
        _state.probe = False
    return _state


def set_precision(mode) -> None:
    """'bf16' (default: bf16 storage + tensor cores, fp32 accumulate) or 'fp32' (SIMT fp32, 1e-4 parity)."""
    d = {"bf16": torch.bfloat16, "fp32": torch.float32, torch.bfloat16: torch.bfloat16,
         torch.float32: torch.float32}.get(mode)
    if d is None:
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    _cfg().dtype = d


def get_precision() -> torch.dtype:
    return _cfg().dtype


@contextlib.contextmanager
def precision(mode):
    old = _cfg().dtype
    set_precision(mode)
    try:
        yield
    finally:
        _cfg().dtype = old


@contextlib.contextmanager
def engine(which: str):
    """Force a kernel engine: 'auto', 'simt' (fp32-math SIMT kernels) or 'tensor' (tcgen05 GEMM, tensor-core
    attention, persistent GRU; raises if a call cannot run there).  Used by tests and benchmarks."""
    e = {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tensor": ENGINE_TCGEN05}[which]
    old = _cfg().engine
    _cfg().engine = e
    try:
        yield
    finally:
        _cfg().engine = old


@contextlib.contextmanager
def shape_probe():
    """Inside this context the drop-in modules return zero tensors of the right shape without launching
    anything, for the reference's CPU shape probing (train_multimodal.py:346-353)."""
    old = _cfg().probe
    _cfg().probe = True
    try:
        yield
    finally:
        _cfg().probe = old


def probing() -> bool:
    return _cfg().probe


def _eng() -> int:
    return _cfg().engine


def last_engine() -> str:
    return {0: "none", 1: "simt", 2: "tensor"}[_lib.load().mar_last_engine()]


def launch_count() -> int:
    return int(_lib.load().mar_launch_count())


def reset_launch_count() -> None:
    _lib.load().mar_launch_count_reset()


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return MAR_F32
    if t.dtype == torch.bfloat16:
        return MAR_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: got a {t.device} tensor. multimodalaggressionrecognition_b200 runs only on sm_100a CUDA "
            "devices; there is no CPU path (use shape_probe() for the reference's CPU shape probing).")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        # kernels are enqueued on torch's CURRENT stream, which belongs to the current device
        raise RuntimeError(f"{what}: tensor lives on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}; "
                           f"call torch.cuda.set_device({idx}) (one process per GPU) or wrap the call in torch.cuda.device({idx})")
    _lib.check_device(idx)


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D view with unit column stride (copy only if needed)."""
    if t.dim() != 2:
        t = t.reshape(-1, t.shape[-1])
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


class _Rng:
    """Device-side (seed, step) + host-side site counter.  See include/mar.h on the mask function."""

    def __init__(self):
        self.states = {}
        self.site = 0
        self.seed = None

    def state(self, device: torch.device) -> torch.Tensor:
        key = device.index if device.index is not None else torch.cuda.current_device()
        st = self.states.get(key)
        if st is None:
            with torch.cuda.device(key):
                st = torch.zeros(2, dtype=torch.int64, device=f"cuda:{key}")
                seed = self.seed if self.seed is not None else torch.initial_seed()
                call("mar_rng_init", st.data_ptr(), seed & 0xFFFFFFFFFFFFFFFF, 0, _stream())
            self.states[key] = st
        return st

    def next_site(self) -> int:
        self.site = (self.site + 1) & 0x7FFFFFFF
        return self.site


_rng = _Rng()


def manual_seed(seed: int) -> None:
    """Seed of the dropout mask function (defaults to torch.initial_seed())."""
    _rng.seed = int(seed)
    _rng.site = 0
    for key, st in _rng.states.items():
        with torch.cuda.device(key):
            call("mar_rng_init", st.data_ptr(), seed & 0xFFFFFFFFFFFFFFFF, 0, _stream())


def rng_advance(device=None) -> None:
    """Bump the device-side step (captured into CUDA graphs so every replay draws fresh masks)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = _rng.state(dev)
    call("mar_rng_advance", st.data_ptr(), _stream())


# ---- compute-dtype copies of the fp32 master weights ---------------------------------------
_wcache = {}
_wepoch = [0]


def clear_weight_cache() -> None:
    _wcache.clear()


def weights_changed() -> None:
    """Parameters were updated behind torch's back (the fused Adam kernel writes the flat buffer through raw
    pointers, a graph replay does so without running any Python): their version counters did not move, so every
    cached compute-dtype copy is stale from now on."""
    _wepoch[0] += 1


# bf16 mirrors maintained by the step driver's Adam kernel (training.FlatParams): parameter address -> (view, weakref, version)
_mirror = {}


def register_weight_mirror(param: torch.Tensor, view: torch.Tensor) -> None:
    _mirror[param.data_ptr()] = [view, weakref.ref(param), param._version]


def compute_weight(w: torch.Tensor, dtype: torch.dtype, need_t: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """(W, Wᵀ) in the compute dtype.  fp32 mode uses the parameter itself.  bf16 copies are cached per
    parameter version (torch optimizers bump it; `weights_changed()` covers updates torch does not see), so eval
    loops cast once.  An entry belongs to ONE tensor object
    (weak reference): a new model whose fresh parameter lands on a freed parameter's address with the same
    version counter must not pick up the old model's copy."""
    if dtype == torch.float32:
        return (w if w.is_contiguous() else w.contiguous()), None
    if dtype == torch.bfloat16 and not need_t:
        # the step driver keeps a bf16 copy current from inside its Adam kernel; it is valid as long as nobody else wrote
        # the parameter (any torch in-place update bumps the version) — then the ordinary cast below takes over
        ent = _mirror.get(w.data_ptr())
        if ent is not None and ent[1]() is w and ent[2] == w._version and tuple(ent[0].shape) == tuple(w.shape):
            return ent[0], None
    key = (w.data_ptr(), tuple(w.shape), dtype)
    ent = _wcache.get(key)
    ver = w._version
    if ent is not None and ent[0] == ver and ent[3]() is w and ent[4] == _wepoch[0] and (ent[2] is not None or not need_t) \
            and not torch.cuda.is_current_stream_capturing():
        return ent[1], ent[2]
    src = w.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    N, K = src.shape
    wc = torch.empty((N, K), dtype=dtype, device=w.device)
    wt = torch.empty((K, N), dtype=dtype, device=w.device) if need_t else None
    call("mar_cast_weight", src.data_ptr(), wc.data_ptr(), _p(wt), N, K, _dt(wc), _stream())
    _wcache[key] = (ver, wc, wt, weakref.ref(w), _wepoch[0])
    return wc, wt


# --------------------------------------------------------------------------------------
# cast
# --------------------------------------------------------------------------------------
class _Cast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
        call("mar_cast", x.data_ptr(), _dt(x), out.data_ptr(), _dt(out), x.numel(), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty(g.shape, dtype=ctx.src_dtype, device=g.device)
        call("mar_cast", g.data_ptr(), _dt(g), out.data_ptr(), _dt(out), g.numel(), _stream())
        return out, None


def to_compute(x: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Activation in the compute dtype (contiguous)."""
    _require_cuda(x, "input")
    dtype = dtype or get_precision()
    if x.dtype == dtype:
        return x if x.is_contiguous() else x.contiguous()
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    return _Cast.apply(x, dtype)


# --------------------------------------------------------------------------------------
# gradient sink: inside a TrainStep every parameter owns a persistent fp32 .grad view into one flat, pre-zeroed
# buffer.  The weight-gradient GEMM / bias / LayerNorm reductions then accumulate STRAIGHT into that view and the
# autograd Function returns None for the parameter, so no temporary, no zero-fill and no AccumulateGrad add kernel
# is launched per parameter.  `notify` tells the step driver that the parameter's gradient is complete (the
# post-accumulate hook autograd would have fired).
# --------------------------------------------------------------------------------------
_sink_cfg = {"on": False, "notify": None}


@contextlib.contextmanager
def grad_sink(notify=None):
    old = dict(_sink_cfg)
    _sink_cfg["on"], _sink_cfg["notify"] = True, notify
    try:
        yield
    finally:
        _sink_cfg.update(old)


import os as _os
# Which gradients are sunk: weights, biases and "ln" = LayerNorm gains/offsets (MAR_SINK overrides, for A/B runs).  In
# round 1 sinking ALL parameters broke CUDA-graph capture of the step (cudaErrorStreamCaptureIsolation, torch 2.11: no
# AccumulateGrad node ran at all); since the step driver registers a post-accumulate hook on every parameter, created
# on the capture stream, the nodes exist and capture works (measured on B200, round 2: 9.16 -> 9.05 ms per step).
_SINK_KINDS = set(_os.environ.get("MAR_SINK", "w,b,ln").split(","))
_DGRAD_WT = _os.environ.get("MAR_DGRAD_WT") == "1"


def _sink_target(param, kind: str = "w") -> Optional[torch.Tensor]:
    if not _sink_cfg["on"] or not isinstance(param, torch.nn.Parameter) or kind not in _SINK_KINDS:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_cuda or not g.is_contiguous():
        return None
    return g


def _sunk(param) -> None:
    if _sink_cfg["notify"] is not None:
        _sink_cfg["notify"](param)


# --------------------------------------------------------------------------------------
# backward hand-offs between neighbouring ops of an encoder sub-layer.  A linear may be given a `handoff` dict (one per
# forward call).  Its NEIGHBOURS in the backward chain then do parts of its backward inside passes they make anyway:
#   * the LayerNorm that consumes  residual + dropout(linear)  writes the linear's dz (dropout backward) next to its own
#     dx and sums the linear's bias gradient (mar_layernorm_bwd_dropout);
#   * the linear that consumes a deferred ReLU(+dropout) output sums the producer's bias gradient in its dgrad GEMM
#     epilogue (dx_colsum of mar_linear_dgrad);
#   * the attention op sums the in-projection's bias gradient where dQ / dK / dV leave TMEM (dqkv_colsum).
# The neighbour's backward runs first (it produces the linear's incoming gradient), leaves its results in the dict, and
# the linear's own backward picks them up instead of launching mar_linear_bwd_epilogue.  Only valid when the linear's
# output has exactly ONE consumer (true inside models.encoder_layer_forward, the only place that passes a handoff).
# --------------------------------------------------------------------------------------
_handoff_cfg = {"on": _os.environ.get("MAR_HANDOFF", "1") != "0"}      # MAR_HANDOFF=0: the separate passes (A/B runs)


@contextlib.contextmanager
def handoffs(enabled: bool):
    """Switch the backward hand-offs off (every linear launches its own epilogue-backward pass) — tests and A/B runs."""
    old = _handoff_cfg["on"]
    _handoff_cfg["on"] = bool(enabled)
    try:
        yield
    finally:
        _handoff_cfg["on"] = old


def new_handoff() -> Optional[dict]:
    return {} if (_handoff_cfg["on"] and torch.is_grad_enabled()) else None


def _handoff_bias_target(h: Optional[dict], n: int, device) -> Optional[torch.Tensor]:
    """fp32 (n) buffer a neighbour accumulates the linear's bias gradient into (the flat-buffer .grad view inside a
    TrainStep, else fresh zeros), or None when the bias needs no gradient / there is no hand-off."""
    if h is None or h.get("bias_done"):
        return None
    bias = h.get("bias")
    if bias is None or not bias.requires_grad or bias.numel() != n:
        return None
    sunk = _sink_target(bias, "b")
    tgt = sunk if sunk is not None else torch.zeros(n, dtype=torch.float32, device=device)
    h["bias_done"], h["dbias"], h["sunk"] = True, tgt, sunk is not None
    return tgt


# --------------------------------------------------------------------------------------
# linear (+ fused epilogue)
# --------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, flags, p, out_dtype, need_dx, fork=False, x_act_scale=None,
                defer_act=False, pool_T=0, handoff=None, x_handoff=None):
        # x (M,K) compute dtype; weight (N,K) fp32 master; bias (N) fp32; residual (M,N) compute dtype
        # fork: also return x itself as a second output.  A consumer that uses x twice (the projection and the
        # residual branch of an encoder sub-layer) takes the residual from that output, so its gradient arrives
        # HERE and is added inside the dgrad GEMM's epilogue (dx = dz·W + d_fork) instead of in a separate pass.
        # x_act_scale: x came out of a ReLU(+dropout) epilogue whose backward this op applies inside its dgrad GEMM
        #   (dx = dz·W where x > 0, times x_act_scale, else 0) — the producer was called with defer_act=True and only
        #   sums its bias gradient.  pool_T: the epilogue's output is mean-pooled over groups of pool_T rows and the
        #   POOLED tensor is returned; backward broadcasts the pooled gradient inside the epilogue kernel.
        M, K = x.shape
        N = weight.shape[0]
        cd = x.dtype
        # dgrad reads W itself as an MN-major tcgen05 operand: no transposed bf16 copy (MAR_DGRAD_WT=1: the round-1 path)
        need_t = cd == torch.bfloat16 and need_dx and _DGRAD_WT
        wc, wt = compute_weight(weight, cd, need_t)
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
        site = 0
        rng = None
        if (flags & EPI_DROPOUT) and p > 0.0:
            rng = _rng.state(x.device)
            site = _rng.next_site()
        else:
            flags &= ~EPI_DROPOUT
        b32 = None
        if bias is not None:
            b32 = bias.detach()
            if b32.dtype != torch.float32:
                b32 = b32.float()
        call("mar_linear_fwd", x.data_ptr(), x.stride(0), wc.data_ptr(), _p(b32), _p(residual),
             0 if residual is None else residual.stride(0), out.data_ptr(), out.stride(0), M, N, K, _dt(x), _dt(out),
             flags, float(p), _p(rng), site, _eng(), _stream())
        ctx.flags, ctx.p, ctx.site, ctx.rng = flags, float(p), site, rng
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.wc, ctx.wt = wc, wt
        ctx.w_shape = tuple(weight.shape)
        ctx.eng = _eng()
        ctx.x_act_scale = x_act_scale
        ctx.defer_act = bool(defer_act)
        ctx.pool_T = int(pool_T)
        need_out = bool(flags & (EPI_RELU_PRE | EPI_RELU_POST)) and not defer_act
        ctx.save_for_backward(x, out if need_out else None)
        ctx.fork = bool(fork)
        ctx.weight_ref, ctx.bias_ref = weight, bias
        ctx.handoff, ctx.x_handoff = handoff, x_handoff
        if handoff is not None:
            handoff["bias"] = bias if isinstance(bias, torch.nn.Parameter) else None
            if flags == EPI_DROPOUT and residual is not None and not pool_T and out_dtype == x.dtype:
                handoff["drop"] = (rng, site, float(p))          # what the consuming LayerNorm needs to regenerate the mask
        if pool_T:
            pooled = torch.empty((M // pool_T, N), dtype=out_dtype, device=x.device)
            call("mar_meanpool_fwd", out.data_ptr(), pooled.data_ptr(), M // pool_T, pool_T, N, _dt(out), _stream())
            return pooled
        if fork:
            return out, x.view_as(x)
        return out

    @staticmethod
    def backward(ctx, dout, dfork=None):
        x, out = ctx.saved_tensors
        M, K = x.shape
        N = ctx.w_shape[0]
        cd = x.dtype
        st = _stream()
        if dout.dtype != cd:
            dout = _Cast.apply(dout, cd)
        dout = dout.contiguous()
        needs_x, needs_w, needs_b, needs_r = ctx.needs_input_grad[0:4]
        dbias = sunk_b = None
        h = ctx.handoff
        pre_dz = None
        bias_done = False
        if h is not None:
            pre_dz = h.pop("dz", None)
            if pre_dz is not None and (h.pop("dx_ptr", None) != dout.data_ptr() or tuple(pre_dz.shape) != (M, N)):
                raise RuntimeError("linear: the LayerNorm that took over this linear's dropout backward is not the only "
                                   "consumer of its output (hand-offs need a single consumer)")
            if h.pop("bias_done", False):                 # a neighbour's kernel already accumulated the bias gradient
                bias_done = True
                dbias = h.pop("dbias")
                sunk_b = dbias if h.pop("sunk") else None
        if ctx.has_bias and needs_b and not bias_done:
            sunk_b = _sink_target(ctx.bias_ref, "b")
            dbias = sunk_b if sunk_b is not None else torch.zeros(N, dtype=torch.float32, device=x.device)
        flags = 0 if ctx.defer_act else ctx.flags        # deferred: the consumer's dgrad GEMM already applied this mask
        if pre_dz is not None:
            dz = pre_dz                                   # written by mar_layernorm_bwd_dropout next to dout
            if dbias is not None and not bias_done:
                call("mar_linear_bwd_epilogue", dz.data_ptr(), None, None, dbias.data_ptr(), M, N, _dt(dz), _dt(dz), 0, 0.0,
                     None, 0, 0, st)
        elif flags != 0 or ctx.pool_T:
            dz = torch.empty((M, N), dtype=cd, device=x.device)
            call("mar_linear_bwd_epilogue", dout.data_ptr(), _p(out), dz.data_ptr(), _p(dbias), M, N, _dt(dout),
                 _dt(out) if out is not None else _dt(dout), flags, ctx.p, _p(ctx.rng), ctx.site, ctx.pool_T, st)
        else:
            dz = dout
            if dbias is not None and not bias_done:
                call("mar_linear_bwd_epilogue", dout.data_ptr(), None, None, dbias.data_ptr(), M, N, _dt(dout),
                     _dt(dout), 0, 0.0, None, 0, 0, st)
        dx = dw = None
        add = None
        if ctx.fork and dfork is not None:
            add = dfork if dfork.dtype == cd else _Cast.apply(dfork, cd)
            add = add.contiguous()
        if needs_x:
            dx = torch.empty((M, K), dtype=cd, device=x.device)
            act = x if ctx.x_act_scale is not None else None
            if act is not None and add is not None:
                raise RuntimeError("linear: x_act_scale cannot be combined with fork (one staged epilogue input per GEMM)")
            # dx is already the producer's dz when its activation backward is applied here: sum its bias gradient too
            colsum = _handoff_bias_target(ctx.x_handoff, K, x.device) if act is not None else None
            call("mar_linear_dgrad", dz.data_ptr(), ctx.wc.data_ptr(), _p(ctx.wt), _p(add), _p(act),
                 float(ctx.x_act_scale or 1.0), dx.data_ptr(), K, _p(colsum), M, N, K, _dt(dz), ctx.eng, st)
        if needs_w:
            sunk_w = _sink_target(ctx.weight_ref)
            if sunk_w is not None and tuple(sunk_w.shape) == ctx.w_shape:
                call("mar_linear_wgrad", dz.data_ptr(), x.data_ptr(), x.stride(0), sunk_w.data_ptr(), M, N, K, _dt(dz), 1,
                     ctx.eng, st)
                _sunk(ctx.weight_ref)
            else:
                dw = torch.empty(ctx.w_shape, dtype=torch.float32, device=x.device)
                call("mar_linear_wgrad", dz.data_ptr(), x.data_ptr(), x.stride(0), dw.data_ptr(), M, N, K, _dt(dz), 0,
                     ctx.eng, st)
        if sunk_b is not None:
            dbias = None
            _sunk(ctx.bias_ref)
        dres = dout if (ctx.has_res and needs_r) else None
        return dx, dw, dbias, dres, None, None, None, None, None, None, None, None, None, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, relu_pre: bool = False, dropout_p: float = 0.0,
           relu_post: bool = False, out_dtype: Optional[torch.dtype] = None, fork: bool = False,
           x_act_scale: Optional[float] = None, defer_act: bool = False, pool_T: int = 0,
           handoff: Optional[dict] = None, x_handoff: Optional[dict] = None):
    """out = residual + relu_post(dropout(relu_pre(x·Wᵀ + b))) on the last dim of x.
    fork=True returns (out, x_fork): use x_fork wherever x is needed again (a residual branch) and its gradient is
    folded into this linear's dgrad GEMM instead of a separate elementwise add.
    defer_act / x_act_scale (a pair): a ReLU(+dropout) producer called with defer_act=True leaves its activation
    backward to its ONLY consumer, a linear called with x_act_scale = the producer's dropout scale 1/(1-p) (1.0
    without dropout): the consumer's dgrad GEMM writes the gradient already masked (zero where x is zero).
    pool_T > 0 (x must be (B, pool_T, K)): returns mean over the pool_T rows of every group, (B, N).
    handoff / x_handoff: see new_handoff() — this linear's own hand-off dict, and the one of the linear that produced x
    (with x_act_scale: its bias gradient is summed in this linear's dgrad epilogue)."""
    shape = x.shape
    x2 = _rows(to_compute(x))
    r2 = None
    if residual is not None:
        r2 = _rows(to_compute(residual, x2.dtype))
    flags = (EPI_RELU_PRE if relu_pre else 0) | (EPI_DROPOUT if dropout_p > 0 else 0) | (EPI_RELU_POST if relu_post else 0)
    need_dx = torch.is_grad_enabled() and x2.requires_grad
    if defer_act and not (relu_pre or relu_post):
        raise ValueError("defer_act needs a ReLU in the epilogue: the consumer reads the mask off the output's zeros")
    if residual is not None and (relu_pre or relu_post):
        raise ValueError("residual cannot be combined with a ReLU epilogue (the saved output would not carry the ReLU mask)")
    if pool_T:
        if fork or residual is not None or x.dim() != 3 or x.shape[1] != pool_T:
            raise ValueError("pool_T needs x of shape (B, pool_T, K), no fork and no residual")
        return _Linear.apply(x2, weight, bias, None, flags, float(dropout_p), out_dtype or x2.dtype, need_dx, False,
                             x_act_scale, defer_act, int(pool_T), None, None)
    if fork and need_dx:
        out, xf = _Linear.apply(x2, weight, bias, r2, flags, float(dropout_p), out_dtype or x2.dtype, need_dx, True,
                                x_act_scale, defer_act, 0, handoff, x_handoff)
        return out.view(*shape[:-1], weight.shape[0]), xf.view(shape)
    out = _Linear.apply(x2, weight, bias, r2, flags, float(dropout_p), out_dtype or x2.dtype, need_dx, False,
                        x_act_scale, defer_act, 0, handoff, x_handoff)
    out = out.view(*shape[:-1], weight.shape[0])
    return (out, x2.view(shape)) if fork else out


# --------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------
class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, key_mask, H, p, qkv_handoff=None):
        B, T, d3 = qkv.shape
        d = d3 // 3
        dh = d // H
        out = torch.empty((B, T, d), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
        rng, site, bits = None, 0, None
        if p > 0.0:
            rng = _rng.state(qkv.device)
            site = _rng.next_site()
            # one keep bit per score, drawn by the call and kept for backward (T = 250: 16 MB)
            bits = torch.empty(int(_lib.load().mar_attention_dropbits_words(B, T, H)), dtype=torch.int32, device=qkv.device)
        call("mar_attention_fwd", qkv.data_ptr(), _p(key_mask), out.data_ptr(), lse.data_ptr(), B, T, H, dh, _dt(qkv),
             float(p), _p(rng), site, _p(bits), _eng(), _stream())
        ctx.dims = (B, T, H, dh)
        ctx.p, ctx.eng = float(p), _eng()
        ctx.qkv_handoff = qkv_handoff
        ctx.save_for_backward(qkv, out, lse, key_mask, bits)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, key_mask, bits = ctx.saved_tensors
        B, T, H, dh = ctx.dims
        if dout.dtype != qkv.dtype:
            dout = _Cast.apply(dout, qkv.dtype)
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        nwork = int(_lib.load().mar_attention_bwd_work_floats(B, T, H, dh))
        work = torch.empty(nwork, dtype=torch.float32, device=qkv.device)
        # the in-projection's bias gradient = column sums of dqkv: taken where dQ / dK / dV leave the kernel
        colsum = _handoff_bias_target(ctx.qkv_handoff, 3 * H * dh, qkv.device)
        call("mar_attention_bwd", qkv.data_ptr(), _p(key_mask), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
             work.data_ptr(), dqkv.data_ptr(), _p(colsum), B, T, H, dh, _dt(qkv), ctx.p, _p(bits), ctx.eng, _stream())
        return dqkv, None, None, None, None


def attention(qkv: torch.Tensor, key_mask: Optional[torch.Tensor], num_heads: int, dropout_p: float,
              qkv_handoff: Optional[dict] = None) -> torch.Tensor:
    """qkv (B,T,3d) packed in-projection output → (B,T,d).  key_mask (B,T) uint8/bool, 1 = ignore key.
    qkv_handoff: the hand-off dict of the linear that produced qkv (its bias gradient is summed in the backward kernel)."""
    qkv = to_compute(qkv)
    if key_mask is not None:
        if key_mask.dtype == torch.bool:
            key_mask = key_mask.view(torch.uint8) if key_mask.is_contiguous() else key_mask.contiguous().view(torch.uint8)
        key_mask = key_mask.contiguous()
    return _Attention.apply(qkv, key_mask, int(num_heads), float(dropout_p), qkv_handoff)


# --------------------------------------------------------------------------------------
# layer norm
# --------------------------------------------------------------------------------------
def _row_map_args(segs):
    """ctypes arrays for mar_layernorm_*_mapped from [(t0, t1, Tout, tout0, ptr), ...] (kept alive by the caller)."""
    import ctypes
    n = len(segs)
    I64 = ctypes.c_int64 * n
    arrs = [I64(*[int(sg[i]) for sg in segs]) for i in range(4)]
    ptrs = (ctypes.c_void_p * n)(*[int(sg[4]) for sg in segs])
    return n, arrs, ptrs


def _strided_rows(t: torch.Tensor, D: int):
    """(B,T,D) tensor readable as rows of D with a batch pitch (no copy), or None."""
    if t.dim() == 3 and t.shape[2] == D and t.stride(2) == 1 and t.stride(1) == D and t.stride(0) % D == 0 \
            and t.stride(0) >= t.shape[1] * D and (t.data_ptr() % 16) == 0:
        return t.stride(0) // D
    return None


class _LayerNorm(torch.autograd.Function):
    """y = LN(x)·gamma + beta on (rows, D).  Plain: one (rows, D) output.  `place` = (buffer (B,Ttot,D), t_off, B, T): the
    output is written straight into buffer[:, t_off:t_off+T] and that (strided) view is returned — an extractor's final
    norm lands in its slice of the fused sequence, torch.cat along T costs nothing (models.py:419).  `split` =
    (B, Ttot, bounds): one contiguous (B, T_k, D) output per (t0, t1) of bounds — the fusion encoder's final norm writes
    every modality's slice (models.py:430) itself.  Backward reads the incoming gradient(s) through the same row map."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, zero_rows, need_grad, gamma_ref=None, beta_ref=None, place=None, split=None,
                x_handoff=None):
        rows, D = x.shape
        ctx.gamma_ref, ctx.beta_ref = gamma_ref, beta_ref
        # x = residual + dropout(linear): this op's backward also writes that linear's dz and sums its bias gradient
        ctx.x_handoff = x_handoff if (x_handoff is not None and x_handoff.get("drop") is not None and need_grad
                                      and place is None and split is None) else None
        if zero_rows is not None and need_grad:
            raise RuntimeError("layer_norm(zero_rows=...) is the eval-only nested-tensor zero fill; no backward")
        mean = torch.empty(rows, dtype=torch.float32, device=x.device) if need_grad else None
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if need_grad else None
        ctx.mode = "plain"
        if place is not None:
            buf, t_off, B, T = place
            assert rows == B * T and buf.dtype == x.dtype and buf.is_contiguous() and buf.shape[0] == B and buf.shape[2] == D
            segs = [(0, T, buf.shape[1], t_off, buf.data_ptr())]
            outs = buf[:, t_off:t_off + T]
            ctx.mode, ctx.Tin = "place", T
        elif split is not None:
            B, Ttot, bounds = split
            assert rows == B * Ttot
            outs = tuple(torch.empty((B, t1 - t0, D), dtype=x.dtype, device=x.device) for t0, t1 in bounds)
            segs = [(t0, t1, t1 - t0, 0, o.data_ptr()) for (t0, t1), o in zip(bounds, outs)]
            ctx.mode, ctx.Tin, ctx.bounds, ctx.B = "split", Ttot, list(bounds), B
        if ctx.mode == "plain":
            y = torch.empty_like(x)
            call("mar_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _p(mean), _p(rstd),
                 _p(zero_rows), rows, D, float(eps), _dt(x), _stream())
            outs = y
        else:
            n, arrs, ptrs = _row_map_args(segs)
            call("mar_layernorm_fwd_mapped", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, _p(mean), _p(rstd),
                 _p(zero_rows), rows, D, float(eps), _dt(x), n, ctx.Tin, arrs[0], arrs[1], arrs[2], arrs[3], ptrs, _stream())
        ctx.save_for_backward(x, gamma, mean, rstd)
        return outs

    @staticmethod
    def backward(ctx, *dys):
        x, gamma, mean, rstd = ctx.saved_tensors
        rows, D = x.shape
        segs = None
        keep = []
        if ctx.mode == "plain":
            dy = dys[0]
            if dy.dtype != x.dtype:
                dy = _Cast.apply(dy, x.dtype)
            dy = dy.contiguous()
        else:
            bounds = [(0, ctx.Tin)] if ctx.mode == "place" else ctx.bounds
            segs = []
            for (t0, t1), g in zip(bounds, dys):
                if g is None:                         # a slice nobody used downstream
                    g = torch.zeros((rows // ctx.Tin, t1 - t0, D), dtype=x.dtype, device=x.device)
                if g.dtype != x.dtype:
                    g = _Cast.apply(g, x.dtype)
                pitch = _strided_rows(g, D)
                if pitch is None:
                    g = g.contiguous()
                    pitch = g.shape[1]
                keep.append(g)
                segs.append((t0, t1, pitch, 0, g.data_ptr()))
        dx = torch.empty_like(x)
        sg, sb = _sink_target(ctx.gamma_ref, "ln"), _sink_target(ctx.beta_ref, "ln")
        sunk = sg is not None and sb is not None
        dgamma = sg if sunk else torch.zeros(D, dtype=torch.float32, device=x.device)
        dbeta = sb if sunk else torch.zeros(D, dtype=torch.float32, device=x.device)
        h = ctx.x_handoff
        if segs is None and h is not None:
            rng, site, p = h["drop"]
            dz = torch.empty_like(x)
            dbias = _handoff_bias_target(h, D, x.device)
            call("mar_layernorm_bwd_dropout", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                 dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), dz.data_ptr(), _p(dbias), rows, D, _dt(x), p,
                 rng.data_ptr(), site, _stream())
            h["dz"], h["dx_ptr"] = dz, dx.data_ptr()
        elif segs is None:
            call("mar_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                 dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), rows, D, _dt(x), _stream())
        else:
            n, arrs, ptrs = _row_map_args(segs)
            call("mar_layernorm_bwd_mapped", None, x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                 dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), rows, D, _dt(x), n, ctx.Tin, arrs[0], arrs[1], arrs[2],
                 arrs[3], ptrs, _stream())
        if sunk:
            _sunk(ctx.gamma_ref); _sunk(ctx.beta_ref)
            return (dx,) + (None,) * 10
        return (dx, dgamma, dbeta) + (None,) * 8


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
               zero_rows: Optional[torch.Tensor] = None, place=None, split=None, x_handoff: Optional[dict] = None):
    """nn.LayerNorm over the last dim.  place = (buffer (B,Ttot,D), t_off): write into buffer[:, t_off:t_off+T] and return
    that view; split = [(t0, t1), ...] (x must be (B, Ttot, D)): return one contiguous (B, t1-t0, D) tensor per slice."""
    shape = x.shape
    x2 = to_compute(x).reshape(-1, shape[-1])
    g = gamma if gamma.dtype == torch.float32 else gamma.float()
    b = beta if beta.dtype == torch.float32 else beta.float()
    need_grad = torch.is_grad_enabled() and (x2.requires_grad or g.requires_grad or b.requires_grad)
    gref, bref = (gamma if g is gamma else None), (beta if b is beta else None)
    if place is not None:
        buf, t_off = place
        return _LayerNorm.apply(x2, g, b, eps, zero_rows, need_grad, gref, bref, (buf, int(t_off), shape[0], shape[1]), None)
    if split is not None:
        return _LayerNorm.apply(x2, g, b, eps, zero_rows, need_grad, gref, bref, None, (shape[0], shape[1], [tuple(map(int, s_)) for s_ in split]))
    return _LayerNorm.apply(x2, g, b, eps, zero_rows, need_grad, gref, bref, None, None, x_handoff).view(shape)


# ---- placement of an extractor's final norm inside the fused sequence ------------------------------------------------
_place_hint = threading.local()


@contextlib.contextmanager
def place_final_norm(buffer: torch.Tensor, t_off: int, T: int):
    """While active, the next sequence encoder whose output is (B, T, D) with B, D of `buffer` writes its FINAL
    LayerNorm into buffer[:, t_off:t_off+T] (models.encoder_forward takes the hint once)."""
    _place_hint.value = (buffer, int(t_off), int(T))
    try:
        yield
    finally:
        _place_hint.value = None


def take_final_norm_placement(B: int, T: int, D: int, dtype: torch.dtype):
    hint = getattr(_place_hint, "value", None)
    if hint is None:
        return None
    buf, t_off, Th = hint
    _place_hint.value = None
    if Th != T or buf.shape[0] != B or buf.shape[2] != D or buf.dtype != dtype:
        return None
    return buf, t_off


# --------------------------------------------------------------------------------------
# pooling / masks / concat
# --------------------------------------------------------------------------------------
class _MeanPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        B, T, D = x.shape
        out = torch.empty((B, D), dtype=x.dtype, device=x.device)
        call("mar_meanpool_fwd", x.data_ptr(), out.data_ptr(), B, T, D, _dt(x), _stream())
        ctx.dims = (B, T, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, T, D = ctx.dims
        dout = dout.contiguous()
        dx = torch.empty((B, T, D), dtype=dout.dtype, device=dout.device)
        call("mar_meanpool_bwd", dout.data_ptr(), dx.data_ptr(), B, T, D, _dt(dout), _stream())
        return dx


def mean_pool(x: torch.Tensor) -> torch.Tensor:
    """(B,T,D) → (B,D): x.mean(dim=1) (SequenceAverageFeatures, models.py:105)."""
    return _MeanPool.apply(to_compute(x))


def rowzero_mask(x: torch.Tensor) -> torch.Tensor:
    """(B,T,D) → (B,T) uint8, 1 where the feature row sums to exactly 0 (models.py:421-422)."""
    x = to_compute(x)
    B, T, D = x.shape
    mask = torch.empty((B, T), dtype=torch.uint8, device=x.device)
    call("mar_rowzero_mask", x.data_ptr(), mask.data_ptr(), B * T, D, _dt(x), _stream())
    return mask


class _ConcatT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        B, _, D = xs[0].shape
        Ts = [x.shape[1] for x in xs]
        total = sum(Ts)
        out = torch.empty((B, total, D), dtype=xs[0].dtype, device=xs[0].device)
        off = 0
        for x, T in zip(xs, Ts):
            call("mar_concat_rows", x.data_ptr(), out.data_ptr(), B, T, total, off, D, _dt(x), 1, _stream())
            off += T
        ctx.Ts, ctx.B, ctx.D = Ts, B, D
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        total = sum(ctx.Ts)
        outs, off = [], 0
        for T in ctx.Ts:
            dx = torch.empty((ctx.B, T, ctx.D), dtype=g.dtype, device=g.device)
            call("mar_concat_rows", g.data_ptr(), dx.data_ptr(), ctx.B, T, total, off, ctx.D, _dt(g), 0, _stream())
            outs.append(dx)
            off += T
        return tuple(outs)


class _AliasConcat(torch.autograd.Function):
    """The blocks already ARE consecutive time slices of one (B, Ttot, D) buffer (every encoder's final norm wrote its
    slice there): torch.cat along T is the buffer itself, and its backward hands every block a strided view of the
    incoming gradient — no copy in either direction."""

    @staticmethod
    def forward(ctx, buffer, *xs):
        ctx.Ts = [x.shape[1] for x in xs]
        return buffer.detach()

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for T in ctx.Ts:
            outs.append(g[:, off:off + T])
            off += T
        return (None,) + tuple(outs)


def fused_slices_of(xs):
    """The shared buffer if `xs` are, in order, the consecutive time slices of ONE contiguous (B, Ttot, D) tensor."""
    tags = [getattr(x, "_mar_fused", None) for x in xs]
    if not xs or any(t is None for t in tags):
        return None
    buf = tags[0][0]
    off = 0
    for x, (b_, t_off) in zip(xs, tags):
        if b_ is not buf or t_off != off or x.dim() != 3 or x.data_ptr() != buf.data_ptr() + off * buf.shape[2] * buf.element_size():
            return None
        off += x.shape[1]
    return buf if off == buf.shape[1] else None


def concat_time(xs) -> torch.Tensor:
    """torch.cat(xs, dim=1) for (B,T_i,D) blocks (models.py:419)."""
    buf = fused_slices_of(xs)
    if buf is not None:
        return _AliasConcat.apply(buf, *xs)
    xs = [to_compute(x) for x in xs]
    return _ConcatT.apply(*xs)


class _SplitT(torch.autograd.Function):
    """All time slices of x (B,total,D) at once.  Backward writes every incoming gradient into ITS rows of ONE buffer:
    no zero-fill and no adds (one Function per slice gave autograd a zero-padded (B,total,D) tensor per slice to sum).
    A slice nobody used downstream arrives as None and its rows are zeroed."""

    @staticmethod
    def forward(ctx, x, *bounds):
        B, total, D = x.shape
        outs = []
        for t0, t1 in zip(bounds[0::2], bounds[1::2]):
            out = torch.empty((B, t1 - t0, D), dtype=x.dtype, device=x.device)
            call("mar_concat_rows", x.data_ptr(), out.data_ptr(), B, t1 - t0, total, t0, D, _dt(x), 0, _stream())
            outs.append(out)
        ctx.dims = (B, total, D)
        ctx.bounds = bounds
        ctx.dtype = x.dtype
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        B, total, D = ctx.dims
        dev = next(g for g in gs if g is not None).device
        covered = sorted((t0, t1) for t0, t1 in zip(ctx.bounds[0::2], ctx.bounds[1::2]))
        full = all(g is not None for g in gs) and covered[0][0] == 0 and covered[-1][1] == total and \
            all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:]))
        dx = (torch.empty if full else torch.zeros)((B, total, D), dtype=ctx.dtype, device=dev)
        for g, t0, t1 in zip(gs, ctx.bounds[0::2], ctx.bounds[1::2]):
            if g is None:
                continue
            if g.dtype != ctx.dtype:
                g = _Cast.apply(g, ctx.dtype)
            g = g.contiguous()
            call("mar_concat_rows", g.data_ptr(), dx.data_ptr(), B, t1 - t0, total, t0, D, _dt(g), 1, _stream())
        return (dx,) + (None,) * len(ctx.bounds)


def split_time(x: torch.Tensor, bounds) -> Tuple[torch.Tensor, ...]:
    """Contiguous copies of x[:, t0:t1] for every (t0, t1) of `bounds` (models.py:430)."""
    flat = [int(v) for b in bounds for v in b]
    return _SplitT.apply(x.contiguous(), *flat)


def slice_time(x: torch.Tensor, t0: int, t1: int) -> torch.Tensor:
    """Contiguous copy of x[:, t0:t1] (models.py:430)."""
    return split_time(x, [(t0, t1)])[0]


# --------------------------------------------------------------------------------------
# classifier loss
# --------------------------------------------------------------------------------------
class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, weight):
        B, C = logits.shape
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        call("mar_cross_entropy_fwd", logits.data_ptr(), labels.data_ptr(), _p(weight), loss.data_ptr(),
             dlogits.data_ptr(), None, B, C, _stream())
        ctx.save_for_backward(dlogits)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None


def cross_entropy(logits: torch.Tensor, labels: torch.Tensor, weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.CrossEntropyLoss (mean; optional class weights).  Rows with label < 0 are ignored."""
    _require_cuda(logits, "cross_entropy")
    if logits.dtype != torch.float32:
        logits = to_compute(logits, torch.float32)
    logits = logits.contiguous()
    labels = labels.to(device=logits.device, dtype=torch.int64).contiguous()
    if weight is not None:
        weight = weight.to(device=logits.device, dtype=torch.float32).contiguous()
    return _CrossEntropy.apply(logits, labels, weight)


class _FocalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, alpha, gamma):
        B, C = logits.shape
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        call("mar_focal_loss_fwd", logits.data_ptr(), labels.data_ptr(), _p(alpha), float(gamma), loss.data_ptr(),
             dlogits.data_ptr(), B, C, _stream())
        ctx.save_for_backward(dlogits)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None, None


def focal_loss(logits: torch.Tensor, labels: torch.Tensor, alpha: Optional[torch.Tensor] = None, gamma: float = 0.0) -> torch.Tensor:
    """Multi-class focal loss with 'mean' reduction (train_multimodal.py:494-510).  Rows with label < 0 are ignored."""
    _require_cuda(logits, "focal_loss")
    if logits.dtype != torch.float32:
        logits = to_compute(logits, torch.float32)
    logits = logits.contiguous()
    labels = labels.to(device=logits.device, dtype=torch.int64).contiguous()
    if alpha is not None:
        alpha = alpha.to(device=logits.device, dtype=torch.float32).contiguous()
    return _FocalLoss.apply(logits, labels, alpha, float(gamma))


def label_weight_sum(labels: torch.Tensor, weight: Optional[torch.Tensor], out: torch.Tensor) -> None:
    """out[0] = Σ over rows with label >= 0 of weight[label] (their count when weight is None): the denominator of
    nn.CrossEntropyLoss(reduction='mean') (models.py:232-263), on the device."""
    _require_cuda(labels, "label_weight_sum")
    if weight is not None:
        weight = weight.to(device=labels.device, dtype=torch.float32).contiguous()
    C = weight.numel() if weight is not None else (1 << 40)
    call("mar_label_weight_sum", labels.data_ptr(), _p(weight), out.data_ptr(), labels.numel(), C, _stream())


def argmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """argmax over classes on the device (trainer.py:170, :726)."""
    logits = to_compute(logits, torch.float32).contiguous()
    B, C = logits.shape
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    preds = torch.empty(B, dtype=torch.int64, device=logits.device)
    labels = torch.full((B,), -1, dtype=torch.int64, device=logits.device)
    call("mar_cross_entropy_fwd", logits.data_ptr(), labels.data_ptr(), None, loss.data_ptr(), None, preds.data_ptr(),
         B, C, _stream())
    return preds


# --------------------------------------------------------------------------------------
# recurrences
# --------------------------------------------------------------------------------------
class _GRU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, w_hh, b_hh, need_grad):
        B, T, H3 = gi.shape
        H = H3 // 3
        cd = gi.dtype
        wc, _ = compute_weight(w_hh, cd, False)
        hseq = torch.empty((B, T, H), dtype=cd, device=gi.device)
        saved = torch.empty((B, T, 5 * H), dtype=torch.float32, device=gi.device) if need_grad else None
        hprev = torch.empty((B, T, H), dtype=cd, device=gi.device) if need_grad else None
        work = torch.empty(B * 5 * H, dtype=torch.float32, device=gi.device)
        b32 = b_hh.detach().float().contiguous()
        call("mar_gru_fwd", gi.data_ptr(), wc.data_ptr(), b32.data_ptr(), hseq.data_ptr(), _p(hprev), _p(saved),
             work.data_ptr(), B, T, H, _dt(gi), _eng(), _stream())
        ctx.dims = (B, T, H)
        ctx.wc = wc
        ctx.eng = _eng()
        ctx.save_for_backward(hprev, saved)
        return hseq

    @staticmethod
    def backward(ctx, dhseq):
        hprev, saved = ctx.saved_tensors
        B, T, H = ctx.dims
        cd = hprev.dtype
        hseq = hprev
        st = _stream()
        if dhseq.dtype != cd:
            dhseq = _Cast.apply(dhseq, cd)
        dhseq = dhseq.contiguous()
        dgi = torch.empty((B, T, 3 * H), dtype=cd, device=hseq.device)
        dgh = torch.empty((B, T, 3 * H), dtype=cd, device=hseq.device)
        work = torch.empty(B * 5 * H, dtype=torch.float32, device=hseq.device)
        call("mar_gru_bwd", dhseq.data_ptr(), saved.data_ptr(), ctx.wc.data_ptr(), dgi.data_ptr(),
             dgh.data_ptr(), work.data_ptr(), B, T, H, _dt(hseq), ctx.eng, st)
        dw = torch.empty((3 * H, H), dtype=torch.float32, device=hseq.device)
        call("mar_linear_wgrad", dgh.data_ptr(), hprev.data_ptr(), H, dw.data_ptr(), B * T, 3 * H, H, _dt(dgh), 0,
             ctx.eng, st)
        db = torch.zeros(3 * H, dtype=torch.float32, device=hseq.device)
        call("mar_linear_bwd_epilogue", dgh.data_ptr(), None, None, db.data_ptr(), B * T, 3 * H, _dt(dgh), _dt(dgh), 0,
             0.0, None, 0, 0, st)
        return dgi, dw, db, None


def gru(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh) -> torch.Tensor:
    """1-layer batch_first GRU with h0 = 0 → (B,T,H) (nn.GRU as used at models.py:110,122)."""
    gi = linear(x, w_ih, b_ih)
    need = torch.is_grad_enabled() and (gi.requires_grad or w_hh.requires_grad or b_hh.requires_grad)
    return _GRU.apply(gi.contiguous(), w_hh, b_hh, need)


class _LSTM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, w_hh, b_hh, need_grad):
        B, T, H4 = gi.shape
        H = H4 // 4
        cd = gi.dtype
        wc, _ = compute_weight(w_hh, cd, False)
        hseq = torch.empty((B, T, H), dtype=cd, device=gi.device)
        saved = torch.empty((B, T, 5 * H), dtype=torch.float32, device=gi.device) if need_grad else None
        hprev = torch.empty((B, T, H), dtype=cd, device=gi.device) if need_grad else None
        work = torch.empty(B * 5 * H, dtype=torch.float32, device=gi.device)
        b32 = b_hh.detach().float().contiguous()
        call("mar_lstm_fwd", gi.data_ptr(), wc.data_ptr(), b32.data_ptr(), hseq.data_ptr(), _p(hprev), _p(saved),
             work.data_ptr(), B, T, H, _dt(gi), _eng(), _stream())
        ctx.dims = (B, T, H)
        ctx.wc = wc
        ctx.eng = _eng()
        ctx.save_for_backward(hprev, saved)
        return hseq

    @staticmethod
    def backward(ctx, dhseq):
        hprev, saved = ctx.saved_tensors
        B, T, H = ctx.dims
        cd = hprev.dtype
        st = _stream()
        if dhseq.dtype != cd:
            dhseq = _Cast.apply(dhseq, cd)
        dhseq = dhseq.contiguous()
        dg = torch.empty((B, T, 4 * H), dtype=cd, device=saved.device)
        work = torch.empty(B * 2 * H, dtype=torch.float32, device=saved.device)
        call("mar_lstm_bwd", dhseq.data_ptr(), saved.data_ptr(), ctx.wc.data_ptr(), dg.data_ptr(), work.data_ptr(),
             B, T, H, _dt(hprev), ctx.eng, st)
        dw = torch.empty((4 * H, H), dtype=torch.float32, device=saved.device)
        call("mar_linear_wgrad", dg.data_ptr(), hprev.data_ptr(), H, dw.data_ptr(), B * T, 4 * H, H, _dt(dg), 0,
             ctx.eng, st)
        db = torch.zeros(4 * H, dtype=torch.float32, device=saved.device)
        call("mar_linear_bwd_epilogue", dg.data_ptr(), None, None, db.data_ptr(), B * T, 4 * H, _dt(dg), _dt(dg), 0,
             0.0, None, 0, 0, st)
        return dg, dw, db, None


def lstm(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh) -> torch.Tensor:
    """1-layer batch_first LSTM with (h0,c0) = 0 → (B,T,H) (train_video_rnn.py:94-106)."""
    gi = linear(x, w_ih, b_ih)
    need = torch.is_grad_enabled() and (gi.requires_grad or w_hh.requires_grad or b_hh.requires_grad)
    return _LSTM.apply(gi.contiguous(), w_hh, b_hh, need)
