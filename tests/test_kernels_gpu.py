"""GPU: every kernel through the C ABI against the CPU oracle's primitive of the same op.

Tolerances (BASELINE.json north_star): fp32 mode 1e-4; bf16 mode ~1e-2 relative against fp32
(measured as ||a-b||/||b||; written in tests/helpers.py)."""
import math

import pytest
import torch

import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import ops
from oracle import oracle as O
from tests.helpers import BF16_TOL, FP32_TOL, assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    O.DROPOUT_ENABLED = False
    yield


def bf16_round(t):
    return t.to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------
# linear: forward / dgrad / wgrad, all engines
# ------------------------------------------------------------------------------------------
LINEAR_SHAPES = [(128, 256, 64), (200, 768, 768), (1000, 2304, 768), (333, 768, 2048), (64, 2, 256), (37, 256, 512),
                 (4096, 1536, 512), (256, 512, 1536), (130, 72, 136)]


def _linear_ref(x, w, b, res, relu_pre, relu_post):
    y = O.linear(x, w, b)
    if relu_pre or relu_post:
        y = O.relu(y)
    if res is not None:
        y = y + res
    return y


@pytest.mark.parametrize("M,N,K", LINEAR_SHAPES)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_linear_fwd_bwd(M, N, K, mode):
    x = torch.randn(M, K)
    w = torch.randn(N, K) / math.sqrt(K)
    b = torch.randn(N)
    res = torch.randn(M, N)
    gy = torch.randn(M, N)
    if mode == "bf16":
        x, res, gy = bf16_round(x), bf16_round(res), bf16_round(gy)
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    for relu_pre, use_res in ((False, False), (True, False), (False, True)):
        xr, wr, br = x.clone().requires_grad_(True), (bf16_round(w) if mode == "bf16" else w.clone()).requires_grad_(True), b.clone().requires_grad_(True)
        rr = res.clone().requires_grad_(True) if use_res else None
        yr = _linear_ref(xr, wr, br, rr, relu_pre, False)
        yr.backward(gy)
        xg = x.to(DEV).requires_grad_(True)
        wg = w.to(DEV).requires_grad_(True)
        bg = b.to(DEV).requires_grad_(True)
        rg = res.to(DEV).requires_grad_(True) if use_res else None
        with mar.precision(mode):
            y = ops.linear(xg, wg, bg, residual=rg, relu_pre=relu_pre)
            y.backward(gy.to(DEV).to(y.dtype))
        what = f"linear {mode} M={M} N={N} K={K} relu={relu_pre} res={use_res}"
        assert_close(y.float().cpu(), yr, tol, what + " y")
        assert_close(xg.grad.cpu(), xr.grad, tol, what + " dx")
        assert_close(wg.grad.cpu(), wr.grad, tol, what + " dw")
        assert_close(bg.grad.cpu(), br.grad, tol, what + " db")
        if use_res:
            assert_close(rg.grad.cpu(), gy, 1e-6, what + " dres")


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 128), (200, 768, 768), (8000, 2304, 768), (1024, 768, 2048),
                                   (130, 72, 136), (4096, 1536, 512), (640, 2048, 768)])
def test_tcgen05_gemm_engine(M, N, K):
    """The tcgen05/TMA kernel specifically (engine='tensor' raises if it cannot run): fwd, dgrad, wgrad
    against an fp32 matmul of the same bf16 operands."""
    x = bf16_round(torch.randn(M, K))
    w = bf16_round(torch.randn(N, K) / math.sqrt(K))
    b = torch.randn(N)
    gy = bf16_round(torch.randn(M, N))
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.linear(xr, wr, br)
    yr.backward(gy)
    xg, wg, bg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    with mar.precision("bf16"), mar.engine("tensor"):
        y = ops.linear(xg, wg, bg)
        assert ops.last_engine() == "tensor"
        y.backward(gy.to(DEV).to(y.dtype))
    assert_close(y.float().cpu(), yr, 4e-3, "tcgen05 fwd")          # one bf16 rounding of the output
    assert_close(xg.grad.cpu(), xr.grad, 4e-3, "tcgen05 dgrad")
    assert_close(wg.grad.cpu(), wr.grad, 1e-4, "tcgen05 wgrad (fp32 out)")
    assert_close(bg.grad.cpu(), br.grad, 1e-4, "bias grad")


def test_tcgen05_matches_simt_bitwise_dropout_mask():
    """Both GEMM engines draw the same dropout mask (same counter-hash function), forward and backward."""
    M, N, K = 512, 768, 256
    x = bf16_round(torch.randn(M, K)).to(DEV)
    w = (torch.randn(N, K) / math.sqrt(K)).to(DEV)
    outs = {}
    for eng in ("simt", "tensor"):
        ops.manual_seed(123)
        with mar.precision("bf16"), mar.engine(eng):
            outs[eng] = ops.linear(x, w, None, dropout_p=0.3).float()
    za, zb = outs["simt"] == 0, outs["tensor"] == 0
    assert torch.equal(za, zb)
    frac = float(za.float().mean())
    assert abs(frac - 0.3) < 0.01
    assert_close(outs["tensor"], outs["simt"], 4e-3, "dropout epilogue values")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_linear_dropout_backward_uses_same_mask(mode):
    M, N, K = 256, 512, 128
    x = torch.randn(M, K, device=DEV, requires_grad=True)
    w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).requires_grad_(True)
    for kw in (dict(relu_pre=True, dropout_p=0.5), dict(dropout_p=0.3, relu_post=True), dict(dropout_p=0.1)):
        with mar.precision(mode):
            y = ops.linear(x, w, None, **kw)
            (gx,) = torch.autograd.grad(y.float().sum(), x, retain_graph=False)
        # d(sum y)/dx = mask_scale @ W  => recompute from the observed zero pattern of y
        p = kw["dropout_p"]
        if kw.get("relu_pre") or kw.get("relu_post"):
            f = (y != 0).float() / (1 - p)
        else:
            f = (y != 0).float() / (1 - p)       # y == 0 exactly only where dropped (measure-zero otherwise)
        wq = w.detach().to(y.dtype).float()
        expect = f @ wq
        assert_close(gx.float(), expect, FP32_TOL if mode == "fp32" else BF16_TOL, f"dropout bwd {kw}")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_ffn_deferred_activation_backward_matches_the_unfused_path(mode, p):
    """linear1 (ReLU + dropout) → linear2 with linear1's activation backward applied inside linear2's dgrad GEMM
    (defer_act / x_act_scale) against the same two linears with the separate epilogue-backward pass: same dropout
    mask (same seed and sites), same forward, same gradients."""
    Mr, d, ff = 300, 256, 512
    x = torch.randn(Mr, d, device=DEV)
    w1, b1 = torch.randn(ff, d, device=DEV) / d ** 0.5, torch.randn(ff, device=DEV) * 0.1
    w2, b2 = torch.randn(d, ff, device=DEV) / ff ** 0.5, torch.randn(d, device=DEV) * 0.1
    go = torch.randn(Mr, d, device=DEV)
    res = {}
    for fused in (False, True):
        mar.manual_seed(99)
        leaves = [t.clone().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
        xx, a1, c1, a2, c2 = leaves
        with mar.precision(mode):
            hid, xr = ops.linear(xx, a1, c1, relu_pre=True, dropout_p=p, fork=True, defer_act=fused)
            y = ops.linear(hid, a2, c2, residual=xr, x_act_scale=(1.0 / (1.0 - p)) if fused else None)
            y.backward(go.to(y.dtype))
        res[fused] = [y.detach().float()] + [t.grad.float() for t in leaves]
    tol = 1e-5 if mode == "fp32" else 2e-2
    for name, a, b in zip(("y", "dx", "dW1", "db1", "dW2", "db2"), res[True], res[False]):
        assert_close(a, b, tol, f"deferred activation backward: {name}")
    assert float(res[True][1].abs().max()) > 0


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("p", [0.0, 0.3])
def test_adaptor_with_fused_mean_pool_matches_linear_then_mean_pool(mode, p):
    """Linear → Dropout → ReLU → mean over T as one node (pool_T) against linear followed by ops.mean_pool."""
    B, T, d, n = 5, 37, 256, 192
    x = torch.randn(B, T, d, device=DEV)
    w, b = torch.randn(n, d, device=DEV) / d ** 0.5, torch.randn(n, device=DEV) * 0.1
    go = torch.randn(B, n, device=DEV)
    res = {}
    for fused in (False, True):
        mar.manual_seed(7)
        xx, ww, bb = [t.clone().requires_grad_(True) for t in (x, w, b)]
        with mar.precision(mode):
            if fused:
                y = ops.linear(xx, ww, bb, dropout_p=p, relu_post=True, pool_T=T)
            else:
                y = ops.mean_pool(ops.linear(xx, ww, bb, dropout_p=p, relu_post=True))
            y.backward(go.to(y.dtype))
        res[fused] = [y.detach().float(), xx.grad.float(), ww.grad.float(), bb.grad.float()]
    tol = 1e-5 if mode == "fp32" else 2e-2
    for name, a, b_ in zip(("pooled", "dx", "dW", "db"), res[True], res[False]):
        assert_close(a, b_, tol, f"fused mean-pool adaptor: {name}")
    with pytest.raises(ValueError):
        ops.linear(x, w, b, relu_post=True, pool_T=T + 1)


# ------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------
def _attn_ref(qkv, mask, H):
    B, T, d3 = qkv.shape
    d = d3 // 3
    dh = d // H
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    sp = lambda t: t.reshape(B, T, H, dh).permute(0, 2, 1, 3)
    s = torch.matmul(sp(q), sp(k).transpose(-1, -2)) / math.sqrt(dh)
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = O.softmax_lastdim_safe(s)
    return torch.matmul(p, sp(v)).permute(0, 2, 1, 3).reshape(B, T, d)


@pytest.mark.parametrize("B,T,H,dh", [(2, 50, 8, 96), (3, 37, 4, 64), (2, 130, 2, 32), (1, 314, 8, 96), (2, 64, 8, 96), (2, 33, 1, 128),
                                      # the 1-token-per-modality sequences of AveragedFeaturesTransformerFusion (models.py:482-503)
                                      (3, 2, 8, 96), (2, 1, 4, 64), (4, 5, 8, 96)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("masked", [False, True])
def test_attention_fwd_bwd(B, T, H, dh, mode, masked):
    d = H * dh
    qkv = torch.randn(B, T, 3 * d)
    go = torch.randn(B, T, d)
    if mode == "bf16":
        qkv, go = bf16_round(qkv), bf16_round(go)
    mask = None
    if masked:
        mask = torch.rand(B, T) < 0.3
        mask[0, : T // 2] = False
        mask[-1, T // 3:] = True          # a long masked tail
    qr = qkv.clone().requires_grad_(True)
    outr = _attn_ref(qr, mask, H)
    outr.backward(go)
    qg = qkv.to(DEV).requires_grad_(True)
    with mar.precision(mode):
        out = ops.attention(qg, None if mask is None else mask.to(DEV), H, 0.0)
        out.backward(go.to(DEV).to(out.dtype))
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    assert_close(out.float().cpu(), outr, tol, "attention out")
    assert_close(qg.grad.float().cpu(), qr.grad, tol, "attention dqkv")


@pytest.mark.timeout(120)
@pytest.mark.parametrize("B,T", [(48, 314), (80, 250), (12, 700), (40, 129), (6, 1100)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_attention_forward_kernels_over_many_ctas(B, T, p, monkeypatch):
    """Both tcgen05 forward kernels (one query tile per CTA up to T = 512, the pair kernel beyond: T = 700 has an odd
    number of query tiles, so every third CTA runs without its second warp-group) on grids of several waves, with a
    key-padding mask per batch element and dropout, against the mma.sync engine on the same keep bits."""
    H, dh = 8, 96
    d = H * dh
    qkv = torch.randn(B, T, 3 * d, device=DEV).to(torch.bfloat16)
    mask = torch.rand(B, T, device=DEV) < 0.15
    mask[:, 0] = False
    mask[1] = False
    outs = {}
    for eng in ("tc", "mma"):
        monkeypatch.setenv("MAR_ATTN_MMA", "1" if eng == "mma" else "0")
        mar.manual_seed(77)
        with mar.precision("bf16"), torch.no_grad():
            outs[eng] = ops.attention(qkv, mask, H, p).float()
    torch.cuda.synchronize()
    assert torch.isfinite(outs["tc"]).all()
    assert_close(outs["tc"], outs["mma"], 1e-2, "tcgen05 forward vs mma.sync engine")
    worst = float((outs["tc"] - outs["mma"]).abs().amax(dim=(1, 2)).max())      # no single item may be off (a stale tile)
    assert worst < 0.1, worst


@pytest.mark.parametrize("T,H", [(1024, 2), (4096, 1)])
@pytest.mark.parametrize("masked", [False, True])
def test_attention_long_sequences_against_oracle(T, H, masked):
    """BASELINE config 5's fused lengths (2T up to 4096): the tcgen05 forward and backward against the CPU oracle,
    not only against this repo's other engine (nothing above T = 314 was checked against the oracle)."""
    B, dh = 1, 96
    d = H * dh
    qkv = bf16_round(torch.randn(B, T, 3 * d))
    go = bf16_round(torch.randn(B, T, d))
    mask = None
    if masked:
        mask = torch.rand(B, T) < 0.3
        mask[0, T - T // 5:] = True
    qr = qkv.clone().requires_grad_(True)
    outr = _attn_ref(qr, mask, H)
    outr.backward(go)
    qg = qkv.to(DEV).requires_grad_(True)
    with mar.precision("bf16"), mar.engine("tensor"):
        out = ops.attention(qg, None if mask is None else mask.to(DEV), H, 0.0)
        out.backward(go.to(DEV).to(out.dtype))
    assert_close(out.float().cpu(), outr, BF16_TOL, f"attention out T={T}")
    for name, sl in (("dQ", slice(0, d)), ("dK", slice(d, 2 * d)), ("dV", slice(2 * d, 3 * d))):
        assert_close(qg.grad.float().cpu()[..., sl], qr.grad[..., sl], BF16_TOL, f"attention {name} T={T}")


@pytest.mark.parametrize("B,T,H,dh", [(2, 128, 2, 96), (2, 129, 2, 64), (1, 256, 8, 96), (2, 700, 2, 128), (1, 1100, 1, 96),
                                      (5, 17, 3, 64)])
@pytest.mark.parametrize("p", [0.0, 0.2])
def test_attention_bwd_tcgen05_matches_mma_engine(B, T, H, dh, p, monkeypatch):
    """The one-kernel tcgen05 backward (fp32 dQ reduction across key tiles) against the mma.sync engine on the same
    forward, the same dropout mask (both regenerate it from the counter hash) and a key-padding mask."""
    d = H * dh
    qkv = (torch.randn(B, T, 3 * d, device=DEV)).to(torch.bfloat16)
    go = torch.randn(B, T, d, device=DEV).to(torch.bfloat16)
    mask = torch.rand(B, T, device=DEV) < 0.2
    mask[0, : max(1, T // 2)] = False
    mask[-1, T - T // 4:] = True
    grads = {}
    for eng in ("tc", "mma"):
        monkeypatch.setenv("MAR_ATTN_BWD_MMA", "1" if eng == "mma" else "0")
        mar.manual_seed(1234)
        q = qkv.clone().requires_grad_(True)
        with mar.precision("bf16"):
            out = ops.attention(q, mask, H, p)
            out.backward(go)
        grads[eng] = q.grad.float()
    assert torch.isfinite(grads["tc"]).all()
    for name, sl in (("dQ", slice(0, d)), ("dK", slice(d, 2 * d)), ("dV", slice(2 * d, 3 * d))):
        assert_close(grads["tc"][..., sl], grads["mma"][..., sl], 6e-3, f"tcgen05 vs mma backward {name}")
    assert float(grads["tc"][-1, T - T // 4:, d:].abs().max()) == 0.0       # masked keys get no dK / dV


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_attention_fully_masked_rows_give_zero(mode):
    """All keys masked (all-EMPTY batch): P = 0, O = 0, finite grads (torch 2.11 safe softmax)."""
    B, T, H, dh = 2, 40, 8, 96
    qkv = torch.randn(B, T, 3 * H * dh, device=DEV, requires_grad=True)
    mask = torch.zeros(B, T, dtype=torch.bool, device=DEV)
    mask[1] = True
    with mar.precision(mode):
        out = ops.attention(qkv, mask, H, 0.0)
        out.float().sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(qkv.grad).all()
    assert float(out[1].abs().max()) == 0.0
    assert float(qkv.grad[1].abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_attention_dropout_statistics_and_grad_consistency(mode):
    """With V = 1 the output is rowsum(P~) whose mean is 1; d(sum O)/dV = P~ᵀ·1 shares the forward mask."""
    B, T, H, dh = 2, 128, 4, 64
    d = H * dh
    qkv = torch.randn(B, T, 3 * d, device=DEV)
    qkv[..., 2 * d:] = 1.0
    qkv.requires_grad_(True)
    with mar.precision(mode):
        out = ops.attention(qkv, None, H, 0.25)
        out.float().sum().backward()
    o = out.float()
    assert abs(float(o.mean()) - 1.0) < 0.02
    assert float(o.std()) > 0.01                       # masks are actually applied
    # column sums of P~ (via dV) must add up to the row sums of P~ (via O)
    dv = qkv.grad[..., 2 * d:].float()                 # (B,T,d): dV[k, c] = Σ_q P~[q,k]
    total_from_dv = dv.reshape(B, T, H, dh)[..., 0].sum(dim=1)      # (B,H)
    total_from_o = o.reshape(B, T, H, dh)[..., 0].sum(dim=1)
    assert_close(total_from_dv, total_from_o, 2e-2 if mode == "bf16" else 1e-4, "dropout mask fwd/bwd consistency")


@pytest.mark.parametrize("p", [0.1, 0.25, 0.5, 0.9])
def test_attention_keep_bits_are_bernoulli_and_shared_by_every_engine(p):
    """The keep bits one attention call draws (include/mar.h, mar_attention_fwd): the fraction of set bits is the
    quantised keep probability m/256 with m = round((1-p)*256), neighbouring bits and neighbouring rows are
    uncorrelated, another site / step draws other bits, and all three engines — SIMT, mma.sync, tcgen05 — produce the
    same P~ from them (same outputs up to their arithmetic)."""
    from multimodalaggressionrecognition_b200 import _lib
    from multimodalaggressionrecognition_b200._lib import call
    B, T, H, dh = 3, 200, 4, 64
    d = H * dh
    qkv = torch.randn(B, T, 3 * d, device=DEV).to(torch.bfloat16)
    nwords = int(_lib.load().mar_attention_dropbits_words(B, T, H))
    W = 4 * ((T + 127) // 128)
    assert nwords == (B * H * T + 256) * W
    st = torch.cuda.current_stream().cuda_stream
    rng = ops._rng.state(torch.device(DEV))
    outs, bits_of = {}, {}
    for name, eng, env in (("simt", _lib.ENGINE_SIMT, None), ("tc", _lib.ENGINE_TCGEN05, None)):
        bits = torch.zeros(nwords, dtype=torch.int32, device=DEV)
        out = torch.empty(B, T, d, device=DEV, dtype=torch.bfloat16)
        lse = torch.empty(B, H, T, device=DEV)
        call("mar_attention_fwd", qkv.data_ptr(), None, out.data_ptr(), lse.data_ptr(), B, T, H, dh, _lib.MAR_BF16, p,
             rng.data_ptr(), 77, bits.data_ptr(), eng, st)
        outs[name], bits_of[name] = out.float(), bits
    assert torch.equal(bits_of["simt"], bits_of["tc"])
    assert_close(outs["tc"], outs["simt"], 2e-2, "tcgen05 vs SIMT under the same keep bits")
    words = bits_of["tc"][: B * H * T * W].view(B * H * T, W).cpu().numpy().astype("uint32")
    import numpy as np
    bitmat = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(B * H * T, W * 32)[:, :T].astype(np.float64)
    m = max(1, min(255, int((1 - p) * 256 + 0.5)))
    frac, n = bitmat.mean(), bitmat.size
    assert abs(frac - m / 256) < 5 * (0.25 / n) ** 0.5 + 1e-4, (frac, m / 256)
    c = bitmat - frac
    for a, b_ in ((c[:, 1:], c[:, :-1]), (c[1:], c[:-1]), (c[:, 32:], c[:, :-32])):   # next key, next row, same bit of the next word
        corr = float((a * b_).mean() / max(frac * (1 - frac), 1e-9))
        assert abs(corr) < 6 / n ** 0.5 + 2e-3, corr
    other = torch.zeros(nwords, dtype=torch.int32, device=DEV)
    call("mar_attention_fwd", qkv.data_ptr(), None, out.data_ptr(), lse.data_ptr(), B, T, H, dh, _lib.MAR_BF16, p,
         rng.data_ptr(), 78, other.data_ptr(), _lib.ENGINE_TCGEN05, st)
    assert not torch.equal(other, bits_of["tc"])
    # the scale of the kept scores makes the expectation exact: V = 1 -> O = rowsum(P~), mean 1
    qkv1 = qkv.clone()
    qkv1[..., 2 * d:] = 1.0
    call("mar_attention_fwd", qkv1.data_ptr(), None, out.data_ptr(), lse.data_ptr(), B, T, H, dh, _lib.MAR_BF16, p,
         rng.data_ptr(), 79, other.data_ptr(), _lib.ENGINE_TCGEN05, st)
    assert abs(float(out.float().mean()) - 1.0) < 0.03


# ------------------------------------------------------------------------------------------
# layer norm / pooling / masks / concat / loss / adam
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("p", [0.0, 0.2])
@pytest.mark.parametrize("B,T,d,H,ff", [(3, 70, 192, 2, 256), (2, 250, 768, 8, 2048), (5, 33, 128, 2, 512)])
def test_backward_handoffs_match_the_separate_passes(mode, p, B, T, d, H, ff):
    """One post-norm encoder layer, backward with the hand-offs (norm1 / norm2 write out_proj's / linear2's dropout backward
    and sum their bias gradients, the attention kernel sums the in-projection's bias gradient, linear2's dgrad GEMM sums
    linear1's) against the same layer with every linear launching its own epilogue-backward pass: same seed, same sites,
    hence the same masks — the same gradients."""
    from multimodalaggressionrecognition_b200.models import encoder_layer_forward
    torch.manual_seed(5)
    layer = torch.nn.TransformerEncoderLayer(d, H, dim_feedforward=ff, dropout=p, batch_first=True).to(DEV)
    layer.train()
    x = torch.randn(B * T, d, device=DEV)
    go = torch.randn(B * T, d, device=DEV)
    key_mask = torch.zeros(B, T, dtype=torch.uint8, device=DEV)
    key_mask[0, T - 7:] = 1
    res = {}
    for on in (False, True):
        mar.manual_seed(31)
        ops.clear_weight_cache()          # both runs cast the weights (the launch counts below are compared)
        for q in layer.parameters():
            q.grad = None
        xx = x.clone().requires_grad_(True)
        with mar.precision(mode), ops.handoffs(on):
            n0 = ops.launch_count()
            y = encoder_layer_forward(layer, ops.to_compute(xx), B, T, key_mask)
            y.backward(go.to(y.dtype))
            launches = ops.launch_count() - n0
        res[on] = ([y.detach().float(), xx.grad.float()] + [q.grad.float().clone() for q in layer.parameters()], launches)
    names = ["y", "dx"] + [n for n, _ in layer.named_parameters()]
    # fp32: same numbers, only summation orders differ.  bf16: the fused kernel masks the fp32 dx and rounds once, the
    # separate pass masks the bf16-rounded dx and rounds again — the GEMM operands differ by one bf16 rounding
    tol = 2e-5 if mode == "fp32" else 6e-3
    for name, a, b in zip(names, res[True][0], res[False][0]):
        assert_close(a, b, tol, f"hand-off backward: {name}")
        assert float(b.abs().max()) > 0, name
    assert torch.equal(res[True][0][0], res[False][0][0])
    # passes that are gone: out_proj's and linear2's dropout backward (dropout on), and — on the tensor-core engines, which
    # sum inside their own epilogues — the bias-only passes of the in-projection and linear1
    assert res[False][1] - res[True][1] >= (2 if p > 0 else 0) + (2 if mode == "bf16" else 0), (res[False][1], res[True][1])


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("rows,D", [(100, 768), (7, 512), (3000, 1280), (33, 64)])
def test_layernorm_backward_with_fused_dropout_backward(rows, D, mode):
    """mar_layernorm_bwd_dropout against mar_layernorm_bwd followed by mar_linear_bwd_epilogue (the same mask function):
    dx bit-identical, the same elements dropped, dz / dbias / dgamma / dbeta to rounding and summation order."""
    from multimodalaggressionrecognition_b200 import _lib
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    code = 0 if mode == "fp32" else 1
    x = (torch.randn(rows, D, device=DEV) * 2 + 0.5).to(dt)
    dy = torch.randn(rows, D, device=DEV).to(dt)
    gamma = torch.randn(D, device=DEV)
    mean = x.float().mean(1).contiguous()
    rstd = (x.float().var(1, unbiased=False) + 1e-5).rsqrt().contiguous()
    rng = torch.tensor([1234, 7], dtype=torch.int64, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    site, p = 17, 0.3
    z = lambda *s_: torch.zeros(*s_, dtype=torch.float32, device=DEV)
    dx0, dg0, db0 = torch.empty_like(x), z(D), z(D)
    _lib.call("mar_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
              dx0.data_ptr(), dg0.data_ptr(), db0.data_ptr(), rows, D, code, st)
    dz0, dbias0 = torch.empty_like(x), z(D)
    _lib.call("mar_linear_bwd_epilogue", dx0.data_ptr(), None, dz0.data_ptr(), dbias0.data_ptr(), rows, D, code, code, 2, p,
              rng.data_ptr(), site, 0, st)
    dx1, dg1, db1, dz1, dbias1 = torch.empty_like(x), z(D), z(D), torch.full_like(x, float("nan")), z(D)
    _lib.call("mar_layernorm_bwd_dropout", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
              dx1.data_ptr(), dg1.data_ptr(), db1.data_ptr(), dz1.data_ptr(), dbias1.data_ptr(), rows, D, code, p,
              rng.data_ptr(), site, st)
    torch.cuda.synchronize()
    assert torch.equal(dx0, dx1)
    assert not torch.isnan(dz1.float()).any()
    kept = float((dz1 != 0).float().mean())
    assert abs(kept - (1 - p)) < 0.02, kept
    if mode == "fp32":
        assert torch.equal(dz0, dz1)
    else:   # the separate pass masks the bf16-ROUNDED dx, the fused kernel the fp32 value: one rounding apart at most
        assert torch.equal(dz0 != 0, dz1 != 0)
        assert_close(dz1.float(), dz0.float(), 6e-3, "dz")
    assert_close(dbias1, dbias0, 2e-3 if mode == "bf16" else 1e-5, "dbias")
    assert_close(dg1, dg0, 1e-5, "dgamma")
    assert_close(db1, db0, 1e-5, "dbeta")
    # dbias = NULL is accepted
    _lib.call("mar_layernorm_bwd_dropout", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
              dx1.data_ptr(), dg1.data_ptr(), db1.data_ptr(), dz1.data_ptr(), None, rows, D, code, p, rng.data_ptr(), site, st)
    torch.cuda.synchronize()


@pytest.mark.parametrize("rows,D", [(100, 768), (7, 512), (1000, 1280), (33, 64), (64, 2048)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_layernorm(rows, D, mode):
    x = torch.randn(rows, D) * 2 + 0.5
    g, b, gy = torch.randn(D), torch.randn(D), torch.randn(rows, D)
    if mode == "bf16":
        x, gy = bf16_round(x), bf16_round(gy)
    xr, gr, br = x.clone().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = O.layer_norm(xr, gr, br)
    yr.backward(gy)
    xg, gg, bg = x.to(DEV).requires_grad_(True), g.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    with mar.precision(mode):
        y = ops.layer_norm(xg, gg, bg)
        y.backward(gy.to(DEV).to(y.dtype))
    tol = FP32_TOL if mode == "fp32" else BF16_TOL
    assert_close(y.float().cpu(), yr, tol, "ln y")
    assert_close(xg.grad.float().cpu(), xr.grad, tol, "ln dx")
    assert_close(gg.grad.cpu(), gr.grad, tol, "ln dgamma")
    assert_close(bg.grad.cpu(), br.grad, tol, "ln dbeta")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_layernorm_row_maps_place_and_split(mode):
    """The final norms write torch.cat / the per-modality slices themselves (mar_layernorm_*_mapped): two encoders'
    norms placed into one fused (B, Ta+Tv, D) buffer == cat of the plain norms, with gradients arriving through strided
    views; a norm split into per-slice tensors == slices of the plain norm, with one incoming gradient per slice (one of
    them None)."""
    B, Ta, Tv, D = 3, 21, 9, 768
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    xa, xv = torch.randn(B, Ta, D, device=DEV), torch.randn(B, Tv, D, device=DEV)
    ga, ba, gv, bv = [torch.randn(D, device=DEV) for _ in range(4)]
    go = torch.randn(B, Ta + Tv, D, device=DEV)
    with mar.precision(mode):
        leaves = [t.clone().requires_grad_(True) for t in (xa, ga, ba, xv, gv, bv)]
        ref = torch.cat([ops.layer_norm(leaves[0], leaves[1], leaves[2]), ops.layer_norm(leaves[3], leaves[4], leaves[5])], dim=1)
        ref.backward(go.to(ref.dtype))
        ref_grads = [t.grad.clone() for t in leaves]
        leaves2 = [t.clone().requires_grad_(True) for t in (xa, ga, ba, xv, gv, bv)]
        buf = torch.full((B, Ta + Tv, D), float("nan"), device=DEV, dtype=dt)
        pa = ops.layer_norm(leaves2[0], leaves2[1], leaves2[2], place=(buf, 0))
        pv = ops.layer_norm(leaves2[3], leaves2[4], leaves2[5], place=(buf, Ta))
        pa._mar_fused, pv._mar_fused = (buf, 0), (buf, Ta)
        assert pa.data_ptr() == buf.data_ptr() and pv.shape == (B, Tv, D)
        cat = ops.concat_time([pa, pv])
        assert cat.data_ptr() == buf.data_ptr()                       # no copy
        assert torch.equal(cat, ref)
        cat.backward(go.to(cat.dtype))
        for a, b_, name in zip([t.grad for t in leaves2], ref_grads, ("dxa", "dga", "dba", "dxv", "dgv", "dbv")):
            assert_close(a.float(), b_.float(), 1e-6, f"placed layer norm {name}")
        # blocks that are NOT slices of one buffer still go through the copying concat
        assert ops.fused_slices_of([pa, ops.layer_norm(xv, gv, bv)]) is None
        # split
        x = torch.randn(B, Ta + Tv, D, device=DEV)
        g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
        l1 = [t.clone().requires_grad_(True) for t in (x, g, b)]
        full = ops.layer_norm(l1[0], l1[1], l1[2])
        full[:, :Ta].float().pow(2).sum().backward()
        l2 = [t.clone().requires_grad_(True) for t in (x, g, b)]
        sa, sv = ops.layer_norm(l2[0], l2[1], l2[2], split=[(0, Ta), (Ta, Ta + Tv)])
        assert sa.is_contiguous() and sv.is_contiguous() and torch.equal(sa, full[:, :Ta]) and torch.equal(sv, full[:, Ta:])
        sa.float().pow(2).sum().backward()                          # the second slice gets no gradient at all
        for a, b_, name in zip([t.grad for t in l2], [t.grad for t in l1], ("dx", "dgamma", "dbeta")):
            assert_close(a.float(), b_.float(), 1e-6, f"split layer norm {name}")
        assert float(l2[0].grad[:, Ta:].abs().max()) == 0.0


def test_layernorm_zero_rows_is_beta():
    x = torch.randn(10, 768, device=DEV)
    g, b = torch.randn(768, device=DEV), torch.randn(768, device=DEV)
    z = torch.zeros(10, dtype=torch.uint8, device=DEV)
    z[3] = 1
    with mar.precision("fp32"), torch.no_grad():
        y = ops.layer_norm(x, g, b, zero_rows=z)
    assert_close(y[3], b, 1e-6, "LN(0) = beta")
    assert_close(y[4].cpu(), O.layer_norm(x[4].cpu(), g.cpu(), b.cpu()), 1e-5, "other rows untouched")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_meanpool_rowzero_concat(mode):
    B, T, D = 5, 37, 768
    x = torch.randn(B, T, D)
    if mode == "bf16":
        x = bf16_round(x)
    x[1, 5] = 0.0
    x[4, 36] = 0.0
    xg = x.to(DEV).requires_grad_(True)
    with mar.precision(mode):
        m = ops.mean_pool(xg)
        m.float().sum().backward()
        mask = ops.rowzero_mask(xg.detach())
    assert_close(m.float().cpu(), x.mean(dim=1), FP32_TOL if mode == "fp32" else 4e-3, "mean_pool")
    assert_close(xg.grad.float().cpu(), torch.full_like(x, 1.0 / T), 4e-3, "mean_pool bwd")
    assert torch.equal(mask.cpu().bool(), x.sum(dim=2) == 0)
    y = torch.randn(B, 11, D)
    if mode == "bf16":
        y = bf16_round(y)
    yg = y.to(DEV).requires_grad_(True)
    with mar.precision(mode):
        cat = ops.concat_time([xg, yg])
        sl = ops.slice_time(cat, T, T + 11)
        (sl.float() * 2).sum().backward()
    assert torch.equal(cat.float().cpu(), torch.cat([x, y], dim=1))
    assert torch.equal(sl.float().cpu(), y)
    assert float((yg.grad.float() - 2).abs().max()) == 0.0


def test_cross_entropy_weighted_and_ignored():
    B, C = 37, 2
    logits = torch.randn(B, C)
    labels = torch.randint(0, C, (B,))
    w = torch.tensor([0.3, 1.7])
    lr = logits.clone().requires_grad_(True)
    ref = O.cross_entropy(lr, labels, w)
    ref.backward()
    lg = logits.to(DEV).requires_grad_(True)
    loss = ops.cross_entropy(lg, labels.to(DEV), w.to(DEV))
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-6
    assert_close(lg.grad.cpu(), lr.grad, 1e-5, "CE grad")
    # ignored rows (label -1 = EMPTY)
    keep = torch.rand(B) < 0.6
    lab2 = labels.masked_fill(~keep, -1)
    ref2 = O.cross_entropy(logits[keep], labels[keep])
    got2 = ops.cross_entropy(logits.to(DEV), lab2.to(DEV))
    assert abs(float(got2) - float(ref2)) < 1e-6
    assert torch.equal(ops.argmax_rows(logits.to(DEV)).cpu(), logits.argmax(dim=1))


@pytest.mark.parametrize("gamma", [0.0, 1.5, 2.0])
@pytest.mark.parametrize("weighted", [False, True])
def test_focal_loss_kernel_and_module(gamma, weighted):
    """mar_focal_loss_fwd + the FocalLoss drop-in (train_multimodal.py:494-510) against the oracle restatement of the
    hub module: value and gradient, ignored rows, the all-ignored batch, and gamma = 0 == alpha-weighted NLL mean."""
    from multimodalaggressionrecognition_b200 import models as M
    B, C = 37, 3
    logits = torch.randn(B, C) * 2
    labels = torch.randint(0, C, (B,))
    labels[::5] = -100
    alpha = torch.tensor([0.2, 1.0, 2.5]) if weighted else None
    lr = logits.clone().requires_grad_(True)
    ref = O.focal_loss(lr, labels, alpha, gamma)
    ref.backward()
    crit = M.FocalLoss(alpha=alpha, gamma=gamma, reduction="mean")
    lg = logits.to(DEV).requires_grad_(True)
    got = crit(lg, labels.to(DEV))
    got.backward()
    assert abs(float(got) - float(ref)) < 2e-6 * max(1.0, abs(float(ref)))
    assert_close(lg.grad.cpu(), lr.grad, 2e-5, "focal grad")
    assert float(lg.grad[::5].abs().max()) == 0.0
    # through the multimodal criterion wrapper (EMPTY rows are masked the same way)
    mm = M.MultiModalCrossEntropyLoss({"phys": crit})
    names = tuple("phys" if i % 7 else "phys_EMPTY" for i in range(B))
    lab2 = labels.clone(); lab2[labels == -100] = 0
    out = mm({"phys": logits.to(DEV)}, [[names, lab2.to(DEV)]])
    keep = torch.tensor([i % 7 != 0 for i in range(B)])
    assert abs(float(out["phys"]) - float(O.focal_loss(logits[keep], lab2[keep], alpha, gamma))) < 2e-6 * max(1.0, abs(float(ref)))
    # all rows ignored -> 0, finite gradient
    z = M.FocalLoss(gamma=gamma)(lg, torch.full((B,), -100, device=DEV))
    assert float(z) == 0.0


# ------------------------------------------------------------------------------------------
# recurrences
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("B,T,I,H", [(8, 16, 512, 512), (3, 7, 64, 96)])
def test_rnn_fwd_bwd(kind, mode, B, T, I, H):
    G = 3 if kind == "gru" else 4
    x = torch.randn(B, T, I)
    k = 1 / math.sqrt(H)
    w_ih, w_hh = (torch.rand(G * H, I) * 2 - 1) * k, (torch.rand(G * H, H) * 2 - 1) * k
    b_ih, b_hh = (torch.rand(G * H) * 2 - 1) * k, (torch.rand(G * H) * 2 - 1) * k
    gy = torch.randn(B, T, H)
    if mode == "bf16":
        x, gy = bf16_round(x), bf16_round(gy)
    ps = [t.clone().requires_grad_(True) for t in (x, w_ih, w_hh, b_ih, b_hh)]
    yr = (O.gru if kind == "gru" else O.lstm)(*ps)
    yr.backward(gy)
    pg = [t.to(DEV).requires_grad_(True) for t in (x, w_ih, w_hh, b_ih, b_hh)]
    with mar.precision(mode):
        y = (ops.gru if kind == "gru" else ops.lstm)(*pg)
        y.backward(gy.to(DEV).to(y.dtype))
    tol = FP32_TOL if mode == "fp32" else 3e-2
    assert_close(y.float().cpu(), yr, tol, f"{kind} hseq")
    for name, a, b in zip(("dx", "dw_ih", "dw_hh", "db_ih", "db_hh"), pg, ps):
        assert_close(a.grad.float().cpu(), b.grad, tol, f"{kind} {name}")


@pytest.mark.parametrize("B,T,H", [(64, 64, 512), (8, 16, 512), (13, 9, 512), (5, 33, 256), (1, 1, 512)])
@pytest.mark.parametrize("train", [False, True])
def test_gru_persistent_cluster_engine_matches_step_engine(B, T, H, train):
    """The one-launch 16-CTA-cluster GRU (W_hh resident in shared memory, DSMEM exchange of h_t) against the
    step-per-launch engine on the same bf16 inputs: hseq, the saved gates and h_{t-1} that backward consumes."""
    from multimodalaggressionrecognition_b200 import _lib
    k = 1 / math.sqrt(H)
    gi = (torch.randn(B, T, 3 * H, device=DEV)).to(torch.bfloat16)
    w = ((torch.rand(3 * H, H, device=DEV) * 2 - 1) * k).to(torch.bfloat16)
    bh = ((torch.rand(3 * H, device=DEV) * 2 - 1) * k).float()
    outs = {}
    for eng in (_lib.ENGINE_TCGEN05, _lib.ENGINE_SIMT):
        hseq = torch.full((B, T, H), float("nan"), device=DEV, dtype=torch.bfloat16)
        hprev = torch.full((B, T, H), float("nan"), device=DEV, dtype=torch.bfloat16) if train else None
        saved = torch.full((B, T, 5 * H), float("nan"), device=DEV) if train else None
        work = torch.empty(int(_lib.load().mar_gru_work_floats(B, T, H)), device=DEV)
        _lib.call("mar_gru_fwd", gi.data_ptr(), w.data_ptr(), bh.data_ptr(), hseq.data_ptr(),
                  None if hprev is None else hprev.data_ptr(), None if saved is None else saved.data_ptr(),
                  work.data_ptr(), B, T, H, _lib.MAR_BF16, eng, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs[eng] = (hseq, hprev, saved)
    a, b = outs[_lib.ENGINE_TCGEN05], outs[_lib.ENGINE_SIMT]
    assert torch.isfinite(a[0].float()).all()
    assert_close(a[0].float(), b[0].float(), 4e-3, "persistent GRU hseq")
    if train:
        assert_close(a[1].float(), b[1].float(), 4e-3, "persistent GRU hprev")
        assert_close(a[2], b[2], 4e-3, "persistent GRU saved gates")


@pytest.mark.parametrize("B,T,H", [(64, 64, 512), (8, 16, 512), (13, 9, 512), (5, 33, 256), (1, 1, 512), (40, 3, 256)])
@pytest.mark.parametrize("last_only", [False, True])
def test_gru_persistent_backward_matches_step_engine(B, T, H, last_only):
    """BPTT in one launch (fp32 carry reduce-scattered over DSMEM) against the step-per-launch engine, fed with the
    gates a real forward saved; `last_only` is the reference heads' pattern (only seq[:, -1] has a gradient)."""
    from multimodalaggressionrecognition_b200 import _lib
    k = 1 / math.sqrt(H)
    gi = torch.randn(B, T, 3 * H, device=DEV).to(torch.bfloat16)
    w = ((torch.rand(3 * H, H, device=DEV) * 2 - 1) * k).to(torch.bfloat16)
    bh = ((torch.rand(3 * H, device=DEV) * 2 - 1) * k).float()
    st = torch.cuda.current_stream().cuda_stream
    hseq = torch.empty(B, T, H, device=DEV, dtype=torch.bfloat16)
    hprev = torch.empty_like(hseq)
    saved = torch.empty(B, T, 5 * H, device=DEV)
    work = torch.empty(int(_lib.load().mar_gru_work_floats(B, T, H)), device=DEV)
    _lib.call("mar_gru_fwd", gi.data_ptr(), w.data_ptr(), bh.data_ptr(), hseq.data_ptr(), hprev.data_ptr(), saved.data_ptr(),
              work.data_ptr(), B, T, H, _lib.MAR_BF16, _lib.ENGINE_SIMT, st)
    dh = torch.randn(B, T, H, device=DEV).to(torch.bfloat16)
    if last_only:
        dh[:, :-1] = 0
    outs = {}
    for eng in (_lib.ENGINE_TCGEN05, _lib.ENGINE_SIMT):
        dgi = torch.full((B, T, 3 * H), float("nan"), device=DEV, dtype=torch.bfloat16)
        dgh = torch.full_like(dgi, float("nan"))
        _lib.call("mar_gru_bwd", dh.data_ptr(), saved.data_ptr(), w.data_ptr(), dgi.data_ptr(), dgh.data_ptr(),
                  work.data_ptr(), B, T, H, _lib.MAR_BF16, eng, st)
        torch.cuda.synchronize()
        outs[eng] = (dgi.float(), dgh.float())
    a, b = outs[_lib.ENGINE_TCGEN05], outs[_lib.ENGINE_SIMT]
    assert torch.isfinite(a[0]).all() and torch.isfinite(a[1]).all()
    assert_close(a[0], b[0], 6e-3, "persistent GRU dgi")
    assert_close(a[1], b[1], 6e-3, "persistent GRU dgh")
    # per-step check as well: the carry must not drift along the sequence
    for t in (0, T // 2, T - 1):
        assert_close(a[0][:, t], b[0][:, t], 1.5e-2, f"persistent GRU dgi at t={t}")


@pytest.mark.parametrize("B,T,H", [(64, 64, 512), (8, 16, 512), (13, 9, 512), (5, 33, 256), (1, 1, 512), (40, 3, 256)])
def test_lstm_persistent_cluster_engine_matches_step_engine(B, T, H):
    """nn.LSTM recurrence (gate order i,f,g,o; train_video_rnn.py:94-106) on the persistent cluster engine: forward
    (hseq, saved i,f,g,o,c, bf16 h_{t-1}) and BPTT (d gates) against the step-per-launch engine."""
    from multimodalaggressionrecognition_b200 import _lib
    k = 1 / math.sqrt(H)
    gi = torch.randn(B, T, 4 * H, device=DEV).to(torch.bfloat16)
    w = ((torch.rand(4 * H, H, device=DEV) * 2 - 1) * k).to(torch.bfloat16)
    bh = ((torch.rand(4 * H, device=DEV) * 2 - 1) * k).float()
    st = torch.cuda.current_stream().cuda_stream
    dh = torch.randn(B, T, H, device=DEV).to(torch.bfloat16)
    outs = {}
    for eng in (_lib.ENGINE_TCGEN05, _lib.ENGINE_SIMT):
        hseq = torch.full((B, T, H), float("nan"), device=DEV, dtype=torch.bfloat16)
        hprev = torch.full_like(hseq, float("nan"))
        saved = torch.full((B, T, 5 * H), float("nan"), device=DEV)
        work = torch.empty(B * 5 * H, device=DEV)
        _lib.call("mar_lstm_fwd", gi.data_ptr(), w.data_ptr(), bh.data_ptr(), hseq.data_ptr(), hprev.data_ptr(),
                  saved.data_ptr(), work.data_ptr(), B, T, H, _lib.MAR_BF16, eng, st)
        dg = torch.full((B, T, 4 * H), float("nan"), device=DEV, dtype=torch.bfloat16)
        _lib.call("mar_lstm_bwd", dh.data_ptr(), saved.data_ptr(), w.data_ptr(), dg.data_ptr(), work.data_ptr(),
                  B, T, H, _lib.MAR_BF16, eng, st)
        torch.cuda.synchronize()
        outs[eng] = (hseq.float(), hprev.float(), saved, dg.float())
    a, b = outs[_lib.ENGINE_TCGEN05], outs[_lib.ENGINE_SIMT]
    for x in a:
        assert torch.isfinite(x).all()
    assert_close(a[0], b[0], 4e-3, "persistent LSTM hseq")
    assert_close(a[1], b[1], 4e-3, "persistent LSTM hprev")
    assert_close(a[2], b[2], 4e-3, "persistent LSTM saved gates / cell state")
    assert_close(a[3], b[3], 1e-2, "persistent LSTM d(gates)")     # the two engines back-propagate their own saved gates


def test_gru_auto_engine_is_persistent_in_bf16():
    x = torch.randn(4, 6, 512, device=DEV)
    ps = [torch.randn(3 * 512, 512, device=DEV) * 0.04, torch.randn(3 * 512, 512, device=DEV) * 0.04,
          torch.zeros(3 * 512, device=DEV), torch.zeros(3 * 512, device=DEV)]
    with mar.precision("bf16"):
        ops.reset_launch_count()
        ops.gru(x, *ps)
        n_bf16 = ops.launch_count()
    with mar.precision("fp32"):
        ops.reset_launch_count()
        ops.gru(x, *ps)
        n_fp32 = ops.launch_count()
    assert n_bf16 < n_fp32 and n_bf16 <= 6, (n_bf16, n_fp32)      # casts + input GEMM + ONE recurrence launch (fp32: one GEMM + one cell kernel per step)


def test_adam_kernel_matches_oracle():
    n = 10007
    p0, g = torch.randn(n), torch.randn(n)
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    pg, mg, vg = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    step = torch.zeros(1, device=DEV)
    from multimodalaggressionrecognition_b200._lib import call
    st = torch.cuda.current_stream().cuda_stream
    for t in range(1, 4):
        O.adam_step([p], [g], [m], [v], t)
        call("mar_adam_tick", step.data_ptr(), st)
        call("mar_adam_step", pg.data_ptr(), g.to(DEV).data_ptr(), mg.data_ptr(), vg.data_ptr(), step.data_ptr(), n,
             1e-3, 0.9, 0.999, 1e-8, st)
    assert_close(pg.cpu(), p, 1e-6, "adam")


def test_segmented_adam_kernel_skips_inactive_parameters_and_counts_steps_per_parameter():
    """mar_adam_step_segments against the oracle's per-parameter Adam (oracle.adam_step with a step LIST and grad None
    for the skipped parameters = torch.optim.Adam, torch/optim/adam.py `_init_group`): 5 steps over 4 parameters whose
    active sets change from step to step."""
    from multimodalaggressionrecognition_b200 import training
    from multimodalaggressionrecognition_b200._lib import call
    torch.manual_seed(3)
    shapes = [(7, 33), (130,), (64, 64), (5,)]
    host = [torch.randn(*s) for s in shapes]
    params = [torch.nn.Parameter(h.clone().to(DEV)) for h in host]
    flat = training.FlatParams(params)
    opt = training.FlatAdam(flat, lr=1e-2)
    m, v = [torch.zeros_like(h) for h in host], [torch.zeros_like(h) for h in host]
    counts = [0] * 4
    pattern = [(1, 1, 1, 1), (1, 0, 1, 0), (0, 0, 1, 1), (1, 0, 0, 0), (1, 1, 1, 1)]
    for active in pattern:
        grads = [torch.randn(*s) for s in shapes]
        flat.zero_grad()
        for i, a in enumerate(active):
            if a:
                params[i].grad.copy_(grads[i].to(DEV))
                counts[i] += 1
            else:
                # garbage where torch would have .grad None: an inactive parameter must not even READ its gradient
                params[i].grad.fill_(float("nan"))
        flat.flags.copy_(torch.tensor(active, dtype=torch.float32) * 2.0)        # > 0 = active (a SUM over ranks may exceed 1)
        opt.step()
        O.adam_step(host, [g if a else None for g, a in zip(grads, active)], m, v, list(counts), lr=1e-2)
    torch.cuda.synchronize()
    for i, (p_, h) in enumerate(zip(params, host)):
        assert_close(p_.detach().cpu(), h, 1e-6, f"param {i}")
    assert opt.seg_steps.tolist() == [float(c) for c in counts]
    assert not bool(torch.isnan(flat.flat).any())
    with pytest.raises(RuntimeError, match="chunk"):
        call("mar_adam_step_segments", flat.flat.data_ptr(), flat.grad.data_ptr(), opt.exp_avg.data_ptr(),
             opt.exp_avg_sq.data_ptr(), flat.chunk_seg.data_ptr(), opt.seg_steps.data_ptr(), flat.flags.data_ptr(),
             opt.seg_coef.data_ptr(), flat.numel, 48, flat.nseg, 1e-3, 0.9, 0.999, 1e-8, None, 0)
    # the bf16 mirror the kernel maintains equals a fresh cast of the updated parameters, and ops.compute_weight serves it
    assert torch.equal(flat.mirror.float(), flat.flat.to(torch.bfloat16).float())
    wc, _ = ops.compute_weight(params[0], torch.bfloat16, False)
    assert wc.data_ptr() == flat.mirror.data_ptr() + 2 * flat.offsets[0]
    with torch.no_grad():
        params[0].add_(1.0)                       # a foreign in-place update: the mirror is no longer trusted
    wc2, _ = ops.compute_weight(params[0], torch.bfloat16, False)
    assert wc2.data_ptr() != wc.data_ptr() and torch.equal(wc2.float(), params[0].detach().to(torch.bfloat16).float())


def test_label_weight_sum_kernel():
    from multimodalaggressionrecognition_b200 import models as M
    y = torch.tensor([0, 1, -1, 1, 1, -100, 0], device=DEV)
    out = torch.zeros(2, device=DEV)
    ops.label_weight_sum(y, None, out[0:])
    ops.label_weight_sum(y, torch.tensor([0.25, 2.0]), out[1:])
    assert out.tolist() == [5.0, 2 * 0.25 + 3 * 2.0]
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(weight=torch.tensor([1.0, 3.0])),
                                         "verb": torch.nn.CrossEntropyLoss()})
    labels = [[("verb", "verb_EMPTY", "verb"), torch.tensor([1, 0, 0])], [("phys", "phys", "phys"), torch.tensor([1, 1, 0])]]
    assert crit.label_weight_sums(labels, DEV, None) == ["phys", "verb"]
    crit.label_weight_sums(labels, DEV, out)
    assert out.tolist() == [7.0, 2.0]
    crit.criterion_dict["verb"] = torch.nn.CrossEntropyLoss(reduction="sum")
    assert crit.label_weight_sums(labels, DEV, None) is None


def test_error_reporting_through_abi():
    from multimodalaggressionrecognition_b200._lib import call
    with pytest.raises(RuntimeError, match="mar_layernorm_fwd"):
        call("mar_layernorm_fwd", None, None, None, None, None, None, None, 4, 768, 1e-5, 0, 0)
    x = torch.randn(4, 100, device=DEV)     # D not a multiple of 8
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.layer_norm(x, torch.ones(100, device=DEV), torch.zeros(100, device=DEV))
