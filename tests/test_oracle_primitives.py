"""CPU: every primitive of the oracle against the torch module whose published algorithm it restates — the
third-party dependency the reference's arithmetic actually lives in (torch.nn) — on random shapes,
forward AND gradients.  Complements tests/test_oracle_golden.py (whole models against the live reference's outputs):
the golden vectors pin the oracle where the reference was run, these pin each building block everywhere else
(odd lengths, masks of every shape, weighted / ignored labels)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import oracle as O


@pytest.fixture(autouse=True)
def _no_dropout():
    old = O.DROPOUT_ENABLED
    O.DROPOUT_ENABLED = False
    yield
    O.DROPOUT_ENABLED = old


def _close(a, b, tol=2e-5, what=""):
    err = float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    assert err <= tol, f"{what}: {err:.2e}"


@pytest.mark.parametrize("shape,D", [((7, 5), 64), ((3,), 768), ((2, 9), 96)])
def test_layer_norm(shape, D):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, D, generator=g, requires_grad=True)
    gamma, beta = torch.randn(D, generator=g, requires_grad=True), torch.randn(D, generator=g, requires_grad=True)
    up = torch.randn(*shape, D, generator=g)
    ref = F.layer_norm(x, (D,), gamma, beta, 1e-5)
    gr = torch.autograd.grad((ref * up).sum(), [x, gamma, beta])
    got = O.layer_norm(x, gamma, beta)
    gg = torch.autograd.grad((got * up).sum(), [x, gamma, beta])
    _close(got, ref, what="LN")
    for a, b, n in zip(gg, gr, ("dx", "dgamma", "dbeta")):
        _close(a, b, 5e-5, n)


@pytest.mark.parametrize("B,T,d,H,mask", [(2, 11, 64, 4, None), (3, 17, 96, 8, "random"), (2, 8, 64, 2, "tail"), (2, 5, 32, 1, "row")])
def test_multi_head_self_attention(B, T, d, H, mask):
    """F.multi_head_attention_forward through nn.MultiheadAttention (train mode, need_weights=False), incl. a fully
    masked row: torch >= 2.5's safe softmax gives zeros there, as the oracle does."""
    g = torch.Generator().manual_seed(2)
    mha = nn.MultiheadAttention(d, H, batch_first=True).train()
    with torch.no_grad():
        for p in mha.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.2)
    x = torch.randn(B, T, d, generator=g, requires_grad=True)
    kpm = None
    if mask == "random":
        kpm = torch.rand(B, T, generator=g) < 0.3
        kpm[:, 0] = False
    elif mask == "tail":
        kpm = torch.zeros(B, T, dtype=torch.bool)
        kpm[0, 5:] = True
    elif mask == "row":
        kpm = torch.zeros(B, T, dtype=torch.bool)
        kpm[1] = True                      # every key of sample 1 masked
    ref, _ = mha(x, x, x, key_padding_mask=kpm, need_weights=False)
    got = O.multi_head_self_attention(x, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias,
                                      H, kpm, 0.0, True)
    _close(got, ref, what="MHA out")
    up = torch.randn(B, T, d, generator=g)
    params = [x, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias]
    gr = torch.autograd.grad((ref * up).sum(), params)
    gg = torch.autograd.grad((got * up).sum(), params)
    for a, b, n in zip(gg, gr, ("dx", "din_w", "din_b", "dout_w", "dout_b")):
        _close(a, b, 1e-4, f"MHA {n}")


@pytest.mark.parametrize("layers,mask,training", [(2, None, True), (1, "tail", True), (1, "tail", False), (2, "middle", False),
                                                  (1, None, False)])
def test_transformer_encoder(layers, mask, training):
    """nn.TransformerEncoder(post-norm ReLU layers, norm=LayerNorm) as models.py:348-352 builds it, dropout off:
    train mode, and eval mode under no_grad where torch takes the nested-tensor path for a left-aligned mask
    (padded tokens re-inserted as zeros before the final norm) but not for a mask in the middle."""
    g = torch.Generator().manual_seed(3)
    d, H, B, T = 64, 4, 3, 10
    enc = nn.TransformerEncoder(nn.TransformerEncoderLayer(d, H, batch_first=True), layers, norm=nn.LayerNorm(d))
    with torch.no_grad():
        for p in enc.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.2 if p.dim() > 1 else 0.5))
    for m in enc.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    x = torch.randn(B, T, d, generator=g)
    kpm = None
    if mask == "tail":
        kpm = torch.zeros(B, T, dtype=torch.bool)
        kpm[0, 6:] = True
        kpm[2, 3:] = True
    elif mask == "middle":
        kpm = torch.zeros(B, T, dtype=torch.bool)
        kpm[1, 2:5] = True
    sd = {k: v.detach() for k, v in enc.state_dict().items()}
    enc.train(training)
    if training:
        ref = enc(x, src_key_padding_mask=kpm)
        got = O.transformer_encoder(x, sd, "", layers, H, kpm, 0.0, True, True)
    else:
        with torch.no_grad():
            ref = enc(x, src_key_padding_mask=kpm)
            got = O.transformer_encoder(x, sd, "", layers, H, kpm, 0.0, False, False)
    _close(got, ref.detach(), 5e-5, f"encoder layers={layers} mask={mask} training={training}")
    if not training and mask == "tail":
        beta = sd["norm.bias"]
        _close(got[0, 7], beta, 1e-6, "padded token = LayerNorm(0) = beta on the nested path")


@pytest.mark.parametrize("kind", ["gru", "lstm"])
@pytest.mark.parametrize("B,T,I,H", [(3, 7, 16, 24), (1, 1, 8, 8), (5, 12, 32, 16)])
def test_recurrences(kind, B, T, I, H):
    g = torch.Generator().manual_seed(4)
    rnn = (nn.GRU if kind == "gru" else nn.LSTM)(I, H, num_layers=1, batch_first=True)
    x = torch.randn(B, T, I, generator=g, requires_grad=True)
    ref, _ = rnn(x)
    params = [rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0]
    got = (O.gru if kind == "gru" else O.lstm)(x, *params)
    _close(got, ref, what=kind)
    up = torch.randn(B, T, H, generator=g)
    gr = torch.autograd.grad((ref * up).sum(), [x] + params)
    gg = torch.autograd.grad((got * up).sum(), [x] + params)
    for a, b, n in zip(gg, gr, ("dx", "dw_ih", "dw_hh", "db_ih", "db_hh")):
        _close(a, b, 1e-4, f"{kind} {n}")


@pytest.mark.parametrize("C,weighted", [(2, False), (2, True), (5, True)])
def test_cross_entropy(C, weighted):
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(40, C, generator=g) * 3).requires_grad_(True)
    y = torch.randint(0, C, (40,), generator=g)
    w = torch.rand(C, generator=g) + 0.2 if weighted else None
    ref = F.cross_entropy(x, y, weight=w)
    got = O.cross_entropy(x, y, w)
    _close(got, ref, 1e-6, "CE")
    _close(torch.autograd.grad(got, x)[0], torch.autograd.grad(ref, x)[0], 1e-5, "dCE")


def test_embedding_and_heads_against_torch_modules():
    """EmbeddingLayer / OutputClassifier / FeatureSequenceProcessing-style MLP heads are Linear-ReLU(-Dropout)-Linear
    stacks over a mean: the oracle's helpers against the same stack written with torch.nn."""
    g = torch.Generator().manual_seed(6)
    seq = nn.Sequential(nn.Linear(24, 16), nn.ReLU(), nn.Dropout(0.0), nn.Linear(16, 2))
    x = torch.randn(4, 9, 24, generator=g)
    sd = {f"h.{k}": v.detach() for k, v in seq.state_dict().items()}
    _close(O.mlp_head(x.mean(dim=1), sd, "h.", 0, 3, 0.3, False), seq(x.mean(dim=1)).detach(), what="mlp head")
    emb = nn.Sequential(nn.Linear(24, 16), nn.ReLU())
    sd = {f"e.embedding.{k}": v.detach() for k, v in emb.state_dict().items()}
    _close(O.embedding_layer(x, sd, "e."), emb(x).detach(), what="embedding layer")
