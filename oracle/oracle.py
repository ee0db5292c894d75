"""CPU oracle for the sequence-classifier hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, as explicit fp32 tensor arithmetic on the CPU, the algorithm that the
reference (cafe1930/MultimodalAggressionRecognition) runs on its hot path.  The reference holds
no arithmetic of its own: every number is produced by torch.nn modules it instantiates
(`nn.TransformerEncoder`, `nn.GRU`, `nn.LSTM`, `nn.Linear`, `nn.LayerNorm`, `nn.CrossEntropyLoss`,
`optim.Adam`; pinned `pytorch=2.1.0` in env_config.yml:115, installed here: torch 2.11).  So each
function below cites (a) the reference call site in /root/reference and (b) the published
algorithm of the torch module it restates.  Nothing here calls nn.TransformerEncoder, nn.GRU,
F.scaled_dot_product_attention, F.layer_norm or F.cross_entropy: only matmul / exp / tanh / sum,
so that it is an independent statement of the math the CUDA kernels must reproduce.

Parity pinning: the reference ships no golden vectors.  The oracle is pinned
against the LIVE reference classes imported from /root/reference in the build container by
`oracle/make_golden.py`, which also writes the fixtures under `tests/golden/`; `tests/
test_oracle_golden.py` re-checks the oracle against those fixtures on any machine.
Exception — PARITY UNPINNED: `focal_loss` restates the torch.hub criterion of train_multimodal.py:494-510
(adeelh/pytorch-multi-class-focal-loss, default branch, not vendored, no network here) from its published
algorithm; it is checked against the same formula written with torch's log_softmax / nll_loss only.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` (its `cpu_baseline` leg and
`--impl reference`) may import this module.  The product package never does.

State-dict key names follow the reference modules, so the same dict of tensors
drives the reference classes, this oracle and the CUDA drop-in modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """y = x·Wᵀ + b with W stored (out, in) — nn.Linear (torch/nn/modules/linear.py:117-134)."""
    y = torch.matmul(x, w.t())
    return y if b is None else y + b


def relu(x: Tensor) -> Tensor:
    return torch.clamp_min(x, 0.0)


# Exact parity is defined with dropout off on both sides (`nn.Dropout.p = 0`,
# `self_attn.dropout = 0.0` on the reference).  Parity checks set this to False; the timed CPU
# baseline leaves it True so that the mask generation cost (37 % of the reference's CPU step) is paid.
DROPOUT_ENABLED = True


def dropout(x: Tensor, p: float, training: bool) -> Tensor:
    """Inverted dropout: keep with prob 1-p, scale by 1/(1-p) (nn.Dropout).  The mask stream is
    torch's CPU generator; a CUDA kernel cannot share it, so exact parity is defined at p=0."""
    if not training or p == 0.0 or not DROPOUT_ENABLED:
        return x
    keep = torch.bernoulli(torch.full_like(x, 1.0 - p))
    return x * keep / (1.0 - p)


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
    """Biased-variance LayerNorm over the last dim (nn.LayerNorm; models.py:352, :403 and
    TransformerEncoderLayer.norm1/norm2)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * gamma + beta


def softmax_lastdim_safe(s: Tensor) -> Tensor:
    """softmax over the last dim; a row that is entirely -inf gives all zeros (torch 2.11's
    safe softmax inside SDPA — SURVEY.md §7 'all-masked rows')."""
    m = s.max(dim=-1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m)
    den = e.sum(dim=-1, keepdim=True)
    return torch.where(den > 0, e / den.clamp_min(1e-38), torch.zeros_like(e))


def multi_head_self_attention(x: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor,
                              num_heads: int, key_padding_mask: Optional[Tensor],
                              attn_dropout_p: float, training: bool) -> Tensor:
    """nn.MultiheadAttention self-attention, batch_first, packed in-projection
    (F.multi_head_attention_forward, torch/nn/functional.py:6244-6695; rows [0:d]=Q, [d:2d]=K,
    [2d:3d]=V of in_proj_weight, :5798).  key_padding_mask: (B,T) bool, True = key ignored
    (canonicalised to additive -inf, transformer.py:431-437)."""
    B, T, d = x.shape
    dh = d // num_heads
    qkv = linear(x, in_w, in_b)                          # (B,T,3d)
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    q = q.reshape(B, T, num_heads, dh).permute(0, 2, 1, 3)   # (B,H,T,dh)
    k = k.reshape(B, T, num_heads, dh).permute(0, 2, 1, 3)
    v = v.reshape(B, T, num_heads, dh).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)  # (B,H,T,T)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    p = softmax_lastdim_safe(s)
    p = dropout(p, attn_dropout_p, training)
    o = torch.matmul(p, v)                                # (B,H,T,dh)
    o = o.permute(0, 2, 1, 3).reshape(B, T, d)
    return linear(o, out_w, out_b)


def encoder_layer(x: Tensor, sd: SD, prefix: str, num_heads: int, key_padding_mask: Optional[Tensor],
                  p: float, training: bool) -> Tensor:
    """Post-norm TransformerEncoderLayer with ReLU (torch/nn/modules/transformer.py:950-982),
    as constructed at models.py:348 and :398 with torch defaults (d_ff=2048, dropout=0.1)."""
    a = multi_head_self_attention(
        x, sd[prefix + "self_attn.in_proj_weight"], sd[prefix + "self_attn.in_proj_bias"],
        sd[prefix + "self_attn.out_proj.weight"], sd[prefix + "self_attn.out_proj.bias"],
        num_heads, key_padding_mask, p, training)
    x1 = layer_norm(x + dropout(a, p, training), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    h = dropout(relu(linear(x1, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])), p, training)
    f = linear(h, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])
    return layer_norm(x1 + dropout(f, p, training), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])


def mask_is_left_aligned(key_padding_mask: Tensor) -> bool:
    """torch._nested_tensor_from_mask_left_aligned on ~mask: every row is valid…valid pad…pad."""
    valid = (~key_padding_mask).to(torch.int64)
    # once a pad is seen no valid may follow  <=>  valid is non-increasing along T
    return bool((valid[:, 1:] <= valid[:, :-1]).all())


def transformer_encoder(x: Tensor, sd: SD, prefix: str, num_layers: int, num_heads: int,
                        key_padding_mask: Optional[Tensor] = None, p: float = 0.1,
                        training: bool = False, grad_enabled: bool = False) -> Tensor:
    """nn.TransformerEncoder(layer, N, norm=LayerNorm) (transformer.py:407-553; models.py:349-352,
    :400-403, call sites :365 and :425).

    Eval asymmetry: when not training, no grad is being recorded, a padding mask
    is given and it is left-aligned, torch converts to a nested tensor: padded tokens are dropped
    and re-inserted as ZEROS before the final LayerNorm (transformer.py:455-548).  Valid tokens
    see exactly the same arithmetic as below with the key mask applied."""
    nested = (not training) and (not grad_enabled) and key_padding_mask is not None \
        and mask_is_left_aligned(key_padding_mask)
    if nested and bool(key_padding_mask.all(dim=1).all()):
        raise RuntimeError("to_padded_tensor: at least one constituent tensor should have non-zero numel")
    out = x
    for i in range(num_layers):
        out = encoder_layer(out, sd, f"{prefix}layers.{i}.", num_heads, key_padding_mask, p, training)
    if nested:
        out = out.masked_fill(key_padding_mask[:, :, None], 0.0)
    return layer_norm(out, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])


def gru(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """1-layer batch_first nn.GRU, h0 = 0, gate row blocks (r,z,n) (torch/nn/modules/rnn.py GRU
    docstring; reference call site models.py:110,122).  Returns the full (B,T,H) sequence."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    gi_all = linear(x, w_ih, b_ih)                        # (B,T,3H)
    outs = []
    for t in range(T):
        gi = gi_all[:, t]
        gh = linear(h, w_hh, b_hh)
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, dim=1)


def lstm(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """1-layer batch_first nn.LSTM, (h0,c0)=0, gate row blocks (i,f,g,o) (rnn.py LSTM docstring;
    train_video_rnn.py:94-106)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    gi_all = linear(x, w_ih, b_ih)
    outs = []
    for t in range(T):
        g = gi_all[:, t] + linear(h, w_hh, b_hh)
        i = torch.sigmoid(g[:, :H])
        f = torch.sigmoid(g[:, H:2 * H])
        gg = torch.tanh(g[:, 2 * H:3 * H])
        o = torch.sigmoid(g[:, 3 * H:])
        c = f * c + i * gg
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def cross_entropy(logits: Tensor, labels: Tensor, weight: Optional[Tensor] = None) -> Tensor:
    """nn.CrossEntropyLoss(), mean reduction; weighted form divides by Σ w[y_i]
    (models.py:258, :293; train_multimodal.py:512-513)."""
    m = logits.max(dim=1, keepdim=True).values
    lse = m.squeeze(1) + torch.log(torch.exp(logits - m).sum(dim=1))
    nll = lse - logits.gather(1, labels[:, None]).squeeze(1)
    if weight is None:
        return nll.mean()
    w = weight[labels]
    return (nll * w).sum() / w.sum()


def focal_loss(logits: Tensor, labels: Tensor, alpha: Optional[Tensor] = None, gamma: float = 0.0,
               ignore_index: int = -100) -> Tensor:
    """The criterion train_multimodal.py:494-510 loads with torch.hub.load('adeelh/pytorch-multi-class-focal-loss',
    'FocalLoss', alpha=class_weights, gamma=..., reduction='mean').  The hub repo is a third-party dependency that is
    unpinned (default branch) and absent offline, so this restates its published algorithm — PARITY UNPINNED:
        log_p = log_softmax(x);  ce = NLLLoss(weight=alpha, reduction='none')(log_p, y) = -alpha[y]·log_p[y]
        pt = exp(log_p[y]);  loss = (1 - pt)^gamma · ce;  'mean' → loss.mean() over the rows with y != ignore_index
    (a plain mean, unlike the alpha-weighted mean of nn.CrossEntropyLoss); no such rows → 0."""
    keep = labels != ignore_index
    labels, logits = labels[keep], logits[keep]
    if labels.numel() == 0:
        return logits.sum() * 0.0
    m = logits.max(dim=1, keepdim=True).values
    log_p = logits - m - torch.log(torch.exp(logits - m).sum(dim=1, keepdim=True))
    log_pt = log_p.gather(1, labels[:, None]).squeeze(1)
    a = torch.ones_like(log_pt) if alpha is None else alpha[labels]
    return ((1.0 - torch.exp(log_pt)) ** gamma * (-a * log_pt)).mean()


def adam_step(params: List[Tensor], grads: List[Optional[Tensor]], exp_avg: List[Tensor],
              exp_avg_sq: List[Tensor], step, lr: float = 1e-3, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam defaults, no weight decay (train_multimodal.py:444):
    m=β1 m+(1-β1)g ; v=β2 v+(1-β2)g² ; p -= lr/(1-β1^t) · m / (sqrt(v)/sqrt(1-β2^t) + eps).
    Parameters whose grad is None are skipped entirely, as torch does (torch/optim/adam.py `_init_group`: only
    parameters with a gradient enter the update) — and t is PER PARAMETER (torch keeps `state['step']` per
    parameter and increments it only when the parameter is updated): the reference's batches are homogeneous in
    aggression type (datasets.py:630-645), so a head, or a whole modality branch, sits out every other step.
    `step`: one int for all parameters, or a list with each parameter's own update count (this update included)."""
    steps = step if isinstance(step, (list, tuple)) else [step] * len(params)
    with torch.no_grad():
        for p, g, m, v, t in zip(params, grads, exp_avg, exp_avg_sq, steps):
            if g is None:
                continue
            bc1 = 1.0 - beta1 ** t
            bc2 = 1.0 - beta2 ** t
            m.mul_(beta1).add_(g, alpha=1.0 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
            p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# modules of models.py, as functions of a state dict
# --------------------------------------------------------------------------------------


def embedding_layer(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """EmbeddingLayer: per-token ReLU(Linear) (models.py:139-150)."""
    return relu(linear(x, sd[prefix + "embedding.0.weight"], sd[prefix + "embedding.0.bias"]))


def transformer_sequence_processor(x: Tensor, sd: SD, prefix: str, num_layers: int, num_heads: int,
                                   extractor: str = "identity", training: bool = False,
                                   p: float = 0.1) -> Tensor:
    """TransformerSequenceProcessor.forward (models.py:362-365): encoder(feature_extractor(x)),
    no mask.  `extractor` ∈ {'identity' (nn.Sequential()), 'embedding' (EmbeddingLayer)}."""
    if extractor == "embedding":
        x = embedding_layer(x, sd, prefix + "feature_extractor.")
    return transformer_encoder(x, sd, prefix + "transformer_squence_processing.", num_layers, num_heads,
                               None, p, training)


def mlp_head(x: Tensor, sd: SD, prefix: str, i0: int, i1: int, p: float, training: bool) -> Tensor:
    """Linear → ReLU → Dropout(p) → Linear, Sequential indices i0 / i1."""
    h = dropout(relu(linear(x, sd[f"{prefix}{i0}.weight"], sd[f"{prefix}{i0}.bias"])), p, training)
    return linear(h, sd[f"{prefix}{i1}.weight"], sd[f"{prefix}{i1}.bias"])


def output_classifier(x: Tensor, sd: SD, prefix: str, training: bool = False) -> Tensor:
    """OutputClassifier (models.py:378-389): mean_T → Linear d→256 → ReLU → Dropout .3 → Linear."""
    return mlp_head(x.mean(dim=1), sd, prefix + "classifier.", 1, 4, 0.3, training)


def feature_sequence_processing(x: Tensor, sd: SD, prefix: str, kind: str, training: bool = False) -> Tensor:
    """FeatureSequenceProcessing (models.py:107-124): sequence_nn → LAST time step → MLP with
    Dropout(0.5).  kind ∈ {'gru','lstm','avg'} ('avg' = AverageFeatureSequence, :91-97)."""
    if kind == "gru":
        seq = gru(x, sd[prefix + "sequence_nn.weight_ih_l0"], sd[prefix + "sequence_nn.weight_hh_l0"],
                  sd[prefix + "sequence_nn.bias_ih_l0"], sd[prefix + "sequence_nn.bias_hh_l0"])
    elif kind == "lstm":
        seq = lstm(x, sd[prefix + "sequence_nn.weight_ih_l0"], sd[prefix + "sequence_nn.weight_hh_l0"],
                   sd[prefix + "sequence_nn.bias_ih_l0"], sd[prefix + "sequence_nn.bias_hh_l0"])
    elif kind == "avg":
        seq = x.mean(dim=1).unsqueeze(1)
    else:
        raise ValueError(kind)
    return mlp_head(seq[:, -1, :], sd, prefix + "output_classifier.", 0, 3, 0.5, training)


def video_multi_nn(x: Tensor, sd: SD, heads: Dict[str, str], training: bool = False) -> Dict[str, Tensor]:
    """VideoMultiNN.forward (models.py:169-175): every head on the same input."""
    return {name: feature_sequence_processing(x, sd, f"models_dict.{name}.", kind, training)
            for name, kind in heads.items()}


def zero_row_key_mask(concat: Tensor) -> Tensor:
    """key_padding_mask = (concat.sum(dim=2) == 0) (models.py:421-422)."""
    return concat.sum(dim=2) == 0


def equal_sized_fusion(feats: Dict[str, Tensor], sd: SD, prefix: str, num_layers: int, num_heads: int,
                       training: bool = False, grad_enabled: bool = False, p: float = 0.1) -> Dict[str, Tensor]:
    """EqualSizedTransformerModalitiesFusion.forward (models.py:405-430)."""
    bounds, start = {}, 0
    for name, t in feats.items():
        bounds[name] = (start, start + t.shape[1])
        start += t.shape[1]
    concat = torch.cat(list(feats.values()), dim=1)
    mask = zero_row_key_mask(concat)
    fused = transformer_encoder(concat, sd, prefix + "modality_fusion_transformer.", num_layers, num_heads,
                                mask, p, training, grad_enabled)
    return {name: fused[:, b0:b1] for name, (b0, b1) in bounds.items()}


def averaged_features_fusion(feats: Dict[str, Tensor], sd: SD, prefix: str, num_layers: int, num_heads: int,
                             training: bool = False, grad_enabled: bool = False, p: float = 0.1) -> Dict[str, Tensor]:
    """AveragedFeaturesTransformerFusion.forward (models.py:482-503): mean-pool each modality to one
    token, then the same fusion."""
    pooled = {k: v.mean(dim=1).unsqueeze(1) for k, v in feats.items()}
    return equal_sized_fusion(pooled, sd, prefix, num_layers, num_heads, training, grad_enabled, p)


def physverb_classifier_concat(feats: Dict[str, Tensor], sd: SD, prefix: str, aggr_types: Sequence[str],
                               training: bool = False, p: float = 0.3) -> Dict[str, Tensor]:
    """PhysVerbClassifierConcatFeatures.forward (models.py:756-770): per modality
    Linear → Dropout → ReLU → mean_T (:743-748), concat over sorted modalities (:757,:764), every
    aggression-type head Linear → ReLU → Dropout → Linear on the same concat (:767-768)."""
    pooled = []
    for name in sorted(feats):
        a = linear(feats[name], sd[f"{prefix}adaptors_dict.{name}.0.weight"], sd[f"{prefix}adaptors_dict.{name}.0.bias"])
        pooled.append(relu(dropout(a, p, training)).mean(dim=1))
    cat = torch.cat(pooled, dim=1)
    return {t: mlp_head(cat, sd, f"{prefix}classifiers_dict.{t}.", 0, 3, p, training) for t in aggr_types}


def physverb_classifier(feats: Dict[str, Tensor], sd: SD, prefix: str, modality2aggr: Dict[str, str],
                        training: bool = False, p: float = 0.3) -> Dict[str, Tensor]:
    """Base PhysVerbClassifier.forward (models.py:718-735): heads see only their own modalities."""
    per_type: Dict[str, Tensor] = {}
    for name in sorted(feats):
        a = linear(feats[name], sd[f"{prefix}adaptors_dict.{name}.0.weight"], sd[f"{prefix}adaptors_dict.{name}.0.bias"])
        a = relu(dropout(a, p, training)).mean(dim=1)
        t = modality2aggr[name]
        per_type[t] = torch.cat([per_type[t], a], dim=1) if t in per_type else a
    return {t: mlp_head(f, sd, f"{prefix}classifiers_dict.{t}.", 0, 3, p, training) for t, f in per_type.items()}


def split_names(names: Sequence[str]) -> Tuple[str, np.ndarray]:
    """'<modality>[_EMPTY]' parsing (models.py:840-846, :244-247)."""
    base = names[0].split("_")[0]
    not_empty = np.array([n.split("_")[-1] for n in names]) != "EMPTY"
    return base, not_empty


def physverb_extract_features(data, sd: SD, cfg: dict, training: bool = False) -> Dict[str, Tensor]:
    """PhysVerbModel.extract_features (models.py:835-863): zeros stub (B,*shape), extractor only on
    the non-EMPTY rows, scatter, dict sorted by modality name."""
    out = {}
    for names, batch in data:
        name, not_empty = split_names(names)
        feat = torch.zeros([batch.shape[0]] + list(cfg["feature_shapes"][name]), dtype=batch.dtype)
        if not_empty.any() and name in cfg["extractors"]:
            e = cfg["extractors"][name]
            idx = torch.from_numpy(not_empty)
            if e["layers"] == 0:      # nn.Sequential() as the extractor (text, train_multimodal.py:365): pass-through
                got = batch[idx]
            else:
                got = transformer_sequence_processor(batch[idx], sd, f"modality_extractors_dict.{name}.",
                                                     e["layers"], e["heads"], e["extractor"], training)
            feat = feat.index_put((idx,), got)
        out[name] = feat
    return dict(sorted(out.items()))


def physverb_model(data, sd: SD, cfg: dict, training: bool = False, grad_enabled: bool = False) -> Dict[str, Tensor]:
    """PhysVerbModel.forward (models.py:865-879) with EqualSizedTransformerModalitiesFusion and
    PhysVerbClassifierConcatFeatures, the C3 assembly of train_multimodal.py:298-420; `cfg` also selects the
    script's alternatives: "fusion": "avg" (AveragedFeaturesTransformerFusion, :375), "classifier": "base"
    (PhysVerbClassifier, models.py:667-735), "top": "old" (MultimodalModel, models.py:505-558)."""
    feats = physverb_extract_features(data, sd, cfg, training)
    fuse = averaged_features_fusion if cfg.get("fusion", "equal") == "avg" else equal_sized_fusion
    fused = fuse(feats, sd, "modality_fusion_module.", cfg["fusion_layers"], cfg["fusion_heads"], training, grad_enabled)
    if cfg.get("top", "physverb") == "old":
        # MultimodalModel.forward (models.py:543-555): one OutputClassifier per modality on its own fused slice
        return {m: output_classifier(fused[m], sd, f"classifiers.{m}.", training) for m in fused}
    if cfg.get("classifier", "concat") == "base":
        return physverb_classifier(fused, sd, "classifiers.", cfg["modality2aggr"], training)
    return physverb_classifier_concat(fused, sd, "classifiers.", cfg["aggr_types"], training)


def multimodal_ce(pred: Dict[str, Tensor], target, weights: Optional[Dict[str, Tensor]] = None,
                  heads: Optional[Sequence[str]] = None) -> Dict[str, Tensor]:
    """MultiModalCrossEntropyLoss.forward (models.py:238-263): one CE per label group that has at
    least one non-EMPTY sample and a configured criterion; EMPTY rows are filtered out."""
    losses = {}
    for names, labels in target:
        name, not_empty = split_names(names)
        if not_empty.any() and (heads is None or name in heads):
            idx = torch.from_numpy(not_empty)
            w = None if weights is None else weights.get(name)
            losses[name] = cross_entropy(pred[name][idx], labels[idx], w)
    return losses


def multi_ce(pred: Dict[str, Tensor], labels: Tensor) -> Dict[str, Tensor]:
    """MultiCrossEntropyLoss.forward (models.py:290-295)."""
    return {k: cross_entropy(v, labels) for k, v in pred.items()}


def audio_text_model(data, sd: SD, cfg: dict, training: bool = False) -> Tensor:
    """AudioTextualModel.forward (models.py:907-926): mean_T(audio) ‖ mean_T(text) → Linear 2d→d,
    ReLU, Dropout .3 → Linear d→256, ReLU, Dropout .3, Linear 256→C."""
    d = {names[0]: t for names, t in data}
    a = transformer_sequence_processor(d["audio"], sd, "audio_extractor.", cfg["audio"]["layers"],
                                       cfg["audio"]["heads"], cfg["audio"]["extractor"], training)
    t = transformer_sequence_processor(d["text"], sd, "text_extractor.", cfg["text"]["layers"],
                                       cfg["text"]["heads"], cfg["text"]["extractor"], training)
    cat = torch.cat([a.mean(dim=1), t.mean(dim=1)], dim=-1)
    f = dropout(relu(linear(cat, sd["modality_fusion_module.0.weight"], sd["modality_fusion_module.0.bias"])), 0.3, training)
    return mlp_head(f, sd, "output_classifier.", 0, 3, 0.3, training)


# --------------------------------------------------------------------------------------
# whole training steps (used by the tests as the checker and by bench.py as the CPU baseline)
# --------------------------------------------------------------------------------------


class OracleTrainer:
    """fwd → dict of losses → one backward per head (LossesDict.backward, models.py:226-230; the sum
    of per-head gradients) → Adam.  `forward_fn(sd, batch, training)` returns {head: logits};
    `loss_fn(pred, target)` returns {head: scalar}."""

    def __init__(self, sd: SD, forward_fn, loss_fn, lr: float = 1e-3):
        self.sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        self.forward_fn, self.loss_fn, self.lr = forward_fn, loss_fn, lr
        self.m = {k: torch.zeros_like(v) for k, v in self.sd.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.sd.items()}
        self.t = 0
        self.steps = {k: 0 for k in self.sd}          # torch.optim.Adam's per-parameter state['step']

    def step(self, data, target, training: bool = True) -> Dict[str, float]:
        for p in self.sd.values():
            p.grad = None
        pred = self.forward_fn(self.sd, data, training)
        losses = self.loss_fn(pred, target)
        items = list(losses.items())
        for i, (_, l) in enumerate(items):
            l.backward(retain_graph=i != len(items) - 1)
        self.t += 1
        keys = list(self.sd)
        for k in keys:
            if self.sd[k].grad is not None:
                self.steps[k] += 1
        adam_step([self.sd[k] for k in keys], [self.sd[k].grad for k in keys],
                  [self.m[k] for k in keys], [self.v[k] for k in keys], [self.steps[k] for k in keys], self.lr)
        return {k: float(v.detach()) for k, v in items}
