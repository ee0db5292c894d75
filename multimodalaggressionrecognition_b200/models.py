"""Drop-in replacements for the hot-path classes of the reference's models.py.

Same class names, constructor signatures, forward signatures, sub-module names and state_dict keys
as /root/reference/models.py (cited per class), so `trainer.py`, `datasets.py` and the train_*.py
scripts work unchanged — but every forward runs hand-written sm_100a kernels through libmar.so
(`ops.py`).  The constructors build the very same torch.nn parameter containers in the same order
as the reference, so "same seed ⇒ same initial weights" and checkpoints interchange; the torch
modules are used as PARAMETER HOLDERS only, their forward() is never called.

Precision: `ops.set_precision('bf16'|'fp32')`.  Activations flow between these modules in the
compute dtype; classifier logits are always fp32.

Out of scope here: frozen video/audio backbones, 1-D CNN, 3-D CNN
classifiers and the reference's dead classes.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops

__all__ = [
    "AverageFeatureSequence", "SequenceAverageFeatures", "FeatureSequenceProcessing", "VideoAverageFeatures",
    "EmbeddingLayer", "VideoMultiNN", "AudioMultiNN", "LossesDict", "MultiModalCrossEntropyLoss",
    "MultiCrossEntropyLoss", "TransformerSequenceProcessor", "OutputClassifier",
    "EqualSizedTransformerModalitiesFusion", "AveragedFeaturesTransformerFusion", "MultimodalModel",
    "PhysVerbClassifier", "PhysVerbClassifierConcatFeatures", "PhysVerbModel", "AudioTextualModel",
]


# --------------------------------------------------------------------------------------
# helpers: run torch parameter containers through the kernels
# --------------------------------------------------------------------------------------
def _split_names(names):
    """'<modality>' / '<modality>_EMPTY' → (modality, bool mask of present samples)  (models.py:840-846)."""
    base = names[0].split("_")[0]
    present = np.array([n.split("_")[-1] for n in names]) != "EMPTY"
    return base, present


_row_cache: Dict = {}


def _present_rows(present: np.ndarray, device: torch.device, kind: str) -> torch.Tensor:
    """Device copy of a per-sample presence pattern: 'idx' → int64 indices of the present rows, 'keep' → bool mask.
    The pattern comes from the batch's NAMES (host strings), so it is uploaded once per distinct pattern and reused:
    no boolean-mask indexing (nonzero() = a device→host sync) and no pageable upload on the step's path — both are
    illegal while a CUDA graph is being captured (an eager step with the same names always precedes a capture)."""
    key = (present.tobytes(), str(device), kind)
    t = _row_cache.get(key)
    if t is None:
        if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("a presence pattern met for the first time during CUDA-graph capture")
        host = torch.from_numpy(np.nonzero(present)[0].astype(np.int64)) if kind == "idx" else torch.from_numpy(present.copy())
        t = host.to(device)
        if len(_row_cache) > 4096:
            _row_cache.clear()
        _row_cache[key] = t
    return t


def _mlp_head(seq: nn.Sequential, i0: int, i1: int, x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    """Linear(i0) → ReLU → Dropout(p) → Linear(i1); logits in fp32."""
    l0, l1 = seq[i0], seq[i1]
    h = ops.linear(x, l0.weight, l0.bias, relu_pre=True, dropout_p=p if training else 0.0)
    return ops.linear(h, l1.weight, l1.bias, out_dtype=torch.float32)


def _dropout_p(seq: nn.Sequential) -> float:
    for m in seq:
        if isinstance(m, nn.Dropout):
            return float(m.p)
    return 0.0


def _check_layer(layer: nn.TransformerEncoderLayer) -> None:
    if layer.norm_first:
        raise NotImplementedError("pre-norm TransformerEncoderLayer is not on the reference's path")
    act = getattr(layer, "activation_relu_or_gelu", 0)
    if act != 1:
        raise NotImplementedError("only the ReLU TransformerEncoderLayer of the reference is implemented")
    if not layer.self_attn.batch_first or not layer.self_attn._qkv_same_embed_dim:
        raise NotImplementedError("only batch_first self-attention with packed in-projection is implemented")


def encoder_layer_forward(layer: nn.TransformerEncoderLayer, h: torch.Tensor, B: int, T: int,
                          key_mask: Optional[torch.Tensor]) -> torch.Tensor:
    """Post-norm encoder layer on (B*T, d) rows (torch/nn/modules/transformer.py:950-982):
    x1 = LN1(x + Drop(OutProj(Attn(QKV(x))))) ; x2 = LN2(x1 + Drop(W2 Drop(ReLU(W1 x1))))."""
    sa = layer.self_attn
    train = layer.training
    d = h.shape[-1]
    # fork=True: the residual branch takes h / x1 from the projection op, so in backward the residual gradient is
    # added inside that projection's dgrad GEMM epilogue instead of by a separate elementwise pass
    # hand-offs (ops.new_handoff): each linear's dropout backward / bias-gradient sum is done by the neighbour that
    # touches the same tensor in backward anyway (norm1 / norm2, the attention kernel, linear2's dgrad GEMM), so no
    # separate pass over a (B*T, ·) gradient is launched for them
    h_in, h_out, h_l1, h_l2 = (ops.new_handoff() for _ in range(4))
    qkv, h_res = ops.linear(h, sa.in_proj_weight, sa.in_proj_bias, fork=True, handoff=h_in)
    o = ops.attention(qkv.view(B, T, 3 * d), key_mask, sa.num_heads, sa.dropout if train else 0.0, qkv_handoff=h_in)
    pre1 = ops.linear(o.view(B * T, d), sa.out_proj.weight, sa.out_proj.bias, residual=h_res,
                      dropout_p=layer.dropout1.p if train else 0.0, handoff=h_out)
    x1 = ops.layer_norm(pre1, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, x_handoff=h_out)
    # linear1's ReLU+dropout backward is applied inside linear2's dgrad GEMM (hid is zero exactly where that derivative
    # is zero): the (B*T, d_ff) gradient is written once, already masked, and linear1 only sums its bias gradient
    p_ff = float(layer.dropout.p) if train else 0.0
    hid, x1_res = ops.linear(x1, layer.linear1.weight, layer.linear1.bias, relu_pre=True, dropout_p=p_ff, fork=True,
                             defer_act=True, handoff=h_l1)
    pre2 = ops.linear(hid, layer.linear2.weight, layer.linear2.bias, residual=x1_res,
                      dropout_p=layer.dropout2.p if train else 0.0, x_act_scale=1.0 / (1.0 - p_ff), handoff=h_l2,
                      x_handoff=h_l1)
    return ops.layer_norm(pre2, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, x_handoff=h_l2)


def _left_aligned(key_mask: torch.Tensor) -> bool:
    valid = key_mask == 0
    return bool((valid[:, 1:] <= valid[:, :-1]).all())


def encoder_forward(enc: nn.TransformerEncoder, x: torch.Tensor, key_mask: Optional[torch.Tensor] = None, place=None,
                    split=None):
    """nn.TransformerEncoder(layer, N, norm=LayerNorm) on (B,T,d) (transformer.py:407-553).

    Reproduces torch's eval-mode asymmetry: when the encoder is in eval mode, no
    grad is recorded, a key padding mask is given and it is left-aligned, torch's nested-tensor
    fast path re-inserts padded tokens as zeros before the final LayerNorm (transformer.py:455-548).

    place = (buffer, t_off): the final norm writes into buffer[:, t_off:t_off+T] (the fused sequence) and that view is
    returned; split = [(t0, t1), ...]: the final norm returns one contiguous tensor per time slice (a tuple).
    """
    B, T, d = x.shape
    for layer in enc.layers:
        _check_layer(layer)
    training = enc.layers[0].training
    grad_on = torch.is_grad_enabled() and any(p.requires_grad for p in enc.layers[0].parameters())
    zero_rows = None
    if key_mask is not None and not training and not grad_on and getattr(enc, "use_nested_tensor", True):
        if _left_aligned(key_mask):      # one host sync, eval only (torch does the same check)
            if bool(key_mask.all()):
                raise RuntimeError("to_padded_tensor: at least one constituent tensor should have non-zero numel")
            zero_rows = key_mask.reshape(-1)
    h = ops.to_compute(x).reshape(B * T, d)
    for layer in enc.layers:
        h = encoder_layer_forward(layer, h, B, T, key_mask)
    if enc.norm is not None:
        if place is not None or split is not None:
            out = ops.layer_norm(h.view(B, T, d), enc.norm.weight, enc.norm.bias, enc.norm.eps, zero_rows=zero_rows,
                                 place=place, split=split)
            if place is not None:
                out._mar_fused = (place[0], place[1])        # lets the fusion module recognise its own slices
            return out
        h = ops.layer_norm(h, enc.norm.weight, enc.norm.bias, enc.norm.eps, zero_rows=zero_rows)
    elif zero_rows is not None:
        h = h.masked_fill(zero_rows.bool()[:, None], 0.0)
    h = h.view(B, T, d)
    if split is not None:
        return ops.split_time(h, split)
    return h


# --------------------------------------------------------------------------------------
# sequence encoders
# --------------------------------------------------------------------------------------
class AverageFeatureSequence(nn.Module):
    """models.py:91-97 (and the duplicate at :298-304): mean over T shaped like an RNN output, `(B,1,d), None`."""

    def __init__(self, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size

    def forward(self, x):
        if ops.probing():
            return x.new_zeros(x.shape[0], 1, x.shape[2]), None
        return ops.mean_pool(x).unsqueeze(1), None


class SequenceAverageFeatures(nn.Module):
    """models.py:99-105: x.mean(dim=1)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size

    def forward(self, x):
        if ops.probing():
            return x.new_zeros(x.shape[0], x.shape[2])
        return ops.mean_pool(x)


class FeatureSequenceProcessing(nn.Module):
    """models.py:107-124: sequence_nn (nn.GRU / nn.LSTM / AverageFeatureSequence built from a
    {'model': cls, 'kwargs': {...}} dict) → LAST time step → Linear H→256, ReLU, Dropout(0.5), Linear 256→C.
    The torch RNN module only holds weight_ih_l0 / weight_hh_l0 / bias_*; the recurrence runs in mar_gru_* /
    mar_lstm_*."""

    def __init__(self, sequence_nn_dict, class_num):
        super().__init__()
        self.sequence_nn = sequence_nn_dict['model'](**sequence_nn_dict['kwargs'])
        self.hidden_size = sequence_nn_dict['kwargs']['hidden_size']
        self.output_classifier = nn.Sequential(
            nn.Linear(self.hidden_size, 256),
            nn.ReLU(),
            nn.Dropout(),
            nn.Linear(256, class_num),
        )

    def _run_sequence(self, x):
        nn_ = self.sequence_nn
        if isinstance(nn_, (nn.GRU, nn.LSTM)):
            if nn_.num_layers != 1 or nn_.bidirectional or not nn_.batch_first or not nn_.bias or getattr(nn_, "proj_size", 0):
                raise NotImplementedError("only the reference's 1-layer unidirectional batch_first GRU/LSTM is implemented")
            fn = ops.gru if isinstance(nn_, nn.GRU) else ops.lstm
            return fn(x, nn_.weight_ih_l0, nn_.weight_hh_l0, nn_.bias_ih_l0, nn_.bias_hh_l0)
        out = nn_(x)
        return out[0] if isinstance(out, tuple) else out

    def forward(self, sequence):
        if ops.probing():
            return sequence.new_zeros(sequence.shape[0], self.output_classifier[3].out_features)
        seq = self._run_sequence(sequence)
        last = seq[:, -1, :]
        return _mlp_head(self.output_classifier, 0, 3, last, self.output_classifier[2].p, self.training)


class VideoAverageFeatures(nn.Module):
    """models.py:126-137: mean over T → Linear, ReLU, Dropout(0.5), Linear."""

    def __init__(self, input_dim, class_num):
        super().__init__()
        self.output_classifier = nn.Sequential(
            nn.Linear(input_dim, 256),
            nn.ReLU(),
            nn.Dropout(),
            nn.Linear(256, class_num),
        )

    def forward(self, x):
        if ops.probing():
            return x.new_zeros(x.shape[0], self.output_classifier[3].out_features)
        return _mlp_head(self.output_classifier, 0, 3, ops.mean_pool(x), self.output_classifier[2].p, self.training)


class EmbeddingLayer(nn.Module):
    """models.py:139-150: per-token ReLU(Linear(in→out)); one GEMM with a fused bias+ReLU epilogue."""

    def __init__(self, input_size, output_size):
        super().__init__()
        self.embedding = nn.Sequential(
            nn.Linear(input_size, output_size),
            nn.ReLU(),
        )

    def forward(self, x):
        lin = self.embedding[0]
        if ops.probing():
            return x.new_zeros(x.shape[0], x.shape[1], lin.out_features)
        return ops.linear(x, lin.weight, lin.bias, relu_pre=True)


class VideoMultiNN(nn.Module):
    """models.py:152-175: run every head on the same features → {name: logits}."""

    def __init__(self, models_dict):
        super().__init__()
        self.models_dict = nn.ModuleDict(models_dict)

    def get_models_names(self):
        return [name for name in self.models_dict.keys()]

    def forward(self, x):
        if not ops.probing():
            x = ops.to_compute(x)      # cast once, shared by all heads
        return {name: model(x) for name, model in self.models_dict.items()}


class AudioMultiNN(nn.Module):
    """models.py:198-223.  The frozen extractor(s) are out-of-scope torch modules and run as given under
    no_grad; the heads run on the kernels."""

    def __init__(self, models_dict, extractor_dict):
        super().__init__()
        self.extractor_dict = nn.ModuleDict(extractor_dict)
        self.extractor_dict.eval()
        self.models_dict = nn.ModuleDict(models_dict)

    def get_models_names(self):
        return [n for n in self.extractor_dict.keys()], [n for n in self.models_dict.keys()]

    def forward(self, x):
        with torch.no_grad():
            for _, extractor in self.extractor_dict.items():
                features = extractor(x)
        if not ops.probing():
            features = ops.to_compute(features)
        return {name: model(features) for name, model in self.models_dict.items()}


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
class LossesDict(dict):
    """models.py:225-230.  The reference back-propagates each head separately with retain_graph (the shared
    trunk is walked once per head and the gradients add up).  One backward over all heads at once gives
    the same sums while walking the trunk once."""

    def backward(self):
        losses = [l for l in self.values() if isinstance(l, torch.Tensor) and l.requires_grad]
        if losses:
            torch.autograd.backward(losses)

    def total(self):
        vals = list(self.values())
        return sum(vals[1:], vals[0]) if vals else None


class FocalLoss(nn.Module):
    """Drop-in for `torch.hub.load('adeelh/pytorch-multi-class-focal-loss', 'FocalLoss', alpha=..., gamma=...,
    reduction='mean')` (train_multimodal.py:494-510): same constructor arguments, fused forward + gradient
    kernel (mar_focal_loss_fwd).  loss_i = -alpha[y_i]·(1 - p_{y_i})^gamma·log p_{y_i}; 'mean' / 'sum' / over the
    rows whose label is not `ignore_index`; an all-ignored batch gives 0."""

    def __init__(self, alpha: Optional[torch.Tensor] = None, gamma: float = 0.0, reduction: str = "mean",
                 ignore_index: int = -100):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise ValueError('Reduction must be one of: "mean", "sum" (per-row "none" is not on the hot path)')
        self.alpha = None if alpha is None else torch.as_tensor(alpha, dtype=torch.float32)
        self.gamma, self.reduction, self.ignore_index = float(gamma), reduction, int(ignore_index)

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.dim() > 2:                      # (N, C, d1, ..) -> (N·d1·.., C), as the hub module does
            c = x.shape[1]
            x = x.permute(0, *range(2, x.dim()), 1).reshape(-1, c)
            y = y.reshape(-1)
        y = y.to(x.device)
        y = y.masked_fill(y == self.ignore_index, -1)
        loss = ops.focal_loss(x, y, self.alpha, self.gamma)
        if self.reduction == "sum":
            loss = loss * (y >= 0).sum().to(loss.dtype)
        return loss


def _apply_criterion(criterion, preds, labels, keep=None):
    """Plain / class-weighted nn.CrossEntropyLoss and this module's FocalLoss run on the fused kernels (ignored
    rows carry label -1); any other criterion object is called as given."""
    if isinstance(criterion, FocalLoss) and preds.is_cuda:
        labels = _effective_labels(criterion, labels, keep)
        loss = ops.focal_loss(preds, labels, criterion.alpha, criterion.gamma)
        return loss if criterion.reduction == "mean" else loss * (labels >= 0).sum().to(loss.dtype)
    if _mean_norm_kind(criterion) == "weighted" and preds.is_cuda:
        return ops.cross_entropy(preds, _effective_labels(criterion, labels, keep), criterion.weight)
    if keep is not None:            # a foreign criterion object: called as given, on the kept rows (not capturable)
        preds, labels = preds[keep], labels[keep]
    return criterion(preds.float(), labels)


def _mean_norm_kind(criterion) -> Optional[str]:
    """'weighted' / 'count': what the criterion's mean divides by; None if it is not a fused mean-reduced criterion."""
    if isinstance(criterion, FocalLoss):
        return "count" if criterion.reduction == "mean" else None
    if isinstance(criterion, nn.CrossEntropyLoss) and criterion.reduction == "mean" and criterion.label_smoothing == 0.0:
        return "weighted"
    return None


def _effective_labels(criterion, labels, keep):
    """Labels as the fused loss kernels read them: -1 on every row the criterion ignores."""
    if keep is not None:
        labels = labels.masked_fill(~keep, -1)
    ii = criterion.ignore_index
    if isinstance(criterion, FocalLoss) or ii >= 0:
        labels = labels.masked_fill(labels == ii, -1)
    return labels


def _label_weight_sum(criterion, labels, keep, out) -> None:
    kind = _mean_norm_kind(criterion)
    labels = _effective_labels(criterion, labels, keep).to(torch.int64).contiguous()
    w = criterion.weight if kind == "weighted" else None
    ops.label_weight_sum(labels, w, out)


class MultiModalCrossEntropyLoss(nn.Module):
    """models.py:232-263: one loss per label group that has a non-EMPTY sample and a configured criterion."""

    def __init__(self, modalities_losses_dict):
        super().__init__()
        self.criterion_dict = modalities_losses_dict
        self.modalities_list = list(modalities_losses_dict.keys())

    def forward(self, output_dict, target):
        losses_dict = LossesDict()
        for names, labels in target:
            name, present = _split_names(names)
            if not present.any() or name not in self.modalities_list:
                continue
            preds = output_dict[name]
            labels = labels.to(preds.device)
            keep = None if present.all() else _present_rows(present, preds.device, "keep")
            losses_dict[name] = _apply_criterion(self.criterion_dict[name], preds, labels, keep)
        return losses_dict

    def label_weight_sums(self, target, device, out: Optional[torch.Tensor]):
        """The denominators of the configured 'mean' criteria on THIS batch, per head in `modalities_list` order:
        Σ class_weight[label] over the rows the loss keeps (their count without class weights).  out=None → the list
        of heads (None when a criterion is not one of the fused mean-reduced ones); otherwise out[i] is written on
        the device (no sync).  A data-parallel TrainStep exchanges these to weigh every rank's gradient so that the
        reduced gradient is the one of the reference loss on the global batch."""
        for c in self.criterion_dict.values():
            if _mean_norm_kind(c) is None:
                return None
        if out is None:
            return list(self.modalities_list)
        out.zero_()
        for names, labels in target:
            name, present = _split_names(names)
            if not present.any() or name not in self.modalities_list:
                continue
            keep = None if present.all() else _present_rows(present, torch.device(device), "keep")
            _label_weight_sum(self.criterion_dict[name], labels.to(device), keep, out[self.modalities_list.index(name):])
        return list(self.modalities_list)


class MultiCrossEntropyLoss(nn.Module):
    """models.py:285-295: the same CE on every head's logits."""

    def __init__(self):
        super().__init__()
        self.criterion = nn.CrossEntropyLoss()

    def forward(self, output_dict, target):
        losses_dict = LossesDict()
        for name, preds in output_dict.items():
            losses_dict[name] = _apply_criterion(self.criterion, preds, target.to(preds.device))
        return losses_dict

    def label_weight_sums(self, target, device, out: Optional[torch.Tensor]):
        """One entry: every head is scored against the same labels with the same plain CE (see
        MultiModalCrossEntropyLoss.label_weight_sums)."""
        if out is None:
            return ["*"]
        _label_weight_sum(self.criterion, target.to(device), None, out)
        return ["*"]


# --------------------------------------------------------------------------------------
# transformer encoders / fusion
# --------------------------------------------------------------------------------------
class TransformerSequenceProcessor(nn.Module):
    """models.py:344-365: feature_extractor → nn.TransformerEncoder(L × post-norm layer, norm=LayerNorm), no mask.
    Attribute name `transformer_squence_processing` (sic) is the reference's and fixes the state_dict keys."""

    def __init__(self, extractor_model, hidden_size, transformer_layer_num, transformer_head_num, class_num):
        super().__init__()
        self.feature_extractor = extractor_model
        transformer_layer = nn.TransformerEncoderLayer(d_model=hidden_size, nhead=transformer_head_num, batch_first=True)
        self.transformer_squence_processing = nn.TransformerEncoder(
            transformer_layer,
            num_layers=transformer_layer_num,
            norm=nn.LayerNorm(hidden_size))

    def forward(self, x):
        features = self.feature_extractor(x)
        if ops.probing():
            return features.new_zeros(features.shape)
        place = None
        if features.dim() == 3 and features.is_cuda:     # a fusion model may have reserved this encoder's slice of its sequence
            place = ops.take_final_norm_placement(features.shape[0], features.shape[1], features.shape[2], ops.get_precision())
        return encoder_forward(self.transformer_squence_processing, features, None, place=place)


class OutputClassifier(nn.Module):
    """models.py:378-389: mean_T → Linear d→256 → ReLU → Dropout(0.3) → Linear 256→C."""

    def __init__(self, input_features, class_num):
        super().__init__()
        self.classifier = nn.Sequential(
            SequenceAverageFeatures(hidden_size=input_features),
            nn.Linear(input_features, 256),
            nn.ReLU(),
            nn.Dropout(0.3),
            nn.Linear(256, class_num),
        )

    def forward(self, x):
        if ops.probing():
            return x.new_zeros(x.shape[0], self.classifier[4].out_features)
        pooled = self.classifier[0](x)
        return _mlp_head(self.classifier, 1, 4, pooled, self.classifier[3].p, self.training)


class EqualSizedTransformerModalitiesFusion(nn.Module):
    """models.py:391-430: concat modalities along T → key mask from all-zero rows → encoder → split."""

    def __init__(self, fusion_transformer_layer_num, fusion_transformer_hidden_size, fusion_transformer_head_num):
        super().__init__()
        transformer_layer = nn.TransformerEncoderLayer(
            d_model=fusion_transformer_hidden_size, nhead=fusion_transformer_head_num, batch_first=True)
        self.modality_fusion_transformer = nn.TransformerEncoder(
            transformer_layer,
            num_layers=fusion_transformer_layer_num,
            norm=nn.LayerNorm(fusion_transformer_hidden_size))

    def _fuse(self, feats: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        bounds, start = {}, 0
        for name, t in feats.items():
            bounds[name] = (start, start + t.size(1))
            start += t.size(1)
        if ops.probing():
            return {name: t.new_zeros(t.shape) for name, t in feats.items()}
        blocks = list(feats.values())
        concat = blocks[0] if len(blocks) == 1 else ops.concat_time(blocks)
        key_mask = ops.rowzero_mask(concat)
        if len(blocks) == 1:
            return {next(iter(feats)): encoder_forward(self.modality_fusion_transformer, concat, key_mask)}
        # the encoder's final norm writes every modality's time slice as its own contiguous tensor (models.py:430)
        parts = encoder_forward(self.modality_fusion_transformer, concat, key_mask, split=list(bounds.values()))
        return dict(zip(bounds.keys(), parts))

    def forward(self, modalities_features_dict):
        return self._fuse(modalities_features_dict)


class AveragedFeaturesTransformerFusion(EqualSizedTransformerModalitiesFusion):
    """models.py:480-503: mean-pool every modality to one token first."""

    def forward(self, modalities_features_dict):
        if ops.probing():
            return {k: v.new_zeros(v.shape[0], 1, v.shape[2]) for k, v in modalities_features_dict.items()}
        pooled = {k: ops.mean_pool(v).unsqueeze(1) for k, v in modalities_features_dict.items()}
        return self._fuse(pooled)


# --------------------------------------------------------------------------------------
# classifier heads
# --------------------------------------------------------------------------------------
class PhysVerbClassifier(nn.Module):
    """models.py:667-735 (the effective definition; the one at :602 is shadowed).  Per modality
    Linear → Dropout → ReLU → mean_T; per aggression type a 2-layer MLP on the concat of ITS modalities.
    `p_droput` (sic) is the reference's keyword."""

    def __init__(self, modalities_list, class_num, modalities_adaptors_inout_sizes_dict,
                 modality2aggr={'video': 'phys', 'text': 'verb', 'audio': 'verb'}, p_droput=0.3):
        super().__init__()
        self.modalities_list = modalities_list
        self.modalities_inout_sizes_dict = modalities_adaptors_inout_sizes_dict
        self.class_num = class_num
        self.modalities_adaptors_inout_sizes_dict = modalities_adaptors_inout_sizes_dict
        self.modality2aggr = modality2aggr
        adaptors_dict, head_in_dims = self.prepare_adaptors(modalities_list, modalities_adaptors_inout_sizes_dict, p_droput)
        classifiers_dict = self.prepare_classifiers(head_in_dims, p_droput, class_num)
        self.classifiers_dict = nn.ModuleDict(classifiers_dict)
        self.adaptors_dict = nn.ModuleDict(adaptors_dict)

    @staticmethod
    def _adaptor(in_features, out_features, p):
        return nn.Sequential(
            nn.Linear(in_features, out_features),
            nn.Dropout(p),
            nn.ReLU(),
            SequenceAverageFeatures(hidden_size=out_features),
        )

    def prepare_adaptors(self, modalities_list, modalities_adaptors_inout_sizes_dict, p_droput):
        adaptors, dims = {}, {}
        for modality in modalities_list:
            fin, fout = modalities_adaptors_inout_sizes_dict[modality]
            adaptors[modality] = self._adaptor(fin, fout, p_droput)
            aggr = self.modality2aggr[modality]
            dims[aggr] = dims.get(aggr, 0) + fout
        return adaptors, dims

    def prepare_classifiers(self, aggr_types_classifier_in_dims_dict, p_droput, class_num):
        heads = {}
        for aggr_type, width in aggr_types_classifier_in_dims_dict.items():
            heads[aggr_type] = nn.Sequential(
                nn.Linear(width, width // 3),
                nn.ReLU(),
                nn.Dropout(p_droput),
                nn.Linear(width // 3, class_num),
            )
        return heads

    def _adapt(self, modality, features):
        """Linear → Dropout → ReLU (dropout BEFORE ReLU, models.py:743-748) fused in one GEMM epilogue, then mean_T."""
        seq = self.adaptors_dict[modality]
        lin = seq[0]
        if ops.probing():
            return features.new_zeros(features.shape[0], lin.out_features)
        # Linear → Dropout → ReLU → mean over T as ONE autograd node: backward broadcasts the pooled gradient inside the
        # epilogue kernel instead of materialising a (B, T, d) tensor for it
        return ops.linear(features, lin.weight, lin.bias, dropout_p=seq[1].p if self.training else 0.0, relu_post=True,
                          pool_T=features.shape[1])

    def _head(self, aggr_type, x):
        seq = self.classifiers_dict[aggr_type]
        if ops.probing():
            return x.new_zeros(x.shape[0], seq[3].out_features)
        return _mlp_head(seq, 0, 3, x, seq[2].p, self.training)

    def forward(self, modalities_features_dict):
        feats = dict(sorted(modalities_features_dict.items()))
        per_type: Dict[str, List[torch.Tensor]] = {}
        for modality, features in feats.items():
            per_type.setdefault(self.modality2aggr[modality], []).append(self._adapt(modality, features))
        return {t: self._head(t, xs[0] if len(xs) == 1 else torch.cat(xs, dim=1)) for t, xs in per_type.items()}


class PhysVerbClassifierConcatFeatures(PhysVerbClassifier):
    """models.py:737-770: every aggression-type head sees the concat of ALL modalities' pooled adaptors."""

    def prepare_adaptors(self, modalities_list, modalities_adaptors_inout_sizes_dict, p_droput):
        adaptors = {}
        for modality in modalities_list:
            fin, fout = modalities_adaptors_inout_sizes_dict[modality]
            adaptors[modality] = self._adaptor(fin, fout, p_droput)
        width = sum(v[1] for k, v in modalities_adaptors_inout_sizes_dict.items() if k in modalities_list)
        dims = {aggr: width for aggr in self.modality2aggr.values()}
        return adaptors, dims

    def forward(self, modalities_features_dict):
        feats = dict(sorted(modalities_features_dict.items()))
        pooled = [self._adapt(m, f) for m, f in feats.items()]
        cat = pooled[0] if len(pooled) == 1 else torch.cat(pooled, dim=1)
        return {t: self._head(t, cat) for t in self.classifiers_dict}


# --------------------------------------------------------------------------------------
# top-level multimodal models
# --------------------------------------------------------------------------------------
class _MultimodalBase(nn.Module):
    def _fused_layout(self, input_data):
        """When the fusion module concatenates the modalities along T (EqualSizedTransformerModalitiesFusion proper) and all
        feature widths agree, the per-modality features can be produced straight inside the fused (B, ΣT, d) sequence:
        {modality: (t_off, T)} in the sorted order the fusion sees, and the buffer's shape.  None otherwise."""
        fusion = getattr(self, "modality_fusion_module", None)
        if type(fusion) is not EqualSizedTransformerModalitiesFusion or ops.probing() or len(input_data) < 2:
            return None
        names = sorted(_split_names(n)[0] for n, _ in input_data)
        shapes = [self.modality_features_shapes_dict.get(n) for n in names]
        if any(s is None or len(s) != 2 for s in shapes) or len({s[1] for s in shapes}) != 1 or len(set(names)) != len(names):
            return None
        layout, off = {}, 0
        for n, sh in zip(names, shapes):
            layout[n] = (off, int(sh[0]))
            off += int(sh[0])
        return layout, off, int(shapes[0][1])

    def extract_features(self, input_data):
        """models.py:835-863 / :513-541: per modality a zeros stub (B,*shape); the extractor runs on the
        non-EMPTY rows only and its output is scattered back; result sorted by modality name."""
        out = {}
        fused = self._fused_layout(input_data) if (input_data and input_data[0][1].is_cuda) else None
        buffer = None
        if fused is not None:
            layout, total, width = fused
            buffer = torch.empty((input_data[0][1].size(0), total, width), device=input_data[0][1].device, dtype=ops.get_precision())
            # the zero stubs of EMPTY modalities are written first: once an extractor's output is a view of the buffer no
            # other view of it may be modified in place (autograd's version check)
            for names, batch in input_data:
                name, present = _split_names(names)
                if not (present.any() and name in self.modality_extractors_dict):
                    t_off, T = layout[name]
                    buffer[:, t_off:t_off + T].zero_()
        for names, batch in input_data:
            name, present = _split_names(names)
            shape = [batch.size(0)] + list(self.modality_features_shapes_dict[name])
            feats = None
            if present.any() and name in self.modality_extractors_dict:
                extractor = self.modality_extractors_dict[name]
                if present.all():
                    if buffer is not None:
                        with ops.place_final_norm(buffer, *layout[name]):
                            feats = extractor(batch)
                    else:
                        feats = extractor(batch)
                else:
                    idx = _present_rows(present, batch.device, "idx")
                    got = extractor(batch.index_select(0, idx))
                    feats = torch.zeros(shape, device=batch.device, dtype=got.dtype).index_copy(0, idx, got)
            if feats is None:
                dtype = batch.dtype if (ops.probing() or not batch.is_cuda) else ops.get_precision()
                if buffer is not None and dtype == buffer.dtype:
                    t_off, T = layout[name]                   # the zero stub of an EMPTY modality, in place (zeroed above)
                    feats = buffer[:, t_off:t_off + T]
                    feats._mar_fused = (buffer, t_off)
                else:
                    feats = torch.zeros(shape, device=batch.device, dtype=dtype)
            out[name] = feats
        return dict(sorted(out.items()))


class PhysVerbModel(_MultimodalBase):
    """models.py:823-886: extractors → fusion → classifiers(dict) → {'phys': logits, 'verb': logits}."""

    def __init__(self, modality_extractors_dict, modality_fusion_module, classifiers, modality_features_shapes_dict,
                 modality2aggr, hidden_size, class_num):
        super().__init__()
        self.modality_extractors_dict = modality_extractors_dict
        self.modality_features_shapes_dict = modality_features_shapes_dict
        self.modality_fusion_module = modality_fusion_module
        self.classifiers = classifiers
        self.modality2aggr = modality2aggr

    def forward(self, input_data):
        feats = self.extract_features(input_data)
        fused = self.modality_fusion_module(feats)
        return self.classifiers(fused)

    def get_output_names(self):
        return list(self.classifiers.classifiers_dict.keys())


class MultimodalModel(_MultimodalBase):
    """models.py:505-558: older variant, `classifiers` is a ModuleDict keyed by the modality it reads."""

    def __init__(self, modality_extractors_dict, modality_fusion_module, classifiers, modality_features_shapes_dict,
                 hidden_size, class_num):
        super().__init__()
        self.modality_extractors_dict = modality_extractors_dict
        self.modality_features_shapes_dict = modality_features_shapes_dict
        self.modality_fusion_module = modality_fusion_module
        self.classifiers = classifiers

    def forward(self, input_data):
        fused = self.modality_fusion_module(self.extract_features(input_data))
        return {k: self.classifiers[k](fused[k]) for k in self.classifiers}

    def get_output_names(self):
        return list(self.classifiers.keys())


class AudioTextualModel(nn.Module):
    """models.py:889-928: mean_T(audio) ‖ mean_T(text) → Linear 2d→d, ReLU, Dropout(.3) → Linear d→256, ReLU,
    Dropout(.3), Linear 256→C."""

    def __init__(self, audio_extractor_model, text_extractor_model, hidden_size, class_num):
        super().__init__()
        self.audio_extractor = audio_extractor_model
        self.text_extractor = text_extractor_model
        self.modality_fusion_module = nn.Sequential(
            nn.Linear(hidden_size * 2, hidden_size),
            nn.ReLU(),
            nn.Dropout(0.3),
        )
        self.output_classifier = nn.Sequential(
            nn.Linear(hidden_size, 256),
            nn.ReLU(),
            nn.Dropout(0.3),
            nn.Linear(256, class_num),
        )

    def forward(self, x):
        data = {names[0]: t for names, t in x}
        audio = self.audio_extractor(data['audio'])
        text = self.text_extractor(data['text'])
        if ops.probing():
            return audio.new_zeros(audio.shape[0], self.output_classifier[3].out_features)
        cat = torch.cat([ops.mean_pool(audio), ops.mean_pool(text)], dim=-1)
        lin = self.modality_fusion_module[0]
        fused = ops.linear(cat, lin.weight, lin.bias, relu_pre=True,
                           dropout_p=self.modality_fusion_module[2].p if self.training else 0.0)
        return _mlp_head(self.output_classifier, 0, 3, fused, self.output_classifier[2].p, self.training)
