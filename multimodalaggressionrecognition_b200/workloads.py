"""The BASELINE.json configurations as builders that work on ANY namespace exposing
the reference's class names: the reference's own `models` module (used only by oracle/make_golden.py
in the build container), or this package's drop-in `models`.  Also the synthetic batches in the exact
`datasets.py` layouts, and the algorithmic FLOP counts the roofline is computed from."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

MODALITY2AGGR = {'video': 'phys', 'text': 'verb', 'audio': 'verb'}


# ---- models ---------------------------------------------------------------------------------
def build_c1(ns, d: int = 768, layers: int = 2, heads: int = 8, classes: int = 2) -> nn.Module:
    """C1 audio Transformer classifier built only from live reference classes."""
    return nn.Sequential(ns.TransformerSequenceProcessor(nn.Sequential(), d, layers, heads, classes),
                         ns.OutputClassifier(d, classes))


def build_c2(ns, d: int = 512, classes: int = 2, heads: Tuple[str, ...] = ("GRU_1L",)) -> nn.Module:
    """C2 video RNN heads (train_video_rnn.py:93-133)."""
    kinds = {
        "GRU_1L": {'model': nn.GRU, 'kwargs': {'input_size': d, 'hidden_size': d, 'num_layers': 1, 'batch_first': True}},
        "LSTM_1L": {'model': nn.LSTM, 'kwargs': {'input_size': d, 'hidden_size': d, 'num_layers': 1, 'batch_first': True}},
        "Avg_features": {'model': ns.AverageFeatureSequence, 'kwargs': {'hidden_size': d}},
    }
    return ns.VideoMultiNN({h: ns.FeatureSequenceProcessing(kinds[h], classes) for h in heads})


def build_c3(ns, t_audio: int = 250, t_video: int = 64, d: int = 768, d_video_in: int = 512, heads: int = 8,
             classes: int = 2) -> nn.Module:
    """C3 audio+video Transformer fusion (train_multimodal.py:298-420 with Transformer extractors)."""
    extractors = nn.ModuleDict({
        'audio': ns.TransformerSequenceProcessor(nn.Sequential(), d, 1, heads, classes),
        'video': ns.TransformerSequenceProcessor(ns.EmbeddingLayer(d_video_in, d), d, 1, heads, classes),
    })
    fusion = ns.EqualSizedTransformerModalitiesFusion(1, d, heads)
    classifiers = ns.PhysVerbClassifierConcatFeatures(
        ['audio', 'video'], classes, {'video': [d, d], 'audio': [d, d], 'text': [d, d]}, dict(MODALITY2AGGR))
    return ns.PhysVerbModel(extractors, fusion, classifiers, {'audio': [t_audio, d], 'video': [t_video, d]},
                            dict(MODALITY2AGGR), d, classes)


def c3_oracle_cfg(t_audio: int = 250, t_video: int = 64, d: int = 768, heads: int = 8) -> dict:
    return {
        "feature_shapes": {'audio': [t_audio, d], 'video': [t_video, d]},
        "extractors": {'audio': {"layers": 1, "heads": heads, "extractor": "identity"},
                       'video': {"layers": 1, "heads": heads, "extractor": "embedding"}},
        "fusion_layers": 1, "fusion_heads": heads, "aggr_types": ['phys', 'verb'],
    }


def build_c3x(ns, modalities: Tuple[str, ...] = ("audio", "video"), fusion: str = "equal", classifier: str = "concat",
              top: str = "physverb", t_audio: int = 250, t_video: int = 64, t_text: int = 48, d: int = 768,
              d_video_in: int = 512, heads: int = 8, classes: int = 2) -> nn.Module:
    """The other assemblies train_multimodal.py can build: any subset of
    audio / text / video (text enters through `nn.Sequential()`, zero-padded RuBERT tokens, train_multimodal.py:365),
    `EqualSizedTransformerModalitiesFusion` or `AveragedFeaturesTransformerFusion` (:374-375),
    `PhysVerbClassifierConcatFeatures` or the base `PhysVerbClassifier` (:406-411), under `PhysVerbModel` or
    the older `MultimodalModel` with one `OutputClassifier` per modality (:425-443)."""
    makers = {
        'audio': lambda: ns.TransformerSequenceProcessor(nn.Sequential(), d, 1, heads, classes),
        'text': lambda: nn.Sequential(),
        'video': lambda: ns.TransformerSequenceProcessor(ns.EmbeddingLayer(d_video_in, d), d, 1, heads, classes),
    }
    shapes = {'audio': [t_audio, d], 'text': [t_text, d], 'video': [t_video, d]}
    extractors = nn.ModuleDict({m: makers[m]() for m in ('audio', 'text', 'video') if m in modalities})
    fusion_cls = {"equal": ns.EqualSizedTransformerModalitiesFusion, "avg": ns.AveragedFeaturesTransformerFusion}[fusion]
    fusion_module = fusion_cls(1, d, heads)
    shapes = {m: s for m, s in shapes.items() if m in modalities}
    if top == "old":
        heads_dict = nn.ModuleDict({m: ns.OutputClassifier(d, classes) for m in ('audio', 'text', 'video') if m in modalities})
        return ns.MultimodalModel(extractors, fusion_module, heads_dict, shapes, d, classes)
    clf_cls = {"concat": ns.PhysVerbClassifierConcatFeatures, "base": ns.PhysVerbClassifier}[classifier]
    classifiers = clf_cls(list(modalities), classes, {'video': [d, d], 'audio': [d, d], 'text': [d, d]}, dict(MODALITY2AGGR))
    return ns.PhysVerbModel(extractors, fusion_module, classifiers, shapes, dict(MODALITY2AGGR), d, classes)


def c3x_oracle_cfg(modalities=("audio", "video"), fusion: str = "equal", classifier: str = "concat", top: str = "physverb",
                   t_audio: int = 250, t_video: int = 64, t_text: int = 48, d: int = 768, heads: int = 8, **_) -> dict:
    ex = {'audio': {"layers": 1, "heads": heads, "extractor": "identity"},
          'text': {"layers": 0, "heads": heads, "extractor": "identity"},      # nn.Sequential(): tokens pass through
          'video': {"layers": 1, "heads": heads, "extractor": "embedding"}}
    shapes = {'audio': [t_audio, d], 'text': [t_text, d], 'video': [t_video, d]}
    aggr = ['phys', 'verb']        # ConcatFeatures builds a head per value of modality2aggr, whatever the modalities
    return {
        "feature_shapes": {m: shapes[m] for m in modalities}, "extractors": {m: ex[m] for m in modalities},
        "fusion_layers": 1, "fusion_heads": heads, "aggr_types": aggr, "fusion": fusion, "classifier": classifier,
        "top": top, "modality2aggr": dict(MODALITY2AGGR),
    }


def build_audio_text(ns, d: int = 768, heads: int = 8, classes: int = 2, text_layers: int = 2) -> nn.Module:
    """AudioTextualModel (models.py:889-928) as train_audio_text.py:158-178 assembles it, with the audio branch's
    out-of-scope CNN front end replaced by pre-computed d-wide features."""
    audio = ns.TransformerSequenceProcessor(nn.Sequential(), d, 1, heads, classes)
    text = ns.TransformerSequenceProcessor(nn.Sequential(), d, text_layers, heads, classes)
    return ns.AudioTextualModel(audio, text, d, classes)


def disable_dropout(model: nn.Module) -> nn.Module:
    """Exact-parity recipe: every nn.Dropout.p = 0 and every self_attn.dropout = 0."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model


def perturb_norms(model: nn.Module, seed: int = 7, scale: float = 0.1) -> nn.Module:
    """Move every LayerNorm off its (γ=1, β=0) initial point, deterministically.

    Why parity tests need it: at initialisation each encoder's output rows are LayerNorm outputs with β = 0,
    whose sum over d is 0 in exact arithmetic and rounding noise in floating point — and the reference's
    fusion masks every key whose feature row sums to EXACTLY 0 (models.py:421-422).  Which tokens hit
    exactly 0.0 then depends on the summation order of the implementation (torch fp32 vs fp64 already
    disagree), so outputs are only comparable once β ≠ 0, as after the first optimizer steps."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.LayerNorm):
                m.weight.add_(scale * torch.randn(m.weight.shape, generator=g).to(m.weight))
                m.bias.add_(scale * torch.randn(m.bias.shape, generator=g).to(m.bias))
    return model


# ---- synthetic batches ------------------------------------------------------------------------
def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed)


def batch_c1(B: int = 32, T: int = 250, d: int = 768, seed: int = 1000):
    g = _gen(seed)
    return torch.randn(B, T, d, generator=g), torch.randint(0, 2, (B,), generator=g)


def batch_c1_learnable(B: int = 32, T: int = 250, d: int = 768, seed: int = 1000, shift: float = 0.06):
    """`batch_c1` with a label the model can learn (the first 32 feature columns are shifted by ±shift according to
    the label), so that a loss curve over an epoch goes somewhere and the trained model's predictions differ
    between clips."""
    x, y = batch_c1(B, T, d, seed)
    x[:, :, :32] += shift * (2.0 * y.float() - 1.0)[:, None, None]
    return x, y


def batch_c2(B: int = 64, T: int = 64, d: int = 512, seed: int = 1000):
    g = _gen(seed)
    return torch.randn(B, T, d, generator=g), torch.randint(0, 2, (B,), generator=g)


def batch_c3(B: int = 256, t_audio: int = 250, t_video: int = 64, d_audio: int = 768, d_video: int = 512,
             seed: int = 1000, empty: Optional[str] = None, zero_pad_audio: int = 0):
    """`[data, labels]` in the MultimodalPhysVerbDataset layout (datasets.py:564-608).
    empty='video' → a verb-only batch (video stub = -1, phys labels -1 / 'phys_EMPTY');
    empty='audio' → a phys-only batch.  zero_pad_audio > 0 zero-fills the last audio frames."""
    g = _gen(seed)
    audio = torch.randn(B, t_audio, d_audio, generator=g)
    video = torch.randn(B, t_video, d_video, generator=g)
    y_verb = torch.randint(0, 2, (B,), generator=g)
    y_phys = torch.randint(0, 2, (B,), generator=g)
    if zero_pad_audio:
        audio[:, t_audio - zero_pad_audio:] = 0.0
    a_name, v_name, verb_name, phys_name = 'audio', 'video', 'verb', 'phys'
    if empty == 'video':
        video = torch.full_like(video, -1.0); v_name = 'video_EMPTY'
        y_phys = torch.full_like(y_phys, -1); phys_name = 'phys_EMPTY'
    elif empty == 'audio':
        audio = torch.full_like(audio, -1.0); a_name = 'audio_EMPTY'
        y_verb = torch.full_like(y_verb, -1); verb_name = 'verb_EMPTY'
    data = [[(a_name,) * B, audio], [(v_name,) * B, video]]
    labels = [[(verb_name,) * B, y_verb], [(phys_name,) * B, y_phys]]
    return data, labels


def _ragged_zero_pad(x: torch.Tensor, g: torch.Generator, min_len: int = 1) -> torch.Tensor:
    """AppendZeroValues (datasets.py:212-231): every sample keeps a random number of leading tokens, the rest is 0."""
    B, T, _ = x.shape
    lens = torch.randint(min_len, T + 1, (B,), generator=g)
    lens[0] = T                                   # at least one full-length sample
    x = x.clone()
    for i in range(B):
        x[i, int(lens[i]):] = 0.0
    return x


def batch_c3x(B: int = 8, modalities=("audio", "video"), t_audio: int = 250, t_video: int = 64, t_text: int = 48,
              d_audio: int = 768, d_video: int = 512, d_text: int = 768, seed: int = 1000, empty: Optional[str] = None,
              ragged_text: bool = True, flat_labels: bool = False, **_):
    """`batch_c3` for any modality subset; text tokens are zero-padded to `t_text` with per-sample lengths (the
    real key-padding masks of the fusion encoder).  flat_labels → `(B,)` labels for the
    older MultimodalModel + MultiCrossEntropyLoss."""
    g = _gen(seed)
    shapes = {'audio': (t_audio, d_audio), 'text': (t_text, d_text), 'video': (t_video, d_video)}
    data = []
    for m in modalities:
        x = torch.randn(B, *shapes[m], generator=g)
        if m == 'text' and ragged_text:
            x = _ragged_zero_pad(x, g)
        name = m
        if empty == m:
            x = torch.full_like(x, -1.0); name = m + '_EMPTY'
        data.append([(name,) * B, x])
    y_verb = torch.randint(0, 2, (B,), generator=g)
    y_phys = torch.randint(0, 2, (B,), generator=g)
    if flat_labels:
        return data, y_verb
    verb_name, phys_name = 'verb', 'phys'
    if empty == 'video':
        y_phys = torch.full_like(y_phys, -1); phys_name = 'phys_EMPTY'
    return data, [[(verb_name,) * B, y_verb], [(phys_name,) * B, y_phys]]


def batch_audio_text(B: int = 4, t_audio: int = 50, t_text: int = 48, d: int = 768, seed: int = 1000):
    """train_audio_text.py batches: [[('audio',)*B, (B,T_a,d)], [('text',)*B, (B,T_t,d) zero-padded]], labels (B,)."""
    g = _gen(seed)
    audio = torch.randn(B, t_audio, d, generator=g)
    text = _ragged_zero_pad(torch.randn(B, t_text, d, generator=g), g)
    return [[('audio',) * B, audio], [('text',) * B, text]], torch.randint(0, 2, (B,), generator=g)


def batch_c3_mixed(B: int = 6, t_audio: int = 250, t_video: int = 64, seed: int = 1000, no_video=(1, 4), no_audio=(2,)):
    """A NON-homogeneous batch (the reference's sampler never builds one, its model and loss handle it row by row,
    models.py:840-860, :244-258): clips `no_video` lack the video modality and their phys label, clips `no_audio`
    lack audio and their verb label."""
    data, labels = batch_c3(B=B, t_audio=t_audio, t_video=t_video, seed=seed)
    names = {m: list(n) for m, (n, _) in zip(("audio", "video"), data)}
    lab = {m: list(n) for m, (n, _) in zip(("verb", "phys"), labels)}
    for i in no_video:
        names["video"][i] = "video_EMPTY"; data[1][1][i] = -1.0
        lab["phys"][i] = "phys_EMPTY"; labels[1][1][i] = -1
    for i in no_audio:
        names["audio"][i] = "audio_EMPTY"; data[0][1][i] = -1.0
        lab["verb"][i] = "verb_EMPTY"; labels[0][1][i] = -1
    data = [[tuple(names["audio"]), data[0][1]], [tuple(names["video"]), data[1][1]]]
    labels = [[tuple(lab["verb"]), labels[0][1]], [tuple(lab["phys"]), labels[1][1]]]
    return data, labels


def to_device(batch, device):
    """Move a (nested list) batch to the device, as datasets.py does at construction (datasets.py:493-561)."""
    if isinstance(batch, torch.Tensor):
        return batch.to(device)
    if isinstance(batch, (list, tuple)) and batch and isinstance(batch[0], str):
        return batch
    if isinstance(batch, (list, tuple)):
        return [to_device(b, device) for b in batch]
    return batch


# ---- algorithmic FLOPs (multiply-add = 2; backward = 2x forward; softmax/LN/elementwise excluded) ----
def encoder_layer_flops(T: int, d: int = 768, d_ff: int = 2048) -> Dict[str, float]:
    return {"gemm": T * (8 * d * d + 4 * d * d_ff), "attn": 4 * T * T * d}


def c1_flops_per_clip(T: int = 250, d: int = 768, layers: int = 2, classes: int = 2) -> float:
    e = encoder_layer_flops(T, d)
    return layers * (e["gemm"] + e["attn"]) + 2 * d * 256 + 2 * 256 * classes


def c2_flops_per_clip(T: int = 64, I: int = 512, H: int = 512, classes: int = 2) -> float:
    return T * (2 * I * 3 * H + 2 * H * 3 * H) + 2 * H * 256 + 2 * 256 * classes


def c3_flops_per_clip(t_audio: int = 250, t_video: int = 64, d: int = 768, d_video_in: int = 512,
                      classes: int = 2) -> Dict[str, float]:
    ea, ev, ef = encoder_layer_flops(t_audio, d), encoder_layer_flops(t_video, d), encoder_layer_flops(t_audio + t_video, d)
    gemm = ea["gemm"] + ev["gemm"] + ef["gemm"] + 2 * t_video * d_video_in * d + 2 * (t_audio + t_video) * d * d \
        + 2 * (2 * (2 * d) * (2 * d // 3) + 2 * (2 * d // 3) * classes)
    attn = ea["attn"] + ev["attn"] + ef["attn"]
    return {"gemm": float(gemm), "attn": float(attn), "total": float(gemm + attn)}
