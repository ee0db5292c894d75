"""Shared test plumbing: build a drop-in model + the matching oracle call for a golden case, and the
comparison metrics (the tolerances of BASELINE.json's north_star are written here)."""
from __future__ import annotations

import torch

from multimodalaggressionrecognition_b200 import workloads as W
from oracle import oracle as O

FP32_TOL = 1e-4      # fp32 mode vs fp32/fp64 reference
BF16_TOL = 2e-2      # bf16 mode vs fp32 reference ("about 1e-2 relative"): measured as ||a-b|| / ||b||
# bf16 gradients: the bar is on the WHOLE gradient (all parameters concatenated, ||Δ||/||g|| ≤ BF16_TOL);
# single tensors that are sums of cancelling terms (2-element logit biases, LayerNorm gains deep in the stack)
# carry a larger relative error in any bf16 implementation — tests/test_models_gpu.py::test_bf16_error_vs_torch_bf16
# calibrates this against torch's own bf16 kernels on the same model.
BF16_TENSOR_TOL = 8e-2
# At the tiny batches of the golden cases (B·T ≤ 200 tokens) bf16 gradients of ANY implementation are dominated by
# ReLU boundary flips (a pre-activation within bf16 rounding of 0 lands on the other side and moves one token's
# whole contribution to a weight-gradient row), so an absolute bar is meaningless there.  `bf16_floor()` measures
# what plain torch bf16 arithmetic does on the very same case (the oracle's math run with bf16 tensors instead of
# fp32) and the bf16 gradient bars are  max(absolute bar, BF16_VS_TORCH × that floor).  The factor is the same 1.5 the
# per-tensor bar uses: on the most ill-conditioned golden case (c3_audio_empty: 4 clips, phys head only) torch's own
# bf16 arithmetic is 14.5 % / 16.6 % off on the two recorded tensors and this repo 15.7 % / 19.5 % — the same noise,
# a different draw of ReLU flips (measured on B200, round 1).
BF16_VS_TORCH = 1.5
# Single tensors get 2x their torch-bf16 floor: a tensor's error is one draw of rounding noise on each side, and some
# tensors are mostly noise by construction — a third of every `in_proj_bias` gradient is the key bias, whose true
# gradient is identically 0 (softmax is invariant to a constant added to all keys), e.g. golden_v2's
# c3x_audio_text_ragged: torch bf16 11.1 % off on the fusion in_proj_bias, this repo 18.5 % (measured on B200).
BF16_TENSOR_VS_TORCH = 2.0
KINDS = {"GRU_1L": "gru", "LSTM_1L": "lstm", "Avg_features": "avg"}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, tol, what=""):
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e} (max-norm {max_err(a, b):.3e})"


def assert_grad_close(a, b, tol, what="", max_flipped_rows=2):
    """Gradient comparison that tolerates ReLU boundary flips: a pre-activation within rounding of 0 may be
    on either side in two correct implementations, which moves ONE row of the following weight gradient
    (oracle/make_golden.py picks seeds where the reference's own fp32 and fp64 runs agree; the CUDA path can
    still flip).  All but `max_flipped_rows` rows must meet `tol`; the whole tensor must meet 50*tol."""
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = rel_err(a, b)
    if e <= tol:
        return
    a2 = a.detach().double().cpu().reshape(a.shape[0], -1) if a.dim() > 1 else a.detach().double().cpu().reshape(-1, 1)
    b2 = b.detach().double().cpu().reshape(a2.shape)
    scale = b2.norm() / (b2.shape[0] ** 0.5)
    row_err = (a2 - b2).norm(dim=1) / scale.clamp_min(1e-30)
    bad = int((row_err > tol * 3).sum())
    assert bad <= max_flipped_rows and e <= 50 * tol, \
        f"{what}: relative error {e:.3e} > {tol:.1e} with {bad} rows off (not a ReLU-flip pattern)"


def global_rel_err(got: dict, ref: dict) -> float:
    """||concat(got) - concat(ref)|| / ||concat(ref)|| over the keys of ref that have a value."""
    num = den = 0.0
    for k, r in ref.items():
        if r is None:
            continue
        g = got[k]
        num += float((g.detach().double().cpu() - r.detach().double().cpu()).pow(2).sum())
        den += float(r.detach().double().cpu().pow(2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def bf16_floor(spec, model, batch_cpu):
    """Error of plain torch bf16 arithmetic on this case: the oracle's math evaluated once with fp32 tensors and
    once with every weight/activation held in bf16 (CPU).  Returns {"global": whole-gradient rel err,
    "tensor": {param: rel err}, "logits": {head: rel err}}."""
    data, labels = batch_cpu

    def run(dt):
        sd = {k: v.detach().cpu().clone().to(dt).requires_grad_(True) for k, v in model.state_dict().items()}
        d = data.to(dt) if isinstance(data, torch.Tensor) else [[n, t.to(dt)] for n, t in data]
        po = {k: v.float() for k, v in oracle_forward(spec, sd, d, True, True).items()}
        sum(oracle_losses(spec, po, labels).values()).backward()
        return {k: (None if v.grad is None else v.grad.float()) for k, v in sd.items()}, po

    ref, pr = run(torch.float32)
    got, pg = run(torch.bfloat16)
    return {"global": global_rel_err(got, ref),
            "tensor": {k: rel_err(got[k], r) for k, r in ref.items() if r is not None and float(r.norm()) > 0},
            "logits": {k: rel_err(pg[k].detach(), pr[k].detach()) for k in pr}}


def assert_bf16_grads(got: dict, ref: dict, floor: dict, what=""):
    """bf16 gradient bar: whole gradient ≤ max(BF16_TOL, BF16_VS_TORCH × torch-bf16 error on the same case); every
    tensor ≤ max(BF16_TENSOR_TOL, BF16_TENSOR_VS_TORCH × torch-bf16 error of that tensor)."""
    e = global_rel_err(got, ref)
    bar = max(BF16_TOL, BF16_VS_TORCH * floor["global"])
    assert e <= bar, f"{what}: whole-gradient relative error {e:.3e} > {bar:.3e} (torch bf16 on this case: {floor['global']:.3e})"
    for k, r in ref.items():
        if r is None or float(r.norm()) == 0:
            continue
        ek = rel_err(got[k], r)
        bk = max(BF16_TENSOR_TOL, BF16_TENSOR_VS_TORCH * floor["tensor"].get(k, 0.0))
        assert ek <= bk, f"{what}/{k}: relative error {ek:.3e} > {bk:.3e} (torch bf16: {floor['tensor'].get(k, 0.0):.3e})"
    return e


def build_case(spec, ns, device=None, init_seed=1234):
    """The case's model (dropout off, norms perturbed exactly like oracle/make_golden.py) and batch."""
    torch.manual_seed(init_seed)
    model = W.perturb_norms(W.disable_dropout(getattr(W, spec["builder"])(ns, **spec["bkw"])))
    batch = getattr(W, spec["batch"])(**spec["dkw"])
    if device is not None:
        model = model.to(device)
        batch = W.to_device(batch, device)
    return model, batch


def _c3_cfg(spec):
    bkw = spec["bkw"]
    if spec["builder"] == "build_c3x":
        return W.c3x_oracle_cfg(**bkw)
    return W.c3_oracle_cfg(bkw.get("t_audio", 250), bkw.get("t_video", 64), bkw.get("d", 768), bkw.get("heads", 8))


def oracle_forward(spec, sd, data, training, grad_enabled):
    b, bkw = spec["builder"], spec["bkw"]
    if b == "build_c1":
        h = O.transformer_sequence_processor(data, sd, "0.", bkw.get("layers", 2), bkw.get("heads", 8), "identity", training)
        return {"logits": O.output_classifier(h, sd, "1.", training)}
    if b == "build_c2":
        return O.video_multi_nn(data, sd, {h: KINDS[h] for h in bkw.get("heads", ("GRU_1L",))}, training)
    if b == "build_audio_text":
        heads = bkw.get("heads", 8)
        cfg = {"audio": {"layers": 1, "heads": heads, "extractor": "identity"},
               "text": {"layers": bkw.get("text_layers", 2), "heads": heads, "extractor": "identity"}}
        return {"logits": O.audio_text_model(data, sd, cfg, training)}
    return O.physverb_model(data, sd, _c3_cfg(spec), training, grad_enabled)


def _ce_weights(spec, like: torch.Tensor):
    w = spec.get("ce_weights")
    return None if w is None else {k: torch.tensor(v, dtype=like.dtype, device=like.device) for k, v in w.items()}


def oracle_losses(spec, pred, labels):
    b = spec["builder"]
    if b in ("build_c1", "build_audio_text"):
        return {"loss": O.cross_entropy(pred["logits"], labels)}
    if b == "build_c2" or spec["bkw"].get("top") == "old":
        return O.multi_ce(pred, labels)
    return O.multimodal_ce(pred, labels, weights=_ce_weights(spec, next(iter(pred.values()))), heads=["phys", "verb"])


def model_losses(spec, ns, model, batch):
    """Drop-in forward + losses through the reference-facing API (same calls trainer.py makes).  `ns` is the
    namespace the classes come from: this package's `models`, or the reference's (oracle/make_golden.py)."""
    data, labels = batch
    pred = model(data)
    b = spec["builder"]
    if b in ("build_c1", "build_audio_text"):
        crit = ns.MultiCrossEntropyLoss()
        return {"logits": pred}, crit({"loss": pred}, labels)
    if b == "build_c2" or spec["bkw"].get("top") == "old":
        return pred, ns.MultiCrossEntropyLoss()(pred, labels)
    w = _ce_weights(spec, next(iter(pred.values()))) or {}
    crit = ns.MultiModalCrossEntropyLoss({k: torch.nn.CrossEntropyLoss(weight=w.get(k)) for k in ("phys", "verb")})
    return pred, crit(pred, labels)


# ---- epoch bookkeeping of the reference's trainers (oracle/make_golden_trainer.py, tests/test_host_logic_cpu.py) ----
TRAINER_STREAMS = {
    # TorchSupervisedTrainer: one tensor of logits, one loss, labels (B,)
    "single": dict(kind="single", steps=7, B=16, seed=11, dataset_size=7 * 16 - 5),
    # RNN_trainer: {head: logits}, {head: loss}, the same labels for every head (train_video_rnn.py)
    "multi_head": dict(kind="multi_head", steps=5, B=12, seed=12, heads=("LSTM_1L", "GRU_1L", "Avg_features"), dataset_size=60),
    # MultimodalTrainer: label groups with whole-group and per-row `_EMPTY` markers (datasets.py:592-608)
    "multimodal": dict(kind="multimodal", steps=9, B=10, seed=13, heads=("phys", "verb"), dataset_size=90),
}


def trainer_metrics():
    """metrics_dict as the train_*.py scripts build it (train_multimodal.py:515-524)."""
    from sklearn import metrics
    return {
        'loss': None,
        'accuracy': metrics.accuracy_score,
        'recall': {'metric': metrics.recall_score, 'kwargs': {'average': None, 'zero_division': 0}},
        'UAR': {'metric': metrics.recall_score, 'kwargs': {'average': 'macro', 'zero_division': 0}},
    }


def trainer_stream(kind, steps, B, seed, heads=("loss",), dataset_size=None, device=None):
    """A seeded stream of (data, losses, pred, labels) as a trainer's step sees them.  For 'multimodal', step i's
    label groups cycle through: all present / phys group all-EMPTY / verb group all-EMPTY / a few EMPTY rows in one
    group (the non-homogeneous case `create_batch_results_dict` filters row-wise, trainer.py:893-905)."""
    g = torch.Generator().manual_seed(seed)
    dev = device or torch.device("cpu")
    out = []
    for i in range(steps):
        b = B if i < steps - 1 else max(1, B - 3)              # a shorter last batch
        if kind == "single":
            logits = torch.randn(b, 2, generator=g)
            out.append((torch.zeros(b, 4, 8), torch.rand((), generator=g).to(dev), logits.to(dev), torch.randint(0, 2, (b,), generator=g).to(dev)))
            continue
        pred = {h: torch.randn(b, 2, generator=g).to(dev) for h in heads}
        losses = {h: torch.rand((), generator=g).to(dev) for h in heads}
        if kind == "multi_head":
            out.append((torch.zeros(b, 4, 8), losses, pred, torch.randint(0, 2, (b,), generator=g).to(dev)))
            continue
        data = [[("audio",) * b, torch.zeros(b, 4, 8)], [("video",) * b, torch.zeros(b, 2, 8)]]
        labels = []
        for h in ("verb", "phys"):
            y = torch.randint(0, 2, (b,), generator=g)
            names = [h] * b
            mode = i % 4
            if (mode == 1 and h == "phys") or (mode == 2 and h == "verb"):
                names = [h + "_EMPTY"] * b
                y = torch.full_like(y, -1)
                losses.pop(h)                               # MultiModalCrossEntropyLoss emits no loss for an all-EMPTY group
            elif mode == 3 and h == "phys":
                for r in range(0, b, 3):
                    names[r] = h + "_EMPTY"
                    y[r] = -1
            labels.append([tuple(names), y.to(dev)])
        out.append((data, losses, pred, labels))
    return out


# ---- host stand-ins for the two device calls of training.py (the package has no CPU arithmetic) --------------------
def host_adam_step(self) -> None:
    """What `mar_adam_step_segments` computes, in torch on the host: per-parameter Adam over the flat buffers, the
    parameters whose header flag is 0 untouched.  Installed by the gloo / host-logic tests only."""
    f = self.flat
    b1, b2 = self.betas
    for i, (p, o) in enumerate(zip(f.params, f.offsets)):
        if not float(f.flags[i]) > 0:
            continue
        self.seg_steps[i] += 1
        t = float(self.seg_steps[i])
        sl = slice(o, o + p.numel())
        g = f.grad[sl]
        self.exp_avg[sl].mul_(b1).add_(g, alpha=1 - b1)
        self.exp_avg_sq[sl].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (self.exp_avg_sq[sl].sqrt() / (1 - b2 ** t) ** 0.5).add_(self.eps)
        f.flat[sl].addcdiv_(self.exp_avg[sl], denom, value=-self.lr / (1 - b1 ** t))


def install_host_stand_ins() -> None:
    """Call at the start of every CPU process that drives training.TrainStep / EpochAccumulator without a GPU."""
    from multimodalaggressionrecognition_b200 import training
    training.FlatAdam.step = host_adam_step
    training.EpochAccumulator._argmax = staticmethod(lambda logits: logits.detach().argmax(dim=1))
