"""GPU: the drop-in modules, through the reference-facing nn.Module API, against (a) the golden vectors
recorded from the live reference and (b) the CPU oracle on fresh seeded inputs, in fp32 and bf16 mode;
plus size-independent properties at BASELINE.json's full sizes."""
import copy

import pytest
import torch

import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import models as M, ops, workloads as W
from oracle import oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CASES = ["c1_small", "c1_odd", "c2_small", "c3_small", "c3_video_empty", "c3_audio_empty", "c3_audio_padded"]
CASES_V2 = ["c3x_audio_text_ragged", "c3x_three_modalities", "c3x_three_video_empty", "c3x_avg_fusion",
            "c3x_avg_fusion_video_empty", "c3x_base_classifier", "c3x_old_multimodal_model", "c3_weighted_ce",
            "audio_text_model"]


@pytest.fixture(autouse=True)
def _no_dropout():
    O.DROPOUT_ENABLED = False
    ops.clear_weight_cache()
    yield


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", CASES + CASES_V2)
def test_golden_parity(golden, name, mode):
    """eval outputs (incl. the nested-tensor zero-fill path), train logits, per-head losses, every parameter
    gradient — against what the reference itself produced."""
    case = golden["cases"][name]
    spec = case["spec"]
    tol = H.FP32_TOL if mode == "fp32" else H.BF16_TOL
    model, batch = H.build_case(spec, M, DEV)
    floor, eval_bar = None, {}
    if mode == "bf16":
        # what torch's own bf16 arithmetic does on this case (logits of a near-initialisation model are small
        # differences of large terms: on some of the tiny golden batches plain torch bf16 is itself 2e-2 off);
        # the bf16 bars are max(the absolute bar, 1.5 x that floor), as for the gradients (tests/helpers.py)
        floor = H.bf16_floor(spec, model, getattr(W, spec["batch"])(**spec["dkw"]))
        eval_bar = {k: max(tol, H.BF16_VS_TORCH * e) for k, e in floor["logits"].items()}
    with mar.precision(mode):
        model.eval()
        if "eval" in case:
            with torch.no_grad():
                pred = model(batch[0])
            pred = pred if isinstance(pred, dict) else {"logits": pred}
            for k, v in case["eval"].items():
                assert pred[k].dtype == torch.float32
                H.assert_close(pred[k].cpu(), v, tol if mode == "fp32" else eval_bar.get(k, tol), f"{name}/{mode}/eval/{k}")
        else:
            with torch.no_grad(), pytest.raises(RuntimeError):
                model(batch[0])
        model.train()
        model.zero_grad()
        pred, losses = H.model_losses(spec, M, model, batch)
        assert set(losses) == set(case["losses"])
        losses.backward()
    for k, v in case["train"].items():
        bar = eval_bar.get(k, tol)
        H.assert_close(pred[k].cpu(), v, bar, f"{name}/{mode}/train/{k}")
    for k, v in case["losses"].items():
        assert abs(float(losses[k]) - v) <= (2e-5 if mode == "fp32" else 5e-3), f"{name}/{mode}/loss/{k}"
    grads = {k: p.grad for k, p in model.named_parameters()}
    for k, n in case["grad_norms"].items():
        g = grads[k]
        if n is None:
            assert g is None or float(g.abs().max()) == 0.0, f"{name}: {k} must have no gradient"
        else:
            assert g is not None, f"{name}: {k} has no gradient"
            lim = 2e-3 * n if mode == "fp32" else H.BF16_TENSOR_TOL * n + 5e-4
            assert abs(float(g.norm()) - n) <= lim, f"{name}/{mode}: |grad {k}| = {float(g.norm())} vs {n}"
    # the norm of the whole gradient
    tot = sum(n * n for n in case["grad_norms"].values() if n is not None) ** 0.5
    got = sum(float(g.norm()) ** 2 for g in grads.values() if g is not None) ** 0.5
    assert abs(got - tot) <= (1e-3 if mode == "fp32" else H.BF16_TOL) * tot
    if mode == "fp32":
        for k, v in case["grads"].items():
            H.assert_grad_close(grads[k].cpu(), v, tol, f"{name}/{mode}/grad/{k}")
    else:
        H.assert_bf16_grads({k: grads[k] for k in case["grads"]}, case["grads"], floor, f"{name}/bf16/grad")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["c1_small", "c2_small", "c3_small"])
def test_all_gradients_against_oracle(golden, name, mode):
    """Every parameter gradient (not only the recorded ones) against the oracle run here on the same weights."""
    spec = golden["cases"][name]["spec"]
    model, batch = H.build_case(spec, M, DEV)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    data, labels = getattr(W, spec["batch"])(**spec["dkw"])
    po = H.oracle_forward(spec, sd, data, True, True)
    sum(H.oracle_losses(spec, po, labels).values()).backward()
    with mar.precision(mode):
        model.train()
        _, losses = H.model_losses(spec, M, model, batch)
        losses.backward()
    got = {k: p.grad for k, p in model.named_parameters()}
    ref = {k: v.grad for k, v in sd.items()}
    if mode == "bf16":
        floor = H.bf16_floor(spec, model, (data, labels))
        e = H.assert_bf16_grads(got, ref, floor, f"{name}/bf16")
        print(f"{name}: whole-gradient bf16 error vs fp32 oracle: ours {e:.3e}, torch bf16 arithmetic {floor['global']:.3e}")
        return
    for k, r in ref.items():
        if r is not None:
            H.assert_grad_close(got[k].cpu(), r, H.FP32_TOL, f"{name}/{mode}/{k}")
    e = H.global_rel_err(got, ref)
    assert e <= H.FP32_TOL * 3, f"{name}/{mode}: whole-gradient relative error {e:.3e}"


def test_bf16_error_vs_torch_bf16():
    """Calibration of the bf16 tolerance: the same C1 model in torch's own bf16 kernels (plain torch.nn modules
    cast to bfloat16 on this GPU) against the fp32 oracle, beside ours.  Ours must not be worse than 1.5x torch's
    (+ a small floor)."""
    import torch.nn as nn
    torch.manual_seed(5)
    model = W.perturb_norms(W.disable_dropout(W.build_c1(M))).to(DEV).train()
    x, y = W.batch_c1(B=8, T=100, seed=1003)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    O.cross_entropy(O.output_classifier(O.transformer_sequence_processor(x, sd, "0.", 2, 8, "identity", True), sd, "1.", True), y).backward()
    ref = {k: v.grad for k, v in sd.items()}
    with mar.precision("bf16"):
        M.MultiCrossEntropyLoss()({"loss": model(x.to(DEV))}, y.to(DEV)).backward()
    ours = H.global_rel_err({k: p.grad for k, p in model.named_parameters()}, ref)

    enc_layer = nn.TransformerEncoderLayer(768, 8, batch_first=True)
    tmodel = nn.Sequential()
    tmodel.enc = nn.TransformerEncoder(enc_layer, 2, norm=nn.LayerNorm(768))
    tmodel.l1, tmodel.l2 = nn.Linear(768, 256), nn.Linear(256, 2)
    remap = {}
    for k, v in model.state_dict().items():
        k2 = k.replace("0.transformer_squence_processing.", "enc.").replace("1.classifier.1.", "l1.").replace("1.classifier.4.", "l2.")
        remap[k2] = v
    tmodel.load_state_dict(remap)
    W.disable_dropout(tmodel)
    tmodel = tmodel.to(DEV).to(torch.bfloat16).train()
    h = tmodel.enc(x.to(DEV).to(torch.bfloat16)).mean(dim=1)
    logits = tmodel.l2(torch.relu(tmodel.l1(h)))
    nn.functional.cross_entropy(logits.float(), y.to(DEV)).backward()
    tg = {}
    for k2, p in tmodel.named_parameters():
        k = k2.replace("enc.", "0.transformer_squence_processing.").replace("l1.", "1.classifier.1.").replace("l2.", "1.classifier.4.")
        tg[k] = p.grad.float()
    theirs = H.global_rel_err(tg, ref)
    print(f"whole-gradient bf16 error vs fp32 oracle: ours {ours:.3e}, torch bf16 {theirs:.3e}")
    assert ours <= 1.5 * theirs + 5e-3


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_adam_loss_curve(golden, mode):
    """3 Adam steps through torch.optim.Adam on the drop-in's parameters reproduce the reference's loss curve."""
    for name in ("c2_small", "c3_small"):
        case = golden["cases"][name]
        spec = case["spec"]
        model, batch = H.build_case(spec, M, DEV)
        model.train()
        opt = torch.optim.Adam(model.parameters())
        with mar.precision(mode):
            for ref_step in case["adam_curve"]:
                opt.zero_grad()
                _, losses = H.model_losses(spec, M, model, batch)
                losses.backward()
                opt.step()
                for k, v in ref_step.items():
                    assert abs(float(losses[k]) - v) <= (1e-3 if mode == "fp32" else 5e-2) * max(1.0, abs(v)), \
                        f"{name}/{mode}: loss curve {k}: {float(losses[k])} vs {v}"


def test_adam_loss_curve_other_assemblies(golden):
    """The 3-step Adam curves of golden_v2 (ragged text masks in the middle of the fused sequence, averaged fusion,
    base classifier heads, the older MultimodalModel, class-weighted CE), fp32 mode."""
    for name in ("c3x_three_modalities", "c3x_avg_fusion", "c3x_base_classifier", "c3x_old_multimodal_model", "c3_weighted_ce"):
        case = golden["cases"][name]
        spec = case["spec"]
        model, batch = H.build_case(spec, M, DEV)
        model.train()
        opt = torch.optim.Adam(model.parameters())
        with mar.precision("fp32"):
            for ref_step in case["adam_curve"]:
                opt.zero_grad()
                _, losses = H.model_losses(spec, M, model, batch)
                losses.backward()
                opt.step()
                for k, v in ref_step.items():
                    assert abs(float(losses[k]) - v) <= 1e-3 * max(1.0, abs(v)), f"{name}: loss curve {k}: {float(losses[k])} vs {v}"


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_c1_epoch_loss_curve_and_predictions(golden_c1_epoch, mode):
    """BASELINE config 1 at full size (B=32, T=250, d=768, 2 layers) over "1 epoch" = 48 Adam steps on 48 different
    batches: the per-step loss curve and the trained model's label predictions on a held-out batch, against what
    the LIVE reference produced (tests/golden/golden_c1_epoch.pt, `python oracle/make_golden.py c1_epoch`)."""
    g = golden_c1_epoch
    torch.manual_seed(g["init_seed"])
    model = W.perturb_norms(W.disable_dropout(W.build_c1(M)))
    chk = float(sum(v.double().abs().sum() for v in model.state_dict().values()))
    assert abs(chk - g["weights_checksum"]) <= 1e-9 * g["weights_checksum"]
    model = model.to(DEV).train()
    opt = torch.optim.Adam(model.parameters())
    crit = M.MultiCrossEntropyLoss()
    curve = []
    with mar.precision(mode):
        for i in range(g["steps"]):
            x, y = W.batch_c1_learnable(seed=g["batch_seed0"] + i)
            opt.zero_grad()
            losses = crit({"loss": model(x.to(DEV))}, y.to(DEV))
            losses.backward()
            opt.step()
            curve.append(losses["loss"].detach())
        x, y = W.batch_c1_learnable(seed=g["eval_seed"])
        model.eval()
        with torch.no_grad():
            logits = model(x.to(DEV)).float().cpu()
    curve = [float(c) for c in curve]
    ref = g["loss_curve"]
    dev = [abs(a - b) for a, b in zip(curve, ref)]
    epoch_loss, ref_epoch_loss = sum(curve) / len(curve), sum(ref) / len(ref)      # the per-epoch loss of trainer.py:258
    print(f"C1 epoch {mode}: max |loss - reference| {max(dev):.3e} at step {dev.index(max(dev))}, mean {sum(dev) / len(dev):.3e}; "
          f"epoch loss {epoch_loss:.5f} vs {ref_epoch_loss:.5f}")
    # The reference's curve: a plateau around 0.7-1.8 for ~15 steps, a sharp learning transition over steps 14-22,
    # then -> 1e-4.  fp32: the CPU oracle port is within 1e-5 of the reference on the plateau and 7e-4 off at the
    # steepest step; bars 2e-3 outside / 2e-2 inside the transition.  bf16: Adam's first updates are lr x sign(g), so
    # rounding noise in near-zero gradient components becomes +-lr kicks and the plateau is a chaotic transient —
    # WHEN the transition starts moves by a step between bf16 variants.  Measured (same weights, batches, fp32 master
    # parameters): the oracle with only weights/inputs rounded to bf16 is up to 6.7e-2 off on the plateau and 1.45e-1
    # at the transition; with ALL arithmetic in bf16 1.4e-2 / 8.9e-2, per-epoch loss +3.1 %; this repo's bf16 mode on
    # B200, two runs of the same build (the fp32 atomics' order differs run to run): 7.7e-2 / 2.6e-1 at step 18,
    # mean 3.5e-2, per-epoch loss +9.7 %, and 1.5e-1 at step 18, mean 2.4e-2, +6.6 %.  The bf16 bars are twice the
    # worst of those; the run must still learn (final loss < 1e-2) and reproduce every label prediction.  (fp32 mode
    # on B200: 6.8e-4 at step 18 — the same as the CPU oracle — mean 2.8e-5, per-epoch loss 0.29129 vs 0.29127.)
    plateau = [d for i, d in enumerate(dev) if i < 14 or i >= 24]
    if mode == "fp32":
        assert max(plateau) <= 2e-3 and max(dev) <= 2e-2
        assert abs(epoch_loss - ref_epoch_loss) <= 2e-3 * ref_epoch_loss
    else:
        assert max(plateau) <= 0.15 and max(dev) <= 0.5 and sum(dev) / len(dev) <= 7e-2
        assert abs(epoch_loss - ref_epoch_loss) <= 0.2 * ref_epoch_loss
    assert curve[-1] < 1e-2                                                          # it learned, like the reference
    ref_logits, ref_pred = g["eval_logits"], g["eval_pred"]
    assert float((ref_logits[:, 1] - ref_logits[:, 0]).abs().min()) > 5.0            # the reference decides every clip clearly
    assert torch.equal(logits.argmax(1), ref_pred), f"{mode}: label predictions differ from the reference's"
    H.assert_close(logits, ref_logits, 2e-2 if mode == "fp32" else 0.25, "held-out logits after the epoch")


def test_train_eval_asymmetry_of_masked_tokens(golden):
    """with a left-aligned padding mask, eval re-inserts padded tokens as zeros before the final
    LayerNorm (output = beta), train computes them as ordinary queries."""
    fusion = M.EqualSizedTransformerModalitiesFusion(1, 768, 8).to(DEV)
    W.perturb_norms(W.disable_dropout(fusion))
    a = torch.randn(2, 10, 768, device=DEV)
    v = torch.zeros(2, 4, 768, device=DEV)      # EMPTY modality → zero rows → masked keys, left-aligned
    beta = fusion.modality_fusion_transformer.norm.bias
    with mar.precision("fp32"):
        fusion.eval()
        with torch.no_grad():
            out_eval = fusion({"audio": a, "video": v})
        fusion.train()
        out_train = fusion({"audio": a, "video": v})
    H.assert_close(out_eval["video"][0, 0], beta, 1e-5, "eval: padded token = LN(0) = beta")
    assert float((out_train["video"][0, 0] - beta).abs().max()) > 1e-2
    H.assert_close(out_eval["audio"], out_train["audio"].detach(), 1e-5, "valid tokens agree between train and eval")


def test_per_head_backward_equals_reference_semantics(golden):
    """LossesDict.backward() (one pass over all heads) = the reference's per-head backward with retain_graph."""
    spec = golden["cases"]["c3_small"]["spec"]
    model, batch = H.build_case(spec, M, DEV)
    model.train()
    with mar.precision("fp32"):
        _, losses = H.model_losses(spec, M, model, batch)
        losses.backward()
        g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
        model.zero_grad()
        _, losses = H.model_losses(spec, M, model, batch)
        items = list(losses.items())
        for i, (_, l) in enumerate(items):
            l.backward(retain_graph=i != len(items) - 1)
    for k, p in model.named_parameters():
        H.assert_close(p.grad, g1[k], 1e-5, k)


def test_dropout_training_mode_runs_and_is_stochastic():
    model = W.build_c3(M, t_audio=24, t_video=8).to(DEV).train()
    batch = W.to_device(W.batch_c3(B=8, t_audio=24, t_video=8), DEV)
    with mar.precision("bf16"):
        a = model(batch[0])
        b = model(batch[0])
        model.eval()
        with torch.no_grad():
            c = model(batch[0])
            d = model(batch[0])
    assert float((a["phys"] - b["phys"]).abs().max()) > 0      # fresh masks per call
    assert torch.equal(c["phys"], d["phys"])                    # eval is deterministic
    assert all(torch.isfinite(t).all() for t in a.values())


def test_mixed_empty_rows_in_one_batch():
    """Non-homogeneous batch (some samples lack video): extractor runs on the present rows only."""
    kw = dict(t_audio=24, t_video=8)
    torch.manual_seed(3)
    model = W.perturb_norms(W.disable_dropout(W.build_c3(M, **kw))).to(DEV).eval()
    data, labels = W.batch_c3(B=6, **kw)
    names = list(data[1][0])
    for i in (1, 4):
        names[i] = "video_EMPTY"
        data[1][1][i] = -1.0
    data[1][0] = tuple(names)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = O.physverb_model(data, sd, W.c3_oracle_cfg(**kw), False, False)
        with mar.precision("fp32"):
            got = model(W.to_device(data, DEV))
    for k in ref:
        H.assert_close(got[k].cpu(), ref[k], H.FP32_TOL, f"mixed EMPTY {k}")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_full_size_c3_properties(mode):
    """BASELINE config 3 at full size (B=256 is too slow for the CPU oracle): size-independent properties.
    (1) batch independence: logits of clip i do not depend on the other clips;
    (2) linearity of the loss gradient in the upstream gradient (grad of 2·loss = 2·grad);
    (3) the first 8 clips reproduce the oracle run on those 8 clips alone."""
    B = 256 if mode == "bf16" else 64
    torch.manual_seed(0)
    model = W.perturb_norms(W.disable_dropout(W.build_c3(M))).to(DEV).train()
    data, labels = W.batch_c3(B=B)
    gdata, glabels = W.to_device(data, DEV), W.to_device(labels, DEV)
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
    with mar.precision(mode):
        pred = model(gdata)
        sub = [[n[:8], t[:8]] for n, t in gdata]
        pred8 = model(sub)
        for k in pred:
            H.assert_close(pred[k][:8], pred8[k], 1e-5 if mode == "fp32" else 1e-2, f"batch independence {k}")
        model.zero_grad()
        crit(pred, glabels).backward()
        g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
        model.zero_grad()
        losses = crit(model(gdata), glabels)
        (2.0 * losses.total()).backward()
        for k, p in model.named_parameters():
            H.assert_close(p.grad, 2.0 * g1[k], 1e-4 if mode == "fp32" else 2e-2, f"linearity {k}")
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = O.physverb_model([[n[:8], t[:8]] for n, t in data], sd, W.c3_oracle_cfg(), True, True)
    for k in ref:
        H.assert_close(pred8[k].float().cpu(), ref[k], H.FP32_TOL if mode == "fp32" else H.BF16_TOL, f"full-size weights, 8 clips {k}")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("B", [8, 64])
def test_full_width_gradient_parity(B):
    """the golden cases hold <= 200 tokens.  Here the full-width C3 model (d = 768, 8 heads, d_ff = 2048,
    T_a = 250, T_v = 64) on 8 clips (2 512 fused tokens) and 64 clips (20 096): every parameter gradient against the
    fp32 CPU oracle.
    fp32 mode: 1e-4 per tensor (rows moved by ReLU boundary flips tolerated, their number scales with tokens x units)
    and 1e-3 on the whole gradient (measured 1.8e-5).
    bf16 mode: the whole-gradient relative error is compared with what PLAIN TORCH bf16 arithmetic makes of the same case
    (the oracle's math on bf16 tensors) and printed: measured on B200 8.3e-2 (torch bf16: 8.4e-2) at 8 clips and 2.5e-2 at
    64 clips.  It does NOT fall to the north star's "about 1e-2" with more tokens: with random labels at initialisation
    the true mean gradient is itself mostly cancellation between clips (58 % of its norm sits in the heads' first layers,
    fed by ONE pooled row per clip), so bf16 rounding is large against it for any implementation."""
    torch.manual_seed(0)
    model = W.perturb_norms(W.disable_dropout(W.build_c3(M))).to(DEV).train()
    data, labels = W.batch_c3(B=B, seed=4242)
    gdata, glabels = W.to_device(data, DEV), W.to_device(labels, DEV)
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})

    def oracle(dt):
        sd = {k: v.detach().cpu().clone().to(dt).requires_grad_(True) for k, v in model.state_dict().items()}
        d = [[n, t.to(dt)] for n, t in data]
        po = {k: v.float() for k, v in O.physverb_model(d, sd, W.c3_oracle_cfg(), True, True).items()}
        sum(O.multimodal_ce(po, labels, heads=["phys", "verb"]).values()).backward()
        return {k: (None if v.grad is None else v.grad.float()) for k, v in sd.items()}, po

    ref, po = oracle(torch.float32)
    floor = H.global_rel_err(oracle(torch.bfloat16)[0], ref)
    errs = {}
    for mode in ("fp32", "bf16"):
        model.zero_grad(set_to_none=True)
        with mar.precision(mode):
            pred = model(gdata)
            crit(pred, glabels).backward()
        got = {k: p.grad for k, p in model.named_parameters()}
        errs[mode] = H.global_rel_err(got, ref)
        for k in po:
            H.assert_close(pred[k].float().cpu(), po[k].detach(), H.FP32_TOL if mode == "fp32" else H.BF16_TOL, f"{mode} logits {k}")
        if mode == "fp32":
            # ReLU boundary flips scale with the number of (token, unit) pairs: a pre-activation within the ~1e-5 the two
            # fp32 evaluations differ by lands on either side of 0 and moves ONE row (unit) of the following weight
            # gradient; 42 of linear1's 2048 rows were measured on B200 (2 000 audio tokens x 2 048 units, P ~ 1e-5)
            for k, r in ref.items():
                if r is not None:
                    H.assert_grad_close(got[k].cpu(), r, H.FP32_TOL, f"full width fp32 {k}",
                                        max_flipped_rows=4 + int(3e-5 * 314 * B * r.shape[0]))
            assert errs[mode] <= 1e-3, errs[mode]
    print(f"full-width C3, {B} clips = {314 * B} fused tokens: whole-gradient relative error vs the fp32 oracle: fp32 mode "
          f"{errs['fp32']:.3e}, bf16 mode {errs['bf16']:.3e}, plain torch bf16 arithmetic {floor:.3e}")
    assert errs["bf16"] <= max(H.BF16_TOL, H.BF16_VS_TORCH * floor), (errs["bf16"], floor)


def test_full_size_c1_c2_against_oracle():
    """C1 (B=32,T=250,d=768, 2 layers) and C2 (B=64,T=64,d=512 GRU) at BASELINE sizes, forward, vs the oracle."""
    torch.manual_seed(0)
    m1 = W.perturb_norms(W.disable_dropout(W.build_c1(M))).to(DEV).eval()
    x, _ = W.batch_c1()
    sd = {k: v.detach().cpu() for k, v in m1.state_dict().items()}
    with torch.no_grad():
        ref = O.output_classifier(O.transformer_sequence_processor(x, sd, "0.", 2, 8), sd, "1.")
        for mode, tol in (("fp32", H.FP32_TOL), ("bf16", H.BF16_TOL)):
            with mar.precision(mode):
                H.assert_close(m1(x.to(DEV)).cpu(), ref, tol, f"C1 {mode}")
    m2 = W.build_c2(M, heads=("GRU_1L", "LSTM_1L", "Avg_features")).to(DEV).eval()
    x, _ = W.batch_c2()
    sd = {k: v.detach().cpu() for k, v in m2.state_dict().items()}
    with torch.no_grad():
        ref = O.video_multi_nn(x, sd, {"GRU_1L": "gru", "LSTM_1L": "lstm", "Avg_features": "avg"})
        for mode, tol in (("fp32", H.FP32_TOL), ("bf16", 3e-2)):
            with mar.precision(mode):
                got = m2(x.to(DEV))
            for k in ref:
                H.assert_close(got[k].cpu(), ref[k], tol, f"C2 {mode} {k}")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("T", [600, 1100])
def test_long_sequences_through_the_model_against_oracle(T):
    """BASELINE config 5's lengths through the module API: TransformerSequenceProcessor + OutputClassifier at T = 600 /
    1100 (the pair kernel's range, several key tiles, a ragged last tile) with d = 192 (head dim 96 as in the reference's
    768 / 8): logits and every parameter gradient against the fp32 oracle, fp32 and bf16 mode."""
    torch.manual_seed(0)
    d, heads, B = 192, 2, 3
    model = W.perturb_norms(W.disable_dropout(W.build_c1(M, d=d, layers=1, heads=heads))).to(DEV).train()
    x, y = W.batch_c1(B=B, T=T, d=d, seed=77)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = O.output_classifier(O.transformer_sequence_processor(x, sd, "0.", 1, heads), sd, "1.")
    O.cross_entropy(ref, y).backward()
    ref_g = {k: v.grad for k, v in sd.items() if v.grad is not None}
    for mode, tol, gtol in (("fp32", H.FP32_TOL, 3e-4), ("bf16", H.BF16_TOL, 4e-2)):
        with mar.precision(mode):
            model.zero_grad()
            out = model(x.to(DEV))
            ops.cross_entropy(out, y.to(DEV)).backward()
        H.assert_close(out.detach().float().cpu(), ref.detach(), tol, f"T={T} {mode} logits")
        num = sum(float((p.grad.double().cpu() - ref_g[k].double()).pow(2).sum()) for k, p in model.named_parameters())
        den = sum(float(g.double().pow(2).sum()) for g in ref_g.values())
        assert (num / den) ** 0.5 <= gtol, f"T={T} {mode}: whole-gradient relative error {(num / den) ** 0.5:.3e}"


def test_reference_state_dict_round_trip():
    """Checkpoints interchange with the reference: same keys, load_state_dict strict."""
    a = W.build_c3(M, t_audio=24, t_video=8)
    b = W.build_c3(M, t_audio=24, t_video=8)
    b.load_state_dict(copy.deepcopy(a.state_dict()), strict=True)
    a, b = a.to(DEV).eval(), b.to(DEV).eval()
    batch = W.to_device(W.batch_c3(B=4, t_audio=24, t_video=8), DEV)
    with torch.no_grad(), mar.precision("fp32"):
        pa, pb = a(batch[0]), b(batch[0])
    assert all(torch.equal(pa[k], pb[k]) for k in pa)


def test_train_step_graphs_are_keyed_by_batch_signature():
    """A graph-captured TrainStep fed a stream of batches whose layout changes (full batch, verb-only batch with the
    video modality EMPTY, a smaller last batch) must follow the eager TrainStep step for step: a batch never
    replays a graph captured for another signature (the `_EMPTY` names steer the model's control flow)."""
    from multimodalaggressionrecognition_b200 import training
    kw = dict(t_audio=24, t_video=8)
    batches = {
        "full": W.batch_c3(B=8, seed=1, **kw),
        "full2": W.batch_c3(B=8, seed=2, **kw),
        "verb_only": W.batch_c3(B=8, seed=3, empty="video", **kw),
        "last": W.batch_c3(B=5, seed=4, **kw),
    }
    # 10 steps: every signature is captured and replayed at least once.  (Longer runs are not comparable step by
    # step: Adam turns a gradient component that is numerical noise around 0 into a +-lr update, so two runs of
    # the SAME driver — eager vs eager as well — occasionally split into two discrete loss curves after ~11 steps;
    # measured on B200: bimodal 6.3e-3 jump of one head at step 11, everything before it agrees to 1e-5.)
    order = ["full", "full2", "full", "full2", "verb_only", "full", "last", "verb_only", "full2", "last"]
    curves = {}
    for graph in (False, True):
        torch.manual_seed(0)
        model = W.perturb_norms(W.disable_dropout(W.build_c3(M, **kw))).to(DEV).train()
        crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
        step = training.TrainStep(model, crit, lr=1e-3, graph=graph, precision="fp32")
        out = []
        for name in order:
            data, labels = batches[name]
            losses = step(W.to_device(data, DEV), W.to_device(labels, DEV))
            out.append({k: float(v) for k, v in losses.items()})
        curves[graph] = out
        if graph:
            assert len(step._graphs) == 3            # full/full2 share one signature; verb_only and last have their own
            step.release_graphs()
    for i, (a, b) in enumerate(zip(curves[False], curves[True])):
        assert set(a) == set(b), f"step {i} ({order[i]}): heads {set(a)} vs {set(b)}"
        for k in a:
            # two runs of the same step differ by the order of the fp32 atomic reductions (split-K wgrad, LayerNorm,
            # dQ reduce); over 10 Adam steps that grows to a few 1e-4.  Replaying a graph captured for ANOTHER
            # signature gives errors of 1e-1 (other heads active, other batch size).
            # (5e-3, the bar of the recorded reference curves: 2.1e-3 was measured once on the verb head at step 7)
            assert abs(a[k] - b[k]) <= 5e-3 * max(1.0, abs(a[k])), f"step {i} ({order[i]}) loss[{k}]: eager {a[k]} vs graph {b[k]}"


def test_epoch_accumulator_on_the_graph_captured_step():
    """training.EpochAccumulator fed from a graph-captured TrainStep (device tensors that the next replay overwrites,
    argmax on the device, ONE read at the end) against the reference's per-step host bookkeeping (`.item()` per head,
    argmax -> numpy per head, trainer.py:718-737) done on the same steps."""
    from multimodalaggressionrecognition_b200 import training
    kw = dict(t_audio=24, t_video=8)
    torch.manual_seed(0)
    model = W.build_c3(M, **kw).to(DEV).train()
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
    step = training.TrainStep(model, crit, lr=1e-3, graph=True, precision="bf16")
    dev_acc = training.EpochAccumulator(H.trainer_metrics())
    host_acc = training.EpochAccumulator(H.trainer_metrics())
    host_acc._argmax = lambda logits: logits.detach().argmax(dim=1)     # the reference's way runs on host tensors
    n = 0
    for i in range(8):
        data, labels = W.batch_c3(B=8, seed=100 + i, empty="video" if i % 3 == 2 else None, **kw)
        losses = step(W.to_device(data, DEV), W.to_device(labels, DEV))
        dev_acc.add(losses, step.last_pred, W.to_device(labels, DEV), data=data)
        # the reference's way: blocking reads every step
        host_acc.add({k: torch.tensor(v.item()) for k, v in losses.items()},
                     {k: v.detach().float().cpu() for k, v in step.last_pred.items()}, labels, data=data)
        n += 8
    got, ref = dev_acc.results(n), host_acc.results(n)
    step.release_graphs()
    assert list(got) == list(ref) and set(got) == {"phys", "verb"}
    for h in ref:
        assert abs(got[h]["loss"] - ref[h]["loss"]) <= 1e-6 * max(1.0, abs(ref[h]["loss"]))
        for m in ("accuracy", "UAR"):
            assert got[h][m] == ref[h][m]
        assert (dev_acc.last_arrays[h]["pred"] == host_acc.last_arrays[h]["pred"]).all()
        assert (dev_acc.last_arrays[h]["true"] == host_acc.last_arrays[h]["true"]).all()
