"""CPU: the oracle restatement against the golden vectors recorded from the LIVE reference
(oracle/make_golden.py).  This is the oracle's pin; it runs anywhere (no /root/reference needed)."""
import pytest
import torch

from oracle import oracle as O
from tests import helpers as H
import multimodalaggressionrecognition_b200.models as mine

CASES = ["c1_small", "c1_odd", "c2_small", "c3_small", "c3_video_empty", "c3_audio_empty", "c3_audio_padded"]
CASES_V2 = ["c3x_audio_text_ragged", "c3x_three_modalities", "c3x_three_video_empty", "c3x_avg_fusion",
            "c3x_avg_fusion_video_empty", "c3x_base_classifier", "c3x_old_multimodal_model", "c3_weighted_ce",
            "audio_text_model"]
CASES_V3 = ["c3_mixed_rows"]


@pytest.fixture(autouse=True)
def _no_dropout():
    old = O.DROPOUT_ENABLED
    O.DROPOUT_ENABLED = False
    yield
    O.DROPOUT_ENABLED = old


@pytest.mark.parametrize("name", CASES + CASES_V2 + CASES_V3)
def test_oracle_matches_reference_golden(golden, name):
    case = golden["cases"][name]
    spec = case["spec"]
    model, (data, labels) = H.build_case(spec, mine)        # same seed => same weights as the reference
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    chk = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(chk - case["weights_checksum"]) <= 1e-9 * case["weights_checksum"], "same seed must give the reference's weights"

    if "eval" in case:
        with torch.no_grad():
            pred = H.oracle_forward(spec, sd, data, False, False)
        for k, v in case["eval"].items():
            H.assert_close(pred[k], v, 2e-5, f"{name}/eval/{k}")
    else:
        with pytest.raises(RuntimeError):
            H.oracle_forward(spec, sd, data, False, False)

    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pred = H.oracle_forward(spec, sdg, data, True, True)
    losses = H.oracle_losses(spec, pred, labels)
    assert set(losses) == set(case["losses"])
    for k, v in case["train"].items():
        H.assert_close(pred[k], v, 2e-5, f"{name}/train/{k}")
    for k, v in case["losses"].items():
        assert abs(float(losses[k].detach()) - v) < 2e-5
    sum(losses.values()).backward()
    for k, n in case["grad_norms"].items():
        g = sdg[k].grad
        if n is None:
            assert g is None or float(g.abs().max()) == 0.0
        else:
            assert abs(float(g.norm()) - n) <= 2e-4 * max(n, 1e-6), f"{name}: grad norm of {k}"
    for k, v in case["grads"].items():
        H.assert_close(sdg[k].grad, v, 1e-4, f"{name}/grad/{k}")


def test_oracle_adam_curve(golden):
    case = golden["cases"]["c2_small"]
    spec = case["spec"]
    model, (data, labels) = H.build_case(spec, mine)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr = O.OracleTrainer(sd, lambda s, d, t: H.oracle_forward(spec, s, d, t, True), lambda p, t: H.oracle_losses(spec, p, t))
    for ref_step in case["adam_curve"]:
        got = tr.step(data, labels, training=True)
        for k, v in ref_step.items():
            assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v))


def test_oracle_c1_epoch_first_steps(golden_c1_epoch):
    """golden_c1_epoch.pt holds the reference's loss curve over BASELINE config 1's epoch (48 Adam steps, full size)
    and the oracle's own curve recorded beside it (1e-5 apart on the plateau, 7e-4 at the steepest step of the
    learning transition; asserted when the fixture was generated); here the oracle re-runs the first 2 steps (a
    full epoch is minutes of CPU)."""
    from multimodalaggressionrecognition_b200 import workloads as W
    g = golden_c1_epoch
    dev = [abs(a - b) for a, b in zip(g["loss_curve"], g["loss_curve_oracle"])]
    assert max(dev) < 2e-3 and max(dev[:14]) < 2e-5
    assert len(g["loss_curve"]) == g["steps"] == 48
    spec = dict(builder="build_c1", bkw={})
    torch.manual_seed(g["init_seed"])
    model = W.perturb_norms(W.disable_dropout(W.build_c1(mine)))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr = O.OracleTrainer(sd, lambda s, d, t: H.oracle_forward(spec, s, d, t, True), lambda p, t: H.oracle_losses(spec, p, t))
    for i in range(2):
        x, y = W.batch_c1_learnable(seed=g["batch_seed0"] + i)
        got = tr.step(x, y, training=True)["loss"]
        assert abs(got - g["loss_curve"][i]) <= 1e-4, f"step {i}: oracle {got} vs reference {g['loss_curve'][i]}"


def test_oracle_alternating_batches(golden_alternating):
    """The reference's normal regime — full / verb-only / phys-only batches in turn (AggrBatchSampler makes batches
    homogeneous in aggression type): torch.optim.Adam skips the parameters of an inactive head or branch and keeps a
    step count PER PARAMETER.  10 steps recorded from the live reference (`make_golden.py alternating`); an Adam with
    one global step count is 3e-2 off on this curve."""
    from multimodalaggressionrecognition_b200 import workloads as W
    g = golden_alternating
    kw = g["kw"]
    spec = dict(builder="build_c3", bkw=kw)
    torch.manual_seed(g["init_seed"])
    model = W.perturb_norms(W.disable_dropout(W.build_c3(mine, **kw)))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr = O.OracleTrainer(sd, lambda s, d, t: H.oracle_forward(spec, s, d, t, True), lambda p, t: H.oracle_losses(spec, p, t))
    for i, kind in enumerate(g["pattern"]):
        data, labels = W.batch_c3(B=g["B"], seed=g["seed0"] + i, empty=None if kind == "full" else kind, **kw)
        got = tr.step(data, labels, training=True)
        assert set(got) == set(g["loss_curve"][i])
        for k, v in g["loss_curve"][i].items():
            assert abs(got[k] - v) <= 1e-5, f"step {i} ({kind}) {k}: oracle {got[k]} vs reference {v}"
    for k, v in g["final_params"].items():
        H.assert_close(tr.sd[k].detach(), v, 1e-4, f"final {k}")
    # an inactive head stood still: its per-parameter step count is the number of batches it saw
    assert tr.steps["classifiers.classifiers_dict.phys.3.bias"] == sum(k != "video" for k in g["pattern"])
    assert tr.steps["classifiers.classifiers_dict.verb.3.bias"] == sum(k != "audio" for k in g["pattern"])
    assert tr.steps["modality_fusion_module.modality_fusion_transformer.norm.bias"] == len(g["pattern"])


def test_adam_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref])
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for t in range(1, 6):
        g = torch.randn(1000)
        p_ref.grad = g.clone()
        opt.step()
        O.adam_step([p], [g], [m], [v], t)
        assert torch.allclose(p, p_ref.detach(), atol=1e-6)


def test_safe_softmax_all_masked_row():
    s = torch.full((2, 4), float("-inf"))
    s[1, 2] = 0.5
    p = O.softmax_lastdim_safe(s)
    assert torch.equal(p[0], torch.zeros(4)) and float(p[1, 2]) == 1.0


def test_oracle_focal_loss_restates_the_hub_module():
    """The focal loss the reference fetches from torch.hub is absent offline: check the restatement against the
    module's published forward written with torch's own log_softmax / nll_loss (parity unpinned, see oracle.py)."""
    import torch.nn.functional as F
    from oracle import oracle as O
    torch.manual_seed(0)
    x = torch.randn(50, 4) * 3
    y = torch.randint(0, 4, (50,))
    y[::6] = -100
    alpha = torch.tensor([0.3, 1.0, 2.0, 0.7])
    for gamma in (0.0, 1.5, 2.0):
        keep = y != -100
        log_p = F.log_softmax(x[keep], dim=-1)
        ce = F.nll_loss(log_p, y[keep], weight=alpha, reduction="none")
        pt = log_p[torch.arange(int(keep.sum())), y[keep]].exp()
        ref = ((1 - pt) ** gamma * ce).mean()
        got = O.focal_loss(x, y, alpha, gamma)
        assert abs(float(got) - float(ref)) < 1e-6
    assert float(O.focal_loss(x, torch.full((50,), -100), alpha, 2.0)) == 0.0
