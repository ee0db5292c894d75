"""CPU: host-side logic of the step driver that does not need a GPU — batch signatures that key the captured
graphs, the gradient-sink switch, and the FocalLoss module's argument handling."""
import pytest
import torch
import torch.nn as nn

from multimodalaggressionrecognition_b200 import models as M, ops, training, workloads as W


def test_batch_signature_separates_shapes_and_names():
    sig = training.TrainStep._signature
    a = W.batch_c3(B=4, t_audio=6, t_video=3, seed=1)
    b = W.batch_c3(B=4, t_audio=6, t_video=3, seed=2)          # other values, same layout
    c = W.batch_c3(B=4, t_audio=6, t_video=3, seed=1, empty="video")
    d = W.batch_c3(B=3, t_audio=6, t_video=3, seed=1)          # last batch of an epoch
    e = W.batch_c3(B=4, t_audio=7, t_video=3, seed=1)
    assert sig(list(a)) == sig(list(b))
    assert len({sig(list(x)) for x in (a, c, d, e)}) == 4
    # the names that steer control flow are part of the key even when every tensor shape agrees
    assert [t.shape for t in training.TrainStep._tensors(list(a))] == [t.shape for t in training.TrainStep._tensors(list(c))]
    hash(sig(list(a)))                                          # usable as a dict key


def test_grad_sink_is_off_by_default_and_never_targets_cpu_parameters():
    p = nn.Parameter(torch.zeros(4, 4))
    p.grad = torch.zeros(4, 4)
    assert ops._sink_target(p) is None                          # switched off outside a TrainStep
    seen = []
    with ops.grad_sink(seen.append):
        assert ops._sink_cfg["on"]
        assert ops._sink_target(p) is None                      # CPU gradient: the kernels cannot write there
        assert ops._sink_target(torch.zeros(3)) is None         # not a Parameter (a slice / a cast of one)
        ops._sunk(p)
    assert seen == [p] and not ops._sink_cfg["on"]


def test_grad_sync_notify_counts_like_the_autograd_hook():
    lin = nn.Linear(3, 2)
    flat = training.FlatParams(list(lin.parameters()))
    sync = training.GradSync(flat, num_buckets=1)
    assert sync.world == 1
    sync.notify(lin.weight)                                     # world 1: nothing to exchange, must not raise
    assert set(sync.index_of) == {id(p) for p in flat.params}
    # ... but the parameter is recorded as active for the per-parameter Adam: finish() writes the header flags
    sync.finish()
    assert flat.flags.tolist() == [1.0, 0.0] and sync.last_active == [True, False]
    lin(torch.randn(2, 3)).sum().backward()                     # both parameters through autograd's hooks
    sync.finish()
    assert flat.flags.tolist() == [1.0, 1.0]
    sync.finish()                                               # a step in which nothing received a gradient
    assert flat.flags.tolist() == [0.0, 0.0]


def test_no_cpu_arithmetic_in_the_step_driver():
    """FlatAdam refuses CPU buffers (README: no CPU fallback); the gloo tests install their own stand-in."""
    flat = training.FlatParams(list(nn.Linear(3, 2).parameters()))
    with pytest.raises(RuntimeError, match="sm_100a"):
        training.FlatAdam(flat).step()


def test_focal_loss_module_arguments():
    with pytest.raises(ValueError):
        M.FocalLoss(reduction="none")
    f = M.FocalLoss(alpha=[0.25, 0.75], gamma=2.0, ignore_index=-1)
    assert f.alpha.dtype == torch.float32 and f.gamma == 2.0 and f.ignore_index == -1
    with pytest.raises(RuntimeError):                           # no CPU path: CPU logits fail loudly
        f(torch.zeros(2, 2), torch.zeros(2, dtype=torch.long))


@pytest.mark.parametrize("name", ["single", "multi_head", "multimodal"])
def test_epoch_accumulator_reproduces_the_reference_trainers(name, monkeypatch):
    """training.EpochAccumulator (one device->host read per epoch) against what the LIVE reference's
    TorchSupervisedTrainer / RNN_trainer / MultimodalTrainer methods returned for the same seeded step stream
    (tests/golden/golden_trainer_epoch.pt, oracle/make_golden_trainer.py) — including the reference's
    `size = len(data[0])` rule, all-EMPTY groups dropped per step and row-wise EMPTY filtering."""
    import os
    import numpy as np
    from multimodalaggressionrecognition_b200 import training
    from tests import helpers as H
    golden = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_trainer_epoch.pt"), weights_only=False)
    case = golden["cases"][name]
    spec = case["spec"]
    assert spec == H.TRAINER_STREAMS[name]
    # the argmax kernel cannot run here: torch's argmax stands in for it (the package itself has no CPU arithmetic)
    monkeypatch.setattr(training.EpochAccumulator, "_argmax", staticmethod(lambda logits: logits.detach().argmax(dim=1)))
    acc = training.EpochAccumulator(H.trainer_metrics())
    for data, losses, pred, labels in H.trainer_stream(**spec):
        acc.add(losses, pred, labels, data=data)
    got = acc.results(spec["dataset_size"])
    ref = case["results"]

    def same(a, b, where):
        assert set(a) == set(b), where
        for k in a:
            assert np.allclose(np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64), rtol=1e-6, atol=1e-9), f"{where}/{k}: {a[k]} vs {b[k]}"

    if name == "single":
        same(got, ref, name)
    else:
        assert list(got) == list(ref)                       # same heads, in the reference's order
        for h in ref:
            same(got[h], ref[h], f"{name}/{h}")
    acc.reset()
    assert acc.results(1) == {}


def test_weight_cache_invalidation(monkeypatch):
    """ops.compute_weight's cache of bf16 weight copies (host logic; the cast kernel is stubbed out): one cast per
    parameter version, a new cast after an update torch cannot see (the fused Adam kernel / a graph replay call
    ops.weights_changed), after a torch in-place update, and for ANOTHER tensor object on the same address."""
    casts = []
    monkeypatch.setattr(ops, "call", lambda name, *a: casts.append(name))
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    ops.clear_weight_cache()
    w = nn.Parameter(torch.randn(4, 3))
    a = ops.compute_weight(w, torch.bfloat16, True)
    b = ops.compute_weight(w, torch.bfloat16, True)
    assert a[0] is b[0] and a[1] is b[1] and casts == ["mar_cast_weight"]
    ops.weights_changed()
    c = ops.compute_weight(w, torch.bfloat16, True)
    assert c[0] is not a[0] and len(casts) == 2
    with torch.no_grad():
        w.add_(1.0)                                             # torch.optim.Adam's kind of update
    ops.compute_weight(w, torch.bfloat16, True)
    assert len(casts) == 3
    w2 = nn.Parameter(w.data)                                   # a fresh parameter on the same address and shape
    assert w2.data_ptr() == w.data_ptr()
    ops.compute_weight(w2, torch.bfloat16, True)
    assert len(casts) == 4
    assert ops.compute_weight(w, torch.float32, True)[0] is w and len(casts) == 4      # fp32 mode: the parameter itself
    ops.clear_weight_cache()


def test_flat_adam_cuda_branch_marks_weights_changed():
    """training.FlatAdam / TrainStep replay call ops.weights_changed (source-level check: the CUDA branches cannot run here)."""
    import inspect
    assert "ops.weights_changed()" in inspect.getsource(training.FlatAdam.step)
    assert "ops.weights_changed()" in inspect.getsource(training.TrainStep._graphed)


def test_backward_handoff_protocol():
    """ops.new_handoff / _handoff_bias_target: the dict a linear shares with the neighbour that does part of its backward
    (DESIGN.md §4.3).  One neighbour at most takes the bias gradient; a bias without a gradient, a foreign size, a missing
    dict or grad mode off hand nothing over; ops.handoffs(False) switches the mechanism off."""
    b = nn.Parameter(torch.zeros(6))
    h = ops.new_handoff()
    assert h == {}
    h["bias"] = b
    tgt = ops._handoff_bias_target(h, 6, torch.device("cpu"))
    assert tgt is not None and tgt.shape == (6,) and tgt.dtype == torch.float32 and float(tgt.abs().sum()) == 0.0
    assert h["bias_done"] and h["dbias"] is tgt and h["sunk"] is False
    assert ops._handoff_bias_target(h, 6, torch.device("cpu")) is None            # already taken by a neighbour
    assert ops._handoff_bias_target(None, 6, torch.device("cpu")) is None
    assert ops._handoff_bias_target({"bias": None}, 6, torch.device("cpu")) is None
    assert ops._handoff_bias_target({"bias": nn.Parameter(torch.zeros(6), requires_grad=False)}, 6, torch.device("cpu")) is None
    assert ops._handoff_bias_target({"bias": b}, 7, torch.device("cpu")) is None   # not this linear's width
    # inside a gradient sink the target is the parameter's flat-buffer .grad view — CUDA buffers only: a CPU .grad is not sunk
    b.grad = torch.zeros(6)
    with ops.grad_sink():
        h2 = {"bias": b}
        assert ops._handoff_bias_target(h2, 6, torch.device("cpu")) is not b.grad and h2["sunk"] is False
    with ops.handoffs(False):
        assert ops.new_handoff() is None
    with torch.no_grad():
        assert ops.new_handoff() is None
    assert ops.new_handoff() == {}


def test_fused_exchange_is_only_chosen_with_peer_memory():
    """GradSync.fused(): the fused peer-memory exchange + Adam kernel needs CUDA buffers, the bf16 wire format, ONE bucket
    and an NCCL group; a CPU / single-process GradSync never has a PeerExchange and TrainStep keeps calling FlatAdam.step
    (source-level check of the branch: the CUDA path cannot run here)."""
    import inspect
    lin = nn.Linear(8, 8)
    flat = training.FlatParams(list(lin.parameters()))
    sync = training.GradSync(flat)
    assert sync.peer is None and not sync.fused() and sync.wire == "fp32"
    src = inspect.getsource(training.TrainStep._eager)
    assert "if not fused:" in src and "self.opt.step()" in src
    assert training.bind_to_gpu_numa_node(0) is None or isinstance(training.bind_to_gpu_numa_node(0), list)
