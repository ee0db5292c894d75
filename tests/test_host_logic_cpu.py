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


def test_focal_loss_module_arguments():
    with pytest.raises(ValueError):
        M.FocalLoss(reduction="none")
    f = M.FocalLoss(alpha=[0.25, 0.75], gamma=2.0, ignore_index=-1)
    assert f.alpha.dtype == torch.float32 and f.gamma == 2.0 and f.ignore_index == -1
    with pytest.raises(RuntimeError):                           # no CPU path: CPU logits fail loudly
        f(torch.zeros(2, 2), torch.zeros(2, dtype=torch.long))
