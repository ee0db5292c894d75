"""CPU (gloo, world_size 2): the host-side data-parallel logic of training.py — flat parameter/gradient
buffers, bucketed gradient all-reduce fired from grad hooks, the fused-Adam arithmetic — checked against a
single-process run on the concatenated batch.  The modules under test here are plain torch layers: the
CUDA drop-in modules refuse CPU tensors by design, and this file tests the plumbing around them."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from multimodalaggressionrecognition_b200 import training
from multimodalaggressionrecognition_b200.models import LossesDict


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class TwoHead(nn.Module):
    """Shared trunk + two heads, like the fusion model's {'phys','verb'} outputs."""

    def __init__(self):
        super().__init__()
        self.trunk = nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 32), nn.ReLU())
        self.a = nn.Linear(32, 2)
        self.b = nn.Linear(32, 2)
        self.unused = nn.Linear(4, 4)      # never reached: gets no gradient (an inactive head)

    def forward(self, x):
        h = self.trunk(x)
        return {"a": self.a(h), "b": self.b(h)}


class _SinkLinear(torch.autograd.Function):
    """CPU stand-in for the gradient-sink protocol of ops._Linear inside a TrainStep: the weight / bias gradients are
    accumulated STRAIGHT into the parameters' flat .grad views, the step driver is told through ops._sunk(), and
    autograd gets None for both parameters."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.refs = (w, b)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, g):
        from multimodalaggressionrecognition_b200 import ops
        x, w = ctx.saved_tensors
        wref, bref = ctx.refs
        if not ops._sink_cfg["on"]:                       # outside a TrainStep: ordinary autograd
            return g @ w, g.t() @ x, g.sum(0), 
        wref.grad.add_(g.t() @ x)
        ops._sunk(wref)
        bref.grad.add_(g.sum(0))
        ops._sunk(bref)
        return g @ w, None, None


class SinkTwoHead(TwoHead):
    """TwoHead whose trunk layers sink their gradients (like the 38 Linear parameters of the fusion model) while the
    heads go through autograd's AccumulateGrad (like its LayerNorm vectors): both ways of counting a parameter's
    gradient as complete meet in the same buckets."""

    def __init__(self):
        super().__init__()
        self.extra = nn.ModuleList([nn.Linear(32, 32) for _ in range(3)])

    def forward(self, x):
        h = torch.relu(_SinkLinear.apply(x, self.trunk[0].weight, self.trunk[0].bias))
        h = torch.relu(_SinkLinear.apply(h, self.trunk[2].weight, self.trunk[2].bias))
        for lin in self.extra:       # several sinking layers in ONE bucket: double counting completes it before the trunk's turn
            h = torch.relu(_SinkLinear.apply(h, lin.weight, lin.bias))
        return {"a": self.a(h), "b": self.b(h)}


def _criterion(pred, labels):
    out = LossesDict()
    ce = nn.CrossEntropyLoss()
    for k, v in pred.items():
        out[k] = ce(v, labels)
    return out


def _make(seed=0, sink=False):
    torch.manual_seed(seed)
    return SinkTwoHead() if sink else TwoHead()


def _worker(rank, world, port, steps, result_q, sink=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _make(sink=sink)
        step = training.TrainStep(model, _criterion, lr=1e-2, num_buckets=1 if sink else 3)
        assert step.sync.world == world and len(step.sync.buckets) >= (1 if sink else 2)
        comm = None
        if sink == "streams":
            # Run GradSync's CUDA branch (side communication stream) without a GPU: fake stream objects that record
            # who waited for whom, gradients "written" alternately on two compute streams, and gloo's missing AVG
            # expressed as SUM / world.
            import contextlib

            class FakeStream:
                def __init__(self, handle):
                    self.cuda_stream, self.waited = handle, []

                def wait_stream(self, other):
                    self.waited.append(other.cuda_stream)

            comm, compute, turn = FakeStream(99), [FakeStream(1), FakeStream(2)], [0]

            def current_stream(*_a):
                turn[0] += 1
                return compute[turn[0] & 1]
            torch.cuda.current_stream = current_stream
            torch.cuda.stream = lambda _s: contextlib.nullcontext()
            real_all_reduce = dist.all_reduce

            def all_reduce(t, op=None, group=None):
                real_all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                t.div_(world)
            training.dist.all_reduce = all_reduce
            step.sync.comm_stream = comm
        g = torch.Generator().manual_seed(123)
        X = torch.randn(steps, 8 * world, 16, generator=g)
        Y = torch.randint(0, 2, (steps, 8 * world), generator=g)
        for s in range(steps):
            xs, ys = X[s, rank * 8:(rank + 1) * 8], Y[s, rank * 8:(rank + 1) * 8]
            losses = step(xs, ys)
            assert set(losses) == {"a", "b"}
        flat = step.flat.flat.detach().clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        grads = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(grads, step.flat.grad.detach().clone())     # the last step's gradients, after the exchange
        if comm is not None:
            # every launch waited for BOTH streams gradients of the bucket were written on, and finish() joined
            assert {1, 2} <= set(comm.waited), comm.waited
            assert any(99 in c.waited for c in compute)
        if rank == 0:
            result_q.put(([g.numpy() for g in gathered], X.numpy(), Y.numpy(), [g.numpy() for g in grads]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("sink", [False, True, "streams"])
def test_two_rank_training_matches_single_process(sink):
    """sink=True: parameters whose gradient is written by a kernel straight into the flat buffer (ops.grad_sink) are
    counted through GradSync.notify AND torch fires their post-accumulate hook as well; counting both launched a
    bucket's all-reduce before its gradients were complete and the ranks drifted apart (found on 2 B200s,
    tools/dp_diag.py) — every rank must end each step with the same gradients and parameters.
    sink="streams": the same through GradSync's side-stream branch, with recording fake streams."""
    world, steps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, q, sink)) for r in range(world)]
    for p in procs:
        p.start()
    flats, X, Y, grads = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f0, f1 = torch.from_numpy(flats[0]), torch.from_numpy(flats[1])
    assert torch.equal(torch.from_numpy(grads[0]), torch.from_numpy(grads[1])), "ranks hold different gradients after the exchange"
    assert torch.equal(f0, f1), "ranks diverged"

    # single process, global batch, torch.optim.Adam as the reference optimizer
    model = _make(sink=sink)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    X, Y = torch.from_numpy(X), torch.from_numpy(Y)
    for s in range(steps):
        opt.zero_grad()
        losses = _criterion(model(X[s]), Y[s])
        losses.backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters() if p.requires_grad and p.grad is not None])
    # compare parameter by parameter through the flat offsets (the unused head must be untouched)
    single = training.FlatParams(list(_make(sink=sink).parameters()))
    off = dict(zip([id(p) for p in single.params], single.offsets))
    m2 = _make(sink=sink)
    for (name, p_ref), p_init in zip(model.named_parameters(), m2.parameters()):
        idx = [i for i, q_ in enumerate(m2.parameters()) if q_ is p_init][0]
        o = single.offsets[idx]
        got = f0[o:o + p_ref.numel()].view(p_ref.shape)
        assert torch.allclose(got, p_ref.detach(), atol=2e-6, rtol=1e-5), f"{name} differs from the single-process run"
    assert ref.numel() > 0


def test_flat_params_keep_module_semantics():
    model = _make()
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    flat = training.FlatParams(list(model.parameters()))
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd_before[k])                         # values preserved, keys unchanged
    p = next(model.parameters())
    assert p.data_ptr() == flat.flat.data_ptr()                      # parameters are views of the flat buffer
    assert p.grad is not None and p.grad.data_ptr() == flat.grad.data_ptr()
    model(torch.randn(4, 16))["a"].sum().backward()
    assert float(flat.grad.abs().sum()) > 0                          # autograd accumulated into the flat gradient
    flat.zero_grad()
    assert float(flat.grad.abs().sum()) == 0


def test_flat_adam_matches_torch_adam_on_cpu():
    m1, m2 = _make(1), _make(1)
    flat = training.FlatParams(list(m1.parameters()))
    opt1 = training.FlatAdam(flat, lr=1e-3)
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    x, y = torch.randn(8, 16), torch.randint(0, 2, (8,))
    for _ in range(4):
        opt1.zero_grad(); opt2.zero_grad()
        _criterion(m1(x), y).backward()
        _criterion(m2(x), y).backward()
        opt1.step(); opt2.step()
    for (n, a), b in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-6), n


def test_bucket_layout_covers_every_parameter_once():
    model = _make()
    flat = training.FlatParams(list(model.parameters()))
    sync = training.GradSync(flat, num_buckets=4)
    covered = []
    for lo, hi, e0, e1 in sync.buckets:
        covered += list(range(lo, hi))
        assert e0 == flat.offsets[lo] and e1 > e0
    assert sorted(covered) == list(range(len(flat.params)))
    assert sync.buckets[0][1] == len(flat.params)                    # first bucket = the LAST parameters (backward order)
