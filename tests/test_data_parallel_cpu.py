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
from tests import helpers as H


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
    H.install_host_stand_ins()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _make(sink=sink)
        step = training.TrainStep(model, _criterion, lr=1e-2, num_buckets=1 if sink else 3)
        assert step.sync.world == world and len(step.sync.buckets) >= (1 if sink else 2)
        comm = None
        if sink == "streams":
            # Run GradSync's CUDA branch (side communication stream) without a GPU: fake stream objects that record
            # who waited for whom, gradients "written" alternately on two compute streams.
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
    assert p.data_ptr() == flat.flat.data_ptr() + 4 * flat.header    # parameters are views of the flat buffer (after the flag header)
    assert p.grad is not None and p.grad.data_ptr() == flat.grad.data_ptr() + 4 * flat.header
    assert flat.header % flat.align == 0 and flat.header >= flat.nseg and flat.chunk_seg.numel() * flat.align == flat.numel
    assert flat.chunk_seg[0] == -1 and flat.chunk_seg[flat.header // flat.align] == 0 and flat.chunk_seg[-1] == flat.nseg - 1
    model(torch.randn(4, 16))["a"].sum().backward()
    assert float(flat.grad.abs().sum()) > 0                          # autograd accumulated into the flat gradient
    flat.zero_grad()
    assert float(flat.grad.abs().sum()) == 0


def test_train_step_follows_torch_adam_when_heads_alternate(monkeypatch):
    """The reference's normal regime (datasets.py:630-645): batches alternate between heads, the inactive head's
    parameters have no gradient and torch.optim.Adam skips them (no moment decay, no step count).  TrainStep's
    tracking of the active parameters + the per-parameter Adam (host stand-in of the kernel here) must follow
    torch.optim.Adam exactly; a flat Adam with one global step count does not (checked below)."""
    monkeypatch.setattr(training.FlatAdam, "step", H.host_adam_step)
    m1, m2 = _make(1), _make(1)
    pattern = ["ab", "a", "b", "a", "ab", "b", "b", "a"]

    def crit(pred, labels):
        out = LossesDict()
        for k in labels[0]:
            out[k] = nn.CrossEntropyLoss()(pred[k], labels[1])
        return out

    step = training.TrainStep(m1, crit, lr=1e-2)
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(5)
    for heads in pattern:
        x, y = torch.randn(8, 16, generator=g), torch.randint(0, 2, (8,), generator=g)
        step(x, (heads, y))
        active = {n for n, f in zip([n for n, _ in m1.named_parameters()], step.sync.last_active) if f}
        assert ("a.weight" in active) == ("a" in heads) and ("b.weight" in active) == ("b" in heads)
        assert "unused.weight" not in active and "trunk.0.weight" in active
        opt2.zero_grad()
        crit(m2(x), (heads, y)).backward()
        opt2.step()
    for (n, a), b in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-6), n
    steps = dict(zip([n for n, _ in m1.named_parameters()], step.opt.seg_steps.tolist()))
    assert steps["a.weight"] == sum("a" in h for h in pattern) and steps["b.bias"] == sum("b" in h for h in pattern)
    assert steps["trunk.0.weight"] == len(pattern) and steps["unused.weight"] == 0


def _uneven_worker(rank, world, port, steps, result_q):
    """Ranks whose slices differ in which heads are present and in how many rows each head keeps."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    H.install_host_stand_ins()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        step = training.TrainStep(_make(), MaskedCE(), lr=1e-2, num_buckets=3)
        assert len(step.sync.buckets) >= 2
        X, Y = _uneven_data(steps, world)
        orders = []
        for s in range(steps):
            sl = slice(rank * 8, (rank + 1) * 8)
            step(X[s, sl], {k: v[s, sl] for k, v in Y.items()})
            orders.append(list(range(len(step.sync.buckets))))
        flat = step.flat.flat.detach().clone()
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        if rank == 0:
            result_q.put([g.numpy() for g in gathered])
    finally:
        dist.destroy_process_group()


class MaskedCE:
    """Two-head criterion over per-head labels with -1 = row absent (the EMPTY rows of MultiModalCrossEntropyLoss):
    a head whose rows are all absent emits no loss.  `label_weight_sums` is the protocol TrainStep uses to weigh the
    ranks' gradients (models.MultiModalCrossEntropyLoss implements it with a kernel)."""
    heads = ["a", "b"]

    def __call__(self, pred, labels):
        out = LossesDict()
        for k in self.heads:
            keep = labels[k] >= 0
            if bool(keep.any()):
                out[k] = nn.CrossEntropyLoss()(pred[k][keep], labels[k][keep])
        return out

    def label_weight_sums(self, labels, device, out):
        if out is not None:
            for i, k in enumerate(self.heads):
                out[i] = float((labels[k] >= 0).sum())
        return list(self.heads)


def _uneven_data(steps, world):
    g = torch.Generator().manual_seed(321)
    X = torch.randn(steps, 8 * world, 16, generator=g)
    Y = {k: torch.randint(0, 2, (steps, 8 * world), generator=g) for k in ("a", "b")}
    Y["a"][0, 8:] = -1          # step 0: head a lives on rank 0 only
    Y["b"][0, :8] = -1          #         head b on rank 1 only   (different active sets per rank)
    Y["a"][1, :5] = -1          # step 1: both heads everywhere, different row counts per rank
    Y["b"][1, 10:] = -1
    Y["b"][2] = -1              # step 2: head b absent on every rank (skipped by Adam, as in a single process)
    Y["a"][3, :8] = -1          # step 3: rank 0 has nothing for head a, 3 rows for b
    Y["b"][3, 3:8] = -1
    return X, Y


@pytest.mark.timeout(180)
def test_ranks_with_different_active_heads_and_row_counts_match_the_global_batch():
    """(i) the collective order does not depend on which parameters a rank's slice activates (bucket b goes only
    after buckets 0..b-1; the old first-complete-first-launched order paired different buckets across ranks — gloo
    aborted with a size mismatch, NCCL would hang); (ii) every rank weighs its loss by local rows / global rows, so
    the SUMMED gradient is the gradient of the mean over the GLOBAL batch; (iii) a parameter is active if any rank
    saw a gradient for it.  Reference: one process, torch.optim.Adam, the whole batch."""
    world, steps = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_uneven_worker, args=(r, world, port, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    flats = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f0, f1 = torch.from_numpy(flats[0]), torch.from_numpy(flats[1])
    assert torch.equal(f0, f1), "ranks diverged"
    model = _make()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    X, Y = _uneven_data(steps, world)
    crit = MaskedCE()
    for s in range(steps):
        opt.zero_grad()
        crit(model(X[s]), {k: v[s] for k, v in Y.items()}).backward()
        opt.step()
    layout = training.FlatParams(list(_make().parameters()))
    for (name, p_ref), o in zip(model.named_parameters(), layout.offsets):
        got = f0[o:o + p_ref.numel()].view(p_ref.shape)
        assert torch.allclose(got, p_ref.detach(), atol=3e-6, rtol=1e-5), f"{name} differs from the single-process run"


def test_bucket_layout_covers_every_parameter_once():
    model = _make()
    flat = training.FlatParams(list(model.parameters()))
    sync = training.GradSync(flat, num_buckets=4)
    covered = []
    for lo, hi, e0, e1 in sync.buckets:
        covered += list(range(lo, hi))
        assert e0 == (flat.offsets[lo] if lo > 0 else 0) and e1 > e0      # parameter 0's bucket carries the flag header
    assert sorted(covered) == list(range(len(flat.params)))
    assert sync.buckets[0][1] == len(flat.params)                    # first bucket = the LAST parameters (backward order)


def test_tail_bucket_is_cut_small():
    """The last bucket to be reduced (the first parameters: the end of backward, nothing left to hide its all-reduce
    behind) is cut just above `tail_elems`; the others split the rest from the end; every parameter exactly once."""
    torch.manual_seed(0)
    model = nn.Sequential(*[nn.Linear(64, 64) for _ in range(6)])
    flat = training.FlatParams(list(model.parameters()))
    sync = training.GradSync(flat, num_buckets=4, tail_elems=3000)
    lo, hi, e0, e1 = sync.buckets[-1]
    assert (lo, hi) == (0, 1) and e0 == 0                      # one 4096-element weight >= 3000, plus the flag header
    covered = [i for lo_, hi_, _, _ in sync.buckets for i in range(lo_, hi_)]
    assert sorted(covered) == list(range(len(flat.params))) and len(sync.buckets) == 4
    assert sync.wire == "fp32"                                    # the bf16 wire is a CUDA / multi-rank option only
    with pytest.raises(ValueError):
        training.GradSync(flat, wire="fp16")


class _AggrLikeSampler:
    """The reference's AggrBatchSampler protocol (datasets.py:620-655): homogeneous batches per group, reshuffled with
    an UNSEEDED random generator after every pass."""

    def __init__(self, groups, batch_size):
        self.groups, self.batch_size = groups, batch_size
        self.batches = self._make()

    def _make(self):
        import random
        out = []
        for idx in self.groups.values():
            idx = list(idx)
            random.seed(None)
            random.shuffle(idx)
            out += [idx[i:i + self.batch_size] for i in range(0, len(idx), self.batch_size)]
        random.seed(None)
        random.shuffle(out)
        return out

    def __iter__(self):
        yield from self.batches
        self.batches = self._make()

    def __len__(self):
        return len(self.batches)


def _sampler_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        groups = {"phys": range(0, 40), "verb": range(40, 103)}
        sampler = training.ShardedBatchSampler(_AggrLikeSampler(groups, 16))
        epochs = [[list(b) for b in sampler] for _ in range(2)]
        q.put((rank, epochs, len(sampler)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_batch_sampler_gives_every_rank_a_slice_of_the_same_homogeneous_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sampler_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (e, n)) for r, e, n in (q.get(timeout=100) for _ in range(world)))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (e0, n0), (e1, n1) = got[0], got[1]
    assert n0 == n1 == len(e0[0]) == len(e1[0])
    seen = []
    for ep in range(2):
        for b0, b1 in zip(e0[ep], e1[ep]):
            both = b0 + b1
            assert len(set(both)) == len(both) and 0 < len(both) <= 16 and abs(len(b0) - len(b1)) <= 1
            assert all(i < 40 for i in both) or all(i >= 40 for i in both)       # one aggression type per global batch
        seen.append(sorted(i for b in e0[ep] + e1[ep] for i in b))
        assert seen[-1] == list(range(103))                                       # every sample exactly once per epoch
    assert e0[0] != e0[1]                                                         # reshuffled between epochs


def test_shard_batch_slices_tensors_and_names():
    from multimodalaggressionrecognition_b200 import workloads as W
    data, labels = W.batch_c3_mixed(B=6, t_audio=4, t_video=2)
    parts = [training.shard_batch([data, labels], r, 3) for r in range(3)]
    for r, (d, l) in enumerate(parts):
        assert d[0][0] == data[0][0][2 * r:2 * r + 2] and torch.equal(d[1][1], data[1][1][2 * r:2 * r + 2])
        assert l[1][0] == labels[1][0][2 * r:2 * r + 2] and torch.equal(l[0][1], labels[0][1][2 * r:2 * r + 2])
