"""Step driver for the hot path: flat parameter/gradient buffers, one fused Adam launch, bucketed gradient
all-reduce overlapped with backward (data parallel, one process per GPU), and optional whole-step CUDA-graph
capture.  This is the sync-free equivalent of `TorchSupervisedTrainer.train_step` (trainer.py:110-163):
zero_grad → model(data) → criterion → loss.backward() → optimizer.step(), without the per-step `.item()` /
`.cpu()` host syncs (trainer.py:731-737, :718-729) — losses and predictions stay on the device until asked for.

The reference has no distributed code; data parallelism is this repo's addition (§8e):
every op on the path is per-clip, so the batch shards across ranks and the only exchange is the parameter
gradient all-reduce (NCCL over NVLink; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.utils.data

from . import ops
from ._lib import call
from .models import LossesDict


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


import os as _os
_NOCOMM = _os.environ.get("MAR_DP_NOCOMM") == "1"


class FlatParams:
    """Moves a model's parameters into ONE contiguous fp32 buffer (params become views, state_dict keys and
    shapes are unchanged) and gives every parameter a persistent .grad view into ONE flat gradient buffer.

    Layout of both buffers: [header | p0 | p1 | ...], every block starting on a multiple of `align` floats.  One
    SEGMENT per parameter; `chunk_seg[i // align]` names the segment of float i (-1: header).  The header of the
    GRADIENT buffer carries the step's per-parameter "received a gradient" flags (`flags`, one float per segment):
    they travel with the last gradient bucket of the data-parallel all-reduce and steer the per-parameter Adam."""

    def __init__(self, params: Sequence[nn.Parameter], align: int = 64):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        assert align >= 4 and align & (align - 1) == 0
        dev = self.params[0].device
        self.align = align
        self.nseg = len(self.params)
        self.header = (self.nseg + align - 1) // align * align
        self.offsets, off = [], self.header
        seg_of_chunk = [-1] * (self.header // align)
        for i, p in enumerate(self.params):
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets.append(off)
            n = (p.numel() + align - 1) // align * align
            seg_of_chunk += [i] * (n // align)
            off += n
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.chunk_seg = torch.tensor(seg_of_chunk, dtype=torch.int32).to(dev)
        self.flags = self.grad[:self.nseg]
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + p.numel()].view(p.shape)
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        # bf16 compute copy of the whole buffer: the Adam kernel rewrites it wherever it updates a parameter, so the
        # forward's weights need no per-step cast kernels (ops.compute_weight finds the 2-D parameters' views here)
        self.mirror = None
        if dev.type == "cuda":
            self.mirror = torch.empty(off, dtype=torch.bfloat16, device=dev)
            call("mar_cast", self.flat.data_ptr(), 0, self.mirror.data_ptr(), 1, off, _stream())
            for p, o in zip(self.params, self.offsets):
                if p.dim() == 2:
                    ops.register_weight_mirror(p, self.mirror[o:o + p.numel()].view(p.shape))

    def zero_grad(self) -> None:
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):       # re-attach if something set .grad = None
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)


class FlatAdam:
    """torch.optim.Adam (train_multimodal.py:444: lr 1e-3, betas (0.9, 0.999), eps 1e-8, no weight decay) over the
    flat buffers with torch's per-parameter semantics: a parameter that received NO gradient this step is skipped —
    no moment decay, no update, no step-count increment — and every parameter keeps its own step count.  That is the
    reference's normal regime: its sampler makes every batch homogeneous in aggression type (datasets.py:630-645) and
    zero_grad leaves the inactive head's gradients None (trainer.py:140-150).  Which parameters are active is read on
    the DEVICE from the gradient buffer's header flags (FlatParams.flags), the step counts live on the device too:
    two launches per step (`mar_adam_step_segments`), CUDA-graph capturable."""

    def __init__(self, flat: FlatParams, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.flat, self.lr, self.betas, self.eps = flat, lr, betas, eps
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.seg_steps = torch.zeros(flat.nseg, dtype=torch.float32, device=flat.flat.device)
        self.seg_coef = torch.zeros(2 * flat.nseg, dtype=torch.float32, device=flat.flat.device)

    def step(self) -> None:
        f = self.flat
        if not f.flat.is_cuda:
            raise RuntimeError("FlatAdam runs on sm_100a devices only: there is no CPU arithmetic in this package "
                               "(the host-logic tests install their own stand-in, tests/helpers.py)")
        call("mar_adam_step_segments", f.flat.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
             self.exp_avg_sq.data_ptr(), f.chunk_seg.data_ptr(), self.seg_steps.data_ptr(), f.flags.data_ptr(),
             self.seg_coef.data_ptr(), f.numel, f.align, f.nseg, self.lr, self.betas[0], self.betas[1], self.eps,
             None if f.mirror is None else f.mirror.data_ptr(), _stream())
        ops.weights_changed()      # raw-pointer update: torch's version counters did not move

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.flat.zero_grad()


class PeerExchange:
    """Symmetric NVLink peer memory of the fused exchange + Adam kernel (`mar_dp_allreduce_adam`, csrc/dp_exchange.cu):
    every rank cudaMallocs one block [arrival flags | bf16 wire copy | bf16 reduced gradient], exports it as a CUDA IPC
    handle, the handles are all-gathered over the process group and every rank maps its peers' blocks.  One node only.
    Raises if any rank fails to set it up (the caller then keeps the NCCL all-reduce)."""

    def __init__(self, n: int, group=None):
        import ctypes
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n = int(n)
        lib = _lib_load()
        self._own = ctypes.c_void_p()
        self._opened: List = []
        self.blocks = (ctypes.c_void_p * self.world)()
        err = None
        handle = ctypes.create_string_buffer(64)
        try:
            call("mar_peer_alloc", ctypes.byref(self._own), int(lib.mar_dp_block_bytes(self.n)))
            call("mar_peer_export", self._own, handle)
        except RuntimeError as e:            # keep the collective below symmetric: every rank learns that one failed
            err = str(e)
        gathered: List = [None] * self.world
        dist.all_gather_object(gathered, (err, bytes(handle.raw), _host_id()), group=group)
        errs = [g[0] for g in gathered if g[0] is not None]
        if errs or len({g[2] for g in gathered}) != 1:
            self.close()
            raise RuntimeError("peer-memory exchange unavailable: " + (errs[0] if errs else "ranks on different hosts"))
        try:
            for r, (_, h, _) in enumerate(gathered):
                if r == self.rank:
                    self.blocks[r] = self._own.value
                else:
                    ptr = ctypes.c_void_p()
                    call("mar_peer_import", ctypes.create_string_buffer(h, 64), ctypes.byref(ptr))
                    self._opened.append(ptr)
                    self.blocks[r] = ptr.value
        except RuntimeError as e:
            err = str(e)
        ok = [None] * self.world
        dist.all_gather_object(ok, err, group=group)
        if any(o is not None for o in ok):
            self.close()
            raise RuntimeError("peer-memory exchange unavailable: " + next(o for o in ok if o is not None))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.ctrl = torch.zeros((int(lib.mar_dp_ctrl_bytes()) + 7) // 8, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group=group)             # every block is zeroed and mapped before anybody's first kernel

    def check(self) -> None:
        """Synchronises the current stream; raises if a wait on a peer inside the kernel ever timed out."""
        call("mar_dp_check", self.ctrl.data_ptr(), _stream())

    def close(self, barrier: bool = False) -> None:
        """Unmap the peers' blocks and free the own one.  A block must not be freed while a peer still has it mapped
        (CUDA IPC), so with barrier=True — a COLLECTIVE call, every rank of the group makes it — the ranks meet between the
        two; without a barrier only the mappings are closed and the own block is left to the context's teardown."""
        for ptr in self._opened:
            try:
                call("mar_peer_close", ptr)
            except RuntimeError:
                pass
        self._opened = []
        if not barrier:
            return
        dist.barrier(group=self.group)
        if self._own is not None and self._own.value:
            try:
                call("mar_peer_free", self._own)
            except RuntimeError:
                pass
            self._own = None


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pin this process (one process per GPU) to the CPUs NVML names as local to GPU `device_index` — before pinned host
    buffers are allocated, so that first touch puts the staging memory of the H2D copies on the GPU's own NUMA node (with 8
    ranks streaming 230 MB of fp32 features per step each, remote-node staging is what the end-to-end step waits for).
    Returns the CPU list, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def _lib_load():
    from . import _lib
    return _lib.load()


def _host_id() -> str:
    import socket
    try:
        return open("/proc/sys/kernel/random/boot_id").read().strip()
    except OSError:
        return socket.gethostname()


class GradSync:
    """Tracks which parameters received a gradient this step and, with more than one rank, exchanges the gradients.

    * Every parameter is counted ONCE per step, by its post-accumulate-grad hook or by `notify` (a kernel accumulated
      the gradient straight into the flat buffer, ops.grad_sink), whichever comes first.  `finish()` writes the
      resulting active flags into the gradient buffer's header (FlatParams.flags) for the per-parameter Adam.
    * Data parallel: bucketed all-reduce (SUM — every rank has already weighed its loss, see TrainStep) over slices
      of the flat gradient buffer on a side stream, overlapping the rest of backward.  Buckets are contiguous slices
      taken from the END of the buffer (backward produces gradients in roughly reverse parameter order).  The
      collective ORDER is the same on every rank whatever its shard of the batch activates: bucket b is launched
      only after buckets 0..b-1, as soon as all of them are complete locally, the remainder in `finish()` (a rank
      whose slice lacks a modality or a head reduces zeros, at the same place in the sequence as its peers).  The
      last bucket holds the header flags and always goes from `finish()`: a parameter is active if ANY rank saw a
      gradient for it, as in a single process on the global batch."""

    def __init__(self, flat: FlatParams, group=None, num_buckets: int = 4, wire: str = "fp32", tail_elems: int = 0):
        """wire: "fp32" (exact) or "bf16" — gradients are rounded to bf16 for the exchange (half the bytes on NVLink; the
        sum is accumulated by NCCL in bf16) and widened again before Adam; the bf16-mode step uses it, fp32 mode never.
        tail_elems > 0: the LAST bucket (the first parameters = the end of backward, whose all-reduce nothing can hide)
        is cut as small as parameter boundaries allow above that many elements."""
        self.flat = flat
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat.grad.is_cuda
        self.comm_stream = torch.cuda.Stream() if (self.cuda and self.world > 1) else None
        if wire not in ("fp32", "bf16"):
            raise ValueError("wire must be 'fp32' or 'bf16'")
        self.wire = wire if (self.cuda and self.world > 1) else "fp32"
        n = len(flat.params)
        total = flat.numel
        # the small tail bucket first (from parameter 0 upwards), then the rest split by element count from the end,
        # all aligned to parameter boundaries
        first = 0
        if tail_elems > 0 and num_buckets > 1:
            acc = 0
            while first < n - 1 and acc < tail_elems:
                acc += flat.params[first].numel()
                first += 1
        rest = sum(flat.params[i].numel() for i in range(first, n))
        bounds, target, acc = [n], rest / max(1, num_buckets - (1 if first else 0)), 0
        for i in range(n - 1, first - 1, -1):
            acc += flat.params[i].numel()
            if acc >= target and i > first:
                bounds.append(i)
                acc = 0
        if first:
            bounds.append(first)
        bounds.append(0)
        self.buckets = []           # (param index range [lo, hi), element range [e0, e1))
        for hi, lo in zip(bounds[:-1], bounds[1:]):
            if lo < hi:
                e0 = flat.offsets[lo] if lo > 0 else 0          # the bucket of parameter 0 also carries the header
                e1 = flat.offsets[hi] if hi < n else total
                self.buckets.append((lo, hi, e0, e1))
        self.bucket_of = {}
        for b, (lo, hi, _, _) in enumerate(self.buckets):
            for i in range(lo, hi):
                self.bucket_of[i] = b
        self.pending = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self._streams = [dict() for _ in self.buckets]
        self.fired = [False] * n
        self.order: List[int] = []
        self.enabled = True          # False: hooks do nothing (backward passes outside a TrainStep, e.g. profiling)
        self.index_of = {id(p): i for i, p in enumerate(flat.params)}
        self._hooks = [self._make_hook(i) for i in range(n)]
        self._masks: Dict[bytes, torch.Tensor] = {}      # active pattern -> device flags (one upload per pattern)
        self._wire_buf = None
        if self.wire == "bf16":
            self._wire_buf = torch.empty(max(e1 - e0 for _, _, e0, e1 in self.buckets), dtype=torch.bfloat16, device=flat.grad.device)
            self._wire_done: List[Optional[torch.cuda.Event]] = [None]
        self.last_active: Optional[List[bool]] = None
        for i, p in enumerate(flat.params):
            p.register_post_accumulate_grad_hook(self._hooks[i])
        # Fused exchange + Adam over NVLink peer memory (one kernel instead of cast + ncclAllReduce + cast + Adam):
        # bf16 wire, ONE bucket (the measured optimum anyway), NCCL process group on one node.  MAR_DP_FUSED=0 keeps NCCL.
        self.peer: Optional[PeerExchange] = None
        self.fused_opt = None            # the FlatAdam whose step the fused kernel performs (set by TrainStep)
        self.fused_write_back = False    # also leave the reduced gradient in the fp32 .grad views
        if (self.wire == "bf16" and len(self.buckets) == 1 and _os.environ.get("MAR_DP_FUSED", "1") != "0" and not _NOCOMM
                and dist.get_backend(group) == "nccl" and flat.align >= 8):
            try:
                self.peer = PeerExchange(flat.numel, group)
            except RuntimeError as e:
                import warnings
                warnings.warn(f"{e}; falling back to the NCCL all-reduce")
        self.reset()

    def notify(self, param) -> None:
        """A kernel accumulated this parameter's gradient straight into the flat buffer (ops.grad_sink) and its
        autograd Function returned None for it: count the parameter here, right after the kernel was enqueued.
        torch still fires the parameter's post-accumulate hook afterwards (AccumulateGrad runs with an undefined
        gradient, measured on torch 2.11) — every parameter is counted ONCE per step, whichever comes first
        (counting both launched the buckets' all-reduces before their gradients were complete: found on 2 B200s
        with tools/dp_diag.py, the ranks' parameters drifted apart)."""
        i = self.index_of.get(id(param))
        if i is not None:
            self._hooks[i](param)

    def reset(self) -> None:
        for b, (lo, hi, _, _) in enumerate(self.buckets):
            self.pending[b] = hi - lo
            self.launched[b] = False
            self._streams[b] = {}
        self.fired = [False] * len(self.flat.params)
        self.order = []

    def _make_hook(self, i: int) -> Callable:
        def hook(_param):
            if not self.enabled or self.fired[i]:
                return
            self.fired[i] = True
            if self.world == 1:
                return
            b = self.bucket_of[i]
            if self.comm_stream is not None:        # the stream this gradient was written on (autograd may run a
                s = torch.cuda.current_stream()     # kept-alive AccumulateGrad node on the stream of an earlier step)
                self._streams[b][s.cuda_stream] = s
            self.pending[b] -= 1
            self._launch_ready()
        return hook

    def _launch_ready(self) -> None:
        """Launch, in index order, every bucket whose predecessors are out and whose gradients are complete —
        except the last one (header flags), which waits for finish()."""
        for b in range(len(self.buckets) - 1):
            if self.launched[b]:
                continue
            if self.pending[b] != 0:
                return
            self._launch(b)

    def _launch(self, b: int) -> None:
        self.launched[b] = True
        self.order.append(b)
        if _NOCOMM:          # diagnostic only (MAR_DP_NOCOMM=1): what the step costs without any exchange — ranks drift apart
            return
        _, _, e0, e1 = self.buckets[b]
        view = self.flat.grad[e0:e1]
        if self.comm_stream is not None:
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            for key, s in self._streams[b].items():     # every stream a gradient of this bucket was written on
                if key != cur.cuda_stream:
                    self.comm_stream.wait_stream(s)
            for bb in range(b):                          # a bucket launched late also covers earlier stragglers' streams
                for key, s in self._streams[bb].items():
                    if key != cur.cuda_stream:
                        self.comm_stream.wait_stream(s)
            with torch.cuda.stream(self.comm_stream):
                if self.wire == "bf16":
                    # fp32 slice -> bf16 staging -> all-reduce -> back into the fp32 slice, all on the communication stream
                    # (one staging buffer: the buckets go through it one after the other in stream order)
                    stage = self._wire_buf[: e1 - e0]
                    st = self.comm_stream.cuda_stream
                    call("mar_cast", view.data_ptr(), 0, stage.data_ptr(), 1, e1 - e0, st)
                    dist.all_reduce(stage, op=dist.ReduceOp.SUM, group=self.group)
                    call("mar_cast", stage.data_ptr(), 1, view.data_ptr(), 0, e1 - e0, st)
                else:
                    dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def _write_flags(self) -> None:
        active = bytes(bytearray(1 if f else 0 for f in self.fired))
        mask = self._masks.get(active)
        if mask is None:
            if self.cuda and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("TrainStep: a batch layout activated a set of parameters that no eager step has "
                                   "seen yet; its flags cannot be uploaded during CUDA-graph capture")
            mask = torch.tensor([float(b) for b in active], dtype=torch.float32).to(self.flat.grad.device)
            self._masks[active] = mask
        self.flat.flags.copy_(mask)
        self.last_active = list(self.fired)

    def fused(self) -> bool:
        return self.peer is not None and self.fused_opt is not None

    def _launch_fused(self) -> None:
        """gradient exchange AND the Adam step, one kernel on the current stream (nothing left to overlap with)."""
        import ctypes
        f, o, pe = self.flat, self.fused_opt, self.peer
        cur = torch.cuda.current_stream()
        for b in range(len(self.buckets)):           # gradients written on other streams (kept-alive AccumulateGrad nodes)
            for key, s in self._streams[b].items():
                if key != cur.cuda_stream:
                    cur.wait_stream(s)
            self.launched[b] = True
        call("mar_dp_allreduce_adam", f.grad.data_ptr(), int(self.fused_write_back), f.numel, pe.world, pe.rank,
             ctypes.cast(pe.blocks, ctypes.c_void_p), pe.ctrl.data_ptr(), f.flat.data_ptr(), o.exp_avg.data_ptr(),
             o.exp_avg_sq.data_ptr(), f.chunk_seg.data_ptr(), o.seg_steps.data_ptr(), f.align, f.nseg, o.lr, o.betas[0],
             o.betas[1], o.eps, None if f.mirror is None else f.mirror.data_ptr(), cur.cuda_stream)
        ops.weights_changed()

    def finish(self) -> None:
        self._write_flags()
        if self.world > 1 and self.fused():
            self._launch_fused()
            self.reset()
            return
        if self.world > 1:
            for b in range(len(self.buckets)):
                if not self.launched[b]:
                    self._launch(b)
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.reset()


class TrainStep:
    """One optimizer step of the hot path as a single call:  losses = step(data, labels).

    model(data) → criterion(pred, labels) → LossesDict.backward() → gradient all-reduce (if world > 1) → Adam.
    Returns the dict of per-head losses as 0-d DEVICE tensors (no host sync).  With `graph=True` the whole step —
    including the bucketed NCCL all-reduces on the side stream when world > 1 — is captured into a CUDA graph on
    first use and replayed afterwards: dropout masks advance on the device, ~150 kernel launches collapse into one
    graph launch.  Inputs go through TWO sets of static buffers (and two captured graphs) used alternately: when
    the batch is given as pinned HOST tensors, the H2D copy of step i+1 runs on a copy stream while the graph of
    step i is still executing, so the transfer is hidden behind compute without any change to the call."""

    def __init__(self, model: nn.Module, criterion: Callable, lr: float = 1e-3, graph: bool = False,
                 group=None, num_buckets: int = 1, precision: Optional[str] = None, wire: Optional[str] = None,
                 tail_elems: int = 0):
        # num_buckets = 1 (ONE all-reduce after backward) is the measured optimum on B200 for this model (19.7 M parameters,
        # 39 MB in bf16): NCCL's CTAs running beside the persistent one-CTA-per-SM GEMMs cost the GEMMs as much as the
        # overlap hides — N = 8: 6 buckets 9.34 ms, 2 buckets 9.29, 1 bucket 9.28, no exchange at all 8.99 ms per step
        # (profiles/r02_dp_matrix_n8.txt); larger models / slower links: raise num_buckets, set tail_elems.
        self.model, self.criterion = model, criterion
        self.flat = FlatParams(list(model.parameters()))
        self.opt = FlatAdam(self.flat, lr=lr)
        # The hooks' AccumulateGrad nodes remember the stream they were created on: create them on the stream the graph-
        # captured step runs on (the warm-up steps and every capture use this ONE side stream), or autograd inserts a
        # cross-stream synchronisation per parameter and warns about the mismatch on every backward.
        self._side = torch.cuda.Stream() if (graph and self.flat.flat.is_cuda) else None
        with (torch.cuda.stream(self._side) if self._side is not None else _null()):
            if wire is None:       # bf16 compute -> bf16 gradient exchange; fp32 mode keeps the exchange exact
                mode = precision if precision is not None else ("fp32" if ops.get_precision() == torch.float32 else "bf16")
                wire = "bf16" if (mode == "bf16" and self.flat.flat.is_cuda) else "fp32"
            self.sync = GradSync(self.flat, group=group, num_buckets=num_buckets, wire=wire, tail_elems=tail_elems)
            self.sync.fused_opt = self.opt       # with peer memory available, the exchange kernel also does the Adam step
        if self.sync.world > 1 and self.flat.flat.is_cuda and ops._rng.seed is None:
            # every rank seeds torch identically (same initial weights), which would also give every rank the SAME
            # dropout masks for its different clips; a single process on the global batch draws independent masks per
            # clip, so each rank gets its own mask stream unless the caller seeded ops.manual_seed() itself
            ops.manual_seed((torch.initial_seed() + 0x9E3779B97F4A7C15 * dist.get_rank(group)) & 0x7FFFFFFFFFFFFFFF)
        self.use_graph = graph
        self.precision = precision
        # captured graphs are keyed by the batch SIGNATURE: tensor shapes/dtypes AND the non-tensor leaves (the
        # per-sample modality / label names whose `_EMPTY` markers steer the model's control flow, datasets.py:564-608)
        # — a batch with another signature (last batch of an epoch, a verb-only batch) never replays a foreign graph
        self._graphs = {}
        self.max_signatures = 4          # further signatures run eagerly (each capture owns its activation pool)
        self._sets = None                # the two buffer sets of the most recent signature (kept for introspection)
        self._calls = 0
        self._copy_stream = None
        self._warm = 0
        self.captured_launches = 0
        self.last_pred = None
        self._norm_local = self._norm_global = None

    # -- helpers ------------------------------------------------------------------------------
    @staticmethod
    def _tensors(batch) -> List[torch.Tensor]:
        out = []
        def walk(x):
            if isinstance(x, torch.Tensor):
                out.append(x)
            elif isinstance(x, (list, tuple)) and not (x and isinstance(x[0], str)):
                for y in x:
                    walk(y)
        walk(batch)
        return out

    @staticmethod
    def _like(batch, tensors):
        it = iter(tensors)
        def walk(x):
            if isinstance(x, torch.Tensor):
                return next(it)
            if isinstance(x, (list, tuple)) and not (x and isinstance(x[0], str)):
                return [walk(y) for y in x]
            return x
        return walk(batch)

    def _loss_weights(self, labels):
        """Data parallel only.  nn.CrossEntropyLoss averages over ITS rows (the sum of their class weights): with a
        different number of non-EMPTY rows per rank — or a head absent from a rank's slice — the mean of the ranks'
        gradients is not the gradient of the reference loss on the global batch.  Every rank therefore weighs each
        head's loss by  local_norm / Σ_ranks norm  and the gradients are SUMMED.  The norms depend on the labels
        alone, so their (tiny) all-reduce is launched on the communication stream before the forward pass and is
        off the critical path.  Returns {head: 0-d device tensor}, or a float for criteria that do not expose
        `label_weight_sums` (uniform 1 / world: equal slices assumed)."""
        world = self.sync.world
        if world == 1:
            return None
        fn = getattr(self.criterion, "label_weight_sums", None)
        heads = fn(labels, self.flat.flat.device, None) if fn is not None else None
        if heads is None:
            return 1.0 / world
        if self._norm_local is None or self._norm_local.numel() != len(heads):
            self._norm_local = torch.zeros(len(heads), dtype=torch.float32, device=self.flat.flat.device)
            self._norm_global = torch.zeros_like(self._norm_local)
        fn(labels, self.flat.flat.device, self._norm_local)
        self._norm_global.copy_(self._norm_local)
        comm = self.sync.comm_stream
        if comm is not None:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_reduce(self._norm_global, op=dist.ReduceOp.SUM, group=self.sync.group)
        else:
            dist.all_reduce(self._norm_global, op=dist.ReduceOp.SUM, group=self.sync.group)
        return heads

    def _backward(self, losses, weights) -> None:
        if weights is None:
            losses.backward()
            return
        items = [(k, l) for k, l in losses.items() if isinstance(l, torch.Tensor) and l.requires_grad]
        if not items:
            return
        if isinstance(weights, float):
            grads = [torch.full_like(l, weights) for _, l in items]
        else:
            if self.sync.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.sync.comm_stream)
            w = self._norm_local / self._norm_global.clamp_min(1e-30)
            grads = [w[weights.index(k)].to(l.dtype).reshape(l.shape) for k, l in items]
        torch.autograd.backward([l for _, l in items], grad_tensors=grads)

    def _eager(self, data, labels):
        self.opt.zero_grad()
        if self.flat.flat.is_cuda:
            ops.rng_advance(self.flat.flat.device)
        weights = self._loss_weights(labels)
        pred = self.model(data)
        losses = self.criterion(pred, labels)
        if isinstance(losses, torch.Tensor):
            losses = LossesDict(loss=losses)
        with ops.grad_sink(self.sync.notify):       # weight / bias / LayerNorm gradients accumulate straight into the flat buffer
            self._backward(losses, weights)
        fused = self.sync.world > 1 and self.sync.fused()
        self.sync.finish()
        if not fused:                 # the fused exchange kernel has already applied Adam
            self.opt.step()
        self.last_pred = pred
        return {k: v.detach() for k, v in losses.items()}

    def __call__(self, data, labels):
        ctx = ops.precision(self.precision) if self.precision else _null()
        with ctx:
            if not self.use_graph:
                dev = self.flat.flat.device
                if dev.type == "cuda":
                    data = self._to_dev(data, dev)
                    labels = self._to_dev(labels, dev)
                return self._eager(data, labels)
            return self._graphed(data, labels)

    @staticmethod
    def _to_dev(batch, dev):
        def walk(x):
            if isinstance(x, torch.Tensor):
                return x.to(dev, non_blocking=True)
            if isinstance(x, (list, tuple)) and not (x and isinstance(x[0], str)):
                return [walk(y) for y in x]
            return x
        return walk(batch)

    @staticmethod
    def _signature(batch):
        sig = []
        def walk(x):
            if isinstance(x, torch.Tensor):
                sig.append((tuple(x.shape), str(x.dtype)))
            elif isinstance(x, (list, tuple)) and not (x and isinstance(x[0], str)):
                sig.append(("[", len(x)))
                for y in x:
                    walk(y)
            elif isinstance(x, (list, tuple)):
                sig.append(tuple(x))
            else:
                sig.append(repr(x))
        walk(batch)
        return tuple(sig)

    def _graphed(self, data, labels):
        dev = self.flat.flat.device
        src = self._tensors([data, labels])
        sig = self._signature([data, labels])
        state = self._graphs.get(sig)
        if state is None:
            if len(self._graphs) >= self.max_signatures:        # too many distinct batch layouts: stay eager for this one
                return self._eager(self._to_dev(data, dev), self._to_dev(labels, dev))
            state = {"sets": [dict(static_in=None, graph=None, static_out=None, static_pred=None, done=None) for _ in range(2)],
                     "calls": 0, "warm": 0}
            self._graphs[sig] = state
        self._sets = state["sets"]
        cur = state["sets"][state["calls"] & 1]
        state["calls"] += 1
        self._calls += 1
        if cur["static_in"] is None:
            cur["static_in"] = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in src]
        main = torch.cuda.current_stream()
        from_host = all(not t.is_cuda for t in src)
        if from_host and self._graph_ready():
            # H2D on the copy stream: it only has to wait for the last replay that READ this buffer set, not for
            # the step that is executing right now on the main stream (that one reads the other set)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            cs = self._copy_stream
            if cur["done"] is not None:
                cs.wait_event(cur["done"])
            with torch.cuda.stream(cs):
                for s_, t in zip(cur["static_in"], src):
                    s_.copy_(t, non_blocking=True)
            main.wait_stream(cs)
        else:
            for s_, t in zip(cur["static_in"], src):
                s_.copy_(t, non_blocking=True)
        sdata, slabels = self._like([data, labels], cur["static_in"])
        if cur["graph"] is None:
            if self._side is None:
                self._side = torch.cuda.Stream()
            side = self._side
            need = 3 if self._warm < 3 else 1       # eager steps before a capture: 3 for the first, 1 for any new signature
            if state["warm"] < need:                # (uploads the signature's index / flag tensors, which capture cannot)
                state["warm"] += 1
                self._warm += 1
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    out = self._eager(sdata, slabels)
                main.wait_stream(side)
                return out
            ops.clear_weight_cache()
            g = torch.cuda.CUDAGraph()
            before = ops.launch_count()
            side.wait_stream(main)
            with torch.cuda.graph(g, stream=side):
                cur["static_out"] = self._eager(sdata, slabels)
            main.wait_stream(side)
            self.captured_launches = ops.launch_count() - before   # libmar kernels per replay
            cur["static_pred"] = self.last_pred     # this graph's own logits buffers
            cur["graph"] = g                        # capture records but does not execute: fall through to replay
        cur["graph"].replay()
        ops.weights_changed()                       # the replay's Adam ran without any Python: cached bf16 copies are stale
        self.last_pred = cur["static_pred"]         # valid until this buffer set's next replay (two steps later)
        if cur["done"] is None:
            cur["done"] = torch.cuda.Event()
        cur["done"].record(main)
        return cur["static_out"]

    def _graph_ready(self) -> bool:
        return self._warm >= 3

    def release_graphs(self) -> None:
        """Drop the captured graphs (and their memory pools) and the peer-memory blocks of the fused exchange (later steps
        fall back to the NCCL all-reduce).  COLLECTIVE with more than one rank: every rank calls it, before
        torch.distributed.destroy_process_group() — a communicator must not be torn down while graphs that captured its
        collectives are still alive, and a peer block must not be freed while another rank still maps it."""
        if self.flat.flat.is_cuda:
            torch.cuda.synchronize()
        if self.sync.peer is not None:
            peer, self.sync.peer = self.sync.peer, None
            peer.close(barrier=True)
        for state in self._graphs.values():
            for s in state["sets"]:
                s["graph"] = None
                s["static_out"] = None
                s["static_pred"] = None
                s["done"] = None
        self._graphs = {}
        if self.flat.flat.is_cuda:
            torch.cuda.synchronize()


def shard_batch(batch, rank: int, world: int):
    """Rank `rank`'s contiguous share of every per-sample sequence of a (nested) batch in the datasets.py layouts:
    tensors are sliced along dim 0, name tuples (the `_EMPTY` markers, datasets.py:592-608) along their length."""
    def cut(n):
        return n * rank // world, n * (rank + 1) // world

    def walk(x):
        if isinstance(x, torch.Tensor):
            lo, hi = cut(x.shape[0])
            return x[lo:hi]
        if isinstance(x, (list, tuple)) and x and isinstance(x[0], str):
            lo, hi = cut(len(x))
            return type(x)(x[lo:hi])
        if isinstance(x, (list, tuple)):
            return [walk(y) for y in x]
        return x
    return walk(batch)


class ShardedBatchSampler(torch.utils.data.Sampler):
    """Data-parallel wrapper for the reference's `AggrBatchSampler` (datasets.py:620-655), as the `batch_sampler` of every
    rank's DataLoader.  That sampler makes every batch homogeneous in aggression type and reshuffles with
    `random.seed(None)` (datasets.py:636,643), so two ranks building their own would draw different batches: here rank 0's
    list of index batches is broadcast once per epoch and every rank yields ITS contiguous slice of the SAME batch — all
    ranks step on one homogeneous batch, i.e. the same set of active heads and modalities, exactly as a single process on
    the global batch (batch_size of the wrapped sampler = the GLOBAL batch)."""

    def __init__(self, base, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.base, self.group = base, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world

    def __iter__(self):
        batches = [list(b) for b in self.base] if self.rank == 0 else None     # (iterating also reshuffles the reference's sampler)
        if self.world > 1:
            box = [batches]
            dist.broadcast_object_list(box, src=0, group=self.group)
            batches = box[0]
        for b in batches:
            lo, hi = len(b) * self.rank // self.world, len(b) * (self.rank + 1) // self.world
            yield b[lo:hi]

    def __len__(self):
        return len(self.base)


class EpochAccumulator:
    """The reference trainers' per-epoch bookkeeping without their per-step host reads.

    `TorchSupervisedTrainer.train_step` / `.test_step` end every step with `loss.item()` per head
    (`compute_batch_loss`, trainer.py:175-177, :731-737) and `torch.max(pred, 1)[1].cpu().numpy()` per head
    (`nn_output_processing`, trainer.py:165-171, :718-729) plus the labels' `.cpu()` (`create_batch_results_dict`,
    trainer.py:179, :888-914): 2 x heads + 1 blocking device->host reads per step.  Here `add()` only enqueues
    device work (a multiply-add into a running loss sum per head, an argmax kernel, references to the label
    tensors); `results()` does ONE device->host transfer per epoch and then builds exactly the dictionary
    `compute_epoch_results` returns (trainer.py:237-283 single model, :739-812 multi-head `RNN_trainer`,
    :916-1007 `MultimodalTrainer`):  {head: {'loss': sum_steps(loss x size) / dataset_size, metric: value, ...}}.

    Kept verbatim from the reference: `size` is `len(data[0])` for list inputs and `len(data)` otherwise
    (trainer.py:155-158) - for the multimodal layout `data[0]` is the `[names, tensor]` pair, so size is 2, not the
    batch size; label groups / heads whose samples are all `_EMPTY` are dropped from that step's losses and
    predictions, partially-EMPTY groups keep only their present rows (trainer.py:893-912); metrics are called as
    `metric(true, pred, **kwargs)` with `metrics_dict` entries being either a callable or
    `{'metric': f, 'kwargs': {...}}`, and any entry named 'loss' (case-insensitive) is skipped."""

    def __init__(self, metrics_dict: Optional[Dict] = None):
        self.metrics_dict = dict(metrics_dict or {})
        self.reset()

    def reset(self) -> None:
        self._loss: Dict[str, torch.Tensor] = {}     # head -> running sum on the device
        self._pred: Dict[str, List[torch.Tensor]] = {}
        self._true: Dict[str, List[torch.Tensor]] = {}
        self._keep: Dict[str, List] = {}             # head -> per step: None (all rows) or a host bool array
        self._flat = False                           # single-output trainer: results() returns {'loss':…, metric:…}
        self.steps = 0

    @staticmethod
    def batch_size_as_reference(data) -> int:
        """trainer.py:155-158."""
        return len(data[0]) if isinstance(data, list) else len(data)

    @staticmethod
    def _argmax(logits: torch.Tensor) -> torch.Tensor:
        return ops.argmax_rows(logits.detach())        # (the host-logic tests install their own stand-in)

    def add(self, losses, pred, labels, data=None, size: Optional[int] = None) -> None:
        """One step's results; nothing here waits for the device.  `losses`: {head: 0-d tensor} (or a tensor),
        `pred`: {head: (B,C) logits} (or a tensor), `labels`: (B,) tensor or the multimodal `[[names, y], ...]`."""
        if size is None:
            size = self.batch_size_as_reference(data) if data is not None else 1
        self._flat = isinstance(pred, torch.Tensor)
        losses = {"loss": losses} if isinstance(losses, torch.Tensor) else losses
        pred = {"loss": pred} if isinstance(pred, torch.Tensor) else pred
        heads, true, keep = list(pred.keys()), {}, {}
        if isinstance(labels, torch.Tensor):
            for h in heads:                            # every head is scored against the same labels (trainer.py:739-812)
                true[h], keep[h] = labels, None
        else:
            present_heads = []
            for names, y in labels:
                head = names[0].split('_')[0]
                present = [n.split('_')[-1] != 'EMPTY' for n in names]
                if head in pred and any(present):
                    present_heads.append(head)
                    true[head] = y
                    keep[head] = None if all(present) else present
            heads = [h for h in heads if h in present_heads]
        for h in heads:
            if h in losses:
                term = losses[h].detach().float() * float(size)
                self._loss[h] = term if h not in self._loss else self._loss[h] + term
            self._pred.setdefault(h, []).append(self._argmax(pred[h]))
            self._true.setdefault(h, []).append(true[h].detach().clone())     # the caller may reuse its label buffers
            self._keep.setdefault(h, []).append(keep[h])
        self.steps += 1

    def results(self, dataset_size: int) -> Dict[str, Dict]:
        """ONE device->host transfer, then the reference's epoch dictionary."""
        import numpy as np
        heads = list(self._pred.keys())
        if not heads:
            return {}
        dev = self._pred[heads[0]][0].device
        parts, layout = [], []
        for h in heads:
            p = torch.cat(self._pred[h]).to(torch.int64)
            t = torch.cat([x.to(dev) for x in self._true[h]]).to(torch.int64)
            parts += [p, t]
            layout.append((h, p.numel()))
        loss_heads = [h for h in heads if h in self._loss]
        packed = torch.cat(parts).double()
        if loss_heads:
            packed = torch.cat([packed, torch.stack([self._loss[h].double() for h in loss_heads])])
        host = packed.cpu().numpy()                    # the epoch's only synchronisation
        out, off = {}, 0
        for h, n in layout:
            p = host[off:off + n].astype(np.int64)
            t = host[off + n:off + 2 * n].astype(np.int64)
            off += 2 * n
            if any(k is not None for k in self._keep[h]):
                rows = np.concatenate([np.ones(x.numel(), dtype=bool) if k is None else np.asarray(k, dtype=bool)
                                       for x, k in zip(self._pred[h], self._keep[h])])
                p, t = p[rows], t[rows]
            out[h] = {"pred": p, "true": t}
        log = {}
        for i, h in enumerate(loss_heads):
            log[h] = {"loss": float(host[off + i]) / dataset_size}
        for h in heads:
            entry = log.setdefault(h, {})
            for name, metric in self.metrics_dict.items():
                if name.lower() == "loss":
                    continue
                fn, kw = (metric["metric"], metric["kwargs"]) if type(metric) is dict else (metric, {})
                entry[name] = fn(out[h]["true"], out[h]["pred"], **kw)
        self.last_arrays = out
        return log["loss"] if self._flat else log


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
