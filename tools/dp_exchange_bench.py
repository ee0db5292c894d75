#!/usr/bin/env python
"""The data-parallel gradient exchange + Adam step alone, on N GPUs of one node, both engines on the same buffers:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_exchange_bench.py [--iters 30]

  fused : mar_dp_allreduce_adam — ONE kernel: fp32 -> bf16 wire copy, reduce-scatter by peer loads, all-gather by peer
          stores, per-parameter Adam (csrc/dp_exchange.cu)
  nccl  : mar_cast + ncclAllReduce(bf16) + mar_cast + mar_adam_step_segments (what GradSync does without peer memory)

Timed with CUDA events on the launching stream after a barrier, max over ranks, median over iterations.  Rank 0 prints
one JSON line; `nvlink_floor_us` = bytes that cross one rank's links per direction (reduce-scatter + all-gather:
2·(N-1)/N of the bf16 buffer) / 770 GB/s (the measured peer-copy bandwidth of this pool, B200_PROFILING.md), `hbm_floor_us` = local bytes (cast + Adam) / measured
copy bandwidth."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import models as M, training, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
mar.set_precision("bf16")
torch.manual_seed(0)
model = W.build_c3(M).to(dev).train()
crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})


def timed(step, fused):
    sync, flat = step.sync, step.flat
    assert sync.fused() == fused, (sync.fused(), fused)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    times = []
    for it in range(args.iters + 5):
        flat.grad.copy_(torch.randn(flat.numel, device=dev, generator=g) * 1e-3)
        sync.fired = [True] * len(flat.params)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sync.finish()
        if not fused:
            step.opt.step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 5:
            times.append(float(t))
    times.sort()
    digest = flat.flat.view(torch.int32).sum(dtype=torch.int64).reshape(1)
    allv = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(allv, digest)
    if sync.peer is not None:
        sync.peer.check()
    return times[len(times) // 2], times[0], all(torch.equal(allv[0], a) for a in allv[1:])


out = {}
for name, env in (("fused", "1"), ("nccl", "0")):
    os.environ["MAR_DP_FUSED"] = env
    step = training.TrainStep(model, crit, graph=False)
    med, best, same = timed(step, env == "1")
    out[name] = {"us_median": round(med * 1e3, 1), "us_best": round(best * 1e3, 1), "ranks_identical": same}
    n = step.flat.numel
    del step
if rank == 0:
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out.update({"n_gpus": world, "params": n, "wire_bytes": 2 * n,
                "nvlink_floor_us": round(2 * (2 * n) * (world - 1) / world / 770e9 * 1e6, 1),   # reduce-scatter + all-gather
                "hbm_floor_us": round((4 * n + 2 * n + 2 * n + 12 * n + 14 * n) / (hbm * 1e9) * 1e6, 1),
                "hbm_gbs_used": hbm})
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
