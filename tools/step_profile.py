#!/usr/bin/env python
"""Exactly K eager C3 train steps (TrainStep, no CUDA graph so every kernel keeps its name) between
cudaProfilerStart/Stop, for a clean per-kernel launch list of ONE step:

    python tools/step_profile.py --steps 2 > gpurun_out/plain.log 2>&1 &&
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/step_profile.py --steps 2
    python tools/launch_summary.py gpurun_out/launches.csv 2

(bench.py's own command line also runs the roofline passes and the other configurations, which pollute a launch list.)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import models as M, training, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--t-audio", type=int, default=250)
ap.add_argument("--t-video", type=int, default=64)
args = ap.parse_args()
dev = torch.device("cuda:0")
mar.set_precision("bf16")
torch.manual_seed(0)
model = W.build_c3(M, args.t_audio, args.t_video).to(dev).train()
crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
data, labels = W.batch_c3(B=args.batch, t_audio=args.t_audio, t_video=args.t_video)
data, labels = W.to_device(data, dev), W.to_device(labels, dev)
step = training.TrainStep(model, crit, graph=False)
for _ in range(args.warmup):
    step(data, labels)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    step(data, labels)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
