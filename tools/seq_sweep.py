#!/usr/bin/env python
"""BASELINE.json configs[4]: multimodal fusion sequence-length sweep T_a = T_v = T in {128..2048} (fused length 2T:
the attention-bound regime), training step and inference, one B200.  One JSON line per T.
usage: python tools/seq_sweep.py [--T 128 256 512 1024 2048] [--tokens 65536] [--steps 8]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import models as M, ops, training, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, nargs="*", default=[128, 256, 512, 1024, 2048])
ap.add_argument("--tokens", type=int, default=65536, help="audio tokens per step (batch = tokens / T)")
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--warmup", type=int, default=4)
args = ap.parse_args()
dev = torch.device("cuda:0")
mar.set_precision("bf16")
peak = 1388.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass

def timed(fn, steps, warmup):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

for T in args.T:
    B = max(1, args.tokens // T)
    torch.manual_seed(0)
    model = W.build_c3(M, T, T).to(dev).train()
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
    data, labels = W.batch_c3(B=B, t_audio=T, t_video=T, seed=1000)
    gdata, glabels = W.to_device(data, dev), W.to_device(labels, dev)
    step = training.TrainStep(model, crit, graph=False)
    ms_train = timed(lambda: step(gdata, glabels), args.steps, args.warmup)
    model.eval()
    def infer():
        with torch.no_grad():
            return model(gdata)
    ms_inf = timed(infer, args.steps, args.warmup)
    fl = W.c3_flops_per_clip(T, T)
    line = {"T": T, "fused_len": 2 * T, "batch": B, "train_ms": round(ms_train, 3), "train_clips_per_s": round(B / ms_train * 1e3, 1),
            "infer_ms": round(ms_inf, 3), "infer_clips_per_s": round(B / ms_inf * 1e3, 1),
            "attn_share_of_flops": round(fl["attn"] / fl["total"], 3),
            "train_tflops": round(3 * fl["total"] * B / ms_train / 1e9, 1), "infer_tflops": round(fl["total"] * B / ms_inf / 1e9, 1),
            "train_frac_of_bf16_peak": round(3 * fl["total"] * B / ms_train / 1e9 / peak, 3),
            "infer_frac_of_bf16_peak": round(fl["total"] * B / ms_inf / 1e9 / peak, 3), "dropout": "on (train)", "dtype": "bf16"}
    print(json.dumps(line), flush=True)
    del model, step, gdata, glabels
    ops.clear_weight_cache()
    torch.cuda.empty_cache()
