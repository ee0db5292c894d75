#!/usr/bin/env python
"""Attention micro-benchmark at the C3 shapes: forward and backward, tcgen05 engine vs the mma.sync engine.
usage: python tools/attn_bench.py [--p 0.1] [--reps 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--p", type=float, default=0.1)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--T", type=int, nargs="*", default=[64, 250, 314, 1024])
ap.add_argument("--engines", nargs="*", default=["tc", "mma"])
args = ap.parse_args()
dev = torch.device("cuda:0")
H, dh = 8, 96
d = H * dh
res = {}
for T in args.T:
    B = args.B if T <= 314 else max(1, args.B * 250 // T // 2)
    qkv = torch.randn(B, T, 3 * d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    go = torch.randn(B, T, d, device=dev, dtype=torch.bfloat16)
    flops_f = 4.0 * T * T * dh * B * H
    for eng in args.engines:
        os.environ["MAR_ATTN_MMA"] = "1" if eng == "mma" else "0"
        with mar.precision("bf16"):
            outs = None
            for it in range(3):
                out = ops.attention(qkv, None, H, args.p)
                out.backward(go)
                qkv.grad = None
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            tf = tb = 0.0
            for it in range(args.reps):
                e[0].record()
                out = ops.attention(qkv, None, H, args.p)
                e[1].record()
                out.backward(go)
                e[2].record()
                torch.cuda.synchronize()
                tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
                qkv.grad = None
            tf /= args.reps; tb /= args.reps
        res[f"T{T}_{eng}"] = {"B": B, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
                              "fwd_tflops": round(flops_f / tf / 1e9, 1), "bwd_tflops": round(2.5 * flops_f / tb / 1e9, 1)}
print(json.dumps({"p_drop": args.p, "results": res}))
