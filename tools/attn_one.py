#!/usr/bin/env python
"""One attention forward+backward on a small case (for compute-sanitizer / debugging)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import ops
B, T, H, dh = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 50, 8, 96))]
p = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
d = H * dh
qkv = torch.randn(B, T, 3 * d, device="cuda", dtype=torch.bfloat16, requires_grad=True)
go = torch.randn(B, T, d, device="cuda", dtype=torch.bfloat16)
with mar.precision("bf16"):
    out = ops.attention(qkv, None, H, p)
    torch.cuda.synchronize(); print("fwd ok")
    out.backward(go)
    torch.cuda.synchronize(); print("bwd ok", float(qkv.grad.float().abs().mean()))
