#!/usr/bin/env python
"""Achieved HBM bandwidth of the step's elementwise / reduction kernels, each timed alone through the C ABI with CUDA
events (L2 flushed between launches), against MEASURED_PEAKS.json hbm_gbs:

    python tools/hbm_bench.py > gpurun_out/hbm_bench.jsonl

bytes = ALGORITHMIC bytes (every tensor the kernel must read or write once)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalaggressionrecognition_b200 import _lib

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6549.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rng = torch.zeros(2, dtype=torch.int64, device=dev)


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, M, N, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"kernel": name, "rows": M, "cols": N, "us": round(ms * 1e3, 1), "GB_s": round(gbs, 1), "frac_of_measured_copy_peak": round(gbs / peak, 3)}), flush=True)


for M in (80384, 64000, 16384):
    for N, flags, what in ((768, 2, "dropout backward + bias sum"), (2304, 0, "bias sum only"), (2048, 0, "bias sum only"), (768, 6, "dropout+ReLU backward + bias sum")):
        dout = torch.randn(M, N, device=dev).bfloat16()
        out = torch.randn(M, N, device=dev).bfloat16()
        dz = torch.empty_like(dout)
        db = torch.zeros(N, device=dev)
        relu = bool(flags & 5)
        fn = lambda: _lib.call("mar_linear_bwd_epilogue", dout.data_ptr(), out.data_ptr() if relu else None, dz.data_ptr() if flags else None,
                               db.data_ptr(), M, N, 1, 1, flags, 0.1, rng.data_ptr(), 3, 0, st)
        nbytes = M * N * 2 * ((1 if not flags else 2) + (1 if relu else 0))
        report(f"bwd_epilogue ({what})", M, N, nbytes, timeit(fn))
    D = 768
    x = torch.randn(M, D, device=dev).bfloat16(); y = torch.empty_like(x); dy = torch.randn_like(x); dx = torch.empty_like(x)
    g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev); dg = torch.zeros(D, device=dev); dbt = torch.zeros(D, device=dev)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    f = lambda: _lib.call("mar_layernorm_fwd", x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), None, M, D, 1e-5, 1, st)
    report("layernorm_fwd", M, D, M * D * 2 * 2, timeit(f))
    f = lambda: _lib.call("mar_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), g.data_ptr(), dx.data_ptr(), dg.data_ptr(), dbt.data_ptr(), M, D, 1, st)
    report("layernorm_bwd", M, D, M * D * 2 * 3, timeit(f))
n = 19_700_000 // 64 * 64
a, c = torch.randn(n, device=dev), torch.empty(n, device=dev, dtype=torch.bfloat16)
report("cast fp32->bf16", n, 1, n * 6, timeit(lambda: _lib.call("mar_cast", a.data_ptr(), 0, c.data_ptr(), 1, n, st)))
