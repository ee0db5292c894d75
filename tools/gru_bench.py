#!/usr/bin/env python
"""C2 video GRU classifier (BASELINE.json configs[1]: B=64, T=64, d=512, bf16): recurrence kernel and whole train step,
persistent cluster engine vs the step-per-launch engine.   usage: python tools/gru_bench.py [--reps 20]"""
import argparse, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalaggressionrecognition_b200 as mar
from multimodalaggressionrecognition_b200 import _lib, models as M, ops, workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--T", type=int, default=64)
args = ap.parse_args()
dev = torch.device("cuda:0")
B, T, H = args.B, args.T, 512
res = {"B": B, "T": T, "H": H}

def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# ---- the recurrence alone through the C ABI
k = 1 / math.sqrt(H)
gi = torch.randn(B, T, 3 * H, device=dev).to(torch.bfloat16)
w = ((torch.rand(3 * H, H, device=dev) * 2 - 1) * k).to(torch.bfloat16)
bh = ((torch.rand(3 * H, device=dev) * 2 - 1) * k).float()
hseq = torch.empty(B, T, H, device=dev, dtype=torch.bfloat16)
hprev = torch.empty_like(hseq); saved = torch.empty(B, T, 5 * H, device=dev)
work = torch.empty(int(_lib.load().mar_gru_work_floats(B, T, H)), device=dev)
st = torch.cuda.current_stream().cuda_stream
for name, eng in (("persistent", _lib.ENGINE_TCGEN05), ("steps", _lib.ENGINE_SIMT)):
    f = lambda: _lib.call("mar_gru_fwd", gi.data_ptr(), w.data_ptr(), bh.data_ptr(), hseq.data_ptr(), hprev.data_ptr(),
                          saved.data_ptr(), work.data_ptr(), B, T, H, _lib.MAR_BF16, eng, st)
    ms = timeit(f, args.reps)
    if name == "persistent":
        f2 = lambda: _lib.call("mar_gru_fwd", gi.data_ptr(), w.data_ptr(), bh.data_ptr(), hseq.data_ptr(), None,
                               None, work.data_ptr(), B, T, H, _lib.MAR_BF16, eng, st)
        res["recurrence_infer_ms_persistent"] = round(timeit(f2, args.reps), 4)
    res[f"recurrence_fwd_ms_{name}"] = round(ms, 4)
    res[f"recurrence_us_per_step_{name}"] = round(ms * 1e3 / T, 3)
flops = T * 2.0 * B * 3 * H * H
res["recurrence_gflops_persistent"] = round(flops / res["recurrence_fwd_ms_persistent"] / 1e6, 1)

# ---- the C2 train step (fwd + CE + bwd) through the drop-in modules
ap2 = os.environ.get("C2_HEADS", "GRU_1L").split(",")
model = W.build_c2(M, heads=tuple(ap2)).to(dev)
res["heads"] = ap2
x, y = W.batch_c2(B, T, 512)
x, y = x.to(dev), y.to(dev)
crit = M.MultiCrossEntropyLoss() if hasattr(M, "MultiCrossEntropyLoss") else None
def step(engine):
    with mar.precision("bf16"), mar.engine(engine):
        model.zero_grad(set_to_none=True)
        out = model(x)
        loss = sum(torch.nn.functional.cross_entropy(v.float(), y) for v in out.values())
        loss.backward()
for name, engine in (("auto", "auto"), ("steps", "simt")):
    try:
        ms = timeit(lambda: step(engine), max(3, args.reps // 4))
        res[f"c2_train_step_ms_{name}"] = round(ms, 3)
        res[f"c2_clips_per_s_{name}"] = round(B / ms * 1e3, 1)
    except Exception as e:  # the simt engine in bf16 mode may refuse some ops; report instead of hiding
        res[f"c2_train_step_{name}_error"] = str(e)[:200]
# ---- the same step through the sync-free step driver (flat Adam, CUDA graph): what train_video_rnn.py would run
from multimodalaggressionrecognition_b200 import training
mar.set_precision("bf16")
step = training.TrainStep(model, M.MultiCrossEntropyLoss(), graph=True)
for _ in range(8):          # 3 eager warm-ups + one capture per input-buffer set happen before the timed region
    step(x, y)
ms = timeit(lambda: step(x, y), args.reps)
res["c2_trainstep_graph_ms"] = round(ms, 3)
res["c2_trainstep_graph_clips_per_s"] = round(B / ms * 1e3, 1)
print(json.dumps(res))
