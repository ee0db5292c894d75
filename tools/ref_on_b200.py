"""The survey's kernel bar: the UNMODIFIED reference modules on the same B200
under the container's torch 2.11 — every device op a PyTorch library call (cuBLAS, cuDNN / flash SDPA, cuDNN RNN, ATen
elementwise) — timed beside this repo's kernels in the same process:

    python tools/ref_on_b200.py [--quick] > gpurun_out/ref_on_b200.jsonl

* C3 fusion train step (B = 256; zero_grad, forward, criterion, the reference's per-head backward, torch.optim.Adam):
  fp32 with TF32 off (the parity configuration), fp32 with TF32 on, bf16 autocast (the speed configuration);
* C2 video-GRU train step (B = 64, T = 64, cuDNN GRU), C5 fused-length sweep in bf16 autocast;
* F.scaled_dot_product_attention forward / forward+backward at the step's attention shapes with dropout 0.1, next to
  this repo's tcgen05 attention; nn.GRU (cuDNN persistent RNN) next to the persistent cluster GRU.

The reference classes come from oracle/_ref/models.py (a git-ignored copy made by oracle/build_ref.py in the build
container; /root/reference does not exist on the GPU box).  One JSON object per line."""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multimodalaggressionrecognition_b200 as mar  # noqa: E402
from multimodalaggressionrecognition_b200 import models as M, ops, training, workloads as W  # noqa: E402

DEV = torch.device("cuda", 0)


def load_reference():
    path = os.path.join(ROOT, "oracle", "_ref", "models.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("reference_models", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def timeit(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def emit(**kw):
    print(json.dumps(kw), flush=True)


def ref_train_step_fn(ref, model, crit, data, labels, autocast):
    opt = torch.optim.Adam(model.parameters())

    def step():
        opt.zero_grad()                                   # trainer.py:140
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            pred = model(data)                            # trainer.py:142
            losses = crit(pred, labels)                   # trainer.py:144
        losses.backward()                                 # trainer.py:147 (LossesDict: one backward per head)
        opt.step()                                        # trainer.py:149
    return step


def ours_train_step_fn(model, crit, data, labels, graph):
    step = training.TrainStep(model, crit, graph=graph, precision="bf16")
    return (lambda: step(data, labels)), step


def c3(ref, B, ta, tv, steps, modes):
    flops = 3 * W.c3_flops_per_clip(ta, tv)["total"] * B
    data, labels = W.batch_c3(B=B, t_audio=ta, t_video=tv)
    data, labels = W.to_device(data, DEV), W.to_device(labels, DEV)
    for mode in modes:
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        torch.manual_seed(0)
        model = W.build_c3(ref, ta, tv).to(DEV).train()
        crit = ref.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
        try:
            ms = timeit(ref_train_step_fn(ref, model, crit, data, labels, mode == "bf16_autocast"), steps)
            emit(what="c3_train_step", impl="reference modules, torch eager", mode=mode, B=B, t_audio=ta, t_video=tv, ms=ms,
                 clips_per_s=B / ms * 1e3, model_tflops=flops / ms / 1e9)
        except torch.OutOfMemoryError as e:               # (not expected on 180 GB)
            emit(what="c3_train_step", impl="reference modules, torch eager", mode=mode, B=B, t_audio=ta, t_video=tv, error=str(e)[:100])
        del model
        torch.cuda.empty_cache()
    torch.manual_seed(0)
    model = W.build_c3(M, ta, tv).to(DEV).train()
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
    fn, step = ours_train_step_fn(model, crit, data, labels, True)
    ms = timeit(fn, steps, warmup=8)
    step.release_graphs()
    emit(what="c3_train_step", impl="this repo (TrainStep, CUDA graph)", mode="bf16", B=B, t_audio=ta, t_video=tv, ms=ms,
         clips_per_s=B / ms * 1e3, model_tflops=flops / ms / 1e9)


def c2(ref, steps):
    x, y = W.batch_c2(B=64, T=64)
    x, y = x.to(DEV), y.to(DEV)
    for mode in ("fp32", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.manual_seed(0)
        model = W.build_c2(ref).to(DEV).train()
        crit = ref.MultiCrossEntropyLoss()
        ms = timeit(ref_train_step_fn(ref, model, crit, x, y, mode == "bf16_autocast"), steps * 3)
        emit(what="c2_train_step", impl="reference modules, torch eager (cuDNN GRU)", mode=mode, B=64, T=64, ms=ms, clips_per_s=64 / ms * 1e3)
    torch.manual_seed(0)
    model = W.build_c2(M).to(DEV).train()
    for graph in (False, True):
        fn, step = ours_train_step_fn(model, M.MultiCrossEntropyLoss(), x, y, graph)
        ms = timeit(fn, steps * 3, warmup=8)
        step.release_graphs()
        emit(what="c2_train_step", impl=f"this repo (TrainStep, graph={graph})", mode="bf16", B=64, T=64, ms=ms, clips_per_s=64 / ms * 1e3)


def attention(steps):
    H, dh, p = 8, 96, 0.1
    for (B, T) in ((256, 64), (256, 250), (256, 314), (32, 1024), (8, 4096)):
        flops_f = 4.0 * B * H * T * T * dh
        q, k, v = [torch.randn(B, H, T, dh, device=DEV, dtype=torch.bfloat16, requires_grad=True) for _ in range(3)]
        go = torch.randn(B, H, T, dh, device=DEV, dtype=torch.bfloat16)
        for backend in ("default", "cudnn", "flash"):
            from torch.nn.attention import SDPBackend, sdpa_kernel
            ctx = {"default": None, "cudnn": SDPBackend.CUDNN_ATTENTION, "flash": SDPBackend.FLASH_ATTENTION}[backend]

            def fwd():
                if ctx is None:
                    return F.scaled_dot_product_attention(q, k, v, dropout_p=p)
                with sdpa_kernel(ctx):
                    return F.scaled_dot_product_attention(q, k, v, dropout_p=p)

            def fwdbwd():
                o = fwd()
                o.backward(go)
            try:
                with torch.no_grad():
                    ms_f = timeit(fwd, steps)
                ms_fb = timeit(fwdbwd, steps)
                emit(what="attention", impl=f"torch SDPA ({backend})", B=B, T=T, H=H, dh=dh, p=p, fwd_ms=ms_f, bwd_ms=ms_fb - ms_f,
                     fwd_tflops=flops_f / ms_f / 1e9, bwd_tflops=2.5 * flops_f / max(ms_fb - ms_f, 1e-6) / 1e9)
            except Exception as e:                         # a backend that does not take this shape / dropout
                emit(what="attention", impl=f"torch SDPA ({backend})", B=B, T=T, H=H, dh=dh, p=p, error=str(e).splitlines()[0][:160])
        qkv = torch.randn(B, T, 3 * H * dh, device=DEV, dtype=torch.bfloat16, requires_grad=True)
        go2 = torch.randn(B, T, H * dh, device=DEV, dtype=torch.bfloat16)
        with mar.precision("bf16"):
            def ofwd():
                return ops.attention(qkv, None, H, p)

            def ofwdbwd():
                ofwd().backward(go2)
            with torch.no_grad():
                ms_f = timeit(ofwd, steps)
            ms_fb = timeit(ofwdbwd, steps)
        emit(what="attention", impl="this repo (tcgen05)", B=B, T=T, H=H, dh=dh, p=p, fwd_ms=ms_f, bwd_ms=ms_fb - ms_f,
             fwd_tflops=flops_f / ms_f / 1e9, bwd_tflops=2.5 * flops_f / max(ms_fb - ms_f, 1e-6) / 1e9)


def gru(steps):
    B, T, Hh = 64, 64, 512
    x = torch.randn(B, T, Hh, device=DEV)
    for mode in ("fp32", "bf16"):
        net = torch.nn.GRU(Hh, Hh, 1, batch_first=True).to(DEV)
        if mode == "bf16":
            net = net.to(torch.bfloat16)
        xi = x.to(torch.bfloat16 if mode == "bf16" else torch.float32).requires_grad_(True)

        def fwdbwd():
            out, _ = net(xi)
            out[:, -1].sum().backward()
        with torch.no_grad():
            ms_f = timeit(lambda: net(xi), steps * 3)
        ms_fb = timeit(fwdbwd, steps * 3)
        emit(what="gru_layer", impl="torch nn.GRU (cuDNN)", mode=mode, B=B, T=T, H=Hh, fwd_ms=ms_f, fwd_bwd_ms=ms_fb)
    net = torch.nn.GRU(Hh, Hh, 1, batch_first=True).to(DEV)
    xi = x.clone().requires_grad_(True)
    with mar.precision("bf16"):
        def ofwd():
            return ops.gru(xi, net.weight_ih_l0, net.weight_hh_l0, net.bias_ih_l0, net.bias_hh_l0)

        def ofwdbwd():
            ofwd()[:, -1].float().sum().backward()
        with torch.no_grad():
            ms_f = timeit(ofwd, steps * 3)
        ms_fb = timeit(ofwdbwd, steps * 3)
    emit(what="gru_layer", impl="this repo (input GEMM + persistent cluster recurrence)", mode="bf16", B=B, T=T, H=Hh, fwd_ms=ms_f, fwd_bwd_ms=ms_fb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    steps = 5 if args.quick else 10
    only = set(args.only.split(",")) if args.only else None
    ref = load_reference()
    emit(what="env", torch=torch.__version__, gpu=torch.cuda.get_device_name(0), reference_modules=ref is not None,
         cudnn=torch.backends.cudnn.version())
    if only is None or "attention" in only:
        attention(steps)
    if only is None or "gru" in only:
        gru(steps)
    if ref is None:
        emit(what="note", text="oracle/_ref/models.py absent: model-level reference timings skipped (python oracle/build_ref.py in the build container)")
        return
    if only is None or "c3" in only:
        c3(ref, 256, 250, 64, steps, ("fp32", "tf32", "bf16_autocast"))
    if only is None or "c2" in only:
        c2(ref, steps)
    if (only is None and not args.quick) or (only and "c5" in only):
        for T, B in ((128, 256), (512, 64), (2048, 8)):
            c3(ref, B, T, T, max(3, steps // 2), ("bf16_autocast",))


if __name__ == "__main__":
    main()
