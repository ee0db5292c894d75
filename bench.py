#!/usr/bin/env python
"""bench.py — train clips/sec of the multimodal Transformer fusion step (BASELINE.json configs[2]/[3]).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo, 1 B200
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # data parallel, one rank per GPU (NCCL)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle port)

A "step" is one pass of the hot path over one batch of synthetic input: forward of the C3 fusion model
(audio 250x768 + video 64x512 per clip, PhysVerbModel assembly of train_multimodal.py:298-420) → two-head
cross-entropy → backward → gradient all-reduce (N > 1) → Adam.  Per-GPU batch is 256 (weak scaling; N = 4
is BASELINE config 4's global batch of 1024).  bf16 compute, fp32 master weights, dropout ON (training
mode, p as the reference's modules define them).

One JSON line on stdout (rank 0).  `value` = clips/s with inputs resident in HBM; `e2e` = the same step
through the public API (`TrainStep`) from pinned HOST buffers, H2D copies and a D2H loss read inside the
timed region.  `roofline` = the tcgen05 GEMM kernel (≥ 93 % of the step's FLOPs), achieved TFLOP/s from CUDA
events around every GEMM launch of one full step, against MEASURED_PEAKS.json.  `cpu_baseline` = the oracle
port of the reference's path on this box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train clips/sec, multimodal transformer fusion"
UNIT = "clips/s"
PER_GPU_BATCH = 256
T_AUDIO, T_VIDEO = 250, 64


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe), running during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception as e:  # nvidia-smi absent: report that instead of inventing clocks
            log("clock sampler unavailable:", e)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the reference arm / cpu baseline: oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def load_reference_models():
    """oracle/_ref/models.py: the reference's own module, copied verbatim by oracle/build_ref.py in the build container
    (git-ignored, travels to the GPU box).  None when absent."""
    path = os.path.join(ROOT, "oracle", "_ref", "models.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_models", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_clips_per_s(steps: int, warmup: int, budget_s: float):
    """The reference's CPU path on this box's host cores, all threads, on a bounded sample of the C3 workload.

    kind "reference": the LIVE reference classes (oracle/_ref/models.py: PhysVerbModel assembly of
    train_multimodal.py:298-420, MultiModalCrossEntropyLoss, per-head backward, torch.optim.Adam — the step of
    trainer.py:140-149), train mode (dropout on), 32 clips/step.
    kind "port": the oracle restatement on 8 clips/step, when oracle/_ref is absent."""
    from multimodalaggressionrecognition_b200 import models as M, workloads as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = load_reference_models()
    torch.manual_seed(0)
    if ref is not None:
        kind, batch = "reference", 32
        model = W.build_c3(ref, T_AUDIO, T_VIDEO).train()
        crit = ref.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
        opt = torch.optim.Adam(model.parameters())
        data, labels = W.batch_c3(B=batch, t_audio=T_AUDIO, t_video=T_VIDEO)

        def one():
            opt.zero_grad()
            losses = crit(model(data), labels)
            losses.backward()
            opt.step()
    else:
        from oracle import oracle as O
        kind, batch = "port", 8
        O.DROPOUT_ENABLED = True
        model = W.build_c3(M, T_AUDIO, T_VIDEO)          # parameter containers only (torch init), never run here
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        cfg = W.c3_oracle_cfg(T_AUDIO, T_VIDEO)
        data, labels = W.batch_c3(B=batch, t_audio=T_AUDIO, t_video=T_VIDEO)
        tr = O.OracleTrainer(sd, lambda s, d, t: O.physverb_model(d, s, cfg, t, True),
                             lambda p, t: O.multimodal_ce(p, t, heads=["phys", "verb"]))

        def one():
            tr.step(data, labels, True)
    t0 = time.perf_counter()
    one()
    first = time.perf_counter() - t0
    # bound the run: shrink the number of timed steps (never below 2) to stay inside the budget
    steps = max(2, min(steps, int(budget_s / max(first, 1e-3)) - warmup))
    for _ in range(max(0, warmup - 1)):
        one()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    what = "live reference classes (oracle/_ref/models.py) + torch.optim.Adam" if kind == "reference" else "oracle port"
    return {"value": batch / (ms / 1e3), "ms_per_step": ms, "steps": steps, "cores": cores, "batch": batch, "kind": kind,
            "sample": f"C3 train step on {batch} clips/step (full config 256), {what}, {steps} timed steps, dropout on, "
                      f"{cores} host threads, torch {torch.__version__} CPU fp32"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_clips_per_s(args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C3 audio+video transformer fusion train step, T_a={T_AUDIO}x768, T_v={T_VIDEO}x512, "
                               f"CPU sample of {r['batch']} clips/step"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import multimodalaggressionrecognition_b200 as mar
    from multimodalaggressionrecognition_b200 import models as M, ops, training, workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NUMA-local pinned staging for the end-to-end arm (MAR_NUMA_BIND=0: leave the affinity alone, for A/B runs)
    numa_cpus = training.bind_to_gpu_numa_node(local) if os.environ.get("MAR_NUMA_BIND", "1") != "0" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    mar.set_precision(args.mode)
    torch.manual_seed(0)                                      # same initial weights on every rank
    model = W.build_c3(M, T_AUDIO, T_VIDEO).to(dev).train()
    crit = M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
    B = args.batch
    data, labels = W.batch_c3(B=B, t_audio=T_AUDIO, t_video=T_VIDEO, seed=1000 + rank)
    use_graph = not args.no_graph           # world > 1: the NCCL all-reduces are captured into the step graph too
    # data-parallel knobs for A/B measurements (defaults = TrainStep's): MAR_WIRE=fp32|bf16, MAR_BUCKETS=n, MAR_TAIL=elements
    dp_kw = {}
    if os.environ.get("MAR_WIRE"):
        dp_kw["wire"] = os.environ["MAR_WIRE"]
    if os.environ.get("MAR_BUCKETS"):
        dp_kw["num_buckets"] = int(os.environ["MAR_BUCKETS"])
    if os.environ.get("MAR_TAIL"):
        dp_kw["tail_elems"] = int(os.environ["MAR_TAIL"])
    step = training.TrainStep(model, crit, graph=use_graph, **dp_kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(run_one, steps):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(steps):
            run_one()
        ev[1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[1])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    # ---- device-resident arm ------------------------------------------------------------------
    gdata, glabels = W.to_device(data, dev), W.to_device(labels, dev)
    out = {}
    def dev_step():
        out["l"] = step(gdata, glabels)
    # initialisation (not warm-up): the step driver runs 3 eager steps, then captures one CUDA graph per input-buffer set
    setup_steps = 6 if use_graph else 0
    for _ in range(setup_steps):
        dev_step()
    for _ in range(max(args.warmup, 3)):                      # the W untimed warm-up steps of the contract (W >= 3)
        dev_step()
    torch.cuda.synchronize()
    ops.reset_launch_count()
    dev_step()
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count()
    if use_graph and launches_per_step == 0:
        # a graph replay does not pass through the C ABI: count the launches recorded at capture time
        launches_per_step = step.captured_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(dev_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    loss_vals = {k: float(v) for k, v in out["l"].items()}

    # ---- end-to-end arm: pinned host inputs, H2D + D2H inside the timed region -------------------
    host_t = [t.pin_memory() for t in training.TrainStep._tensors([data, labels])]
    hdata, hlabels = training.TrainStep._like([data, labels], host_t)
    h2d = sum(t.numel() * t.element_size() for t in host_t)
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    def e2e_step():
        l = step(hdata, hlabels)
        loss_host.copy_(torch.stack([l["phys"], l["verb"]]), non_blocking=True)   # D2H of the step's result
    for _ in range(4):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    # ---- the other BASELINE.json configurations, time-bounded, inside the same line ---------------------------
    other = None
    if not args.no_extras:
        other = other_configs(args, world, rank, dev, timed, M, ops, training, W)

    # ---- roofline of the dominant kernel: events around every GEMM launch of one eager step --------
    roof = None
    if rank == 0:
        step.sync.enabled = False          # this extra backward runs on rank 0 only: no collectives
        side = step._side if step._side is not None else torch.cuda.current_stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # the stream the parameters' AccumulateGrad nodes live on
            roof = gemm_roofline(model, crit, gdata, glabels, args, ops, training)
        torch.cuda.current_stream().wait_stream(side)

    def shutdown():
        # graphs that captured NCCL collectives must go before the communicator; a watchdog makes sure a stuck
        # teardown can never keep the job (and the other ranks) alive after the result line is out
        sys.stdout.flush()
        t = threading.Timer(45.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        step.release_graphs()
        if world > 1:
            dist.destroy_process_group()

    # ---- data-parallel proof: every rank holds bit-identical parameters after the timed steps --------------------
    ranks_identical = None
    if world > 1:
        bits = step.flat.flat.view(torch.int32)
        digest = torch.stack([bits.sum(dtype=torch.int64), (bits[::7].to(torch.int64) * 31).sum(), bits.to(torch.int64).abs().max()])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        ranks_identical = all(torch.equal(gathered[0], g_) for g_ in gathered[1:])
        if step.sync.peer is not None:
            step.sync.peer.check()          # raises if a wait on a peer inside the fused exchange kernel ever timed out
        dist.barrier()
    if rank != 0:
        shutdown()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_clips_per_s(steps=3, warmup=1, budget_s=25.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    torch_b200 = None
    if world == 1 and not args.no_extras:
        torch_b200 = torch_on_b200(dev, W, gdata, glabels)

    global_batch = B * world
    line = {
        "metric": METRIC, "value": global_batch / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"C3 audio+video transformer fusion train step (fwd+loss+bwd+allreduce+Adam), "
                               f"T_a={T_AUDIO}x768, T_v={T_VIDEO}x512, d=768, 8 heads, d_ff=2048, 19.7M params, dropout on",
                   "global_batch": global_batch, "per_gpu_batch": B, "parallelism": f"dp{world}",
                   "cuda_graph": bool(use_graph), "setup_steps_before_warmup": setup_steps,
                   "host_cpus_bound_to_gpu_numa_node": None if numa_cpus is None else len(numa_cpus),
                   "grad_exchange": None if world == 1 else {"wire": step.sync.wire, "buckets": len(step.sync.buckets),
                                                              "engine": "one kernel over NVLink peer memory: bf16 reduce-scatter (peer loads) + all-gather (peer stores) + Adam"
                                                              if step.sync.fused() else "ncclAllReduce + Adam kernel",
                                                              "bucket_elems": [e1 - e0 for _, _, e0, e1 in step.sync.buckets]},
                   "l2": "inputs (229 MB/step fp32) and activations (> 3 GB/step) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": global_batch / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "ranks_identical": ranks_identical,
        "other_configs": other,
        "torch_b200": torch_b200,
        "losses_last_step": loss_vals,
        "flops_per_clip_train": 3 * W.c3_flops_per_clip(T_AUDIO, T_VIDEO)["total"],
        "model_tflops": 3 * W.c3_flops_per_clip(T_AUDIO, T_VIDEO)["total"] * global_batch / (ms_dev / 1e3) / 1e12,
    }
    peaks = measured_peaks()
    line["model_frac_of_bf16_peak"] = line["model_tflops"] / (peaks["bf16_tflops_sustained"] * world)
    print(json.dumps(line), flush=True)
    shutdown()


def other_configs(args, world, rank, dev, timed, M, ops, training, W):
    """BASELINE.json configs[1], [3], [4] as sub-results of the same line (the headline stays configs[2], 256 clips/GPU):
    * c4_global1024 — the C3 model data-parallel at a GLOBAL batch of 1024 (1024 / N clips per GPU, strong scaling in N);
    * c2 — the video GRU classifier, B = 64, T = 64, d = 512 (N = 1 only: the config names one B200);
    * c5 — the fusion model at T_a = T_v = T in {128, 512, 2048} (fused length 2T: the attention-bound regime), 65 536
      audio tokens per GPU per step (weak scaling), train step and inference forward.
    Every entry: CUDA-event time over `--extra-steps` steps after warm-up, max over ranks, whole-job clips/s, and the
    model FLOP rate as a fraction of N x the measured sustained bf16 peak."""
    peaks = measured_peaks()
    peak = peaks["bf16_tflops_sustained"] * world
    K = args.extra_steps
    out = {}

    def crit_c3():
        return M.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})

    def fusion(ta, tv, B, infer):
        torch.manual_seed(0)
        model = W.build_c3(M, ta, tv).to(dev).train()
        data, labels = W.batch_c3(B=B, t_audio=ta, t_video=tv, seed=2000 + rank)
        gd, gl = W.to_device(data, dev), W.to_device(labels, dev)
        st = training.TrainStep(model, crit_c3(), graph=not args.no_graph)
        for _ in range(6):
            st(gd, gl)
        ms = timed(lambda: st(gd, gl), K)
        fl = W.c3_flops_per_clip(ta, tv)
        res = {"per_gpu_batch": B, "global_batch": B * world, "train_ms": ms, "train_clips_per_s": B * world / ms * 1e3,
               "train_frac_of_bf16_peak": 3 * fl["total"] * B * world / ms / 1e9 / peak,
               "attn_share_of_flops": fl["attn"] / fl["total"]}
        st.release_graphs()
        if infer:
            model.eval()

            def fwd():
                with torch.no_grad():
                    model(gd)
            for _ in range(3):
                fwd()
            ms_i = timed(fwd, K)
            res.update({"infer_ms": ms_i, "infer_clips_per_s": B * world / ms_i * 1e3,
                        "infer_frac_of_bf16_peak": fl["total"] * B * world / ms_i / 1e9 / peak})
        del st, model, gd, gl
        ops.clear_weight_cache()
        torch.cuda.empty_cache()
        return res

    if 1024 % world == 0:
        out["c4_global1024"] = fusion(T_AUDIO, T_VIDEO, 1024 // world, False)
    out["c5"] = {f"T{T}": fusion(T, T, max(1, 65536 // T), True) for T in (128, 512, 2048)}
    if world == 1:
        torch.manual_seed(0)
        model = W.build_c2(M).to(dev).train()
        x, y = W.batch_c2(64, 64, 512)
        x, y = x.to(dev), y.to(dev)
        st = training.TrainStep(model, M.MultiCrossEntropyLoss(), graph=not args.no_graph)
        for _ in range(8):
            st(x, y)
        ms = timed(lambda: st(x, y), 4 * K)
        fl = W.c2_flops_per_clip()
        out["c2"] = {"batch": 64, "T": 64, "d": 512, "head": "GRU_1L", "train_ms": ms, "train_clips_per_s": 64 / ms * 1e3,
                     "model_tflops": 3 * fl * 64 / ms / 1e9}
        st.release_graphs()
        with torch.no_grad():
            gi = torch.randn(64, 64, 3 * 512, device=dev, dtype=torch.bfloat16)
            gru = model.models_dict["GRU_1L"].sequence_nn
            with ops.precision("bf16"):
                rec = lambda: ops._GRU.apply(gi, gru.weight_hh_l0, gru.bias_hh_l0, False)
                for _ in range(3):
                    rec()
                ms_r = timed(rec, 4 * K)
        out["c2"]["recurrence_fwd_us_per_time_step"] = ms_r * 1e3 / 64
        del st, model
        ops.clear_weight_cache()
        torch.cuda.empty_cache()
    return out


def torch_on_b200(dev, W, data, labels, steps=5):
    """The survey's kernel bar, in the same run: the UNMODIFIED reference modules
    (oracle/_ref/models.py) on this B200 under torch eager — every device op a PyTorch library call (cuBLAS, cuDNN /
    flash SDPA, ATen) — on the same C3 batch: bf16 autocast and fp32 with TF32 off; plus F.scaled_dot_product_attention
    at the step's two large attention shapes.  tools/ref_on_b200.py has the full table (C2 / cuDNN GRU, C5 lengths)."""
    import torch.nn.functional as F
    ref = load_reference_models()
    out = {"torch": torch.__version__}

    def timeit(fn, n, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    if ref is not None:
        B = data[0][1].shape[0]
        for mode in ("bf16_autocast", "fp32"):
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            torch.manual_seed(0)
            model = W.build_c3(ref, T_AUDIO, T_VIDEO).to(dev).train()
            crit = ref.MultiModalCrossEntropyLoss({"phys": torch.nn.CrossEntropyLoss(), "verb": torch.nn.CrossEntropyLoss()})
            opt = torch.optim.Adam(model.parameters())

            def step():
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16_autocast"):
                    losses = crit(model(data), labels)
                losses.backward()
                opt.step()
            ms = timeit(step, steps if mode == "bf16_autocast" else 3)
            out[f"c3_step_reference_modules_{mode}"] = {"ms": ms, "clips_per_s": B / ms * 1e3}
            del model, opt
            torch.cuda.empty_cache()
    else:
        out["c3_step_reference_modules"] = "oracle/_ref/models.py absent (python oracle/build_ref.py in the build container)"
    for T in (T_AUDIO, T_AUDIO + T_VIDEO):
        q, k, v = [torch.randn(PER_GPU_BATCH, 8, T, 96, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3)]
        go = torch.randn_like(q)
        with torch.no_grad():
            f = timeit(lambda: F.scaled_dot_product_attention(q, k, v, dropout_p=0.1), 10)
        fb = timeit(lambda: F.scaled_dot_product_attention(q, k, v, dropout_p=0.1).backward(go), 10)
        out[f"sdpa_T{T}_B{PER_GPU_BATCH}_H8_dh96_p0.1"] = {"fwd_ms": f, "bwd_ms": fb - f}
    return out


def gemm_roofline(model, crit, gdata, glabels, args, ops, training):
    """Achieved TFLOP/s of the tcgen05 GEMM kernel: CUDA events around every mar_linear_{fwd,dgrad,wgrad} call of
    one full eager train step (on the launching stream), algorithmic FLOPs = 2·M·N·K per launch.  The GPU is kept busy
    (a spin kernel of ~25 ms) while the host enqueues the step, so that the launches execute back to back as they do
    inside the captured graph and an event pair brackets the kernel's execution, not the host's launch latency (an
    eager launch submitted to an idle GPU adds 3-8 us of host time between the two events: a third of a 20 us video-branch
    GEMM).  MAR_ROOFLINE_QUEUE=0 restores the unqueued measurement."""
    from multimodalaggressionrecognition_b200 import _lib
    recs = []
    orig = _lib.call
    gemm_names = {"mar_linear_fwd": (8, 9, 10), "mar_linear_dgrad": (9, 10, 11), "mar_linear_wgrad": (4, 5, 6)}

    def timed_call(name, *a):
        if name in gemm_names:
            i, j, k = gemm_names[name]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig(name, *a)
            e1.record()
            recs.append((name, 2.0 * a[i] * a[j] * a[k], e0, e1, _lib.load().mar_last_engine(), (a[i], a[j], a[k])))
        else:
            orig(name, *a)

    ops.call = timed_call
    training.call = timed_call
    try:
        for p in model.parameters():
            p.grad = None
        passes = []
        for it in range(4):                    # first pass warms; three recorded passes, per-launch MEDIAN (an eager
            recs.clear()                       # launch can sit behind a host hiccup that a graph replay never sees)
            if os.environ.get("MAR_ROOFLINE_QUEUE", "1") != "0":
                torch.cuda._sleep(48_000_000)                  # ~25 ms at 1.9 GHz: the host runs ahead of the GPU
            ops.rng_advance()
            losses = crit(model(gdata), glabels)
            losses.backward()
            torch.cuda.synchronize()
            if it > 0:
                passes.append([(nm, f, e0.elapsed_time(e1), eng, mnk) for (nm, f, e0, e1, eng, mnk) in recs])
    finally:
        ops.call = orig
        training.call = orig
    n = min(len(p_) for p_ in passes)
    med = []
    for i in range(n):
        nm, f, _, eng, mnk = passes[0][i]
        med.append((nm, f, statistics.median(p_[i][2] for p_ in passes), eng, mnk))
    if os.environ.get("MAR_GEMM_TABLE"):          # per-launch table on stderr (M, N, K as the C ABI names them)
        for (nm, f, ms_, eng, mnk) in med:
            log(f"{nm:18s} M={mnk[0]:6d} N={mnk[1]:5d} K={mnk[2]:5d} engine={eng} {ms_ * 1e3:8.1f} us {f / ms_ / 1e9:8.1f} TFLOP/s")
    tc = [(f, ms_) for (_, f, ms_, eng, _mnk) in med if eng == 2]
    if not tc:
        return None
    flops = sum(f for f, _ in tc)
    ms = sum(t for _, t in tc)
    peaks = measured_peaks()
    achieved = flops / (ms / 1e3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    traffic = None                     # DRAM bytes per launch of this kernel from the committed ncu capture of the same command
    for name in ("r02b_gemm_dram_traffic.json", "r02_gemm_dram_traffic.json", "r01_gemm_dram_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            break
    return {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, mean over the step's 49 launches)",
            "launches": len(tc), "gemm_ms_per_step": ms,
            "gemm_flops_per_step": flops,
            "timing": "CUDA events around every GEMM launch of an eager step queued behind a 25 ms spin kernel (back-to-back execution, "
                      "median of 3 steps)",
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); kernels timed inside a full step "
                           f"(CUDA events around every GEMM launch of an eager step, per-launch median of 3 steps)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="clips per GPU per step")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c2 / c4_global1024 / c5 sub-results")
    ap.add_argument("--extra-steps", type=int, default=5)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
