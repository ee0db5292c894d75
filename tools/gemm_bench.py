"""Micro-benchmark of the tcgen05 GEMM through the C ABI (CUDA events, L2 flushed between launches).
    python tools/gemm_bench.py [--shapes qkv,ffn1,ffn2,out] [--iters 20] [--mode fwd|dgrad|wgrad]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalaggressionrecognition_b200 import _lib

SHAPES = {"qkv": (80384, 2304, 768), "ffn1": (80384, 2048, 768), "ffn2": (80384, 768, 2048), "out": (80384, 768, 768),
          "qkv_a": (64000, 2304, 768), "emb": (16384, 768, 512), "sq8k": (8192, 8192, 8192), "vid_out": (16384, 768, 768), "vid_qkv": (16384, 2304, 768)}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="qkv,ffn1,ffn2,out")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--mode", default="fwd")
    ap.add_argument("--epi", default="bias", help="fwd: none|bias|res (bias+residual)|full (bias+dropout+residual); "
                    "dgrad: none|res (+ fork gradient)|aux (activation mask)|auxsum (mask + column sums)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rng = torch.zeros(2, dtype=torch.int64, device=dev)
    out = {}
    for name in args.shapes.split(","):
        M, N, K = SHAPES[name]
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        wt = w.t().contiguous()
        b = torch.randn(N, device=dev)
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = torch.randn(M, N, device=dev).bfloat16() if args.epi in ("full", "res") else None
        dw = torch.empty(N, K, device=dev)
        xres = torch.randn(M, K, device=dev).bfloat16() if args.mode == "dgrad" else None
        csum = torch.zeros(K, device=dev)
        def run():
            if args.mode == "fwd":
                flags, p = (2, 0.1) if args.epi == "full" else (0, 0.0)
                _lib.call("mar_linear_fwd", x.data_ptr(), K, w.data_ptr(), None if args.epi == "none" else b.data_ptr(),
                          None if res is None else res.data_ptr(), N, y.data_ptr(), N, M, N, K, 1, 1, flags, p,
                          rng.data_ptr(), 1, 2, st)
            elif args.mode == "dgrad":
                _lib.call("mar_linear_dgrad", y.data_ptr(), w.data_ptr(), wt.data_ptr(), xres.data_ptr() if args.epi == "res" else None,
                          xres.data_ptr() if args.epi in ("aux", "auxsum") else None, 1.1, x.data_ptr(), K,
                          csum.data_ptr() if args.epi == "auxsum" else None, M, N, K, 1, 2, st)
            else:
                _lib.call("mar_linear_wgrad", y.data_ptr(), x.data_ptr(), K, dw.data_ptr(), M, N, K, 1, 0, 2, st)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]
        tf = 2.0 * M * N * K / (ms / 1e3) / 1e12
        out[name] = {"ms": round(ms, 4), "tflops": round(tf, 1), "best_tflops": round(2.0 * M * N * K / (times[0] / 1e3) / 1e12, 1)}
        # cuBLAS (library) beside it, same shape
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(x, w.t(), out=y) if args.mode == "fwd" else torch.matmul(x, w.t()); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out[name]["cublas_tflops"] = round(2.0 * M * N * K / (ts[len(ts) // 2] / 1e3) / 1e12, 1)
    print(json.dumps({"mode": args.mode, "epi": args.epi, "env": {k: v for k, v in os.environ.items() if k.startswith("MAR_")}, "results": out}))

if __name__ == "__main__":
    main()
