mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py --mode wgrad --shapes qkv,ffn1,ffn2,out,qkv_a,emb 2>&1 | tail -1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "linear or gemm" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','model_frac_of_bf16_peak')}, d['roofline']['achieved'], d['roofline']['gemm_ms_per_step'], d['e2e']['value'])"
