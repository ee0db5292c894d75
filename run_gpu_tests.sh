mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short --maxfail=20 -k "attention" > gpurun_out/t_attn.log 2>&1; echo "pytest attn rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/attn_bench.py --p 0.1 > gpurun_out/attn_bench.jsonl 2> gpurun_out/attn_bench.err; echo "attn_bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/attn_bench.py --p 0.0 >> gpurun_out/attn_bench.jsonl 2>> gpurun_out/attn_bench.err
cat gpurun_out/summary.txt; tail -n 30 gpurun_out/t_attn.log; cat gpurun_out/attn_bench.jsonl; tail -n 5 gpurun_out/attn_bench.err
