mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/gemm_bench.jsonl
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short --maxfail=8 > gpurun_out/t_kernels.log 2>&1; echo "pytest kernels rc=$?" >> gpurun_out/summary.txt
for epi in none bias full; do timeout 120 python tools/gemm_bench.py --epi $epi >> gpurun_out/gemm_bench.jsonl 2>> gpurun_out/gemm_bench.err; done
timeout 120 python tools/gemm_bench.py --mode dgrad >> gpurun_out/gemm_bench.jsonl 2>> gpurun_out/gemm_bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_kernels.log; cat gpurun_out/gemm_bench.jsonl; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['roofline'] and (round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],2)), d['clocks'])
"
