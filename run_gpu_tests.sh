mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short --maxfail=15 -k attention > gpurun_out/t_attn.log 2>&1; echo "pytest attention rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --tb=short --maxfail=15 -s > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2> gpurun_out/bench_nograph.err; echo "bench nograph rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/t_attn.log; grep -n "^FAILED\|passed\|failed\|whole-gradient" gpurun_out/t_gpu.log | tail -n 25; tail -n 3 gpurun_out/smoke.log; python -c "
import json
for f in ('bench','bench_nograph'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['roofline'] and (d['roofline']['achieved'], d['roofline']['gemm_ms_per_step']), d['clocks'])
"
