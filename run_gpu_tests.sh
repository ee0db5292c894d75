mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -q -m gpu --tb=short --maxfail=30 > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 12 gpurun_out/t_gpu.log; tail -n 3 gpurun_out/smoke.log; cut -c1-1800 gpurun_out/bench.json
