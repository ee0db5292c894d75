# What the round-end driver runs on the GPU box, in one script (use with: gpurun --timeout 1800 -- 'bash run_gpu_tests.sh')
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench.json
