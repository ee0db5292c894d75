mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -n 4 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
