# What the round-end driver runs on the GPU box, in one script (use with: gpurun --timeout 1800 -- 'bash run_gpu_tests.sh')
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/t_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -n 3 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench.json
# Two-GPU checks of the data-parallel step (not part of the driver's 1-GPU run):
#   gpurun --gpus 2 --timeout 300 -- 'python -m pytest tests/test_z_data_parallel_gpu.py -q -m gpu'
#   gpurun --gpus 2 --timeout 120 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_diag.py'
#   (every "N of 56 tensors differ" line must read 0 — ranks hold bit-identical gradients and parameters after each step)
