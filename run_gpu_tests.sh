timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "gru or rnn" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_models_gpu.py -q -m gpu --tb=line -k "c2" 2>&1 | tail -3
for B in 32 64; do timeout 300 python tools/gru_bench.py --B $B 2>&1 | tail -1 | cut -c230-600; done
