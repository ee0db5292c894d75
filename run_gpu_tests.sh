timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -x -k "gru or rnn or lstm" 2>&1 | tail -8
timeout 300 python -m pytest tests/test_models_gpu.py -q -m gpu --tb=line -k "c2" 2>&1 | tail -3
timeout 300 python tools/gru_bench.py --B 64 2>&1 | tail -1 | cut -c230-700
C2_HEADS=LSTM_1L,GRU_1L,Avg_features timeout 300 python tools/gru_bench.py --B 64 2>&1 | tail -1 | cut -c230-700
