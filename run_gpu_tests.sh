mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu launches rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_tcgen05 -s 300 -c 98 --csv --log-file gpurun_out/gemm_dram.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu gemm dram rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 320 -c 2 -o gpurun_out/prof_gemm_step -f python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "ncu gemm full rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/attn_bench.py --p 0.1 > gpurun_out/attn_bench.jsonl 2> gpurun_out/attn_bench.err
timeout 300 python tools/gru_bench.py > gpurun_out/gru_bench.jsonl 2> gpurun_out/gru_bench.err
C2_HEADS=LSTM_1L,GRU_1L,Avg_features timeout 300 python tools/gru_bench.py >> gpurun_out/gru_bench.jsonl 2>> gpurun_out/gru_bench.err
timeout 300 python tools/gemm_bench.py --mode wgrad --shapes qkv,ffn1,ffn2,out > gpurun_out/gemm_wgrad.jsonl 2>&1
cat gpurun_out/summary.txt; cat gpurun_out/bench.json | cut -c1-400
