mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 3 -c 1 -o gpurun_out/prof_attn_bwd3 -f python tools/attn_bench.py --T 250 --engines tc --reps 2 > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
