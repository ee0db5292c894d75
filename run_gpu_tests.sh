mkdir -p gpurun_out
date +%s > /tmp/t0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$? elapsed $(( $(date +%s) - $(cat /tmp/t0) ))s"
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['config']['cuda_graph'])"
tail -3 gpurun_out/bench_n2.err
