# Data-parallel A/B matrix: NG=<ranks> VARIANTS="name:ENV=VAL,ENV=VAL name2:..." bash tools/dp_matrix.sh
# (every variant is one bench.py run under torchrun; prints ms/step device-resident, end-to-end, ranks_identical, SM clock)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 5 --no-extras > gpurun_out/r02_dpm_${NG}_$name.json 2> gpurun_out/r02_dpm_${NG}_$name.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_dpm_${NG}_$name.json').read().strip().splitlines()[-1]);print('$name', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['ranks_identical'], d['clocks']['sm_mhz'], round(d['value']))"; }
for v in ${VARIANTS:-default:X=1 nocomm:MAR_DP_NOCOMM=1 b1:MAR_BUCKETS=1}; do
  name=${v%%:*}; envs=${v#*:}
  run $name $(echo $envs | tr ',' ' ')
done
