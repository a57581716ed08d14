# default bench at N GPUs (with dp_parity); bash scripts/gpu_r2_scale.sh N [dpcheck]
set -x
N=$1
mkdir -p gpurun_out
if [ "$2" = dpcheck ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 pgm-vae_b200/tools/dp_check.py > gpurun_out/dp_check_n$N.log 2>&1; echo "dp_check rc=$?"
grep -v "^W\|OMP_NUM\|\*\*\*" gpurun_out/dp_check_n$N.log | tail -3
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2_n$N.json'))
print('N $N value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'],d['pll_eval']['variable_sharded_value'], d['clocks'])
print('dp_parity',d['dp_parity'])
PY
