# A/B of the SM reservation for NCCL at N GPUs: bash scripts/gpu_r2_dp_ab.sh N "8 0 16"
set -x
N=$1
mkdir -p gpurun_out
for R in $2; do
PGMVAE_COMM_SMS=$R timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --no-dp-parity > gpurun_out/bench_r2_n${N}_sms$R.json 2> gpurun_out/bench_r2_n${N}_sms$R.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_r2_n${N}_sms$R.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2_n${N}_sms$R.json'))
print('COMM_SMS $R N $N value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'],d['pll_eval']['variable_sharded_value'])
PY
done
