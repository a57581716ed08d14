# exchange / compute interference experiments at N GPUs: bash scripts/gpu_r2_i.sh N "ENV1" "ENV2" ...
N=$1; shift
mkdir -p gpurun_out
i=0
for E in "$@"; do
i=$((i+1))
env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $N --steps 5 --no-cpu-baseline --no-microbench --no-secondary --no-dp-parity > gpurun_out/bench_i_$i.json 2> gpurun_out/bench_i_$i.err; echo "== $E rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_i_$i.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),d['clocks']['sm_mhz'])
print(' '.join(f"{k['name'].replace('dense_','').replace('_bf16','')}={k['ms_per_step']:.2f}" for k in d['roofline']['kernels'][:7]))
PY
done
