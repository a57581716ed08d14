set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "chain" > gpurun_out/pytest_chain.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_chain.log
tail -5 gpurun_out/pytest_chain.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
PGMVAE_VQ_SUB=3 timeout 300 python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 5 | cut -c1-250
