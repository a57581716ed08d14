set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "chain" > gpurun_out/pytest_chain.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_chain.log
tail -5 gpurun_out/pytest_chain.log
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for nch in 3 2; do
PGMVAE_CHAINS=$nch timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'])
for k in d['roofline']['kernels'][:4]: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
done
