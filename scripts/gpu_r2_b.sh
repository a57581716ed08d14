# bf16 tests (operators + model) and a short cfg3 bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -s > gpurun_out/pytest_bf16_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16_all.log
grep -v "^$" gpurun_out/pytest_bf16_all.log | grep -v "^\.bf16 G=\|^bf16 G=" | tail -30
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_cfg3_quick.json 2> gpurun_out/bench_cfg3_quick.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_cfg3_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cfg3_quick.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'], d['config']['achieved_tflops'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
