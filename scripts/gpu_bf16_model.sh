# bf16 model-level tests + a short cfg3 bench in bf16
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -s -k "model or multi_group or three_steps" > gpurun_out/pytest_bf16_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16_model.log
tail -40 gpurun_out/pytest_bf16_model.log
timeout 600 python bench.py --workload cfg3 --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/bench_cfg3_bf16.json 2> gpurun_out/bench_cfg3_bf16.err
tail -5 gpurun_out/bench_cfg3_bf16.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cfg3_bf16.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'], d['config']['achieved_tflops'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
