# model-level bf16 tests, experimental sorted scatter test, default bench (cfg3, bf16)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -s -k "model or multi_group or three_steps" > gpurun_out/pytest_bf16_model.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16_model.log
grep -v "^$" gpurun_out/pytest_bf16_model.log | tail -30
PGMVAE_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s -k "sorted or scatter" > gpurun_out/pytest_sorted.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sorted.log
tail -8 gpurun_out/pytest_sorted.log
timeout 900 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r2_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_default.json'))
print('value',d['value'],'ms',d['ms_per_step'],'R',d['repeats'],'e2e',d['e2e']['value'],'pll',d['pll_eval'], d['loss_after'], d['config']['achieved_tflops'], d['cpu_baseline'])
print('cfg2', d['cfg2']); print('vq', d['vq_assign']); print('hbm', d['hbm_stages'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
