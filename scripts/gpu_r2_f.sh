# cta_group::2 schedule: operator tests first (short timeout: a hang must not eat the budget), then everything, then A/B
set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -k "sigmoid" > gpurun_out/pytest_2sm_ops.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_2sm_ops.log
tail -15 gpurun_out/pytest_2sm_ops.log
if [ $rc -ne 0 ]; then echo "2sm operator tests failed: stopping"; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/pytest_bf16_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16_all.log
grep -v "^$" gpurun_out/pytest_bf16_all.log | tail -15
for MODE in quad; do
if [ $MODE = mc ]; then export PGMVAE_BF16_NO_QUAD=1; else unset PGMVAE_BF16_NO_QUAD; fi
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_cfg3_$MODE.json 2> gpurun_out/bench_cfg3_$MODE.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_cfg3_$MODE.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_$MODE.json'))
print('MODE $MODE value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'], d['config']['achieved_tflops'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
done
