# bf16 tests (operators + model) and short cfg3 benches at two group sizes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -s > gpurun_out/pytest_bf16_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16_all.log
grep -v "^$" gpurun_out/pytest_bf16_all.log | grep -v "^\.bf16 G=\|^bf16 G=\|^\.dense bf16\|^dense bf16" | tail -25
for GV in 296 148; do
PGMVAE_GROUP_VARS=$GV timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_cfg3_gv$GV.json 2> gpurun_out/bench_cfg3_gv$GV.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_cfg3_gv$GV.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_gv$GV.json'))
print('GV $GV value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'], d['config']['achieved_tflops'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
done
