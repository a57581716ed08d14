# the whole GPU suite, smoke(), the default bench line and the reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; echo "ref rc=$?"
cut -c1-600 gpurun_out/bench_r2_reference.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_default.json'))
print('value',d['value'],'ms',d['ms_per_step'],'R',d['repeats'],'e2e',d['e2e']['value'],'pll',d['pll_eval'], d['loss_after'], d['config']['achieved_tflops'], d['cpu_baseline'], d['clocks'])
print('cfg2', d['cfg2']); print('vq', d['vq_assign']); print('hbm', d['hbm_stages'])
print({k:v for k,v in d['roofline'].items() if k!='kernels'})
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
