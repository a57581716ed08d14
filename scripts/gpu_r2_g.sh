# epilogue rework (bias through shared memory, double-buffered output tiles): operator + model tests, bench, then
# the main-loop experiment: big GEMMs against ring depth and pairing schedule
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/pytest_g.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/pytest_g.log
grep -v "^$" gpurun_out/pytest_g.log | tail -15
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_g.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_g.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'], d['config']['achieved_tflops'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
for PAIR in "" 1 3; do for ST in 8 3 2; do
echo "== PAIR=$PAIR STAGES=$ST"
PGMVAE_BF16_PAIR=$PAIR PGMVAE_BF16_STAGES=$ST timeout 200 python pgm-vae_b200/tools/bf16_microbench.py 148 4096 1556x400 400x1556 2>>gpurun_out/mb_g.err | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['shape'],r['orient'],r['kernel'],round(r['ms'],4),round(r['TFLOPs'] or 0,1))
"
done; done
