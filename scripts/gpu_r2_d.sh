# bf16 tests incl. forced 2-CTA pair schedules; cfg3 bench with pairs (auto) vs without
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q > gpurun_out/pytest_bf16_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16_all.log
grep -v "^$" gpurun_out/pytest_bf16_all.log | tail -25
for PAIR in auto 0; do
if [ $PAIR = auto ]; then unset PGMVAE_BF16_PAIR; else export PGMVAE_BF16_PAIR=$PAIR; fi
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_cfg3_pair$PAIR.json 2> gpurun_out/bench_cfg3_pair$PAIR.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_cfg3_pair$PAIR.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg3_pair$PAIR.json'))
print('PAIR $PAIR value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'], d['config']['achieved_tflops'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
done
