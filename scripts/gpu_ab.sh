# A/B of the step: GPU parity tests, then the quick bench with the merged / per-layer weight-gradient launches.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for mode in merged per_layer; do
if [ $mode = per_layer ]; then export PGMVAE_WGRAD_PER_LAYER=1; else unset PGMVAE_WGRAD_PER_LAYER; fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err
tail -2 gpurun_out/bench_$mode.err
python - $mode <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_%s.json'%sys.argv[1]))
print(sys.argv[1],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['loss_after'])
for k in d['roofline']['kernels'][:5]: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
done
