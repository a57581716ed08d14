# sweep of an environment knob over the quick bench: bash scripts/gpu_sweep.sh VAR v1 v2 ...
set -x
mkdir -p gpurun_out
VAR=$1; shift
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -2
for v in "$@"; do
export $VAR=$v
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err
tail -2 gpurun_out/bench_sweep.err
python - $VAR $v <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_sweep.json'))
print(sys.argv[1],sys.argv[2],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'pll',round(d['pll_eval']['value']), ' '.join('%s=%.4f'%(k['name'],k['ms_per_step']) for k in d['roofline']['kernels'][:3]))
PY
done
