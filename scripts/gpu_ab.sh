# A/B over an environment switch: bash scripts/gpu_ab.sh VAR   (unset vs =1); full GPU parity tests first
set -x
mkdir -p gpurun_out
VAR=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in unset 1; do
if [ $v = unset ]; then unset $VAR; else export $VAR=1; fi
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
tail -2 gpurun_out/bench_ab_$v.err
python - $VAR $v <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_ab_%s.json'%sys.argv[2]))
print(sys.argv[1],sys.argv[2],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'pll',round(d['pll_eval']['value']), d['loss_after']['loss'], ' '.join('%s=%.4f'%(k['name'],k['ms_per_step']) for k in d['roofline']['kernels'][:4]))
PY
done
