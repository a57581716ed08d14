# whole GPU suite + sorted-scatter A/B on the stand-alone HBM stages
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for S in 0 1; do
if [ $S = 1 ]; then export PGMVAE_SCATTER_SORTED=1; else unset PGMVAE_SCATTER_SORTED; fi
timeout 600 python bench.py --workload cfg2 --precision tf32 --steps 20 --no-cpu-baseline > gpurun_out/bench_cfg2_sorted$S.json 2> gpurun_out/bench_cfg2_sorted$S.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_cfg2_sorted$S.json'))
print('SORTED $S cfg2 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'])
print(d['hbm_stages'])
PY
done
