# same-box A/B at 1 GPU: bash scripts/gpu_r2_j.sh "ENV1" "ENV2" ... (each: short bench + launch timeline)
mkdir -p gpurun_out
i=0
for E in "$@"; do
i=$((i+1))
env $E PGMVAE_PROF_TIMELINE=gpurun_out/tlj_$i timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_j_$i.json 2> gpurun_out/bench_j_$i.err; echo "== $E rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_j_$i.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],2),d['clocks']['sm_mhz'])
print(' '.join(f"{k['name'].replace('dense_','').replace('_bf16','')}={k['ms_per_step']:.2f}" for k in d['roofline']['kernels'][:7]))
PY
done
