# same-box A/B of two builds (lib/libpgmvae_prev.so, lib/libpgmvae_new.so) on the quick bench
set -x
mkdir -p gpurun_out
for rep in 1 2; do
for v in prev new; do
  PGMVAE_LIB=$PWD/pgm-vae_b200/lib/libpgmvae_$v.so timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_lib_$v.json 2> gpurun_out/bench_lib_$v.err
  python - $v <<'PY'
import json,sys
d=json.load(open('gpurun_out/bench_lib_%s.json'%sys.argv[1]))
print(sys.argv[1],'ms',round(d['ms_per_step'],4),'pll',round(d['pll_eval']['value']),' '.join('%s=%.4f'%(k['name'],k['ms_per_step']) for k in d['roofline']['kernels'][:2]))
PY
done
done
