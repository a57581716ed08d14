# same-box A/B of two builds of the library on the cfg4-shaped VQ micro-benchmark
set -x
mkdir -p gpurun_out
for rep in 1 2; do
for v in prev new; do
  PGMVAE_LIB=$PWD/pgm-vae_b200/lib/libpgmvae_$v.so timeout 300 python pgm-vae_b200/tools/vq_microbench.py --n 4194304 --prec f16 --reps 5 > gpurun_out/vqmb_ab_$v.json 2> gpurun_out/vqmb_ab_$v.err
  echo "$v $(cut -c100-260 gpurun_out/vqmb_ab_$v.json)"; tail -2 gpurun_out/vqmb_ab_$v.err
done
done
