# Round evidence, part B: the bench line (default flags), the reference arm, the cfg4 VQ micro-benchmark
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref rc=$?"
python pgm-vae_b200/tools/vq_microbench.py --n 16777216 --prec f16 --reps 3 --fused > gpurun_out/vqmb_f16_16M_fused.json 2>&1
python pgm-vae_b200/tools/vq_microbench.py --n 16777216 --prec f16 --reps 3 > gpurun_out/vqmb_f16_16M.json 2>&1
python pgm-vae_b200/tools/vq_microbench.py --n 4194304 --prec f16 --reps 3 --clustered > gpurun_out/vqmb_f16_clustered.json 2>&1
tail -2 gpurun_out/bench_r1.err; cut -c1-400 gpurun_out/bench_r1.json; cut -c1-300 gpurun_out/bench_r1_ref.json
cut -c1-260 gpurun_out/vqmb_f16_16M_fused.json; cut -c1-260 gpurun_out/vqmb_f16_16M.json; cut -c1-260 gpurun_out/vqmb_f16_clustered.json
