# VQ assignment: parity tests, then the cfg4-shaped micro-benchmark (fp16 tcgen05 kernel, 2 / 3 sub-tiles, fused scatter, clustered data)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -m gpu -x -q -k "vq or assign" > gpurun_out/pytest_vq.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_vq.log
tail -4 gpurun_out/pytest_vq.log
for sub in 2 3; do
  PGMVAE_VQ_SUB=$sub timeout 300 python pgm-vae_b200/tools/vq_microbench.py --n 4194304 --prec f16 --reps 5 > gpurun_out/vqmb_f16_sub$sub.json 2> gpurun_out/vqmb_f16_sub$sub.err
  cut -c1-330 gpurun_out/vqmb_f16_sub$sub.json; tail -3 gpurun_out/vqmb_f16_sub$sub.err
done
timeout 300 python pgm-vae_b200/tools/vq_microbench.py --n 4194304 --prec f16 --reps 5 --fused > gpurun_out/vqmb_f16_fused.json 2>&1
cut -c1-330 gpurun_out/vqmb_f16_fused.json
timeout 300 python pgm-vae_b200/tools/vq_microbench.py --n 4194304 --prec f16 --reps 5 --clustered > gpurun_out/vqmb_f16_clustered.json 2>&1
cut -c1-330 gpurun_out/vqmb_f16_clustered.json
