set -x
mkdir -p gpurun_out
export PGMVAE_WGRAD_ORIENT=d
timeout 300 python pgm-vae_b200/tools/bf16_microbench.py 148 4096 1556x400 > gpurun_out/mb_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 6 -c 3 -f -o gpurun_out/prof_wgrad python pgm-vae_b200/tools/bf16_microbench.py 148 4096 1556x400 > gpurun_out/ncu_wgrad.log 2>&1
tail -3 gpurun_out/ncu_wgrad.log
cat gpurun_out/mb_plain.log
ls -la gpurun_out/
