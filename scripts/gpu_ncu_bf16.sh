# ncu --set full of the bf16 GEMM kernels at the two big cfg3 layer shapes (operator microbench, G = 32)
set -x
mkdir -p gpurun_out
timeout 300 python pgm-vae_b200/tools/bf16_microbench.py 32 4096 1556x400 400x1556 > gpurun_out/mb_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 10 -c 10 -f -o gpurun_out/prof_bf16 python pgm-vae_b200/tools/bf16_microbench.py 32 4096 1556x400 400x1556 > gpurun_out/ncu_bf16.log 2>&1
tail -3 gpurun_out/ncu_bf16.log
cat gpurun_out/mb_plain.log
