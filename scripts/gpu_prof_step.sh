# ncu --set full of one training step's three big kernels (forward chain, backward chain, merged wgrad)
set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|wgrad_multi" -s 12 -c 3 \
    -f -o gpurun_out/prof_r1_step2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
