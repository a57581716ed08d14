# ncu evidence for the default bench command (cfg3, bf16): launch list + --set full of the GEMM kernels and the VQ kernel
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-microbench --no-secondary"
timeout 600 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|vq_assign_f16" -s 330 -c 34 -f -o gpurun_out/prof_r2_step $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/prof_r2_step.ncu-rep gpurun_out/r2_launches.csv
