# ncu evidence of the final build (cfg3, bf16, 1 GPU): launch list of the bench command + --set full of the forward half of a
# variable group, fd9 wgrad and fd9 dgrad (13 launches: the report has to stay under the 64 MiB that travel back)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-microbench --no-secondary"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|vq_assign_f16" -s 330 -c 13 -f -o gpurun_out/prof_r2_final $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
ncu -i gpurun_out/prof_r2_final.ncu-rep --page raw --csv > gpurun_out/prof_r2_final_raw.csv 2>/dev/null
ls -la gpurun_out/
