# Round evidence: bench line, ncu launch list of the same command, --set full captures of the top kernels.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|dense_tc_kernel" -s 18 -c 12 \
    -o gpurun_out/prof_r1_step python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/ncu2.log 2>&1
python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 2 --fused > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vq_assign_f16 -s 2 -c 1 \
    -o gpurun_out/prof_r1_vq_f16 python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 2 --fused > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/bench_r1.err; cut -c1-300 gpurun_out/bench_r1.json
