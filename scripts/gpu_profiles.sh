# Round evidence, part A: parity tests, smoke, the ncu launch list of the bench command and --set full captures of
# the top kernels (training step: chain_train + merged wgrad; cfg4-shaped fp16 VQ assignment; encode chain).
# Part B (scripts/gpu_bench.sh) takes the bench lines once profiles/r1_dram_traffic.json has the new captures.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"chain_kernel|wgrad_multi" -s 8 -c 3 \
    -f -o gpurun_out/prof_r1_step python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-microbench > gpurun_out/ncu2.log 2>&1
python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 2 --fused > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vq_assign_f16 -s 2 -c 1 \
    -f -o gpurun_out/prof_r1_vq_f16 python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 2 --fused > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1.csv
