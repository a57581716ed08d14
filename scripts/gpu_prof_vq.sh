set -x
mkdir -p gpurun_out
python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vq_assign_f16 -s 2 -c 1 \
    -f -o gpurun_out/prof_r1_vq_f16 python pgm-vae_b200/tools/vq_microbench.py --n 1048576 --prec f16 --reps 2 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
