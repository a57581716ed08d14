# bf16 GEMM operator tests + kernel-only throughput at the cfg3 layer shapes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -x -q -s > gpurun_out/pytest_bf16.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bf16.log
tail -30 gpurun_out/pytest_bf16.log
timeout 300 python pgm-vae_b200/tools/bf16_microbench.py 16 4096 > gpurun_out/bf16_microbench.jsonl 2> gpurun_out/bf16_microbench.err
tail -5 gpurun_out/bf16_microbench.err
cat gpurun_out/bf16_microbench.jsonl
