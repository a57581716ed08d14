# data-parallel check + bench on N GPUs: bash scripts/gpu_dp.sh N [both]   (default schedule; "both": also PGMVAE_DP_BUCKETS=1)
set -x
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 pgm-vae_b200/tools/dp_check.py 2>&1 | grep -v "^W\|OMP_NUM\|\*\*\*" | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
cut -c1-260 gpurun_out/bench_n$N.json
if [ "$2" = both ]; then
PGMVAE_DP_BUCKETS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_n${N}_buckets.json 2> gpurun_out/bench_n${N}_buckets.err
cut -c1-260 gpurun_out/bench_n${N}_buckets.json
fi
