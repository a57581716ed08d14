set -x
N=4
mkdir -p gpurun_out
for v in default buckets; do
if [ $v = buckets ]; then export PGMVAE_DP_BUCKETS=1; else unset PGMVAE_DP_BUCKETS; fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-microbench > gpurun_out/bench_n4_$v.json 2> gpurun_out/bench_n4_$v.err
echo "$v $(cut -c1-200 gpurun_out/bench_n4_$v.json)"
done
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-microbench 2>&1 | grep -i -E "NVLS|algo|Channel|Connected" | head -12
