# sharded peer-to-peer exchange fused with Adam: bash scripts/gpu_r2_h.sh N  (single-GPU operator tests, DP parity tool, bench at N)
N=$1
mkdir -p gpurun_out
[ -n "$SKIP_PYTEST" ] || timeout 600 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_h.log
[ -n "$SKIP_DPCHECK" ] || timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 pgm-vae_b200/tools/dp_check.py > gpurun_out/dp_check_h_n$N.log 2>&1; echo "dp_check rc=$?"
grep -v "^W\|OMP_NUM\|\*\*\*" gpurun_out/dp_check_h_n$N.log | cut -c1-420 | tail -24
for SH in ${SHARDS:-1 0}; do
PGMVAE_P2P_SHARD=$SH timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu-baseline --no-microbench --no-secondary > gpurun_out/bench_h_n${N}_s$SH.json 2> gpurun_out/bench_h_n${N}_s$SH.err; echo "bench shard=$SH rc=$?"
tail -2 gpurun_out/bench_h_n${N}_s$SH.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_h_n${N}_s$SH.json'))
print('N $N shard $SH value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'pll',d['pll_eval']['value'], d['clocks'])
print('dp_parity',d['dp_parity'])
for k in d['roofline']['kernels']: print(k['name'],round(k['ms_per_step'],4),round(k['GBps']),round(k['TFLOPs'],1))
PY
done
