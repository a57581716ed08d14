set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 120 python __graft_entry__.py --smoke > gpurun_out/dbg_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_smoke.log
tail -5 gpurun_out/dbg_smoke.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke > gpurun_out/dbg_san.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_san.log
grep -v "^$" gpurun_out/dbg_san.log | head -60
