mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
timeout 900 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2_dp2_tests.log 2>&1; tail -3 gpurun_out/r2_dp2_tests.log; grep -n "FAILED on rank" gpurun_out/r2_dp2_tests.log | head -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/bench_dp_parts.py 64 4096 8192 > gpurun_out/r2_dp2_parts.txt 2>&1; grep "^{" gpurun_out/r2_dp2_parts.txt
P3D_P2P_GRAD=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 tools/bench_dp_parts.py 4096 > gpurun_out/r2_dp2_parts_nccl.txt 2>&1; grep "^{" gpurun_out/r2_dp2_parts_nccl.txt
