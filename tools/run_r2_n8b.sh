mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=30
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 tools/bench_dp_parts.py 64 4096 32768 > gpurun_out/r2_n8_parts.txt 2>&1; grep "^{" gpurun_out/r2_n8_parts.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 tools/dp_check.py > gpurun_out/r2_n8_dpcheck.txt 2>&1; grep "DP CHECK\|FAILED on rank 0" gpurun_out/r2_n8_dpcheck.txt | head -5
