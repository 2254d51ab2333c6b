# 2-GPU data-parallel check + step timing (gpurun --gpus 2 -- bash tools/run_dp2.sh)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
P3D_DP_OVERLAP=0 timeout 200 $TR tools/dp_check.py 2>&1 | grep -E "DP CHECK|FAILED on" > gpurun_out/dp2_check_ov0.log
P3D_DP_OVERLAP=1 timeout 200 $TR tools/dp_check.py 2>&1 | grep -E "DP CHECK|FAILED on" > gpurun_out/dp2_check_ov1.log
P3D_P2P=0 timeout 200 $TR tools/dp_check.py 2>&1 | grep -E "DP CHECK|FAILED on" > gpurun_out/dp2_check_nccl.log
timeout 200 $TR tools/bench_train_dp.py 64 4096 > gpurun_out/dp2_b.log 2>&1
head -5 gpurun_out/dp2_check_ov0.log gpurun_out/dp2_check_ov1.log gpurun_out/dp2_check_nccl.log; grep us_per gpurun_out/dp2_b.log
