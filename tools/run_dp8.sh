# 8-GPU data-parallel check + step timing (gpurun --gpus 8 -- bash tools/run_dp8.sh)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tools/dp_check.py 2>&1 | grep -E "DP CHECK|FAILED on" > gpurun_out/dp${N}_check.log
P3D_DP_OVERLAP=1 timeout 200 $TR tools/bench_train_dp.py 64 4096 32768 2>&1 | grep us_per > gpurun_out/dp${N}_ov1.log
P3D_DP_OVERLAP=0 timeout 200 $TR tools/bench_train_dp.py 64 4096 32768 2>&1 | grep us_per > gpurun_out/dp${N}_ov0.log
head -3 gpurun_out/dp${N}_check.log; cat gpurun_out/dp${N}_ov1.log gpurun_out/dp${N}_ov0.log
