mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
timeout 120 python tools/diag_fused_phases.py 4096 1 > gpurun_out/r2e_phases_4096.txt 2>&1; cat gpurun_out/r2e_phases_4096.txt
for B in 64 512 4096; do
  timeout 90 python tools/train_steps.py $B bf16 20 > gpurun_out/r2e_train_${B}.txt 2>&1; tail -1 gpurun_out/r2e_train_${B}.txt
done
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp_golden.py -x -q > gpurun_out/r2e_tests_train.log 2>&1; tail -4 gpurun_out/r2e_tests_train.log
