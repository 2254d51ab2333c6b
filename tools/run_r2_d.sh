mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
timeout 120 python tools/diag_fused_phases.py 4096 1 > gpurun_out/r2d_phases_4096.txt 2>&1; cat gpurun_out/r2d_phases_4096.txt
timeout 120 python tools/diag_fused_phases.py 512 1 > gpurun_out/r2d_phases_512.txt 2>&1; cat gpurun_out/r2d_phases_512.txt
for B in 64 512 4096; do
  timeout 90 python tools/train_steps.py $B bf16 20 > gpurun_out/r2d_train_${B}.txt 2>&1; tail -1 gpurun_out/r2d_train_${B}.txt
done
