# Round 2, GPU call B (1 GPU): where does the fused training step at batch 4096 spend its time?
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_train4096_fused_launches.csv \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2b_ncu1.log 2>&1; tail -2 gpurun_out/r2b_ncu1.log
P3D_TRAIN_FUSED=0 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_train4096_unfused_launches.csv \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2b_ncu2.log 2>&1; tail -2 gpurun_out/r2b_ncu2.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel<3" -s 6 -c 1 -o gpurun_out/r2b_fused_fwd \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2b_ncu3.log 2>&1; tail -2 gpurun_out/r2b_ncu3.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel<4" -s 6 -c 1 -o gpurun_out/r2b_fused_bwd \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2b_ncu4.log 2>&1; tail -2 gpurun_out/r2b_ncu4.log
ls -la gpurun_out/*.ncu-rep
