# Round 2, GPU call C (1 GPU): restructured fused training epilogues (phase 0 under the mainloop, TMEM scratch).
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
timeout 900 python -m pytest tests/test_gpu_mlp_golden.py -x -q > gpurun_out/r2c_tests_train.log 2>&1; tail -5 gpurun_out/r2c_tests_train.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_train4096_fused_launches.csv \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2c_ncu1.log 2>&1; tail -1 gpurun_out/r2c_ncu1.log
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_kernel<\(int\)3' -s 6 -c 1 -o gpurun_out/r2c_fused_fwd \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2c_ncu3.log 2>&1; tail -1 gpurun_out/r2c_ncu3.log
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_kernel<\(int\)4' -s 6 -c 1 -o gpurun_out/r2c_fused_bwd \
  python tools/train_steps.py 4096 bf16 3 > gpurun_out/r2c_ncu4.log 2>&1; tail -1 gpurun_out/r2c_ncu4.log
timeout 300 python -m pytest tests/test_gpu_geometry.py -x -q > gpurun_out/r2c_tests_geo.log 2>&1; tail -2 gpurun_out/r2c_tests_geo.log
ls -la gpurun_out/*.ncu-rep
