# Round-2 opener (written at the end of round 1 without GPU access): the TMA-store epilogue of tc_gemm
# (P3D_GEMM_TMASTORE=1, csrc/tc_gemm.cu template parameter TS) on one B200 - exact-product diagnostics on that path,
# then the default path for the A/B, training parity tests on the TS path, training-step A/B on the same box.
#   gpurun --timeout 600 -- 'bash tools/run_r2_tmastore.sh'
set -x
mkdir -p gpurun_out
P3D_GEMM_TMASTORE=1 timeout 150 python tools/diag_tcgemm.py > gpurun_out/r2_ts_diag.txt 2>&1; grep -c MISMATCH gpurun_out/r2_ts_diag.txt; grep -A1 "M=4096\|M=32768\|M=1024 N=1024 K=4096" gpurun_out/r2_ts_diag.txt; tail -2 gpurun_out/r2_ts_diag.txt
timeout 150 python tools/diag_tcgemm.py > gpurun_out/r2_default_diag.txt 2>&1; grep -A1 "M=4096\|M=32768\|M=1024 N=1024 K=4096" gpurun_out/r2_default_diag.txt; tail -1 gpurun_out/r2_default_diag.txt
for B in 64 4096 32768; do
  timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/r2_train_${B}_default.txt 2>&1; tail -1 gpurun_out/r2_train_${B}_default.txt
  P3D_GEMM_TMASTORE=1 timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/r2_train_${B}_ts.txt 2>&1; tail -1 gpurun_out/r2_train_${B}_ts.txt
done
P3D_GEMM_TMASTORE=1 timeout 250 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r2_train_tests_ts.log 2>&1; tail -3 gpurun_out/r2_train_tests_ts.log
