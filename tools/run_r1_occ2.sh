# Two-CTAs-per-SM tc_gemm instantiations (P3D_GEMM_OCC2=1) on one B200: exact-product diagnostics + timing, training parity
# tests on that path, training-step A/B on the same box; then the default training tests (incl. the pair-variant subprocess test).
set -x
mkdir -p gpurun_out
P3D_GEMM_OCC2=1 timeout 120 python tools/diag_tcgemm.py > gpurun_out/last_occ2_diag.txt 2>&1; tail -3 gpurun_out/last_occ2_diag.txt
for B in 4096 32768; do
  timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_occ1.txt 2>&1; tail -1 gpurun_out/last_train_${B}_occ1.txt
  P3D_GEMM_OCC2=1 timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_occ2.txt 2>&1; tail -1 gpurun_out/last_train_${B}_occ2.txt
done
P3D_GEMM_OCC2=1 timeout 200 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/last_train_tests_occ2.log 2>&1; tail -3 gpurun_out/last_train_tests_occ2.log
timeout 200 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/last_train_tests_default.log 2>&1; tail -3 gpurun_out/last_train_tests_default.log
