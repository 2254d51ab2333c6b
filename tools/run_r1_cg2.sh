# CTA-pair tc_gemm (P3D_GEMM_CG2=1) on one B200: exact-product diagnostics + timing, training parity tests on the pair
# path, training-step A/B on the same box.
set -x
mkdir -p gpurun_out
P3D_GEMM_CG2=1 timeout 120 python tools/diag_tcgemm.py > gpurun_out/last_cg2_diag.txt 2>&1; tail -4 gpurun_out/last_cg2_diag.txt
P3D_GEMM_CG2=1 timeout 200 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/last_train_tests_cg2.log 2>&1; tail -3 gpurun_out/last_train_tests_cg2.log
for B in 4096 32768; do
  timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_cg1.txt 2>&1; tail -1 gpurun_out/last_train_${B}_cg1.txt
  P3D_GEMM_CG2=1 timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_cg2.txt 2>&1; tail -1 gpurun_out/last_train_${B}_cg2.txt
done
