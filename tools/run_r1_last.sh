# Last GPU call of round 1 (a few minutes): the GPU tests added in this session, the training tests on the default
# GEMM path, then the CTA-pair tc_gemm variant (P3D_GEMM_CG2=1): exact-product diagnostics, training parity tests, step timing A/B.
set -x
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_eval.py tests/test_gpu_checkpoint.py -x -q > gpurun_out/last_new_tests.log 2>&1; tail -3 gpurun_out/last_new_tests.log
P3D_GEMM_CG2=1 timeout 120 python tools/diag_tcgemm.py > gpurun_out/last_cg2_diag.txt 2>&1; tail -4 gpurun_out/last_cg2_diag.txt
timeout 100 python tools/diag_tcgemm.py > gpurun_out/last_cg1_diag.txt 2>&1; tail -2 gpurun_out/last_cg1_diag.txt
for B in 4096 32768; do
  timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_cg1.txt 2>&1; tail -1 gpurun_out/last_train_${B}_cg1.txt
  P3D_GEMM_CG2=1 timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_cg2.txt 2>&1; tail -1 gpurun_out/last_train_${B}_cg2.txt
done
P3D_GEMM_CG2=1 timeout 200 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/last_train_tests_cg2.log 2>&1; tail -3 gpurun_out/last_train_tests_cg2.log
timeout 200 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp.py -x -q > gpurun_out/last_train_mlp_tests.log 2>&1; tail -3 gpurun_out/last_train_mlp_tests.log
