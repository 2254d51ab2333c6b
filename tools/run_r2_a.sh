# Round 2, GPU call A (1 GPU): the pruned tc_gemm + grid-synchronised fused training epilogues.
#   gpurun --timeout 1500 -- 'bash tools/run_r2_a.sh'
mkdir -p gpurun_out
export P3D_SYNC_TIMEOUT_S=20
timeout 150 python tools/diag_tcgemm.py > gpurun_out/r2a_diag.txt 2>&1; grep -c MISMATCH gpurun_out/r2a_diag.txt; tail -2 gpurun_out/r2a_diag.txt
for B in 64 512 1024 4096 32768; do
  timeout 90 python tools/train_steps.py $B bf16 20 > gpurun_out/r2a_train_${B}.txt 2>&1; tail -1 gpurun_out/r2a_train_${B}.txt
  P3D_TRAIN_FUSED=0 timeout 90 python tools/train_steps.py $B bf16 20 > gpurun_out/r2a_train_${B}_unfused.txt 2>&1; tail -1 gpurun_out/r2a_train_${B}_unfused.txt
done
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp_golden.py -x -q > gpurun_out/r2a_tests_train.log 2>&1; tail -15 gpurun_out/r2a_tests_train.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_train.py --deselect tests/test_gpu_mlp_golden.py > gpurun_out/r2a_tests_rest.log 2>&1; tail -5 gpurun_out/r2a_tests_rest.log
