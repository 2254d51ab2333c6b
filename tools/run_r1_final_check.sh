# After the column-sum local-memory fix of the plain tc_gemm epilogues: exactness + timing, training tests, step times.
set -x
mkdir -p gpurun_out
timeout 100 python tools/diag_tcgemm.py > gpurun_out/final_diag.txt 2>&1; grep -A1 "M=32768\|M=4096 N=1024 K=1024 a_mn=0 b_mn=1" gpurun_out/final_diag.txt; tail -1 gpurun_out/final_diag.txt
timeout 120 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp.py -x -q -k "not pair_gemm" > gpurun_out/final_train_mlp_tests.log 2>&1; tail -2 gpurun_out/final_train_mlp_tests.log
for B in 4096 32768 64; do timeout 40 python tools/train_steps.py $B bf16 20 > gpurun_out/final_train_${B}.txt 2>&1; tail -1 gpurun_out/final_train_${B}.txt; done
