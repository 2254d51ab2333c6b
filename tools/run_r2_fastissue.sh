# Round-2 opener (written at the end of round 1 without GPU access): lean MMA issue path of tc_gemm
# (P3D_GEMM_FASTISSUE=1, template parameter FI: descriptors advanced by addition instead of rebuilt per k-block).
# Exact-product diagnostics, inference + training parity tests on that path, then the A/B where a k-block's fixed cost
# matters: the layered inference range (batch 2..4096) and the batch-64 / 4096 training step.
#   gpurun --timeout 900 -- 'bash tools/run_r2_fastissue.sh'
set -x
mkdir -p gpurun_out
P3D_GEMM_FASTISSUE=1 timeout 150 python tools/diag_tcgemm.py > gpurun_out/r2_fi_diag.txt 2>&1; grep -c MISMATCH gpurun_out/r2_fi_diag.txt; grep -A1 "M=4096\|M=32768" gpurun_out/r2_fi_diag.txt; tail -2 gpurun_out/r2_fi_diag.txt
P3D_GEMM_FASTISSUE=1 timeout 300 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_train.py -x -q -k "not pair_gemm" > gpurun_out/r2_fi_tests.log 2>&1; tail -3 gpurun_out/r2_fi_tests.log
timeout 120 python tools/bench_sweep.py 1024 2 13 > gpurun_out/r2_sweep_default.jsonl 2>&1
P3D_GEMM_FASTISSUE=1 timeout 120 python tools/bench_sweep.py 1024 2 13 > gpurun_out/r2_sweep_fi.jsonl 2>&1
paste -d'\n' gpurun_out/r2_sweep_default.jsonl gpurun_out/r2_sweep_fi.jsonl | cut -c1-60
for B in 64 4096; do
  timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/r2_fi_train_${B}_default.txt 2>&1; tail -1 gpurun_out/r2_fi_train_${B}_default.txt
  P3D_GEMM_FASTISSUE=1 timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/r2_fi_train_${B}_fi.txt 2>&1; tail -1 gpurun_out/r2_fi_train_${B}_fi.txt
done
