# Last GPU calls of round 1 (a few minutes, ordered by priority; every step bounded by its own timeout):
#  1 GPU tests added in this session; 2 default-path tests touching the edited files (tc_gemm.cu, api.cu);
#  3 CTA-pair tc_gemm (P3D_GEMM_CG2=1): exact-product diagnostics + timing, training parity tests on the pair path,
#    training-step A/B (also the A-multicast variant); 4 bench line with the tapered pipeline schedule on / off.
set -x
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_eval.py tests/test_gpu_checkpoint.py -x -q > gpurun_out/last_new_tests.log 2>&1; tail -3 gpurun_out/last_new_tests.log
timeout 170 python -m pytest tests/test_gpu_train.py tests/test_gpu_mlp.py -x -q -k "bf16_tensor_core or trajectories or ragged or pinned or predict_14 or without_target" > gpurun_out/last_default_tests.log 2>&1; tail -3 gpurun_out/last_default_tests.log
P3D_GEMM_CG2=1 timeout 120 python tools/diag_tcgemm.py > gpurun_out/last_cg2_diag.txt 2>&1; tail -4 gpurun_out/last_cg2_diag.txt
P3D_GEMM_CG2=1 timeout 200 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/last_train_tests_cg2.log 2>&1; tail -3 gpurun_out/last_train_tests_cg2.log
for B in 4096 32768; do
  timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_cg1.txt 2>&1; tail -1 gpurun_out/last_train_${B}_cg1.txt
  P3D_GEMM_CG2=1 timeout 60 python tools/train_steps.py $B bf16 20 > gpurun_out/last_train_${B}_cg2.txt 2>&1; tail -1 gpurun_out/last_train_${B}_cg2.txt
done
P3D_GEMM_MCAST=1 timeout 60 python tools/train_steps.py 4096 bf16 20 > gpurun_out/last_train_4096_mcast.txt 2>&1; tail -1 gpurun_out/last_train_4096_mcast.txt
timeout 100 python tools/diag_tcgemm.py > gpurun_out/last_cg1_diag.txt 2>&1; tail -2 gpurun_out/last_cg1_diag.txt
timeout 90 python bench.py --steps 5 --no-secondary --no-cpu-baseline > gpurun_out/last_bench_taper.json 2> gpurun_out/last_bench_taper.err; echo "bench rc=$?"
P3D_PIPE_TAPER=0 timeout 90 python bench.py --steps 5 --no-secondary --no-cpu-baseline > gpurun_out/last_bench_plain.json 2> gpurun_out/last_bench_plain.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("plain", "taper"):
    try:
        d = json.load(open(f"gpurun_out/last_bench_{n}.json"))
        print(n, "value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 1), "pred-only", round(d["e2e"]["predictions_only"]["value"] / 1e6, 1), d["e2e"]["outputs_match_device_path"])
    except Exception as e:
        print(n, "unreadable:", e)
PY
