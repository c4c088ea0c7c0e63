#!/usr/bin/env bash
# Run on the GPU box (under gpurun): GPU test groups in separate processes with timeouts (a kernel fault poisons the
# CUDA context of its own process only), logs under gpurun_out/.  Usage: tools/gpu_ci.sh [group ...]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host_cores.txt
status=0
run() {  # name file kexpr
  local name=$1 file=$2 kexpr=$3
  timeout ${VQ_TEST_TIMEOUT:-600} python -m pytest "$file" -m gpu -q -s --no-header -p no:cacheprovider -k "$kexpr" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
  [ $rc -ne 0 ] && status=1
}
groups=${*:-kernels vq_tc gemm_fp32 gemm_tc enc_fp32 enc_tc decoder data}
for g in $groups; do
  case $g in
    kernels)   run kernels tests/test_gpu_kernels.py "test" ;;
    vq_tc)     run vq_tc tests/test_gpu_vq_tc.py "test" ;;
    gemm_fp32) run gemm_fp32 tests/test_gpu_gemm.py "fp32" ;;
    gemm_tc)   run gemm_tc tests/test_gpu_gemm.py "not fp32" ;;
    gemm_tc0)  VQCPC_GEMM_TILE=0 run gemm_tc0 tests/test_gpu_gemm.py "not fp32" ;;
    gemm_tc1)  VQCPC_GEMM_TILE=1 run gemm_tc1 tests/test_gpu_gemm.py "not fp32" ;;
    gemm_tc2)  VQCPC_GEMM_TILE=2 run gemm_tc2 tests/test_gpu_gemm.py "not fp32" ;;
    gemm_tc3)  VQCPC_GEMM_TILE=3 run gemm_tc3 tests/test_gpu_gemm.py "not fp32" ;;
    gemm_tc4)  VQCPC_GEMM_TILE=4 run gemm_tc4 tests/test_gpu_gemm.py "not fp32" ;;
    enc_fp32)  run enc_fp32 tests/test_gpu_encoder.py "fp32 or micro or eval or first_call or fused" ;;
    enc_tc)    run enc_tc tests/test_gpu_encoder.py "bf16x3" ;;
    decoder)   run decoder tests/test_gpu_decoder.py "test" ;;
    data)      run data tests/test_gpu_dataloader.py "test" ;;
  esac
done
exit $status
