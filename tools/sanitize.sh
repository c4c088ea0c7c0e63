#!/usr/bin/env bash
# compute-sanitizer pass over the hand-written kernels (SURVEY.md section 5 "Race detection"): memcheck, racecheck and
# synccheck on a selection of small kernel tests that together launch every kernel family of libvqcpc_b200.so (tcgen05
# GEMMs incl. the CTA-pair / multicast variants, attention fwd/bwd, the tcgen05 code search, the cluster/DSMEM backward,
# the streaming InfoNCE kernels, LayerNorm, GRU gates, optimizer).  Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Writes gpurun_out/sanitize_<tool>.log and a one-line-per-tool summary to gpurun_out/sanitize_summary.txt
# (copied to profiles/ once reviewed).  Each tool is bounded by its own timeout so a hang cannot eat the GPU budget.
set -uo pipefail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL=${SANITIZE_SELECT:-"test_layernorm_fwd_bwd or test_attention_fwd_bwd_vs_oracle or test_attention_kept_rows_and_split_io or test_quantizer_bit_exact_vs_reference_golden or test_infonce_fwd_bwd_vs_oracle or test_clip_adam_matches_oracle or test_embed_fwd_bwd or test_posemb_fill_and_bwd or test_subsample_colsum_selu_reduce or (test_native_gru_matches_torch_gru and 5-16) or (test_vq_tc_scores_and_indices_small and 3-16-1000) or (test_vq_fused_bwd_matches_autograd and 3-5000) or test_gemm_epilogues or test_gemm_presplit_operands_split_outputs_and_bf16_mask or (idxpath and small)"}
FILES=""
for f in tests/test_gpu_kernels.py tests/test_gpu_vq_tc.py tests/test_gpu_gemm.py tests/test_gpu_idxpath.py tests/test_gpu_student.py; do [[ -f $f ]] && FILES="$FILES $f"; done
TOOLS=${SANITIZE_TOOLS:-"memcheck racecheck synccheck"}
PER_TOOL_TIMEOUT=${SANITIZE_TIMEOUT:-420}
: > gpurun_out/sanitize_summary.txt
for tool in $TOOLS; do
  log=gpurun_out/sanitize_${tool}.log
  extra=""
  [[ $tool == memcheck ]] && extra="--leak-check no"
  timeout "$PER_TOOL_TIMEOUT" compute-sanitizer --tool "$tool" $extra --print-limit 20 --error-exitcode 99 \
      python -m pytest $FILES -m gpu -x -q -k "$SEL" -p no:cacheprovider > "$log" 2>&1
  rc=$?
  errs=$(grep -c "^========= .*\(Invalid\|Race\|Barrier error\|hazard\)" "$log" || true)
  summ=$(grep "ERROR SUMMARY\|RACECHECK SUMMARY" "$log" | tail -1)
  tests=$(grep -E "passed|failed" "$log" | tail -1)
  echo "$tool: exit=$rc findings=$errs | ${summ:-no summary line} | pytest: ${tests:-n/a}" | tee -a gpurun_out/sanitize_summary.txt
done
