#!/bin/bash
# compute-sanitizer passes over the kernels that use mbarriers, bulk async copies, system-scope fences and a host-polled flag
# (SURVEY.md §5): small-N GPU tests of the fused tower-train kernel, the tiled wgrad, the tensor-core forward, the host step
# with mapped result blocks (incl. the chunk-released variant whose CTAs poll host memory, and its rollback), the dependent
# launches of the fused update with the optimizer step's own grid barrier, and the fused policy rollout.  Run under gpurun; summaries land in gpurun_out/ and are copied to
# profiles/r2_sanitizer_{memcheck,racecheck,synccheck}.txt.
#   bash profiles/sanitizer.sh [tool ...]        (default: memcheck racecheck synccheck)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
SEL='tests/test_train_fused_gpu.py::test_mn_major_descriptor_probe
tests/test_train_fused_gpu.py::test_wgrad_tiled_matches_torch[128]
tests/test_train_fused_gpu.py::test_wgrad_tiled_matches_torch[1024]
tests/test_train_fused_gpu.py::test_fused_minibatch_matches_unfused[ball3d-64-8-100]
tests/test_train_fused_gpu.py::test_fused_minibatch_matches_unfused[ball3d-512-32-4096]
tests/test_rollout_fused_gpu.py::test_fused_rollout_is_bit_identical_to_per_step_calls[ball3d-512-32-bf16]
tests/test_envs_gpu.py::test_host_step_contract
tests/test_envs_gpu.py::test_chunked_host_step_matches_device_path_and_rolls_back[ball3d]
tests/test_envs_gpu.py::test_chunked_host_step_matches_device_path_and_rolls_back[gridworld]
tests/test_tc_gpu.py::test_pipelined_rollout_forward_is_bit_identical_to_the_classic_kernel
tests/test_ppo_gpu.py::test_adam_clip_vs_torch
tests/test_ppo_gpu.py::test_adam_zero_grads_and_operand_image_refresh'
for tool in $TOOLS; do
  out=gpurun_out/sanitizer_${tool}.log
  timeout 1500 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 0 \
      python -m pytest -x -q -p no:cacheprovider $SEL > "$out" 2>&1
  echo "== $tool rc=$?" >> "$out"
  { echo "== compute-sanitizer --tool $tool ($(date -u +%FT%TZ))"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" "$out" | tail -12; } > gpurun_out/sanitizer_${tool}_summary.txt
done
cat gpurun_out/sanitizer_*_summary.txt
