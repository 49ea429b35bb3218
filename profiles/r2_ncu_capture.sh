#!/bin/bash
# Round-2 ncu evidence for the PPO kernels (run on a B200 through gpurun; reports land in gpurun_out/, summaries are made from
# them with `ncu -i ... --page raw --csv` and committed under profiles/).
#   1. launch list (gpu__time_duration.sum) of one rollout + one epoch of BASELINE-size minibatches (profiles/ppo_profile.py)
#   2. one `--set full` capture each of the tower training kernel, the dual weight-gradient kernel, the cooperative optimizer
#      step and the pipelined rollout forward kernel
set -u
mkdir -p gpurun_out
export T=8 MB=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_ppo_iteration.csv \
    python profiles/ppo_profile.py > gpurun_out/r2_ncu_launches.log 2>&1
for k in tc_tower_train_kernel tc_wgrad_tiled_kernel opt_step_kernel tc_tower_forward_pipe_dual_kernel; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2_ncu_$k \
        python profiles/ppo_profile.py > gpurun_out/r2_ncu_$k.log 2>&1
    tail -2 gpurun_out/r2_ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
