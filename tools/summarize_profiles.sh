#!/bin/bash
# Turns the .ncu-rep files and bench JSONs collect_profiles.sh left in gpurun_out/ into the tracked summaries
# under profiles/.   usage: bash tools/summarize_profiles.sh <tag>
set -u
T=${1:-rX}
for k in step step_n8 rollout deal pack; do
  [ -f gpurun_out/${T}_${k}_full.ncu-rep ] && python tools/ncu_summary.py gpurun_out/${T}_${k}_full.ncu-rep > profiles/${T}_${k}_ncu_full_summary.csv
done
cp gpurun_out/${T}_bench_*.json gpurun_out/${T}_launches.csv profiles/ 2>/dev/null
L="python tools/ncu_lines.py"
{ $L gpurun_out/${T}_step_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n4.o 'step_kernel<4, 0, 1>' --mangled step_kernelILi4ELb0ELb1E --file skyjo_step.cuh --top 16
  $L gpurun_out/${T}_step_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n4.o 'step_kernel<4, 0, 1>' --mangled step_kernelILi4ELb0ELb1E --file skyjo_core.cuh --top 30
  $L gpurun_out/${T}_step_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n4.o 'step_kernel<4, 0, 1>' --mangled step_kernelILi4ELb0ELb1E --opcodes --top 16; } > profiles/${T}_step_lines.txt 2>&1
{ $L gpurun_out/${T}_step_n8_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n8.o 'step_kernel<8, 0, 1>' --mangled step_kernelILi8ELb0ELb1E --file skyjo_step.cuh --top 16
  $L gpurun_out/${T}_step_n8_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n8.o 'step_kernel<8, 0, 1>' --mangled step_kernelILi8ELb0ELb1E --file skyjo_core.cuh --top 30
  $L gpurun_out/${T}_step_n8_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n8.o 'step_kernel<8, 0, 1>' --mangled step_kernelILi8ELb0ELb1E --opcodes --top 16; } > profiles/${T}_step_n8_lines.txt 2>&1
{ $L gpurun_out/${T}_rollout_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n4.o 'rollout_kernel<4, 0>' --mangled rollout_kernelILi4ELb0E --file skyjo_step.cuh --top 16
  $L gpurun_out/${T}_rollout_full.ncu-rep skyjo_rl_b200/build/skyjo_step_n4.o 'rollout_kernel<4, 0>' --mangled rollout_kernelILi4ELb0E --opcodes --top 16; } > profiles/${T}_rollout_lines.txt 2>&1
{ $L gpurun_out/${T}_deal_full.ncu-rep skyjo_rl_b200/build/skyjo_capi.o 'deal_kernel' --mangled deal_kernel --top 16
  $L gpurun_out/${T}_deal_full.ncu-rep skyjo_rl_b200/build/skyjo_capi.o 'deal_kernel' --mangled deal_kernel --opcodes --top 12; } > profiles/${T}_deal_lines.txt 2>&1
