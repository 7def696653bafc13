#!/bin/bash
# ncu counters for the megakernel-vs-wavefront comparison (divergence, issue slots, FP32 pipe)
set -x
mkdir -p gpurun_out
M="smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active"
for scene in "" "--glass"; do
  tag=std; [ -n "$scene" ] && tag=glass
  CMD="python scripts/sweep.py --width 1200 --spp 8 --reps 1 $scene --configs"
  $CMD wavefront:2:16 > /dev/null 2>&1
  ncu --metrics $M --clock-control none -k regex:"wf_intersect|wf_shade|wf_generate" -s 60 -c 60 --csv --log-file gpurun_out/m_wf_$tag.csv $CMD wavefront:2:16 > /dev/null 2>&1
  ncu --metrics $M --clock-control none -k regex:rz_path_kernel -s 1 -c 1 --csv --log-file gpurun_out/m_mega_$tag.csv $CMD mega:2:16 > /dev/null 2>&1
  ncu --metrics $M --clock-control none -k regex:rz_bvh_kernel -s 1 -c 1 --csv --log-file gpurun_out/m_bvh_$tag.csv $CMD bvh:1:16 > /dev/null 2>&1
done
ls -la gpurun_out/m_*
