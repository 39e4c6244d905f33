#!/bin/bash
# ncu evidence for profiles/: launch lists (C2, C3) + one `--set full` capture of the head-of-pass launch of each hot kernel (C2).
# Usage: gpurun -- bash tools/gpu_profile.sh <tag>.  Numbers printed by runs under ncu are never bench values.
tag=${1:-r1}
mkdir -p gpurun_out
M=gpu__time_duration.sum,launch__grid_size,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_c2_launches.csv python tools/profile_run.py C2 2 > gpurun_out/${tag}_ncu_c2.log 2>&1
timeout 500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${tag}_c3_launches.csv python tools/profile_run.py C3 1 > gpurun_out/${tag}_ncu_c3.log 2>&1
for k in k_camera_shade k_camera_merge_coop k_camera_connect_deferred k_light_bounce k_shadow_trace k_trace_closest k_camera_continue; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o /tmp/p_$k python tools/profile_run.py C2 1 > gpurun_out/${tag}_ncu_full_$k.log 2>&1
  if [ -f /tmp/p_$k.ncu-rep ]; then
    ncu -i /tmp/p_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_c2_$k.raw.csv 2>/dev/null
  fi
done
# source-level view of the dominant kernel only (kept small)
if [ -f /tmp/p_k_camera_merge_coop.ncu-rep ]; then
  ncu -i /tmp/p_k_camera_merge_coop.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/${tag}_c2_k_camera_merge_coop.source.csv
fi
ls -la gpurun_out | tail -20
