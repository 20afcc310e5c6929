#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): the ncu launch list of two eager UNet steps + one decode, and one
# `ncu --set full` capture per hot kernel (tools/prof_kernels.py launches it alone), each summarised with tools/ncu_regions.py.
# Every command runs once WITHOUT ncu first (B200_PROFILING.md).  Outputs: gpurun_out/r2_*.
set -u
out=gpurun_out
mkdir -p $out
python tools/one_step.py > $out/r2_one_step_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2_launches_one_step.csv python tools/one_step.py > $out/r2_one_step_ncu.log 2>&1
python tools/launch_summary.py $out/r2_launches_one_step.csv > $out/r2_launches_one_step.summary.txt 2>&1
for spec in gemm_ab:2:gemm_tc_kernel gemm_c:2:gemm_tc_kernel mlp:0:mlp_fused_kernel mlp:1:mlp_fused_kernel gconv:0:gconv_halo_kernel gconv:2:gconv_halo_kernel norm:0:norm_film norm:2:norm_film attn:0:window_attention_tc_kernel attn:1:window_attention_tc_kernel; do
  IFS=: read which lvl pat <<< "$spec"
  name=r2_ncu_${which}_level${lvl}
  WHICH=$which:$lvl python tools/prof_kernels.py > $out/$name.plain.log 2>&1 &&
  WHICH=$which:$lvl ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 -f -o $out/$name python tools/prof_kernels.py > $out/$name.ncu.log 2>&1
  python tools/ncu_regions.py $out/$name.ncu-rep > $out/$name.txt 2>&1
done
ls -la $out | tail -30
