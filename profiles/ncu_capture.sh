#!/bin/bash
# ncu --set full capture of one headline kernel, reduced on the GPU box to the text summaries kept under profiles/
# (the .ncu-rep files are too large to bring back in bulk).   usage: profiles/ncu_capture.sh <what> <kernel-regex> <tag>
what=$1; regex=$2; tag=$3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 1 -c 1 -o /tmp/${tag} python profiles/prof_driver.py $what > gpurun_out/ncu_${tag}.log 2>&1
{
  ncu -i /tmp/${tag}.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_summary.py
  ncu -i /tmp/${tag}.ncu-rep --page raw --csv 2>/dev/null | python profiles/ncu_l1_breakdown.py 2>/dev/null | tail -32
  ncu -i /tmp/${tag}.ncu-rep --page source --csv 2>/dev/null | python profiles/ncu_source_top.py 2>/dev/null | head -40
} > gpurun_out/${tag}.txt
rm -f /tmp/${tag}.ncu-rep
