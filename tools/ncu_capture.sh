#!/bin/bash
# Captures `ncu --set full` for one kernel and leaves only CSV summaries (raw + details pages) in gpurun_out/.
# usage: tools/ncu_capture.sh <name> <kernel-regex> <skip> <command...>
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/$name -f "$@" > gpurun_out/$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page details --csv > gpurun_out/$name.details.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
rm -f /tmp/$name.ncu-rep
