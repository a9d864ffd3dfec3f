#!/bin/bash
# Run on the GPU box (via gpurun): collects the ncu evidence that profiles/ summarises.
#   1. launch list of one eager step (per-launch device time; cold-cache, serialised -> compare shares)
#   2. DRAM traffic + duration of the vocoder's conv launches (one vocoder-only pass)
#   3. ncu --set full of representative vocoder conv launches, summarised in place
#      (gpurun_out is capped at 64 MiB: only one .ncu-rep is kept, the rest become text summaries)
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/profile_step.py > gpurun_out/${TAG}_prof_step.log 2>&1
STEPS=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:conv_ --csv --log-file gpurun_out/${TAG}_vocoder_traffic.csv python tools/profile_step.py --vocoder-only \
    > gpurun_out/${TAG}_prof_voc.log 2>&1
for cfg in "128 7 640000" "64 7 1920000" "32 11 3840000" "256 11 128000"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:conv_ -s 8 -c 1 -o gpurun_out/${TAG}_conv_c$1k$2 \
      python tools/prof_conv.py $1 $2 $3 res > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_conv_c$1k$2.ncu-rep 16 > gpurun_out/${TAG}_conv_c$1k$2_summary.txt 2>&1
  if [ "$1" != "128" ]; then rm -f gpurun_out/${TAG}_conv_c$1k$2.ncu-rep; fi
done
ls -la gpurun_out | tail -14
