#!/bin/bash
# Run on the GPU box (via gpurun): collects the ncu evidence that profiles/ summarises.
#   1. launch list of one eager step (per-launch device time; cold-cache, serialised -> compare shares)
#   2. DRAM traffic + duration of the vocoder's launches (one vocoder-only pass)
#   3. ncu --set full of the fused resblock-pair kernel (one launch per channel count) and of the
#      stage-0 implicit-GEMM conv, summarised in place (gpurun_out is capped at 64 MiB: one .ncu-rep is kept)
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/profile_step.py > gpurun_out/${TAG}_prof_step.log 2>&1
STEPS=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'conv_|resblock_pair' --csv --log-file gpurun_out/${TAG}_vocoder_traffic.csv python tools/profile_step.py --vocoder-only \
    > gpurun_out/${TAG}_prof_voc.log 2>&1
for cfg in "128 7 3" "64 7 3" "32 7 3" "64 11 5"; do
  set -- $cfg
  ncu --set full --clock-control none --import-source on -k regex:resblock_pair -s 2 -c 1 -o gpurun_out/${TAG}_pair_c$1k$2d$3 \
      python tools/prof_pair.py 2 $1 $2 $3 > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/${TAG}_pair_c$1k$2d$3.ncu-rep 16 > gpurun_out/${TAG}_pair_c$1k$2d$3_summary.txt 2>&1
  if [ "$1$2" != "647" ]; then rm -f gpurun_out/${TAG}_pair_c$1k$2d$3.ncu-rep; fi
done
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 8 -c 1 -o gpurun_out/${TAG}_conv_c256k11 \
    python tools/prof_conv.py 256 11 128000 res > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_conv_c256k11.ncu-rep 16 > gpurun_out/${TAG}_conv_c256k11_summary.txt 2>&1
rm -f gpurun_out/${TAG}_conv_c256k11.ncu-rep
# 4. the fused InstanceNorm + AdaIN kernel (HBM roofline entry of bench.py): full capture of one launch + per-launch DRAM bytes
ncu --set full --clock-control none --import-source on -k regex:adain_ring -s 8 -c 1 -o gpurun_out/${TAG}_adain \
    python tools/prof_adain.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_adain.ncu-rep 16 > gpurun_out/${TAG}_adain_summary.txt 2>&1
rm -f gpurun_out/${TAG}_adain.ncu-rep
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:adain_ring -c 40 --csv \
    --log-file gpurun_out/${TAG}_adain.csv python tools/prof_adain.py > /dev/null 2>&1
ls -la gpurun_out | tail -14
