#!/bin/bash
# Profile evidence for profiles/: run on the GPU box (gpurun -- 'bash tools/capture_profiles.sh TAG').
# 1) launch list of bench.py (per-launch durations, serialised, cold cache)   2) `--set full` captures of the chained linear
# kernel, the attention kernel and the UNet's convolution GEMM   3) launch list of a UNet step   4) text exports.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 66 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 6 --warmup 3 --skip-configs > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:linear_chain \
  -s 9 -c 1 -o $OUT/${TAG}_chain -f python tools/gpu_b64_steps.py 3 > $OUT/${TAG}_ncu_chain.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:attention_pair \
  -s 9 -c 1 -o $OUT/${TAG}_attention -f python tools/gpu_b64_steps.py 3 > $OUT/${TAG}_ncu_attention.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/${TAG}_unet_launches.csv python tools/gpu_unet_step.py 2 > $OUT/${TAG}_ncu_unet.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:linear2_kernel \
  -s 30 -c 1 -o $OUT/${TAG}_unet_conv -f python tools/gpu_unet_step.py 2 > $OUT/${TAG}_ncu_unet_conv.log 2>&1
for k in chain attention unet_conv; do
  ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$k.ncu-rep --page details > $OUT/${TAG}_${k}_details.txt 2>/dev/null
done
ls -la $OUT | grep $TAG
