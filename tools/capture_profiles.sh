#!/bin/bash
# Profile evidence for profiles/: run on the GPU box (gpurun -- 'bash tools/capture_profiles.sh TAG').
# 1) launch list of bench.py (per-launch durations, serialised)   2) one `--set full` capture of the linear
# layers of one encoder layer (qkv, out_proj, ffn1, ffn2), of the attention kernel and of the attention backward
# 3) launch list of reconstruction-guided steps   4) text exports.
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 130 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 6 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:linear2_kernel \
  -s 69 -c 4 -o $OUT/${TAG}_linear2 -f python bench.py --steps 4 --warmup 3 > $OUT/${TAG}_ncu_linear2.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:attention_persistent \
  -s 17 -c 1 -o $OUT/${TAG}_attention -f python bench.py --steps 4 --warmup 3 > $OUT/${TAG}_ncu_attention.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:attention_bwd_tc \
  -s 16 -c 2 -o $OUT/${TAG}_attention_bwd -f python tools/gpu_guided_probe.py 2 > $OUT/${TAG}_ncu_attention_bwd.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file $OUT/${TAG}_guided_launches.csv python tools/gpu_guided_probe.py 3 > $OUT/${TAG}_ncu_guided.log 2>&1
for k in linear2 attention attention_bwd; do
  ncu -i $OUT/${TAG}_$k.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_$k.ncu-rep --page details > $OUT/${TAG}_${k}_details.txt 2>/dev/null
done
ls -la $OUT | grep $TAG
