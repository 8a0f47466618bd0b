#!/bin/bash
# Round profile capture (run on the GPU box through gpurun; keeps gpurun_out/ small: summaries + one full report).
#   1. per-kernel HBM / tensor-pipe metrics of one joint layer + one ViT layer (fwd+bwd) + the optimiser, batch 8
#   2. ncu --set full of the four dominant GEMM classes at the bench's batch-32 shapes (reports kept: ~25 MB)
#   3. launch list (gpu__time_duration) of one timed bench step
#   usage: tools/ncu_capture.sh [all|layer] [tag=r02]   (layer = step 1 only, ~2 min)
set -u
MODE=${1:-all}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 500 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/prof_layer python tools/ncu_layer.py 8 9 > $OUT/ncu_layer.log 2>&1
ncu -i /tmp/prof_layer.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_layer_raw.csv 2>/dev/null
python tools/ncu_summarize.py $OUT/${TAG}_ncu_layer_raw.csv > $OUT/${TAG}_ncu_layer.md 2>> $OUT/ncu_layer.log
if [ "$MODE" = "layer" ]; then gzip -f $OUT/${TAG}_ncu_layer_raw.csv; ls -la $OUT; exit 0; fi
for c in geglu dgrad wgrad down; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -f -o $OUT/prof_${TAG}_gemm_$c python tools/one_gemm.py $c 2 > $OUT/ncu_gemm_$c.log 2>&1
  python tools/ncu_summarize.py $OUT/prof_${TAG}_gemm_$c.ncu-rep > $OUT/${TAG}_ncu_gemm_$c.md 2>> $OUT/ncu_gemm_$c.log
done
PI05_CUDA_PROFILER=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err
python tools/launches_summarize.py $OUT/launches_${TAG}.csv > $OUT/${TAG}_launches_summary.md
gzip -f $OUT/launches_${TAG}.csv $OUT/${TAG}_ncu_layer_raw.csv
du -sh $OUT; ls -la $OUT
