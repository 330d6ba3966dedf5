#!/bin/bash
# Round-2 profiling pass (GPU box, one GPU): regenerates the ncu evidence under gpurun_out/prof/ that profiles/r2_* summarise.
#   1. launch list of the bench command (eager iteration: every kernel is its own launch), gpu__time_duration per launch
#   2. DRAM traffic / tensor-pipe table for the dominant kernels
#   3. one --set full capture of the top kernel (largest BN-backward dz launch)
# Numbers printed by a run under ncu are never bench values.
O=gpurun_out/prof; mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --cuda-graph 0 --no-extras --no-cpu-baseline"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $B > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv 40 > $O/launches_summary.txt 2>&1
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size \
  --clock-control none -k regex:"bn_bwd_dz_bf16|bn_bwd_sums_bf16|pw_tc_kernel|pw_tc_ws_kernel|pw_wgrad_tc|affine_act_bf16|dw_tile_kernel|dw_wgrad_tile|c3_tc_kernel|resize_bwd" \
  -c 2500 --csv --log-file $O/traffic.csv $B > $O/traffic.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:bn_bwd_dz_bf16 -s 30 -c 2 -f -o $O/r2_top_kernel $B > $O/full.log 2>&1
gzip -f $O/launches.csv $O/traffic.csv
ls -la $O
