#!/bin/bash
O=gpurun_out/c30; mkdir -p $O
for T in 1184 4096 1000000000; do
  NASB_PDL_MAX_CTAS=$T timeout -k 10 300 python bench.py --workload search --steps 3 --warmup 1 > $O/search_$T.json 2> $O/search_$T.err; echo "search $T rc=$?" >> $O/rc.txt
done
NASB_PDL_MAX_CTAS=4096 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_4096.json 2> $O/bench_4096.err; echo "bench 4096 rc=$?" >> $O/rc.txt
cat $O/rc.txt
