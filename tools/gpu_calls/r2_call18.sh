#!/bin/bash
O=gpurun_out/c18; mkdir -p $O
run() { # name, env...
  n=$1; shift
  env "$@" timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err; echo "bench $n rc=$?" >> $O/rc.txt
  env "$@" timeout -k 10 400 python tools/pdl_sweep.py > $O/sweep_$n.json 2> $O/sweep_$n.err; echo "sweep $n rc=$?" >> $O/rc.txt
}
run off NASB_PDL=0
run t148 NASB_PDL_MAX_CTAS=148
run t1184 NASB_PDL_MAX_CTAS=1184
run inf NASB_PDL_MAX_CTAS=1000000000
timeout -k 10 400 python tools/search_breakdown.py > $O/breakdown.json 2> $O/breakdown.err; echo "breakdown rc=$?" >> $O/rc.txt
NASB_PDL=0 timeout -k 10 400 python tools/search_breakdown.py > $O/breakdown_off.json 2> $O/breakdown_off.err; echo "breakdown off rc=$?" >> $O/rc.txt
cat $O/rc.txt
