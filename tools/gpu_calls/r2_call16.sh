#!/bin/bash
O=gpurun_out/c16; mkdir -p $O
nproc > $O/host.txt; cat /sys/fs/cgroup/cpu.max >> $O/host.txt 2>/dev/null; uptime >> $O/host.txt
timeout -k 10 900 python bench.py --workload search --steps 3 --warmup 1 > $O/search_alone.json 2> $O/search_alone.err; echo "search alone rc=$?" >> $O/rc.txt
uptime >> $O/host.txt
timeout -k 10 1800 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
uptime >> $O/host.txt
cat $O/rc.txt
