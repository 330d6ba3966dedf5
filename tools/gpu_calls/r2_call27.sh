#!/bin/bash
O=gpurun_out/c27; mkdir -p $O
timeout -k 10 1500 python bench.py --steps 20 --warmup 5 --profile-out $O/percall.txt > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" >> $O/rc.txt
cat $O/rc.txt
