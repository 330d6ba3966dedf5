#!/bin/bash
O=gpurun_out/c15; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout -k 10 200 python bench.py --metric-only > $O/metric.json 2> $O/metric.err; echo "metric rc=$?" >> $O/rc.txt
timeout -k 10 1800 python bench.py --steps 20 --warmup 5 --profile-out $O/per_call.txt > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" >> $O/rc.txt
cat $O/rc.txt
