#!/bin/bash
O=gpurun_out/c17; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout -k 10 1800 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
NASB_PDL=0 timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_nopdl.json 2> $O/bench_nopdl.err; echo "bench nopdl rc=$?" >> $O/rc.txt
NASB_PDL=0 timeout -k 10 600 python bench.py --workload search --steps 3 --warmup 1 > $O/search_nopdl.json 2> $O/search_nopdl.err; echo "search nopdl rc=$?" >> $O/rc.txt
cat $O/rc.txt
