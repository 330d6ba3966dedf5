#!/bin/bash
O=gpurun_out/c29; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
B="python bench.py --steps 1 --warmup 3 --cuda-graph 0 --no-extras --no-cpu-baseline"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $B > $O/launches.log 2>&1; echo "ncu rc=$?" >> $O/rc.txt
python tools/summarize_launches.py $O/launches.csv 40 > $O/launches_summary.txt 2>&1
gzip -f $O/launches.csv
cat $O/rc.txt
