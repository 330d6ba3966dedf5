#!/bin/bash
O=gpurun_out/c14; mkdir -p $O
timeout -k 10 300 python tools/kbench.py sepconv > $O/kb_sep.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 200 python bench.py --metric-only > $O/metric.json 2> $O/metric.err; echo "metric rc=$?" >> $O/rc.txt
timeout -k 10 900 python bench.py --workload search --steps 3 --warmup 1 > $O/search.json 2> $O/search.err; echo "search rc=$?" >> $O/rc.txt
cat $O/rc.txt
