#!/bin/bash
O=gpurun_out/c24; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python tools/kbench.py pw_tc_wgrad > $O/kbench.txt 2>&1; echo "kbench rc=$?" >> $O/rc.txt
timeout -k 10 400 python tools/pdl_sweep.py > $O/sweep.json 2> $O/sweep.err; echo "sweep rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --workload search --steps 3 --warmup 1 > $O/search.json 2> $O/search.err; echo "search rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
