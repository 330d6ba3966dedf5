#!/bin/bash
O=gpurun_out/c19; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python tools/kbench.py conv3_tc > $O/kbench_conv3.txt 2>&1; echo "kbench rc=$?" >> $O/rc.txt
timeout -k 10 400 python tools/search_profile.py $O/search_profile.txt > $O/search_profile.log 2>&1; echo "search_profile rc=$?" >> $O/rc.txt
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:c3_tc_kernel -c 2 -o $O/c3_full python tools/kbench.py conv3_tc > $O/ncu_c3.log 2>&1; echo "ncu rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --workload search --steps 3 --warmup 1 > $O/search.json 2> $O/search.err; echo "search rc=$?" >> $O/rc.txt
cat $O/rc.txt
