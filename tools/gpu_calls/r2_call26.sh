#!/bin/bash
O=gpurun_out/c26; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python tools/kbench.py dwconv_tile conv3_tc > $O/kbench.txt 2>&1; echo "kbench rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --workload search --steps 3 --warmup 1 > $O/search.json 2> $O/search.err; echo "search rc=$?" >> $O/rc.txt
cat $O/rc.txt
