#!/bin/bash
O=gpurun_out/c20; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python tools/kbench.py conv3_tc > $O/kbench_conv3.txt 2>&1; echo "kbench rc=$?" >> $O/rc.txt
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/search_launches.csv python tools/search_profile.py --graphs > $O/search_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/rc.txt
gzip -f $O/search_launches.csv
timeout -k 10 400 python tools/search_cprofile.py $O/search_cprofile.txt > $O/search_cprofile.log 2>&1; echo "cprofile rc=$?" >> $O/rc.txt
cat $O/rc.txt
