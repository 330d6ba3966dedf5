#!/bin/bash
O=gpurun_out/c10; mkdir -p $O
timeout -k 10 600 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
bash tools/profile_r2.sh > $O/profile.log 2>&1; echo "profile rc=$?" >> $O/rc.txt
cat $O/rc.txt
