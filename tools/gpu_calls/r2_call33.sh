#!/bin/bash
O=gpurun_out/c33; mkdir -p $O
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
cat $O/rc.txt
