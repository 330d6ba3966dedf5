#!/bin/bash
O=gpurun_out/c34; mkdir -p $O
timeout -k 10 400 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt
