#!/bin/bash
O=gpurun_out/c31; mkdir -p $O
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 200 python bench.py --augment-only > $O/augment.json 2> $O/augment.err; echo "augment rc=$?" >> $O/rc.txt
cat $O/rc.txt
