#!/bin/bash
O=gpurun_out/c32; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_augment.py -m gpu -q --timeout 200 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt
