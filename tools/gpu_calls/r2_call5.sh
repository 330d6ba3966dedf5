#!/bin/bash
O=gpurun_out/c5; mkdir -p $O
timeout -k 10 300 python tools/debug_capture.py > $O/debug_capture.txt 2>&1
timeout -k 10 300 python tools/kbench.py resize_bwd > $O/kb_resize.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
cat $O/rc.txt
