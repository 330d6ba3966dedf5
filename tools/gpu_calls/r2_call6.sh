#!/bin/bash
O=gpurun_out/c6; mkdir -p $O
timeout -k 10 300 python tools/debug_capture.py > $O/debug_capture.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out $O/per_call.txt > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
