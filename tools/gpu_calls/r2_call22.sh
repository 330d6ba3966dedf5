#!/bin/bash
O=gpurun_out/c22; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python tools/kbench.py pw_tc conv3_tc sepconv > $O/kbench.txt 2>&1; echo "kbench rc=$?" >> $O/rc.txt
timeout -k 10 400 python bench.py --batch 32 --height 350 --width 350 --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out $O/percall_350.txt > $O/bench_350.json 2> $O/bench_350.err; echo "bench350 rc=$?" >> $O/rc.txt
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --profile-out $O/percall.txt > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
