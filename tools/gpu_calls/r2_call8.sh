#!/bin/bash
O=gpurun_out/c8; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -q -k "segmenter_matches" > $O/pytest_iso1.log 2>&1; echo "iso segmenter_matches rc=$?" >> $O/rc.txt
timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -q -k "graph and segmenter_matches" > $O/pytest_iso2.log 2>&1; echo "iso graph only rc=$?" >> $O/rc.txt
timeout -k 10 600 python -m pytest tests/test_gpu_engine.py -q > $O/pytest_engine.log 2>&1; echo "engine file rc=$?" >> $O/rc.txt
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 200 python bench.py --metric-only > $O/metric.json 2> $O/metric.err; echo "metric rc=$?" >> $O/rc.txt
cat $O/rc.txt
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 --profile-out $O/per_call.txt > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
NASB_ASYNC_WGRAD=0 timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_sync_wgrad.json 2> $O/bench_sync_wgrad.err; echo "bench sync rc=$?" >> $O/rc.txt
cat $O/rc.txt
