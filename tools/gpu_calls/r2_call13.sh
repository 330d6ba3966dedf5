#!/bin/bash
O=gpurun_out/c13; mkdir -p $O
CUDA_VISIBLE_DEVICES=0 timeout -k 10 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
CUDA_VISIBLE_DEVICES=0 timeout -k 10 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" >> $O/rc.txt
cat $O/rc.txt
