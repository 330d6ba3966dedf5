#!/bin/bash
O=gpurun_out/c12; mkdir -p $O
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload search --steps 2 --warmup 1 > $O/search_n2.json 2> $O/search_n2.err; echo "search n2 rc=$?" >> $O/rc.txt
CUDA_VISIBLE_DEVICES=0 timeout -k 10 900 python bench.py --workload search --steps 2 --warmup 1 > $O/search_n1.json 2> $O/search_n1.err; echo "search n1 rc=$?" >> $O/rc.txt
OMP_NUM_THREADS=1 CUDA_VISIBLE_DEVICES=0 timeout -k 10 900 python bench.py --workload search --steps 2 --warmup 1 > $O/search_n1_omp1.json 2> $O/search_n1_omp1.err; echo "search n1 omp1 rc=$?" >> $O/rc.txt
cat $O/rc.txt
