#!/bin/bash
# round-2 first GPU call: validate the opt-in switches written blind in round 1 (run them or delete them)
O=gpurun_out/c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
NASB_FULLSIZE_TESTS=1 timeout -k 10 600 python -m pytest tests -m gpu -q > $O/pytest_default.log 2>&1; echo "default rc=$?" >> $O/rc.txt
timeout -k 10 300 python tools/kbench.py pw_tc_fwd bn_finalize bn_act_bwd pool3x3 > $O/kb_default.txt 2>&1; echo "kb default rc=$?" >> $O/rc.txt
for SW in NASB_PW_WS NASB_BN_COOP NASB_POOL_PACK; do
  env $SW=1 timeout -k 10 400 python -m pytest tests -m gpu -q > $O/pytest_$SW.log 2>&1; echo "$SW pytest rc=$?" >> $O/rc.txt
  env $SW=1 timeout -k 10 300 python tools/kbench.py pw_tc_fwd bn_finalize bn_act_bwd pool3x3 > $O/kb_$SW.txt 2>&1; echo "$SW kb rc=$?" >> $O/rc.txt
done
timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" >> $O/rc.txt
NASB_PW_WS=1 timeout -k 10 600 python bench.py --steps 10 --warmup 3 > $O/bench_ws.json 2> $O/bench_ws.err; echo "bench ws rc=$?" >> $O/rc.txt
cat $O/rc.txt
