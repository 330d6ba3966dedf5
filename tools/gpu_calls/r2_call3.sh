#!/bin/bash
O=gpurun_out/c3; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -s > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/rc.txt
timeout -k 10 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/rc.txt
python - > $O/h2d.txt 2>&1 <<'P'
import torch, time
x = torch.empty(218103808, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("pinned H2D 218 MB: %.3f ms  %.1f GB/s" % (ms, 218.1 / ms))
y = torch.empty(8, 3, 1024, 2048)
t0 = time.time(); yp = y.pin_memory(); print("pin_memory of 201 MB: %.1f ms" % ((time.time() - t0) * 1e3))
P
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 --profile-out $O/per_call.txt > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/rc.txt
