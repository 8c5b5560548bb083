#!/bin/bash
# 2-GPU pass: multi-GPU tests on two real devices, sharded VI tuning + waits, bench --gpus 2
TAG=${1:-r2n2}
N=${2:-2}
O=gpurun_out
mkdir -p $O
if [ "$N" = "2" ]; then (timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -6) > $O/tests_multi_$TAG.log 2>&1; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
(timeout 600 $TR tools/vi_shard_bench.py grid4096 f32fast 2>&1 | tail -3) > $O/vishard_$TAG.log 2>&1
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 > $O/bench_n${N}_$TAG.json 2> $O/bench_n${N}_$TAG.err
cat $O/tests_multi_$TAG.log 2>/dev/null | tail -3; tail -2 $O/vishard_$TAG.log | cut -c1-1500; tail -c 300 $O/bench_n${N}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_n${N}_$TAG.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g (%.3f) ms %.2f / %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["value"]/d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
    print("parity", json.dumps(d["parity"])[:1200])
    print("vi_sharded", {k:(round(v["us_per_sweep"],2)) for k,v in d["vi_sharded"].items() if isinstance(v,dict) and "us_per_sweep" in v})
    print("exchange", d["exchange"]["ms_per_step"], d["exchange"]["isolated_ms"])
except Exception as e:
    print("bench unreadable", e)
PY
