#!/bin/bash
# Multi-GPU visit (under gpurun --gpus N): sharded-vs-unsharded parity test, then the bench at N ranks for both splits.
TAG=${1:-m}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
for SH in utterances frames; do
  NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --shard $SH --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_${SH}.log 2> gpurun_out/${TAG}_bench_${SH}.err
  grep '^{' gpurun_out/${TAG}_bench_${SH}.log | tail -1 > gpurun_out/${TAG}_bench_${SH}.json
  grep -c "AllGather\|ncclAllGather" gpurun_out/${TAG}_bench_${SH}.log gpurun_out/${TAG}_bench_${SH}.err | head -2
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${SH}.json"))
    print("$SH n_gpus", d["n_gpus"], "value %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", d.get("e2e"))
except Exception as e:
    print("$SH bench parse failed", e); print(open("gpurun_out/${TAG}_bench_${SH}.err").read()[-1500:])
PY
done
