#!/bin/bash
# Multi-GPU evidence for round 2, one gpurun call:
#   /usr/local/graft/bin/gpurun --gpus N --timeout 1500 -- 'bash tools/round2_multi.sh N > gpurun_out/r2_multi_nN.log 2>&1'
# (1) NCCL decomposition tests (slabs, x/y/PETSC_DECIDE boxes; full, matrix-free, per-GP and symmetric operators)
#     with the mailbox all-reduce (default) -- and once more, shortened, with MACROC_ALLREDUCE=nccl
# (2) bench.py at N ranks with both all-reduce paths: cg_iteration_ms is the A/B
N=${1:-2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MACROC_TEST_WORLD=$N MACROC_TEST_CASES=${CASES:-quick} timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q 2>&1 | tail -5
MACROC_ALLREDUCE=nccl MACROC_TEST_WORLD=$N MACROC_TEST_CASES=boxes timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q -k nccl 2>&1 | tail -3
for ar in mailbox nccl; do
    MACROC_ALLREDUCE=$ar timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 --no-extras > gpurun_out/r2_bench_n${N}_${ar}.json 2> gpurun_out/r2_bench_n${N}_${ar}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n${N}_${ar}.json"))
    print("$ar", "value", d["value"], "ms/step", d["ms_per_step"], "its", d["cg_iterations_per_step"], "cg_iteration_ms", d["cg_iteration_ms"],
          "apply_ms", d["roofline"]["launch_ms"], "full:", (d.get("assembled_full") or {}).get("cg_iteration_ms"))
except Exception as e:
    print("$ar", "bench failed", e)
PY
    tail -3 gpurun_out/r2_bench_n${N}_${ar}.err
done
