#!/bin/bash
# Opening measurements for the next round, one gpurun call on 2 GPUs:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_open.sh > gpurun_out/round2_open.log 2>&1'
# (1) parity of the symmetric-storage operator on several ranks (guarded path)
# (2) a probe build of the symmetric SpMV that drops classes of neighbour-block gathers (results
#     are wrong, timings only): 1 = as shipped, 257 = without the gathers whose block this CTA
#     streamed itself, 513 = without the ones streamed by another CTA / an earlier z segment,
#     769 = without any gather (the stream alone)
set -x
cd "$(dirname "$0")/.."
MACROC_SYM_MULTIRANK=1 MACROC_TEST_WORLD=2 timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q -k nccl
mkdir -p gpurun_out
make -C macroc_b200/csrc OUT="$PWD/gpurun_out/libprobe.so" EXTRA=-DMACROC_SYM_PROBE "$PWD/gpurun_out/libprobe.so"
python tools/sym_only.py 256 full
for h in 1 257 513 769; do
    echo "probe hint $h"
    MACROC_B200_LIB="$PWD/gpurun_out/libprobe.so" MACROC_SYM_HINT=$h timeout 120 python tools/sym_only.py 256
done
