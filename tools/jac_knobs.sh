#!/bin/bash
# A/B of the element-Jacobian traversal (MACROC_ASM_COLBLOCK: tiles per column block, 0 = linear) and store policy
for cb in ${CBS:-0 8 32 64 256}; do for st in ${STS:-1}; do
  echo "colblock $cb stream $st"; MACROC_ASM_COLBLOCK=$cb MACROC_ASM_STREAM=$st python tools/jac_probe.py 256 2>&1 | grep "element-kernel" | awk '{print "   material", $2, $6, "ms"}'
done; done
