#!/bin/bash
# SASS evidence for profiles/: per kernel the opcode histogram and the excerpt around the first TMA / mbarrier /
# cp.async / DFMA instructions (cuobjdump -sass of the shipped library; arch must read sm_100a).
#   bash tools/dump_sass.sh > profiles/r2_sass_excerpts.txt
LIB=${1:-macroc_b200/lib/libmacroc_b200.so}
echo "# $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
cuobjdump -sass $LIB > /tmp/all.sass
grep -m1 "arch =" /tmp/all.sass
for pat in 'k_spmv_symILi8ELi3ELi8ELb1E' 'k_spmv_tmaILi8ELi4ELb1E' 'k_assemble_nodes_uniformILb0E' 'k_assemble_nodes_uniformILb1E' 'k_assemble_nodes_pergpILb0E' 'k_apply_mf_marchILb1E' 'k_ab_dmmaILb0E' '14k_cg_update_xrE' 'k_cg_reduce_iter_mbox'; do
    awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on' /tmp/all.sass > /tmp/one.sass
    echo; echo "==== $(grep -m1 'Function :' /tmp/one.sass)"
    echo "-- opcode histogram (static):"
    grep -E "^\s+/\*[0-9a-f]+\*/" /tmp/one.sass | awk '{print $2}' | sed 's/;//' | sed -E 's/^(@!?U?P[0-9T]+)$/PRED/' | sort | uniq -c | sort -rn | head -16 | awk '{printf "   %6d %s\n", $1, $2}'
    echo "-- lines with TMA bulk copies (UBLKCP), mbarrier ops (SYNCS), cp.async (LDGSTS), system-scope ld/st, DMMA, global atomics (last-block ticket):"
    grep -nE "UBLKCP|SYNCS|LDGSTS|\.SYS|ELECT|UTMALDG|DMMA|ATOMG|RED\." /tmp/one.sass | head -14 | cut -c1-150
    echo "-- first DFMA run:"
    grep -n "DFMA" /tmp/one.sass | head -8 | cut -c1-120
done
