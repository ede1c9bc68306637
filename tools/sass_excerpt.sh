#!/bin/bash
# SASS evidence of tcgen05 / TMEM / TMA in the shipped library (CPU only).  Usage: bash tools/sass_excerpt.sh > profiles/r02_sass_excerpt.txt
SO=perm_equiv_graph_neural_cdes_b200/libpegncde.so
cuobjdump -sass $SO > /tmp/peg_sass.txt 2>/dev/null
echo "# SASS evidence that the contraction runs on tcgen05 / TMEM / TMA (cuobjdump -sass $SO, sm_100a), round 2 final build"
echo "# instruction counts over the whole library:"
for m in UTCHMMA UTCQMMA LDTM UTMALDG UTMASTG UTCBAR SYNCS "LDG.E.128" "STS.128" F2FP STL LDL; do printf "%-12s %6d\n" $m $(grep -c "[ .]$m" /tmp/peg_sass.txt); done
echo
echo "# per kernel (k_tc_contract<KIND, FMT>: KIND 0 fwd, 1 adjoint, 2 light adjoint; FMT 0 3xTF32 kind::tf32, 1 bf16x2, 2 fp16x2 kind::f16):"
echo "#  count  mnemonic  function"
awk '/Function :/{fn=$3} { for(i=1;i<=NF;i++){ if($i ~ /^(UTCHMMA|LDTM|UTMALDG|STL|LDL|UTCBAR)/){ split($i,b,"."); cnt[fn" "b[1]]++ } } } END{for(k in cnt) print cnt[k], k}' /tmp/peg_sass.txt | grep "k_tc_\|k_small" | sort -k2,2 -k3,3 | awk '{printf "%6d  %-8s %s\n",$1,$3,$2}'
echo
echo "# first tcgen05 MMA / TMEM load / TMA load lines of the default adjoint kernel k_tc_contract<1,2>:"
awk '/Function :/{p=($3 ~ /k_tc_contractILi1ELi2E/)} p && /UTCHMMA|LDTM|UTMALDG/{print; n++} n>=16{exit}' /tmp/peg_sass.txt | sed 's/^ *//' | cut -c1-150
