#!/bin/bash
# SASS listings of the hot kernels of the built library -> profiles/r02_sass_<kernel>.txt(.gz), plus a mnemonic histogram
so=isaacgyminsertion_b200/libigi_b200.so
out=profiles/r02_sass_summary.txt
: > $out
for k in tac_contactILi1 tac_geom pcl_compact_kernel fps_sorted_kernel; do
  sym=$(cuobjdump -elf $so 2>/dev/null | grep -o "_ZN[A-Za-z0-9_]*${k}[A-Za-z0-9_]*" | sort -u | head -1)
  cuobjdump -sass -fun "$sym" $so > profiles/r02_sass_$k.txt 2>/dev/null
  n=$(grep -cE "^\s+/\*[0-9a-f]{4}\*/" profiles/r02_sass_$k.txt)
  echo "== $k ($sym): $n instructions" >> $out
  grep -E "^\s+/\*[0-9a-f]{4}\*/" profiles/r02_sass_$k.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -24 | awk '{printf "%s %s; ", $2, $1} END {print ""}' >> $out
  echo "   bulk-copy / barrier / cluster mnemonics: $(grep -oE "UBLKCP[.A-Z]*|SYNCS[.A-Z]*|UCGABAR[.A-Z_]*|REDUX[.A-Z0-9]*|FFMA2|FADD2|FMUL2|CCTL[.A-Z]*|ATOMS[.A-Z0-9]*|MUFU[.A-Z0-9]*" profiles/r02_sass_$k.txt | sort | uniq -c | awk '{printf "%s x%s  ", $2, $1}')" >> $out
  gzip -f profiles/r02_sass_$k.txt
done
cat $out
