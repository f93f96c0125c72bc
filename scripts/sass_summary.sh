#!/bin/bash
# SASS evidence per kernel of the in-tree libcgq.so: counts of the mnemonics that prove the Blackwell-native paths
# (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk copy,
# HMMA/IMMA = legacy warp MMA, LDSM = ldmatrix, SYNCS = mbarrier, UCGABAR = cluster barrier)
lib=${1:-chatglm_q_b200/libcgq.so}
cuobjdump -sass "$lib" 2>/dev/null | awk '
  /Function :/ { fn=$3; sub(/^_ZN3cgq[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_/, "", fn); names[fn]=1; next }
  { for (i=1;i<=NF;i++) { t=$i; sub(/\..*/, "", t);
      if (t ~ /^(UTC[A-Z]*MMA|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|IMMA|LDSM|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|BAR|REDG|ATOMG|RED|LDGSTS)$/) c[fn" "t]++ } }
  END { for (k in c) print k, c[k] }' | sort | awk '{ if ($1!=last) { if (last!="") print line; line=$1 ":"; last=$1 } line=line " " $2 "=" $3 } END { print line }' | c++filt 2>/dev/null | cut -c1-400
