#!/bin/bash
# SASS evidence per object file of libpolyphemus_b200: Blackwell-native instructions (tcgen05 -> UTC*MMA, tcgen05.ld -> LDTM,
# TMA -> UTMALDG, bulk L2 prefetch -> UBLKPF, cp.async -> LDGSTS, mbarrier -> SYNCS) and the absence of legacy tensor paths / float atomics.
OBJ=polyphemus_b200/lib/obj
echo "# cuobjdump -sass of every object (sm_100a), instruction counts"
printf "%-22s %8s %8s %8s %8s %8s %8s %8s %8s %10s %8s\n" object UTCHMMA UTCxMMA LDTM UTMALDG UBLKPF LDGSTS UTCBAR SYNCS "ATOM/RED.F" HMMA
for o in $OBJ/*.o; do
  s=$(cuobjdump -sass $o 2>/dev/null)
  c() { echo "$s" | grep -cE "$1"; }
  printf "%-22s %8s %8s %8s %8s %8s %8s %8s %8s %10s %8s\n" $(basename $o) $(c 'UTCHMMA') $(c 'UTC[A-Z]*MMA') $(c 'LDTM') $(c 'UTMALDG') $(c 'UBLKPF') $(c 'LDGSTS') $(c 'UTCBAR') $(c 'SYNCS') \
    $(c '(ATOM|RED)[A-Z.]*\.(F32|F16|BF16|ADD\.F)') $(c '[^C]HMMA\.')
done
