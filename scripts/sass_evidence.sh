#!/bin/bash
# Static evidence that the conv kernels are tcgen05 / TMEM / TMA code: per kernel, the count of the SASS mnemonics
# B200_PROFILING.md names (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA tensor
# load / store / reduce, SYNCS = mbarrier, UTCBAR = tcgen05.commit).   scripts/sass_evidence.sh > profiles/sass_r01.txt
cd "$(dirname "$0")/.."
cuobjdump -sass peclr_b200/build/conv_tc.o | awk '
  /Function :/ { name = $3; sub(/_ZN5peclr[0-9]*/, "", name); sub(/EvNS_.*/, "", name) }
  /^ +\/\*[0-9a-f]+\*\/ / { split($2, a, "."); op = a[1];
    if (op ~ /^(UTCHMMA|UTCQMMA|UTCMMA|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|LDTM|STTM|UTCBAR|SYNCS|UTCATOMSWS|UTMACMDFLUSH|REDG|UCGABAR_ARV|UCGABAR_WAIT|UTMAPF)$/) cnt[name " " op]++ }
  END { for (k in cnt) printf "%-44s %s\n", k, cnt[k] }' | sort
