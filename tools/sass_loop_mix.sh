#!/bin/bash
# Histogram of SASS opcodes between the first VABSDIFF4 and the last VIADDMNMX before tile end of the fast kernel (the round loop body).
cuobjdump -sass "$1" | awk '/Function : .*me_u8_tile/ {f=1} /Function : / && !/me_u8_tile/ {f=0} f && /^ +\/\*[0-9a-f]{4}\*\// {print}' > /tmp/fast.sass
first=$(grep -n "VABSDIFF4" /tmp/fast.sass | head -1 | cut -d: -f1)
last=$(grep -n "BAR.SYNC" /tmp/fast.sass | tail -1 | cut -d: -f1)
start=$((first-40)); [ $start -lt 1 ] && start=1
sed -n "${start},$((last+140))p" /tmp/fast.sass | awk '{op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); c[op]++; n++} END {for (k in c) print c[k], k; print n, "TOTAL(loop approx)"}' | sort -rn | head -${2:-30}
