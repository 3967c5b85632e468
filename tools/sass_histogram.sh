#!/bin/bash
# Opcode histogram of the shipped library, per kernel: the SASS mnemonics that prove a Blackwell-native path
# (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk,
#  SYNCS = mbarrier, UCGABAR = cluster barrier) beside the legacy ones that must NOT carry the hot path (HMMA).
# Usage: bash tools/sass_histogram.sh > profiles/r02_sass_opcode_histogram.txt   (no GPU needed)
LIB=catre_b200/libcatre_b200.so
echo "# cuobjdump -sass $LIB ($(stat -c %s $LIB) bytes), $(date -u +%F)"
echo "# kernel | UTCHMMA | LDTM | UTMALDG | UTMASTG | UBLKCP | SYNCS | UCGABAR | HMMA | FFMA | MUFU"
cuobjdump -sass $LIB | awk '
/Function :/ { if (name != "") print_row(); name=$3; for (k in c) delete c[k]; next }
{ for (i=1;i<=NF;i++) { op=$i; sub(/\..*/, "", op);
    if (op=="UTCHMMA"||op=="LDTM"||op=="UTMALDG"||op=="UTMASTG"||op=="UBLKCP"||op=="SYNCS"||op=="UCGABAR_ARV"||op=="UCGABAR_WAIT"||op=="HMMA"||op=="FFMA"||op=="MUFU") { if (op ~ /UCGABAR/) op="UCGABAR"; c[op]++ } } }
function print_row() { printf "%s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d\n", name, c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UBLKCP"], c["SYNCS"], c["UCGABAR"], c["HMMA"], c["FFMA"], c["MUFU"] }
END { print_row() }' | while IFS= read -r line; do
  m=$(echo "$line" | cut -d'|' -f1 | tr -d ' ')
  d=$(echo "$m" | c++filt 2>/dev/null | sed 's/(.*//' | cut -c1-70)
  echo "$d |$(echo "$line" | cut -d'|' -f2-)"
done | sort
