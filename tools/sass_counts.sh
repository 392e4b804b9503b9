#!/bin/bash
# SASS evidence that the kernels are Blackwell-native (tcgen05 / TMEM / TMA), per object file of the library:
#   tools/sass_counts.sh > profiles/sass_counts.txt
# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG =
# cp.async.bulk.tensor load / store / reduce, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA. = legacy mma.sync (must be 0).
set -e
cd "$(dirname "$0")/.."
python -m pafuse_b200.build >/dev/null 2>&1 || true
echo "# cuobjdump -sass of pafuse_b200/_lib/*.o (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a), $(git rev-parse --short HEAD)"
printf "%-22s %8s %8s %6s %6s %8s %8s %9s %7s %6s %6s\n" object UTCHMMA .2CTA LDTM STTM UTMALDG UTMASTG UTMAREDG UTCBAR SYNCS HMMA.
for o in pafuse_b200/_lib/*.o; do
  s=$(cuobjdump -sass "$o")
  c() { echo "$s" | grep -c -- "$1" || true; }
  legacy=$(echo "$s" | grep -v UTCHMMA | grep -c "HMMA\." || true)
  printf "%-22s %8s %8s %6s %6s %8s %8s %9s %7s %6s %6s\n" "$(basename $o)" "$(c UTCHMMA)" "$(c 'UTCHMMA.2CTA')" "$(c LDTM)" "$(c STTM)" "$(c UTMALDG)" "$(c UTMASTG)" "$(c UTMAREDG)" "$(c UTCBAR)" "$(c SYNCS)" "$legacy"
done
echo "# kernels with tensor-memory code:"
cuobjdump -sass pafuse_b200/_lib/libpafuse_b200.so | awk '/Function :/{f=$3} /UTCHMMA/{n[f]++} END{for(k in n) printf "#   %6d UTCHMMA  %s\n", n[k], k}' | sort -k2 -n -r | cut -c1-220
