#!/bin/bash
# Per-kernel SASS mnemonic counts of the built library (evidence that the hot kernels are tcgen05 / TMEM / TMA code).
# Usage: tools/sass_summary.sh > profiles/rN_sass_summary.txt     (CPU only; needs cuobjdump)
SO="$(dirname "$0")/../transformer-gan_b200/tgan_b200/libtgan_b200.so"
echo "# cuobjdump -sass $(basename "$SO") (sm_100a), mnemonic counts per kernel"
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,"
echo "# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async, REDG / ATOMG = global reductions,
# HMMA / LDSM = warp-level mma.sync / ldmatrix (the 64-token discriminator attention tiles)"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn = $3; order[++n] = fn }
  fn != "" {
    if ($0 ~ /UTCHMMA\.2CTA/) c[fn,"UTCHMMA.2CTA"]++; else if ($0 ~ /UTCHMMA/) c[fn,"UTCHMMA"]++;
    if ($0 ~ /LDTM/) c[fn,"LDTM"]++; if ($0 ~ /STTM/) c[fn,"STTM"]++;
    if ($0 ~ /UTMALDG/) c[fn,"UTMALDG"]++; if ($0 ~ /UTMASTG/) c[fn,"UTMASTG"]++; if ($0 ~ /UTMAREDG/) c[fn,"UTMAREDG"]++;
    if ($0 ~ /UTCBAR/) c[fn,"UTCBAR"]++; if ($0 ~ /SYNCS/) c[fn,"SYNCS"]++; if ($0 ~ /LDGSTS/) c[fn,"LDGSTS"]++;
    if ($0 ~ /REDG|RED\.E/) c[fn,"REDG"]++; if ($0 ~ /ATOMG/) c[fn,"ATOMG"]++;
    if ($0 ~ /HMMA/ && $0 !~ /UTCHMMA/) c[fn,"HMMA"]++; if ($0 ~ /LDSM/) c[fn,"LDSM"]++;
    if ($0 ~ /^ +\/\*[0-9a-f]+\*\/ /) c[fn,"total"]++;
  }
  END {
    split("UTCHMMA UTCHMMA.2CTA LDTM STTM UTMALDG UTMASTG UTMAREDG UTCBAR SYNCS LDGSTS REDG ATOMG HMMA LDSM total", k, " ");
    for (i = 1; i <= n; i++) { fn = order[i]; line = "";
      for (j = 1; j <= 15; j++) if (c[fn,k[j]] > 0) line = line " " k[j] "=" c[fn,k[j]];
      print fn ":" line }
  }' | c++filt | sed -e 's/(anonymous namespace):://g' -e 's/^void //' -e 's/(.*):/:/' | sort
