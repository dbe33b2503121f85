#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMA / TMEM use (profiles/r2_sass_counts.txt).
# usage: tools/sass_counts.sh   (after __graft_entry__.build(); needs cuobjdump and c++filt, no GPU)
cuobjdump -sass lightweight_openpose_b200/_lwop.so > /tmp/lwop.sass
for m in UTCHMMA 'UTCHMMA\.2CTA' LDTM STTM UTMALDG 'UTMALDG\.[0-9]D\.IM2COL' UTMASTG UTCBAR FFMA2; do
  printf "%-28s %6d\n" "$m" "$(grep -cE "\b$m" /tmp/lwop.sass)"
done
