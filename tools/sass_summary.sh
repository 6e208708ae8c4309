#!/bin/bash
# Tracked proof of the Blackwell-native claim: tcgen05 / TMEM / TMA / mbarrier SASS mnemonics per kernel of the in-tree library.
# usage: tools/sass_summary.sh > profiles/sass_summary.txt   (cuobjdump from the CUDA toolkit; no GPU needed)
SO=real-time-video-quality-analysis_b200/libvqa_b200.so
echo "# cuobjdump -sass $SO  ($(cuobjdump --version | tail -1); built $(date -u -r $SO +%Y-%m-%dT%H:%MZ))"
echo "# arch of every cubin: $(cuobjdump -lelf $SO | sed 's/.*\.\(sm_[0-9a-z]*\)\..*/\1/' | sort | uniq -c | tr '\n' ' ')"
cuobjdump -sass $SO > /tmp/vqa_sass.txt
echo "# whole library:"
for m in UTCHMMA UTCQMMA UTMALDG UTMASTG UTMAPF LDTM STTM UTCBAR UTCCP SYNCS "LDGSTS" "ATOMS" "RED\." "REDUX" "SHFL" "MATCH" "UBLKCP" "TEX\b" ; do
  printf "%-10s %6d\n" "$m" "$(grep -c -E "\b$m" /tmp/vqa_sass.txt)"
done
echo "# per kernel (count of UTCHMMA / UTMALDG / LDTM / UTCBAR / SYNCS / LDGSTS / ATOMS, then the function):"
awk '
/Function :/ { if (name != "") printf "%4d %4d %4d %4d %4d %4d %4d  %s\n", a, b, c, d, e, f, g, name; name=$3; a=b=c=d=e=f=g=0 }
/UTCHMMA/ {a++} /UTMALDG/ {b++} /LDTM/ {c++} /UTCBAR/ {d++} /SYNCS/ {e++} /LDGSTS/ {f++} /ATOMS/ {g++}
END { printf "%4d %4d %4d %4d %4d %4d %4d  %s\n", a, b, c, d, e, f, g, name }' /tmp/vqa_sass.txt | c++filt | sed 's/(.*)//' | sort -k8
