#!/bin/bash
# run the reference-over-the-seam loop-back many times; print every run that does not come out clean
cd "$(dirname "$0")/.." || exit 1
kind=${1:-gpu}; n=${2:-25}
fails=0
for i in $(seq 1 $n); do
  for cfg in "1 0" "8 3" "32 19" "40 35"; do
    set -- $cfg
    port=$((21000 + (i * 17 + $1) % 9000))
    out=$(timeout 120 ./oracle/_ref/ref_seam_loopback_$kind $port $1 5 $2 11 /tmp/cap_$kind.bin 2>/tmp/seam_err.txt); rc=$?
    if [ $rc -ne 0 ]; then fails=$((fails+1)); echo "run $i cfg $cfg rc=$rc $out"; grep -a "cm256 (sdrd" /tmp/seam_err.txt | head -3; fi
  done
done
echo "failures: $fails of $((n*4))"
