#!/bin/bash
# profiles/artest_dropin.sh -- the reference's own test program (artest.c, unmodified, compiled by oracle/Makefile) run on its
# own sources (artest_ref, CPU) and on libresampler_b200.so (artest_b200, GPU): same counts, same levels, same round-trip error.
D=oracle/_ref
for opts in "-3 -c2 -n60 -s44100 -d48000 -i" "-3 -c2 -n60 -s44100 -d48000 -e -i" "-3 -c2 -n20 -s44100 -d48000 -x -a -i" \
            "-3 -c2 -n20 -s44100 -d48000 -v" "-1 -c1 -n60 -s44100 -d48000" "-4 -c64 -n4 -s96000 -d44100 -l20000" "-2 -c8 -n20 -s48000 -d48005"; do
  echo "=== artest $opts"
  for which in ref b200; do
    s=$(date +%s.%N)
    out=$($D/artest_$which $opts 2>&1 > /dev/null | grep -E "output|diff|fatal|info")
    e=$(date +%s.%N)
    echo "$out" | sed "s/^/[$which] /"
    echo "[$which] wall $(python3 -c "print(round($e - $s, 2))") s"
  done
done
