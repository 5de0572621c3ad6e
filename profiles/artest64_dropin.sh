#!/bin/bash
# the reference's artest.c compiled with -DPATH_WIDTH=64 (its artest64 target), once on the reference's own sources and once on
# libresampler_b200_64.so (oracle/Makefile builds both): same options, outputs side by side
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for opts in "-1 -c1 -n4 -s44100 -d48000" "-3 -c2 -n4 -s44100 -d48000 -i" "-2 -c2 -n4 -s48000 -d44100 -e -i" "-4 -c8 -n2 -s96000 -d44100 -l20000" "-2 -c8 -n4 -s48000 -d48005 -v"; do
  for bin in artest64_ref artest64_b200; do
    echo "== $bin $opts"
    /usr/bin/env time -f "   wall %e s" oracle/_ref/$bin $opts 2>&1 | grep -E "count =|wall|extra sample" || oracle/_ref/$bin $opts 2>&1 | grep -E "count =|extra sample"
  done
done
