#!/bin/bash
# same-box A/B of two builds of libhfg.so: tools/ab/libhfg_old.so, tools/ab/libhfg_new.so
cp flagger_b200/libhfg.so /tmp/keep.so
for rep in 1 2 3; do
  for v in old new; do
    cp tools/ab/libhfg_$v.so flagger_b200/libhfg.so
    echo "$v: $(python tools/em_mean.py ${1:-cfg2})"
  done
done
cp /tmp/keep.so flagger_b200/libhfg.so
