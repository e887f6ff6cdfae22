#!/bin/bash
# A/B of two builds of the library on ONE box (box-to-box variation is ~1 %): head = sast_b200/libsast_b200_head.so,
# new = sast_b200/libsast_b200.so; alternating runs of the headline bench
for k in 1 2 3; do
  for l in libsast_b200_head.so libsast_b200.so; do
    printf "%s " $l
    SAST_B200_LIB=sast_b200/$l python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
  done
done
