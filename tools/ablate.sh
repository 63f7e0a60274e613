#!/bin/bash
# Profiling aid: run the SqueezeDet forward with pieces of the tensor-core pipeline disabled
# (library built with `make EXTRA=-DSQDET_ABLATE`).  Results are wrong by construction; only
# the per-op times matter.  bits: 1 splitter, 2 B loads, 4 A loads, 8 drain ld, 16 epilogue, 32 1-of-3 MMAs
for a in 0 1 2 4 6 8 16 24 32 33 63; do
  echo "== ablate $a"
  SQDET_TC_ABLATE=$a timeout 90 python tests/debug_forward.py squeezeDet 20 2>&1 | grep -E "^fire(3|5|10|11)|^conv12|^total" | tail -6
done
