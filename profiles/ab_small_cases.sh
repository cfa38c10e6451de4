#!/bin/bash
for so in build_variants/sp*.so; do
  echo "== $(basename $so .so)"
  DABRY_B200_LIBRARY=$PWD/$so python profiles/small_case_latency.py movor three_vortices obstacle 2>&1 | tail -6
done
