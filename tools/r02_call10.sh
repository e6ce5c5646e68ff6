#!/bin/bash
# round 2, GPU call 10: self-join schedule A/B on one rank's share of C3 (pair vs single CTA), alternating
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for rep in 1 2; do for P in 0 1; do
  echo "MMRS_SJ_PAIR=$P"
  MMRS_SJ_PAIR=$P timeout 300 python tools/c3_one_rank.py 10000000 0 2>&1 | tail -1
  nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
done; done | tee gpurun_out/r02_selfjoin_ab.log
