#!/bin/bash
# same-box A/B of the self-join tensor-core kernel variants on one rank's share of C3 (run under gpurun).
# A = the library in the tree; B, C = variant libraries linked from the tree's objects with another
# selfjoin_mma.cu (B: the single-CTA kernel of commit 93ac352, C: the same with the warp-uniform MMA
# issue loop), built into tools/_ab/ (git-ignored) with the flags of <pkg>/build.py.  Result:
# profiles/r01_selfjoin_ab.log; it led to the duration-based choice in sjm_pair_mode().
PKG=multi-modal-retrieval-system-image-search-and-data-governance_b200
cp $PKG/lib/libmmrs_b200.so /tmp/lib_a.so
run() { cp $2 $PKG/lib/libmmrs_b200.so; echo "variant $1"; timeout 300 python tools/c3_one_rank.py 10000000 0 2>&1 | tail -1; nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader; }
run A-pair /tmp/lib_a.so
run B-old tools/_ab/libmmrs_b.so
run C-single-uniform tools/_ab/libmmrs_c.so
run A-pair /tmp/lib_a.so
cp /tmp/lib_a.so $PKG/lib/libmmrs_b200.so
