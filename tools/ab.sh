# usage: [QP_GENES=8192] tools/ab.sh variant...   (scdd_b200/libscdd_b200_<variant>.so); two interleaved rounds, 6 chunks each.
# Use enough genes that every warp runs ~100 fit-sets per chunk: at 1024 genes a chunk is 14 fit-sets per warp and its
# time is set by the slowest few (a throughput change of several percent is invisible).
for round in 1 2; do for v in "$@"; do echo -n "== $v : "; QP_CHUNKS=6 SCDD_LIB_OVERRIDE=$PWD/scdd_b200/libscdd_b200_$v.so timeout 300 python tools/quick_perf.py ${QP_GENES:-8192} 2000 16 2>&1 | grep -E "BEST"; done; done
