# usage: tools/ab.sh variant...   (libscdd_b200_<variant>.so); two interleaved rounds, 6 chunks each
for round in 1 2; do for v in "$@"; do echo -n "== $v : "; QP_CHUNKS=6 SCDD_LIB_OVERRIDE=$PWD/scdd_b200/libscdd_b200_$v.so timeout 300 python tools/quick_perf.py 1024 2000 16 2>&1 | grep -E "BEST"; done; done
