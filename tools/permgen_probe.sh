# rng_kernel_ms of quick_perf (8192 genes x 2000 cells x 16 permutations, 4 chunks) for group sizes / warps per block
for gs in 8 32 4; do for wpb in 4 2 1; do echo -n "GS=$gs WPB=$wpb : "; SCDD_PERMGEN_GS=$gs SCDD_PERMGEN_WPB=$wpb QP_CHUNKS=4 python tools/quick_perf.py 8192 2000 16 2>&1 | grep -o "'rng_kernel_ms': [0-9.]*"; done; done
