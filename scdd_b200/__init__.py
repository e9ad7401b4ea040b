"""scdd_b200 -- B200 (sm_100a) implementation of scDD's mixture-fit / permutation-test / KS hot path.

csrc/       CUDA kernels + the C ABI (include/scdd_b200.h) -> libscdd_b200.so
hotpath.py  numpy marshalling onto the C ABI
api.py      host-side mirror of the reference's R interface (scDD, testKS, results)
simulate.py simulateSet-style synthetic inputs for the throughput harness
"""
from .api import SingleCellExperiment, scDD, testKS, results, normcounts, p_adjust_BH  # noqa: F401
