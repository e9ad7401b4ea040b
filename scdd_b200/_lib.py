"""ctypes binding of libscdd_b200.so (the C ABI declared in include/scdd_b200.h).

The library is the product: there is no Python/CPU implementation of the hot path here.  If the shared
object is missing it is built with nvcc; if it cannot be built or no CUDA device is usable, the compute
entry points raise -- they never fall back.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

GMAX = 5
_LIB = None

ERRORS = {-1: "SCDD_ERR_NO_DEVICE", -2: "SCDD_ERR_BAD_ARG", -3: "SCDD_ERR_CUDA", -4: "SCDD_ERR_TOO_LARGE",
          -5: "SCDD_ERR_FIT_FAILED", -6: "SCDD_ERR_INTERRUPTED"}
ERR_FIT_FAILED = -5
ERR_INTERRUPTED = -6


class ScddError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Cfg(C.Structure):
    _fields_ = [("alpha", C.c_double), ("m0", C.c_double), ("s0", C.c_double), ("a0", C.c_double),
                ("b0", C.c_double), ("min_size", C.c_int32), ("restrict_", C.c_int32),
                ("extra_mstep", C.c_int32), ("ks_old_pvalue", C.c_int32)]


class Fit(C.Structure):
    _fields_ = [("G", C.c_int32), ("model_v", C.c_int32), ("n", C.c_int32), ("df", C.c_int32),
                ("bic", C.c_double), ("loglik", C.c_double),
                ("mean", C.c_double * GMAX), ("var", C.c_double * GMAX)]


INTERRUPT_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p)


class PermOpts(C.Structure):
    _fields_ = [("seed_mode", C.c_int32), ("reserved", C.c_int32), ("jp_terms", C.c_void_p),
                ("interrupt", INTERRUPT_FN), ("interrupt_arg", C.c_void_p)]


STATS_FIELDS = ["triples", "point_iters", "fit_sets", "em_iters", "iter_cap_hits", "tol_near", "bic_near",
                "const_sets", "failed_sets", "fit_kernel_ms", "fit_kernel_launches", "rng_kernel_ms", "total_ms",
                "h2d_bytes", "d2h_bytes", "debug_violations"]


class Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in STATS_FIELDS]

    def as_dict(self):
        return {n: getattr(self, n) for n in STATS_FIELDS}


FIT_DTYPE = np.dtype([("G", np.int32), ("model_v", np.int32), ("n", np.int32), ("df", np.int32),
                      ("bic", np.float64), ("loglik", np.float64), ("mean", np.float64, (GMAX,)),
                      ("var", np.float64, (GMAX,))])

class GlmOpts(C.Structure):
    """scdd_glm_opts of include/scdd_b200.h"""
    _fields_ = [("prior_scale", C.c_double), ("prior_scale_for_intercept", C.c_double), ("prior_df", C.c_double),
                ("epsilon", C.c_double), ("maxit", C.c_int32), ("reserved", C.c_int32)]


class ClassifyOpts(C.Structure):
    """scdd_classify_opts of include/scdd_b200.h"""
    _fields_ = [("ndraws", C.c_int32), ("reserved", C.c_int32), ("seed", C.c_uint64), ("comp_off", C.c_void_p),
                ("comp_detail", C.c_void_p), ("comp_cap", C.c_int64)]


CATEGORIES = ("NC", "DE", "DP", "DM", "DB")      # SCDD_CAT_*


EXPORTS = ["scdd_abi_version", "scdd_last_error", "scdd_device_count", "scdd_set_devices", "scdd_default_cfg",
           "scdd_fit_observed", "scdd_perm_bf", "scdd_perm_bf_ex", "scdd_release_cache", "scdd_ks", "scdd_rng_gene_seeds", "scdd_rng_sample_perms",
           "scdd_rng_mt_sample_perms", "scdd_plan_create", "scdd_plan_run_chunk", "scdd_plan_bf_device_ptr",
           "scdd_plan_stats", "scdd_plan_gene_costs", "scdd_plan_destroy", "scdd_selftest_x87_cumsum", "scdd_selftest_exp", "scdd_selftest_softmax", "scdd_selftest_log",
           "scdd_measure_fp64_peak", "scdd_dense_scan", "scdd_dense_to_csr", "scdd_dense_cell_covariate", "scdd_labels_to_dense", "scdd_cluster_counts", "scdd_fedp", "scdd_selftest_student_p",
           "scdd_default_glm_opts", "scdd_test_zeroes", "scdd_default_classify_opts", "scdd_classify",
           "scdd_selftest_student_t"]


def lib_path():
    return _build.LIB


def lib():
    """Loads (building first if needed) libscdd_b200.so and declares the prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("SCDD_LIB_OVERRIDE") or _build.build()   # override: A/B testing of kernel builds
    L = C.CDLL(path)
    dp, ip, lp, up = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint32))
    vp = C.c_void_p
    L.scdd_abi_version.restype = C.c_int32
    L.scdd_last_error.restype = C.c_char_p
    L.scdd_device_count.restype = C.c_int32
    L.scdd_set_devices.argtypes = [ip, C.c_int32]
    L.scdd_default_cfg.argtypes = [C.POINTER(Cfg)]
    L.scdd_default_cfg.restype = None
    L.scdd_fit_observed.argtypes = [vp, vp, vp, C.c_int32, C.POINTER(Cfg), vp, vp, vp, vp, vp, C.POINTER(Stats)]
    L.scdd_perm_bf.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, C.POINTER(Cfg), vp, vp, C.POINTER(Stats)]
    L.scdd_perm_bf_ex.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp, vp, vp, C.POINTER(Cfg), C.POINTER(PermOpts), vp, vp,
                                  C.POINTER(Stats)]
    L.scdd_release_cache.restype = None
    L.scdd_ks.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, vp, vp, C.POINTER(Stats)]
    L.scdd_rng_gene_seeds.argtypes = [C.c_uint32, C.c_int32, vp]
    L.scdd_rng_sample_perms.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.scdd_rng_mt_sample_perms.argtypes = [C.c_uint32, C.c_int32, C.c_int32, vp]
    L.scdd_plan_create.argtypes = [C.c_int32, vp, vp, vp, C.c_int32, C.POINTER(Cfg), C.c_int32, vp, vp]
    L.scdd_plan_create.restype = vp
    L.scdd_plan_run_chunk.argtypes = [vp, vp, C.c_int32]
    L.scdd_plan_bf_device_ptr.argtypes = [vp]
    L.scdd_plan_bf_device_ptr.restype = vp
    L.scdd_plan_stats.argtypes = [vp, C.POINTER(Stats)]
    L.scdd_plan_gene_costs.argtypes = [vp, vp]
    L.scdd_plan_destroy.argtypes = [vp]
    L.scdd_plan_destroy.restype = None
    L.scdd_selftest_x87_cumsum.argtypes = [vp, C.c_int32, vp]
    L.scdd_selftest_exp.argtypes = [vp, C.c_int32, vp]
    L.scdd_selftest_softmax.argtypes = [vp, C.c_int32, vp, vp]
    L.scdd_selftest_log.argtypes = [vp, C.c_int32, vp]
    L.scdd_measure_fp64_peak.argtypes = [C.c_int32, dp]
    L.scdd_measure_fp64_peak.restype = C.c_double
    i64 = C.c_int64
    L.scdd_dense_scan.argtypes = [vp, i64, i64, C.c_int32, vp, vp, vp, vp, vp, vp]
    L.scdd_dense_to_csr.argtypes = [vp, i64, i64, C.c_int32, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    L.scdd_dense_cell_covariate.argtypes = [vp, i64, i64, C.c_int32, vp, vp, C.c_int32, vp, vp]
    L.scdd_labels_to_dense.argtypes = [vp, i64, i64, C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, vp, vp]
    L.scdd_cluster_counts.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp]
    L.scdd_fedp.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int32, vp, vp, vp, C.POINTER(Stats)]
    L.scdd_selftest_student_p.argtypes = [vp, vp, C.c_int32, vp]
    L.scdd_default_classify_opts.argtypes = [C.POINTER(ClassifyOpts)]
    L.scdd_default_classify_opts.restype = None
    L.scdd_classify.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int32, vp, C.POINTER(Cfg), C.POINTER(ClassifyOpts), vp,
                                C.POINTER(Stats)]
    L.scdd_selftest_student_t.argtypes = [C.c_double, C.c_int32, C.c_uint64, vp]
    L.scdd_default_glm_opts.argtypes = [C.POINTER(GlmOpts)]
    L.scdd_default_glm_opts.restype = None
    L.scdd_test_zeroes.argtypes = [vp, i64, i64, C.c_int32, vp, vp, i64, C.POINTER(GlmOpts), vp, vp, C.POINTER(Stats)]
    _LIB = L
    return L


def last_error():
    return lib().scdd_last_error().decode()


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise ScddError(rc, last_error())
    return rc


def ptr(a):
    return None if a is None else a.ctypes.data


def make_cfg(alpha=0.10, mu0=0.0, s0=0.01, a0=0.01, b0=0.01, min_size=3, restrict=True, extra_mstep=True,
             ks_old_pvalue=False):
    c = Cfg()
    lib().scdd_default_cfg(C.byref(c))
    c.alpha, c.m0, c.s0, c.a0, c.b0 = alpha, mu0, s0, a0, b0
    c.min_size = int(min_size)
    c.restrict_ = int(restrict)
    c.extra_mstep = int(extra_mstep)
    c.ks_old_pvalue = int(ks_old_pvalue)
    return c
