"""The R `.Call` glue (r/src/scdd_glue.c) compiled and EXECUTED against a mock of R's C API (tests/mock_R).

R exists neither in the build image nor on the GPU box, so this is how the glue meets a compiler: gcc -Wall -Wextra
-Werror, the routine table's arities against the definitions, every routine called with R-shaped arguments, PROTECT
balance after every call, the error paths (wrong vector types -- e.g. a double vector where INTEGER() is applied --
no device, interrupt) and, on a GPU, the same numbers as the ctypes binding of the same library.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from scdd_b200 import _lib, hotpath as hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_R")
GLUE = os.path.join(ROOT, "r", "src", "scdd_glue.c")
OUT = os.path.join(MOCK, "_build", "scdd_glue_mock.so")
PRI = np.array([0.01, 0.0, 0.01, 0.01, 0.01])


@pytest.fixture(scope="module")
def glue():
    lib = _lib.lib_path()
    _lib.lib()                                               # builds libscdd_b200.so if needed
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [GLUE, os.path.join(MOCK, "mock_runtime.c")]
    if not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs + [lib]):
        cmd = ["gcc", "-std=gnu11", "-O1", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-fPIC", "-shared",
               "-I", MOCK, "-I", os.path.join(ROOT, "include"), "-o", OUT] + srcs + \
              ["-L", os.path.dirname(lib), "-lscdd_b200", "-Wl,-rpath," + os.path.dirname(lib)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
    L = C.CDLL(OUT)
    vp = C.c_void_p
    for name, res, args in [("mock_real", vp, [vp, C.c_long]), ("mock_int", vp, [vp, C.c_long]),
                            ("mock_real_matrix", vp, [vp, C.c_int, C.c_int]), ("mock_int_matrix", vp, [vp, C.c_int, C.c_int]),
                            ("mock_lgl", vp, [C.c_int]), ("mock_nil", vp, []), ("mock_type", C.c_int, [vp]),
                            ("mock_len", C.c_long, [vp]), ("mock_nrow", C.c_int, [vp]), ("mock_ncol", C.c_int, [vp]),
                            ("mock_data", vp, [vp]), ("mock_list_get", vp, [vp, C.c_char_p]),
                            ("mock_protect_depth", C.c_int, []), ("mock_protect_underflow", C.c_int, []),
                            ("mock_last_error", C.c_char_p, []), ("mock_interrupt_after", None, [C.c_int]),
                            ("mock_interrupt_checks", C.c_int, []), ("mock_registered", C.c_int, [C.c_char_p]),
                            ("mock_call", vp, [C.c_char_p, C.c_int, C.POINTER(vp)])]:
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return R(L)


class R:
    """just enough of an R session: vectors in, .Call, vectors out"""

    def __init__(self, L):
        self.L = L

    def real(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return self.L.mock_real(a.ctypes.data, a.size)

    def int(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return self.L.mock_int(a.ctypes.data, a.size)

    def matrix(self, a):                                       # R matrices are column-major
        a = np.asfortranarray(a, dtype=np.float64)
        return self.L.mock_real_matrix(a.ctypes.data, a.shape[0], a.shape[1])

    def lgl(self, v):
        return self.L.mock_lgl(int(v))

    @property
    def nil(self):
        return self.L.mock_nil()

    def call(self, name, *args):
        arr = (C.c_void_p * max(1, len(args)))(*args)
        out = self.L.mock_call(name.encode(), len(args), arr)
        assert self.L.mock_protect_depth() == 0 and not self.L.mock_protect_underflow(), "PROTECT imbalance in " + name
        if not out:
            raise RuntimeError(self.L.mock_last_error().decode())
        return out

    def value(self, s):
        t, n = self.L.mock_type(s), self.L.mock_len(s)
        if t == 0:
            return None
        dt = {13: np.int32, 10: np.int32, 14: np.float64, 24: np.uint8}[t]
        a = np.ctypeslib.as_array(C.cast(self.L.mock_data(s), C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()
        nr, nc = self.L.mock_nrow(s), self.L.mock_ncol(s)
        return a.reshape((nc, nr)).T if nr * nc == n and nr > 0 else a

    def get(self, lst, name):
        s = self.L.mock_list_get(lst, name.encode())
        assert s, name
        return self.value(s)


def _example(golden_dir, ngenes=12):
    d = np.load(os.path.join(golden_dir, "scDatExSim.npz"), allow_pickle=True)
    return d["normcounts"][:ngenes], d["condition"]


def test_registered_arities_match_the_definitions(glue):
    src = re.sub(r"/\*.*?\*/", "", open(GLUE).read(), flags=re.S)
    defs = {m.group(1): (0 if m.group(2).strip() in ("void", "") else m.group(2).count("SEXP"))
            for m in re.finditer(r"^SEXP (C_scdd_\w+)\(([^)]*)\)\s*\{", src, flags=re.M)}
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(C_scdd_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', src))
    assert defs == table and len(defs) >= 10
    for name, n in defs.items():
        assert glue.L.mock_registered(name.encode()) == n
    assert glue.L.mock_registered(b"C_nope") == -1


def test_host_side_routines_match_the_ctypes_binding(glue, golden_dir):
    X, cond = _example(golden_dir)
    c01 = (cond != cond[0]).astype(np.int32)
    out = glue.call("C_scdd_gene_csr", glue.matrix(X), glue.int([X.shape[0]]), glue.int([X.shape[1]]), glue.int(c01),
                    glue.lgl(True), glue.lgl(False))
    vals, c, off = glue.get(out, "vals"), glue.get(out, "cond01"), glue.get(out, "off")
    rv, rc, ro = hp.gene_csr(X, cond)
    assert np.array_equal(vals, rv) and np.array_equal(c, rc) and np.array_equal(off, ro.astype(np.float64))
    labels = (np.arange(vals.size) % 3 + 1).astype(np.int32)
    for which in (-1, 0, 1):
        cnt = glue.value(glue.call("C_scdd_cluster_counts", glue.int(labels), glue.int(c), glue.real(off), glue.int([3]),
                                   glue.int([which])))
        assert np.array_equal(cnt, hp.cluster_counts(labels, rc, ro, min_size=3, which=which))
        Z = glue.value(glue.call("C_scdd_zhat", glue.matrix(X), glue.int([X.shape[0]]), glue.int([X.shape[1]]), glue.int(c01),
                                 glue.int([which]), glue.int(np.arange(1, X.shape[0] + 1)), glue.real(off), glue.int(labels)))
        assert np.array_equal(Z, hp.labels_to_dense(X, cond, labels, ro, which=which))
    sc = glue.call("C_scdd_dense_scan", glue.matrix(X), glue.int([X.shape[0]]), glue.int([X.shape[1]]), glue.int(c01))
    n0, n1, k0, k1, neg = hp.dense_scan(X, cond)
    assert np.array_equal(glue.get(sc, "nnz0"), n0) and np.array_equal(glue.get(sc, "nnz1"), n1)
    assert np.array_equal(glue.get(sc, "const0").astype(bool), k0) and glue.get(sc, "n.negative")[0] == neg
    Cv = (X > 0).mean(axis=0)
    cd = glue.value(glue.call("C_scdd_cell_covariate", glue.matrix(X), glue.int([X.shape[0]]), glue.int([X.shape[1]]),
                              glue.real(Cv), glue.real(off)))
    assert np.array_equal(cd, np.broadcast_to(Cv, X.shape)[X > 0]) and np.array_equal(cd, hp.cell_covariate(X, Cv, ro))
    assert glue.value(glue.call("C_scdd_device_count"))[0] == hp.device_count()
    glue.call("C_scdd_release_cache")


def test_argument_type_errors_are_r_errors_not_crashes(glue, golden_dir):
    X, cond = _example(golden_dir, 4)
    vals, c, off = hp.gene_csr(X, cond)
    good = dict(logy=glue.real(vals), off=glue.real(off), c=glue.int(c), B=glue.int([2]), pri=glue.real(PRI), ms=glue.int([3]))
    # a double vector where the glue applies INTEGER()
    with pytest.raises(RuntimeError, match="perm_idx must be integer"):
        glue.call("C_scdd_perm_bf", good["logy"], good["off"], good["c"], good["B"], glue.real(np.ones(vals.size * 2)),
                  glue.nil, glue.lgl(False), glue.nil, good["pri"], good["ms"])
    with pytest.raises(RuntimeError, match="seed6 must be a 6 x 4"):
        glue.call("C_scdd_perm_bf", good["logy"], good["off"], good["c"], good["B"], glue.nil, glue.int(np.zeros(6)),
                  glue.lgl(False), glue.nil, good["pri"], good["ms"])
    with pytest.raises(RuntimeError, match="cond01 must be integer"):
        glue.call("C_scdd_ks", good["logy"], good["off"], glue.real(c.astype(float)))
    with pytest.raises(RuntimeError, match="prior must be"):
        glue.call("C_scdd_fit_observed", good["logy"], good["off"], good["c"], glue.real(PRI[:3]), good["ms"], glue.lgl(True))
    with pytest.raises(RuntimeError, match="registered with 3 args"):
        glue.call("C_scdd_ks", good["logy"], good["off"])


def test_compute_routines_fail_loudly_without_gpu(glue, golden_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X, cond = _example(golden_dir, 4)
    vals, c, off = hp.gene_csr(X, cond)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        glue.call("C_scdd_fit_observed", glue.real(vals), glue.real(off), glue.int(c), glue.real(PRI), glue.int([3]), glue.lgl(True))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        glue.call("C_scdd_ks", glue.real(vals), glue.real(off), glue.int(c))


@pytest.mark.gpu
def test_glue_on_gpu_matches_ctypes_binding(glue, golden_dir):
    X, cond = _example(golden_dir, 20)
    vals, c, off = hp.gene_csr(X, cond)
    cfg = hp.make_cfg(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)
    fits, cls_oa, cls_c, bf, den, _ = hp.fit_observed(vals, c, off, cfg=cfg)
    out = glue.call("C_scdd_fit_observed", glue.real(vals), glue.real(off), glue.int(c), glue.real(PRI), glue.int([3]), glue.lgl(True))
    assert np.array_equal(glue.get(out, "G"), fits["G"].ravel())
    assert np.array_equal(glue.get(out, "class.oa"), cls_oa) and np.array_equal(glue.get(out, "class.c"), cls_c)
    assert np.array_equal(glue.get(out, "bf"), bf) and np.array_equal(glue.get(out, "den"), den)
    assert np.array_equal(glue.get(out, "bic"), fits["bic"].ravel())
    m = glue.get(out, "mean")                                  # GMAX x (3 * ngenes)
    assert np.array_equal(m[:, 4][:fits[1, 1]["G"]], fits[1, 1]["mean"][:fits[1, 1]["G"]])
    # permutations: per-gene streams, per-permutation streams, adjusted
    B = 20
    seeds = hp.gene_seeds(5, 20)
    ref, _ = hp.perm_bf(vals, c, off, B, seed6=seeds, cfg=cfg)
    seed_mat = glue.L.mock_int_matrix(np.ascontiguousarray(seeds.view(np.int32)).ctypes.data, 6, 20)   # 6 x ngenes, column-major
    out = glue.call("C_scdd_perm_bf", glue.real(vals), glue.real(off), glue.int(c), glue.int([B]), glue.nil, seed_mat,
                    glue.lgl(False), glue.nil, glue.real(PRI), glue.int([3]))
    assert np.array_equal(glue.get(out, "bf.perm"), ref.T)     # B x ngenes: column g = bf.perm[[g]]
    assert glue.get(out, "failed.sets")[0] == 0
    pseeds = hp.gene_seeds(5, B)
    ref1, _ = hp.perm_bf(vals, c, off, B, seed6=pseeds, seed_mode=1, cfg=cfg)
    out = glue.call("C_scdd_perm_bf", glue.real(vals), glue.real(off), glue.int(c), glue.int([B]), glue.nil,
                    glue.L.mock_int_matrix(np.ascontiguousarray(pseeds.view(np.int32)).ctypes.data, 6, B),
                    glue.lgl(True), glue.nil, glue.real(PRI), glue.int([3]))
    assert np.array_equal(glue.get(out, "bf.perm"), ref1.T)
    # KS and feDP
    rv, rc, ro = hp.gene_csr(X, cond, log=False)
    ks = glue.call("C_scdd_ks", glue.real(rv), glue.real(ro), glue.int(rc))
    D, p = hp.ks(rv, rc, ro)
    assert np.array_equal(glue.get(ks, "D"), D) and np.array_equal(glue.get(ks, "p"), p)
    fe = glue.call("C_scdd_fedp", glue.real(vals), glue.real(off), glue.int(c), glue.int(cls_oa), glue.int(cls_c), glue.nil, glue.int([3]))
    ps, pf, cnt, _ = hp.fedp(vals, c, off, cls_oa, cls_c)
    assert np.array_equal(glue.get(fe, "p.shift"), ps, equal_nan=True) and np.array_equal(glue.get(fe, "counts"), cnt.T)
    # classifyDD: 1-based genes in, category codes out -- the same call through ctypes gives the same labels
    sel = np.array([2, 5, 11, 20])
    cd = glue.call("C_scdd_classify", glue.real(vals), glue.real(off), glue.int(c), glue.int(cls_oa), glue.int(cls_c),
                   glue.int(sel), glue.real(ps), glue.real(PRI), glue.int([3]), glue.int([7]))
    names = hp.classify(vals, c, off, sel - 1, cls_oa, cls_c, p_shift=ps, cfg=cfg, seed=7)[0]
    assert [("NC", "DE", "DP", "DM", "DB")[k] for k in glue.value(cd)] == names.tolist()
    # testZeroes: 1-based rows in, NA_real_ (payload 1954) for a gene without zeros
    Xz = np.array(X, dtype=np.float64)
    Xz[2] = np.abs(Xz[2]) + 1.0
    these = np.array([1, 3, 20, 7])
    tz = glue.call("C_scdd_test_zeroes", glue.matrix(Xz), glue.int([Xz.shape[0]]), glue.int([Xz.shape[1]]),
                   glue.int(hp.factor_level2(cond)), glue.int(these))
    pz, _ = hp.test_zeroes(Xz, cond, these=these - 1)
    got = glue.value(tz)
    assert np.array_equal(got, pz, equal_nan=True) and np.isnan(got[1])
    assert got[1:2].view(np.uint64)[0] & 0xFFFFFFFF == 1954
    # Ctrl-C: R_CheckUserInterrupt() fires at the second poll; the library unwinds, THEN the glue raises the R error
    glue.L.mock_interrupt_after(1)
    with pytest.raises(RuntimeError, match="interrupted"):
        glue.call("C_scdd_perm_bf", glue.real(vals), glue.real(off), glue.int(c), glue.int([100]), glue.nil, seed_mat,
                  glue.lgl(False), glue.nil, glue.real(PRI), glue.int([3]))
    assert glue.L.mock_interrupt_checks() == 2
    glue.L.mock_interrupt_after(-1)
    glue.call("C_scdd_set_devices", glue.int([0]))


def _strip_r(src):
    """R source without comments and string contents (delimiters kept), good enough for delimiter / call checks"""
    out, i, n = [], 0, len(src)
    while i < n:
        ch = src[i]
        if ch == "#":
            while i < n and src[i] != "\n":
                i += 1
        elif ch in "\"'":
            q = ch
            out.append(q)
            i += 1
            while i < n and src[i] != q:
                i += 2 if src[i] == "\\" else 1
            out.append(q)
            i += 1
        else:
            out.append(ch)
            i += 1
    return "".join(out)


def test_r_files_are_balanced_and_call_registered_routines_with_the_right_arity():
    """No R interpreter exists here; this is the static part of what R CMD check would say: delimiters balance and
    every .Call(C_xxx, ...) names a registered routine and passes as many arguments as its registration declares."""
    glue_src = re.sub(r"/\*.*?\*/", "", open(GLUE).read(), flags=re.S)
    table = dict((m.group(1), int(m.group(2))) for m in re.finditer(r'\{"(C_scdd_\w+)", \(DL_FUNC\)&\w+, (\d+)\}', glue_src))
    calls_seen = set()
    for name in ("scDD_gpu.R", "scDD_b200.R"):
        src = _strip_r(open(os.path.join(ROOT, "r", "R", name)).read())
        stack = []
        pairs = {")": "(", "]": "[", "}": "{"}
        for ch in src:
            if ch in "([{":
                stack.append(ch)
            elif ch in ")]}":
                assert stack and stack.pop() == pairs[ch], f"unbalanced {ch} in {name}"
        assert not stack, f"unclosed {stack[-1]} in {name}"
        for m in re.finditer(r"\.Call\(\s*(C_scdd_\w+)", src):
            routine = m.group(1)
            assert routine in table, f"{name}: {routine} is not registered in scdd_glue.c"
            depth, nargs, i = 1, 0, m.end()
            while depth > 0:
                ch = src[i]
                if ch in "([{":
                    depth += 1
                elif ch in ")]}":
                    depth -= 1
                elif ch == "," and depth == 1:
                    nargs += 1
                i += 1
            assert nargs == table[routine], f"{name}: .Call({routine}) passes {nargs} arguments, registered with {table[routine]}"
            calls_seen.add(routine)
    assert calls_seen == set(table), f"registered but never called from R: {set(table) - calls_seen}"
