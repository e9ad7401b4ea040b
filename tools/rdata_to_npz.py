#!/usr/bin/env python3
"""Convert the reference's bundled example data (`data/*.RData`) into small
`.npz` fixtures under `tests/golden/`.

R is not installed in the build container, so this is a minimal reader for the
R serialization format (XDR, `RDX2`/`RDX3`), just enough for the three bundled
objects (SingleCellExperiment S4 objects and a list of matrices).  It is a
build-time tool: it reads `/root/reference/data` (which does not exist on the
GPU box) and writes fixtures that DO travel with the repo.

Run:  python tools/rdata_to_npz.py [/root/reference/data] [tests/golden]

Layout facts used: the expression matrix of a
SingleCellExperiment lives at @assays@data@listData[["normcounts"]] (REALSXP,
column-major, dim = genes x cells); the condition vector at
@colData@listData[["condition"]].
"""
import bz2
import lzma
import gzip
import struct
import sys
import os

import numpy as np


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v


class RObj:
    """Value plus R attributes (attributes is a dict tag -> value)."""

    def __init__(self, value, attrs=None, kind=None):
        self.value = value
        self.attrs = attrs or {}
        self.kind = kind

    def __repr__(self):
        return "RObj(%s, attrs=%s)" % (self.kind, list(self.attrs))


def _pairlist_to_dict(pl):
    out = {}
    for tag, val in pl:
        out[tag] = val
    return out


def _read_item(r):
    flags = r.i32()
    t = flags & 0xFF
    has_attr = bool(flags & (1 << 9))
    has_tag = bool(flags & (1 << 10))

    if t == 254:   # NILVALUE
        return None
    if t == 253 or t == 242 or t == 241 or t == 251 or t == 250:
        return RObj(None, kind="env%d" % t)
    if t == 255:   # REFSXP
        idx = flags >> 8
        if idx == 0:
            idx = r.i32()
        return r.refs[idx - 1]
    if t == 1:     # SYMSXP
        name = _read_item(r)
        r.refs.append(name)
        return name
    if t == 249 or t == 247:   # NAMESPACESXP / PACKAGESXP
        r.i32()
        n = r.i32()
        info = [_read_item(r) for _ in range(n)]
        o = RObj(info, kind="namespace")
        r.refs.append(o)
        return o
    if t == 4:     # ENVSXP
        o = RObj(None, kind="env")
        r.refs.append(o)
        r.i32()  # locked
        _read_item(r)  # enclos
        _read_item(r)  # frame
        _read_item(r)  # hashtab
        _read_item(r)  # attrib
        return o
    if t in (2, 6, 5):   # LISTSXP / LANGSXP / PROMSXP : pairlist, iterative on CDR
        items = []
        attrs = None
        while True:
            if has_attr:
                attrs = _read_item(r)
            tag = _read_item(r) if has_tag else None
            car = _read_item(r)
            items.append((tag, car))
            # CDR
            flags = r.i32()
            t2 = flags & 0xFF
            if t2 == 254:
                break
            if t2 not in (2, 6, 5):
                # dotted tail; unread the header by re-parsing in place
                r.p -= 4
                items.append((None, _read_item(r)))
                break
            has_attr = bool(flags & (1 << 9))
            has_tag = bool(flags & (1 << 10))
        return items
    if t == 9:     # CHARSXP
        n = r.i32()
        if n == -1:
            return None
        return r.raw(n).decode("utf-8", "replace")
    if t == 238:   # ALTREP
        info = _read_item(r)
        state = _read_item(r)
        _attr = _read_item(r)
        cls = info[0][1] if info else None
        # compact_intseq / compact_realseq: state = (n, start, step)
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = [float(x) for x in state.value]
            arr = start + step * np.arange(int(n))
            if cls == "compact_intseq":
                arr = arr.astype(np.int32)
            return RObj(arr, kind=cls)
        if cls == "deferred_string":
            # state is pairlist (arg, scalar)
            arg = state[0][1]
            vals = arg.value
            return RObj([("%d" % v if float(v).is_integer() else repr(v)) for v in vals], kind="str")
        if cls in ("wrap_real", "wrap_integer", "wrap_logical", "wrap_string"):
            return state.value[0] if isinstance(state, RObj) else state[0][1]
        raise NotImplementedError("ALTREP class %r" % (cls,))

    # vectors
    if t == 10 or t == 13:
        n = r.length()
        v = np.frombuffer(r.raw(4 * n), dtype=">i4").astype(np.int32)
        o = RObj(v, kind="lgl" if t == 10 else "int")
    elif t == 14:
        n = r.length()
        v = np.frombuffer(r.raw(8 * n), dtype=">f8").astype(np.float64)
        o = RObj(v, kind="real")
    elif t == 16:
        n = r.length()
        o = RObj([_read_item(r) for _ in range(n)], kind="str")
    elif t == 19 or t == 20:
        n = r.length()
        o = RObj([_read_item(r) for _ in range(n)], kind="list")
    elif t == 25:  # S4SXP: attributes only
        o = RObj(None, kind="S4")
    elif t == 24:  # RAWSXP
        n = r.length()
        o = RObj(r.raw(n), kind="raw")
    elif t == 3:   # CLOSXP
        attrs = _read_item(r) if has_attr else None
        env = _read_item(r)
        formals = _read_item(r)
        body = _read_item(r)
        return RObj(None, kind="closure")
    else:
        raise NotImplementedError("SEXP type %d at offset %d" % (t, r.p))
    if has_attr:
        a = _read_item(r)
        o.attrs = _pairlist_to_dict(a) if a else {}
    return o


def read_rdata(path):
    raw = open(path, "rb").read()
    for dec in (lzma.decompress, bz2.decompress, gzip.decompress):
        try:
            raw = dec(raw)
            break
        except Exception:
            continue
    assert raw[:5] in (b"RDX3\n", b"RDX2\n"), raw[:8]
    ver3 = raw[:5] == b"RDX3\n"
    r = _Reader(raw)
    r.p = 5
    assert r.raw(2) == b"X\n"
    r.i32(); r.i32(); r.i32()
    if ver3:
        n = r.i32()
        r.raw(n)
    top = _read_item(r)
    return _pairlist_to_dict(top)


def _slot(o, name):
    return o.attrs[name]


def _named_list(o):
    names = o.attrs.get("names")
    if names is None:
        return {}
    return dict(zip(names.value, o.value))


def sce_to_arrays(sce):
    assays = _slot(sce, "assays")
    data = _slot(assays, "data")
    ld = _named_list(_slot(data, "listData"))
    m = ld["normcounts"]
    dim = m.attrs["dim"].value
    mat = np.asarray(m.value, dtype=np.float64).reshape((int(dim[1]), int(dim[0]))).T.copy()
    dimnames = m.attrs.get("dimnames")
    genes = np.array(dimnames.value[0].value) if dimnames is not None and dimnames.value[0] is not None else \
        np.array(["g%d" % i for i in range(mat.shape[0])])
    coldata = _slot(sce, "colData")
    cl = _named_list(_slot(coldata, "listData"))
    cond = np.asarray(cl["condition"].value, dtype=np.float64)
    rn = coldata.attrs.get("rownames")
    cells = np.array(rn.value) if rn is not None and rn.value is not None else \
        np.array(["c%d" % i for i in range(mat.shape[1])])
    return mat, cond, genes, cells


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    for name in ("scDatExSim", "scDatEx"):
        top = read_rdata(os.path.join(src, name + ".RData"))
        mat, cond, genes, cells = sce_to_arrays(top[name])
        print(name, mat.shape, "cond", np.unique(cond, return_counts=True),
              "zero frac %.3f" % (mat == 0).mean())
        np.savez_compressed(os.path.join(dst, name + ".npz"), normcounts=mat, condition=cond,
                            genes=genes, cells=cells)
    top = read_rdata(os.path.join(src, "scDatExList.RData"))
    lst = _named_list(top["scDatExList"])
    out = {}
    for k, m in lst.items():
        dim = m.attrs["dim"].value
        out[k] = np.asarray(m.value, dtype=np.float64).reshape((int(dim[1]), int(dim[0]))).T.copy()
        print("scDatExList", k, out[k].shape)
    np.savez_compressed(os.path.join(dst, "scDatExList.npz"), **out)


if __name__ == "__main__":
    main()
