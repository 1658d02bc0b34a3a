"""TEST INFRASTRUCTURE ONLY (oracle/): reader for R's `.rda` fixtures.

The reference ships its fixtures as bzip2-compressed R serialisation v2 (XDR)
files: /root/reference/data/{pbmc,pbmcPlot,bmmc}.rda (described in
/root/reference/R/data.R:1-24).  There is no R in the build container, so this
module decodes exactly the subset of SEXP types those three files use and
returns plain Python objects.  It is used only by tests/golden/make_golden.py
(run in the build container, where /root/reference exists) -- never by the
product path.

Returned representation
  * atomic vectors  -> numpy arrays wrapped in `RObj(value, attrs)`
  * generic vectors -> `RObj(list, attrs)`
  * S4 objects      -> `RObj(None, attrs)` with slots in `attrs`
  * pairlists       -> dict (tag -> value) when tagged, else list
"""
import bz2
import gzip
import lzma
import struct

import numpy as np


class RObj:
    __slots__ = ("value", "attrs")

    def __init__(self, value, attrs=None):
        self.value = value
        self.attrs = attrs or {}

    def __getitem__(self, key):
        if isinstance(key, str):
            if key in self.attrs:
                return self.attrs[key]
            names = self.attrs.get("names")
            if names is not None:
                nm = list(names.value)
                return self.value[nm.index(key)]
            raise KeyError(key)
        return self.value[key]

    def names(self):
        n = self.attrs.get("names")
        return list(n.value) if n is not None else None

    def __repr__(self):
        v = self.value
        if isinstance(v, np.ndarray):
            v = f"ndarray{v.shape}{v.dtype}"
        elif isinstance(v, list):
            v = f"list[{len(v)}]"
        return f"RObj({v}, attrs={list(self.attrs)})"


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def int(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.int()
        if n == -1:
            hi = self.int()
            lo = self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == 254:      # NILVALUE_SXP
            return None
        if t in (242, 241, 253, 252, 251, 247):   # env singletons / missing
            return ("<special>", t)
        if t == 255:      # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (249, 250, 248):  # NAMESPACESXP / PACKAGESXP / PERSISTSXP
            self.int()
            n = self.int()
            info = [self.item() for _ in range(n)]
            obj = ("<namespace>", info)
            self.refs.append(obj)
            return obj
        if t == 4:        # ENVSXP
            self.int()
            env = {"__env__": True}
            self.refs.append(env)
            env["enclos"] = self.item()
            env["frame"] = self.item()
            env["hashtab"] = self.item()
            env["attrib"] = self.item()
            return env
        if t in (2, 6, 239, 240):   # LISTSXP / LANGSXP (+ attr variants)
            return self.pairlist(flags)
        if t == 9:        # CHARSXP
            n = self.int()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", "replace")
        if t == 238:      # ALTREP
            info = self.item()
            state = self.item()
            attr = self.item()
            return self.altrep(info, state, attr)
        if t == 25:       # S4SXP
            attrs = self.item() if has_attr else {}
            return RObj(None, attrs if isinstance(attrs, dict) else {})
        if t in (10, 13):  # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:      # REALSXP
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif t == 16:      # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t in (19, 20):  # VECSXP / EXPRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == 24:      # RAWSXP
            n = self.length()
            v = np.frombuffer(self.raw(n), dtype=np.uint8).copy()
        elif t == 3:       # CLOSXP: attr?, env, formals, body
            attrs = self.item() if has_attr else {}
            self.item(); self.item(); self.item()
            return ("<closure>",)
        elif t == 21:      # BCODESXP -- not present in the fixtures
            raise NotImplementedError("bytecode")
        elif t == 22:      # EXTPTRSXP
            obj = ("<extptr>",)
            self.refs.append(obj)
            self.item(); self.item()
            if has_attr:
                self.item()
            return obj
        else:
            raise NotImplementedError(f"SEXP type {t} at {self.p}")
        attrs = self.item() if has_attr else {}
        return RObj(v, attrs if isinstance(attrs, dict) else {})

    def pairlist(self, flags):
        # iterative over cdr to avoid deep recursion
        tagged = {}
        untagged = []
        any_tag = False
        while True:
            t = flags & 0xFF
            if t == 254:
                break
            if t not in (2, 6, 239, 240):
                # dotted pair tail; rewind not possible -> treat as value
                raise NotImplementedError("dotted pairlist")
            has_attr = bool(flags & 0x200) or t in (239, 240)
            has_tag = bool(flags & 0x400)
            if has_attr:
                self.item()
            tag = self.item() if has_tag else None
            car = self.item()
            if tag is not None:
                any_tag = True
                tagged[tag] = car
            untagged.append(car)
            flags = self.int()
        return tagged if any_tag else untagged

    def altrep(self, info, state, attr):
        cls = info[0] if isinstance(info, list) else None
        attrs = attr if isinstance(attr, dict) else {}
        if cls == "compact_intseq":
            n, start, step = (int(x) for x in state.value)
            return RObj(np.arange(start, start + n * step, step, dtype=np.int32), attrs)
        if cls == "compact_realseq":
            n, start, step = state.value
            return RObj(start + step * np.arange(int(n), dtype=np.float64), attrs)
        if cls in ("wrap_real", "wrap_integer", "wrap_logical", "wrap_string", "wrap_list"):
            inner = state[0] if isinstance(state, (list, RObj)) else state
            if isinstance(inner, RObj):
                merged = dict(inner.attrs)
                merged.update(attrs)
                return RObj(inner.value, merged)
            return inner
        if cls == "deferred_string":
            src = state[0] if isinstance(state, list) else state
            vals = src.value
            return RObj([str(int(x)) if float(x).is_integer() else repr(float(x)) for x in vals], attrs)
        raise NotImplementedError(f"ALTREP class {cls}")


def read_rda(path):
    """Return {name: object} for an .rda file."""
    raw = open(path, "rb").read()
    for mod in (bz2, gzip, lzma):
        try:
            raw = mod.decompress(raw)
            break
        except Exception:
            continue
    assert raw[:5] == b"RDX2\n", "not an RDX2 file"
    assert raw[5:7] == b"X\n", "only XDR serialisation supported"
    r = _Reader(raw)
    r.p = 7
    version = r.int()
    r.int()
    r.int()
    if version == 3:
        n = r.int()
        r.raw(n)
    top = r.item()
    assert isinstance(top, dict)
    return top


def dgc_to_csc(obj):
    """dgCMatrix S4 -> (indptr int32, indices int32, data float64, shape, dimnames)."""
    a = obj.attrs
    dim = tuple(int(x) for x in a["Dim"].value)
    dn = a.get("Dimnames")
    dimnames = [None, None]
    if dn is not None:
        for i in range(2):
            e = dn.value[i]
            dimnames[i] = list(e.value) if e is not None else None
    return (a["p"].value.astype(np.int32), a["i"].value.astype(np.int32),
            a["x"].value.astype(np.float64), dim, dimnames)


def dense(obj):
    """R matrix (REALSXP + dim, column-major) -> (ndarray[rows, cols], dimnames)."""
    dim = tuple(int(x) for x in obj.attrs["dim"].value)
    m = np.asarray(obj.value, dtype=np.float64).reshape(dim, order="F")
    dn = obj.attrs.get("dimnames")
    dimnames = [None, None]
    if dn is not None:
        for i in range(2):
            e = dn.value[i]
            dimnames[i] = list(e.value) if e is not None else None
    return m, dimnames
