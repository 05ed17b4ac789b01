"""Minimal reader for R `.rda` files (gzip/bzip2/xz, XDR serialization v2/v3).

TEST INFRASTRUCTURE ONLY. Used by `make_golden.py` to turn the reference's bundled fixture
(`/root/reference/data/example_sce.rda`, `segList_ensemblGeneID.rda`; described in the reference at
`R/data.R:1-23`) into the small `.npz` fixtures committed next to this file.  R itself is not
available in this environment, so the serialization format (R internals manual, "Serialization
Formats") is decoded directly.  Nothing in the product path imports this module.

Objects are returned as `RObj(kind, value, attr)`:
  vectors -> numpy arrays / lists of str, pairlists -> list of (tag, value), S4/env -> dict-like.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np


class RObj:
    __slots__ = ("kind", "value", "attr")

    def __init__(self, kind, value=None, attr=None):
        self.kind, self.value, self.attr = kind, value, attr

    def attrs(self) -> dict:
        out = {}
        if self.attr is not None and self.attr.kind == "pairlist":
            for tag, val in self.attr.value:
                out[tag] = val
        return out

    def __repr__(self):
        return f"RObj({self.kind})"


_REFSXP, _NILVALUE, _GLOBALENV, _UNBOUND, _MISSING, _BASENS = 255, 254, 253, 252, 251, 250
_NAMESPACE, _PACKAGE, _PERSIST, _CLASSREF, _GENERICREF = 249, 248, 247, 246, 245
_BCREPDEF, _BCREPREF, _EMPTYENV, _BASEENV, _ATTRLANG, _ATTRLIST, _ALTREP = 244, 243, 242, 241, 240, 239, 238


class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.p, self.refs = buf, 0, []

    def int(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self) -> int:
        n = self.int()
        if n == -1:
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def array(self, dtype: str, n: int) -> np.ndarray:
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.p)
        self.p += n * a.itemsize
        return a.astype(a.dtype.newbyteorder("="))

    # ---- items -------------------------------------------------------------------------
    def item(self) -> RObj:
        flags = self.int()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == _NILVALUE:
            return RObj("null")
        if t in (_GLOBALENV, _EMPTYENV, _BASEENV, _BASENS, _UNBOUND, _MISSING):
            return RObj("special", t)
        if t == _REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t in (_NAMESPACE, _PACKAGE, _PERSIST):
            self.int()
            n = self.int()
            o = RObj("namespace", [self.item().value for _ in range(n)])
            self.refs.append(o)
            return o
        if t == 1:  # SYMSXP
            o = RObj("symbol", self.item().value)
            self.refs.append(o)
            return o
        if t in (2, 6, 3, 5, 17, _ATTRLANG, _ATTRLIST):  # pairlist-like
            items, attr0, first = [], None, True
            while True:
                attr = self.item() if has_attr else None
                tag = self.item().value if has_tag else None
                if first:
                    attr0, first = attr, False
                items.append((tag, self.item()))
                flags = self.int()
                t2 = flags & 0xFF
                if t2 not in (2, 6):
                    self.p -= 4
                    tail = self.item()
                    if tail.kind != "null":
                        items.append(("__cdr__", tail))
                    break
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
            return RObj("pairlist", items, attr0)
        if t == 4:  # ENVSXP
            self.int()
            o = RObj("env", {})
            self.refs.append(o)
            enclos, frame, hashtab, attr = self.item(), self.item(), self.item(), self.item()
            del enclos
            if frame.kind == "pairlist":
                for tag, val in frame.value:
                    o.value[tag] = val
            if hashtab.kind == "list":
                for bucket in hashtab.value:
                    if bucket.kind == "pairlist":
                        for tag, val in bucket.value:
                            o.value[tag] = val
            o.attr = attr if attr.kind != "null" else None
            return o
        if t == 22:  # EXTPTRSXP
            o = RObj("extptr")
            self.refs.append(o)
            self.item(), self.item()
            if has_attr:
                o.attr = self.item()
            return o
        if t == 23:  # WEAKREFSXP
            o = RObj("weakref")
            self.refs.append(o)
            if has_attr:
                o.attr = self.item()
            return o
        if t in (7, 8):
            o = RObj("builtin", self.raw(self.int()).decode())
        elif t == 9:  # CHARSXP
            n = self.int()
            o = RObj("char", None if n == -1 else self.raw(n).decode("utf-8", "replace"))
        elif t in (10, 13):
            o = RObj("logical" if t == 10 else "integer", self.array(">i4", self.length()))
        elif t == 14:
            o = RObj("real", self.array(">f8", self.length()))
        elif t == 15:
            o = RObj("complex", self.array(">c16", self.length()))
        elif t == 16:
            o = RObj("character", [self.item().value for _ in range(self.length())])
        elif t in (19, 20):
            o = RObj("list", [self.item() for _ in range(self.length())])
        elif t == 24:
            o = RObj("raw", self.raw(self.length()))
        elif t == 25:
            o = RObj("S4")
        elif t == 21:  # BCODESXP
            reps = [None] * self.int()
            o = self._bc1(reps)
        elif t == _ALTREP:
            info, state, attr = self.item(), self.item(), self.item()
            o = self._altrep(info, state)
            if attr.kind != "null":
                o.attr = attr
            return o
        elif t in (_CLASSREF, _GENERICREF):
            raise NotImplementedError("class/generic refs")
        else:
            raise NotImplementedError(f"SEXP type {t} at {self.p}")
        if has_attr:
            o.attr = self.item()
        return o

    def _bc1(self, reps) -> RObj:
        code = self.item()
        consts = []
        for _ in range(self.int()):
            t = self.int()
            if t == 21:
                consts.append(self._bc1(reps))
            elif t in (6, 2, _BCREPDEF, _BCREPREF, _ATTRLANG, _ATTRLIST):
                consts.append(self._bclang(t, reps))
            else:
                consts.append(self.item())
        return RObj("bytecode", (code, consts))

    def _bclang(self, t, reps) -> RObj:
        if t == _BCREPREF:
            return reps[self.int()]
        if t in (_BCREPDEF, 6, 2, _ATTRLANG, _ATTRLIST):
            pos = -1
            if t == _BCREPDEF:
                pos, t = self.int(), self.int()
            o = RObj("bclang", [])
            if pos >= 0:
                reps[pos] = o
            if t in (_ATTRLANG, _ATTRLIST):
                o.attr = self.item()
            tag = self.item()
            car = self._bclang(self.int(), reps)
            cdr = self._bclang(self.int(), reps)
            o.value = [tag, car, cdr]
            return o
        return self.item()

    @staticmethod
    def _altrep(info: RObj, state: RObj) -> RObj:
        cls = info.value[0][1].value if info.kind == "pairlist" else None
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = state.value
            seq = start + step * np.arange(int(n))
            return RObj("integer" if cls == "compact_intseq" else "real",
                        seq.astype(np.int32 if cls == "compact_intseq" else np.float64))
        if cls in ("wrap_real", "wrap_integer", "wrap_logical", "wrap_string", "wrap_list", "wrap_complex", "wrap_raw"):
            return state.value[0][1] if state.kind == "pairlist" else state.value[0]
        if cls == "deferred_string":
            src = state.value[0][1] if state.kind == "pairlist" else state.value[0]
            return RObj("character", [str(x) for x in src.value])
        raise NotImplementedError(f"ALTREP class {cls}")


def read_rda(path: str) -> dict:
    """Return {object name: RObj} for every object saved in an `.rda` file."""
    blob = open(path, "rb").read()
    for mod in (gzip, bz2, lzma):
        try:
            blob = mod.decompress(blob)
            break
        except Exception:
            continue
    if blob[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError("not an RDX2/RDX3 file")
    r = _Reader(blob)
    r.p = 5
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR serialization is supported")
    version = r.int()
    r.int(), r.int()
    if version == 3:
        r.raw(r.int())
    top = r.item()
    return {tag: val for tag, val in top.value}


def factor_labels(obj: RObj) -> list:
    levels = obj.attrs()["levels"].value
    return [levels[i - 1] for i in obj.value]
