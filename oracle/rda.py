"""TEST INFRASTRUCTURE ONLY (oracle side): decode the reference's data/mushroom.rda.

The file is a bzip2 stream holding R's XDR serialisation ("RDX2") of one
pairlist node whose tag is the symbol `mushroom` and whose value is a
data.frame (VECSXP) of 23 character columns x 8124 rows (SURVEY.md Appendix
A.2; documented in /root/reference/R/data.R:1-66).  Only the SEXP types that
occur in that file are handled.  Used by tests/golden/make_golden.py, in THIS
container only -- /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: list = []

    def i32(self) -> int:
        (v,) = struct.unpack_from(">i", self.b, self.p)
        self.p += 4
        return v

    def f64(self) -> float:
        (v,) = struct.unpack_from(">d", self.b, self.p)
        self.p += 8
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.p : self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if typ == 254:  # NILVALUE_SXP
            return None
        if typ == 255:  # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            sym = ("sym", name)
            self.refs.append(sym)
            return sym
        if typ == 2:  # LISTSXP (pairlist): attr, tag, car, cdr
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag[1] if tag else None, car, attr))
                nflags = struct.unpack_from(">i", self.b, self.p)[0]
                ntyp = nflags & 0xFF
                if ntyp == 2:
                    self.p += 4
                    has_attr = bool(nflags & (1 << 9))
                    has_tag = bool(nflags & (1 << 10))
                    continue
                tail = self.item()
                assert tail is None, "dotted pairlists are not expected here"
                return out
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            return self.raw(n).decode("utf-8")
        if typ == 16:  # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif typ == 19:  # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif typ == 13 or typ == 10:  # INTSXP / LGLSXP
            n = self.i32()
            v = [self.i32() for _ in range(n)]
        elif typ == 14:  # REALSXP
            n = self.i32()
            v = [self.f64() for _ in range(n)]
        else:
            raise ValueError(f"unhandled SEXP type {typ} at byte {self.p}")
        attrs = self.item() if has_attr else None
        if attrs:
            return {"value": v, "attr": {t: c for (t, c, _a) in attrs}}
        return v


def read_rda(path: str):
    """Return {object_name: decoded value} for an RDX2 .rda file."""
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    assert raw[:5] == b"RDX2\n", raw[:5]
    assert raw[5:7] == b"X\n", "only the XDR flavour is handled"
    r = _Reader(raw[7:])
    version, _writer, _minreader = r.i32(), r.i32(), r.i32()
    assert version == 2, version
    top = r.item()
    return {tag: car for (tag, car, _a) in top}


def read_mushroom(path: str = "/root/reference/data/mushroom.rda"):
    """-> (column_names, columns) with columns[j] a list of str | None (None = NA)."""
    obj = read_rda(path)["mushroom"]
    names = obj["attr"]["names"]
    cols = obj["value"]
    cols = [c["value"] if isinstance(c, dict) else c for c in cols]
    return list(names), cols
