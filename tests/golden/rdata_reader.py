"""Minimal reader for R's XDR serialization format (version 2), enough for the
chromVAR fixtures (`data/*.rda`, `inst/extdata/deviations_test`).

Test infrastructure only: it is used once, by `make_golden.py`, inside the build
container where `/root/reference` is mounted, to turn the reference's bundled
fixtures into the small `.npz` files committed next to this script.  There is no R
interpreter in the image, hence the hand-written parser.

Format notes (R Internals, "Serialization Formats"): every item starts with an
int32 of flags: bits 0-7 SEXP type, bit 8 object, bit 9 has-attributes, bit 10
has-tag.  Pairlists carry attr, tag, car, cdr in that order; other types carry
their payload followed by the attribute pairlist.
"""
import bz2
import gzip
import struct

import numpy as np

NILVALUE, GLOBALENV, EMPTYENV, BASEENV, BASENAMESPACE = 254, 253, 242, 241, 247
MISSINGARG, UNBOUND = 251, 252
REFSXP, PERSISTSXP, PACKAGESXP, NAMESPACESXP = 255, 247, 250, 249
BCREPREF, BCREPDEF, ATTRLANGSXP, ATTRLISTSXP = 243, 244, 240, 239
ALTREP = 238


class RObj:
    """Generic node: `kind`, `value`, `attr` (dict)."""

    def __init__(self, kind, value=None, attr=None):
        self.kind = kind
        self.value = value
        self.attr = attr or {}

    def __repr__(self):
        return "RObj(%s, attrs=%s)" % (self.kind, list(self.attr))


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def attrs_to_dict(self, pl):
        d = {}
        while pl is not None and isinstance(pl, RObj) and pl.kind == "pairlist":
            for tag, car in pl.value:
                d[tag] = car
            break
        return d

    def read_pairlist(self, flags, kind="pairlist"):
        # iterative over cdr to avoid deep recursion
        items = []
        first_attr = None
        while True:
            typ = flags & 0xFF
            if typ == NILVALUE:
                break
            if typ not in (2, 6, ATTRLANGSXP, ATTRLISTSXP):
                # dotted tail; keep it as a final item
                items.append((None, self.read_item(flags)))
                break
            has_attr = bool(flags & (1 << 9)) or typ in (ATTRLANGSXP, ATTRLISTSXP)
            has_tag = bool(flags & (1 << 10))
            a = self.read_item() if has_attr else None
            if first_attr is None:
                first_attr = a
            tag = None
            if has_tag:
                t = self.read_item()
                tag = t.value if isinstance(t, RObj) and t.kind == "symbol" else t
            car = self.read_item()
            items.append((tag, car))
            flags = self.i32()
        return RObj(kind, items)

    def read_bc_consts(self, reps):
        n = self.i32()
        out = []
        for _ in range(n):
            typ = self.i32()
            if typ == 21:
                out.append(self.read_bc1(reps))
            elif typ in (2, 6, ATTRLANGSXP, ATTRLISTSXP, BCREPDEF, BCREPREF):
                out.append(self.read_bclang(typ, reps))
            else:
                out.append(self.read_item())
        return out

    def read_bclang(self, typ, reps):
        if typ == BCREPREF:
            return reps[self.i32()]
        if typ in (2, 6, ATTRLANGSXP, ATTRLISTSXP, BCREPDEF):
            pos = -1
            hasattr_ = False
            if typ == BCREPDEF:
                pos = self.i32()
                typ = self.i32()
            if typ in (ATTRLANGSXP, ATTRLISTSXP):
                hasattr_ = True
            node = RObj("bclang", [])
            if pos >= 0:
                reps[pos] = node
            if hasattr_:
                self.read_item()
            tag = self.read_item()
            car = self.read_bclang(self.i32(), reps)
            cdr = self.read_bclang(self.i32(), reps)
            node.value = (tag, car, cdr)
            return node
        return self.read_item()

    def read_bc1(self, reps):
        code = self.read_item()
        consts = self.read_bc_consts(reps)
        return RObj("bytecode", (code, consts))

    def read_item(self, flags=None):
        if flags is None:
            flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        if typ == NILVALUE:
            return None
        if typ in (GLOBALENV, EMPTYENV, BASEENV, BASENAMESPACE, MISSINGARG, UNBOUND):
            return RObj("const%d" % typ)
        if typ == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ in (PACKAGESXP, NAMESPACESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            names = [self.read_item() for _ in range(n)]
            o = RObj("namespace", names)
            self.refs.append(o)
            return o
        if typ == 1:  # SYMSXP
            s = self.read_item()
            o = RObj("symbol", s)
            self.refs.append(o)
            return o
        if typ == 4:  # ENVSXP
            o = RObj("env", {})
            self.refs.append(o)
            self.i32()  # locked
            enclos = self.read_item()
            frame = self.read_item()
            hashtab = self.read_item()
            attr = self.read_item()
            env = {}
            if frame is not None and frame.kind == "pairlist":
                for tag, car in frame.value:
                    env[tag] = car
            if hashtab is not None and hashtab.kind == "list":
                for bucket in hashtab.value:
                    if bucket is not None and bucket.kind == "pairlist":
                        for tag, car in bucket.value:
                            env[tag] = car
            o.value = env
            o.attr = self.attrs_to_dict(attr)
            return o
        if typ in (2, 6, ATTRLANGSXP, ATTRLISTSXP):
            return self.read_pairlist(flags, "pairlist" if typ == 2 else "lang")
        if typ in (3, 5):  # CLOSXP, PROMSXP
            has_tag = bool(flags & (1 << 10))
            a = self.read_item() if has_attr else None
            env = self.read_item() if has_tag else None
            formals = self.read_item()
            body = self.read_item()
            return RObj("closure", (env, formals, body), self.attrs_to_dict(a))
        if typ in (7, 8):  # SPECIALSXP, BUILTINSXP
            n = self.i32()
            return RObj("builtin", self.raw(n).decode())
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", "replace")
        if typ == 21:  # BCODESXP
            nreps = self.i32()
            reps = [None] * nreps
            o = self.read_bc1(reps)
            if has_attr:
                self.read_item()
            return o
        if typ == 25:  # S4SXP
            a = self.read_item() if has_attr else None
            return RObj("S4", None, self.attrs_to_dict(a))
        if typ == 22:  # EXTPTRSXP
            o = RObj("extptr")
            self.refs.append(o)
            self.read_item()
            self.read_item()
            if has_attr:
                self.read_item()
            return o
        # vectors
        if typ == 10 or typ == 13:
            n = self.i32()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
            kind = "lgl" if typ == 10 else "int"
        elif typ == 14:
            n = self.i32()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
            kind = "real"
        elif typ == 16:
            n = self.i32()
            v = [self.read_item() for _ in range(n)]
            kind = "str"
        elif typ in (19, 20):
            n = self.i32()
            v = [self.read_item() for _ in range(n)]
            kind = "list"
        elif typ == 24:
            n = self.i32()
            v = self.raw(n)
            kind = "raw"
        else:
            raise ValueError("unsupported SEXP type %d at offset %d" % (typ, self.o))
        a = self.read_item() if has_attr else None
        return RObj(kind, v, self.attrs_to_dict(a))


def _decompress(path):
    b = open(path, "rb").read()
    if b[:3] == b"BZh":
        return bz2.decompress(b)
    if b[:2] == b"\x1f\x8b":
        return gzip.decompress(b)
    if b[:6] == b"\xfd7zXZ\x00":
        import lzma
        return lzma.decompress(b)
    return b


def read_rds(path):
    b = _decompress(path)
    assert b[:2] == b"X\n", b[:10]
    r = _Reader(b)
    r.o = 2
    version = r.i32()
    r.i32()
    r.i32()
    assert version == 2, version
    return r.read_item()


def read_rda(path):
    """Returns {name: object}."""
    b = _decompress(path)
    assert b[:5] == b"RDX2\n", b[:10]
    r = _Reader(b)
    r.o = 5
    assert r.raw(2) == b"X\n"
    version = r.i32()
    r.i32()
    r.i32()
    assert version == 2
    top = r.read_item()
    return {tag: car for tag, car in top.value}


def walk(obj, pred, path="", seen=None, out=None):
    """Collect (path, node) for every node satisfying pred."""
    if out is None:
        out, seen = [], set()
    if obj is None or id(obj) in seen:
        return out
    if not isinstance(obj, RObj):
        return out
    seen.add(id(obj))
    if pred(obj):
        out.append((path, obj))
    for k, v in obj.attr.items():
        walk(v, pred, path + "@" + str(k), seen, out)
    val = obj.value
    if isinstance(val, dict):
        for k, v in val.items():
            walk(v, pred, path + "$" + str(k), seen, out)
    elif isinstance(val, (list, tuple)):
        for i, v in enumerate(val):
            if isinstance(v, tuple):
                for j, w in enumerate(v):
                    walk(w, pred, path + "[%d.%s]" % (i, v[0] if j == 1 else j), seen, out)
            else:
                walk(v, pred, path + "[%d]" % i, seen, out)
    return out
