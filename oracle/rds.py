"""Minimal reader for R's XDR serialization format (.rds / readRDS-style .rda).

TEST INFRASTRUCTURE ONLY (oracle/): used by oracle/make_golden.py to turn the
reference's bundled fixtures (inst/extdata/*.rda, *.rds; loaded by
R/io.R:10-26 `load_a549`) into .npz files under tests/golden/.  Nothing in the
product package imports this module.

Format notes (R internals, serialize.c): gzip'd stream, header "X\\n", then
version (2 or 3), writer version, min reader version, (v3: native encoding),
followed by one serialized SEXP.  Each item starts with a 32-bit flags word:
type = flags & 0xff, is_obj = bit 8, has_attr = bit 9, has_tag = bit 10.
"""
import gzip
import struct

import numpy as np

NILVALUE_SXP = 254
GLOBALENV_SXP = 253
EMPTYENV_SXP = 242
BASEENV_SXP = 241
MISSINGARG_SXP = 251
UNBOUNDVALUE_SXP = 252
REFSXP = 255
NAMESPACESXP = 249
PACKAGESXP = 250
PERSISTSXP = 247
ALTREP_SXP = 238
ATTRLISTSXP = 239
ATTRLANGSXP = 240
BASENAMESPACE_SXP = 248

NA_INT = -2147483648


class RObj:
    """A decoded R object: `.value` (python/numpy payload) + `.attr` dict."""

    def __init__(self, rtype, value=None, attr=None):
        self.rtype = rtype
        self.value = value
        self.attr = attr or {}

    def __repr__(self):
        return "RObj(type=%d, attr=%s)" % (self.rtype, list(self.attr))


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

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP,
                 UNBOUNDVALUE_SXP, BASENAMESPACE_SXP):
            return RObj(t)
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            sym = RObj(1, name.value if isinstance(name, RObj) else name)
            self.refs.append(sym)
            return sym
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            info = [self.item() for _ in range(n)]
            o = RObj(t, info)
            self.refs.append(o)
            return o
        if t == 4:  # ENVSXP
            self.i32()  # locked
            o = RObj(4, {})
            self.refs.append(o)
            self.item()  # enclos
            self.item()  # frame
            self.item()  # hashtab
            self.item()  # attrib
            return o
        if t in (2, 6, 5, 17, ATTRLISTSXP, ATTRLANGSXP):  # pairlist-like
            if t in (ATTRLISTSXP, ATTRLANGSXP):
                has_attr = True
            # iterative to avoid deep recursion on long pairlists
            items = []
            attr = None
            cur_t, cur_attr, cur_tag = t, has_attr, has_tag
            while True:
                if cur_attr:
                    attr = self.item()
                tag = self.item() if cur_tag else None
                car = self.item()
                items.append((tag.value if isinstance(tag, RObj) else tag, car))
                # peek at cdr
                flags2 = self.i32()
                t2 = flags2 & 0xFF
                if t2 in (2, 6, 5, 17) and t2 == cur_t:
                    cur_attr = bool(flags2 & (1 << 9))
                    cur_tag = bool(flags2 & (1 << 10))
                    continue
                self.o -= 4
                cdr = self.item()
                if cdr is not None:
                    items.append(("__cdr__", cdr))
                break
            return RObj(t, items, {"__attr__": attr} if attr is not None else None)
        if t == ALTREP_SXP:
            info = self.item()
            state = self.item()
            self.item()  # attr
            cls = info.value[0][1].value if isinstance(info, RObj) else None
            return self._altrep(cls, state)
        # vector-like and others
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                val = None
            else:
                val = self.raw(n).decode("utf-8", "replace")
            o = RObj(9, val)
        elif t == 10 or t == 13:  # LGLSXP / INTSXP
            n = self.length()
            val = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
            o = RObj(t, val)
        elif t == 14:  # REALSXP
            n = self.length()
            val = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
            o = RObj(t, val)
        elif t == 16:  # STRSXP
            n = self.length()
            val = [self.item().value for _ in range(n)]
            o = RObj(t, val)
        elif t == 19 or t == 20:  # VECSXP / EXPRSXP
            n = self.length()
            val = [self.item() for _ in range(n)]
            o = RObj(t, val)
        elif t == 24:  # RAWSXP
            n = self.length()
            o = RObj(t, self.raw(n))
        elif t == 25:  # S4SXP
            o = RObj(t)
        elif t == 3:  # CLOSXP
            attr = self.item() if has_attr else None
            env, formals, body = self.item(), self.item(), self.item()
            return RObj(3, (formals, body))
        elif t == 8 or t == 7:  # BUILTIN / SPECIAL
            n = self.i32()
            o = RObj(t, self.raw(n).decode())
        elif t == 21:  # BCODESXP -- not needed by any fixture
            raise NotImplementedError("bytecode")
        else:
            raise NotImplementedError("SEXP type %d at offset %d" % (t, self.o))
        if has_attr:
            a = self.item()
            if a is not None:
                for tag, car in a.value:
                    o.attr[tag] = car
        return o

    def _altrep(self, cls, state):
        cname = cls if isinstance(cls, str) else str(cls)
        if cname == "compact_intseq":
            n, start, step = state.value
            return RObj(13, (start + step * np.arange(int(n))).astype(np.int32))
        if cname == "compact_realseq":
            n, start, step = state.value
            return RObj(14, start + step * np.arange(int(n), dtype=np.float64))
        if cname in ("deferred_string",):
            # state is pairlist (vec, scalar); convert numbers to strings
            vec = state.value[0][1]
            vals = vec.value
            return RObj(16, [str(int(v)) if float(v).is_integer() else repr(float(v)) for v in vals])
        if cname in ("wrap_real", "wrap_integer", "wrap_string", "wrap_logical"):
            return state.value[0] if isinstance(state.value, list) and isinstance(state.value[0], RObj) else state.value[0][1]
        raise NotImplementedError("ALTREP class %s" % cname)


def read_rds(path):
    """Decode one serialized R object from `path` (gzip or plain)."""
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:2] != b"X\n":
        raise ValueError("only XDR serialization is supported")
    r = _Reader(raw)
    r.o = 2
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    return r.item()


def dgcmatrix(obj):
    """Extract (p, i, x, dim, rownames, colnames) from a decoded dgCMatrix S4 object."""
    a = obj.attr
    p = a["p"].value.astype(np.int32)
    i = a["i"].value.astype(np.int32)
    x = a["x"].value.astype(np.float64)
    dim = tuple(int(v) for v in a["Dim"].value)
    dn = a["Dimnames"].value
    rn = dn[0].value if dn[0] is not None else None
    cn = dn[1].value if dn[1] is not None else None
    return p, i, x, dim, rn, cn


def dataframe_rownames(obj):
    rn = obj.attr.get("row.names")
    if rn is None:
        return None
    if rn.rtype == 16:
        return rn.value
    return [str(v) for v in rn.value]
