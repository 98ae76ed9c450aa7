"""ctypes front end for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle("port")``      -> oracle/liboracle.so, the C restatement (fastani_oracle.c)
* ``Oracle("reference")`` -> oracle/_ref/libfastani_ref.so, the reference's own headers
  compiled in place by oracle/Makefile (ref_driver.cpp)

Both expose the same calls so tests can diff one against the other and against
the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module; the product (pyfastani_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("window", C.c_int32), ("frag_len", C.c_int32), ("alphabet", C.c_int32),
                ("min_fraction", C.c_float), ("pct_identity", C.c_float), ("p_value", C.c_double),
                ("ref_size", C.c_uint64)]


class Cand(C.Structure):
    _fields_ = [("frag", C.c_int32), ("seq", C.c_int32), ("start", C.c_int32), ("end", C.c_int32)]


class Mapping(C.Structure):
    _fields_ = [("frag", C.c_int32), ("seq", C.c_int32), ("ref_start", C.c_int32), ("shared", C.c_int32),
                ("sketch", C.c_int32), ("identity", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("ref_genome", C.c_int32), ("matches", C.c_int32), ("fragments", C.c_int32),
                ("identity", C.c_float)]


class Contig(C.Structure):
    _fields_ = [("data", C.c_void_p), ("unit_bytes", C.c_int32), ("len", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fragments", "seeds", "candidates", "scanned", "mappings", "sketch_sum")]


CAND_DT = np.dtype([("frag", "<i4"), ("seq", "<i4"), ("start", "<i4"), ("end", "<i4")])
MAP_DT = np.dtype([("frag", "<i4"), ("seq", "<i4"), ("ref_start", "<i4"), ("shared", "<i4"),
                   ("sketch", "<i4"), ("identity", "<f4")])
HIT_DT = np.dtype([("ref_genome", "<i4"), ("matches", "<i4"), ("fragments", "<i4"), ("identity", "<f4")])


def build(want_ref=True):
    """(Re)build the checkers with oracle/Makefile; returns the paths that exist."""
    subprocess.run(["make", "-s", "-C", HERE, "liboracle.so"] + (["ref"] if want_ref else []), check=True)
    return available()


def available():
    out = {}
    p = os.path.join(HERE, "liboracle.so")
    if os.path.exists(p):
        out["port"] = p
    r = os.path.join(HERE, "_ref", "libfastani_ref.so")
    if os.path.exists(r):
        out["reference"] = r
    return out


def make_params(k=16, fragment_length=3000, minimum_fraction=0.2, p_value=1e-3, percentage_identity=80.0,
                reference_size=5_000_000, window=None, protein=False):
    # protein: alphabet 20 and window 1 (pyx:548-550)
    p = Params(k, 0, fragment_length, 20 if protein else 4, minimum_fraction, percentage_identity, p_value, reference_size)
    return p, (1 if protein and window is None else window)


def _as_buf(seq):
    """(keepalive, pointer, unit_bytes, length) for bytes / str / numpy input, like pyx:633-645."""
    if isinstance(seq, str):
        try:
            b = seq.encode("latin-1")
            unit = 1
        except UnicodeEncodeError:
            m = max(map(ord, seq))
            unit = 2 if m < 65536 else 4
            b = np.array([ord(c) for c in seq], dtype="<u%d" % unit).tobytes()
        arr = np.frombuffer(b, dtype=np.uint8)
        return arr, arr.ctypes.data, unit, len(seq)
    arr = np.frombuffer(seq, dtype=np.uint8) if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq).view(np.uint8)
    return arr, (arr.ctypes.data if arr.size else 0), 1, arr.size


class Oracle:
    def __init__(self, kind="port"):
        libs = available()
        if kind not in libs:
            libs = build(want_ref=(kind == "reference"))
        if kind not in libs:
            raise RuntimeError("oracle %r is not built" % kind)
        self.kind = kind
        self.pfx = "orc_" if kind == "port" else "ref_"
        self.lib = C.CDLL(libs[kind])
        L, f = self.lib, self._f
        f("recommended_window").restype = C.c_int
        f("sketch_new").restype = C.c_void_p
        f("sketch_new").argtypes = [C.POINTER(Params)]
        f("sketch_free").argtypes = [C.c_void_p]
        f("sketch_add_contig").restype = C.c_int64
        f("sketch_add_contig").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64]
        f("sketch_end_genome").argtypes = [C.c_void_p]
        f("sketch_index").argtypes = [C.c_void_p]
        for n in ("sketch_size", "sketch_unique"):
            f(n).restype = C.c_uint64
            f(n).argtypes = [C.c_void_p]
        f("sketch_copy").argtypes = [C.c_void_p] + [C.c_void_p] * 3
        f("minimizers").restype = C.c_int64
        f("minimizers").argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        f("l2_stat").restype = C.c_int
        f("l2_stat").argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_float)]
        f("query").restype = C.c_int64
        args = [C.c_void_p, C.POINTER(Contig), C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int32),
                C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                C.c_int64, C.c_int64, C.POINTER(Stats)]
        if kind == "reference":
            args.append(C.c_int)
        f("query").argtypes = args
        mh = "estimate_minimum_hits_relaxed" if kind == "port" else "minimum_hits_relaxed"
        f(mh).restype = C.c_int
        f(mh).argtypes = [C.c_int, C.c_int, C.c_float]
        self._mh = f(mh)
        f("hash").restype = C.c_uint32
        f("hash").argtypes = [C.c_char_p, C.c_int]

    def _f(self, name):
        return getattr(self.lib, self.pfx + name)

    # ---- scalar helpers -------------------------------------------------
    def recommended_window(self, **kw):
        p, _ = make_params(**kw)
        return self._f("recommended_window")(C.byref(p))

    def minimum_hits(self, s, k=16, pid=80.0):
        return self._mh(s, k, pid)

    def l2_stat(self, shared, s, k=16, pid=80.0):
        ident = C.c_float()
        ok = self._f("l2_stat")(shared, s, k, pid, C.byref(ident))
        return bool(ok), ident.value

    def hash(self, kmer):
        return self._f("hash")(kmer, len(kmer))

    def minimizers(self, seq, k=16, w=24, seq_id=0, protein=False):
        keep, ptr, unit, n = _as_buf(seq)
        cap = max(n, 1)
        h = np.empty(cap, np.uint32); s = np.empty(cap, np.int32); wp = np.empty(cap, np.int32)
        if protein:         # (C port only)
            fn = self._f("minimizers_alpha")
            fn.restype = C.c_int64
            fn.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int32, C.c_int,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
            m = fn(ptr, unit, n, k, w, seq_id, 20, h.ctypes.data, s.ctypes.data, wp.ctypes.data, cap)
        else:
            m = self._f("minimizers")(ptr, unit, n, k, w, seq_id, h.ctypes.data, s.ctypes.data, wp.ctypes.data, cap)
        return h[:m].copy(), s[:m].copy(), wp[:m].copy()

    # ---- sketch / query -------------------------------------------------
    def sketch(self, **kw):
        return OracleSketch(self, **kw)


class OracleSketch:
    """Mirrors Sketch.add_draft/add_genome + Mapper.query_draft at the C level."""

    def __init__(self, orc, **kw):
        self.o = orc
        self.params, window = make_params(**kw)
        self.params.window = window if window is not None else orc._f("recommended_window")(C.byref(self.params))
        self.h = orc._f("sketch_new")(C.byref(self.params))
        self.names = []
        self.warnings = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.o._f("sketch_free")(self.h)
            self.h = None

    def add_draft(self, name, contigs):
        for c in contigs:
            keep, ptr, unit, n = _as_buf(c)
            if self.o._f("sketch_add_contig")(self.h, ptr, unit, n) < 0:
                self.warnings += 1
        self.o._f("sketch_end_genome")(self.h)
        self.names.append(name)
        return self

    def add_genome(self, name, seq):
        return self.add_draft(name, (seq,))

    def add_genomes(self, names, seqs, threads=1):
        """Many single-contig genomes; the compiled reference sketches them on `threads` threads (set-up of the CPU
        baseline), the port one by one.  Same result as add_genome in a loop."""
        if self.o.kind != "reference" or threads <= 1:
            for n, s in zip(names, seqs):
                self.add_genome(n, s)
            return self
        keeps, arr = [], (Contig * max(len(seqs), 1))()
        for i, c in enumerate(seqs):
            keep, ptr, unit, n = _as_buf(c)
            keeps.append(keep)
            arr[i] = Contig(ptr, unit, n)
        fn = self.o._f("sketch_add_genomes")
        fn.argtypes = [C.c_void_p, C.POINTER(Contig), C.c_int32, C.c_int]
        fn.restype = None
        fn(self.h, arr, len(seqs), threads)
        self.names.extend(names)
        return self

    def index(self):
        self.o._f("sketch_index")(self.h)
        return self

    def minimizers(self):
        n = self.o._f("sketch_size")(self.h)
        h = np.empty(n, np.uint32); s = np.empty(n, np.int32); w = np.empty(n, np.int32)
        self.o._f("sketch_copy")(self.h, h.ctypes.data, s.ctypes.data, w.ctypes.data)
        return h, s, w

    def unique(self):
        return self.o._f("sketch_unique")(self.h)

    def query_draft(self, contigs, dump=False, frag_first=0, frag_count=-1, threads=1):
        """Returns (hits, info); hits is a HIT_DT array already filtered and sorted (pyx:1121-1135)."""
        keeps, arr = [], (Contig * max(len(contigs), 1))()
        for i, c in enumerate(contigs):
            keep, ptr, unit, n = _as_buf(c)
            keeps.append(keep)
            arr[i] = Contig(ptr, unit, n)
        hits = np.zeros(max(len(self.names), 1), HIT_DT)
        shorts, ncand, nmap, st = C.c_int32(), C.c_int64(), C.c_int64(), Stats()
        extra = [threads] if self.o.kind == "reference" else []

        def call(cands, maps):
            return self.o._f("query")(self.h, arr, len(contigs), hits.ctypes.data, len(hits), C.byref(shorts),
                                      cands.ctypes.data if cands is not None else None,
                                      len(cands) if cands is not None else 0, C.byref(ncand),
                                      maps.ctypes.data if maps is not None else None,
                                      len(maps) if maps is not None else 0, C.byref(nmap),
                                      frag_first, frag_count, C.byref(st), *extra)

        nh = call(None, None)
        info = {"short_contigs": shorts.value, "stats": {n: getattr(st, n) for n, _ in Stats._fields_}}
        if dump:
            cands = np.zeros(max(ncand.value, 1), CAND_DT)
            maps = np.zeros(max(nmap.value, 1), MAP_DT)
            nh = call(cands, maps)
            info["candidates"] = cands[:ncand.value]
            info["mappings"] = maps[:nmap.value]
        return hits[:nh].copy(), info

    def query_genome(self, seq, **kw):
        return self.query_draft((seq,), **kw)
