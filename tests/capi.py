"""ctypes binding of the C ABI (include/fastani_b200.h) used by the parity tests so that
they call libfastani_b200.so exactly as a foreign host (cgo / JNI / Cython) would."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "pyfastani_b200", "lib", "libfastani_b200.so")
HEADER = os.path.join(ROOT, "include", "fastani_b200.h")


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("window", C.c_int32), ("frag_len", C.c_int32), ("alphabet", C.c_int32),
                ("min_fraction", C.c_float), ("pct_identity", C.c_float), ("p_value", C.c_double),
                ("ref_size", C.c_uint64)]


class Contig(C.Structure):
    _fields_ = [("data", C.c_void_p), ("unit_bytes", C.c_int32), ("on_device", C.c_int32), ("len", C.c_int64)]


class Hit(C.Structure):
    _fields_ = [("ref_genome", C.c_int32), ("matches", C.c_int32), ("fragments", C.c_int32), ("identity", C.c_float)]


class QueryInfo(C.Structure):
    _fields_ = ([(n, C.c_uint64) for n in ("fragments", "sketch_sum", "seeds", "candidates", "scanned", "mappings")]
                + [("short_contigs", C.c_int32), ("kernel_launches", C.c_int32)]
                + [(n, C.c_float) for n in ("ms_h2d", "ms_sketch", "ms_lookup", "ms_seed_sort", "ms_l1", "ms_l2",
                                            "ms_cgi", "ms_d2h", "ms_total")]
                + [("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("l2_fallback", C.c_uint64), ("events", C.c_uint64)]
                + [(n, C.c_float) for n in ("ms_l2_prep", "ms_l2_events", "ms_l2_slide")]
                + [("l1_sorted_fragments", C.c_uint32), ("l1_small_fragments", C.c_uint32), ("events_replayed", C.c_uint64),
                   ("ms_batch", C.c_float), ("l1_parts", C.c_uint32), ("l1_tiny_fragments", C.c_uint32), ("ms_exchange", C.c_float)])

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


HIT_DT = np.dtype([("ref_genome", "<i4"), ("matches", "<i4"), ("fragments", "<i4"), ("identity", "<f4")])
CAND_DT = np.dtype([("frag", "<i4"), ("seq", "<i4"), ("start", "<i4"), ("end", "<i4")])
MAP_DT = np.dtype([("frag", "<i4"), ("seq", "<i4"), ("ref_start", "<i4"), ("shared", "<i4"),
                   ("sketch", "<i4"), ("identity", "<f4")])

_lib = None


class FaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fa error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.fa_last_error.restype = C.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise FaError(rc, lib().fa_last_error().decode())


def make_params(k=16, fragment_length=3000, minimum_fraction=0.2, p_value=1e-3, percentage_identity=80.0,
                reference_size=5_000_000, window=0, protein=False):
    # protein: alphabet 20; window 0 lets the library choose (1 in protein mode, pyx:548-550)
    return Params(k, window, fragment_length, 20 if protein else 4, minimum_fraction, percentage_identity, p_value, reference_size)


def as_buf(seq):
    """(keepalive, pointer, unit_bytes, length) like the reference's input adaptation (pyx:633-645)."""
    if isinstance(seq, str):
        try:
            b = seq.encode("latin-1")
            unit = 1
        except UnicodeEncodeError:
            m = max(map(ord, seq))
            unit = 2 if m < 65536 else 4
            b = np.array([ord(c) for c in seq], dtype="<u%d" % unit).tobytes()
        arr = np.frombuffer(b, dtype=np.uint8)
        return arr, (arr.ctypes.data if arr.size else 0), unit, len(seq)
    arr = np.frombuffer(seq, dtype=np.uint8) if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq).view(np.uint8)
    return arr, (arr.ctypes.data if arr.size else 0), 1, arr.size


class PackedStruct(C.Structure):
    _fields_ = [("bits", C.c_void_p), ("run_pos", C.c_void_p), ("run_len", C.c_void_p), ("run_byte", C.c_void_p),
                ("n_runs", C.c_uint64)]


FA_UNIT_PACKED2 = -2


class Packed:
    """A sequence packed by fa_pack_2bit (include/fastani_b200.h fa_packed); pass it wherever a contig goes."""

    def __init__(self, data):
        src = np.frombuffer(bytes(data), dtype=np.uint8)
        self.length = src.size
        self.bits = np.zeros((src.size + 3) // 4, dtype=np.uint8)
        cap, n = 16, C.c_uint64()
        while True:
            self.run_pos, self.run_len = np.zeros(cap, dtype=np.uint32), np.zeros(cap, dtype=np.uint32)
            self.run_byte = np.zeros(cap, dtype=np.uint8)
            check(lib().fa_pack_2bit(C.c_void_p(src.ctypes.data if src.size else 0), C.c_uint64(src.size),
                                     C.c_void_p(self.bits.ctypes.data), C.c_void_p(self.run_pos.ctypes.data),
                                     C.c_void_p(self.run_len.ctypes.data), C.c_void_p(self.run_byte.ctypes.data),
                                     C.c_uint64(cap), C.byref(n)))
            if n.value <= cap:
                break
            cap = n.value
        self.n_runs = n.value
        self.struct = PackedStruct(self.bits.ctypes.data, self.run_pos.ctypes.data, self.run_len.ctypes.data,
                                   self.run_byte.ctypes.data, self.n_runs)

    def unpack(self):
        out = np.zeros(self.length, dtype=np.uint8)
        check(lib().fa_unpack_2bit(C.byref(self.struct), C.c_uint64(self.length), C.c_void_p(out.ctypes.data)))
        return out.tobytes()


class DeviceContig:
    """A contig already in device memory (fa_contig.on_device), e.g. a record of fa_fasta_parse."""

    def __init__(self, ptr, length, keep=None):
        self.ptr, self.length, self.keep = ptr, length, keep


class Fasta:
    """fa_fasta_parse: FASTA text -> device-resident records."""

    def __init__(self, text, device=0):
        text = bytes(text)
        self.h = C.c_void_p()
        check(lib().fa_fasta_parse(C.c_int32(device), text, C.c_uint64(len(text)), C.byref(self.h)))
        nr, nb = C.c_uint64(), C.c_uint64()
        check(lib().fa_fasta_counts(self.h, C.byref(nr), C.byref(nb)))
        self.n_bases = nb.value
        arr = (Contig * max(nr.value, 1))()
        ib, il = (C.c_uint64 * max(nr.value, 1))(), (C.c_uint64 * max(nr.value, 1))()
        check(lib().fa_fasta_records(self.h, arr, ib, il))
        self.ids = [text[ib[i]:ib[i] + il[i]].decode("latin-1") for i in range(nr.value)]
        self.contigs = [DeviceContig(arr[i].data, arr[i].len, self) for i in range(nr.value)]
        self.device = device

    def download(self, i):
        c = self.contigs[i]
        out = np.zeros(c.length, dtype=np.uint8)
        if c.length:
            check(lib().fa_device_download(C.c_int32(self.device), C.c_void_p(out.ctypes.data), C.c_void_p(c.ptr), C.c_uint64(c.length)))
        return out.tobytes()

    def __del__(self):
        if getattr(self, "h", None):
            lib().fa_fasta_free(self.h)
            self.h = None


def contig_array(contigs):
    keeps, arr = [], (Contig * max(len(contigs), 1))()
    for i, c in enumerate(contigs):
        if isinstance(c, Packed):
            keeps.append(c)
            arr[i] = Contig(C.addressof(c.struct), FA_UNIT_PACKED2, 0, c.length)
            continue
        if isinstance(c, DeviceContig):
            keeps.append(c)
            arr[i] = Contig(c.ptr, 1, 1, c.length)
            continue
        keep, ptr, unit, n = as_buf(c)
        keeps.append(keep)
        arr[i] = Contig(ptr, unit, 0, n)
    return keeps, arr


def mem_info(device=0):
    f, t = C.c_uint64(), C.c_uint64()
    check(lib().fa_device_mem_info(device, C.byref(f), C.byref(t)))
    return f.value, t.value


def recommended_window(**kw):
    p = make_params(**kw)
    w = C.c_int32()
    check(lib().fa_recommended_window(C.byref(p), C.byref(w)))
    return w.value


def stat_minimum_hits(s, k=16, pid=80.0):
    out = C.c_int32()
    check(lib().fa_stat_minimum_hits(s, k, C.c_float(pid), C.byref(out)))
    return out.value


def stat_l2(shared, s, k=16, pid=80.0):
    ident, ok = C.c_float(), C.c_int32()
    check(lib().fa_stat_l2(shared, s, k, C.c_float(pid), C.byref(ident), C.byref(ok)))
    return bool(ok.value), ident.value


class Sketch:
    def __init__(self, device=0, batched=True, **kw):
        self.params = make_params(**kw)
        self.h = C.c_void_p()
        check(lib().fa_sketch_create(C.byref(self.params), device, C.byref(self.h)))
        self.names, self.warnings, self.batched = [], 0, batched

    def __del__(self):
        if getattr(self, "h", None):
            lib().fa_sketch_free(self.h)
            self.h = None

    def add_draft(self, name, contigs):
        contigs = list(contigs)
        if self.batched:
            keeps, arr = contig_array(contigs)
            glen, nshort = C.c_uint64(), C.c_int32()
            check(lib().fa_sketch_add_genome(self.h, arr, len(contigs), C.byref(glen), C.byref(nshort)))
            self.warnings += nshort.value
        else:
            for c in contigs:
                keep, ptr, unit, n = as_buf(c)
                added = C.c_int64()
                check(lib().fa_sketch_add_contig(self.h, C.c_void_p(ptr), unit, n, C.byref(added)))
                if added.value < 0:
                    self.warnings += 1
            check(lib().fa_sketch_end_genome(self.h, None))
        self.names.append(name)
        return self

    def add_genome(self, name, seq):
        return self.add_draft(name, (seq,))

    def add_many(self, names, genomes):
        """fa_sketch_add_genomes: `genomes` = list of contig lists."""
        flat = [c for g in genomes for c in g]
        keeps, arr = contig_array(flat)
        counts = (C.c_int32 * max(len(genomes), 1))(*[len(g) for g in genomes])
        glen = (C.c_uint64 * max(len(genomes), 1))()
        nshort = C.c_int32()
        check(lib().fa_sketch_add_genomes(self.h, arr, counts, len(genomes), glen, C.byref(nshort)))
        self.warnings += nshort.value
        self.names.extend(names)
        return list(glen)[:len(genomes)]

    def meta(self):
        n = self.counts()[2]
        sbg = (C.c_int32 * max(n, 1))(); gl = (C.c_uint64 * max(n, 1))()
        check(lib().fa_sketch_copy_meta(self.h, sbg, gl))
        return list(sbg)[:n], list(gl)[:n]

    def counts(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().fa_sketch_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def minimizers(self):
        n = self.counts()[0]
        h = np.empty(n, np.uint32); s = np.empty(n, np.int32); w = np.empty(n, np.int32)
        check(lib().fa_sketch_copy_minimizers(self.h, 0, n, h.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                              w.ctypes.data_as(C.c_void_p)))
        return h, s, w

    def index(self):
        ix = C.c_void_p()
        check(lib().fa_sketch_index(self.h, C.byref(ix)))
        m = Index(ix, list(self.names))
        self.names = []
        return m


class Index:
    def __init__(self, h, names):
        self.h, self.names = h, names

    def __del__(self):
        if getattr(self, "h", None):
            lib().fa_index_free(self.h)
            self.h = None

    def counts(self):
        v = [C.c_uint64() for _ in range(4)]
        check(lib().fa_index_counts(self.h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def params(self):
        p = Params()
        check(lib().fa_index_params(self.h, C.byref(p)))
        return p

    def minimizers(self):
        n = self.counts()[0]
        h = np.empty(n, np.uint32); s = np.empty(n, np.int32); w = np.empty(n, np.int32)
        check(lib().fa_index_copy_minimizers(self.h, 0, n, h.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                             w.ctypes.data_as(C.c_void_p)))
        return h, s, w

    def keys(self):
        n = self.counts()[1]
        k = np.empty(n, np.uint32)
        check(lib().fa_index_copy_keys(self.h, 0, n, k.ctypes.data_as(C.c_void_p)))
        return k

    def lookup(self, hash_):
        n = C.c_uint64()
        check(lib().fa_index_lookup(self.h, C.c_uint32(hash_), None, None, 0, C.byref(n)))
        s = np.empty(n.value, np.int32); w = np.empty(n.value, np.int32)
        if n.value:
            check(lib().fa_index_lookup(self.h, C.c_uint32(hash_), s.ctypes.data_as(C.c_void_p),
                                        w.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return s, w

    def set_l1_seed_cap(self, cap):
        """Test hook: fragments with more seeds than `cap` take the radix-sort L1 path (-1 = default)."""
        check(lib().fa_debug_set_l1_seed_cap(self.h, C.c_int64(cap)))

    def set_l1_small_cap(self, cap):
        """Test hook: on-chip fragments with more seeds than `cap` take the large shape of the L1 kernel (-1 = default)."""
        check(lib().fa_debug_set_l1_small_cap(self.h, C.c_int64(cap)))

    def set_l1_tiny_cap(self, cap):
        check(lib().fa_debug_set_l1_tiny_cap(self.h, C.c_int64(cap)))

    def set_l1_parts(self, parts, part_cap=-1):
        check(lib().fa_debug_set_l1_parts(self.h, C.c_int32(parts), C.c_int64(part_cap)))

    def set_l1_small_shape(self, shape):
        """Test hook: smallest of the three shared-memory sizes the small L1 shape may use (-1 = default)."""
        check(lib().fa_debug_set_l1_small_shape(self.h, C.c_int32(shape)))

    def query_draft(self, contigs, dump=False):
        contigs = list(contigs)
        keeps, arr = contig_array(contigs)
        hits = np.zeros(max(len(self.names), 1), HIT_DT)
        n, info = C.c_uint64(), QueryInfo()
        check(lib().fa_query(self.h, arr, len(contigs), hits.ctypes.data_as(C.c_void_p), len(hits), C.byref(n), C.byref(info)))
        out = {"short_contigs": info.short_contigs, "info": info.as_dict()}
        if dump:
            m = C.c_uint64()
            check(lib().fa_debug_last_candidates(self.h, None, 0, C.byref(m)))
            cands = np.zeros(max(m.value, 1), CAND_DT)
            check(lib().fa_debug_last_candidates(self.h, cands.ctypes.data_as(C.c_void_p), m.value, C.byref(m)))
            out["candidates"] = cands[:m.value]
            maps = np.zeros(max(m.value, 1), MAP_DT)
            k = C.c_uint64()
            check(lib().fa_debug_last_mappings(self.h, maps.ctypes.data_as(C.c_void_p), m.value, C.byref(k)))
            out["mappings"] = maps[:k.value]
        return hits[:n.value].copy(), out

    def query_genome(self, seq, **kw):
        return self.query_draft((seq,), **kw)

    def query_batch(self, queries):
        """fa_query_batch: `queries` = list of contig lists -> (list of HIT_DT arrays, summed info dict)."""
        flat = [c for q in queries for c in q]
        keeps, arr = contig_array(flat)
        counts = (C.c_int32 * max(len(queries), 1))(*[len(q) for q in queries])
        offs = (C.c_uint64 * (len(queries) + 1))()
        hits = np.zeros(max(len(self.names), 1) * max(len(queries), 1), HIT_DT)
        info = QueryInfo()
        check(lib().fa_query_batch(self.h, arr, counts, len(queries), hits.ctypes.data_as(C.c_void_p), len(hits), offs, C.byref(info)))
        return [hits[offs[q]:offs[q + 1]].copy() for q in range(len(queries))], info.as_dict()
