# coding: utf-8
# cython: language_level=3, binding=False
"""Drop-in mirror of ``pyfastani._fastani`` (src/pyfastani/_fastani.pyx in the reference)
whose every computing call goes to ``libfastani_b200.so`` through the C ABI of
``include/fastani_b200.h`` -- CUDA kernels for sm_100a, no CPU fallback.

The class surface follows src/pyfastani/_fastani.pyi:1-116: `Sketch.add_genome/add_draft/
index` -> `Mapper.query_genome/query_draft` -> `Hit(name, identity, matches, fragments)`,
plus the `Minimizers`, `MinimizerInfo`, `MinimizerIndex`, `Position` views and pickling.
This module only marshals: input adaptation (str of any kind / any contiguous byte buffer,
pyx:633-645), warnings (pyx:671-677, 1063-1069), argument errors (pyx:523-539, 1049-1050),
names and result objects.
"""

cimport cython
from cpython.unicode cimport (
    PyUnicode_DATA,
    PyUnicode_KIND,
    PyUnicode_GET_LENGTH,
    PyUnicode_1BYTE_KIND,
    PyUnicode_2BYTE_KIND,
    PyUnicode_FromKindAndData,
)
from libc.stdint cimport int32_t, int64_t, uint8_t, uint32_t, uint64_t, uintptr_t
from libc.stdlib cimport malloc, free
from libc.string cimport memset
from cpython.bytes cimport PyBytes_FromStringAndSize

import threading
import warnings


cdef extern from "fastani_b200.h" nogil:
    ctypedef struct fa_params:
        int32_t k
        int32_t window
        int32_t frag_len
        int32_t alphabet
        float min_fraction
        float pct_identity
        double p_value
        uint64_t ref_size
    ctypedef struct fa_contig:
        const void* data
        int32_t unit_bytes
        int32_t on_device
        int64_t len
    ctypedef struct fa_hit:
        int32_t ref_genome
        int32_t matches
        int32_t fragments
        float identity
    ctypedef struct fa_query_info:
        uint64_t fragments
        uint64_t sketch_sum
        uint64_t seeds
        uint64_t candidates
        uint64_t scanned
        uint64_t mappings
        int32_t short_contigs
        int32_t kernel_launches
        float ms_h2d
        float ms_sketch
        float ms_lookup
        float ms_seed_sort
        float ms_l1
        float ms_l2
        float ms_cgi
        float ms_d2h
        float ms_total
        uint64_t h2d_bytes
        uint64_t d2h_bytes
        uint64_t l2_fallback
        uint64_t events
        float ms_l2_prep
        float ms_l2_events
        float ms_l2_slide
        uint32_t l1_sorted_fragments
        uint32_t l1_small_fragments
        uint64_t events_replayed
        float ms_batch
        uint32_t l1_parts
        uint32_t l1_tiny_fragments
        float ms_exchange
    ctypedef struct fa_packed:
        const uint8_t* bits
        const uint32_t* run_pos
        const uint32_t* run_len
        const uint8_t* run_byte
        uint64_t n_runs
    int FA_UNIT_PACKED2
    ctypedef struct fa_sketch
    ctypedef struct fa_index
    ctypedef struct fa_comm
    ctypedef struct fa_fasta

    const char* fa_last_error()
    int fa_device_count(int32_t* n)
    int fa_recommended_window(const fa_params* p, int32_t* w)
    int fa_sketch_create(const fa_params* p, int32_t device, fa_sketch** out)
    void fa_sketch_free(fa_sketch* s)
    int fa_sketch_add_genome(fa_sketch* s, const fa_contig* contigs, int32_t n, uint64_t* glen, int32_t* n_short)
    int fa_sketch_add_genomes(fa_sketch* s, const fa_contig* contigs, const int32_t* contigs_per_genome, int32_t n_genomes,
                              uint64_t* glen, int32_t* n_short)
    int fa_sketch_build_stats(const fa_sketch* s, double* ms_sketch, uint64_t* bases)
    int fa_index_build_stats(const fa_index* ix, float* ms_build, float* ms_sort)
    int fa_sketch_clear(fa_sketch* s)
    int fa_sketch_counts(const fa_sketch* s, uint64_t* n_min, uint64_t* n_contigs, uint64_t* n_genomes)
    int fa_sketch_copy_minimizers(const fa_sketch* s, uint64_t first, uint64_t n, uint32_t* h, int32_t* sq, int32_t* w)
    int fa_sketch_copy_meta(const fa_sketch* s, int32_t* seqs_by_genome, uint64_t* genome_len)
    int fa_sketch_restore(fa_sketch* s, const uint32_t* h, const int32_t* sq, const int32_t* w, uint64_t n,
                          const int32_t* seqs_by_genome, const uint64_t* genome_len, uint64_t n_genomes, uint64_t n_contigs)
    int fa_sketch_index(fa_sketch* s, fa_index** out)
    void fa_index_free(fa_index* ix)
    int fa_index_counts(const fa_index* ix, uint64_t* n_min, uint64_t* n_unique, uint64_t* n_contigs, uint64_t* n_genomes)
    int fa_index_copy_minimizers(const fa_index* ix, uint64_t first, uint64_t n, uint32_t* h, int32_t* sq, int32_t* w)
    int fa_index_copy_meta(const fa_index* ix, int32_t* seqs_by_genome, uint64_t* genome_len)
    int fa_index_copy_keys(const fa_index* ix, uint64_t first, uint64_t n, uint32_t* keys)
    int fa_index_lookup(const fa_index* ix, uint32_t hash, int32_t* sq, int32_t* w, uint64_t cap, uint64_t* n)
    int fa_index_has_key(const fa_index* ix, uint32_t hash, int32_t* found)
    int fa_index_set_lookup(fa_index* ix, uint32_t hash, const int32_t* sq, const int32_t* w, uint64_t n)
    int fa_index_del_lookup(fa_index* ix, uint32_t hash, int32_t* found)
    int fa_index_occurrence_threshold(const fa_index* ix, int32_t* out)
    int fa_query(fa_index* ix, const fa_contig* contigs, int32_t n, fa_hit* out, uint64_t cap, uint64_t* n_out,
                 fa_query_info* info)
    int fa_query_batch(fa_index* ix, const fa_contig* contigs, const int32_t* contigs_per_query, int32_t n_queries,
                       fa_hit* out, uint64_t cap, uint64_t* hit_offsets, fa_query_info* info)
    int fa_comm_unique_id(unsigned char* id)
    int fa_comm_create(const unsigned char* id, int32_t world, int32_t rank, int32_t device, fa_comm** out)
    void fa_comm_free(fa_comm* c)
    int fa_comm_info(const fa_comm* c, int32_t* world, int32_t* rank, int32_t* nccl_version, uint64_t* collectives,
                     uint64_t* bytes_gathered)
    int fa_gather_hits(fa_comm* c, const fa_hit* rows, const uint64_t* hit_offsets, int32_t n_queries,
                       const int32_t* genome_offsets, fa_hit* out, uint64_t cap, uint64_t* out_offsets)
    int fa_query_batch_sharded(fa_index* ix, fa_comm* comm, const fa_contig* contigs, const int32_t* contigs_per_query,
                               int32_t n_queries, const int32_t* genome_offsets, fa_hit* out, uint64_t cap,
                               uint64_t* hit_offsets, fa_query_info* info)
    int fa_pack_2bit(const uint8_t* data, uint64_t n, uint8_t* bits, uint32_t* run_pos, uint32_t* run_len,
                     uint8_t* run_byte, uint64_t run_cap, uint64_t* n_runs)
    int fa_unpack_2bit(const fa_packed* p, uint64_t n, uint8_t* out)
    int fa_fasta_parse(int32_t device, const void* text, uint64_t n, fa_fasta** out)
    void fa_fasta_free(fa_fasta* f)
    int fa_fasta_counts(const fa_fasta* f, uint64_t* n_records, uint64_t* n_bases)
    int fa_fasta_records(const fa_fasta* f, fa_contig* contigs, uint64_t* id_begin, uint64_t* id_len)
    int fa_device_alloc(int32_t device, uint64_t nbytes, void** dptr)
    int fa_device_upload(int32_t device, void* dptr, const void* src, uint64_t nbytes)
    int fa_device_free(int32_t device, void* dptr)


cdef extern from *:
    """
    #define _MAX_KMER_SIZE 2048
    """
    const size_t _MAX_KMER_SIZE

MAX_KMER_SIZE = _MAX_KMER_SIZE
__version__ = "0.1.0"


class CudaError(RuntimeError):
    """Raised when the CUDA library reports a failure (there is no CPU path to fall back to)."""


cdef int _check(int rc) except -1:
    if rc == 0:
        return 0
    msg = fa_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    elif rc == 3:
        raise MemoryError(msg)
    elif rc == 4:
        raise NotImplementedError(msg)
    raise CudaError(msg)


def device_count():
    """Number of CUDA devices visible to the library."""
    cdef int32_t n = 0
    _check(fa_device_count(&n))
    return n


# --- multi-GPU ------------------------------------------------------------------

HIT_FIELDS = [("ref_genome", "<i4"), ("matches", "<i4"), ("fragments", "<i4"), ("identity", "<f4")]


cdef object _hit_rows(const fa_hit* rows, const uint64_t* offs, int32_t nq):
    """`fa_hit` rows of nq queries -> a list of numpy structured arrays (slices of one array), no `Hit` objects."""
    import numpy
    cdef uint64_t total = offs[nq]
    data = bytearray(PyBytes_FromStringAndSize(<const char*> rows, total * sizeof(fa_hit))) if total else bytearray()
    allrows = numpy.frombuffer(data, dtype=numpy.dtype(HIT_FIELDS))
    return [allrows[offs[q]:offs[q + 1]] for q in range(nq)]


cdef class Communicator:
    """One rank of a group of GPUs that map against shards of a reference set (not in the reference API; the analogue of
    upstream FastANI's per-thread reference split, computeCoreIdentity.hpp:454-484).  Wraps an NCCL communicator owned
    by the library: `unique_id()` on one rank, its 128 bytes handed to the others by the caller, then
    `Communicator(id, world_size, rank, device)` on every rank (collective)."""
    cdef fa_comm* _c
    cdef readonly int world_size
    cdef readonly int rank
    cdef readonly int device

    def __cinit__(self):
        self._c = NULL

    def __init__(self, bytes unique_id, int world_size, int rank, int device=0):
        cdef const unsigned char* idp = unique_id
        cdef int rc
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes returned by Communicator.unique_id()")
        with nogil:
            rc = fa_comm_create(idp, world_size, rank, device, &self._c)
        _check(rc)
        self.world_size, self.rank, self.device = world_size, rank, device

    def __dealloc__(self):
        if self._c != NULL:
            fa_comm_free(self._c)
            self._c = NULL

    @staticmethod
    def unique_id():
        cdef unsigned char buf[128]
        _check(fa_comm_unique_id(buf))
        return PyBytes_FromStringAndSize(<const char*> buf, 128)

    @property
    def info(self):
        cdef int32_t w = 0, r = 0, v = 0
        cdef uint64_t n = 0, b = 0
        _check(fa_comm_info(self._c, &w, &r, &v, &n, &b))
        return {"world_size": w, "rank": r, "nccl_version": v, "collectives": n, "bytes_gathered": b}

    def gather_hits(self, object rows_per_query, object genome_offsets):
        """Collective: this rank's rows (structured arrays with LOCAL genome ids, one per query) -> on every rank one
        array per query with the rows of all ranks, GLOBAL ids, in the reference's order (`fa_gather_hits`)."""
        import numpy
        cdef list items = [numpy.ascontiguousarray(r, dtype=numpy.dtype(HIT_FIELDS)) for r in rows_per_query]
        cdef int32_t nq = <int32_t> len(items)
        cdef list offs_l = list(genome_offsets)
        if len(offs_l) != self.world_size + 1:
            raise ValueError("genome_offsets needs world_size + 1 entries")
        cdef uint64_t cap = <uint64_t> max(nq, 1) * <uint64_t> max(int(offs_l[-1]) - int(offs_l[0]), 1)
        cdef uint64_t* in_offs = <uint64_t*> malloc((nq + 1) * sizeof(uint64_t))
        cdef uint64_t* out_offs = <uint64_t*> malloc((nq + 1) * sizeof(uint64_t))
        cdef int32_t* goffs = <int32_t*> malloc((self.world_size + 1) * sizeof(int32_t))
        cdef fa_hit* out = <fa_hit*> malloc(cap * sizeof(fa_hit))
        cdef const unsigned char[::1] view
        cdef const fa_hit* inp = NULL
        cdef int rc
        cdef int32_t q
        try:
            if in_offs == NULL or out_offs == NULL or goffs == NULL or out == NULL:
                raise MemoryError()
            flat = numpy.concatenate(items) if nq else numpy.zeros(0, dtype=numpy.dtype(HIT_FIELDS))
            in_offs[0] = 0
            for q in range(nq):
                in_offs[q + 1] = in_offs[q] + <uint64_t> len(items[q])
            for q in range(self.world_size + 1):
                goffs[q] = offs_l[q]
            if len(flat):
                view = flat.view(numpy.uint8)
                inp = <const fa_hit*> &view[0]
            with nogil:
                rc = fa_gather_hits(self._c, inp, in_offs, nq, goffs, out, cap, out_offs)
            _check(rc)
            return _hit_rows(out, out_offs, nq)
        finally:
            free(in_offs); free(out_offs); free(goffs); free(out)


# --- input adaptation ---------------------------------------------------------

cdef class DeviceSequence:
    """A byte sequence resident in GPU memory, accepted wherever a contig is.

    Not part of the reference API: it lets callers (bench.py) keep inputs in HBM so that
    the device-only throughput can be timed next to the host-buffer path.
    """
    cdef void*          _ptr
    cdef readonly int64_t length
    cdef readonly int   device
    cdef object         _owner
    cdef bint           _owned

    def __cinit__(self):
        self._ptr = NULL
        self._owned = False

    def __dealloc__(self):
        if self._owned and self._ptr != NULL:
            fa_device_free(self.device, self._ptr)

    def __len__(self):
        return self.length

    @property
    def pointer(self):
        return <uintptr_t> self._ptr

    @staticmethod
    def from_host(object data, int device=0):
        cdef const unsigned char[::1] view = data
        cdef DeviceSequence d = DeviceSequence.__new__(DeviceSequence)
        d.device = device
        d.length = view.shape[0]
        _check(fa_device_alloc(device, d.length, &d._ptr))
        d._owned = True
        if d.length:
            _check(fa_device_upload(device, d._ptr, &view[0], d.length))
        return d

    @staticmethod
    def from_pointer(uintptr_t pointer, int64_t length, int device=0, object owner=None):
        """Wrap device memory owned by someone else (e.g. a torch tensor's ``data_ptr()``)."""
        cdef DeviceSequence d = DeviceSequence.__new__(DeviceSequence)
        d.device = device
        d.length = length
        if length < 0 or (length > 0 and pointer == 0):
            raise ValueError("from_pointer: a null pointer or a negative length")
        if device < 0 or device >= device_count():
            raise ValueError(f"from_pointer: device {device} out of range")
        d._ptr = <void*> pointer
        d._owner = owner
        return d


cdef class PackedSequence:
    """A sequence at two bits per base (`fa_packed`), accepted wherever a contig is.

    Not part of the reference API (SURVEY.md 8(f)-2): a quarter of the bytes cross host memory and PCIe, the GPU
    expands them in front of the sketch kernel.  ``A C G T`` (either case) take two bits; every other byte -- ``N``,
    IUPAC codes -- is kept exactly, as runs, so results are identical to those of the plain bytes.
    """
    cdef readonly int64_t length
    cdef readonly object  bits       # numpy uint8, (length + 3) // 4
    cdef readonly object  run_pos    # numpy uint32
    cdef readonly object  run_len    # numpy uint32
    cdef readonly object  run_byte   # numpy uint8
    cdef fa_packed _pk

    def __init__(self, int64_t length, object bits, object run_pos, object run_len, object run_byte):
        import numpy
        self.length = length
        self.bits = numpy.ascontiguousarray(bits, dtype=numpy.uint8)
        self.run_pos = numpy.ascontiguousarray(run_pos, dtype=numpy.uint32)
        self.run_len = numpy.ascontiguousarray(run_len, dtype=numpy.uint32)
        self.run_byte = numpy.ascontiguousarray(run_byte, dtype=numpy.uint8)
        if length < 0 or self.bits.shape[0] < (length + 3) // 4:
            raise ValueError("PackedSequence: `bits` is shorter than the length needs")
        if not (self.run_pos.shape[0] == self.run_len.shape[0] == self.run_byte.shape[0]):
            raise ValueError("PackedSequence: run arrays of different lengths")
        cdef const uint8_t[::1] b = self.bits
        cdef const uint32_t[::1] rp = self.run_pos
        cdef const uint32_t[::1] rl = self.run_len
        cdef const uint8_t[::1] rb = self.run_byte
        self._pk.bits = &b[0] if b.shape[0] else NULL
        self._pk.n_runs = rp.shape[0]
        self._pk.run_pos = &rp[0] if rp.shape[0] else NULL
        self._pk.run_len = &rl[0] if rp.shape[0] else NULL
        self._pk.run_byte = &rb[0] if rp.shape[0] else NULL

    def __len__(self):
        return self.length

    def __reduce__(self):
        return PackedSequence, (self.length, self.bits, self.run_pos, self.run_len, self.run_byte)

    @property
    def nbytes(self):
        return self.bits.nbytes + self.run_pos.nbytes + self.run_len.nbytes + self.run_byte.nbytes

    @staticmethod
    def pack(object data):
        """Pack `bytes`-like data (or an ASCII `str`)."""
        import numpy
        if isinstance(data, str):
            data = data.encode("ascii")
        cdef const unsigned char[::1] view = data
        cdef uint64_t n = view.shape[0], n_runs = 0, cap = 1024
        bits = numpy.zeros((n + 3) // 4, dtype=numpy.uint8)
        cdef uint8_t[::1] b = bits
        cdef uint32_t[::1] rp, rl
        cdef uint8_t[::1] rb
        while True:
            pos = numpy.empty(cap, dtype=numpy.uint32); ln = numpy.empty(cap, dtype=numpy.uint32)
            byt = numpy.empty(cap, dtype=numpy.uint8)
            rp = pos; rl = ln; rb = byt
            _check(fa_pack_2bit(&view[0] if n else NULL, n, &b[0] if n else NULL, &rp[0], &rl[0], &rb[0], cap, &n_runs))
            if n_runs <= cap:
                break
            cap = n_runs
        return PackedSequence(n, bits, pos[:n_runs].copy(), ln[:n_runs].copy(), byt[:n_runs].copy())

    def unpack(self):
        """The bytes this sequence stands for (lower-case ``acgt`` come back as capitals)."""
        cdef bytearray out = bytearray(self.length)
        cdef unsigned char[::1] o = out
        if self.length:
            _check(fa_unpack_2bit(&self._pk, self.length, &o[0]))
        return bytes(out)


cdef class DeviceFasta:
    """FASTA text parsed on the GPU: the records of the reference's ``Parser`` (``_fasta.pyx:41-103``) with their
    sequences resident in device memory.

    ``ids[i]`` / ``sequences[i]`` are the identifier and the `DeviceSequence` of record ``i``; pass ``sequences`` to
    ``Sketch.add_draft`` / ``Mapper.query_draft``.  The text is uploaded once, header lines and newlines are dropped
    by a stream compaction on the device, and no per-record host string is ever built.
    """
    cdef fa_fasta* _f
    cdef readonly int   device
    cdef readonly list  ids
    cdef readonly list  sequences
    cdef readonly uint64_t bases

    def __cinit__(self):
        self._f = NULL

    def __dealloc__(self):
        if self._f != NULL:
            fa_fasta_free(self._f)

    def __init__(self, object text, int device=0):
        """`text`: the content of a FASTA file (bytes-like), or a path to read it from."""
        import os
        if isinstance(text, (str, os.PathLike)):
            with open(text, "rb") as fh:
                text = fh.read()
        cdef const unsigned char[::1] view = text
        cdef uint64_t n = view.shape[0], n_rec = 0, n_bases = 0, i
        cdef const unsigned char* p = &view[0] if n else NULL
        cdef int rc
        with nogil:
            rc = fa_fasta_parse(device, p, n, &self._f)
        _check(rc)
        self.device = device
        _check(fa_fasta_counts(self._f, &n_rec, &n_bases))
        self.bases = n_bases
        cdef fa_contig* c = <fa_contig*> malloc(max(n_rec, 1) * sizeof(fa_contig))
        cdef uint64_t* ib = <uint64_t*> malloc(max(n_rec, 1) * sizeof(uint64_t))
        cdef uint64_t* il = <uint64_t*> malloc(max(n_rec, 1) * sizeof(uint64_t))
        cdef DeviceSequence d
        try:
            _check(fa_fasta_records(self._f, c, ib, il))
            self.ids = []
            self.sequences = []
            for i in range(n_rec):
                self.ids.append(PyUnicode_FromKindAndData(PyUnicode_1BYTE_KIND, p + ib[i], il[i]))
                d = DeviceSequence.__new__(DeviceSequence)
                d.device = device; d.length = c[i].len; d._ptr = <void*> c[i].data; d._owner = self
                self.sequences.append(d)
        finally:
            free(c); free(ib); free(il)

    def __len__(self):
        return len(self.ids)

    def __iter__(self):
        return iter(zip(self.ids, self.sequences))


cdef class _Contigs:
    """A C array of `fa_contig` built from Python sequences, keeping the buffers alive."""
    cdef fa_contig* arr
    cdef int32_t    n
    cdef list       keep

    def __cinit__(self):
        self.arr = NULL
        self.n = 0
        self.keep = []

    def __dealloc__(self):
        free(self.arr)

    cdef int fill(self, object contigs, int device) except -1:
        cdef const unsigned char[::1] view
        cdef list items = list(contigs)
        cdef object contig
        cdef int kind
        cdef Py_ssize_t i
        cdef DeviceSequence dev
        cdef PackedSequence pk
        self.n = <int32_t> len(items)
        self.arr = <fa_contig*> malloc(max(self.n, 1) * sizeof(fa_contig))
        if self.arr == NULL:
            raise MemoryError()
        memset(self.arr, 0, max(self.n, 1) * sizeof(fa_contig))
        for i, contig in enumerate(items):
            if isinstance(contig, str):
                # any unicode kind is read in place (pyx:633-637)
                kind = PyUnicode_KIND(contig)
                self.arr[i].data = PyUnicode_DATA(contig)
                self.arr[i].len = PyUnicode_GET_LENGTH(contig)
                self.arr[i].unit_bytes = 1 if kind == PyUnicode_1BYTE_KIND else (2 if kind == PyUnicode_2BYTE_KIND else 4)
                self.keep.append(contig)
            elif isinstance(contig, DeviceSequence):
                dev = contig
                # the kernels dereference the pointer on the Sketch / Mapper's own device
                if dev.device != device:
                    raise ValueError(f"DeviceSequence lives on device {dev.device}, expected device {device}")
                self.arr[i].data = dev._ptr
                self.arr[i].len = dev.length
                self.arr[i].unit_bytes = 1
                self.arr[i].on_device = 1
                self.keep.append(contig)
            elif isinstance(contig, PackedSequence):
                pk = contig
                self.arr[i].data = &pk._pk
                self.arr[i].len = pk.length
                self.arr[i].unit_bytes = FA_UNIT_PACKED2
                self.keep.append(contig)
            else:
                # anything exposing a contiguous byte buffer (pyx:638-645)
                view = contig
                self.arr[i].len = view.shape[0]
                self.arr[i].unit_bytes = 1
                if view.shape[0] != 0:
                    self.arr[i].data = <const void*> &view[0]
                self.keep.append(view)
        return 0


# --- parameters ---------------------------------------------------------------

cdef class _Parameterized:
    """A base class for types wrapping a `skch::Parameters` equivalent (pyx:364-446)."""

    cdef fa_params _param
    cdef int       _device

    def __cinit__(self):
        memset(&self._param, 0, sizeof(fa_params))
        self._device = 0

    def __getstate__(self):
        return {
            "kmerSize": self._param.k,
            "windowSize": self._param.window,
            "minReadLength": self._param.frag_len,
            "minFraction": self._param.min_fraction,
            "threads": 1,
            "alphabetSize": self._param.alphabet,
            "referenceSize": self._param.ref_size,
            "percentageIdentity": self._param.pct_identity,
            "p_value": self._param.p_value,
        }

    def __setstate__(self, state):
        self._param.k = state["kmerSize"]
        self._param.window = state["windowSize"]
        self._param.frag_len = state["minReadLength"]
        self._param.min_fraction = state["minFraction"]
        self._param.alphabet = state["alphabetSize"]
        self._param.ref_size = state["referenceSize"]
        self._param.pct_identity = state["percentageIdentity"]
        self._param.p_value = state["p_value"]

    @property
    def k(self):
        """`int`: The k-mer size used for sketching."""
        return self._param.k

    @property
    def window_size(self):
        """`int`: The window size used for sketching."""
        return self._param.window

    @property
    def fragment_length(self):
        """`int`: The minimum read length to use for mapping."""
        return self._param.frag_len

    @property
    def minimum_fraction(self):
        """`float`: The minimum genome fraction required to trust ANI values."""
        return self._param.min_fraction

    @property
    def percentage_identity(self):
        """`float`: The identity threshold for similarity when estimating hits."""
        return self._param.pct_identity

    @property
    def p_value(self):
        """`float`: The p-value threshold for similarity when estimating hits."""
        return self._param.p_value

    @property
    def protein(self):
        """`bool`: Whether or not the object expects peptides or nucleotides."""
        return self._param.alphabet == 20

    @property
    def device(self):
        """`int`: The CUDA device holding this object's data (not in the reference)."""
        return self._device


# --- Sketch -------------------------------------------------------------------

@cython.final
cdef class Sketch(_Parameterized):
    """An index computing minimizers over the reference genomes (pyx:449-806)."""

    cdef          fa_sketch* _sk
    cdef          list       _names
    cdef readonly Minimizers minimizers
    cdef readonly object     _lock

    def __cinit__(self):
        self._sk = NULL
        self._names = []
        self.minimizers = Minimizers.__new__(Minimizers)
        self.minimizers._owner = self

    def __init__(
        self,
        *,
        unsigned int k=16,
        unsigned int fragment_length=3000,
        float minimum_fraction=0.2,
        double p_value=1e-03,
        float percentage_identity=80.0,
        uint64_t reference_size=5_000_000,
        bint protein=False,
        int device=0,
    ):
        """__init__(self, *, k=16, fragment_length=3000, minimum_fraction=0.2, p_value=1e-03, percentage_identity=80, reference_size=5e6, protein=False, device=0)\n--

        Create a new FastANI sequence sketch on CUDA device ``device``.  Arguments and
        errors as in the reference (pyx:484-539).
        """
        cdef int32_t w = 0
        if minimum_fraction > 1 or minimum_fraction < 0:
            raise ValueError(f"minimum_fraction must be between 0 and 1, got {minimum_fraction!r}")
        if fragment_length <= 0:
            raise ValueError(f"fragment_length must be strictly positive, got {fragment_length!r}")
        if p_value <= 0:
            raise ValueError(f"p_value must be positive, got {p_value!r}")
        if percentage_identity > 100 or percentage_identity < 0:
            raise ValueError(f"percentage_identity must be between 0 and 100, got {percentage_identity!r}")
        if k <= 0:
            raise ValueError(f"k must be strictly positive, got {k!r}")
        elif k > _MAX_KMER_SIZE:
            raise BufferError(f"k must be smaller than {_MAX_KMER_SIZE}, got {k}")
        elif k > 16:
            warnings.warn(
                f"Using k-mer size greater than 16 ({k!r}), accuracy will be degraded.",
                UserWarning,
            )
        self._param.k = k
        self._param.frag_len = fragment_length
        self._param.min_fraction = minimum_fraction
        self._param.p_value = p_value
        self._param.pct_identity = percentage_identity
        self._param.ref_size = reference_size
        if protein:
            # alphabet 20, forward strand only, window 1 (pyx:548-550)
            self._param.alphabet = 20
            self._param.window = 1
        else:
            self._param.alphabet = 4
            self._param.window = 0
            _check(fa_recommended_window(&self._param, &w))
            self._param.window = w
        self._device = device

        self._lock = threading.Lock()
        # (re)create the device-side sketch; __init__ may be called more than once
        if self._sk != NULL:
            fa_sketch_free(self._sk)
            self._sk = NULL
        _check(fa_sketch_create(&self._param, device, &self._sk))
        self._names = []

    def __dealloc__(self):
        if self._sk != NULL:
            fa_sketch_free(self._sk)
            self._sk = NULL

    def __getstate__(self):
        cdef uint64_t n_min = 0, n_contigs = 0, n_genomes = 0
        _check(fa_sketch_counts(self._sk, &n_min, &n_contigs, &n_genomes))
        cdef int32_t*  seqs = <int32_t*> malloc(max(n_genomes, 1) * sizeof(int32_t))
        cdef uint64_t* lens = <uint64_t*> malloc(max(n_genomes, 1) * sizeof(uint64_t))
        try:
            _check(fa_sketch_copy_meta(self._sk, seqs, lens))
            return {
                "parameters": _Parameterized.__getstate__(self),
                "counter": n_contigs,
                "lengths": [lens[i] for i in range(n_genomes)],
                "names": list(self._names),
                "device": self._device,
                "sketch": {
                    "sequencesByFileInfo": [seqs[i] for i in range(n_genomes)],
                    "minimizers": self.minimizers.__getstate__(),
                },
            }
        finally:
            free(seqs)
            free(lens)

    def __setstate__(self, state):
        _Parameterized.__setstate__(self, state["parameters"])
        self._device = state.get("device", 0)
        self._lock = threading.Lock()
        if self._sk != NULL:
            fa_sketch_free(self._sk)
            self._sk = NULL
        _check(fa_sketch_create(&self._param, self._device, &self._sk))
        self._names = list(state["names"])
        _restore_sketch(self._sk, state["sketch"], state["lengths"], state["counter"])

    def save(self, path):
        """save(self, path)\n--

        Write the sketch to `path` (the on-disk form of what `pickle` would carry: parameters, names, genome lengths,
        sequencesByFileInfo and the minimizer columns), so that a later run can `Sketch.load` it instead of sketching
        the genomes again."""
        cdef uint64_t n_min = 0, n_contigs = 0, n_genomes = 0
        _check(fa_sketch_counts(self._sk, &n_min, &n_contigs, &n_genomes))
        cdef int32_t*  seqs = <int32_t*> malloc(max(n_genomes, 1) * sizeof(int32_t))
        cdef uint64_t* lens = <uint64_t*> malloc(max(n_genomes, 1) * sizeof(uint64_t))
        try:
            _check(fa_sketch_copy_meta(self._sk, seqs, lens))
            _save_sketch_file(path, "sketch", _Parameterized.__getstate__(self), list(self._names), n_contigs,
                              [lens[i] for i in range(n_genomes)], [seqs[i] for i in range(n_genomes)], self.minimizers)
        finally:
            free(seqs)
            free(lens)

    @classmethod
    def load(cls, path, device=None):
        """load(cls, path, device=None)\n--

        A `Sketch` with the content of a file written by `Sketch.save` or `Mapper.save`, on `device` (default: 0)."""
        header, cols = _load_sketch_file(path)
        cdef Sketch sk = cls.__new__(cls)
        sk.__setstate__({"parameters": header["parameters"], "device": 0 if device is None else device, "names": [],
                         "lengths": [], "counter": 0,
                         "sketch": {"sequencesByFileInfo": [], "minimizers": {"hashes": [], "ids": [], "offsets": [], "length": 0}}})
        _restore_sketch_columns(sk._sk, cols, header.get("counter", 0))
        sk._names = list(header["names"])
        return sk

    @property
    def occurences_threshold(self):
        """`int`: The occurence threshold above which minimizers are ignored."""
        return 2147483647

    @property
    def names(self):
        """`list` of `str`: The names of the sequences currently sketched."""
        return self._names[:]

    cdef int _add_draft(self, object name, object contigs) except 1:
        cdef _Contigs c = _Contigs.__new__(_Contigs)
        cdef uint64_t glen = 0
        cdef int32_t  n_short = 0
        cdef int      rc
        c.fill(contigs, self._device)
        with nogil:
            rc = fa_sketch_add_genome(self._sk, c.arr, c.n, &glen, &n_short)
        _check(rc)
        for _ in range(n_short):
            warnings.warn(
                (
                    "Sketch received a short contig relative to parameters, "
                    "minimizers will not be added."
                ),
                UserWarning,
            )
        self._names.append(name)
        return 0

    cpdef Sketch add_draft(self, object name, object contigs):
        """add_draft(self, name, contigs)\n--

        Add a reference draft genome to the sketcher (pyx:692-717)."""
        with self._lock:
            self._add_draft(name, contigs)
        return self

    cpdef Sketch add_genome(self, object name, object sequence):
        """add_genome(self, name, sequence)\n--

        Add a reference genome to the sketcher (pyx:719-744)."""
        with self._lock:
            self._add_draft(name, (sequence,))
        return self

    def add_many(self, object names, object genomes):
        """add_many(self, names, genomes)\n--

        Add many reference genomes with one call into the library (`fa_sketch_add_genomes`; not in the
        reference API, which adds genome by genome): ``genomes[i]`` is one sequence (a complete genome) or a
        list / tuple of contigs (a draft), named ``names[i]``.  Same result as `add_genome` / `add_draft` in a
        loop; whole genomes share launch sequences (about 256 MB of bases each)."""
        cdef _Contigs c = _Contigs.__new__(_Contigs)
        cdef list     name_l = list(names)
        cdef list     items = list(genomes)
        cdef list     flat = []
        cdef int32_t  ng = <int32_t> len(items)
        cdef int32_t  n_short = 0
        cdef int32_t* counts
        cdef int      rc
        cdef int32_t  g
        if len(name_l) != len(items):
            raise ValueError("`names` and `genomes` differ in length")
        counts = <int32_t*> malloc(max(ng, 1) * sizeof(int32_t))
        if counts == NULL:
            raise MemoryError()
        try:
            for g in range(ng):
                if isinstance(items[g], (list, tuple)):
                    counts[g] = <int32_t> len(items[g])
                    flat.extend(items[g])
                else:
                    counts[g] = 1
                    flat.append(items[g])
            c.fill(flat, self._device)
            with self._lock:
                with nogil:
                    rc = fa_sketch_add_genomes(self._sk, c.arr, counts, ng, NULL, &n_short)
                _check(rc)
                self._names.extend(name_l)
        finally:
            free(counts)
        for _ in range(n_short):
            warnings.warn(
                (
                    "Sketch received a short contig relative to parameters, "
                    "minimizers will not be added."
                ),
                UserWarning,
            )
        return self

    @property
    def build_stats(self):
        """`dict`: CUDA-event time of the device work of all `add_*` calls so far and the bases they saw."""
        cdef double ms = 0
        cdef uint64_t bases = 0
        _check(fa_sketch_build_stats(self._sk, &ms, &bases))
        return {"ms_sketch": ms, "bases": bases}

    cpdef Sketch clear(self):
        """clear(self)\n--

        Reset the `Sketch`, removing any reference genome it may contain (pyx:746-767)."""
        self._names.clear()
        if self._sk != NULL:
            _check(fa_sketch_clear(self._sk))
        return self

    cpdef Mapper index(self):
        """index(self)\n--

        Index the reference genomes for fast lookups using the minimizers (pyx:769-806).
        Ownership of the data moves to the returned `Mapper`; the sketch is left empty.
        """
        cdef Mapper mapper = Mapper.__new__(Mapper)
        cdef int rc
        with nogil:
            rc = fa_sketch_index(self._sk, &mapper._ix)
        _check(rc)
        mapper._param = self._param
        mapper._device = self._device
        mapper._names = self._names.copy()
        self._names = []
        return mapper


cdef int _restore_sketch(fa_sketch* sk, dict sketch_state, object lengths, uint64_t counter) except -1:
    cdef dict   mins    = sketch_state["minimizers"]
    cdef list   hashes  = mins["hashes"]
    cdef list   ids     = mins["ids"]
    cdef list   offsets = mins["offsets"]
    cdef size_t n       = mins["length"]
    cdef list   sbf     = list(sketch_state["sequencesByFileInfo"])
    cdef list   lens    = list(lengths)
    cdef size_t g       = len(sbf)
    cdef size_t i
    cdef uint32_t* h = <uint32_t*> malloc(max(n, 1) * sizeof(uint32_t))
    cdef int32_t*  s = <int32_t*> malloc(max(n, 1) * sizeof(int32_t))
    cdef int32_t*  w = <int32_t*> malloc(max(n, 1) * sizeof(int32_t))
    cdef int32_t*  q = <int32_t*> malloc(max(g, 1) * sizeof(int32_t))
    cdef uint64_t* l = <uint64_t*> malloc(max(g, 1) * sizeof(uint64_t))
    try:
        for i in range(n):
            h[i] = hashes[i]
            s[i] = ids[i]
            w[i] = offsets[i]
        for i in range(g):
            q[i] = sbf[i]
            l[i] = lens[i] if i < len(lens) else 0
        if g and counter < <uint64_t> q[g - 1]:
            counter = q[g - 1]
        _check(fa_sketch_restore(sk, h, s, w, n, q, l, g, counter))
    finally:
        free(h); free(s); free(w); free(q); free(l)
    return 0


# --- On-disk sketch (SURVEY.md 8f-2) ---------------------------------------------
# One uncompressed .npz: the minimizer triples as three columns (what `Sketch` / `Mapper` pickle as Python lists,
# pyx:572-591, 842-865), sequencesByFileInfo, the genome lengths, and a JSON header with the parameters and names.
# Loading uploads the columns and -- for a Mapper -- rebuilds the lookup index on the GPU, as unpickling does.

_SKETCH_FORMAT = "pyfastani_b200.sketch/1"


def _save_sketch_file(path, kind, dict parameters, list names, uint64_t counter, lengths, seqs_by_genome, Minimizers mins):
    import json
    import numpy
    try:
        header = json.dumps({"format": _SKETCH_FORMAT, "kind": kind, "parameters": parameters, "names": names, "counter": counter})
    except TypeError:
        raise TypeError("the on-disk sketch stores genome names as JSON (str / int / float / None); pickle the object for other names") from None
    h, s, w = mins.arrays()
    with open(path, "wb") as f:
        numpy.savez(f, header=numpy.frombuffer(header.encode("utf-8"), dtype=numpy.uint8), hashes=h, ids=s, offsets=w,
                    lengths=numpy.asarray(lengths, dtype=numpy.uint64), sequences_by_file=numpy.asarray(seqs_by_genome, dtype=numpy.int32))


def _load_sketch_file(path):
    import json
    import numpy
    with numpy.load(path, allow_pickle=False) as z:
        header = json.loads(bytes(z["header"]).decode("utf-8"))
        if header.get("format") != _SKETCH_FORMAT:
            raise ValueError("not a pyfastani_b200 sketch file: {!r}".format(header.get("format")))
        cols = {k: numpy.ascontiguousarray(z[k]) for k in ("hashes", "ids", "offsets", "lengths", "sequences_by_file")}
    if not (len(cols["hashes"]) == len(cols["ids"]) == len(cols["offsets"])) or len(cols["lengths"]) != len(cols["sequences_by_file"]):
        raise ValueError("inconsistent column lengths in sketch file")
    if len(header["names"]) != len(cols["lengths"]):
        raise ValueError("sketch file holds {} names for {} genomes".format(len(header["names"]), len(cols["lengths"])))
    return header, cols


cdef int _restore_sketch_columns(fa_sketch* sk, dict cols, uint64_t counter) except -1:
    cdef uint32_t[::1] h = cols["hashes"].astype("uint32", copy=False)
    cdef int32_t[::1]  s = cols["ids"].astype("int32", copy=False)
    cdef int32_t[::1]  w = cols["offsets"].astype("int32", copy=False)
    cdef int32_t[::1]  q = cols["sequences_by_file"].astype("int32", copy=False)
    cdef uint64_t[::1] l = cols["lengths"].astype("uint64", copy=False)
    cdef uint64_t n = h.shape[0], g = q.shape[0]
    cdef uint32_t h0 = 0
    cdef int32_t  i0 = 0
    cdef uint64_t l0 = 0
    if g and counter < <uint64_t> q[g - 1]:
        counter = q[g - 1]
    _check(fa_sketch_restore(sk, &h[0] if n else &h0, &s[0] if n else &i0, &w[0] if n else &i0, n,
                             &q[0] if g else &i0, &l[0] if g else &l0, g, counter))
    return 0


# --- Mapper -------------------------------------------------------------------

@cython.final
cdef class Mapper(_Parameterized):
    """A genome mapper using Murmur3 hashes and k-mers to compute ANI (pyx:809-1200)."""

    cdef          fa_index*  _ix
    cdef          list       _names
    cdef readonly Minimizers minimizers
    cdef readonly dict       last_query_info

    def __cinit__(self):
        self._ix = NULL
        self._names = []
        self.minimizers = Minimizers.__new__(Minimizers)
        self.minimizers._owner = self
        self.last_query_info = {}

    def __init__(self, *args, **kwargs):
        raise TypeError("Mapper cannot be instantiated, use `Sketch.index` instead.")

    def __dealloc__(self):
        if self._ix != NULL:
            fa_index_free(self._ix)
            self._ix = NULL

    def __getstate__(self):
        cdef uint64_t n_min = 0, n_unique = 0, n_contigs = 0, n_genomes = 0
        _check(fa_index_counts(self._ix, &n_min, &n_unique, &n_contigs, &n_genomes))
        cdef int32_t*  seqs = <int32_t*> malloc(max(n_genomes, 1) * sizeof(int32_t))
        cdef uint64_t* lens = <uint64_t*> malloc(max(n_genomes, 1) * sizeof(uint64_t))
        try:
            _check(fa_index_copy_meta(self._ix, seqs, lens))
            return {
                "parameters": _Parameterized.__getstate__(self),
                "lengths": [lens[i] for i in range(n_genomes)],
                "names": list(self._names),
                "device": self._device,
                "counter": n_contigs,
                "sketch": {
                    "sequencesByFileInfo": [seqs[i] for i in range(n_genomes)],
                    "minimizers": self.minimizers.__getstate__(),
                },
            }
        finally:
            free(seqs)
            free(lens)

    def __setstate__(self, state):
        # restore the minimizers into a scratch sketch, then rebuild the lookup index on the
        # GPU, like the reference rebuilds its hash table (pyx:853-865)
        cdef fa_sketch* sk = NULL
        _Parameterized.__setstate__(self, state["parameters"])
        self._device = state.get("device", 0)
        self._names = list(state["names"])
        _check(fa_sketch_create(&self._param, self._device, &sk))
        try:
            _restore_sketch(sk, state["sketch"], state["lengths"], state.get("counter", 0))
            if self._ix != NULL:
                fa_index_free(self._ix)
                self._ix = NULL
            _check(fa_sketch_index(sk, &self._ix))
        finally:
            fa_sketch_free(sk)

    def save(self, path):
        """save(self, path)\n--

        Write the indexed sketch to `path` (same file format as `Sketch.save`); `Mapper.load` rebuilds the lookup index
        on the GPU from the minimizer columns, as unpickling does (pyx:853-865), without sketching anything."""
        cdef uint64_t n_min = 0, n_unique = 0, n_contigs = 0, n_genomes = 0
        _check(fa_index_counts(self._ix, &n_min, &n_unique, &n_contigs, &n_genomes))
        cdef int32_t*  seqs = <int32_t*> malloc(max(n_genomes, 1) * sizeof(int32_t))
        cdef uint64_t* lens = <uint64_t*> malloc(max(n_genomes, 1) * sizeof(uint64_t))
        try:
            _check(fa_index_copy_meta(self._ix, seqs, lens))
            _save_sketch_file(path, "mapper", _Parameterized.__getstate__(self), list(self._names), n_contigs,
                              [lens[i] for i in range(n_genomes)], [seqs[i] for i in range(n_genomes)], self.minimizers)
        finally:
            free(seqs)
            free(lens)

    @classmethod
    def load(cls, path, device=None):
        """load(cls, path, device=None)\n--

        A `Mapper` over the content of a file written by `Mapper.save` or `Sketch.save`, indexed on `device`."""
        return Sketch.load(path, device).index()

    @property
    def lookup_index(self):
        """`MinimizerIndex`: The index of initial minimizer positions."""
        cdef MinimizerIndex index = MinimizerIndex.__new__(MinimizerIndex)
        index.owner = self
        return index

    @property
    def names(self):
        """`list`: The names of the indexed reference genomes (not in the reference API)."""
        return self._names[:]

    @property
    def build_stats(self):
        """`dict`: CUDA-event time of the index build (`Sketch.index`): all of it / the radix sort inside it."""
        cdef float ms_build = 0, ms_sort = 0
        _check(fa_index_build_stats(self._ix, &ms_build, &ms_sort))
        return {"ms_build": ms_build, "ms_sort": ms_sort}

    cdef list _query_draft(self, object contigs, int threads=0):
        cdef _Contigs      c = _Contigs.__new__(_Contigs)
        cdef uint64_t      cap = len(self._names)
        cdef uint64_t      n_out = 0
        cdef fa_hit*       out
        cdef fa_query_info info
        cdef int           rc
        cdef list          hits = []
        cdef uint64_t      i

        # `threads` is the reference's CPU worker count (pyx:1043-1050); fragments are mapped
        # by GPU threads here, so only its validation is kept
        if threads < 0:
            raise ValueError(f"`threads` must be positive or null, got {threads!r}")
        c.fill(contigs, self._device)
        out = <fa_hit*> malloc(max(cap, 1) * sizeof(fa_hit))
        if out == NULL:
            raise MemoryError()
        try:
            with nogil:
                rc = fa_query(self._ix, c.arr, c.n, out, cap, &n_out, &info)
            _check(rc)
            for _ in range(info.short_contigs):
                warnings.warn(
                    (
                        "Mapper received a short sequence relative to parameters, "
                        "mapping will not be computed."
                    ),
                    UserWarning,
                )
            for i in range(min(n_out, cap)):
                hits.append(_make_hit(self._names[out[i].ref_genome], out[i].identity, out[i].matches, out[i].fragments))
            self.last_query_info = {
                "fragments": info.fragments, "sketch_sum": info.sketch_sum, "seeds": info.seeds,
                "candidates": info.candidates, "scanned": info.scanned, "mappings": info.mappings,
                "kernel_launches": info.kernel_launches,
                "ms_h2d": info.ms_h2d, "ms_sketch": info.ms_sketch, "ms_lookup": info.ms_lookup,
                "ms_seed_sort": info.ms_seed_sort, "ms_l1": info.ms_l1, "ms_l2": info.ms_l2,
                "ms_cgi": info.ms_cgi, "ms_d2h": info.ms_d2h, "ms_total": info.ms_total,
                "h2d_bytes": info.h2d_bytes, "d2h_bytes": info.d2h_bytes, "l2_fallback": info.l2_fallback, "events": info.events,
                "ms_l2_prep": info.ms_l2_prep, "ms_l2_events": info.ms_l2_events, "ms_l2_slide": info.ms_l2_slide,
                "l1_sorted_fragments": info.l1_sorted_fragments, "l1_small_fragments": info.l1_small_fragments, "events_replayed": info.events_replayed,
                "ms_batch": info.ms_batch, "l1_parts": info.l1_parts, "l1_tiny_fragments": info.l1_tiny_fragments, "ms_exchange": info.ms_exchange, "l1_tiny_fragments": info.l1_tiny_fragments, "ms_exchange": info.ms_exchange, "queries": 1,
            }
        finally:
            free(out)
        return hits

    def query_many(self, object queries, int threads=0, bint rows=False, Communicator comm=None, object genome_offsets=None):
        """query_many(self, queries, threads=0, rows=False, comm=None, genome_offsets=None)\n--

        Map many queries with one call into the library (`fa_query_batch`; not in the reference
        API, which loops over `query_draft` in Python).  Each item of `queries` is either one
        sequence (`str`, bytes-like, `DeviceSequence`: a complete genome) or a list / tuple of
        contigs (a draft).  Returns one list of `Hit` per query, each exactly what
        `query_genome` / `query_draft` returns for that item.  Releases the GIL for the whole batch.

        ``rows=True`` returns numpy structured arrays ``(ref_genome, matches, fragments, identity)`` instead of
        `Hit` objects (``ref_genome`` = position in `names`).  With ``comm`` and ``genome_offsets`` (world + 1
        entries) the call is collective: this `Mapper` holds the reference genomes
        ``[genome_offsets[rank], genome_offsets[rank + 1])`` and every rank receives, per query, the rows of ALL
        ranks with global genome ids in the reference's order (`fa_query_batch_sharded`; implies ``rows``)."""
        cdef _Contigs      c = _Contigs.__new__(_Contigs)
        cdef list          flat = []
        cdef list          items = list(queries)
        cdef int32_t       nq = <int32_t> len(items)
        cdef uint64_t      n_refs = <uint64_t> max(len(self._names), 1)
        cdef uint64_t      cap
        cdef int32_t*      counts = NULL
        cdef int32_t*      goffs = NULL
        cdef uint64_t*     offs = NULL
        cdef fa_hit*       out = NULL
        cdef fa_comm*      cc = NULL
        cdef fa_query_info info
        cdef int           rc
        cdef list          result = []
        cdef list          hits
        cdef list          offs_l
        cdef uint64_t      i
        cdef int32_t       q

        if threads < 0:
            raise ValueError(f"`threads` must be positive or null, got {threads!r}")
        if comm is not None:
            if genome_offsets is None:
                raise ValueError("`genome_offsets` is required with `comm`")
            offs_l = list(genome_offsets)
            if len(offs_l) != comm.world_size + 1:
                raise ValueError("genome_offsets needs world_size + 1 entries")
            n_refs = <uint64_t> max(int(offs_l[-1]) - int(offs_l[0]), 1)
            cc = comm._c
        cap = n_refs * <uint64_t> max(nq, 1)
        counts = <int32_t*> malloc(max(nq, 1) * sizeof(int32_t))
        offs = <uint64_t*> malloc((nq + 1) * sizeof(uint64_t))
        out = <fa_hit*> malloc(cap * sizeof(fa_hit))
        try:
            if counts == NULL or offs == NULL or out == NULL:
                raise MemoryError()
            for q in range(nq):
                if isinstance(items[q], (list, tuple)):
                    counts[q] = <int32_t> len(items[q])
                    flat.extend(items[q])
                else:
                    counts[q] = 1
                    flat.append(items[q])
            c.fill(flat, self._device)
            if cc != NULL:
                goffs = <int32_t*> malloc((comm.world_size + 1) * sizeof(int32_t))
                if goffs == NULL:
                    raise MemoryError()
                for q in range(comm.world_size + 1):
                    goffs[q] = offs_l[q]
                with nogil:
                    rc = fa_query_batch_sharded(self._ix, cc, c.arr, counts, nq, goffs, out, cap, offs, &info)
            else:
                with nogil:
                    rc = fa_query_batch(self._ix, c.arr, counts, nq, out, cap, offs, &info)
            _check(rc)
            for _ in range(info.short_contigs):
                warnings.warn(
                    (
                        "Mapper received a short sequence relative to parameters, "
                        "mapping will not be computed."
                    ),
                    UserWarning,
                )
            if rows or cc != NULL:
                result = _hit_rows(out, offs, nq)
            else:
                for q in range(nq):
                    hits = []
                    for i in range(offs[q], offs[q + 1]):
                        hits.append(_make_hit(self._names[out[i].ref_genome], out[i].identity, out[i].matches, out[i].fragments))
                    result.append(hits)
            self.last_query_info = {
                "fragments": info.fragments, "seeds": info.seeds, "candidates": info.candidates, "mappings": info.mappings,
                "scanned": info.scanned, "sketch_sum": info.sketch_sum,
                "kernel_launches": info.kernel_launches, "ms_total": info.ms_total, "h2d_bytes": info.h2d_bytes,
                "d2h_bytes": info.d2h_bytes, "events": info.events, "events_replayed": info.events_replayed, "queries": nq,
                "ms_h2d": info.ms_h2d, "ms_sketch": info.ms_sketch, "ms_lookup": info.ms_lookup,
                "ms_seed_sort": info.ms_seed_sort, "ms_l1": info.ms_l1, "ms_l2": info.ms_l2, "ms_cgi": info.ms_cgi,
                "ms_d2h": info.ms_d2h, "ms_l2_prep": info.ms_l2_prep, "ms_l2_events": info.ms_l2_events,
                "ms_l2_slide": info.ms_l2_slide, "l1_small_fragments": info.l1_small_fragments,
                "l1_sorted_fragments": info.l1_sorted_fragments, "l2_fallback": info.l2_fallback, "ms_batch": info.ms_batch, "l1_parts": info.l1_parts, "l1_tiny_fragments": info.l1_tiny_fragments, "ms_exchange": info.ms_exchange,
            }
        finally:
            free(counts)
            free(offs)
            free(out)
            free(goffs)
        return result

    cpdef list query_draft(self, object contigs, int threads=0):
        """query_draft(self, contigs, threads=0)\n--

        Query the mapper for a draft genome (pyx:1138-1168).  Reentrant; releases the GIL."""
        return self._query_draft(contigs, threads=threads)

    cpdef list query_genome(self, object sequence, int threads=0):
        """query_genome(self, sequence, threads=0)\n--

        Query the mapper for a complete genome (pyx:1170-1200)."""
        return self._query_draft((sequence,), threads=threads)


# --- views ----------------------------------------------------------------------

cdef class Minimizers:
    """A read-only view over the minimizers of a `Sketch` or a `Mapper` (pyx:1203-1268).

    The data lives in GPU memory; elements are copied to the host on access.
    """

    cdef object _owner

    def __cinit__(self):
        self._owner = None

    cdef uint64_t _size(self) except? 0:
        cdef uint64_t n = 0
        cdef Sketch sk
        cdef Mapper mp
        if isinstance(self._owner, Sketch):
            sk = self._owner
            if sk._sk != NULL:
                _check(fa_sketch_counts(sk._sk, &n, NULL, NULL))
        elif isinstance(self._owner, Mapper):
            mp = self._owner
            if mp._ix != NULL:
                _check(fa_index_counts(mp._ix, &n, NULL, NULL, NULL))
        return n

    cdef int _copy(self, uint64_t first, uint64_t n, uint32_t* h, int32_t* s, int32_t* w) except -1:
        cdef Sketch sk
        cdef Mapper mp
        if isinstance(self._owner, Sketch):
            sk = self._owner
            _check(fa_sketch_copy_minimizers(sk._sk, first, n, h, s, w))
        else:
            mp = self._owner
            _check(fa_index_copy_minimizers(mp._ix, first, n, h, s, w))
        return 0

    def __len__(self):
        return self._size()

    def __getitem__(self, ssize_t index):
        cdef ssize_t  length = self._size()
        cdef ssize_t  index_ = index
        cdef uint32_t h = 0
        cdef int32_t  s = 0, w = 0
        if index_ < 0:
            index_ += length
        if index_ < 0 or index_ >= length:
            raise IndexError(index)
        self._copy(index_, 1, &h, &s, &w)
        return MinimizerInfo(h, s, w)

    def arrays(self):
        """arrays(self)\n--

        The minimizers as three NumPy arrays ``(hashes: uint32, sequence ids: int32, window positions: int32)`` --
        one device-to-host copy per column instead of one Python object per minimizer (the SoA form of the on-disk
        sketch, `Sketch.save`)."""
        import numpy
        cdef uint64_t n = self._size()
        h = numpy.empty(n, dtype=numpy.uint32)
        s = numpy.empty(n, dtype=numpy.int32)
        w = numpy.empty(n, dtype=numpy.int32)
        cdef uint32_t[::1] hv = h
        cdef int32_t[::1]  sv = s
        cdef int32_t[::1]  wv = w
        if n:
            self._copy(0, n, &hv[0], &sv[0], &wv[0])
        return h, s, w

    cpdef dict __getstate__(self):
        cdef uint64_t  n = self._size()
        cdef uint64_t  i
        cdef uint32_t* h = <uint32_t*> malloc(max(n, 1) * sizeof(uint32_t))
        cdef int32_t*  s = <int32_t*> malloc(max(n, 1) * sizeof(int32_t))
        cdef int32_t*  w = <int32_t*> malloc(max(n, 1) * sizeof(int32_t))
        try:
            if n:
                self._copy(0, n, h, s, w)
            return {
                "hashes": [h[i] for i in range(n)],
                "ids": [s[i] for i in range(n)],
                "offsets": [w[i] for i in range(n)],
                "length": n,
            }
        finally:
            free(h); free(s); free(w)


cdef class Hit:
    """A single hit found when querying a `Mapper` with a genome (pyx:1271-1324)."""

    cdef readonly object name
    cdef readonly int    matches
    cdef readonly int    fragments
    cdef readonly float  identity

    def __init__(self, object name, float identity, int matches, int fragments):
        """__init__(self, name, identity, matches, fragments)\n--

        Create a new `Hit` instance with the given parameters."""
        self.name = name
        self.matches = matches
        self.fragments = fragments
        self.identity = identity

    def __repr__(self):
        cdef str ty = type(self).__name__
        return "{}(name={!r}, identity={!r}, matches={!r}, fragments={!r})".format(
            ty, self.name, self.identity, self.matches, self.fragments
        )

    def __eq__(self, Hit other):
        return (
                self.name == other.name
            and self.matches == other.matches
            and self.fragments == other.fragments
            and self.identity == other.identity
        )

    def __reduce__(self):
        return (Hit, (self.name, self.identity, self.matches, self.fragments))


cdef class MinimizerInfo:
    """The information about a single minimizer (pyx:1327-1379)."""

    cdef readonly uint32_t hash
    cdef readonly int      sequence_id
    cdef readonly int      window_position

    def __init__(self, uint32_t hash, int sequence_id, int window_position):
        """__init__(self, hash, sequence_id, window_position)\n--

        Create a new `MinimizerInfo` with the given parameters."""
        self.hash = hash
        self.sequence_id = sequence_id
        self.window_position = window_position

    def __repr__(self):
        cdef str ty = type(self).__name__
        return "{}(hash={!r}, sequence_id={!r}, window_position={!r})".format(
            ty, self.hash, self.sequence_id, self.window_position
        )

    def __eq__(self, MinimizerInfo other):
        return (
                self.hash == other.hash
            and self.sequence_id == other.sequence_id
            and self.window_position == other.window_position
        )

    def __reduce__(self):
        return (MinimizerInfo, (self.hash, self.sequence_id, self.window_position))


cdef class Position:
    """A (sequence, window) position of a minimizer in the references (pyx:1382-1428)."""

    cdef readonly int sequence_id
    cdef readonly int window_position

    def __init__(self, int sequence_id, int window_position):
        """__init__(self, sequence_id, window_position)\n--

        Create a new `Position` instance with the given parameters."""
        self.sequence_id = sequence_id
        self.window_position = window_position

    def __repr__(self):
        cdef str ty = type(self).__name__
        return "{}(sequence_id={!r}, window_position={!r})".format(
            ty, self.sequence_id, self.window_position
        )

    def __eq__(self, Position other):
        return (
                self.sequence_id == other.sequence_id
            and self.window_position == other.window_position
        )

    def __reduce__(self):
        return (Position, (self.sequence_id, self.window_position))


cdef class MinimizerIndex:
    """The index mapping minimizer hash values to their positions (pyx:1431-1539).

    ``Mapper.lookup_index`` is a view over the CSR lookup table held in GPU memory: ``len``, iteration,
    ``in``, ``[]`` and ``items()`` read it; ``index[h] = positions`` and ``del index[h]`` rebuild the
    table around that entry (`fa_index_set_lookup` / `fa_index_del_lookup`), so the next query seeds
    from the edited table, as it does with the reference's host hash table.  Positions assigned must
    be positions of minimizers of the sketch (the table stores them as indices of the minimizer
    array); anything else raises `ValueError`.  A `MinimizerIndex()` created on its own is a plain
    host-side table, as in the reference.
    """

    cdef object owner
    cdef dict   _local

    def __cinit__(self):
        self.owner = None
        self._local = None

    def __init__(self):
        self._local = {}

    cdef fa_index* _index(self) except NULL:
        cdef Mapper mp
        if self.owner is None:
            raise ValueError("MinimizerIndex is not attached to a Mapper")
        mp = self.owner
        return mp._ix

    def __len__(self):
        cdef uint64_t n_unique = 0
        if self._local is not None:
            return len(self._local)
        _check(fa_index_counts(self._index(), NULL, &n_unique, NULL, NULL))
        return n_unique

    def _keys(self):
        cdef uint64_t  n
        cdef uint64_t  i
        cdef uint32_t* k
        if self._local is not None:
            return list(self._local)
        n = len(self)
        k = <uint32_t*> malloc(max(n, 1) * sizeof(uint32_t))
        try:
            if n:
                _check(fa_index_copy_keys(self._index(), 0, n, k))
            return [k[i] for i in range(n)]
        finally:
            free(k)

    def __iter__(self):
        return iter(self._keys())

    def __contains__(self, uint32_t item):
        cdef int32_t found = 0
        if self._local is not None:
            return item in self._local
        _check(fa_index_has_key(self._index(), item, &found))
        return found != 0

    def __getitem__(self, uint32_t item):
        cdef uint64_t n = 0
        cdef uint64_t i
        cdef int32_t* s
        cdef int32_t* w
        if self._local is not None:
            return list(self._local[item])
        if item not in self:
            raise KeyError(item)
        _check(fa_index_lookup(self._index(), item, NULL, NULL, 0, &n))
        if n == 0:
            return []
        s = <int32_t*> malloc(n * sizeof(int32_t))
        w = <int32_t*> malloc(n * sizeof(int32_t))
        try:
            _check(fa_index_lookup(self._index(), item, s, w, n, &n))
            return [Position(s[i], w[i]) for i in range(n)]
        finally:
            free(s); free(w)

    def __setitem__(self, uint32_t item, object value):
        cdef Position position
        cdef list     positions = []
        cdef uint64_t n, i
        cdef int32_t* s
        cdef int32_t* w
        for position in value:                    # (a non-Position raises TypeError here, as in the reference)
            positions.append(position)
        if self._local is not None:
            self._local[item] = positions
            return
        n = len(positions)
        s = <int32_t*> malloc(max(n, 1) * sizeof(int32_t))
        w = <int32_t*> malloc(max(n, 1) * sizeof(int32_t))
        try:
            for i in range(n):
                position = positions[i]
                s[i] = position.sequence_id
                w[i] = position.window_position
            _check(fa_index_set_lookup(self._index(), item, s, w, n))
        finally:
            free(s); free(w)

    def __delitem__(self, uint32_t item):
        cdef int32_t found = 0
        if self._local is not None:
            del self._local[item]
            return
        _check(fa_index_del_lookup(self._index(), item, &found))
        if not found:
            raise KeyError(item)

    def __reduce__(self):
        return (MinimizerIndex, (), None, None, self.items())

    def items(self):
        for key in self._keys():
            yield key, self[key]


cdef inline Hit _make_hit(object name, float identity, int matches, int fragments):
    """A `Hit` without the keyword-argument call of `Hit.__init__` (a query against a thousand genomes builds a
    thousand of them inside the timed path)."""
    cdef Hit h = Hit.__new__(Hit)
    h.name = name
    h.identity = identity
    h.matches = matches
    h.fragments = fragments
    return h
