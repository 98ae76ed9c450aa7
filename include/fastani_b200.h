/*
 * fastani_b200.h -- C ABI of libfastani_b200.so, the B200 (sm_100a) implementation of
 * pyfastani's FastANI/MashMap mapping path.
 *
 * This is the drop-in boundary: each entry point replaces what the reference's Cython
 * layer (src/pyfastani/_fastani.pyx, "pyx" below) does through `cdef extern` bindings to
 * the skch::/cgi:: C++ headers (vendor/FastANI/src, "FA/" below).  Plain pointers and
 * sizes only; status-code returns (0 = ok), message via fa_last_error(); no C++
 * exceptions cross the boundary.  Threading: fa_sketch_* calls on one sketch must be
 * externally serialised (the reference holds a threading.Lock, pyx:563,715,742);
 * fa_query* may be called concurrently on one fa_index (pyx:1158-1161).
 *
 * There is no CPU fallback: every call that computes runs CUDA kernels and fails with
 * FA_ERR_CUDA when no device is usable.
 */
#ifndef FASTANI_B200_H
#define FASTANI_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define FA_API __attribute__((visibility("default")))
#else
#define FA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum {
    FA_OK = 0,
    FA_ERR_INVALID = 1,      /* bad argument (maps to ValueError, pyx:523-539) */
    FA_ERR_CUDA = 2,         /* CUDA runtime failure / no device */
    FA_ERR_NOMEM = 3,
    FA_ERR_UNSUPPORTED = 4,  /* parameter combination not implemented on the device path */
    FA_ERR_STATE = 5         /* call order (e.g. query on an empty index) */
};

/* skch::Parameters, FA/map/include/map_parameters.hpp:21-38 (the fields pyfastani sets,
 * pyx:374-378, 542-560). */
typedef struct fa_params {
    int32_t  k;             /* kmerSize */
    int32_t  window;        /* windowSize (0 = derive with fa_recommended_window) */
    int32_t  frag_len;      /* minReadLength */
    int32_t  alphabet;      /* alphabetSize: 4 (nucleotide) */
    float    min_fraction;  /* minFraction */
    float    pct_identity;  /* percentageIdentity */
    double   p_value;       /* p_value */
    uint64_t ref_size;      /* referenceSize */
} fa_params;

/* One sequence as the reference receives it (pyx:633-645, 1071-1095): `unit_bytes` is 1
 * for bytes / UCS1 str, 2 or 4 for UCS2 / UCS4 str.  `on_device` != 0 means `data` is a
 * device pointer on the index's device (unit_bytes must be 1); used to time the path
 * with inputs already resident in HBM.  `unit_bytes` == FA_UNIT_PACKED2 means `data` points
 * to an fa_packed (host memory, see below) holding `len` bases at two bits each. */
typedef struct fa_contig {
    const void *data;
    int32_t     unit_bytes;
    int32_t     on_device;
    int64_t     len;
} fa_contig;

/* A sequence packed to two bits per base for staging: a quarter of the bytes cross host memory and
 * PCIe, the device expands them in front of the sketch kernel (unpack_2bit in csrc/fa_map.cu).  Base i sits in bits
 * 2 (i & 3) .. 2 (i & 3) + 1 of bits[i >> 2]: A = 0, C = 1, G = 2, T = 3; a, c, g, t pack like their capitals (the
 * path upper-cases before it hashes, pyx:116-153).  Every other byte value -- N, IUPAC codes, anything -- is kept
 * exactly, as runs of one value: run r covers positions [run_pos[r], run_pos[r] + run_len[r]) with byte run_byte[r]
 * (ascending, disjoint; the two bits under a run are zero).  Results are identical to the unpacked bytes. */
typedef struct fa_packed {
    const uint8_t  *bits;       /* (len + 3) / 4 bytes */
    const uint32_t *run_pos;
    const uint32_t *run_len;
    const uint8_t  *run_byte;
    uint64_t        n_runs;
} fa_packed;
enum { FA_UNIT_PACKED2 = -2 };

/* cgi::CGI_Results, FA/cgi/include/cgid_types.hpp:68-80, after the min-fraction filter
 * and identity-descending stable sort of pyx:1121-1135. */
typedef struct fa_hit {
    int32_t ref_genome;     /* index into the genomes added to the sketch */
    int32_t matches;        /* countSeq */
    int32_t fragments;      /* totalQueryFragments */
    float   identity;
} fa_hit;

/* Realised workload counters and per-stage device times of one fa_query call
 * (the figures the roofline is computed from).  Times are CUDA-event
 * milliseconds on the library's own stream. */
typedef struct fa_query_info {
    uint64_t fragments;        /* F: query fragments mapped */
    uint64_t sketch_sum;       /* sum of per-fragment sketch sizes s */
    uint64_t seeds;            /* sum H: seed hits */
    uint64_t candidates;       /* sum C: L1 candidate regions */
    uint64_t scanned;          /* sum R: reference minimizers visited by L2 */
    uint64_t mappings;         /* P: L2 mappings that passed the identity filter */
    int32_t  short_contigs;    /* contigs that raise the "short sequence" warning, pyx:1062-1070 */
    int32_t  kernel_launches;  /* kernels launched by this call */
    float    ms_h2d, ms_sketch, ms_lookup, ms_seed_sort, ms_l1, ms_l2, ms_cgi, ms_d2h, ms_total;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t l2_fallback;      /* candidates taken by the exact fallback L2 kernel (long regions, huge sketches, bucket overflow) */
    uint64_t events;           /* insert/delete events replayed by the L2 slide kernel */
    float    ms_l2_prep, ms_l2_events, ms_l2_slide;   /* the three kernels inside ms_l2 */
    uint32_t l1_sorted_fragments;                     /* fragments whose seeds took the device-wide radix sort instead of the on-chip L1 */
    uint32_t l1_small_fragments;                      /* of the on-chip ones: fragments mapped by the small shape of the L1 kernel (256 threads, several CTAs per SM) */
    uint64_t events_replayed;  /* events the slide kernel went through before its early stop (<= events) */
    float    ms_batch;         /* ONE event pair around the whole call on the library's stream (fa_query: == ms_total) */
    uint32_t l1_parts;         /* parts the fragments of the large L1 class were cut into (0 = mapped whole) */
    uint32_t l1_tiny_fragments;/* fragments mapped by the warp-per-fragment shape of the L1 kernel (a few hundred hits at most) */
    float    ms_exchange;      /* fa_query_batch_sharded: device time of the sketch exchanges (own stream, one group of queries ahead of the mapping) */
} fa_query_info;

typedef struct fa_sketch fa_sketch;   /* skch::Sketch under construction (pyx:465-470) */
typedef struct fa_index  fa_index;    /* indexed skch::Sketch owned by a Mapper (pyx:821-824) */

/* -- library -------------------------------------------------------------------------- */
FA_API const char *fa_last_error(void);                  /* thread-local message of the last failure */
FA_API int fa_device_count(int32_t *n_out);
FA_API int fa_version(void);

/* Stat::recommendedWindowSize, FA/map/include/map_stats.hpp:226-256 (pyx:553-560). */
FA_API int fa_recommended_window(const fa_params *p, int32_t *w_out);
/* Stat::estimateMinimumHitsRelaxed (map_stats.hpp:142-167) and the identity / 90 % CI
 * filter of Map::doL2Mapping (computeMap.hpp:371-380); exported for the parity tests. */
FA_API int fa_stat_minimum_hits(int32_t s, int32_t k, float pct_identity, int32_t *out);
FA_API int fa_stat_l2(int32_t shared, int32_t s, int32_t k, float pct_identity, float *identity, int32_t *pass);
/* The row the device kernels read for sketch size s (tabulated once per (k, identity), csrc/fa_stat.cpp): minHits of
 * computeMap.hpp:312-313 and the smallest shared count that passes the filter of :380.  *irregular (optional) counts
 * the rows of the table, built for sizes 1..s_max, where a bisection had to fall back to the reference's linear walk. */
FA_API int fa_stat_table_row(int32_t s, int32_t s_max, int32_t k, float pct_identity, int32_t *min_hits, int32_t *min_shared,
                      int32_t *irregular);

/* -- Sketch (pyx:449-806) ---------------------------------------------------------------- */
FA_API int fa_sketch_create(const fa_params *p, int32_t device, fa_sketch **out);            /* pyx:476 */
FA_API void fa_sketch_free(fa_sketch *s);                                                    /* pyx:569-570 */
/* One iteration of the contig loop of Sketch._add_draft (pyx:629-683): sketches the contig
 * on the GPU (if long enough), consumes one sequence id.  *n_added (optional) receives the
 * number of minimizers appended, or -1 for the "short contig" warning case. */
FA_API int fa_sketch_add_contig(fa_sketch *s, const void *data, int32_t unit_bytes, int64_t len, int64_t *n_added);
/* pyx:686-690: closes the current genome; returns its fragment-rounded length. */
FA_API int fa_sketch_end_genome(fa_sketch *s, uint64_t *genome_len_out);
/* Batched form of the two calls above for one genome (one H2D + one launch sequence for
 * all contigs).  *n_short (optional) counts contigs that raise the warning. */
FA_API int fa_sketch_add_genome(fa_sketch *s, const fa_contig *contigs, int32_t n_contigs,
                         uint64_t *genome_len_out, int32_t *n_short);
/* Many genomes in one call (not in the reference API, which adds genome by genome from Python): genome g owns the next
 * contigs_per_genome[g] entries of `contigs`.  Whole genomes are sketched about 256 MB of bases per launch sequence.
 * genome_len_out (optional) has n_genomes entries. */
FA_API int fa_sketch_add_genomes(fa_sketch *s, const fa_contig *contigs, const int32_t *contigs_per_genome, int32_t n_genomes,
                          uint64_t *genome_len_out, int32_t *n_short);
/* CUDA-event time of the device work of all add calls so far (staging copies + kernels) and the bases they saw. */
FA_API int fa_sketch_build_stats(const fa_sketch *s, double *ms_sketch, uint64_t *bases);
FA_API int fa_sketch_clear(fa_sketch *s);                                                    /* pyx:746-767 */
FA_API int fa_sketch_counts(const fa_sketch *s, uint64_t *n_minimizers, uint64_t *n_contigs, uint64_t *n_genomes);
/* Minimizers view, pyx:1222-1254: copy [first, first+n) of (hash, seqId, wpos). */
FA_API int fa_sketch_copy_minimizers(const fa_sketch *s, uint64_t first, uint64_t n,
                              uint32_t *hash, int32_t *seq, int32_t *wpos);
/* Pickle support (pyx:572-591): bookkeeping out / everything back in. */
FA_API int fa_sketch_copy_meta(const fa_sketch *s, int32_t *seqs_by_genome, uint64_t *genome_len);
FA_API int fa_sketch_restore(fa_sketch *s, const uint32_t *hash, const int32_t *seq, const int32_t *wpos, uint64_t n,
                      const int32_t *seqs_by_genome, const uint64_t *genome_len, uint64_t n_genomes,
                      uint64_t n_contigs);
/* Sketch.index(), pyx:769-806 (Sketch::index + computeFreqHist, FA/map/include/
 * winSketch.hpp:177-244): builds the lookup index on the GPU; the data moves to *out and
 * the sketch is left empty but usable. */
FA_API int fa_sketch_index(fa_sketch *s, fa_index **out);

/* -- Mapper (pyx:809-1200) --------------------------------------------------------------- */
FA_API void fa_index_free(fa_index *ix);                                                     /* pyx:839-840 */
FA_API int fa_index_counts(const fa_index *ix, uint64_t *n_minimizers, uint64_t *n_unique,
                    uint64_t *n_contigs, uint64_t *n_genomes);                        /* pyx:1222, 1454-1456 */
FA_API int fa_index_params(const fa_index *ix, fa_params *out);
/* CUDA-event time of fa_sketch_index's device work: the whole build / the radix sort of (hash, position) inside it. */
FA_API int fa_index_build_stats(const fa_index *ix, float *ms_build, float *ms_sort);
FA_API int fa_index_copy_minimizers(const fa_index *ix, uint64_t first, uint64_t n,
                             uint32_t *hash, int32_t *seq, int32_t *wpos);           /* pyx:1225-1254 */
FA_API int fa_index_copy_meta(const fa_index *ix, int32_t *seqs_by_genome, uint64_t *genome_len);
/* MinimizerIndex view (pyx:1431-1539): keys in ascending hash order; positions of one hash
 * in insertion order (winSketch.hpp:180-185).  *n receives the bucket size (0 = KeyError). */
FA_API int fa_index_copy_keys(const fa_index *ix, uint64_t first, uint64_t n, uint32_t *keys);
FA_API int fa_index_lookup(const fa_index *ix, uint32_t hash, int32_t *seq, int32_t *wpos, uint64_t cap, uint64_t *n);
/* MinimizerIndex.__contains__ (pyx:1468-1471): a hash whose list was set to no positions is still a key. */
FA_API int fa_index_has_key(const fa_index *ix, uint32_t hash, int32_t *found);
/* MinimizerIndex.__setitem__ (pyx:1480-1497): the position list of `hash` becomes the n (seq, wpos) pairs given, in that
 * order; the hash is added when absent.  The reference edits the unordered_map that L1 seeding reads
 * (Sketch::minimizerPosLookupIndex); here the CSR table in device memory is rebuilt around the entry, so the next query
 * seeds from the new list.  A position that is not the position of a minimizer of the sketch, or given twice, is
 * FA_ERR_INVALID (the table stores positions as indices of the minimizer array). */
FA_API int fa_index_set_lookup(fa_index *ix, uint32_t hash, const int32_t *seq, const int32_t *wpos, uint64_t n);
/* MinimizerIndex.__delitem__ (pyx:1499-1507): *found = 0 when the hash was not a key (KeyError), nothing changes then. */
FA_API int fa_index_del_lookup(fa_index *ix, uint32_t hash, int32_t *found);
FA_API int fa_index_occurrence_threshold(const fa_index *ix, int32_t *out);                  /* getFreqThreshold, pyx:596-600 */
/* Mapper._query_draft, pyx:1006-1136: fragments the contigs, sketches them, L1 seeding,
 * L2 sliding Jaccard, computeCGI, min-fraction filter, sort.  Writes at most `cap` rows;
 * *n_out receives the number of hits.  `info` is optional. */
FA_API int fa_query(fa_index *ix, const fa_contig *contigs, int32_t n_contigs,
             fa_hit *out, uint64_t cap, uint64_t *n_out, fa_query_info *info);
/* Many queries in one call -- the throughput entry for many-to-many runs (BASELINE configs 3-5; the reference loops
 * over queries in Python, benches/mapping/bench.py:55-66): query q owns the next contigs_per_query[q] entries of
 * `contigs`; its hits are out[hit_offsets[q] .. hit_offsets[q + 1]) (hit_offsets has n_queries + 1 entries), in
 * the order fa_query returns them.  FA_ERR_INVALID if `cap` rows do not hold all hits (n_queries x genomes always
 * do).  `info` (may be NULL) receives the counters and stage times summed over the queries. */
FA_API int fa_query_batch(fa_index *ix, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries,
                   fa_hit *out, uint64_t cap, uint64_t *hit_offsets, fa_query_info *info);
/* -- multi-GPU: reference-sharded mapping (BASELINE configs[4]) ----------------------------------------
 * One process per GPU.  Whole reference genomes are dealt to the ranks in contiguous blocks (genome_offsets, world + 1
 * entries: rank r indexes genomes [genome_offsets[r], genome_offsets[r + 1])) and every rank maps every query against
 * its shard -- what upstream FastANI does per thread (splitReferenceGenomes / correctRefGenomeIds,
 * FA/cgi/include/computeCoreIdentity.hpp:454-484).  Two exchange steps over NCCL, both inside fa_query_batch_sharded:
 * the query sketches are made ONCE across the ranks (rank r sketches 1/world of the fragments of a group of queries, one
 * ncclAllGather of the packed sketches per group, issued a group ahead of the mapping on its own stream), and the final
 * per-genome rows are gathered at the end of the call (one small ncclAllGather; a second one only when a rank holds
 * more than 2048 rows for the batch).  NCCL is bound at run time (libnccl.so.2).  Calls on one communicator -- and
 * sharded calls on one index -- must not overlap: collectives are matched by issue order. */
typedef struct fa_comm fa_comm;
enum { FA_COMM_ID_BYTES = 128 };
/* ncclGetUniqueId on one rank; the 128 bytes reach the other ranks by whatever channel the host has (a socket, a file,
 * MPI, torch.distributed), then every rank calls fa_comm_create (ncclCommInitRank; collective). */
FA_API int fa_comm_unique_id(uint8_t *id);
FA_API int fa_comm_create(const uint8_t *id, int32_t world, int32_t rank, int32_t device, fa_comm **out);
FA_API void fa_comm_free(fa_comm *c);
FA_API int fa_comm_info(const fa_comm *c, int32_t *world, int32_t *rank, int32_t *nccl_version, uint64_t *collectives,
                 uint64_t *bytes_gathered);
/* Collective.  This rank's rows of n_queries queries (rows[hit_offsets[q] .. hit_offsets[q + 1]), LOCAL genome ids, in
 * the order fa_query returns them) -> on every rank the rows of all ranks with GLOBAL genome ids, per query in the order
 * of pyx:1135 (identity descending, stable in ascending genome id): out[out_offsets[q] .. out_offsets[q + 1]). */
FA_API int fa_gather_hits(fa_comm *c, const fa_hit *rows, const uint64_t *hit_offsets, int32_t n_queries,
                   const int32_t *genome_offsets, fa_hit *out, uint64_t cap, uint64_t *out_offsets);
/* Collective.  fa_query_batch against this rank's shard -- with the query sketches shared between the ranks as described
 * above (FA_NO_SKETCH_EXCHANGE=1 in the environment of ALL ranks: every rank sketches every query) -- followed by
 * fa_gather_hits: the whole reference-sharded step without the host language in between.  Every rank passes the same
 * queries in the same order. */
FA_API int fa_query_batch_sharded(fa_index *ix, fa_comm *comm, const fa_contig *contigs, const int32_t *contigs_per_query,
                           int32_t n_queries, const int32_t *genome_offsets, fa_hit *out, uint64_t cap,
                           uint64_t *hit_offsets, fa_query_info *info);
/* Intermediates of the last fa_query on this index, for the bit-exact parity tests
 * (SURVEY.md 7.1 step 0): L1 candidates as (frag, seq, start, end) rows and L2 mappings as
 * (frag, seq, refStartPos, shared, sketch, identity-bits) rows of int32. */
FA_API int fa_debug_last_candidates(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n);
FA_API int fa_debug_last_mappings(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n);
/* Test hook: cap on the seeds per fragment the on-chip L1 kernel accepts (fragments above it take the
 * device-wide radix-sort path); -1 restores the default (whatever fits in shared memory). */
FA_API int fa_debug_set_l1_seed_cap(fa_index *ix, int64_t cap);
/* Test hook: cap on the seeds per fragment the small shape of the on-chip L1 kernel accepts (0 = every on-chip
 * fragment takes the large shape); -1 restores the default (what leaves room for four CTAs per SM). */
FA_API int fa_debug_set_l1_small_cap(fa_index *ix, int64_t cap);
/* Test hook: the small shape of the on-chip L1 kernel has three shared-memory sizes (four, three or two CTAs per SM; the
 * larger ones serve indexes whose chunk histogram leaves no room otherwise, i.e. thousands of genomes); `shape` = 0, 1, 2
 * is the smallest one it may pick, -1 restores the default. */
FA_API int fa_debug_set_l1_small_shape(fa_index *ix, int32_t shape);
/* -- ingestion --------------------------------------------------------------------------------------
 * Host helper: packs `len` bytes into the fa_packed layout.  `bits` has (len + 3) / 4 bytes; the run arrays have
 * `run_cap` entries; *n_runs receives the number of runs the sequence has -- when it exceeds run_cap only the first
 * run_cap were stored (call again with larger arrays).  len < 2^32. */
FA_API int fa_pack_2bit(const uint8_t *data, uint64_t len, uint8_t *bits, uint32_t *run_pos, uint32_t *run_len,
                 uint8_t *run_byte, uint64_t run_cap, uint64_t *n_runs);
/* The inverse on the host (pickling, tests). */
FA_API int fa_unpack_2bit(const fa_packed *p, uint64_t len, uint8_t *out);
/* FASTA text parsed on the device: what the reference's Parser does line by line on the host
 * (src/pyfastani/_fasta.pyx:41-103 -- a record starts at a line that begins with '>', its id is the rest of that line,
 * its sequence the following lines up to the next such line with their '\n' removed and letters upper-cased; a text
 * that does not start with '>' has no records).  The text is uploaded once; header lines and newlines are dropped by a
 * stream compaction on the GPU and the records stay in device memory, ready to be passed as on_device contigs to
 * fa_sketch_add_genome / fa_query without ever existing as host strings.
 * Differences from the reference parser, both outside what a FASTA file holds: its 2048-byte line buffer treats a '>'
 * at a multiple of 2047 bytes into a longer line as a header, and its SSE2 upper-casing clears bit 5 of non-letters
 * (digits, punctuation) in the 16-byte blocks of a line. */
typedef struct fa_fasta fa_fasta;
FA_API int fa_fasta_parse(int32_t device, const void *text, uint64_t len, fa_fasta **out);
FA_API void fa_fasta_free(fa_fasta *f);
FA_API int fa_fasta_counts(const fa_fasta *f, uint64_t *n_records, uint64_t *n_bases);
/* contigs[r] = the sequence of record r as a device-resident contig (valid until fa_fasta_free); id_begin[r] / id_len[r]
 * locate its identifier in the text the caller passed (the bytes between '>' and the end of the line). */
FA_API int fa_fasta_records(const fa_fasta *f, fa_contig *contigs, uint64_t *id_begin, uint64_t *id_len);

/* Test hook: cap on the hits per fragment the warp-per-fragment shape of the L1 kernel accepts (0 = off, at most 256);
 * -1 restores the default (256). */
FA_API int fa_debug_set_l1_tiny_cap(fa_index *ix, int64_t cap);
/* How the on-chip L1 kernel maps fragments with tens of thousands of hits.  parts = 0: whole, one CTA per fragment (the
 * default -- measured faster on BASELINE configs[1]); -1: cut at genome boundaries into as many parts as the workload
 * asks for, each mapped by its own small CTA; n: exactly n parts.  part_cap: most hits a part may hold before its
 * fragment falls back to the whole shape (-1 = what fits in shared memory).  Same results either way. */
FA_API int fa_debug_set_l1_parts(fa_index *ix, int32_t parts, int64_t part_cap);
/* Device buffers for callers that want inputs resident in HBM before the timed region
 * (fa_contig.on_device). */
FA_API int fa_device_alloc(int32_t device, uint64_t bytes, void **dptr);
FA_API int fa_device_upload(int32_t device, void *dptr, const void *src, uint64_t bytes);
FA_API int fa_device_download(int32_t device, void *dst, const void *dptr, uint64_t bytes);
FA_API int fa_device_free(int32_t device, void *dptr);
FA_API int fa_device_mem_info(int32_t device, uint64_t *free_bytes, uint64_t *total_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FASTANI_B200_H */
