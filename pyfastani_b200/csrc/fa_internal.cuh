// fa_internal.cuh -- the two opaque handles of the C ABI and the per-query workspace.
#pragma once
#include <mutex>

#include "fa_common.cuh"
#include "fa_stat.h"

namespace fa {

// L1 candidate region (L1_candidateLocus_t, FA/map/include/computeMap.hpp:40-49) in device form.
// The range is kept by the two seeds that define it, as indices of reference minimizers:
// rangeStartPos = max(0, wpos[hint] - fragLen + 1), rangeEndPos = wpos[tail] (computeMap.hpp:325-347) -- L2 turns
// them into index ranges with the per-minimizer offsets of fa_index.cu (slide_order_kernel), without a search.
struct Cand {
    int32_t  frag;     // query fragment (querySeqId)
    uint32_t hint;     // last seed of the pair that opened the region
    uint32_t tail;     // first seed of the last pair merged into it
    uint32_t spare;
};

// L2 result per candidate slot (L2_mapLocus_t + the MappingResult fields computeCGI reads,
// computeMap.hpp:52-59, base_types.hpp:89-102); shared < 0 marks "filtered out".
struct Mapping {
    int32_t seq;       // refSeqId
    int32_t ref_start; // refStartPos = meanOptimalPos
    int32_t shared;    // conservedSketches
    float   identity;  // nucIdentity
};

struct Workspace {
    SketchScratch sk;
    PinBuf stage;                       // pinned host staging for query bytes
    std::vector<SeqDesc> h_seqs;
    std::vector<int32_t> h_fragq;       // several queries in one pass: the query of every fragment
    DevBuf<int32_t>  frag_q;
    DevBuf<uint32_t> qhash;             // per-fragment minimizer hashes -> sorted unique sketches in place
    DevBuf<int32_t>  qs;                // per-fragment sketch size s
    DevBuf<uint32_t> hit_start, hit_cnt;   // per query hash: bucket in pos_idx
    DevBuf<uint64_t> frag_seeds;        // per fragment: seed count, then exclusive prefix (F + 1)
    DevBuf<uint64_t> seeds_a, seeds_b;  // (frag << bits | ref index) keys of the radix-sort L1 path, double buffered
    DevBuf<uint64_t> fb_seeds;          // seed prefix (F + 1) over the fragments that take the radix-sort path
    DevBuf<Cand>     cand_tmp;          // l1_fused_kernel: regions of fragment f at its seed offset
    DevBuf<uint32_t> l1_split, part_off, part_cands, l1_over;   // L1 in parts (fa_map.cu L1Parts)
    DevBuf<uint32_t> chunk_hist;        // hits per chunk of a sample of the heavy fragments (where to cut the parts)
    PinBuf h_chunk_hist;
    PinBuf hfs;                         // host copy of the seed prefixes
    DevBuf<uint8_t>  cub_tmp;
    DevBuf<uint32_t> frag_cands;        // per fragment: candidate count, then exclusive prefix (F + 1)
    DevBuf<uint32_t> work_base;         // per fragment: L2 work items, exclusive prefix (F + 1)
    DevBuf<Cand>     cands;
    DevBuf<Mapping>  maps;
    DevBuf<uint4>    prep;              // per candidate: the index searches of L2 (fa_map.cu Prep)
    DevBuf<uint64_t> ev_off;            // per candidate: padded event count, then exclusive prefix (C + 1)
    DevBuf<uint4>    jobs;              // per candidate: slide descriptor (fa_map.cu SlideJob, two entries each)
    DevBuf<uint32_t> mid;               // per candidate: begin / end index of the window its slide starts from
    DevBuf<uint32_t> seq_cnt;           // per fragment: minimizers the sketch kernel emitted into its slot of qhash
    DevBuf<uint16_t> events;            // the merged, classified insert/delete events of all candidates
    DevBuf<uint32_t> cells;             // per (ref contig, bin): best identity bits (computeCGI pass 2)
    DevBuf<float>    g_identity;        // per genome
    DevBuf<int32_t>  g_count;
    DevBuf<unsigned long long> counters;   // misc device counters (see fa_map.cu)
    PinBuf hres;                        // pinned result staging
    cudaEvent_t ev[12] = {};
    bool ev_ready = false;
    uint64_t last_cands = 0, last_frags = 0;
    // every buffer above, in one place: a member added to the struct is released here or nowhere
    void release()
    {
        sk.bytes.release(); sk.seqs.release(); sk.tile_status.release(); sk.counters.release(); sk.seq_first.release(); sk.drops.release();
        stage.release(); frag_q.release(); qhash.release(); qs.release(); hit_start.release(); hit_cnt.release();
        frag_seeds.release(); seeds_a.release(); seeds_b.release(); fb_seeds.release(); cand_tmp.release(); hfs.release();
        l1_split.release(); part_off.release(); part_cands.release(); l1_over.release(); chunk_hist.release(); h_chunk_hist.release();
        cub_tmp.release(); frag_cands.release(); work_base.release(); cands.release(); maps.release(); prep.release();
        ev_off.release(); jobs.release(); mid.release(); seq_cnt.release(); events.release(); cells.release(); g_identity.release();
        g_count.release(); counters.release(); hres.release();
        if (ev_ready) for (auto &e : ev) cudaEventDestroy(e);
        ev_ready = false;
    }
};

// The sketches of a group of queries as the reference-sharded layout exchanges them (fa_comm.cu sketch_exchange): every
// rank sketches `per` fragments of the group, all ranks gather all shares.  Fragment g of the group: `size` sorted
// unique hashes at recv[(g / per) * block + (g % per) * stride], size at recv[(g / per) * block + per * stride + g % per]
// (negative: the fragment could not travel -- see import_sketch_kernel).
struct PreSketch {
    const uint32_t *recv = nullptr;
    uint32_t per = 0, stride = 0;
    uint64_t block = 0;                 // 32-bit words per rank: per * (stride + 1)
    uint64_t first_frag = 0;            // first fragment of the pass inside the group
};
enum { FA_RETRY_PLAIN = 100 };          // internal status of run_queries: a pre-sketched fragment did not fit its slot, sketch this pass here

// Scratch of the sketch exchange (fa_comm.cu): it runs on its own stream, one group of queries ahead of the mapping, so it
// owns everything it touches -- staging, sketch scratch, this rank's packed share, and two receive buffers (the passes of
// group g read one while the shares of group g + 1 land in the other).
struct ExchScratch {
    SketchScratch sk;
    PinBuf stage;
    std::vector<SeqDesc> h_seqs;
    DevBuf<uint32_t> qhash, seq_cnt, send, recv[2];
    DevBuf<int32_t> qs;
    DevBuf<unsigned long long> counters;
    cudaStream_t st = nullptr;
    cudaEvent_t done[2] = {}, t0 = nullptr, t1 = nullptr;
    void release()
    {
        sk.bytes.release(); sk.seqs.release(); sk.tile_status.release(); sk.counters.release(); sk.seq_first.release(); sk.drops.release();
        stage.release(); qhash.release(); seq_cnt.release(); send.release(); recv[0].release(); recv[1].release(); qs.release(); counters.release();
        for (auto &e : done) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        if (st) cudaStreamDestroy(st);
        t0 = t1 = nullptr; st = nullptr;
    }
};

// One host or device buffer to place at `off` in the batch byte buffer (unit: fa_contig.unit_bytes).
struct Upload { const void *ptr; int32_t unit; int32_t on_device; int64_t len; uint64_t off; };
// The bytes of one query staged ahead of its turn (fa_query_batch): while query q is mapped, a helper thread copies
// query q + 1 through its own pinned buffer and copy stream into `bytes`; run_query then swaps `bytes` with the
// workspace's batch buffer and waits for `done` on its stream instead of staging.
struct Prefetch {
    DevBuf<uint8_t> bytes;
    PinBuf stage;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    const fa_contig *contigs = nullptr;
    int32_t n_contigs = 0;
    uint64_t total = 0, h2d_bytes = 0;
    bool valid = false;
    void release();
};
}  // namespace fa

struct fa_sketch {
    fa_params prm;
    int device = 0;
    cudaStream_t st = nullptr;
    fa::DevBuf<fa::RefMini> ref;                 // minimizerIndex, winSketch.hpp:93
    uint64_t n = 0;
    std::vector<int32_t>  seqs_by_genome;        // sequencesByFileInfo, winSketch.hpp:75
    std::vector<uint64_t> genome_len;            // Sketch._lengths, pyx:467
    uint64_t counter = 0;                        // Sketch._counter: sequence ids consumed (pyx:466, 683)
    uint64_t cur_len = 0;
    fa::SketchScratch sc;
    fa::PinBuf stage;
    std::vector<fa::SeqDesc> h_seqs;
    cudaEvent_t ev[2] = {};                      // around the device work of every add call
    bool ev_ready = false;
    double ms_sketch = 0;                        // staging + sketch kernels, summed over the add calls (CUDA events)
    uint64_t bases = 0;                          // bases seen by those calls
};

struct fa_index {
    fa_params prm;
    int device = 0;
    cudaStream_t st = nullptr;
    // position-ordered minimizers (minimizerIndex) and the hash -> positions table
    // (minimizerPosLookupIndex, winSketch.hpp:83-84) as CSR over sorted unique hashes
    fa::DevBuf<fa::RefMini> ref;
    fa::DevBuf<uint2> hw;                        // (hash, wpos | has-duplicate-nearby << 31): the 8-byte stream L2 reads
    fa::DevBuf<uint32_t> fb;                     // per minimizer: elements one fragment length behind (low 16 bits) / ahead (high 16 bits)
    fa::DevBuf<uint2> hl;                        // (hash, slide order word): the 8-byte stream the L2 events kernel reads (fa_index.cu slide_order_kernel)
    fa::DevBuf<uint32_t> gpos;                   // running coordinate for the L1 proximity test (fa_index.cu gpos_delta_kernel)
    fa::DevBuf<uint32_t> irr;                    // one bit per 1024 minimizers: some step of gpos inside is longer than the window
    uint64_t n = 0, n_unique = 0;
    fa::DevBuf<uint32_t> pos_idx;                // ref indices grouped by hash, insertion order inside a group
    fa::DevBuf<uint32_t> ukeys, uoff;            // unique hashes, group offsets (n_unique + 1)
    fa::DevBuf<uint32_t> dir;                    // directory over the top dir_bits of the hash (2^bits + 1)
    int dir_bits = 0;
    fa::DevBuf<uint32_t> contig_off;             // first ref index of each contig (n_contigs + 1)
    fa::DevBuf<int32_t>  genome_of_seq;          // reviseRefIdToGenomeId, computeCoreIdentity.hpp:31-42
    fa::DevBuf<uint32_t> bin_base;               // first (contig, bin) cell of each contig (n_contigs + 1)
    fa::DevBuf<uint32_t> genome_cell;            // first cell of each genome (n_genomes + 1)
    uint64_t n_cells = 0;
    std::vector<int32_t>  seqs_by_genome;
    std::vector<uint64_t> genome_len;
    uint64_t n_contigs = 0;
    // statistics tables (fa_stat.h)
    int s_max = 0;
    int max_min_hits = 1;                        // largest entry of d_min_hits
    fa::DevBuf<int32_t>  d_min_hits, d_min_shared;
    fa::DevBuf<uint32_t> d_id_off;
    fa::DevBuf<float>    d_identity;
    float ms_build = 0, ms_sort = 0;             // build_index on its stream (CUDA events): everything / the radix sort of (hash, index)
    long long l1_small_cap = -1;                 // test hook: most seeds per fragment for the small shape of the on-chip L1 (-1 = default)
    int l1_small_shape = -1;                     // test hook: smallest entry of L1S_SMEM the small shape may use (-1 = 0: the first that fits)
    int l1_parts = 0;                            // 0 = the large class is mapped whole (the default: measured faster), -1 = parts chosen from the workload, n = exactly n parts
    long long l1_part_cap = -1;                  // test hook: most hits a part may hold (-1 = what fits)
    std::vector<uint32_t> genome_first;          // first reference index of every genome (+ n), filled by the first query that maps in parts
    long long l1_tiny_cap = -1;                  // test hook: most hits per fragment for the warp-per-fragment L1 shape (-1 = 256, 0 = off)
    long long l1_seed_cap = -1;                  // test hook: most seeds per fragment for the on-chip L1 (-1 = what fits)
    std::mutex mtx;                              // serialises queries on the single workspace
    fa::Workspace ws;
    fa::ExchScratch xs;                          // sketch exchange of the reference-sharded layout (one sharded batch at a time)
    std::mutex pre_mtx;                          // one fa_query_batch at a time stages ahead (others map without)
    fa::Prefetch pre[2];
};

namespace fa {
// what every entry point accepts as a contig (include/fastani_b200.h fa_contig)
inline int check_contig(const fa_contig &ct, int c)
{
    if (ct.len < 0 || (ct.len > 0 && !ct.data)) { set_error("contig %d: bad buffer", c); return FA_ERR_INVALID; }
    if (ct.unit_bytes != 1 && ct.unit_bytes != 2 && ct.unit_bytes != 4 && ct.unit_bytes != FA_UNIT_PACKED2) {
        set_error("unit_bytes must be 1, 2, 4 or FA_UNIT_PACKED2"); return FA_ERR_INVALID;
    }
    if (ct.on_device && ct.unit_bytes != 1) { set_error("device-resident contigs must be bytes"); return FA_ERR_INVALID; }
    if (ct.unit_bytes == FA_UNIT_PACKED2 && ct.len > 0) {
        const fa_packed *pk = (const fa_packed *)ct.data;
        if (!pk->bits || (pk->n_runs && (!pk->run_pos || !pk->run_len || !pk->run_byte))) { set_error("contig %d: bad packed buffer", c); return FA_ERR_INVALID; }
    }
    return FA_OK;
}
inline bool contig_prenormalised(const fa_contig &ct) { return ct.unit_bytes == 2 || ct.unit_bytes == 4; }   // SeqDesc.raw
int build_index(fa_index *ix, int *launches);
// MinimizerIndex.__setitem__ / __delitem__: replace / insert / erase the position list of one hash (fa_index.cu)
int edit_lookup(fa_index *ix, uint32_t hash, const int32_t *seq, const int32_t *wpos, uint64_t m, bool erase, int *missing);
// the uploads of a query (whole fragments of every contig that is long enough, pyx:1059-1105) and their layout
void plan_uploads(const fa_params &P, const fa_contig *contigs, int32_t n_contigs, std::vector<Upload> &ups, uint64_t *total);
int prefetch_query(fa_index *ix, Prefetch &pf, const fa_contig *contigs, int32_t n_contigs);
int run_query(fa_index *ix, const fa_contig *contigs, int32_t n_contigs, fa_hit *out, uint64_t cap, uint64_t *n_out,
              fa_query_info *info, Prefetch *pf = nullptr);
int run_queries(fa_index *ix, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries, fa_hit *out,
                uint64_t cap, uint64_t *hit_offsets, fa_query_info *info, Prefetch *pf = nullptr, const PreSketch *ps = nullptr);
// reference-sharded layout: sketch this rank's share of the fragments of a group of queries (fragments
// [rank * per, rank * per + per) of all of them) and leave [per * stride hashes | per sizes] in ws.x_send
int exchange_stride(const fa_params &P);
int sketch_share(fa_index *ix, const fa_contig *contigs, int32_t n_contigs, int world, int rank, uint32_t stride, uint32_t *per_out,
                 uint64_t *frags_out, fa_query_info *qi);      // (on ix->xs.st, into ix->xs.send)
// fa_query_batch with an optional communicator: with one, the query sketches are made once across the ranks (fa_comm.cu)
int query_batch_impl(fa_index *ix, fa_comm *comm, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries,
                     fa_hit *out, uint64_t cap, uint64_t *hit_offsets, fa_query_info *info);
int comm_world(const fa_comm *c);
int sketch_exchange(fa_index *ix, fa_comm *comm, const fa_contig *contigs, int32_t n_contigs, int slot, PreSketch *ps, fa_query_info *qi);
// shared by the sketch and query paths: narrow/copy the uploads into the batch byte buffer
int stage_sequences(cudaStream_t st, DevBuf<uint8_t> &bytes, PinBuf &stage, const std::vector<Upload> &ups, uint64_t total,
                    uint64_t *h2d_bytes, int workers = 0);
int debug_candidates(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n);
int debug_mappings(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n);
}
