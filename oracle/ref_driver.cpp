// ref_driver.cpp -- thin C shim over the REFERENCE's own headers, compiled in
// place from /root/reference (never copied into this repo) by oracle/Makefile
// into oracle/_ref/libfastani_ref.so.
//
// TEST INFRASTRUCTURE ONLY: used to validate oracle/fastani_oracle.c and as the
// "reference" CPU baseline of bench.py.  The product never links this.
//
// Everything algorithmic below is a call into the reference:
//   skch::CommonFunc::addMinimizers      FA/map/include/commonFunc.hpp:91-167
//   skch::Sketch::index/computeFreqHist  FA/map/include/winSketch.hpp:177-244
//   skch::Map::computeL1CandidateRegions FA/map/include/computeMap.hpp:310-350
//   skch::Map::doL2Mapping               FA/map/include/computeMap.hpp:360-406
//   skch::Stat::*                        FA/map/include/map_stats.hpp
//   cgi::computeCGI                      FA/cgi/include/computeCoreIdentity.hpp:163-295
// The per-fragment driver mirrors Mapper._query_fragment / _do_l1_mappings
// (src/pyfastani/_fastani.pyx:885-1004) and the fragment loop + hit filter of
// Mapper._query_draft (pyx:1006-1136), with the ThreadPool replaced by
// std::thread workers over fragments (same decomposition, pyx:1099-1102).
//
// Caveat: upstream addMinimizers complements ACGT only and
// upper-cases a-z only, whereas pyfastani's Cython copy uses an IUPAC table and
// an SSE2 `& ~0x20`; the two agree on ACGTN/acgtn input.  Golden vectors from
// the real pyfastani (tests/golden) cover the rest.
//
// Build: g++ -O3 -std=c++11 -fno-access-control -DUSE_BOOST=1 -DBOOST_MATH_STANDALONE=1
//        -I<ref>/src/FastANI (stub omp.h) -I<ref>/vendor/FastANI/src -I<ref>/vendor/boost-math/include
#include <cstdint>
#include <cstring>
#include <vector>
#include <tuple>
#include <string>
#include <chrono>
#include <fstream>
#include <iostream>
#include <limits>
#include <functional>
#include <algorithm>
#include <thread>
#include <mutex>
#include <atomic>

#include "map/include/base_types.hpp"
#include "map/include/map_parameters.hpp"
#include "map/include/commonFunc.hpp"
#include "map/include/winSketch.hpp"
#include "map/include/map_stats.hpp"
#include "map/include/computeMap.hpp"
#include "cgi/include/computeCoreIdentity.hpp"

// the reference silences its logging the same way (src/FastANI/omp.cpp:5-13)
extern "C" {
int omp_get_thread_num(void) { return 1; }
int omp_get_num_threads(void) { return 1; }
void omp_set_num_threads(int) {}
}

extern "C" {

typedef struct {
    int32_t k, window, frag_len, alphabet;
    float min_fraction, pct_identity;
    double p_value;
    uint64_t ref_size;
} ref_params;

typedef struct { int32_t frag, seq, start, end; } ref_cand;
typedef struct { int32_t frag, seq, ref_start, shared, sketch; float identity; } ref_mapping;
typedef struct { int32_t ref_genome, matches, fragments; float identity; } ref_hit;
typedef struct { const void *data; int32_t unit_bytes; int64_t len; } ref_contig;
typedef struct { uint64_t fragments, seeds, candidates, scanned, mappings, sketch_sum; } ref_stats;

struct ref_sketch {
    skch::Parameters param;
    skch::Sketch *sk;
    std::vector<uint64_t> lengths;   // Sketch._lengths, pyx:467
    size_t counter;                  // Sketch._counter, pyx:466
    uint64_t cur_len;
};

static void fill_params(skch::Parameters &p, const ref_params *in)
{
    p.kmerSize = in->k; p.windowSize = in->window; p.minReadLength = in->frag_len;
    p.alphabetSize = in->alphabet; p.minFraction = in->min_fraction;
    p.percentageIdentity = in->pct_identity; p.p_value = in->p_value; p.referenceSize = in->ref_size;
    p.threads = 1; p.reportAll = true; p.visualize = false; p.matrixOutput = false;   // pyx:374-378
    p.outFileName = "/dev/null";
    p.querySequences.assign(1, "/dev/null");   // lets the unpatched Map ctor run on an empty file
}

int ref_recommended_window(const ref_params *in)
{
    return skch::Stat::recommendedWindowSize(in->p_value, in->k, in->alphabet, in->pct_identity,
                                             in->frag_len, in->ref_size);
}
int ref_minimum_hits_relaxed(int s, int k, float pid) { return skch::Stat::estimateMinimumHitsRelaxed(s, k, pid); }
int ref_l2_stat(int shared, int s, int k, float pid, float *identity)
{   // computeMap.hpp:371-380
    float mash_dist = skch::Stat::j2md(1.0 * shared / s, k);
    float lower = skch::Stat::md_lower_bound(mash_dist, s, k, 0.9);
    float nucIdentity = 100 * (1 - mash_dist);
    float upper = 100 * (1 - lower);
    if (identity) *identity = nucIdentity;
    return upper >= pid;
}
uint32_t ref_hash(const char *seq, int len) { return skch::CommonFunc::getHash(seq, len); }

ref_sketch *ref_sketch_new(const ref_params *in)
{
    ref_sketch *s = new ref_sketch();
    fill_params(s->param, in);
    s->sk = new skch::Sketch(s->param);     // refSequences is empty: build() reads nothing
    s->sk->minimizerPosLookupIndex.clear();
    s->sk->minimizerFreqHistogram.clear();
    s->sk->freqThreshold = std::numeric_limits<int>::max();
    s->counter = 0; s->cur_len = 0;
    return s;
}
void ref_sketch_free(ref_sketch *s) { if (s) { delete s->sk; delete s; } }

static void narrow(std::string &dst, const void *data, int unit, int64_t len)
{
    dst.resize(len);
    if (unit == 1) memcpy(&dst[0], data, len);
    else for (int64_t i = 0; i < len; i++) {
        uint32_t cp = unit == 2 ? ((const uint16_t *)data)[i] : ((const uint32_t *)data)[i];
        dst[i] = (char)cp;
    }
}

int64_t ref_sketch_add_contig(ref_sketch *s, const void *data, int unit, int64_t slen)
{   // Sketch._add_draft loop body, pyx:629-683, via the upstream addMinimizers
    int64_t n = -1;
    if (slen >= s->param.windowSize && slen >= s->param.kmerSize) {
        std::string buf; narrow(buf, data, unit, slen);       // addMinimizers upper-cases in place
        kseq_t ks; memset(&ks, 0, sizeof ks);
        ks.seq.s = &buf[0]; ks.seq.l = slen;
        size_t before = s->sk->minimizerIndex.size();
        skch::CommonFunc::addMinimizers(s->sk->minimizerIndex, &ks, s->param.kmerSize, s->param.windowSize,
                                        s->param.alphabetSize, (skch::seqno_t)s->counter);
        n = (int64_t)(s->sk->minimizerIndex.size() - before);
    }
    s->cur_len += (uint64_t)(slen / s->param.minReadLength) * s->param.minReadLength;
    s->counter++;
    return n;
}
// Set-up helper for the CPU baseline of bench.py: many single-contig genomes at once, each sketched by the reference's
// addMinimizers into a private vector on its own thread, then appended in genome order -- the same minimizers, ids and
// bookkeeping as calling ref_sketch_add_contig + ref_sketch_end_genome genome by genome.
void ref_sketch_add_genomes(ref_sketch *s, const ref_contig *genomes, int32_t n_genomes, int threads)
{
    std::vector<std::vector<skch::MinimizerInfo>> parts(n_genomes);
    std::atomic<int32_t> next(0);
    const size_t base = s->counter;
    auto work = [&]() {
        for (;;) {
            const int32_t g = next.fetch_add(1);
            if (g >= n_genomes) break;
            const int64_t slen = genomes[g].len;
            if (slen >= s->param.windowSize && slen >= s->param.kmerSize) {
                std::string buf; narrow(buf, genomes[g].data, genomes[g].unit_bytes, slen);
                kseq_t ks; memset(&ks, 0, sizeof ks);
                ks.seq.s = &buf[0]; ks.seq.l = slen;
                skch::CommonFunc::addMinimizers(parts[g], &ks, s->param.kmerSize, s->param.windowSize,
                                                s->param.alphabetSize, (skch::seqno_t)(base + g));
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, std::min(threads, (int)n_genomes)); t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    for (int32_t g = 0; g < n_genomes; g++) {
        s->sk->minimizerIndex.insert(s->sk->minimizerIndex.end(), parts[g].begin(), parts[g].end());
        std::vector<skch::MinimizerInfo>().swap(parts[g]);
        s->cur_len = (uint64_t)(genomes[g].len / s->param.minReadLength) * s->param.minReadLength;
        s->counter++;
        s->lengths.push_back(s->cur_len); s->cur_len = 0;
        s->sk->sequencesByFileInfo.push_back((skch::seqno_t)s->counter);
    }
}
void ref_sketch_end_genome(ref_sketch *s)
{   // pyx:686-690
    s->lengths.push_back(s->cur_len); s->cur_len = 0;
    s->sk->sequencesByFileInfo.push_back((skch::seqno_t)s->counter);
}
void ref_sketch_index(ref_sketch *s)
{   // pyx:790-791
    s->sk->minimizerPosLookupIndex.clear();
    s->sk->minimizerFreqHistogram.clear();
    s->sk->index();
    s->sk->computeFreqHist();
}
uint64_t ref_sketch_size(const ref_sketch *s) { return s->sk->minimizerIndex.size(); }
uint64_t ref_sketch_unique(const ref_sketch *s) { return s->sk->minimizerPosLookupIndex.size(); }
void ref_sketch_copy(const ref_sketch *s, uint32_t *hash, int32_t *seq, int32_t *wpos)
{
    size_t i = 0;
    for (auto &e : s->sk->minimizerIndex) { hash[i] = e.hash; seq[i] = e.seqId; wpos[i] = e.wpos; i++; }
}

int64_t ref_minimizers(const void *data, int unit, int64_t slen, int k, int w, int32_t seq,
                       uint32_t *hash, int32_t *seqs, int32_t *wpos, int64_t cap)
{
    std::vector<skch::MinimizerInfo> v;
    if (slen >= w && slen >= k) {
        std::string buf; narrow(buf, data, unit, slen);
        kseq_t ks; memset(&ks, 0, sizeof ks);
        ks.seq.s = &buf[0]; ks.seq.l = slen;
        skch::CommonFunc::addMinimizers(v, &ks, k, w, 4, seq);
    }
    for (size_t i = 0; i < v.size() && (int64_t)i < cap; i++) { hash[i] = v[i].hash; seqs[i] = v[i].seqId; wpos[i] = v[i].wpos; }
    return (int64_t)v.size();
}

int64_t ref_query(ref_sketch *s, const ref_contig *contigs, int32_t n_contigs,
                  ref_hit *hits, int64_t hit_cap, int32_t *short_contigs,
                  ref_cand *cand_out, int64_t cand_cap, int64_t *n_cand,
                  ref_mapping *map_out, int64_t map_cap, int64_t *n_map,
                  int64_t frag_first, int64_t frag_count, ref_stats *stats, int threads)
{
    typedef skch::QueryMetaData<kseq_t *, skch::Sketch::MI_Type> Q_t;
    skch::Parameters &param = s->param;
    uint64_t dummy = 0;
    skch::Map map(param, *s->sk, dummy, 0);                  // pyx:1053-1054
    skch::MappingResultsVector_t final_mappings;             // pyx:1055 (atomic_vector)
    std::vector<ref_cand> all_cands;
    std::mutex mtx;
    std::atomic<uint64_t> nseeds(0), ssum(0);
    uint64_t total_frags = 0, total_len = 0;
    int32_t shorts = 0;
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;

    for (int32_t c = 0; c < n_contigs; c++) {
        int64_t slen = contigs[c].len;
        if (slen < std::min(std::min(param.windowSize, param.kmerSize), param.minReadLength)) { shorts++; continue; }
        std::string buf; narrow(buf, contigs[c].data, contigs[c].unit_bytes, slen);
        int64_t nfrag = slen / param.minReadLength;          // pyx:1097
        std::atomic<int64_t> next(0);
        uint64_t base = total_frags;
        auto worker = [&]() {
            std::vector<ref_cand> my_cands;
            skch::MappingResultsVector_t my_maps;
            std::string frag;
            for (;;) {
                int64_t i = next.fetch_add(1);
                if (i >= nfrag) break;
                int64_t gid = (int64_t)base + i;
                if (frag_count >= 0 && (gid < frag_first || gid >= frag_first + frag_count)) continue;
                frag.assign(buf, i * param.minReadLength, param.minReadLength);
                kseq_t ks; memset(&ks, 0, sizeof ks);
                ks.seq.s = &frag[0]; ks.seq.l = param.minReadLength;
                Q_t Q; Q.kseq = &ks; Q.seqCounter = (skch::seqno_t)gid;                 // pyx:982-985
                // _do_l1_mappings, pyx:906-952
                skch::CommonFunc::addMinimizers(Q.minimizerTableQuery, Q.kseq, param.kmerSize, param.windowSize,
                                                param.alphabetSize, 0);
                std::sort(Q.minimizerTableQuery.begin(), Q.minimizerTableQuery.end(), skch::MinimizerInfo::lessByHash);
                auto uniq_end = std::unique(Q.minimizerTableQuery.begin(), Q.minimizerTableQuery.end(),
                                            skch::MinimizerInfo::equalityByHash);
                Q.sketchSize = std::distance(Q.minimizerTableQuery.begin(), uniq_end);
                if (Q.sketchSize == 0) continue;
                ssum += (uint64_t)Q.sketchSize;
                std::vector<skch::MinimizerMetaData> seeds;
                for (auto it = Q.minimizerTableQuery.begin(); it != uniq_end; ++it) {
                    auto f = s->sk->minimizerPosLookupIndex.find(it->hash);
                    if (f != s->sk->minimizerPosLookupIndex.end() && (int)f->second.size() < s->sk->getFreqThreshold())
                        seeds.insert(seeds.end(), f->second.begin(), f->second.end());
                }
                nseeds += seeds.size();
                int min_hits = skch::Stat::estimateMinimumHitsRelaxed(Q.sketchSize, param.kmerSize, param.percentageIdentity);
                std::vector<skch::Map::L1_candidateLocus_t> l1;
                map.computeL1CandidateRegions(Q, seeds, min_hits, l1);
                for (auto &e : l1) my_cands.push_back(ref_cand{(int32_t)gid, e.seqId, e.rangeStartPos, e.rangeEndPos});
                map.doL2Mapping(Q, l1, my_maps);                                        // pyx:998-1002
            }
            std::lock_guard<std::mutex> g(mtx);
            all_cands.insert(all_cands.end(), my_cands.begin(), my_cands.end());
            final_mappings.insert(final_mappings.end(), my_maps.begin(), my_maps.end());
        };
        if (threads == 1) worker();
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++) pool.emplace_back(worker);
            for (auto &t : pool) t.join();
        }
        total_frags += (uint64_t)nfrag;                      // pyx:1104
        total_len += (uint64_t)slen;                         // pyx:1105
    }

    // deterministic dumps regardless of thread interleaving
    std::sort(all_cands.begin(), all_cands.end(), [](const ref_cand &a, const ref_cand &b) {
        return std::tie(a.frag, a.seq, a.start) < std::tie(b.frag, b.seq, b.start); });
    std::vector<ref_mapping> maps;
    for (auto &e : final_mappings)
        maps.push_back(ref_mapping{e.querySeqId, e.refSeqId, e.refStartPos, e.conservedSketches, e.sketchSize, e.nucIdentity});
    std::sort(maps.begin(), maps.end(), [](const ref_mapping &a, const ref_mapping &b) {
        return std::tie(a.frag, a.seq, a.ref_start) < std::tie(b.frag, b.seq, b.ref_start); });

    std::vector<cgi::CGI_Results> results;
    std::string fname;
    cgi::computeCGI(param, final_mappings, map, *s->sk, total_frags, 0, fname, results);   // pyx:1108-1118

    std::vector<ref_hit> out;
    for (auto &r : results) {                                // pyx:1121-1132
        uint64_t min_length = std::min(total_len, s->lengths[r.refGenomeId]);
        uint64_t shared_length = (uint64_t)r.countSeq * param.minReadLength;
        if (shared_length >= min_length * param.minFraction)
            out.push_back(ref_hit{r.refGenomeId, r.countSeq, r.totalQueryFragments, r.identity});
    }
    std::stable_sort(out.begin(), out.end(), [](const ref_hit &a, const ref_hit &b) { return a.identity > b.identity; });  // pyx:1135

    for (size_t i = 0; i < out.size() && (int64_t)i < hit_cap; i++) hits[i] = out[i];
    if (short_contigs) *short_contigs = shorts;
    if (n_cand) *n_cand = (int64_t)all_cands.size();
    if (cand_out) for (size_t i = 0; i < all_cands.size() && (int64_t)i < cand_cap; i++) cand_out[i] = all_cands[i];
    if (n_map) *n_map = (int64_t)maps.size();
    if (map_out) for (size_t i = 0; i < maps.size() && (int64_t)i < map_cap; i++) map_out[i] = maps[i];
    if (stats) {
        stats->fragments = total_frags; stats->seeds = nseeds; stats->candidates = all_cands.size();
        stats->scanned = 0; stats->mappings = maps.size(); stats->sketch_sum = ssum;
    }
    return (int64_t)out.size();
}

} // extern "C"
