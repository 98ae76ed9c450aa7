// fa_api.cu -- the extern "C" surface declared in include/fastani_b200.h.
#include <cstdarg>
#include <cstring>
#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include <cstdlib>
#include "fa_internal.cuh"

namespace fa {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

namespace {

__global__ void unpack_kernel(const RefMini *ref, uint64_t first, uint64_t n, uint32_t *hash, int32_t *seq, int32_t *wpos)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RefMini e = ref[first + i];
    hash[i] = e.x; wpos[i] = (int32_t)e.y; seq[i] = (int32_t)e.z;
}

__global__ void pack_kernel(RefMini *ref, uint64_t n, const uint32_t *hash, const int32_t *seq, const int32_t *wpos)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ref[i] = make_uint4(hash[i], (uint32_t)wpos[i], (uint32_t)seq[i], 0u);
}

__global__ void lookup_one_kernel(const uint32_t *ukeys, const uint32_t *uoff, uint32_t n_unique, uint32_t h, uint32_t *out)
{
    uint32_t lo = 0, hi = n_unique;
    while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (ukeys[mid] < h) lo = mid + 1; else hi = mid; }
    if (lo < n_unique && ukeys[lo] == h) { out[0] = uoff[lo]; out[1] = uoff[lo + 1] - uoff[lo]; }
    else { out[0] = 0; out[1] = 0; }
}

__global__ void has_key_kernel(const uint32_t *ukeys, uint32_t n_unique, uint32_t h, uint32_t *out)
{
    uint32_t lo = 0, hi = n_unique;
    while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (ukeys[mid] < h) lo = mid + 1; else hi = mid; }
    out[0] = (lo < n_unique && ukeys[lo] == h) ? 1u : 0u;
}

__global__ void gather_kernel(const RefMini *ref, const uint32_t *pos_idx, uint32_t start, uint32_t n, int32_t *seq, int32_t *wpos)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RefMini e = ref[pos_idx[start + i]];
    seq[i] = (int32_t)e.z; wpos[i] = (int32_t)e.y;
}

int check_params(const fa_params *p)
{
    if (!p) { set_error("params is NULL"); return FA_ERR_INVALID; }
    if (p->k <= 0 || p->k > 2048) { set_error("k must be in [1, 2048], got %d", p->k); return FA_ERR_INVALID; }
    if (p->frag_len <= 0) { set_error("fragment_length must be strictly positive, got %d", p->frag_len); return FA_ERR_INVALID; }
    if (p->min_fraction < 0 || p->min_fraction > 1) { set_error("minimum_fraction must be between 0 and 1"); return FA_ERR_INVALID; }
    if (!(p->p_value > 0)) { set_error("p_value must be positive"); return FA_ERR_INVALID; }
    if (p->pct_identity < 0 || p->pct_identity > 100) { set_error("percentage_identity must be between 0 and 100"); return FA_ERR_INVALID; }
    // 4 = nucleotides; anything else is the reference's protein mode (pyx:548-550, 650-668): forward strand only
    if (p->alphabet < 2) { set_error("alphabet size must be at least 2, got %d", p->alphabet); return FA_ERR_INVALID; }
    if (p->window < 0) { set_error("window must be >= 0"); return FA_ERR_INVALID; }
    return FA_OK;
}

int copy_minimizers(int device, cudaStream_t st, const RefMini *ref, uint64_t total, uint64_t first, uint64_t n,
                    uint32_t *hash, int32_t *seq, int32_t *wpos)
{
    if (first > total || n > total - first) { set_error("minimizer range out of bounds"); return FA_ERR_INVALID; }
    if (n == 0) return FA_OK;
    FA_CUDA(cudaSetDevice(device));
    TmpBuf<uint32_t> dh; TmpBuf<int32_t> ds, dw;
    int rc = dh.reserve(n);
    if (rc == FA_OK) rc = ds.reserve(n);
    if (rc == FA_OK) rc = dw.reserve(n);
    if (rc == FA_OK) {
        unpack_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(ref, first, n, dh.p, ds.p, dw.p);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(hash, dh.p, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(seq, ds.p, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(wpos, dw.p, n * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("copy_minimizers: %s", cudaGetErrorString(e)); rc = FA_ERR_CUDA; }
    }
    return rc;
}

// Sketch one batch of contigs (each consumes a sequence id, pyx:683) and append the minimizers.
// With `contigs_per_genome` the batch holds n_genomes whole genomes (pyx:686-690 after each: length and id bookkeeping).
int sketch_add_batch(fa_sketch *s, const fa_contig *contigs, int32_t n_contigs, int32_t *n_short, int64_t *n_added,
                     const int32_t *contigs_per_genome = nullptr, int32_t n_genomes = 0, uint64_t *genome_len_out = nullptr)
{
    FA_CUDA(cudaSetDevice(s->device));
    const fa_params &P = s->prm;
    std::vector<Upload> ups;
    s->h_seqs.clear();
    uint64_t off = 0, worst = 0;
    int64_t tiles = 0;
    int32_t shorts = 0;
    // ids and lengths consumed by this batch: committed to the sketch only when the GPU work has succeeded, so a failed
    // call (out of memory, CUDA error) leaves the sketch as it was
    uint64_t counter = s->counter, cur_len = s->cur_len;
    for (int32_t c = 0; c < n_contigs; c++) {
        const fa_contig &ct = contigs[c];
        FA_TRY(check_contig(ct, c));
        if (ct.len > 0x7FFFFF00ll) { set_error("contigs longer than 2^31 bases are not supported"); return FA_ERR_UNSUPPORTED; }
    }
    uint64_t bases = 0;
    std::vector<uint64_t> g_len;            // per genome of the batch: its length and the ids consumed when it ends
    std::vector<int32_t> g_end;
    int32_t g = 0, g_left = n_genomes ? contigs_per_genome[0] : 0;
    auto close_genomes = [&]() {            // (genomes without contigs close at once)
        while (g < n_genomes && g_left == 0) {
            g_len.push_back(cur_len); g_end.push_back((int32_t)counter); cur_len = 0;
            g++;
            g_left = g < n_genomes ? contigs_per_genome[g] : 0;
        }
    };
    close_genomes();
    for (int32_t c = 0; c < n_contigs; c++) {
        const fa_contig &ct = contigs[c];
        const int32_t id = (int32_t)counter;
        bases += (uint64_t)ct.len;
        if (ct.len >= P.window && ct.len >= P.k) {                                    // pyx:648
            const int nk = (int)ct.len - P.k + 1;
            SeqDesc d;
            d.off = off; d.len = (int32_t)ct.len; d.id = id; d.raw = contig_prenormalised(ct); d.tile0 = (int32_t)tiles;
            s->h_seqs.push_back(d);
            ups.push_back(Upload{ct.data, ct.unit_bytes, ct.on_device, ct.len, off});
            off += ((uint64_t)ct.len + 15) & ~15ull;
            tiles += (nk + SK_TILE - 1) / SK_TILE;
            if (nk >= P.window) worst += (uint64_t)(nk - P.window + 1);
        } else shorts++;                                                              // pyx:670-677
        cur_len += (uint64_t)(ct.len / P.frag_len) * (uint64_t)P.frag_len;            // pyx:680
        counter++;                                                                    // pyx:683
        if (n_genomes) { g_left--; close_genomes(); }
    }
    auto commit = [&]() {
        s->counter = counter; s->cur_len = cur_len;
        for (size_t i = 0; i < g_len.size(); i++) {
            s->genome_len.push_back(g_len[i]); s->seqs_by_genome.push_back(g_end[i]);
            if (genome_len_out) genome_len_out[i] = g_len[i];
        }
    };
    if (n_short) *n_short = shorts;
    if (n_added) *n_added = shorts == n_contigs && n_contigs == 1 ? -1 : 0;
    const int n_seqs = (int)s->h_seqs.size();
    if (n_seqs == 0) { commit(); return FA_OK; }
    if (tiles > 0x7FFFFF00ll) { set_error("batch too large"); return FA_ERR_UNSUPPORTED; }
    cudaStream_t st = s->st;
    int launches = 0;
    if (!s->ev_ready) {
        FA_CUDA(cudaEventCreate(&s->ev[0])); FA_CUDA(cudaEventCreate(&s->ev[1]));
        s->ev_ready = true;
    }
    FA_CUDA(cudaEventRecord(s->ev[0], st));
    NvtxStages nv;
    nv.next("fa:sketch add (stage + sketch kernels)");
    FA_TRY(stage_sequences(st, s->sc.bytes, s->stage, ups, off, nullptr));
    FA_TRY(s->sc.seqs.reserve(n_seqs)); FA_TRY(s->sc.tile_status.reserve((size_t)tiles));
    FA_TRY(s->sc.counters.reserve(4)); FA_TRY(s->sc.seq_first.reserve(n_seqs)); FA_TRY(s->sc.drops.reserve(n_seqs));
    FA_TRY(s->ref.reserve(s->n + worst + 1, true, st));
    FA_CUDA(cudaMemcpyAsync(s->sc.seqs.p, s->h_seqs.data(), (size_t)n_seqs * sizeof(SeqDesc), cudaMemcpyHostToDevice, st));
    FA_TRY(launch_sketch(st, s->sc, n_seqs, (int)tiles, P.k, P.window, P.alphabet != 4, s->ref.p, nullptr, s->n, &launches));
    FA_TRY(launch_quirk_find(st, s->sc, n_seqs, s->ref.p + s->n, &launches));
    unsigned long long h_ct[4];
    FA_CUDA(cudaMemcpyAsync(h_ct, s->sc.counters.p, sizeof h_ct, cudaMemcpyDeviceToHost, st));
    FA_CUDA(cudaStreamSynchronize(st));
    uint64_t added = h_ct[1];
    if (h_ct[2] > 0) FA_TRY(quirk_compact(st, s->sc, (unsigned int)h_ct[2], s->ref.p + s->n, added, &added, &launches));
    FA_CUDA(cudaEventRecord(s->ev[1], st));
    FA_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0;
    cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
    s->ms_sketch += ms; s->bases += bases;
    s->n += added;
    commit();
    if (n_added) *n_added = (int64_t)added;
    return FA_OK;
}

void free_sketch_scratch(SketchScratch &sc)
{
    sc.bytes.release(); sc.seqs.release(); sc.tile_status.release(); sc.counters.release(); sc.seq_first.release(); sc.drops.release();
}

}  // namespace
}  // namespace fa

using namespace fa;

extern "C" {

const char *fa_last_error(void) { return g_err; }
int fa_version(void) { return 100; }

int fa_device_count(int32_t *n_out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); if (n_out) *n_out = 0; return FA_ERR_CUDA; }
    if (n_out) *n_out = n;
    return FA_OK;
}

int fa_recommended_window(const fa_params *p, int32_t *w_out)
{
    FA_TRY(check_params(p));
    *w_out = recommended_window(p->p_value, p->k, p->alphabet, p->pct_identity, p->frag_len, p->ref_size);
    return FA_OK;
}

int fa_stat_minimum_hits(int32_t s, int32_t k, float pid, int32_t *out) { *out = minimum_hits_relaxed(s, k, pid); return FA_OK; }

int fa_stat_l2(int32_t shared, int32_t s, int32_t k, float pid, float *identity, int32_t *pass)
{
    float id = 0;
    bool ok = l2_pass(shared, s, k, pid, &id);
    if (identity) *identity = id;
    if (pass) *pass = ok ? 1 : 0;
    return FA_OK;
}

int fa_stat_table_row(int32_t s, int32_t s_max, int32_t k, float pid, int32_t *min_hits, int32_t *min_shared, int32_t *irregular)
{
    if (s < 1 || s > s_max || s_max > 4096 || k < 1) { set_error("bad arguments"); return FA_ERR_INVALID; }
    const StatTable &t = stat_table(k, pid, s_max);
    if (min_hits) *min_hits = t.min_hits[s];
    if (min_shared) *min_shared = t.min_shared[s];
    if (irregular) *irregular = t.irregular + t.irregular_l2;
    return FA_OK;
}

int fa_sketch_create(const fa_params *p, int32_t device, fa_sketch **out)
{
    FA_TRY(check_params(p));
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { set_error("no CUDA device is available (%s); libfastani_b200 has no CPU path", cudaGetErrorString(e)); return FA_ERR_CUDA; }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(device));
    fa_sketch *s = new (std::nothrow) fa_sketch();
    if (!s) return FA_ERR_NOMEM;
    s->prm = *p;
    if (s->prm.window == 0)      // protein mode (alphabet != 4) maps with window 1, pyx:548-550
        s->prm.window = p->alphabet == 4 ? recommended_window(p->p_value, p->k, p->alphabet, p->pct_identity, p->frag_len, p->ref_size) : 1;
    s->device = device;
    e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); delete s; return FA_ERR_CUDA; }
    *out = s;
    return FA_OK;
}

void fa_sketch_free(fa_sketch *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    s->ref.release(); free_sketch_scratch(s->sc); s->stage.release();
    if (s->ev_ready) { cudaEventDestroy(s->ev[0]); cudaEventDestroy(s->ev[1]); }
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

int fa_sketch_add_contig(fa_sketch *s, const void *data, int32_t unit_bytes, int64_t len, int64_t *n_added)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    fa_contig c{data, unit_bytes, 0, len};
    int32_t shorts = 0;
    int64_t added = 0;
    FA_TRY(sketch_add_batch(s, &c, 1, &shorts, &added));
    if (n_added) *n_added = shorts ? -1 : added;
    return FA_OK;
}

int fa_sketch_end_genome(fa_sketch *s, uint64_t *genome_len_out)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    s->genome_len.push_back(s->cur_len);                                  // pyx:687
    s->seqs_by_genome.push_back((int32_t)s->counter);                 // pyx:690
    if (genome_len_out) *genome_len_out = s->cur_len;
    s->cur_len = 0;
    return FA_OK;
}

int fa_sketch_add_genome(fa_sketch *s, const fa_contig *contigs, int32_t n_contigs, uint64_t *genome_len_out, int32_t *n_short)
{
    if (!s || n_contigs < 0 || (n_contigs > 0 && !contigs)) { set_error("bad arguments"); return FA_ERR_INVALID; }
    FA_TRY(sketch_add_batch(s, contigs, n_contigs, n_short, nullptr));
    return fa_sketch_end_genome(s, genome_len_out);
}

int fa_sketch_add_genomes(fa_sketch *s, const fa_contig *contigs, const int32_t *contigs_per_genome, int32_t n_genomes,
                          uint64_t *genome_len_out, int32_t *n_short)
{
    if (!s || n_genomes < 0 || (n_genomes > 0 && !contigs_per_genome)) { set_error("bad arguments"); return FA_ERR_INVALID; }
    if (n_short) *n_short = 0;
    // whole genomes per launch sequence, about 256 MB of bases at a time: one H2D stream, one sketch launch, one sync
    constexpr int64_t BATCH_BYTES = 256ll << 20;
    int64_t c0 = 0;
    int32_t g0 = 0;
    while (g0 < n_genomes) {
        int32_t g1 = g0;
        int64_t c1 = c0, bytes = 0;
        while (g1 < n_genomes) {
            if (contigs_per_genome[g1] < 0) { set_error("genome %d: negative contig count", g1); return FA_ERR_INVALID; }
            int64_t gb = 0;
            for (int64_t c = c1; c < c1 + contigs_per_genome[g1]; c++) gb += contigs[c].len > 0 ? contigs[c].len : 0;
            if (g1 > g0 && (bytes + gb > BATCH_BYTES || c1 + contigs_per_genome[g1] - c0 > 0x3FFFFFFF)) break;
            bytes += gb; c1 += contigs_per_genome[g1]; g1++;
        }
        int32_t shorts = 0;
        FA_TRY(sketch_add_batch(s, contigs ? contigs + c0 : nullptr, (int32_t)(c1 - c0), &shorts, nullptr, contigs_per_genome + g0, g1 - g0,
                                genome_len_out ? genome_len_out + g0 : nullptr));
        if (n_short) *n_short += shorts;
        c0 = c1; g0 = g1;
    }
    return FA_OK;
}

int fa_sketch_build_stats(const fa_sketch *s, double *ms_sketch, uint64_t *bases)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    if (ms_sketch) *ms_sketch = s->ms_sketch;
    if (bases) *bases = s->bases;
    return FA_OK;
}

int fa_index_build_stats(const fa_index *ix, float *ms_build, float *ms_sort)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    if (ms_build) *ms_build = ix->ms_build;
    if (ms_sort) *ms_sort = ix->ms_sort;
    return FA_OK;
}

int fa_sketch_clear(fa_sketch *s)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    s->n = 0; s->cur_len = 0;
    s->seqs_by_genome.clear(); s->genome_len.clear(); s->counter = 0;
    return FA_OK;
}

int fa_sketch_counts(const fa_sketch *s, uint64_t *n_minimizers, uint64_t *n_contigs, uint64_t *n_genomes)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    if (n_minimizers) *n_minimizers = s->n;
    if (n_contigs) *n_contigs = s->counter;
    if (n_genomes) *n_genomes = s->genome_len.size();
    return FA_OK;
}

int fa_sketch_copy_minimizers(const fa_sketch *s, uint64_t first, uint64_t n, uint32_t *hash, int32_t *seq, int32_t *wpos)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    return copy_minimizers(s->device, s->st, s->ref.p, s->n, first, n, hash, seq, wpos);
}

int fa_sketch_copy_meta(const fa_sketch *s, int32_t *seqs_by_genome, uint64_t *genome_len)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    if (seqs_by_genome) std::copy(s->seqs_by_genome.begin(), s->seqs_by_genome.end(), seqs_by_genome);
    if (genome_len) std::copy(s->genome_len.begin(), s->genome_len.end(), genome_len);
    return FA_OK;
}

int fa_sketch_restore(fa_sketch *s, const uint32_t *hash, const int32_t *seq, const int32_t *wpos, uint64_t n,
                      const int32_t *seqs_by_genome, const uint64_t *genome_len, uint64_t n_genomes,
                      uint64_t n_contigs)
{
    if (!s) { set_error("sketch is NULL"); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(s->device));
    fa_sketch_clear(s);
    s->seqs_by_genome.assign(seqs_by_genome, seqs_by_genome + n_genomes);
    s->genome_len.assign(genome_len, genome_len + n_genomes);
    s->counter = n_contigs;
    if (n == 0) return FA_OK;
    FA_TRY(s->ref.reserve(n + 1));
    TmpBuf<uint32_t> dh; TmpBuf<int32_t> ds, dw;
    FA_TRY(dh.reserve(n)); FA_TRY(ds.reserve(n)); FA_TRY(dw.reserve(n));
    FA_CUDA(cudaMemcpyAsync(dh.p, hash, n * 4, cudaMemcpyHostToDevice, s->st));
    FA_CUDA(cudaMemcpyAsync(ds.p, seq, n * 4, cudaMemcpyHostToDevice, s->st));
    FA_CUDA(cudaMemcpyAsync(dw.p, wpos, n * 4, cudaMemcpyHostToDevice, s->st));
    pack_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, s->st>>>(s->ref.p, n, dh.p, ds.p, dw.p);
    FA_CUDA(cudaGetLastError());
    FA_CUDA(cudaStreamSynchronize(s->st));
    s->n = n;
    return FA_OK;
}

int fa_sketch_index(fa_sketch *s, fa_index **out)
{
    if (!s || !out) { set_error("bad arguments"); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(s->device));
    if (s->prm.frag_len > 32767) { set_error("fragment_length > 32767 is not supported on the device path"); return FA_ERR_UNSUPPORTED; }
    if (s->prm.frag_len <= 20) { set_error("fragment_length <= 20 is not supported (the reference divides by fragment_length - 20)"); return FA_ERR_UNSUPPORTED; }
    fa_index *ix = new (std::nothrow) fa_index();
    if (ix) { const char *e = getenv("FA_L1_PARTS"); if (e && *e) ix->l1_parts = atoi(e); }   // (experiments: see fa_debug_set_l1_parts)
    if (!ix) return FA_ERR_NOMEM;
    ix->prm = s->prm; ix->device = s->device;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    if (e != cudaSuccess) { set_error("fa_sketch_index: %s", cudaGetErrorString(e)); fa_index_free(ix); return FA_ERR_CUDA; }
    // The index is built on the sketch's minimizers in place; ownership moves to it only when the build has succeeded
    // (pyx:795-804: the Mapper takes the Sketch_t, the Sketch gets a fresh one).  A failed build -- out of memory in the
    // sort buffers, more than 2^32 minimizers, a CUDA error -- leaves the sketch as it was.
    ix->ref = s->ref; ix->n = s->n;
    ix->seqs_by_genome = s->seqs_by_genome;
    ix->genome_len = s->genome_len;
    ix->n_contigs = s->counter;
    int rc = FA_OK;
    if (!ix->ref.p) { rc = ix->ref.reserve(1); if (rc == FA_OK) s->ref = ix->ref; }
    int launches = 0;
    if (rc == FA_OK) rc = build_index(ix, &launches);
    if (rc != FA_OK) {
        ix->ref = DevBuf<RefMini>();          // still the sketch's
        fa_index_free(ix);
        return rc;
    }
    s->ref = DevBuf<RefMini>();
    fa_sketch_clear(s);
    *out = ix;
    return FA_OK;
}

void fa_index_free(fa_index *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->device);
    ix->ref.release(); ix->hw.release(); ix->hl.release(); ix->fb.release(); ix->gpos.release(); ix->irr.release(); ix->pos_idx.release(); ix->ukeys.release(); ix->uoff.release(); ix->dir.release();
    ix->contig_off.release(); ix->genome_of_seq.release(); ix->bin_base.release(); ix->genome_cell.release();
    ix->pre[0].release(); ix->pre[1].release();
    ix->d_min_hits.release(); ix->d_min_shared.release(); ix->d_id_off.release(); ix->d_identity.release();
    ix->ws.release(); ix->xs.release();
    if (ix->st) cudaStreamDestroy(ix->st);
    delete ix;
}

int fa_index_counts(const fa_index *ix, uint64_t *n_minimizers, uint64_t *n_unique, uint64_t *n_contigs, uint64_t *n_genomes)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    if (n_minimizers) *n_minimizers = ix->n;
    if (n_unique) *n_unique = ix->n_unique;
    if (n_contigs) *n_contigs = ix->n_contigs;
    if (n_genomes) *n_genomes = ix->genome_len.size();
    return FA_OK;
}

int fa_index_params(const fa_index *ix, fa_params *out)
{
    if (!ix || !out) { set_error("bad arguments"); return FA_ERR_INVALID; }
    *out = ix->prm;
    return FA_OK;
}

int fa_index_copy_minimizers(const fa_index *ix, uint64_t first, uint64_t n, uint32_t *hash, int32_t *seq, int32_t *wpos)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    return copy_minimizers(ix->device, ix->st, ix->ref.p, ix->n, first, n, hash, seq, wpos);
}

int fa_index_copy_meta(const fa_index *ix, int32_t *seqs_by_genome, uint64_t *genome_len)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    if (seqs_by_genome) std::copy(ix->seqs_by_genome.begin(), ix->seqs_by_genome.end(), seqs_by_genome);
    if (genome_len) std::copy(ix->genome_len.begin(), ix->genome_len.end(), genome_len);
    return FA_OK;
}

int fa_index_copy_keys(const fa_index *ix, uint64_t first, uint64_t n, uint32_t *keys)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    if (first > ix->n_unique || n > ix->n_unique - first) { set_error("key range out of bounds"); return FA_ERR_INVALID; }
    if (!n) return FA_OK;
    FA_CUDA(cudaSetDevice(ix->device));
    FA_CUDA(cudaMemcpy(keys, ix->ukeys.p + first, n * 4, cudaMemcpyDeviceToHost));
    return FA_OK;
}

int fa_index_lookup(const fa_index *ix, uint32_t hash, int32_t *seq, int32_t *wpos, uint64_t cap, uint64_t *n)
{
    if (!ix || !n) { set_error("bad arguments"); return FA_ERR_INVALID; }
    *n = 0;
    if (ix->n_unique == 0) return FA_OK;
    FA_CUDA(cudaSetDevice(ix->device));
    uint32_t h_out[2] = {0, 0};
    TmpBuf<uint32_t> d_out;
    FA_TRY(d_out.reserve(2));
    lookup_one_kernel<<<1, 1, 0, ix->st>>>(ix->ukeys.p, ix->uoff.p, (uint32_t)ix->n_unique, hash, d_out.p);
    FA_CUDA(cudaMemcpyAsync(h_out, d_out.p, 8, cudaMemcpyDeviceToHost, ix->st));
    FA_CUDA(cudaStreamSynchronize(ix->st));
    *n = h_out[1];
    uint32_t m = (uint32_t)std::min<uint64_t>(cap, h_out[1]);
    if (m && seq && wpos) {
        TmpBuf<int32_t> ds, dw;
        FA_TRY(ds.reserve(m)); FA_TRY(dw.reserve(m));
        gather_kernel<<<(m + 127) / 128, 128, 0, ix->st>>>(ix->ref.p, ix->pos_idx.p, h_out[0], m, ds.p, dw.p);
        FA_CUDA(cudaMemcpyAsync(seq, ds.p, m * 4, cudaMemcpyDeviceToHost, ix->st));
        FA_CUDA(cudaMemcpyAsync(wpos, dw.p, m * 4, cudaMemcpyDeviceToHost, ix->st));
        FA_CUDA(cudaStreamSynchronize(ix->st));
    }
    return FA_OK;
}

int fa_index_has_key(const fa_index *ix, uint32_t hash, int32_t *found)
{
    if (!ix || !found) { set_error("bad arguments"); return FA_ERR_INVALID; }
    *found = 0;
    if (ix->n_unique == 0) return FA_OK;
    FA_CUDA(cudaSetDevice(ix->device));
    uint32_t h_out[2] = {0, 0};
    TmpBuf<uint32_t> d_out;
    FA_TRY(d_out.reserve(2));
    // (lookup_one_kernel reports entries; a key with an empty list -- possible after fa_index_set_lookup -- has none)
    has_key_kernel<<<1, 1, 0, ix->st>>>(ix->ukeys.p, (uint32_t)ix->n_unique, hash, d_out.p);
    FA_CUDA(cudaMemcpyAsync(h_out, d_out.p, 4, cudaMemcpyDeviceToHost, ix->st));
    FA_CUDA(cudaStreamSynchronize(ix->st));
    *found = (int32_t)h_out[0];
    return FA_OK;
}

int fa_index_set_lookup(fa_index *ix, uint32_t hash, const int32_t *seq, const int32_t *wpos, uint64_t n)
{
    if (!ix || (n && (!seq || !wpos))) { set_error("bad arguments"); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(ix->device));
    std::lock_guard<std::mutex> guard(ix->mtx);
    return edit_lookup(ix, hash, seq, wpos, n, false, nullptr);
}

int fa_index_del_lookup(fa_index *ix, uint32_t hash, int32_t *found)
{
    if (!ix || !found) { set_error("bad arguments"); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(ix->device));
    std::lock_guard<std::mutex> guard(ix->mtx);
    int missing = 0;
    const int rc = edit_lookup(ix, hash, nullptr, nullptr, 0, true, &missing);
    *found = missing ? 0 : 1;
    return rc;
}

int fa_index_occurrence_threshold(const fa_index *ix, int32_t *out)
{
    (void)ix;
    *out = 2147483647;      // freqThreshold stays INT_MAX: percentageThreshold == 0 (winSketch.hpp:52, 208-231)
    return FA_OK;
}

static void add_info(fa_query_info &sum, const fa_query_info &qi)
{
    sum.fragments += qi.fragments; sum.sketch_sum += qi.sketch_sum; sum.seeds += qi.seeds; sum.candidates += qi.candidates;
    sum.scanned += qi.scanned; sum.mappings += qi.mappings; sum.short_contigs += qi.short_contigs;
    sum.kernel_launches += qi.kernel_launches; sum.h2d_bytes += qi.h2d_bytes; sum.d2h_bytes += qi.d2h_bytes;
    sum.l2_fallback += qi.l2_fallback; sum.events += qi.events; sum.events_replayed += qi.events_replayed;
    sum.l1_sorted_fragments += qi.l1_sorted_fragments; sum.l1_small_fragments += qi.l1_small_fragments;
    sum.ms_h2d += qi.ms_h2d; sum.ms_sketch += qi.ms_sketch; sum.ms_lookup += qi.ms_lookup; sum.ms_seed_sort += qi.ms_seed_sort;
    sum.ms_l1 += qi.ms_l1; sum.ms_l2 += qi.ms_l2; sum.ms_cgi += qi.ms_cgi; sum.ms_d2h += qi.ms_d2h; sum.ms_total += qi.ms_total;
    sum.ms_l2_prep += qi.ms_l2_prep; sum.ms_l2_events += qi.ms_l2_events; sum.ms_l2_slide += qi.ms_l2_slide;
    sum.ms_batch += qi.ms_batch;
    sum.l1_parts = std::max(sum.l1_parts, qi.l1_parts); sum.l1_tiny_fragments += qi.l1_tiny_fragments; sum.ms_exchange += qi.ms_exchange;
}

static int query_checked(fa_index *ix, const fa_contig *contigs, int32_t n_contigs, fa_hit *out, uint64_t cap, uint64_t *n_out,
                         fa_query_info *info, Prefetch *pf)
{
    if (!ix || n_contigs < 0 || (n_contigs > 0 && !contigs) || !n_out) { set_error("bad arguments"); return FA_ERR_INVALID; }
    for (int32_t c = 0; c < n_contigs; c++)
        if (contigs[c].len < 0 || (contigs[c].len > 0 && !contigs[c].data)) { set_error("contig %d: bad buffer", c); return FA_ERR_INVALID; }
    return run_query(ix, contigs, n_contigs, out, cap, n_out, info, pf);
}

int fa_query(fa_index *ix, const fa_contig *contigs, int32_t n_contigs, fa_hit *out, uint64_t cap, uint64_t *n_out,
             fa_query_info *info)
{
    return query_checked(ix, contigs, n_contigs, out, cap, n_out, info, nullptr);
}

// Many queries in one call (no host language between them): query q owns contigs
// [sum(contigs_per_query[:q]), +contigs_per_query[q]); its hits are out[hit_offsets[q] .. hit_offsets[q + 1]).
// Counters and stage times of `info` are summed over the passes.
//  * Light queries share a pass of the pipeline (run_queries, fa_map.cu): a many-to-many query has a few thousand
//    candidates, far too few to fill the GPU's lanes in the L2 slide, and its twenty kernel launches and four host
//    syncs cost as much as its kernels.  The first pass takes one query; afterwards the seeds and events per fragment
//    seen so far size the next pass (at most 64 queries / 96 k fragments / the budgets below), so a heavy query --
//    config 2: 86 M seeds, 1.8 G events -- still runs alone.
//  * While a pass is mapped, a helper thread stages the bytes of the next one (its own pinned buffer, copy stream and
//    device buffer, fa_map.cu prefetch_query), so host copy and H2D overlap the kernels.
int fa_query_batch(fa_index *ix, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries,
                   fa_hit *out, uint64_t cap, uint64_t *hit_offsets, fa_query_info *info)
{
    return query_batch_impl(ix, nullptr, contigs, contigs_per_query, n_queries, out, cap, hit_offsets, info);
}

}  // extern "C"

// `comm` (reference-sharded layout, fa_query_batch_sharded): the queries are cut into groups by their sizes alone -- the
// same groups on every rank -- and the sketches of a group are made once across the ranks (sketch_exchange); the passes,
// which every rank sizes from what ITS shard made of the queries so far, stay inside a group and read the gathered
// sketches.  No collective depends on anything a rank measured.
int fa::query_batch_impl(fa_index *ix, fa_comm *comm, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries,
                         fa_hit *out, uint64_t cap, uint64_t *hit_offsets, fa_query_info *info)
{
    if (!ix || n_queries < 0 || !hit_offsets || (n_queries > 0 && !contigs_per_query)) { set_error("bad arguments"); return FA_ERR_INVALID; }
    fa_query_info sum;
    memset(&sum, 0, sizeof sum);
    hit_offsets[0] = 0;
    int64_t n_contigs = 0;
    for (int32_t q = 0; q < n_queries; q++) {
        if (contigs_per_query[q] < 0) { set_error("query %d: negative contig count", q); return FA_ERR_INVALID; }
        n_contigs += contigs_per_query[q];
    }
    if (n_contigs > 0 && !contigs) { set_error("bad arguments"); return FA_ERR_INVALID; }
    for (int64_t c = 0; c < n_contigs; c++)
        if (contigs[c].len < 0 || (contigs[c].len > 0 && !contigs[c].data)) { set_error("contig %lld: bad buffer", (long long)c); return FA_ERR_INVALID; }
    const int L = std::max(ix->prm.frag_len, 1);
    std::vector<int64_t> first(n_queries + 1, 0);
    std::vector<uint64_t> frags(n_queries, 0);
    for (int32_t q = 0; q < n_queries; q++) {
        first[q + 1] = first[q] + contigs_per_query[q];
        for (int64_t c = first[q]; c < first[q + 1]; c++) frags[q] += (uint64_t)(contigs[c].len / L);
    }
    // one event pair around the whole call on the library's stream: the device-side time of the batch, gaps between
    // the passes included (the per-stage timers of `info` only cover the passes themselves)
    struct Pair {
        cudaEvent_t a = nullptr, b = nullptr;
        ~Pair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } outer;
    FA_CUDA(cudaSetDevice(ix->device));
    FA_CUDA(cudaEventCreate(&outer.a)); FA_CUDA(cudaEventCreate(&outer.b));
    FA_CUDA(cudaEventRecord(outer.a, ix->st));
    const bool exchange = comm && comm_world(comm) > 1 && exchange_stride(ix->prm) > 0 && !getenv("FA_NO_SKETCH_EXCHANGE");
    std::unique_lock<std::mutex> pre_lock(ix->pre_mtx, std::try_to_lock);      // a second concurrent batch maps without staging ahead
    const bool ahead = pre_lock.owns_lock() && n_queries > 1 && !exchange;
    if (pre_lock.owns_lock()) ix->pre[0].valid = ix->pre[1].valid = false;       // nothing staged by an earlier call is ours

    constexpr uint64_t PASS_QUERIES = 64, PASS_FRAGS = 96 * 1024, PASS_SEEDS = 384ull << 20, PASS_EVENTS = 1024ull << 20;
    double seeds_per_frag = -1.0, events_per_frag = -1.0;      // largest seen so far in this call (-1: nothing seen)
    int32_t g0 = 0, g1 = n_queries;                              // the group of queries whose sketches are at hand (exchange)
    auto group_end = [&](int32_t q0) {
        int32_t q1 = q0 + 1;
        uint64_t f = frags[q0];
        while (q1 < n_queries && (uint64_t)(q1 - q0) < PASS_QUERIES && f + frags[q1] <= PASS_FRAGS) f += frags[q1++];
        return q1;
    };
    auto pass_end = [&](int32_t q0) {
        int32_t q1 = q0 + 1;
        if (seeds_per_frag < 0) return q1;
        uint64_t f = frags[q0];
        while (q1 < g1 && (uint64_t)(q1 - q0) < PASS_QUERIES) {
            const uint64_t nf = f + frags[q1];
            if (nf > PASS_FRAGS || (double)nf * seeds_per_frag * 1.5 > (double)PASS_SEEDS ||
                (double)nf * events_per_frag * 1.5 > (double)PASS_EVENTS) break;
            f = nf; q1++;
        }
        return q1;
    };

    uint64_t used = 0;
    std::vector<uint64_t> offs(PASS_QUERIES + 1);
    int slot = 0;
    // exchange: the groups are fixed up front; the sketches of group g + 1 are made and gathered (on the exchange stream)
    // while the passes of group g run
    std::vector<int32_t> g_end;
    if (exchange) for (int32_t q = 0; q < n_queries;) { q = group_end(q); g_end.push_back(q); }
    PreSketch shares[2], group;
    float ms_exchange = 0;
    size_t gi = 0;
    auto exchange_time = [&]() { float ms = 0; if (cudaEventSynchronize(ix->xs.t1) == cudaSuccess && cudaEventElapsedTime(&ms, ix->xs.t0, ix->xs.t1) == cudaSuccess) ms_exchange += ms; };
    if (exchange) {
        g1 = 0;                                                   // (the first group begins inside the loop)
        if (!g_end.empty()) FA_TRY(sketch_exchange(ix, comm, contigs, (int32_t)first[g_end[0]], 0, &shares[0], &sum));
    }
    int32_t q0 = 0, q1 = n_queries && !exchange ? pass_end(0) : 0;
    while (q0 < n_queries) {
        if (exchange && q0 == g1) {
            // the next group: its shares have been on their way since the previous group began
            g0 = g1; g1 = g_end[gi];
            FA_CUDA(cudaEventSynchronize(ix->xs.done[gi & 1]));
            exchange_time();
            group = shares[gi & 1];
            if (gi + 1 < g_end.size())
                FA_TRY(sketch_exchange(ix, comm, contigs + first[g1], (int32_t)(first[g_end[gi + 1]] - first[g1]), (int)((gi + 1) & 1), &shares[(gi + 1) & 1], &sum));
            gi++;
            q1 = pass_end(q0);
        }
        const int32_t nq = q1 - q0;
        fa_query_info qi;
        if (exchange) { group.first_frag = 0; for (int32_t q = g0; q < q0; q++) group.first_frag += frags[q]; }
        // plan the pass after this one with what is known now, and stage it while this one runs
        const int32_t q2 = q1 < n_queries ? pass_end(q1) : q1;
        std::thread helper;
        if (ahead && q1 < n_queries && first[q2] > first[q1]) {
            const fa_contig *nx = contigs + first[q1];
            const int32_t nx_n = (int32_t)(first[q2] - first[q1]);
            Prefetch *pfs = &ix->pre[slot ^ 1];
            helper = std::thread([ix, pfs, nx, nx_n]() { if (prefetch_query(ix, *pfs, nx, nx_n) != FA_OK) pfs->valid = false; });
        }
        int rc = run_queries(ix, contigs ? contigs + first[q0] : nullptr, contigs_per_query + q0, nq, out ? out + used : nullptr,
                             cap - used, offs.data(), &qi, ahead ? &ix->pre[slot] : nullptr, exchange ? &group : nullptr);
        if (helper.joinable()) helper.join();
        if (rc == FA_RETRY_PLAIN)          // a sketch did not fit its slot of the exchange: this pass sketches its own queries
            rc = run_queries(ix, contigs ? contigs + first[q0] : nullptr, contigs_per_query + q0, nq, out ? out + used : nullptr,
                             cap - used, offs.data(), &qi, nullptr, nullptr);
        if ((rc == FA_ERR_NOMEM || rc == FA_ERR_UNSUPPORTED) && nq > 1) {
            // the pass was sized from lighter queries (or holds more queries than the cell tables of this index allow):
            // map its queries one by one
            cudaGetLastError();
            memset(&qi, 0, sizeof qi);
            uint64_t u2 = 0;
            offs[0] = 0;
            rc = FA_OK;
            for (int32_t q = q0; q < q1 && rc == FA_OK; q++) {
                fa_query_info one;
                uint64_t o2[2];
                const uint64_t room = cap - used > u2 ? cap - used - u2 : 0;
                rc = run_queries(ix, contigs + first[q], contigs_per_query + q, 1, out && room ? out + used + u2 : nullptr, room, o2, &one, nullptr);
                u2 += o2[1];
                offs[q - q0 + 1] = u2;
                add_info(qi, one);
            }
        }
        if (rc != FA_OK) return rc;
        if (offs[nq] > cap - used) { set_error("queries %d..%d: %llu hits do not fit the output (capacity %llu)", q0, q1 - 1, (unsigned long long)(used + offs[nq]), (unsigned long long)cap); return FA_ERR_INVALID; }
        for (int32_t q = 0; q < nq; q++) hit_offsets[q0 + q + 1] = used + offs[q + 1];
        used += offs[nq];
        add_info(sum, qi);
        if (qi.fragments) {
            seeds_per_frag = std::max(seeds_per_frag, (double)qi.seeds / (double)qi.fragments);
            events_per_frag = std::max(events_per_frag, (double)qi.events / (double)qi.fragments);
        }
        // the staged pass was planned before this one's counters were known: keep it as planned
        q0 = q1; q1 = q2; slot ^= 1;
        if (q0 < n_queries && q1 == q0 && !(exchange && q0 == g1)) q1 = pass_end(q0);
    }
    sum.ms_exchange = ms_exchange;                                // (off the critical path except for the first group)
    FA_CUDA(cudaEventRecord(outer.b, ix->st));
    FA_CUDA(cudaEventSynchronize(outer.b));
    sum.ms_batch = 0;
    cudaEventElapsedTime(&sum.ms_batch, outer.a, outer.b);
    if (info) *info = sum;
    return FA_OK;
}

extern "C" {

int fa_debug_last_candidates(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n) { return debug_candidates(ix, rows, cap, n); }
int fa_debug_last_mappings(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n) { return debug_mappings(ix, rows, cap, n); }
int fa_debug_set_l1_seed_cap(fa_index *ix, int64_t cap)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    std::lock_guard<std::mutex> guard(ix->mtx);
    ix->l1_seed_cap = cap < 0 ? -1 : (long long)cap;
    return FA_OK;
}

int fa_debug_set_l1_small_cap(fa_index *ix, int64_t cap)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    std::lock_guard<std::mutex> guard(ix->mtx);
    ix->l1_small_cap = cap < 0 ? -1 : (long long)cap;
    return FA_OK;
}

int fa_debug_set_l1_small_shape(fa_index *ix, int32_t shape)
{
    if (!ix || shape > 2) { set_error("bad arguments"); return FA_ERR_INVALID; }
    std::lock_guard<std::mutex> guard(ix->mtx);
    ix->l1_small_shape = shape < 0 ? -1 : shape;
    return FA_OK;
}

int fa_debug_set_l1_tiny_cap(fa_index *ix, int64_t cap)
{
    if (!ix) { set_error("index is NULL"); return FA_ERR_INVALID; }
    std::lock_guard<std::mutex> guard(ix->mtx);
    ix->l1_tiny_cap = cap < 0 ? -1 : (long long)cap;
    return FA_OK;
}

int fa_debug_set_l1_parts(fa_index *ix, int32_t parts, int64_t part_cap)
{
    if (!ix) { set_error("bad arguments"); return FA_ERR_INVALID; }
    std::lock_guard<std::mutex> guard(ix->mtx);
    ix->l1_parts = parts < 0 ? -1 : parts;      // (0 is the default)
    ix->l1_part_cap = part_cap < 0 ? -1 : part_cap;
    return FA_OK;
}

int fa_device_alloc(int32_t device, uint64_t bytes, void **dptr)
{
    if (!dptr) { set_error("bad arguments"); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(device));
    FA_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return FA_OK;
}
int fa_device_upload(int32_t device, void *dptr, const void *src, uint64_t bytes)
{
    FA_CUDA(cudaSetDevice(device));
    FA_CUDA(cudaMemcpy(dptr, src, bytes, cudaMemcpyHostToDevice));
    return FA_OK;
}
int fa_device_download(int32_t device, void *dst, const void *dptr, uint64_t bytes)
{
    FA_CUDA(cudaSetDevice(device));
    FA_CUDA(cudaMemcpy(dst, dptr, bytes, cudaMemcpyDeviceToHost));
    return FA_OK;
}
int fa_device_free(int32_t device, void *dptr)
{
    FA_CUDA(cudaSetDevice(device));
    FA_CUDA(cudaFree(dptr));
    return FA_OK;
}
int fa_device_mem_info(int32_t device, uint64_t *free_bytes, uint64_t *total_bytes)
{
    size_t f = 0, t = 0;
    FA_CUDA(cudaSetDevice(device));
    FA_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return FA_OK;
}

}  // extern "C"
