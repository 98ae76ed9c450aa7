// fa_map.cu -- the query path: K1 on query fragments, K3 (L1 seeding + candidate regions),
// K4 (L2 sliding-window winnowed-MinHash Jaccard), K5 (core-genome identity).
//
// Replaces Mapper._query_draft / _query_fragment / _do_l1_mappings (src/pyfastani/_fastani.pyx:
// 885-1136), skch::Map::computeL1CandidateRegions / doL2Mapping / computeL2MappedRegions
// (FA/map/include/computeMap.hpp:310-493), skch::SlideMapper (slidingMap.hpp),
// skch::MIIteratorL2 (MIIteratorL2.hpp:54-96) and cgi::computeCGI
// (FA/cgi/include/computeCoreIdentity.hpp:163-295).
//
// The reference maps one fragment at a time on a CPU thread; here every stage runs over all
// fragments of the query at once:
//   sketch      one tile of k-mers per CTA (fa_sketch.cu), then sort+unique per fragment
//   lookup      one CTA per fragment: directory + binary search over the unique hashes
//   seeds       (fragment, reference index) keys, one device-wide radix sort
//   candidates  one CTA per fragment streaming over its sorted seeds
//   L2          one THREAD per candidate: the super-window slide with O(1) state updates
//   CGI         best per (fragment, genome), atomicMax per (contig, bin), ordered float sum
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <thread>

#include <cstdlib>
#include "fa_internal.cuh"

namespace fa {

namespace {

// device counters (Workspace::counters)
enum { CT_MAXS = 0, CT_ERR = 1, CT_WORK = 2, CT_SCANNED = 3, CT_MAPPINGS = 4, CT_SKETCH_SUM = 5, CT_REDO = 6, CT_WORK2 = 7, CT_WORK3 = 8, CT_REPLAYED = 9, CT_N = 10 };
enum { ERR_SORT_CAP = 1, ERR_S_MAX = 2, ERR_EXCH = 4 };

constexpr int L2_THREADS = 64;          // lanes per L2 CTA: each lane slides one candidate at a time
constexpr int L2_ITEM = 128;            // candidates of one fragment per work item (lanes pull them one by one)
constexpr int L2_FB_STATE = 32 * 1024;  // u16 state of the exact fallback kernel
constexpr int L2_TAB_BITS = 12;         // classification table over the top bits of the hash
constexpr int L2_TAB = 1 << L2_TAB_BITS;
constexpr int L2_QPAD = 8;              // sentinel entries behind the staged query sketch
constexpr int32_t L2_REDO = INT32_MIN;  // Mapping.ref_start marker: redo this candidate in the fallback kernel

// state words (four one-byte buckets each) a sketch of size s needs per lane, and the lanes that fit
__host__ __device__ inline int l2_words_for(int s) { return (s + 4) >> 2; }
__host__ __device__ inline int l2_fb_lanes_for(int s)
{
    int t = L2_FB_STATE / ((s + 1) * 2);
    return t > L2_THREADS ? L2_THREADS : (t < 1 ? 1 : t);
}

template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                     // s_warp may still be read from a previous call
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < THREADS / 32; q++) { uint32_t t = s_warp[q]; if (q < wid) base += t; tot += t; }
    *total = tot;
    return base + incl - v;
}

// ---- per-fragment sort + unique of the query minimizer hashes (pyx:929-938) -----------------
__global__ void sort_unique_kernel(uint32_t *qhash, const uint64_t *seq_first, const uint32_t *seq_cnt,
                                   int n_frags, int cap, int s_max, int32_t *qs, unsigned long long *counters)
{
    extern __shared__ uint32_t s_h[];
    __shared__ uint32_t s_warp[8];
    const int f = blockIdx.x, tid = threadIdx.x;
    const uint64_t b = seq_first[f];                     // the fragment's slot (launch_sketch with slot_cap)
    const int n = (int)seq_cnt[f];
    if (n > cap) {
        if (tid == 0) { atomicOr(&counters[CT_ERR], (unsigned long long)ERR_SORT_CAP); qs[f] = 0; }
        return;
    }
    int p2 = 1;
    while (p2 < n) p2 <<= 1;
    for (int i = tid; i < p2; i += blockDim.x) s_h[i] = i < n ? qhash[b + i] : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < p2; i += blockDim.x) {
                int l = i ^ j;
                if (l > i) {
                    uint32_t a = s_h[i], c = s_h[l];
                    bool up = (i & k) == 0;
                    if ((a > c) == up) { s_h[i] = c; s_h[l] = a; }
                }
            }
            __syncthreads();
        }
    // unique (stable compaction of run heads)
    uint32_t carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + tid;
        uint32_t head = (i < n && (i == 0 || s_h[i] != s_h[i - 1])) ? 1u : 0u;
        uint32_t tot, off = block_excl_scan<256>(head, s_warp, &tot);
        if (head) qhash[b + carry + off] = s_h[i];
        carry += tot;
    }
    if (tid == 0) {
        qs[f] = (int32_t)carry;
        atomicMax(&counters[CT_MAXS], (unsigned long long)carry);
        atomicAdd(&counters[CT_SKETCH_SUM], (unsigned long long)carry);
        if ((int)carry > s_max) atomicOr(&counters[CT_ERR], (unsigned long long)ERR_S_MAX);
    }
}

// ---- sketch exchange (reference-sharded layout, fa_comm.cu) -------------------------------------
// This rank's share, packed for the all-gather: fragment lf -> [stride hashes] at send + lf * stride, its size at
// send[per * stride + lf].  Size -1: more minimizers than the on-chip sort holds; -3: the sketch does not fit `stride`
// (the pass that meets such a fragment sketches its queries itself).
__global__ void pack_sketch_kernel(const uint32_t *qhash, const uint64_t *seq_first, const uint32_t *seq_cnt, const int32_t *qs, int n_local,
                                   int sort_cap, uint32_t per, uint32_t stride, uint32_t *send)
{
    const uint32_t lf = blockIdx.x;
    int s = 0;
    if ((int)lf < n_local) {
        s = qs[lf];
        if ((int)seq_cnt[lf] > sort_cap) s = -1;
        else if (s > (int)stride) s = -3;
        const uint64_t b = seq_first[lf];
        for (int i = threadIdx.x; i < s; i += blockDim.x) send[(size_t)lf * stride + i] = qhash[b + i];
    }
    if (threadIdx.x == 0) send[(size_t)per * stride + lf] = (uint32_t)s;
}

// The fragments of one pass out of the gathered shares, into the layout the rest of the pipeline reads (slot f * stride
// of qhash), with the counters sort_unique_kernel keeps.
__global__ void import_sketch_kernel(const uint32_t *recv, uint32_t per, uint32_t stride, uint64_t block, uint64_t first, int s_max,
                                     uint32_t *qhash, uint64_t *seq_first, int32_t *qs, unsigned long long *counters)
{
    const int f = blockIdx.x;
    const uint64_t g = first + (uint64_t)f, r = g / per, lf = g % per;
    const uint32_t *src = recv + r * block + lf * stride;
    int s = (int)recv[r * block + (uint64_t)per * stride + lf];
    if (s < 0) {
        if (threadIdx.x == 0) atomicOr(&counters[CT_ERR], (unsigned long long)(s == -1 ? ERR_SORT_CAP : ERR_EXCH));
        s = 0;
    }
    for (int i = threadIdx.x; i < s; i += blockDim.x) qhash[(size_t)f * stride + i] = src[i];
    if (threadIdx.x == 0) {
        seq_first[f] = (uint64_t)f * stride;
        qs[f] = s;
        atomicMax(&counters[CT_MAXS], (unsigned long long)s);
        atomicAdd(&counters[CT_SKETCH_SUM], (unsigned long long)s);
        if (s > s_max) atomicOr(&counters[CT_ERR], (unsigned long long)ERR_S_MAX);
    }
}

// ---- lookup of every sketch hash in the reference index (pyx:941-948) ------------------------
__global__ void lookup_kernel(const uint32_t *qhash, const uint64_t *seq_first, const int32_t *qs, int n_frags,
                              const uint32_t *dir, int dir_bits, const uint32_t *ukeys, const uint32_t *uoff,
                              uint32_t *hit_start, uint32_t *hit_cnt, uint64_t *frag_seeds)
{
    __shared__ unsigned long long s_sum;
    const int f = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_sum = 0;
    __syncthreads();
    const uint64_t b = seq_first[f];
    const int s = qs[f];
    unsigned long long mine = 0;
    for (int i = tid; i < s; i += blockDim.x) {
        const uint32_t h = qhash[b + i];
        const uint32_t d = h >> (32 - dir_bits);
        uint32_t lo = dir[d], hi = dir[d + 1];
        while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (ukeys[mid] < h) lo = mid + 1; else hi = mid; }
        uint32_t st = 0, cnt = 0;
        if (lo < dir[d + 1] && ukeys[lo] == h) { st = uoff[lo]; cnt = uoff[lo + 1] - st; }
        hit_start[b + i] = st; hit_cnt[b + i] = cnt;
        mine += cnt;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_sum, mine);
    __syncthreads();
    if (tid == 0) frag_seeds[f] = s_sum;
    if (f == 0 && tid == 0) frag_seeds[n_frags] = 0;
}

// ---- seed hits: the concatenated position lists, tagged with the fragment -------------------
__global__ void fill_seeds_kernel(const uint64_t *seq_first, const int32_t *qs, const uint32_t *hit_start,
                                  const uint32_t *hit_cnt, const uint64_t *seed_base, const uint32_t *pos_idx,
                                  int shift, uint64_t *seeds)
{
    __shared__ uint32_t s_warp[4];
    __shared__ uint32_t s_off[128], s_st[128], s_cnt[128];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t b = seq_first[f];
    const int s = qs[f];
    uint64_t out = seed_base[f];
    if (seed_base[f + 1] == out) return;                 // not a fragment of this path (or no seeds)
    const uint64_t tag = (uint64_t)f << shift;
    for (int base = 0; base < s; base += 128) {
        const int i = base + tid;
        uint32_t cnt = i < s ? hit_cnt[b + i] : 0, tot;
        uint32_t off = block_excl_scan<128>(cnt, s_warp, &tot);
        s_off[tid] = off; s_cnt[tid] = cnt; s_st[tid] = i < s ? hit_start[b + i] : 0;
        __syncthreads();
        for (int q = wid; q < 128; q += 4) {
            const uint32_t c = s_cnt[q], st = s_st[q];
            const uint64_t o = out + s_off[q];
            for (uint32_t t = lane; t < c; t += 32) seeds[o + t] = tag | pos_idx[st + t];
        }
        out += tot;
        __syncthreads();
    }
}

// ---- L1 candidate regions (computeMap.hpp:310-350) -------------------------------------------
// Seeds of a fragment are sorted by reference index == (seqId, wpos) order.  For the seed at t
// and the one m-1 further: same contig and closer than a fragment => raw candidate
// [max(0, wpos_b - L + 1), wpos_a].  Raw candidates are merged while the previous end reaches the
// next start; starts and ends are non-decreasing, so a merged region begins exactly where
// `previous raw end < start` (or the contig changes).
template <bool FILL>
__global__ void __launch_bounds__(256)
candidates_kernel(const uint64_t *seeds, const uint64_t *seed_base, const int32_t *qs, const int32_t *min_hits,
                  const RefMini *ref, uint64_t idx_mask, int frag_len, uint32_t *frag_cands,
                  const uint32_t *cand_base, Cand *cands)
{
    __shared__ uint32_t s_warp[8];
    __shared__ int s_wlast[8];
    __shared__ int s_seq[256], s_end[256];
    __shared__ int s_carry_seq, s_carry_end, s_carry_valid;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t b = seed_base[f], e = seed_base[f + 1];
    const int s = qs[f];
    if (s <= 0 || b == e) return;                         // fragments without seeds here belong to l1_fused_kernel
    const int m = min_hits[s];
    if (tid == 0) { s_carry_valid = 0; s_carry_seq = -1; s_carry_end = 0; }
    uint32_t heads_before = 0;
    const uint32_t out0 = FILL ? cand_base[f] : 0;
    __syncthreads();
    for (uint64_t base = b; base < e; base += 256) {
        const uint64_t t = base + tid;
        bool valid = false;
        int seq = -1, start = 0, end = 0;
        uint32_t ja = 0, jb = 0;
        if (t + (uint64_t)(m - 1) < e) {
            ja = (uint32_t)(seeds[t] & idx_mask);
            jb = (uint32_t)(seeds[t + m - 1] & idx_mask);
            const RefMini ra = ref[ja], rb = ref[jb];
            if (ra.z == rb.z && (int)(rb.y - ra.y) < frag_len) {
                valid = true; seq = (int)ra.z; end = (int)ra.y;
                start = max(0, (int)rb.y - frag_len + 1);
            }
        }
        s_seq[tid] = seq; s_end[tid] = end;
        // index (within the chunk) of the last valid raw candidate before this one, -1 = none in chunk
        int last = valid ? tid : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xFFFFFFFFu, last, o); if (lane >= o) last = max(last, v); }
        if (lane == 31) s_wlast[wid] = last;
        __syncthreads();
        int prev = __shfl_up_sync(0xFFFFFFFFu, last, 1);
        if (lane == 0) prev = -1;
        for (int q = 0; q < wid; q++) prev = max(prev, s_wlast[q]);
        bool head = false;
        if (valid) {
            int pseq, pend, pvalid;
            if (prev >= 0) { pseq = s_seq[prev]; pend = s_end[prev]; pvalid = 1; }
            else { pseq = s_carry_seq; pend = s_carry_end; pvalid = s_carry_valid; }
            head = !pvalid || pseq != seq || pend < start;
        }
        uint32_t tot, off = block_excl_scan<256>(head ? 1u : 0u, s_warp, &tot);
        if (FILL) {
            const uint32_t slot = out0 + heads_before + off;            // heads: their own slot
            if (head) cands[slot] = Cand{f, jb, ja, 0u};
            __syncthreads();
            if (valid && !head) atomicMax(&cands[slot - 1].tail, ja);     // members: the open region (indices grow with positions)
        }
        heads_before += tot;
        __syncthreads();
        if (tid == 255 || t + 1 == e) {
            // carry the last valid raw candidate of this chunk into the next one
            int lastv = max(prev, valid ? tid : -1);
            if (lastv >= 0) { s_carry_seq = s_seq[lastv]; s_carry_end = s_end[lastv]; s_carry_valid = 1; }
        }
        __syncthreads();
    }
    if (!FILL && tid == 0) frag_cands[f] = heads_before;
}


// ---- L1 on chip: seeds -> candidate regions without leaving the SM -----------------------------
// The seed hits of one fragment (pyx:941-948: the concatenated position lists of its sketch
// hashes) must be visited in (seqId, wpos) = reference-index order (computeMap.hpp:316).  Each
// position list is already sorted and the hits of a fragment cluster in a few hundred loci, so a
// device-wide 64-bit radix sort (the general path above) moves far more bytes than needed.  Here
// one CTA owns a fragment:
//   A  count the hits per chunk of 2^16 reference minimizers (shared-memory histogram)
//   B  scatter them, as 16-bit offsets, into chunk order (the histogram becomes the cursors)
//   C  sort every chunk bucket by one warp: a 1024-bit bitmap when the bucket spans < 1024
//      indices (a locus: ~50 hits over ~250 minimizers), an in-place bitonic network otherwise
//   D  tiles of 4096 sorted hits: gather the running coordinate gpos (fa_index.cu), test the
//      pairs (t, t + minHits - 1) of computeMap.hpp:320-334, find region heads / ends of the
//      merge step (:338-347) from the bitmap of valid pairs -- no sequential carry but one
//      (index, gpos) pair per tile -- and write the regions to a per-fragment scratch range.
// Seeds never touch HBM; the only traffic is pos_idx (twice, the second time from L2) and the
// gathers of gpos / hw.  Fragments whose hits do not fit the CTA's shared memory take the
// radix-sort path.
// Two shapes of the same kernel: 1024 threads and tiles of 4096 hits for fragments with tens of thousands of hits
// (one CTA per SM, its shared memory holds the hits), 256 threads and tiles of 1024 for fragments with a few
// thousand (many-to-many workloads: several CTAs per SM hide each other's barriers and gathers).
// Parts (round 2): a fragment with tens of thousands of hits (every reference genome related to the query) needs the
// 1024-thread shape with all of its hits in one CTA's shared memory -- one CTA per SM, every barrier exposed.  Candidate
// regions never span two genomes, so such a fragment is cut at genome boundaries into L1_PARTS_MAX or fewer parts
// that are mapped by independent CTAs of the small shape (four per SM).  split_lists_kernel cuts every position list of
// the fragment at the part boundaries (the lists are ascending reference indices) and counts the hits per part.
constexpr int L1_SHIFT = 16;             // chunks of 2^16 reference minimizers
constexpr int L1_PARTS_MAX = 8;
struct PartCfg { uint32_t first[L1_PARTS_MAX + 1]; };       // first reference index of every part (and one past the last)
__device__ __forceinline__ uint32_t part_first(const PartCfg &c, int p)
{
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q <= L1_PARTS_MAX; q++) r = q == p ? c.first[q] : r;    // (constant indices: the table stays in the parameter bank)
    return r;
}
__global__ void __launch_bounds__(256)
split_lists_kernel(const uint64_t *seq_first, const int32_t *qs, const uint32_t *hit_start, const uint32_t *hit_cnt,
                   const uint64_t *seed_base, const uint32_t *pos_idx, PartCfg cfg, int n_parts,
                   uint32_t seed_lo, uint32_t seed_cap, int s_stride, uint32_t *split, uint32_t *part_off)
{
    __shared__ uint32_t s_cnt[L1_PARTS_MAX + 1];
    const int f = blockIdx.x, tid = threadIdx.x;
    const uint64_t nf = seed_base[f + 1] - seed_base[f];
    if (nf > (uint64_t)seed_cap || nf < (uint64_t)seed_lo) return;
    const int s = qs[f];
    const uint64_t qb = seq_first[f];
    if (tid <= L1_PARTS_MAX) s_cnt[tid] = 0u;
    __syncthreads();
    // one thread per (list, inner boundary): the first entry at or behind the boundary
    for (int i = tid; i < s * (n_parts + 1); i += blockDim.x) {
        const int q = i / (n_parts + 1), p = i - q * (n_parts + 1);
        const uint32_t st = hit_start[qb + q], c = hit_cnt[qb + q];
        uint32_t lo = st, hi = st + c;
        if (p == 0) hi = lo;
        else if (p == n_parts) lo = hi;
        else {
            const uint32_t r = part_first(cfg, p);
            while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (__ldg(pos_idx + mid) < r) lo = mid + 1; else hi = mid; }
        }
        split[((size_t)f * s_stride + q) * (n_parts + 1) + p] = lo;
        // entries before boundary p, summed over the lists: hits of the parts below p
        if (p > 0 && lo > st) atomicAdd(&s_cnt[p], lo - st);
    }
    __syncthreads();
    if (tid <= n_parts) part_off[(size_t)f * (n_parts + 1) + tid] = s_cnt[tid];     // (s_cnt[0] = 0, s_cnt[n_parts] = all hits)
}

// Where the hits of the heavy fragments lie: hits per chunk of 2^16 reference minimizers, summed over a sample of the
// fragments that have at least `min_seeds` hits.  The host cuts the index into parts of about equal hit counts with it
// (related genomes are not spread evenly: the references of BASELINE configs[1] are ordered by identity).
constexpr int L1_SAMPLE = 32, L1_SAMPLE_SLICES = 16;    // sampled fragments; CTAs that share the lists of one (the walk is latency bound)
__global__ void __launch_bounds__(256)
sample_chunks_kernel(const uint64_t *seq_first, const int32_t *qs, const uint32_t *hit_start, const uint32_t *hit_cnt,
                     const uint64_t *frag_seed_counts, int n_frags, uint32_t min_seeds, const uint32_t *pos_idx, uint32_t *chunk_hist)
{
    const int n_samp = gridDim.x / L1_SAMPLE_SLICES, slice = blockIdx.x % L1_SAMPLE_SLICES;
    const int f = (int)(((long long)(blockIdx.x / L1_SAMPLE_SLICES) * n_frags) / n_samp);
    if (frag_seed_counts[f] < (uint64_t)min_seeds) return;
    const int s = qs[f], lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t qb = seq_first[f];
    for (int q = slice * 8 + wid; q < s; q += 8 * L1_SAMPLE_SLICES) {
        const uint32_t st = hit_start[qb + q], c = hit_cnt[qb + q];
        for (uint32_t t = lane; t < c; t += 32) atomicAdd(&chunk_hist[__ldg(pos_idx + st + t) >> L1_SHIFT], 1u);
    }
}

constexpr int L1L_THREADS = 1024, L1L_TILE = 4096;     // the large shape
constexpr int L1S_THREADS = 256, L1S_TILE = 1024;      // the small shape
// its shared memory per CTA: four CTAs per SM; three or two when the chunk histogram of a large index (4 B per 2^16
// reference minimizers: 46 KB for 2 000 genomes) leaves no room for the hits otherwise
constexpr size_t L1S_SMEM[3] = {54 * 1024, 72 * 1024, 110 * 1024};
constexpr int L1_BM = 32;                     // bitmap words per warp in phase C (+ as many prefix words, + as many stray keys)
// look-ahead of the pair test: minHits - 1 <= THREADS; staged hits per tile = TILE + THREADS (also holds the s position lists)
__host__ __device__ constexpr int l1_stage(int threads, int tile) { return tile + threads; }

__host__ __device__ inline size_t l1_fixed_smem(uint32_t n_chunks, int stage)
{
    return (size_t)((n_chunks + 1 + 3) & ~3u) * 4 + (size_t)stage * 8;
}

// in-place sort of k[0, n) by one warp
__device__ __forceinline__ void l1_sort_bucket(uint16_t *k, int n, uint32_t *bm, int lane)
{
    if (n <= 64) {
        // the usual bucket (one locus): both keys of a lane stay in registers; rank = set bits below in a 1024-bit map
        const uint32_t v0 = lane < n ? (uint32_t)k[lane] : 0xFFFFFFFFu, v1 = lane + 32 < n ? (uint32_t)k[lane + 32] : 0xFFFFFFFFu;
        const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, min(v0, v1));
        const uint32_t hi = __reduce_max_sync(0xFFFFFFFFu, max(lane < n ? v0 : 0u, lane + 32 < n ? v1 : 0u));
        if (hi - lo < (uint32_t)(32 * L1_BM)) {
            bm[lane] = 0u;
            __syncwarp();
            const uint32_t d0 = v0 - lo, d1 = v1 - lo;
            if (lane < n) atomicOr(&bm[d0 >> 5], 1u << (d0 & 31u));
            if (lane + 32 < n) atomicOr(&bm[d1 >> 5], 1u << (d1 & 31u));
            __syncwarp();
            const uint32_t wv = bm[lane], cnt = __popc(wv);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            bm[L1_BM + lane] = incl - cnt;
            __syncwarp();
            if (lane < n) k[bm[L1_BM + (d0 >> 5)] + __popc(bm[d0 >> 5] & ((1u << (d0 & 31u)) - 1u))] = (uint16_t)v0;
            if (lane + 32 < n) k[bm[L1_BM + (d1 >> 5)] + __popc(bm[d1 >> 5] & ((1u << (d1 & 31u)) - 1u))] = (uint16_t)v1;
            __syncwarp();
            return;
        }
        if (n <= 16) {
            // a handful of chance hits spread over the chunk (the usual bucket of a fragment that is unrelated to the
            // genomes of this chunk): rank = keys below, by shuffles
            int r = 0;
            for (int j = 0; j < n; j++) r += __shfl_sync(0xFFFFFFFFu, v0, j) < v0;
            __syncwarp();                                        // (every lane has read its key)
            if (lane < n) k[r] = (uint16_t)v0;
            __syncwarp();
            return;
        }
    }
    if (n <= 256) {
        // up to eight keys per lane in registers, rank = set bits below in a 1024-bit map (as above)
        uint32_t v[8], lo = 0xFFFFu, hi = 0u;
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (32 * u >= n) break;
            const int i = lane + 32 * u;
            if (i < n) { v[u] = (uint32_t)k[i]; lo = min(lo, v[u]); hi = max(hi, v[u]); }
        }
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (hi - lo < (uint32_t)(32 * L1_BM)) {
            bm[lane] = 0u;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (32 * u >= n) break;
                if (lane + 32 * u < n) { const uint32_t d = v[u] - lo; atomicOr(&bm[d >> 5], 1u << (d & 31u)); }
            }
            __syncwarp();
            const uint32_t cnt = __popc(bm[lane]);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            bm[L1_BM + lane] = incl - cnt;
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (32 * u >= n) break;
                if (lane + 32 * u < n) { const uint32_t d = v[u] - lo; k[bm[L1_BM + (d >> 5)] + __popc(bm[d >> 5] & ((1u << (d & 31u)) - 1u))] = (uint16_t)v[u]; }
            }
            __syncwarp();
            return;
        }
        // A locus and a few stray hits far from it (a k-mer that occurs twice in the genome, a second locus at the other
        // end of the chunk): the 1024 positions from the smallest key up, or up to the largest, whichever hold more keys,
        // go through the bitmap; the rest -- all on one side of them -- are ranked among themselves.
        int c_lo = 0, c_hi = 0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (32 * u >= n) break;
            const bool ok = lane + 32 * u < n;
            c_lo += __popc(__ballot_sync(0xFFFFFFFFu, ok && v[u] - lo < (uint32_t)(32 * L1_BM)));
            c_hi += __popc(__ballot_sync(0xFFFFFFFFu, ok && hi - v[u] < (uint32_t)(32 * L1_BM)));
        }
        const bool at_lo = c_lo >= c_hi;
        const int n_in = at_lo ? c_lo : c_hi, n_out = n - n_in;
        if (n_out <= 32) {
            const uint32_t w0 = at_lo ? lo : hi - (uint32_t)(32 * L1_BM - 1);
            uint32_t *stray = bm + 2 * L1_BM;
            bm[lane] = 0u;
            __syncwarp();
            int seen = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (32 * u >= n) break;
                const bool ok = lane + 32 * u < n;
                const uint32_t d = v[u] - w0;
                const bool in = ok && d < (uint32_t)(32 * L1_BM);
                if (in) atomicOr(&bm[d >> 5], 1u << (d & 31u));
                const unsigned om = __ballot_sync(0xFFFFFFFFu, ok && !in);
                if (ok && !in) stray[seen + __popc(om & ((1u << lane) - 1u))] = v[u];
                seen += __popc(om);
            }
            __syncwarp();
            const uint32_t cnt = __popc(bm[lane]);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            bm[L1_BM + lane] = incl - cnt;
            __syncwarp();
            const int base_in = at_lo ? 0 : n_out, base_out = at_lo ? n_in : 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (32 * u >= n) break;
                if (lane + 32 * u < n) {
                    const uint32_t d = v[u] - w0;
                    int pos;
                    if (d < (uint32_t)(32 * L1_BM)) pos = base_in + (int)(bm[L1_BM + (d >> 5)] + __popc(bm[d >> 5] & ((1u << (d & 31u)) - 1u)));
                    else {
                        pos = base_out;
                        for (int j = 0; j < n_out; j++) pos += stray[j] < v[u];
                    }
                    k[pos] = (uint16_t)v[u];
                }
            }
            __syncwarp();
            return;
        }
    } else {
        uint32_t lo = 0xFFFFu, hi = 0u;
        for (int i = lane; i < n; i += 32) { const uint32_t v = k[i]; lo = min(lo, v); hi = max(hi, v); }
        lo = __reduce_min_sync(0xFFFFFFFFu, lo);
        hi = __reduce_max_sync(0xFFFFFFFFu, hi);
        if (hi - lo < (uint32_t)(32 * L1_BM)) {
            bm[lane] = 0u;
            __syncwarp();
            for (int i = lane; i < n; i += 32) { const uint32_t d = (uint32_t)k[i] - lo; atomicOr(&bm[d >> 5], 1u << (d & 31u)); }
            __syncwarp();
            uint32_t wv = bm[lane];
            const uint32_t cnt = __popc(wv);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            int pos = (int)(incl - cnt);
            const uint32_t base = lo + (uint32_t)lane * 32u;
            while (wv) { const int bit = __ffs(wv) - 1; wv &= wv - 1u; k[pos++] = (uint16_t)(base + (uint32_t)bit); }
            __syncwarp();
            return;
        }
    }
    // bitonic network with ascending comparators only (first step of every merge mirrors the
    // upper half), so slots at or beyond n act as +infinity and are never touched
    int lg = 1;
    while ((1 << lg) < n) lg++;
    const int half = 1 << (lg - 1);
    for (int ks = 1; ks <= lg; ks++) {                       // merge size 2^ks
        const int hs = ks - 1, hm = (1 << hs) - 1;
        for (int idx = lane; idx < half; idx += 32) {
            const int blk = idx >> hs, off = idx & hm;
            const int i = (blk << ks) + off, l = (blk << ks) + (1 << ks) - 1 - off;
            if (l < n) { const uint16_t a = k[i], b = k[l]; if (a > b) { k[i] = b; k[l] = a; } }
        }
        __syncwarp();
        for (int js = hs - 1; js >= 0; js--) {               // half-cleaners of distance 2^js
            const int jm = (1 << js) - 1;
            for (int idx = lane; idx < half; idx += 32) {
                const int i = ((idx >> js) << (js + 1)) + (idx & jm), l = i + (1 << js);
                if (l < n) { const uint16_t a = k[i], b = k[l]; if (a > b) { k[i] = b; k[l] = a; } }
            }
            __syncwarp();
        }
    }
}

// What the parts mode hands to the kernel (PARTS) and what the large shape needs to step in for a fragment one of whose
// parts did not fit (n_parts > 0 with PARTS == false: only fragments flagged in `overflow` are mapped).
struct L1Parts {
    int n_parts, s_stride;
    const uint32_t *split;        // [fragment][list][n_parts + 1]: the cuts of every position list
    const uint32_t *part_off;     // [fragment][n_parts + 1]: hits below every part
    uint32_t *part_cands;         // [fragment][n_parts]: regions found by every part
    uint32_t *overflow;           // [fragment]: some part held more hits than a CTA of the small shape stages
    PartCfg cfg;
};

template <int L1_THREADS, int L1_TILE, bool PARTS>
__global__ void __launch_bounds__(L1_THREADS, 1024 / L1_THREADS)
l1_fused_kernel(const uint64_t *seq_first, const int32_t *qs, const uint32_t *hit_start, const uint32_t *hit_cnt,
                const uint64_t *seed_base, const uint32_t *pos_idx, const uint32_t *gpos, const uint32_t *irr, const uint2 *hw,
                const int32_t *min_hits, int frag_len, uint32_t d_near, uint32_t n_chunks, uint32_t seed_lo, uint32_t seed_cap,
                uint32_t key_cap, Cand *tmp, uint32_t *frag_cands, const L1Parts pt)
{
    constexpr int L1_PER = L1_TILE / L1_THREADS;
    constexpr int L1_STAGE = l1_stage(L1_THREADS, L1_TILE);
    extern __shared__ __align__(16) uint8_t l1_smem[];
    uint32_t *hist = reinterpret_cast<uint32_t *>(l1_smem);                 // [n_chunks]: counts -> cursors -> bucket ends
    uint32_t *s_j = hist + ((n_chunks + 1 + 3) & ~3u);                       // phase D: reference index per staged hit
    uint32_t *s_g = s_j + L1_STAGE;                                          //          its gpos
    uint32_t *s_bm = s_j;                                                    // phase C: per-warp bitmaps (aliases s_j)
    uint32_t *s_lst = s_j, *s_lcnt = s_g;                                    // phases A, B: the position lists (alias)
    uint16_t *keys = reinterpret_cast<uint16_t *>(s_g + L1_STAGE);           // [key_cap], key_cap a multiple of 32
    uint16_t *blk_chunk = keys + key_cap;                                    // chunk holding sorted hit 32 * q
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_valid[L1_TILE / 32], s_head[L1_TILE / 32], s_hpre[L1_TILE / 32 + 1];
    __shared__ uint32_t s_blk, s_nlist, s_cj, s_cf;
    __shared__ int s_cvalid;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int f = PARTS ? (int)(blockIdx.x / (unsigned)pt.n_parts) : (int)blockIdx.x;
    const int part = PARTS ? (int)blockIdx.x - f * pt.n_parts : 0;
    uint64_t sb = seed_base[f];
    const uint64_t nf64 = seed_base[f + 1] - sb;
    if (nf64 > (uint64_t)seed_cap) return;                                   // the radix-sort path takes this fragment
    if (nf64 < (uint64_t)seed_lo) return;                                    // the other shape of this kernel does
    if (!PARTS && pt.n_parts > 0 && pt.overflow[f] == 0u) return;            // (large shape behind the parts: flagged fragments only)
    const int s = qs[f];
    if (s <= 0 || nf64 == 0) { if (tid == 0) frag_cands[f] = 0; return; }
    uint32_t n = (uint32_t)nf64;
    const uint64_t qb = seq_first[f];
    uint32_t c_base = 0;                                                     // first chunk of this CTA's share of the index
    if (PARTS) {
        const uint32_t *po = pt.part_off + (size_t)f * (pt.n_parts + 1) + part;
        sb += po[0];
        n = po[1] - po[0];
        const uint32_t r0 = part_first(pt.cfg, part), r1 = part_first(pt.cfg, part + 1);
        c_base = r0 >> L1_SHIFT;
        // a part that does not fit (its genomes hold more of the hits than their share): the whole fragment goes to the
        // large shape, launched behind this kernel
        if (n > key_cap || (r1 > r0 && ((r1 - 1u) >> L1_SHIFT) - c_base + 1u > n_chunks)) {
            if (tid == 0) { pt.overflow[f] = 1u; pt.part_cands[(size_t)f * pt.n_parts + part] = 0u; }
            return;
        }
        if (n == 0) { if (tid == 0) pt.part_cands[(size_t)f * pt.n_parts + part] = 0u; return; }
    }

    for (uint32_t i = tid; i <= n_chunks; i += L1_THREADS) hist[i] = 0u;
    if (PARTS) {
        const uint32_t *sp = pt.split + ((size_t)f * pt.s_stride) * (pt.n_parts + 1) + part;
        for (int q = tid; q < s; q += L1_THREADS) { const uint32_t a = sp[(size_t)q * (pt.n_parts + 1)]; s_lst[q] = a; s_lcnt[q] = sp[(size_t)q * (pt.n_parts + 1) + 1] - a; }
    } else {
        for (int q = tid; q < s; q += L1_THREADS) { s_lst[q] = hit_start[qb + q]; s_lcnt[q] = hit_cnt[qb + q]; }
    }
    if (tid == 0) { s_blk = 0u; s_nlist = 0u; s_cvalid = 0; s_cj = 0u; s_cf = 0u; }
    __syncthreads();

    // The small shape serves fragments whose position lists hold a handful of entries each (a hash occurs once per
    // related genome): a warp per list would leave most lanes idle and walk ~30 lists one DRAM round trip after the
    // other.  There the hits are numbered through all lists instead -- s_lcnt becomes the exclusive prefix of the
    // list lengths -- and thread t fetches hit t, t + THREADS, ...: every load of a phase is in flight at once.
    constexpr bool FLAT = L1_THREADS <= 256 && !PARTS;      // (the lists of a part hold dozens of entries: a warp per list)
    constexpr bool LISTC = true;                            // phase C takes the buckets from a list, one by one (false: blocks of 32 chunks per warp)
    auto flat_hit = [&](uint32_t t) -> uint32_t {             // the reference index of hit t
        int lo = 0, hi = s - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_lcnt[mid] <= t) lo = mid; else hi = mid - 1; }
        return __ldg(pos_idx + s_lst[lo] + (t - s_lcnt[lo]));
    };
    if (FLAT) {
        const int per = (s + L1_THREADS - 1) / L1_THREADS;
        const int i0 = min(tid * per, s), i1 = min(i0 + per, s);
        uint32_t sum = 0;
        for (int i = i0; i < i1; i++) sum += s_lcnt[i];
        uint32_t tot, x = block_excl_scan<L1_THREADS>(sum, s_warp, &tot);
        for (int i = i0; i < i1; i++) { const uint32_t v = s_lcnt[i]; s_lcnt[i] = x; x += v; }
        __syncthreads();
    }
    // ---- A: histogram over chunks ------------------------------------------------------------
    if (FLAT) {
        for (uint32_t t0 = tid; t0 < n; t0 += 4 * L1_THREADS) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const uint32_t t = t0 + u * L1_THREADS; v[u] = t < n ? flat_hit(t) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < 4; u++) if (v[u] != 0xFFFFFFFFu) atomicAdd(&hist[v[u] >> L1_SHIFT], 1u);
        }
    } else {
        for (int q = wid; q < s; q += L1_THREADS / 32) {
            const uint32_t st = s_lst[q], c = s_lcnt[q];
            for (uint32_t t0 = 0; t0 < c; t0 += 128) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const uint32_t t = t0 + u * 32 + lane; v[u] = t < c ? __ldg(pos_idx + st + t) : 0xFFFFFFFFu; }
#pragma unroll
                for (int u = 0; u < 4; u++) if (v[u] != 0xFFFFFFFFu) atomicAdd(&hist[(v[u] >> L1_SHIFT) - c_base], 1u);
            }
        }
    }
    __syncthreads();
    // ---- exclusive scan: hist[c] = first slot of bucket c --------------------------------------
    {
        const uint32_t per = (n_chunks + L1_THREADS - 1) / L1_THREADS;
        const uint32_t i0 = min((uint32_t)tid * per, n_chunks), i1 = min(i0 + per, n_chunks);
        uint32_t sum = 0;
        for (uint32_t i = i0; i < i1; i++) sum += hist[i];
        uint32_t tot, x = block_excl_scan<L1_THREADS>(sum, s_warp, &tot);
        for (uint32_t i = i0; i < i1; i++) { const uint32_t v = hist[i]; hist[i] = x; x += v; }
    }
    __syncthreads();
    // ---- B: scatter the low 16 bits into chunk order; afterwards hist[c] = end of bucket c -------
    if (FLAT) {
        for (uint32_t t0 = tid; t0 < n; t0 += 4 * L1_THREADS) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const uint32_t t = t0 + u * L1_THREADS; v[u] = t < n ? flat_hit(t) : 0xFFFFFFFFu; }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (v[u] != 0xFFFFFFFFu) { const uint32_t slot = atomicAdd(&hist[v[u] >> L1_SHIFT], 1u); keys[slot] = (uint16_t)v[u]; }
        }
    } else {
        for (int q = wid; q < s; q += L1_THREADS / 32) {
            const uint32_t st = s_lst[q], c = s_lcnt[q];
            for (uint32_t t0 = 0; t0 < c; t0 += 128) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const uint32_t t = t0 + u * 32 + lane; v[u] = t < c ? __ldg(pos_idx + st + t) : 0xFFFFFFFFu; }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (v[u] != 0xFFFFFFFFu) { const uint32_t slot = atomicAdd(&hist[(v[u] >> L1_SHIFT) - c_base], 1u); keys[slot] = (uint16_t)v[u]; }
            }
        }
    }
    __syncthreads();
    // ---- C: sort the buckets ------------------------------------------------------------------------
    if (LISTC) {
        // The hits of a many-to-many fragment sit in a few dozen chunks next to each other (the genomes of its genus):
        // blocks of 32 chunks would hand all of them to two or three warps.  The buckets that need sorting are listed
        // first (s_g is free since phase B) and the warps take them one by one.
        uint32_t *list = s_g;
        for (uint32_t c = tid; c < n_chunks; c += L1_THREADS) {
            const uint32_t b0 = c ? hist[c - 1] : 0u, sz = hist[c] - b0;
            for (uint32_t t0 = (b0 + 31u) & ~31u; t0 < b0 + sz; t0 += 32u) blk_chunk[t0 >> 5] = (uint16_t)c;
            if (sz >= 2u) { const uint32_t i = atomicAdd(&s_nlist, 1u); if (i < (uint32_t)L1_STAGE) list[i] = c; }
        }
        __syncthreads();
        const uint32_t n_list = s_nlist;
        if (n_list <= (uint32_t)L1_STAGE) {
            for (;;) {
                uint32_t i = 0;
                if (lane == 0) i = atomicAdd(&s_blk, 1u);
                i = __shfl_sync(0xFFFFFFFFu, i, 0);
                if (i >= n_list) break;
                const uint32_t c = list[i];
                const uint32_t b0 = c ? hist[c - 1] : 0u;
                l1_sort_bucket(keys + b0, (int)(hist[c] - b0), s_bm + wid * (3 * L1_BM), lane);
            }
        }
    }
    if (!LISTC || s_nlist > (uint32_t)L1_STAGE) {           // (the list overflowed: blocks of 32 chunks, as in the large shape)
        for (;;) {
            uint32_t blk = 0;
            if (lane == 0) blk = atomicAdd(&s_blk, 1u);
            blk = __shfl_sync(0xFFFFFFFFu, blk, 0);
            const uint32_t c = blk * 32u + (uint32_t)lane;
            if (blk * 32u >= n_chunks) break;
            uint32_t b0 = 0, sz = 0;
            if (c < n_chunks) { b0 = c ? hist[c - 1] : 0u; sz = hist[c] - b0; }
            for (uint32_t t0 = (b0 + 31u) & ~31u; t0 < b0 + sz; t0 += 32u) blk_chunk[t0 >> 5] = (uint16_t)c;
            unsigned todo = __ballot_sync(0xFFFFFFFFu, sz >= 2u);
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1u;
                l1_sort_bucket(keys + __shfl_sync(0xFFFFFFFFu, b0, l), (int)__shfl_sync(0xFFFFFFFFu, sz, l), s_bm + wid * (3 * L1_BM), lane);
            }
        }
    }
    __syncthreads();

    // ---- D: candidate regions -------------------------------------------------------------------
    const int m = min_hits[s];
    const uint32_t L = (uint32_t)frag_len;
    // "Reference minimizers a <= b lie closer than a fragment" (same contig, position distance < L; computeMap.hpp:325-330
    // and :338-340) is gpos[b] - gpos[a] < L.  Almost every time the index distance decides it without touching gpos:
    // positions grow by at least one per minimizer, so b - a >= L is far; and where every step is at most one window
    // (all blocks of 1024 minimizers between a and b unmarked in `irr`), (b - a) * window < L is near.  The two gathers
    // are left for what falls between (and for hits next to a contig end).
    // The mark of its block is staged with every hit, so the tests themselves read shared memory only.
    auto l1_near = [&](uint32_t a, uint32_t fa, uint32_t b2, uint32_t fb2) -> bool {
        const uint32_t d = b2 - a;
        if (d >= L) return false;
        if (d <= d_near && (fa | fb2) == 0u) return true;                        // (d_near < 1024: at most two blocks)
        return __ldg(gpos + b2) - __ldg(gpos + a) < L;
    };
    Cand *out = tmp + sb;
    uint32_t heads_before = 0;
    for (uint32_t base = 0; base < n; base += L1_TILE) {
        // D1: stage the reference indices of the tile and of the m - 1 hits behind it, each with the marks of its blocks
        // (a table of one bit per 1024 minimizers: it stays in L2, and all loads of a thread are in flight at once)
        {
            uint32_t jv[L1_PER + 1], w0[L1_PER + 1];
#pragma unroll
            for (int u = 0; u <= L1_PER; u++) {
                const int e0 = u * L1_THREADS + wid * 32;
                const uint32_t t = base + (uint32_t)(e0 + lane);
                jv[u] = 0xFFFFFFFFu;
                if (e0 < L1_TILE + m - 1 && t < n) {
                    // the chunk of hit t: the first bucket that ends behind it, between the chunks of the two hits that
                    // bracket its block of 32 (chance hits of a large index lie thousands of empty chunks apart)
                    const uint32_t q = (base + (uint32_t)e0) >> 5;
                    uint32_t c = blk_chunk[q], hi = ((q + 1u) << 5) < n ? (uint32_t)blk_chunk[q + 1u] : n_chunks - 1u;
                    while (c < hi) { const uint32_t mid = (c + hi) >> 1; if (hist[mid] <= t) c = mid + 1u; else hi = mid; }
                    jv[u] = ((c + c_base) << L1_SHIFT) | (uint32_t)keys[t];
                }
            }
#pragma unroll
            for (int u = 0; u <= L1_PER; u++) w0[u] = jv[u] != 0xFFFFFFFFu ? __ldg(irr + (jv[u] >> 15)) : 0u;
#pragma unroll
            for (int u = 0; u <= L1_PER; u++)
                if (jv[u] != 0xFFFFFFFFu) { const int e = u * L1_THREADS + tid; s_j[e] = jv[u]; s_g[e] = (w0[u] >> ((jv[u] >> 10) & 31u)) & 1u; }
        }
        __syncthreads();
        // D2: valid pairs (computeMap.hpp:325-330) as a bitmap
        uint32_t jb[L1_PER];
        bool valid[L1_PER];
#pragma unroll
        for (int u = 0; u < L1_PER; u++) {
            const int e = u * L1_THREADS + tid;
            const uint32_t t = base + (uint32_t)e;
            valid[u] = false; jb[u] = 0;
            if (t + (uint32_t)(m - 1) < n) {
                jb[u] = s_j[e + m - 1];
                valid[u] = l1_near(s_j[e], s_g[e], jb[u], s_g[e + m - 1]);
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, valid[u]);
            if (lane == 0) s_valid[e >> 5] = bal;
        }
        __syncthreads();
        // D3: heads: a valid pair whose predecessor among the valid pairs ends before its start (:338-340)
        uint32_t pj[L1_PER];
        bool head[L1_PER], pv[L1_PER];
#pragma unroll
        for (int u = 0; u < L1_PER; u++) {
            const int e = u * L1_THREADS + tid;
            head[u] = false; pv[u] = false; pj[u] = 0;
            if (valid[u]) {
                int w = e >> 5;
                uint32_t x = s_valid[w] & ((1u << lane) - 1u);
                while (x == 0u && w > 0) x = s_valid[--w];
                uint32_t pf;
                if (x) { const int p = w * 32 + 31 - __clz(x); pj[u] = s_j[p]; pf = s_g[p]; pv[u] = true; }
                else { pj[u] = s_cj; pf = s_cf; pv[u] = s_cvalid != 0; }
                head[u] = !pv[u] || !l1_near(pj[u], pf, jb[u], s_g[e + m - 1]);
            }
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, head[u]);
            if (lane == 0) s_head[e >> 5] = bal;
        }
        __syncthreads();
        if (wid == 0) {                                                   // exclusive prefix of the head counts per word
            uint32_t c4[L1_TILE / 1024], sum = 0;
#pragma unroll
            for (int q = 0; q < L1_TILE / 1024; q++) { c4[q] = __popc(s_head[lane * (L1_TILE / 1024) + q]); sum += c4[q]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
            uint32_t x = incl - sum;
#pragma unroll
            for (int q = 0; q < L1_TILE / 1024; q++) { s_hpre[lane * (L1_TILE / 1024) + q] = x; x += c4[q]; }
            if (lane == 31) s_hpre[L1_TILE / 32] = incl;
        }
        __syncthreads();
        // D4: write the heads; a head also closes the region before it (its end is the last member's wpos, :341-346)
#pragma unroll
        for (int u = 0; u < L1_PER; u++) {
            if (head[u]) {
                const int e = u * L1_THREADS + tid;
                const uint32_t slot = heads_before + s_hpre[e >> 5] + __popc(s_head[e >> 5] & ((1u << lane) - 1u));
                Cand *o = out + slot;
                o->frag = f; o->hint = jb[u]; o->spare = 0u;       // (tail: written by whoever closes the region, never here)
                if (pv[u]) out[slot - 1].tail = pj[u];
            }
        }
        heads_before += s_hpre[L1_TILE / 32];
        // carry: the last valid pair of the tile
        if (tid == 0) {
            int w = L1_TILE / 32 - 1;
            while (w >= 0 && s_valid[w] == 0u) w--;
            if (w >= 0) { const int p = w * 32 + 31 - __clz(s_valid[w]); s_cj = s_j[p]; s_cf = s_g[p]; s_cvalid = 1; }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (heads_before) out[heads_before - 1].tail = s_cj;
        if (PARTS) {
            pt.part_cands[(size_t)f * pt.n_parts + part] = heads_before;
            atomicAdd(&frag_cands[f], heads_before);             // (zeroed before the launch)
        } else {
            frag_cands[f] = heads_before;
            if (pt.n_parts > 0) {                                // in place of the parts of this fragment: everything is "part 0"
                for (int p2 = 0; p2 < pt.n_parts; p2++) pt.part_cands[(size_t)f * pt.n_parts + p2] = p2 ? 0u : heads_before;
            }
        }
    }
}

// The third shape: one WARP per fragment, for fragments with at most L1_TINY hits -- what a fragment of a many-to-many run
// meets in one reference shard (a few related genomes and the chance hits of its 32-bit hashes).  The two CTA shapes
// pay for a histogram over all chunks of the index and a dozen barriers whatever the hit count; here the hits are
// expanded into 1 KB of shared memory, sorted by the warp (bitonic network over the full 32-bit reference indices) and
// the pair tests / region heads of computeMap.hpp:320-347 run on ballots, with dozens of fragments in flight per SM.
constexpr int L1_TINY = 256, L1_TINY_WARPS = 4;
__global__ void __launch_bounds__(32 * L1_TINY_WARPS, 8)
l1_tiny_kernel(const uint64_t *seq_first, const int32_t *qs, const uint32_t *hit_start, const uint32_t *hit_cnt, const uint64_t *seed_base,
               const uint32_t *pos_idx, const uint32_t *gpos, const uint32_t *irr, const int32_t *min_hits, int n_frags, int frag_len,
               uint32_t d_near, uint32_t tiny_cap, Cand *tmp, uint32_t *frag_cands)
{
    __shared__ uint32_t s_k[L1_TINY_WARPS][L1_TINY];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = blockIdx.x * L1_TINY_WARPS + wid;
    if (f >= n_frags) return;
    const uint64_t sb = seed_base[f], nf64 = seed_base[f + 1] - sb;
    if (nf64 > (uint64_t)tiny_cap) return;                                   // another shape takes this fragment
    const int s = qs[f], n = (int)nf64;
    if (s <= 0 || n == 0) { if (lane == 0) frag_cands[f] = 0; return; }
    const uint64_t qb = seq_first[f];
    uint32_t *K = s_k[wid];
    // 1. the position lists, one after the other
    int base = 0;
    for (int q0 = 0; q0 < s; q0 += 32) {
        const int q = q0 + lane;
        const uint32_t c = q < s ? hit_cnt[qb + q] : 0u, st = q < s ? hit_start[qb + q] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        const int off = base + (int)(incl - c);
        for (uint32_t j = 0; j < c; j++) K[off + (int)j] = __ldg(pos_idx + st + j);
        base += (int)__shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    int p2 = 32;
    while (p2 < n) p2 <<= 1;
    for (int i = n + lane; i < p2; i += 32) K[i] = 0xFFFFFFFFu;
    __syncwarp();
    // 2. bitonic sort of K[0, p2)
    for (int k = 2; k <= p2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (p2 >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;      // the t-th comparator of this step
                const uint32_t a = K[i], b = K[l];
                if ((a > b) == ((i & k) == 0)) { K[i] = b; K[l] = a; }
            }
            __syncwarp();
        }
    // 3. valid pairs, region heads and ends
    const int m = min_hits[s];
    const uint32_t L = (uint32_t)frag_len;
    auto marked = [&](uint32_t j) -> uint32_t { return (__ldg(irr + (j >> 15)) >> ((j >> 10) & 31u)) & 1u; };
    auto near = [&](uint32_t a, uint32_t b2) -> bool {                        // l1_near of l1_fused_kernel
        const uint32_t d = b2 - a;
        if (d >= L) return false;
        if (d <= d_near && (marked(a) | marked(b2)) == 0u) return true;
        return __ldg(gpos + b2) - __ldg(gpos + a) < L;
    };
    Cand *out = tmp + sb;
    uint32_t heads = 0, carry = 0;
    bool have = false;                                                       // a valid pair seen so far (its first seed: carry)
    for (int t0 = 0; t0 + m - 1 < n; t0 += 32) {
        const int t = t0 + lane;
        const bool in = t + m - 1 < n;
        const uint32_t ja = in ? K[t] : 0u, jb = in ? K[t + m - 1] : 0u;
        const bool valid = in && near(ja, jb);
        const unsigned vm = __ballot_sync(0xFFFFFFFFu, valid);
        // the valid pair before this one: in this group of 32, or the carry
        const unsigned below = vm & ((1u << lane) - 1u);
        const int pl = below ? 31 - __clz(below) : -1;
        const uint32_t pj_in = __shfl_sync(0xFFFFFFFFu, ja, pl < 0 ? 0 : pl);
        const bool pv = pl >= 0 || have;
        const uint32_t pj = pl >= 0 ? pj_in : carry;
        const bool head = valid && (!pv || !near(pj, jb));
        const unsigned hm = __ballot_sync(0xFFFFFFFFu, head);
        if (head) {
            const uint32_t slot = heads + __popc(hm & ((1u << lane) - 1u));
            Cand *o = out + slot;
            o->frag = f; o->hint = jb; o->spare = 0u;
            if (pv) out[slot - 1].tail = pj;
        }
        heads += __popc(hm);
        if (vm) { carry = __shfl_sync(0xFFFFFFFFu, ja, 31 - __clz(vm)); have = true; }
        __syncwarp();
    }
    if (lane == 0) {
        if (heads) out[heads - 1].tail = carry;
        frag_cands[f] = heads;
    }
}

// scratch ranges of l1_fused_kernel -> the candidate array in fragment order
// (a fragment mapped in parts left one run of regions per part, each at the seed offset of its part)
__global__ void compact_cands_kernel(const Cand *tmp, const uint64_t *seed_base, const uint32_t *cand_base, uint32_t seed_cap, Cand *cands,
                                     uint32_t parts_lo, int n_parts, const uint32_t *part_off, const uint32_t *part_cands)
{
    const int f = blockIdx.x;
    const uint64_t sb = seed_base[f], nf = seed_base[f + 1] - sb;
    if (nf > (uint64_t)seed_cap) return;
    const uint32_t c0 = cand_base[f], n = cand_base[f + 1] - c0;
    uint4 *dst = reinterpret_cast<uint4 *>(cands + c0);
    if (n_parts > 0 && nf >= (uint64_t)parts_lo) {
        // (when a part overflowed the large shape mapped the fragment: its regions are "part 0", at the fragment's own offset)
        uint32_t done = 0;
        for (int p = 0; p < n_parts && done < n; p++) {
            const uint32_t np = part_cands[(size_t)f * n_parts + p];
            const uint4 *src = reinterpret_cast<const uint4 *>(tmp + sb + part_off[(size_t)f * (n_parts + 1) + p]);
            for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) dst[done + i] = src[i];
            done += np;
        }
        return;
    }
    const uint4 *src = reinterpret_cast<const uint4 *>(tmp + sb);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

__global__ void work_items_kernel(const uint32_t *frag_cands_counts, int n_frags, uint32_t *work)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > n_frags) return;
    work[f] = f == n_frags ? 0u : (frag_cands_counts[f] + L2_ITEM - 1) / L2_ITEM;
}

// ---- L2: sliding super-window Jaccard -------------------------------------------------------
// State per candidate (SURVEY.md A.5 in incremental form).  Q = sorted query sketch q_1 < ... < q_s.
// A reference hash that is in Q toggles a match bit M(i); one that is not falls in bucket
// b = #{q < h} and toggles a distinct-hash count cnt[b].  With
//   istar = number of query hashes inside the bottom-s of (Q u window),
//   sigma = number of window-only hashes of bucket `istar` inside it,
// (istar + sum_{b<istar} cnt[b] + sigma == s), every insert/delete moves (istar, sigma) by one
// step and `shared` = #{i <= istar : M(i)} is updated in O(1) -- the std::map pivot walk of
// slidingMap.hpp:137-284 without the tree.  Duplicate hashes inside a window are resolved with
// the distances stored at index time (RefMini.w), mirroring the wposR bookkeeping of
// slidingMap.hpp:150-155, 178-205.
//
// The work is split so that only the inherently sequential part runs one lane per candidate:
//   l2_prep_kernel    per candidate: the index searches of computeMap.hpp:421-433, event counts
//   l2_events_kernel  per candidate (one warp): classify every reference minimizer of the region
//                     against the query sketch, merge the insert and delete streams by event time
//                     (merge path), resolve duplicates, mark the ends of the time groups
//                     -> a list of 16-bit events in HBM
//   l2_slide_kernel   one lane per candidate replays its event list through the state machine
//   l2_fallback_kernel  exact, slow variant for the candidates the fast path does not take

__device__ __forceinline__ uint32_t lb_hw(const uint2 *hw, uint32_t lo, uint32_t hi, int target)
{
    while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if ((int)(hw[mid].y & 0x7FFFFFFFu) < target) lo = mid + 1; else hi = mid; }
    return lo;
}

// lower bound near a guess: gallop from the guess, then bisect the bracket -- the probes stay within a few
// sectors of the answer instead of walking a wide range (the searches below are DRAM-sector bound)
__device__ __forceinline__ uint32_t lb_near(const uint2 *hw, uint32_t lo, uint32_t hi, int target, long long guess)
{
    if (lo >= hi) return lo;
    const uint32_t g = (uint32_t)min(max(guess, (long long)lo), (long long)hi - 1);
    uint32_t l, r, step = 4;
    if ((int)(hw[g].y & 0x7FFFFFFFu) < target) {                      // answer in (g, hi]
        l = g + 1;
        r = min(hi, g + step);
        while (r < hi && (int)(hw[r].y & 0x7FFFFFFFu) < target) { l = r + 1; step *= 2; r = (hi - r > step) ? r + step : hi; }
    } else {                                                          // answer in [lo, g]
        r = g;
        l = (g - lo > step) ? g - step : lo;
        while (l > lo && (int)(hw[l].y & 0x7FFFFFFFu) >= target) { r = l; step *= 2; l = (l - lo > step) ? l - step : lo; }
    }
    return lb_hw(hw, l, r, target);
}

// Events: bits 0-1 and 8-14 hold the state byte the event touches, already in the lane-
// interleaved layout of the slide kernel (four one-byte buckets per word, words of one lane 256
// bytes apart): offset = (idx >> 2) << 8 | (idx & 3) with idx = #{q < h} + match.
constexpr uint32_t EV_MATCH = 4;      // the hash is in the query sketch: toggles M(idx)
constexpr uint32_t EV_ONLY = 8;       // the hash is not in the sketch: counts in bucket idx (neither bit: no state change)
constexpr uint32_t EV_DEL = 16;       // delete (window begin advances) / insert
constexpr uint32_t EV_GRP = 32;       // last event of its time group: evaluate the window after it
constexpr uint32_t EV_AOFF = 0x7F03;
constexpr int EV_MAX_S = 508;         // largest sketch the 16-bit events address
constexpr int EV_RMAX = 1024;         // most reference minimizers of a candidate region on the event path
constexpr int EVK_THREADS = 128;
constexpr int EV_UNROLL = 4;          // elements per lane and trip of the classification loop (the loop body has to stay in the instruction cache)
constexpr int EV_LIST_BYTES = (2 * EV_RMAX + 16) * 2;   // per-warp staging of one event list
constexpr int EV_HIST_BYTES = 512;                      // per-warp byte state of the start window (EV_MAX_S + 4 buckets)
constexpr int EV_WARP_BYTES = EV_LIST_BYTES + EV_HIST_BYTES;
// 16-bit units of the state block that precedes a candidate's event list: one byte per bucket, padded to 16 bytes
__host__ __device__ inline int ev_state_u16(int s) { return ((((s + 4) >> 2) * 4 + 15) & ~15) >> 1; }

__host__ __device__ inline uint32_t ev_aoff(int idx) { return ((uint32_t)(idx & ~3) << 6) | (uint32_t)(idx & 3); }

// The three Sketch::searchIndex calls of computeL2MappedRegions (computeMap.hpp:421-433) for every
// candidate, restricted to the candidate's contig.  Minimizer positions grow strictly inside a
// contig (one minimizer per window at most), so an index distance never exceeds the position
// distance and each search runs over a short range.
struct Prep {
    uint32_t beg, last;         // first reference minimizer of the region; the slide stops when the window end reaches `last`
    int32_t  seq;               // refSeqId
    uint32_t n_del;             // delete events before the slide stops (insert events: last - 1 - beg) | elements of the first window << 16
};

__global__ void __launch_bounds__(256)
l2_prep_kernel(const Cand *cands, const uint32_t *cand_base, int n_frags, const int32_t *qs, const RefMini *ref, const uint2 *hw,
               const uint2 *hl, const uint32_t *fb, const uint32_t *contig_off, int frag_len, int cmw, int dens_num, int dens_den, Prep *prep,
               uint32_t *mid, unsigned long long *ev_cnt, Mapping *maps, unsigned long long *counters)
{
    const uint32_t n = cand_base[n_frags];
    unsigned long long scanned = 0;
    unsigned int redo = 0;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c <= n; c += gridDim.x * blockDim.x) {
        if (c == n) { ev_cnt[c] = 0; break; }
        const Cand cd = cands[c];
        // the region [max(0, wpos[hint] - L + 1), wpos[tail]] as an index range: the first element at or past its start
        // and the first one at or past wpos[tail] + L (computeMap.hpp:421-433), both known to the index
        const uint32_t fb_h = fb[cd.hint], fb_t = fb[cd.tail];
        const int seq = (int)ref[cd.hint].z;
        const uint32_t c0 = contig_off[seq], c1 = contig_off[seq + 1];
        const uint32_t beg = cd.hint - (fb_h & 0xFFFFu);
        const int wpos_beg = (int)(hw[beg].y & 0x7FFFFFFFu);
        // first index with wpos >= wpos[beg] + cmw: the index knows it for the element before beg (fa_index.cu
        // slide_order_kernel: first index at or past wpos[j + 1] + cmw - 1, and whether it sits exactly there)
        uint32_t end0;
        const uint32_t ow_b = (beg > c0 && cmw >= 2) ? hl[beg - 1].y : 0xFFFFFFFFu;
        if (((ow_b >> 16) & 0x7FFFu) != 0x7FFFu) end0 = min(beg - 1 + ((ow_b >> 16) & 0x7FFFu) + ((ow_b >> 15) & 1u), c1);
        else end0 = lb_near(hw, beg, c1, wpos_beg + cmw, (long long)beg + (long long)cmw * dens_num / dens_den);
        const uint32_t last = max(end0, cd.tail + (fb_t >> 16));
        scanned += max(end0, last) - beg;
        Prep pp{beg, last, seq, 0u};
        unsigned long long cnt = 0;
        Mapping mp{seq, 0, -1, 0.0f};
        if (end0 < last) {
            // the slide stops before the insert of element last - 1; deletes at or after that time are not reached
            const int t_stop = (int)(hw[last - 1].y & 0x7FFFFFFFu) - cmw + 1;
            // first index in [beg + 1, last) with wpos >= t_stop: the index knows the first one past t_stop
            uint32_t dstop;
            const uint32_t ow_l = hl[last - 1].y;
            if ((ow_l & 0x7FFFu) != 0x7FFFu && cmw >= 2) {
                const uint32_t a = last - 1 - (ow_l & 0x7FFFu);                     // first index of the contig with wpos > t_stop
                dstop = (a > c0 && (int)(hw[a - 1].y & 0x7FFFFFFFu) == t_stop) ? a - 1 : a;
                dstop = min(max(dstop, beg + 1), last);
            } else dstop = lb_near(hw, beg + 1, last, t_stop, (long long)last - 1 - (long long)cmw * dens_num / dens_den);
            pp.n_del = dstop - 1 - beg;
            const bool fast = last - beg <= (uint32_t)EV_RMAX && qs[cd.frag] <= EV_MAX_S && cmw >= 2;
            if (fast) {
                // The slide starts in the MIDDLE of the region -- at the window centred on the seeds that define it -- and
                // runs to the right, then to the left (l2_slide_kernel).  d_m = begin index of that window relative to beg
                // (= deletes before it), e_m = its end index (= inserts before it); the state after delete d_m - 1 (and the
                // insert that shares its time, if any) is an evaluation point of the forward order.
                const int nI = (int)(last - 1 - beg), nD = (int)pp.n_del, y0 = (int)(end0 - beg);
                const int ctr = (int)(((long long)cd.hint + (long long)cd.tail) / 2 - (long long)beg);
                const int dm = min(max(ctr - y0 / 2, 0), nD);
                int em = y0;
                if (dm > 0) {
                    const uint32_t ow = hl[beg + (uint32_t)dm - 1u].y;
                    const int lead = (int)((ow >> 16) & 0x7FFFu);
                    em = min(dm - 1 + lead, nI);
                    em += (((ow >> 15) & 1u) && dm - 1 + lead < nI) ? 1 : 0;
                }
                const int pad0 = (8 - ((dm + em - y0) & 7)) & 7;                // the start state sits on a 16-byte boundary
                cnt = (unsigned long long)(ev_state_u16(qs[cd.frag]) + ((pad0 + nI + nD - y0 + 7) & ~7));   // (the first window is never replayed)
                mid[c] = (uint32_t)dm | ((uint32_t)em << 16);
                pp.n_del |= (end0 - beg) << 16;
            }
            else { mp.ref_start = L2_REDO; redo++; }
        }
        prep[c] = pp;
        ev_cnt[c] = cnt;
        maps[c] = mp;                         // final for candidates without a window; overwritten by the slide otherwise
    }
    for (int o = 16; o > 0; o >>= 1) scanned += __shfl_xor_sync(0xFFFFFFFFu, scanned, o);
    redo = __reduce_add_sync(0xFFFFFFFFu, redo);
    if ((threadIdx.x & 31) == 0) {
        if (scanned) atomicAdd(&counters[CT_SCANNED], scanned);
        if (redo) atomicAdd(&counters[CT_REDO], (unsigned long long)redo);
    }
}

// Minimizer hashes are window minima, so they crowd towards zero (density ~ (1 - u)^(2w - 1)): a
// table over the plain top bits would put most of a sketch into a few slots.  The slot function is
// piecewise linear instead: the hash range is cut into equal pieces of 2^p (about the half-life of
// that density) and piece k gets L2_TAB / 2 >> k slots, which spreads a sketch about evenly.  Monotone.
// (4096 slots for ~240 hashes: nine sketches in ten have at most two hashes in any slot.)
__host__ __device__ inline int l2_tab_shift(int w)
{
    const double half = 4294967296.0 * 0.6931471805599453 / (2.0 * (w < 1 ? 1 : w));
    int p = L2_TAB_BITS - 1;
    while (p < 30 && (double)(1u << p) < half) p++;
    return p;
}
__device__ __forceinline__ uint32_t l2_slot(uint32_t h, int p)
{
    const uint32_t k = min(h >> p, (uint32_t)L2_TAB_BITS);
    return (uint32_t)L2_TAB - ((uint32_t)L2_TAB >> k) + ((h & ((1u << p) - 1u)) >> (p - (L2_TAB_BITS - 1) + k));
}

// Per candidate descriptor for the slide kernel (32 bytes).
struct SlideJob {
    unsigned long long ev_off;  // the candidate's state block in the event buffer (16-bit units); its event list follows
    uint32_t n_list;            // events of the list, leading padding included, PLUS ONE; 0 = nothing to slide (no window)
    uint32_t k_mid;             // list position of the start state (a multiple of 8)
    uint16_t s, d_m;            // sketch size; begin index of the start window relative to the region
    uint16_t room_r, room_l;    // elements in the sketch at or right of the start window's begin / left of its end
    uint16_t istar, sigma, shared, flags;   // pivot and shared count of the start window; flags bit 0: a bucket count overflowed
};

// ---- events --------------------------------------------------------------------------------
// One CTA per work item (<= L2_ITEM candidates of one fragment): the fragment's sketch and the
// classification table are staged once, then each warp takes candidates one by one.
//   classification: slot table (above) + a few compares against the staged sketch
//   placement: the order of the inserts and deletes of a slide does not depend on the query, only on
//          the minimizer positions, so the index carries per element how many elements lie one window
//          behind / ahead (fa_index.cu slide_order_kernel).  Insert i follows the deletes of the
//          elements that left the window before it (i - lag), delete i the inserts of those that
//          entered before it (i + lead), deletes first inside a time group (MIIteratorL2.hpp:74-96):
//          every element is classified and both of its events go straight to their place in a staged
//          list -- no merge, no position array; the list leaves in whole 16-byte chunks.
//   start state: the slide starts at the window in the middle of the region (l2_prep_kernel) and replays the
//          list from there in both directions, so the bucket bytes of that window are built here, in
//          parallel (one shared-memory atomic per element of the window), together with its pivot and shared
//          count (ev_pivot); they travel in front of the list.  The events of the first window are never
//          replayed and are not written.
struct EvCtx {
    const uint32_t *s_q; const uint16_t *s_tab; uint16_t *w_ev; uint32_t *w_hist;
    const RefMini *ref; const uint2 *hl; uint16_t *ev; SlideJob *jobs;
    int s, tab_p, maxn;
};

// Pivot (istar, sigma) and shared count of the window whose bucket bytes (count << 1 | match bit) are in `hist`:
// T(x) = x + sum of the counts of the buckets below x grows strictly with x; istar is the largest x with T(x) <= s,
// sigma = s - T(istar) (<= the count of bucket istar), shared = match bits of the buckets 1..istar.  One warp, every
// lane owns 16 buckets.
__device__ __forceinline__ void ev_pivot(const uint32_t *hist, int s, int lane, int &istar, int &sigma, int &shared)
{
    constexpr unsigned full = 0xFFFFFFFFu;
    const uint4 v = reinterpret_cast<const uint4 *>(hist)[lane];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int csum = 0, msum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        csum += (int)__dp4a((w[j] >> 1) & 0x7F7F7F7Fu, 0x01010101u, 0u);
        msum += __popc(w[j] & 0x01010101u);
    }
    int cpre = csum, mpre = msum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(full, cpre, o), b = __shfl_up_sync(full, mpre, o);
        if (lane >= o) { cpre += a; mpre += b; }
    }
    cpre -= csum; mpre -= msum;
    int t = 16 * lane + cpre, m = mpre;                         // T(16 lane), match bits below bucket 16 lane
    const unsigned ok = __ballot_sync(full, t <= s);            // (lane 0: T(0) = 0)
    const int src = 31 - __clz(ok);
    int bi = 16 * lane, bt = t, bm = m;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int byte = (int)((w[j >> 2] >> (8 * (j & 3))) & 0xFFu);
        if (t <= s) { bi = 16 * lane + j; bt = t; bm = m + (byte & 1); }
        t += 1 + (byte >> 1); m += byte & 1;
    }
    istar = __shfl_sync(full, bi, src);
    sigma = s - __shfl_sync(full, bt, src);
    shared = __shfl_sync(full, bm, src);
}

// The tail of a candidate: pivot of the start window, state block and event list to global memory, descriptor.
__device__ __forceinline__ void ev_finish(const EvCtx &X, uint32_t c, unsigned long long off, int n_list, int kmid_list, int dm,
                                       int room_r, int room_l, int ovf, int lane)
{
    uint16_t *w_ev = X.w_ev;
    const int s = X.s, sb = ev_state_u16(s);
    if (lane < 8) w_ev[n_list + lane] = (uint16_t)0;        // the tail of the last chunk is padding: no-ops
    __syncwarp();
    int istar, sigma, shared;
    ev_pivot(X.w_hist, s, lane, istar, sigma, shared);
    if (lane == 0) {
        SlideJob jb;
        jb.ev_off = off; jb.n_list = (uint32_t)n_list + 1u; jb.k_mid = (uint32_t)kmid_list;
        jb.s = (uint16_t)s; jb.d_m = (uint16_t)dm; jb.room_r = (uint16_t)room_r; jb.room_l = (uint16_t)room_l;
        jb.istar = (uint16_t)istar; jb.sigma = (uint16_t)sigma; jb.shared = (uint16_t)shared; jb.flags = (uint16_t)(ovf ? 1 : 0);
        const uint4 *src = reinterpret_cast<const uint4 *>(&jb);
        uint4 *dst = reinterpret_cast<uint4 *>(X.jobs + c);
        dst[0] = src[0]; dst[1] = src[1];
    }
    // state block: the bucket bytes, 16 per lane (the rest of the histogram is zero)
    if (lane < (sb >> 3)) reinterpret_cast<uint4 *>(X.ev + off)[lane] = reinterpret_cast<const uint4 *>(X.w_hist)[lane];
    // the list, in 16-byte chunks
    {
        const uint4 *srcv = reinterpret_cast<const uint4 *>(w_ev);
        uint4 *dst = reinterpret_cast<uint4 *>(X.ev + off + sb);
        const int n_chunks = (n_list + 7) >> 3;
        for (int t = lane; t < n_chunks; t += 32) dst[t] = srcv[t];
    }
}

// MAXN > 0: every table slot holds at most MAXN sketch hashes; MAXN == 0: bisect inside the slot.
// The classification loop of one candidate (one warp): events into the staged list, the start window into the histogram.
// Returns (packed) the elements in the sketch at or right of the start window's begin / left of its end, and whether a
// bucket count of the start window overflowed.
template <int MAXN>
__device__ __forceinline__ uint32_t ev_candidate(const EvCtx &X, const Prep &pp, int dm, int em, int shift, int lane)
{
    const uint32_t *s_q = X.s_q; const uint16_t *s_tab = X.s_tab; uint16_t *w_ev = X.w_ev; uint32_t *w_hist = X.w_hist;
    const RefMini *ref = X.ref; const uint2 *hl = X.hl;
    const int s = X.s, tab_p = X.tab_p;
    const int R = (int)(pp.last - pp.beg), nI = R - 1, nD = (int)(pp.n_del & 0xFFFFu), y0 = (int)(pp.n_del >> 16);
    int n_r = 0, n_l = 0;                              // elements in the sketch right / left of the start window's edges
    uint32_t ovf = 0;                                  // OR of the bucket words the adds found
    // (hash, order word) of EV_UNROLL elements per lane and trip, the next trip's loads in flight while this one is classified
    uint2 nx[EV_UNROLL];
#pragma unroll
    for (int u = 0; u < EV_UNROLL; u++) {
        const int i = u * 32 + lane;
        nx[u] = i < R ? __ldg(hl + pp.beg + i) : make_uint2(0u, 0u);
    }
#pragma unroll 1
    for (int i0 = 0; i0 < R; i0 += 32 * EV_UNROLL) {
        uint2 xs[EV_UNROLL];
#pragma unroll
        for (int u = 0; u < EV_UNROLL; u++) {
            xs[u] = nx[u];
            const int i = i0 + (EV_UNROLL + u) * 32 + lane;
            if (i < R) nx[u] = __ldg(hl + pp.beg + i);
        }
#pragma unroll
        for (int u = 0; u < EV_UNROLL; u++) {
            const int i = i0 + u * 32 + lane;
            if (i < R) {
                const uint32_t h = xs[u].x;
                const uint32_t slot = l2_slot(h, tab_p);
                int l = (int)s_tab[slot], match;
                if (MAXN > 0) {
                    int lt = 0, eq = 0;
#pragma unroll
                    for (int q = 0; q < MAXN; q++) { const uint32_t qv = s_q[l + q]; lt += qv < h ? 1 : 0; eq |= qv == h ? 1 : 0; }
                    l += lt;
                    match = l < s ? eq : 0;
                } else {
                    int r = (int)s_tab[slot + 1];
                    while (l < r) { const int mid2 = (l + r) >> 1; if (s_q[mid2] < h) l = mid2 + 1; else r = mid2; }
                    match = (l < s && s_q[l] == h) ? 1 : 0;
                }
                n_r += (match && i >= dm) ? 1 : 0;
                n_l += (match && i < em) ? 1 : 0;
                const int idx = l + match;
                const uint32_t code = ev_aoff(idx) | (match ? EV_MATCH : EV_ONLY);
                const uint32_t ow = xs[u].y;
                uint32_t dd = 0u;
                if (__builtin_expect((ow >> 31) != 0u, 0)) dd = ref[pp.beg + (uint32_t)i].w;   // a same-hash neighbour exists (rare)
                const uint32_t dp = dd & 0xFFFFu;
                // start window [dm, em): one count per distinct hash
                if (i >= dm && i < em && !(dp && i - (int)dp >= dm)) {
                    // (a byte cannot wrap before an add has seen its top bit set: counts from 64 on go to the exact kernel)
                    ovf |= atomicAdd(&w_hist[idx >> 2], (match ? 1u : 2u) << (8u * (uint32_t)(idx & 3)));
                }
                // insert i: after the deletes of the region's elements that left before it entered
                if (i >= y0 && i < nI) {
                    const int x = min(max(i - (int)(ow & 0x7FFFu) - 1, 0), nD);
                    const bool skip = dp && i - (int)dp >= x;              // already present (REV)
                    // insert times are distinct: every insert after the first window ends its time group
                    w_ev[i + x + shift] = (uint16_t)((skip ? 0u : code) | EV_GRP);
                }
                // delete i: after the inserts of the elements that entered before it leaves
                if (i < nD) {
                    const int lead = (int)((ow >> 16) & 0x7FFFu);
                    const int y = min(i + lead, nI);
                    const bool twin = ((ow >> 15) & 1u) && i + lead < nI;      // an insert of the same time follows: same group
                    const uint32_t dn = dd >> 16;
                    const bool skip = dn && i + (int)dn < y;               // a later copy stays (NOOP)
                    w_ev[i + y + shift] = (uint16_t)((skip ? 0u : code) | EV_DEL | (twin ? 0u : EV_GRP));
                }
            }
        }
    }
    // (upper bounds of the inserts that set a match bit -- same-hash copies and the last element are counted, too --
    // are all the early stops of the slide need)
    return (uint32_t)n_r | ((uint32_t)n_l << 12) | ((ovf & 0x80808080u) ? 1u << 24 : 0u);     // (at most EV_RMAX = 1024 elements per lane sum)
}

__global__ void __launch_bounds__(EVK_THREADS, 8)
l2_events_kernel(const Prep *prep, const uint32_t *mid, const unsigned long long *ev_off, const uint32_t *cand_base, const uint32_t *work_base,
                 int n_frags, const uint32_t *qhash, const uint64_t *seq_first, const int32_t *qs,
                 const RefMini *ref, const uint2 *hl, int tab_p, uint16_t *ev, SlideJob *jobs,
                 unsigned long long *counters, int q_cap)
{
    extern __shared__ __align__(16) uint8_t ev_smem[];
    uint32_t *s_q = reinterpret_cast<uint32_t *>(ev_smem);                  // q_cap entries: sketch + sentinels
    uint16_t *s_tab = reinterpret_cast<uint16_t *>(s_q + q_cap);            // L2_TAB + 2 entries
    uint8_t *s_warp = reinterpret_cast<uint8_t *>(s_tab + L2_TAB + 2) + 12;   // 16-byte aligned (q_cap is a multiple of 4)
    __shared__ uint32_t s_item, s_next;
    __shared__ int s_maxn, s_cached_f;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint16_t *w_ev = reinterpret_cast<uint16_t *>(s_warp + wid * EV_WARP_BYTES);  // the event list of the warp's candidate
    uint32_t *w_hist = reinterpret_cast<uint32_t *>(s_warp + wid * EV_WARP_BYTES + EV_LIST_BYTES);   // bucket bytes of its start window
    const uint32_t n_work = work_base[n_frags];
    if (tid == 0) s_cached_f = -1;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (uint32_t)atomicAdd(&counters[CT_WORK3], 1ull);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_work) break;
        int flo = 0, fhi = n_frags - 1;
        while (flo < fhi) { int mid2 = (flo + fhi + 1) >> 1; if (work_base[mid2] <= item) flo = mid2; else fhi = mid2 - 1; }
        const int f = flo;
        const int s = qs[f];
        const uint32_t c_lo = cand_base[f] + (item - work_base[f]) * L2_ITEM;
        const uint32_t c_hi = min(c_lo + (uint32_t)L2_ITEM, cand_base[f + 1]);
        if (s_cached_f != f && s <= EV_MAX_S) {
            const uint64_t qb = seq_first[f];
            for (int i = tid; i < s + L2_QPAD; i += EVK_THREADS) s_q[i] = i < s ? qhash[qb + i] : 0xFFFFFFFFu;
            if (tid == 0) s_maxn = 0;
            __syncthreads();
            // tab[x] = first sketch index whose slot is >= x (tab[L2_TAB] = s)
            for (int i = tid; i <= s; i += EVK_THREADS) {
                const uint32_t lo = i == 0 ? 0u : l2_slot(s_q[i - 1], tab_p) + 1u;
                const uint32_t hi = i == s ? (uint32_t)L2_TAB : l2_slot(s_q[i], tab_p);
                for (uint32_t x = lo; x <= hi; x++) s_tab[x] = (uint16_t)i;
            }
            __syncthreads();
            int ml = 0;
            for (int x = tid; x < L2_TAB; x += EVK_THREADS) ml = max(ml, (int)s_tab[x + 1] - (int)s_tab[x]);
            atomicMax(&s_maxn, ml);
        }
        if (tid == 0) { s_next = c_lo + EVK_THREADS / 32; s_cached_f = s <= EV_MAX_S ? f : -1; }
        __syncthreads();
        const int maxn = s_maxn;
        EvCtx X;
        X.s_q = s_q; X.s_tab = s_tab; X.w_ev = w_ev; X.w_hist = w_hist; X.ref = ref; X.hl = hl; X.ev = ev; X.jobs = jobs;
        X.s = s; X.tab_p = tab_p; X.maxn = maxn;

        // candidates are handed out one ahead, so the next descriptor is in flight while this one is processed
        uint32_t c = c_lo + (uint32_t)wid;
        Prep pp{};
        uint32_t md = 0;
        unsigned long long off = 0, off1 = 0;
        if (c < c_hi) { pp = prep[c]; md = mid[c]; off = ev_off[c]; off1 = ev_off[c + 1]; }
        while (c < c_hi) {
            uint32_t c_nx = 0;
            if (lane == 0) c_nx = atomicAdd(&s_next, 1u);
            c_nx = __shfl_sync(0xFFFFFFFFu, c_nx, 0);
            Prep pp_nx{};
            uint32_t md_nx = 0;
            unsigned long long off_nx = 0, off1_nx = 0;
            if (c_nx < c_hi) { pp_nx = prep[c_nx]; md_nx = mid[c_nx]; off_nx = ev_off[c_nx]; off1_nx = ev_off[c_nx + 1]; }

            if (off1 == off) {                              // nothing to slide (no window, or left to the exact kernel)
                if (lane == 0) { uint4 *dst = reinterpret_cast<uint4 *>(jobs + c); dst[0] = make_uint4(0, 0, 0, 0); dst[1] = make_uint4(0, 0, 0, 0); }
            } else {
                const int nI = (int)(pp.last - pp.beg) - 1, nD = (int)(pp.n_del & 0xFFFFu), y0 = (int)(pp.n_del >> 16);
                const int dm = (int)(md & 0xFFFFu), em = (int)(md >> 16);
                const int pad0 = (8 - ((dm + em - y0) & 7)) & 7;
                __syncwarp();                                  // (the copy-out of the previous list is done)
                reinterpret_cast<uint4 *>(w_hist)[lane] = make_uint4(0, 0, 0, 0);
                if (lane < pad0) w_ev[lane] = (uint16_t)0;
                __syncwarp();
                uint32_t r;
                switch (maxn <= 2 ? 2 : (maxn <= 4 ? 4 : (maxn <= L2_QPAD - 1 ? L2_QPAD - 1 : 0))) {
                case 2: r = ev_candidate<2>(X, pp, dm, em, pad0 - y0, lane); break;
                case 4: r = ev_candidate<4>(X, pp, dm, em, pad0 - y0, lane); break;
                case L2_QPAD - 1: r = ev_candidate<L2_QPAD - 1>(X, pp, dm, em, pad0 - y0, lane); break;
                default: r = ev_candidate<0>(X, pp, dm, em, pad0 - y0, lane); break;
                }
                const int n_r = (int)__reduce_add_sync(0xFFFFFFFFu, r & 0xFFFu), n_l = (int)__reduce_add_sync(0xFFFFFFFFu, (r >> 12) & 0xFFFu);
                const int ovf = __any_sync(0xFFFFFFFFu, (r >> 24) != 0u) ? 1 : 0;
                ev_finish(X, c, off, pad0 + nI + nD - y0, dm + em - y0 + pad0, dm, n_r, n_l, ovf, lane);
            }
            c = c_nx; pp = pp_nx; md = md_nx; off = off_nx; off1 = off1_nx;
        }
    }
}

// ---- slide ---------------------------------------------------------------------------------
// One lane per candidate, eight events (one 16-byte chunk) per loop trip, the next chunk in
// flight.  Every event is applied with selects only, so lanes do not diverge whatever mix of
// inserts / deletes / matches they replay -- in either direction: the state is a function of the
// window content, so an event is undone by applying its opposite.  A lane starts from the state of
// the window in the middle of its region (built by the events kernel), replays the list to the
// right until no later window can reach the best one, takes the start state again and replays the
// list to the left until the same holds there.  State is one byte per bucket (7-bit count + match
// bit), four buckets per 32-bit word, words interleaved by lane so that every lane owns a
// shared-memory bank.  A count that would pass 127 sends the candidate to the fallback.
//   Nothing a lane waits for is fetched by that lane alone while the other 31 idle:
//   * candidates are claimed one ahead (one atomic per warp and trip), so the descriptor of the next
//     candidate is in registers when the current one ends;
//   * the start state (one 16-byte piece per lane of the warp, a single coalesced request) is put
//     into the owner's column by the whole warp -- when a candidate starts and when it turns round;
//   * the two position gathers of a finished candidate are consumed one trip after they were issued.
struct SlideLane {
    int sigma, shared, best, nb;
    int pa, pb;                // begin index of the optimum: pa is set when it improves, pb also when it is equalled
                               // (forward: first / last position of computeMap.hpp:467-481, backward: last / first)
    uint32_t a;                // cached state byte of bucket istar
    uint32_t ioff;             // ev_aoff(istar)
    uint32_t mx;               // largest state byte written (> 255: a bucket count overflowed)
    int room;                  // elements in the sketch still inside or ahead of the window: no later window shares more
};

// shared-memory bytes by 32-bit shared address (keeps the window base out of the per-event code)
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v)); }

// window evaluation at the end of a time group (computeMap.hpp:467-481), by begin index
__device__ __forceinline__ void slide_eval(SlideLane &L, bool grp)
{
    const bool gt = grp && L.shared > L.best, ge = grp && L.shared >= L.best;
    L.best = gt ? L.shared : L.best;
    L.pa = gt ? L.nb : L.pa;
    L.pb = ge ? L.nb : L.pb;
}

// state byte of a bucket: count << 1 | match bit, so an event adds +-(EV_MATCH ? 1 : EV_ONLY ? 2 : 0).
// The event sits in bits SH..SH+15 of w (two events per word; the flags are tested in place).  xdel2 = EV_DEL in both
// halves undoes the events (backward replay: a delete puts its element back); gw = the word whose EV_GRP flag (same
// half) asks for an evaluation after this event (forward: the event itself; backward: the one undone next, which ends
// its time group with the state this event leaves); dir = +1 / -1, the way the begin index moves.
template <int SH>
__device__ __forceinline__ void slide_event(SlideLane &L, uint32_t st, uint32_t w, uint32_t xdel2, uint32_t gw, int dir)
{
    const uint32_t aoff = (w >> SH) & EV_AOFF;
    const uint32_t pa = st + aoff;
    const uint32_t v = lds_u8(pa);
    const bool del = ((w ^ xdel2) & (EV_DEL << SH)) != 0;                        // the element leaves the window
    const int sgn = del ? -1 : 1;
    const uint32_t noff = (L.ioff + (del ? 0xFDu : 0xFFFFFFFFu)) & ~0xFCu;       // bucket istar + 1 / istar - 1
    const uint32_t pn = lds_u8(st + noff);                                      // (slack rows on both sides)
    const uint32_t v2 = v + (uint32_t)(sgn * (int)((w >> (SH + 2)) & 3u));
    sts_u8(pa, v2);
    L.mx = max(L.mx, v2);
    const bool below = aoff < L.ioff, at_p = aoff == L.ioff;
    L.a = at_p ? v2 : L.a;
    const int cc = (int)(L.a >> 1);
    const bool on = (w & (EV_ONLY << SH)) != 0;
    const bool in_ins = on && !del && below;
    const bool in_del = on && del && (below || (at_p && L.sigma > cc));
    const bool mv_dn = in_ins && L.sigma == 0;
    const bool mv_up = in_del && (at_p || L.sigma >= cc);
    const bool mv = mv_dn || mv_up;
    const bool mt = (w & (EV_MATCH << SH)) != 0;
    L.shared += (mt && aoff <= L.ioff) ? sgn : 0;
    L.room -= (mt && del) ? 1 : 0;
    const uint32_t a_nb = (aoff == noff) ? v2 : pn;                  // (only possible when moving down)
    const int dsh = mv_up ? (int)(a_nb & 1u) : (mv_dn ? -(int)(L.a & 1u) : 0);
    L.shared += dsh;
    L.a = mv ? a_nb : L.a;
    L.ioff = mv ? noff : L.ioff;
    const int sig_n = L.sigma + (in_del ? 1 : 0) - (in_ins ? 1 : 0);
    L.sigma = mv_dn ? (int)(L.a >> 1) : (mv_up ? 0 : sig_n);
    L.nb += (w & (EV_DEL << SH)) ? dir : 0;
    slide_eval(L, (gw & (EV_GRP << SH)) != 0);
}

// 4 bytes global -> shared without a register in between (LDGSTS); n = 0 writes zeros
__device__ __forceinline__ void cp_async_u32(uint32_t dst_sh, const void *src, int n)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_sh), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int L2_BATCH = 32;            // candidates a warp claims with one atomic

__global__ void __launch_bounds__(L2_THREADS)
l2_slide_kernel(const SlideJob *jobs, const Prep *prep, uint32_t n_cands, const uint2 *hw, const uint16_t *ev,
                const int32_t *min_shared, const uint32_t *id_off, const float *id_tab,
                Mapping *maps, unsigned long long *counters, int rows_above)
{
    extern __shared__ __align__(16) uint8_t l2_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr int pitch = L2_THREADS * 4;
    uint8_t *const st = l2_smem + pitch + tid * 4;                 // one row of slack below bucket 0
    const uint32_t st_sh = (uint32_t)__cvta_generic_to_shared(st);
    const uint32_t st_warp_sh = st_sh - (uint32_t)lane * 4u;       // column of lane 0 of this warp
    constexpr unsigned wmask = 0xFFFFFFFFu;
    const uint4 *jobv = reinterpret_cast<const uint4 *>(jobs);
    const unsigned long long n_c = n_cands;

    // fills the columns of the lanes in `lm` with their start states: piece `lane` (16 bytes = four rows) of every
    // block, zeros above the sketch (the slack rows); the copies land while the other lanes replay a chunk
    auto issue_states = [&](unsigned lm, const uint16_t *blk, int s) {
        while (lm) {
            const int t = __ffs(lm) - 1;
            lm &= lm - 1u;
            const unsigned long long bp = __shfl_sync(wmask, (unsigned long long)blk, t);
            const int nw = __shfl_sync(wmask, l2_words_for(s), t);
            const int w0 = 4 * lane;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(bp) + w0;
            const uint32_t dst = st_warp_sh + (uint32_t)t * 4u + (uint32_t)w0 * (uint32_t)pitch;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (w0 + j < rows_above) cp_async_u32(dst + (uint32_t)(j * pitch), w0 + j < nw ? src + j : src - w0, w0 + j < nw ? 4 : 0);
        }
    };

    bool have = false, wait = false, pend = false;
    uint32_t c = 0, xdel = 0;
    uint32_t pv0 = 0, sh0 = 0;                                     // pivot word and shared count of the start window
    int s = 0, kc = 0, kmc = 0, n_chunks = 0, room_l = 0, d_m = 0, dir = 1;
    int ms_s = 0;                                                  // min_shared[s]
    uint32_t ido_s = 0, pp_beg = 0;                                // id_off[s]; first reference minimizer of the region
    int pp_seq = 0;
    // result of the candidate that ended last trip: three gathers in flight
    uint32_t c_pend = 0, fpos = 0, lpos = 0;
    int seq_pend = 0, sh_pend = 0;
    float id_pend = 0.0f;
    unsigned long long replayed = 0;
    const uint16_t *blk = nullptr;
    const uint4 *evp = nullptr;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    SlideLane L{};

    // ---- the warp's queue: a batch of L2_BATCH candidates in use, the next batch claimed (its base may still be in flight
    // in lane 0); every lane holds the candidate after its current one with the descriptor in flight --------------------
    unsigned long long q_cur = 0, q_end = 0, q_next_raw = 0;
    if (lane == 0) {
        q_cur = atomicAdd(&counters[CT_WORK], (unsigned long long)(2 * L2_BATCH));
        q_next_raw = q_cur + L2_BATCH;
    }
    q_cur = __shfl_sync(wmask, q_cur, 0);
    q_end = q_cur + L2_BATCH;
    unsigned long long nc = q_cur + (unsigned long long)lane;     // (L2_BATCH == 32: the first batch is used up at once)
    q_cur = q_end;
    uint4 nj0 = make_uint4(0, 0, 0, 0), nj1 = make_uint4(0, 0, 0, 0);
    if (nc < n_c) { nj0 = __ldg(jobv + 2 * nc); nj1 = __ldg(jobv + 2 * nc + 1); }
    bool alive = true;

    for (;;) {
        bool fin = false, overflow = false;
        // ---- start states issued last trip have landed -----------------------------------------------
        cp_async_wait_all();
        __syncwarp();
        if (wait) {
            L.ioff = ev_aoff((int)(pv0 & 0xFFFFu)); L.sigma = (int)(pv0 >> 16); L.shared = (int)sh0; L.nb = d_m;
            L.a = lds_u8(st_sh + L.ioff);
            wait = false;
        }
        // ---- the result of the candidate that ended last trip ------------------------------------------
        if (pend) {
            Mapping mp;
            mp.seq = seq_pend;
            mp.ref_start = ((int)(fpos & 0x7FFFFFFFu) + (int)(lpos & 0x7FFFFFFFu)) / 2;    // computeMap.hpp:492
            mp.shared = sh_pend;
            mp.identity = sh_pend >= 0 ? id_pend : 0.0f;
            maps[c_pend] = mp;
            pend = false;
        }
        // ---- lanes without a candidate start the one they hold and take the one after it from the warp's queue ----
        const unsigned need = __ballot_sync(wmask, !have && alive);
        if (need) {
            const unsigned long long used = (unsigned long long)__popc(need);
            unsigned long long idx = q_cur + (unsigned long long)__popc(need & ((1u << lane) - 1u));
            q_cur += used;
            if (q_cur > q_end) {                                     // (warp-uniform) into the next batch; claim the one after it
                const unsigned long long q_next = __shfl_sync(wmask, q_next_raw, 0);
                if (idx >= q_end) idx = q_next + (idx - q_end);
                q_cur = q_next + (q_cur - q_end);
                q_end = q_next + L2_BATCH;
                if (lane == 0) q_next_raw = atomicAdd(&counters[CT_WORK], (unsigned long long)L2_BATCH);
            }
            if (!have && alive) {
                if (nc >= n_c) alive = false;
                else {
                    c = (uint32_t)nc;
                    const uint4 j0 = nj0, j1 = nj1;
                    nc = idx;
                    if (nc < n_c) { nj0 = __ldg(jobv + 2 * nc); nj1 = __ldg(jobv + 2 * nc + 1); }
                    if (j0.z) {                                      // (else: nothing to slide, ask again next trip)
                        s = (int)(j1.x & 0xFFFFu); d_m = (int)(j1.x >> 16);
                        blk = ev + (((unsigned long long)j0.y << 32) | j0.x);
                        evp = reinterpret_cast<const uint4 *>(blk + ev_state_u16(s));
                        n_chunks = (int)((j0.z + 6u) >> 3);                  // (j0.z = events + 1)
                        kmc = (int)(j0.w >> 3);
                        pv0 = j1.z; sh0 = j1.w & 0xFFFFu;
                        // the start window is an evaluation point (computeMap.hpp:467-481)
                        L.best = (int)sh0; L.pa = d_m; L.pb = d_m; L.mx = 0;
                        L.room = (int)(j1.y & 0xFFFFu); room_l = (int)(j1.y >> 16);
                        xdel = 0; dir = 1; kc = kmc;
                        have = true;
                        overflow = (j1.w >> 16) != 0u;
                        if (overflow) fin = true;
                        else if (kmc >= n_chunks) {                   // no window to the right of the start window
                            if (kmc == 0 || room_l < L.best) fin = true;
                            else { xdel = EV_DEL | (EV_DEL << 16); dir = -1; kc = kmc - 1; L.room = room_l; }
                        }
                        if (!fin) { nxt = __ldg(evp + kc); wait = true; }
                        // what the end of the candidate needs (all in flight until then)
                        const Prep pp = prep[c];
                        pp_beg = pp.beg; pp_seq = pp.seq;
                        ms_s = __ldg(min_shared + s); ido_s = __ldg(id_off + s);
                    }
                }
            }
            __syncwarp();
            issue_states(__ballot_sync(wmask, wait), blk, s);
        }
        if (!__any_sync(wmask, alive)) break;
        bool turn = false;
        if (have && !fin && !wait) {
            const uint4 cur = nxt;
            nxt = __ldg(evp + kc + dir);                             // (a chunk of slack behind the last list; the state block in front)
            // backward: the eight events of the chunk in reverse order.  The state in front of an event that ends a time group
            // is an evaluation point: the flag of an event is looked at after the one undone before it -- for the first
            // event of a chunk that is now, for the last one of the list (in front of it sits the first window) at the end.
            const bool bw = xdel != 0u;
            const uint32_t e0 = bw ? __byte_perm(cur.w, 0u, 0x1032) : cur.x, e1 = bw ? __byte_perm(cur.z, 0u, 0x1032) : cur.y;
            const uint32_t e2 = bw ? __byte_perm(cur.y, 0u, 0x1032) : cur.z, e3 = bw ? __byte_perm(cur.x, 0u, 0x1032) : cur.w;
            const uint32_t g0 = bw ? __funnelshift_r(e0, e1, 16) : e0, g1 = bw ? __funnelshift_r(e1, e2, 16) : e1;
            const uint32_t g2 = bw ? __funnelshift_r(e2, e3, 16) : e2, g3 = bw ? e3 >> 16 : e3;
            slide_eval(L, bw && (e0 & EV_GRP) != 0u);
            slide_event<0>(L, st_sh, e0, xdel, g0, dir);
            slide_event<16>(L, st_sh, e0, xdel, g0, dir);
            slide_event<0>(L, st_sh, e1, xdel, g1, dir);
            slide_event<16>(L, st_sh, e1, xdel, g1, dir);
            slide_event<0>(L, st_sh, e2, xdel, g2, dir);
            slide_event<16>(L, st_sh, e2, xdel, g2, dir);
            slide_event<0>(L, st_sh, e3, xdel, g3, dir);
            slide_event<16>(L, st_sh, e3, xdel, g3, dir);
            replayed += 8;
            overflow = L.mx > 255u;                                  // (checked once per chunk: the pivot strays at most eight buckets)
            // early stops: the shared count of a window is at most its number of elements in the sketch, and the windows
            // still to come on this side hold at most `room` of them -- below the best so far they change neither the
            // optimum nor its first / last position (computeMap.hpp:467-481)
            if (overflow) fin = true;
            else if (!bw) {
                if (kc + 1 >= n_chunks || L.room < L.best) {
                    if (kmc == 0 || room_l < L.best) fin = true;
                    else {                                           // turn round: the start state again, then to the left
                        xdel = EV_DEL | (EV_DEL << 16); dir = -1; kc = kmc - 1; L.room = room_l;
                        const int t = L.pa; L.pa = L.pb; L.pb = t;
                        nxt = __ldg(evp + kc);
                        turn = true;
                    }
                } else kc++;
            } else {
                if (kc == 0) { slide_eval(L, true); fin = true; }    // the first window
                else if (L.room < L.best) fin = true;
                else kc--;
            }
        }
        {
            const unsigned tm = __ballot_sync(wmask, turn);
            if (tm) { __syncwarp(); issue_states(tm, blk, s); wait = wait || turn; }   // (the owner's own stores of this trip come first)
        }
        if (fin) {
            if (overflow) {
                Mapping mp;
                mp.seq = pp_seq; mp.ref_start = L2_REDO; mp.shared = -1; mp.identity = 0.0f;
                maps[c] = mp;
                atomicAdd(&counters[CT_REDO], 1ull);
            } else {
                const bool bw = xdel != 0u;
                const int first_nb = bw ? L.pb : L.pa, last_nb = bw ? L.pa : L.pb;
                fpos = __ldg(&hw[pp_beg + (uint32_t)first_nb].y);
                lpos = __ldg(&hw[pp_beg + (uint32_t)last_nb].y);
                id_pend = __ldg(id_tab + ido_s + (uint32_t)L.best);
                sh_pend = L.best >= ms_s ? L.best : -1 - L.best;      // computeMap.hpp:371-380 via the table
                c_pend = c; seq_pend = pp_seq;
                pend = true;
            }
            have = false; wait = false;
        }
    }
    for (int o = 16; o > 0; o >>= 1) replayed += __shfl_xor_sync(wmask, replayed, o);
    if (lane == 0 && replayed) atomicAdd(&counters[CT_REPLAYED], replayed);
}

// ---- exact fallback (16-bit bucket counts, binary-search classification) -------------------------
struct SlideState {
    uint16_t *st;        // st[b * stride]: bit 15 = M(b) (b >= 1), bits 0..14 = cnt[b]
    int stride, s;
    int istar, sigma, shared;
    __device__ __forceinline__ uint16_t &at(int b) { return st[b * stride]; }
    __device__ __forceinline__ void ins_only(int b)
    {
        at(b) += 1;
        if (b < istar) {
            if (sigma > 0) sigma--;
            else { shared -= at(istar) >> 15; istar--; sigma = at(istar) & 0x7FFF; }
        }
    }
    __device__ __forceinline__ void del_only(int b)
    {
        at(b) -= 1;
        const int c = at(istar) & 0x7FFF;
        if (b < istar) {
            if (sigma < c) sigma++;
            else { istar++; shared += at(istar) >> 15; sigma = 0; }
        } else if (b == istar && sigma > c) { istar++; shared += at(istar) >> 15; sigma = 0; }
    }
    __device__ __forceinline__ void ins_match(int i) { at(i) |= 0x8000; if (i <= istar) shared++; }
    __device__ __forceinline__ void del_match(int i) { at(i) &= 0x7FFF; if (i <= istar) shared--; }
};

__global__ void __launch_bounds__(L2_THREADS)
l2_fallback_kernel(const Prep *prep, const uint32_t *cand_base, const uint32_t *work_base, int n_frags,
          const uint32_t *qhash, const uint64_t *seq_first, const int32_t *qs,
          const RefMini *ref, uint32_t n_ref, int cmw,
          const int32_t *min_shared, const uint32_t *id_off, const float *id_tab,
          Mapping *maps, unsigned long long *counters, int q_cap)
{
    extern __shared__ __align__(16) uint8_t l2_smem[];
    uint32_t *s_q = reinterpret_cast<uint32_t *>(l2_smem);
    uint16_t *s_state = reinterpret_cast<uint16_t *>(s_q + q_cap);
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x;
    const uint32_t n_work = work_base[n_frags];
    if (counters[CT_REDO] == 0) return;        // nothing left for the exact path (the usual case)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (uint32_t)atomicAdd(&counters[CT_WORK2], 1ull);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_work) break;
        // work item -> fragment (last f with work_base[f] <= item)
        int lo = 0, hi = n_frags - 1;
        while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (work_base[mid] <= item) lo = mid; else hi = mid - 1; }
        const int f = lo;
        const int s = qs[f];
        const int tpc = l2_fb_lanes_for(s);
        const uint32_t c_lo = cand_base[f] + (item - work_base[f]) * L2_ITEM;
        const uint32_t c_hi = min(c_lo + (uint32_t)L2_ITEM, cand_base[f + 1]);
        bool staged = false;
        for (uint32_t c0 = c_lo; c0 < c_hi; c0 += tpc) {
            const uint32_t c = c0 + tid;
            const bool active = tid < tpc && c < c_hi && maps[c].ref_start == L2_REDO;
            if (!__syncthreads_or(active ? 1 : 0)) continue;
            if (!staged) {
                const uint64_t qb = seq_first[f];
                for (int i = tid; i < s; i += L2_THREADS) s_q[i] = qhash[qb + i];
                staged = true;
            }
            for (int i = tid; i < (s + 1) * tpc; i += L2_THREADS) s_state[i] = 0;
            __syncthreads();
            if (!active) continue;

            const Prep pp = prep[c];
            const uint32_t beg = pp.beg, last = pp.last;
            const RefMini rbeg = ref[beg];
            uint32_t end0 = beg;                                              // first index with wpos >= wpos[beg] + cmw
            {
                uint32_t l = beg, r = last;
                while (l < r) { uint32_t mid = l + ((r - l) >> 1); if ((int)ref[mid].y < (int)rbeg.y + cmw) l = mid + 1; else r = mid; }
                end0 = l;
            }

            SlideState S;
            S.st = s_state + tid; S.stride = tpc; S.s = s; S.istar = s; S.sigma = 0; S.shared = 0;

            auto classify = [&](uint32_t h, int &b) -> bool {      // b = #{q < h}; true if q_{b+1} == h
                int l = 0, r = s;
                while (l < r) { int mid = (l + r) >> 1; if (s_q[mid] < h) l = mid + 1; else r = mid; }
                b = l;
                return l < s && s_q[l] == h;
            };
            auto insert = [&](uint32_t j, const RefMini &e, uint32_t win_beg) {      // window is [win_beg, j)
                const uint32_t dprev = e.w & 0xFFFFu;
                if (dprev && j >= dprev && j - dprev >= win_beg) return;      // same hash already present (REV)
                int b;
                if (classify(e.x, b)) S.ins_match(b + 1); else S.ins_only(b);
            };
            auto remove = [&](uint32_t j, const RefMini &e, uint32_t win_end) {      // window is [j, win_end)
                const uint32_t dnext = e.w >> 16;
                if (dnext && j + dnext < win_end) return;                     // a later copy keeps the hash present (NOOP)
                int b;
                if (classify(e.x, b)) S.del_match(b + 1); else S.del_only(b);
            };

            for (uint32_t j = beg; j < end0; j++) insert(j, ref[j], beg);     // first super-window, computeMap.hpp:446

            // MIIteratorL2 (MIIteratorL2.hpp:54-96) + the slide loop of computeMap.hpp:453-488
            uint32_t sw_beg = beg, sw_end = end0, prev_end = end0;
            int sw_pos = (int)rbeg.y;
            int best = 0, first_pos = 0, last_pos = 0;
            const bool runs = end0 < last;
            RefMini e_b = rbeg, e_b1 = runs ? ref[beg + 1] : rbeg, e_e = runs ? ref[end0] : rbeg, old_b = rbeg, old_e = rbeg;
            bool adv_b = false, adv_e = false;
            while (sw_end < last) {
                if (adv_b) remove(sw_beg - 1, old_b, prev_end);               // :459-460
                if (adv_e) insert(sw_end - 1, old_e, sw_beg);                 // :463-464
                const int wb = (int)e_b.y;
                if (S.shared > best) { best = S.shared; first_pos = last_pos = wb; }     // :467-476
                else if (S.shared == best) last_pos = wb;                                // :477-481
                const int d1 = (int)e_b1.y - sw_pos;
                const int d2 = (int)e_e.y - (sw_pos + cmw - 1);
                const int adv = min(d1, d2);
                sw_pos += adv;
                adv_b = adv == d1; adv_e = adv == d2;
                prev_end = sw_end;
                if (adv_b) { old_b = e_b; e_b = e_b1; sw_beg++; if (sw_beg + 1 < n_ref) e_b1 = ref[sw_beg + 1]; }
                if (adv_e) { old_e = e_e; sw_end++; if (sw_end < last) e_e = ref[sw_end]; }
            }
            Mapping mp;
            mp.seq = pp.seq;
            mp.ref_start = (first_pos + last_pos) / 2;                        // computeMap.hpp:492
            const bool pass = s > 0 && best >= min_shared[s];                 // computeMap.hpp:371-380 via the table
            mp.shared = pass ? best : -1 - best;
            mp.identity = pass ? id_tab[id_off[s] + best] : 0.0f;
            maps[c] = mp;
        }
    }
}

// ---- K5: core-genome identity (computeCoreIdentity.hpp:163-295) --------------------------------
// Pass 1 (:210-231): per (genome, fragment) keep the best mapping by (identity, refSeqId,
// refStartPos).  Candidates of a fragment are ordered by reference index, genome ids are monotone
// in it, so each group is a contiguous run of candidate slots; its first slot does the work.
// Pass 2 (:234-255): per (ref contig, bin) keep the best identity -> atomicMax on the bit
// pattern (identities are positive floats, so the unsigned order is the float order).
__global__ void cgi_best_kernel(const Cand *cands, const Mapping *maps, const uint32_t *cand_base, int n_frags,
                                const int32_t *genome_of_seq, const uint32_t *bin_base, int bin_w, uint32_t *cells,
                                const int32_t *frag_q, uint32_t n_cells, unsigned long long *counters)
{
    const uint32_t n = cand_base[n_frags];
    unsigned int passed = 0;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const Mapping m = maps[c];
        passed += m.shared >= 0 ? 1u : 0u;
        const int f = cands[c].frag;
        const int g = genome_of_seq[m.seq];
        const uint32_t cb = cand_base[f], ce = cand_base[f + 1];
        if (c > cb && genome_of_seq[maps[c - 1].seq] == g) continue;      // not the first slot of its group
        bool have = false;
        Mapping best = m;
        for (uint32_t p = c; p < ce; p++) {
            const Mapping q = maps[p];
            if (genome_of_seq[q.seq] != g) break;
            if (q.shared < 0) continue;
            if (!have || q.identity > best.identity ||
                (q.identity == best.identity && (q.seq > best.seq || (q.seq == best.seq && q.ref_start > best.ref_start)))) {
                best = q; have = true;
            }
        }
        // (several queries in one pass: every query has its own cell table)
        if (have) atomicMax(&cells[(frag_q ? (uint32_t)frag_q[f] * n_cells : 0u) + bin_base[best.seq] + (uint32_t)(best.ref_start / bin_w)],
                            __float_as_uint(best.identity));
    }
    passed = __reduce_add_sync(0xFFFFFFFFu, passed);
    if ((threadIdx.x & 31) == 0 && passed) atomicAdd(&counters[CT_MAPPINGS], (unsigned long long)passed);
}

// Per genome (:268-294): float32 sum of the surviving identities in (refSeqId, bin) order, count,
// mean.  One warp per genome; the adds are sequential on purpose.  Cells are
// cleared on the way so the table is ready for the next query.
__global__ void cgi_sum_kernel(uint32_t *cells, const uint32_t *genome_cell, int n_genomes, int n_queries, uint32_t n_cells,
                               int32_t *g_count, float *g_identity)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;    // (query, genome) pair
    if (g >= n_genomes * n_queries) return;
    const uint32_t qb = (uint32_t)(g / n_genomes) * n_cells;
    const uint32_t b = qb + genome_cell[g % n_genomes], e = qb + genome_cell[g % n_genomes + 1];
    // The float32 sum runs over the non-empty cells in (contig, bin) order (computeCoreIdentity.hpp:264-294), one add
    // after the other.  Empty cells hold +0.0f and x + 0.0f == x bit for bit (identities are positive), so lane 0 adds
    // every cell of a 128-cell batch straight from shared memory -- a chain of plain FADDs, no shuffles or bit scans --
    // while the loads of the next batch are in flight.
    __shared__ __align__(16) float s_v[8][128];
    float *sv = s_v[(threadIdx.x >> 5) & 7];
    float sum = 0.0f;
    int cnt = 0;
    uint32_t v[4], nx[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const uint32_t i = b + u * 32 + lane; nx[u] = i < e ? cells[i] : 0u; }
    for (uint32_t base = b; base < e; base += 128) {
        int in_batch = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u] = nx[u];
            const uint32_t i = base + u * 32 + lane;
            if (v[u]) cells[i] = 0u;
            const uint32_t j = i + 128;
            nx[u] = j < e ? cells[j] : 0u;
            in_batch += __popc(__ballot_sync(0xFFFFFFFFu, v[u] != 0u));
            sv[u * 32 + lane] = __uint_as_float(v[u]);
        }
        cnt += in_batch;
        __syncwarp();
        if (lane == 0 && in_batch) {
#pragma unroll 8
            for (int j = 0; j < 128; j += 4) {
                const float4 x = *reinterpret_cast<const float4 *>(sv + j);
                sum += x.x; sum += x.y; sum += x.z; sum += x.w;
            }
        }
        __syncwarp();
    }
    if (lane == 0) { g_count[g] = cnt; g_identity[g] = cnt ? sum / (float)cnt : 0.0f; }
}

template <typename T>
int excl_scan(cudaStream_t st, DevBuf<uint8_t> &tmp, const T *in, T *out, int64_t n, int *launches)
{
    size_t bytes = 0;
    FA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
    FA_TRY(tmp.reserve(bytes + 16));
    FA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, n, st));
    if (launches) *launches += 2;
    return FA_OK;
}

// Device-resident contigs of a draft (hundreds per query) into the batch buffer with ONE launch instead of one
// cudaMemcpyAsync each: thread t owns the 16 bytes at 16 t; uploads start at ascending 16-byte aligned offsets.
// kind 1: the source is a 2-bit packed sequence (fa_packed.bits, already in device memory behind the batch bytes) and
// the 16 bytes are expanded from one 32-bit word of it.
struct DevCopy { const uint8_t *src; uint64_t off; int64_t len; int64_t kind; };
__global__ void gather_contigs_kernel(const DevCopy *tab, int n, uint8_t *dst, uint64_t n_chunks)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_chunks) return;
    const uint64_t c = t * 16;
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tab[mid].off <= c) lo = mid; else hi = mid - 1; }
    const DevCopy d = tab[lo];
    if (c < d.off) return;
    const int64_t rel = (int64_t)(c - d.off);
    if (rel >= d.len) return;                                // padding between two uploads
    if (d.kind == 1) {
        // 16 bases = 4 packed bytes (the packed copy is padded to a multiple of 4 bytes); A C G T = 0 1 2 3
        const uint32_t wd = *reinterpret_cast<const uint32_t *>(d.src + (rel >> 2));
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t x = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t v = (wd >> (8 * j + 2 * i)) & 3u;
                x |= ((0x54474341u >> (8 * v)) & 0xFFu) << (8 * i);   // "ACGT" little-endian
            }
            o[j] = x;
        }
        *reinterpret_cast<uint4 *>(dst + c) = make_uint4(o[0], o[1], o[2], o[3]);   // (the slack behind an upload is ours: offsets are 16-aligned)
        return;
    }
    const uint8_t *sp = d.src + rel;
    if (d.len - rel >= 16 && ((uintptr_t)sp & 15) == 0) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(sp);
    else {
        const int m = (int)min((int64_t)16, d.len - rel);
        for (int i = 0; i < m; i++) dst[c + i] = sp[i];
    }
}

// The bytes of a packed sequence that are not A, C, G or T (fa_packed runs): one warp per run writes its value over the
// expanded bases.
struct RunDesc { uint64_t off; uint32_t len; uint32_t byte; };
__global__ void packed_runs_kernel(const RunDesc *runs, uint64_t n_runs, uint8_t *dst)
{
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_runs) return;
    const RunDesc d = runs[r];
    for (uint32_t i = threadIdx.x & 31; i < d.len; i += 32) dst[d.off + i] = (uint8_t)d.byte;
}

inline int bits_for(uint64_t n) { int b = 1; while (b < 63 && (1ull << b) < n) b++; return b; }

}  // namespace

// Narrow / copy the uploads into the pinned staging buffer (host sources), then H2D copies and one D2D per
// device-resident source.  The staged bytes leave in pieces of about 1 MiB, so the copy engine works on one piece
// while the host fills the next.  `workers` > 0: that many helper threads fill the pieces and this thread only issues
// the copies, in order, as pieces complete -- used when a whole pass of fa_query_batch (tens of MB) is staged ahead;
// for a lone 5 MB query spawning the threads costs more than the 0.3 ms they save.
// 2-bit packed sources (fa_packed): their bits and their run table travel to a region behind the batch bytes -- a
// quarter of the bytes through host memory and PCIe -- and gather_contigs_kernel / packed_runs_kernel expand them.
int stage_sequences(cudaStream_t st, DevBuf<uint8_t> &bytes, PinBuf &stage, const std::vector<Upload> &ups, uint64_t total,
                    uint64_t *h2d_bytes, int workers)
{
    // device-resident sources: more than a few go through one gather launch, its table behind the bytes
    std::vector<DevCopy> dcopy;
    struct Packed { const Upload *u; const fa_packed *pk; uint64_t bits_off, bits_bytes; };
    std::vector<Packed> packed;
    std::vector<RunDesc> runs;
    size_t n_dev = 0;
    for (const Upload &u : ups) {
        if (u.len <= 0) continue;
        if (u.unit == FA_UNIT_PACKED2) packed.push_back(Packed{&u, (const fa_packed *)u.ptr, 0, ((uint64_t)u.len + 3) / 4});
        else if (u.on_device) n_dev++;
    }
    const bool gather = n_dev > 4 || !packed.empty();
    const uint64_t tab_off = (total + 64 + 15) & ~15ull, tab_bytes = gather ? (n_dev + packed.size()) * sizeof(DevCopy) : 0;
    uint64_t end = tab_off + tab_bytes;
    for (Packed &pk : packed) {
        pk.bits_off = (end + 15) & ~15ull;
        end = pk.bits_off + ((pk.bits_bytes + 3) & ~3ull);
        for (uint64_t r = 0; r < pk.pk->n_runs; r++) {
            const uint64_t pos = pk.pk->run_pos[r];
            if (pos >= (uint64_t)pk.u->len) break;                       // (a query stages whole fragments only)
            runs.push_back(RunDesc{pk.u->off + pos, (uint32_t)std::min<uint64_t>(pk.pk->run_len[r], (uint64_t)pk.u->len - pos), pk.pk->run_byte[r]});
        }
    }
    const uint64_t run_off = (end + 15) & ~15ull;
    end = run_off + runs.size() * sizeof(RunDesc);
    FA_TRY(bytes.reserve(end));
    // (source, destination offset, bytes, unit): units 2 / 4 are narrowed on the way
    struct Piece { const void *src; uint64_t dst; int64_t n; int unit; };
    std::vector<Piece> pieces;
    const int64_t piece = 1ll << 20;
    for (const Upload &u : ups) {
        if (u.on_device || u.len <= 0 || u.unit == FA_UNIT_PACKED2) continue;
        for (int64_t o = 0; o < u.len; o += piece)
            pieces.push_back(Piece{(const uint8_t *)u.ptr + o * u.unit, u.off + (uint64_t)o, std::min<int64_t>(piece, u.len - o), u.unit});
    }
    for (const Packed &pk : packed)
        for (uint64_t o = 0; o < pk.bits_bytes; o += (uint64_t)piece)
            pieces.push_back(Piece{pk.pk->bits + o, pk.bits_off + o, (int64_t)std::min<uint64_t>((uint64_t)piece, pk.bits_bytes - o), 1});
    if (!pieces.empty() || gather) FA_TRY(stage.reserve(end));
    if (!pieces.empty()) {
        uint8_t *const sp = stage.p;
        auto fill = [sp](const Piece &pc) {
            uint8_t *dst = sp + pc.dst;
            if (pc.unit == 1) memcpy(dst, pc.src, (size_t)pc.n);
            else {
                // pyx:147-148: (char)toupper(code point); glibc's toupper leaves values outside
                // [-128, 255] unchanged
                for (int64_t i = 0; i < pc.n; i++) {
                    uint32_t cp = pc.unit == 2 ? ((const uint16_t *)pc.src)[i] : ((const uint32_t *)pc.src)[i];
                    if (cp >= 'a' && cp <= 'z') cp -= 32;
                    dst[i] = (uint8_t)cp;
                }
            }
        };
        // stage.p[lo .. hi) is filled and not yet on its way; pieces come at increasing offsets, a gap (the slots of
        // device-resident contigs, the table) starts a new copy
        uint64_t lo = pieces[0].dst, hi = lo;
        auto flush = [&]() -> int {
            if (hi > lo) {
                FA_CUDA(cudaMemcpyAsync(bytes.p + lo, sp + lo, hi - lo, cudaMemcpyHostToDevice, st));
                if (h2d_bytes) *h2d_bytes += hi - lo;
            }
            lo = hi;
            return FA_OK;
        };
        const size_t np = pieces.size();
        const int nw = np >= 16 ? std::min<int>(workers, 8) : 0;
        std::vector<std::thread> pool;
        std::unique_ptr<std::atomic<int>[]> ready;
        std::atomic<size_t> next{0};
        if (nw > 0) {
            ready.reset(new std::atomic<int>[np]);
            for (size_t i = 0; i < np; i++) ready[i].store(0, std::memory_order_relaxed);
            for (int t = 0; t < nw; t++)
                pool.emplace_back([&]() {
                    for (;;) {
                        const size_t i = next.fetch_add(1, std::memory_order_relaxed);
                        if (i >= np) break;
                        fill(pieces[i]);
                        ready[i].store(1, std::memory_order_release);
                    }
                });
        }
        int rc = FA_OK;
        for (size_t i = 0; i < np; i++) {
            if (nw > 0) { while (!ready[i].load(std::memory_order_acquire)) std::this_thread::yield(); }
            else fill(pieces[i]);
            if (rc != FA_OK) continue;                                   // (keep draining the workers after an error)
            if (pieces[i].dst > hi + 64) { rc = flush(); lo = hi = pieces[i].dst; }
            hi = pieces[i].dst + (uint64_t)pieces[i].n;
            if (rc == FA_OK && hi - lo >= (uint64_t)piece) rc = flush();
        }
        for (auto &t : pool) t.join();
        FA_TRY(rc);
        FA_TRY(flush());
    }
    if (gather) {
        for (const Upload &u : ups)
            if (u.len > 0 && (u.on_device || u.unit == FA_UNIT_PACKED2)) dcopy.push_back(DevCopy{(const uint8_t *)u.ptr, u.off, u.len, 0});
        size_t ip = 0;
        for (DevCopy &d : dcopy)
            if (ip < packed.size() && d.off == packed[ip].u->off && d.src == (const uint8_t *)packed[ip].pk) {
                d.src = bytes.p + packed[ip].bits_off; d.kind = 1; ip++;
            }
        memcpy(stage.p + tab_off, dcopy.data(), (size_t)tab_bytes);
        FA_CUDA(cudaMemcpyAsync(bytes.p + tab_off, stage.p + tab_off, (size_t)tab_bytes, cudaMemcpyHostToDevice, st));
        const uint64_t n_chunks = (total + 15) / 16;
        gather_contigs_kernel<<<(unsigned int)((n_chunks + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const DevCopy *>(bytes.p + tab_off), (int)dcopy.size(), bytes.p, n_chunks);
        FA_CUDA(cudaGetLastError());
        if (!runs.empty()) {
            memcpy(stage.p + run_off, runs.data(), runs.size() * sizeof(RunDesc));
            FA_CUDA(cudaMemcpyAsync(bytes.p + run_off, stage.p + run_off, runs.size() * sizeof(RunDesc), cudaMemcpyHostToDevice, st));
            if (h2d_bytes) *h2d_bytes += runs.size() * sizeof(RunDesc);
            packed_runs_kernel<<<(unsigned int)((runs.size() * 32 + 255) / 256), 256, 0, st>>>(
                reinterpret_cast<const RunDesc *>(bytes.p + run_off), runs.size(), bytes.p);
            FA_CUDA(cudaGetLastError());
        }
    } else {
        for (const Upload &u : ups)
            if (u.on_device && u.len > 0) FA_CUDA(cudaMemcpyAsync(bytes.p + u.off, u.ptr, (size_t)u.len, cudaMemcpyDeviceToDevice, st));
    }
    return FA_OK;
}

// Whole fragments of every contig that is long enough (pyx:1059-1105), at 16-byte aligned offsets of the batch buffer.
void plan_uploads(const fa_params &P, const fa_contig *contigs, int32_t n_contigs, std::vector<Upload> &ups, uint64_t *total)
{
    const int L = P.frag_len;
    const int lim = std::min(std::min(P.window, P.k), L);
    uint64_t off = 0;
    ups.clear();
    for (int32_t c = 0; c < n_contigs; c++) {
        const int64_t slen = contigs[c].len;
        if (slen < lim) continue;
        const int64_t nfrag = slen / L;
        if (nfrag > 0) {
            ups.push_back(Upload{contigs[c].data, contigs[c].unit_bytes, contigs[c].on_device, nfrag * L, off});
            off += ((uint64_t)(nfrag * L) + 15) & ~15ull;
        }
    }
    *total = off;
}

// The slot a sketch travels in: twice the expected size of a fragment's sketch (2 L / (w + 1) minimizers) and some, never
// more than a fragment can hold.  0 = packing would not make the sketches smaller than their slots already are.
int exchange_stride(const fa_params &P)
{
    const int L = P.frag_len, w = std::max(P.window, 1), k = P.k;
    const int cmw = L - (w - 1) - (k - 1);
    if (cmw <= 0) return 0;
    int want = ((4 * L / (w + 1) + 64) + 31) & ~31;
    if (const char *e = getenv("FA_EXCHANGE_STRIDE")) { if (*e && atoi(e) > 0) want = (atoi(e) + 31) & ~31; }   // (test hook: slots too small for the sketches)
    return want < cmw ? want : 0;
}

int sketch_share(fa_index *ix, const fa_contig *contigs, int32_t n_contigs, int world, int rank, uint32_t stride, uint32_t *per_out,
                 uint64_t *frags_out, fa_query_info *qi)
{
    FA_CUDA(cudaSetDevice(ix->device));
    ExchScratch &ws = ix->xs;
    if (!ws.st) {
        FA_CUDA(cudaStreamCreateWithFlags(&ws.st, cudaStreamNonBlocking));
        for (auto &e : ws.done) FA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        FA_CUDA(cudaEventCreate(&ws.t0)); FA_CUDA(cudaEventCreate(&ws.t1));
    }
    cudaStream_t st = ws.st;
    const fa_params &P = ix->prm;
    const int L = P.frag_len, k = P.k, w = P.window;
    const int lim = std::min(std::min(w, k), L);
    const int nk = L - k + 1;
    const int tiles_per_frag = nk > 0 ? (nk + SK_TILE - 1) / SK_TILE : 0;
    const int cmw = L - (w - 1) - (k - 1);
    if (L <= 0 || L > 32767 || tiles_per_frag <= 0 || cmw <= 0) { set_error("sketch exchange: unsupported fragment length"); return FA_ERR_UNSUPPORTED; }
    uint64_t F = 0;
    for (int32_t c = 0; c < n_contigs; c++) {
        if (contigs[c].len < lim) continue;
        FA_TRY(check_contig(contigs[c], c));
        F += (uint64_t)(contigs[c].len / L);
    }
    *frags_out = F;
    const uint64_t per = std::max<uint64_t>((F + (uint64_t)world - 1) / (uint64_t)world, 1);
    *per_out = (uint32_t)per;
    if (per > 0x7FFFFFFFull / (uint64_t)std::max(tiles_per_frag, 1)) { set_error("sketch exchange: group too large"); return FA_ERR_UNSUPPORTED; }
    const uint64_t f0 = std::min<uint64_t>(F, (uint64_t)rank * per), f1 = std::min<uint64_t>(F, f0 + per);
    const int n_loc = (int)(f1 - f0);
    // the bytes of fragments [f0, f1): the slice of every contig that holds some of them (a packed contig from its start)
    std::vector<Upload> ups;
    ws.h_seqs.clear();
    uint64_t a = 0, off = 0;
    for (int32_t c = 0; c < n_contigs && a < f1; c++) {
        const fa_contig &ct = contigs[c];
        if (ct.len < lim) continue;
        const uint64_t nfrag = (uint64_t)(ct.len / L);
        const uint64_t x = std::max(a, f0), y = std::min(a + nfrag, f1);
        if (x < y) {
            const bool packed = ct.unit_bytes == FA_UNIT_PACKED2;
            const uint64_t skip = packed ? 0 : x - a;                        // fragments of the contig left out in front
            const uint8_t *src = (const uint8_t *)ct.data + (packed ? 0 : skip * (uint64_t)L * (uint64_t)ct.unit_bytes);
            ups.push_back(Upload{src, ct.unit_bytes, ct.on_device, (int64_t)((y - a - skip) * (uint64_t)L), off});
            for (uint64_t i = x; i < y; i++) {
                SeqDesc d;
                d.off = off + (i - a - skip) * (uint64_t)L; d.len = L; d.id = (int32_t)(i - f0);
                d.raw = contig_prenormalised(ct); d.tile0 = (int32_t)((i - f0) * (uint64_t)tiles_per_frag);
                ws.h_seqs.push_back(d);
            }
            off += (((y - a - skip) * (uint64_t)L) + 15) & ~15ull;
        }
        a += nfrag;
    }
    FA_CUDA(cudaEventRecord(ws.t0, st));
    FA_TRY(ws.send.reserve((size_t)per * (stride + 1)));
    FA_TRY(ws.qs.reserve(std::max(n_loc, 1))); FA_TRY(ws.seq_cnt.reserve(std::max(n_loc, 1))); FA_TRY(ws.sk.seq_first.reserve(std::max(n_loc, 1)));
    FA_TRY(ws.counters.reserve(CT_N));
    int launches = 0, sort_cap = 0;
    if (n_loc > 0) {
        uint64_t h2d = 0;
        FA_TRY(stage_sequences(st, ws.sk.bytes, ws.stage, ups, off, &h2d));
        const int n_tiles = n_loc * tiles_per_frag;
        FA_TRY(ws.sk.seqs.reserve(n_loc)); FA_TRY(ws.sk.tile_status.reserve(n_tiles)); FA_TRY(ws.sk.counters.reserve(4));
        FA_TRY(ws.qhash.reserve((uint64_t)n_loc * (uint64_t)cmw));
        FA_CUDA(cudaMemcpyAsync(ws.sk.seqs.p, ws.h_seqs.data(), (size_t)n_loc * sizeof(SeqDesc), cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemsetAsync(ws.counters.p, 0, CT_N * sizeof(unsigned long long), st));
        FA_TRY(launch_sketch(st, ws.sk, n_loc, n_tiles, k, w, P.alphabet != 4, nullptr, ws.qhash.p, 0, &launches, cmw, ws.seq_cnt.p, tiles_per_frag));
        int p2 = 1; while (p2 < cmw) p2 <<= 1;
        sort_cap = std::min(p2, 32768);
        const size_t smem = (size_t)sort_cap * 4;
        if (smem > 48 * 1024) FA_CUDA(cudaFuncSetAttribute(sort_unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sort_unique_kernel<<<n_loc, 256, smem, st>>>(ws.qhash.p, ws.sk.seq_first.p, ws.seq_cnt.p, n_loc, sort_cap, ix->s_max, ws.qs.p, ws.counters.p);
        FA_CUDA(cudaGetLastError()); launches++;
        if (qi) qi->h2d_bytes += h2d + (uint64_t)n_loc * sizeof(SeqDesc);
    }
    pack_sketch_kernel<<<(unsigned int)per, 128, 0, st>>>(ws.qhash.p, ws.sk.seq_first.p, ws.seq_cnt.p, ws.qs.p, n_loc, sort_cap, (uint32_t)per, stride,
                                                         ws.send.p);
    FA_CUDA(cudaGetLastError()); launches++;
    if (qi) qi->kernel_launches += launches;
    return FA_OK;
}

void Prefetch::release()
{
    bytes.release(); stage.release();
    if (done) cudaEventDestroy(done);
    if (st) cudaStreamDestroy(st);
    done = nullptr; st = nullptr; valid = false;
}

// Stage the bytes of a query that is still waiting for its turn (called from a helper thread of fa_query_batch).
// Anything unusual -- bad arguments, nothing to upload -- leaves the slot invalid and run_query stages as usual.
int prefetch_query(fa_index *ix, Prefetch &pf, const fa_contig *contigs, int32_t n_contigs)
{
    pf.valid = false;
    if (ix->prm.frag_len <= 0 || ix->prm.frag_len > 32767) return FA_OK;
    for (int32_t c = 0; c < n_contigs; c++) {
        if (check_contig(contigs[c], c) != FA_OK) return FA_OK;
    }
    FA_CUDA(cudaSetDevice(ix->device));
    if (!pf.st) FA_CUDA(cudaStreamCreateWithFlags(&pf.st, cudaStreamNonBlocking));
    if (!pf.done) FA_CUDA(cudaEventCreateWithFlags(&pf.done, cudaEventDisableTiming));
    std::vector<Upload> ups;
    uint64_t total = 0;
    plan_uploads(ix->prm, contigs, n_contigs, ups, &total);
    if (!total) return FA_OK;
    pf.h2d_bytes = 0;
    FA_TRY(stage_sequences(pf.st, pf.bytes, pf.stage, ups, total, &pf.h2d_bytes, 3));
    FA_CUDA(cudaEventRecord(pf.done, pf.st));
    pf.contigs = contigs; pf.n_contigs = n_contigs; pf.total = total;
    pf.valid = true;
    return FA_OK;
}

int run_query(fa_index *ix, const fa_contig *contigs, int32_t n_contigs, fa_hit *out, uint64_t cap, uint64_t *n_out,
              fa_query_info *info, Prefetch *pf)
{
    uint64_t offs[2] = {0, 0};
    const int rc = run_queries(ix, contigs, &n_contigs, 1, out, cap, offs, info, pf);
    *n_out = offs[1];
    return rc;
}

// One pass of the pipeline over the fragments of `n_queries` queries (query q owns the next contigs_per_query[q]
// contigs).  Sketch, lookup, L1 and L2 see one flat list of fragments; only the core-genome step knows the queries:
// every query has its own (contig, bin) cell table and per-genome sums.  hit_offsets[q + 1] - hit_offsets[q] hits of
// query q are written at out + hit_offsets[q] (all counted, only those below `cap` stored).
int run_queries(fa_index *ix, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries, fa_hit *out,
                uint64_t cap, uint64_t *hit_offsets, fa_query_info *info, Prefetch *pf, const PreSketch *ps)
{
    std::lock_guard<std::mutex> guard(ix->mtx);
    FA_CUDA(cudaSetDevice(ix->device));
    cudaStream_t st = ix->st;
    Workspace &ws = ix->ws;
    const fa_params &P = ix->prm;
    fa_query_info qi;
    memset(&qi, 0, sizeof qi);
    int launches = 0;
    if (!ws.ev_ready) {
        for (auto &e : ws.ev) FA_CUDA(cudaEventCreate(&e));
        ws.ev_ready = true;
    }
    ws.last_cands = 0; ws.last_frags = 0;
    const uint32_t B = (uint32_t)std::max(n_queries, 0);
    for (uint32_t q = 0; q <= B; q++) hit_offsets[q] = 0;

    // ---- fragments (pyx:1059-1105) -----------------------------------------------------------
    const int L = P.frag_len, k = P.k, w = P.window;
    const uint32_t d_near = std::min<uint32_t>((uint32_t)(std::max(L - 1, 0) / std::max(w, 1)), 1023u);   // l1_fused_kernel: index distance that is near for sure
    const int lim = std::min(std::min(w, k), L);
    std::vector<Upload> ups;
    ws.h_seqs.clear(); ws.h_fragq.clear();
    std::vector<uint64_t> q_len(B, 0), q_frags(B, 0);
    uint64_t total_frags = 0, off = 0;
    int32_t n_contigs = 0;
    const int nk = L - k + 1;
    const int tiles_per_frag = nk > 0 ? (nk + SK_TILE - 1) / SK_TILE : 0;
    for (uint32_t q = 0; q < B; q++) {
        for (int32_t c = n_contigs; c < n_contigs + contigs_per_query[q]; c++) {
            const int64_t slen = contigs[c].len;
            if (slen < lim) { qi.short_contigs++; continue; }                  // pyx:1062-1070
            FA_TRY(check_contig(contigs[c], c));
            const int64_t nfrag = slen / L;                                      // pyx:1097
            if (nfrag > 0) {
                if (!ps) {                                                       // (pre-sketched fragments: nothing to stage or sketch)
                    ups.push_back(Upload{contigs[c].data, contigs[c].unit_bytes, contigs[c].on_device, nfrag * L, off});
                    for (int64_t i = 0; i < nfrag; i++) {
                        SeqDesc d;
                        d.off = off + (uint64_t)i * L; d.len = L; d.id = (int32_t)(total_frags + i);
                        d.raw = contig_prenormalised(contigs[c]); d.tile0 = (int32_t)((total_frags + i) * tiles_per_frag);
                        ws.h_seqs.push_back(d);
                    }
                }
                if (B > 1) ws.h_fragq.insert(ws.h_fragq.end(), (size_t)nfrag, (int32_t)q);
                off += ((uint64_t)(nfrag * L) + 15) & ~15ull;
            }
            total_frags += (uint64_t)nfrag; q_frags[q] += (uint64_t)nfrag;       // pyx:1104
            q_len[q] += (uint64_t)slen;                                          // pyx:1105
        }
        n_contigs += contigs_per_query[q];
    }
    const int F = (int)total_frags;
    const uint32_t G = (uint32_t)ix->seqs_by_genome.size();
    qi.fragments = total_frags;
    if ((uint64_t)F != total_frags || (uint64_t)F * tiles_per_frag > 0x7FFFFFF0ull) { set_error("query too large for one call"); return FA_ERR_UNSUPPORTED; }
    if (L > 32767) { set_error("fragment_length > 32767 is not supported on the device path"); return FA_ERR_UNSUPPORTED; }

    std::vector<int32_t> h_count;
    std::vector<float> h_ident;
    NvtxStages nv;
    FA_CUDA(cudaEventRecord(ws.ev[0], st));
    nv.next("fa:query stage + h2d");
    if (F > 0 && tiles_per_frag > 0 && ix->n > 0 && G > 0) {
        // ---- upload + sketch the fragments ---------------------------------------------------
        if (ps) {
            // (sketched across the ranks and gathered: nothing to stage)
        } else if (pf && pf->valid && pf->contigs == contigs && pf->n_contigs == n_contigs && pf->total == off) {
            // staged ahead by fa_query_batch: take its buffer (it gets ours, idle since the previous query returned)
            std::swap(ws.sk.bytes.p, pf->bytes.p); std::swap(ws.sk.bytes.cap, pf->bytes.cap);
            FA_CUDA(cudaStreamWaitEvent(st, pf->done, 0));
            qi.h2d_bytes += pf->h2d_bytes;
            pf->valid = false;
        } else {
            FA_TRY(stage_sequences(st, ws.sk.bytes, ws.stage, ups, off, &qi.h2d_bytes));
        }
        const int n_tiles = F * tiles_per_frag;
        const int cmw = L - (w - 1) - (k - 1);                                // minimizer windows per fragment
        const uint64_t emit_cap = (uint64_t)F * (uint64_t)(ps ? ps->stride : (uint32_t)std::max(cmw, 1));
        FA_TRY(ws.sk.seqs.reserve(F)); FA_TRY(ws.sk.tile_status.reserve(n_tiles));
        FA_TRY(ws.sk.counters.reserve(4)); FA_TRY(ws.sk.seq_first.reserve(F));
        FA_TRY(ws.qhash.reserve(emit_cap)); FA_TRY(ws.hit_start.reserve(emit_cap)); FA_TRY(ws.hit_cnt.reserve(emit_cap));
        FA_TRY(ws.qs.reserve(F)); FA_TRY(ws.frag_seeds.reserve((size_t)F + 1));
        FA_TRY(ws.frag_cands.reserve((size_t)F + 1)); FA_TRY(ws.work_base.reserve((size_t)F + 1));
        FA_TRY(ws.counters.reserve(CT_N));
        const uint64_t n_cells = ix->n_cells ? ix->n_cells : 1;
        if ((uint64_t)B * n_cells > 0xFFFFFFFFull) { set_error("too many queries in one pass"); return FA_ERR_UNSUPPORTED; }
        if (ws.cells.cap < (size_t)B * n_cells) {
            FA_TRY(ws.cells.reserve((size_t)B * n_cells));
            FA_CUDA(cudaMemsetAsync(ws.cells.p, 0, ws.cells.cap * sizeof(uint32_t), st));   // kept clean by cgi_sum_kernel afterwards
        }
        FA_TRY(ws.g_count.reserve((size_t)B * G)); FA_TRY(ws.g_identity.reserve((size_t)B * G));
        if (B > 1) {
            FA_TRY(ws.frag_q.reserve(F));
            FA_CUDA(cudaMemcpyAsync(ws.frag_q.p, ws.h_fragq.data(), (size_t)F * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            qi.h2d_bytes += (uint64_t)F * sizeof(int32_t);
        }
        if (!ps) {
            FA_CUDA(cudaMemcpyAsync(ws.sk.seqs.p, ws.h_seqs.data(), (size_t)F * sizeof(SeqDesc), cudaMemcpyHostToDevice, st));
            qi.h2d_bytes += (uint64_t)F * sizeof(SeqDesc);
        }
        FA_CUDA(cudaMemsetAsync(ws.counters.p, 0, CT_N * sizeof(unsigned long long), st));
        FA_CUDA(cudaEventRecord(ws.ev[1], st));
        nv.next("fa:query sketch");
        FA_TRY(ws.seq_cnt.reserve(F));
        if (ps) {
            import_sketch_kernel<<<F, 128, 0, st>>>(ps->recv, ps->per, ps->stride, ps->block, ps->first_frag, ix->s_max, ws.qhash.p,
                                                    ws.sk.seq_first.p, ws.qs.p, ws.counters.p);
            FA_CUDA(cudaGetLastError()); launches++;
        } else {
            FA_TRY(launch_sketch(st, ws.sk, F, n_tiles, k, w, P.alphabet != 4, nullptr, ws.qhash.p, 0, &launches, std::max(cmw, 1), ws.seq_cnt.p, tiles_per_frag));
            int p2 = 1; while (p2 < cmw) p2 <<= 1;
            int sort_cap = std::min(p2, 32768);
            size_t smem = (size_t)sort_cap * 4;
            if (smem > 48 * 1024) FA_CUDA(cudaFuncSetAttribute(sort_unique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sort_unique_kernel<<<F, 256, smem, st>>>(ws.qhash.p, ws.sk.seq_first.p, ws.seq_cnt.p, F, sort_cap, ix->s_max,
                                                     ws.qs.p, ws.counters.p);
            FA_CUDA(cudaGetLastError()); launches++;
        }
        FA_CUDA(cudaEventRecord(ws.ev[2], st));
        nv.next("fa:query lookup");
        // ---- lookup + seed counts ------------------------------------------------------------
        lookup_kernel<<<F, 128, 0, st>>>(ws.qhash.p, ws.sk.seq_first.p, ws.qs.p, F, ix->dir.p, ix->dir_bits, ix->ukeys.p,
                                         ix->uoff.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p);
        FA_CUDA(cudaGetLastError()); launches++;
        // (for L1 in parts: where the hits of the heavy fragments lie -- returns at once when no sampled fragment is heavy)
        const uint32_t n_chunks_ix = (uint32_t)((ix->n + (1ull << L1_SHIFT) - 1) >> L1_SHIFT);
        const bool sample_parts = ix->l1_parts != 0 && G >= 2;
        if (sample_parts) {
            FA_TRY(ws.chunk_hist.reserve(n_chunks_ix)); FA_TRY(ws.h_chunk_hist.reserve((size_t)n_chunks_ix * 4));
            FA_CUDA(cudaMemsetAsync(ws.chunk_hist.p, 0, (size_t)n_chunks_ix * 4, st));
            sample_chunks_kernel<<<std::min(F, L1_SAMPLE) * L1_SAMPLE_SLICES, 256, 0, st>>>(ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p, F,
                                                                       4096u, ix->pos_idx.p, ws.chunk_hist.p);
            FA_CUDA(cudaGetLastError()); launches++;
            FA_CUDA(cudaMemcpyAsync(ws.h_chunk_hist.p, ws.chunk_hist.p, (size_t)n_chunks_ix * 4, cudaMemcpyDeviceToHost, st));
        }
        FA_TRY(excl_scan<uint64_t>(st, ws.cub_tmp, ws.frag_seeds.p, ws.frag_seeds.p, (int64_t)F + 1, &launches));
        FA_TRY(ws.hres.reserve(128 + (size_t)B * G * 8));
        FA_TRY(ws.hfs.reserve(((size_t)F + 1) * 16));
        unsigned long long *h_ct = reinterpret_cast<unsigned long long *>(ws.hres.p);
        uint64_t *h_fs = reinterpret_cast<uint64_t *>(ws.hfs.p);              // per-fragment seed prefix (F + 1), then the slow-path prefix
        FA_CUDA(cudaMemcpyAsync(h_ct, ws.counters.p, CT_N * 8, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaMemcpyAsync(h_fs, ws.frag_seeds.p, ((size_t)F + 1) * 8, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaStreamSynchronize(st));                                   // sync 1: seed counts, max sketch, errors
        qi.d2h_bytes += ((uint64_t)F + 1) * 8;
        if (h_ct[CT_ERR] & ERR_EXCH) return FA_RETRY_PLAIN;                   // (a gathered sketch did not fit its slot: the caller maps this pass the plain way)
        if (h_ct[CT_ERR] & ERR_SORT_CAP) { set_error("a fragment produced more minimizers than the on-chip sort holds"); return FA_ERR_UNSUPPORTED; }
        if (h_ct[CT_ERR] & ERR_S_MAX) { set_error("sketch size above %d is not supported on the device path", ix->s_max); return FA_ERR_UNSUPPORTED; }
        const int max_s = (int)h_ct[CT_MAXS];
        const uint64_t S = h_fs[F];
        qi.seeds = S; qi.sketch_sum = h_ct[CT_SKETCH_SUM];
        FA_CUDA(cudaEventRecord(ws.ev[3], st));
        nv.next("fa:query seed sort");
        uint64_t C = 0;
        if (S > 0) {
            int dev_sms = 148, smem_optin = 48 * 1024;
            cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, ix->device);
            cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ix->device);
            // ---- which fragments fit the on-chip L1 (l1_fused_kernel), which take the radix sort ----
            const uint32_t n_chunks = (uint32_t)((ix->n + (1ull << L1_SHIFT) - 1) >> L1_SHIFT);
            auto *const l1_large = l1_fused_kernel<L1L_THREADS, L1L_TILE, false>;
            auto *const l1_small = l1_fused_kernel<L1S_THREADS, L1S_TILE, false>;
            auto *const l1_parts = l1_fused_kernel<L1S_THREADS, L1S_TILE, true>;
            cudaFuncAttributes l1_attr;
            FA_CUDA(cudaFuncGetAttributes(&l1_attr, l1_large));
            const size_t l1_fixed = l1_fixed_smem(n_chunks, l1_stage(L1L_THREADS, L1L_TILE)), l1_room = (size_t)smem_optin - l1_attr.sharedSizeBytes;
            uint64_t seed_cap = 0;
            if (l1_fixed + 256 < l1_room && ix->max_min_hits - 1 <= L1L_THREADS) seed_cap = ((l1_room - l1_fixed - 128) * 16 / 33) & ~31ull;   // 2 + 1/16 bytes per hit
            if (ix->l1_seed_cap >= 0) seed_cap = std::min<uint64_t>(seed_cap, (uint64_t)ix->l1_seed_cap);
            seed_cap = std::min<uint64_t>(seed_cap, 0x7FFFFFFFull);
            // the small shape takes the fragments whose hits leave room for four CTAs per SM (and whose sketch and
            // minHits fit its staging area); it is not used when hardly any fragment qualifies
            const size_t l1s_fixed = l1_fixed_smem(n_chunks, l1_stage(L1S_THREADS, L1S_TILE));
            uint64_t small_cap = 0;
            size_t l1s_smem = 0;
            for (int i = std::max(ix->l1_small_shape, 0); i < 3; i++)
                if (!l1s_smem && l1s_fixed + 8 * 1024 <= L1S_SMEM[i]) l1s_smem = L1S_SMEM[i];
            if (l1s_smem && ix->max_min_hits - 1 <= L1S_THREADS && max_s <= l1_stage(L1S_THREADS, L1S_TILE))
                small_cap = std::min<uint64_t>(seed_cap, ((l1s_smem - l1s_fixed - 128) * 16 / 33) & ~31ull);
            if (ix->l1_small_cap >= 0) small_cap = std::min<uint64_t>(small_cap, (uint64_t)ix->l1_small_cap);
            // the warp-per-fragment shape takes the fragments with a few hundred hits at most
            const uint64_t tiny_cap = std::min<uint64_t>(ix->l1_tiny_cap < 0 ? (uint64_t)L1_TINY : std::min<uint64_t>((uint64_t)ix->l1_tiny_cap, (uint64_t)L1_TINY), seed_cap);
            uint64_t max_fast = 0, max_small = 0, S_slow = 0;
            uint32_t n_slow = 0, n_small = 0, n_tiny = 0;
            uint64_t *h_fb = h_fs + F + 1;
            h_fb[0] = 0;
            for (int f = 0; f < F; f++) {
                const uint64_t nf = h_fs[f + 1] - h_fs[f];
                const bool slow = nf > seed_cap;
                if (slow) { n_slow++; S_slow += nf; }
                else if (tiny_cap && nf <= tiny_cap) n_tiny++;
                else if (small_cap && nf <= small_cap) { n_small++; max_small = std::max(max_small, nf); }
                else max_fast = std::max(max_fast, nf);
                h_fb[f + 1] = h_fb[f] + (slow ? nf : 0);
            }
            if (n_tiny * 16u < (uint32_t)F) {
                // hardly any fragment is that light: not worth a launch over all of them
                for (int f = 0; f < F && n_tiny; f++) {
                    const uint64_t nf = h_fs[f + 1] - h_fs[f];
                    if (nf > tiny_cap || nf > seed_cap) continue;
                    if (small_cap && nf <= small_cap) { n_small++; max_small = std::max(max_small, nf); } else max_fast = std::max(max_fast, nf);
                }
                n_tiny = 0;
            }
            if (n_small * 8u < (uint32_t)F - n_tiny) { max_fast = std::max(max_fast, max_small); n_small = 0; small_cap = 0; max_small = 0; }
            const uint32_t n_large = (uint32_t)F - n_slow - n_small - n_tiny;
            const uint32_t tiny_lo = n_tiny ? (uint32_t)tiny_cap + 1u : 0u;      // the CTA shapes start above the tiny class
            // ---- the large class in parts: cut at genome boundaries into shares of about equal size, each mapped by a CTA
            // of the small shape at four CTAs per SM (see split_lists_kernel).  The fewest parts whose expected share of the
            // largest fragment (with a quarter of head-room: the hits need not be spread evenly) fits.
            L1Parts pt;
            memset(&pt, 0, sizeof pt);
            uint32_t part_chunks = 0;
            uint64_t part_cap = 0;
            if (n_large && ix->l1_parts != 0 && G >= 2 && ix->max_min_hits - 1 <= L1S_THREADS && max_s <= l1_stage(L1S_THREADS, L1S_TILE)) {
                if (ix->genome_first.empty()) {               // first reference index of every genome (once per index)
                    std::vector<uint32_t> co((size_t)ix->n_contigs + 1);
                    FA_CUDA(cudaMemcpyAsync(co.data(), ix->contig_off.p, co.size() * 4, cudaMemcpyDeviceToHost, st));
                    FA_CUDA(cudaStreamSynchronize(st));
                    ix->genome_first.resize((size_t)G + 1);
                    for (uint32_t g = 0; g <= G; g++) ix->genome_first[g] = co[g == 0 ? 0 : std::min<uint64_t>((uint64_t)ix->seqs_by_genome[g - 1], ix->n_contigs)];
                    ix->genome_first[G] = (uint32_t)ix->n;
                }
                const std::vector<uint32_t> &gf = ix->genome_first;
                // cumulative hits of the sampled fragments per chunk; with no heavy fragment in the sample: by index size
                const uint32_t *ch = reinterpret_cast<const uint32_t *>(ws.h_chunk_hist.p);
                std::vector<uint64_t> cum((size_t)n_chunks + 1, 0);
                for (uint32_t c = 0; c < n_chunks; c++) cum[c + 1] = cum[c] + (sample_parts ? ch[c] : 0u);
                const bool sampled = cum[n_chunks] > 0;
                if (!sampled) for (uint32_t c = 0; c <= n_chunks; c++) cum[c] = c;
                auto mass_below = [&](uint32_t ref_idx) -> double {     // sampled hits below a reference index (linear inside a chunk)
                    const uint32_t c = std::min(ref_idx >> L1_SHIFT, n_chunks - 1);
                    return (double)cum[c] + (double)(cum[c + 1] - cum[c]) * (double)(ref_idx - (c << L1_SHIFT)) / 65536.0;
                };
                const double total = (double)cum[n_chunks];
                for (int np = 2; np <= std::min<int>(L1_PARTS_MAX, (int)G) && !pt.n_parts; np++) {
                    if (ix->l1_parts > 0 && np != std::min<int>(std::min<int>(ix->l1_parts, L1_PARTS_MAX), (int)G)) continue;
                    // genome boundaries next to the points that cut the sampled hits into np equal shares
                    uint32_t gb[L1_PARTS_MAX + 1];
                    gb[0] = 0; gb[np] = G;
                    for (int q = 1; q < np; q++) {
                        const double want = total * q / np;
                        uint32_t lo = gb[q - 1], hi = G;                 // first genome whose start holds `want` hits below it
                        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (mass_below(gf[mid]) < want) lo = mid + 1; else hi = mid; }
                        if (lo > gb[q - 1] + 1 && lo <= G && want - mass_below(gf[lo - 1]) < mass_below(gf[std::min(lo, G)]) - want) lo--;
                        gb[q] = std::min(std::max(lo, gb[q - 1]), G);
                    }
                    uint32_t span = 1;
                    double share = 0;
                    for (int q = 0; q < np; q++) {
                        if (gf[gb[q + 1]] > gf[gb[q]]) span = std::max(span, ((gf[gb[q + 1]] - 1u) >> L1_SHIFT) - (gf[gb[q]] >> L1_SHIFT) + 1u);
                        share = std::max(share, (mass_below(gf[gb[q + 1]]) - mass_below(gf[gb[q]])) / std::max(total, 1.0));
                    }
                    const size_t fixed = l1_fixed_smem(span, l1_stage(L1S_THREADS, L1S_TILE));
                    if (fixed + 8 * 1024 > L1S_SMEM[0]) continue;
                    const uint64_t cap = ((L1S_SMEM[0] - fixed - 128) * 16 / 33) & ~31ull;
                    // (an eighth of head-room: a fragment need not follow the sample exactly; what does not fit falls back)
                    if (ix->l1_parts < 0 && (double)max_fast * share * 1.125 > (double)cap) continue;
                    pt.n_parts = np; part_chunks = span; part_cap = cap;
                    for (int q = 0; q <= np; q++) pt.cfg.first[q] = gf[gb[q]];
                }
                if (pt.n_parts && ix->l1_part_cap >= 0) part_cap = std::min<uint64_t>(part_cap, (uint64_t)ix->l1_part_cap) & ~31ull;
            }
            qi.l1_parts = (uint32_t)pt.n_parts;
            qi.l1_small_fragments = n_small;
            qi.l1_tiny_fragments = n_tiny;
            qi.l1_sorted_fragments = n_slow;
            const int shift = bits_for(ix->n), fbits = bits_for((uint64_t)F);
            const uint64_t idx_mask = (1ull << shift) - 1;
            if (n_slow) {
                // ---- general path: fill + sort by (fragment, reference index) -------------------
                FA_TRY(ws.fb_seeds.reserve((size_t)F + 1));
                FA_CUDA(cudaMemcpyAsync(ws.fb_seeds.p, h_fb, ((size_t)F + 1) * 8, cudaMemcpyHostToDevice, st));
                qi.h2d_bytes += ((uint64_t)F + 1) * 8;
                FA_TRY(ws.seeds_a.reserve(S_slow)); FA_TRY(ws.seeds_b.reserve(S_slow));
                fill_seeds_kernel<<<F, 128, 0, st>>>(ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.fb_seeds.p,
                                                     ix->pos_idx.p, shift, ws.seeds_a.p);
                FA_CUDA(cudaGetLastError()); launches++;
                size_t sort_bytes = 0;
                FA_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, ws.seeds_a.p, ws.seeds_b.p, (int64_t)S_slow, 0, shift + fbits, st));
                FA_TRY(ws.cub_tmp.reserve(sort_bytes + 16));
                FA_CUDA(cub::DeviceRadixSort::SortKeys(ws.cub_tmp.p, sort_bytes, ws.seeds_a.p, ws.seeds_b.p, (int64_t)S_slow, 0, shift + fbits, st));
                launches += 2 + (shift + fbits + 7) / 8;
            }
            FA_CUDA(cudaEventRecord(ws.ev[4], st));
            nv.next("fa:query L1");
            // ---- L1 candidates -------------------------------------------------------------------
            if (n_slow < (uint32_t)F) {
                FA_TRY(ws.cand_tmp.reserve(S));
                const L1Parts no_parts{};
                if (pt.n_parts) FA_CUDA(cudaMemsetAsync(ws.frag_cands.p, 0, ((size_t)F + 1) * 4, st));   // (the parts add their counts)
                if (n_tiny) {           // fragments with [0, tiny_cap] hits: a warp each
                    l1_tiny_kernel<<<(F + L1_TINY_WARPS - 1) / L1_TINY_WARPS, 32 * L1_TINY_WARPS, 0, st>>>(
                        ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p, ix->pos_idx.p, ix->gpos.p, ix->irr.p,
                        ix->d_min_hits.p, F, L, d_near, (uint32_t)tiny_cap, ws.cand_tmp.p, ws.frag_cands.p);
                    FA_CUDA(cudaGetLastError()); launches++;
                }
                if (n_small) {          // fragments with (tiny_cap, small_cap] hits
                    const uint64_t key_cap = (max_small + 31) & ~31ull;
                    const size_t smem = l1s_fixed + 2 * key_cap + 2 * (key_cap / 32 + 8);
                    FA_CUDA(cudaFuncSetAttribute(l1_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    l1_small<<<F, L1S_THREADS, smem, st>>>(ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p,
                                                           ix->pos_idx.p, ix->gpos.p, ix->irr.p, ix->hw.p, ix->d_min_hits.p, L, d_near, n_chunks,
                                                           tiny_lo, (uint32_t)small_cap, (uint32_t)key_cap, ws.cand_tmp.p, ws.frag_cands.p, no_parts);
                    FA_CUDA(cudaGetLastError()); launches++;
                }
                if (n_large) {          // fragments with (small_cap, seed_cap] hits (all of them when the small shape is off)
                    const uint32_t large_lo = n_small ? (uint32_t)small_cap + 1u : tiny_lo;
                    if (pt.n_parts) {
                        const int np = pt.n_parts;
                        pt.s_stride = std::max(max_s, 1);
                        FA_TRY(ws.l1_split.reserve((size_t)F * pt.s_stride * (np + 1))); FA_TRY(ws.part_off.reserve((size_t)F * (np + 1)));
                        FA_TRY(ws.part_cands.reserve((size_t)F * np)); FA_TRY(ws.l1_over.reserve(F));
                        pt.split = ws.l1_split.p; pt.part_off = ws.part_off.p; pt.part_cands = ws.part_cands.p; pt.overflow = ws.l1_over.p;
                        FA_CUDA(cudaMemsetAsync(ws.l1_over.p, 0, (size_t)F * 4, st));
                        split_lists_kernel<<<F, 256, 0, st>>>(ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p, ix->pos_idx.p,
                                                             pt.cfg, np, large_lo, (uint32_t)seed_cap, pt.s_stride, ws.l1_split.p,
                                                             ws.part_off.p);
                        FA_CUDA(cudaGetLastError()); launches++;
                        const size_t psmem = L1S_SMEM[0];
                        FA_CUDA(cudaFuncSetAttribute(l1_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
                        l1_parts<<<(unsigned int)F * (unsigned int)np, L1S_THREADS, psmem, st>>>(
                            ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p, ix->pos_idx.p, ix->gpos.p, ix->irr.p, ix->hw.p,
                            ix->d_min_hits.p, L, d_near, part_chunks, large_lo, (uint32_t)seed_cap, (uint32_t)part_cap, ws.cand_tmp.p, ws.frag_cands.p, pt);
                        FA_CUDA(cudaGetLastError()); launches++;
                    }
                    // (behind the parts: only the fragments one of whose parts did not fit)
                    const uint64_t key_cap = (max_fast + 31) & ~31ull;
                    const size_t smem = l1_fixed + 2 * key_cap + 2 * (key_cap / 32 + 8);
                    FA_CUDA(cudaFuncSetAttribute(l1_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    l1_large<<<F, L1L_THREADS, smem, st>>>(ws.sk.seq_first.p, ws.qs.p, ws.hit_start.p, ws.hit_cnt.p, ws.frag_seeds.p,
                                                           ix->pos_idx.p, ix->gpos.p, ix->irr.p, ix->hw.p, ix->d_min_hits.p, L, d_near, n_chunks,
                                                           large_lo, (uint32_t)seed_cap, (uint32_t)key_cap,
                                                           ws.cand_tmp.p, ws.frag_cands.p, pt);
                    FA_CUDA(cudaGetLastError()); launches++;
                }
            }
            if (n_slow) {
                candidates_kernel<false><<<F, 256, 0, st>>>(ws.seeds_b.p, ws.fb_seeds.p, ws.qs.p, ix->d_min_hits.p, ix->ref.p,
                                                            idx_mask, L, ws.frag_cands.p, nullptr, nullptr);
                FA_CUDA(cudaGetLastError()); launches++;
            }
            work_items_kernel<<<(F + 1 + 255) / 256, 256, 0, st>>>(ws.frag_cands.p, F, ws.work_base.p);
            FA_CUDA(cudaGetLastError()); launches++;
            FA_CUDA(cudaMemsetAsync(ws.frag_cands.p + F, 0, 4, st));
            FA_TRY(excl_scan<uint32_t>(st, ws.cub_tmp, ws.frag_cands.p, ws.frag_cands.p, (int64_t)F + 1, &launches));
            FA_TRY(excl_scan<uint32_t>(st, ws.cub_tmp, ws.work_base.p, ws.work_base.p, (int64_t)F + 1, &launches));
            uint32_t *h_u = reinterpret_cast<uint32_t *>(h_ct + CT_N + 1);
            FA_CUDA(cudaMemcpyAsync(h_u, ws.frag_cands.p + F, 4, cudaMemcpyDeviceToHost, st));
            FA_CUDA(cudaMemcpyAsync(h_u + 1, ws.work_base.p + F, 4, cudaMemcpyDeviceToHost, st));
            FA_CUDA(cudaStreamSynchronize(st));                               // sync 2: candidate + work totals
            C = h_u[0];
            const uint32_t W = h_u[1];
            if (C > 0) {
                FA_TRY(ws.cands.reserve(C)); FA_TRY(ws.maps.reserve(C));
                if (n_slow < (uint32_t)F) {
                    compact_cands_kernel<<<F, 128, 0, st>>>(ws.cand_tmp.p, ws.frag_seeds.p, ws.frag_cands.p, (uint32_t)seed_cap, ws.cands.p,
                                                            n_small ? (uint32_t)small_cap + 1u : tiny_lo, pt.n_parts, ws.part_off.p, ws.part_cands.p);
                    FA_CUDA(cudaGetLastError()); launches++;
                }
                if (n_slow) {
                    candidates_kernel<true><<<F, 256, 0, st>>>(ws.seeds_b.p, ws.fb_seeds.p, ws.qs.p, ix->d_min_hits.p, ix->ref.p,
                                                               idx_mask, L, nullptr, ws.frag_cands.p, ws.cands.p);
                    FA_CUDA(cudaGetLastError()); launches++;
                }
                FA_CUDA(cudaEventRecord(ws.ev[5], st));
                nv.next("fa:query L2 prep");
                // ---- L2 -----------------------------------------------------------------------
                FA_TRY(ws.prep.reserve(C)); FA_TRY(ws.ev_off.reserve(C + 1)); FA_TRY(ws.jobs.reserve(2 * (size_t)C)); FA_TRY(ws.mid.reserve(C));
                l2_prep_kernel<<<std::min<uint32_t>((uint32_t)((C + 256) / 256), (uint32_t)dev_sms * 8u), 256, 0, st>>>(
                    ws.cands.p, ws.frag_cands.p, F, ws.qs.p, ix->ref.p, ix->hw.p, ix->hl.p, ix->fb.p, ix->contig_off.p, L, cmw, 2, w + 1,
                    reinterpret_cast<Prep *>(ws.prep.p), ws.mid.p, reinterpret_cast<unsigned long long *>(ws.ev_off.p), ws.maps.p, ws.counters.p);
                FA_CUDA(cudaGetLastError()); launches++;
                FA_TRY(excl_scan<uint64_t>(st, ws.cub_tmp, ws.ev_off.p, ws.ev_off.p, (int64_t)C + 1, &launches));
                FA_CUDA(cudaEventRecord(ws.ev[9], st));
                nv.next("fa:query L2 events");
                uint64_t *h_ev = reinterpret_cast<uint64_t *>(h_u + 2);
                FA_CUDA(cudaMemcpyAsync(h_ev, ws.ev_off.p + C, 8, cudaMemcpyDeviceToHost, st));
                FA_CUDA(cudaStreamSynchronize(st));                           // sync 3: size of the event lists
                const uint64_t n_ev = h_ev[0];
                qi.events = n_ev;
                FA_TRY(ws.events.reserve(n_ev + 64));
                const int q_cap = (std::max(max_s, 1) + L2_QPAD + 3) & ~3;
                if (n_ev > 0) {
                    {
                        const size_t smem = (size_t)q_cap * 4 + (size_t)(L2_TAB + 2) * 2 + 12 +
                                            (size_t)(EVK_THREADS / 32) * EV_WARP_BYTES;
                        FA_CUDA(cudaFuncSetAttribute(l2_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        int per_sm = 1;
                        FA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, l2_events_kernel, EVK_THREADS, smem));
                        const uint32_t grid = std::min<uint32_t>(W, (uint32_t)dev_sms * (uint32_t)std::max(per_sm, 1));
                        l2_events_kernel<<<grid, EVK_THREADS, smem, st>>>(
                            reinterpret_cast<const Prep *>(ws.prep.p), ws.mid.p, reinterpret_cast<const unsigned long long *>(ws.ev_off.p),
                            ws.frag_cands.p, ws.work_base.p, F, ws.qhash.p, ws.sk.seq_first.p, ws.qs.p, ix->ref.p, ix->hl.p,
                            l2_tab_shift(w), ws.events.p, reinterpret_cast<SlideJob *>(ws.jobs.p), ws.counters.p, q_cap);
                        FA_CUDA(cudaGetLastError()); launches++;
                    }
                    FA_CUDA(cudaEventRecord(ws.ev[10], st));
                    nv.next("fa:query L2 slide");
                    {
                        const int rows = l2_words_for(std::min(std::max(max_s, 1), EV_MAX_S) + 9) + 1;  // slack: one row below, the pivot may stray eight buckets above
                        const size_t smem = (size_t)rows * L2_THREADS * 4;
                        auto slide_fn = l2_slide_kernel;
                        if (smem > 48 * 1024) FA_CUDA(cudaFuncSetAttribute(slide_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        int per_sm = 1;
                        FA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, slide_fn, L2_THREADS, smem));
                        const uint64_t want = (C + L2_THREADS - 1) / L2_THREADS;
                        const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)dev_sms * (uint64_t)std::max(per_sm, 1));
                        slide_fn<<<grid, L2_THREADS, smem, st>>>(
                            reinterpret_cast<const SlideJob *>(ws.jobs.p), reinterpret_cast<const Prep *>(ws.prep.p), (uint32_t)C, ix->hw.p,
                            ws.events.p, ix->d_min_shared.p, ix->d_id_off.p, ix->d_identity.p, ws.maps.p, ws.counters.p, rows - 1);
                        FA_CUDA(cudaGetLastError()); launches++;
                    }
                }
                {   // exact path for what the event path did not take: long regions, very large sketches, bucket overflow
                    const size_t smem = (size_t)q_cap * 4 + (size_t)L2_FB_STATE + 64;
                    if (smem > 48 * 1024) FA_CUDA(cudaFuncSetAttribute(l2_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    const uint32_t grid = std::min<uint32_t>(W, (uint32_t)dev_sms * 4u);
                    l2_fallback_kernel<<<grid, L2_THREADS, smem, st>>>(reinterpret_cast<const Prep *>(ws.prep.p), ws.frag_cands.p, ws.work_base.p, F,
                                                                       ws.qhash.p, ws.sk.seq_first.p, ws.qs.p, ix->ref.p, (uint32_t)ix->n, cmw,
                                                                       ix->d_min_shared.p, ix->d_id_off.p, ix->d_identity.p, ws.maps.p,
                                                                       ws.counters.p, q_cap);
                    FA_CUDA(cudaGetLastError()); launches++;
                }
                FA_CUDA(cudaEventRecord(ws.ev[6], st));
                nv.next("fa:query CGI");
                // ---- CGI ----------------------------------------------------------------------
                cgi_best_kernel<<<std::min<uint32_t>((uint32_t)((C + 255) / 256), 148u * 8u), 256, 0, st>>>(
                    ws.cands.p, ws.maps.p, ws.frag_cands.p, F, ix->genome_of_seq.p, ix->bin_base.p, L - 20, ws.cells.p,
                    B > 1 ? ws.frag_q.p : nullptr, (uint32_t)n_cells, ws.counters.p);
                FA_CUDA(cudaGetLastError()); launches++;
            } else {
                FA_CUDA(cudaEventRecord(ws.ev[5], st));
                FA_CUDA(cudaEventRecord(ws.ev[6], st));
            }
        } else {
            for (int i = 4; i <= 6; i++) FA_CUDA(cudaEventRecord(ws.ev[i], st));
        }
        cgi_sum_kernel<<<(B * G * 32 + 255) / 256, 256, 0, st>>>(ws.cells.p, ix->genome_cell.p, (int)G, (int)B, (uint32_t)n_cells,
                                                                 ws.g_count.p, ws.g_identity.p);
        FA_CUDA(cudaGetLastError()); launches++;
        FA_CUDA(cudaEventRecord(ws.ev[7], st));
        nv.next("fa:query d2h");
        // ---- results back ----------------------------------------------------------------------
        int32_t *h_c = reinterpret_cast<int32_t *>(ws.hres.p + 128);
        float *h_i = reinterpret_cast<float *>(ws.hres.p + 128 + (size_t)B * G * 4);
        FA_CUDA(cudaMemcpyAsync(h_c, ws.g_count.p, (size_t)B * G * 4, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaMemcpyAsync(h_i, ws.g_identity.p, (size_t)B * G * 4, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaMemcpyAsync(h_ct, ws.counters.p, CT_N * 8, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaEventRecord(ws.ev[8], st));
        FA_CUDA(cudaStreamSynchronize(st));                                   // sync 3: results
        qi.d2h_bytes = (uint64_t)B * G * 8 + CT_N * 8 * 2 + 16;
        qi.candidates = C; qi.scanned = h_ct[CT_SCANNED]; qi.mappings = h_ct[CT_MAPPINGS]; qi.l2_fallback = h_ct[CT_REDO];
        qi.events_replayed = h_ct[CT_REPLAYED];
        h_count.assign(h_c, h_c + (size_t)B * G); h_ident.assign(h_i, h_i + (size_t)B * G);
        ws.last_cands = C; ws.last_frags = (uint64_t)F;
        float ms;
        cudaEventElapsedTime(&ms, ws.ev[0], ws.ev[1]); qi.ms_h2d = ms;
        cudaEventElapsedTime(&ms, ws.ev[1], ws.ev[2]); qi.ms_sketch = ms;
        cudaEventElapsedTime(&ms, ws.ev[2], ws.ev[3]); qi.ms_lookup = ms;
        cudaEventElapsedTime(&ms, ws.ev[3], ws.ev[4]); qi.ms_seed_sort = ms;
        cudaEventElapsedTime(&ms, ws.ev[4], ws.ev[5]); qi.ms_l1 = ms;
        cudaEventElapsedTime(&ms, ws.ev[5], ws.ev[6]); qi.ms_l2 = ms;
        if (qi.events) {
            cudaEventElapsedTime(&ms, ws.ev[5], ws.ev[9]); qi.ms_l2_prep = ms;
            cudaEventElapsedTime(&ms, ws.ev[9], ws.ev[10]); qi.ms_l2_events = ms;
            cudaEventElapsedTime(&ms, ws.ev[10], ws.ev[6]); qi.ms_l2_slide = ms;
        }
        cudaEventElapsedTime(&ms, ws.ev[6], ws.ev[7]); qi.ms_cgi = ms;
        cudaEventElapsedTime(&ms, ws.ev[7], ws.ev[8]); qi.ms_d2h = ms;
        cudaEventElapsedTime(&ms, ws.ev[0], ws.ev[8]); qi.ms_total = ms; qi.ms_batch = ms;
    }

    // ---- hit filter + sort (pyx:1121-1135), query by query ----------------------------------------
    uint64_t used = 0;
    std::vector<fa_hit> hits;
    for (uint32_t q = 0; q < B; q++) {
        hits.clear();
        if (!h_count.empty())
            for (uint32_t g = 0; g < G; g++) {
                const int32_t cnt = h_count[(size_t)q * G + g];
                if (cnt <= 0) continue;
                const uint64_t ref_len = ix->genome_len[g];
                const uint64_t min_length = std::min(q_len[q], ref_len);
                const uint64_t shared_length = (uint64_t)cnt * (uint64_t)L;
                if ((float)shared_length >= (float)min_length * P.min_fraction)
                    hits.push_back(fa_hit{(int32_t)g, cnt, (int32_t)q_frags[q], h_ident[(size_t)q * G + g]});
            }
        std::stable_sort(hits.begin(), hits.end(), [](const fa_hit &a, const fa_hit &b) { return a.identity > b.identity; });
        for (size_t i = 0; i < hits.size(); i++)
            if (out && used + i < cap) out[used + i] = hits[i];
        used += hits.size();
        hit_offsets[q + 1] = used;
    }
    qi.kernel_launches = launches;
    if (info) *info = qi;
    return FA_OK;
}

int debug_candidates(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n)
{
    std::lock_guard<std::mutex> guard(ix->mtx);
    FA_CUDA(cudaSetDevice(ix->device));
    Workspace &ws = ix->ws;
    *n = ws.last_cands;
    const uint64_t m = std::min(cap, ws.last_cands);
    if (!m) return FA_OK;
    std::vector<Cand> h(m);
    FA_CUDA(cudaMemcpy(h.data(), ws.cands.p, m * sizeof(Cand), cudaMemcpyDeviceToHost));
    std::vector<RefMini> e(2);
    for (uint64_t i = 0; i < m; i++) {      // (test hook: two small copies per candidate)
        FA_CUDA(cudaMemcpy(&e[0], ix->ref.p + h[i].hint, sizeof(RefMini), cudaMemcpyDeviceToHost));
        FA_CUDA(cudaMemcpy(&e[1], ix->ref.p + h[i].tail, sizeof(RefMini), cudaMemcpyDeviceToHost));
        rows[4 * i] = h[i].frag; rows[4 * i + 1] = (int32_t)e[0].z;
        rows[4 * i + 2] = std::max(0, (int32_t)e[0].y - ix->prm.frag_len + 1);      // rangeStartPos, computeMap.hpp:332
        rows[4 * i + 3] = (int32_t)e[1].y;                                          // rangeEndPos
    }
    return FA_OK;
}

int debug_mappings(fa_index *ix, int32_t *rows, uint64_t cap, uint64_t *n)
{
    std::lock_guard<std::mutex> guard(ix->mtx);
    FA_CUDA(cudaSetDevice(ix->device));
    Workspace &ws = ix->ws;
    const uint64_t C = ws.last_cands;
    *n = 0;
    if (!C) return FA_OK;
    std::vector<Cand> hc(C);
    std::vector<Mapping> hm(C);
    std::vector<int32_t> hs(ws.last_frags);
    FA_CUDA(cudaMemcpy(hc.data(), ws.cands.p, C * sizeof(Cand), cudaMemcpyDeviceToHost));
    FA_CUDA(cudaMemcpy(hm.data(), ws.maps.p, C * sizeof(Mapping), cudaMemcpyDeviceToHost));
    FA_CUDA(cudaMemcpy(hs.data(), ws.qs.p, ws.last_frags * 4, cudaMemcpyDeviceToHost));
    uint64_t k = 0;
    for (uint64_t i = 0; i < C; i++) {
        if (hm[i].shared < 0) continue;
        if (k < cap) {
            int32_t *r = rows + 6 * k;
            r[0] = hc[i].frag; r[1] = hm[i].seq; r[2] = hm[i].ref_start; r[3] = hm[i].shared; r[4] = hs[hc[i].frag];
            memcpy(&r[5], &hm[i].identity, 4);
        }
        k++;
    }
    *n = k;
    return FA_OK;
}

}  // namespace fa
