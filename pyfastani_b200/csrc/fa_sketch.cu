// fa_sketch.cu -- K1: byte normalisation + reverse complement + MurmurHash3 k-mer hashing +
// sliding-window minimizer winnowing, one tile of SK_TILE k-mer positions per CTA.
//
// Replaces, for both reference contigs and query fragments:
//   copy_upper / reverse_complement    src/pyfastani/_sequtils/{sse2.c:4-21, sequtils.cpp:68-79, complement.h:5-26}
//   _read_nucl / _add_minimizers_nucl  src/pyfastani/_fastani.pyx:116-222
//   getHash / MurmurHash3_x64_128      FA/map/include/commonFunc.hpp:71-81, FA/common/murmur3.h:226-303
//
// The reference's deque algorithm is sequential; here every position is independent
// (SURVEY.md Appendix A.3): key_i = (min(hf,hb) << 32) | ~i for non-symmetric k-mers, +inf
// for symmetric ones (hf == hb, skipped entirely by pyx:202); the window minimum of the keys
// picks the smallest hash and, on ties, the right-most position (the deque pops `>=`,
// pyx:211); a minimizer is emitted at window i-w+1 when the selected position differs from
// the one selected at the previous non-symmetric position.  Output order (sequence-major,
// position-minor) is kept with a single-pass decoupled look-back scan across tiles.
//
// Roofline: ALU bound by construction (two 128-bit Murmur finalisations per base, ~300
// integer instructions) -- about 2 B/base of compulsory HBM traffic.
#include "fa_common.cuh"

#include <algorithm>

namespace fa {

namespace {

constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_MASK = (1ull << 62) - 1;
constexpr unsigned long long KEY_INF = ~0ull;

__constant__ uint8_t c_comp[128];

// 64-bit rotate by a constant 0 < r < 64, r != 32, as two funnel shifts (the shift / or form compiles to six instructions)
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    uint32_t nlo, nhi;
    if (r < 32) { nhi = __funnelshift_l(lo, hi, r); nlo = __funnelshift_l(hi, lo, r); }
    else { nhi = __funnelshift_l(hi, lo, r - 32); nlo = __funnelshift_l(lo, hi, r - 32); }
    return ((uint64_t)nhi << 32) | nlo;
}

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

constexpr uint64_t MC1 = 0x87c37b91114253d5ULL, MC2 = 0x4cf5ad432745937fULL;

__device__ __forceinline__ void mm_block(uint64_t &h1, uint64_t &h2, uint64_t k1, uint64_t k2)
{
    k1 *= MC1; k1 = rotl64(k1, 31); k1 *= MC2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= MC2; k2 = rotl64(k2, 33); k2 *= MC1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
}

__device__ __forceinline__ uint32_t mm_finish(uint64_t h1, uint64_t h2, int len)
{
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return (uint32_t)h1;      // getHash keeps the first four bytes of the 128-bit digest
}

// k == 16: exactly one block, no tail.
__device__ __forceinline__ uint32_t murmur16(uint64_t k1, uint64_t k2)
{
    uint64_t h1 = 42, h2 = 42;
    mm_block(h1, h2, k1, k2);
    return mm_finish(h1, h2, 16);
}

// Any k: bytes come from shared memory; `step` is +1 (forward strand, fwd bytes) or -1
// (reverse strand: complement bytes read right to left).
__device__ __forceinline__ uint32_t murmur_any(const uint8_t *s, int first, int step, int len)
{
    uint64_t h1 = 42, h2 = 42;
    int nblocks = len >> 4, p = first;
    for (int b = 0; b < nblocks; b++) {
        uint64_t k1 = 0, k2 = 0;
#pragma unroll
        for (int m = 0; m < 8; m++) { k1 |= (uint64_t)s[p] << (8 * m); p += step; }
#pragma unroll
        for (int m = 0; m < 8; m++) { k2 |= (uint64_t)s[p] << (8 * m); p += step; }
        mm_block(h1, h2, k1, k2);
    }
    int t = len & 15;
    if (t) {
        uint64_t k1 = 0, k2 = 0;
        for (int m = 0; m < t; m++) {
            uint64_t v = s[p]; p += step;
            if (m < 8) k1 |= v << (8 * m); else k2 |= v << (8 * (m - 8));
        }
        if (t > 8) { k2 *= MC2; k2 = rotl64(k2, 33); k2 *= MC1; h2 ^= k2; }
        k1 *= MC1; k1 = rotl64(k1, 31); k1 *= MC2; h1 ^= k1;
    }
    return mm_finish(h1, h2, len);
}

__device__ __forceinline__ uint8_t upper_c(uint8_t b) { return (b >= 'a' && b <= 'z') ? (uint8_t)(b - 32) : b; }

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <bool K16>
__global__ void __launch_bounds__(SK_THREADS)
sketch_kernel(const uint8_t *__restrict__ bytes, const SeqDesc *__restrict__ seqs, int n_seqs, int n_tiles,
              int k, int w, int fwd_only, int nb_cap, int nk_cap,
              unsigned long long *status, unsigned long long *counters, uint64_t *seq_first,
              RefMini *out_ref, uint32_t *out_hash, uint64_t out_base, int slot_cap, uint32_t *seq_cnt, int uniform_tiles)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *sf = smem;                                        // normalised forward bytes
    uint8_t *sc = sf + nb_cap;                                 // their complements
    unsigned long long *A = reinterpret_cast<unsigned long long *>(sc + nb_cap);
    unsigned long long *B = A + nk_cap;
    uint32_t *vbits = reinterpret_cast<uint32_t *>(B + nk_cap);
    __shared__ uint8_t s_comp[128];
    __shared__ int s_tile;
    __shared__ uint32_t s_warp[SK_THREADS / 32];
    __shared__ unsigned long long s_excl;

    const int tid = threadIdx.x;
    if (tid == 0) s_tile = (int)atomicAdd(&counters[0], 1ull);  // ticket: lower tiles are already running
    if (tid < 128) s_comp[tid] = c_comp[tid];
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) return;

    // tile -> sequence (last descriptor whose first tile is <= tile).  Query fragments all have the same number of tiles:
    // a division instead of a chain of fifteen dependent loads at the start of every CTA (a fifth of the kernel's samples)
    int lo_s = 0, hi_s = n_seqs - 1;
    if (uniform_tiles > 0) lo_s = tile / uniform_tiles;
    else
        while (lo_s < hi_s) {
            int mid = (lo_s + hi_s + 1) >> 1;
            if (seqs[mid].tile0 <= tile) lo_s = mid; else hi_s = mid - 1;
        }
    const SeqDesc sd = seqs[lo_s];
    const int len = sd.len;
    const int nk = len - k + 1;                                // number of k-mers
    const int t0 = (tile - sd.tile0) * SK_TILE;                // first k-mer position owned by this tile
    const int halo = 2 * w - 2;
    const int lo = max(0, t0 - halo);                          // first k-mer whose key is needed
    const int hi = min(nk, t0 + SK_TILE);                      // one past the last owned k-mer
    const int nkeys = hi - lo;
    const int nbytes = nkeys + k - 1;

    // ---- 1. stage + normalise bytes [lo, lo + nbytes) with 16-byte loads ------------------
    {
        const uint64_t g0 = sd.off + (uint64_t)lo;             // sd.off is 16-byte aligned
        const uint64_t a0 = g0 & ~15ull;
        const int shift = (int)(g0 - a0);                      // sf[j] holds byte a0 + j
        const int n16 = (shift + nbytes + 15) >> 4;
        for (int v = tid; v < n16; v += SK_THREADS) {
            uint4 q = *reinterpret_cast<const uint4 *>(bytes + a0 + 16ull * v);
            uint32_t wds[4] = {q.x, q.y, q.z, q.w}, fo[4], co[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t f = 0, c = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    uint8_t raw = (uint8_t)(wds[j] >> (8 * b));
                    int p = lo - shift + 16 * v + 4 * j + b;   // position in the sequence
                    uint8_t u = raw;
                    if (!sd.raw) {
                        // sse2_copy_upper on each 2048-byte block (pyx:135-145): `& ~0x20` on the
                        // full 16-byte chunks, toupper on the block's tail
                        int blk = p & ~2047;
                        int blen = min(2048, len - blk);
                        u = ((p - blk) < (blen & ~15)) ? (uint8_t)(raw & 0xDF) : upper_c(raw);
                    }
                    f |= (uint32_t)u << (8 * b);
                    c |= (uint32_t)s_comp[u & 0x7F] << (8 * b);
                }
                fo[j] = f; co[j] = c;
            }
            reinterpret_cast<uint4 *>(sf)[v] = make_uint4(fo[0], fo[1], fo[2], fo[3]);
            reinterpret_cast<uint4 *>(sc)[v] = make_uint4(co[0], co[1], co[2], co[3]);
        }
        for (int v = tid; v < (nk_cap + 31) / 32; v += SK_THREADS) vbits[v] = 0;
        __syncthreads();
        sf += shift; sc += shift;                              // now sf[j] = byte at position lo + j
    }

    // ---- 2. canonical hash keys ------------------------------------------------------------
    {
        const int per = (nkeys + SK_THREADS - 1) / SK_THREADS;
        const int x0 = tid * per, x1 = min(nkeys, x0 + per);
        uint64_t f1 = 0, f2 = 0, r1 = 0, r2 = 0;
        if (K16 && x0 < x1) {
#pragma unroll
            for (int m = 0; m < 8; m++) {
                f1 |= (uint64_t)sf[x0 + m] << (8 * m);
                f2 |= (uint64_t)sf[x0 + 8 + m] << (8 * m);
                r1 |= (uint64_t)sc[x0 + 15 - m] << (8 * m);
                r2 |= (uint64_t)sc[x0 + 7 - m] << (8 * m);
            }
        }
        uint32_t vword = 0;
        int vbase = x0 & ~31;
        for (int x = x0; x < x1; x++) {
            uint32_t hf, hb = 0xFFFFFFFFu;
            if (K16) {
                hf = murmur16(f1, f2);
                if (!fwd_only) hb = murmur16(r1, r2);
                uint64_t nf = sf[x + 16], nc = sc[x + 16];    // (one byte past the tile on the last step: staged padding)
                f1 = (f1 >> 8) | (f2 << 56); f2 = (f2 >> 8) | (nf << 56);
                r2 = (r2 << 8) | (r1 >> 56); r1 = (r1 << 8) | nc;
            } else {
                hf = murmur_any(sf, x, 1, k);
                if (!fwd_only) hb = murmur_any(sc, x + k - 1, -1, k);
            }
            unsigned long long key = KEY_INF;
            // symmetric k-mers are skipped, pyx:202; proteins hash the forward strand only and skip nothing, pyx:290
            if (hf != hb || fwd_only) {
                key = ((unsigned long long)min(hf, hb) << 32) | (0xFFFFFFFFu - (uint32_t)x);
                if ((x & ~31) != vbase) { if (vword) atomicOr(&vbits[vbase >> 5], vword); vword = 0; vbase = x & ~31; }
                vword |= 1u << (x & 31);
            }
            A[x] = key;
        }
        if (vword) atomicOr(&vbits[vbase >> 5], vword);
    }
    __syncthreads();

    // ---- 3. sliding-window minimum by doubling: m_s[x] = min(key[x-s+1 .. x]) --------------
    // (Tried in round 2: per-thread suffix / prefix minima over the block of keys a thread hashed -- three or four minima
    // per key and one barrier instead of log2(w) + 1 passes.  Fewer instructions, but two serial chains of dependent
    // 64-bit minima per thread: query sketching 0.090 -> 0.097 ms per 4.5 Mbp genome.  The passes below are fully parallel.)
    unsigned long long *src = A, *dst = B;
    int span = 1;
    while (span * 2 <= w) {
        for (int x = tid; x < nkeys; x += SK_THREADS) {
            unsigned long long a = src[x], b = x >= span ? src[x - span] : KEY_INF;
            dst[x] = a < b ? a : b;
        }
        __syncthreads();
        unsigned long long *t = src; src = dst; dst = t;
        span *= 2;
    }
    if (span < w) {
        const int d = w - span;
        for (int x = tid; x < nkeys; x += SK_THREADS) {
            unsigned long long a = src[x], b = x >= d ? src[x - d] : KEY_INF;
            dst[x] = a < b ? a : b;
        }
        __syncthreads();
        unsigned long long *t = src; src = dst; dst = t;
    }
    const unsigned long long *mw = src;                         // mw[x]: min key of the window ending at lo + x

    // ---- 4. emission flags for the owned positions ------------------------------------------
    uint32_t e_hash[SK_PER_THREAD];
    uint32_t e_mask = 0;
    int e_cnt = 0;
    const int i0 = t0 + tid * SK_PER_THREAD;
#pragma unroll
    for (int j = 0; j < SK_PER_THREAD; j++) {
        const int i = i0 + j;
        e_hash[j] = 0;
        if (i < hi && i >= w - 1) {
            const int x = i - lo;
            if ((vbits[x >> 5] >> (x & 31)) & 1u) {
                const unsigned long long mk = mw[x];
                // previous non-symmetric position that owns a full window; if it lies w or more
                // positions back its minimum is outside this window, so the selection changed
                bool emit = true;
                const int stop = max(max(w - 1, i - w + 1), lo) - lo;
                for (int y = x - 1; y >= stop; y--) {
                    if ((vbits[y >> 5] >> (y & 31)) & 1u) { emit = (mw[y] != mk); break; }
                }
                if (emit) { e_hash[j] = (uint32_t)(mk >> 32); e_mask |= 1u << j; e_cnt++; }
            }
        }
    }

    // ---- 5. block scan + decoupled look-back across tiles -----------------------------------
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t incl = (uint32_t)e_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0, agg = 0;
#pragma unroll
    for (int q = 0; q < SK_THREADS / 32; q++) { if (q < wid) wbase += s_warp[q]; agg += s_warp[q]; }
    const uint32_t local_off = wbase + incl - (uint32_t)e_cnt;

    // slot_cap > 0 (query fragments): every sequence writes into its own slot of slot_cap entries, so a tile needs the
    // counts of the earlier tiles of ITS sequence only -- the first tile of a sequence publishes a prefix at once and
    // the look-back of the others ends there, two tiles back at most.  slot_cap == 0 (reference genomes): one
    // (sequence, position)-ordered array, look-back over all tiles.
    const bool local = slot_cap > 0;
    if (wid == 0) {
        unsigned long long excl = 0;
        if (tile == 0 || (local && t0 == 0)) {
            if (lane == 0) atomicExch(&status[tile], ST_PREFIX | (unsigned long long)agg);
        } else {
            if (lane == 0) atomicExch(&status[tile], ST_AGG | (unsigned long long)agg);
            int p = tile - 1;
            while (true) {
                const int idx = p - lane;
                unsigned long long v = idx >= 0 ? ld_status(&status[idx]) : ST_PREFIX;
                const unsigned ready = __ballot_sync(0xFFFFFFFFu, (v >> 62) != 0);
                const unsigned isp = __ballot_sync(0xFFFFFFFFu, (v >> 62) == 2);
                const int first_nr = (~ready) ? __ffs(~ready) - 1 : 32;
                const int first_p = isp ? __ffs(isp) - 1 : 32;
                const int take = first_p < first_nr ? first_p + 1 : first_nr;   // lanes [0, take) are usable
                unsigned long long part = lane < take ? (v & ST_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
                excl += part;
                if (first_p < first_nr) break;
                if (take == 0) __nanosleep(64);
                p -= take;
            }
            if (lane == 0) atomicExch(&status[tile], ST_PREFIX | (excl + agg));
        }
        if (lane == 0) {
            const unsigned long long slot = local ? (unsigned long long)lo_s * (unsigned long long)slot_cap : 0ull;
            s_excl = slot + excl;
            if (t0 == 0) seq_first[lo_s] = slot + excl;
            if (local && hi == nk) seq_cnt[lo_s] = (uint32_t)(excl + agg);
            if (tile == n_tiles - 1) counters[1] = excl + agg;
        }
    }
    __syncthreads();

    // ---- 6. write ---------------------------------------------------------------------------
    uint64_t o = out_base + s_excl + local_off;
#pragma unroll
    for (int j = 0; j < SK_PER_THREAD; j++) {
        if (e_mask & (1u << j)) {
            const int i = i0 + j;
            if (out_ref) out_ref[o] = make_uint4(e_hash[j], (uint32_t)(i - w + 1), (uint32_t)sd.id, 0u);
            else out_hash[o] = e_hash[j];
            o++;
        }
    }
}

// ---- first-window quirk (reference sketches only) -------------------------------------------
// pyx:219-222 compares the deque front, whose wpos is still the placeholder 0, with the last
// emitted triple: after a minimizer emitted at window 0, later fronts with the same hash are
// suppressed until one with a different hash is emitted.  The kernel above ignores this; the
// pass below finds the (rare) affected runs so they can be compacted away.
__global__ void quirk_find_kernel(const RefMini *ref, const unsigned long long *counters, const uint64_t *seq_first,
                                  const SeqDesc *seqs, int n_seqs, ulonglong2 *drops, unsigned long long *n_drops,
                                  unsigned int drop_cap)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seqs) return;
    const uint64_t n = counters[1];
    uint64_t f = seq_first[s];
    if (f >= n) return;
    RefMini e = ref[f];
    if ((int)e.z != seqs[s].id || e.y != 0u) return;
    uint64_t r = f + 1;
    while (r < n) {
        RefMini q = ref[r];
        if ((int)q.z != seqs[s].id || q.x != e.x) break;
        r++;
    }
    if (r > f + 1) {
        unsigned int slot = (unsigned int)atomicAdd(n_drops, 1ull);
        if (slot < drop_cap) drops[slot] = make_ulonglong2(f + 1, r);
    }
}

// drops are sorted by start; removed_before[i] = elements dropped by ranges [0, i)
__global__ void quirk_compact_kernel(const RefMini *src, RefMini *dst, uint64_t first, uint64_t n,
                                     const ulonglong2 *drops, const uint64_t *removed_before, int n_drops)
{
    uint64_t j = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int lo = 0, hi = n_drops;                      // number of ranges with start <= j
    while (lo < hi) { int mid = (lo + hi) >> 1; if (drops[mid].x <= j) lo = mid + 1; else hi = mid; }
    uint64_t removed = 0;
    if (lo > 0) {
        if (j < drops[lo - 1].y) return;           // inside a dropped run
        removed = removed_before[lo];
    }
    dst[j - removed] = src[j];
}

}  // namespace

static int init_tables()
{
    static bool done = false;     // per device context would be stricter; one device per process here
    static int device = -1;
    int cur = 0;
    FA_CUDA(cudaGetDevice(&cur));
    if (done && device == cur) return FA_OK;
    // complement.h:5-26: identity except the IUPAC pairs (both cases); entries 0x0b and 0x1b
    // hold 0x00 and 0x01 in the reference's table.
    uint8_t t[128];
    for (int i = 0; i < 128; i++) t[i] = (uint8_t)i;
    const char *pairs = "ATCGBVDHKMRY";
    for (int i = 0; i < 12; i++) {
        t[(int)pairs[i]] = (uint8_t)pairs[i ^ 1];
        t[(int)pairs[i] + 32] = (uint8_t)(pairs[i ^ 1] + 32);
    }
    t[0x0b] = 0x00; t[0x1b] = 0x01;
    FA_CUDA(cudaMemcpyToSymbol(c_comp, t, 128));
    done = true; device = cur;
    return FA_OK;
}

int launch_sketch(cudaStream_t st, const SketchScratch &sc, int n_seqs, int n_tiles, int k, int w, int fwd_only,
                  RefMini *out_ref, uint32_t *out_hash, uint64_t out_base, int *launches, int slot_cap, uint32_t *seq_cnt, int uniform_tiles)
{
    FA_TRY(init_tables());
    const int halo = 2 * w - 2;
    const int nk_cap = SK_TILE + halo;
    const int nb_cap = ((nk_cap + k - 1 + 15 + 16 + 15) / 16) * 16;
    const size_t smem = 2 * (size_t)nb_cap + 2 * (size_t)nk_cap * 8 + ((nk_cap + 31) / 32) * 4 + 16;
    if (smem > 200 * 1024) { set_error("window size %d / k %d need %zu bytes of shared memory per tile", w, k, smem); return FA_ERR_UNSUPPORTED; }
    FA_CUDA(cudaMemsetAsync(sc.tile_status.p, 0, (size_t)n_tiles * sizeof(unsigned long long), st));
    FA_CUDA(cudaMemsetAsync(sc.counters.p, 0, 4 * sizeof(unsigned long long), st));
    FA_CUDA(cudaMemsetAsync(sc.seq_first.p, 0xFF, (size_t)n_seqs * sizeof(uint64_t), st));
    auto kern = (k == 16) ? sketch_kernel<true> : sketch_kernel<false>;
    if (smem > 48 * 1024) FA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, SK_THREADS, smem, st>>>(sc.bytes.p, sc.seqs.p, n_seqs, n_tiles, k, w, fwd_only, nb_cap, nk_cap,
                                            sc.tile_status.p, sc.counters.p, sc.seq_first.p, out_ref, out_hash, out_base,
                                            seq_cnt ? slot_cap : 0, seq_cnt, uniform_tiles);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    return FA_OK;
}

int launch_quirk_find(cudaStream_t st, const SketchScratch &sc, int n_seqs, const RefMini *batch_ref, int *launches)
{
    if (n_seqs == 0) return FA_OK;
    quirk_find_kernel<<<(n_seqs + 127) / 128, 128, 0, st>>>(batch_ref, sc.counters.p, sc.seq_first.p, sc.seqs.p, n_seqs,
                                                            sc.drops.p, sc.counters.p + 2, (unsigned int)n_seqs);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    return FA_OK;
}

int quirk_compact(cudaStream_t st, const SketchScratch &sc, unsigned int nd, RefMini *batch_ref, uint64_t n,
                  uint64_t *n_out, int *launches)
{
    std::vector<ulonglong2> drops(nd);
    FA_CUDA(cudaMemcpy(drops.data(), sc.drops.p, nd * sizeof(ulonglong2), cudaMemcpyDeviceToHost));
    std::sort(drops.begin(), drops.end(), [](const ulonglong2 &a, const ulonglong2 &b) { return a.x < b.x; });
    std::vector<uint64_t> before(nd + 1, 0);
    for (unsigned int i = 0; i < nd; i++) before[i + 1] = before[i] + (drops[i].y - drops[i].x);
    uint64_t *d_before = nullptr;
    RefMini *tmp = nullptr;
    const uint64_t first = drops[0].x;
    FA_CUDA(cudaMalloc((void **)&d_before, (nd + 1) * sizeof(uint64_t)));
    FA_CUDA(cudaMalloc((void **)&tmp, (size_t)n * sizeof(RefMini)));
    FA_CUDA(cudaMemcpyAsync(sc.drops.p, drops.data(), nd * sizeof(ulonglong2), cudaMemcpyHostToDevice, st));
    FA_CUDA(cudaMemcpyAsync(d_before, before.data(), (nd + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const uint64_t m = n - first;
    quirk_compact_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, st>>>(batch_ref, tmp, first, n, sc.drops.p, d_before, (int)nd);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    const uint64_t kept = m - before[nd];
    FA_CUDA(cudaMemcpyAsync(batch_ref + first, tmp + first, (size_t)kept * sizeof(RefMini), cudaMemcpyDeviceToDevice, st));
    FA_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_before); cudaFree(tmp);
    *n_out = n - before[nd];
    return FA_OK;
}

}  // namespace fa
