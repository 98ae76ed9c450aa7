// fa_index.cu -- K2: reference sketch index build on the GPU.
//
// Replaces skch::Sketch::index (FA/map/include/winSketch.hpp:177-189: an
// unordered_map<hash, vector<(seqId, wpos)>> filled one push_back at a time), the lookup
// side of Sketch::searchIndex (:255-266) and reviseRefIdToGenomeId
// (FA/cgi/include/computeCoreIdentity.hpp:31-42).  computeFreqHist (:195-244) leaves
// freqThreshold at INT_MAX because percentageThreshold is 0 (:52), so no histogram is built.
//
// Layout: a stable LSD radix sort of (hash, position index) gives
// `pos_idx` grouped by hash with insertion order kept inside each group; run-length encoding
// gives the CSR (ukeys, uoff); a directory over the top bits of the hash narrows each lookup
// to a few keys.  The sort is cub::DeviceRadixSort (library code); the other kernels are ours.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "fa_internal.cuh"

namespace fa {

namespace {

__global__ void extract_keys_kernel(const RefMini *ref, uint64_t n, uint32_t *keys, uint32_t *vals)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ref[i].x; vals[i] = (uint32_t)i; }
}

// dir[b] = first unique key whose top `bits` bits are >= b
__global__ void directory_kernel(const uint32_t *ukeys, uint32_t n_unique, int bits, uint32_t *dir)
{
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nb = 1u << bits;
    if (b > nb) return;
    if (b == nb) { dir[b] = n_unique; return; }
    uint32_t target = b << (32 - bits);
    uint32_t lo = 0, hi = n_unique;
    while (lo < hi) { uint32_t mid = lo + ((hi - lo) >> 1); if (ukeys[mid] < target) lo = mid + 1; else hi = mid; }
    dir[b] = lo;
}

// Neighbouring entries of one hash group that lie in the same contig are duplicates a sliding
// L2 window may hold at the same time; record their distance on both elements.
__global__ void dup_delta_kernel(const uint32_t *sorted_keys, const uint32_t *pos_idx, uint64_t n, RefMini *ref)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (p >= n) return;
    if (sorted_keys[p] != sorted_keys[p - 1]) return;
    uint32_t a = pos_idx[p - 1], b = pos_idx[p];
    if (ref[a].z != ref[b].z) return;
    uint32_t d = b - a;                        // stable sort: b > a
    if (d > 65535u) return;
    atomicOr(&ref[b].w, d);                    // previous duplicate of b
    atomicOr(&ref[a].w, d << 16);              // next duplicate of a
}

__global__ void make_hw_kernel(const RefMini *ref, uint64_t n, uint2 *hw)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RefMini e = ref[i];
    hw[i] = make_uint2(e.x, e.y | (e.w ? 0x80000000u : 0u));
}

// gpos: a 32-bit running coordinate over the position-ordered minimizers in which two elements are
// closer than a fragment exactly when they lie in the same contig less than frag_len apart.  Every
// step adds min(wpos difference, frag_len) (frag_len across a contig border), so the test of
// computeL1CandidateRegions (computeMap.hpp:325-330: same seqId, wpos distance < fragment length)
// becomes one unsigned subtraction of two 4-byte gathers (fa_map.cu l1_fused_kernel).  Sums wrap
// mod 2^32; differences of elements fewer than frag_len indices apart never do.
// Order of the slide events, independent of the query.  L2 slides a window of cmw = fragLen - (w - 1) - (k - 1)
// positions over a region of one contig (computeMap.hpp:415-488, MIIteratorL2.hpp:54-96): element j enters the
// window at time wpos[j] - (cmw - 1) and leaves when the begin reaches element j + 1, at time wpos[j + 1].  Merging
// the two event streams by time only asks, per element, how many elements of the contig lie cmw - 1 behind / ahead:
//   lag(j)  = j - A(j),  A(j) = first index with wpos > wpos[j] - (cmw - 1): the elements at or before A(j) have
//             left the window when j enters (their deletes precede the insert of j)
//   lead(j) = B(j) - j,  B(j) = first index with wpos >= wpos[j + 1] + (cmw - 1): the elements before B(j) have
//             entered when j leaves; twin(j): element B(j) enters at exactly that time (same group, after the delete)
// packed as lag (15 bits, saturating) | twin << 15 | lead << 16 (15 bits, saturating) | has-duplicate-nearby << 31
// and stored next to the hash.  Saturated values only occur in windows of more than 32 767 minimizers, which the
// event path of L2 (<= 1024 per region) never takes.
// The same with a whole fragment length L for the L1 regions: fb[j] = (j - P(j)) | (Q(j) - j) << 16 with P(j) the
// first index with wpos >= wpos[j] - (L - 1) (where a region that seed j opens begins) and Q(j) the first index
// with wpos >= wpos[j] + L (where a region that seed j closes ends).  At most L minimizers lie in L positions and
// L <= 32767, so both fit 16 bits.
__global__ void slide_order_kernel(const RefMini *ref, const uint32_t *contig_off, uint64_t n, int cmw1, int frag_len, uint2 *hl,
                                   uint32_t *fb)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const RefMini e = ref[j];
    const uint32_t c0 = contig_off[e.z], c1 = contig_off[e.z + 1];
    const int pos = (int)e.y;
    // A(j): positions grow strictly inside a contig, so the answer is at most cmw1 elements back
    uint32_t lo = (uint64_t)cmw1 < j - c0 ? (uint32_t)j - (uint32_t)cmw1 : c0, hi = (uint32_t)j;       // answer in [lo, j]
    const int ta = pos - cmw1;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((int)ref[mid].y <= ta) lo = mid + 1; else hi = mid; }
    const uint32_t lag = min((uint32_t)j - lo, 32767u);
    uint32_t lead = 0, twin = 0;
    if (j + 1 < c1) {
        const int tb = (int)ref[j + 1].y + cmw1;
        uint32_t l = (uint32_t)j + 1, h = (uint64_t)cmw1 + 1 < c1 - (j + 1) ? (uint32_t)j + 2 + (uint32_t)cmw1 : c1;   // answer in [j + 1, h]
        while (l < h) { const uint32_t mid = l + ((h - l) >> 1); if ((int)ref[mid].y < tb) l = mid + 1; else h = mid; }
        lead = min(l - (uint32_t)j, 32767u);
        twin = (l < c1 && (int)ref[l].y == tb) ? 1u : 0u;
    }
    hl[j] = make_uint2(e.x, lag | (twin << 15) | (lead << 16) | (e.w ? 0x80000000u : 0u));
    {
        const uint32_t Lm1 = (uint32_t)(frag_len - 1);
        uint32_t l = (uint64_t)Lm1 < j - c0 ? (uint32_t)j - Lm1 : c0, h = (uint32_t)j;                       // P(j) in [l, j]
        const int tp = pos - (int)Lm1;
        while (l < h) { const uint32_t mid = l + ((h - l) >> 1); if ((int)ref[mid].y < tp) l = mid + 1; else h = mid; }
        const uint32_t back = (uint32_t)j - l;
        uint32_t l2 = (uint32_t)j + 1, h2 = (uint64_t)frag_len < c1 - j ? (uint32_t)j + (uint32_t)frag_len : c1;   // Q(j) in [j + 1, h2]
        const int tq = pos + frag_len;
        while (l2 < h2) { const uint32_t mid = l2 + ((h2 - l2) >> 1); if ((int)ref[mid].y < tq) l2 = mid + 1; else h2 = mid; }
        fb[j] = min(back, 65535u) | (min(l2 - (uint32_t)j, 65535u) << 16);
    }
}

__global__ void gpos_delta_kernel(const RefMini *ref, uint64_t n, uint32_t frag_len, uint32_t window, uint32_t *gpos, uint32_t *irr)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t d = 0;
    if (i > 0) {
        const RefMini a = ref[i - 1], b = ref[i];
        d = frag_len;
        if (a.z == b.z) d = min(b.y - a.y, frag_len);          // (an unordered pair wraps to a huge value: frag_len)
    }
    gpos[i] = d;
    // Winnowing leaves at most `window` positions between neighbouring minimizers, except across contig ends and runs
    // of skipped k-mers.  Blocks of 1024 minimizers in which every step obeys that bound let the L1 kernel decide
    // "closer than a fragment" from the index distance alone (fa_map.cu l1_near); the others are marked here.
    // (A step into the first minimizer of a block also marks the block before it, so two minimizers less than 1024
    // apart are covered by the marks of the blocks they sit in.)
    const bool bad = d > window;
    const unsigned any = __ballot_sync(__activemask(), bad);
    if (bad && (any & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) atomicOr(&irr[i >> 15], 1u << ((i >> 10) & 31u));
    if (bad && (i & 1023u) == 0u && i > 0) atomicOr(&irr[(i - 1) >> 15], 1u << (((i - 1) >> 10) & 31u));
}

// contig_off[s] = first ref index with seqId >= s (contigs without minimizers get empty ranges)
__global__ void contig_off_kernel(const RefMini *ref, uint64_t n, uint32_t n_contigs, uint32_t *contig_off)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t cur = i < n ? ref[i].z : n_contigs;
    uint32_t prev = i == 0 ? 0xFFFFFFFFu : ref[i - 1].z;      // -1: everything up to cur starts here
    for (uint32_t s = prev + 1; s <= cur && s <= n_contigs; s++) contig_off[s] = (uint32_t)i;
    if (i == 0) contig_off[0] = 0;
}

// (contig, bin) cells of computeCGI's second pass (computeCoreIdentity.hpp:191, 234-255): a mapping
// starts at the mean of two minimizer positions of its contig, so bins up to the one holding the
// contig's last minimizer suffice.
__global__ void contig_bins_kernel(const RefMini *ref, const uint32_t *contig_off, uint32_t n_contigs, int bin_w, uint32_t *nbins)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_contigs) return;
    uint32_t v = 0;
    if (s < n_contigs) {
        const uint32_t lo = contig_off[s], hi = contig_off[s + 1];
        if (hi > lo) v = ref[hi - 1].y / (uint32_t)bin_w + 1;
    }
    nbins[s] = v;
}

__global__ void genome_cells_kernel(const uint32_t *bin_base, const int32_t *first_contig, uint32_t n_genomes, uint32_t *genome_cell)
{
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= n_genomes) genome_cell[g] = bin_base[first_contig[g]];
}

}  // namespace

int build_index(fa_index *ix, int *launches)
{
    cudaStream_t st = ix->st;
    const uint64_t n = ix->n;
    const uint32_t n_contigs = (uint32_t)ix->n_contigs;
    const uint32_t n_genomes = (uint32_t)ix->seqs_by_genome.size();
    if (n >= 0xFFFFFFF0ull) { set_error("index of %llu minimizers exceeds the 32-bit position index", (unsigned long long)n); return FA_ERR_UNSUPPORTED; }
    struct Ev {                                  // build timers: whole build, and the radix sort inside it
        cudaEvent_t e[4] = {};
        ~Ev() { for (auto &x : e) if (x) cudaEventDestroy(x); }
    } ev;
    for (auto &x : ev.e) FA_CUDA(cudaEventCreate(&x));
    FA_CUDA(cudaEventRecord(ev.e[0], st));
    // FA_BUILD_TRACE=1: host-clock phase times on stderr (each mark drains the stream first)
    const bool trace = getenv("FA_BUILD_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    NvtxStages nv;
    nv.next("fa:index build");
    auto mark = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[fa build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };

    // ---- host-side tables: contig -> genome, (contig, bin) cells ---------------------------
    std::vector<int32_t> genome_of(n_contigs ? n_contigs : 1, 0);
    {
        uint32_t g = 0;
        for (uint32_t s = 0; s < n_contigs; s++) {
            while (g < n_genomes && (uint32_t)ix->seqs_by_genome[g] <= s) g++;   // upper_bound, computeCoreIdentity.hpp:36-40
            genome_of[s] = (int32_t)g;
        }
    }
    // first contig of each genome (+ the total), for the per-genome cell ranges
    std::vector<int32_t> first_contig(n_genomes + 1, 0);
    for (uint32_t g = 0; g < n_genomes; g++) first_contig[g + 1] = std::min<int32_t>(ix->seqs_by_genome[g], (int32_t)n_contigs);
    first_contig[n_genomes] = (int32_t)n_contigs;
    TmpBuf<int32_t> d_first;
    FA_TRY(d_first.reserve(first_contig.size()));
    FA_TRY(ix->genome_of_seq.reserve(genome_of.size()));
    FA_TRY(ix->bin_base.reserve((size_t)n_contigs + 1));
    FA_TRY(ix->genome_cell.reserve((size_t)n_genomes + 1));
    FA_CUDA(cudaMemcpyAsync(ix->genome_of_seq.p, genome_of.data(), genome_of.size() * 4, cudaMemcpyHostToDevice, st));
    FA_CUDA(cudaMemcpyAsync(d_first.p, first_contig.data(), first_contig.size() * 4, cudaMemcpyHostToDevice, st));
    FA_CUDA(cudaStreamSynchronize(st));       // the host vectors go out of scope

    mark("contig tables to the device");
    // ---- statistics tables -------------------------------------------------------------------
    {
        int cmw = ix->prm.frag_len - (ix->prm.window - 1) - (ix->prm.k - 1);     // windows per fragment
        int s_max = cmw < 1 ? 1 : cmw;
        if (s_max > 4096) s_max = 4096;       // larger sketches are reported as unsupported at query time
        const StatTable &t = stat_table(ix->prm.k, ix->prm.pct_identity, s_max);
        if (t.irregular_l2) {
            set_error("the identity filter is not monotone in the shared-sketch count for k=%d, identity=%g (%d sketch sizes): "
                      "not supported on the device path", ix->prm.k, (double)ix->prm.pct_identity, t.irregular_l2);
            return FA_ERR_UNSUPPORTED;
        }
        ix->s_max = s_max;
        ix->max_min_hits = 1;
        for (int32_t v : t.min_hits) ix->max_min_hits = std::max(ix->max_min_hits, (int)v);
        FA_TRY(ix->d_min_hits.reserve(t.min_hits.size()));
        FA_TRY(ix->d_min_shared.reserve(t.min_shared.size()));
        FA_TRY(ix->d_id_off.reserve(t.id_off.size()));
        FA_TRY(ix->d_identity.reserve(t.identity.size()));
        FA_CUDA(cudaMemcpyAsync(ix->d_min_hits.p, t.min_hits.data(), t.min_hits.size() * 4, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemcpyAsync(ix->d_min_shared.p, t.min_shared.data(), t.min_shared.size() * 4, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemcpyAsync(ix->d_id_off.p, t.id_off.data(), t.id_off.size() * 4, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemcpyAsync(ix->d_identity.p, t.identity.data(), t.identity.size() * 4, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaStreamSynchronize(st));
    }

    mark("statistics tables");
    FA_TRY(ix->contig_off.reserve((size_t)n_contigs + 1));
    {
        uint64_t m = n + 1;
        contig_off_kernel<<<(unsigned int)((m + 255) / 256), 256, 0, st>>>(ix->ref.p, n, n_contigs, ix->contig_off.p);
        FA_CUDA(cudaGetLastError());
        if (launches) *launches += 1;
    }
    {
        const int bin_w = ix->prm.frag_len - 20;                                   // computeCoreIdentity.hpp:191
        TmpBuf<uint8_t> tmp0;
        contig_bins_kernel<<<(n_contigs + 1 + 255) / 256, 256, 0, st>>>(ix->ref.p, ix->contig_off.p, n_contigs, bin_w, ix->bin_base.p);
        FA_CUDA(cudaGetLastError());
        size_t sb = 0;
        FA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, ix->bin_base.p, ix->bin_base.p, (int64_t)n_contigs + 1, st));
        FA_TRY(tmp0.reserve(sb + 16));
        FA_CUDA(cub::DeviceScan::ExclusiveSum(tmp0.p, sb, ix->bin_base.p, ix->bin_base.p, (int64_t)n_contigs + 1, st));
        genome_cells_kernel<<<(n_genomes + 1 + 255) / 256, 256, 0, st>>>(ix->bin_base.p, d_first.p, n_genomes, ix->genome_cell.p);
        FA_CUDA(cudaGetLastError());
        if (launches) *launches += 4;
        uint32_t cells = 0;
        FA_CUDA(cudaMemcpyAsync(&cells, ix->bin_base.p + n_contigs, 4, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaStreamSynchronize(st));
        ix->n_cells = cells;
        tmp0.release(); d_first.release();
    }
    ix->n_unique = 0;
    ix->dir_bits = 0;
    if (n == 0) {
        FA_TRY(ix->dir.reserve(2));
        FA_CUDA(cudaMemsetAsync(ix->dir.p, 0, 8, st));
        FA_TRY(ix->uoff.reserve(1));
        FA_CUDA(cudaMemsetAsync(ix->uoff.p, 0, 4, st));
        FA_CUDA(cudaStreamSynchronize(st));
        return FA_OK;
    }

    mark("contig offsets, bins, cells");
    // ---- sort (hash, index) ------------------------------------------------------------------
    // Every array of 4 or 8 bytes per minimizer is a cudaMalloc of gigabytes, and those cost more than the kernels
    // (config 2: 37 ms of kernels inside a 196 ms build).  So the transient arrays live in the index's own buffers
    // until these are filled: the sort input in `hl`, the sorted keys and the run lengths in `hw`, the untrimmed
    // unique keys in `gpos`.
    FA_TRY(ix->pos_idx.reserve(n));
    FA_TRY(ix->hw.reserve(n + 8));            // L2 reads whole 32-byte chunks and one chunk ahead
    FA_TRY(ix->hl.reserve(n + 8)); FA_TRY(ix->fb.reserve(n + 8));
    FA_TRY(ix->gpos.reserve(n));
    mark("cudaMalloc of the index arrays");
    uint32_t *keys_a = reinterpret_cast<uint32_t *>(ix->hl.p), *vals_a = keys_a + n;
    uint32_t *keys_b = reinterpret_cast<uint32_t *>(ix->hw.p), *run_len = keys_b + n;
    uint32_t *ukeys_all = ix->gpos.p;
    extract_keys_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(ix->ref.p, n, keys_a, vals_a);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    size_t tmp_bytes = 0;
    FA_CUDA(cudaEventRecord(ev.e[2], st));
    FA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, vals_a, ix->pos_idx.p, (int64_t)n, 0, 32, st));
    TmpBuf<uint8_t> tmp;
    FA_TRY(tmp.reserve(tmp_bytes + 16));
    FA_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys_a, keys_b, vals_a, ix->pos_idx.p, (int64_t)n, 0, 32, st));
    if (launches) *launches += 5;
    FA_CUDA(cudaEventRecord(ev.e[3], st));

    mark("extract + radix sort");
    // duplicate distances for the L2 sliding window
    dup_delta_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(keys_b, ix->pos_idx.p, n, ix->ref.p);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;

    // ---- CSR over unique hashes --------------------------------------------------------------
    TmpBuf<uint64_t> d_nruns;
    FA_TRY(d_nruns.reserve(1));
    size_t rle_bytes = 0;
    FA_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, rle_bytes, keys_b, ukeys_all, run_len, d_nruns.p, (int64_t)n, st));
    FA_TRY(tmp.reserve(rle_bytes + 16));
    FA_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, rle_bytes, keys_b, ukeys_all, run_len, d_nruns.p, (int64_t)n, st));
    if (launches) *launches += 2;
    uint64_t n_unique = 0;
    FA_CUDA(cudaMemcpyAsync(&n_unique, d_nruns.p, 8, cudaMemcpyDeviceToHost, st));
    FA_CUDA(cudaStreamSynchronize(st));
    ix->n_unique = n_unique;
    mark("dup distances + run lengths");
    FA_TRY(ix->uoff.reserve(n_unique + 1));
    FA_TRY(ix->ukeys.reserve(n_unique ? n_unique : 1));
    size_t scan_bytes = 0;
    FA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, run_len, ix->uoff.p, (int64_t)n_unique, st));
    FA_TRY(tmp.reserve(scan_bytes + 16));
    FA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, run_len, ix->uoff.p, (int64_t)n_unique, st));
    if (launches) *launches += 2;
    const uint32_t n32 = (uint32_t)n;
    FA_CUDA(cudaMemcpyAsync(ix->uoff.p + n_unique, &n32, 4, cudaMemcpyHostToDevice, st));
    FA_CUDA(cudaMemcpyAsync(ix->ukeys.p, ukeys_all, n_unique * 4, cudaMemcpyDeviceToDevice, st));

    mark("CSR offsets, unique keys");
    // ---- the streams of the query path (over the transient arrays, which are done) -------------
    FA_CUDA(cudaMemsetAsync(ix->hw.p + n, 0, 8 * sizeof(uint2), st));
    make_hw_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(ix->ref.p, n, ix->hw.p);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    {
        const int cmw1 = ix->prm.frag_len - (ix->prm.window - 1) - (ix->prm.k - 1) - 1;
        FA_CUDA(cudaMemsetAsync(ix->hl.p + n, 0, 8 * sizeof(uint2), st));
        slide_order_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(ix->ref.p, ix->contig_off.p, n, cmw1 < 0 ? 0 : cmw1,
                                                                              ix->prm.frag_len, ix->hl.p, ix->fb.p);
        FA_CUDA(cudaGetLastError());
        if (launches) *launches += 1;
    }
    FA_TRY(ix->irr.reserve((size_t)(n >> 15) + 2));
    FA_CUDA(cudaMemsetAsync(ix->irr.p, 0, ((size_t)(n >> 15) + 2) * 4, st));
    gpos_delta_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(ix->ref.p, n, (uint32_t)ix->prm.frag_len, (uint32_t)ix->prm.window,
                                                                         ix->gpos.p, ix->irr.p);
    FA_CUDA(cudaGetLastError());
    {
        size_t sb = 0;
        FA_CUDA(cub::DeviceScan::InclusiveSum(nullptr, sb, ix->gpos.p, ix->gpos.p, (int64_t)n, st));
        FA_TRY(tmp.reserve(sb + 16));
        FA_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, sb, ix->gpos.p, ix->gpos.p, (int64_t)n, st));
    }
    if (launches) *launches += 3;

    mark("hw, slide order, gpos");
    // ---- directory ---------------------------------------------------------------------------
    int bits = 1;
    while (bits < 24 && (1ull << bits) < n_unique) bits++;      // about one key per directory slot, <= 64 MiB
    ix->dir_bits = bits;
    FA_TRY(ix->dir.reserve((1u << bits) + 1));
    directory_kernel<<<((1u << bits) + 1 + 255) / 256, 256, 0, st>>>(ix->ukeys.p, (uint32_t)n_unique, bits, ix->dir.p);
    FA_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    FA_CUDA(cudaEventRecord(ev.e[1], st));
    FA_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ix->ms_build, ev.e[0], ev.e[1]);
    cudaEventElapsedTime(&ix->ms_sort, ev.e[2], ev.e[3]);
    d_nruns.release(); tmp.release();
    return FA_OK;
}

// ---- mutation of the hash -> positions table (MinimizerIndex.__setitem__ / __delitem__, pyx:1480-1507) ----------
// The reference edits one entry of an unordered_map on the host.  Here the table is three device arrays in CSR form,
// so an edit rebuilds them around the entry -- device-to-device copies of the untouched parts and one pass over the
// offsets -- and then the directory.  Positions are stored as indices into the position-ordered minimizer array, so a
// position has to be one a minimizer of the sketch sits at (any hash); anything else is FA_ERR_INVALID.
namespace {

__global__ void find_positions_kernel(const uint2 *hw, const uint32_t *contig_off, uint32_t n_contigs, const int32_t *seq,
                                      const int32_t *wpos, uint32_t m, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t r = 0xFFFFFFFFu;
    const int32_t sq = seq[i], wp = wpos[i];
    if (sq >= 0 && (uint32_t)sq < n_contigs && wp >= 0) {
        uint32_t lo = contig_off[sq], hi = contig_off[sq + 1];
        const uint32_t end = hi;
        while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((int32_t)(hw[mid].y & 0x7FFFFFFFu) < wp) lo = mid + 1; else hi = mid; }
        if (lo < end && (int32_t)(hw[lo].y & 0x7FFFFFFFu) == wp) r = lo;
    }
    out[i] = r;
}

// out = {slot (lower bound of h among the keys), found, first entry, entries}
__global__ void key_slot_kernel(const uint32_t *ukeys, const uint32_t *uoff, uint32_t n_unique, uint32_t h, uint32_t *out)
{
    uint32_t lo = 0, hi = n_unique;
    while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (ukeys[mid] < h) lo = mid + 1; else hi = mid; }
    const bool found = lo < n_unique && ukeys[lo] == h;
    out[0] = lo; out[1] = found ? 1u : 0u; out[2] = uoff[lo]; out[3] = found ? uoff[lo + 1] - uoff[lo] : 0u;
}

// offsets after the edit: entries up to `slot` keep theirs; later ones come from old entry i + skew, moved by delta
__global__ void edit_offsets_kernel(const uint32_t *old_off, uint32_t *new_off, uint32_t new_n, uint32_t slot, int skew, int delta)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > new_n) return;
    new_off[i] = i <= slot ? old_off[i] : (uint32_t)((int64_t)old_off[(int64_t)i + skew] + delta);
}

}  // namespace

int edit_lookup(fa_index *ix, uint32_t hash, const int32_t *seq, const int32_t *wpos, uint64_t m, bool erase, int *missing)
{
    cudaStream_t st = ix->st;
    if (missing) *missing = 0;
    if (m >= 0x7FFFFFFFull) { set_error("too many positions"); return FA_ERR_INVALID; }
    const uint32_t nu = (uint32_t)ix->n_unique;
    uint32_t h_slot[4] = {0, 0, 0, 0};
    TmpBuf<uint32_t> d_slot;
    FA_TRY(d_slot.reserve(4));
    if (nu) {
        key_slot_kernel<<<1, 1, 0, st>>>(ix->ukeys.p, ix->uoff.p, nu, hash, d_slot.p);
        FA_CUDA(cudaMemcpyAsync(h_slot, d_slot.p, 16, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaStreamSynchronize(st));
    }
    const uint32_t slot = h_slot[0], start = h_slot[2], old_cnt = h_slot[3];
    const bool found = h_slot[1] != 0;
    if (erase && !found) { if (missing) *missing = 1; return FA_OK; }
    uint32_t old_total = 0;
    if (nu) {
        FA_CUDA(cudaMemcpyAsync(&old_total, ix->uoff.p + nu, 4, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaStreamSynchronize(st));
    }
    // the new list as indices of the minimizer array
    TmpBuf<uint32_t> d_list;
    const uint32_t new_cnt = erase ? 0u : (uint32_t)m;
    if (new_cnt) {
        TmpBuf<int32_t> d_sq, d_wp;
        FA_TRY(d_sq.reserve(new_cnt)); FA_TRY(d_wp.reserve(new_cnt)); FA_TRY(d_list.reserve(new_cnt));
        FA_CUDA(cudaMemcpyAsync(d_sq.p, seq, (size_t)new_cnt * 4, cudaMemcpyHostToDevice, st));
        FA_CUDA(cudaMemcpyAsync(d_wp.p, wpos, (size_t)new_cnt * 4, cudaMemcpyHostToDevice, st));
        find_positions_kernel<<<(new_cnt + 127) / 128, 128, 0, st>>>(ix->hw.p, ix->contig_off.p, (uint32_t)ix->n_contigs, d_sq.p, d_wp.p,
                                                                      new_cnt, d_list.p);
        FA_CUDA(cudaGetLastError());
        std::vector<uint32_t> idx(new_cnt);
        FA_CUDA(cudaMemcpyAsync(idx.data(), d_list.p, (size_t)new_cnt * 4, cudaMemcpyDeviceToHost, st));
        FA_CUDA(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < new_cnt; i++)
            if (idx[i] == 0xFFFFFFFFu) {
                set_error("Position(%d, %d) is not the position of a minimizer of this sketch: the device lookup table stores "
                          "positions as indices of the minimizer array", seq[i], wpos[i]);
                return FA_ERR_INVALID;
            }
        std::sort(idx.begin(), idx.end());
        for (uint32_t i = 1; i < new_cnt; i++)
            if (idx[i] == idx[i - 1]) { set_error("a position occurs twice in the list"); return FA_ERR_INVALID; }
    }
    const uint64_t new_total64 = (uint64_t)old_total - old_cnt + new_cnt;
    if (new_total64 >= 0xFFFFFFF0ull) { set_error("lookup table too large"); return FA_ERR_UNSUPPORTED; }
    const uint32_t new_total = (uint32_t)new_total64;
    const uint32_t new_nu = nu + (found ? 0u : 1u) - (erase ? 1u : 0u);
    // entries
    DevBuf<uint32_t> pos2, keys2, off2;
    int rc = pos2.reserve(new_total ? new_total : 1);
    if (rc == FA_OK) rc = keys2.reserve(new_nu ? new_nu : 1);
    if (rc == FA_OK) rc = off2.reserve((size_t)new_nu + 1);
    auto fail = [&](int code) { pos2.release(); keys2.release(); off2.release(); return code; };
    if (rc != FA_OK) return fail(rc);
    cudaError_t e = cudaSuccess;
    auto copy = [&](uint32_t *dst, const uint32_t *src, uint64_t cnt) {
        if (cnt && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, cnt * 4, cudaMemcpyDeviceToDevice, st);
    };
    copy(pos2.p, ix->pos_idx.p, start);
    copy(pos2.p + start, d_list.p, new_cnt);
    copy(pos2.p + start + new_cnt, ix->pos_idx.p + start + old_cnt, (uint64_t)old_total - start - old_cnt);
    // keys
    copy(keys2.p, ix->ukeys.p, slot);
    if (!erase && e == cudaSuccess) e = cudaMemcpyAsync(keys2.p + slot, &hash, 4, cudaMemcpyHostToDevice, st);
    copy(keys2.p + slot + (erase ? 0u : 1u), ix->ukeys.p + slot + (found ? 1u : 0u), (uint64_t)nu - slot - (found ? 1u : 0u));
    if (e != cudaSuccess) { set_error("lookup edit: %s", cudaGetErrorString(e)); return fail(FA_ERR_CUDA); }
    // offsets
    if (nu == 0) {
        const uint32_t two[2] = {0u, new_cnt};
        e = cudaMemcpyAsync(off2.p, two, (size_t)(new_nu + 1) * 4, cudaMemcpyHostToDevice, st);
    } else {
        const int skew = erase ? 1 : (found ? 0 : -1);
        const int delta = erase ? -(int)old_cnt : (found ? (int)new_cnt - (int)old_cnt : (int)new_cnt);
        edit_offsets_kernel<<<(new_nu + 1 + 255) / 256, 256, 0, st>>>(ix->uoff.p, off2.p, new_nu, slot, skew, delta);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("lookup edit: %s", cudaGetErrorString(e)); return fail(FA_ERR_CUDA); }
    ix->pos_idx.release(); ix->ukeys.release(); ix->uoff.release();
    ix->pos_idx = pos2; ix->ukeys = keys2; ix->uoff = off2;
    ix->n_unique = new_nu;
    // directory
    int bits = 1;
    while (bits < 24 && (1ull << bits) < new_nu) bits++;
    ix->dir_bits = bits;                                     // (an emptied table: one bit, all three slots 0)
    FA_TRY(ix->dir.reserve((1u << bits) + 1));
    directory_kernel<<<((1u << bits) + 1 + 255) / 256, 256, 0, st>>>(ix->ukeys.p, new_nu, bits, ix->dir.p);
    FA_CUDA(cudaGetLastError());
    FA_CUDA(cudaStreamSynchronize(st));
    return FA_OK;
}

}  // namespace fa
