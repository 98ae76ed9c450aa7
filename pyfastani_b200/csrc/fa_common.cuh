// fa_common.cuh -- shared declarations of libfastani_b200 (host + device).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/fastani_b200.h"

namespace fa {

void set_error(const char *fmt, ...);

#define FA_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            fa::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return FA_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define FA_TRY(call)                 \
    do {                             \
        int rc_ = (call);            \
        if (rc_ != FA_OK) return rc_; \
    } while (0)

// Position-ordered reference minimizer (MinimizerInfo of
// FA/map/include/base_types.hpp:22-53) widened to one 16-byte vector so L1 and L2 fetch a
// whole element with a single LDG.128:
//   x = hash, y = wpos, z = seqId,
//   w = (delta to the next element of the same contig with the same hash) << 16
//       | (delta to the previous one); 0 = none within 65535 elements.
typedef uint4 RefMini;

// Growable device array.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n, bool keep = false, cudaStream_t st = 0)
    {
        if (n <= cap) return FA_OK;
        size_t ncap = n;                       // a first allocation is exact (index arrays are gigabytes); a buffer that
        if (cap) {                             // grows is a workspace: leave head-room
            ncap = cap;
            while (ncap < n) ncap += ncap / 2 + 256;
        }
        T *np = nullptr;
        cudaError_t e = cudaMalloc((void **)&np, ncap * sizeof(T));
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes): %s", ncap * sizeof(T), cudaGetErrorString(e)); return FA_ERR_NOMEM; }
        if (keep && p && cap) {
            e = cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(np); set_error("grow copy: %s", cudaGetErrorString(e)); return FA_ERR_CUDA; }
        }
        if (p) cudaFree(p);
        p = np; cap = ncap;
        return FA_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// A DevBuf that frees itself: for temporaries of one call (DevBuf itself is a plain handle that is moved between owners).
template <typename T>
struct TmpBuf : DevBuf<T> {
    TmpBuf() = default;
    TmpBuf(const TmpBuf &) = delete;
    TmpBuf &operator=(const TmpBuf &) = delete;
    ~TmpBuf() { this->release(); }
    DevBuf<T> take() { DevBuf<T> d = *this; this->p = nullptr; this->cap = 0; return d; }   // hand the memory to a long-lived owner
};

// Pinned host staging buffer.
struct PinBuf {
    uint8_t *p = nullptr;
    size_t cap = 0;
    int reserve(size_t n)
    {
        if (n <= cap) return FA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t ncap = n + n / 4 + 4096;
        cudaError_t e = cudaMallocHost((void **)&p, ncap);
        if (e != cudaSuccess) { set_error("cudaMallocHost(%zu): %s", ncap, cudaGetErrorString(e)); return FA_ERR_NOMEM; }
        cap = ncap;
        return FA_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// NVTX ranges for the stages of a call (tracing): one range open at a time, closed on scope exit
// whatever the return path.  Header-only NVTX3: nothing happens unless a profiler is attached.
struct NvtxStages {
    bool open = false;
    void next(const char *name) { if (open) nvtxRangePop(); nvtxRangePushA(name); open = true; }
    void close() { if (open) nvtxRangePop(); open = false; }
    ~NvtxStages() { close(); }
};

// ---- sketching (fa_sketch.cu) ---------------------------------------------------------
// One sequence of a batch: `off` is its 16-byte-aligned start in the batch byte buffer.
struct SeqDesc {
    uint64_t off;
    int32_t  len;
    int32_t  id;        // seqId stamped on the minimizers (reference contigs) / fragment number (queries)
    int32_t  raw;       // 1 = bytes are already normalised on the host (UCS2/UCS4 input, pyx:147-148)
    int32_t  tile0;     // first tile of this sequence in the batch
};

struct SketchScratch {
    DevBuf<uint8_t>  bytes;      // batch bytes
    DevBuf<SeqDesc>  seqs;
    DevBuf<unsigned long long> tile_status;   // decoupled look-back words
    DevBuf<unsigned long long> counters;      // [0] tile ticket, [1] emitted total, [2] quirk runs found
    DevBuf<uint64_t> seq_first;  // per sequence: exclusive prefix of its first tile
    DevBuf<ulonglong2> drops;    // first-window quirk: runs [first, last) to remove
};

constexpr int SK_THREADS = 128;     // threads per sketch CTA
constexpr int SK_PER_THREAD = 8;    // k-mer positions per thread
constexpr int SK_TILE = SK_THREADS * SK_PER_THREAD;

// Launch the sketch kernel over `n_tiles` tiles of the batch; minimizers are appended in
// (sequence, position) order to out_ref[out_base ...] (reference mode: RefMini) or to
// out_hash[...] (query mode: hash only).  counters[1] receives the total emitted.
// slot_cap / seq_cnt (query fragments): sequence f writes its hashes at out_hash[f * slot_cap ...] and its count to
// seq_cnt[f]; without them the output is one array in (sequence, position) order.
int launch_sketch(cudaStream_t st, const SketchScratch &sc, int n_seqs, int n_tiles, int k, int w, int fwd_only,
                  RefMini *out_ref, uint32_t *out_hash, uint64_t out_base, int *launches, int slot_cap = 0, uint32_t *seq_cnt = nullptr,
                  int uniform_tiles = 0);     // uniform_tiles > 0: every sequence has that many tiles (query fragments)

// The reference's `wpos == 0` comparison quirk (pyx:219-222) suppresses some
// minimizers right after window 0 of a contig.  launch_quirk_find records the affected runs of
// the batch just written at batch_ref (counters[2] = how many); quirk_compact removes them.
int launch_quirk_find(cudaStream_t st, const SketchScratch &sc, int n_seqs, const RefMini *batch_ref, int *launches);
int quirk_compact(cudaStream_t st, const SketchScratch &sc, unsigned int n_runs, RefMini *batch_ref, uint64_t n,
                  uint64_t *n_out, int *launches);

}  // namespace fa
