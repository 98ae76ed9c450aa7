// fa_comm.cu -- the exchange steps of the reference-sharded multi-GPU layout: reference genomes are dealt to
// the GPUs as whole genomes, every GPU maps every query against its own shard and produces FINAL hit rows for its genomes
// (the per-thread split of upstream FastANI, FA/cgi/include/computeCoreIdentity.hpp:454-484: splitReferenceGenomes,
// correctRefGenomeIds).  Two things travel over NCCL / NVLink:
//   * the query sketches (sketch_exchange): every rank sketches 1/world of the fragments of a group of queries and one
//     ncclAllGather hands every rank all of them, on a stream of its own, one group ahead of the mapping;
//   * the hit rows of a batch of queries (fa_gather_hits): a few KB per query, bound by launch latency, which is why the
//     counts and the first rows of every rank travel in ONE collective; merged into the order of pyx:1135 (identity
//     descending, stable in ascending genome id).
// No collective runs inside the mapping kernels, and none depends on anything a rank measured (pass sizes differ between
// the ranks; the groups of the sketch exchange depend on the query sizes alone).
//
// NCCL is bound at run time (dlopen of libnccl.so.2 at the first fa_comm_* call): a process that already carries an NCCL
// -- torch.distributed in bench.py -- shares that copy instead of loading a second one, and single-GPU users need none.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "fa_internal.cuh"

namespace fa {
namespace {

struct Nccl {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

int load_nccl(const Nccl **out)
{
    static std::mutex mtx;
    static Nccl api;
    static bool tried = false, ok = false;
    std::lock_guard<std::mutex> g(mtx);
    if (!tried) {
        tried = true;
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) {
            api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(api.handle, "ncclGetVersion"));
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
            api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
            ok = api.GetVersion && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
        }
    }
    if (!ok) { set_error("NCCL is not available: %s", api.handle ? "libnccl.so.2 lacks a required symbol" : dlerror()); return FA_ERR_UNSUPPORTED; }
    *out = &api;
    return FA_OK;
}

#define FA_NCCL(api, call)                                                                              \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) {                                                                        \
            fa::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, (api)->GetErrorString(r_)); \
            return FA_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

constexpr uint32_t INLINE_ROWS = 2048;      // rows of a rank that travel with the counts (32 KB): one collective for the usual batch

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace
}  // namespace fa

using namespace fa;

struct fa_comm {
    const Nccl *api = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = 0;
    cudaStream_t st = nullptr;
    DevBuf<uint8_t> d_send, d_recv;
    PinBuf h_send, h_recv;
    std::mutex mtx;
    uint64_t collectives = 0, bytes_gathered = 0;
};

// The second exchange of the reference-sharded layout: the query sketches.  Every rank maps every query, so every rank
// would sketch every query -- the largest stage of a 10 000 x 10 000 run (a third of it).  Instead rank r sketches
// fragments [r * per, (r + 1) * per) of a group of queries and one ncclAllGather on the mapping stream hands every rank
// all sketches, packed to `stride` hashes per fragment.  The exchange has its own stream and scratch and runs one group
// ahead of the mapping (query_batch_impl): its kernels fill the gaps of the mapping kernels, and a rank that arrives late
// at the collective (the one that holds the genus of the current queries) delays the NEXT group's sketches, not a pass.
int fa::sketch_exchange(fa_index *ix, fa_comm *c, const fa_contig *contigs, int32_t n_contigs, int slot, PreSketch *ps, fa_query_info *qi)
{
    const int stride = exchange_stride(ix->prm);
    if (stride <= 0) { set_error("sketch exchange is not available for these parameters"); return FA_ERR_UNSUPPORTED; }
    uint32_t per = 0;
    uint64_t frags = 0;
    FA_TRY(sketch_share(ix, contigs, n_contigs, c->world, c->rank, (uint32_t)stride, &per, &frags, qi));
    std::lock_guard<std::mutex> guard(c->mtx);
    ExchScratch &xs = ix->xs;
    const size_t block = (size_t)per * ((size_t)stride + 1);
    FA_TRY(xs.recv[slot].reserve(block * (size_t)c->world));
    FA_NCCL(c->api, c->api->AllGather(xs.send.p, xs.recv[slot].p, block, ncclUint32, c->comm, xs.st));
    FA_CUDA(cudaEventRecord(xs.t1, xs.st));
    FA_CUDA(cudaEventRecord(xs.done[slot], xs.st));
    c->collectives++; c->bytes_gathered += block * 4 * (size_t)c->world;
    ps->recv = xs.recv[slot].p; ps->per = per; ps->stride = (uint32_t)stride; ps->block = block; ps->first_frag = 0;
    return FA_OK;
}

int fa::comm_world(const fa_comm *c) { return c ? c->world : 1; }

extern "C" {

int fa_comm_unique_id(uint8_t *id)
{
    if (!id) { set_error("bad arguments"); return FA_ERR_INVALID; }
    const Nccl *api = nullptr;
    FA_TRY(load_nccl(&api));
    static_assert(sizeof(ncclUniqueId) == FA_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId u;
    FA_NCCL(api, api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return FA_OK;
}

int fa_comm_create(const uint8_t *id, int32_t world, int32_t rank, int32_t device, fa_comm **out)
{
    if (!id || !out || world < 1 || rank < 0 || rank >= world) { set_error("bad arguments"); return FA_ERR_INVALID; }
    const Nccl *api = nullptr;
    FA_TRY(load_nccl(&api));
    FA_CUDA(cudaSetDevice(device));
    fa_comm *c = new (std::nothrow) fa_comm();
    if (!c) return FA_ERR_NOMEM;
    c->api = api; c->world = world; c->rank = rank; c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); delete c; return FA_ERR_CUDA; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank(world %d, rank %d): %s", world, rank, api->GetErrorString(r));
        cudaStreamDestroy(c->st);
        delete c;
        return FA_ERR_CUDA;
    }
    *out = c;
    return FA_OK;
}

void fa_comm_free(fa_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->comm) c->api->CommDestroy(c->comm);
    c->d_send.release(); c->d_recv.release(); c->h_send.release(); c->h_recv.release();
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int fa_comm_info(const fa_comm *c, int32_t *world, int32_t *rank, int32_t *nccl_version, uint64_t *collectives, uint64_t *bytes_gathered)
{
    if (!c) { set_error("comm is NULL"); return FA_ERR_INVALID; }
    if (world) *world = c->world;
    if (rank) *rank = c->rank;
    if (nccl_version) { int v = 0; c->api->GetVersion(&v); *nccl_version = v; }
    if (collectives) *collectives = c->collectives;
    if (bytes_gathered) *bytes_gathered = c->bytes_gathered;
    return FA_OK;
}

// One all-gather of `bytes` per rank: pinned host -> device -> ncclAllGather -> pinned host, on the communicator's stream.
static int gather_block(fa_comm *c, size_t bytes)
{
    const size_t all = bytes * (size_t)c->world;
    FA_TRY(c->d_send.reserve(bytes)); FA_TRY(c->d_recv.reserve(all)); FA_TRY(c->h_recv.reserve(all));
    FA_CUDA(cudaMemcpyAsync(c->d_send.p, c->h_send.p, bytes, cudaMemcpyHostToDevice, c->st));
    FA_NCCL(c->api, c->api->AllGather(c->d_send.p, c->d_recv.p, bytes, ncclChar, c->comm, c->st));
    FA_CUDA(cudaMemcpyAsync(c->h_recv.p, c->d_recv.p, all, cudaMemcpyDeviceToHost, c->st));
    FA_CUDA(cudaStreamSynchronize(c->st));
    c->collectives++; c->bytes_gathered += all;
    return FA_OK;
}

int fa_gather_hits(fa_comm *c, const fa_hit *rows, const uint64_t *hit_offsets, int32_t n_queries, const int32_t *genome_offsets,
                   fa_hit *out, uint64_t cap, uint64_t *out_offsets)
{
    if (!c || n_queries < 0 || !hit_offsets || !genome_offsets || !out_offsets) { set_error("bad arguments"); return FA_ERR_INVALID; }
    std::lock_guard<std::mutex> guard(c->mtx);
    FA_CUDA(cudaSetDevice(c->device));
    const uint32_t nq = (uint32_t)n_queries, W = (uint32_t)c->world;
    const uint64_t total = hit_offsets[nq] - hit_offsets[0];
    if (total > 0 && !rows) { set_error("bad arguments"); return FA_ERR_INVALID; }
    const fa_hit *mine = rows ? rows + hit_offsets[0] : nullptr;

    // ---- phase 1: [total][counts of the queries] + the first INLINE_ROWS rows of every rank -------------------------
    const size_t head = align16(8 + 4 * (size_t)nq), b1 = head + (size_t)INLINE_ROWS * sizeof(fa_hit);
    FA_TRY(c->h_send.reserve(b1));
    memset(c->h_send.p, 0, head);
    uint64_t *h_tot = reinterpret_cast<uint64_t *>(c->h_send.p);
    uint32_t *h_cnt = reinterpret_cast<uint32_t *>(c->h_send.p + 8);
    *h_tot = total;
    for (uint32_t q = 0; q < nq; q++) h_cnt[q] = (uint32_t)(hit_offsets[q + 1] - hit_offsets[q]);
    const uint64_t inl = std::min<uint64_t>(total, INLINE_ROWS);
    if (inl) memcpy(c->h_send.p + head, mine, (size_t)inl * sizeof(fa_hit));
    FA_TRY(gather_block(c, b1));
    std::vector<uint64_t> r_total(W);
    std::vector<std::vector<fa_hit>> r_rows(W);
    uint64_t max_total = 0, grand = 0;
    for (uint32_t r = 0; r < W; r++) {
        const uint8_t *blk = c->h_recv.p + (size_t)r * b1;
        r_total[r] = *reinterpret_cast<const uint64_t *>(blk);
        max_total = std::max(max_total, r_total[r]); grand += r_total[r];
        r_rows[r].resize((size_t)r_total[r]);
        const uint64_t n = std::min<uint64_t>(r_total[r], INLINE_ROWS);
        if (n) memcpy(r_rows[r].data(), blk + head, (size_t)n * sizeof(fa_hit));
    }
    std::vector<uint32_t> r_cnt((size_t)W * nq);
    for (uint32_t r = 0; r < W; r++)
        if (nq) memcpy(&r_cnt[(size_t)r * nq], c->h_recv.p + (size_t)r * b1 + 8, 4 * (size_t)nq);
    // ---- phase 2 (only when some rank holds more): the remaining rows, padded to the largest remainder -----------------
    if (max_total > INLINE_ROWS) {
        const uint64_t width = max_total - INLINE_ROWS;
        const size_t b2 = (size_t)width * sizeof(fa_hit);
        FA_TRY(c->h_send.reserve(b2));
        const uint64_t rest = total > INLINE_ROWS ? total - INLINE_ROWS : 0;
        if (rest) memcpy(c->h_send.p, mine + INLINE_ROWS, (size_t)rest * sizeof(fa_hit));
        if (rest < width) memset(c->h_send.p + (size_t)rest * sizeof(fa_hit), 0, (size_t)(width - rest) * sizeof(fa_hit));
        FA_TRY(gather_block(c, b2));
        for (uint32_t r = 0; r < W; r++)
            if (r_total[r] > INLINE_ROWS)
                memcpy(r_rows[r].data() + INLINE_ROWS, c->h_recv.p + (size_t)r * b2, (size_t)(r_total[r] - INLINE_ROWS) * sizeof(fa_hit));
    }
    // ---- merge: per query, the rows of rank 0, 1, ... one after the other (ascending global genome id among equal
    // identities, since every rank's rows are already in that order), then the stable sort of pyx:1135 ------------------
    if (grand > cap && out) { set_error("%llu gathered hits do not fit the output (capacity %llu)", (unsigned long long)grand, (unsigned long long)cap); return FA_ERR_INVALID; }
    std::vector<uint64_t> pos(W, 0);
    uint64_t used = 0;
    out_offsets[0] = 0;
    for (uint32_t q = 0; q < nq; q++) {
        const uint64_t q0 = used;
        for (uint32_t r = 0; r < W; r++) {
            const uint32_t n = r_cnt[(size_t)r * nq + q];
            if (pos[r] + n > r_total[r]) { set_error("rank %u sent inconsistent counts", r); return FA_ERR_STATE; }
            for (uint32_t i = 0; i < n; i++) {
                fa_hit h = r_rows[r][(size_t)pos[r] + i];
                h.ref_genome += genome_offsets[r];              // correctRefGenomeIds, computeCoreIdentity.hpp:477-484
                if (out) out[used] = h;
                used++;
            }
            pos[r] += n;
        }
        if (out) std::stable_sort(out + q0, out + used, [](const fa_hit &a, const fa_hit &b) { return a.identity > b.identity; });
        out_offsets[q + 1] = used;
    }
    return FA_OK;
}

int fa_query_batch_sharded(fa_index *ix, fa_comm *comm, const fa_contig *contigs, const int32_t *contigs_per_query, int32_t n_queries,
                           const int32_t *genome_offsets, fa_hit *out, uint64_t cap, uint64_t *hit_offsets, fa_query_info *info)
{
    if (!ix || !comm || !genome_offsets || !hit_offsets || n_queries < 0) { set_error("bad arguments"); return FA_ERR_INVALID; }
    const uint64_t n_local = ix->genome_len.size();
    if ((uint64_t)(genome_offsets[comm->rank + 1] - genome_offsets[comm->rank]) != n_local) {
        set_error("rank %d holds %llu genomes but its shard is [%d, %d)", comm->rank, (unsigned long long)n_local,
                  genome_offsets[comm->rank], genome_offsets[comm->rank + 1]);
        return FA_ERR_INVALID;
    }
    std::vector<fa_hit> local((size_t)std::max<uint64_t>(n_local, 1) * (size_t)std::max(n_queries, 1));
    std::vector<uint64_t> offs((size_t)n_queries + 1, 0);
    FA_TRY(query_batch_impl(ix, comm, contigs, contigs_per_query, n_queries, local.data(), local.size(), offs.data(), info));
    return fa_gather_hits(comm, local.data(), offs.data(), n_queries, genome_offsets, out, cap, hit_offsets);
}

}  // extern "C"
