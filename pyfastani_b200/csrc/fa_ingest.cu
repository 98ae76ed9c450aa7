// fa_ingest.cu -- the front of the path: 2-bit packing of sequences for staging and a FASTA text
// parser that runs on the device (what src/pyfastani/_fasta.pyx:41-103 does line by line on the host).
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <cstring>

#include "fa_internal.cuh"

using namespace fa;

// ---- 2-bit packing (host side of fa_packed; the device side is gather_contigs_kernel in fa_map.cu) ------------------
namespace {
struct PackTable {
    uint8_t code[256];                                    // 0..3, or 4 = keep the byte in a run
    PackTable()
    {
        memset(code, 4, sizeof code);
        code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2; code['T'] = code['t'] = 3;
    }
};
const PackTable PACK;
}  // namespace

extern "C" int fa_pack_2bit(const uint8_t *data, uint64_t len, uint8_t *bits, uint32_t *run_pos, uint32_t *run_len,
                            uint8_t *run_byte, uint64_t run_cap, uint64_t *n_runs)
{
    if ((len && (!data || !bits)) || !n_runs) { set_error("fa_pack_2bit: null argument"); return FA_ERR_INVALID; }
    if (len >= 0xFFFFFFFFull) { set_error("fa_pack_2bit: sequences of 2^32 bases or more are not supported"); return FA_ERR_UNSUPPORTED; }
    uint64_t nr = 0;
    bool open = false;                                    // the previous byte belongs to run nr - 1
    uint8_t open_byte = 0;
    for (uint64_t i = 0; i < len; i += 4) {
        if (i + 4 <= len) {
            // four plain bases: the usual case, no branch per byte
            const uint8_t c0 = PACK.code[data[i]], c1 = PACK.code[data[i + 1]], c2 = PACK.code[data[i + 2]], c3 = PACK.code[data[i + 3]];
            if (!((c0 | c1 | c2 | c3) & 4)) {
                bits[i >> 2] = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
                open = false;
                continue;
            }
        }
        uint32_t b = 0;
        const int m = (int)std::min<uint64_t>(4, len - i);
        for (int j = 0; j < m; j++) {
            const uint8_t ch = data[i + j], c = PACK.code[ch];
            if (c < 4) { b |= (uint32_t)c << (2 * j); open = false; continue; }
            if (open && ch == open_byte) {
                if (nr <= run_cap) run_len[nr - 1]++;
                continue;
            }
            if (nr < run_cap) { run_pos[nr] = (uint32_t)(i + j); run_len[nr] = 1; run_byte[nr] = ch; }
            nr++;
            open = true; open_byte = ch;
        }
        bits[i >> 2] = (uint8_t)b;
    }
    *n_runs = nr;
    return FA_OK;
}

extern "C" int fa_unpack_2bit(const fa_packed *p, uint64_t len, uint8_t *out)
{
    if (!p || (len && (!p->bits || !out))) { set_error("fa_unpack_2bit: null argument"); return FA_ERR_INVALID; }
    for (uint64_t i = 0; i < len; i++) out[i] = (uint8_t)"ACGT"[(p->bits[i >> 2] >> (2 * (i & 3))) & 3];
    for (uint64_t r = 0; r < p->n_runs; r++) {
        if (p->run_pos[r] >= len) break;
        memset(out + p->run_pos[r], p->run_byte[r], (size_t)std::min<uint64_t>(p->run_len[r], len - p->run_pos[r]));
    }
    return FA_OK;
}

// ---- FASTA text on the device --------------------------------------------------------------------------------------
struct fa_fasta {
    int device = 0;
    DevBuf<uint8_t> bases;                                // the sequences of all records, one after the other
    std::vector<uint64_t> id_begin, id_len, seq_off, seq_len;
    uint64_t n_bases = 0;
};

namespace {

struct IsHeaderStart {
    const uint8_t *t;
    __device__ bool operator()(uint64_t i) const { return t[i] == '>' && (i == 0 || t[i - 1] == '\n'); }
};

__global__ void count_headers_kernel(const uint8_t *t, uint64_t len, unsigned long long *n)
{
    const IsHeaderStart is{t};
    unsigned int c = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x) c += is(i);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n, (unsigned long long)c);
}

// one thread per header line: the position of the '\n' that ends it (or the end of the text)
__global__ void header_end_kernel(const uint8_t *t, uint64_t len, const uint64_t *hs, uint64_t n, uint64_t *he)
{
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint64_t i = hs[r];
    while (i < len && t[i] != '\n') i++;
    he[r] = i;
}

// a byte stays when it is not a newline and does not lie on a header line; headers are few, the bisection is in cache
struct KeepByte {
    const uint8_t *t;
    const uint64_t *hs, *he;
    uint64_t n;
    __device__ bool operator()(uint64_t i) const
    {
        if (t[i] == '\n') return false;
        uint64_t lo = 0, hi = n;                          // last header that starts at or before i
        while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (hs[mid] <= i) lo = mid; else hi = mid; }
        return i > he[lo];
    }
};

// toupper of the C locale, as copy_upper's scalar tail does (sequtils.cpp:68-79)
struct Upper {
    const uint8_t *t;
    __device__ uint8_t operator()(uint64_t i) const { const uint8_t c = t[i]; return c >= 'a' && c <= 'z' ? (uint8_t)(c - 32) : c; }
};

// per record: newlines between the end of its header and the start of the next one (one warp per record)
__global__ void count_newlines_kernel(const uint8_t *t, uint64_t len, const uint64_t *hs, const uint64_t *he, uint64_t n, uint64_t *nl)
{
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint64_t b = min(he[r] + 1, len), e = r + 1 < n ? hs[r + 1] : len;
    uint64_t c = 0;
    for (uint64_t i = b + (threadIdx.x & 31); i < e; i += 32) c += t[i] == '\n';
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) nl[r] = c;
}

}  // namespace

extern "C" int fa_fasta_parse(int32_t device, const void *text, uint64_t len, fa_fasta **out)
{
    if (!out || (len && !text)) { set_error("fa_fasta_parse: null argument"); return FA_ERR_INVALID; }
    FA_CUDA(cudaSetDevice(device));
    fa_fasta *f = new fa_fasta;
    f->device = device;
    *out = f;
    const uint8_t *h = (const uint8_t *)text;
    if (len == 0 || h[0] != '>') return FA_OK;            // _fasta.pyx:77-78: the first line is not a header -- no records
    auto fail = [&](int rc) { fa_fasta_free(f); *out = nullptr; return rc; };
    TmpBuf<uint8_t> d_text, tmp;
    TmpBuf<uint64_t> d_hs, d_he, d_n;
    int rc;
    if ((rc = d_text.reserve(len)) != FA_OK || (rc = d_n.reserve(2)) != FA_OK) return fail(rc);
#define FA_F(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_error("%s: %s", #call, cudaGetErrorString(e_)); return fail(FA_ERR_CUDA); } } while (0)
    cudaStream_t st = 0;
    FA_F(cudaMemcpyAsync(d_text.p, text, len, cudaMemcpyHostToDevice, st));
    // 1. header lines: count, then list in order
    thrust::counting_iterator<uint64_t> idx(0);
    uint64_t n_rec = 0;
    FA_F(cudaMemsetAsync(d_n.p, 0, 2 * sizeof(uint64_t), st));
    count_headers_kernel<<<148 * 8, 256, 0, st>>>(d_text.p, len, (unsigned long long *)d_n.p);
    FA_F(cudaMemcpyAsync(&n_rec, d_n.p, sizeof n_rec, cudaMemcpyDeviceToHost, st));
    FA_F(cudaStreamSynchronize(st));
    if ((rc = d_hs.reserve(n_rec)) != FA_OK) return fail(rc);
    {
        size_t bytes = 0;
        auto flags = thrust::make_transform_iterator(idx, IsHeaderStart{d_text.p});
        FA_F(cub::DeviceSelect::Flagged(nullptr, bytes, idx, flags, d_hs.p, d_n.p, (int64_t)len, st));
        if ((rc = tmp.reserve(bytes + 16)) != FA_OK) return fail(rc);
        FA_F(cub::DeviceSelect::Flagged(tmp.p, bytes, idx, flags, d_hs.p, d_n.p, (int64_t)len, st));
    }
    if ((rc = d_he.reserve(2 * n_rec)) != FA_OK) return fail(rc);
    header_end_kernel<<<(unsigned int)((n_rec + 127) / 128), 128, 0, st>>>(d_text.p, len, d_hs.p, n_rec, d_he.p);
    count_newlines_kernel<<<(unsigned int)((n_rec * 32 + 255) / 256), 256, 0, st>>>(d_text.p, len, d_hs.p, d_he.p, n_rec, d_he.p + n_rec);
    FA_F(cudaGetLastError());
    // 2. everything else that is not a newline, upper-cased, in order
    if ((rc = f->bases.reserve(len + 16)) != FA_OK) return fail(rc);
    {
        size_t bytes = 0;
        auto vals = thrust::make_transform_iterator(idx, Upper{d_text.p});
        auto keep = thrust::make_transform_iterator(idx, KeepByte{d_text.p, d_hs.p, d_he.p, n_rec});
        FA_F(cub::DeviceSelect::Flagged(nullptr, bytes, vals, keep, f->bases.p, d_n.p + 1, (int64_t)len, st));
        if ((rc = tmp.reserve(bytes + 16)) != FA_OK) return fail(rc);
        FA_F(cub::DeviceSelect::Flagged(tmp.p, bytes, vals, keep, f->bases.p, d_n.p + 1, (int64_t)len, st));
    }
    std::vector<uint64_t> hs(n_rec), he(2 * n_rec);
    FA_F(cudaMemcpyAsync(hs.data(), d_hs.p, n_rec * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    FA_F(cudaMemcpyAsync(he.data(), d_he.p, 2 * n_rec * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    FA_F(cudaMemcpyAsync(&f->n_bases, d_n.p + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    FA_F(cudaStreamSynchronize(st));
#undef FA_F
    // 3. records: the sequence of record r is what lies between its header and the next one, minus the newlines
    uint64_t off = 0;
    for (uint64_t r = 0; r < n_rec; r++) {
        const uint64_t b = std::min(he[r] + 1, len), e = r + 1 < n_rec ? hs[r + 1] : len;
        const uint64_t sl = e - b - he[n_rec + r];
        f->id_begin.push_back(hs[r] + 1); f->id_len.push_back(he[r] - hs[r] - 1);
        f->seq_off.push_back(off); f->seq_len.push_back(sl);
        off += sl;
    }
    if (off != f->n_bases) { set_error("fa_fasta_parse: %llu bases kept, %llu expected", (unsigned long long)f->n_bases, (unsigned long long)off); return fail(FA_ERR_STATE); }
    return FA_OK;
}

extern "C" void fa_fasta_free(fa_fasta *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    f->bases.release();
    delete f;
}

extern "C" int fa_fasta_counts(const fa_fasta *f, uint64_t *n_records, uint64_t *n_bases)
{
    if (!f) { set_error("null fasta"); return FA_ERR_INVALID; }
    if (n_records) *n_records = f->seq_len.size();
    if (n_bases) *n_bases = f->n_bases;
    return FA_OK;
}

extern "C" int fa_fasta_records(const fa_fasta *f, fa_contig *contigs, uint64_t *id_begin, uint64_t *id_len)
{
    if (!f) { set_error("null fasta"); return FA_ERR_INVALID; }
    for (size_t r = 0; r < f->seq_len.size(); r++) {
        if (contigs) contigs[r] = fa_contig{f->bases.p + f->seq_off[r], 1, 1, (int64_t)f->seq_len[r]};
        if (id_begin) id_begin[r] = f->id_begin[r];
        if (id_len) id_len[r] = f->id_len[r];
    }
    return FA_OK;
}
