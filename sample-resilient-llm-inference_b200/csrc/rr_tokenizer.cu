// rr_tokenizer.cu — K2: prompt token count + tokenisation of a batch of chat messages on the device.
//
// The library's tokenizer is byte level (no tokenizer files exist on the box): id 1 = BOS, then
// 3 + byte value (0 / 1 / 2 reserved: pad / bos / eos), folded into the model's vocabulary when vocab < 259.  So the
// token count of a text is n_bytes + 1 and the ids of text i start at ids_off[i] = text_off[i] + i in the packed output:
// no scan is needed, every 16-byte piece of every text is independent -- pure HBM-bound byte work: one CTA per
// (text, 4 KB chunk), 16-byte vector loads, 64-byte vector stores (4 x int4 per thread), coalesced both ways.
// The counts are what the ADMIT events carry into the rpm / tpm check of K1 (rr_router.cu).
//
// Replaces litellm.token_counter as used for tpm accounting (no call site in the reference tree; the tpm values it
// serves are reference config/config.yaml:42,50,57,65,72,80,87,94).
#include "rr_ptx.cuh"
#include "rr_launch.cuh"
#include "rr_kernels.h"

#include <mutex>
#include <string.h>

#define RR_API extern "C" __attribute__((visibility("default")))

namespace rr {
void note_cuda_error(cudaError_t e);

constexpr int TOK_CHUNK = 4096;          // bytes per CTA
constexpr int TOK_THREADS = 256;         // 16 bytes per thread

__device__ __forceinline__ int32_t tok_of_byte(uint32_t b, int vocab) {
    const int32_t t = 3 + (int32_t)b;
    return t < vocab ? t : 3 + (t - 3) % (vocab - 3);
}

__global__ void __launch_bounds__(TOK_THREADS)
tokenize_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ text_off, int n_texts, int vocab,
                int32_t* __restrict__ ids, int32_t* __restrict__ counts, int64_t* __restrict__ ids_off) {
    griddep_launch();
    griddep_wait();
    const int ti = blockIdx.y;
    if (ti >= n_texts) return;
    const int64_t t0 = text_off[ti], t1 = text_off[ti + 1];
    const int64_t len = t1 - t0;
    const int64_t o0 = t0 + ti;                                  // ids of text ti start here: one BOS per earlier text
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[ti] = (int32_t)len + 1;
        if (ids_off) { ids_off[ti] = o0; if (ti == n_texts - 1) ids_off[n_texts] = t1 + n_texts; }
        if (ids) ids[o0] = 1;                                    // BOS
    }
    if (!ids) return;
    const int64_t c0 = (int64_t)blockIdx.x * TOK_CHUNK;
    if (c0 >= len) return;
    const int64_t c1 = c0 + TOK_CHUNK < len ? c0 + TOK_CHUNK : len;
    const uint8_t* src = text + t0;
    int32_t* dst = ids + o0 + 1;
    // peel to 16-byte alignment of the SOURCE, then vector body, then tail
    int64_t i = c0 + threadIdx.x;
    const int64_t mis = (16 - ((uintptr_t)(src + c0) & 15)) & 15;
    const int64_t head_end = c0 + mis < c1 ? c0 + mis : c1;
    if (i < head_end) dst[i] = tok_of_byte(src[i], vocab);
    const int64_t nvec = (c1 - head_end) / 16;
    for (int64_t v = threadIdx.x; v < nvec; v += TOK_THREADS) {
        const int64_t p = head_end + v * 16;
        const uint4 w = ldg_nc_v4(src + p);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
        int32_t out[16];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) out[4 * k + b] = tok_of_byte((ws[k] >> (8 * b)) & 0xffu, vocab);
        if (((uintptr_t)(dst + p) & 15) == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                *reinterpret_cast<int4*>(dst + p + 4 * k) = make_int4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[p + k] = out[k];
        }
    }
    const int64_t tail0 = head_end + nvec * 16;
    i = tail0 + threadIdx.x;
    if (i < c1) dst[i] = tok_of_byte(src[i], vocab);
}

struct TokScratch {                      // per-process staging (grown on demand), serialised by `mu`
    std::mutex mu;
    int device = -1;
    cudaStream_t stream = nullptr;
    uint8_t *h_text = nullptr, *d_text = nullptr;
    int64_t *h_off = nullptr, *d_off = nullptr, *d_ids_off = nullptr;
    int32_t *d_ids = nullptr, *d_counts = nullptr, *h_ids = nullptr, *h_counts = nullptr;
    int64_t *h_ids_off = nullptr;
    size_t cap_bytes = 0, cap_texts = 0;
};
static TokScratch g_tok;
// forget the buffers of another device (they stay allocated there: a process normally tokenizes on one device)
static TokScratch& TokScratchReset(TokScratch& s) {
    s.stream = nullptr;
    s.h_text = s.d_text = nullptr; s.h_off = s.d_off = s.d_ids_off = s.h_ids_off = nullptr;
    s.d_counts = s.h_counts = s.d_ids = s.h_ids = nullptr;
    s.cap_bytes = s.cap_texts = 0;
    return s;
}

static int tok_reserve(TokScratch& s, size_t bytes, size_t n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (s.device != dev) {                                       // first use (or the caller switched devices): start over
        TokScratchReset(s);
        s.device = dev;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return RR_CUDA_ERROR;
    }
    if (bytes <= s.cap_bytes && n <= s.cap_texts) return RR_OK;
    // grow: free everything, allocate for the new capacities (the scratch is only ever used under s.mu, synchronously)
    const size_t cb = bytes > s.cap_bytes ? bytes * 2 + 4096 : s.cap_bytes;
    const size_t ct = n > s.cap_texts ? n * 2 + 64 : s.cap_texts;
    if (s.h_text) cudaFreeHost(s.h_text);
    if (s.d_text) cudaFree(s.d_text);
    if (s.h_off) cudaFreeHost(s.h_off);
    if (s.d_off) cudaFree(s.d_off);
    if (s.d_ids_off) cudaFree(s.d_ids_off);
    if (s.d_counts) cudaFree(s.d_counts);
    if (s.h_counts) cudaFreeHost(s.h_counts);
    if (s.h_ids_off) cudaFreeHost(s.h_ids_off);
    if (s.d_ids) cudaFree(s.d_ids);
    if (s.h_ids) cudaFreeHost(s.h_ids);
    s.h_text = s.d_text = nullptr; s.h_off = s.d_off = s.d_ids_off = s.h_ids_off = nullptr;
    s.d_counts = s.h_counts = s.d_ids = s.h_ids = nullptr;
    s.cap_bytes = s.cap_texts = 0;
    const size_t ni = cb + ct + 16;
    if (cudaMallocHost(&s.h_text, cb) != cudaSuccess || cudaMalloc(&s.d_text, cb) != cudaSuccess ||
        cudaMallocHost(&s.h_off, (ct + 1) * 8) != cudaSuccess || cudaMalloc(&s.d_off, (ct + 1) * 8) != cudaSuccess ||
        cudaMalloc(&s.d_ids_off, (ct + 1) * 8) != cudaSuccess || cudaMalloc(&s.d_counts, ct * 4) != cudaSuccess ||
        cudaMallocHost(&s.h_counts, ct * 4) != cudaSuccess || cudaMallocHost(&s.h_ids_off, (ct + 1) * 8) != cudaSuccess ||
        cudaMalloc(&s.d_ids, ni * 4) != cudaSuccess || cudaMallocHost(&s.h_ids, ni * 4) != cudaSuccess)
        return RR_CUDA_ERROR;
    s.cap_bytes = cb; s.cap_texts = ct;
    return RR_OK;
}

int launch_tokenize(const uint8_t* d_text, const int64_t* d_off, int n_texts, int64_t max_len, int vocab, int32_t* d_ids,
                    int32_t* d_counts, int64_t* d_ids_off, cudaStream_t st) {
    if (n_texts <= 0) return RR_OK;
    const int chunks = (int)((max_len + TOK_CHUNK - 1) / TOK_CHUNK);
    dim3 grid(chunks > 0 ? chunks : 1, n_texts);
    cudaError_t e = launch_pdl(tokenize_kernel, grid, dim3(TOK_THREADS), 0, st, d_text, d_off, n_texts, vocab, d_ids, d_counts, d_ids_off);
    return e == cudaSuccess ? RR_OK : RR_CUDA_ERROR;
}
}  // namespace rr

using namespace rr;

// Host buffers in and out; the H2D copy, the kernel and the D2H copies happen inside.
//   text        n_texts messages back to back, text_off[n_texts + 1] byte offsets
//   counts_out  [n_texts] token counts (n_bytes + 1)
//   ids_out     optional packed ids, capacity ids_capacity; ids_off_out[n_texts + 1] optional offsets into it
RR_API int rr_tokenize_batch(const uint8_t* text, const int64_t* text_off, int n_texts, int32_t vocab, int32_t* counts_out,
                             int32_t* ids_out, int64_t ids_capacity, int64_t* ids_off_out) {
    if (n_texts < 0 || !text_off || !counts_out || vocab < 4) return RR_INVALID_ARGUMENT;
    if (n_texts == 0) return RR_OK;
    const int64_t total = text_off[n_texts] - text_off[0];
    if (total < 0 || (!text && total > 0) || n_texts > 65535) return RR_INVALID_ARGUMENT;
    int64_t max_len = 0;
    for (int i = 0; i < n_texts; ++i) {
        const int64_t l = text_off[i + 1] - text_off[i];
        if (l < 0) return RR_INVALID_ARGUMENT;
        if (l > max_len) max_len = l;
    }
    if (ids_out && ids_capacity < total + n_texts) return RR_INVALID_ARGUMENT;
    TokScratch& s = g_tok;
    std::lock_guard<std::mutex> lk(s.mu);
    int rc = tok_reserve(s, (size_t)total, (size_t)n_texts);
    if (rc != RR_OK) { note_cuda_error(cudaGetLastError()); return rc; }
    const int64_t base = text_off[0];
    if (total > 0) memcpy(s.h_text, text + base, (size_t)total);
    for (int i = 0; i <= n_texts; ++i) s.h_off[i] = text_off[i] - base;
    cudaStream_t st = s.stream;
    cudaError_t e = cudaSuccess;
    if (total > 0) e = cudaMemcpyAsync(s.d_text, s.h_text, (size_t)total, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.d_off, s.h_off, (size_t)(n_texts + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { note_cuda_error(e); return RR_CUDA_ERROR; }
    rc = launch_tokenize(s.d_text, s.d_off, n_texts, max_len, vocab, ids_out ? s.d_ids : nullptr, s.d_counts, s.d_ids_off, st);
    if (rc != RR_OK) { note_cuda_error(cudaGetLastError()); return rc; }
    e = cudaMemcpyAsync(s.h_counts, s.d_counts, (size_t)n_texts * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && ids_out)
        e = cudaMemcpyAsync(s.h_ids, s.d_ids, (size_t)(total + n_texts) * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && ids_off_out)
        e = cudaMemcpyAsync(s.h_ids_off, s.d_ids_off, (size_t)(n_texts + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { note_cuda_error(e); return RR_CUDA_ERROR; }
    memcpy(counts_out, s.h_counts, (size_t)n_texts * 4);
    if (ids_out) memcpy(ids_out, s.h_ids, (size_t)(total + n_texts) * 4);
    if (ids_off_out) memcpy(ids_off_out, s.h_ids_off, (size_t)(n_texts + 1) * 8);
    return RR_OK;
}

// Device buffers already resident; asynchronous on `stream`.
RR_API int rr_tokenize_batch_device(const uint8_t* d_text, const int64_t* d_text_off, int n_texts, int64_t max_text_bytes,
                                    int32_t vocab, int32_t* d_counts, int32_t* d_ids, int64_t* d_ids_off, void* stream) {
    if (n_texts < 0 || !d_text_off || !d_counts || vocab < 4 || n_texts > 65535 || max_text_bytes < 0) return RR_INVALID_ARGUMENT;
    const int rc = launch_tokenize(d_text, d_text_off, n_texts, max_text_bytes, vocab, d_ids, d_counts, d_ids_off, (cudaStream_t)stream);
    if (rc != RR_OK) note_cuda_error(cudaGetLastError());
    return rc;
}
