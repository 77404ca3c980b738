// rr_api.cu — extern "C" surface of librr_b200.so for the kernel-level entry points
// (include/rr_b200.h §2, §3).  Router: rr_router.cu.  Engine: rr_engine.cu.
#include "rr_kernels.h"

#include <string.h>

#define RR_API extern "C" __attribute__((visibility("default")))

namespace rr {
thread_local char g_last_cuda_error[256] = "";
void note_cuda_error(cudaError_t e) {
    if (e != cudaSuccess) {
        strncpy(g_last_cuda_error, cudaGetErrorString(e), sizeof(g_last_cuda_error) - 1);
    }
}
int check_last(void) {
    cudaError_t e = cudaGetLastError();
    note_cuda_error(e);
    return e == cudaSuccess ? RR_OK : RR_CUDA_ERROR;
}
}  // namespace rr

using namespace rr;

RR_API const char* rr_version(void) { return "rr_b200 0.1 (sm_100a)"; }

RR_API const char* rr_strerror(int rc) {
    switch (rc) {
        case RR_OK: return "ok";
        case RR_RATE_LIMITED: return "rate limited (429)";
        case RR_NO_GROUP: return "unknown model group";
        case RR_INTERNAL: return "internal error";
        case RR_INVALID_ARGUMENT: return "invalid argument";
        case RR_CUDA_ERROR: return "CUDA error";
        case RR_TIMEOUT: return "timeout";
        case RR_BACKEND_FAILED: return "backend failed";
    }
    return "unknown";
}

RR_API const char* rr_last_cuda_error(void) { return g_last_cuda_error; }

// ---------------------------------------------------------------- tokenizer (K2)
// Byte-level: id 1 = BOS, then 3 + byte value (0/1/2 reserved: pad/bos/eos), folded into the
// model's vocabulary by modulo when vocab < 259.
RR_API int rr_count_tokens(const uint8_t* text, size_t n_bytes, int32_t* n_tokens) {
    if (!n_tokens || (!text && n_bytes)) return RR_INVALID_ARGUMENT;
    *n_tokens = (int32_t)n_bytes + 1;
    return RR_OK;
}

RR_API int rr_tokenize(const uint8_t* text, size_t n_bytes, int32_t vocab, int32_t* ids,
                       int32_t max_ids, int32_t* n_ids) {
    if (!ids || !n_ids || vocab < 4 || (!text && n_bytes)) return RR_INVALID_ARGUMENT;
    if ((size_t)max_ids < n_bytes + 1) return RR_INVALID_ARGUMENT;
    ids[0] = 1;
    for (size_t i = 0; i < n_bytes; ++i) {
        int32_t t = 3 + (int32_t)text[i];
        ids[i + 1] = t < vocab ? t : 3 + (t - 3) % (vocab - 3);
    }
    *n_ids = (int32_t)n_bytes + 1;
    return RR_OK;
}

// ---------------------------------------------------------------- kernels
RR_API int rr_gemm_bf16(const void* A, int rowsA, int ldA, const void* B, int rowsB, int ldB, int K,
                        void* out, int ldo, int ld_rows, int splits, int mode, int bn,
                        void* stream) {
    GemmPlan p;
    int rc = gemm_plan_init(&p, A, rowsA, ldA, B, rowsB, ldB, K, out, ldo, ld_rows, splits, mode, bn);
    if (rc != RR_OK) return rc;
    rc = gemm_launch(p, (cudaStream_t)stream);
    if (rc != RR_OK) note_cuda_error(cudaPeekAtLastError());
    return rc;
}
