// rr_attn_tc.cu — prefill attention (K7) on tcgen05: causal flash attention with both contractions on
// the 5th-gen tensor cores, S and per-tile O in TMEM, K/V tiles staged by TMA.
//
// Persistent kernel, one CTA per SM, 12 warps = 3 warpgroups (one warp of each per SM sub-partition):
//   warp 0      TMA producer  : Q tiles (128 rows x 128 dims, two query heads that share a kv head) and a
//                               3-deep ring of K/V tiles (64 keys x 128 dims each) from the KV cache
//   warp 1      MMA issuer    : S_h = Q_h K^T   (M128 N64  K16 x 8, K-major smem operands)
//                               O_h = P_h V     (M128 N128 K16 x 4, P K-major from smem, V MN-major as
//                               it lies in the cache: keys are rows, the head dim is contiguous)
//   warps 2-3   idle (fill warpgroup 0, which drops to 24 registers per thread: setmaxnreg)
//   warps 4-7   softmax, head 0: thread = one query row (TMEM lane): scale, causal mask, running max,
//   warps 8-11  softmax, head 1: exp2, row sum, P -> bf16 -> swizzled smem; folds the previous tile's
//                               O from TMEM into fp32 registers (acc = acc * alpha + O), normalises and
//                               stores at the end.  240 registers per thread (setmaxnreg.inc).
// S is double-buffered in TMEM (2 x 64 columns per head) so QK^T of tile j+1 runs under the softmax of
// tile j; the two heads alternate on the tensor pipe.  TMEM: 2 x (64 + 64 + 128) = 512 columns.
//
// Work item = (sequence, head pair, 128-row query tile), heaviest (last) query tiles first, strided
// over the CTAs.  Only even GQA group sizes take this path (the pair must share its kv head); the
// mma.sync kernel in rr_attn.cu serves the rest (MHA models).
//
// Replaces the remote bedrock:InvokeModel call (reference iam/policy.json:8).
#include "rr_ptx.cuh"
#include "rr_launch.cuh"
#include "rr_kernels.h"
#include <cstdlib>

namespace rr {
namespace {

constexpr int TQ = 128;                 // query rows per head per work item
constexpr int TKV = 64;                 // keys per tile
constexpr int NH = 2;                   // query heads per CTA
constexpr int NS = 3;                   // K/V ring depth
constexpr int TC_THREADS = 128 + 128 * NH;
constexpr uint32_t Q_BYTES = 32768;     // [2 dim halves][128 rows][128 B]
constexpr uint32_t P_BYTES = 16384;     // [128 rows][128 B]  (64 keys)
constexpr uint32_t KV_BYTES = 16384;    // [2 dim halves][64 keys][128 B]
constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_P = OFF_Q + NH * Q_BYTES;
constexpr uint32_t OFF_KV = OFF_P + NH * 2 * P_BYTES;
constexpr uint32_t OFF_BAR = OFF_KV + NS * 2 * KV_BYTES;
constexpr uint32_t TC_SMEM = OFF_BAR + 256 + 1024;   // + barriers + manual 1024 B alignment slack

struct Bars {
    uint64_t q_full, q_empty;
    uint64_t kv_full[NS], kv_empty[NS];
    uint64_t s_full[NH][2], s_empty[NH][2];
    uint64_t p_full[NH][2];
    uint64_t o_full[NH], o_empty[NH];
    uint32_t tmem;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), split in 32-bit halves so the issuing
// thread only carries the low words.  High word (same for every operand here): SBO = 1024 B (8 rows of
// 128 B) [32,46), version 1 [46,48), SWIZZLE_128B [61,64).  Low word: start address >> 4 [0,14), LBO >> 4
// [16,30).  K-major operands (Q, K, P): LBO unused (1).  V is the MN-major B operand exactly as it lies
// in the cache -- 64 dims (128 B) contiguous, 8 keys per 1024 B swizzle atom (SBO), the second 64-dim
// half one TMA box (8192 B) further (LBO): canonical layout Sw<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_kmajor(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mnmajor(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16); }
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
        : "memory");
}
constexpr uint32_t IDESC_S = umma_idesc_bf16_f32(TQ, TKV);
constexpr uint32_t IDESC_PV = umma_idesc_bf16_f32(TQ, 128) | (1u << 16);   // b_major = MN

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Work {
    int qt, seq, pair, tok0, len, n_kt;
};
__device__ __forceinline__ bool decode_work(const PrefillAttnArgs& a, int w, int n_pairs, int max_qt, Work& k) {
    const int per_qt = a.n_seqs * n_pairs;
    const int qrev = w / per_qt, rem = w - qrev * per_qt;
    k.qt = max_qt - 1 - qrev;
    k.seq = rem / n_pairs;
    k.pair = rem - k.seq * n_pairs;
    k.tok0 = a.seq_start[k.seq];
    k.len = a.seq_start[k.seq + 1] - k.tok0;
    if (k.qt * TQ >= k.len) return false;
    const int causal = 2 * k.qt + 2, avail = (k.len + TKV - 1) / TKV;
    k.n_kt = causal < avail ? causal : avail;
    return true;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
prefill_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const PrefillAttnArgs a) {
    extern __shared__ uint8_t tc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    griddep_launch();
    const int tr_slot = trace_begin(TR_ATTN_PF);
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(&bars->q_full, 1); mbar_init(&bars->q_empty, 1);
        for (int s = 0; s < NS; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
        for (int h = 0; h < NH; ++h) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(&bars->s_full[h][b], 1); mbar_init(&bars->s_empty[h][b], 128);
                mbar_init(&bars->p_full[h][b], 128);
            }
            mbar_init(&bars->o_full[h], 1); mbar_init(&bars->o_empty[h], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&bars->tmem);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = bars->tmem;

    const int n_pairs = a.n_heads / NH;
    const int max_qt = (a.max_len + TQ - 1) / TQ;
    const int n_work = max_qt * a.n_seqs * n_pairs;
    const int G = a.n_heads / a.n_kv_heads;

    griddep_wait();
    trace_dep(tr_slot);

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t kv_it = 0, q_it = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                Work k;
                if (!decode_work(a, w, n_pairs, max_qt, k)) continue;
                const int head0 = k.pair * NH, kvh = head0 / G;
                const int kv_row0 = (a.seq_slot[k.seq] * a.n_kv_heads + kvh) * a.ctx_max;
                mbar_wait(&bars->q_empty, (q_it & 1) ^ 1);
                mbar_arrive_expect_tx(&bars->q_full, NH * Q_BYTES);
                for (int h = 0; h < NH; ++h)
                    for (int c = 0; c < 2; ++c)
                        tma_load_2d(smem + OFF_Q + h * Q_BYTES + c * 16384, &tmQ, &bars->q_full,
                                    (head0 + h) * 128 + c * 64, k.tok0 + k.qt * TQ);
                ++q_it;
                for (int j = 0; j < k.n_kt; ++j, ++kv_it) {
                    const uint32_t s = kv_it % NS;
                    mbar_wait(&bars->kv_empty[s], ((kv_it / NS) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->kv_full[s], 2 * KV_BYTES);
                    uint8_t* dst = smem + OFF_KV + s * 2 * KV_BYTES;
                    for (int c = 0; c < 2; ++c) {
                        tma_load_2d(dst + c * 8192, &tmK, &bars->kv_full[s], c * 64, kv_row0 + j * TKV);
                        tma_load_2d(dst + KV_BYTES + c * 8192, &tmV, &bars->kv_full[s], c * 64, kv_row0 + j * TKV);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t kv_it = 0, q_it = 0, s_it = 0, o_it = 0;
            const uint32_t q_u = smem_u32(smem + OFF_Q), p_u = smem_u32(smem + OFF_P), kv_u = smem_u32(smem + OFF_KV);
            auto pv = [&](uint32_t kvit_t, uint32_t sit_t) {
                const uint32_t v_u = kv_u + (kvit_t % NS) * 2 * KV_BYTES + KV_BYTES, b = sit_t & 1;
                for (int h = 0; h < NH; ++h) {
                    mbar_wait(&bars->p_full[h][b], (sit_t >> 1) & 1);
                    mbar_wait(&bars->o_empty[h], (o_it & 1) ^ 1);
                    tcgen05_fence_after();
                    const uint32_t pa = p_u + (h * 2 + b) * P_BYTES;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_lo(tmem + h * 256 + 128, desc_lo_kmajor(pa) + ks * 2, desc_lo_mnmajor(v_u) + ks * 128,
                                IDESC_PV, ks > 0);
                    umma_commit(&bars->o_full[h]);
                }
                umma_commit(&bars->kv_empty[kvit_t % NS]);
                ++o_it;
            };
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                Work k;
                if (!decode_work(a, w, n_pairs, max_qt, k)) continue;
                mbar_wait(&bars->q_full, q_it & 1);
                ++q_it;
                for (int j = 0; j < k.n_kt; ++j, ++kv_it, ++s_it) {
                    const uint32_t s = kv_it % NS, b = s_it & 1;
                    mbar_wait(&bars->kv_full[s], (kv_it / NS) & 1);
                    const uint32_t k_u = kv_u + s * 2 * KV_BYTES;
                    for (int h = 0; h < NH; ++h) {
                        mbar_wait(&bars->s_empty[h][b], ((s_it >> 1) & 1) ^ 1);
                        tcgen05_fence_after();
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_lo(tmem + h * 256 + b * 64,
                                    desc_lo_kmajor(q_u + h * Q_BYTES) + (ks >> 2) * 1024 + (ks & 3) * 2,
                                    desc_lo_kmajor(k_u) + (ks >> 2) * 512 + (ks & 3) * 2, IDESC_S, ks > 0);
                        umma_commit(&bars->s_full[h][b]);
                    }
                    if (j == k.n_kt - 1) umma_commit(&bars->q_empty);
                    if (j > 0) pv(kv_it - 1, s_it - 1);
                }
                pv(kv_it - 1, s_it - 1);
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ------------------------------------------------------------------ softmax + output
        const int h = (warp - 4) >> 2, quarter = warp & 3, row = quarter * 32 + lane;
        const uint32_t t_base = tmem + ((uint32_t)(quarter * 32) << 16) + h * 256;
        const float sc = a.scale * 1.4426950408889634f;
        uint32_t s_it = 0, o_it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            Work k;
            if (!decode_work(a, w, n_pairs, max_qt, k)) continue;
            const int head = k.pair * NH + h;
            const int qi = k.qt * TQ + row;                       // query index inside the sequence
            const int last_key = qi < k.len ? qi : k.len - 1;     // causal / ragged bound (inclusive)
            float acc[128];
#pragma unroll
            for (int i = 0; i < 128; ++i) acc[i] = 0.f;
            float m = -INFINITY, l = 0.f, alpha_prev = 0.f;

            auto fold = [&]() {      // acc = acc * alpha(prev tile) + O(prev tile)
                mbar_wait(&bars->o_full[h], o_it & 1);
                tcgen05_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_base + 128 + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[c * 32 + i] = fmaf(acc[c * 32 + i], alpha_prev, __uint_as_float(v[i]));
                }
                tcgen05_fence_before();
                mbar_arrive(&bars->o_empty[h]);
                ++o_it;
            };

            for (int j = 0; j < k.n_kt; ++j, ++s_it) {
                const uint32_t b = s_it & 1;
                mbar_wait(&bars->s_full[h][b], (s_it >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t tS = t_base + b * 64;
                const int key0 = j * TKV;
                const bool need_mask = (key0 + TKV - 1 > k.qt * TQ) || (key0 + TKV > k.len);
                const int lim = last_key - key0;                  // keep columns c <= lim
                // ---- pass 1: row max
                float mx = -INFINITY;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tS + half * 32, v);
                    tmem_ld_wait();
                    if (need_mask) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (half * 32 + i <= lim) mx = fmaxf(mx, __uint_as_float(v[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                }
                const float m_new = fmaxf(m, mx * sc);
                const float ms = m_new == -INFINITY ? 0.f : m_new;
                const float alpha = ex2(m - ms);
                m = m_new;
                // ---- pass 2: P = exp2(S * sc - m), row sum, bf16 -> swizzled smem (K-major A operand)
                float ps = 0.f;
                uint8_t* prow = smem + OFF_P + (h * 2 + b) * P_BYTES + row * 128;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tS + half * 32, v);
                    tmem_ld_wait();
                    if (half == 1) {                              // S(j) fully read: release the buffer
                        tcgen05_fence_before();
                        mbar_arrive(&bars->s_empty[h][b]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float e[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            e[i] = ex2(fmaf(__uint_as_float(v[u * 8 + i]), sc, -ms));
                            if (need_mask && half * 32 + u * 8 + i > lim) e[i] = 0.f;
                            ps += e[i];
                        }
                        uint4 pk;
                        pk.x = pack_bf16(e[0], e[1]);
                        pk.y = pack_bf16(e[2], e[3]);
                        pk.z = pack_bf16(e[4], e[5]);
                        pk.w = pack_bf16(e[6], e[7]);
                        *reinterpret_cast<uint4*>(prow + (((half * 4 + u) ^ (row & 7)) << 4)) = pk;
                    }
                }
                l = l * alpha + ps;
                fence_proxy_async();
                mbar_arrive(&bars->p_full[h][b]);
                if (j > 0) fold();
                alpha_prev = alpha;
            }
            fold();
            // ---- normalise + store (row = one token, 128 contiguous dims of this head)
            if (qi < k.len) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
                __nv_bfloat16* dst = a.out + (size_t)(k.tok0 + qi) * a.n_heads * a.head_dim + (size_t)head * a.head_dim;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    if (u * 8 >= a.head_dim) break;               // dims beyond the true head dim are padding
                    uint4 pk;
                    pk.x = pack_bf16(acc[u * 8 + 0] * inv, acc[u * 8 + 1] * inv);
                    pk.y = pack_bf16(acc[u * 8 + 2] * inv, acc[u * 8 + 3] * inv);
                    pk.z = pack_bf16(acc[u * 8 + 4] * inv, acc[u * 8 + 5] * inv);
                    pk.w = pack_bf16(acc[u * 8 + 6] * inv, acc[u * 8 + 7] * inv);
                    *reinterpret_cast<uint4*>(dst + u * 8) = pk;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem);
    }
    trace_end(tr_slot);
}

}  // namespace

int g_use_attn_tc = -1;

bool prefill_attn_tc_eligible(const PrefillAttnArgs& a) {
    if (g_use_attn_tc < 0) g_use_attn_tc = std::getenv("RR_NO_ATTN_TC") ? 0 : 1;
    const int G = a.n_kv_heads > 0 ? a.n_heads / a.n_kv_heads : 0;
    return g_use_attn_tc && a.has_maps && G >= 2 && G % 2 == 0 && a.ctx_max % TKV == 0 && a.head_dim % 8 == 0;
}

// q: [q_rows, n_heads*128]; caches: [kv_rows = n_slots*n_kv_heads*ctx_max, 128].  Boxes: Q 64 x 128 rows,
// K/V 64 x 64 rows, 128-byte swizzle.
int prefill_attn_make_maps(PrefillAttnArgs* a, long long q_rows, long long kv_rows) {
    a->has_maps = 0;
    if (q_rows <= 0 || kv_rows <= 0 || q_rows > 0x7fffffffLL || kv_rows > 0x7fffffffLL) return RR_ERR_ARG;
    int rc = make_tmap_bf16_2d(&a->tmQ, a->q, (int)q_rows, a->n_heads * 128, a->n_heads * 128, TQ);
    if (rc != RR_OK) return rc;
    rc = make_tmap_bf16_2d(&a->tmK, a->k_cache, (int)kv_rows, 128, 128, TKV);
    if (rc != RR_OK) return rc;
    rc = make_tmap_bf16_2d(&a->tmV, a->v_cache, (int)kv_rows, 128, 128, TKV);
    if (rc != RR_OK) return rc;
    a->has_maps = 1;
    return RR_OK;
}

int launch_prefill_attn_tc(const PrefillAttnArgs& a, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(prefill_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess)
            return RR_ERR_CUDA;
        attr = true;
    }
    const int n_work = ((a.max_len + TQ - 1) / TQ) * a.n_seqs * (a.n_heads / NH);
    const int grid = n_work < num_sms() ? n_work : num_sms();
    cudaError_t e = launch_pdl(prefill_attn_tc_kernel, dim3(grid), dim3(TC_THREADS), (size_t)TC_SMEM, st,
                               a.tmQ, a.tmK, a.tmV, a);
    return e == cudaSuccess ? RR_OK : RR_ERR_CUDA;
}

void rr_trace_set_attn_tc(unsigned long long* p) { rr_trace_set_local(p); }

}  // namespace rr
