// rr_attn.cu — attention over the slot-contiguous KV cache  [slot][kv_head][ctx_max][128] bf16.
//
//  * decode_attn (K8): one query token per row, GQA group of G query heads shares each KV head.
//    HBM-bound (reads ctx * 512 B per (row, kv_head)); K/V tiles of 64 tokens are staged with
//    cp.async (LDGSTS, 16 B, coalesced) into double-buffered shared memory; fp32 online softmax;
//    optional split-KV with a combine pass for small batches.
//  * prefill_attn (K7): causal flash attention, 64 query rows x one head per CTA, bf16
//    mma.sync.m16n8k16 with ldmatrix from XOR-swizzled shared memory.  (<1% of prefill FLOPs at
//    512-token prompts — SURVEY.md §8d; the dense contraction of the path, the projections,
//    runs on tcgen05 in rr_gemm.cu.)
//
// Replaces the remote bedrock:InvokeModel call (reference iam/policy.json:8).
#include "rr_ptx.cuh"
#include "rr_kernels.h"

namespace rr {

constexpr int HD = 128;          // head_dim
constexpr int DT = 64;           // tokens per tile
constexpr int KS_STRIDE = 136;   // padded K row (bf16 elements): 272 B -> conflict-free 16 B reads
constexpr int DEC_THREADS = 128;

template <int G>
struct DecSmem {
    __nv_bfloat16 k[2][DT][KS_STRIDE];
    __nv_bfloat16 v[2][DT][HD];
    float q[G][HD];
    float s_part[2][G][DT];
    float p[DT][G];
    float alpha[G];
    float o_red[G][HD];
};

template <int G>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_kernel(DecodeAttnArgs a) {
    extern __shared__ __align__(16) uint8_t dec_smem_raw[];
    DecSmem<G>& S = *reinterpret_cast<DecSmem<G>*>(dec_smem_raw);
    const int kvh = blockIdx.x, row = blockIdx.y, split = blockIdx.z;
    const int slot = a.slot[row];
    if (slot < 0) return;
    const int ctx = a.pos[row] + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // token range of this split (multiples of DT)
    const int n_tiles_all = (ctx + DT - 1) / DT;
    const int tiles_per = (n_tiles_all + a.kv_splits - 1) / a.kv_splits;
    const int tile0 = split * tiles_per;
    const int tile1 = min(n_tiles_all, tile0 + tiles_per);

    const __nv_bfloat16* kbase = a.k_cache + ((size_t)slot * a.n_kv_heads + kvh) * a.ctx_max * HD;
    const __nv_bfloat16* vbase = a.v_cache + ((size_t)slot * a.n_kv_heads + kvh) * a.ctx_max * HD;

    // q (G heads x 128) -> fp32 smem, pre-scaled by scale*log2(e)
    const float qs = a.scale * 1.4426950408889634f;
    for (int i = tid; i < G * HD; i += DEC_THREADS) {
        const int g = i / HD, d = i % HD;
        S.q[g][d] = __bfloat162float(a.q[(size_t)row * a.n_heads * HD + (kvh * G + g) * HD + d]) * qs;
    }

    auto issue_tile = [&](int tile, int buf) {
        const int t0 = tile * DT;
        // 64 rows x 256 B = 1024 x 16 B chunks for K, same for V; 128 threads -> 8 + 8 each
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = tid + i * DEC_THREADS;       // chunk id
            const int r = c >> 4, cc = c & 15;
            const bool ok = (t0 + r) < ctx;
            const size_t goff = (size_t)(t0 + (ok ? r : 0)) * HD + cc * 8;
            cp_async_16_zfill(&S.k[buf][r][cc * 8], kbase + goff, ok);
            cp_async_16_zfill(&S.v[buf][r][cc * 8], vbase + goff, ok);
        }
        cp_async_commit();
    };

    float m_run = -INFINITY, l_run = 0.f;   // per warp: head g = warp (+4 for G = 8), lane-replicated
    float m_run2 = -INFINITY, l_run2 = 0.f;
    float acc[G][2];
#pragma unroll
    for (int g = 0; g < G; ++g) acc[g][0] = acc[g][1] = 0.f;

    if (tile0 < tile1) issue_tile(tile0, 0);
    for (int tile = tile0; tile < tile1; ++tile) {
        const int buf = (tile - tile0) & 1;
        if (tile + 1 < tile1) {
            issue_tile(tile + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // ---- QK^T: thread = (token, half of the head dim), all G heads
        {
            const int tok = tid & 63, half = tid >> 6;
            float dot[G];
#pragma unroll
            for (int g = 0; g < G; ++g) dot[g] = 0.f;
            const __nv_bfloat16* kr = &S.k[buf][tok][half * 64];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 kv = *reinterpret_cast<const uint4*>(kr + c * 8);
                const float k0 = bf16_lo(kv.x), k1 = bf16_hi(kv.x), k2 = bf16_lo(kv.y), k3 = bf16_hi(kv.y);
                const float k4 = bf16_lo(kv.z), k5 = bf16_hi(kv.z), k6 = bf16_lo(kv.w), k7 = bf16_hi(kv.w);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float4 qa = *reinterpret_cast<const float4*>(&S.q[g][half * 64 + c * 8]);
                    const float4 qb = *reinterpret_cast<const float4*>(&S.q[g][half * 64 + c * 8 + 4]);
                    dot[g] += k0 * qa.x + k1 * qa.y + k2 * qa.z + k3 * qa.w + k4 * qb.x + k5 * qb.y +
                              k6 * qb.z + k7 * qb.w;
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) S.s_part[half][g][tok] = dot[g];
        }
        __syncthreads();

        // ---- online softmax: warp w owns head w (and w + 4 when G == 8)
        {
            const int t0 = tile * DT;
#pragma unroll
            for (int rep = 0; rep < (G + 3) / 4; ++rep) {
                const int g = warp + rep * 4;
                if (g < G) {
                    float s0 = S.s_part[0][g][lane] + S.s_part[1][g][lane];
                    float s1 = S.s_part[0][g][lane + 32] + S.s_part[1][g][lane + 32];
                    if (t0 + lane >= ctx) s0 = -INFINITY;
                    if (t0 + lane + 32 >= ctx) s1 = -INFINITY;
                    float mx = fmaxf(s0, s1);
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    float& mr = rep ? m_run2 : m_run;
                    float& lr = rep ? l_run2 : l_run;
                    const float m_new = fmaxf(mr, mx);          // finite: tile has >= 1 valid token
                    const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
                    float ps = p0 + p1;
                    for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
                    const float al = exp2f(mr - m_new);          // 0 on the first tile (mr = -inf)
                    lr = lr * al + ps;
                    mr = m_new;
                    S.p[lane][g] = p0;
                    S.p[lane + 32][g] = p1;
                    if (lane == 0) S.alpha[g] = al;
                }
            }
        }
        __syncthreads();

        // ---- P V: thread = (dim pair, token half)
        {
            const int dp = tid & 63, th = tid >> 6;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float al = S.alpha[g];
                acc[g][0] *= al;
                acc[g][1] *= al;
            }
#pragma unroll 8
            for (int t = 0; t < 32; ++t) {
                const int tok = th * 32 + t;
                const uint32_t vv = *reinterpret_cast<const uint32_t*>(&S.v[buf][tok][dp * 2]);
                const float v0 = bf16_lo(vv), v1 = bf16_hi(vv);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float p = S.p[tok][g];
                    acc[g][0] += p * v0;
                    acc[g][1] += p * v1;
                }
            }
        }
        __syncthreads();
    }

    // ---- combine the two token halves, normalise, write
    {
        const int dp = tid & 63, th = tid >> 6;
        if (th == 1) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                S.o_red[g][dp * 2] = acc[g][0];
                S.o_red[g][dp * 2 + 1] = acc[g][1];
            }
        }
        // publish per-head (m, l): warp g lane 0
#pragma unroll
        for (int rep = 0; rep < (G + 3) / 4; ++rep) {
            const int g = warp + rep * 4;
            if (g < G && lane == 0) {
                S.s_part[0][g][0] = rep ? m_run2 : m_run;
                S.s_part[0][g][1] = rep ? l_run2 : l_run;
            }
        }
        __syncthreads();
        if (th == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float o0 = acc[g][0] + S.o_red[g][dp * 2];
                const float o1 = acc[g][1] + S.o_red[g][dp * 2 + 1];
                const float m = S.s_part[0][g][0], l = S.s_part[0][g][1];
                const int head = kvh * G + g;
                if (a.kv_splits == 1) {
                    const float inv = l > 0.f ? 1.f / l : 0.f;
                    *reinterpret_cast<uint32_t*>(a.out + (size_t)row * a.n_heads * HD + head * HD + dp * 2) =
                        pack_bf16(o0 * inv, o1 * inv);
                } else {
                    float* w = a.ws + (((size_t)row * a.n_heads + head) * a.kv_splits + split) * (HD + 2);
                    w[dp * 2] = o0;
                    w[dp * 2 + 1] = o1;
                    if (dp == 0) { w[HD] = m; w[HD + 1] = l; }
                }
            }
        }
    }
}

// combine split-KV partials: grid (n_heads, rows), 64 threads (dim pairs)
__global__ void decode_attn_combine_kernel(DecodeAttnArgs a) {
    const int head = blockIdx.x, row = blockIdx.y;
    if (a.slot[row] < 0) return;
    const float* w = a.ws + ((size_t)row * a.n_heads + head) * a.kv_splits * (HD + 2);
    float m = -INFINITY;
    for (int s = 0; s < a.kv_splits; ++s) m = fmaxf(m, w[s * (HD + 2) + HD]);
    float l = 0.f, o0 = 0.f, o1 = 0.f;
    const int dp = threadIdx.x;
    for (int s = 0; s < a.kv_splits; ++s) {
        const float* ws = w + s * (HD + 2);
        const float ms = ws[HD];
        if (ms == -INFINITY) continue;                // empty split
        const float sc = exp2f(ms - m);
        l += ws[HD + 1] * sc;
        o0 += ws[dp * 2] * sc;
        o1 += ws[dp * 2 + 1] * sc;
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    *reinterpret_cast<uint32_t*>(a.out + (size_t)row * a.n_heads * HD + head * HD + dp * 2) =
        pack_bf16(o0 * inv, o1 * inv);
}

size_t decode_attn_ws_bytes(int rows, int n_heads, int kv_splits) {
    return kv_splits > 1 ? (size_t)rows * n_heads * kv_splits * (HD + 2) * sizeof(float) : 0;
}

template <int G>
static void launch_dec(const DecodeAttnArgs& a, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(decode_attn_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(DecSmem<G>));
        attr = true;
    }
    dim3 grid(a.n_kv_heads, a.rows, a.kv_splits);
    decode_attn_kernel<G><<<grid, DEC_THREADS, sizeof(DecSmem<G>), st>>>(a);
    if (a.kv_splits > 1) decode_attn_combine_kernel<<<dim3(a.n_heads, a.rows), 64, 0, st>>>(a);
}

void launch_decode_attn(const DecodeAttnArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return;
    const int G = a.n_heads / a.n_kv_heads;
    switch (G) {
        case 1: launch_dec<1>(a, st); break;
        case 2: launch_dec<2>(a, st); break;
        case 4: launch_dec<4>(a, st); break;
        case 8: launch_dec<8>(a, st); break;
        default: break;
    }
}

// =================================================================================================
// Prefill: causal flash attention with mma.sync (bf16 in, fp32 accumulate).
// CTA = 4 warps, 64 query rows of one head of one sequence; KV tiles of 64 tokens read from the
// KV cache (RoPE'd K already appended by rope_kv_kernel).  smem rows are 256 B = 16 chunks of 16 B,
// physical chunk = chunk ^ (row & 7)  -> conflict-free ldmatrix.
constexpr int PF_THREADS = 128;
constexpr int PF_Q = 64;

__device__ __forceinline__ uint32_t swz(int row, int chunk) {   // byte offset inside a [rows][128] bf16 tile
    return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(PF_THREADS)
prefill_attn_kernel(PrefillAttnArgs a) {
    extern __shared__ __align__(128) uint8_t pf_smem[];
    uint8_t* sQ = pf_smem;                       // 64 x 256 B
    uint8_t* sK = pf_smem + 16384;               // 2 x 64 x 256 B
    uint8_t* sV = pf_smem + 16384 + 32768;       // 2 x 64 x 256 B

    const int seq = blockIdx.z, head = blockIdx.y;
    // heaviest (last) q tiles first
    const int tok0 = a.seq_start[seq];
    const int len = a.seq_start[seq + 1] - tok0;
    const int n_qt = (len + PF_Q - 1) / PF_Q;
    const int qt = n_qt - 1 - (int)blockIdx.x;
    if (qt < 0) return;
    const int slot = a.seq_slot[seq];
    const int G = a.n_heads / a.n_kv_heads;
    const int kvh = head / G;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    const __nv_bfloat16* kbase = a.k_cache + ((size_t)slot * a.n_kv_heads + kvh) * a.ctx_max * HD;
    const __nv_bfloat16* vbase = a.v_cache + ((size_t)slot * a.n_kv_heads + kvh) * a.ctx_max * HD;
    const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);

    // ---- load Q tile (64 rows x 16 chunks)
    for (int c = tid; c < PF_Q * 16; c += PF_THREADS) {
        const int r = c >> 4, cc = c & 15;
        const int qi = qt * PF_Q + r;
        const bool ok = qi < len;
        const __nv_bfloat16* src = a.q + (size_t)(tok0 + (ok ? qi : 0)) * a.n_heads * HD + head * HD + cc * 8;
        cp_async_16_zfill(sQ + swz(r, cc), src, ok);
    }
    auto issue_kv = [&](int tile, int buf) {
        const int t0 = tile * DT;
        for (int c = tid; c < DT * 16; c += PF_THREADS) {
            const int r = c >> 4, cc = c & 15;
            const bool ok = (t0 + r) < len;
            const size_t goff = (size_t)(t0 + (ok ? r : 0)) * HD + cc * 8;
            cp_async_16_zfill(sK + buf * 16384 + swz(r, cc), kbase + goff, ok);
            cp_async_16_zfill(sV + buf * 16384 + swz(r, cc), vbase + goff, ok);
        }
        cp_async_commit();
    };
    const int n_kt = qt + 1;                       // causal: kv tiles 0..qt
    issue_kv(0, 0);                                // group 0 = Q + KV tile 0

    uint32_t qf[8][4];
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float sc = a.scale * 1.4426950408889634f;
    const int q_row0 = qt * PF_Q + warp * 16 + g;  // rows q_row0 and q_row0 + 8

    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) {
            issue_kv(kt + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int r = warp * 16 + (lane & 15);
                ldmatrix_x4(qf[ks], sQ_u + swz(r, ks * 2 + (lane >> 4)));
            }
        }
        // ---- S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {        // pairs of n-tiles (16 tokens)
                uint32_t kb[4];
                const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                ldmatrix_x4(kb, sK_u + buf * 16384 + swz(r, ks * 2 + ((lane >> 3) & 1)));
                mma_bf16_16816(s[np * 2], qf[ks], kb[0], kb[1]);
                mma_bf16_16816(s[np * 2 + 1], qf[ks], kb[2], kb[3]);
            }
        }
        // ---- mask + online softmax
        const int kv0 = kt * DT;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kv = kv0 + nt * 8 + t * 2 + (e & 1);
                const int qr = q_row0 + ((e >> 1) << 3);
                float v = s[nt][e] * sc;
                if (kv > qr || kv >= len) v = -INFINITY;
                s[nt][e] = v;
                if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        // rows past the end of the sequence are fully masked: keep them at exp2(-inf - 0) = 0
        const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;
        const float al0 = exp2f(m0 - ms0), al1 = exp2f(m1 - ms1);
        m0 = mn0; m1 = mn1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - ms0); s[nt][1] = exp2f(s[nt][1] - ms0);
            s[nt][2] = exp2f(s[nt][2] - ms1); s[nt][3] = exp2f(s[nt][3] - ms1);
            ps0 += s[nt][0] + s[nt][1];
            ps1 += s[nt][2] + s[nt][3];
        }
        ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
        ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
        l0 = l0 * al0 + ps0;
        l1 = l1 * al1 + ps1;
#pragma unroll
        for (int i = 0; i < 16; ++i) { o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1; }
        // ---- O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {           // 16 tokens per k-step
            uint32_t pa[4];
            pa[0] = pack_bf16(s[ks * 2][0], s[ks * 2][1]);
            pa[1] = pack_bf16(s[ks * 2][2], s[ks * 2][3]);
            pa[2] = pack_bf16(s[ks * 2 + 1][0], s[ks * 2 + 1][1]);
            pa[3] = pack_bf16(s[ks * 2 + 1][2], s[ks * 2 + 1][3]);
#pragma unroll
            for (int dpair = 0; dpair < 8; ++dpair) {   // pairs of d n-tiles (16 dims)
                uint32_t vb[4];
                const int r = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                ldmatrix_x4_trans(vb, sV_u + buf * 16384 + swz(r, dpair * 2 + (lane >> 4)));
                mma_bf16_16816(o[dpair * 2], pa, vb[0], vb[1]);
                mma_bf16_16816(o[dpair * 2 + 1], pa, vb[2], vb[3]);
            }
        }
        __syncthreads();
    }
    // ---- normalise + store
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    const int qi0 = q_row0, qi1 = q_row0 + 8;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
        const int d = nt * 8 + t * 2;
        if (qi0 < len)
            *reinterpret_cast<uint32_t*>(a.out + (size_t)(tok0 + qi0) * a.n_heads * HD + head * HD + d) =
                pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
        if (qi1 < len)
            *reinterpret_cast<uint32_t*>(a.out + (size_t)(tok0 + qi1) * a.n_heads * HD + head * HD + d) =
                pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
    }
}

void launch_prefill_attn(const PrefillAttnArgs& a, cudaStream_t st) {
    if (a.n_seqs <= 0 || a.max_len <= 0) return;
    constexpr int smem = 16384 + 2 * 32768;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(prefill_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr = true;
    }
    dim3 grid((a.max_len + PF_Q - 1) / PF_Q, a.n_heads, a.n_seqs);
    prefill_attn_kernel<<<grid, PF_THREADS, smem, st>>>(a);
}

}  // namespace rr
