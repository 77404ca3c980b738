// rr_attn.cu — attention over the slot-contiguous KV cache  [slot][kv_head][ctx_max][128] bf16.
//
//  * decode attention (K8) lives in rr_attn_decode.cu.
//  * prefill attention (K7) for even GQA group sizes runs on tcgen05 (rr_attn_tc.cu); launch_prefill_attn() below
//    dispatches to it.
//  * prefill_attn_kernel here: the mma.sync version -- causal flash attention, 64 query rows x one head per CTA, bf16
//    mma.sync.m16n8k16 with ldmatrix from XOR-swizzled shared memory.  Serves MHA models (group size 1: the tcgen05
//    kernel pairs two query heads on one kv head) and RR_NO_ATTN_TC=1.
//
// Replaces the remote bedrock:InvokeModel call (reference iam/policy.json:8).
#include "rr_ptx.cuh"
#include "rr_launch.cuh"
#include "rr_kernels.h"

namespace rr {

constexpr int HD = 128;          // head_dim
constexpr int DT = 64;           // KV tokens per tile

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
    griddep_launch();
    const int tr_slot = trace_begin(TR_ATTN_PF);
    griddep_wait();
    trace_dep(tr_slot);

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
        if (d >= a.head_dim) break;                 // columns beyond the true head dim are zero padding
        if (qi0 < len)
            *reinterpret_cast<uint32_t*>(a.out + (size_t)(tok0 + qi0) * a.n_heads * a.head_dim + head * a.head_dim + d) =
                pack_bf16(o[nt][0] * inv0, o[nt][1] * inv0);
        if (qi1 < len)
            *reinterpret_cast<uint32_t*>(a.out + (size_t)(tok0 + qi1) * a.n_heads * a.head_dim + head * a.head_dim + d) =
                pack_bf16(o[nt][2] * inv1, o[nt][3] * inv1);
    }
    trace_end(tr_slot);
}

void launch_prefill_attn(const PrefillAttnArgs& a, cudaStream_t st) {
    if (a.n_seqs <= 0 || a.max_len <= 0) return;
    if (prefill_attn_tc_eligible(a)) { launch_prefill_attn_tc(a, st); return; }
    constexpr int smem = 16384 + 2 * 32768;
    static std::atomic<uint64_t> attr{0};
    if (ensure_dyn_smem(prefill_attn_kernel, smem, attr) != cudaSuccess) return;
    dim3 grid((a.max_len + PF_Q - 1) / PF_Q, a.n_heads, a.n_seqs);
    launch_pdl(prefill_attn_kernel, grid, dim3(PF_THREADS), (size_t)smem, st, a);
}

void rr_trace_set_attn(unsigned long long* p) { rr_trace_set_local(p); }

}  // namespace rr
