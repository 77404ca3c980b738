// rr_chain.cu — persistent "chain" kernel for the decode step: everything between two attention kernels,
//
//     O-proj GEMM -> [add + RMSNorm] -> gate/up GEMM (+SiLU*mul) -> down GEMM -> [add + RMSNorm] -> next GEMM
//                                                                                         (QKV of layer l+1, or lm_head)
//
// in ONE launch of one CTA per SM.  Why: at 64 rows every GEMM of the step is a 10-40 us weight stream, and each
// kernel boundary costs ~4 us of drain + launch + ramp (profiles/: in-graph timeline), i.e. ~20 % of the step.
// Here the phases are separated by grid-wide barriers (one atomic arrive per CTA, acquire-spin by the few threads
// that need the data) and, more importantly, the TMA producer warp never stops: it walks straight into the next
// phase and requests that phase's WEIGHT tiles (constant data) while the previous phase is still draining; only the
// activation tiles wait for the barrier.  The tcgen05 pipeline (smem ring, TMEM double buffer, warp roles) is the
// one of rr_gemm.cu, its state simply carries over from phase to phase.
//
// Replaces (together with rr_attn_decode.cu) the remote bedrock:InvokeModel call (reference iam/policy.json:8).
#include "rr_gemm_dev.cuh"

namespace rr {

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void grid_wait(const unsigned* ctr, unsigned target) {
    while (ld_acquire_u32(ctr) < target) {
    }
}
// call after a CTA-level barrier that covers every thread whose writes must be published
__device__ __forceinline__ void grid_arrive(unsigned* ctr) {
    __threadfence();
    atomicAdd(ctr, 1u);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
decode_chain_kernel(const __grid_constant__ ChainArgs a) {
    using Cfg = GemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + kStages * Cfg::kStageBytesA;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* red = reinterpret_cast<float*>(tmem_ptr + 2);                               // [4] norm reduction
    float* stage_silu = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes + 256);   // 16 KB

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_cta = gridDim.x;
    griddep_launch();
    const int tr_slot = trace_begin(TR_MISC);

    if (warp == 0 && elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { tma_prefetch_desc(&a.g[j].tmA); tma_prefetch_desc(&a.g[j].tmB); }
    }
    if (warp == 1) {
        if (elect_one()) {
            for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // barrier counters: 0 after G0 (O), 1 after N0, 2 after G1 (gate/up), 3 after G2 (down), 4 after N1
    // GEMM phase j reads activations published by: j=0 previous kernel (PDL), j=1 ctr[1], j=2 ctr[2], j=3 ctr[4]
    if (warp == 0) {
        // ===================== TMA producer: runs ahead across phases =====================
        if (elect_one()) {
            const uint64_t polA = l2_policy_evict_first(), polB = l2_policy_evict_last();
            int stage = 0;
            uint32_t phase = 0;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const ChainGemm& G = a.g[j];
                WorkSched sched;
                sched.init(G.rowsA, a.rows, G.K, G.splits, BN);
                WorkItem t;
                bool have = sched.next(t);
                // weights of the first `pre` k-blocks are requested before the activations exist
                int pre = 0, pstage = stage;
                uint32_t pphase = phase;
                if (have) {
                    pre = min(kStages, t.kb1 - t.kb0);
                    for (int i = 0; i < pre; ++i) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        tma_load_2d_hint(smemA + stage * Cfg::kStageBytesA, &G.tmA, &full_bar[stage], (t.kb0 + i) * BLOCK_K,
                                         t.a_tile * BLOCK_A, polA);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
                if (j == 0) griddep_wait();
                else grid_wait(&a.counters[j == 1 ? 1 : (j == 2 ? 2 : 4)], n_cta);
                asm volatile("fence.proxy.async;" ::: "memory");        // generic-proxy writes of other SMs -> TMA reads
                if (j == 0) trace_dep(tr_slot);
                if (have) {
                    for (int i = 0; i < pre; ++i) {
                        tma_load_2d_hint(smemB + pstage * Cfg::kStageBytesB, &G.tmB, &full_bar[pstage], (t.kb0 + i) * BLOCK_K, 0, polB);
                        if (++pstage == kStages) { pstage = 0; pphase ^= 1; }
                    }
                    t.kb0 += pre;
                }
                while (have) {
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        tma_load_2d_hint(smemA + stage * Cfg::kStageBytesA, &G.tmA, &full_bar[stage], kb * BLOCK_K,
                                         t.a_tile * BLOCK_A, polA);
                        tma_load_2d_hint(smemB + stage * Cfg::kStageBytesB, &G.tmB, &full_bar[stage], kb * BLOCK_K, 0, polB);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    have = sched.next(t);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(BLOCK_A, BN);
            int stage = 0, it = 0;
            uint32_t phase = 0;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const ChainGemm& G = a.g[j];
                WorkSched sched;
                sched.init(G.rowsA, a.rows, G.K, G.splits, BN);
                WorkItem t;
                for (; sched.next(t); ++it) {
                    const int acc = it & 1;
                    mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(smemA + stage * Cfg::kStageBytesA));
                        const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(smemB + stage * Cfg::kStageBytesB));
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tmem_full[acc]);
                }
            }
        }
    } else {
        // ===================== epilogue warps: GEMM epilogues + the two norm phases =====================
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const int etid = (warp - 2) * 32 + lane;            // 0..127
        int it = 0;
        griddep_wait();

        auto gemm_epilogue = [&](const ChainGemm& G) {
            WorkSched sched;
            sched.init(G.rowsA, a.rows, G.K, G.splits, BN);
            WorkItem t;
            for (; sched.next(t); ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_full[acc], (it >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
                if (G.mode == OUT_TRANSPOSED_SILU) {
                    const bool is_up = quarter >= 2;
                    const int r64 = (quarter & 1) * 32 + lane;
                    const int n = t.a_tile * 64 + r64;
                    __nv_bfloat16* act = reinterpret_cast<__nv_bfloat16*>(G.out);
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32) {
                        float* buf = stage_silu + ((c >> 5) & 1) * (32 * 64);
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr0 + c, v);
                        tmem_ld_wait();
                        if (is_up) {
#pragma unroll
                            for (int jx = 0; jx < 32; ++jx) buf[jx * 64 + r64] = __uint_as_float(v[jx]);
                        }
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        if (!is_up && 2 * n < G.rowsA) {
#pragma unroll
                            for (int jx = 0; jx < 32; ++jx) {
                                const int b = c + jx;
                                if (b < a.rows)
                                    act[(size_t)b * G.ldo + n] = __float2bfloat16(silu_mul(__uint_as_float(v[jx]), buf[jx * 64 + r64]));
                            }
                        }
                    }
                } else {
                    const int a_row = t.a_tile * BLOCK_A + row_in_tile;
                    float* dst = reinterpret_cast<float*>(G.out);
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr0 + c, v);
                        tmem_ld_wait();
                        if (a_row < G.rowsA) {
#pragma unroll
                            for (int jx = 0; jx < 32; ++jx) {
                                const int b = c + jx;
                                if (b < a.rows) dst[((size_t)t.z * a.rows + b) * G.ldo + a_row] = __uint_as_float(v[jx]);
                            }
                        }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
        };
        auto publish = [&](int c) {                          // all 128 epilogue threads are done writing
            asm volatile("bar.sync 3, 128;" ::: "memory");
            if (etid == 0) grid_arrive(&a.counters[c]);
        };
        auto norm_phase = [&](const ChainNorm& N, int wait_c) {
            // x[row] += sum_z part[z][row]; xn[row] = rmsnorm(x[row]) * w  — one row per CTA (128 threads)
            if ((int)blockIdx.x < a.rows) {
                if (etid == 0) grid_wait(&a.counters[wait_c], n_cta);      // one poller per CTA, then CTA-level release
                asm volatile("bar.sync 3, 128;" ::: "memory");
            }
            for (int row = blockIdx.x; row < a.rows; row += n_cta) {
                float* xr = a.x + (size_t)row * a.hidden;
                float4 v[16];
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = (etid + i * 128) * 4;
                    if (c < a.hidden) {
                        v[i] = *reinterpret_cast<const float4*>(xr + c);
                        const float* pb = N.part + (size_t)row * a.hidden + c;
#pragma unroll 4
                        for (int z = 0; z < N.n_splits; ++z) {
                            const float4 p = *reinterpret_cast<const float4*>(pb + (size_t)z * N.split_stride);
                            v[i].x += p.x; v[i].y += p.y; v[i].z += p.z; v[i].w += p.w;
                        }
                        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
                    }
                }
                for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                if (lane == 0) red[warp - 2] = ss;
                asm volatile("bar.sync 3, 128;" ::: "memory");
                const float tot = red[0] + red[1] + red[2] + red[3];
                asm volatile("bar.sync 3, 128;" ::: "memory");
                const float inv = rsqrtf(tot / (float)a.hidden + a.eps);
                __nv_bfloat16* out = a.xn + (size_t)row * a.hidden;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = (etid + i * 128) * 4;
                    if (c < a.hidden) {
                        *reinterpret_cast<float4*>(xr + c) = v[i];
                        const uint2 wv = *reinterpret_cast<const uint2*>(N.w + c);
                        uint2 o;
                        o.x = pack_bf16(v[i].x * inv * bf16_lo(wv.x), v[i].y * inv * bf16_hi(wv.x));
                        o.y = pack_bf16(v[i].z * inv * bf16_lo(wv.y), v[i].w * inv * bf16_hi(wv.y));
                        *reinterpret_cast<uint2*>(out + c) = o;
                    }
                }
            }
        };

        auto mark = [&](int k) { if (etid == 0) trace_mark(k); };
        mark(20);
        gemm_epilogue(a.g[0]);  mark(21); publish(0);
        norm_phase(a.n[0], 0);  mark(22); publish(1);
        gemm_epilogue(a.g[1]);  mark(23); publish(2);
        gemm_epilogue(a.g[2]);  mark(24); publish(3);
        norm_phase(a.n[1], 3);  mark(25); publish(4);
        gemm_epilogue(a.g[3]);  mark(26);
    }

    tcgen05_fence_before();
    __syncthreads();
    trace_end(tr_slot);
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

template <int BN>
static int launch_chain_bn(const ChainArgs& a, cudaStream_t st) {
    constexpr int smem = GemmCfg<BN>::kSmemBytes + GemmCfg<BN>::kSiluStageBytes;
    static std::atomic<uint64_t> attr{0};
    if (ensure_dyn_smem(decode_chain_kernel<BN>, (int)smem, attr) != cudaSuccess) return RR_ERR_CUDA;
    cudaError_t e = launch_pdl(decode_chain_kernel<BN>, dim3(num_sms()), dim3(GEMM_THREADS), (size_t)smem, st, a);
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? RR_OK : RR_ERR_CUDA;
}

// Requires: all CTAs co-resident (grid = SM count, 1 CTA/SM by shared memory), counters zeroed before the launch.
int launch_decode_chain(const ChainArgs& a, int bn, cudaStream_t st) {
    switch (bn) {
        case 32: return launch_chain_bn<32>(a, st);
        case 64: return launch_chain_bn<64>(a, st);
        case 128: return launch_chain_bn<128>(a, st);
        case 256: return launch_chain_bn<256>(a, st);
    }
    return RR_ERR_ARG;
}

int chain_gemm_init(ChainGemm* g, const void* W, int rowsA, int K, const void* act, int rows, void* out, int ldo,
                    int splits, int mode, int bn) {
    if (K % 8 || rows > bn) return RR_ERR_ARG;
    const int kblocks = (K + BLOCK_K - 1) / BLOCK_K;
    if (splits < 1) splits = 1;
    if (splits > kblocks) splits = kblocks;
    if (mode == OUT_TRANSPOSED_SILU && (splits != 1 || rowsA % 128)) return RR_ERR_ARG;
    g->out = out; g->rowsA = rowsA; g->K = K; g->splits = splits; g->ldo = ldo; g->mode = mode;
    int rc = make_tmap_bf16_2d(&g->tmA, W, rowsA, K, K, BLOCK_A);
    if (rc != RR_OK) return rc;
    return make_tmap_bf16_2d(&g->tmB, act, rows, K, K, bn);
}

void rr_trace_set_chain(unsigned long long* p) { rr_trace_set_local(p); }

}  // namespace rr
