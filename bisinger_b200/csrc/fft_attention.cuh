// Self-attention of the FastSpeech FFT blocks (modules/commons/common_layers.py:199-420 MultiheadAttention with self_attention=True,
// bias=False -> F.multi_head_attention_forward; used by EncSALayer :664-731 inside FFTBlocks, modules/fastspeech/tts_modules.py:253-310)
// as ONE tcgen05 kernel per layer: O = softmax(Q K^T / sqrt(d) + key_padding_mask) V for every (utterance, head, 128-query tile).
//
//   Q, K   fp16 [B][T][3C] (the in-projection's output as it stands: q at columns h*128, k at C + h*128), read by TMA
//   Vt     fp16 [B*H][128][Tp]  V transposed (keys contiguous), so that P V is a K-major x K-major tcgen05 MMA like every other here
//   O      bf16 hi/lo [B*T][C] at columns h*128: the A operand of the out-projection GEMM (bf16x3)
//
// Two passes over the key tiles, both on the tensor cores: pass 1 computes S = Q K^T tile by tile and keeps the running row maximum and
// row sum in registers (one query row per thread, as tcgen05.ld hands the accumulator over); pass 2 recomputes S, writes
// P = exp(S - max) / sum as fp16 straight into shared memory in the SWIZZLE_128B K-major layout (the same trick as the fused ResBlock
// kernel's intermediate tile) and accumulates O += P Vt in TMEM -- no rescaling of the O accumulator, at the price of computing Q K^T twice
// (cheap: the kernel is not on the 100-step path).  S is double-buffered in TMEM, the K / Vt tiles in shared memory; the MMA warp issues
// S(i+1) before P V(i), so the softmax of tile i overlaps the next Q K^T.
// Warps: 0 = TMA producer + TMEM allocator, 1 = MMA issuer (whole warp, elected lane), 2..5 = softmax / epilogue (TMEM lane quadrant = warp % 4).
#pragma once
#include "resblock_fused.cuh"

namespace b200 {

struct AttnArgs {
    CUtensorMap qkv;          // fp16 [B][T][3C], box = 64 columns x 128 rows
    CUtensorMap vt;           // fp16 [B*H][128][Tp], box = 64 keys x 128 rows (d)
    const uint32_t* keybits;  // [B][ceil(T / 32)]: bit = 1 for a key that may be attended to (not padding, < T)
    __nv_bfloat16* o_hi;      // [B*T][C]
    __nv_bfloat16* o_lo;
    int B, T, H, C;           // C = H * 128
    int q_tiles;              // ceil(T / 128)
    float scale_log2e;        // head_dim^-0.5 * log2(e)
};

constexpr int kAttnThreads = 32 * 6;
struct AttnSmem {
    static constexpr int kTile = 128 * 128 * 2;     // 128 rows x 128 fp16 = two 16 KB slabs
    static constexpr int kOffQ = 0;
    static constexpr int kOffK = kTile;             // 2 stages
    static constexpr int kOffV = 3 * kTile;         // 2 stages
    static constexpr int kOffP = 5 * kTile;
    static constexpr int kOffBar = 6 * kTile;
    static constexpr int kTotal = kOffBar + 256 + 1024;
    static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

__global__ void __launch_bounds__(kAttnThreads, 1) fft_attn_kernel(const __grid_constant__ AttnArgs args) {
    using S = AttnSmem;
    constexpr uint32_t kIdesc = umma_idesc_f16(128, 128, /*fp16=*/true);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOffBar);
    uint64_t* qfull = bars;            // 1
    uint64_t* kfull = bars + 1;        // 2
    uint64_t* kempty = bars + 3;       // 2
    uint64_t* vfull = bars + 5;        // 2
    uint64_t* vempty = bars + 7;       // 2
    uint64_t* sfull = bars + 9;        // 2
    uint64_t* sempty = bars + 11;      // 2
    uint64_t* pfull = bars + 13;       // 1
    uint64_t* pempty = bars + 14;      // 1
    uint64_t* ofull = bars + 15;       // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int qt = static_cast<int>(blockIdx.x) % args.q_tiles;
    const int bh = static_cast<int>(blockIdx.x) / args.q_tiles;
    const int b = bh / args.H, h = bh % args.H;
    const int nk = (args.T + 127) / 128;
    const int n_it = 2 * nk;

    if (threadIdx.x == 32) {
        mbar_init(qfull, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kfull[s], 1); mbar_init(&kempty[s], 1); mbar_init(&vfull[s], 1); mbar_init(&vempty[s], 1);
            mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 4);
        }
        mbar_init(pfull, 4); mbar_init(pempty, 1); mbar_init(ofull, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        mbar_arrive_expect_tx(qfull, S::kTile);
        for (int s = 0; s < 2; ++s) tma_load_3d(smem + S::kOffQ + s * (S::kTile / 2), &args.qkv, qfull, h * 128 + s * 64, qt * 128, b);
        for (int i = 0; i < n_it; ++i) {
            const int j = i % nk, st = i & 1;
            mbar_wait(&kempty[st], ((i >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&kfull[st], S::kTile);
            for (int s = 0; s < 2; ++s)
                tma_load_3d(smem + S::kOffK + st * S::kTile + s * (S::kTile / 2), &args.qkv, &kfull[st], args.C + h * 128 + s * 64, j * 128, b);
            if (i >= nk) {
                const int iv = i - nk, vs = iv & 1;
                mbar_wait(&vempty[vs], ((iv >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&vfull[vs], S::kTile);
                for (int s = 0; s < 2; ++s)
                    tma_load_3d(smem + S::kOffV + vs * S::kTile + s * (S::kTile / 2), &args.vt, &vfull[vs], j * 128 + s * 64, 0, bh);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint64_t q_desc = umma_smem_desc<128>(sbase + S::kOffQ);
        auto issue_s = [&](int i) {
            const int st = i & 1;
            mbar_wait_warp(smem_u32(&kfull[st]), (i >> 1) & 1);
            mbar_wait_warp(smem_u32(&sempty[st]), ((i >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint64_t k_desc = umma_smem_desc<128>(sbase + S::kOffK + st * S::kTile);
            const uint32_t tacc = tmem_base + st * 128;
            mma_f16_x4(tacc, q_desc, k_desc, kIdesc, 0u);                                        // d 0..63
            mma_f16_x4(tacc, q_desc + (S::kTile / 2 >> 4), k_desc + (S::kTile / 2 >> 4), kIdesc, 1u);   // d 64..127
            umma_commit_elect_s(smem_u32(&kempty[st]));
            umma_commit_elect_s(smem_u32(&sfull[st]));
        };
        auto issue_pv = [&](int iv) {
            const int vs = iv & 1;
            mbar_wait_warp(smem_u32(pfull), iv & 1);
            mbar_wait_warp(smem_u32(&vfull[vs]), (iv >> 1) & 1);
            tc_fence_after();
            const uint64_t p_desc = umma_smem_desc<128>(sbase + S::kOffP);
            const uint64_t v_desc = umma_smem_desc<128>(sbase + S::kOffV + vs * S::kTile);
            const uint32_t tacc = tmem_base + 256;
            mma_f16_x4(tacc, p_desc, v_desc, kIdesc, iv > 0 ? 1u : 0u);                          // keys 0..63 of the tile
            mma_f16_x4(tacc, p_desc + (S::kTile / 2 >> 4), v_desc + (S::kTile / 2 >> 4), kIdesc, 1u);
            umma_commit_elect_s(smem_u32(&vempty[vs]));
            umma_commit_elect_s(smem_u32(pempty));
        };
        mbar_wait_warp(smem_u32(qfull), 0);
        issue_s(0);
        for (int i = 0; i < n_it; ++i) {
            if (i + 1 < n_it) issue_s(i + 1);
            if (i >= nk) issue_pv(i - nk);
        }
        umma_commit_elect_s(smem_u32(ofull));
    } else if (warp >= 2) {
        // ================= softmax / epilogue: one query row per thread =================
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const int t = qt * 128 + r;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t* kb = args.keybits + static_cast<size_t>(b) * ((args.T + 31) / 32);
        float m = -3.0e38f, l = 0.0f, inv_l = 0.0f;
        const uint32_t prow = sbase + S::kOffP + r * 128;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        for (int i = 0; i < n_it; ++i) {
            const int j = i % nk, st = i & 1;
            const bool pass2 = i >= nk;
            if (i == nk) inv_l = 1.0f / l;
            mbar_wait(&sfull[st], (i >> 1) & 1);
            tc_fence_after();
            if (pass2) mbar_wait(pempty, ((i - nk) & 1) ^ 1);   // the previous P V has read the tile
            const uint32_t tacc = tmem_base + lane_addr + st * 128;
#pragma unroll 1
            for (int c = 0; c < 128; c += 32) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld32(tacc + c, v);
                tmem_ld_wait32(v);
                const int w_idx = j * 4 + (c >> 5);
                const uint32_t bits = (j * 128 + c < args.T) ? __ldg(kb + w_idx) : 0u;
                float s[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) s[e] = ((bits >> e) & 1u) ? __uint_as_float(v[e]) * args.scale_log2e : -3.0e38f;
                if (!pass2) {
                    float mx = m;
#pragma unroll
                    for (int e = 0; e < 32; ++e) mx = fmaxf(mx, s[e]);
                    float acc = 0.0f;
#pragma unroll
                    for (int e = 0; e < 32; ++e) acc += exp2f(s[e] - mx);   // masked keys: exp2(-3e38 - mx) = 0
                    l = l * exp2f(m - mx) + acc;
                    m = mx;
                } else {
                    uint32_t p[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) p[e] = pack_h2(exp2f(s[2 * e] - m) * inv_l, exp2f(s[2 * e + 1] - m) * inv_l);
                    const uint32_t slab = prow + (c >> 6) * (S::kTile / 2);
                    const uint32_t j0 = static_cast<uint32_t>((c & 63) >> 3);
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts128u(slab + (((j0 + q) ^ sw) << 4), p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
                }
            }
            tc_fence_before();
            if (pass2) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sempty[st]);
                if (pass2) mbar_arrive(pfull);
            }
        }
        // ---- O -> bf16 hi/lo rows
        mbar_wait(ofull, 0);
        tc_fence_after();
        const bool ok = t < args.T;
        const long long row = static_cast<long long>(b) * args.T + t;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(tmem_base + lane_addr + 256 + c, v);
            tmem_ld_wait32(v);
            if (ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = __uint_as_float(v[8 * g + 2 * e]), a1 = __uint_as_float(v[8 * g + 2 * e + 1]);
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
                        hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
                        const __nv_bfloat162 ll = __floats2bfloat162_rn(a0 - __uint_as_float(hi[e] << 16), a1 - __uint_as_float(hi[e] & 0xffff0000u));
                        lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    const long long off = row * args.C + h * 128 + c + 8 * g;
                    *reinterpret_cast<uint4*>(args.o_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(args.o_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200
