// Persistent warp-specialised implicit-GEMM convolution for sm_100a (tcgen05 + TMEM + TMA).
//
// One kernel template serves every 1-D convolution on the hot path.  Activations are stored
// channels-last ([B][L][C] bf16), so a convolution tap is a *row shift* of the A operand:
//     Y[b, t, n] = sum_seg sum_c  A_seg[b, t + shift_seg, c] * W[n, wcol_seg + c]
// A tiles (128 rows x 64 channels) are fetched by 3-D TMA with the shifted row coordinate; rows that
// fall outside [0, L) are zero-filled by the TMA unit, which is exactly the reference's zero padding
// (nn.Conv1d padding=dilation, usr/diff/net.py:61; get_padding, modules/hifigan/hifigan.py:26-27).
// Weights are packed K-major ([N][Ktot] bf16) and fetched by 2-D TMA.  tcgen05.mma (M=128, N=N_TILE,
// K=16) accumulates in TMEM (fp32); two accumulator buffers let the epilogue of tile i overlap the
// MMAs of tile i+1.
//
// TERMS == 3 is the "bf16x3" contraction: every operand is carried as hi = bf16(x), lo = bf16(x - hi)
// and each k-block issues A_hi*W_hi + A_lo*W_hi + A_hi*W_lo into the same fp32 accumulator (~16
// mantissa bits).  The 100-step sampler needs it to stay within the 1e-2 mel tolerance.
//
// Warp roles (192 threads): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer (one thread),
// warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>

#include "ptx.cuh"

namespace b200 {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kMaxSeg = 12;
constexpr int kGemmThreads = 192;
constexpr int kSmemBudget = 200 * 1024;

struct Segment {
    int a_src;      // index of the A tensor-map pair (hi = 2*a_src, lo = 2*a_src + 1)
    int row_shift;  // tap offset in rows (may be negative)
    int a_col0;     // first channel of A used by this segment
    int n_kb;       // number of 64-channel k-blocks
    int w_col0;     // first K column of the packed weight matrix
};

// Epilogue selector
enum : int {
    EPI_F32 = 0,        // out_f32[row][n] = acc + bias[n]                       (unit tests)
    EPI_INPROJ = 1,     // relu(acc+bias) -> xres f32 ; (+dvec) -> xa hi/lo      (net.py:116-118)
    EPI_GATE = 2,       // sigmoid(g+bg) * tanh(f+bf) -> z hi/lo                 (net.py:71-74)
    EPI_RES_SKIP = 3,   // n_tile 0: x=(x+r+b)/sqrt2 ; n_tile 1: skip += s+b     (net.py:76-78,126)
    EPI_RELU_BF16 = 4,  // relu(acc+bias) -> hi/lo                               (net.py:127-128)
    EPI_POSTERIOR = 5,  // eps=acc+bias -> DDPM posterior update of x_t          (shallow_diffusion_tts.py:149-166)
    EPI_BIAS_ACT = 6,   // HiFi-GAN: y = acc+bias (+res) ; writes f32 and/or lrelu(y) bf16
};

struct EpiParams {
    const float* bias;         // [N_total]
    float* f32_a;              // EPI_F32: out ; INPROJ/RES: residual stream x [rows][256] ; POSTERIOR: x_t [B][M][T]
    float* f32_b;              // RES_SKIP: skip accumulator [rows][256] ; POSTERIOR: mel_out [B][T][M] (last step) or null
    __nv_bfloat16* out_hi;     // bf16 operand written for the next GEMM
    __nv_bfloat16* out_lo;     // (TERMS==3 consumers) low part, may be null
    __nv_bfloat16* out2_hi;    // RES_SKIP: head input (last layer) ; BIAS_ACT: second bf16 output
    __nv_bfloat16* out2_lo;
    const float* dvec;         // step-embedding vector added before the bf16 split (next layer's d), or null
    const float* aux0;         // POSTERIOR: injected noise [B][M][T] or null ; BIAS_ACT: residual f32 in
    const float* aux1;         // POSTERIOR: spec_min[M]
    const float* aux2;         // POSTERIOR: spec_max[M]
    const int64_t* mel2ph;     // POSTERIOR: [B][T] or null
    int out_pitch;             // elements per row of out_hi/out_lo/f32 outputs
    int out_col0;              // column offset added to n (transposed-conv phases)
    int act_pitch;             // BIAS_ACT: elements per row of the bf16 activation output (channel-padded buffers)
    int flags;                 // RES_SKIP: bit0 = first layer (skip = ...), bit1 = last layer ; BIAS_ACT: see hifigan
    float c0, c1, c2, c3, c4;  // POSTERIOR: sqrt_recip, sqrt_recipm1, coef1, coef2, sigma (0 at t==0) ; BIAS_ACT: scale, slope
    const unsigned long long* seed_ptr;  // POSTERIOR: device-resident Philox seed (read when aux0 == null)
    unsigned int step;         // POSTERIOR: Philox stream offset (executed step index)
};

struct ConvGemmArgs {
    CUtensorMap amap[6];   // A sources: (hi, lo) pairs
    CUtensorMap wmap[2];   // packed weights hi / lo
    int B, L;              // batches, rows per batch
    int tiles_per_batch;   // ceil(L / 128)
    int n_tiles_n;         // N tiles per row tile
    int num_tiles;
    int w_row0;            // first weight row of this launch
    int n_seg;
    Segment seg[kMaxSeg];
    EpiParams epi;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG + Box-Muller (production noise path; parity tests inject noise)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u = (static_cast<float>(a) + 0.5f) * 2.3283064365386963e-10f;   // (0,1)
    const float v = (static_cast<float>(b) + 0.5f) * 2.3283064365386963e-10f;
    const float r = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(6.283185307179586f * v, &s, &c);
    return make_float2(r * c, r * s);
}
// standard normal for element `idx` of executed step `step`
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t step, uint64_t idx) {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(idx >> 2), static_cast<uint32_t>(idx >> 34), step, 0x5eedu),
                                  make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
    const uint32_t w = static_cast<uint32_t>(idx & 3);
    const float2 n = (w < 2) ? box_muller(r.x, r.y) : box_muller(r.z, r.w);
    return (w & 1) ? n.y : n.x;
}

// ---------------------------------------------------------------------------------------------
// Epilogue helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_tanh(float x) {
    // tanh(x) = 1 - 2/(1+e^{2x}); ex2/rcp approximations (rel err ~1e-6), saturates correctly.
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, 1.0f + e);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// store 32 floats as bf16 hi (and lo) rows: 64 B each, 16-byte vector stores
__device__ __forceinline__ void store_split32(const float (&v)[32], __nv_bfloat16* hi_row, __nv_bfloat16* lo_row) {
    uint32_t ph[16], pl[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float h0, h1;
        __nv_bfloat16 bh0, bl0, bh1, bl1;
        split_bf16(v[2 * i], h0, bh0, bl0);
        split_bf16(v[2 * i + 1], h1, bh1, bl1);
        ph[i] = static_cast<uint32_t>(__bfloat16_as_ushort(bh0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bh1)) << 16);
        pl[i] = static_cast<uint32_t>(__bfloat16_as_ushort(bl0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bl1)) << 16);
    }
    uint4* dh = reinterpret_cast<uint4*>(hi_row);
#pragma unroll
    for (int i = 0; i < 4; ++i) dh[i] = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
    if (lo_row != nullptr) {
        uint4* dl = reinterpret_cast<uint4*>(lo_row);
#pragma unroll
        for (int i = 0; i < 4; ++i) dl[i] = make_uint4(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3]);
    }
}
__device__ __forceinline__ void load_f32x32(const float* p, float (&v)[32]) {
    const float4* s = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 q = s[i];
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
}
__device__ __forceinline__ void store_f32x32(float* p, const float (&v)[32]) {
    float4* d = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void ld_acc32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after any divergent store path
    tmem_ld32(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld_acc16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    __syncwarp();
    tmem_ld16(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// BIAS_ACT flags (HiFi-GAN epilogue)
enum : int {
    BA_ADD_RES = 1,      // y += aux0[row][n]            (ResBlock1 residual, hifigan.py:60)
    BA_WRITE_F32 = 2,    // f32_a[row][n] = y
    BA_ACCUM_F32B = 4,   // f32_b[row][n] (+)= y * c0    (MRF sum / num_kernels, hifigan.py:161-168)
    BA_ACCUM_INIT = 8,   // with BA_ACCUM_F32B: '=' instead of '+='
    BA_WRITE_ACT = 16,   // out_hi[row][n] = bf16(lrelu(y, c1))
    BA_ACT_FROM_B = 32,  // the bf16 activation is taken from the accumulated f32_b value instead of y
};

template <int N_TILE, int EPI>
__device__ __forceinline__ void run_epilogue(const ConvGemmArgs& args, uint32_t tacc, int b, int t, int n_tile) {
    const EpiParams& e = args.epi;
    const bool row_ok = t < args.L;
    const long long row = static_cast<long long>(b) * args.L + t;

    if constexpr (EPI == EPI_F32) {
#pragma unroll 1
        for (int c = 0; c < N_TILE; c += 32) {
            float v[32];
            ld_acc32(tacc + c, v);
            if (row_ok) {
                const int n = n_tile * N_TILE + c;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(e.bias + n + i);
                store_f32x32(e.f32_a + row * e.out_pitch + n, v);
            }
        }
    } else if constexpr (EPI == EPI_INPROJ) {
#pragma unroll 1
        for (int c = 0; c < N_TILE; c += 32) {
            float v[32];
            ld_acc32(tacc + c, v);
            if (row_ok) {
                const int n = n_tile * N_TILE + c;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + __ldg(e.bias + n + i), 0.0f);
                store_f32x32(e.f32_a + row * e.out_pitch + n, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += __ldg(e.dvec + n + i);
                store_split32(v, e.out_hi + row * e.out_pitch + n, e.out_lo ? e.out_lo + row * e.out_pitch + n : nullptr);
            }
        }
    } else if constexpr (EPI == EPI_GATE) {
        // tile columns [0, N_TILE/2) = gate pre-activations, [N_TILE/2, N_TILE) = filter pre-activations of
        // the same N_TILE/2 channels (weight rows permuted by the packer).
        constexpr int HALF = N_TILE / 2;
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
            float g[32], f[32];
            ld_acc32(tacc + c, g);
            ld_acc32(tacc + HALF + c, f);
            if (row_ok) {
                const int nb = n_tile * N_TILE + c;   // bias index of gate col c ; filter bias at +HALF
                const int ch = n_tile * HALF + c;     // output channel
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float gg = g[i] + __ldg(e.bias + nb + i);
                    const float ff = f[i] + __ldg(e.bias + nb + HALF + i);
                    g[i] = fast_sigmoid(gg) * fast_tanh(ff);
                }
                store_split32(g, e.out_hi + row * e.out_pitch + ch, e.out_lo ? e.out_lo + row * e.out_pitch + ch : nullptr);
            }
        }
    } else if constexpr (EPI == EPI_RES_SKIP) {
        static_assert(N_TILE == 256, "residual/skip split assumes 256 channels per tile");
#pragma unroll 1
        for (int c = 0; c < N_TILE; c += 32) {
            float v[32];
            ld_acc32(tacc + c, v);
            if (!row_ok) continue;
            const int n = n_tile * N_TILE + c;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(e.bias + n + i);
            if (n_tile == 0) {
                float* xr = e.f32_a + row * e.out_pitch + c;
                float x[32];
                load_f32x32(xr, x);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = (x[i] + v[i]) * 0.70710678118654752440f;
                store_f32x32(xr, v);
                if (e.dvec != nullptr) {   // not needed after the last layer
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __ldg(e.dvec + c + i);
                    store_split32(v, e.out_hi + row * e.out_pitch + c, e.out_lo ? e.out_lo + row * e.out_pitch + c : nullptr);
                }
            } else {
                float* sk = e.f32_b + row * e.out_pitch + c;
                if (!(e.flags & 1)) {
                    float s[32];
                    load_f32x32(sk, s);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += s[i];
                }
                if (e.flags & 2) {   // last layer: hand sum/sqrt(L) to the head GEMM as bf16 operand
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= e.c0;
                    store_split32(v, e.out2_hi + row * e.out_pitch + c, e.out2_lo ? e.out2_lo + row * e.out_pitch + c : nullptr);
                } else {
                    store_f32x32(sk, v);
                }
            }
        }
    } else if constexpr (EPI == EPI_RELU_BF16) {
#pragma unroll 1
        for (int c = 0; c < N_TILE; c += 32) {
            float v[32];
            ld_acc32(tacc + c, v);
            if (row_ok) {
                const int n = n_tile * N_TILE + c;
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + __ldg(e.bias + n + i), 0.0f);
                store_split32(v, e.out_hi + row * e.out_pitch + n, e.out_lo ? e.out_lo + row * e.out_pitch + n : nullptr);
            }
        }
    } else if constexpr (EPI == EPI_POSTERIOR) {
        // N_TILE == mel bins (80).  x_t lives in the reference layout [B][M][T] (T contiguous): lanes hold
        // consecutive t, so the per-channel accesses below are coalesced.
        static_assert(N_TILE % 16 == 0 && N_TILE <= 96, "posterior epilogue expects <= 96 mel bins");
        const int M = N_TILE;
        float xnew[N_TILE];
        {
            float v[32];
#pragma unroll
            for (int c = 0; c + 32 <= N_TILE; c += 32) {
                ld_acc32(tacc + c, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) xnew[c + i] = v[i];
            }
            if constexpr (N_TILE % 32 != 0) {
                float w[16];
                ld_acc16(tacc + (N_TILE / 32) * 32, w);
#pragma unroll
                for (int i = 0; i < 16; ++i) xnew[(N_TILE / 32) * 32 + i] = w[i];
            }
        }
        if (row_ok && (e.flags & 1)) {
            // eps-only mode (bsg_diffnet_forward): write the denoiser output in the reference layout [B][M][T]
            float* eo = e.f32_a + (static_cast<long long>(b) * M) * args.L + t;
#pragma unroll
            for (int c = 0; c < N_TILE; ++c) eo[static_cast<long long>(c) * args.L] = xnew[c] + __ldg(e.bias + c);
        } else if (row_ok) {
            float* xt = e.f32_a + (static_cast<long long>(b) * M) * args.L + t;
            const float* nz = e.aux0 ? e.aux0 + (static_cast<long long>(b) * M) * args.L + t : nullptr;
#pragma unroll
            for (int c = 0; c < N_TILE; ++c) {
                const float eps = xnew[c] + __ldg(e.bias + c);
                const float x = xt[static_cast<long long>(c) * args.L];
                float x0 = e.c0 * x - e.c1 * eps;                       // predict_start_from_noise (:134-138)
                x0 = fminf(fmaxf(x0, -1.0f), 1.0f);                      // clamp_ (:153-154)
                float mean = e.c2 * x0 + e.c3 * x;                       // q_posterior (:140-147)
                float z;
                if (nz != nullptr) z = nz[static_cast<long long>(c) * args.L];
                else z = philox_normal(__ldg(e.seed_ptr), e.step, (static_cast<uint64_t>(b) * M + c) * args.L + t);
                const float xn = mean + e.c4 * z;                        // (:166), c4 = 0 at t == 0
                xt[static_cast<long long>(c) * args.L] = xn;
                xnew[c] = xn;
            }
            // bf16 operand copy [B][T][M] for the next step's input projection
            if (e.out_hi != nullptr) {
                __nv_bfloat16* hrow = e.out_hi + row * e.out_pitch;
                __nv_bfloat16* lrow = e.out_lo ? e.out_lo + row * e.out_pitch : nullptr;
#pragma unroll
                for (int c = 0; c < N_TILE; c += 8) {
                    uint32_t ph[4], pl[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float h0, h1;
                        __nv_bfloat16 bh0, bl0, bh1, bl1;
                        split_bf16(xnew[c + 2 * i], h0, bh0, bl0);
                        split_bf16(xnew[c + 2 * i + 1], h1, bh1, bl1);
                        ph[i] = static_cast<uint32_t>(__bfloat16_as_ushort(bh0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bh1)) << 16);
                        pl[i] = static_cast<uint32_t>(__bfloat16_as_ushort(bl0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bl1)) << 16);
                    }
                    *reinterpret_cast<uint4*>(hrow + c) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    if (lrow) *reinterpret_cast<uint4*>(lrow + c) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                }
            }
            // last step: mel_out = denorm_spec(x) * (mel2ph > 0)   (shallow_diffusion_tts.py:268-272,278-279)
            if (e.f32_b != nullptr) {
                const float mask = (e.mel2ph == nullptr || e.mel2ph[row] > 0) ? 1.0f : 0.0f;
                float* mo = e.f32_b + row * M;
#pragma unroll
                for (int c = 0; c < N_TILE; c += 4) {
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float mn = __ldg(e.aux1 + c + i), mx = __ldg(e.aux2 + c + i);
                        o[i] = ((xnew[c + i] + 1.0f) / 2.0f * (mx - mn) + mn) * mask;
                    }
                    *reinterpret_cast<float4*>(mo + c) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    } else if constexpr (EPI == EPI_BIAS_ACT) {
#pragma unroll 1
        for (int c = 0; c < N_TILE; c += 32) {
            float v[32];
            ld_acc32(tacc + c, v);
            if (!row_ok) continue;
            const int n = n_tile * N_TILE + c;            // bias / logical channel index
            const long long o = row * e.out_pitch + e.out_col0 + n;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += __ldg(e.bias + n + i);
            if (e.flags & BA_ADD_RES) {
                float r[32];
                load_f32x32(e.aux0 + o, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += r[i];
            }
            if (e.flags & BA_WRITE_F32) store_f32x32(e.f32_a + o, v);
            if (e.flags & BA_ACCUM_F32B) {
                float s[32];
                if (e.flags & BA_ACCUM_INIT) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) s[i] = v[i] * e.c0;
                } else {
                    load_f32x32(e.f32_b + o, s);
#pragma unroll
                    for (int i = 0; i < 32; ++i) s[i] += v[i] * e.c0;
                }
                store_f32x32(e.f32_b + o, s);
                if (e.flags & BA_ACT_FROM_B) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = s[i];
                }
            }
            if (e.flags & BA_WRITE_ACT) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.0f ? v[i] : v[i] * e.c1;
                const long long oa = row * e.act_pitch + n;
                store_split32(v, e.out_hi + oa, e.out_lo ? e.out_lo + oa : nullptr);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
template <int N_TILE, int TERMS>
struct GemmSmem {
    static constexpr int kABytes = kTileM * kBlockK * 2;
    static constexpr int kBBytes = N_TILE * kBlockK * 2;
    static constexpr int kStageBytes = (TERMS == 3 ? 2 : 1) * (kABytes + kBBytes);
    static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kBarBytes = 256;
    static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;  // +1024 for manual alignment
    static_assert(kStages >= 2, "need at least a double-buffered pipeline");
    static_assert(kBBytes % 1024 == 0, "B tile must keep 1024-byte swizzle alignment");
};

__host__ __device__ constexpr int tmem_cols_for(int n) {
    return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : (n <= 256 ? 256 : 512)));
}

template <int N_TILE, int TERMS, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmArgs args) {
    using S = GemmSmem<N_TILE, TERMS>;
    static_assert(N_TILE % 16 == 0 && N_TILE >= 16 && N_TILE <= 256, "UMMA N constraint for M=128");
    constexpr int kTmemCols = tmem_cols_for(2 * N_TILE);
    constexpr uint32_t kIdesc = umma_idesc_bf16(kTileM, N_TILE);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + S::kStages;
    uint64_t* tfull_bar = bars + 2 * S::kStages;
    uint64_t* tempty_bar = bars + 2 * S::kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 32) {
        for (int s = 0; s < S::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_seg = args.n_seg;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
            const int n_tile = tile % args.n_tiles_n;
            const int m = tile / args.n_tiles_n;
            const int b = m / args.tiles_per_batch;
            const int t0 = (m % args.tiles_per_batch) * kTileM;
            const int wrow = args.w_row0 + n_tile * N_TILE;
            for (int s = 0; s < n_seg; ++s) {
                const Segment sg = args.seg[s];
                for (int kb = 0; kb < sg.n_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* st = smem + stage * S::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
                    tma_load_3d(st, &args.amap[2 * sg.a_src], &full_bar[stage], sg.a_col0 + kb * kBlockK, t0 + sg.row_shift, b);
                    if (TERMS == 3)
                        tma_load_3d(st + S::kABytes, &args.amap[2 * sg.a_src + 1], &full_bar[stage], sg.a_col0 + kb * kBlockK,
                                    t0 + sg.row_shift, b);
                    uint8_t* sb = st + (TERMS == 3 ? 2 : 1) * S::kABytes;
                    tma_load_2d(sb, &args.wmap[0], &full_bar[stage], sg.w_col0 + kb * kBlockK, wrow);
                    if (TERMS == 3) tma_load_2d(sb + S::kBBytes, &args.wmap[1], &full_bar[stage], sg.w_col0 + kb * kBlockK, wrow);
                    if (++stage == S::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer (single thread) =================
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + acc * N_TILE;
            uint32_t accumulate = 0;
            for (int s = 0; s < n_seg; ++s) {
                const int n_kb = args.seg[s].n_kb;
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + stage * S::kStageBytes);
                    const uint32_t a_lo = a_hi + S::kABytes;
                    const uint32_t b_hi = a_hi + (TERMS == 3 ? 2 : 1) * S::kABytes;
                    const uint32_t b_lo = b_hi + S::kBBytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint64_t da = umma_smem_desc<128>(a_hi + k * 32);
                        const uint64_t db = umma_smem_desc<128>(b_hi + k * 32);
                        umma_f16(tacc, da, db, kIdesc, accumulate);
                        accumulate = 1;
                        if (TERMS == 3) {
                            umma_f16(tacc, umma_smem_desc<128>(a_lo + k * 32), db, kIdesc, 1);
                            umma_f16(tacc, da, umma_smem_desc<128>(b_lo + k * 32), kIdesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem slot when these MMAs retire
                    if (++stage == S::kStages) { stage = 0; phase ^= 1; }
                }
            }
            umma_commit(&tfull_bar[acc]);             // accumulator complete -> epilogue
        }
    } else if (warp >= 2) {
        // ================= Epilogue warps =================
        const int quad = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int n_tile = tile % args.n_tiles_n;
            const int m = tile / args.n_tiles_n;
            const int b = m / args.tiles_per_batch;
            const int t = (m % args.tiles_per_batch) * kTileM + quad * 32 + lane;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * N_TILE;
            run_epilogue<N_TILE, EPI>(args, tacc, b, t, n_tile);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace b200
