// FastSpeech FFT blocks (the mel-rate decoder of FastSpeech2 / FastSpeech2MIDI) + mel_out: the conditioner's handoff to the sampler,
// SURVEY.md section 8f-3.
//
// Reference (relative to /root/reference/train_bisinger/):
//   modules/fastspeech/tts_modules.py:253-310 (FFTBlocks.forward), :340-347 (FastspeechDecoder), :18-33 (TransformerEncoderLayer)
//   modules/commons/common_layers.py:664-731 (EncSALayer), :598-644 (TransformerFFNLayer), :199-420 (MultiheadAttention, bias=False),
//                                    :106-158 (SinusoidalPositionalEmbedding), utils/__init__.py:146-158 (make_positions)
//   modules/fastspeech/fs2.py:236-240 (run_decoder: decoder -> mel_out -> * tgt_nonpadding)
//
// Per layer (pre-LN transformer block, eval mode):
//     x += out_proj( MHA( LN1(x) ) ) ; x *= nonpad ;  x += ffn_2( act( ffn_1_conv_k9( LN2(x) ) * k^-0.5 ) ) ; x *= nonpad
// Every projection / convolution is one launch of conv_gemm_kernel<256, 3, EPI_BIAS_ACT> (bf16x3: ~fp32 accuracy; the whole decoder is about one
// diffusion step of FLOPs and runs once per batch), the attention is fft_attn_kernel (fft_attention.cuh: tcgen05, fp16 operands), what sits
// between them (padding mask, positional embedding, LayerNorm, operand split) runs in warp-per-row kernels.  Layout: channels-last,
// rows = b*T + t.
//   X    f32  [rows][C]      residual stream
//   A    bf16 [rows][C]      hi / lo operand of the next GEMM (LayerNorm output, attention output)
//   QKV  fp16 [rows][3C]     in-projection output; Vt fp16 [B*H][128][Tp] its V third transposed for the P V MMA
//   F    bf16 [rows][4C]     hi / lo of the FFN's hidden activation
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>

#include "plans.h"
#include "fft_attention.cuh"

namespace b200 {

namespace {
constexpr int kC = 256;

__device__ __forceinline__ void f_split8(const float (&v)[8], __nv_bfloat16* hi, __nv_bfloat16* lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        const float hx = __uint_as_float(h[i] << 16), hy = __uint_as_float(h[i] & 0xffff0000u);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hx, v[2 * i + 1] - hy);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void f_load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void f_store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float f_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// nonpad[row] = (sum |x[row]| != 0)   (tts_modules.py:291: padding_mask = x.abs().sum(-1).eq(0))
__global__ void fft_nonpad_kernel(const float* __restrict__ x, long long rows, float* __restrict__ nonpad) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long r = warp; r < rows; r += nwarps) {
        float v[8];
        f_load8(x + r * kC + lane * 8, v);
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += fabsf(v[i]);
        s = f_warp_sum(s);
        if (lane == 0) nonpad[r] = s == 0.0f ? 0.0f : 1.0f;
    }
}
// nonpad[row] = !padding_mask[row]: FFTBlocks.forward's explicit mask (tts_modules.py:291-292; the encoder passes txt_tokens.eq(0), :333)
__global__ void fft_nonpad_from_mask_kernel(const uint8_t* __restrict__ mask, long long rows, float* __restrict__ nonpad) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < rows) nonpad[i] = mask[i] ? 0.0f : 1.0f;
}
// keybits[b][w] bit e = key 32 w + e of utterance b may be attended to (key_padding_mask == 0 and inside [0, T))
__global__ void fft_keybits_kernel(const float* __restrict__ nonpad, int B, int T, uint32_t* __restrict__ bits) {
    const int words = (T + 31) / 32;
    const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long w = g >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= static_cast<long long>(B) * words) return;
    const int b = static_cast<int>(w / words), t = static_cast<int>(w % words) * 32 + lane;
    const bool on = t < T && nonpad[static_cast<long long>(b) * T + t] != 0.0f;
    const uint32_t m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[w] = m;
}
// pos[b][t] = running count of frames with x[b][t][0] != 0, or 0 where it is zero (make_positions on x[..., 0], padding_idx = 0: tts_modules.py:294)
__global__ void fft_pos_kernel(const float* __restrict__ x, int T, int* __restrict__ pos) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        const int f = (t < T && x[(static_cast<long long>(b) * T + t) * kC] != 0.0f) ? 1 : 0;
        int v = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        int base = carry_s;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        if (t < T) pos[static_cast<long long>(b) * T + t] = f ? base + v : 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < nw; ++w) tot += wsum[w];
            carry_s += tot;
        }
        __syncthreads();
    }
}

enum : int { FROW_POS = 0, FROW_LN = 1, FROW_LN_FINAL = 2 };
struct FRowArgs {
    const float* in;       // POS: the decoder input ; LN*: the residual stream X
    float* x;              // POS: X out ; LN (mask_first): X written back masked
    __nv_bfloat16* a_hi;   // LN*: LayerNorm output as the next GEMM's operand pair
    __nv_bfloat16* a_lo;
    float* hidden;         // LN_FINAL: LayerNorm(x) * nonpad, f32 (the decoder's return value), or null
    const float* gamma;
    const float* beta;
    const float* freq;     // POS: [C/2]
    const float* nonpad;   // [rows]
    const int* pos;
    long long rows;
    float alpha;           // POS: pos_embed_alpha
    int use_pos;
    int mask_first;        // LN: x *= nonpad before the LayerNorm, written back (EncSALayer :714,722: the mask follows each residual add)
};
template <int MODE>
__global__ void __launch_bounds__(256) fft_row_kernel(const FRowArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int c0 = lane * 8;
    float p0[8], p1[8];
    if (MODE == FROW_POS) {
        const float4 fa = __ldg(reinterpret_cast<const float4*>(a.freq + (c0 & (kC / 2 - 1)))), fb = __ldg(reinterpret_cast<const float4*>(a.freq + (c0 & (kC / 2 - 1)) + 4));
        p0[0] = fa.x; p0[1] = fa.y; p0[2] = fa.z; p0[3] = fa.w; p0[4] = fb.x; p0[5] = fb.y; p0[6] = fb.z; p0[7] = fb.w;
    } else {
        f_load8(a.gamma + c0, p0);
        f_load8(a.beta + c0, p1);
    }
    for (long long r = warp; r < a.rows; r += nwarps) {
        float v[8];
        f_load8(a.in + r * kC + c0, v);
        const float m = a.nonpad[r];
        if (MODE == FROW_POS) {      // tts_modules.py:293-298: x = (x + alpha * embed_positions(x[..., 0])) * nonpadding
            const int p = a.use_pos ? a.pos[r] : 0;
            if (p != 0) {
                const float pf = static_cast<float>(p);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float ang = pf * p0[i];
                    v[i] = fmaf(a.alpha, c0 < kC / 2 ? sinf(ang) : cosf(ang), v[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= m;
            f_store8(a.x + r * kC + c0, v);
            continue;
        }
        if (a.mask_first) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= m;
            f_store8(a.x + r * kC + c0, v);
        }
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
        const float mean = f_warp_sum(s) * (1.0f / kC);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
        const float rstd = rsqrtf(f_warp_sum(q) * (1.0f / kC) + 1e-5f);       // nn.LayerNorm default eps (common_layers.py:87-95)
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i] * rstd, p0[i], p1[i]);
        if (MODE == FROW_LN_FINAL) {   // tts_modules.py:303-304: layer_norm(x) * nonpadding
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= m;
            if (a.hidden != nullptr) f_store8(a.hidden + r * kC + c0, v);
        }
        f_split8(v, a.a_hi + r * kC + c0, a.a_lo + r * kC + c0);
    }
}

// Vt[bh][d][key] = QKV[b][key][2C + h*128 + d]: 32 x 32 tiles through shared memory (coalesced both ways)
__global__ void fft_vt_kernel(const __half* __restrict__ qkv, int B, int T, int Tp, int H, __half* __restrict__ vt) {
    __shared__ __half tile[32][34];
    const int bh = blockIdx.z, b = bh / H, h = bh % H;
    const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int key = k0 + i;
        tile[i][tx] = key < T ? qkv[(static_cast<long long>(b) * T + key) * (3 * H * 128) + 2 * H * 128 + h * 128 + d0 + tx] : __float2half(0.0f);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int key = k0 + tx;
        if (key < Tp) vt[(static_cast<long long>(bh) * 128 + d0 + i) * Tp + key] = tile[tx][i];
    }
}

// mel_out[row][m] = y[row][m] * tgt_nonpad[row]   (fs2.py:239-240); y is the mel_out GEMM's 256-column scratch
__global__ void fft_mel_kernel(const float* __restrict__ y, const float* __restrict__ tgt_nonpad, long long rows, int M, float* __restrict__ mel) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= rows * M) return;
    const long long r = i / M;
    const int m = static_cast<int>(i % M);
    mel[i] = y[r * kC + m] * (tgt_nonpad ? tgt_nonpad[r] : 1.0f);
}

std::vector<float> f_take(const float*& p, const float* end, size_t n) {
    B200_CHECK(p + n <= end, "weight blob too short");
    std::vector<float> v(p, p + n);
    p += n;
    return v;
}
}  // namespace

struct FftDecoderPlan::Workspace {
    int B = 0, T = 0, Tp = 0;
    DevBuf nonpad, keybits, pos, X, a_hi, a_lo, qkv, vt, f_hi, f_lo, Y;
    unsigned long long last_use = 0;
};

FftDecoderPlan::FftDecoderPlan(const bsg_fft_config& c, const float* w, size_t n_w, int device) : cfg(c), device(device) {
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    B200_CHECK(prop.major == 10, "bisinger_b200 requires an sm_100 (B200) device -- there is no fallback path");
    B200_CHECK(c.hidden_size == kC, "the FFT-block kernels are built for hidden_size 256");
    B200_CHECK(c.num_heads >= 1 && c.hidden_size == c.num_heads * 128, "the attention kernel is built for 128-wide heads (hidden 256, 2 heads)");
    B200_CHECK(c.num_layers >= 1 && c.ffn_kernel % 2 == 1 && c.ffn_kernel <= kMaxTaps, "bad layer count / FFN kernel size (odd, <= 11)");
    B200_CHECK(c.out_dims >= 0 && c.out_dims <= kC, "out_dims must be <= 256");
    const float* p = w;
    const float* end = w + n_w;
    auto pack = [&](Conv& cv, int cout, int cin, int k, bool has_bias, int cout_pad = 0) {
        auto wt = f_take(p, end, static_cast<size_t>(cout) * cin * k);
        std::vector<float> bs = has_bias ? f_take(p, end, cout) : std::vector<float>(cout, 0.0f);
        const int co = cout_pad ? cout_pad : cout;
        std::vector<float> m(static_cast<size_t>(co) * k * cin, 0.0f);
        for (int o = 0; o < cout; ++o)
            for (int ci = 0; ci < cin; ++ci)
                for (int kk = 0; kk < k; ++kk) m[(static_cast<size_t>(o) * k + kk) * cin + ci] = wt[(static_cast<size_t>(o) * cin + ci) * k + kk];
        bs.resize(co, 0.0f);
        cv.w.pack(m, co, k * cin);
        upload(cv.bias, bs);
        cv.cin = cin; cv.cout = co; cv.k = k;
    };
    auto vec = [&](DevBuf& d, size_t n) { upload(d, f_take(p, end, n)); };
    pos_alpha = f_take(p, end, 1)[0];
    vec(pos_freq, kC / 2);
    layers.resize(c.num_layers);
    for (auto& ly : layers) {
        vec(ly.ln1_g, kC); vec(ly.ln1_b, kC);
        pack(ly.in_proj, 3 * kC, kC, 1, false);
        pack(ly.out_proj, kC, kC, 1, false);
        vec(ly.ln2_g, kC); vec(ly.ln2_b, kC);
        pack(ly.ffn1, 4 * kC, kC, c.ffn_kernel, true);
        pack(ly.ffn2, kC, 4 * kC, 1, true);
    }
    vec(ln_g, kC); vec(ln_b, kC);
    if (c.out_dims > 0) pack(mel_out, c.out_dims, kC, 1, true, kC);
    B200_CHECK(p == end, "weight blob has " + std::to_string(n_w) + " floats, consumed " + std::to_string(p - w));
    ConvGemmArgs none{};
    launch_conv_gemm(256, 3, EPI_BIAS_ACT, none, nullptr);
    launch_conv_gemm(256, 3, EPI_BIAS_ACT, none, nullptr, 1);
    B200_CUDA(cudaFuncSetAttribute(fft_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kTotal));
}

FftDecoderPlan::~FftDecoderPlan() = default;

FftDecoderPlan::Workspace& FftDecoderPlan::workspace(int B, int T) {
    const auto key = std::make_pair(B, T);
    auto it = ws.find(key);
    if (it != ws.end()) {
        it->second->last_use = ++use_clock;
        return *it->second;
    }
    while (ws.size() >= 3) {   // least recently used shape out
        auto lru = ws.begin();
        for (auto jt = ws.begin(); jt != ws.end(); ++jt)
            if (jt->second->last_use < lru->second->last_use) lru = jt;
        ws.erase(lru);
    }
    auto w = std::make_unique<Workspace>();
    w->last_use = ++use_clock;
    w->B = B; w->T = T; w->Tp = (T + 7) / 8 * 8;
    const size_t rows = static_cast<size_t>(B) * T;
    w->nonpad.alloc(rows * 4);
    w->keybits.alloc(static_cast<size_t>(B) * ((T + 31) / 32) * 4);
    w->pos.alloc(rows * 4);
    w->X.alloc(rows * kC * 4);
    w->Y.alloc(rows * kC * 4);
    w->a_hi.alloc(rows * kC * 2);
    w->a_lo.alloc(rows * kC * 2);
    w->qkv.alloc(rows * 3 * kC * 2);
    w->vt.alloc(static_cast<size_t>(B) * cfg.num_heads * 128 * w->Tp * 2);
    w->f_hi.alloc(rows * 4 * kC * 2);
    w->f_lo.alloc(rows * 4 * kC * 2);
    auto& ref = *w;
    ws[key] = std::move(w);
    return ref;
}

void FftDecoderPlan::forward(const float* x, const float* tgt_nonpad, int B, int T, float* hidden_out, float* mel, cudaStream_t st,
                             const uint8_t* padding_mask) {
    B200_CHECK(B > 0 && T > 0, "empty batch");
    B200_CHECK(x != nullptr && (hidden_out != nullptr || mel != nullptr), "x and at least one output are required");
    B200_CHECK(mel == nullptr || cfg.out_dims > 0, "this plan was built without the mel_out projection");
    B200_CUDA(cudaSetDevice(device));
    Workspace& w = workspace(B, T);
    const long long rows = static_cast<long long>(B) * T;
    const int row_blocks = static_cast<int>(std::min<long long>((rows + 7) / 8, static_cast<long long>(device_sm_count()) * 8));
    const int H = cfg.num_heads;
    auto count = [&](int n) { launches += n; g_launch_count += n; };

    // one projection / convolution on the bf16x3 GEMM: A = (a_hi, a_lo) [rows][cin]
    auto run_conv = [&](Conv& cv, const void* a_hi, const void* a_lo, const EpiParams& epi) {
        ConvGemmArgs a{};
        const int pair = rows >= 4096 ? 1 : 0;
        set_geometry(a, B, T, cv.cout, 256, pair != 0);
        std::vector<int> shifts;
        for (int j = 0; j < cv.k; ++j) shifts.push_back(j - (cv.k - 1) / 2);   // padding = kernel_size // 2 (common_layers.py:613-615)
        const int rows_box = set_taps(a, 0, 0, cv.cin / kBlockK, shifts.data(), cv.k, cv.cin);
        a.amap[0] = make_act_tmap(a_hi, B, T, cv.cin, cv.cin, rows_box);
        a.amap[1] = make_act_tmap(a_lo, B, T, cv.cin, cv.cin, rows_box);
        cv.w.maps(pair ? 128 : 256, a.wmap[0], a.wmap[1]);
        a.epi = epi;
        a.epi.bias = cv.bias.as<float>();
        a.epi.flags |= BA_ROWS | BA_ROWS_RMW;
        launch_conv_gemm(256, 3, EPI_BIAS_ACT, a, st, pair);
        count(1);
    };
    FRowArgs ra{};
    ra.nonpad = w.nonpad.as<float>();
    ra.pos = w.pos.as<int>();
    ra.a_hi = w.a_hi.as<__nv_bfloat16>();
    ra.a_lo = w.a_lo.as<__nv_bfloat16>();
    ra.rows = rows;

    // ---- padding mask, positions, positional embedding (tts_modules.py:291-298)
    if (padding_mask != nullptr)
        fft_nonpad_from_mask_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, st>>>(padding_mask, rows, w.nonpad.as<float>());
    else
        fft_nonpad_kernel<<<row_blocks, 256, 0, st>>>(x, rows, w.nonpad.as<float>());
    fft_keybits_kernel<<<static_cast<unsigned>((static_cast<long long>(B) * ((T + 31) / 32) * 32 + 255) / 256), 256, 0, st>>>(
        w.nonpad.as<float>(), B, T, w.keybits.as<uint32_t>());
    count(2);
    if (cfg.use_pos_embed) {
        fft_pos_kernel<<<B, 1024, 0, st>>>(x, T, w.pos.as<int>());
        count(1);
    }
    ra.in = x; ra.x = w.X.as<float>(); ra.freq = pos_freq.as<float>(); ra.alpha = pos_alpha; ra.use_pos = cfg.use_pos_embed;
    fft_row_kernel<FROW_POS><<<row_blocks, 256, 0, st>>>(ra);
    count(1);

    for (size_t li = 0; li < layers.size(); ++li) {
        Layer& ly = layers[li];
        // ---- self-attention sub-layer (common_layers.py:703-714)
        ra.in = w.X.as<float>(); ra.x = w.X.as<float>(); ra.gamma = ly.ln1_g.as<float>(); ra.beta = ly.ln1_b.as<float>();
        ra.mask_first = li > 0 ? 1 : 0;          // the previous layer's FFN residual is masked here (:722); layer 0's input is already masked
        fft_row_kernel<FROW_LN><<<row_blocks, 256, 0, st>>>(ra);
        count(1);
        {
            EpiParams e{};                       // q, k, v = in_proj(LN1(x)), no bias -> fp16 [rows][3C]
            e.flags = BA_WRITE_ACT | BA_ACT_F16;
            e.c1 = 1.0f;
            e.out_hi = w.qkv.as<__nv_bfloat16>();
            e.act_pitch = 3 * kC; e.out_pitch = 3 * kC;
            run_conv(ly.in_proj, w.a_hi.p, w.a_lo.p, e);
        }
        fft_vt_kernel<<<dim3((w.Tp + 31) / 32, 4, B * H), 256, 0, st>>>(w.qkv.as<__half>(), B, T, w.Tp, H, w.vt.as<__half>());
        count(1);
        {
            AttnArgs a{};
            a.qkv = make_act_tmap(w.qkv.p, B, T, 3 * kC, 3 * kC, 128);
            a.vt = make_act_tmap(w.vt.p, B * H, 128, w.Tp, w.Tp, 128);
            a.keybits = w.keybits.as<uint32_t>();
            a.o_hi = w.a_hi.as<__nv_bfloat16>();
            a.o_lo = w.a_lo.as<__nv_bfloat16>();
            a.B = B; a.T = T; a.H = H; a.C = kC;
            a.q_tiles = (T + 127) / 128;
            a.scale_log2e = (1.0f / std::sqrt(128.0f)) * 1.4426950408889634f;
            cudaLaunchConfig_t lc{};
            lc.gridDim = dim3(static_cast<unsigned>(B * H * a.q_tiles));
            lc.blockDim = dim3(kAttnThreads);
            lc.dynamicSmemBytes = AttnSmem::kTotal;
            lc.stream = st;
            B200_CUDA(cudaLaunchKernelEx(&lc, fft_attn_kernel, a));
            count(1);
        }
        {
            EpiParams e{};                       // x = x + out_proj(attn)   (the mask of :714 is applied by the next row kernel)
            e.flags = BA_ADD_RES | BA_WRITE_F32;
            e.aux0 = w.X.as<float>(); e.f32_a = w.X.as<float>();
            e.out_pitch = kC; e.act_pitch = kC;
            run_conv(ly.out_proj, w.a_hi.p, w.a_lo.p, e);
        }
        // ---- conv-FFN sub-layer (:716-722, :626-644)
        ra.gamma = ly.ln2_g.as<float>(); ra.beta = ly.ln2_b.as<float>(); ra.mask_first = 1;
        fft_row_kernel<FROW_LN><<<row_blocks, 256, 0, st>>>(ra);
        count(1);
        {
            EpiParams e{};                       // act(ffn_1(LN2(x)) * k^-0.5) -> bf16 hi/lo [rows][4C]
            e.flags = BA_WRITE_ACT | (cfg.ffn_act == 1 ? BA_RELU_SCALED : BA_GELU);
            e.c0 = 1.0f / std::sqrt(static_cast<float>(cfg.ffn_kernel));
            e.c1 = 1.0f;
            e.out_hi = w.f_hi.as<__nv_bfloat16>(); e.out_lo = w.f_lo.as<__nv_bfloat16>();
            e.act_pitch = 4 * kC; e.out_pitch = 4 * kC;
            run_conv(ly.ffn1, w.a_hi.p, w.a_lo.p, e);
        }
        {
            EpiParams e{};                       // x = x + ffn_2(.)
            e.flags = BA_ADD_RES | BA_WRITE_F32;
            e.aux0 = w.X.as<float>(); e.f32_a = w.X.as<float>();
            e.out_pitch = kC; e.act_pitch = kC;
            run_conv(ly.ffn2, w.f_hi.p, w.f_lo.p, e);
        }
    }
    // ---- last mask + final LayerNorm (tts_modules.py:300-304), mel_out (fs2.py:238-240)
    ra.in = w.X.as<float>(); ra.x = w.X.as<float>(); ra.gamma = ln_g.as<float>(); ra.beta = ln_b.as<float>(); ra.mask_first = 1;
    ra.hidden = hidden_out;
    fft_row_kernel<FROW_LN_FINAL><<<row_blocks, 256, 0, st>>>(ra);
    count(1);
    if (mel != nullptr) {
        EpiParams e{};
        e.flags = BA_WRITE_F32;
        e.f32_a = w.Y.as<float>();
        e.out_pitch = kC; e.act_pitch = kC;
        run_conv(mel_out, w.a_hi.p, w.a_lo.p, e);
        const long long n = rows * cfg.out_dims;
        fft_mel_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(w.Y.as<float>(), tgt_nonpad, rows, cfg.out_dims, mel);
        count(1);
    }
    B200_CUDA(cudaGetLastError());
}

}  // namespace b200
