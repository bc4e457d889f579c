// HiFi-GAN generator with NSF harmonic source on the conv_gemm kernel.
//
// Reference (relative to /root/reference/train_bisinger/):
//   modules/hifigan/hifigan.py:144-173 (HifiGanGenerator.forward), :54-61 (ResBlock1.forward), :104-142 (ctor)
//   modules/parallel_wavegan/models/source.py:45-74,105-138 (SineGen), :386-399 (SourceModuleHnNSF.forward)
//
// Layout: every activation is channels-last, rows = b*L_i + l.  Per up-sampling stage i (C_i channels, L_i samples):
//   X0  f32  [rows][C]   stage input x (transposed conv + noise branch), residual of the first ResBlock iteration
//   A0  bf16 [rows][C]   lrelu(X0): A operand shared by the three ResBlocks
//   Xr  f32  [rows][C]   running residual x of the current ResBlock
//   Ar  bf16 [rows][C]   lrelu(Xr)
//   Tb  bf16 [rows][C]   lrelu(convs1 output)
//   S   f32  [rows][C]   MRF accumulator sum_j ResBlock_j(x) / num_kernels
// For C < 64 the TMA box (64 channels) is wider than the tensor: the out-of-bounds half of every swizzle row is
// zero-filled by the TMA unit and the packed weights carry matching zero columns, so no padded copies are stored.
// ConvTranspose1d (stride u) is run as u independent convolutions, one per output phase p: rows of the input, taps
// delta with 0 <= delta*u + p + pad < k, written to columns [p*Cout, (p+1)*Cout) of the [rows_in][u*Cout] view of the
// output (which is the same memory as [rows_in*u][Cout]).
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>

#include "plans.h"

namespace b200 {

namespace {
constexpr float kLrelu = 0.1f;   // LRELU_SLOPE, hifigan.py:11

// mel f32 [B][M][T] -> bf16 [B*T][M] (A operand of conv_pre)
__global__ void mel_prep_kernel(const float* __restrict__ mel, int B, int M, int T, __nv_bfloat16* __restrict__ out) {
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= static_cast<long long>(B) * T) return;
    const int b = static_cast<int>(r / T), t = static_cast<int>(r % T);
    for (int c = 0; c < M; ++c) out[r * M + c] = __float2bfloat16_rn(mel[(static_cast<long long>(b) * M + c) * T + t]);
}

// Phase at the start of each frame.  source.py:51-74: rad = (f0*h/sr) % 1 per sample, rad[0] += rand_ini, phase = cumsum(rad)
// (the reference subtracts 1 at every wrap only to keep the fp32 running sum small; sin(2*pi*x) is invariant to it).
// f0 is constant over a frame (nearest up-sampling, hifigan.py:113,147), so the cumulative phase at sample i of frame t is
// P[t] + (i+1)*rad[t]; P is accumulated here in fp64 and stored modulo 1.
__global__ void nsf_phase_kernel(const float* __restrict__ f0, const float* __restrict__ rand_ini, int B, int T, int hop, int dim,
                                 float sr, double* __restrict__ phase0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * dim) return;
    const int b = idx / dim, h = idx % dim;
    double p = (h == 0 || rand_ini == nullptr) ? 0.0 : static_cast<double>(rand_ini[b * dim + h]);   // rand_ini[:,0] = 0 (:56)
    for (int t = 0; t < T; ++t) {
        phase0[(static_cast<long long>(b) * T + t) * dim + h] = p;
        const float fh = f0[static_cast<long long>(b) * T + t] * static_cast<float>(h + 1);
        float r = fh / sr;
        r = r - floorf(r);
        p += static_cast<double>(hop) * static_cast<double>(r);
        p -= floor(p);
    }
}

// har_source[b][l] = tanh( Linear_9->1( sine*uv + noise_amp*noise ) )   source.py:121-137,394-395
__global__ void nsf_source_kernel(const float* __restrict__ f0, const double* __restrict__ phase0, const float* __restrict__ noise,
                                  const unsigned long long* __restrict__ seed_ptr, const float* __restrict__ lin, int B, int T, int hop,
                                  int dim, float sr, float* __restrict__ har) {
    const long long L = static_cast<long long>(T) * hop;
    const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= static_cast<long long>(B) * L) return;
    const int b = static_cast<int>(g / L);
    const long long l = g % L;
    const int t = static_cast<int>(l / hop), i = static_cast<int>(l % hop);
    const float f = f0[static_cast<long long>(b) * T + t];
    const float uv = f > 0.0f ? 1.0f : 0.0f;                      // voiced_threshold = 0 (:38-43)
    const float namp = uv * 0.003f + (1.0f - uv) * 0.1f / 3.0f;   // noise_std, sine_amp/3 (:131)
    float acc = lin[dim];                                         // l_linear bias
    for (int h = 0; h < dim; ++h) {
        const float fh = f * static_cast<float>(h + 1);
        float r = fh / sr;
        r = r - floorf(r);
        double ph = phase0[(static_cast<long long>(b) * T + t) * dim + h] + static_cast<double>(i + 1) * static_cast<double>(r);
        ph -= floor(ph);
        const float sine = sinf(static_cast<float>(ph) * 2.0f * 3.14159265358979323846f) * 0.1f;   // sine_amp (:121)
        const long long ni = (static_cast<long long>(b) * L + l) * dim + h;
        const float z = noise ? noise[ni] : philox_normal(__ldg(seed_ptr), 0x4E5Fu, static_cast<uint64_t>(ni));
        acc = fmaf(lin[h], sine * uv + namp * z, acc);
    }
    har[g] = tanhf(acc);
}

// x += layer_norm_C( relu( noise_conv(har) ) ); a = lrelu(x)     hifigan.py:155-160
// noise_conv: Conv1d(1 -> C, kernel ksz, stride s, padding pad).  One warp per output row (grid-stride): lane owns
// channels lane, lane+32, ...; the conv weights sit in shared memory as [k][c] (conflict-free), the <= 32 source
// samples of the row's window are held one per lane and broadcast with shuffles; LayerNorm statistics by warp
// shuffles (two-pass, as F.layer_norm); x is read-modified-written with 128-byte coalesced accesses.
template <int CPL>   // channels per lane = C / 32
__global__ void __launch_bounds__(256) noise_branch_kernel(const float* __restrict__ har, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int B, long long Lout, long long Lhar, int ksz,
                                                           int stride, int pad, int has_source, float* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ act, int act_pitch) {
    constexpr int C = CPL * 32;
    extern __shared__ float wsm[];   // [ksz][C] then bias [C]
    if (has_source) {
        for (int i = threadIdx.x; i < ksz * C; i += blockDim.x) wsm[(i % ksz) * C + i / ksz] = w[i];   // w is [C][ksz]
        for (int i = threadIdx.x; i < C; i += blockDim.x) wsm[ksz * C + i] = bias[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp_g = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long rows = static_cast<long long>(B) * Lout;
    for (long long row = warp_g; row < rows; row += n_warps) {
        float v[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) v[j] = 0.0f;
        if (has_source) {
            const int b = static_cast<int>(row / Lout);
            const long long l = row % Lout;
            const long long h0 = l * stride - pad;
            float hs = 0.0f;
            if (lane < ksz) {
                const long long hl = h0 + lane;
                if (hl >= 0 && hl < Lhar) hs = har[static_cast<long long>(b) * Lhar + hl];
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) v[j] = wsm[ksz * C + lane + 32 * j];
            for (int k = 0; k < ksz; ++k) {
                const float hk = __shfl_sync(0xffffffffu, hs, k);
#pragma unroll
                for (int j = 0; j < CPL; ++j) v[j] = fmaf(wsm[k * C + lane + 32 * j], hk, v[j]);
            }
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) { v[j] = fmaxf(v[j], 0.0f); sum += v[j]; }
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum / static_cast<float>(C);
            float q = 0.0f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) { v[j] -= mean; q += v[j] * v[j]; }
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q / static_cast<float>(C) + 1e-5f);
#pragma unroll
            for (int j = 0; j < CPL; ++j) v[j] *= rstd;
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int c = lane + 32 * j;
            const float xv = x[row * C + c] + v[j];
            x[row * C + c] = xv;
            act[row * act_pitch + c] = __float2bfloat16_rn(xv > 0.0f ? xv : xv * kLrelu);
        }
    }
}

// wav = tanh( conv_post( lrelu(x, 0.01) ) )   hifigan.py:169-171 ; conv_post: Conv1d(C -> 1, k, padding (k-1)/2), C == 32.
// A block stages blockDim + k - 1 rows of lrelu(x) in shared memory (pitch 33: conflict-free) with coalesced loads.
__global__ void __launch_bounds__(256) conv_post_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int B, long long L, int ksz,
                                                        float* __restrict__ wav) {
    constexpr int C = 32, P = 33;
    extern __shared__ float sm[];
    float* wsm = sm;                  // [ksz][C]
    float* xs = sm + ksz * C;         // [blockDim + ksz - 1][P]
    const int half = (ksz - 1) / 2;
    for (int i = threadIdx.x; i < ksz * C; i += blockDim.x) wsm[i] = w[i];
    const long long g0 = static_cast<long long>(blockIdx.x) * blockDim.x;   // first output sample (flattened b*L + l) of this block
    const long long total = static_cast<long long>(B) * L;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nrows = blockDim.x + ksz - 1;
    for (int r = warp; r < nrows; r += nwarp) {
        const long long g = g0 + r - half;          // flattened source row
        float v = 0.0f;
        if (g >= 0 && g < total) {
            v = x[g * C + lane];
            v = v > 0.0f ? v : v * 0.01f;
        }
        xs[r * P + lane] = v;
    }
    __syncthreads();
    const long long g = g0 + threadIdx.x;
    if (g >= total) return;
    const long long l = g % L;
    float acc = bias[0];
    for (int k = 0; k < ksz; ++k) {
        const long long ll = l + k - half;
        if (ll < 0 || ll >= L) continue;              // zero padding at the ends of every batch item
        const float* xr = xs + (threadIdx.x + k) * P;
        const float* ww = wsm + k * C;
#pragma unroll
        for (int c = 0; c < C; ++c) acc = fmaf(ww[c], xr[c], acc);
    }
    wav[g] = tanhf(acc);
}

std::vector<float> take(const float*& p, size_t n) {
    std::vector<float> v(p, p + n);
    p += n;
    return v;
}
int ntile_for(int cout) { return cout >= 256 ? 256 : cout; }
}  // namespace

struct HifiganPlan::Workspace {
    int B = 0, T = 0;
    DevBuf mel16, har, phase0, upin[2], X0, Xr, S, A0, Ar, Tb;
    // graph path (no injected noise): plan-owned copies of the caller's buffers, so that the captured launches see fixed addresses
    DevBuf in_mel, in_f0, out_wav;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};   // [has_src]
    unsigned long long graph_nodes[2] = {0, 0};
    ~Workspace() {
        for (auto& g : graph)
            if (g) cudaGraphExecDestroy(g);
    }
};

// ---------------------------------------------------------------------------------------------
HifiganPlan::HifiganPlan(const bsg_hifigan_config& c, const float* w, size_t n_w, int device) : cfg(c), device(device) {
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    B200_CHECK(prop.major == 10, "bisinger_b200 requires an sm_100 (B200) device -- there is no fallback path");
    B200_CHECK(c.num_upsamples >= 1 && c.num_upsamples <= BSG_MAX_UPSAMPLES, "bad num_upsamples");
    B200_CHECK(c.num_kernels >= 1 && c.num_kernels <= BSG_MAX_RESBLOCK_KERNELS, "bad num_kernels");
    B200_CHECK(c.num_dilations >= 1 && c.num_dilations <= BSG_MAX_RESBLOCK_DILATIONS, "bad num_dilations");
    B200_CHECK(c.num_mels % 8 == 0, "num_mels must be a multiple of 8");
    const int C0 = c.upsample_initial_channel;
    const int dim = c.harmonic_num + 1;
    hop = 1;
    for (int i = 0; i < c.num_upsamples; ++i) hop *= c.upsample_rates[i];

    const float* p = w;
    const float* end = w + n_w;
    auto need = [&](size_t n) { B200_CHECK(p + n <= end, "weight blob too short"); };

    if (c.use_pitch_embed) {
        need(dim + 1);
        auto lw = take(p, dim);
        auto lb = take(p, 1);
        lw.push_back(lb[0]);
        upload(src_lin, lw);
    }
    // a [Cout][Cin][k] conv weight -> K-major [Cout][k * Cp]
    auto pack_conv = [&](Conv& cv, const std::vector<float>& wt, const std::vector<float>& bs, int cout, int cin, int k, int dil) {
        const int cp = ((cin + kBlockK - 1) / kBlockK) * kBlockK;
        std::vector<float> m(static_cast<size_t>(cout) * k * cp, 0.0f);
        for (int o = 0; o < cout; ++o)
            for (int ci = 0; ci < cin; ++ci)
                for (int kk = 0; kk < k; ++kk) m[(static_cast<size_t>(o) * k + kk) * cp + ci] = wt[(static_cast<size_t>(o) * cin + ci) * k + kk];
        cv.w.pack(m, cout, k * cp);
        upload(cv.bias, bs);
        cv.cin = cin; cv.cout = cout; cv.k = k; cv.dilation = dil;
    };
    {
        need(static_cast<size_t>(C0) * c.num_mels * 7 + C0);
        auto wt = take(p, static_cast<size_t>(C0) * c.num_mels * 7);
        auto bs = take(p, C0);
        pack_conv(conv_pre, wt, bs, C0, c.num_mels, 7, 1);
    }
    stages.resize(c.num_upsamples);
    for (int i = 0; i < c.num_upsamples; ++i) {
        Stage& s = stages[i];
        s.cin = C0 >> i;
        s.cout = C0 >> (i + 1);
        s.rate = c.upsample_rates[i];
        s.ksize = c.upsample_kernel_sizes[i];
        B200_CHECK(s.cin % kBlockK == 0, "transposed-conv input channels must be a multiple of 64");
        B200_CHECK(s.cout % 32 == 0 && s.cout >= 32 && (s.cout >= 256 ? s.cout % 256 == 0 : (s.cout & (s.cout - 1)) == 0), "unsupported channel count");
        need(static_cast<size_t>(s.cin) * s.cout * s.ksize + s.cout);
        auto wt = take(p, static_cast<size_t>(s.cin) * s.cout * s.ksize);   // [Cin][Cout][k]
        auto bs = take(p, s.cout);
        const int pad = (s.ksize - s.rate) / 2;
        s.up_phase.resize(s.rate);
        s.up_shifts.resize(s.rate);
        for (int ph = 0; ph < s.rate; ++ph) {
            // taps delta with kernel index kk = delta*u + ph + pad in [0, k); input row = q - delta
            std::vector<int> deltas;
            for (int d = -s.ksize; d <= s.ksize; ++d) {
                const int kk = d * s.rate + ph + pad;
                if (kk >= 0 && kk < s.ksize) deltas.push_back(d);
            }
            B200_CHECK(!deltas.empty() && static_cast<int>(deltas.size()) <= kMaxTaps, "unsupported transposed-conv geometry");
            const int nt = static_cast<int>(deltas.size());
            std::vector<float> wk(static_cast<size_t>(s.cout) * s.cin * nt);   // as a [Cout][Cin][nt] conv
            for (int o = 0; o < s.cout; ++o)
                for (int ci = 0; ci < s.cin; ++ci)
                    for (int j = 0; j < nt; ++j) {
                        const int kk = deltas[j] * s.rate + ph + pad;
                        wk[(static_cast<size_t>(o) * s.cin + ci) * nt + j] = wt[(static_cast<size_t>(ci) * s.cout + o) * s.ksize + kk];
                    }
            pack_conv(s.up_phase[ph], wk, bs, s.cout, s.cin, nt, 1);
            for (int d : deltas) s.up_shifts[ph].push_back(-d);
        }
    }
    if (c.use_pitch_embed) {
        for (int i = 0; i < c.num_upsamples; ++i) {
            Stage& s = stages[i];
            int stride = 1;
            for (int j = i + 1; j < c.num_upsamples; ++j) stride *= c.upsample_rates[j];
            if (i + 1 < c.num_upsamples) { s.noise_k = 2 * stride; s.noise_stride = stride; s.noise_pad = stride / 2; }
            else { s.noise_k = 1; s.noise_stride = 1; s.noise_pad = 0; }
            need(static_cast<size_t>(s.cout) * s.noise_k + s.cout);
            upload(s.noise_w, take(p, static_cast<size_t>(s.cout) * s.noise_k));
            upload(s.noise_b, take(p, s.cout));
        }
    }
    for (int i = 0; i < c.num_upsamples; ++i) {
        Stage& s = stages[i];
        const int C = s.cout;
        s.convs1.resize(c.num_kernels * c.num_dilations);
        s.convs2.resize(c.num_kernels * c.num_dilations);
        for (int j = 0; j < c.num_kernels; ++j) {
            const int k = c.resblock_kernel_sizes[j];
            B200_CHECK(k % 2 == 1 && k <= kMaxTaps, "resblock kernel size must be odd and <= 11");
            for (int m = 0; m < c.num_dilations; ++m) {
                need(static_cast<size_t>(C) * C * k + C);
                auto wt = take(p, static_cast<size_t>(C) * C * k);
                auto bs = take(p, C);
                pack_conv(s.convs1[j * c.num_dilations + m], wt, bs, C, C, k, c.resblock_dilation_sizes[j][m]);
            }
            for (int m = 0; m < c.num_dilations; ++m) {
                need(static_cast<size_t>(C) * C * k + C);
                auto wt = take(p, static_cast<size_t>(C) * C * k);
                auto bs = take(p, C);
                pack_conv(s.convs2[j * c.num_dilations + m], wt, bs, C, C, k, 1);
            }
        }
    }
    {
        const int C = stages.back().cout;
        need(static_cast<size_t>(C) * 7 + 1);
        auto wt = take(p, static_cast<size_t>(C) * 7);   // [1][C][7]
        auto bs = take(p, 1);
        std::vector<float> m(static_cast<size_t>(7) * C);
        for (int ci = 0; ci < C; ++ci)
            for (int kk = 0; kk < 7; ++kk) m[kk * C + ci] = wt[ci * 7 + kk];
        upload(post_w, m);
        upload(post_b, bs);
    }
    B200_CHECK(p == end, "weight blob has " + std::to_string(n_w) + " floats, consumed " + std::to_string(p - w));
    d_seed.alloc(sizeof(unsigned long long));
    ConvGemmArgs none{};
    for (int nt : {256, 128, 64, 32}) launch_conv_gemm(nt, 1, EPI_BIAS_ACT, none, nullptr);
    if (const char* np = std::getenv("BSG_VOC_PAIR")) pair_mode = np[0] == '1';
    if (const char* ng = std::getenv("BSG_VOC_GRAPH")) use_graphs = ng[0] == '1';
    if (pair_mode) for (int nt : {256, 128}) launch_conv_gemm(nt, 1, EPI_BIAS_ACT, none, nullptr, 1);
}

HifiganPlan::~HifiganPlan() = default;

HifiganPlan::Workspace& HifiganPlan::workspace(int B, int T) {
    const auto key = std::make_pair(B, T);
    auto it = ws.find(key);
    if (it != ws.end()) return *it->second;
    ws.clear();   // one shape at a time: the stage buffers are large
    auto w = std::make_unique<Workspace>();
    w->B = B;
    w->T = T;
    const size_t BT = static_cast<size_t>(B) * T;
    const int dim = cfg.harmonic_num + 1;
    w->mel16.alloc(BT * cfg.num_mels * 2);
    w->har.alloc(BT * hop * 4);
    w->phase0.alloc(BT * dim * 8);
    size_t max_f32 = 0, max_b16 = 0, max_up = BT * cfg.upsample_initial_channel * 2;
    size_t rows = BT;
    for (size_t i = 0; i < stages.size(); ++i) {
        rows *= stages[i].rate;
        const size_t C = stages[i].cout;
        max_f32 = std::max(max_f32, rows * C * 4);
        max_b16 = std::max(max_b16, rows * C * 2);
        if (i + 1 < stages.size()) max_up = std::max(max_up, rows * C * 2);
    }
    w->upin[0].alloc(max_up);
    w->upin[1].alloc(max_up);
    w->X0.alloc(max_f32);
    w->Xr.alloc(max_f32);
    w->S.alloc(max_f32);
    w->A0.alloc(max_b16);
    w->Ar.alloc(max_b16);
    w->Tb.alloc(max_b16);
    auto& ref = *w;
    ws[key] = std::move(w);
    return ref;
}

void HifiganPlan::run_source(Workspace& w, const float* f0, const float* rand_ini, const float* src_noise, unsigned long long seed, int B,
                             int T, cudaStream_t st) {
    const int dim = cfg.harmonic_num + 1;
    B200_CHECK(cfg.use_pitch_embed, "this generator was built without the NSF source (use_pitch_embed = 0)");
    nsf_phase_kernel<<<(B * dim + 63) / 64, 64, 0, st>>>(f0, rand_ini, B, T, hop, dim, static_cast<float>(cfg.audio_sample_rate),
                                                         w.phase0.as<double>());
    const long long n = static_cast<long long>(B) * T * hop;
    nsf_source_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(f0, w.phase0.as<double>(), src_noise,
                                                                              d_seed.as<unsigned long long>(), src_lin.as<float>(), B, T,
                                                                              hop, dim, static_cast<float>(cfg.audio_sample_rate),
                                                                              w.har.as<float>());
    launches += 2, g_launch_count += 2;
    B200_CUDA(cudaGetLastError());
}

void HifiganPlan::source(const float* f0, const float* rand_ini, const float* src_noise, unsigned long long seed, int B, int T, float* har,
                         cudaStream_t st) {
    B200_CHECK(B > 0 && T > 0, "empty batch");
    B200_CUDA(cudaSetDevice(device));
    Workspace& w = workspace(B, T);
    B200_CUDA(cudaMemcpyAsync(d_seed.p, &seed, sizeof(seed), cudaMemcpyHostToDevice, st));
    run_source(w, f0, rand_ini, src_noise, seed, B, T, st);
    B200_CUDA(cudaMemcpyAsync(har, w.har.p, static_cast<size_t>(B) * T * hop * 4, cudaMemcpyDeviceToDevice, st));
}

void HifiganPlan::forward(const float* mel, const float* f0, const float* rand_ini, const float* src_noise, unsigned long long seed, int B,
                          int T, float* wav, cudaStream_t st) {
    B200_CHECK(B > 0 && T > 0, "empty batch");
    B200_CUDA(cudaSetDevice(device));
    Workspace& w = workspace(B, T);
    const size_t BT = static_cast<size_t>(B) * T;
    const int has_src = f0 != nullptr ? 1 : 0;
    B200_CUDA(cudaMemcpyAsync(d_seed.p, &seed, sizeof(seed), cudaMemcpyHostToDevice, st));
    if (!use_graphs || rand_ini != nullptr || src_noise != nullptr) {   // injected randomness (parity tests): plain launches
        enqueue(w, mel, f0, rand_ini, src_noise, seed, B, T, wav, st);
        return;
    }
    // production path: the ~110 launches of one forward are captured once per shape and replayed (no launch gaps at small batches)
    w.in_mel.ensure(BT * cfg.num_mels * 4);
    w.out_wav.ensure(BT * hop * 4);
    if (has_src) w.in_f0.ensure(BT * 4);
    B200_CUDA(cudaMemcpyAsync(w.in_mel.p, mel, BT * cfg.num_mels * 4, cudaMemcpyDeviceToDevice, st));
    if (has_src) B200_CUDA(cudaMemcpyAsync(w.in_f0.p, f0, BT * 4, cudaMemcpyDeviceToDevice, st));
    if (!w.graph[has_src]) {
        cudaStream_t cs;
        B200_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        const unsigned long long before = launches;
        B200_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        try {
            enqueue(w, w.in_mel.as<float>(), has_src ? w.in_f0.as<float>() : nullptr, nullptr, nullptr, seed, B, T, w.out_wav.as<float>(), cs);
        } catch (...) {
            cudaStreamEndCapture(cs, &g);
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            throw;
        }
        B200_CUDA(cudaStreamEndCapture(cs, &g));
        w.graph_nodes[has_src] = launches - before;
        launches = before;
        g_launch_count -= w.graph_nodes[has_src];
        B200_CUDA(cudaGraphInstantiate(&w.graph[has_src], g, 0));
        cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
    }
    B200_CUDA(cudaGraphLaunch(w.graph[has_src], st));
    launches += w.graph_nodes[has_src], g_launch_count += w.graph_nodes[has_src];
    B200_CUDA(cudaMemcpyAsync(wav, w.out_wav.p, BT * hop * 4, cudaMemcpyDeviceToDevice, st));
}

// the launches of one forward pass on stream st (everything but the seed upload)
void HifiganPlan::enqueue(Workspace& w, const float* mel, const float* f0, const float* rand_ini, const float* src_noise,
                          unsigned long long seed, int B, int T, float* wav, cudaStream_t st) {
    const size_t BT = static_cast<size_t>(B) * T;
    const bool has_src = f0 != nullptr;
    if (has_src) run_source(w, f0, rand_ini, src_noise, seed, B, T, st);

    mel_prep_kernel<<<static_cast<unsigned>((BT + 127) / 128), 128, 0, st>>>(mel, B, cfg.num_mels, T, w.mel16.as<__nv_bfloat16>());
    launches += 1, g_launch_count += 1;

    auto run_conv = [&](Conv& cv, const void* a_ptr, int a_pitch, int Lrows, const std::vector<int>& shifts, const EpiParams& epi) {
        ConvGemmArgs a{};
        const int nt = ntile_for(cv.cout);
        // wide stages on 2-CTA tiles (M = 256): a single CTA reads 12 KB (N = 256) / 8 KB (N = 128) of shared memory per MMA and is
        // bound by that; a pair reads 8 / 6 KB per SM
        const int pair = (pair_mode && nt >= 128 && static_cast<long long>(B) * Lrows >= 4096) ? 1 : 0;
        set_geometry(a, B, Lrows, cv.cout, nt, pair != 0);
        const int cp = ((cv.cin + kBlockK - 1) / kBlockK) * kBlockK;
        const int rows_box = set_taps(a, 0, 0, cp / kBlockK, shifts.data(), static_cast<int>(shifts.size()), cp);
        a.amap[0] = make_act_tmap(a_ptr, B, Lrows, cv.cin, a_pitch, rows_box);
        a.amap[1] = a.amap[0];
        cv.w.maps(pair ? nt / 2 : nt, a.wmap[0], a.wmap[1]);
        a.k_steps = cv.cin >= kBlockK ? 0 : (cv.cin + 15) / 16;
        a.w_resident = (!pair && a.n_tiles_n == 1 && static_cast<int>(shifts.size()) * (cp / kBlockK) <= conv_gemm_weight_slots(nt, 1)) ? 1 : 0;
        a.epi = epi;
        a.epi.bias = cv.bias.as<float>();
        launch_conv_gemm(nt, 1, EPI_BIAS_ACT, a, st, pair);
        launches += 1, g_launch_count += 1;
    };
    auto same_shifts = [](int k, int dil) {
        std::vector<int> s;
        for (int j = 0; j < k; ++j) s.push_back((j - (k - 1) / 2) * dil);   // padding (k*d - d)/2, hifigan.py:26-27
        return s;
    };

    {   // conv_pre (hifigan.py:151) + the lrelu of the first up-sampling stage (:153)
        EpiParams e{};
        e.flags = BA_WRITE_ACT;
        e.c1 = kLrelu;
        e.out_hi = w.upin[0].as<__nv_bfloat16>();
        e.act_pitch = conv_pre.cout;
        e.out_pitch = conv_pre.cout;
        run_conv(conv_pre, w.mel16.p, cfg.num_mels, T, same_shifts(7, 1), e);
    }
    int Lcur = T;
    int up_sel = 0;
    const int nk = cfg.num_kernels, nd = cfg.num_dilations;
    for (size_t i = 0; i < stages.size(); ++i) {
        Stage& s = stages[i];
        const int C = s.cout, Cp = C;
        const bool last_stage = i + 1 == stages.size();
        // ---- transposed convolution, one launch per output phase (hifigan.py:154)
        for (int ph = 0; ph < s.rate; ++ph) {
            EpiParams e{};
            e.flags = BA_WRITE_F32;
            e.f32_a = w.X0.as<float>();
            e.out_pitch = s.rate * C;
            e.out_col0 = ph * C;
            run_conv(s.up_phase[ph], w.upin[up_sel].p, s.cin, Lcur, s.up_shifts[ph], e);
        }
        const long long Lout = static_cast<long long>(Lcur) * s.rate;
        const long long rows = static_cast<long long>(B) * Lout;
        __nv_bfloat16* A0 = w.A0.as<__nv_bfloat16>();
        __nv_bfloat16* Ar = w.Ar.as<__nv_bfloat16>();
        __nv_bfloat16* Tb = w.Tb.as<__nv_bfloat16>();
        // ---- harmonic-source branch + lrelu (hifigan.py:155-160, ResBlock1's first leaky_relu :56)
        {
            B200_CHECK(s.noise_k <= 32, "noise branch kernel holds the source window in one warp (kernel size <= 32)");
            const int blocks = device_sm_count() * 8;
            const size_t sm_bytes = (static_cast<size_t>(s.noise_k) * C + C) * sizeof(float);
            const float* nw = has_src ? s.noise_w.as<float>() : nullptr;
            const float* nb = has_src ? s.noise_b.as<float>() : nullptr;
            const long long Lhar = static_cast<long long>(T) * hop;
#define B200_NOISE(CPL)                                                                                                              \
    noise_branch_kernel<CPL><<<blocks, 256, sm_bytes, st>>>(w.har.as<float>(), nw, nb, B, Lout, Lhar, s.noise_k, s.noise_stride, \
                                                           s.noise_pad, has_src ? 1 : 0, w.X0.as<float>(), A0, Cp)
            switch (C / 32) {
                case 1: B200_NOISE(1); break;
                case 2: B200_NOISE(2); break;
                case 4: B200_NOISE(4); break;
                case 8: B200_NOISE(8); break;
                case 16: B200_NOISE(16); break;
                default: throw Error("noise branch: unsupported channel count");
            }
#undef B200_NOISE
            launches += 1, g_launch_count += 1;
            B200_CUDA(cudaGetLastError());
        }
        // ---- MRF: sum_j ResBlock1_j(x) / num_kernels (hifigan.py:161-168, :54-61)
        for (int j = 0; j < nk; ++j) {
            const int k = cfg.resblock_kernel_sizes[j];
            for (int m = 0; m < nd; ++m) {
                Conv& c1 = s.convs1[j * nd + m];
                Conv& c2 = s.convs2[j * nd + m];
                {
                    EpiParams e{};
                    e.flags = BA_WRITE_ACT;
                    e.c1 = kLrelu;
                    e.out_hi = Tb;
                    e.act_pitch = Cp;
                    e.out_pitch = C;
                    run_conv(c1, m == 0 ? A0 : Ar, Cp, static_cast<int>(Lout), same_shifts(k, c1.dilation), e);
                }
                {
                    EpiParams e{};
                    e.flags = BA_ADD_RES;
                    e.aux0 = m == 0 ? w.X0.as<float>() : w.Xr.as<float>();
                    e.out_pitch = C;
                    e.act_pitch = Cp;
                    e.c1 = kLrelu;
                    if (m + 1 < nd) {
                        e.flags |= BA_WRITE_F32 | BA_WRITE_ACT;
                        e.f32_a = w.Xr.as<float>();
                        e.out_hi = Ar;
                    } else {
                        e.flags |= BA_ACCUM_F32B | (j == 0 ? BA_ACCUM_INIT : 0);
                        e.f32_b = w.S.as<float>();
                        e.c0 = 1.0f / static_cast<float>(nk);
                        if (j + 1 == nk && !last_stage) {   // stage output -> lrelu -> next transposed conv (:153)
                            e.flags |= BA_WRITE_ACT | BA_ACT_FROM_B;
                            e.out_hi = w.upin[up_sel ^ 1].as<__nv_bfloat16>();
                            e.act_pitch = C;
                        }
                    }
                    run_conv(c2, Tb, Cp, static_cast<int>(Lout), same_shifts(k, 1), e);
                }
            }
        }
        up_sel ^= 1;
        Lcur = static_cast<int>(Lout);
    }
    {   // lrelu(0.01) -> conv_post -> tanh (hifigan.py:169-171)
        const int C = stages.back().cout;
        const long long n = static_cast<long long>(B) * Lcur;
        B200_CHECK(C == 32, "conv_post kernel is specialised for 32 input channels");
        conv_post_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, (7 * 32 + (256 + 6) * 33) * sizeof(float), st>>>(
            w.S.as<float>(), post_w.as<float>(), post_b.as<float>(), B, Lcur, 7, wav);
        launches += 1, g_launch_count += 1;
        B200_CUDA(cudaGetLastError());
    }
}

}  // namespace b200
