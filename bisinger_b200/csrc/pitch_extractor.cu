// PitchExtractor (mel -> f0) between the sampler and the vocoder, SURVEY.md section 8f-2.
//
// Reference (relative to /root/reference/train_bisinger/):
//   modules/fastspeech/pe.py:120-150 (PitchExtractor), :8-42 (Prenet), :45-78 (ConvBlock, norm='gn'), :81-117 (ConvStacks)
//   modules/fastspeech/tts_modules.py:194-237 (PitchPredictor), :39-59 (LayerNorm, eps 1e-12)
//   modules/commons/common_layers.py:106-158 (SinusoidalPositionalEmbedding), utils/__init__.py:146-158 (make_positions)
//   utils/pitch_utils.py:63-76 (denorm_f0)
//
// Every convolution / Linear is one launch of conv_gemm_kernel<256, 3, EPI_BIAS_ACT> (bf16 hi/lo split operands, three MMAs per
// product, fp32 accumulation in TMEM): f0 drives the NSF phase accumulator of the vocoder (source.py:51-74), so this stage is
// kept at ~fp32 accuracy; it is 0.25 % of the path's FLOPs.  What sits between two GEMMs (ReLU, BatchNorm(eval), padding mask,
// GroupNorm, residual, positional embedding, LayerNorm, the final Linear -> 2 and denorm_f0) runs in warp-per-row kernels
// (C = 256: eight channels per lane, two float4 per lane and row) that read the GEMM's fp32 output once and write the next GEMM's
// bf16 hi/lo operand pair.  Layout: channels-last, rows = b*T + t; the input is the sampler's mel_out [B][T][80] as it stands.
//   Y   f32  [rows][C]   GEMM output (bias added)
//   X   f32  [rows][C]   residual stream of the ConvStacks encoder
//   A   bf16 [rows][C]   hi / lo operand of the next GEMM
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>

#include "plans.h"

namespace b200 {

namespace {
constexpr int kC = 256;          // PitchExtractor.hidden_size (pe.py:123)
constexpr int kGnChunk = 128;    // rows per GroupNorm partial sum

__device__ __forceinline__ void split8(const float (&v)[8], __nv_bfloat16* hi, __nv_bfloat16* lo) {
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
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ldg8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mel f32 [rows][M] -> bf16 hi/lo [rows][M] and nonpad[row] = (sum |mel[row]| != 0)   (pe.py:30-31,145)
__global__ void pe_prep_kernel(const float* __restrict__ mel, long long rows, int M, __nv_bfloat16* __restrict__ hi,
                               __nv_bfloat16* __restrict__ lo, float* __restrict__ nonpad) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long r = warp; r < rows; r += nwarps) {
        float s = 0.0f;
        for (int c = lane; c < M; c += 32) {
            const float v = mel[r * M + c];
            s += fabsf(v);
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[r * M + c] = h;
            lo[r * M + c] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
        s = warp_sum(s);
        if (lane == 0) nonpad[r] = s == 0.0f ? 0.0f : 1.0f;
    }
}

// partial sums of one GroupNorm: part[b][chunk][g] = (sum, sum of squares) over the chunk's rows and the 16 channels of group g
__global__ void pe_gn_partial_kernel(const float* __restrict__ y, int T, int n_chunks, double* __restrict__ part) {
    __shared__ double sm[8][16][2];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = chunk * kGnChunk, t1 = min(T, t0 + kGnChunk);
    float s = 0.0f, q = 0.0f;
    for (int t = t0 + warp; t < t1; t += 8) {
        float v[8];
        load8(y + (static_cast<long long>(b) * T + t) * kC + lane * 8, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += v[i]; q = fmaf(v[i], v[i], q); }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);   // a group = 16 channels = two lanes
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    if ((lane & 1) == 0) { sm[warp][lane >> 1][0] = s; sm[warp][lane >> 1][1] = q; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x >> 1, w = threadIdx.x & 1;
        double a = 0.0;
        for (int i = 0; i < 8; ++i) a += sm[i][g][w];
        part[((static_cast<long long>(b) * n_chunks + chunk) * 16 + g) * 2 + w] = a;
    }
}
// stats[b][g] = (mean, 1/sqrt(var + eps)), biased variance over T x 16 values (torch.nn.GroupNorm)
__global__ void pe_gn_final_kernel(const double* __restrict__ part, int T, int n_chunks, float eps, float* __restrict__ stats) {
    const int b = blockIdx.x, g = threadIdx.x;
    if (g >= 16) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        s += part[((static_cast<long long>(b) * n_chunks + c) * 16 + g) * 2];
        q += part[((static_cast<long long>(b) * n_chunks + c) * 16 + g) * 2 + 1];
    }
    const double n = 16.0 * T, mean = s / n;
    double var = q / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    stats[(b * 16 + g) * 2] = static_cast<float>(mean);
    stats[(b * 16 + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// pos[b][t] = running count of frames with y[b][t][0] != 0, or 0 where it is zero (make_positions, padding_idx = 0)
__global__ void pe_pos_kernel(const float* __restrict__ y, int T, int* __restrict__ pos) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        const int f = (t < T && y[(static_cast<long long>(b) * T + t) * kC] != 0.0f) ? 1 : 0;
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

enum : int { ROW_BN_MASK = 0, ROW_MASK = 1, ROW_GN_RES = 2, ROW_POS = 3, ROW_LN = 4, ROW_LN_FINAL = 5 };

struct RowArgs {
    const float* y;            // [rows][C] GEMM output
    float* x;                  // GN_RES: residual stream [rows][C] (read-modify-write)
    float* y_out;              // MASK: masked rows are also written back here (they feed the position scan when there is no encoder), or null
    __nv_bfloat16* a_hi;       // next operand (not written by LN_FINAL)
    __nv_bfloat16* a_lo;
    const float* p0;           // BN scale / GN gamma / LN gamma / POS frequencies[C/2]
    const float* p1;           // BN shift / GN beta / LN beta
    const float* nonpad;       // [rows]
    const float* gn_stats;     // [B][16][2]
    const int* pos;            // [rows]
    const float* lin;          // LN_FINAL: weight [2][C] then bias [2]
    float* pitch_pred;         // [rows][2]
    float* f0;                 // [rows]
    long long rows;
    int T;
    float alpha;               // POS: pos_embed_alpha
    float ln_eps;
    int residual;              // GN_RES: ConvStacks.res
    int pitch_norm;            // 0 log, 1 standard, 2 none
    int use_uv;
    float f0_mean, f0_std;
};

template <int MODE>
__global__ void __launch_bounds__(256) pe_row_kernel(const RowArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int c0 = lane * 8;
    float p0[8], p1[8];
    if (MODE == ROW_BN_MASK || MODE == ROW_GN_RES || MODE == ROW_LN || MODE == ROW_LN_FINAL) {
        ldg8(a.p0 + c0, p0);
        ldg8(a.p1 + c0, p1);
    }
    if (MODE == ROW_POS) ldg8(a.p0 + (c0 & (kC / 2 - 1)), p0);   // sin half / cos half share the frequencies (common_layers.py:136)
    float w0[8], w1[8];
    if (MODE == ROW_LN_FINAL) {
        ldg8(a.lin + c0, w0);
        ldg8(a.lin + kC + c0, w1);
    }
    for (long long r = warp; r < a.rows; r += nwarps) {
        float v[8];
        load8(a.y + r * kC + c0, v);
        if (MODE == ROW_BN_MASK) {            // pe.py:15-17,36: conv -> ReLU -> BatchNorm1d(eval), * nonpadding
            const float m = a.nonpad[r];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(fmaxf(v[i], 0.0f), p0[i], p1[i]) * m;
        } else if (MODE == ROW_MASK) {        // pe.py:40-41
            const float m = a.nonpad[r];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= m;
            if (a.y_out != nullptr) store8(a.y_out + r * kC + c0, v);
        } else if (MODE == ROW_GN_RES) {      // pe.py:67-76,109-111: x (+)= ReLU(GroupNorm(conv(x)))
            const int b = static_cast<int>(r / a.T);
            const float2 st = __ldg(reinterpret_cast<const float2*>(a.gn_stats) + b * 16 + (lane >> 1));
            float x[8];
            if (a.residual) load8(a.x + r * kC + c0, x);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float g = fmaxf(fmaf((v[i] - st.x) * st.y, p0[i], p1[i]), 0.0f);
                v[i] = a.residual ? x[i] + g : g;
            }
            store8(a.x + r * kC + c0, v);
        } else if (MODE == ROW_POS) {         // tts_modules.py:230-231
            const int p = a.pos[r];
            if (p != 0) {
                const float pf = static_cast<float>(p);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float ang = pf * p0[i];
                    v[i] = fmaf(a.alpha, c0 < kC / 2 ? sinf(ang) : cosf(ang), v[i]);
                }
            }
        } else {                              // tts_modules.py:216-217: ReLU -> LayerNorm over channels
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] = fmaxf(v[i], 0.0f); s += v[i]; }
            const float mean = warp_sum(s) * (1.0f / kC);
            float q = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] -= mean; q = fmaf(v[i], v[i], q); }
            const float rstd = rsqrtf(warp_sum(q) * (1.0f / kC) + a.ln_eps);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i] * rstd, p0[i], p1[i]);
        }
        if (MODE == ROW_LN_FINAL) {           // tts_modules.py:236 Linear -> 2, pe.py:144-149 denorm_f0
            float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { d0 = fmaf(v[i], w0[i], d0); d1 = fmaf(v[i], w1[i], d1); }
            d0 = warp_sum(d0) + __ldg(a.lin + 2 * kC);
            d1 = warp_sum(d1) + __ldg(a.lin + 2 * kC + 1);
            if (lane == 0) {
                a.pitch_pred[r * 2] = d0;
                a.pitch_pred[r * 2 + 1] = d1;
                float f = d0;
                if (a.pitch_norm == 1) f = fmaf(f, a.f0_std, a.f0_mean);
                if (a.pitch_norm == 0) f = exp2f(f);
                if (a.use_uv && d1 > 0.0f) f = 0.0f;
                if (a.nonpad[r] == 0.0f) f = 0.0f;
                a.f0[r] = f;
            }
        } else {
            split8(v, a.a_hi + r * kC + c0, a.a_lo + r * kC + c0);
        }
    }
}
}  // namespace

struct PitchExtractorPlan::Workspace {
    int B = 0, T = 0, n_chunks = 0;
    DevBuf mel_hi, mel_lo, nonpad, Y, X, a_hi, a_lo, gn_part, gn_stats, pos;
    DevBuf in_mel, out_pred, out_f0;   // plan-owned copies of the caller's buffers: the captured launches see fixed addresses
    cudaGraphExec_t graph = nullptr;
    unsigned long long graph_nodes = 0;
    ~Workspace() {
        if (graph) cudaGraphExecDestroy(graph);
    }
};

static std::vector<float> take_n(const float*& p, const float* end, size_t n) {
    B200_CHECK(p + n <= end, "weight blob too short");
    std::vector<float> v(p, p + n);
    p += n;
    return v;
}

PitchExtractorPlan::PitchExtractorPlan(const bsg_pe_config& c, const float* w, size_t n_w, int device) : cfg(c), device(device) {
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    B200_CHECK(prop.major == 10, "bisinger_b200 requires an sm_100 (B200) device -- there is no fallback path");
    B200_CHECK(c.hidden_size == kC && c.predictor_hidden == kC, "the PitchExtractor kernels are built for 256 channels (pe.py:123)");
    B200_CHECK(c.n_mel_bins % 8 == 0 && c.n_mel_bins >= 8, "n_mel_bins must be a multiple of 8");
    B200_CHECK(c.gn_group_size == 16, "GroupNorm groups are 16 channels wide (pe.py:54)");
    B200_CHECK(c.prenet_layers >= 1 && c.conv_layers >= 0 && c.predictor_layers >= 1, "bad layer counts");
    for (int k : {c.kernel_size, c.predictor_kernel}) B200_CHECK(k % 2 == 1 && k >= 1 && k <= kMaxTaps, "kernel sizes must be odd and <= 11");
    const float* p = w;
    const float* end = w + n_w;
    // a [Cout][Cin][k] conv weight -> K-major [Cout][k * Cp], bf16 hi/lo
    auto pack = [&](Conv& cv, int cout, int cin, int k) {
        auto wt = take_n(p, end, static_cast<size_t>(cout) * cin * k);
        auto bs = take_n(p, end, cout);
        const int cp = ((cin + kBlockK - 1) / kBlockK) * kBlockK;
        std::vector<float> m(static_cast<size_t>(cout) * k * cp, 0.0f);
        for (int o = 0; o < cout; ++o)
            for (int ci = 0; ci < cin; ++ci)
                for (int kk = 0; kk < k; ++kk) m[(static_cast<size_t>(o) * k + kk) * cp + ci] = wt[(static_cast<size_t>(o) * cin + ci) * k + kk];
        cv.w.pack(m, cout, k * cp);
        upload(cv.bias, bs);
        cv.cin = cin; cv.cout = cout; cv.k = k;
    };
    auto vec = [&](DevBuf& d, size_t n) { upload(d, take_n(p, end, n)); };
    prenet.resize(c.prenet_layers);
    for (int i = 0; i < c.prenet_layers; ++i) {
        pack(prenet[i].conv, kC, i == 0 ? c.n_mel_bins : kC, c.kernel_size);
        vec(prenet[i].p0, kC);   // BatchNorm(eval) scale = weight / sqrt(running_var + eps)
        vec(prenet[i].p1, kC);   //                 shift = bias - running_mean * scale
    }
    pack(prenet_out, kC, kC, 1);
    if (c.conv_layers > 0) {
        pack(enc_in, kC, kC, 1);
        encoder.resize(c.conv_layers);
        for (int j = 0; j < c.conv_layers; ++j) {
            pack(encoder[j].conv, kC, kC, c.kernel_size);
            vec(encoder[j].p0, kC);
            vec(encoder[j].p1, kC);
        }
        pack(enc_out, kC, kC, 1);
    }
    pos_alpha = take_n(p, end, 1)[0];
    vec(pos_freq, kC / 2);
    predictor.resize(c.predictor_layers);
    for (int i = 0; i < c.predictor_layers; ++i) {
        pack(predictor[i].conv, kC, kC, c.predictor_kernel);
        vec(predictor[i].p0, kC);
        vec(predictor[i].p1, kC);
    }
    vec(lin, 2 * kC + 2);
    B200_CHECK(p == end, "weight blob has " + std::to_string(n_w) + " floats, consumed " + std::to_string(p - w));
    ConvGemmArgs none{};
    launch_conv_gemm(256, 3, EPI_BIAS_ACT, none, nullptr);
    if (const char* np = std::getenv("BSG_PE_PAIR")) pair_mode = np[0] == '1';
    if (const char* ng = std::getenv("BSG_PE_GRAPH")) use_graphs = ng[0] == '1';
    if (const char* nr = std::getenv("BSG_ROWS_EPI")) rows_epi = nr[0] == '1';
    if (pair_mode) launch_conv_gemm(256, 3, EPI_BIAS_ACT, none, nullptr, 1);
}

PitchExtractorPlan::~PitchExtractorPlan() = default;

PitchExtractorPlan::Workspace& PitchExtractorPlan::workspace(int B, int T) {
    const auto key = std::make_pair(B, T);
    auto it = ws.find(key);
    if (it != ws.end()) return *it->second;
    ws.clear();
    auto w = std::make_unique<Workspace>();
    w->B = B;
    w->T = T;
    w->n_chunks = (T + kGnChunk - 1) / kGnChunk;
    const size_t rows = static_cast<size_t>(B) * T;
    w->mel_hi.alloc(rows * cfg.n_mel_bins * 2);
    w->mel_lo.alloc(rows * cfg.n_mel_bins * 2);
    w->nonpad.alloc(rows * 4);
    w->Y.alloc(rows * kC * 4);
    w->X.alloc(rows * kC * 4);
    w->a_hi.alloc(rows * kC * 2);
    w->a_lo.alloc(rows * kC * 2);
    w->gn_part.alloc(static_cast<size_t>(B) * w->n_chunks * 16 * 2 * 8);
    w->gn_stats.alloc(static_cast<size_t>(B) * 16 * 2 * 4);
    w->pos.alloc(rows * 4);
    w->in_mel.alloc(rows * cfg.n_mel_bins * 4);
    w->out_pred.alloc(rows * 2 * 4);
    w->out_f0.alloc(rows * 4);
    auto& ref = *w;
    ws[key] = std::move(w);
    return ref;
}

void PitchExtractorPlan::forward(const float* mel, int B, int T, float* pitch_pred, float* f0, cudaStream_t st) {
    B200_CHECK(B > 0 && T > 0, "empty batch");
    B200_CUDA(cudaSetDevice(device));
    Workspace& w = workspace(B, T);
    if (!use_graphs) {
        enqueue(w, mel, B, T, pitch_pred, f0, st);
        return;
    }
    // the 31 launches of one forward are captured once per shape and replayed
    const size_t rows = static_cast<size_t>(B) * T;
    B200_CUDA(cudaMemcpyAsync(w.in_mel.p, mel, rows * cfg.n_mel_bins * 4, cudaMemcpyDeviceToDevice, st));
    if (!w.graph) {
        cudaStream_t cs;
        B200_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        const unsigned long long before = launches;
        B200_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        try {
            enqueue(w, w.in_mel.as<float>(), B, T, w.out_pred.as<float>(), w.out_f0.as<float>(), cs);
        } catch (...) {
            cudaStreamEndCapture(cs, &g);
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            throw;
        }
        B200_CUDA(cudaStreamEndCapture(cs, &g));
        w.graph_nodes = launches - before;
        launches = before;
        g_launch_count -= w.graph_nodes;
        B200_CUDA(cudaGraphInstantiate(&w.graph, g, 0));
        cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
    }
    B200_CUDA(cudaGraphLaunch(w.graph, st));
    launches += w.graph_nodes, g_launch_count += w.graph_nodes;
    B200_CUDA(cudaMemcpyAsync(pitch_pred, w.out_pred.p, rows * 2 * 4, cudaMemcpyDeviceToDevice, st));
    B200_CUDA(cudaMemcpyAsync(f0, w.out_f0.p, rows * 4, cudaMemcpyDeviceToDevice, st));
}

void PitchExtractorPlan::enqueue(Workspace& w, const float* mel, int B, int T, float* pitch_pred, float* f0, cudaStream_t st) {
    const long long rows = static_cast<long long>(B) * T;
    const int row_blocks = static_cast<int>(std::min<long long>((rows + 7) / 8, static_cast<long long>(device_sm_count()) * 8));
    auto count = [&](int n) { launches += n; g_launch_count += n; };

    // one convolution / Linear: A = (a_hi, a_lo) [rows][cin] -> y = conv + bias, written as f32 and / or as the next operand pair
    auto run_conv = [&](Conv& cv, const void* a_hi, const void* a_lo, bool left_pad, float* out_f32, bool write_act) {
        ConvGemmArgs a{};
        const int nt = 256;
        const int pair = (pair_mode && rows >= 4096) ? 1 : 0;
        set_geometry(a, B, T, cv.cout, nt, pair != 0);
        std::vector<int> shifts;
        for (int j = 0; j < cv.k; ++j) shifts.push_back(left_pad ? j - (cv.k - 1) : j - (cv.k - 1) / 2);   // tts_modules.py:212-214
        const int cp = ((cv.cin + kBlockK - 1) / kBlockK) * kBlockK;
        const int rows_box = set_taps(a, 0, 0, cp / kBlockK, shifts.data(), cv.k, cp);
        a.amap[0] = make_act_tmap(a_hi, B, T, cv.cin, cv.cin, rows_box);
        a.amap[1] = make_act_tmap(a_lo, B, T, cv.cin, cv.cin, rows_box);
        cv.w.maps(pair ? nt / 2 : nt, a.wmap[0], a.wmap[1]);
        a.k_steps = 0;
        a.epi.bias = cv.bias.as<float>();
        a.epi.out_pitch = cv.cout;
        a.epi.act_pitch = cv.cout;
        if (out_f32) { a.epi.flags |= BA_WRITE_F32; a.epi.f32_a = out_f32; }
        if (write_act) {   // identity activation (slope 1) -> hi/lo operand of the next GEMM
            a.epi.flags |= BA_WRITE_ACT;
            a.epi.c1 = 1.0f;
            a.epi.out_hi = w.a_hi.as<__nv_bfloat16>();
            a.epi.out_lo = w.a_lo.as<__nv_bfloat16>();
        }
        if (rows_epi) a.epi.flags |= BA_ROWS;   // write-only epilogues, 256-channel rows: row-per-thread 256-bit stores
        launch_conv_gemm(nt, 3, EPI_BIAS_ACT, a, st, pair);
        count(1);
    };
    RowArgs ra{};
    ra.y = w.Y.as<float>();
    ra.x = w.X.as<float>();
    ra.a_hi = w.a_hi.as<__nv_bfloat16>();
    ra.a_lo = w.a_lo.as<__nv_bfloat16>();
    ra.nonpad = w.nonpad.as<float>();
    ra.gn_stats = w.gn_stats.as<float>();
    ra.pos = w.pos.as<int>();
    ra.rows = rows;
    ra.T = T;
    ra.ln_eps = 1e-12f;
    ra.residual = 1;
    ra.pitch_norm = cfg.pitch_norm;
    ra.use_uv = cfg.use_uv;
    ra.f0_mean = cfg.f0_mean;
    ra.f0_std = cfg.f0_std;
    ra.pitch_pred = pitch_pred;
    ra.f0 = f0;
#define B200_ROW(MODE)                                         \
    do {                                                       \
        pe_row_kernel<MODE><<<row_blocks, 256, 0, st>>>(ra);   \
        count(1);                                              \
    } while (0)

    pe_prep_kernel<<<row_blocks, 256, 0, st>>>(mel, rows, cfg.n_mel_bins, w.mel_hi.as<__nv_bfloat16>(), w.mel_lo.as<__nv_bfloat16>(),
                                               w.nonpad.as<float>());
    count(1);
    // ---- Prenet (pe.py:32-42)
    for (size_t i = 0; i < prenet.size(); ++i) {
        run_conv(prenet[i].conv, i == 0 ? w.mel_hi.p : w.a_hi.p, i == 0 ? w.mel_lo.p : w.a_lo.p, false, w.Y.as<float>(), false);
        ra.p0 = prenet[i].p0.as<float>();
        ra.p1 = prenet[i].p1.as<float>();
        B200_ROW(ROW_BN_MASK);
    }
    run_conv(prenet_out, w.a_hi.p, w.a_lo.p, false, w.Y.as<float>(), false);
    ra.y_out = encoder.empty() ? w.Y.as<float>() : nullptr;
    B200_ROW(ROW_MASK);
    // ---- ConvStacks encoder (pe.py:105-117)
    if (!encoder.empty()) {
        run_conv(enc_in, w.a_hi.p, w.a_lo.p, false, w.X.as<float>(), true);
        for (size_t j = 0; j < encoder.size(); ++j) {
            run_conv(encoder[j].conv, w.a_hi.p, w.a_lo.p, false, w.Y.as<float>(), false);
            pe_gn_partial_kernel<<<dim3(w.n_chunks, B), 256, 0, st>>>(w.Y.as<float>(), T, w.n_chunks, w.gn_part.as<double>());
            pe_gn_final_kernel<<<B, 32, 0, st>>>(w.gn_part.as<double>(), T, w.n_chunks, 1e-5f, w.gn_stats.as<float>());
            count(2);
            ra.p0 = encoder[j].p0.as<float>();
            ra.p1 = encoder[j].p1.as<float>();
            B200_ROW(ROW_GN_RES);
        }
        run_conv(enc_out, w.a_hi.p, w.a_lo.p, false, w.Y.as<float>(), false);
    }
    // ---- PitchPredictor (tts_modules.py:224-237) + denorm_f0 (pe.py:144-149)
    pe_pos_kernel<<<B, 1024, 0, st>>>(w.Y.as<float>(), T, w.pos.as<int>());
    count(1);
    ra.p0 = pos_freq.as<float>();
    ra.alpha = pos_alpha;
    B200_ROW(ROW_POS);
    for (size_t i = 0; i < predictor.size(); ++i) {
        run_conv(predictor[i].conv, w.a_hi.p, w.a_lo.p, cfg.left_padding != 0, w.Y.as<float>(), false);
        ra.p0 = predictor[i].p0.as<float>();
        ra.p1 = predictor[i].p1.as<float>();
        if (i + 1 < predictor.size()) {
            B200_ROW(ROW_LN);
        } else {
            ra.lin = lin.as<float>();
            B200_ROW(ROW_LN_FINAL);
        }
    }
#undef B200_ROW
    B200_CUDA(cudaGetLastError());
}

}  // namespace b200
