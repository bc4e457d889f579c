// One DiffNet ResidualBlock (usr/diff/net.py:58-78) as ONE persistent 2-CTA kernel: the dilated-conv gate GEMM of both
// channel halves and the residual half of the output projection of a 256-row tile run as a stream of three GEMM "ops" on
// the same producer / MMA / epilogue pipeline as conv_gemm_kernel:
//
//     G(n,0)  G(n,1)   : acc[256 rows x 256] = conv_dilated(xa)[:, gate|filter rows of channel half h]   (K = 3 x 256)
//                        epilogue EPI_GATE: + conditioner projection, sigmoid*tanh -> z (fp16, all-layer z matrix)
//     R(n)             : acc[256 rows x 256] = z[rows, layer columns] * W_res^T                            (K = 256)
//                        epilogue EPI_RES_SKIP: x = (x + acc + b)/sqrt2 -> x f32, fp16(x + d_next) -> next layer's conv input
//
// R(n) reads the z rows the same CTA has just written: they go to global memory anyway (the skip sum of the step is one
// K = L*C GEMM over that matrix), come back through L2 by TMA and need no shared-memory tile; the epilogue warps order
// their global stores before the async-proxy reads with fence.proxy.async + an mbarrier the TMA producer waits on.
// Ops are issued in the order  G(0,0) G(0,1) | G(1,0) R(0) G(1,1) | G(2,0) R(1) G(2,1) | ... | R(last)  so that the gate
// epilogue of tile n (which R(n) depends on) overlaps the first gate GEMM of tile n+1; TMEM holds two 256-column
// accumulators that alternate op by op exactly like the tile double-buffering of conv_gemm_kernel.
// Versus two launches per layer this removes the residual kernel's launch, its HBM-bound pass (it was 1/3 of the layer
// time with the tensor pipe at 21 %, profiles/r01_g) and one full read of z.
#pragma once
#include "conv_gemm.cuh"

namespace b200 {

struct LayerArgs {
    CUtensorMap xa;        // conv input fp16 [B][T][C], box = 64 channels x a_rows rows
    CUtensorMap z;         // all-layer gated activations fp16 [B][T][L*C], box = 64 x 128
    CUtensorMap wg[2];     // dilated-conv weights fp16 hi / lo [2C (gate/filter permuted)][3C], box = 64 x 128
    CUtensorMap wr[2];     // residual half of the output projection fp16 hi / lo [C][C], box = 64 x 128
    int B, T;
    int tiles_per_batch;   // ceil(T / 256)
    int n_row_tiles;       // B * tiles_per_batch
    int dilation;
    int a_rows;            // rows of the xa halo box (128 + 2 * max dilation, multiple of 8)
    int z_col0;            // first column of this layer in the z matrix
    EpiParams gate;        // EPI_GATE parameters   (aux0 = conditioner projection, out_hi = z, ...)
    EpiParams res;         // EPI_RES_SKIP parameters (f32_a = x, out_hi = next xa, dvec = next step embedding, ...)
    unsigned long long* trace;
};

// op j of a CTA pair that owns cnt row tiles (3 * cnt ops): kind 0 = gate GEMM of channel half h, 1 = residual GEMM, of local row tile n
struct LayerOp {
    int kind, n, h;
};
__device__ __forceinline__ LayerOp layer_op(int j, int cnt) {
    if (j < 2) return LayerOp{0, 0, j};
    const int q = (j - 2) / 3, r = (j - 2) % 3;
    if (q == cnt - 1 || r == 1) return LayerOp{1, q, 0};
    return LayerOp{0, q + 1, r == 0 ? 0 : 1};
}


// ---------------------------------------------------------------------------------------------
// Epilogues of the fused layer kernel.  Same arithmetic and data ownership as EPI_GATE / EPI_RES_SKIP of conv_gemm.cuh
// (lane = two neighbouring columns of 16 rows after the shared-memory transpose), but software-pipelined ACROSS chunks
// and ops: the fp32 operands a chunk needs from global memory (conditioner projection / residual stream) are loaded
// into registers while the previous chunk -- possibly of the previous op -- is still doing its math and stores, so
// the L2 round trip is off the critical path (the unfused epilogues exposed four of them per op and paced the kernel:
// profiles/r01_g).
// ---------------------------------------------------------------------------------------------
#ifndef B200_GATE_MATH
#define B200_GATE_MATH 2
#endif
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// sigmoid(g) * tanh(f)   (net.py:73-74)
__device__ __forceinline__ float gate_act(float g, float f) {
#if B200_GATE_MATH == 2
    // two MUFU ops: sigmoid(g) = 0.5 * tanh(g / 2) + 0.5 ; tanh.approx.f32 has a relative error of 2^-11, the size of the
    // fp16 rounding z gets anyway
    return fmaf(0.5f, tanh_approx(0.5f * g), 0.5f) * tanh_approx(f);
#elif B200_GATE_MATH == 1
    // three MUFU ops: a = e^-g, b = e^-2f, z = (1 - b) / ((1 + a)(1 + b)); arguments clamped so that nothing overflows to inf/inf
    const float a = __expf(-fmaxf(g, -80.0f));
    const float b = __expf(-2.0f * fmaxf(f, -40.0f));
    return __fdividef(1.0f - b, (1.0f + a) * (1.0f + b));
#else
    return fast_sigmoid(g) * fast_tanh(f);
#endif
}

// The epilogue warps walk a flat sequence of 16-column chunks (gate op: 4 per warp, residual op: 8 per warp).
constexpr int kChunkCols = 16;
constexpr int kGateChunks = 4, kResChunks = 8;
constexpr int kStage16Pitch = 20;                              // floats; conflict-free 16-byte row writes, <= 2-way on the reads
constexpr int kStage16Bytes = 32 * kStage16Pitch * 4;          // per warp
struct EpiChunk {
    int kind;          // 0 = gate op, 1 = residual op, -1 = past the end
    int j, h, c;       // op index, channel half (gate ops), chunk index within the op
    long long row_w;   // global row of the warp's lane 0
    int rows_left;     // T - t_warp - (lane's row within a group of four): row 4*rp + r0 is valid iff 4*rp < rows_left
};
// after the transpose lane l owns columns cc, cc+1 of rows 4*rp + r0 (rp = 0..7): a warp instruction covers four 64-byte row segments
struct LanePos16 {
    int r0, cc;
};
__device__ __forceinline__ LanePos16 lane_pos16(int lane) { return LanePos16{lane >> 3, (lane & 7) * 2}; }

// 32 rows x 16 accumulator columns starting at column c -> transposed ownership, scaled
__device__ __forceinline__ void ld_chunk16_t(uint32_t tacc, int c, uint32_t stage_s, int lane, const LanePos16& lp, float2 (&o)[8], float scale) {
    float v[16];
    ld_acc16(tacc + c, v, scale);
    const uint32_t wr = stage_s + lane * (kStage16Pitch * 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) sts128(wr + j * 16, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const uint32_t rd = stage_s + (lp.r0 * kStage16Pitch + lp.cc) * 4;
#pragma unroll
    for (int rp = 0; rp < 8; ++rp) o[rp] = lds64(rd + rp * (4 * kStage16Pitch * 4));
    __syncwarp();
}

// issue the global loads of chunk k: gate = conditioner projection of the gate / filter columns, residual = x
__device__ __forceinline__ void layer_epi_prefetch(const LayerArgs& a, const EpiChunk& k, int grp, const LanePos16& lp, float2 (&pa)[8],
                                                   float2 (&pb)[8]) {
    if (a.gate.flags & 1) {   // timing ablation: no global operand loads (results wrong)
#pragma unroll
        for (int rp = 0; rp < 8; ++rp) { pa[rp] = make_float2(0.f, 0.f); pb[rp] = make_float2(0.f, 0.f); }
        return;
    }
    if (k.kind == 0) {
        const int col = k.h * 256 + grp * 64 + k.c * kChunkCols + lp.cc;
        const float* p = a.gate.aux0 + (k.row_w + lp.r0) * static_cast<long long>(a.gate.out_pitch) + col;
        const long long st = 4LL * a.gate.out_pitch;
#pragma unroll
        for (int rp = 0; rp < 8; ++rp) {
            if (4 * rp < k.rows_left) { pa[rp] = ld2(p + rp * st); pb[rp] = ld2(p + 128 + rp * st); }
        }
    } else if (k.kind == 1) {
        const int col = grp * 128 + k.c * kChunkCols + lp.cc;
        const float* p = a.res.f32_a + (k.row_w + lp.r0) * 256LL + col;
#pragma unroll
        for (int rp = 0; rp < 8; ++rp) {
            if (4 * rp < k.rows_left) pa[rp] = ld2(p + rp * 1024);
        }
    }
}

// LO8: the weight-correction term of the gate GEMM runs on the fp8 pipe (see below); 0 = two fp16 MMAs per product
constexpr int kLayerSmemBytes = GemmSmem<256, 2, true>::kTotal + kEpiWarps * kStage16Bytes - GemmSmem<256, 2, true>::kXposeBytes;
template <int LO8>
__global__ void __launch_bounds__(kGemmThreads, 1) diffnet_layer_kernel(const __grid_constant__ LayerArgs args) {
    using S = GemmSmem<256, 2, true>;
    constexpr int C = 256;
    constexpr uint32_t kIdesc = umma_idesc_f16(2 * kTileM, 256, /*fp16=*/true);
    constexpr int kTileRows = 2 * kTileM;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + S::kAStages * S::kASlotBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kOperandBytes);
    uint64_t* afull_bar = bars;
    uint64_t* aempty_bar = afull_bar + S::kAStages;
    uint64_t* bfull_bar = aempty_bar + S::kAStages;
    uint64_t* bempty_bar = bfull_bar + S::kBStages;
    uint64_t* tfull_bar = bempty_bar + S::kBStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* zfull_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zfull_bar + 1);
    float* xpose = reinterpret_cast<float*>(smem + S::kOperandBytes + S::kBarBytes);
    static_assert((2 * S::kAStages + 2 * S::kBStages + 5) * 8 + 8 <= S::kBarBytes, "barrier area too small");
    static_assert(kEpiWarps * kStage16Bytes <= S::kXposeBytes + 1024 + 1024, "transpose staging: see kLayerSmemBytes");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int worker = static_cast<int>(blockIdx.x >> 1);
    const int n_workers = static_cast<int>(gridDim.x >> 1);
    const int cnt = (args.n_row_tiles - worker + n_workers - 1) / n_workers;   // row tiles of this CTA pair

    if (threadIdx.x == 32) {
        for (int s = 0; s < S::kAStages; ++s) { mbar_init(&afull_bar[s], 1); mbar_init(&aempty_bar[s], 1); }
        for (int s = 0; s < S::kBStages; ++s) { mbar_init(&bfull_bar[s], 1); mbar_init(&bempty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * kEpiWarps); }
        mbar_init(zfull_bar, 2 * kEpiWarps);   // both gate epilogues of a row tile, every epilogue warp of THIS CTA
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    const bool tr_on = args.trace != nullptr;

    if (warp == 0 && lane == 0) {
        // ================= TMA producer (one per CTA) =================
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        long long w_a = 0, w_b = 0, w_z = 0;
        const long long t_begin = tr_on ? clock64() : 0;
        const uint32_t b_bytes = S::kBSlotBytes * 2;
        for (int j = 0; j < 3 * cnt; ++j) {
            const LayerOp op = layer_op(j, cnt);
            const int kind = op.kind, n = op.n, h = op.h;
            const int m = worker + n * n_workers;
            const int b = m / args.tiles_per_batch;
            const int t0 = (m % args.tiles_per_batch) * kTileRows + rank * kTileM;
            const int n_taps = kind == 0 ? 3 : 1;
            const CUtensorMap* amap = kind == 0 ? &args.xa : &args.z;
            const CUtensorMap* wmap = kind == 0 ? args.wg : args.wr;
            const uint32_t a_bytes = (kind == 0 ? static_cast<uint32_t>(args.a_rows) : static_cast<uint32_t>(kTileM)) * kBlockK * 2 * 2;
            const int a_col0 = kind == 0 ? 0 : args.z_col0;
            const int a_row = kind == 0 ? t0 - args.dilation : t0;
            const int wrow = (kind == 0 ? h * 256 : 0) + rank * 128;
            if (kind == 1) mbar_wait_tr(zfull_bar, static_cast<uint32_t>(n & 1), tr_on, w_z);   // this CTA's z rows of tile n are in global memory
            for (int kb = 0; kb < C / kBlockK; ++kb) {
                mbar_wait_tr(&aempty_bar[as], aph ^ 1, tr_on, w_a);
                if (rank == 0) mbar_arrive_expect_tx(&afull_bar[as], a_bytes);
                tma_load_3d_pair(smem_a + as * S::kASlotBytes, amap, &afull_bar[as], a_col0 + kb * kBlockK, a_row, b);
                if (++as == S::kAStages) { as = 0; aph ^= 1; }
                for (int tp = 0; tp < n_taps; ++tp) {
                    mbar_wait_tr(&bempty_bar[bs], bph ^ 1, tr_on, w_b);
                    uint8_t* sb = smem_b + bs * S::kBSlotBytes;
                    const int wc = tp * C + kb * kBlockK;
                    if (rank == 0) mbar_arrive_expect_tx(&bfull_bar[bs], b_bytes);
                    tma_load_2d_pair(sb, &wmap[0], &bfull_bar[bs], wc, wrow);
                    tma_load_2d_pair(sb + S::kBPartBytes, &wmap[1], &bfull_bar[bs], wc, wrow);
                    if (++bs == S::kBStages) { bs = 0; bph ^= 1; }
                }
            }
        }
        if (tr_on) {
            unsigned long long* t = args.trace + blockIdx.x * 16;
            t[4] = clock64() - t_begin; t[5] = w_a; t[6] = w_b; t[9] = w_z;
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ================= MMA issuer (leader CTA) =================
        int as = 0, bs = 0, it = 0;
        uint32_t aph = 0, bph = 0;
        long long w_t = 0, w_a = 0, w_b = 0;
        const long long t_begin = tr_on ? clock64() : 0;
        for (int j = 0; j < 3 * cnt; ++j) {
            const int kind = layer_op(j, cnt).kind;
            const int acc = it & 1;
            mbar_wait_tr(&tempty_bar[acc], ((it >> 1) & 1) ^ 1, tr_on, w_t);
            tc_fence_after();
            const uint32_t tacc = tmem_base + acc * 256;
            const int n_taps = kind == 0 ? 3 : 1;
            const uint32_t tap_stride = kind == 0 ? static_cast<uint32_t>(args.dilation) * (kBlockK * 2) : 0;   // taps = row offsets 0, d, 2d of the halo tile
            uint32_t accumulate = 0;
            for (int kb = 0; kb < C / kBlockK; ++kb) {
                mbar_wait_tr(&afull_bar[as], aph, tr_on, w_a);
                tc_fence_after();
                const uint32_t a_slot = smem_u32(smem_a + as * S::kASlotBytes);
                for (int tp = 0; tp < n_taps; ++tp) {
                    mbar_wait_tr(&bfull_bar[bs], bph, tr_on, w_b);
                    tc_fence_after();
                    const uint32_t a_op = a_slot + tp * tap_stride;
                    const uint32_t b_hi = smem_u32(smem_b + bs * S::kBSlotBytes);
                    const uint32_t b_lo = b_hi + S::kBPartBytes;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint64_t da = umma_smem_desc<128>(a_op + k * 32);
                        umma_f16_pair(tacc, da, umma_smem_desc<128>(b_hi + k * 32), kIdesc, accumulate);
                        umma_f16_pair(tacc, da, umma_smem_desc<128>(b_lo + k * 32), kIdesc, 1);
                        accumulate = 1;
                    }
                    umma_commit_pair(&bempty_bar[bs]);
                    if (++bs == S::kBStages) { bs = 0; bph ^= 1; }
                }
                umma_commit_pair(&aempty_bar[as]);
                if (++as == S::kAStages) { as = 0; aph ^= 1; }
            }
            umma_commit_pair(&tfull_bar[acc]);
            ++it;
        }
        if (tr_on) {
            unsigned long long* t = args.trace + blockIdx.x * 16;
            t[0] = clock64() - t_begin; t[1] = w_t; t[2] = w_a; t[3] = w_b; t[10] = cnt;
        }
    } else if (warp >= 2 && warp < 2 + kEpiWarps) {
        // ================= Epilogue warps =================
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        const uint32_t stage_s = smem_u32(reinterpret_cast<uint8_t*>(xpose) + (warp - 2) * kStage16Bytes);
        const LanePos16 lp = lane_pos16(lane);
        const uint32_t tempty_remote0 = map_to_cta(smem_u32(&tempty_bar[0]), 0);
        const bool tr = tr_on && warp == 2 && lane == 0;
        long long w_f = 0, t_gate = 0, t_res = 0, t_op = 0;
        const long long t_begin = tr ? clock64() : 0;
        const int n_ops = 3 * cnt;
        const float gsc = args.gate.acc_scale != 0.0f ? args.gate.acc_scale : 1.0f;
        const float rsc = args.res.acc_scale != 0.0f ? args.res.acc_scale : 1.0f;
        const float rs2 = 0.70710678118654752440f;

        auto chunk_at = [&](int j, int c) {
            EpiChunk k;
            k.j = j; k.c = c; k.h = 0; k.row_w = 0; k.rows_left = 0; k.kind = -1;
            if (j >= n_ops) return k;
            const LayerOp op = layer_op(j, cnt);
            const int m = worker + op.n * n_workers;
            const int b = m / args.tiles_per_batch;
            const int t_warp = (m % args.tiles_per_batch) * kTileRows + rank * kTileM + quad * 32;
            k.kind = op.kind; k.h = op.h;
            k.row_w = static_cast<long long>(b) * args.T + t_warp;
            k.rows_left = args.T - t_warp - lp.r0;
            return k;
        };
        auto next_of = [&](const EpiChunk& k) {
            const int n_chunks = k.kind == 0 ? kGateChunks : kResChunks;
            return (k.c + 1 < n_chunks) ? chunk_at(k.j, k.c + 1) : chunk_at(k.j + 1, 0);
        };
        // one chunk: its global operands are already in (pa, pb)
        auto do_chunk = [&](const EpiChunk& k, float2 (&pa)[8], float2 (&pb)[8]) {
            const int acc = k.j & 1;
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
            if (k.c == 0) {
                // L2 prefetch of the fp32 rows the same kind of op will read for the next row tile
                const LayerOp op = layer_op(k.j, cnt);
                if (op.n + 1 < cnt) {
                    const int m2 = worker + (op.n + 1) * n_workers;
                    const int b2 = m2 / args.tiles_per_batch;
                    const int t2 = (m2 % args.tiles_per_batch) * kTileRows + rank * kTileM + quad * 32;
                    const int rows = min(32, args.T - t2);
                    const long long row0 = static_cast<long long>(b2) * args.T + t2;
                    for (int idx = lane; idx < rows * 4; idx += 32) {
                        const int r = idx >> 2, q = idx & 3;
                        const float* p = k.kind == 0
                            ? args.gate.aux0 + (row0 + r) * static_cast<long long>(args.gate.out_pitch) + op.h * 256 + grp * 64 + (q >> 1) * 128 + (q & 1) * 32
                            : args.res.f32_a + (row0 + r) * 256LL + grp * 128 + q * 32;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                    }
                }
                mbar_wait_tr(&tfull_bar[acc], (k.j >> 1) & 1, tr, w_f);
                tc_fence_after();
                if (tr) t_op = clock64();
            }
            if (k.kind == 0) {
                // ---- gate: columns [cg, cg+16) = gate pre-activations, +128 = filter pre-activations of the same channels
                const int cg = grp * 64 + k.c * kChunkCols;
                float2 g[8], f[8];
                ld_chunk16_t(tacc, cg, stage_s, lane, lp, g, gsc);
                ld_chunk16_t(tacc, 128 + cg, stage_s, lane, lp, f, gsc);
                __nv_bfloat16* zp = args.gate.out_hi + (k.row_w + lp.r0) * static_cast<long long>(args.gate.act_pitch) + args.gate.out_col0 +
                                    k.h * 128 + cg + lp.cc;
                const long long st = 4LL * args.gate.act_pitch;
#pragma unroll
                for (int rp = 0; rp < 8; ++rp) {
                    if (4 * rp < k.rows_left && !(args.gate.flags & 4))
                        st_half2(make_float2(gate_act(g[rp].x + pa[rp].x, f[rp].x + pb[rp].x), gate_act(g[rp].y + pa[rp].y, f[rp].y + pb[rp].y)),
                                 zp + rp * st);
                }
            } else {
                // ---- residual: x <- (x + W_res z + b) / sqrt(2)  (net.py:76-78); fp16(x + d_next) is the next layer's conv input
                const int cl = grp * 128 + k.c * kChunkCols + lp.cc;
                float2 o[8];
                ld_chunk16_t(tacc, grp * 128 + k.c * kChunkCols, stage_s, lane, lp, o, rsc);
                const float2 bias = ldg2(args.res.bias + cl);
                float2 d = make_float2(0.f, 0.f);
                if (args.res.dvec != nullptr) d = ldg2(args.res.dvec + cl);
                const long long off0 = (k.row_w + lp.r0) * 256LL + cl;
#pragma unroll
                for (int rp = 0; rp < 8; ++rp) {
                    if (4 * rp < k.rows_left && !(args.gate.flags & 4)) {
                        const long long off = off0 + rp * 1024;
                        const float2 y = make_float2((pa[rp].x + o[rp].x + bias.x) * rs2, (pa[rp].y + o[rp].y + bias.y) * rs2);
                        st2(args.res.f32_a + off, y);
                        if (args.res.dvec != nullptr) st_half2(make_float2(y.x + d.x, y.y + d.y), args.res.out_hi + off);
                    }
                }
            }
            if (k.c + 1 == (k.kind == 0 ? kGateChunks : kResChunks)) {   // op done: hand the accumulator back
                if (k.kind == 0) {
                    // the z rows written above are read back by this CTA's TMA loads of R(n): order the generic-proxy stores
                    // before the async proxy, then signal the producer
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    __threadfence_block();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (k.kind == 0) mbar_arrive(zfull_bar);
                    mbar_arrive_remote(tempty_remote0 + acc * 8);
                }
                if (tr) { if (k.kind == 0) t_gate += clock64() - t_op; else t_res += clock64() - t_op; }
            }
        };

        // ping-pong operand registers: the loads of chunk q+1 are issued before chunk q is processed
        float2 a0[8], b0[8], a1[8], b1[8];
        EpiChunk k0 = chunk_at(0, 0);
        layer_epi_prefetch(args, k0, grp, lp, a0, b0);
        while (k0.kind >= 0) {
            const EpiChunk k1 = next_of(k0);
            layer_epi_prefetch(args, k1, grp, lp, a1, b1);
            do_chunk(k0, a0, b0);
            if (k1.kind < 0) break;
            k0 = next_of(k1);
            layer_epi_prefetch(args, k0, grp, lp, a0, b0);
            do_chunk(k1, a1, b1);
        }
        if (tr) {
            unsigned long long* t = args.trace + blockIdx.x * 16;
            t[7] = clock64() - t_begin; t[8] = w_f; t[11] = t_gate; t[12] = t_res;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace b200
