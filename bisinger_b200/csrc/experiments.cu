// Hardware experiments (not on the product path): UMMA descriptor semantics that the guides do not pin down.
//   bsg_experiment_rowoffset: A tile of 256 rows x 64 bf16 (SWIZZLE_128B, one TMA box); D = A[r : r+128] * B^T computed by
//   pointing the A descriptor at row r (start address + r*128 B) with base_offset = 0 (mode 0) or ((addr >> 7) & 7) (mode 1).
#include "plans.h"

namespace b200 {

__global__ void __launch_bounds__(128, 1) rowoffset_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                          int row_off, int mode, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA = smem;                 // 256 rows x 128 B = 32 KB
    uint8_t* sB = smem + 32768;         // 64 rows x 128 B = 8 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 32768 + 8192);
        tma_load_2d(sA, &amap, bar, 0, 0);
        tma_load_2d(sB, &bmap, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA) + row_off * 128;
        const uint32_t b0 = smem_u32(sB);
        for (int k = 0; k < 4; ++k) {
            uint64_t da = umma_smem_desc<128>(a0 + k * 32);
            if (mode == 1) da |= static_cast<uint64_t>((a0 >> 7) & 7) << 49;
            const uint64_t db = umma_smem_desc<128>(b0 + k * 32);
            umma_f16(tmem_base, da, db, umma_idesc_bf16(128, 64), k > 0);
        }
        umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    float v[32];
    for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
    }
    (void)v;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// bsg_experiment_f8: D = A16[r : r+128] * B16^T (fp16, kind::f16, K = 64) + A8[r : r+128] * B8^T (A e4m3, B e5m2, kind::f8f6f4,
// K = 128) accumulated into the SAME TMEM accumulator; the 8-bit tiles use the same 128-byte swizzled rows and the same
// row-offset descriptors as the 16-bit ones.  mode bit 0: skip the 16-bit MMAs, bit 1: skip the 8-bit MMAs.
__global__ void __launch_bounds__(128, 1) f8_kernel(const __grid_constant__ CUtensorMap a16, const __grid_constant__ CUtensorMap b16,
                                                   const __grid_constant__ CUtensorMap a8, const __grid_constant__ CUtensorMap b8,
                                                   int row_off, int mode, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* sA16 = smem;                     // 256 x 128 B
    uint8_t* sA8 = smem + 32768;              // 256 x 128 B
    uint8_t* sB16 = smem + 65536;             // 64 x 128 B
    uint8_t* sB8 = smem + 65536 + 8192;       // 64 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 16384);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 65536 + 16384);
        tma_load_2d(sA16, &a16, bar, 0, 0);
        tma_load_2d(sA8, &a8, bar, 0, 0);
        tma_load_2d(sB16, &b16, bar, 0, 0);
        tma_load_2d(sB8, &b8, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        uint32_t acc = 0;
        if (!(mode & 1))
            for (int k = 0; k < 4; ++k) {
                umma_f16(tmem_base, umma_smem_desc<128>(smem_u32(sA16) + row_off * 128 + k * 32), umma_smem_desc<128>(smem_u32(sB16) + k * 32),
                         umma_idesc_f16(128, 64, true), acc);
                acc = 1;
            }
        if (!(mode & 2))
            for (int k = 0; k < 4; ++k) {
                umma_f8(tmem_base, umma_smem_desc<128>(smem_u32(sA8) + row_off * 128 + k * 32), umma_smem_desc<128>(smem_u32(sB8) + k * 32),
                        umma_idesc_f8(128, 64, 0, 1), acc);
                acc = 1;
            }
        umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

}  // namespace b200

extern "C" int bsg_experiment_f8(const void* a16_dev /*[256][64] fp16*/, const void* b16_dev /*[64][64] fp16*/, const void* a8_dev /*[256][128] e4m3*/,
                                 const void* b8_dev /*[64][128] e5m2*/, int row_off, int mode, float* out_dev /*[128][64]*/) {
    using namespace b200;
    try {
        const uint64_t d16a[2] = {64, 256}, d16b[2] = {64, 64}, d8a[2] = {128, 256}, d8b[2] = {128, 64}, str[1] = {128};
        const uint32_t bx16a[2] = {64, 256}, bx16b[2] = {64, 64}, bx8a[2] = {128, 256}, bx8b[2] = {128, 64};
        CUtensorMap a16 = make_tmap_bf16(a16_dev, 2, d16a, str, bx16a);
        CUtensorMap b16 = make_tmap_bf16(b16_dev, 2, d16b, str, bx16b);
        CUtensorMap a8 = make_tmap_bf16(a8_dev, 2, d8a, str, bx8a, true);
        CUtensorMap b8 = make_tmap_bf16(b8_dev, 2, d8b, str, bx8b, true);
        const int smem = 65536 + 16384 + 64 + 1024;
        B200_CUDA(cudaFuncSetAttribute(f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        f8_kernel<<<1, 128, smem>>>(a16, b16, a8, b8, row_off, mode, out_dev);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaDeviceSynchronize());
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "experiment failed: %s\n", e.what());
        return 1;
    }
}

extern "C" int bsg_experiment_rowoffset(const void* a_bf16_dev /*[256][64]*/, const void* b_bf16_dev /*[64][64]*/, int row_off, int mode,
                                        float* out_dev /*[128][64]*/) {
    using namespace b200;
    try {
        const uint64_t adims[2] = {64, 256}, astr[1] = {128};
        const uint32_t abox[2] = {64, 256};
        const uint64_t bdims[2] = {64, 64};
        const uint32_t bbox[2] = {64, 64};
        CUtensorMap am = make_tmap_bf16(a_bf16_dev, 2, adims, astr, abox);
        CUtensorMap bm = make_tmap_bf16(b_bf16_dev, 2, bdims, astr, bbox);
        const int smem = 32768 + 8192 + 64 + 1024;
        B200_CUDA(cudaFuncSetAttribute(rowoffset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rowoffset_kernel<<<1, 128, smem>>>(am, bm, row_off, mode, out_dev);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaDeviceSynchronize());
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "experiment failed: %s\n", e.what());
        return 1;
    }
}
