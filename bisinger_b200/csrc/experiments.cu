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

}  // namespace b200

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
