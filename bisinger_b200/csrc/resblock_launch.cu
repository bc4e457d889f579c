// Instantiations + host dispatcher of the fused ResBlock1 iteration kernel (resblock_fused.cuh).
#include <mutex>

#include "runtime.h"
#include "resblock_fused.cuh"

namespace b200 {

// One fused ResBlock1 iteration (resblock_fused.cuh): persistent CTAs over the output-row tiles.  args.num_tiles == 0 only sets the
// kernel attribute up (outside of any stream capture).
template <int C>
static void launch_resblock_inst(const ResblockArgs& args, cudaStream_t stream) {
    using S = ResblockSmem<C>;
    auto kern = resblock_iter_kernel<C>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal); });
    B200_CUDA(attr_err);
    if (args.num_tiles <= 0) return;
    B200_CHECK(args.a_rows % 8 == 0 && args.a_rows * 128 <= S::kASlotBytes, "halo box does not fit the shared-memory slot");
    B200_CHECK(args.ntaps >= 1 && args.ntaps <= kMaxTaps && (args.ntaps & 1), "odd kernel size <= 11 expected");
    B200_CHECK(args.w_slots >= 2 && args.w_slots <= S::kWSlots && args.w_resident >= 0 && args.w_resident <= 2, "weight ring depth");
    const int sms = device_sm_count();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(args.num_tiles < sms ? args.num_tiles : sms);
    cfg.blockDim = dim3(kRbThreads);
    cfg.dynamicSmemBytes = S::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (use_pdl()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
    B200_CUDA(cudaGetLastError());
}
void launch_resblock_iter(int channels, const ResblockArgs& args, cudaStream_t stream) {
    switch (channels) {
        case 32: return launch_resblock_inst<32>(args, stream);
        case 64: return launch_resblock_inst<64>(args, stream);
        case 128: return launch_resblock_inst<128>(args, stream);
        default: throw Error("fused ResBlock kernel: unsupported channel count " + std::to_string(channels));
    }
}
int resblock_weight_slots(int channels) {
    switch (channels) {
        case 32: return ResblockSmem<32>::kWSlots;
        case 64: return ResblockSmem<64>::kWSlots;
        case 128: return ResblockSmem<128>::kWSlots;
        default: return 0;
    }
}

}  // namespace b200
