// placeholder -- replaced by the HiFi-GAN/NSF implementation
#include "plans.h"
namespace b200 {
struct HifiganPlan::Workspace {};
HifiganPlan::HifiganPlan(const bsg_hifigan_config& c, const float*, size_t, int device) : cfg(c), device(device) {
    throw Error("HiFi-GAN plan not built yet");
}
HifiganPlan::~HifiganPlan() = default;
void HifiganPlan::forward(const float*, const float*, const float*, const float*, unsigned long long, int, int, float*, cudaStream_t) {}
void HifiganPlan::source(const float*, const float*, const float*, unsigned long long, int, int, float*, cudaStream_t) {}
}  // namespace b200
