// Recurrent policy forward (sb3-contrib MlpLstmPolicy) -- placeholder translation unit; the kernels land next.
#include <cuda_runtime.h>
#include <string>
#include "../../include/myo_b200.h"
namespace myo { void set_error(const std::string& msg); }
struct myo_policy { int dummy; };
extern "C" {
int myo_policy_create(const myo_policy_cfg*, int, int, myo_policy**) { myo::set_error("policy kernels not built yet"); return MYO_E_UNSUPPORTED; }
void myo_policy_destroy(myo_policy*) {}
int myo_policy_set_weight(myo_policy*, const char*, const float*, int64_t, void*) { return MYO_E_UNSUPPORTED; }
int myo_policy_forward(myo_policy*, int, const float*, float*, float*, const float*, const float*, float*, float*, float*, void*) { return MYO_E_UNSUPPORTED; }
int64_t myo_policy_launch_count(const myo_policy*) { return 0; }
}
