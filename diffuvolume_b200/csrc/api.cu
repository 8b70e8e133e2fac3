// api.cu — library introspection entry points of the C-ABI (include/dv_b200.h).
#include "common.cuh"

namespace dv {
std::atomic<int64_t> g_launch_count{0};
}

extern "C" int dv_version(void) { return 10000 * 0 + 100 * 4 + 0; }   // 0.4.0: warp backward, fused backward passes, refinement-input assembly, softmax uncertainty vote, tcgen05 all-pairs; 0.3.0: caller-owned tile counters; 0.2.0: packed geo pyramid, f4, bf16 volumes

extern "C" int dv_built_for_sm(void) { return 100; }

extern "C" int64_t dv_launch_count(void) { return dv::g_launch_count.load(std::memory_order_relaxed); }

extern "C" const char *dv_status_string(int status) {
    switch (status) {
        case DV_OK: return "ok";
        case DV_ERR_BAD_SHAPE: return "bad shape";
        case DV_ERR_BAD_DTYPE: return "bad dtype";
        case DV_ERR_MISALIGNED: return "misaligned pointer";
        case DV_ERR_LAUNCH: return "kernel launch failed";
        case DV_ERR_NULL: return "required pointer is NULL";
        case DV_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown status";
    }
}
