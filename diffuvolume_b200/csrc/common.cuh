// common.cuh — shared device/host helpers for the DiffuVolume B200 hot-path kernels (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdlib>
#include "../../include/dv_b200.h"

namespace dv {

extern std::atomic<int64_t> g_launch_count;

inline int finish_launch(int nlaunches = 1) {
    g_launch_count.fetch_add(nlaunches, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? DV_OK : DV_ERR_LAUNCH;
}

// Kernel-variant switches for tuning runs (scripts/tune_kernels.py); unset = the shipped default.  The environment is
// read ONCE per process and call site (function-local static), never on the launch path.
inline int read_env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}
#define DV_TUNE(name, dflt)                                      \
    ([&]() -> int {                                              \
        static const int dv_tune_v = ::dv::read_env_int(name, dflt); \
        return dv_tune_v;                                        \
    }())

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// SM count of the current device (B200: 148 = 2 dies x 74), queried once per device ordinal.
inline int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ---- mbarrier + bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    // make the barrier initialisation visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned).
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 1-D bulk copy shared -> global (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    // order generic-proxy smem writes before async-proxy (bulk store) reads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tensor-map TMA (cp.async.bulk.tensor, SASS: UTMALDG) ------------------------------------
// The driver's encoder is fetched through the runtime (no link-time dependency on libcuda).
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// Dense fp32 tensor of `rank` dims (dims[0] fastest, contiguous), tile box `box`; no swizzle; OOB reads give 0.
inline bool make_tensor_map_f32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint32_t *box) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[5], gstride[5];
    cuuint32_t bdim[5], estride[5];
    uint64_t stride = 4;
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estride[i] = 1;
        stride *= dims[i];
        if (i + 1 < rank) {
            if (stride % 16 != 0) return false;
            gstride[i] = stride;   // byte stride of dim i+1
        }
    }
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<void *>(base), gdim, gstride,
               bdim, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// ---- streaming 128-bit global accesses ------------------------------------------------------
__device__ __forceinline__ void stg_cs(float4 *p, const float4 &v) {
    // evict-first streaming store: the volume is written once and consumed by a later kernel
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// 4 consecutive volume elements in the output dtype: fp32 = one 128-bit streaming store, bf16 (round-to-nearest-even,
// what torch's .to(torch.bfloat16) does) = one 64-bit streaming store; a warp still writes one contiguous run.
__device__ __forceinline__ void store4_cs(float *p, const float4 &v) { stg_cs(reinterpret_cast<float4 *>(p), v); }
__device__ __forceinline__ void store4_cs(__nv_bfloat16 *p, const float4 &v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(*reinterpret_cast<const uint32_t *>(&lo)),
                 "r"(*reinterpret_cast<const uint32_t *>(&hi))
                 : "memory");
}

__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ float ldg_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// n = ((clamp(xt + shift, -s, s) / s) + 1) / 2 in the dtype of xt, then float()
// (SceneFlow/models/acv_ddim.py:256-260)
template <typename T>
__device__ __noinline__ T divide_noinline(T a, T b) {
    return a / b;
}

template <typename T>
__device__ __forceinline__ T filter_n(T xt, float shift, T s) {
    T v = xt + static_cast<T>(shift);
    v = v < -s ? -s : (v > s ? s : v);
    // v / s is the identity for the reference's scale (self.scale = 1.0, acv_ddim.py:131) and (r + 1) / 2 == (r + 1) * 0.5
    // exactly: no division is issued on the hot path (an fp64 division is ~30 instructions; ncu r01b showed the fused DDIM
    // step executing 237 instructions per element because of five of them).
    // (ptxas if-converts `s == 1 ? v : v / s` and runs the inlined division sequence regardless — 19 DFMA + 2 MUFU per
    // element in ncu r01c — so the general case lives behind a call that cannot be speculated)
    T r = v;
    if (s != static_cast<T>(1)) r = divide_noinline(v, s);
    return (r + static_cast<T>(1)) * static_cast<T>(0.5);
}

// x / b with a precomputed inv = 1 / b: one Newton correction on the residual reproduces the correctly rounded quotient
// (Markstein) in 3 FMAs instead of a full division sequence.
__device__ __forceinline__ double div_by_const(double x, double b, double inv) {
    const double q0 = x * inv;
    const double rem = fma(-b, q0, x);
    return fma(rem, inv, q0);
}

// concat_stream.cu: TMA-fed streaming producer; DV_ERR_UNSUPPORTED when tensor maps cannot be built
int launch_concat_stream(const float *ref, const float *tgt, float *out, int B, int C, int HW, int W, int D, int mask_left,
                         const float *wts, const float *nf, int *tile_counters, cudaStream_t st);

// allpairs_tcgen05.cu: a14 on tcgen05 / TMEM; DV_ERR_UNSUPPORTED when the shape does not fit its tiling
int launch_allpairs_tcgen05(const float *f1, const float *f2, float *out, float *pooled, int64_t B, int64_t C, int64_t H,
                            int64_t W1, int64_t W2, cudaStream_t st);

}  // namespace dv
