// Micro-benchmark (not product code): pure tensor-map TMA read rate of a [B, D, HW] fp32 volume as a function of the box
// shape — how wide must a tile row be, and how many CTAs per SM, before HBM reads reach the sequential-read rate?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

constexpr int MAXST = 16;
// tile t -> (b, dchunk, span): span fastest
__global__ void tma_read(const __grid_constant__ CUtensorMap map, int stages, int stage_bytes, int spans, int dchunks, int bw, int bh, int ntiles, int touch, float *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[MAXST], empty[MAXST];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], blockDim.x / 32 - 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (warp == 0) {
        if (lane != 0) return;
        int it = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % stages, k = it / stages;
            if (k > 0) mbar_wait(&empty[s], (k & 1) ^ 1);
            int r = t; const int sp = r % spans; r /= spans; const int dk = r % dchunks; const int b = r / dchunks;
            mbar_expect_tx(&full[s], stage_bytes);
            tma_load_3d(smem + (size_t)s * stage_bytes, &map, sp * bw, dk * bh, b, &full[s]);
        }
        return;
    }
    float acc = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % stages, k = it / stages;
        mbar_wait(&full[s], k & 1);
        if (touch) {
            const float4 *p = reinterpret_cast<const float4 *>(smem + (size_t)s * stage_bytes);
            for (int i = threadIdx.x - 32; i < stage_bytes / 16; i += blockDim.x - 32) { float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 123.456f) *sink = acc;
}

typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int HW = 540 * 960, D = 192, B = 8;
    const size_t n = (size_t)B * D * HW;
    float *buf, *sink; cudaMalloc(&buf, n * 4); cudaMalloc(&sink, 4); cudaMemset(buf, 0, n * 4);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    cudaFuncSetAttribute(tma_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const double gb = n * 4 / 1e9;
    struct Cfg { int bw, bh, cps, warps, touch; };
    std::vector<Cfg> cfgs;
    for (int touch : {0, 1}) for (int cps : {1, 2}) for (int bw : {32, 64, 128, 256}) for (int bh : {192, 96, 48, 24}) cfgs.push_back({bw, bh, cps, 9, touch});
    for (auto c : cfgs) {
        const int stage_bytes = c.bw * c.bh * 4;
        int stages = (200 * 1024 / c.cps) / stage_bytes; if (stages > MAXST) stages = MAXST; if (stages < 2) continue;
        CUtensorMap map;
        cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)D, (cuuint64_t)B}, gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 4 * D};
        cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, 1}, es[3] = {1, 1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
        const int spans = HW / c.bw, dchunks = D / c.bh, ntiles = spans * dchunks * B;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        std::vector<float> ts;
        for (int i = 0; i < 8; ++i) {
            cudaEventRecord(a);
            tma_read<<<148 * c.cps, 32 * c.warps, (size_t)stages * stage_bytes>>>(map, stages, stage_bytes, spans, dchunks, c.bw, c.bh, ntiles, c.touch, sink);
            cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (i >= 3) ts.push_back(ms);
        }
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        std::sort(ts.begin(), ts.end());
        printf("tma_read box=%3dx%3d stages=%2d cta/sm=%d touch=%d  %7.4f ms %8.1f GB/s\n", c.bw, c.bh, stages, c.cps, c.touch, ts[ts.size() / 2], gb / ts[ts.size() / 2] * 1e3);
        fflush(stdout);
    }
    return 0;
}
