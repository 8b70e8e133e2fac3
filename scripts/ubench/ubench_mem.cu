// Micro-benchmarks of HBM access patterns on B200 (not product code): which write / read structure reaches the
// plain fill / reduce rate?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mem ubench_mem.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void stg_cs(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_def(float4 *p, float4 v) { *p = v; }

// A: sequential grid-stride fill
template <bool CS>
__global__ void fill_seq(float4 *out, int64_t n4) {
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        if (CS) stg_cs(out + i, v); else stg_def(out + i, v);
    }
}
// B: volume pattern. out [NP planes][HW]; CTA = (span of SQ quads, group of planes); thread (q, slot) writes planes slot, slot+SLOTS...
// each plane-write of a warp covers min(SQ,32)*16 contiguous bytes.
template <bool CS>
__global__ void fill_planes(float *out, int HW, int planes_per_cta, int SQ) {
    const int q = threadIdx.x % SQ, slot = threadIdx.x / SQ, slots = blockDim.x / SQ;
    const int p = (blockIdx.x * SQ + q) * 4;
    if (p >= HW) return;
    const int64_t pl0 = (int64_t)blockIdx.y * planes_per_cta;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int k = slot; k < planes_per_cta; k += slots) {
        float4 *o = reinterpret_cast<float4 *>(out + (pl0 + k) * HW + p);
        if (CS) stg_cs(o, v); else stg_def(o, v);
    }
}
// C: like B, but a thread walks consecutive planes (d inner) — the concat kernel's order: thread = (q, channel), loop over D planes
template <bool CS>
__global__ void fill_planes_dinner(float *out, int HW, int D, int chans_per_cta, int SQ) {
    const int q = threadIdx.x % SQ, slot = threadIdx.x / SQ, slots = blockDim.x / SQ;
    const int p = (blockIdx.x * SQ + q) * 4;
    if (p >= HW) return;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int c = blockIdx.y * chans_per_cta + slot; c < (blockIdx.y + 1) * chans_per_cta; c += slots) {
        float *o = out + ((int64_t)c * D) * HW + p;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            if (CS) stg_cs(reinterpret_cast<float4 *>(o + (int64_t)d * HW), v); else stg_def(reinterpret_cast<float4 *>(o + (int64_t)d * HW), v);
        }
    }
}
// R: sequential read-reduce
__global__ void read_seq(const float4 *in, int64_t n4, float *sink) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(in + i));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}
// R2: sequential read, 8 independent loads per thread per iteration
__global__ void read_seq_u8(const float4 *in, int64_t n4, float *sink) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w) : "l"(in + i + k * stride));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    for (; i < n4; i += stride) { float4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) *sink = acc;
}


// P: persistent CTAs: tile = (plane group of `ppt` planes, span of 128 px); thread (q, slot) stores planes slot, slot+nslots, ...
__global__ void fill_persistent(float *out, int HW, int ppt, int ntiles, int spans) {
    const int q = threadIdx.x & 31, slot = threadIdx.x >> 5, nslots = blockDim.x >> 5;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int sp = t % spans, pg = t / spans;
        const int p = (sp * 32 + q) * 4;
        if (p >= HW) continue;
        for (int k = slot; k < ppt; k += nslots)
            stg_cs(reinterpret_cast<float4 *>(out + ((int64_t)pg * ppt + k) * HW + p), v);
    }
}


// P2: persistent with a dynamic (atomic) tile counter
__global__ void fill_persistent_dyn(float *out, int HW, int ppt, int ntiles, int spans, int *counter) {
    const int q = threadIdx.x & 31, slot = threadIdx.x >> 5, nslots = blockDim.x >> 5;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    __shared__ int st;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) st = atomicAdd(counter, 1);
        __syncthreads();
        const int t = st;
        if (t >= ntiles) break;
        const int sp = t % spans, pg = t / spans;
        const int p = (sp * 32 + q) * 4;
        if (p >= HW) continue;
        for (int k = slot; k < ppt; k += nslots)
            stg_cs(reinterpret_cast<float4 *>(out + ((int64_t)pg * ppt + k) * HW + p), v);
    }
}
// P3: each CTA does `n` consecutive tiles (grid = ntiles / n): CTA lifetime sweep
__global__ void fill_chunked(float *out, int HW, int ppt, int ntiles, int spans, int n) {
    const int q = threadIdx.x & 31, slot = threadIdx.x >> 5, nslots = blockDim.x >> 5;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int i = 0; i < n; ++i) {
        const int t = blockIdx.x * n + i;
        if (t >= ntiles) break;
        const int sp = t % spans, pg = t / spans;
        const int p = (sp * 32 + q) * 4;
        if (p >= HW) continue;
        for (int k = slot; k < ppt; k += nslots)
            stg_cs(reinterpret_cast<float4 *>(out + ((int64_t)pg * ppt + k) * HW + p), v);
    }
}
// P4: like P3 but tile i of the CTA is blockIdx.x + i * gridDim.x (strided, persistent-like order, short-lived CTAs)
__global__ void fill_strided(float *out, int HW, int ppt, int ntiles, int spans, int n) {
    const int q = threadIdx.x & 31, slot = threadIdx.x >> 5, nslots = blockDim.x >> 5;
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int i = 0; i < n; ++i) {
        const int t = blockIdx.x + i * gridDim.x;
        if (t >= ntiles) break;
        const int sp = t % spans, pg = t / spans;
        const int p = (sp * 32 + q) * 4;
        if (p >= HW) continue;
        for (int k = slot; k < ppt; k += nslots)
            stg_cs(reinterpret_cast<float4 *>(out + ((int64_t)pg * ppt + k) * HW + p), v);
    }
}

template <typename F>
float time_ms(F f, int iters = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    std::vector<float> ts;
    for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); ts.push_back(ms); }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main() {
    const int HW = 135 * 240, D = 48, C = 64, B = 8;
    const int64_t planes = (int64_t)B * C * D;
    const int64_t n = planes * HW;
    float *buf, *sink; CK(cudaMalloc(&buf, n * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, n * 4));
    const double gb = n * 4 / 1e9;
    auto rep = [&](const char *name, float ms) { printf("%-40s %8.4f ms  %8.1f GB/s\n", name, ms, gb / ms * 1e3); fflush(stdout); };
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        char nm[64];
        snprintf(nm, 64, "fill_seq default grid=%d", g); rep(nm, time_ms([&] { fill_seq<false><<<g, 256>>>((float4 *)buf, n / 4); }));
        snprintf(nm, 64, "fill_seq .cs     grid=%d", g); rep(nm, time_ms([&] { fill_seq<true><<<g, 256>>>((float4 *)buf, n / 4); }));
    }
    rep("cudaMemsetAsync", time_ms([&] { cudaMemsetAsync(buf, 0, n * 4); }));
    {
        int *ctr; cudaMalloc(&ctr, 4);
        const int ppt = 48, spans = (HW / 4 + 31) / 32; const int ntiles = spans * (int)(planes / ppt);
        for (int thr : {128, 256, 384}) for (int cps : {1, 2, 3}) for (int pp : {48, 192, 768}) { char nm[64]; snprintf(nm, 64, "fill_persistent_dyn thr=%d cta/sm=%d ppt=%d", thr, cps, pp);
            const int nt = spans * (int)(planes / pp);
            rep(nm, time_ms([&] { cudaMemsetAsync(ctr, 0, 4); fill_persistent_dyn<<<148 * cps, thr>>>(buf, HW, pp, nt, spans, ctr); })); }
        for (int n : {1, 2, 4, 8, 16, 32, 64}) { char nm[64]; snprintf(nm, 64, "fill_chunked n=%d", n);
            rep(nm, time_ms([&] { fill_chunked<<<(ntiles + n - 1) / n, 256>>>(buf, HW, ppt, ntiles, spans, n); })); }
        for (int n : {2, 8, 32}) { char nm[64]; snprintf(nm, 64, "fill_strided n=%d", n);
            rep(nm, time_ms([&] { fill_strided<<<(ntiles + n - 1) / n, 256>>>(buf, HW, ppt, ntiles, spans, n); })); }
    }
    for (int thr : {256}) for (int cps : {4}) for (int ppt : {48}) {
        if (thr * cps > 2048) continue;
        const int spans = (HW / 4 + 31) / 32; const int ntiles = spans * (int)(planes / ppt);
        char nm[64]; snprintf(nm, 64, "fill_persistent thr=%d cta/sm=%d ppt=%d", thr, cps, ppt);
        rep(nm, time_ms([&] { fill_persistent<<<148 * cps, thr>>>(buf, HW, ppt, ntiles, spans); }));
    }
    for (int SQ : {32}) {
        for (int ppc : {48, 96, 384, 768}) {
            dim3 grid((HW / 4 + SQ - 1) / SQ, (unsigned)(planes / ppc));
            char nm[64];
            snprintf(nm, 64, "fill_planes cs SQ=%d ppc=%d", SQ, ppc); rep(nm, time_ms([&] { fill_planes<true><<<grid, 256>>>(buf, HW, ppc, SQ); }));
            snprintf(nm, 64, "fill_planes df SQ=%d ppc=%d", SQ, ppc); rep(nm, time_ms([&] { fill_planes<false><<<grid, 256>>>(buf, HW, ppc, SQ); }));
        }
    }
    for (int SQ : {32, 64}) {
        for (int cpc : {8, 16}) {
            dim3 grid((HW / 4 + SQ - 1) / SQ, (unsigned)(B * C / cpc));
            char nm[64];
            snprintf(nm, 64, "fill_dinner cs SQ=%d cpc=%d", SQ, cpc); rep(nm, time_ms([&] { fill_planes_dinner<true><<<grid, 256>>>(buf, HW, D, cpc, SQ); }));
            snprintf(nm, 64, "fill_dinner df SQ=%d cpc=%d", SQ, cpc); rep(nm, time_ms([&] { fill_planes_dinner<false><<<grid, 256>>>(buf, HW, D, cpc, SQ); }));
        }
    }
    for (int g : {148 * 8, 148 * 16, 148 * 32}) {
        char nm[64];
        snprintf(nm, 64, "read_seq grid=%d", g); rep(nm, time_ms([&] { read_seq<<<g, 256>>>((const float4 *)buf, n / 4, sink); }));
        snprintf(nm, 64, "read_seq_u8 grid=%d", g); rep(nm, time_ms([&] { read_seq_u8<<<g, 256>>>((const float4 *)buf, n / 4, sink); }));
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
