// concat_stream.cu — TMA-fed, dynamically scheduled streaming producer of the (attention-weighted, DDIM-filtered)
// concatenation volume (a3 + a4 + a9) for sm_100a.  This is the kernel behind dv_concat_volume_weighted_f32.
//
// The op writes 398 MB per pair from 21 MB of inputs, so it lives or dies by the HBM *write* stream.  Two
// measurements on B200 (scripts/ubench/ubench_mem.cu, gpurun_out/ubench_mem*.log) shaped the design:
//   1. a write-only stream reaches the memset rate (7.3 TB/s) only while the tiles being written at any moment
//      stay a compact, in-order window of the output: statically strided persistent CTAs drift apart and fall to
//      5.3 TB/s, an atomic in-order tile counter with 1-2 CTAs per SM holds 7.0-7.4 TB/s;
//   2. with plain loads the stores of a warp wait on two dependent L2 round trips (factor quads, then the tgt
//      window) and the SM cannot keep ~1 store per 20 cycles in flight.
// So: one persistent CTA per SM; a producer warp takes tiles from an atomic counter IN ORDER and lands the tile's
// inputs in shared memory with tensor-map TMA loads (cp.async.bulk.tensor -> SASS UTMALDG), 2 stages deep; twelve
// consumer warps read them with LDS.128 (30-cycle latency instead of ~700) and do nothing but FMUL + 128-bit
// streaming stores.
//
// Tile = (batch b, channel group of CGT channels, chunk of <= 48 disparities, span of 128 px of the flattened
// plane).  Stage layout (floats):  w[48][128] | n[48][128] | features[CGT][176].
//   left-half channels : features row = ref[c][p0 .. p0+128)
//   right-half channels: features row = tgt[c][p0-48(k+1) .. p0-48k+128), k = disparity-chunk index; consumer thread
//                        (quad q, group ds of 4 disparities, d0 = 48k + 4ds) needs tgt[p-d0-j .. +3], j < 4, which are
//                        compile-time permutations of two ALIGNED quads of that row (cur at 48+4q-4ds, prev 4 before).
// Whatever the window holds left of the row start (previous plane / TMA zero fill) is only ever selected away
// (x < d -> 0), exactly as in the reference (SceneFlow/models/submodule.py:180-191, KITTI12/models/submodule.py:86-97).
#include "common.cuh"

namespace dv {

constexpr int kCsSpan = 128;       // px per tile
constexpr int kCsDC = 48;          // disparities per tile
constexpr int kCsWin = kCsSpan + kCsDC;   // tgt window per right-half channel
constexpr int kCsConsumerWarps = kCsDC / 4;   // 12: one warp per group of 4 disparities
constexpr int kCsThreads = 32 * (kCsConsumerWarps + 1);
constexpr int kCsSlots = 256;      // concurrent launches per device that may share the counter pool

// Tile counters {next, done}.  The caller normally owns them (the `tile_counters` argument of the C-ABI: 2 zeroed ints in
// device memory, one pair per stream that may run this kernel concurrently; the kernel re-arms them to zero before it
// exits).  With tile_counters == NULL the launch draws a pair from this library-owned pool round-robin: at most
// kCsSlots launches of this kernel may then overlap on a device, and a captured CUDA graph keeps its slot forever.
static __device__ int g_cs_ctr[kCsSlots][2];
static std::atomic<unsigned> g_cs_slot{0};

template <int CGT>
struct CsStage {
    static constexpr int kFactor = kCsDC * kCsSpan;           // floats per factor tile
    static constexpr int kFeat = CGT * kCsWin;                // floats reserved for the feature rows
    static constexpr int kFloats = 2 * kFactor + ((kFeat + 31) / 32) * 32;   // keep every stage 128-byte aligned
};

template <int CGT, bool HAS_W, bool HAS_N, int kCsStages, int MINB, typename OutT>
__global__ void __launch_bounds__(kCsThreads, MINB)
concat_stream_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_n,
                     const __grid_constant__ CUtensorMap map_ref, const __grid_constant__ CUtensorMap map_tgt,
                     OutT *__restrict__ out, int C, int HW, int W, int D, int mask_left, int spans, int ndchunks,
                     int ntiles, int *ctr_arg, int slot, int sync_tiles) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full_bar[kCsStages], empty_bar[kCsStages];
    __shared__ int tile_id[kCsStages];
    constexpr int kStageFloats = CsStage<CGT>::kFloats;
    constexpr int kFactor = CsStage<CGT>::kFactor;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *const ctr = ctr_arg ? ctr_arg : g_cs_ctr[slot];
    const int ngroups = (2 * C) / CGT;
    const int dcbox = D < kCsDC ? D : kCsDC;     // rows of the factor boxes

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kCsStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kCsConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kCsConsumerWarps) {
        // ---------------- producer warp (lane 0 does the work)
        if (lane != 0) return;
        for (int it = 0;; ++it) {
            const int s = it % kCsStages, k = it / kCsStages;
            if (k > 0) mbar_wait(&empty_bar[s], (k & 1) ^ 1);
            const int t = atomicAdd(&ctr[0], 1);
            if (t >= ntiles) {
                tile_id[s] = -1;
                mbar_arrive(&full_bar[s]);
                // the last CTA to run dry re-arms the counter pair for the next launch that draws this slot
                if (atomicAdd(&ctr[1], 1) == static_cast<int>(gridDim.x) - 1) {
                    ctr[0] = 0;
                    ctr[1] = 0;
                    __threadfence();
                }
                return;
            }
            tile_id[s] = t;
            // t = ((b * ngroups + g) * ndchunks + dk) * spans + sp   (span fastest: neighbours in the output are
            // written at the same time)
            int r = t;
            const int sp = r % spans; r /= spans;
            const int dk = r % ndchunks; r /= ndchunks;
            const int g = r % ngroups;
            const int b = r / ngroups;
            const int p0 = sp * kCsSpan, c0 = g * CGT;
            const bool left = c0 < C;
            float *st = smem + s * kStageFloats;
            uint32_t bytes = (left ? CGT * kCsSpan : CGT * kCsWin) * 4u;
            if (HAS_W) bytes += static_cast<uint32_t>(dcbox) * kCsSpan * 4u;
            if (HAS_N) bytes += static_cast<uint32_t>(dcbox) * kCsSpan * 4u;
            mbar_expect_tx(&full_bar[s], bytes);
            if (HAS_W) tma_load_3d(st, &map_w, p0, dk * kCsDC, b, &full_bar[s]);
            if (HAS_N) tma_load_3d(st + kFactor, &map_n, p0, dk * kCsDC, b, &full_bar[s]);
            if (left)
                tma_load_3d(st + 2 * kFactor, &map_ref, p0, c0, b, &full_bar[s]);
            else
                tma_load_3d(st + 2 * kFactor, &map_tgt, p0 - kCsDC * (dk + 1), c0 - C, b, &full_bar[s]);
        }
    }

    // ---------------- consumer warps: thread = (quad q, disparity group ds)
    const int q = lane, ds = warp;
    for (int it = 0;; ++it) {
        const int s = it % kCsStages, k = it / kCsStages;
        mbar_wait(&full_bar[s], k & 1);
        const int t = tile_id[s];
        if (t < 0) return;
        int r = t;
        const int sp = r % spans; r /= spans;
        const int dk = r % ndchunks; r /= ndchunks;
        const int g = r % ngroups;
        const int b = r / ngroups;
        const int p = sp * kCsSpan + 4 * q;
        const int c0 = g * CGT;
        const int d0 = dk * kCsDC + 4 * ds;
        if (p < HW && d0 < D) {
            const float *st = smem + s * kStageFloats;
            float4 w4[4], n4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (HAS_W) w4[j] = *reinterpret_cast<const float4 *>(st + (4 * ds + j) * kCsSpan + 4 * q);
                if (HAS_N) n4[j] = *reinterpret_cast<const float4 *>(st + kFactor + (4 * ds + j) * kCsSpan + 4 * q);
            }
            int xs[4];
            xs[0] = p % W;
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                xs[i] = xs[i - 1] + 1;
                if (xs[i] >= W) xs[i] -= W;
            }
            const bool left = c0 < C;
            const bool masked = !left || mask_left;
            const float *sf = st + 2 * kFactor;
            OutT *op = out + ((static_cast<int64_t>(b) * 2 * C + c0) * D + d0) * HW + p;
            const int64_t cstride = static_cast<int64_t>(D) * HW;
#pragma unroll 2
            for (int kc = 0; kc < CGT; ++kc) {
                float4 cur, prev;
                if (left) {
                    cur = *reinterpret_cast<const float4 *>(sf + kc * kCsSpan + 4 * q);
                    prev = cur;
                } else {
                    const float *row = sf + kc * kCsWin + kCsDC + 4 * q - 4 * ds;
                    cur = *reinterpret_cast<const float4 *>(row);
                    prev = *reinterpret_cast<const float4 *>(row - 4);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = d0 + j;
                    if (d < D) {
                        float4 o;
                        if (left || j == 0) o = cur;
                        else if (j == 1) o = make_float4(prev.w, cur.x, cur.y, cur.z);
                        else if (j == 2) o = make_float4(prev.z, prev.w, cur.x, cur.y);
                        else o = make_float4(prev.y, prev.z, prev.w, cur.x);
                        if (masked) {
                            o.x = xs[0] >= d ? o.x : 0.0f; o.y = xs[1] >= d ? o.y : 0.0f;
                            o.z = xs[2] >= d ? o.z : 0.0f; o.w = xs[3] >= d ? o.w : 0.0f;
                        }
                        if (HAS_W) { o.x *= w4[j].x; o.y *= w4[j].y; o.z *= w4[j].z; o.w *= w4[j].w; }
                        if (HAS_N) { o.x *= n4[j].x; o.y *= n4[j].y; o.z *= n4[j].z; o.w *= n4[j].w; }
                        store4_cs(op + kc * cstride + static_cast<int64_t>(j) * HW, o);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (sync_tiles) asm volatile("bar.sync 1, 384;" ::: "memory");   // keep the 12 store streams of the CTA in step
    }
}

template <int CGT, bool HAS_W, bool HAS_N, int STAGES, int MINB, typename OutT>
static int launch_cs2(const CUtensorMap &mw, const CUtensorMap &mn, const CUtensorMap &mr, const CUtensorMap &mt, OutT *out,
                      int B, int C, int HW, int W, int D, int mask_left, int *ctr, cudaStream_t st) {
    const size_t smem = sizeof(float) * STAGES * CsStage<CGT>::kFloats;
    auto kern = concat_stream_kernel<CGT, HAS_W, HAS_N, STAGES, MINB, OutT>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    const int spans = (HW + kCsSpan - 1) / kCsSpan;
    const int ndchunks = (D + kCsDC - 1) / kCsDC;
    const int64_t ntiles = static_cast<int64_t>(B) * ((2 * C) / CGT) * ndchunks * spans;
    if (ntiles > INT32_MAX - 4 * num_sms()) return DV_ERR_UNSUPPORTED;
    const int grid = static_cast<int>(ntiles < MINB * num_sms() ? ntiles : MINB * num_sms());
    const int slot = static_cast<int>(g_cs_slot.fetch_add(1, std::memory_order_relaxed) % kCsSlots);
    kern<<<grid, kCsThreads, smem, st>>>(mw, mn, mr, mt, out, C, HW, W, D, mask_left, spans, ndchunks,
                                        static_cast<int>(ntiles), ctr, slot, DV_TUNE("DV_CS_BAR", 1));
    return finish_launch();
}
template <int CGT, bool HAS_W, bool HAS_N>
static int launch_cs(const CUtensorMap &mw, const CUtensorMap &mn, const CUtensorMap &mr, const CUtensorMap &mt, float *out,
                     int B, int C, int HW, int W, int D, int mask_left, int *ctr, cudaStream_t st) {
    const int v = DV_TUNE("DV_CS_SHAPE", 21);   // stages*10 + CTAs per SM
    if (v == 22) return launch_cs2<CGT, HAS_W, HAS_N, 2, 2, float>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st);
    if (v == 31) return launch_cs2<CGT, HAS_W, HAS_N, 3, 1, float>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st);
    return launch_cs2<CGT, HAS_W, HAS_N, 2, 1, float>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st);
}

// Returns DV_ERR_UNSUPPORTED when the tensor maps cannot be built (the caller then takes the LDG kernel).
int launch_concat_stream(const float *ref, const float *tgt, float *out, int B, int C, int HW, int W, int D, int mask_left,
                         const float *wts, const float *nf, int *ctr, cudaStream_t st) {
    CUtensorMap mw, mn, mr, mt;
    const uint64_t fdims[3] = {static_cast<uint64_t>(HW), static_cast<uint64_t>(D), static_cast<uint64_t>(B)};
    const uint32_t fbox[3] = {kCsSpan, static_cast<uint32_t>(D < kCsDC ? D : kCsDC), 1u};
    const uint64_t cdims[3] = {static_cast<uint64_t>(HW), static_cast<uint64_t>(C), static_cast<uint64_t>(B)};
    int cgt = DV_TUNE("DV_CS_CGT", 8);
    while (cgt > 1 && C % cgt != 0) cgt /= 2;
    const uint32_t lbox[3] = {kCsSpan, static_cast<uint32_t>(cgt), 1u};
    const uint32_t rbox[3] = {kCsWin, static_cast<uint32_t>(cgt), 1u};
    // a NULL factor still needs a valid map object for the launch; alias the ref map (never dereferenced)
    if (!make_tensor_map_f32(&mr, ref, 3, cdims, lbox) || !make_tensor_map_f32(&mt, tgt, 3, cdims, rbox))
        return DV_ERR_UNSUPPORTED;
    mw = mr;
    mn = mr;
    if (wts && !make_tensor_map_f32(&mw, wts, 3, fdims, fbox)) return DV_ERR_UNSUPPORTED;
    if (nf && !make_tensor_map_f32(&mn, nf, 3, fdims, fbox)) return DV_ERR_UNSUPPORTED;
#define DV_CS(CG)                                                                                              \
    (wts && nf ? launch_cs<CG, true, true>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st)                  \
     : wts     ? launch_cs<CG, true, false>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st)                 \
     : nf      ? launch_cs<CG, false, true>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st)                 \
               : launch_cs<CG, false, false>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st))
    switch (cgt) {
        case 16: return DV_CS(16);
        case 8: return DV_CS(8);
        case 4: return DV_CS(4);
        case 2: return DV_CS(2);
        default: return DV_CS(1);
    }
#undef DV_CS
}

// bf16 volume: the same producer with one rounding at the store (channel groups of 8 or 4 only)
int launch_concat_stream_bf16(const float *ref, const float *tgt, __nv_bfloat16 *out, int B, int C, int HW, int W, int D,
                              int mask_left, const float *wts, const float *nf, int *ctr, cudaStream_t st) {
    CUtensorMap mw, mn, mr, mt;
    const uint64_t fdims[3] = {static_cast<uint64_t>(HW), static_cast<uint64_t>(D), static_cast<uint64_t>(B)};
    const uint32_t fbox[3] = {kCsSpan, static_cast<uint32_t>(D < kCsDC ? D : kCsDC), 1u};
    const uint64_t cdims[3] = {static_cast<uint64_t>(HW), static_cast<uint64_t>(C), static_cast<uint64_t>(B)};
    int cgt = 8;
    while (cgt > 1 && C % cgt != 0) cgt /= 2;
    if (cgt != 8 && cgt != 4) return DV_ERR_UNSUPPORTED;
    const uint32_t lbox[3] = {kCsSpan, static_cast<uint32_t>(cgt), 1u};
    const uint32_t rbox[3] = {kCsWin, static_cast<uint32_t>(cgt), 1u};
    if (!make_tensor_map_f32(&mr, ref, 3, cdims, lbox) || !make_tensor_map_f32(&mt, tgt, 3, cdims, rbox))
        return DV_ERR_UNSUPPORTED;
    mw = mr;
    mn = mr;
    if (wts && !make_tensor_map_f32(&mw, wts, 3, fdims, fbox)) return DV_ERR_UNSUPPORTED;
    if (nf && !make_tensor_map_f32(&mn, nf, 3, fdims, fbox)) return DV_ERR_UNSUPPORTED;
#define DV_CSB(CG)                                                                                                          \
    (wts && nf ? launch_cs2<CG, true, true, 2, 1, __nv_bfloat16>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st)        \
     : wts     ? launch_cs2<CG, true, false, 2, 1, __nv_bfloat16>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st)       \
     : nf      ? launch_cs2<CG, false, true, 2, 1, __nv_bfloat16>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st)       \
               : launch_cs2<CG, false, false, 2, 1, __nv_bfloat16>(mw, mn, mr, mt, out, B, C, HW, W, D, mask_left, ctr, st))
    return cgt == 8 ? DV_CSB(8) : DV_CSB(4);
#undef DV_CSB
}

}  // namespace dv

extern "C" int dv_concat_volume_weighted_bf16(const float *ref, const float *tgt, void *out, int64_t B, int64_t C, int64_t H,
                                              int64_t W, int64_t D, int mask_left, const float *att_weights, const float *n,
                                              void *tile_counters, void *stream) {
    using namespace dv;
    if (!ref || !tgt || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || 2 * C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    auto ok16 = [](const void *p) { return !p || aligned16(p); };
    if (!((HW % 4 == 0) && W >= 4 && aligned16(ref) && aligned16(tgt) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0 &&
          ok16(att_weights) && ok16(n)))
        return DV_ERR_MISALIGNED;
    return launch_concat_stream_bf16(ref, tgt, static_cast<__nv_bfloat16 *>(out), static_cast<int>(B), static_cast<int>(C),
                                     static_cast<int>(HW), static_cast<int>(W), static_cast<int>(D), mask_left, att_weights, n,
                                     static_cast<int *>(tile_counters), static_cast<cudaStream_t>(stream));
}
