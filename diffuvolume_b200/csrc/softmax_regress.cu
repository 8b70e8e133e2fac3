// softmax_regress.cu — softmax over D + disparity regression, with the renewal-mask
// uncertainty, vote and ensemble accumulation fused (a6, a11, a13) for sm_100a.
//
// Replaces F.softmax(cost, dim=1) + disparity_regression (SceneFlow/models/submodule.py:173-177;
// call sites acv_ddim.py:269-270, pwcnet_ddim.py:483-484, igev_stereo_ddim.py:384-385) and the
// second full pass over the probability volume that ddim_sample makes for the renewal mask
// (acv_ddim.py:320-331: |disp-used| < 1  &  sum_d |disp-d| p[d] < 3).  The reference writes
// the 398 MB softmax, writes a 398 MB product, reads it back, then builds and reduces another
// [B,192,H,W] temporary; here the cost volume is read from HBM exactly once and only
// [B,H,W] maps are written.
//
// CTA = 32 quads of pixels x 8 disparity slices (256 threads).  A thread owns DPT
// disparities (d = j*8 + slice) of one pixel quad and keeps them in registers (DPT float4,
// all loads issued up front: 24 independent 128-bit loads per thread at D=192).  Max, sum,
// sum d*e and sum |disp-d|*e are combined across the 8 slices through shared memory
// (three block barriers); the warp of slice 0 writes the outputs.  exp is exp2 of a
// pre-scaled argument (one FFMA + MUFU.EX2 per element).
#include "common.cuh"

namespace dv {

constexpr int kSrLanes = 32;

template <int V>
struct alignas(V * 4) Vec {
    float v[V];
};

template <int V>
__device__ __forceinline__ Vec<V> load_vec(const float *p) {
    Vec<V> r;
    if constexpr (V == 4) {
        const float4 t = ldg_stream(reinterpret_cast<const float4 *>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else if constexpr (V == 2) {
        float2 t;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(t.x), "=f"(t.y) : "l"(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}
template <int V>
__device__ __forceinline__ void store_vec(float *p, const Vec<V> &r) {
    if constexpr (V == 4)
        *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    else if constexpr (V == 2)
        *reinterpret_cast<float2 *>(p) = make_float2(r.v[0], r.v[1]);
    else
        *p = r.v[0];
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    // volatile: under register pressure ptxas otherwise REMATERIALISES the exponentials in the uncertainty
    // pass instead of keeping them (2x MUFU + FFMA, seen in the ncu source page of round 1)
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.EX2; exp2(-inf) = +0
    return y;
}

// cost [B,D,HW];  pixel-vector index pv in [0, HW/V)
template <int DPT, int V, int SL, int MINB, bool FULLD>
__global__ void __launch_bounds__(SL * kSrLanes, MINB)
softmax_regress_kernel(const float *__restrict__ cost, int D, int HW, float *__restrict__ disp_out,
                       float *__restrict__ prob_out, const float *__restrict__ used, const float *__restrict__ disp_ext,
                       float *__restrict__ unc_out,
                       float *__restrict__ vote_out, float thr_dif, float thr_unc, float *__restrict__ ens_acc,
                       float ens_coef, int ens_init) {
    constexpr int kSrSlices = SL;   // disparity slices per CTA; thread (lane, slice) owns d = j*SL + slice
    __shared__ Vec<V> red[3][kSrSlices][kSrLanes];
    const int lane = threadIdx.x % kSrLanes;
    const int slice = threadIdx.x / kSrLanes;
    const int b = blockIdx.y;
    const int64_t pv = blockIdx.x * static_cast<int64_t>(kSrLanes) + lane;
    const bool live = pv * V < HW;
    const float *cp = cost + static_cast<int64_t>(b) * D * HW + pv * V;
    constexpr float kLog2e = 1.4426950408889634f;

    Vec<V> x[DPT];
    {
        const float *lp = cp + static_cast<int64_t>(slice) * HW;
        const int64_t step = static_cast<int64_t>(kSrSlices) * HW;
#pragma unroll
        for (int j = 0; j < DPT; ++j, lp += step) {
            const bool ok = FULLD ? live : (live && (j * kSrSlices + slice < D));   // FULLD: D == SL * DPT
            if (ok) {
                x[j] = load_vec<V>(lp);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) x[j].v[i] = -INFINITY;
            }
        }
    }
    // ---- max over D
    Vec<V> m;
#pragma unroll
    for (int i = 0; i < V; ++i) m.v[i] = -INFINITY;
#pragma unroll
    for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) m.v[i] = fmaxf(m.v[i], x[j].v[i]);
    red[0][slice][lane] = m;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kSrSlices; ++s) {
        const Vec<V> o = red[0][s][lane];
#pragma unroll
        for (int i = 0; i < V; ++i) m.v[i] = fmaxf(m.v[i], o.v[i]);
    }
    // ---- e = exp(x - max); S = sum e; Wd = sum d * e
    Vec<V> mL, S, Wd;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        mL.v[i] = live ? m.v[i] * kLog2e : 0.0f;
        S.v[i] = 0.0f;
        Wd.v[i] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const float df = static_cast<float>(j * kSrSlices + slice);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float e = fast_exp2(fmaf(x[j].v[i], kLog2e, -mL.v[i]));  // exp2(-inf) = 0 for d >= D
            x[j].v[i] = e;
            S.v[i] += e;
            Wd.v[i] = fmaf(df, e, Wd.v[i]);
        }
    }
    red[1][slice][lane] = S;
    red[2][slice][lane] = Wd;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
        S.v[i] = 0.0f;
        Wd.v[i] = 0.0f;
    }
#pragma unroll
    for (int s = 0; s < kSrSlices; ++s) {
        const Vec<V> a = red[1][s][lane], w = red[2][s][lane];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            S.v[i] += a.v[i];
            Wd.v[i] += w.v[i];
        }
    }
    Vec<V> rS, disp;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        rS.v[i] = 1.0f / S.v[i];
        disp.v[i] = Wd.v[i] * rS.v[i];
    }
    if (prob_out && live) {
        float *pp = prob_out + static_cast<int64_t>(b) * D * HW + pv * V;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = j * kSrSlices + slice;
            if (FULLD || d < D) {
                Vec<V> p;
#pragma unroll
                for (int i = 0; i < V; ++i) p.v[i] = x[j].v[i] * rS.v[i];
                store_vec<V>(pp + static_cast<int64_t>(d) * HW, p);
            }
        }
    }
    const bool need_unc = (unc_out != nullptr) || (vote_out != nullptr);
    Vec<V> U;
#pragma unroll
    for (int i = 0; i < V; ++i) U.v[i] = 0.0f;
    // the disparity the uncertainty / vote are taken around: the regression itself, or an external map (PWCNet's refined
    // disparity against the pre-refinement distribution, pwcnet_ddim.py:553-570)
    Vec<V> dq = disp;
    if (need_unc && disp_ext && live) dq = load_vec<V>(disp_ext + static_cast<int64_t>(b) * HW + pv * V);
    if (need_unc) {  // uniform across the block
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const float df = static_cast<float>(j * kSrSlices + slice);
#pragma unroll
            for (int i = 0; i < V; ++i) U.v[i] = fmaf(fabsf(dq.v[i] - df), x[j].v[i], U.v[i]);
        }
        __syncthreads();  // red[0] is free again only after everyone has read the max
        red[0][slice][lane] = U;
        __syncthreads();
        if (slice == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) U.v[i] = 0.0f;
#pragma unroll
            for (int s = 0; s < kSrSlices; ++s) {
                const Vec<V> a = red[0][s][lane];
#pragma unroll
                for (int i = 0; i < V; ++i) U.v[i] += a.v[i];
            }
#pragma unroll
            for (int i = 0; i < V; ++i) U.v[i] *= rS.v[i];
        }
    }
    if (slice == 0 && live) {
        const int64_t o = static_cast<int64_t>(b) * HW + pv * V;
        if (disp_out) store_vec<V>(disp_out + o, disp);
        if (unc_out) store_vec<V>(unc_out + o, U);
        if (vote_out) {
            Vec<V> vt;
            if (used) {
                const Vec<V> u0 = load_vec<V>(used + o);
#pragma unroll
                for (int i = 0; i < V; ++i)
                    vt.v[i] = (fabsf(dq.v[i] - u0.v[i]) < thr_dif && U.v[i] < thr_unc) ? 1.0f : 0.0f;
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) vt.v[i] = U.v[i] < thr_unc ? 1.0f : 0.0f;
            }
            store_vec<V>(vote_out + o, vt);
        }
        if (ens_acc) {
            Vec<V> a;
            if (ens_init) {
#pragma unroll
                for (int i = 0; i < V; ++i) a.v[i] = 0.0f;
            } else {
                a = load_vec<V>(ens_acc + o);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) a.v[i] = fmaf(ens_coef, disp.v[i], a.v[i]);
            store_vec<V>(ens_acc + o, a);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Pipelined variant (the shipped path for 16-byte-aligned maps with D <= 192).
//
// A persistent CTA (2 per SM) walks tiles of [D x 64 px] of one batch item.  A dedicated producer warp lands each
// tile in shared memory with ONE tensor-map TMA load (cp.async.bulk.tensor.3d -> SASS UTMALDG, mbarrier
// transaction count), STAGES tiles deep, so HBM requests stay in flight while the 8 consumer warps do the
// arithmetic of earlier tiles — the register-resident kernel above alternates between "all loads" and "all math"
// per warp and reached 4.5 TB/s; per-row 1-D bulk copies (UBLKCP, 256-512 B each) were measured at 2.3 TB/s and
// dropped.  Consumers copy their slice of the tile to registers (conflict-free LDS.128: a warp reads two adjacent
// rows = 512 contiguous bytes) and release the stage at once.  Thread (quad q of 16, slice ds of NDS) owns
// d = NDS j + ds; the constant NDS j folds into FFMA/FADD immediates, the per-thread ds enters once per tile.
// Cross-slice reductions: one shuffle + (NDS/2)-way shared memory combine behind a named barrier of the consumer
// warps (the producer warp never joins it).
constexpr int kSrSlots = 256;
// {next, done} tile counters: caller-owned through the C-ABI's `tile_counters` argument, else a pair of this pool (see
// concat_stream.cu for the contract)
static __device__ int g_sr_ctr[kSrSlots][2];
static std::atomic<unsigned> g_sr_slot{0};

template <int NJ, int NDS, int STAGES, int SPAN, int MINB, bool DYN, bool FULLD>
__global__ void __launch_bounds__(SPAN / 4 * NDS + 32, MINB)
softmax_regress_tma_kernel(const __grid_constant__ CUtensorMap tmap, int D, int HW, int spans_per_b, int ntiles, int *ctr_arg, int slot,
                           float *__restrict__ disp_out, float *__restrict__ prob_out, const float *__restrict__ used,
                           const float *__restrict__ disp_ext, float *__restrict__ unc_out, float *__restrict__ vote_out,
                           float thr_dif, float thr_unc, float *__restrict__ ens_acc, float ens_coef, int ens_init) {
    constexpr int SQ = SPAN / 4, NCONS = SQ * NDS, NW = NCONS / 32;   // quads per span, consumer threads / warps
    extern __shared__ __align__(128) float smem[];              // [STAGES][D][SPAN]
    __shared__ float4 red[4][NW][SQ];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ int tile_id[STAGES];
    const int stage_floats = D * SPAN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *const ctr = ctr_arg ? ctr_arg : g_sr_ctr[slot];

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ---------------- producer warp (lane 0): one tensor-map TMA per tile, box [SPAN px, D rows, 1] of the
        // [HW, D, B] view (OOB px read as 0).  Tiles are handed out IN ORDER by an atomic counter (DYN) so that the
        // spans being read at any moment stay a compact window of the volume (DRAM page locality; the statically
        // strided assignment lets CTAs drift apart).
        if (lane != 0) return;
        for (int it = 0;; ++it) {
            const int s = it % STAGES, k = it / STAGES;
            if (k > 0) mbar_wait(&empty_bar[s], (k & 1) ^ 1);
            const int t = DYN ? atomicAdd(&ctr[0], 1) : static_cast<int>(blockIdx.x + it * gridDim.x);
            if (t >= ntiles) {
                tile_id[s] = -1;
                mbar_arrive(&full_bar[s]);
                if (DYN && atomicAdd(&ctr[1], 1) == static_cast<int>(gridDim.x) - 1) {
                    ctr[0] = 0;   // last CTA out re-arms the counter pair
                    ctr[1] = 0;
                    __threadfence();
                }
                return;
            }
            tile_id[s] = t;
            const int b = t / spans_per_b;
            const int p0 = (t - b * spans_per_b) * SPAN;
            mbar_expect_tx(&full_bar[s], static_cast<uint32_t>(D) * SPAN * 4u);
            tma_load_3d(smem + s * stage_floats, &tmap, p0, 0, b, &full_bar[s]);
        }
    }

    // ---------------- consumer warps
    const int q = threadIdx.x & (SQ - 1);
    const int ds = threadIdx.x / SQ;           // 0..NDS-1
    const float dsf = static_cast<float>(ds);
    constexpr float kLog2e = 1.4426950408889634f;
    const bool fin = ds == 0;                  // lanes 0..15 of warp 0 write the span's outputs
    for (int it = 0;; ++it) {
        const int s = it % STAGES, k = it / STAGES;
        mbar_wait(&full_bar[s], k & 1);
        const int t = tile_id[s];
        if (t < 0) return;
        const int b = t / spans_per_b;
        const int p0 = (t - b * spans_per_b) * SPAN;
        const int p = p0 + 4 * q;
        const bool live = p < HW;
        const int64_t o = static_cast<int64_t>(b) * HW + p;
        float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = make_float4(0.f, 0.f, 0.f, 0.f), e4 = u4;
        if (disp_ext && live) e4 = __ldg(reinterpret_cast<const float4 *>(disp_ext + o));   // every slice needs it
        if (fin && live) {   // issue the small map reads before waiting for the tile
            if (vote_out && used) u4 = ldg_stream(reinterpret_cast<const float4 *>(used + o));
            if (ens_acc && !ens_init) a4 = *reinterpret_cast<const float4 *>(ens_acc + o);
        }
        float4 x[NJ];
        {
            const float *st = smem + s * stage_floats + ds * SPAN + 4 * q;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                if (FULLD || NDS * j + ds < D)
                    x[j] = *reinterpret_cast<const float4 *>(st + NDS * j * SPAN);
                else
                    x[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);   // the tile now lives in registers

        // ---- max over D
        float4 m = x[0];
#pragma unroll
        for (int j = 1; j < NJ; ++j) {
            m.x = fmaxf(m.x, x[j].x); m.y = fmaxf(m.y, x[j].y); m.z = fmaxf(m.z, x[j].z); m.w = fmaxf(m.w, x[j].w);
        }
        if (SQ < 32) {
        m.x = fmaxf(m.x, __shfl_xor_sync(0xffffffffu, m.x, 16));
        m.y = fmaxf(m.y, __shfl_xor_sync(0xffffffffu, m.y, 16));
        m.z = fmaxf(m.z, __shfl_xor_sync(0xffffffffu, m.z, 16));
        m.w = fmaxf(m.w, __shfl_xor_sync(0xffffffffu, m.w, 16));
        }
        if (SQ == 32 || lane < SQ) red[0][warp][q] = m;
        asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float4 v = red[0][w][q];
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
        // ---- e = exp(x - max), S = sum e, Wd = sum d e   (d = 16 j + ds)
        const float4 mL = make_float4(m.x * kLog2e, m.y * kLog2e, m.z * kLog2e, m.w * kLog2e);
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Wj = S;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float dj = static_cast<float>(NDS * j);
            x[j].x = fast_exp2(fmaf(x[j].x, kLog2e, -mL.x)); S.x += x[j].x; Wj.x = fmaf(dj, x[j].x, Wj.x);
            x[j].y = fast_exp2(fmaf(x[j].y, kLog2e, -mL.y)); S.y += x[j].y; Wj.y = fmaf(dj, x[j].y, Wj.y);
            x[j].z = fast_exp2(fmaf(x[j].z, kLog2e, -mL.z)); S.z += x[j].z; Wj.z = fmaf(dj, x[j].z, Wj.z);
            x[j].w = fast_exp2(fmaf(x[j].w, kLog2e, -mL.w)); S.w += x[j].w; Wj.w = fmaf(dj, x[j].w, Wj.w);
        }
        float4 Wd = make_float4(fmaf(dsf, S.x, Wj.x), fmaf(dsf, S.y, Wj.y), fmaf(dsf, S.z, Wj.z), fmaf(dsf, S.w, Wj.w));
        if (SQ < 32) {
        S.x += __shfl_xor_sync(0xffffffffu, S.x, 16); S.y += __shfl_xor_sync(0xffffffffu, S.y, 16);
        S.z += __shfl_xor_sync(0xffffffffu, S.z, 16); S.w += __shfl_xor_sync(0xffffffffu, S.w, 16);
        Wd.x += __shfl_xor_sync(0xffffffffu, Wd.x, 16); Wd.y += __shfl_xor_sync(0xffffffffu, Wd.y, 16);
        Wd.z += __shfl_xor_sync(0xffffffffu, Wd.z, 16); Wd.w += __shfl_xor_sync(0xffffffffu, Wd.w, 16);
        }
        if (SQ == 32 || lane < SQ) {
            red[1][warp][q] = S;
            red[2][warp][q] = Wd;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
        S = make_float4(0.f, 0.f, 0.f, 0.f);
        Wd = S;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float4 a = red[1][w][q], c = red[2][w][q];
            S.x += a.x; S.y += a.y; S.z += a.z; S.w += a.w;
            Wd.x += c.x; Wd.y += c.y; Wd.z += c.z; Wd.w += c.w;
        }
        const float4 rS = make_float4(1.0f / S.x, 1.0f / S.y, 1.0f / S.z, 1.0f / S.w);
        const float4 disp = make_float4(Wd.x * rS.x, Wd.y * rS.y, Wd.z * rS.z, Wd.w * rS.w);
        if (prob_out && live) {
            float *pp = prob_out + static_cast<int64_t>(b) * D * HW + p + static_cast<int64_t>(ds) * HW;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (FULLD || NDS * j + ds < D)
                    *reinterpret_cast<float4 *>(pp + static_cast<int64_t>(NDS * j) * HW) =
                        make_float4(x[j].x * rS.x, x[j].y * rS.y, x[j].z * rS.z, x[j].w * rS.w);
        }
        float4 U = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 dq = disp_ext ? e4 : disp;   // the map the uncertainty / vote are taken around (see the register kernel)
        if (unc_out || vote_out) {   // uniform across the grid
            const float4 t0 = make_float4(dq.x - dsf, dq.y - dsf, dq.z - dsf, dq.w - dsf);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float dj = static_cast<float>(NDS * j);
                U.x = fmaf(fabsf(t0.x - dj), x[j].x, U.x);
                U.y = fmaf(fabsf(t0.y - dj), x[j].y, U.y);
                U.z = fmaf(fabsf(t0.z - dj), x[j].z, U.z);
                U.w = fmaf(fabsf(t0.w - dj), x[j].w, U.w);
            }
            if (SQ < 32) {
            U.x += __shfl_xor_sync(0xffffffffu, U.x, 16); U.y += __shfl_xor_sync(0xffffffffu, U.y, 16);
            U.z += __shfl_xor_sync(0xffffffffu, U.z, 16); U.w += __shfl_xor_sync(0xffffffffu, U.w, 16);
            }
            if (SQ == 32 || lane < SQ) red[3][warp][q] = U;
            asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
            if (fin) {
                U = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const float4 a = red[3][w][q];
                    U.x += a.x; U.y += a.y; U.z += a.z; U.w += a.w;
                }
                U.x *= rS.x; U.y *= rS.y; U.z *= rS.z; U.w *= rS.w;
            }
        }
        if (fin && live) {
            if (disp_out) *reinterpret_cast<float4 *>(disp_out + o) = disp;
            if (unc_out) *reinterpret_cast<float4 *>(unc_out + o) = U;
            if (vote_out) {
                float4 vt;
                const bool nu = used == nullptr;
                vt.x = ((nu || fabsf(dq.x - u4.x) < thr_dif) && U.x < thr_unc) ? 1.0f : 0.0f;
                vt.y = ((nu || fabsf(dq.y - u4.y) < thr_dif) && U.y < thr_unc) ? 1.0f : 0.0f;
                vt.z = ((nu || fabsf(dq.z - u4.z) < thr_dif) && U.z < thr_unc) ? 1.0f : 0.0f;
                vt.w = ((nu || fabsf(dq.w - u4.w) < thr_dif) && U.w < thr_unc) ? 1.0f : 0.0f;
                *reinterpret_cast<float4 *>(vote_out + o) = vt;
            }
            if (ens_acc) {
                a4.x = fmaf(ens_coef, disp.x, a4.x); a4.y = fmaf(ens_coef, disp.y, a4.y);
                a4.z = fmaf(ens_coef, disp.z, a4.z); a4.w = fmaf(ens_coef, disp.w, a4.w);
                *reinterpret_cast<float4 *>(ens_acc + o) = a4;
            }
        }
    }
}

template <int NJ, int NDS, int STAGES, int SPAN, int MINB>
static int launch_sr_tma(const float *cost, int B, int D, int HW, float *disp_out, float *prob_out, const float *used,
                         const float *disp_ext, float *unc_out, float *vote_out, float thr_dif, float thr_unc, float *ens_acc,
                         float ens_coef, int ens_init, int *ctr, cudaStream_t st) {
    const size_t smem = sizeof(float) * STAGES * static_cast<size_t>(D) * SPAN;
    const int spans = (HW + SPAN - 1) / SPAN;
    const int64_t ntiles = static_cast<int64_t>(spans) * B;
    const int grid = static_cast<int>(ntiles < 1LL * MINB * num_sms() ? ntiles : 1LL * MINB * num_sms());
    CUtensorMap tmap;
    const uint64_t dims[3] = {static_cast<uint64_t>(HW), static_cast<uint64_t>(D), static_cast<uint64_t>(B)};
    const uint32_t box[3] = {static_cast<uint32_t>(SPAN), static_cast<uint32_t>(D), 1u};
    if (!make_tensor_map_f32(&tmap, cost, 3, dims, box)) return DV_ERR_UNSUPPORTED;
    const int slot = static_cast<int>(g_sr_slot.fetch_add(1, std::memory_order_relaxed) % kSrSlots);
#define DV_LAUNCH(FULLD)                                                                                               \
    {                                                                                                                  \
        auto kern = softmax_regress_tma_kernel<NJ, NDS, STAGES, SPAN, MINB, true, FULLD>;                              \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=         \
            cudaSuccess)                                                                                               \
            return DV_ERR_LAUNCH;                                                                                      \
        kern<<<grid, SPAN / 4 * NDS + 32, smem, st>>>(tmap, D, HW, spans, static_cast<int>(ntiles), ctr, slot, disp_out,    \
                                                      prob_out, used, disp_ext, unc_out, vote_out, thr_dif, thr_unc, ens_acc, \
                                                      ens_coef, ens_init);                                             \
    }
    if (D == NJ * NDS) DV_LAUNCH(true) else DV_LAUNCH(false)
#undef DV_LAUNCH
    return DV_OK;
}

// Any D: one thread per pixel, three passes over D (the re-reads hit L2).
__global__ void softmax_regress_generic_kernel(const float *__restrict__ cost, int D, int HW,
                                               float *__restrict__ disp_out, float *__restrict__ prob_out,
                                               const float *__restrict__ used, const float *__restrict__ disp_ext,
                                               float *__restrict__ unc_out, float *__restrict__ vote_out, float thr_dif,
                                               float thr_unc, float *__restrict__ ens_acc, float ens_coef, int ens_init,
                                               int64_t total) {
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t b = idx / HW, p = idx % HW;
        const float *cp = cost + b * D * HW + p;
        float m = -INFINITY;
        for (int d = 0; d < D; ++d) m = fmaxf(m, cp[static_cast<int64_t>(d) * HW]);
        float S = 0.0f, Wd = 0.0f;
        for (int d = 0; d < D; ++d) {
            const float e = expf(cp[static_cast<int64_t>(d) * HW] - m);
            S += e;
            Wd = fmaf(static_cast<float>(d), e, Wd);
        }
        const float rS = 1.0f / S, disp = Wd * rS;
        const float dq = disp_ext ? disp_ext[idx] : disp;
        float U = 0.0f;
        if (prob_out || unc_out || vote_out) {
            for (int d = 0; d < D; ++d) {
                const float pr = expf(cp[static_cast<int64_t>(d) * HW] - m) * rS;
                if (prob_out) prob_out[b * D * HW + static_cast<int64_t>(d) * HW + p] = pr;
                U = fmaf(fabsf(dq - static_cast<float>(d)), pr, U);
            }
        }
        if (disp_out) disp_out[idx] = disp;
        if (unc_out) unc_out[idx] = U;
        if (vote_out) vote_out[idx] = ((!used || fabsf(dq - used[idx]) < thr_dif) && U < thr_unc) ? 1.0f : 0.0f;
        if (ens_acc) ens_acc[idx] = fmaf(ens_coef, disp, ens_init ? 0.0f : ens_acc[idx]);
    }
}

// disparity_regression alone: out[b,p] = sum_d d * x[b,d,p] (sequential over d like the reference's sum)
template <int V>
__global__ void __launch_bounds__(256)
disparity_regression_kernel(const float *__restrict__ x, float *__restrict__ out, int D, int HW) {
    const int b = blockIdx.y;
    const int64_t pv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (pv * V >= HW) return;
    const float *xp = x + static_cast<int64_t>(b) * D * HW + pv * V;
    Vec<V> acc;
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = 0.0f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const Vec<V> t = load_vec<V>(xp + static_cast<int64_t>(d) * HW);
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] = fmaf(t.v[i], static_cast<float>(d), acc.v[i]);
    }
    store_vec<V>(out + static_cast<int64_t>(b) * HW + pv * V, acc);
}

// a11 when the disparity is NOT the regression of `prob` (PCWNet: the refined disparity is compared with
// the pre-refinement distribution, pwcnet_ddim.py:553-570): unc = sum_d |disp - d| * prob[d].
template <int V>
__global__ void __launch_bounds__(256)
uncertainty_vote_kernel(const float *__restrict__ prob, const float *__restrict__ disp, const float *__restrict__ used,
                        float *__restrict__ unc_out, float *__restrict__ vote_out, float thr_dif, float thr_unc, int D,
                        int HW) {
    const int b = blockIdx.y;
    const int64_t pv = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (pv * V >= HW) return;
    const int64_t o = static_cast<int64_t>(b) * HW + pv * V;
    const float *pp = prob + static_cast<int64_t>(b) * D * HW + pv * V;
    const Vec<V> dsp = load_vec<V>(disp + o);
    Vec<V> U;
#pragma unroll
    for (int i = 0; i < V; ++i) U.v[i] = 0.0f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const Vec<V> t = load_vec<V>(pp + static_cast<int64_t>(d) * HW);
#pragma unroll
        for (int i = 0; i < V; ++i) U.v[i] = fmaf(fabsf(dsp.v[i] - static_cast<float>(d)), t.v[i], U.v[i]);
    }
    if (unc_out) store_vec<V>(unc_out + o, U);
    if (vote_out) {
        Vec<V> vt;
        if (used) {
            const Vec<V> u0 = load_vec<V>(used + o);
#pragma unroll
            for (int i = 0; i < V; ++i)
                vt.v[i] = (fabsf(dsp.v[i] - u0.v[i]) < thr_dif && U.v[i] < thr_unc) ? 1.0f : 0.0f;
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) vt.v[i] = U.v[i] < thr_unc ? 1.0f : 0.0f;
        }
        store_vec<V>(vote_out + o, vt);
    }
}

template <int DPT, int V, int SL, int MINB>
static void launch_sr(const float *cost, int B, int D, int HW, float *disp_out, float *prob_out, const float *used,
                      const float *disp_ext, float *unc_out, float *vote_out, float thr_dif, float thr_unc, float *ens_acc, float ens_coef,
                      int ens_init, cudaStream_t st) {
    const int pvs = (HW + V - 1) / V;
    dim3 grid((pvs + kSrLanes - 1) / kSrLanes, B);
    if (D == SL * DPT)
        softmax_regress_kernel<DPT, V, SL, MINB, true><<<grid, SL * kSrLanes, 0, st>>>(
            cost, D, HW, disp_out, prob_out, used, disp_ext, unc_out, vote_out, thr_dif, thr_unc, ens_acc, ens_coef, ens_init);
    else
        softmax_regress_kernel<DPT, V, SL, MINB, false><<<grid, SL * kSrLanes, 0, st>>>(
            cost, D, HW, disp_out, prob_out, used, disp_ext, unc_out, vote_out, thr_dif, thr_unc, ens_acc, ens_coef, ens_init);
}

static int softmax_regress_impl(const float *cost, int64_t B, int64_t D, int64_t H, int64_t W, float *disp_out,
                                float *prob_out, const float *used, const float *disp_ext, float *unc_out, float *vote_out,
                                float thr_dif, float thr_unc, float *ens_acc, float ens_coef, int ens_init,
                                void *tile_counters, void *stream) {
    if (!cost) return DV_ERR_NULL;
    if (vote_out && !used && !disp_ext) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || D > INT32_MAX) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto all_aligned = [&](uintptr_t mask) {
        auto ok = [mask](const void *p) { return !p || (reinterpret_cast<uintptr_t>(p) & mask) == 0; };
        return ok(cost) && ok(disp_out) && ok(prob_out) && ok(used) && ok(disp_ext) && ok(unc_out) && ok(vote_out) && ok(ens_acc);
    };
    const bool vec4 = (HW % 4 == 0) && all_aligned(15);
    const bool vec2 = (HW % 2 == 0) && all_aligned(7);
    // D <= 8*DPT.  Tuning switch DV_SR_VARIANT (scripts/tune_kernels.py) selects the D = 192 layout.
    const int variant = DV_TUNE("DV_SR_VARIANT", 4);
    if (vec4 && D <= 192 && DV_TUNE("DV_SR_TMA", 1) && static_cast<int64_t>((HW + 63) / 64) * B <= INT32_MAX) {
        int rc;
#define DV_SR_TMA_ARGS cost, B, D, HW, disp_out, prob_out, used, disp_ext, unc_out, vote_out, thr_dif, thr_unc, ens_acc, ens_coef, ens_init, static_cast<int *>(tile_counters), st
        // thread = (quad of 16, slice of 8): 4 consumer warps per CTA, 2 CTAs per SM.  Measured at D = 192, B = 8
        // (gpurun_out/bench_sr3.log): 8 slices x 24 d per thread 0.496 ms (6.59 TB/s) vs 16 slices x 12 d 0.586 ms — half as
        // many partials to merge per pixel and half the per-tile overhead per element; 3 CTAs x 1 stage spills.
        const int shape = DV_TUNE("DV_SR_SHAPE", 2);
        if (D <= 48) rc = launch_sr_tma<6, 8, 4, 64, 2>(DV_SR_TMA_ARGS);
        else if (D <= 96) rc = launch_sr_tma<12, 8, 4, 64, 2>(DV_SR_TMA_ARGS);
        else if (shape == 0) rc = launch_sr_tma<12, 16, 2, 64, 2>(DV_SR_TMA_ARGS);
        else rc = launch_sr_tma<24, 8, 2, 64, 2>(DV_SR_TMA_ARGS);
#undef DV_SR_TMA_ARGS
        if (rc == DV_OK) return finish_launch();
        if (rc != DV_ERR_UNSUPPORTED) return rc;   // no tensor-map encoder: fall through to the register kernel
    }
#define DV_SR_ARGS cost, B, D, HW, disp_out, prob_out, used, disp_ext, unc_out, vote_out, thr_dif, thr_unc, ens_acc, ens_coef, ens_init, st
    if (D <= 48) {
        if (vec4) launch_sr<6, 4, 8, 4>(DV_SR_ARGS);
        else if (vec2) launch_sr<6, 2, 8, 4>(DV_SR_ARGS);
        else launch_sr<6, 1, 8, 4>(DV_SR_ARGS);
    } else if (D <= 96) {
        if (vec4) launch_sr<12, 4, 8, 2>(DV_SR_ARGS);
        else if (vec2) launch_sr<12, 2, 8, 3>(DV_SR_ARGS);
        else launch_sr<12, 1, 8, 4>(DV_SR_ARGS);
    } else if (D <= 192) {
        if (vec4 && variant == 0) launch_sr<24, 4, 8, 1>(DV_SR_ARGS);
        else if (vec4 && variant == 1) launch_sr<12, 4, 16, 1>(DV_SR_ARGS);
        else if (vec4 && variant == 5) launch_sr<12, 4, 16, 2>(DV_SR_ARGS);
        else if (vec2 && variant == 2) launch_sr<24, 2, 8, 2>(DV_SR_ARGS);
        else if (vec2 && variant == 3) launch_sr<12, 2, 16, 2>(DV_SR_ARGS);
        else if (vec2) launch_sr<24, 2, 8, 3>(DV_SR_ARGS);   // measured best of the six layouts (r01 tuning)
        else launch_sr<24, 1, 8, 4>(DV_SR_ARGS);
    } else {
        const int64_t total = B * HW;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
        softmax_regress_generic_kernel<<<grid, 256, 0, st>>>(cost, static_cast<int>(D), static_cast<int>(HW), disp_out,
                                                             prob_out, used, disp_ext, unc_out, vote_out, thr_dif, thr_unc,
                                                             ens_acc, ens_coef, ens_init, total);
    }
#undef DV_SR_ARGS
#undef DV_SR
    return finish_launch();
}

}  // namespace dv

extern "C" int dv_softmax_regress_f32(const float *cost, int64_t B, int64_t D, int64_t H, int64_t W, float *disp_out,
                                      float *prob_out, const float *used, float *unc_out, float *vote_out,
                                      float thr_dif, float thr_unc, float *ens_acc, float ens_coef, int ens_init,
                                      void *tile_counters, void *stream) {
    if (vote_out && !used) return DV_ERR_NULL;
    return dv::softmax_regress_impl(cost, B, D, H, W, disp_out, prob_out, used, nullptr, unc_out, vote_out, thr_dif, thr_unc,
                                    ens_acc, ens_coef, ens_init, tile_counters, stream);
}

extern "C" int dv_softmax_uncertainty_vote_f32(const float *cost, const float *disp, const float *used, int64_t B, int64_t D,
                                               int64_t H, int64_t W, float thr_dif, float thr_unc, float *unc_out,
                                               float *vote_out, void *tile_counters, void *stream) {
    if (!disp || (!unc_out && !vote_out)) return DV_ERR_NULL;
    return dv::softmax_regress_impl(cost, B, D, H, W, nullptr, nullptr, used, disp, unc_out, vote_out, thr_dif, thr_unc, nullptr,
                                    0.0f, 0, tile_counters, stream);
}

extern "C" int dv_disparity_regression_f32(const float *x, float *out, int64_t B, int64_t D, int64_t H, int64_t W,
                                           void *stream) {
    using namespace dv;
    if (!x || !out) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((HW % 4 == 0) && aligned16(x) && aligned16(out)) {
        dim3 grid(static_cast<unsigned>((HW / 4 + 255) / 256), static_cast<unsigned>(B));
        disparity_regression_kernel<4><<<grid, 256, 0, st>>>(x, out, static_cast<int>(D), static_cast<int>(HW));
    } else {
        dim3 grid(static_cast<unsigned>((HW + 255) / 256), static_cast<unsigned>(B));
        disparity_regression_kernel<1><<<grid, 256, 0, st>>>(x, out, static_cast<int>(D), static_cast<int>(HW));
    }
    return finish_launch();
}

extern "C" int dv_uncertainty_vote_f32(const float *prob, const float *disp, const float *used, int64_t B, int64_t D,
                                       int64_t H, int64_t W, float thr_dif, float thr_unc, float *unc_out,
                                       float *vote_out, void *stream) {
    using namespace dv;
    if (!prob || !disp) return DV_ERR_NULL;
    if (!unc_out && !vote_out) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto ok16 = [](const void *p) { return !p || aligned16(p); };
    if ((HW % 4 == 0) && ok16(prob) && ok16(disp) && ok16(used) && ok16(unc_out) && ok16(vote_out)) {
        dim3 grid(static_cast<unsigned>((HW / 4 + 255) / 256), static_cast<unsigned>(B));
        uncertainty_vote_kernel<4><<<grid, 256, 0, st>>>(prob, disp, used, unc_out, vote_out, thr_dif, thr_unc,
                                                         static_cast<int>(D), static_cast<int>(HW));
    } else {
        dim3 grid(static_cast<unsigned>((HW + 255) / 256), static_cast<unsigned>(B));
        uncertainty_vote_kernel<1><<<grid, 256, 0, st>>>(prob, disp, used, unc_out, vote_out, thr_dif, thr_unc,
                                                         static_cast<int>(D), static_cast<int>(HW));
    }
    return finish_launch();
}
