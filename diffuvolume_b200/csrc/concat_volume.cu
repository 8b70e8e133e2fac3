// concat_volume.cu — concatenation volume with the ACV attention weights and the DDIM filter
// multiply fused into the single producing pass (a3, a4, a9) for sm_100a.
//
// Replaces build_concat_volume (variant M: SceneFlow/models/submodule.py:180-191,
// KITTI15/core/submodule.py:206-217; variant T: SceneFlow/submodule.py:137-148,
// KITTI12/models/submodule.py:86-97), `F.softmax(att_weights, dim=2) * concat_volume`
// (SceneFlow/models/acv_ddim.py:390, acv.py:203) and the filter multiply
// `volume * noise.unsqueeze(1).float()` (acv_ddim.py:254-260).  The reference runs these as
// memset + 2*D strided copies, a softmax, a 398 MB read+write multiply and (per DDIM step)
// another 398 MB read+write multiply.  Here the [B,2C,D,H,W] volume is written once,
// straight from the 1/4-resolution features (20 MB), with 128-bit streaming stores.
//
// The op is write-bound (398 MB written per 20 MB read), so there is no reuse to stage:
// CTA = (span of 32 quads of the flattened plane, channel group, batch).  It first builds
// the per-(d,pixel) factors of its span in shared memory (softmax over D of the attention
// logits and/or the filter factor n), then every thread = (quad, channel slot) streams the
// D planes of its channels: left half = one float4 of ref reused for all d; right half =
// a sliding window over tgt (one new aligned float4 per 4 disparities).
#include "common.cuh"

namespace dv {

constexpr int kConcatSQ = 32;           // quads per span  (one warp wide)
constexpr int kConcatSpan = kConcatSQ * 4;
constexpr int kConcatSlots = 8;         // channel slots per CTA
constexpr int kConcatThreads = kConcatSQ * kConcatSlots;

template <typename XT>
__device__ __forceinline__ void fill_filter_factor(float *sn, const XT *__restrict__ xt, const float *__restrict__ shift,
                                                   XT scale, int b, int D, int HW, int p0) {
    for (int e = threadIdx.x; e < D * kConcatSpan; e += kConcatThreads) {
        const int d = e / kConcatSpan, px = e % kConcatSpan;
        const int p = p0 + px;
        float v = 0.0f;
        if (p < HW) {
            const XT x = xt[(static_cast<int64_t>(b) * D + d) * HW + p];
            const float sh = shift ? shift[b * D + d] : 0.0f;
            v = static_cast<float>(filter_n<XT>(x, sh, scale));
        }
        sn[e] = v;
    }
}

template <bool HAS_ATT, bool HAS_N>
__global__ void __launch_bounds__(kConcatThreads)
concat_volume_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, float *__restrict__ out,
                     int C, int HW, int W, int D, int mask_left, int chans_per_cta,
                     const float *__restrict__ att, const void *__restrict__ xt, int xt_is_f64,
                     const float *__restrict__ shift, double scale) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                                       // [D][span] softmax weights
    float *sn = smem + (HAS_ATT ? D * kConcatSpan : 0);     // [D][span] filter factor
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * kConcatSpan;
    const int c_begin = blockIdx.y * chans_per_cta;
    const int c_end = min(c_begin + chans_per_cta, 2 * C);

    if (HAS_ATT) {
        // stage the raw attention logits of the span: all threads, coalesced, independent loads
        for (int e = threadIdx.x; e < D * kConcatSpan; e += kConcatThreads) {
            const int d = e / kConcatSpan, px = e % kConcatSpan;
            const int p = p0 + px;
            sw[e] = p < HW ? __ldg(att + (static_cast<int64_t>(b) * D + d) * HW + p) : 0.0f;
        }
    }
    if (HAS_N) {
        if (xt_is_f64)
            fill_filter_factor<double>(sn, static_cast<const double *>(xt), shift, scale, b, D, HW, p0);
        else
            fill_filter_factor<float>(sn, static_cast<const float *>(xt), shift, static_cast<float>(scale), b, D, HW, p0);
    }
    if (HAS_ATT) {
        __syncthreads();
        // softmax over D of att[b,0,:,p] (F.softmax(att_weights, dim=2)): exp(a - max) / sum, in place
        if (threadIdx.x < kConcatSpan) {
            float *col = sw + threadIdx.x;
            float mx = -INFINITY;
            for (int d = 0; d < D; ++d) mx = fmaxf(mx, col[d * kConcatSpan]);
            float sum = 0.0f;
            for (int d = 0; d < D; ++d) {
                const float e = expf(col[d * kConcatSpan] - mx);
                col[d * kConcatSpan] = e;
                sum += e;
            }
            for (int d = 0; d < D; ++d) col[d * kConcatSpan] = col[d * kConcatSpan] / sum;
        }
    }
    if (HAS_ATT || HAS_N) __syncthreads();

    const int q = threadIdx.x % kConcatSQ;
    const int slot = threadIdx.x / kConcatSQ;
    const int p = p0 + 4 * q;
    if (p >= HW) return;
    int xs[4];
    xs[0] = p % W;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        xs[i] = xs[i - 1] + 1;
        if (xs[i] >= W) xs[i] -= W;
    }

    for (int c = c_begin + slot; c < c_end; c += kConcatSlots) {
        float *op = out + ((static_cast<int64_t>(b) * 2 * C + c) * D) * HW + p;
        if (c < C) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(ref + (static_cast<int64_t>(b) * C + c) * HW + p));
#pragma unroll 4
            for (int d = 0; d < D; ++d) {
                float4 o = v;
                if (mask_left) {
                    o.x = xs[0] >= d ? o.x : 0.0f;
                    o.y = xs[1] >= d ? o.y : 0.0f;
                    o.z = xs[2] >= d ? o.z : 0.0f;
                    o.w = xs[3] >= d ? o.w : 0.0f;
                }
                if (HAS_ATT) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(sw + d * kConcatSpan + 4 * q);
                    o.x *= w4.x; o.y *= w4.y; o.z *= w4.z; o.w *= w4.w;
                }
                if (HAS_N) {
                    const float4 n4 = *reinterpret_cast<const float4 *>(sn + d * kConcatSpan + 4 * q);
                    o.x *= n4.x; o.y *= n4.y; o.z *= n4.z; o.w *= n4.w;
                }
                stg_cs(reinterpret_cast<float4 *>(op + static_cast<int64_t>(d) * HW), o);
            }
        } else {
            // right half: out[d] = tgt[p - d + i]; aligned blocks blk(m) = tgt[p-4m .. p-4m+3]
            const int64_t base = (static_cast<int64_t>(b) * C + (c - C)) * HW + p;  // flat index into tgt
            float4 cur = __ldg(reinterpret_cast<const float4 *>(tgt + base));
            for (int m = 0; 4 * m < D; ++m) {
                const int64_t nb = base - 4 * (m + 1);
                // only the first plane of the tensor can reach below index 0; x < d there
                float4 prev = nb >= 0 ? __ldg(reinterpret_cast<const float4 *>(tgt + nb)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int d = 4 * m + j;
                    if (d < D) {
                        float4 o;
                        if (j == 0) o = cur;
                        else if (j == 1) o = make_float4(prev.w, cur.x, cur.y, cur.z);
                        else if (j == 2) o = make_float4(prev.z, prev.w, cur.x, cur.y);
                        else o = make_float4(prev.y, prev.z, prev.w, cur.x);
                        o.x = xs[0] >= d ? o.x : 0.0f;
                        o.y = xs[1] >= d ? o.y : 0.0f;
                        o.z = xs[2] >= d ? o.z : 0.0f;
                        o.w = xs[3] >= d ? o.w : 0.0f;
                        if (HAS_ATT) {
                            const float4 w4 = *reinterpret_cast<const float4 *>(sw + d * kConcatSpan + 4 * q);
                            o.x *= w4.x; o.y *= w4.y; o.z *= w4.z; o.w *= w4.w;
                        }
                        if (HAS_N) {
                            const float4 n4 = *reinterpret_cast<const float4 *>(sn + d * kConcatSpan + 4 * q);
                            o.x *= n4.x; o.y *= n4.y; o.z *= n4.z; o.w *= n4.w;
                        }
                        stg_cs(reinterpret_cast<float4 *>(op + static_cast<int64_t>(d) * HW), o);
                    }
                }
                cur = prev;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Weighted producer (the shipped path of the DDIM loop): the per-(b,d,pixel) factors arrive PRECOMPUTED as
// fp32 maps — w = softmax_D(att logits) (dv_att_softmax_f32, once per pair) and n = the filter factor
// (dv_filter_factor_f32 / the n_next output of dv_ddim_step, 6 MB per pair and step) — so the volume-sized
// pass has no prologue, no shared memory and no barrier: thread = (quad q of a 128-px span, slot) walks
// d = slot, slot+8, ...; per d it loads w and n once (2 x LDG.128, L2-resident) and produces CG channels
// (left: the ref quad, held in registers for the whole CTA; right: two aligned LDG.128 of tgt + a warp-uniform
// select).  Small CTAs (CG x D planes x 512 B) keep the grid >= 25 waves, which the micro-benchmarks
// (scripts/ubench/ubench_mem.cu) show is what separates 6.4 from 7.3 TB/s on a write-only stream.
template <int CG, bool HAS_W, bool HAS_N, int MINB>
__global__ void __launch_bounds__(256, MINB)
concat_weighted_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, float *__restrict__ out, int C,
                       int HW, int W, int D, int mask_left, const float *__restrict__ wts,
                       const float *__restrict__ nf) {
    const int q = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const int b = blockIdx.z;
    const int p = (blockIdx.x * 32 + q) * 4;
    if (p >= HW) return;
    const int c0 = blockIdx.y * CG;
    int xs[4];
    xs[0] = p % W;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        xs[i] = xs[i - 1] + 1;
        if (xs[i] >= W) xs[i] -= W;
    }
    // left-half channels of this CTA: one quad each, reused for every d
    float4 lv[CG];
#pragma unroll
    for (int k = 0; k < CG; ++k) {
        const int c = c0 + k;
        lv[k] = (c < C) ? __ldg(reinterpret_cast<const float4 *>(ref + (static_cast<int64_t>(b) * C + c) * HW + p))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t fbase = static_cast<int64_t>(b) * D * HW + p;
    for (int d = slot; d < D; d += 8) {
        float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), n4 = w4;
        if (HAS_W) w4 = __ldg(reinterpret_cast<const float4 *>(wts + fbase + static_cast<int64_t>(d) * HW));
        if (HAS_N) n4 = __ldg(reinterpret_cast<const float4 *>(nf + fbase + static_cast<int64_t>(d) * HW));
        const bool k0 = xs[0] >= d, k1 = xs[1] >= d, k2 = xs[2] >= d, k3 = xs[3] >= d;
        const int sh = d & 3;            // warp-uniform
        const int dal = d - sh;          // aligned part of the shift
#pragma unroll
        for (int k = 0; k < CG; ++k) {
            const int c = c0 + k;
            if (c >= 2 * C) break;
            float4 o;
            if (c < C) {
                o = lv[k];
                if (mask_left) {
                    o.x = k0 ? o.x : 0.0f; o.y = k1 ? o.y : 0.0f; o.z = k2 ? o.z : 0.0f; o.w = k3 ? o.w : 0.0f;
                }
            } else {
                // out = tgt[p-d .. p-d+3] = last `sh` floats of block (p-dal-4) ++ first 4-sh floats of block (p-dal)
                const int64_t base = (static_cast<int64_t>(b) * C + (c - C)) * HW + p - dal;   // flat index into tgt
                // only the first rows of the very first plane can reach below 0; x < d there
                const float4 cur = base >= 0 ? __ldg(reinterpret_cast<const float4 *>(tgt + base)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sh != 0 && base >= 4) prev = __ldg(reinterpret_cast<const float4 *>(tgt + base - 4));
                if (sh == 0) o = cur;
                else if (sh == 1) o = make_float4(prev.w, cur.x, cur.y, cur.z);
                else if (sh == 2) o = make_float4(prev.z, prev.w, cur.x, cur.y);
                else o = make_float4(prev.y, prev.z, prev.w, cur.x);
                o.x = k0 ? o.x : 0.0f; o.y = k1 ? o.y : 0.0f; o.z = k2 ? o.z : 0.0f; o.w = k3 ? o.w : 0.0f;
            }
            if (HAS_W) { o.x *= w4.x; o.y *= w4.y; o.z *= w4.z; o.w *= w4.w; }
            if (HAS_N) { o.x *= n4.x; o.y *= n4.y; o.z *= n4.z; o.w *= n4.w; }
            stg_cs(reinterpret_cast<float4 *>(out + ((static_cast<int64_t>(b) * 2 * C + c) * D + d) * HW + p), o);
        }
    }
}

// Variant with all loads hoisted: thread = (quad q, group ds of 4 consecutive disparities d0 = 4 ds).  The two
// factor quads of each of its 4 disparities are fetched first (8 independent LDG.128), then per channel the
// thread needs ONE ref quad (left half) or TWO aligned tgt quads (right half: d0 % 4 == 0 makes the four shifted
// windows compile-time permutations of {prev, cur}) for FOUR 128-bit stores.
template <int CG, bool HAS_W, bool HAS_N>
__global__ void __launch_bounds__(384)
concat_weighted4_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, float *__restrict__ out, int C,
                        int HW, int W, int D, int mask_left, const float *__restrict__ wts,
                        const float *__restrict__ nf) {
    const int q = threadIdx.x & 31, slot = threadIdx.x >> 5, nslots = blockDim.x >> 5;
    const int b = blockIdx.z;
    const int p = (blockIdx.x * 32 + q) * 4;
    if (p >= HW) return;
    const int c0 = blockIdx.y * CG;
    int xs[4];
    xs[0] = p % W;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        xs[i] = xs[i - 1] + 1;
        if (xs[i] >= W) xs[i] -= W;
    }
    const int64_t fbase = static_cast<int64_t>(b) * D * HW + p;
    for (int d0 = 4 * slot; d0 < D; d0 += 4 * nslots) {
        float4 w4[4], n4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            w4[j] = make_float4(1.f, 1.f, 1.f, 1.f);
            n4[j] = w4[j];
            if (d0 + j < D) {
                if (HAS_W) w4[j] = __ldg(reinterpret_cast<const float4 *>(wts + fbase + static_cast<int64_t>(d0 + j) * HW));
                if (HAS_N) n4[j] = __ldg(reinterpret_cast<const float4 *>(nf + fbase + static_cast<int64_t>(d0 + j) * HW));
            }
        }
        float4 cur[CG], prev[CG];
#pragma unroll
        for (int k = 0; k < CG; ++k) {
            const int c = c0 + k;
            cur[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            prev[k] = cur[k];
            if (c < C) {
                cur[k] = __ldg(reinterpret_cast<const float4 *>(ref + (static_cast<int64_t>(b) * C + c) * HW + p));
            } else if (c < 2 * C) {
                const int64_t base = (static_cast<int64_t>(b) * C + (c - C)) * HW + p - d0;   // flat index into tgt
                // only the first rows of the very first plane can reach below 0; x < d there
                if (base >= 0) cur[k] = __ldg(reinterpret_cast<const float4 *>(tgt + base));
                if (base >= 4) prev[k] = __ldg(reinterpret_cast<const float4 *>(tgt + base - 4));
            }
        }
#pragma unroll
        for (int k = 0; k < CG; ++k) {
            const int c = c0 + k;
            if (c >= 2 * C) break;
            float *op = out + ((static_cast<int64_t>(b) * 2 * C + c) * D + d0) * HW + p;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = d0 + j;
                if (d >= D) break;
                float4 o;
                if (c < C || j == 0) o = cur[k];
                else if (j == 1) o = make_float4(prev[k].w, cur[k].x, cur[k].y, cur[k].z);
                else if (j == 2) o = make_float4(prev[k].z, prev[k].w, cur[k].x, cur[k].y);
                else o = make_float4(prev[k].y, prev[k].z, prev[k].w, cur[k].x);
                if (c >= C || mask_left) {
                    o.x = xs[0] >= d ? o.x : 0.0f; o.y = xs[1] >= d ? o.y : 0.0f;
                    o.z = xs[2] >= d ? o.z : 0.0f; o.w = xs[3] >= d ? o.w : 0.0f;
                }
                if (HAS_W) { o.x *= w4[j].x; o.y *= w4[j].y; o.z *= w4[j].z; o.w *= w4[j].w; }
                if (HAS_N) { o.x *= n4[j].x; o.y *= n4[j].y; o.z *= n4[j].z; o.w *= n4[j].w; }
                stg_cs(reinterpret_cast<float4 *>(op + static_cast<int64_t>(j) * HW), o);
            }
        }
    }
}

// softmax over D of the attention logits [B,1,D,HW] -> weights [B,D,HW] (F.softmax(att_weights, dim=2),
// acv_ddim.py:390): thread = one quad of pixels, three passes over D (the 2nd and 3rd hit L1/L2).
__global__ void __launch_bounds__(128)
att_softmax_kernel(const float *__restrict__ att, float *__restrict__ wts, int D, int HW) {
    const int b = blockIdx.y;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p >= HW) return;
    const float *ap = att + static_cast<int64_t>(b) * D * HW + p;
    float *wp = wts + static_cast<int64_t>(b) * D * HW + p;
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int d = 0; d < D; ++d) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(ap + static_cast<int64_t>(d) * HW));
        mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
    }
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 0; d < D; ++d) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(ap + static_cast<int64_t>(d) * HW));
        sum.x += expf(v.x - mx.x); sum.y += expf(v.y - mx.y); sum.z += expf(v.z - mx.z); sum.w += expf(v.w - mx.w);
    }
    for (int d = 0; d < D; ++d) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(ap + static_cast<int64_t>(d) * HW));
        float4 o;
        o.x = expf(v.x - mx.x) / sum.x; o.y = expf(v.y - mx.y) / sum.y;
        o.z = expf(v.z - mx.z) / sum.z; o.w = expf(v.w - mx.w) / sum.w;
        *reinterpret_cast<float4 *>(wp + static_cast<int64_t>(d) * HW) = o;
    }
}
// D known at compile time: thread = one pixel, the D logits are loaded ONCE (D independent coalesced loads in flight) and
// stay in registers for the max / exp-sum / normalise passes — the three-pass kernel above issues 3 D dependent loads per
// thread with 14 warps per SM and ran at 2 TB/s.  Same arithmetic in the same order (exp(v - max) / sum, sum accumulated in
// d order), so the result is bit-identical.
template <int DT>
__global__ void __launch_bounds__(128)
att_softmax_reg_kernel(const float *__restrict__ att, float *__restrict__ wts, int HW) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const float *ap = att + static_cast<int64_t>(b) * DT * HW + p;
    float *wp = wts + static_cast<int64_t>(b) * DT * HW + p;
    float v[DT];
#pragma unroll
    for (int d = 0; d < DT; ++d) v[d] = __ldg(ap + static_cast<int64_t>(d) * HW);
    float mx = -INFINITY;
#pragma unroll
    for (int d = 0; d < DT; ++d) mx = fmaxf(mx, v[d]);
    float sum = 0.0f;
#pragma unroll
    for (int d = 0; d < DT; ++d) {
        v[d] = expf(v[d] - mx);
        sum += v[d];
    }
#pragma unroll
    for (int d = 0; d < DT; ++d) wp[static_cast<int64_t>(d) * HW] = v[d] / sum;
}
__global__ void att_softmax_generic_kernel(const float *__restrict__ att, float *__restrict__ wts, int D, int HW,
                                           int64_t total) {
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t b = idx / HW, p = idx % HW;
        const float *ap = att + b * D * HW + p;
        float mx = -INFINITY;
        for (int d = 0; d < D; ++d) mx = fmaxf(mx, ap[static_cast<int64_t>(d) * HW]);
        float sum = 0.0f;
        for (int d = 0; d < D; ++d) sum += expf(ap[static_cast<int64_t>(d) * HW] - mx);
        for (int d = 0; d < D; ++d)
            wts[b * D * HW + static_cast<int64_t>(d) * HW + p] = expf(ap[static_cast<int64_t>(d) * HW] - mx) / sum;
    }
}

// n[b,d,p] = float(((clamp(xt + shift[b,d], -s, s) / s) + 1) / 2)   (acv_ddim.py:256-258)
template <typename XT>
__global__ void filter_factor_kernel(const XT *__restrict__ xt, const float *__restrict__ shift, XT scale,
                                     float *__restrict__ nf, XT *__restrict__ n_native, int HW, int64_t total) {
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t bd = idx / HW;
        const XT n = filter_n<XT>(xt[idx], shift ? shift[bd] : 0.0f, scale);
        if (nf) nf[idx] = static_cast<float>(n);
        if (n_native) n_native[idx] = n;
    }
}

// 4 consecutive pixels of one (b, d) plane per thread, (b, d) in blockIdx.y: no 64-bit division, 128-bit accesses
template <typename XT>
__global__ void __launch_bounds__(256)
filter_factor_quad_kernel(const XT *__restrict__ xt, const float *__restrict__ shift, XT scale, float *__restrict__ nf,
                          XT *__restrict__ n_native, int HW) {
    const int bd = blockIdx.y;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p >= HW) return;
    const float sh = shift ? __ldg(shift + bd) : 0.0f;
    const int64_t o = static_cast<int64_t>(bd) * HW + p;
    XT x[4];
    if constexpr (sizeof(XT) == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(xt + o));
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
        const double2 a = *reinterpret_cast<const double2 *>(xt + o), b = *reinterpret_cast<const double2 *>(xt + o + 2);
        x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
    }
    XT n[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) n[i] = filter_n<XT>(x[i], sh, scale);
    if (nf)
        *reinterpret_cast<float4 *>(nf + o) = make_float4(static_cast<float>(n[0]), static_cast<float>(n[1]),
                                                          static_cast<float>(n[2]), static_cast<float>(n[3]));
    if (n_native) {
#pragma unroll
        for (int i = 0; i < 4; ++i) n_native[o + i] = n[i];
    }
}

template <int CG, bool HAS_W, bool HAS_N>
static int launch_weighted(const float *ref, const float *tgt, float *out, int B, int C, int HW, int W, int D,
                           int mask_left, const float *wts, const float *nf, cudaStream_t st) {
    dim3 grid((HW / 4 + 31) / 32, (2 * C + CG - 1) / CG, B);
    if (DV_TUNE("DV_CONCAT_V", 0) == 1) {
        int slots = (D + 3) / 4;
        if (slots > 12) slots = 12;
        concat_weighted4_kernel<CG, HAS_W, HAS_N><<<grid, 32 * slots, 0, st>>>(ref, tgt, out, C, HW, W, D, mask_left, wts, nf);
        return finish_launch();
    }
    const int minb = DV_TUNE("DV_CONCAT_MINB", 4);
    if (minb == 8)
        concat_weighted_kernel<CG, HAS_W, HAS_N, 8><<<grid, 256, 0, st>>>(ref, tgt, out, C, HW, W, D, mask_left, wts, nf);
    else if (minb == 6)
        concat_weighted_kernel<CG, HAS_W, HAS_N, 6><<<grid, 256, 0, st>>>(ref, tgt, out, C, HW, W, D, mask_left, wts, nf);
    else
        concat_weighted_kernel<CG, HAS_W, HAS_N, 4><<<grid, 256, 0, st>>>(ref, tgt, out, C, HW, W, D, mask_left, wts, nf);
    return finish_launch();
}

// Shape-agnostic path ((H*W) % 4 != 0 or unaligned pointers): one thread per output element.
template <typename XT>
__global__ void concat_volume_generic_kernel(const float *__restrict__ ref, const float *__restrict__ tgt,
                                             float *__restrict__ out, int C, int HW, int W, int D, int mask_left,
                                             const float *__restrict__ att, const XT *__restrict__ xt,
                                             const float *__restrict__ shift, XT scale, int64_t total) {
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % HW);
        int64_t t = idx / HW;
        const int d = static_cast<int>(t % D);
        t /= D;
        const int c = static_cast<int>(t % (2 * C));
        const int64_t b = t / (2 * C);
        const int x = p % W;
        float v;
        if (c < C)
            v = (!mask_left || x >= d) ? ref[(b * C + c) * HW + p] : 0.0f;
        else
            v = x >= d ? tgt[(b * C + (c - C)) * HW + p - d] : 0.0f;
        if (att) {
            const float *ap = att + b * D * HW + p;
            float mx = -INFINITY;
            for (int k = 0; k < D; ++k) mx = fmaxf(mx, ap[static_cast<int64_t>(k) * HW]);
            float sum = 0.0f;
            for (int k = 0; k < D; ++k) sum += expf(ap[static_cast<int64_t>(k) * HW] - mx);
            v *= expf(ap[static_cast<int64_t>(d) * HW] - mx) / sum;
        }
        if (xt) {
            const float sh = shift ? shift[b * D + d] : 0.0f;
            v *= static_cast<float>(filter_n<XT>(xt[(b * D + d) * HW + p], sh, scale));
        }
        out[idx] = v;
    }
}

template <bool HAS_ATT, bool HAS_N>
static int launch_concat(const float *ref, const float *tgt, float *out, int B, int C, int HW, int W, int D,
                         int mask_left, const float *att, const void *xt, int xt_is_f64, const float *shift,
                         double scale, cudaStream_t st) {
    const size_t smem = sizeof(float) * static_cast<size_t>(D) * kConcatSpan * ((HAS_ATT ? 1 : 0) + (HAS_N ? 1 : 0));
    if (smem > 200 * 1024) return DV_ERR_UNSUPPORTED;
    auto kern = concat_volume_kernel<HAS_ATT, HAS_N>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return DV_ERR_LAUNCH;
    }
    const int spans = (HW + kConcatSpan - 1) / kConcatSpan;
    // channels per CTA: amortise the per-CTA factor build, keep >= ~8 waves of CTAs
    int cpc = DV_TUNE("DV_CONCAT_CPC", HAS_N ? 16 : 8);
    while (cpc > 8 && static_cast<int64_t>(spans) * B * ((2 * C + cpc - 1) / cpc) < 8LL * num_sms() * 4) cpc /= 2;
    dim3 grid(spans, (2 * C + cpc - 1) / cpc, B);
    kern<<<grid, kConcatThreads, smem, st>>>(ref, tgt, out, C, HW, W, D, mask_left, cpc, att, xt, xt_is_f64, shift,
                                            scale);
    return finish_launch();
}

}  // namespace dv

extern "C" int dv_concat_volume_f32(const float *ref, const float *tgt, float *out, int64_t B, int64_t C, int64_t H,
                                    int64_t W, int64_t D, int mask_left, const float *att_logits, const void *xt,
                                    int xt_is_f64, const float *shift, double scale, void *stream) {
    using namespace dv;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!ref || !tgt || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
    if (xt && xt_is_f64 != 0 && xt_is_f64 != 1) return DV_ERR_BAD_DTYPE;
    if (xt && !(scale > 0.0)) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || 2 * C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    const bool fast = (HW % 4 == 0) && aligned16(ref) && aligned16(tgt) && aligned16(out) && W >= 4;
    if (fast) {
        int rc;
        if (att_logits && xt)
            rc = launch_concat<true, true>(ref, tgt, out, B, C, HW, W, D, mask_left, att_logits, xt, xt_is_f64, shift, scale, st);
        else if (att_logits)
            rc = launch_concat<true, false>(ref, tgt, out, B, C, HW, W, D, mask_left, att_logits, xt, xt_is_f64, shift, scale, st);
        else if (xt)
            rc = launch_concat<false, true>(ref, tgt, out, B, C, HW, W, D, mask_left, att_logits, xt, xt_is_f64, shift, scale, st);
        else
            rc = launch_concat<false, false>(ref, tgt, out, B, C, HW, W, D, mask_left, att_logits, xt, xt_is_f64, shift, scale, st);
        if (rc != DV_ERR_UNSUPPORTED) return rc;
    }
    const int64_t total = B * 2 * C * D * HW;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
    if (xt && xt_is_f64)
        concat_volume_generic_kernel<double><<<grid, threads, 0, st>>>(
            ref, tgt, out, static_cast<int>(C), static_cast<int>(HW), static_cast<int>(W), static_cast<int>(D), mask_left,
            att_logits, static_cast<const double *>(xt), shift, scale, total);
    else
        concat_volume_generic_kernel<float><<<grid, threads, 0, st>>>(
            ref, tgt, out, static_cast<int>(C), static_cast<int>(HW), static_cast<int>(W), static_cast<int>(D), mask_left,
            att_logits, static_cast<const float *>(xt), shift, static_cast<float>(scale), total);
    return finish_launch();
}

extern "C" int dv_att_softmax_f32(const float *att_logits, float *weights, int64_t B, int64_t D, int64_t H, int64_t W,
                                  void *stream) {
    using namespace dv;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!att_logits || !weights) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || D > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if (D == 48 && DV_TUNE("DV_ATT_REG", 1)) {          // every reference configuration: maxdisp / 4
        dim3 grid(static_cast<unsigned>((HW + 127) / 128), static_cast<unsigned>(B));
        att_softmax_reg_kernel<48><<<grid, 128, 0, st>>>(att_logits, weights, static_cast<int>(HW));
    } else if ((HW % 4 == 0) && aligned16(att_logits) && aligned16(weights)) {
        dim3 grid(static_cast<unsigned>((HW / 4 + 127) / 128), static_cast<unsigned>(B));
        att_softmax_kernel<<<grid, 128, 0, st>>>(att_logits, weights, static_cast<int>(D), static_cast<int>(HW));
    } else {
        const int64_t total = B * HW;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
        att_softmax_generic_kernel<<<grid, 256, 0, st>>>(att_logits, weights, static_cast<int>(D), static_cast<int>(HW), total);
    }
    return finish_launch();
}

extern "C" int dv_filter_factor(const void *xt, int xt_is_f64, const float *shift, double scale, float *n_out_f32,
                                void *n_out_native, int64_t B, int64_t D, int64_t H, int64_t W, void *stream) {
    using namespace dv;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!xt || (!n_out_f32 && !n_out_native)) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || !(scale > 0.0)) return DV_ERR_BAD_SHAPE;
    if (xt_is_f64 != 0 && xt_is_f64 != 1) return DV_ERR_BAD_DTYPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX) return DV_ERR_BAD_SHAPE;
    const int64_t total = B * D * HW;
    auto al = [](const void *p, uintptr_t m) { return !p || (reinterpret_cast<uintptr_t>(p) & m) == 0; };
    if (HW % 4 == 0 && B * D <= 65535 && al(xt, 15) && al(n_out_f32, 15) && al(n_out_native, 15) && DV_TUNE("DV_FF_QUAD", 1)) {
        dim3 qgrid(static_cast<unsigned>((HW / 4 + 255) / 256), static_cast<unsigned>(B * D));
        if (xt_is_f64)
            filter_factor_quad_kernel<double><<<qgrid, 256, 0, st>>>(static_cast<const double *>(xt), shift, scale, n_out_f32,
                                                                   static_cast<double *>(n_out_native), static_cast<int>(HW));
        else
            filter_factor_quad_kernel<float><<<qgrid, 256, 0, st>>>(static_cast<const float *>(xt), shift, static_cast<float>(scale),
                                                                  n_out_f32, static_cast<float *>(n_out_native), static_cast<int>(HW));
        return finish_launch();
    }
    const int64_t blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 16 ? blocks : static_cast<int64_t>(num_sms()) * 16);
    if (xt_is_f64)
        filter_factor_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double *>(xt), shift, scale, n_out_f32,
                                                           static_cast<double *>(n_out_native), static_cast<int>(HW), total);
    else
        filter_factor_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(xt), shift, static_cast<float>(scale),
                                                          n_out_f32, static_cast<float *>(n_out_native), static_cast<int>(HW), total);
    return finish_launch();
}

extern "C" int dv_filter_factor_f32(const void *xt, int xt_is_f64, const float *shift, double scale, float *n_out,
                                    int64_t B, int64_t D, int64_t H, int64_t W, void *stream) {
    return dv_filter_factor(xt, xt_is_f64, shift, scale, n_out, nullptr, B, D, H, W, stream);
}

extern "C" int dv_concat_volume_weighted_f32(const float *ref, const float *tgt, float *out, int64_t B, int64_t C,
                                             int64_t H, int64_t W, int64_t D, int mask_left, const float *att_weights,
                                             const float *n, void *tile_counters, void *stream) {
    using namespace dv;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!ref || !tgt || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || 2 * C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    auto ok16 = [](const void *p) { return !p || aligned16(p); };
    if (!((HW % 4 == 0) && W >= 4 && aligned16(ref) && aligned16(tgt) && aligned16(out) && ok16(att_weights) && ok16(n)))
        return DV_ERR_MISALIGNED;   // callers fall back to dv_concat_volume_f32 + dv_volume_filter_f32
    if (DV_TUNE("DV_CONCAT_STREAM", 1)) {
        const int rc = launch_concat_stream(ref, tgt, out, static_cast<int>(B), static_cast<int>(C), static_cast<int>(HW),
                                            static_cast<int>(W), static_cast<int>(D), mask_left, att_weights, n,
                                            static_cast<int *>(tile_counters), st);
        if (rc != DV_ERR_UNSUPPORTED) return rc;
    }
    const int cg = DV_TUNE("DV_CONCAT_CG", 2);
#define DV_W(CGV)                                                                                                      \
    (att_weights && n ? launch_weighted<CGV, true, true>(ref, tgt, out, B, C, HW, W, D, mask_left, att_weights, n, st)  \
     : att_weights   ? launch_weighted<CGV, true, false>(ref, tgt, out, B, C, HW, W, D, mask_left, att_weights, n, st) \
     : n             ? launch_weighted<CGV, false, true>(ref, tgt, out, B, C, HW, W, D, mask_left, att_weights, n, st) \
                     : launch_weighted<CGV, false, false>(ref, tgt, out, B, C, HW, W, D, mask_left, att_weights, n, st))
    if (cg == 1) return DV_W(1);
    if (cg == 2) return DV_W(2);
    if (cg == 8) return DV_W(8);
    return DV_W(4);
#undef DV_W
}
