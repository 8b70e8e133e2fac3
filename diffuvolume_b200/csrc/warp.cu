// warp.cu — disparity-driven horizontal warp of the right features to the left view.
//
// Replaces warp(x, disp) of KITTI12/models/submodule.py:137-176 (= SceneFlow/submodule.py:188-227), the step right
// before build_corrleation_volume in PCWNet's refinement (pwcnet_ddim.py:493-494).  The reference builds two [B,1,H,W]
// mesh grids, normalises them, runs F.grid_sample twice (features, and a ones tensor for the validity mask) and
// thresholds / multiplies the mask: ~15 launches and four full-resolution temporaries.  Here: one pass, thread =
// (b, y, x) computes the sampling position and the four tap weights once and streams the C channels.
//
// Quirk reproduced: the grid is normalised as 2 v / (size - 1) - 1 (align_corners=True convention) but F.grid_sample
// runs with its default align_corners=False, so ix = (x - disp) * W / (W - 1) - 0.5 and iy = y * H / (H - 1) - 0.5
// (the warp is NOT purely horizontal: rows are resampled too).  Zero padding; mask = [sum of in-bounds weights >= 0.999].
#include "common.cuh"

namespace dv {

constexpr int kWarpCg = 16;  // channels per thread: all 4 x 16 tap loads of a thread are issued before the first blend

// a / b for an integer-valued b <= 4096 from y = RN(1/b): RN(a*y) corrected by one FMA step equals __fdiv_rn(a, b) for every
// normal a (exhaustively checked per divisor, see geo_lookup.cu) — the kernel was issue-bound (ncu: 81 % issue slots) and
// the IEEE division sequence with its slow-path branches was a third of the per-pixel header.
__device__ __forceinline__ float warp_div(float a, float b, float y, bool use_rcp) {
    if (!use_rcp) return __fdiv_rn(a, b);
    const float q = __fmul_rn(a, y);
    return __fmaf_rn(__fmaf_rn(-q, b, a), y, q);
}

// Sampling position, tap weights (zero for out-of-image taps), clamped tap offsets and validity of one output pixel, in the
// reference's fp32 op sequence (no contraction across its separate tensor ops).
struct WarpTaps {
    float w00, w01, w10, w11, m;      // tap weights with zero padding applied; m = [sum of in-bounds weights >= 0.999]
    float wx0, wx1, wy0, wy1;         // raw bilinear weights (the backward needs them unmasked)
    int o00, o01, o10, o11;           // clamped flat offsets inside a channel plane (H*W < 2^31)
    bool okx0, okx1, oky0, oky1;
    __device__ __forceinline__ WarpTaps(float d, int px, int y, int H, int W, float rcp_w, float rcp_h, int use_rcp) {
        const float gx = __fsub_rn(warp_div(__fmul_rn(2.0f, __fsub_rn(static_cast<float>(px), d)), static_cast<float>(max(W - 1, 1)), rcp_w, use_rcp), 1.0f);
        const float gy = __fsub_rn(warp_div(__fmul_rn(2.0f, static_cast<float>(y)), static_cast<float>(max(H - 1, 1)), rcp_h, use_rcp), 1.0f);
        const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(W)), 1.0f), 0.5f);   // "/ 2": exact
        const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(H)), 1.0f), 0.5f);
        const float x0f = floorf(ix), y0f = floorf(iy);
        const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f), x1 = x0 + 1, y1 = y0 + 1;
        wx1 = ix - x0f; wx0 = (x0f + 1.0f) - ix; wy1 = iy - y0f; wy0 = (y0f + 1.0f) - iy;
        okx0 = x0 >= 0 && x0 < W; okx1 = x1 >= 0 && x1 < W; oky0 = y0 >= 0 && y0 < H; oky1 = y1 >= 0 && y1 < H;
        w00 = (okx0 && oky0) ? wx0 * wy0 : 0.0f; w01 = (okx1 && oky0) ? wx1 * wy0 : 0.0f;
        w10 = (okx0 && oky1) ? wx0 * wy1 : 0.0f; w11 = (okx1 && oky1) ? wx1 * wy1 : 0.0f;
        const float msum = ((w00 + w01) + w10) + w11;
        m = msum < 0.999f ? 0.0f : 1.0f;
        const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
        const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
        o00 = cy0 * W + cx0; o01 = cy0 * W + cx1; o10 = cy1 * W + cx0; o11 = cy1 * W + cx1;
    }
};

template <bool ASSEMBLE, int CG>
__global__ void __launch_bounds__(256)
warp_kernel(const float *__restrict__ x, const float *__restrict__ disp, float *__restrict__ out, int C, int H, int W,
            int cgroups, float rcp_w, float rcp_h, int use_rcp, const float *__restrict__ ref, float *__restrict__ diff_out,
            int64_t diff_bstride, float *__restrict__ copy_out, int64_t copy_bstride) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z / cgroups, cbeg = (blockIdx.z % cgroups) * CG;
    const int cend = min(cbeg + CG, C);
    if (px >= W) return;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const WarpTaps t(disp[static_cast<int64_t>(b) * HW + static_cast<int64_t>(y) * W + px], px, y, H, W, rcp_w, rcp_h, use_rcp);
    const float w00 = t.w00, w01 = t.w01, w10 = t.w10, w11 = t.w11, m = t.m;
    const int o00 = t.o00, o01 = t.o01, o10 = t.o10, o11 = t.o11;
    const float *xp = x + static_cast<int64_t>(b) * C * HW;
    const int64_t pofs = static_cast<int64_t>(y) * W + px;
    float *op = out + static_cast<int64_t>(b) * C * HW + pofs;
    // refinement-input assembly (pwcnet_ddim.py:497-499): `ref - warp(x)` and the copy of `ref` go straight into channel
    // slices of the caller's concat buffer
    const float *rp = (ASSEMBLE && ref) ? ref + static_cast<int64_t>(b) * C * HW + pofs : nullptr;
    float *dp = (ASSEMBLE && diff_out) ? diff_out + b * diff_bstride + pofs : nullptr;
    float *cp = (ASSEMBLE && copy_out) ? copy_out + b * copy_bstride + pofs : nullptr;
    if (m == 0.0f) {
        for (int c = cbeg; c < cend; ++c) {
            op[static_cast<int64_t>(c) * HW] = 0.0f;
            if (ASSEMBLE && rp) {
                const float r = __ldg(rp + static_cast<int64_t>(c) * HW);
                if (dp) dp[static_cast<int64_t>(c) * HW] = __fsub_rn(r, 0.0f);
                if (cp) cp[static_cast<int64_t>(c) * HW] = r;
            }
        }
        return;
    }
    // (measured, B = 8 at 384x1248: `ref` loads issued with the taps 0.72 ms; one dependent load per channel inside the
    // blend loop 1.28 ms; hoisted above the masked-pixel branch 0.85 ms)
    float t00[CG], t01[CG], t10[CG], t11[CG], rr[ASSEMBLE ? CG : 1];
#pragma unroll
    for (int k = 0; k < CG; ++k) {
        const float *pc = xp + static_cast<int64_t>(min(cbeg + k, C - 1)) * HW;
        t00[k] = __ldg(pc + o00); t01[k] = __ldg(pc + o01); t10[k] = __ldg(pc + o10); t11[k] = __ldg(pc + o11);
        if (ASSEMBLE) rr[k] = rp ? __ldg(rp + static_cast<int64_t>(min(cbeg + k, C - 1)) * HW) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < CG; ++k) {
        float v = __fmul_rn(t00[k], w00);
        v = __fadd_rn(v, __fmul_rn(t01[k], w01));
        v = __fadd_rn(v, __fmul_rn(t10[k], w10));
        v = __fadd_rn(v, __fmul_rn(t11[k], w11));
        if (cbeg + k < cend) {
            op[static_cast<int64_t>(cbeg + k) * HW] = v;
            if (ASSEMBLE && rp) {
                const float r = rr[k];
                if (dp) dp[static_cast<int64_t>(cbeg + k) * HW] = __fsub_rn(r, v);
                if (cp) cp[static_cast<int64_t>(cbeg + k) * HW] = r;
            }
        }
    }
}

// Backward of warp (KITTI12/models/submodule.py:137-176 under autograd, as PCWNet's training
// differentiates it).  The validity mask is piecewise constant (two in-place index assignments), so
//   grad_x[c, tap]  += w_tap * m * g[c, p]                                     (grid_sample's own scatter-add, RED.ADD.F32)
//   grad_disp[p]     = -2 / (W-1) * (W/2) * sum_c m g[c,p] ((t01 - t00) wy0 + (t11 - t10) wy1)    (in-bounds taps only)
// thread = (pixel, 16-channel group); grad_x and grad_disp must be zero on entry (the entry point clears them).
template <bool NEED_X, bool NEED_D>
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float *__restrict__ g, const float *__restrict__ x, const float *__restrict__ disp,
                float *__restrict__ gx, float *__restrict__ gdisp, int C, int H, int W, int cgroups, float rcp_w, float rcp_h,
                int use_rcp, float half_w, float wm1) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z / cgroups, cbeg = (blockIdx.z % cgroups) * kWarpCg;
    if (px >= W) return;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const int64_t pofs = static_cast<int64_t>(y) * W + px;
    const WarpTaps t(disp[static_cast<int64_t>(b) * HW + pofs], px, y, H, W, rcp_w, rcp_h, use_rcp);
    if (t.m == 0.0f) return;
    const float *gp = g + static_cast<int64_t>(b) * C * HW + pofs;
    const float *xp = x + static_cast<int64_t>(b) * C * HW;
    float *gxp = NEED_X ? gx + static_cast<int64_t>(b) * C * HW : nullptr;
    const bool in00 = t.okx0 && t.oky0, in01 = t.okx1 && t.oky0, in10 = t.okx0 && t.oky1, in11 = t.okx1 && t.oky1;
    float go[kWarpCg], t00[NEED_D ? kWarpCg : 1], t01[NEED_D ? kWarpCg : 1], t10[NEED_D ? kWarpCg : 1], t11[NEED_D ? kWarpCg : 1];
#pragma unroll
    for (int k = 0; k < kWarpCg; ++k) {
        const int64_t co = static_cast<int64_t>(min(cbeg + k, C - 1)) * HW;
        go[k] = cbeg + k < C ? __ldg(gp + co) : 0.0f;
        if (NEED_D) {
            const float *pc = xp + co;
            t00[k] = in00 ? __ldg(pc + t.o00) : 0.0f; t01[k] = in01 ? __ldg(pc + t.o01) : 0.0f;
            t10[k] = in10 ? __ldg(pc + t.o10) : 0.0f; t11[k] = in11 ? __ldg(pc + t.o11) : 0.0f;
        }
    }
    float gix = 0.0f;
#pragma unroll
    for (int k = 0; k < kWarpCg; ++k) {
        if (NEED_X && cbeg + k < C) {
            float *pc = gxp + static_cast<int64_t>(cbeg + k) * HW;
            if (in00) atomicAdd(pc + t.o00, t.w00 * go[k]);
            if (in01) atomicAdd(pc + t.o01, t.w01 * go[k]);
            if (in10) atomicAdd(pc + t.o10, t.w10 * go[k]);
            if (in11) atomicAdd(pc + t.o11, t.w11 * go[k]);
        }
        if (NEED_D) {
            gix -= t00[k] * t.wy0 * go[k];
            gix += t01[k] * t.wy0 * go[k];
            gix -= t10[k] * t.wy1 * go[k];
            gix += t11[k] * t.wy1 * go[k];
        }
    }
    if (NEED_D) {
        const float gd = -(((half_w * gix) / wm1) * 2.0f);
        if (cgroups == 1) gdisp[static_cast<int64_t>(b) * HW + pofs] = gd;
        else atomicAdd(gdisp + static_cast<int64_t>(b) * HW + pofs, gd);
    }
}

}  // namespace dv

template <bool ASSEMBLE, int CG>
static int warp_launch(const float *x, const float *disp, float *out, const float *ref, float *diff_out, int64_t diff_bstride,
                       float *copy_out, int64_t copy_bstride, int64_t B, int64_t C, int64_t H, int64_t W, cudaStream_t st) {
    using namespace dv;
    const int64_t cgroups = (C + CG - 1) / CG;
    if (H * W > INT32_MAX || B * cgroups > 65535 || H > 65535 || C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    dim3 grid(static_cast<unsigned>((W + 255) / 256), static_cast<unsigned>(H), static_cast<unsigned>(B * cgroups));
    const int use_rcp = (W - 1 <= 4096 && H - 1 <= 4096) ? 1 : 0;     // the range the reciprocal division was verified on
    const float rcp_w = 1.0f / static_cast<float>(W > 1 ? W - 1 : 1), rcp_h = 1.0f / static_cast<float>(H > 1 ? H - 1 : 1);
    warp_kernel<ASSEMBLE, CG><<<grid, 256, 0, st>>>(x, disp, out, static_cast<int>(C), static_cast<int>(H), static_cast<int>(W),
                                                    static_cast<int>(cgroups), rcp_w, rcp_h, use_rcp, ref, diff_out, diff_bstride,
                                                    copy_out, copy_bstride);
    return finish_launch();
}

static int warp_impl(const float *x, const float *disp, float *out, const float *ref, float *diff_out, int64_t diff_bstride,
                     float *copy_out, int64_t copy_bstride, int64_t B, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace dv;
    if (!x || !disp || !out) return DV_ERR_NULL;
    if ((diff_out || copy_out) && !ref) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!ref) {
        switch (DV_TUNE("DV_WARP_CG", 16)) {
            case 4: return warp_launch<false, 4>(x, disp, out, nullptr, nullptr, 0, nullptr, 0, B, C, H, W, st);
            case 8: return warp_launch<false, 8>(x, disp, out, nullptr, nullptr, 0, nullptr, 0, B, C, H, W, st);
            default: return warp_launch<false, kWarpCg>(x, disp, out, nullptr, nullptr, 0, nullptr, 0, B, C, H, W, st);
        }
    }
    // the assembling variant keeps the `ref` loads in flight too: fewer channels per thread buy back the occupancy
    switch (DV_TUNE("DV_WARP_ASM_CG", 4)) {
        case 2: return warp_launch<true, 2>(x, disp, out, ref, diff_out, diff_bstride, copy_out, copy_bstride, B, C, H, W, st);
        case 4: return warp_launch<true, 4>(x, disp, out, ref, diff_out, diff_bstride, copy_out, copy_bstride, B, C, H, W, st);
        case 16: return warp_launch<true, 16>(x, disp, out, ref, diff_out, diff_bstride, copy_out, copy_bstride, B, C, H, W, st);
        default: return warp_launch<true, 8>(x, disp, out, ref, diff_out, diff_bstride, copy_out, copy_bstride, B, C, H, W, st);
    }
}

extern "C" int dv_warp_f32(const float *x, const float *disp, float *out, int64_t B, int64_t C, int64_t H, int64_t W,
                           void *stream) {
    return warp_impl(x, disp, out, nullptr, nullptr, 0, nullptr, 0, B, C, H, W, stream);
}

extern "C" int dv_warp_assemble_f32(const float *x, const float *disp, const float *ref, float *warp_out, float *diff_out,
                                    int64_t diff_batch_stride, float *copy_out, int64_t copy_batch_stride, int64_t B, int64_t C,
                                    int64_t H, int64_t W, void *stream) {
    return warp_impl(x, disp, warp_out, ref, diff_out, diff_batch_stride, copy_out, copy_batch_stride, B, C, H, W, stream);
}

extern "C" int dv_warp_bwd_f32(const float *grad_out, const float *x, const float *disp, float *grad_x, float *grad_disp,
                               int64_t B, int64_t C, int64_t H, int64_t W, void *stream) {
    using namespace dv;
    if (!grad_out || !x || !disp || (!grad_x && !grad_disp)) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t cgroups = (C + kWarpCg - 1) / kWarpCg;
    if (H * W > INT32_MAX || B * cgroups > 65535 || H > 65535 || C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // scatter-add targets; masked pixels and (for several channel groups) grad_disp rely on the zero fill
    if (grad_x && cudaMemsetAsync(grad_x, 0, sizeof(float) * B * C * H * W, st) != cudaSuccess) return DV_ERR_LAUNCH;
    if (grad_disp && cudaMemsetAsync(grad_disp, 0, sizeof(float) * B * H * W, st) != cudaSuccess) return DV_ERR_LAUNCH;
    dim3 grid(static_cast<unsigned>((W + 255) / 256), static_cast<unsigned>(H), static_cast<unsigned>(B * cgroups));
    const int use_rcp = (W - 1 <= 4096 && H - 1 <= 4096) ? 1 : 0;
    const float rcp_w = 1.0f / static_cast<float>(W > 1 ? W - 1 : 1), rcp_h = 1.0f / static_cast<float>(H > 1 ? H - 1 : 1);
    const float half_w = static_cast<float>(W) / 2.0f, wm1 = static_cast<float>(W > 1 ? W - 1 : 1);
#define DV_WARP_BWD(NX, ND)                                                                                                   \
    warp_bwd_kernel<NX, ND><<<grid, 256, 0, st>>>(grad_out, x, disp, grad_x, grad_disp, static_cast<int>(C), static_cast<int>(H), \
                                                  static_cast<int>(W), static_cast<int>(cgroups), rcp_w, rcp_h, use_rcp, half_w, wm1)
    if (grad_x && grad_disp) DV_WARP_BWD(true, true);
    else if (grad_x) DV_WARP_BWD(true, false);
    else DV_WARP_BWD(false, true);
#undef DV_WARP_BWD
    return finish_launch();
}
