// upsample_regress.cu — trilinear upsample fused into softmax + disparity regression.
//
// Replaces the sequence (SceneFlow/models/acv_ddim.py:267-270, :320-331; KITTI12/models/pwcnet_ddim.py:480-484)
//     cost = F.upsample(cost_q, [maxdisp, H, W], mode='trilinear')      # [B,1,Dq,h,w] -> [B,1,D,H,W], 398 MB written
//     pred = F.softmax(squeeze(cost, 1), dim=1); disp = disparity_regression(pred, maxdisp)   # 398 MB read, 2x398 MB written
//     uncertainty / renewal vote over the probability volume                                  # another 398 MB pass
// by one kernel that reads only the 6 MB quarter-resolution cost: the full-resolution logits never exist.  The op
// turns from HBM-bound into MUFU/issue-bound (192 exponentials per output pixel; twice that when the uncertainty is
// requested, because the exponentials are recomputed rather than kept: 192 live registers per thread are not
// available), so there is no bandwidth roofline to quote for it — bench.py reports it as a separate mode.
//
// Interpolation follows ATen's upsample_trilinear3d: per axis src = scale * (dst + 0.5) - 0.5 clamped at 0
// (align_corners=False, scale = in/out) or src = dst * (in-1)/(out-1) (align_corners=True); i0 = int(src),
// i1 = i0 + (i0 < in-1), lambda = src - i0.  Trilinear interpolation is separable, so a thread (= one output pixel)
// first blends the 4 spatial taps of each of the Dq planes (values v[0..Dq)) and then walks D as linear blends of
// neighbouring v.  The softmax reference value is max_d' v[d'] >= every interpolated logit (blends are convex), which
// is mathematically the same softmax and cannot overflow.
//
// FAST path (DQ = 48, D = 4 DQ, align_corners = False — every ACVNet call): v lives in registers, pre-scaled by log2(e),
// and the four outputs of each source interval use compile-time blend weights {0.625, 0.875 | 0.125, 0.375}.
// GENERIC path (any sizes, both align modes): v lives in shared memory, per-d taps come from a table built per CTA.
#include "common.cuh"

namespace dv {

struct AxisTap {
    int i0, i1;
    float l1;
};
__device__ __forceinline__ AxisTap axis_tap(int dst, int n_in, int n_out, int align_corners) {
    float src;
    if (align_corners) {
        const float scale = n_out > 1 ? static_cast<float>(n_in - 1) / static_cast<float>(n_out - 1) : 0.0f;
        src = scale * static_cast<float>(dst);
    } else {
        const float scale = static_cast<float>(n_in) / static_cast<float>(n_out);
        src = fmaxf(scale * (static_cast<float>(dst) + 0.5f) - 0.5f, 0.0f);
    }
    AxisTap t;
    t.i0 = min(static_cast<int>(src), n_in - 1);
    t.i1 = min(t.i0 + 1, n_in - 1);
    t.l1 = fminf(fmaxf(src - static_cast<float>(t.i0), 0.0f), 1.0f);
    return t;
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct UpsArgs {
    const float *cost;   // [B, Dq, h, w]
    int Dq, h, w, D, H, W, align_corners;
    float *disp_out;
    const float *used;
    float *unc_out, *vote_out;
    float thr_dif, thr_unc;
    float *ens_acc;
    float ens_coef;
    int ens_init;
};

__device__ __forceinline__ void ups_store(const UpsArgs &a, int64_t o, float disp, float U, bool need_unc) {
    if (a.disp_out) a.disp_out[o] = disp;
    if (a.unc_out) a.unc_out[o] = U;
    if (a.vote_out) a.vote_out[o] = (fabsf(disp - a.used[o]) < a.thr_dif && U < a.thr_unc) ? 1.0f : 0.0f;
    if (a.ens_acc) a.ens_acc[o] = fmaf(a.ens_coef, disp, a.ens_init ? 0.0f : a.ens_acc[o]);
    (void)need_unc;
}

// CTA = 32 x 8 output pixels; thread = one output pixel.
template <int DQ>
__global__ void __launch_bounds__(256, 3)
upsample_regress_fast_kernel(const UpsArgs a) {
    const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
    if (X >= a.W || Y >= a.H) return;
    const AxisTap ty = axis_tap(Y, a.h, a.H, 0), tx = axis_tap(X, a.w, a.W, 0);
    const int hw = a.h * a.w;
    const float *base = a.cost + static_cast<int64_t>(b) * DQ * hw;
    const float *p00 = base + ty.i0 * a.w + tx.i0, *p01 = base + ty.i0 * a.w + tx.i1;
    const float *p10 = base + ty.i1 * a.w + tx.i0, *p11 = base + ty.i1 * a.w + tx.i1;
    constexpr float kLog2e = 1.4426950408889634f;
    float v[DQ];
    float m = -INFINITY;
#pragma unroll
    for (int d = 0; d < DQ; ++d) {
        const float s00 = __ldg(p00 + d * hw), s01 = __ldg(p01 + d * hw), s10 = __ldg(p10 + d * hw), s11 = __ldg(p11 + d * hw);
        const float r0 = fmaf(tx.l1, s01 - s00, s00), r1 = fmaf(tx.l1, s11 - s10, s10);
        v[d] = fmaf(ty.l1, r1 - r0, r0);
        m = fmaxf(m, v[d]);
    }
    const float mL = m * kLog2e;
#pragma unroll
    for (int d = 0; d < DQ; ++d) v[d] = fmaf(v[d], kLog2e, -mL);   // blends are linear: interpolate in the scaled domain
    // ---- S = sum e, Wd = sum d e over the 4*DQ interpolated logits
    float S = 0.0f, Wd = 0.0f;
#pragma unroll
    for (int i = 0; i < DQ; ++i) {
        const float lo = v[i > 0 ? i - 1 : 0], mid = v[i], hi = v[i + 1 < DQ ? i + 1 : DQ - 1];
        const float dl = mid - lo, dh = hi - mid;
        const float e0 = ex2f(fmaf(0.625f, dl, lo)), e1 = ex2f(fmaf(0.875f, dl, lo));
        const float e2 = ex2f(fmaf(0.125f, dh, mid)), e3 = ex2f(fmaf(0.375f, dh, mid));
        S += (e0 + e1) + (e2 + e3);
        Wd = fmaf(static_cast<float>(4 * i), e0, Wd);
        Wd = fmaf(static_cast<float>(4 * i + 1), e1, Wd);
        Wd = fmaf(static_cast<float>(4 * i + 2), e2, Wd);
        Wd = fmaf(static_cast<float>(4 * i + 3), e3, Wd);
    }
    const float rS = 1.0f / S;
    const float disp = Wd * rS;
    float U = 0.0f;
    const bool need_unc = a.unc_out || a.vote_out;
    if (need_unc) {   // uniform across the grid
        // Recompute, do not keep: make v opaque so that the 192 exponentials of the pass above are not treated as
        // common subexpressions (ptxas would otherwise hold all of them live across the reduction and spill).
#pragma unroll
        for (int d = 0; d < DQ; ++d) asm volatile("" : "+f"(v[d]));
#pragma unroll
        for (int i = 0; i < DQ; ++i) {
            const float lo = v[i > 0 ? i - 1 : 0], mid = v[i], hi = v[i + 1 < DQ ? i + 1 : DQ - 1];
            const float dl = mid - lo, dh = hi - mid;
            U = fmaf(fabsf(disp - static_cast<float>(4 * i)), ex2f(fmaf(0.625f, dl, lo)), U);
            U = fmaf(fabsf(disp - static_cast<float>(4 * i + 1)), ex2f(fmaf(0.875f, dl, lo)), U);
            U = fmaf(fabsf(disp - static_cast<float>(4 * i + 2)), ex2f(fmaf(0.125f, dh, mid)), U);
            U = fmaf(fabsf(disp - static_cast<float>(4 * i + 3)), ex2f(fmaf(0.375f, dh, mid)), U);
        }
        U *= rS;
    }
    ups_store(a, (static_cast<int64_t>(b) * a.H + Y) * a.W + X, disp, U, need_unc);
}

// Same kernel with the quarter-resolution footprint of the CTA's 32 x 8 output tile ([DQ] x 4 rows x 10 columns, 7.7 KB)
// staged in shared memory first: the 192 tap reads of a thread become LDS with compile-time offsets (d * 40 floats) instead
// of LDG with 64-bit address arithmetic — ncu on the kernel above: issue slots 76 % busy, ~400 of the 2 380 instructions per
// pixel are those address computations — and every quarter-res value is fetched once per CTA instead of ~16 times.
template <int DQ>
__global__ void __launch_bounds__(256, 3)
upsample_regress_tile_kernel(const UpsArgs a) {
    constexpr int TR = 4, TC = 10, TP = TR * TC;
    __shared__ float st[DQ * TP];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int X0 = blockIdx.x * 32, Y0 = blockIdx.y * 8, b = blockIdx.z;
    const int r0 = axis_tap(Y0, a.h, a.H, 0).i0, c0 = axis_tap(X0, a.w, a.W, 0).i0;
    const int hw = a.h * a.w;
    const float *base = a.cost + static_cast<int64_t>(b) * DQ * hw;
    for (int e = tid; e < DQ * TP; e += 256) {
        const int d = e / TP, rem = e - d * TP, r = rem / TC, c = rem - r * TC;
        st[e] = __ldg(base + d * hw + min(r0 + r, a.h - 1) * a.w + min(c0 + c, a.w - 1));
    }
    __syncthreads();
    const int X = X0 + threadIdx.x, Y = Y0 + threadIdx.y;
    if (X >= a.W || Y >= a.H) return;
    const AxisTap ty = axis_tap(Y, a.h, a.H, 0), tx = axis_tap(X, a.w, a.W, 0);
    const float *p00 = st + (ty.i0 - r0) * TC + (tx.i0 - c0), *p01 = st + (ty.i0 - r0) * TC + (tx.i1 - c0);
    const float *p10 = st + (ty.i1 - r0) * TC + (tx.i0 - c0), *p11 = st + (ty.i1 - r0) * TC + (tx.i1 - c0);
    constexpr float kLog2e = 1.4426950408889634f;
    float v[DQ];
    float m = -INFINITY;
#pragma unroll
    for (int d = 0; d < DQ; ++d) {
        const float s00 = p00[d * TP], s01 = p01[d * TP], s10 = p10[d * TP], s11 = p11[d * TP];
        const float r0v = fmaf(tx.l1, s01 - s00, s00), r1v = fmaf(tx.l1, s11 - s10, s10);
        v[d] = fmaf(ty.l1, r1v - r0v, r0v);
        m = fmaxf(m, v[d]);
    }
    const float mL = m * kLog2e;
#pragma unroll
    for (int d = 0; d < DQ; ++d) v[d] = fmaf(v[d], kLog2e, -mL);
    float S = 0.0f, Wd = 0.0f;
#pragma unroll
    for (int i = 0; i < DQ; ++i) {
        const float lo = v[i > 0 ? i - 1 : 0], mid = v[i], hi = v[i + 1 < DQ ? i + 1 : DQ - 1];
        const float dl = mid - lo, dh = hi - mid;
        const float e0 = ex2f(fmaf(0.625f, dl, lo)), e1 = ex2f(fmaf(0.875f, dl, lo));
        const float e2 = ex2f(fmaf(0.125f, dh, mid)), e3 = ex2f(fmaf(0.375f, dh, mid));
        S += (e0 + e1) + (e2 + e3);
        Wd = fmaf(static_cast<float>(4 * i), e0, Wd);
        Wd = fmaf(static_cast<float>(4 * i + 1), e1, Wd);
        Wd = fmaf(static_cast<float>(4 * i + 2), e2, Wd);
        Wd = fmaf(static_cast<float>(4 * i + 3), e3, Wd);
    }
    const float rS = 1.0f / S;
    const float disp = Wd * rS;
    float U = 0.0f;
    const bool need_unc = a.unc_out || a.vote_out;
    if (need_unc) {   // uniform across the grid; recompute, do not keep (see the kernel above)
#pragma unroll
        for (int d = 0; d < DQ; ++d) asm volatile("" : "+f"(v[d]));
#pragma unroll
        for (int i = 0; i < DQ; ++i) {
            const float lo = v[i > 0 ? i - 1 : 0], mid = v[i], hi = v[i + 1 < DQ ? i + 1 : DQ - 1];
            const float dl = mid - lo, dh = hi - mid;
            U = fmaf(fabsf(disp - static_cast<float>(4 * i)), ex2f(fmaf(0.625f, dl, lo)), U);
            U = fmaf(fabsf(disp - static_cast<float>(4 * i + 1)), ex2f(fmaf(0.875f, dl, lo)), U);
            U = fmaf(fabsf(disp - static_cast<float>(4 * i + 2)), ex2f(fmaf(0.125f, dh, mid)), U);
            U = fmaf(fabsf(disp - static_cast<float>(4 * i + 3)), ex2f(fmaf(0.375f, dh, mid)), U);
        }
        U *= rS;
    }
    ups_store(a, (static_cast<int64_t>(b) * a.H + Y) * a.W + X, disp, U, need_unc);
}

// Any (Dq, h, w) -> (D, H, W), both align modes.  CTA = 32 x 4 output pixels; v[Dq] per thread in shared memory
// ([d'][thread], conflict-free), per-d taps in a shared table.
__global__ void __launch_bounds__(128)
upsample_regress_generic_kernel(const UpsArgs a) {
    extern __shared__ float sm[];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    float *sv = sm;                                   // [Dq][128]
    int *ti0 = reinterpret_cast<int *>(sm + a.Dq * 128);   // [D]
    int *ti1 = ti0 + a.D;
    float *tl1 = reinterpret_cast<float *>(ti1 + a.D);
    for (int d = tid; d < a.D; d += 128) {
        const AxisTap t = axis_tap(d, a.Dq, a.D, a.align_corners);
        ti0[d] = t.i0; ti1[d] = t.i1; tl1[d] = t.l1;
    }
    __syncthreads();
    const int X = blockIdx.x * 32 + threadIdx.x, Y = blockIdx.y * 4 + threadIdx.y, b = blockIdx.z;
    if (X >= a.W || Y >= a.H) return;
    const AxisTap ty = axis_tap(Y, a.h, a.H, a.align_corners), tx = axis_tap(X, a.w, a.W, a.align_corners);
    const int hw = a.h * a.w;
    const float *base = a.cost + static_cast<int64_t>(b) * a.Dq * hw;
    const float *p00 = base + ty.i0 * a.w + tx.i0, *p01 = base + ty.i0 * a.w + tx.i1;
    const float *p10 = base + ty.i1 * a.w + tx.i0, *p11 = base + ty.i1 * a.w + tx.i1;
    constexpr float kLog2e = 1.4426950408889634f;
    float m = -INFINITY;
    for (int d = 0; d < a.Dq; ++d) {
        const int64_t off = static_cast<int64_t>(d) * hw;
        const float s00 = __ldg(p00 + off), s01 = __ldg(p01 + off), s10 = __ldg(p10 + off), s11 = __ldg(p11 + off);
        const float r0 = fmaf(tx.l1, s01 - s00, s00), r1 = fmaf(tx.l1, s11 - s10, s10);
        const float v = fmaf(ty.l1, r1 - r0, r0);
        sv[d * 128 + tid] = v;
        m = fmaxf(m, v);
    }
    const float mL = m * kLog2e;
    for (int d = 0; d < a.Dq; ++d) sv[d * 128 + tid] = fmaf(sv[d * 128 + tid], kLog2e, -mL);
    float S = 0.0f, Wd = 0.0f;
    for (int d = 0; d < a.D; ++d) {
        const float lo = sv[ti0[d] * 128 + tid], hi = sv[ti1[d] * 128 + tid];
        const float e = ex2f(fmaf(tl1[d], hi - lo, lo));
        S += e;
        Wd = fmaf(static_cast<float>(d), e, Wd);
    }
    const float rS = 1.0f / S;
    const float disp = Wd * rS;
    float U = 0.0f;
    const bool need_unc = a.unc_out || a.vote_out;
    if (need_unc) {
        for (int d = 0; d < a.D; ++d) {
            const float lo = sv[ti0[d] * 128 + tid], hi = sv[ti1[d] * 128 + tid];
            U = fmaf(fabsf(disp - static_cast<float>(d)), ex2f(fmaf(tl1[d], hi - lo, lo)), U);
        }
        U *= rS;
    }
    ups_store(a, (static_cast<int64_t>(b) * a.H + Y) * a.W + X, disp, U, need_unc);
}

}  // namespace dv

extern "C" int dv_upsample_softmax_regress_f32(const float *cost_q, int64_t B, int64_t Dq, int64_t h, int64_t w,
                                               int64_t D, int64_t H, int64_t W, int align_corners, float *disp_out,
                                               const float *used, float *unc_out, float *vote_out, float thr_dif,
                                               float thr_unc, float *ens_acc, float ens_coef, int ens_init,
                                               void *stream) {
    using namespace dv;
    if (!cost_q) return DV_ERR_NULL;
    if (vote_out && !used) return DV_ERR_NULL;
    if (B <= 0 || Dq <= 0 || h <= 0 || w <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    if (B > 65535 || H * W > INT32_MAX || Dq * h * w > INT32_MAX || D > 4096) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    UpsArgs a;
    a.cost = cost_q;
    a.Dq = static_cast<int>(Dq); a.h = static_cast<int>(h); a.w = static_cast<int>(w);
    a.D = static_cast<int>(D); a.H = static_cast<int>(H); a.W = static_cast<int>(W);
    a.align_corners = align_corners ? 1 : 0;
    a.disp_out = disp_out; a.used = used; a.unc_out = unc_out; a.vote_out = vote_out;
    a.thr_dif = thr_dif; a.thr_unc = thr_unc; a.ens_acc = ens_acc; a.ens_coef = ens_coef; a.ens_init = ens_init;
    if (Dq == 48 && D == 4 * Dq && !align_corners && DV_TUNE("DV_UPS_FAST", 1)) {
        dim3 grid(static_cast<unsigned>((W + 31) / 32), static_cast<unsigned>((H + 7) / 8), static_cast<unsigned>(B));
        // 4 x upsampling exactly (H == 4 h, W == 4 w): the tile footprint is at most 4 x 10 quarter-res pixels
        // (measured at B = 8, 540x960: disparity only 0.316 vs 0.356 ms; with the uncertainty pass the staged variant is
        // slower, 0.501 vs 0.475 ms, so that launch keeps the direct-load kernel)
        if (H == 4 * h && W == 4 * w && !unc_out && !vote_out && DV_TUNE("DV_UPS_TILE", 1))
            upsample_regress_tile_kernel<48><<<grid, dim3(32, 8), 0, st>>>(a);
        else upsample_regress_fast_kernel<48><<<grid, dim3(32, 8), 0, st>>>(a);
    } else {
        const size_t smem = sizeof(float) * (static_cast<size_t>(Dq) * 128 + 3 * static_cast<size_t>(D));
        if (smem > 200 * 1024) return DV_ERR_UNSUPPORTED;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(upsample_regress_generic_kernel,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     static_cast<int>(smem)) != cudaSuccess)
            return DV_ERR_LAUNCH;
        dim3 grid(static_cast<unsigned>((W + 31) / 32), static_cast<unsigned>((H + 3) / 4), static_cast<unsigned>(B));
        upsample_regress_generic_kernel<<<grid, dim3(32, 4), smem, st>>>(a);
    }
    return finish_launch();
}
