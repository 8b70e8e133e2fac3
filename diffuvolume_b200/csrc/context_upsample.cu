// context_upsample.cu — IGEV's 9-tap convex upsampling of the quarter-resolution disparity.
//
// Replaces context_upsample(disp_low, up_weights) of KITTI15/core/submodule.py:241-253 (call sites
// igev_stereo_ddim.py:209,462; igev_stereo.py:146,220).  The reference unfolds the [B,1,h,w] map into its 3x3
// neighbourhood ([B,9,h,w]), nearest-upsamples that x4 ([B,9,4h,4w], 9x the output size, materialised), multiplies by
// the weights and sums over the 9 taps: four launches and ~28 output-sized tensors of traffic for an op whose
// compulsory traffic is the 9 weight planes in and one plane out.  Here: one pass, thread = 4 horizontally adjacent
// full-resolution pixels (they share one low-resolution neighbourhood), 9 x 128-bit weight loads, one 128-bit store.
//   out[b, Y, X] = sum_{k = ky*3+kx} disp_low[b, Y/4 + ky - 1, X/4 + kx - 1] * w[b, k, Y, X]      (zero padding)
// The 9 products are accumulated in tap order k = 0..8 in fp32.
#include "common.cuh"

namespace dv {

__global__ void __launch_bounds__(256)
context_upsample_kernel(const float *__restrict__ low, const float *__restrict__ wts, float *__restrict__ out, int h, int w) {
    const int W = 4 * w, H = 4 * h;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;  // low-res column == quad index along X
    const int Y = blockIdx.y, b = blockIdx.z;
    if (x >= w) return;
    const int y = Y >> 2;
    const float *lp = low + static_cast<int64_t>(b) * h * w;
    float nb[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            nb[ky * 3 + kx] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(lp + static_cast<int64_t>(yy) * w + xx) : 0.0f;
        }
    const int64_t HW = static_cast<int64_t>(H) * W;
    const int64_t o = static_cast<int64_t>(Y) * W + 4 * x;
    const float *wp = wts + static_cast<int64_t>(b) * 9 * HW + o;
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 wv = ldg_stream(reinterpret_cast<const float4 *>(wp + k * HW));
        acc.x = __fadd_rn(acc.x, __fmul_rn(nb[k], wv.x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(nb[k], wv.y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(nb[k], wv.z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(nb[k], wv.w));
    }
    *reinterpret_cast<float4 *>(out + static_cast<int64_t>(b) * HW + o) = acc;
}

// backward: d_w[b,k,Y,X] = g[b,Y,X] * nb_k(Y/4, X/4)  (one pass, same thread mapping);
//           d_low[b,y,x] = sum_k sum_{(Y,X) in the 4x4 block of low pixel (y - ky + 1, x - kx + 1)} g * w_k
__global__ void __launch_bounds__(256)
context_upsample_bwd_w_kernel(const float *__restrict__ g, const float *__restrict__ low, float *__restrict__ gw, int h,
                              int w) {
    const int W = 4 * w, H = 4 * h;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = blockIdx.y, b = blockIdx.z;
    if (x >= w) return;
    const int y = Y >> 2;
    const float *lp = low + static_cast<int64_t>(b) * h * w;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const int64_t o = static_cast<int64_t>(Y) * W + 4 * x;
    const float4 gv = *reinterpret_cast<const float4 *>(g + static_cast<int64_t>(b) * HW + o);
    float *gp = gw + static_cast<int64_t>(b) * 9 * HW + o;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            const float v = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(lp + static_cast<int64_t>(yy) * w + xx) : 0.0f;
            stg_cs(reinterpret_cast<float4 *>(gp + (ky * 3 + kx) * HW), make_float4(gv.x * v, gv.y * v, gv.z * v, gv.w * v));
        }
}
__global__ void __launch_bounds__(128)
context_upsample_bwd_low_kernel(const float *__restrict__ g, const float *__restrict__ wts, float *__restrict__ glow, int h,
                                int w) {
    const int W = 4 * w, H = 4 * h;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, b = blockIdx.z;
    if (x >= w) return;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const float *gb = g + static_cast<int64_t>(b) * HW;
    const float *wb = wts + static_cast<int64_t>(b) * 9 * HW;
    float acc = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            // low pixel (y, x) is tap (ky, kx) of the full-res block owned by low pixel (y - ky + 1, x - kx + 1)
            const int yo = y - ky + 1, xo = x - kx + 1;
            if (yo < 0 || yo >= h || xo < 0 || xo >= w) continue;
            const float *wk = wb + (ky * 3 + kx) * HW;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int64_t o = static_cast<int64_t>(4 * yo + r) * W + 4 * xo;
                const float4 gv = __ldg(reinterpret_cast<const float4 *>(gb + o));
                const float4 wv = __ldg(reinterpret_cast<const float4 *>(wk + o));
                acc = fmaf(gv.x, wv.x, acc); acc = fmaf(gv.y, wv.y, acc);
                acc = fmaf(gv.z, wv.z, acc); acc = fmaf(gv.w, wv.w, acc);
            }
        }
    glow[(static_cast<int64_t>(b) * h + y) * w + x] = acc;
}

}  // namespace dv

extern "C" int dv_context_upsample_f32(const float *disp_low, const float *up_weights, float *out, int64_t B, int64_t h,
                                       int64_t w, void *stream) {
    using namespace dv;
    if (!disp_low || !up_weights || !out) return DV_ERR_NULL;
    if (B <= 0 || h <= 0 || w <= 0 || B > 65535 || 4 * h > 65535 || 16 * h * w > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if (!aligned16(up_weights) || !aligned16(out)) return DV_ERR_MISALIGNED;
    dim3 grid(static_cast<unsigned>((w + 255) / 256), static_cast<unsigned>(4 * h), static_cast<unsigned>(B));
    context_upsample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(disp_low, up_weights, out,
                                                                                static_cast<int>(h), static_cast<int>(w));
    return finish_launch();
}

extern "C" int dv_context_upsample_bwd_f32(const float *grad_out, const float *disp_low, const float *up_weights,
                                           float *grad_low, float *grad_weights, int64_t B, int64_t h, int64_t w,
                                           void *stream) {
    using namespace dv;
    if (!grad_out || !disp_low || !up_weights) return DV_ERR_NULL;
    if (!grad_low && !grad_weights) return DV_ERR_NULL;
    if (B <= 0 || h <= 0 || w <= 0 || B > 65535 || 4 * h > 65535 || 16 * h * w > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if (!aligned16(up_weights) || !aligned16(grad_out) || (grad_weights && !aligned16(grad_weights))) return DV_ERR_MISALIGNED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int n = 0;
    if (grad_weights) {
        dim3 grid(static_cast<unsigned>((w + 255) / 256), static_cast<unsigned>(4 * h), static_cast<unsigned>(B));
        context_upsample_bwd_w_kernel<<<grid, 256, 0, st>>>(grad_out, disp_low, grad_weights, static_cast<int>(h),
                                                            static_cast<int>(w));
        ++n;
    }
    if (grad_low) {
        dim3 grid(static_cast<unsigned>((w + 127) / 128), static_cast<unsigned>(h), static_cast<unsigned>(B));
        context_upsample_bwd_low_kernel<<<grid, 128, 0, st>>>(grad_out, up_weights, grad_low, static_cast<int>(h),
                                                              static_cast<int>(w));
        ++n;
    }
    return finish_launch(n);
}
