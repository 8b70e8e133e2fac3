// fused_backward.cu — backward passes of the FUSED forward ops (second half) for sm_100a:
//
//   * softmax over D + disparity regression (F.softmax(cost, 1) + disparity_regression, SceneFlow/models/acv_ddim.py:460-480,
//     SceneFlow/main.py:154): d_cost[b,d,p] = softmax_d(cost)[b,d,p] * (d - disp[b,p]) * g[b,p], the probability volume is
//     neither saved by the forward nor materialised here (autograd of the two separate ops saves the 398 MB softmax and
//     makes four volume-sized passes);
//   * the ACV attention volume with the optional DDIM filter factor, out = (concat(cl, cr) * softmax_d(att)) * n
//     (acv_ddim.py:388-390 and, in the training branch, :446-451): gradients with respect to the concat features AND the
//     attention logits from two passes over grad_out (autograd: softmax backward + mul backward + the D slice-assign
//     backwards of build_concat_volume, six volume-sized passes and a saved 398 MB concat volume).
//
// The gradient of the filter multiply itself (vol * n, acv_ddim.py:260 / pwcnet_ddim.py:472) with respect to the volume
// is the forward kernel applied to grad_out (dv_volume_filter_f32); n carries no gradient in the reference (it is
// re-wrapped with torch.tensor(), acv_ddim.py:449).
#include "common.cuh"

namespace dv {

// -------------------------------------------------------------------------------------------------------------------
// softmax + regression backward.  CTA = 32 pixel vectors x SL disparity slices; thread (lane, slice) keeps its DPT logits
// (d = j*SL + slice) in registers: cost is read from HBM once, d_cost written once.
// -------------------------------------------------------------------------------------------------------------------
template <int DPT, int V, int SL>
__global__ void __launch_bounds__(SL * 32)
softmax_regress_bwd_kernel(const float *__restrict__ cost, const float *__restrict__ gdisp, float *__restrict__ gcost, int D,
                           int HW) {
    __shared__ float red[3][SL][32][V];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int64_t pv = blockIdx.x * 32ll + lane;
    const bool live = pv * V < HW;
    const int64_t base = static_cast<int64_t>(b) * D * HW + pv * V;
    constexpr float kLog2e = 1.4426950408889634f;
    float x[DPT][V];
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = j * SL + slice;
        if (live && d < D) {
            if constexpr (V == 4) {
                const float4 t = ldg_stream(reinterpret_cast<const float4 *>(cost + base + static_cast<int64_t>(d) * HW));
                x[j][0] = t.x; x[j][1] = t.y; x[j][2] = t.z; x[j][3] = t.w;
            } else {
                x[j][0] = ldg_stream_f32(cost + base + static_cast<int64_t>(d) * HW);
            }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) x[j][i] = -INFINITY;
        }
    }
    float m[V];
#pragma unroll
    for (int i = 0; i < V; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int i = 0; i < V; ++i) m[i] = fmaxf(m[i], x[j][i]);
#pragma unroll
    for (int i = 0; i < V; ++i) red[0][slice][lane][i] = m[i];
    __syncthreads();
#pragma unroll
    for (int s = 0; s < SL; ++s)
#pragma unroll
        for (int i = 0; i < V; ++i) m[i] = fmaxf(m[i], red[0][s][lane][i]);
    float S[V], Wd[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        m[i] = live ? m[i] : 0.0f;
        S[i] = 0.0f;
        Wd[i] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const float df = static_cast<float>(j * SL + slice);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float e = exp2f((x[j][i] - m[i]) * kLog2e);   // exp2(-inf) = 0 for d >= D
            x[j][i] = e;
            S[i] += e;
            Wd[i] = fmaf(df, e, Wd[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        red[1][slice][lane][i] = S[i];
        red[2][slice][lane][i] = Wd[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) {
        S[i] = 0.0f;
        Wd[i] = 0.0f;
    }
#pragma unroll
    for (int s = 0; s < SL; ++s)     // fixed order: deterministic
#pragma unroll
        for (int i = 0; i < V; ++i) {
            S[i] += red[1][s][lane][i];
            Wd[i] += red[2][s][lane][i];
        }
    if (!live) return;
    float g[V], rS[V], disp[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        g[i] = gdisp[static_cast<int64_t>(b) * HW + pv * V + i];
        rS[i] = 1.0f / S[i];
        disp[i] = Wd[i] * rS[i];
        g[i] *= rS[i];
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = j * SL + slice;
        if (d >= D) continue;
        const float df = static_cast<float>(d);
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = x[j][i] * (df - disp[i]) * g[i];
        float *op = gcost + base + static_cast<int64_t>(d) * HW;
        if constexpr (V == 4) stg_cs(reinterpret_cast<float4 *>(op), make_float4(o[0], o[1], o[2], o[3]));
        else op[0] = o[0];
    }
}

// any D: one thread per pixel, three passes (re-reads hit L2 for reasonable D)
__global__ void softmax_regress_bwd_generic_kernel(const float *__restrict__ cost, const float *__restrict__ gdisp,
                                                   float *__restrict__ gcost, int D, int HW, int64_t total) {
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
        const float rS = 1.0f / S, disp = Wd * rS, g = gdisp[idx] * rS;
        float *op = gcost + b * D * HW + p;
        for (int d = 0; d < D; ++d)
            op[static_cast<int64_t>(d) * HW] = expf(cp[static_cast<int64_t>(d) * HW] - m) * (static_cast<float>(d) - disp) * g;
    }
}

// -------------------------------------------------------------------------------------------------------------------
// ACV volume backward, pass 1: gradients of the concat features.
//   out[b,c,d,p]   = cl[b,c,p]   * f[b,d,p]      (c < C; variant T: only x >= d)
//   out[b,C+c,d,p] = cr[b,c,p-d] * f[b,d,p]      (x >= d)              f = w * n   (a NULL factor is 1)
//   d_cl[b,c,p] = sum_d g[b,c,d,p] f[b,d,p]          d_cr[b,c,q] = sum_{d : x(q)+d < W} g[b,C+c,d,q+d] f[b,d,q+d]
// thread = (b, channel of the 2C, pixel); the factor maps (6 MB per pair, shared by all 2C channels) stay in L2.
// -------------------------------------------------------------------------------------------------------------------
template <bool HAS_W, bool HAS_N>
__global__ void __launch_bounds__(256)
acv_bwd_features_kernel(const float *__restrict__ go, const float *__restrict__ w, const float *__restrict__ n,
                        float *__restrict__ gcl, float *__restrict__ gcr, int C, int HW, int W, int D, int mask_left) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int c = blockIdx.y, b = blockIdx.z;      // c in [0, 2C)
    const int x = p % W;
    const float *gp = go + ((static_cast<int64_t>(b) * 2 * C + c) * D) * HW;
    const float *wp = w + static_cast<int64_t>(b) * D * HW;
    const float *np_ = n + static_cast<int64_t>(b) * D * HW;
    float acc = 0.0f;
    const bool left = c < C;
    if (left ? !gcl : !gcr) return;
    const int dmax = left ? (mask_left ? min(D - 1, x) : D - 1) : min(D - 1, W - 1 - x);
#pragma unroll 4
    for (int d = 0; d <= dmax; ++d) {
        const int64_t o = static_cast<int64_t>(d) * HW + p + (left ? 0 : d);
        float f = 1.0f;
        if (HAS_W) f = __ldg(wp + o);
        if (HAS_N) f *= __ldg(np_ + o);
        acc = fmaf(ldg_stream_f32(gp + o), f, acc);
    }
    if (left) gcl[(static_cast<int64_t>(b) * C + c) * HW + p] = acc;
    else gcr[(static_cast<int64_t>(b) * C + (c - C)) * HW + p] = acc;
}

// -------------------------------------------------------------------------------------------------------------------
// ACV volume backward, pass 2: gradient of the attention logits.
//   s[d,p]  = sum_c g[c,d,p] cl[c,p] [variant T: x >= d] + sum_c g[C+c,d,p] cr[c,p-d] [x >= d]       (= d out / d f)
//   dw[d,p] = n[d,p] s[d,p]           d_att[d,p] = w[d,p] (dw[d,p] - sum_d' w[d',p] dw[d',p])         (softmax backward)
// CTA = 32 pixels x 16 disparity slices; thread (lane, slice) owns d = slice + 16 k, k < DPT; the dot product over D is
// combined through shared memory in a fixed order.
// -------------------------------------------------------------------------------------------------------------------
template <int DPT>
__global__ void __launch_bounds__(512)
acv_bwd_att_kernel(const float *__restrict__ go, const float *__restrict__ cl, const float *__restrict__ cr,
                   const float *__restrict__ w, const float *__restrict__ n, float *__restrict__ gatt, int C, int HW, int W,
                   int D, int mask_left) {
    __shared__ float red[16][32];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane, b = blockIdx.y;
    const bool live = p < HW;
    const int x = live ? p % W : 0;
    float s[DPT];
#pragma unroll
    for (int k = 0; k < DPT; ++k) s[k] = 0.0f;
    const float *gb = go + static_cast<int64_t>(b) * 2 * C * D * HW + p;
    if (live) {
        for (int c = 0; c < C; ++c) {
            const float a = __ldg(cl + (static_cast<int64_t>(b) * C + c) * HW + p);
            const float *gc = gb + static_cast<int64_t>(c) * D * HW;
#pragma unroll
            for (int k = 0; k < DPT; ++k) {
                const int d = slice + 16 * k;
                if (d < D && (!mask_left || x >= d)) s[k] = fmaf(ldg_stream_f32(gc + static_cast<int64_t>(d) * HW), a, s[k]);
            }
        }
        for (int c = 0; c < C; ++c) {
            const float *rp = cr + (static_cast<int64_t>(b) * C + c) * HW + p;
            const float *gc = gb + static_cast<int64_t>(C + c) * D * HW;
#pragma unroll
            for (int k = 0; k < DPT; ++k) {
                const int d = slice + 16 * k;
                if (d < D && x >= d) s[k] = fmaf(ldg_stream_f32(gc + static_cast<int64_t>(d) * HW), __ldg(rp - d), s[k]);
            }
        }
    }
    float wv[DPT], dw[DPT], part = 0.0f;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int d = slice + 16 * k;
        wv[k] = 0.0f;
        dw[k] = 0.0f;
        if (live && d < D) {
            const int64_t o = (static_cast<int64_t>(b) * D + d) * HW + p;
            wv[k] = w[o];
            dw[k] = n ? s[k] * n[o] : s[k];
            part = fmaf(wv[k], dw[k], part);
        }
    }
    red[slice][lane] = part;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int q = 0; q < 16; ++q) tot += red[q][lane];
    if (!live) return;
#pragma unroll
    for (int k = 0; k < DPT; ++k) {
        const int d = slice + 16 * k;
        if (d < D) gatt[(static_cast<int64_t>(b) * D + d) * HW + p] = wv[k] * (dw[k] - tot);
    }
}

}  // namespace dv

extern "C" int dv_softmax_regress_bwd_f32(const float *cost, const float *grad_disp, float *grad_cost, int64_t B, int64_t D,
                                          int64_t H, int64_t W, void *stream) {
    using namespace dv;
    if (!cost || !grad_disp || !grad_cost) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || D > INT32_MAX) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (HW % 4 == 0) && aligned16(cost) && aligned16(grad_cost) && aligned16(grad_disp);
    const int iD = static_cast<int>(D), iHW = static_cast<int>(HW);
#define DV_SRB(DPT, V)                                                                                          \
    {                                                                                                           \
        const int64_t pvs = (HW + V - 1) / V;                                                                   \
        dim3 grid(static_cast<unsigned>((pvs + 31) / 32), static_cast<unsigned>(B));                            \
        softmax_regress_bwd_kernel<DPT, V, 8><<<grid, 256, 0, st>>>(cost, grad_disp, grad_cost, iD, iHW);       \
    }
    if (D <= 48) {
        if (vec4) DV_SRB(6, 4) else DV_SRB(6, 1)
    } else if (D <= 96) {
        if (vec4) DV_SRB(12, 4) else DV_SRB(12, 1)
    } else if (D <= 192) {
        if (vec4) DV_SRB(24, 4) else DV_SRB(24, 1)
    } else {
        const int64_t total = B * HW;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
        softmax_regress_bwd_generic_kernel<<<grid, 256, 0, st>>>(cost, grad_disp, grad_cost, iD, iHW, total);
    }
#undef DV_SRB
    return finish_launch();
}

extern "C" int dv_acv_volume_bwd_f32(const float *grad_out, const float *cl, const float *cr, const float *att_weights,
                                     const float *n, float *grad_cl, float *grad_cr, float *grad_att_logits, int64_t B,
                                     int64_t C, int64_t H, int64_t W, int64_t D, int mask_left, void *stream) {
    using namespace dv;
    if (!grad_out || (!grad_cl && !grad_cr && !grad_att_logits)) return DV_ERR_NULL;
    if (grad_att_logits && (!att_weights || !cl || !cr)) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || 2 * C > 65535) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int launches = 0;
    if (grad_cl || grad_cr) {
        dim3 grid(static_cast<unsigned>((HW + 255) / 256), static_cast<unsigned>(2 * C), static_cast<unsigned>(B));
#define DV_AF(HW_, HN_)                                                                                                  \
    acv_bwd_features_kernel<HW_, HN_><<<grid, 256, 0, st>>>(grad_out, att_weights, n, grad_cl, grad_cr, static_cast<int>(C), \
                                                            static_cast<int>(HW), static_cast<int>(W), static_cast<int>(D), \
                                                            mask_left)
        if (att_weights && n) DV_AF(true, true);
        else if (att_weights) DV_AF(true, false);
        else if (n) DV_AF(false, true);
        else DV_AF(false, false);
#undef DV_AF
        ++launches;
    }
    if (grad_att_logits) {
        if (D > 16 * 12) return DV_ERR_UNSUPPORTED;
        dim3 grid(static_cast<unsigned>((HW + 31) / 32), static_cast<unsigned>(B));
#define DV_AA(DPT)                                                                                                     \
    acv_bwd_att_kernel<DPT><<<grid, 512, 0, st>>>(grad_out, cl, cr, att_weights, n, grad_att_logits, static_cast<int>(C), \
                                                  static_cast<int>(HW), static_cast<int>(W), static_cast<int>(D), mask_left)
        if (D <= 16) DV_AA(1);
        else if (D <= 48) DV_AA(3);
        else if (D <= 96) DV_AA(6);
        else DV_AA(12);
#undef DV_AA
        ++launches;
    }
    return finish_launch(launches);
}
