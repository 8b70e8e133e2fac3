// volume_filter.cu — the DiffuVolume filter multiply at its reference op boundary (a9).
//
// Replaces the head of model_predictions (SceneFlow/models/acv_ddim.py:254-260,
// KITTI12/models/pwcnet_ddim.py:466-472):
//     noise = time_embedding(noise, t)            # + shift[b,d]   (head.py:74-77)
//     noise = clamp(noise, -s, s); noise = ((noise / s) + 1) / 2
//     volume = volume * noise.unsqueeze(1).float()
// which the reference runs as 6 small elementwise kernels plus one volume-sized multiply.
// Here it is one pass: thread = (quad of pixels, d); the factor n is computed once in the
// dtype of x_t (fp32 on the first DDIM step, fp64 afterwards), converted to float, kept in
// registers and applied to all C channels (128-bit streaming loads and stores, UNROLL
// independent loads in flight per thread).
#include "common.cuh"

namespace dv {

template <typename XT, int UNROLL>
__global__ void __launch_bounds__(256)
volume_filter_kernel(const float *__restrict__ vol, float *__restrict__ out, int C, int D, int HW4,
                     const XT *__restrict__ xt, const float *__restrict__ shift, XT scale, XT *__restrict__ n_out) {
    // grid: x = quads of (d, p) flattened (D*HW4), y = batch
    const int b = blockIdx.y;
    const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;  // quad index in [0, D*HW4)
    if (e >= static_cast<int64_t>(D) * HW4) return;
    const int d = static_cast<int>(e / HW4);
    const float sh = shift ? shift[b * D + d] : 0.0f;
    const int64_t plane_q = static_cast<int64_t>(D) * HW4;  // quads per channel
    const XT *xp = xt + (static_cast<int64_t>(b) * plane_q + e) * 4;
    XT nx[4];
    float n[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        nx[i] = filter_n<XT>(xp[i], sh, scale);
        n[i] = static_cast<float>(nx[i]);
    }
    if (n_out) {
        XT *np = n_out + (static_cast<int64_t>(b) * plane_q + e) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) np[i] = nx[i];
    }
    const float4 *vp = reinterpret_cast<const float4 *>(vol) + static_cast<int64_t>(b) * C * plane_q + e;
    float4 *op = reinterpret_cast<float4 *>(out) + static_cast<int64_t>(b) * C * plane_q + e;
    int c = 0;
    for (; c + UNROLL <= C; c += UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(vp + static_cast<int64_t>(c + u) * plane_q);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            v[u].x *= n[0]; v[u].y *= n[1]; v[u].z *= n[2]; v[u].w *= n[3];
            stg_cs(op + static_cast<int64_t>(c + u) * plane_q, v[u]);
        }
    }
    for (; c < C; ++c) {
        float4 v = ldg_stream(vp + static_cast<int64_t>(c) * plane_q);
        v.x *= n[0]; v.y *= n[1]; v.z *= n[2]; v.w *= n[3];
        stg_cs(op + static_cast<int64_t>(c) * plane_q, v);
    }
}

template <typename XT>
__global__ void volume_filter_generic_kernel(const float *__restrict__ vol, float *__restrict__ out, int C, int D,
                                             int HW, const XT *__restrict__ xt, const float *__restrict__ shift,
                                             XT scale, XT *__restrict__ n_out, int64_t total_bdp) {
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total_bdp;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t dp = idx % (static_cast<int64_t>(D) * HW);
        const int64_t b = idx / (static_cast<int64_t>(D) * HW);
        const int d = static_cast<int>(dp / HW);
        const float sh = shift ? shift[b * D + d] : 0.0f;
        const XT nx = filter_n<XT>(xt[idx], sh, scale);
        if (n_out) n_out[idx] = nx;
        const float n = static_cast<float>(nx);
        for (int c = 0; c < C; ++c) {
            const int64_t o = (b * C + c) * static_cast<int64_t>(D) * HW + dp;
            out[o] = vol[o] * n;
        }
    }
}

template <typename XT>
static int launch_filter(const float *vol, float *out, int64_t B, int64_t C, int64_t D, int64_t HW, const XT *xt,
                         const float *shift, XT scale, XT *n_out, cudaStream_t st) {
    const bool fast = (HW % 4 == 0) && aligned16(vol) && aligned16(out) && aligned16(xt) && (!n_out || aligned16(n_out));
    if (fast) {
        const int64_t quads = D * (HW / 4);
        dim3 grid(static_cast<unsigned>((quads + 255) / 256), static_cast<unsigned>(B));
        volume_filter_kernel<XT, 8><<<grid, 256, 0, st>>>(vol, out, static_cast<int>(C), static_cast<int>(D),
                                                          static_cast<int>(HW / 4), xt, shift, scale, n_out);
    } else {
        const int64_t total = B * D * HW;
        const int64_t blocks = (total + 255) / 256;
        const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
        volume_filter_generic_kernel<XT><<<grid, 256, 0, st>>>(vol, out, static_cast<int>(C), static_cast<int>(D),
                                                               static_cast<int>(HW), xt, shift, scale, n_out, total);
    }
    return finish_launch();
}

}  // namespace dv

extern "C" int dv_volume_filter_f32(const float *vol, float *out, int64_t B, int64_t C, int64_t D, int64_t H, int64_t W,
                                    const void *xt, int xt_is_f64, const float *shift, double scale, void *n_out,
                                    void *stream) {
    using namespace dv;
    if (!vol || !out || !xt) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || !(scale > 0.0)) return DV_ERR_BAD_SHAPE;
    if (B > 65535 || D * H * W > INT32_MAX) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (xt_is_f64 == 1)
        return launch_filter<double>(vol, out, B, C, D, H * W, static_cast<const double *>(xt), shift, scale,
                                     static_cast<double *>(n_out), st);
    if (xt_is_f64 == 0)
        return launch_filter<float>(vol, out, B, C, D, H * W, static_cast<const float *>(xt), shift,
                                    static_cast<float>(scale), static_cast<float *>(n_out), st);
    return DV_ERR_BAD_DTYPE;
}
