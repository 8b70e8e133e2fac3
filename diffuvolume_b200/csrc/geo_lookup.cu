// geo_lookup.cu — IGEV "Combined Geo Encoding Volume": all-pairs 1-D correlation, pyramid
// construction and the radius-r bilinear lookup with the DiffuVolume noise multiply folded in
// (a14, a15) for sm_100a.
//
// Replaces Combined_Geo_Encoding_Volume (KITTI15/core/geometry_ddim.py:7-80, geometry.py:6-68)
// and bilinear_sampler (KITTI15/core/utils/utils.py:59-77).  Per GRU iteration (64 per stereo
// pair) the reference multiplies the whole geo pyramid by the noise pyramid (46+23 MB read and
// written), runs 4 grid_sample launches, a cat and a permute().contiguous().  Here one kernel
// gathers only the <= 2r+2 hypotheses each tap pair needs, multiplies by the noise on the fly
// (same rounding: product first, then the bilinear blend) and writes [B,162,h,w] directly.
#include "common.cuh"

namespace dv {

// ---- a14: all-pairs correlation, one (b, y) row pair per blockIdx.z --------------------------
// out[b,y,x1,x2] = sum_c f1[b,c,y,x1] * f2[b,c,y,x2]; fp32 FMA (TF32 would break the 1e-4 bound).
constexpr int kApTile = 64;  // 64 x 64 output tile, 256 threads, 4 x 4 per thread
constexpr int kApKc = 32;    // channels per shared-memory chunk
__global__ void __launch_bounds__(256)
corr1d_allpairs_kernel(const float *__restrict__ f1, const float *__restrict__ f2, float *__restrict__ out, int C,
                       int H, int W1, int W2) {
    __shared__ __align__(16) float sA[kApKc][kApTile];
    __shared__ __align__(16) float sB[kApKc][kApTile];
    const int by = blockIdx.z;  // b * H + y
    const int b = by / H, y = by % H;
    const int x1_0 = blockIdx.y * kApTile, x2_0 = blockIdx.x * kApTile;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4] = {};
    for (int c0 = 0; c0 < C; c0 += kApKc) {
        for (int e = threadIdx.x; e < kApKc * kApTile; e += 256) {
            const int k = e / kApTile, j = e % kApTile;
            const int c = c0 + k;
            const int x1 = x1_0 + j, x2 = x2_0 + j;
            sA[k][j] = (c < C && x1 < W1) ? f1[((static_cast<int64_t>(b) * C + c) * H + y) * W1 + x1] : 0.0f;
            sB[k][j] = (c < C && x2 < W2) ? f2[((static_cast<int64_t>(b) * C + c) * H + y) * W2 + x2] : 0.0f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kApKc; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&sA[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&sB[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x1 = x1_0 + ty * 4 + i;
        if (x1 >= W1) continue;
        float *op = out + (static_cast<int64_t>(by) * W1 + x1) * W2 + x2_0 + tx * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x2_0 + tx * 4 + j < W2) op[j] = acc[i][j];
    }
}

// ---- geo [B,C,D,h,w] -> rows [B*h*w, C, D]  (permute(0,3,4,1,2), geometry_ddim.py:19) ---------
__global__ void __launch_bounds__(256)
geo_permute_kernel(const float *__restrict__ geo, float *__restrict__ rows, int C, int D, int hw) {
    extern __shared__ float tile[];  // [D][33]
    const int b = blockIdx.z, c = blockIdx.y, p0 = blockIdx.x * 32;
    for (int e = threadIdx.x; e < D * 32; e += 256) {
        const int d = e / 32, pl = e % 32;
        const int p = p0 + pl;
        tile[d * 33 + pl] = p < hw ? geo[((static_cast<int64_t>(b) * C + c) * D + d) * hw + p] : 0.0f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * 32; e += 256) {
        const int pl = e / D, d = e % D;
        const int p = p0 + pl;
        if (p < hw) rows[((static_cast<int64_t>(b) * hw + p) * C + c) * D + d] = tile[d * 33 + pl];
    }
}

// ---- rows [N, L] -> [N, L/2]  (F.avg_pool2d(x, [1,2], stride=[1,2]), geometry_ddim.py:24-30) ---
__global__ void avgpool_w2_kernel(const float *__restrict__ rows, float *__restrict__ pooled, int64_t N, int L) {
    const int Lo = L / 2;
    const int64_t total = N * Lo;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t n = i / Lo;
        const int j = static_cast<int>(i % Lo);
        const float2 v = *reinterpret_cast<const float2 *>(rows + n * L + 2 * j);  // even offset: L*n + 2j
        pooled[i] = __fadd_rn(v.x, v.y) / 2.0f;
    }
}
__global__ void avgpool_w2_kernel_unaligned(const float *__restrict__ rows, float *__restrict__ pooled, int64_t N,
                                            int L) {
    const int Lo = L / 2;
    const int64_t total = N * Lo;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t n = i / Lo;
        const int j = static_cast<int>(i % Lo);
        pooled[i] = __fadd_rn(rows[n * L + 2 * j], rows[n * L + 2 * j + 1]) / 2.0f;
    }
}

// ---- a15: the lookup ---------------------------------------------------------------------------
struct GeoLookupArgs {
    const float *geo[4];   // level i: [N, C, D >> i]
    const float *corr[4];  // level i: [N, W2 >> i]
    const float *noisy;    // [N, D] raw reinterpretation, or NULL
    const float *disp;     // [N]
    const float *coords;   // [N]
    float *out;            // [B, levels*(C+1)*(2r+1), hw]
    int C, D, hw, W2, levels, radius;
    int64_t N;
};

// level-i noise at column j: avg-pooled i times from the raw noisy row
__device__ __forceinline__ float noise_at(const float *__restrict__ row, int level, int j) {
    if (level == 0) return row[j];
    if (level == 1) return __fadd_rn(row[2 * j], row[2 * j + 1]) / 2.0f;
    const float a = noise_at(row, level - 1, 2 * j), b = noise_at(row, level - 1, 2 * j + 1);
    return __fadd_rn(a, b) / 2.0f;
}

// grid_sample(align_corners=True, zero padding) source coordinate for pixel coordinate x on a row of
// width Wl, following bilinear_sampler's normalise (utils.py:64) and ATen's unnormalise.
__device__ __forceinline__ float grid_coord(float x, int Wl) {
    const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, x), static_cast<float>(Wl - 1)), 1.0f);
    return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), static_cast<float>(Wl - 1));
}

__global__ void __launch_bounds__(128)
geo_lookup_kernel(const GeoLookupArgs a) {
    const int64_t n = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (n >= a.N) return;
    const int64_t b = n / a.hw;
    const int p = static_cast<int>(n % a.hw);
    const int taps = 2 * a.radius + 1;
    const int chan_per_level = (a.C + 1) * taps;
    const float disp = a.disp[n], coord = a.coords[n];
    const float *nrow = a.noisy ? a.noisy + n * a.D : nullptr;
    float *obase = a.out + b * static_cast<int64_t>(a.levels) * chan_per_level * a.hw + p;

    float scale = 1.0f;  // 2^i
    for (int lvl = 0; lvl < a.levels; ++lvl, scale *= 2.0f) {
        const int Dl = a.D >> lvl, Wl = a.W2 >> lvl;
        const float dl = disp / scale;
        const float cl = __fsub_rn(coord / scale, dl);
        const float *grow = a.geo[lvl] + n * static_cast<int64_t>(a.C) * Dl;
        const float *crow = a.corr[lvl] + n * static_cast<int64_t>(Wl);
        float *ol = obase + static_cast<int64_t>(lvl) * chan_per_level * a.hw;
        for (int t = 0; t < taps; ++t) {
            const float dx = static_cast<float>(t - a.radius);
            // geo part: x0 = dx + disp / 2^i
            {
                const float ix = grid_coord(__fadd_rn(dx, dl), Dl);
                const float fl = floorf(ix);
                const int i0 = static_cast<int>(fl), i1 = i0 + 1;
                const float w0 = __fsub_rn(static_cast<float>(i1), ix), w1 = __fsub_rn(ix, fl);
                const bool in0 = i0 >= 0 && i0 < Dl, in1 = i1 >= 0 && i1 < Dl;
                float n0 = 1.0f, n1 = 1.0f;
                if (nrow) {
                    n0 = in0 ? noise_at(nrow, lvl, i0) : 0.0f;
                    n1 = in1 ? noise_at(nrow, lvl, i1) : 0.0f;
                }
                for (int c = 0; c < a.C; ++c) {
                    float acc = 0.0f;
                    if (in0) {
                        const float g0 = grow[c * Dl + i0];
                        acc = __fmul_rn(nrow ? __fmul_rn(g0, n0) : g0, w0);
                    }
                    if (in1) {
                        const float g1 = grow[c * Dl + i1];
                        acc = __fadd_rn(acc, __fmul_rn(nrow ? __fmul_rn(g1, n1) : g1, w1));
                    }
                    ol[static_cast<int64_t>(c * taps + t) * a.hw] = acc;
                }
            }
            // corr part: init_x0 = coords / 2^i - disp / 2^i + dx
            {
                const float ix = grid_coord(__fadd_rn(cl, dx), Wl);
                const float fl = floorf(ix);
                const int i0 = static_cast<int>(fl), i1 = i0 + 1;
                const float w0 = __fsub_rn(static_cast<float>(i1), ix), w1 = __fsub_rn(ix, fl);
                float acc = 0.0f;
                if (i0 >= 0 && i0 < Wl) acc = __fmul_rn(crow[i0], w0);
                if (i1 >= 0 && i1 < Wl) acc = __fadd_rn(acc, __fmul_rn(crow[i1], w1));
                ol[static_cast<int64_t>(a.C * taps + t) * a.hw] = acc;
            }
        }
    }
}

}  // namespace dv

extern "C" int dv_corr1d_allpairs_f32(const float *fmap1, const float *fmap2, float *out, int64_t B, int64_t C,
                                      int64_t H, int64_t W1, int64_t W2, void *stream) {
    using namespace dv;
    if (!fmap1 || !fmap2 || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W1 <= 0 || W2 <= 0 || B * H > INT32_MAX) return DV_ERR_BAD_SHAPE;
    dim3 grid(static_cast<unsigned>((W2 + kApTile - 1) / kApTile), static_cast<unsigned>((W1 + kApTile - 1) / kApTile),
              static_cast<unsigned>(B * H));
    if (grid.y > 65535 || B * H > 65535 * 1LL) {
        // z is limited to 65535: fold (b, y) planes in chunks
        for (int64_t z0 = 0; z0 < B * H; z0 += 65535) {
            // not reached for any reference configuration (B*H = 96 per pair); keep it simple
            (void)z0;
        }
        return DV_ERR_UNSUPPORTED;
    }
    corr1d_allpairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        fmap1, fmap2, out, static_cast<int>(C), static_cast<int>(H), static_cast<int>(W1), static_cast<int>(W2));
    return finish_launch();
}

extern "C" int dv_geo_permute_f32(const float *geo, float *rows, int64_t B, int64_t C, int64_t D, int64_t h, int64_t w,
                                  void *stream) {
    using namespace dv;
    if (!geo || !rows) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0 || B > 65535 || C > 65535 || h * w > INT32_MAX || D > 1024)
        return DV_ERR_BAD_SHAPE;
    const int hw = static_cast<int>(h * w);
    dim3 grid((hw + 31) / 32, static_cast<unsigned>(C), static_cast<unsigned>(B));
    geo_permute_kernel<<<grid, 256, sizeof(float) * D * 33, static_cast<cudaStream_t>(stream)>>>(
        geo, rows, static_cast<int>(C), static_cast<int>(D), hw);
    return finish_launch();
}

extern "C" int dv_avgpool_w2_f32(const float *rows, float *pooled, int64_t N, int64_t L, void *stream) {
    using namespace dv;
    if (!rows || !pooled) return DV_ERR_NULL;
    if (N <= 0 || L < 2 || L > INT32_MAX) return DV_ERR_BAD_SHAPE;
    const int64_t total = N * (L / 2);
    const int64_t blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(kNumSMs) * 16 ? blocks : static_cast<int64_t>(kNumSMs) * 16);
    if (L % 2 == 0 && (reinterpret_cast<uintptr_t>(rows) & 7u) == 0)
        avgpool_w2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, pooled, N, static_cast<int>(L));
    else
        avgpool_w2_kernel_unaligned<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(rows, pooled, N, static_cast<int>(L));
    return finish_launch();
}

extern "C" int dv_geo_lookup_f32(const float *const *geo_pyr, const float *const *corr_pyr, const float *noisy,
                                 const float *disp, const float *coords, float *out, int64_t B, int64_t C, int64_t D,
                                 int64_t h, int64_t w, int64_t W2, int num_levels, int radius, void *stream) {
    using namespace dv;
    if (!geo_pyr || !corr_pyr || !disp || !coords || !out) return DV_ERR_NULL;
    if (num_levels < 1 || num_levels > 4 || radius < 0 || radius > 16) return DV_ERR_UNSUPPORTED;
    if (B <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0 || W2 <= 0 || h * w > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if ((D >> (num_levels - 1)) < 2 || (W2 >> (num_levels - 1)) < 2) return DV_ERR_BAD_SHAPE;
    GeoLookupArgs a;
    for (int i = 0; i < 4; ++i) {
        a.geo[i] = i < num_levels ? geo_pyr[i] : nullptr;
        a.corr[i] = i < num_levels ? corr_pyr[i] : nullptr;
        if (i < num_levels && (!a.geo[i] || !a.corr[i])) return DV_ERR_NULL;
    }
    a.noisy = noisy; a.disp = disp; a.coords = coords; a.out = out;
    a.C = static_cast<int>(C); a.D = static_cast<int>(D); a.hw = static_cast<int>(h * w); a.W2 = static_cast<int>(W2);
    a.levels = num_levels; a.radius = radius; a.N = B * h * w;
    const int64_t blocks = (a.N + 127) / 128;
    if (blocks > INT32_MAX) return DV_ERR_BAD_SHAPE;
    geo_lookup_kernel<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return finish_launch();
}
