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
// 64 x 64 output tile per CTA, 64 threads, 8 x 8 register tile per thread: per channel a thread reads 4 LDS.128 (two for
// its 8 rows, two for its 8 columns) for 64 FMA — the 4 x 4 tile of the first version (2 LDS.128 per 16 FMA) left the
// FMA pipe waiting on shared memory (ncu: 66 % issue slots, 98 M shared wavefronts).  A thread's rows/columns are the two
// float4 at 4*t and 32 + 4*t, so the 8 lanes that differ in tx read 32 consecutive floats: conflict-free.
constexpr int kApTile = 64;  // 64 x 64 output tile
constexpr int kApKc = 32;    // channels per shared-memory chunk
__global__ void __launch_bounds__(64)
corr1d_allpairs_kernel(const float *__restrict__ f1, const float *__restrict__ f2, float *__restrict__ out, int C,
                       int H, int W1, int W2) {
    __shared__ __align__(16) float sA[kApKc][kApTile];
    __shared__ __align__(16) float sB[kApKc][kApTile];
    const int by = blockIdx.z;  // b * H + y
    const int b = by / H, y = by % H;
    const int x1_0 = blockIdx.y * kApTile, x2_0 = blockIdx.x * kApTile;
    const int tx = threadIdx.x % 8, ty = threadIdx.x / 8;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    const float *a0 = f1 + (static_cast<int64_t>(b) * C * H + y) * W1;  // + c*H*W1 + x
    const float *b0 = f2 + (static_cast<int64_t>(b) * C * H + y) * W2;
    const bool vecA = (W1 % 4 == 0) && ((reinterpret_cast<uintptr_t>(f1) & 15u) == 0) && x1_0 + kApTile <= W1;
    const bool vecB = (W2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(f2) & 15u) == 0) && x2_0 + kApTile <= W2;
    for (int c0 = 0; c0 < C; c0 += kApKc) {
        // 32 x 64 floats per operand: thread t loads the float4 (k = e / 16, j4 = e % 16), e = t, t + 64, ...
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = threadIdx.x + 64 * u;
            const int k = e / 16, j = (e % 16) * 4;
            const int c = c0 + k;
            float4 va = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vb = va;
            if (c < C) {
                const float *pa = a0 + static_cast<int64_t>(c) * H * W1 + x1_0 + j;
                const float *pb = b0 + static_cast<int64_t>(c) * H * W2 + x2_0 + j;
                if (vecA) va = __ldg(reinterpret_cast<const float4 *>(pa));
                else {
                    va.x = x1_0 + j + 0 < W1 ? __ldg(pa + 0) : 0.0f; va.y = x1_0 + j + 1 < W1 ? __ldg(pa + 1) : 0.0f;
                    va.z = x1_0 + j + 2 < W1 ? __ldg(pa + 2) : 0.0f; va.w = x1_0 + j + 3 < W1 ? __ldg(pa + 3) : 0.0f;
                }
                if (vecB) vb = __ldg(reinterpret_cast<const float4 *>(pb));
                else {
                    vb.x = x2_0 + j + 0 < W2 ? __ldg(pb + 0) : 0.0f; vb.y = x2_0 + j + 1 < W2 ? __ldg(pb + 1) : 0.0f;
                    vb.z = x2_0 + j + 2 < W2 ? __ldg(pb + 2) : 0.0f; vb.w = x2_0 + j + 3 < W2 ? __ldg(pb + 3) : 0.0f;
                }
            }
            *reinterpret_cast<float4 *>(&sA[k][j]) = va;
            *reinterpret_cast<float4 *>(&sB[k][j]) = vb;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kApKc; ++k) {
            const float4 a_lo = *reinterpret_cast<const float4 *>(&sA[k][4 * ty]);
            const float4 a_hi = *reinterpret_cast<const float4 *>(&sA[k][32 + 4 * ty]);
            const float4 b_lo = *reinterpret_cast<const float4 *>(&sB[k][4 * tx]);
            const float4 b_hi = *reinterpret_cast<const float4 *>(&sB[k][32 + 4 * tx]);
            const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
            const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool vecO = (W2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int x1 = x1_0 + (i < 4 ? 4 * ty + i : 32 + 4 * ty + i - 4);
        if (x1 >= W1) continue;
        float *orow = out + (static_cast<int64_t>(by) * W1 + x1) * W2;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int x2 = x2_0 + 32 * hh + 4 * tx;
            if (vecO && x2 + 3 < W2) {
                *reinterpret_cast<float4 *>(orow + x2) = make_float4(acc[i][4 * hh], acc[i][4 * hh + 1], acc[i][4 * hh + 2], acc[i][4 * hh + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (x2 + j < W2) orow[x2 + j] = acc[i][4 * hh + j];
            }
        }
    }
}

// ---- a14 on the tensor cores: 3xTF32 ------------------------------------------------------------
// ncu on the FFMA kernel: DRAM 9 %, issue slots 66 % — the one compute-bound op of the path (AI ~ 41 flop/B), so the
// north_star's "tensor cores only if compute-bound" clause applies.  Plain TF32 (10-bit mantissa) misses the 1e-4 bound;
// each fp32 operand is therefore split x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and the product is accumulated as
// lo*hi + hi*lo + hi*hi in fp32 (the dropped lo*lo term is ~2^-22 relative): fp32-level accuracy from three
// mma.sync.m16n8k8 TF32 instructions.  K = C = 96 is far too short for a tcgen05/TMEM pipeline to amortise its
// prologue on a 312 x 312 x 96 problem per row pair; warp-level MMA with register fragments fits the shape.
// CTA = 64 x 64 output tile, 4 warps (2 x 2), warp tile 32 x 32 = 2 x 4 MMA tiles; operands staged [k][64 + 8] so that
// the fragment reads (address = tig * pitch + gid) hit 32 distinct banks.
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
constexpr int kApPitch = kApTile + 8;
__global__ void __launch_bounds__(128)
corr1d_allpairs_mma_kernel(const float *__restrict__ f1, const float *__restrict__ f2, float *__restrict__ out,
                           float *__restrict__ pooled, int C, int H, int W1, int W2) {
    // pooled (optional): level 1 of the correlation pyramid, avg_pool2d([1,2]) of `out` (geometry_ddim.py:27-30), written
    // from the accumulators — a thread's fragment holds the two columns of a pair, (a + b) / 2 as the pooling kernel does
    __shared__ __align__(16) float sA[kApKc][kApPitch];
    __shared__ __align__(16) float sB[kApKc][kApPitch];
    const int by = blockIdx.z;
    const int b = by / H, y = by % H;
    const int x1_0 = blockIdx.y * kApTile, x2_0 = blockIdx.x * kApTile;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int gid = lane >> 2, tig = lane & 3;
    const int m_base = (warp >> 1) * 32, n_base = (warp & 1) * 32;
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.0f;
    const float *a0 = f1 + (static_cast<int64_t>(b) * C * H + y) * W1;
    const float *b0 = f2 + (static_cast<int64_t>(b) * C * H + y) * W2;
    const bool vecA = (W1 % 4 == 0) && ((reinterpret_cast<uintptr_t>(f1) & 15u) == 0) && x1_0 + kApTile <= W1;
    const bool vecB = (W2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(f2) & 15u) == 0) && x2_0 + kApTile <= W2;
    for (int c0 = 0; c0 < C; c0 += kApKc) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + 128 * u;
            const int k = e / 16, j = (e % 16) * 4;
            const int c = c0 + k;
            float4 va = make_float4(0.0f, 0.0f, 0.0f, 0.0f), vb = va;
            if (c < C) {
                const float *pa = a0 + static_cast<int64_t>(c) * H * W1 + x1_0 + j;
                const float *pb = b0 + static_cast<int64_t>(c) * H * W2 + x2_0 + j;
                if (vecA) va = __ldg(reinterpret_cast<const float4 *>(pa));
                else {
                    va.x = x1_0 + j + 0 < W1 ? __ldg(pa + 0) : 0.0f; va.y = x1_0 + j + 1 < W1 ? __ldg(pa + 1) : 0.0f;
                    va.z = x1_0 + j + 2 < W1 ? __ldg(pa + 2) : 0.0f; va.w = x1_0 + j + 3 < W1 ? __ldg(pa + 3) : 0.0f;
                }
                if (vecB) vb = __ldg(reinterpret_cast<const float4 *>(pb));
                else {
                    vb.x = x2_0 + j + 0 < W2 ? __ldg(pb + 0) : 0.0f; vb.y = x2_0 + j + 1 < W2 ? __ldg(pb + 1) : 0.0f;
                    vb.z = x2_0 + j + 2 < W2 ? __ldg(pb + 2) : 0.0f; vb.w = x2_0 + j + 3 < W2 ? __ldg(pb + 3) : 0.0f;
                }
            }
            *reinterpret_cast<float4 *>(&sA[k][j]) = va;
            *reinterpret_cast<float4 *>(&sB[k][j]) = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k0 = 0; k0 < kApKc; k0 += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int m = m_base + mt * 16 + gid;
                const float v[4] = {sA[k0 + tig][m], sA[k0 + tig][m + 8], sA[k0 + tig + 4][m], sA[k0 + tig + 4][m + 8]};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ah[mt][i] = to_tf32(v[i]);
                    al[mt][i] = to_tf32(__fsub_rn(v[i], __uint_as_float(ah[mt][i])));
                }
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n = n_base + nt * 8 + gid;
                const float v[2] = {sB[k0 + tig][n], sB[k0 + tig + 4][n]};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    bh[nt][i] = to_tf32(v[i]);
                    bl[nt][i] = to_tf32(__fsub_rn(v[i], __uint_as_float(bh[nt][i])));
                }
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mma_tf32(acc[mt][nt], al[mt], bh[nt]);
                    mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
                    mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
                }
        }
        __syncthreads();
    }
    const bool vec2 = (W2 % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7u) == 0);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int x1 = x1_0 + m_base + mt * 16 + gid + 8 * hh;
            if (x1 >= W1) continue;
            float *orow = out + (static_cast<int64_t>(by) * W1 + x1) * W2;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int x2 = x2_0 + n_base + nt * 8 + 2 * tig;
                const float v0 = acc[mt][nt][2 * hh], v1 = acc[mt][nt][2 * hh + 1];
                if (vec2 && x2 + 1 < W2) *reinterpret_cast<float2 *>(orow + x2) = make_float2(v0, v1);
                else {
                    if (x2 < W2) orow[x2] = v0;
                    if (x2 + 1 < W2) orow[x2 + 1] = v1;
                }
                if (pooled && x2 + 1 < W2)
                    pooled[(static_cast<int64_t>(by) * W1 + x1) * (W2 / 2) + x2 / 2] = __fadd_rn(v0, v1) / 2.0f;
            }
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

// Level-LVL noise of flat entry nj = n * D_l + j when D % 2^LVL == 0: the raw values [nj 2^LVL, (nj + 1) 2^LVL) reduced by the
// same pairwise tree as noise_at — but inlined (noise_at is recursive, i.e. a CALL per element, and a call between two
// loads keeps the second one from being issued before the first returns).
template <int LVL>
__device__ __forceinline__ float noise_flat(const float *__restrict__ p) {
    if constexpr (LVL == 0) {
        return __ldg(p);
    } else {
        return __fadd_rn(noise_flat<LVL - 1>(p), noise_flat<LVL - 1>(p + (1 << (LVL - 1)))) / 2.0f;
    }
}

// level-i noise at column j: avg-pooled i times from the raw noisy row (levels 0-3, the C-ABI's limit) — a switch over the
// inlined trees, not a recursion
__device__ __forceinline__ float noise_at(const float *__restrict__ row, int level, int j) {
    switch (level) {
        case 0: return noise_flat<0>(row + j);
        case 1: return noise_flat<1>(row + 2 * j);
        case 2: return noise_flat<2>(row + 4 * j);
        default: return noise_flat<3>(row + 8 * j);
    }
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


// ---- packed pyramid: hypothesis-major rows ------------------------------------------------------
// The reference keeps geo as [N, C, D] (geometry_ddim.py:19), so the 2r+2 hypotheses one lookup needs are C
// separate 40-byte windows 4*D bytes apart: ~2 sectors per channel, half of every sector unused, and a warp of
// 32 pixels touches 32 different 128-byte lines per load.  The packed layout is [N, D_l, C]: the window of one
// pixel and level is ONE contiguous run of (2r+2)*C floats, and for C = 8 every hypothesis is exactly one 32-byte
// sector.  geo_pack_kernel builds every level in one pass over the [B,C,D,h,w] volume (permute + avg-pool fused:
// the volume is read once; the reference layout is never materialised).
__device__ __forceinline__ const float *pick4(const float *const (&arr)[4], int i) {
    return i == 0 ? arr[0] : i == 1 ? arr[1] : i == 2 ? arr[2] : arr[3];  // no dynamically indexed kernel-parameter copy
}
__device__ __forceinline__ float *pick4w(float *const (&arr)[4], int i) {
    return i == 0 ? arr[0] : i == 1 ? arr[1] : i == 2 ? arr[2] : arr[3];
}
struct GeoPackArgs {
    float *rows[4];  // level l: [N, D >> l, C]
    int C, D, hw, levels;
};

constexpr int kPackDch = 16;  // hypotheses per CTA (multiple of 2^(levels-1) for levels <= 4 ... 8 | 16)

__device__ __forceinline__ float pooled_at(const float *__restrict__ tile, int C, int c, int pl, int level, int j) {
    // tile[(d_local * C + c) * 33 + pl]; level-l value j = avg_pool2d applied l times (pairwise (a + b) / 2, floor)
    if (level == 0) return tile[(j * C + c) * 33 + pl];
    const float a = pooled_at(tile, C, c, pl, level - 1, 2 * j), b = pooled_at(tile, C, c, pl, level - 1, 2 * j + 1);
    return __fadd_rn(a, b) / 2.0f;
}

// CTA = (32 pixels, kPackDch hypotheses, batch b): a warp reads 32 consecutive pixels of one (c, d) plane (128 bytes),
// the tile is transposed through shared memory, and every pixel's (d-chunk x C) block leaves as one contiguous run
// (512 bytes at level 0 for C = 8, 256 at level 1): one warp per pixel, lanes along the run.
// CT > 0: C == CT at compile time and only full kPackDch chunks take this instantiation (every index split is a shift).
template <int CT>
__global__ void __launch_bounds__(256)
geo_pack_kernel(const float *__restrict__ geo, const GeoPackArgs a) {
    extern __shared__ float tile[];  // [kPackDch * C][33], slot (d_local * C + c)
    const int C = CT > 0 ? CT : a.C;
    const int b = blockIdx.z, d0 = blockIdx.y * kPackDch, p0 = blockIdx.x * 32;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int nd = CT > 0 ? kPackDch : min(kPackDch, a.D - d0);
    const int p = p0 + lane;
    // 8 independent 128-byte row reads in flight per warp (a single outstanding load per warp caps the kernel at
    // ~1.3 TB/s by Little's law)
    const float *gb = geo + (static_cast<int64_t>(b) * C * a.D + d0) * a.hw + p;
    const int nrows = C * nd;
    for (int r0 = warp; r0 < nrows; r0 += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + 8 * u;
            const int c = r / nd, dd = r % nd;
            v[u] = (r < nrows && p < a.hw) ? __ldg(gb + (static_cast<int64_t>(c) * a.D + dd) * a.hw) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + 8 * u;
            if (r < nrows) tile[((r % nd) * C + r / nd) * 33 + lane] = v[u];
        }
    }
    __syncthreads();
    const int npx = min(32, a.hw - p0);
    for (int lvl = 0; lvl < a.levels; ++lvl) {
        const int Dl = a.D >> lvl;
        const int j0 = d0 >> lvl;
        const int nj = min((d0 + kPackDch) >> lvl, Dl) - j0;  // pooled entries whose sources all lie in this chunk
        if (nj <= 0) break;
        const int run = nj * C;
        float *base = pick4w(a.rows, lvl) + ((static_cast<int64_t>(b) * a.hw + p0) * Dl + j0) * C;
        for (int px = warp; px < npx; px += 8) {
            float *dst = base + static_cast<int64_t>(px) * Dl * C;
            if (lvl == 0) {
                for (int j = lane; j < run; j += 32) dst[j] = tile[j * 33 + px];
            } else if (lvl == 1) {
                for (int j = lane; j < run; j += 32) {
                    const int c = j % C, jj = j / C;
                    dst[j] = __fadd_rn(tile[((2 * jj) * C + c) * 33 + px], tile[((2 * jj + 1) * C + c) * 33 + px]) / 2.0f;
                }
            } else {
                for (int j = lane; j < run; j += 32) dst[j] = pooled_at(tile, C, j % C, px, lvl, j / C);
            }
        }
    }
}

// a9 for IGEV (geometry_ddim.py:37-43,56): geo_l * noise_l on the packed pyramid, every level in one launch.
// out_l[n, j, c] = in_l[n, j, c] * noise_l[n, j], noise_l = the raw [N, D] rows avg-pooled l times.
struct GeoFilterArgs {
    const float *in[4];
    float *out[4];
    const float *noisy;
    int C, D, levels;
    int fast;      // every level takes the vector path (geo_filter_level)
    int64_t N;
};
// Vector path of the filter (C % 4 == 0, D % 2^LVL == 0, < 2^31 float4 per level): a CTA owns a compact tile of 4 x 256
// float4, every thread issues its four volume loads and four noise gathers before the first multiply (the one-load loop
// below ran at 94 % occupancy with 25 long-scoreboard stall cycles per issue: 32 KB in flight per SM), no index divisions.
template <int LVL>
__device__ __forceinline__ void geo_filter_level(const GeoFilterArgs &a) {
    const float4 *in = reinterpret_cast<const float4 *>(pick4(a.in, LVL));
    float4 *out = reinterpret_cast<float4 *>(pick4w(a.out, LVL));
    const unsigned c4 = static_cast<unsigned>(a.C / 4);
    const unsigned total4 = static_cast<unsigned>(a.N * (a.D >> LVL) * c4);
    for (unsigned q0 = blockIdx.x * 1024u + threadIdx.x; q0 < total4; q0 += gridDim.x * 1024u) {
        float4 v[4];
        float nz[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned q = q0 + 256u * u;
            if (q < total4) {
                v[u] = __ldg(in + q);
                nz[u] = noise_flat<LVL>(a.noisy + (static_cast<int64_t>(q / c4) << LVL));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned q = q0 + 256u * u;
            if (q < total4)
                out[q] = make_float4(__fmul_rn(v[u].x, nz[u]), __fmul_rn(v[u].y, nz[u]), __fmul_rn(v[u].z, nz[u]), __fmul_rn(v[u].w, nz[u]));
        }
    }
}

__global__ void __launch_bounds__(256)
geo_filter_packed_kernel(const GeoFilterArgs a) {
    const int lvl = blockIdx.y;
    if (a.fast) {
        if (lvl == 0) geo_filter_level<0>(a);
        else if (lvl == 1) geo_filter_level<1>(a);
        else if (lvl == 2) geo_filter_level<2>(a);
        else geo_filter_level<3>(a);
        return;
    }
    const int Dl = a.D >> lvl;
    const float *in = pick4(a.in, lvl);
    float *out = pick4w(a.out, lvl);
    const int64_t per_row = static_cast<int64_t>(Dl) * a.C;
    const int64_t total = a.N * per_row;
    if (a.C % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
        const int c4 = a.C / 4;
        const int64_t total4 = total / 4;
        for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < total4;
             q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            const int64_t nj = q / c4;  // n * Dl + j
            const int64_t n = nj / Dl;
            const int j = static_cast<int>(nj % Dl);
            const float nz = noise_at(a.noisy + n * a.D, lvl, j);
            float4 v = __ldg(reinterpret_cast<const float4 *>(in) + q);
            v.x = __fmul_rn(v.x, nz); v.y = __fmul_rn(v.y, nz); v.z = __fmul_rn(v.z, nz); v.w = __fmul_rn(v.w, nz);
            reinterpret_cast<float4 *>(out)[q] = v;
        }
    } else {
        for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
             e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
            const int64_t nj = e / a.C;
            const int64_t n = nj / Dl;
            const int j = static_cast<int>(nj % Dl);
            out[e] = __fmul_rn(in[e], noise_at(a.noisy + n * a.D, lvl, j));
        }
    }
}

struct GeoLookupPackedArgs {
    const float *geo[4];   // level l: [N, D >> l, C]
    const float *corr[4];  // level l: [N, W2 >> l]
    const float *noisy, *disp, *coords;
    float *out;
    int C, D, hw, W2, levels, radius;
    int64_t N;
    float rcpD[4], rcpW[4];  // RN(1 / (D_l - 1)), RN(1 / (W2_l - 1)), computed on the host
};

// a / b correctly rounded from y = RN(1/b): q = RN(a*y), r = a - q*b (exact, FMA), RN(q + r*y).  Equal to __fdiv_rn for
// every integer-valued b <= 4096 and every normal a (checked exhaustively over the 2^23 significands per divisor; scaling
// by powers of two is exact), without the slow-path branches of the IEEE division sequence.
__device__ __forceinline__ float div_by_rcp(float a, float b, float y) {
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-q, b, a);
    return __fmaf_rn(r, y, q);
}
__device__ __forceinline__ float grid_coord_rcp(float x, float wm1, float rcp) {
    const float g = __fsub_rn(div_by_rcp(__fmul_rn(2.0f, x), wm1, rcp), 1.0f);
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), wm1);
}

// Window kernel (compile-time radius R, C == 4*C4): C4 adjacent lanes share a pixel, each owns 4 channels.  All 2R+2
// hypothesis vectors of the window [i0(tap 0), i0(tap 0) + 2R + 1] are fetched up front with clamped, unconditional
// loads (every gather of a pixel is in flight at once; one LDG.128 per hypothesis and lane, the C4 lanes of a pixel
// read one contiguous run) and each tap then blends registers.  No noise operand: the DDIM filter is applied to the
// packed pyramid once per timestep (geo_filter_packed_kernel) — the 32 GRU iterations of a step share one noise.  A tap whose
// own floor() disagrees with the window position (possible only through rounding at exact integers) takes a
// direct-load slow path with the same arithmetic.
// COOP: the window gather is done by the WARP, not by the pixel's own lanes.  In the direct form one LDG.128 instruction
// touches 16 pixels x one 32-byte sector each — 16 separate 32-byte requests to L2 — although the K hypothesis vectors of a
// pixel are one contiguous run of K*C4 16-byte chunks.  Here consecutive lanes copy consecutive chunks of the same run
// (lane -> (pixel, chunk) = divmod(i*32 + lane, K*C4); the pixel's window origin comes by shuffle from its owner lane) with
// LDGSTS into shared memory, so a run arrives as 3-4 full 128-byte lines; each thread then reads its own chunks back
// (pixel stride 22 chunks: conflict-free for a quarter warp).  Diagnostic builds had shown the gather half of this kernel
// moving its bytes at only 3.2 TB/s.  0.086 -> 0.078 ms at B = 8, 96x312 with 8 CTAs/SM.  (The same treatment of the
// 40-byte correlation windows — 4 chunks per pixel — measured slower, 0.082 ms, and was dropped.)
template <int C4, int R, int MINB, bool COOP = false>
__global__ void __launch_bounds__(128, MINB)
geo_lookup_window_kernel(const GeoLookupPackedArgs a) {
    constexpr int TAPS = 2 * R + 1, K = TAPS + 1, CC = 4 * C4, PIX = 128 / C4;
    constexpr int RUN = K * C4, PSTRIDE = RUN + 2;        // chunks per pixel run / per pixel slot in shared memory
    extern __shared__ __align__(16) float4 win[];         // COOP only: [PIX][PSTRIDE]
    const int lvl = blockIdx.y;
    const int Dl = a.D >> lvl, Wl = a.W2 >> lvl;
    const int64_t n0 = blockIdx.x * static_cast<int64_t>(PIX);
    const bool valid = n0 + threadIdx.x / C4 < a.N;
    const int64_t n = valid ? n0 + threadIdx.x / C4 : a.N - 1;  // tail lanes shadow the last pixel, they never store
    const int sub = threadIdx.x % C4;
    const int64_t b = n / a.hw;
    const int p = static_cast<int>(n % a.hw);
    constexpr int chan_per_level = (CC + 1) * TAPS;
    const float scale = static_cast<float>(1 << lvl);
    const float dm1 = static_cast<float>(Dl - 1), wm1 = static_cast<float>(Wl - 1);
    const float rD = lvl == 0 ? a.rcpD[0] : lvl == 1 ? a.rcpD[1] : lvl == 2 ? a.rcpD[2] : a.rcpD[3];
    const float rW = lvl == 0 ? a.rcpW[0] : lvl == 1 ? a.rcpW[1] : lvl == 2 ? a.rcpW[2] : a.rcpW[3];
    const float dl = __ldg(a.disp + n) / scale;                    // power of two: exact
    const float cl = __fsub_rn(__ldg(a.coords + n) / scale, dl);
    const float4 *grow = reinterpret_cast<const float4 *>(pick4(a.geo, lvl) + n * static_cast<int64_t>(CC) * Dl) + sub;
    const float *crow = pick4(a.corr, lvl) + n * static_cast<int64_t>(Wl);
    float *ol = a.out + (b * a.levels + lvl) * static_cast<int64_t>(chan_per_level) * a.hw + p;

    // Tap t samples at grid_coord(dx_t + dl); only the window origin and the corr tap indices are needed before the
    // loads are issued — the weights are recomputed at blend time (a dozen ALU ops per tap) instead of living in ~50
    // registers across the memory latency.
    auto geo_tap = [&](int t, int &i0, float &w0, float &w1) {
        const float ix = grid_coord_rcp(__fadd_rn(static_cast<float>(t - R), dl), dm1, rD);
        const float fl = floorf(ix);
        i0 = static_cast<int>(fl);
        w0 = __fsub_rn(static_cast<float>(i0 + 1), ix);
        w1 = __fsub_rn(ix, fl);
    };
    auto corr_tap = [&](int t, int &j0, float &w0, float &w1) {
        const float jx = grid_coord_rcp(__fadd_rn(cl, static_cast<float>(t - R)), wm1, rW);
        const float fj = floorf(jx);
        j0 = static_cast<int>(fj);
        w0 = __fsub_rn(static_cast<float>(j0 + 1), jx);
        w1 = __fsub_rn(jx, fj);
    };
    // window fetch: hypothesis i00 + k, clamped (out-of-range ones are never blended)
    float4 hyp[COOP ? 1 : K];
    int i00;
    {
        float w0, w1;
        geo_tap(0, i00, w0, w1);
    }
    if (COOP) {
        static_assert(!COOP || (32 % C4 == 0), "whole pixels per warp");
        constexpr int PPW = 32 / C4;                      // pixels per warp
        const int lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * PPW;      // first pixel slot of this warp
        const float *glvl = pick4(a.geo, lvl);
#pragma unroll
        for (int i = 0; i < (PPW * RUN + 31) / 32; ++i) {
            const int e = i * 32 + lane;
            const int pl = min(e / RUN, PPW - 1), ch = e - (e / RUN) * RUN;       // pixel of the warp, chunk of its run
            const int org = __shfl_sync(0xffffffffu, i00, pl * C4);               // that pixel's window origin
            const long long nn = __shfl_sync(0xffffffffu, static_cast<long long>(n), pl * C4);
            if (e < PPW * RUN) {
                const int hy = min(max(org + ch / C4, 0), Dl - 1);
                const float4 *src = reinterpret_cast<const float4 *>(glvl + nn * static_cast<int64_t>(CC) * Dl) + hy * C4 + (ch % C4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&win[(wbase + pl) * PSTRIDE + ch])), "l"(src)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = min(max(i00 + k, 0), Dl - 1);
            hyp[k] = __ldg(grow + static_cast<int64_t>(i) * C4);
        }
    }
    // corr taps owned by this lane: t % C4 == sub
    float cv0[TAPS], cv1[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
        if (t % C4 == sub) {
            int j0;
            float w0, w1;
            corr_tap(t, j0, w0, w1);
            cv0[t] = __ldg(crow + min(max(j0, 0), Wl - 1));
            cv1[t] = __ldg(crow + min(max(j0 + 1, 0), Wl - 1));
        }
    }
    if (COOP) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();                                      // the chunks of a pixel were copied by other lanes of the warp
    }
    if (!valid) return;
    const float4 *mine = win + (threadIdx.x / C4) * PSTRIDE + sub;
    float4 vnext = COOP ? mine[0] : hyp[0];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
        int i0;
        float gw0, gw1;
        geo_tap(t, i0, gw0, gw1);
        const int i1 = i0 + 1;
        const bool in0 = i0 >= 0 && i0 < Dl, in1 = i1 >= 0 && i1 < Dl;
        float4 v0 = vnext, v1 = COOP ? mine[(t + 1) * C4] : hyp[COOP ? 0 : t + 1];
        vnext = v1;
        if (i0 != i00 + t) {  // rounding at an exact integer moved this tap off the window grid
            v0 = __ldg(grow + static_cast<int64_t>(min(max(i0, 0), Dl - 1)) * C4);
            v1 = __ldg(grow + static_cast<int64_t>(min(max(i1, 0), Dl - 1)) * C4);
        }
        const float a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float acc = in0 ? __fmul_rn(a0[c], gw0) : 0.0f;
            if (in1) acc = __fadd_rn(acc, __fmul_rn(a1[c], gw1));
            ol[static_cast<int64_t>((4 * sub + c) * TAPS + t) * a.hw] = acc;
        }
        if (t % C4 == sub) {
            int j0;
            float cw0, cw1;
            corr_tap(t, j0, cw0, cw1);
            const int j1 = j0 + 1;
            float acc = (j0 >= 0 && j0 < Wl) ? __fmul_rn(cv0[t], cw0) : 0.0f;
            if (j1 >= 0 && j1 < Wl) acc = __fadd_rn(acc, __fmul_rn(cv1[t], cw1));
            ol[static_cast<int64_t>(CC * TAPS + t) * a.hw] = acc;
        }
    }
}

// thread = (pixel n, level blockIdx.y).  C4 > 0: C == 4*C4 at compile time (hypothesis vectors in registers, the upper
// hypothesis of tap t is reused as the lower one of tap t+1 when the indices agree); C4 == 0: any C, scalar loads.
template <int C4, int R>
__global__ void __launch_bounds__(128)
geo_lookup_packed_kernel(const GeoLookupPackedArgs a) {
    const int64_t n = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (n >= a.N) return;
    const int lvl = blockIdx.y;
    const int64_t b = n / a.hw;
    const int p = static_cast<int>(n % a.hw);
    // R >= 0: radius known at compile time -> the tap loop is fully unrolled and every gather of a pixel is in flight
    // at once (the loads depend only on disp/coords); R < 0: run-time radius
    const int radius = R >= 0 ? R : a.radius;
    const int taps = 2 * radius + 1;
    const int chan_per_level = (a.C + 1) * taps;
    const float scale = static_cast<float>(1 << lvl);
    const int Dl = a.D >> lvl, Wl = a.W2 >> lvl;
    const float dl = a.disp[n] / scale;
    const float cl = __fsub_rn(a.coords[n] / scale, dl);
    const float *nrow = a.noisy ? a.noisy + n * a.D : nullptr;
    const float *grow = pick4(a.geo, lvl) + n * static_cast<int64_t>(a.C) * Dl;
    const float *crow = pick4(a.corr, lvl) + n * static_cast<int64_t>(Wl);
    float *ol = a.out + (b * a.levels + lvl) * static_cast<int64_t>(chan_per_level) * a.hw + p;

    constexpr int CR = C4 > 0 ? 4 * C4 : 1;
    float carry[CR];  // hypothesis `carry_i`, already multiplied by its noise
    int carry_i = INT32_MIN;
#pragma unroll
    for (int t = 0; t < taps; ++t) {
        const float dx = static_cast<float>(t - radius);
        {
            const float ix = grid_coord(__fadd_rn(dx, dl), Dl);
            const float fl = floorf(ix);
            const int i0 = static_cast<int>(fl), i1 = i0 + 1;
            const float w0 = __fsub_rn(static_cast<float>(i1), ix), w1 = __fsub_rn(ix, fl);
            const bool in0 = i0 >= 0 && i0 < Dl, in1 = i1 >= 0 && i1 < Dl;
            const float n0 = (nrow && in0) ? noise_at(nrow, lvl, i0) : 1.0f;
            const float n1 = (nrow && in1) ? noise_at(nrow, lvl, i1) : 1.0f;
            if constexpr (C4 > 0) {
                float v0[CR], v1[CR];
                if (in0) {
                    if (i0 == carry_i) {
#pragma unroll
                        for (int c = 0; c < CR; ++c) v0[c] = carry[c];
                    } else {
#pragma unroll
                        for (int q = 0; q < C4; ++q) {
                            const float4 g = __ldg(reinterpret_cast<const float4 *>(grow + static_cast<int64_t>(i0) * CR) + q);
                            v0[4 * q + 0] = __fmul_rn(g.x, n0); v0[4 * q + 1] = __fmul_rn(g.y, n0);
                            v0[4 * q + 2] = __fmul_rn(g.z, n0); v0[4 * q + 3] = __fmul_rn(g.w, n0);
                        }
                    }
                }
                if (in1) {
#pragma unroll
                    for (int q = 0; q < C4; ++q) {
                        const float4 g = __ldg(reinterpret_cast<const float4 *>(grow + static_cast<int64_t>(i1) * CR) + q);
                        v1[4 * q + 0] = __fmul_rn(g.x, n1); v1[4 * q + 1] = __fmul_rn(g.y, n1);
                        v1[4 * q + 2] = __fmul_rn(g.z, n1); v1[4 * q + 3] = __fmul_rn(g.w, n1);
                    }
                }
#pragma unroll
                for (int c = 0; c < CR; ++c) {
                    float acc = in0 ? __fmul_rn(v0[c], w0) : 0.0f;
                    if (in1) acc = __fadd_rn(acc, __fmul_rn(v1[c], w1));
                    ol[static_cast<int64_t>(c * taps + t) * a.hw] = acc;
                }
                if (in1) {
#pragma unroll
                    for (int c = 0; c < CR; ++c) carry[c] = v1[c];
                    carry_i = i1;
                }
            } else {
                for (int c = 0; c < a.C; ++c) {
                    float acc = 0.0f;
                    if (in0) acc = __fmul_rn(__fmul_rn(grow[static_cast<int64_t>(i0) * a.C + c], n0), w0);
                    if (in1) acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(grow[static_cast<int64_t>(i1) * a.C + c], n1), w1));
                    ol[static_cast<int64_t>(c * taps + t) * a.hw] = acc;
                }
            }
        }
        {
            const float ix = grid_coord(__fadd_rn(cl, dx), Wl);
            const float fl = floorf(ix);
            const int i0 = static_cast<int>(fl), i1 = i0 + 1;
            const float w0 = __fsub_rn(static_cast<float>(i1), ix), w1 = __fsub_rn(ix, fl);
            float acc = 0.0f;
            if (i0 >= 0 && i0 < Wl) acc = __fmul_rn(__ldg(crow + i0), w0);
            if (i1 >= 0 && i1 < Wl) acc = __fadd_rn(acc, __fmul_rn(__ldg(crow + i1), w1));
            ol[static_cast<int64_t>(a.C * taps + t) * a.hw] = acc;
        }
    }
}

}  // namespace dv

static int corr1d_allpairs_impl(const float *fmap1, const float *fmap2, float *out, float *pooled, int64_t B, int64_t C,
                                int64_t H, int64_t W1, int64_t W2, void *stream) {
    using namespace dv;
    if (!fmap1 || !fmap2 || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W1 <= 0 || W2 <= 0 || B * H > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if (pooled && W2 < 2) return DV_ERR_BAD_SHAPE;
    dim3 grid(static_cast<unsigned>((W2 + kApTile - 1) / kApTile), static_cast<unsigned>((W1 + kApTile - 1) / kApTile),
              static_cast<unsigned>(B * H));
    if (grid.y > 65535 || B * H > 65535 * 1LL) return DV_ERR_UNSUPPORTED;   // B*H = 96 per pair in every reference configuration
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (DV_TUNE("DV_ALLPAIRS_TCGEN05", 1)) {
        // 5th-generation tensor cores (tcgen05.mma.kind::tf32, TMEM accumulators, TMA loads and stores); shapes outside
        // its tiling rules fall through to the warp-level kernels
        const int rc = launch_allpairs_tcgen05(fmap1, fmap2, out, pooled, B, C, H, W1, W2, st);
        if (rc != DV_ERR_UNSUPPORTED) return rc;
    }
    if (DV_TUNE("DV_ALLPAIRS_MMA", 1) || pooled) {
        corr1d_allpairs_mma_kernel<<<grid, 128, 0, st>>>(fmap1, fmap2, out, pooled, static_cast<int>(C), static_cast<int>(H),
                                                         static_cast<int>(W1), static_cast<int>(W2));
        return finish_launch();
    }
    corr1d_allpairs_kernel<<<grid, 64, 0, st>>>(fmap1, fmap2, out, static_cast<int>(C), static_cast<int>(H),
                                                static_cast<int>(W1), static_cast<int>(W2));
    return finish_launch();
}

extern "C" int dv_corr1d_allpairs_f32(const float *fmap1, const float *fmap2, float *out, int64_t B, int64_t C,
                                      int64_t H, int64_t W1, int64_t W2, void *stream) {
    return corr1d_allpairs_impl(fmap1, fmap2, out, nullptr, B, C, H, W1, W2, stream);
}

extern "C" int dv_corr1d_allpairs_pooled_f32(const float *fmap1, const float *fmap2, float *out, float *pooled, int64_t B,
                                             int64_t C, int64_t H, int64_t W1, int64_t W2, void *stream) {
    if (!pooled) return DV_ERR_NULL;
    return corr1d_allpairs_impl(fmap1, fmap2, out, pooled, B, C, H, W1, W2, stream);
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
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 16 ? blocks : static_cast<int64_t>(num_sms()) * 16);
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

extern "C" int dv_geo_pack_f32(const float *geo, float *const *rows_pyr, int64_t B, int64_t C, int64_t D, int64_t h,
                               int64_t w, int num_levels, void *stream) {
    using namespace dv;
    if (!geo || !rows_pyr) return DV_ERR_NULL;
    if (num_levels < 1 || num_levels > 4) return DV_ERR_UNSUPPORTED;
    if (B <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0 || B > 65535 || h * w > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if ((D >> (num_levels - 1)) < 1) return DV_ERR_BAD_SHAPE;
    const size_t smem = sizeof(float) * static_cast<size_t>(C) * kPackDch * 33;
    if (smem > 200 * 1024 || (D + kPackDch - 1) / kPackDch > 65535) return DV_ERR_UNSUPPORTED;
    GeoPackArgs a;
    for (int i = 0; i < 4; ++i) {
        a.rows[i] = i < num_levels ? rows_pyr[i] : nullptr;
        if (i < num_levels && !a.rows[i]) return DV_ERR_NULL;
    }
    a.C = static_cast<int>(C); a.D = static_cast<int>(D); a.hw = static_cast<int>(h * w); a.levels = num_levels;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(geo_pack_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    dim3 grid(static_cast<unsigned>((a.hw + 31) / 32), static_cast<unsigned>((D + kPackDch - 1) / kPackDch),
              static_cast<unsigned>(B));
    if (C == 8 && D % kPackDch == 0) {
        geo_pack_kernel<8><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(geo, a);
        return finish_launch();
    }
    geo_pack_kernel<0><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(geo, a);
    return finish_launch();
}

extern "C" int dv_geo_lookup_packed_f32(const float *const *geo_pyr, const float *const *corr_pyr, const float *noisy,
                                        const float *disp, const float *coords, float *out, int64_t B, int64_t C,
                                        int64_t D, int64_t h, int64_t w, int64_t W2, int num_levels, int radius,
                                        void *stream) {
    using namespace dv;
    if (!geo_pyr || !corr_pyr || !disp || !coords || !out) return DV_ERR_NULL;
    if (num_levels < 1 || num_levels > 4 || radius < 0 || radius > 16) return DV_ERR_UNSUPPORTED;
    if (B <= 0 || C <= 0 || D <= 0 || h <= 0 || w <= 0 || W2 <= 0 || h * w > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if ((D >> (num_levels - 1)) < 2 || (W2 >> (num_levels - 1)) < 2) return DV_ERR_BAD_SHAPE;
    GeoLookupPackedArgs a;
    bool al16 = true;
    for (int i = 0; i < 4; ++i) {
        a.geo[i] = i < num_levels ? geo_pyr[i] : nullptr;
        a.corr[i] = i < num_levels ? corr_pyr[i] : nullptr;
        if (i < num_levels && (!a.geo[i] || !a.corr[i])) return DV_ERR_NULL;
        if (i < num_levels) al16 = al16 && aligned16(a.geo[i]);
    }
    a.noisy = noisy; a.disp = disp; a.coords = coords; a.out = out;
    a.C = static_cast<int>(C); a.D = static_cast<int>(D); a.hw = static_cast<int>(h * w); a.W2 = static_cast<int>(W2);
    a.levels = num_levels; a.radius = radius; a.N = B * h * w;
    const int64_t blocks = (a.N + 127) / 128;
    if (blocks > INT32_MAX) return DV_ERR_BAD_SHAPE;
    dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(num_levels));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool small_divisors = true;
    for (int i = 0; i < 4; ++i) {
        const int64_t dm1 = (D >> i) - 1, wm1 = (W2 >> i) - 1;
        a.rcpD[i] = dm1 > 0 ? 1.0f / static_cast<float>(dm1) : 0.0f;
        a.rcpW[i] = wm1 > 0 ? 1.0f / static_cast<float>(wm1) : 0.0f;
        if (i < num_levels && (dm1 > 4096 || wm1 > 4096)) small_divisors = false;  // range div_by_rcp was verified on
    }
    if (al16 && C == 8 && radius == 4 && small_divisors && !noisy && DV_TUNE("DV_GEO_WINDOW", 1)) {  // IGEV
        const int64_t wblocks = (a.N + 63) / 64;  // 64 pixels x 2 lanes per block
        if (wblocks > INT32_MAX) return DV_ERR_BAD_SHAPE;
        const size_t smem = 0;
        const dim3 wgrid(static_cast<unsigned>(wblocks), static_cast<unsigned>(num_levels));
        const int minb = DV_TUNE("DV_GEO_MINB", 6);
        const int coop = DV_TUNE("DV_GEO_COOP", 8);
        const size_t csm = sizeof(float4) * 64 * (10 * 2 + 2);
        if (coop == 8) geo_lookup_window_kernel<2, 4, 8, true><<<wgrid, 128, csm, st>>>(a);
        else if (coop == 6) geo_lookup_window_kernel<2, 4, 6, true><<<wgrid, 128, csm, st>>>(a);
        else if (coop == 10) geo_lookup_window_kernel<2, 4, 10, true><<<wgrid, 128, csm, st>>>(a);
        else if (minb == 4) geo_lookup_window_kernel<2, 4, 4><<<wgrid, 128, smem, st>>>(a);
        else if (minb == 8) geo_lookup_window_kernel<2, 4, 8><<<wgrid, 128, smem, st>>>(a);
        else geo_lookup_window_kernel<2, 4, 6><<<wgrid, 128, smem, st>>>(a);
    } else if (al16 && C == 8 && radius == 4) geo_lookup_packed_kernel<2, 4><<<grid, 128, 0, st>>>(a);
    else if (al16 && C == 8) geo_lookup_packed_kernel<2, -1><<<grid, 128, 0, st>>>(a);
    else if (al16 && C == 4) geo_lookup_packed_kernel<1, -1><<<grid, 128, 0, st>>>(a);
    else if (al16 && C == 16) geo_lookup_packed_kernel<4, -1><<<grid, 128, 0, st>>>(a);
    else geo_lookup_packed_kernel<0, -1><<<grid, 128, 0, st>>>(a);
    return finish_launch();
}

extern "C" int dv_geo_filter_packed_f32(const float *const *rows_in, const float *noisy, float *const *rows_out, int64_t N,
                                        int64_t C, int64_t D, int num_levels, void *stream) {
    using namespace dv;
    if (!rows_in || !rows_out || !noisy) return DV_ERR_NULL;
    if (num_levels < 1 || num_levels > 4) return DV_ERR_UNSUPPORTED;
    if (N <= 0 || C <= 0 || D <= 0 || C > INT32_MAX || D > INT32_MAX || (D >> (num_levels - 1)) < 1) return DV_ERR_BAD_SHAPE;
    GeoFilterArgs a;
    for (int i = 0; i < 4; ++i) {
        a.in[i] = i < num_levels ? rows_in[i] : nullptr;
        a.out[i] = i < num_levels ? rows_out[i] : nullptr;
        if (i < num_levels && (!a.in[i] || !a.out[i])) return DV_ERR_NULL;
    }
    a.noisy = noisy; a.C = static_cast<int>(C); a.D = static_cast<int>(D); a.levels = num_levels; a.N = N;
    const int64_t work = (N * D * C + 3) / 4;
    bool fast = C % 4 == 0 && D % (1 << (num_levels - 1)) == 0 && work < 0x7fffffffLL && aligned16(noisy) &&
                DV_TUNE("DV_GEO_FILTER_FAST", 1);
    for (int i = 0; i < num_levels; ++i) fast = fast && aligned16(a.in[i]) && aligned16(a.out[i]);
    a.fast = fast ? 1 : 0;
    const int64_t blocks = fast ? (work + 1023) / 1024 : (work + 255) / 256;
    const int64_t cap = fast ? 0x7fffffffLL : static_cast<int64_t>(num_sms()) * 16;
    const unsigned gx = static_cast<unsigned>(blocks < cap ? blocks : cap);
    geo_filter_packed_kernel<<<dim3(gx, static_cast<unsigned>(num_levels)), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return finish_launch();
}
