// patch_chain.cu — ACVNet's "patch" depth-wise (1,3,3) convolutions on the gwc volume.
//
// Replaces the module chain of SceneFlow/models/acv_ddim.py:181-188,377-381 (same in acv.py):
//     g  = patch(gwc)                      nn.Conv3d(40,40,(1,3,3), groups=40, dilation 1, padding (0,1,1), no bias)
//     l1 = patch_l1(g[:, :8])   dil 1      l2 = patch_l2(g[:, 8:24])   dil 2      l3 = patch_l3(g[:, 24:40])   dil 3
//     patch_volume = cat(l1, l2, l3)
// i.e. per (b, c, d) plane two chained 3x3 stencils, each with ZERO PADDING OF ITS OWN INPUT (the intermediate is zero
// outside the image, it is not the first stencil evaluated out there).  cuDNN runs them as five volume-sized passes
// (conv, 3 grouped convs on strided slices, cat: ~6 x 249 MB per pair).  Here one kernel per dilation class reads an
// input tile with a (dil1 + dil2)-pixel halo into shared memory, builds the intermediate tile in shared memory and
// writes the final channels in place of the cat: the volume is read once and written once.
//   CTA = (tile of one [H,W] plane, plane (c,d), b); thread = strips of 4 horizontal pixels.
#include "common.cuh"

namespace dv {

constexpr int kPcTH = 32, kPcTW = 64;  // maximum output tile; the launch balances the actual tile to the plane

__global__ void __launch_bounds__(256)
depthwise_chain_kernel(const float *__restrict__ in, const float *__restrict__ w1, const float *__restrict__ w2,
                       float *__restrict__ out, int C, int D, int H, int W, int c0, int nc, int dil1, int dil2, int th,
                       int tw, int tiles_x) {
    extern __shared__ float sm[];
    const int r2 = w2 ? dil2 : 0, halo = dil1 + r2;
    const int iw = tw + 2 * halo, ih = th + 2 * halo;  // staged input tile
    const int mw = tw + 2 * r2, mh = th + 2 * r2;      // intermediate tile
    float *s_in = sm, *s_mid = sm + ih * iw;
    const int tile = blockIdx.x, ty0 = (tile / tiles_x) * th, tx0 = (tile % tiles_x) * tw;
    const int cd = blockIdx.y;                         // (c - c0) * D + d
    const int c = c0 + cd / D, d = cd % D, b = blockIdx.z;
    const int64_t plane = ((static_cast<int64_t>(b) * C + c) * D + d) * H * W;
    const float *ip = in + plane;
    float k1[9], k2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        k1[i] = __ldg(w1 + c * 9 + i);
        k2[i] = w2 ? __ldg(w2 + c * 9 + i) : 0.0f;
    }
    // stage the input tile (zero outside the image)
    for (int e = threadIdx.x; e < ih * iw; e += 256) {
        const int yy = e / iw, xx = e % iw;
        const int y = ty0 - halo + yy, x = tx0 - halo + xx;
        s_in[e] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(ip + static_cast<int64_t>(y) * W + x) : 0.0f;
    }
    __syncthreads();
    if (w2) {
        // first stencil on the (th + 2 dil2) x (tw + 2 dil2) region; zero where the intermediate lies outside the image
        for (int e = threadIdx.x; e < mh * mw; e += 256) {
            const int yy = e / mw, xx = e % mw;
            const int y = ty0 - r2 + yy, x = tx0 - r2 + xx;
            float acc = 0.0f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float *sp = s_in + (yy + dil1) * iw + xx + dil1;  // centre in the staged tile
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) acc = fmaf(k1[ky * 3 + kx], sp[(ky - 1) * dil1 * iw + (kx - 1) * dil1], acc);
            }
            s_mid[e] = acc;
        }
        __syncthreads();
    }
    const float *src = w2 ? s_mid : s_in;
    const int sw = w2 ? mw : iw, dl = w2 ? dil2 : dil1;
    float kk[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) kk[i] = w2 ? k2[i] : k1[i];
    float *op = out + plane;
    const int strips = (tw + 3) / 4;
    for (int e = threadIdx.x; e < th * strips; e += 256) {
        const int yy = e / strips, xs = (e % strips) * 4;
        const int y = ty0 + yy;
        if (y >= H) continue;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const float *sp = src + (yy + dl) * sw + xs + dl;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float *row = sp + (ky - 1) * dl * sw;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float wv = kk[ky * 3 + kx];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, row[(kx - 1) * dl + i], acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = tx0 + xs + i;
            if (xs + i < tw && x < W) op[static_cast<int64_t>(y) * W + x] = acc[i];
        }
    }
}


// ---- 128-bit path: W % 4 == 0, 16-byte aligned planes, dilations <= 4 ---------------------------------------------------
// Everything is laid out in quads (4 horizontal pixels): the staged input tile starts 8 columns left of the output tile,
// the intermediate tile 4 columns left, so the 3 taps x 4 outputs of one row always lie in three aligned float4 of the
// source tile (columns [m, m+12) around the quad at m+4) whatever the dilation: 9 LDS.128 + 36 FMA per output quad and
// stencil.  Because W % 4 == 0 a quad is either entirely inside the image or entirely outside (zero padding per quad).
__device__ __forceinline__ void stencil_quad(const float *__restrict__ row0, int pitch, int dil, const float (&k)[9],
                                             float (&acc)[4]) {
    // row0 points at column m of the tap row ky = 0; outputs sit at columns m+4 .. m+7
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const float *r = row0 + ky * dil * pitch;
        const float4 a = *reinterpret_cast<const float4 *>(r), b = *reinterpret_cast<const float4 *>(r + 4),
                     c = *reinterpret_cast<const float4 *>(r + 8);
        const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float wv = k[ky * 3 + kx];
            // tap column for output i: 4 + i + (kx-1)*dil  (dil <= 4 keeps it inside [0, 12))
            if (dil == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, v[4 + i + (kx - 1) * 1], acc[i]);
            } else if (dil == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, v[4 + i + (kx - 1) * 2], acc[i]);
            } else if (dil == 3) {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, v[4 + i + (kx - 1) * 3], acc[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(wv, v[4 + i + (kx - 1) * 4], acc[i]);
            }
        }
    }
}

// Row-blocked variant for compile-time dilations: a thread owns R vertically adjacent quads.  It walks the R + 2*DIL
// source rows once (3 LDS.128 each) and every source row feeds the up to three output rows it is a tap row of, so the
// shared-memory traffic per output quad drops from 9 LDS.128 to 3 (R + 2 DIL) / R (4.5 at DIL = 1, R = 4) and the index
// arithmetic is amortised over R quads — the runtime-dilation kernel below is issue-bound at 171 instructions per quad.
template <int DIL, int R>
__device__ __forceinline__ void stencil_rows(const float *__restrict__ row0, int pitch, const float (&k)[9], float (&acc)[R][4]) {
    // row0: column m of source row (first output row - DIL); outputs sit at columns m+4 .. m+7
#pragma unroll
    for (int sr = 0; sr < R + 2 * DIL; ++sr) {
        const float *r = row0 + sr * pitch;
        const float4 a = *reinterpret_cast<const float4 *>(r), b = *reinterpret_cast<const float4 *>(r + 4),
                     c = *reinterpret_cast<const float4 *>(r + 8);
        const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int orow = sr - ky * DIL;          // source row sr is tap row ky of output row sr - ky*DIL
            if (orow < 0 || orow >= R) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[orow][i] = fmaf(k[ky * 3 + kx], v[4 + i + (kx - 1) * DIL], acc[orow][i]);
        }
    }
}

template <bool CHAIN, int D1, int D2>
__global__ void __launch_bounds__(256)
depthwise_chain_rows_kernel(const float *__restrict__ in, const float *__restrict__ w1, const float *__restrict__ w2,
                            float *__restrict__ out, int C, int D, int H, int W, int c0, int th, int tw, int tiles_x) {
    constexpr int R = 4;
    extern __shared__ __align__(16) float sm[];
    constexpr int r2 = CHAIN ? D2 : 0;
    // tile heights are rounded up to whole row blocks in shared memory so that no thread reads outside its tile
    const int out_rb = (th + R - 1) / R, mid_rb = (th + 2 * r2 + R - 1) / R;
    // (the last output row block may look 2*D2 rows past the last whole mid row block: rows that only feed unstored outputs)
    const int mh = CHAIN ? max(mid_rb * R, out_rb * R + 2 * r2) : 0, mpitch = tw + 8;
    const int ih = (CHAIN ? mh : out_rb * R) + 2 * D1, ipitch = tw + (CHAIN ? 16 : 8);
    float *s_in = sm, *s_mid = sm + ih * ipitch;
    const int tile = blockIdx.x, ty0 = (tile / tiles_x) * th, tx0 = (tile % tiles_x) * tw;
    const int cd = blockIdx.y;
    const int c = c0 + cd / D, d = cd % D, b = blockIdx.z;
    const int64_t plane = ((static_cast<int64_t>(b) * C + c) * D + d) * H * W;
    const float *ip = in + plane;
    float k1[9], k2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        k1[i] = __ldg(w1 + c * 9 + i);
        k2[i] = CHAIN ? __ldg(w2 + c * 9 + i) : 0.0f;
    }
    {
        const int qpr = ipitch / 4, xl = tx0 - (CHAIN ? 8 : 4), yt = ty0 - D1 - r2;
        for (int e = threadIdx.x; e < ih * qpr; e += 256) {
            const int yy = e / qpr, qx = e - yy * qpr;
            const int y = yt + yy, x = xl + 4 * qx;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const float4 *>(ip + static_cast<int64_t>(y) * W + x));
            *reinterpret_cast<float4 *>(s_in + e * 4) = v;
        }
    }
    __syncthreads();
    if (CHAIN) {
        const int qpr = mpitch / 4;
        for (int e = threadIdx.x; e < mid_rb * qpr; e += 256) {
            const int rb = e / qpr, qx = e - rb * qpr;
            const int x = tx0 - 4 + 4 * qx;
            float acc[R][4];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[r][i] = 0.0f;
            // mid row rb*R + r  <->  input-tile row rb*R + r + D1 (centre); source rows start D1 above the first centre
            stencil_rows<D1, R>(s_in + (rb * R) * ipitch + 4 * qx, ipitch, k1, acc);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int y = ty0 - r2 + rb * R + r;
                const bool inside = y >= 0 && y < H && x >= 0 && x < W;   // the intermediate is ZERO outside the image
                *reinterpret_cast<float4 *>(s_mid + (rb * R + r) * mpitch + 4 * qx) =
                    inside ? make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        __syncthreads();
    }
    float *op = out + plane;
    const int qpr = tw / 4;
    for (int e = threadIdx.x; e < out_rb * qpr; e += 256) {
        const int rb = e / qpr, qx = e - rb * qpr;
        const int x = tx0 + 4 * qx;
        if (x >= W) continue;
        float acc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[r][i] = 0.0f;
        if (CHAIN) stencil_rows<(CHAIN ? D2 : 1), R>(s_mid + (rb * R) * mpitch + 4 * qx, mpitch, k2, acc);
        else stencil_rows<D1, R>(s_in + (rb * R) * ipitch + 4 * qx, ipitch, k1, acc);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int yy = rb * R + r, y = ty0 + yy;
            if (yy < th && y < H)
                *reinterpret_cast<float4 *>(op + static_cast<int64_t>(y) * W + x) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
    }
}

template <bool CHAIN, int D1, int D2>
static int launch_rows(const float *in, const float *w1, const float *w2, float *out, int B, int C, int D, int H, int W, int c0,
                       int c1, cudaStream_t st) {
    constexpr int R = 4, r2 = CHAIN ? D2 : 0;
    const int tiles_y = (H + 31) / 32, tiles_x = (W + 127) / 128;
    const int th = (H + tiles_y - 1) / tiles_y;
    const int tw = (((W + tiles_x - 1) / tiles_x) + 3) / 4 * 4;
    const int tiles_x2 = (W + tw - 1) / tw;
    const int out_rb = (th + R - 1) / R, mid_rb = (th + 2 * r2 + R - 1) / R;
    const int mh = CHAIN ? (mid_rb * R > out_rb * R + 2 * r2 ? mid_rb * R : out_rb * R + 2 * r2) : 0;
    const int ih = (CHAIN ? mh : out_rb * R) + 2 * D1;
    const size_t smem = sizeof(float) * (static_cast<size_t>(ih) * (tw + (CHAIN ? 16 : 8)) + static_cast<size_t>(mh) * (tw + 8));
    auto kern = depthwise_chain_rows_kernel<CHAIN, D1, D2>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    dim3 grid(static_cast<unsigned>(tiles_y * tiles_x2), static_cast<unsigned>((c1 - c0) * D), static_cast<unsigned>(B));
    kern<<<grid, 256, smem, st>>>(in, w1, w2, out, C, D, H, W, c0, th, tw, tiles_x2);
    return finish_launch();
}

// ---- streaming kernel: one WARP walks one [H,W] plane top to bottom -------------------------------------------------------
// The tile kernels above run at ~300 instructions per output quad (run-time tile shapes, e / qpr index splits, phase barriers,
// 1.3x halo recompute of the intermediate) and are issue-bound (ncu: 70 % issue slots, DRAM 32 %).  Here a plane is a
// stream of rows:
//   * lane l owns PX adjacent columns; input rows arrive through an NI-deep ring of LDGSTS copies (one contiguous W*4-byte
//     row per warp instruction pair, zero-filled below the image), NPF rows ahead of their use;
//   * the first stencil (dilation 1) keeps rows m-1, m in REGISTERS and loads only row m+1 ((PX+8)/4 LDS.128), the
//     intermediate row goes to an MR-deep shared ring (zero outside the image, like the reference's second zero padding);
//   * the second stencil reads its three tap rows m-2 D2, m-D2, m from that ring and stores output row m-D2 (one
//     contiguous row per warp);  no vertical halo is ever recomputed, and the only synchronisation is __syncwarp.
// All shared offsets are compile-time (PITCH = 32 PX + 8: 4 zero columns left and right of the row).
__device__ __forceinline__ void pc_cp_async16(void *dst_smem, const void *src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(sz) : "memory");
}

template <int D2, int PX, int NI>
__global__ void __launch_bounds__(32)
patch_chain_stream_kernel(const float *__restrict__ in, const float *__restrict__ w1, const float *__restrict__ w2,
                          float *__restrict__ out, int C, int D, int H, int W, int c0) {
    static_assert(PX % 4 == 0 && D2 >= 1 && D2 <= 3 && (NI & (NI - 1)) == 0, "quads; taps within the 4-column pad");
    constexpr int NPF = NI - 2, MR = D2 == 1 ? 4 : 8, PITCH = 32 * PX + 8, NV = PX + 8, Q = PX / 4;
    extern __shared__ __align__(16) float sm[];
    float *rin = sm, *rmid = sm + NI * PITCH;
    const int lane = threadIdx.x;
    const int c = c0 + blockIdx.x / D, d = blockIdx.x % D, b = blockIdx.y;
    const int64_t plane = ((static_cast<int64_t>(b) * C + c) * D + d) * H * W;
    const float *ip = in + plane;
    float *op = out + plane;
    float k1[9], k2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        k1[i] = __ldg(w1 + c * 9 + i);
        k2[i] = __ldg(w2 + c * 9 + i);
    }
    for (int e = lane; e < (NI + MR) * PITCH / 4; e += 32) reinterpret_cast<float4 *>(sm)[e] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncwarp();
    const int col0 = lane * PX;
    // 16-byte chunk q of a row (chunk 0 = the left zero pad) lives at physical chunk sw(q): lanes read windows of NV/4 chunks
    // at a stride of PX/4 chunks, which without the XOR puts lanes l and l+4 (PX = 8) / l+2 (PX = 16) of a quarter-warp on
    // the same banks — ncu: 2/3 of the shared wavefronts were conflict replays, LSU pipe 90 % busy
    auto sw = [](int q) { return PX == 8 ? (q ^ ((q >> 3) & 1)) : PX == 16 ? (q ^ ((q >> 3) & 3)) : q; };
    int off[NV / 4];      // float offsets of this lane's window chunks inside a row; own data = chunks 1 .. Q
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) off[j] = 4 * sw(lane * Q + j);
    bool qok[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) qok[j] = col0 + 4 * j < W;
    auto issue_row = [&](int y) {   // rows >= H are zero rows of the padding
        float *dst = rin + (y & (NI - 1)) * PITCH;
        const float *src = ip + static_cast<int64_t>(y) * W + col0;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            const bool ok = y < H && qok[j];
            pc_cp_async16(dst + off[1 + j], ok ? src + 4 * j : in, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_row = [&](const float *row, float (&v)[NV]) {   // columns col0-4 .. col0+PX+3
#pragma unroll
        for (int j = 0; j < NV / 4; ++j) {
            const float4 t = *reinterpret_cast<const float4 *>(row + off[j]);
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        }
    };
    for (int y = 0; y <= NPF; ++y) issue_row(y);
    float ra[NV], rb[NV], rc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) ra[i] = 0.0f;                  // row -1
    asm volatile("cp.async.wait_group %0;" ::"n"(NPF) : "memory");
    __syncwarp();
    load_row(rin, rb);                                            // row 0
    const int M = H + D2;
    // one step: intermediate row m from input rows (r0, r1, r2) = (m-1, m, m+1), then output row m - D2
    auto step = [&](int m, const float (&r0)[NV], const float (&r1)[NV], float (&r2)[NV]) {
        issue_row(m + 1 + NPF);
        asm volatile("cp.async.wait_group %0;" ::"n"(NPF) : "memory");
        __syncwarp();
        load_row(rin + ((m + 1) & (NI - 1)) * PITCH, r2);
        float acc[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) acc[i] = 0.0f;
        if (m < H) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] = fmaf(k1[kx], r0[3 + i + kx], acc[i]);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] = fmaf(k1[3 + kx], r1[3 + i + kx], acc[i]);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[i] = fmaf(k1[6 + kx], r2[3 + i + kx], acc[i]);
        }
        float *mrow = rmid + (m & (MR - 1)) * PITCH;
#pragma unroll
        for (int j = 0; j < Q; ++j)     // the intermediate is ZERO outside the image (second zero padding)
            *reinterpret_cast<float4 *>(mrow + off[1 + j]) = (m < H && qok[j])
                ? make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        __syncwarp();
        const int o = m - D2;
        if (o < 0) return;
        float res[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) res[i] = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float v[NV];
            load_row(rmid + ((o + (ky - 1) * D2) & (MR - 1)) * PITCH, v);   // rows < 0: slots still hold their initial zeros
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int i = 0; i < PX; ++i) res[i] = fmaf(k2[ky * 3 + kx], v[4 + i + (kx - 1) * D2], res[i]);
        }
        float *orow = op + static_cast<int64_t>(o) * W + col0;
#pragma unroll
        for (int j = 0; j < Q; ++j)
            if (qok[j]) *reinterpret_cast<float4 *>(orow + 4 * j) = make_float4(res[4 * j], res[4 * j + 1], res[4 * j + 2], res[4 * j + 3]);
    };
    for (int m = 0; m < M; m += 3) {
        step(m, ra, rb, rc);
        if (m + 1 < M) step(m + 1, rb, rc, ra);
        if (m + 2 < M) step(m + 2, rc, ra, rb);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int D2, int PX, int NI = 8>
static int launch_stream(const float *in, const float *w1, const float *w2, float *out, int B, int C, int D, int H, int W,
                         int c0, int c1, cudaStream_t st) {
    constexpr int MR = D2 == 1 ? 4 : 8, PITCH = 32 * PX + 8;
    const size_t smem = sizeof(float) * (NI + MR) * PITCH;
    auto kern = patch_chain_stream_kernel<D2, PX, NI>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    dim3 grid(static_cast<unsigned>((c1 - c0) * D), static_cast<unsigned>(B));
    kern<<<grid, 32, smem, st>>>(in, w1, w2, out, C, D, H, W, c0);
    return finish_launch();
}

template <int D2>
static int launch_stream_px(const float *in, const float *w1, const float *w2, float *out, int B, int C, int D, int H, int W,
                            int c0, int c1, cudaStream_t st) {
    if (W <= 256) {
        // 4 ring slots (2 rows ahead) fit 17 warps per SM instead of 13: 0.878 vs 0.896 ms at B = 8, 135x240
        if (DV_TUNE("DV_PATCH_NI", 4) == 8) return launch_stream<D2, 8, 8>(in, w1, w2, out, B, C, D, H, W, c0, c1, st);
        return launch_stream<D2, 8, 4>(in, w1, w2, out, B, C, D, H, W, c0, c1, st);
    }
    if (W <= 384) return launch_stream<D2, 12>(in, w1, w2, out, B, C, D, H, W, c0, c1, st);
    return launch_stream<D2, 16>(in, w1, w2, out, B, C, D, H, W, c0, c1, st);
}

template <bool CHAIN>
__global__ void __launch_bounds__(256)
depthwise_chain_quad_kernel(const float *__restrict__ in, const float *__restrict__ w1, const float *__restrict__ w2,
                            float *__restrict__ out, int C, int D, int H, int W, int c0, int dil1, int dil2, int th, int tw,
                            int tiles_x) {
    extern __shared__ __align__(16) float sm[];
    const int r2 = CHAIN ? dil2 : 0;
    const int ih = th + 2 * (dil1 + r2), ipitch = tw + (CHAIN ? 16 : 8);   // input tile starts (CHAIN ? 8 : 4) columns left
    const int mh = th + 2 * r2, mpitch = tw + 8;                          // intermediate tile starts 4 columns left
    float *s_in = sm, *s_mid = sm + ih * ipitch;
    const int tile = blockIdx.x, ty0 = (tile / tiles_x) * th, tx0 = (tile % tiles_x) * tw;
    const int cd = blockIdx.y;
    const int c = c0 + cd / D, d = cd % D, b = blockIdx.z;
    const int64_t plane = ((static_cast<int64_t>(b) * C + c) * D + d) * H * W;
    const float *ip = in + plane;
    float k1[9], k2[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        k1[i] = __ldg(w1 + c * 9 + i);
        k2[i] = CHAIN ? __ldg(w2 + c * 9 + i) : 0.0f;
    }
    {   // stage: quads of the input tile
        const int qpr = ipitch / 4, xl = tx0 - (CHAIN ? 8 : 4), yt = ty0 - dil1 - r2;
        for (int e = threadIdx.x; e < ih * qpr; e += 256) {
            const int yy = e / qpr, qx = e - yy * qpr;
            const int y = yt + yy, x = xl + 4 * qx;
            float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const float4 *>(ip + static_cast<int64_t>(y) * W + x));
            *reinterpret_cast<float4 *>(s_in + e * 4) = v;
        }
    }
    __syncthreads();
    if (CHAIN) {
        const int qpr = mpitch / 4;
        for (int e = threadIdx.x; e < mh * qpr; e += 256) {
            const int yy = e / qpr, qx = e - yy * qpr;
            const int y = ty0 - r2 + yy, x = tx0 - 4 + 4 * qx;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            // intermediate quad at mid column 4*qx == input column 4*qx + 4; centre row yy + dil1 of the input tile
            if (y >= 0 && y < H && x >= 0 && x < W) stencil_quad(s_in + yy * ipitch + 4 * qx, ipitch, dil1, k1, acc);
            *reinterpret_cast<float4 *>(s_mid + e * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        __syncthreads();
    }
    float *op = out + plane;
    const int qpr = tw / 4;
    for (int e = threadIdx.x; e < th * qpr; e += 256) {
        const int yy = e / qpr, qx = e - yy * qpr;
        const int y = ty0 + yy, x = tx0 + 4 * qx;
        if (y >= H || x >= W) continue;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (CHAIN) stencil_quad(s_mid + yy * mpitch + 4 * qx, mpitch, dil2, k2, acc);
        else stencil_quad(s_in + yy * ipitch + 4 * qx, ipitch, dil1, k1, acc);
        *reinterpret_cast<float4 *>(op + static_cast<int64_t>(y) * W + x) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

}  // namespace dv

extern "C" int dv_depthwise3x3_chain_f32(const float *in, const float *w1, const float *w2, float *out, int64_t B, int64_t C,
                                         int64_t D, int64_t H, int64_t W, int64_t c0, int64_t c1, int dil1, int dil2,
                                         void *stream) {
    using namespace dv;
    if (!in || !w1 || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || c0 < 0 || c1 > C || c0 >= c1) return DV_ERR_BAD_SHAPE;
    if (dil1 < 1 || dil1 > 8 || (w2 && (dil2 < 1 || dil2 > 8))) return DV_ERR_UNSUPPORTED;
    if (B > 65535 || (c1 - c0) * D > 65535 || H * W > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if (in == out) return DV_ERR_UNSUPPORTED;  // tiles read their neighbours' inputs
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the ACVNet chain (patch dil 1 -> patch_l* dil 1..3) on planes up to 512 columns: the row-streaming kernel
    if (w2 && dil1 == 1 && dil2 >= 1 && dil2 <= 3 && W % 4 == 0 && W >= 64 && W <= 512 && aligned16(in) && aligned16(out) &&
        DV_TUNE("DV_PATCH_STREAM", 1)) {
        const int Bi = static_cast<int>(B), Ci = static_cast<int>(C), Di = static_cast<int>(D), Hi = static_cast<int>(H),
                  Wi = static_cast<int>(W), a = static_cast<int>(c0), z = static_cast<int>(c1);
        if (dil2 == 1) return launch_stream_px<1>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        if (dil2 == 2) return launch_stream_px<2>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        return launch_stream_px<3>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
    }
    if (W % 4 == 0 && aligned16(in) && aligned16(out) && DV_TUNE("DV_PATCH_ROWS", 1)) {
        // compile-time dilations (everything ACVNet uses): the row-blocked kernel
        const int Bi = static_cast<int>(B), Ci = static_cast<int>(C), Di = static_cast<int>(D), Hi = static_cast<int>(H),
                  Wi = static_cast<int>(W), a = static_cast<int>(c0), z = static_cast<int>(c1);
        if (w2 && dil1 == 1 && dil2 == 1) return launch_rows<true, 1, 1>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        if (w2 && dil1 == 1 && dil2 == 2) return launch_rows<true, 1, 2>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        if (w2 && dil1 == 1 && dil2 == 3) return launch_rows<true, 1, 3>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        if (!w2 && dil1 == 1) return launch_rows<false, 1, 1>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        if (!w2 && dil1 == 2) return launch_rows<false, 2, 1>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
        if (!w2 && dil1 == 3) return launch_rows<false, 3, 1>(in, w1, w2, out, Bi, Ci, Di, Hi, Wi, a, z, st);
    }
    if (W % 4 == 0 && aligned16(in) && aligned16(out) && dil1 <= 4 && (!w2 || dil2 <= 4) && DV_TUNE("DV_PATCH_QUAD", 1)) {
        // tiles: up to 32 rows x 128 columns, balanced to the plane, widths in quads
        const int tiles_y = static_cast<int>((H + 31) / 32), tiles_x = static_cast<int>((W + 127) / 128);
        const int th = static_cast<int>((H + tiles_y - 1) / tiles_y);
        const int tw = static_cast<int>((((W + tiles_x - 1) / tiles_x) + 3) / 4 * 4);
        const int tiles_x2 = static_cast<int>((W + tw - 1) / tw);
        const int r2 = w2 ? dil2 : 0;
        const size_t smem = sizeof(float) * (static_cast<size_t>(th + 2 * (dil1 + r2)) * (tw + (w2 ? 16 : 8)) +
                                             (w2 ? static_cast<size_t>(th + 2 * r2) * (tw + 8) : 0));
        dim3 grid(static_cast<unsigned>(tiles_y * tiles_x2), static_cast<unsigned>((c1 - c0) * D), static_cast<unsigned>(B));
        if (w2) {
            if (smem > 48 * 1024 && cudaFuncSetAttribute(depthwise_chain_quad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         static_cast<int>(smem)) != cudaSuccess)
                return DV_ERR_LAUNCH;
            depthwise_chain_quad_kernel<true><<<grid, 256, smem, st>>>(in, w1, w2, out, static_cast<int>(C), static_cast<int>(D),
                                                                      static_cast<int>(H), static_cast<int>(W), static_cast<int>(c0),
                                                                      dil1, dil2, th, tw, tiles_x2);
        } else {
            if (smem > 48 * 1024 && cudaFuncSetAttribute(depthwise_chain_quad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         static_cast<int>(smem)) != cudaSuccess)
                return DV_ERR_LAUNCH;
            depthwise_chain_quad_kernel<false><<<grid, 256, smem, st>>>(in, w1, w2, out, static_cast<int>(C), static_cast<int>(D),
                                                                       static_cast<int>(H), static_cast<int>(W), static_cast<int>(c0),
                                                                       dil1, dil2, th, tw, tiles_x2);
        }
        return finish_launch();
    }
    const int tiles_y = static_cast<int>((H + kPcTH - 1) / kPcTH), tiles_x = static_cast<int>((W + kPcTW - 1) / kPcTW);
    const int th = static_cast<int>((H + tiles_y - 1) / tiles_y), tw = static_cast<int>((((W + tiles_x - 1) / tiles_x) + 3) / 4 * 4);
    const int r2 = w2 ? dil2 : 0, halo = dil1 + r2;
    const size_t smem = sizeof(float) * (static_cast<size_t>(th + 2 * halo) * (tw + 2 * halo) +
                                         (w2 ? static_cast<size_t>(th + 2 * r2) * (tw + 2 * r2) : 0));
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(depthwise_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    const int tiles_x2 = static_cast<int>((W + tw - 1) / tw);
    dim3 grid(static_cast<unsigned>(tiles_y * tiles_x2), static_cast<unsigned>((c1 - c0) * D), static_cast<unsigned>(B));
    depthwise_chain_kernel<<<grid, 256, smem, st>>>(
        in, w1, w2, out, static_cast<int>(C), static_cast<int>(D), static_cast<int>(H), static_cast<int>(W),
        static_cast<int>(c0), static_cast<int>(c1 - c0), dil1, dil2, th, tw, tiles_x2);
    return finish_launch();
}
