// volume_backward.cu — backward passes of the volume ops for sm_100a.
//
// The reference's training scripts differentiate through build_gwc_volume / build_concat_volume /
// build_corrleation_volume / groupwise_correlation / disparity_regression (SceneFlow/main.py:154 ->
// SceneFlow/models/acv_ddim.py:424-482; KITTI12/main.py, KITTI15/train_stereo.py); autograd runs their backward as
// D x {slice, mul, sum, index_put} chains over [B,C,H,W-d] temporaries.  Here each gradient is ONE pass:
//
//   gwc  : d_ref[c,p] = 1/cpg * sum_d  g[grp,d,p]   * tgt[c,p-d]   (x(p) >= d)
//          d_tgt[c,p] = 1/cpg * sum_d  g[grp,d,p+d] * ref[c,p+d]   (x(p)+d < W)
//          thread = (b, group, pixel) keeps the 2 x cpg channel accumulators in registers, so the gradient volume is
//          read exactly twice (once per operand; coalesced along x), the features come from L1/L2.
//   concat: d_ref[c,p] = sum_d g[c,d,p] (x >= d when the left half is masked); d_tgt[c,p] = sum_d g[C+c,d,p+d]
//          every gradient element is read exactly once (pure HBM read stream, D independent loads per thread).
//   regression: d_x[b,d,p] = d * g[b,p]   (pure 128-bit write stream).
//
// All planes are addressed flattened (p = y*W + x); a shift that leaves the row is a masked term, exactly mirroring the
// forward kernels (gwc_volume.cu, concat_volume.cu).  No atomics: every output element has one owner thread.
#include "common.cuh"

namespace dv {

// grad_out [B,G,Dtot,HW]; non-negative shifts d in [0,D) live at planes dofs+d; when mneg > 0 the planes [0,mneg)
// hold the negative shifts of build_corrleation_volume (slot s <-> k = mneg - s; only columns x < k are live and pair
// ref[x] with tgt[x + max(W-k,0)], KITTI12/models/submodule.py:128-131).
template <int CPG>
__global__ void __launch_bounds__(128)
gwc_bwd_kernel(const float *__restrict__ go, const float *__restrict__ ref, const float *__restrict__ tgt,
               float *__restrict__ gref, float *__restrict__ gtgt, int C, int HW, int W, int D, int G, int Dtot, int dofs,
               int mneg) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int g = blockIdx.y, b = blockIdx.z;
    const int x = p % W;
    const float *gp = go + (static_cast<int64_t>(b) * G + g) * Dtot * HW;
    const int64_t fb = (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * CPG) * HW;
    const float *rp = ref + fb, *tp = tgt + fb;
    float ar[CPG], at[CPG];
#pragma unroll
    for (int k = 0; k < CPG; ++k) ar[k] = at[k] = 0.0f;
    const int dmax_r = min(D - 1, x);            // x >= d
    const int dmax_t = min(D - 1, W - 1 - x);    // x + d < W
    if (gref) {
#pragma unroll 4
        for (int d = 0; d <= dmax_r; ++d) {
            const float gv = __ldg(gp + static_cast<int64_t>(dofs + d) * HW + p);
#pragma unroll
            for (int k = 0; k < CPG; ++k) ar[k] = fmaf(gv, __ldg(tp + static_cast<int64_t>(k) * HW + p - d), ar[k]);
        }
        for (int s = 0; s < mneg; ++s) {
            const int kk = mneg - s;
            if (x < kk) {
                const int off = max(W - kk, 0);
                const float gv = __ldg(gp + static_cast<int64_t>(s) * HW + p);
#pragma unroll
                for (int k = 0; k < CPG; ++k) ar[k] = fmaf(gv, __ldg(tp + static_cast<int64_t>(k) * HW + p + off), ar[k]);
            }
        }
    }
    if (gtgt) {
#pragma unroll 4
        for (int d = 0; d <= dmax_t; ++d) {
            const float gv = __ldg(gp + static_cast<int64_t>(dofs + d) * HW + p + d);
#pragma unroll
            for (int k = 0; k < CPG; ++k) at[k] = fmaf(gv, __ldg(rp + static_cast<int64_t>(k) * HW + p + d), at[k]);
        }
        for (int s = 0; s < mneg; ++s) {
            const int kk = mneg - s;
            const int off = max(W - kk, 0);
            const int xs = x - off;              // the ref column paired with this tgt column
            if (xs >= 0 && xs < kk) {
                const float gv = __ldg(gp + static_cast<int64_t>(s) * HW + p - off);
#pragma unroll
                for (int k = 0; k < CPG; ++k) at[k] = fmaf(gv, __ldg(rp + static_cast<int64_t>(k) * HW + p - off), at[k]);
            }
        }
    }
    constexpr float inv = 1.0f / CPG;
#pragma unroll
    for (int k = 0; k < CPG; ++k) {
        if (gref) gref[fb + static_cast<int64_t>(k) * HW + p] = ar[k] * inv;
        if (gtgt) gtgt[fb + static_cast<int64_t>(k) * HW + p] = at[k] * inv;
    }
}

// ---- quad kernel (plain gwc volume, HW % 4 == 0, 16-byte aligned) ------------------------------------------------------
// Same tiling idea as the forward kernel (gwc_volume.cu): CTA = (span of SQ quads, group x channel chunk, batch); the CK
// channel rows of ref (span + Dpad) and of the tgt window (Dpad + span) are staged once in shared memory by bulk copies;
// thread = one quad (4 pixels), 2 x CK x 4 accumulators in registers, the D loop runs in blocks of 4 disparities so that
// every feature operand is a static pick from two aligned float4 of shared memory:
//     d_ref[k][p+i] += g[d][p+i]   * tgt[k][p+i-d]      window  tgt[p-d0-4 .. p-d0+3], element 4 + i - (d-d0)
//     d_tgt[k][p+i] += g[d][p+i+d] * ref[k][p+i+d]      window  ref[p+d0 .. p+d0+7],   element i + (d-d0)
// g is read straight from global (128-bit, L1-cached: the shifted d_tgt reads of a CTA overlap its d_ref reads).  Per 4
// disparities a thread issues 4*CK LDS.128 + 12 LDG.128 for 32*CK FMA (the thread-per-pixel kernel: 8 LDG per FMA pair).
template <int CK, int SQ>
__global__ void __launch_bounds__(2 * SQ)
gwc_bwd_quad_kernel(const float *__restrict__ go, const float *__restrict__ ref, const float *__restrict__ tgt,
                    float *__restrict__ gref, float *__restrict__ gtgt, int C, int HW, int W, int D, int G, int cpg, int Dpad,
                    int Dtot, int dofs, int64_t go_elems) {
    constexpr int SPAN = SQ * 4;
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int pitch = SPAN + Dpad;
    float *sR = smem;               // ref rows  [CK][SPAN + Dpad], element 0 = flat index p0
    float *sT = smem + CK * pitch;  // tgt rows  [CK][Dpad + SPAN], element Dpad = flat index p0
    const int chunks = cpg / CK;
    const int g = blockIdx.y / chunks, kc = blockIdx.y % chunks, b = blockIdx.z;
    const int p0 = blockIdx.x * SPAN;
    const int len = min(SPAN, HW - p0);
    const int64_t fb = (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg + kc * CK) * HW;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int rlen = min(len + Dpad, HW - p0);        // ref: do not run past the plane (masked terms only)
        // tgt window start of channel k is toff + k*HW; it lies before the tensor for the first channels of (b=0, g=0) whenever
        // k*HW + p0 < Dpad (tiny planes: several channels) — clip each copy to the tensor
        const int64_t toff = fb + p0 - Dpad;
        auto skip_of = [&](int k) -> int {
            const int64_t o = toff + static_cast<int64_t>(k) * HW;
            return o < 0 ? static_cast<int>(o < -(len + Dpad) ? len + Dpad : -o) : 0;
        };
        if (threadIdx.x == 0) {
            int skipped = 0;
            for (int k = 0; k < CK; ++k) skipped += skip_of(k);
            mbar_expect_tx(&bar, 4u * (CK * (rlen + len + Dpad) - skipped));
        }
        __syncwarp();
        for (int k = threadIdx.x; k < CK; k += 32) {
            bulk_g2s(sR + k * pitch, ref + fb + static_cast<int64_t>(k) * HW + p0, 4u * rlen, &bar);
            const int skip = skip_of(k);
            if (skip < len + Dpad)
                bulk_g2s(sT + k * pitch + skip, tgt + toff + static_cast<int64_t>(k) * HW + skip, 4u * (len + Dpad - skip), &bar);
        }
    }
    {   // the parts of the windows no copy lands in (plane tail of ref, tensor head of tgt) are only ever multiplied by a
        // masked (zero) gradient: clear them so that stale shared memory cannot inject NaNs
        const int rlen = min(len + Dpad, HW - p0);
        const int64_t toff = fb + p0 - Dpad;
        for (int k = 0; k < CK; ++k) {
            for (int e = rlen + threadIdx.x; e < pitch; e += 2 * SQ) sR[k * pitch + e] = 0.0f;
            for (int e = len + Dpad + threadIdx.x; e < pitch; e += 2 * SQ) sT[k * pitch + e] = 0.0f;
            const int64_t o = toff + static_cast<int64_t>(k) * HW;
            const int skip = o < 0 ? static_cast<int>(o < -(len + Dpad) ? len + Dpad : -o) : 0;
            for (int e = threadIdx.x; e < skip; e += 2 * SQ) sT[k * pitch + e] = 0.0f;
        }
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    // warps [0, SQ/32) accumulate d_ref, warps [SQ/32, 2 SQ/32) accumulate d_tgt: the two gradients share no operand
    // besides the staged rows, and splitting them halves the registers and the serial work per thread
    const int q = threadIdx.x % SQ;
    const bool tgt_pass = threadIdx.x >= SQ;
    float *gdst = tgt_pass ? gtgt : gref;
    const int p = p0 + 4 * q;
    if (p >= HW || !gdst) return;
    int xs[4];
    xs[0] = p % W;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        xs[i] = xs[i - 1] + 1;
        if (xs[i] >= W) xs[i] -= W;
    }
    float acc[CK][4];
#pragma unroll
    for (int k = 0; k < CK; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[k][i] = 0.0f;
    // non-negative shifts live at planes dofs .. dofs+D-1 of the Dtot planes of (b, g) (two-sided volume: dofs = m)
    const int64_t gbase = ((static_cast<int64_t>(b) * G + g) * Dtot + dofs) * HW + p;   // flat index of gp[0]
    const float *gp = go + gbase;
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (!tgt_pass) {
        for (int d0 = 0; d0 < D; d0 += 4) {
            float m[4][4];  // masked gradient: the d_ref term needs x >= d
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = d0 + j;
                const float4 gv = d < D ? __ldg(reinterpret_cast<const float4 *>(gp + static_cast<int64_t>(d) * HW)) : zero4;
                const float v[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) m[j][i] = xs[i] >= d ? v[i] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < CK; ++k) {
                const float *tp = sT + k * pitch + Dpad + 4 * q - d0 - 4;
                const float4 t0 = *reinterpret_cast<const float4 *>(tp), t1 = *reinterpret_cast<const float4 *>(tp + 4);
                const float tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[k][i] = fmaf(m[j][i], tw[4 + i - j], acc[k][i]);
            }
        }
    } else {
        for (int d0 = 0; d0 < D; d0 += 4) {
            float m[4][4];  // masked shifted gradient g[d][p+i+d]: the d_tgt term needs x + d < W
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int d = d0 + j;
                const int64_t o = static_cast<int64_t>(d) * HW + d0;
                // shifted reads may run past the plane (masked terms) but must not run past the tensor
                const float4 ga = (d < D && gbase + o + 3 < go_elems) ? __ldg(reinterpret_cast<const float4 *>(gp + o)) : zero4;
                const float4 gb = (d < D && gbase + o + 7 < go_elems) ? __ldg(reinterpret_cast<const float4 *>(gp + o + 4)) : zero4;
                const float v[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) m[j][i] = (d < D && xs[i] + d < W) ? v[i + j] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < CK; ++k) {
                const float *rp = sR + k * pitch + 4 * q + d0;
                const float4 r0 = *reinterpret_cast<const float4 *>(rp), r1 = *reinterpret_cast<const float4 *>(rp + 4);
                const float rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[k][i] = fmaf(m[j][i], rw[i + j], acc[k][i]);
            }
        }
    }
    const float inv = 1.0f / static_cast<float>(cpg);
#pragma unroll
    for (int k = 0; k < CK; ++k)
        *reinterpret_cast<float4 *>(gdst + fb + static_cast<int64_t>(k) * HW + p) =
            make_float4(acc[k][0] * inv, acc[k][1] * inv, acc[k][2] * inv, acc[k][3] * inv);
}

// ---- ring kernel: gradient volume staged through a cp.async ring, mask-free inner loops ----------------------------------
// ncu on the quad kernel (profiles/r02_gwc_bwd.md): 26 % warps active with long-scoreboard stalls on the gradient LDG.128s,
// 332 instructions per block of 128 FMAs (64-bit row addressing, 32 mask instructions, run-time smem pitch), LSU wavefronts
// at 70 % of peak (the d_tgt warps re-read the tile the d_ref warps read, through L1).  This kernel, for W % 4 == 0:
//  * the gradient rows of (b, g) stream through an NST-deep ring of [RS rows][PITCH] tiles filled by LDGSTS (cp.async.cg 16 B,
//    zero-fill past D / past the tensor); loads run NST-1 stages ahead of the FMAs and both roles read the SAME staged tile;
//  * NO masks: a quad starts at a column x0 % 4 == 0, so for d_ref (needs x >= d) every block of 4 disparities d0 < x0 is
//    fully live, the block d0 == x0 is the static triangle i >= j, and later blocks are dead — the loop just ends; d_tgt
//    (needs x + d < W) is the mirror image with lim = W - x0 - 4 and the triangle i + j < 4;
//  * the feature window slides in registers: a block re-uses half of the previous block's window (1 LDS.128 per channel
//    instead of 2), ping-ponged between two register sets so that no moves are needed;
//  * compile-time pitch (LDS with immediate offsets), features staged by the same cp.async group as the first tile.
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kRingSQ = 64, kRingSpan = 4 * kRingSQ, kRingDpad = 64, kRingPitch = kRingSpan + kRingDpad;

template <int CK, bool TGT>
struct RingBlock {
    // one block of 4 disparities d0 .. d0+3 for one quad.  `gs` = staged gradient rows of the block at the quad's column,
    // `fw` = this role's staged feature rows at the quad (sT + DPAD + 4q for d_ref, sR + 4q for d_tgt), `o` = the window
    // half kept from the previous block, `n` = the half loaded here (kept for the next block).
    static __device__ __forceinline__ void full(float (&acc)[CK][4], const float *gs, const float *fw, int d0, float4 (&o)[CK],
                                                float4 (&n)[CK]) {
        float m[4][4];
        if constexpr (!TGT) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 gv = *reinterpret_cast<const float4 *>(gs + j * kRingPitch);
                m[j][0] = gv.x; m[j][1] = gv.y; m[j][2] = gv.z; m[j][3] = gv.w;
            }
#pragma unroll
            for (int k = 0; k < CK; ++k) {   // window tgt[p-d0-4 .. p-d0+3] = {n, o}
                n[k] = *reinterpret_cast<const float4 *>(fw + k * kRingPitch - d0 - 4);
                const float tw[8] = {n[k].x, n[k].y, n[k].z, n[k].w, o[k].x, o[k].y, o[k].z, o[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[k][i] = fmaf(m[j][i], tw[4 + i - j], acc[k][i]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {    // g[d0+j][p + d0 + j + i]
                const float *gr = gs + j * kRingPitch + d0;
                const float4 ga = *reinterpret_cast<const float4 *>(gr), gb = *reinterpret_cast<const float4 *>(gr + 4);
                const float v[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) m[j][i] = v[i + j];
            }
#pragma unroll
            for (int k = 0; k < CK; ++k) {   // window ref[p+d0 .. p+d0+7] = {o, n}
                n[k] = *reinterpret_cast<const float4 *>(fw + k * kRingPitch + d0 + 4);
                const float rw[8] = {o[k].x, o[k].y, o[k].z, o[k].w, n[k].x, n[k].y, n[k].z, n[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[k][i] = fmaf(m[j][i], rw[i + j], acc[k][i]);
            }
        }
    }
};

template <int CK, int RS, int NST>
__global__ void __launch_bounds__(2 * kRingSQ)
gwc_bwd_ring_kernel(const float *__restrict__ go, const float *__restrict__ ref, const float *__restrict__ tgt,
                    float *__restrict__ gref, float *__restrict__ gtgt, int C, int HW, int W, int D, int G, int cpg, int chunks,
                    int Dtot, int dofs) {
    constexpr int SQ = kRingSQ, SPAN = kRingSpan, NT = 2 * SQ, PITCH = kRingPitch, DPAD = kRingDpad, P4 = PITCH / 4;
    static_assert(RS == 8, "a stage holds two blocks of 4 disparities (the window ping-pong)");
    static_assert(NT == 16 * RS && P4 % 16 == 0, "stage copy: 16 threads per gradient row");
    extern __shared__ __align__(16) float smem[];
    float *sR = smem;                   // ref rows  [CK][PITCH], element 0 = flat index p0
    float *sT = smem + CK * PITCH;      // tgt rows  [CK][PITCH], element DPAD = flat index p0
    float *sG = smem + 2 * CK * PITCH;  // gradient ring [NST][RS][PITCH], element 0 = column p0 of the row's plane
    const int kc = blockIdx.x % chunks, p0 = (blockIdx.x / chunks) * SPAN;   // chunk fastest: CTAs sharing a tile are neighbours
    const int g = blockIdx.y, b = blockIdx.z;
    const int64_t fb = (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg + kc * CK) * HW;
    const int tid = threadIdx.x;
    const float *gplane = go + ((static_cast<int64_t>(b) * G + g) * Dtot + dofs) * HW;   // row 0 of this (b, g)
    const int nstage = (D + RS - 1) / RS;
    // stage copy: 16 threads per row, 5 chunks each (columns tid%16 + 16 i).  Columns past the plane read the next row (dead
    // terms); only the last row of the last (b, g) would leave the tensor there, so that row is clipped to the plane.
    const int srow = tid / 16, scol = 4 * (tid % 16);
    const bool last_bg = b == static_cast<int>(gridDim.z) - 1 && g == G - 1 && dofs + D == Dtot;
    auto issue_stage = [&](int s) {
        const int d = s * RS + srow;
        float *dst = sG + ((s % NST) * RS + srow) * PITCH + scol;
        const float *src = gplane + static_cast<int64_t>(d) * HW + p0 + scol;
        const int ncol = (last_bg && d == D - 1) ? HW - p0 - scol : PITCH;   // columns of this row that may be read
#pragma unroll
        for (int i = 0; i < P4 / 16; ++i) {
            const bool ok = d < D && 64 * i + 3 < ncol;
            cp_async16(dst + 64 * i, ok ? src + 64 * i : go, ok);
        }
    };
    {   // features: ref columns [p0, p0 + PITCH), tgt columns [p0 - DPAD, p0 + SPAN); what lies outside the plane is only ever
        // paired with dead gradient terms and is zero-filled.  NT / (2 CK) threads per row.
        constexpr int TPR = NT / (2 * CK), CPT = (P4 + TPR - 1) / TPR;
        const int row = tid / TPR, c0 = tid % TPR;       // rows [0, CK) = ref, [CK, 2 CK) = tgt
        const bool is_t = row >= CK;
        const int colbase = p0 - (is_t ? DPAD : 0);      // flat index of the row's element 0
        const float *src = (is_t ? tgt : ref) + fb + static_cast<int64_t>(is_t ? row - CK : row) * HW + colbase;
        float *dst = smem + row * PITCH;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = c0 + TPR * i;
            if ((P4 % TPR != 0 && c >= P4) || (NT % (2 * CK) != 0 && row >= 2 * CK)) break;
            const int col = colbase + 4 * c;
            const bool ok = col >= 0 && col + 3 < HW;
            cp_async16(dst + 4 * c, ok ? src + 4 * c : go, ok);
        }
    }
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nstage) issue_stage(s);
        cp_async_commit();
    }
    const int q = tid % SQ;
    const bool tgt_pass = tid >= SQ;
    float *gdst = tgt_pass ? gtgt : gref;
    const int p = p0 + 4 * q;
    const bool active = p < HW && gdst != nullptr;
    const int x0 = p % W;                                  // % 4 == 0
    // blocks d0 < lim are fully live; d0 == lim_raw is the boundary triangle (done after the loop, all lanes at once: inside
    // the loop it would run for one lane at a time); rows past D are zero-filled and their blocks skipped
    const int lim_raw = tgt_pass ? W - x0 - 4 : x0, Dc = (D + 3) / 4 * 4;
    const int lim = active ? min(lim_raw, Dc) : 0;
    const float *fw = tgt_pass ? sR + 4 * q : sT + DPAD + 4 * q;
    float acc[CK][4];
    float4 wa[CK], wb[CK];
#pragma unroll
    for (int k = 0; k < CK; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[k][i] = 0.0f;
    for (int s = 0; s < nstage; ++s) {
        cp_async_wait<NST - 2>();
        __syncthreads();   // stage s (and, for s == 0, the features) landed for every thread; stage s-1's buffer is free
        if (s + NST - 1 < nstage) issue_stage(s + NST - 1);
        cp_async_commit();
        const int d0 = s * RS;
        if (d0 >= lim) continue;
        const float *gs = sG + (s % NST) * RS * PITCH + 4 * q;
        if (s == 0) {
#pragma unroll
            for (int k = 0; k < CK; ++k) wa[k] = *reinterpret_cast<const float4 *>(fw + k * PITCH);   // window at d0 = 0
        }
        if (!tgt_pass) {
            RingBlock<CK, false>::full(acc, gs, fw, d0, wa, wb);
            if (d0 + 4 < lim) RingBlock<CK, false>::full(acc, gs + 4 * PITCH, fw, d0 + 4, wb, wa);
        } else {
            RingBlock<CK, true>::full(acc, gs, fw, d0, wa, wb);
            if (d0 + 4 < lim) RingBlock<CK, true>::full(acc, gs + 4 * PITCH, fw, d0 + 4, wb, wa);
        }
    }
    cp_async_wait<0>();
    if (!active) return;
    if (lim_raw < Dc) {   // boundary triangle: gradient rows lim_raw .. +3 straight from global (rows >= D count as zero)
        const float *gq = gplane + static_cast<int64_t>(lim_raw) * HW + p + (tgt_pass ? lim_raw : 0);
        float4 gv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            gv[j] = lim_raw + j < D ? __ldg(reinterpret_cast<const float4 *>(gq + static_cast<int64_t>(j) * HW))
                                    : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float *fo = fw + (tgt_pass ? lim_raw : -lim_raw);
#pragma unroll
        for (int k = 0; k < CK; ++k) {
            const float4 o = *reinterpret_cast<const float4 *>(fo + k * PITCH);
            const float w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float v[4] = {gv[j].x, gv[j].y, gv[j].z, gv[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (tgt_pass) {          // terms i + j < 4:  g[d][p + d + i] * ref[p + d + i]
                        if (i + j < 4) acc[k][i] = fmaf(v[i + j], w[i + j], acc[k][i]);
                    } else {                 // terms i >= j:     g[d][p + i] * tgt[p + i - d]
                        if (i >= j) acc[k][i] = fmaf(v[i], w[i - j], acc[k][i]);
                    }
                }
            }
        }
    }
    const float inv = 1.0f / static_cast<float>(cpg);
#pragma unroll
    for (int k = 0; k < CK; ++k)
        *reinterpret_cast<float4 *>(gdst + fb + static_cast<int64_t>(k) * HW + p) =
            make_float4(acc[k][0] * inv, acc[k][1] * inv, acc[k][2] * inv, acc[k][3] * inv);
}

// Negative shifts of the two-sided volume (slot s <-> k = m - s pairs ref[x] with tgt[x + W - k] for x < k, KITTI12/models/
// submodule.py:128-131): only the first m columns of d_ref and the last m columns of d_tgt receive terms.  One thread per
// (b, c, y, j < m) ADDS them to the gradients the quad kernel has written (same stream, one owner per element).
__global__ void __launch_bounds__(128)
corr_negative_bwd_kernel(const float *__restrict__ go, const float *__restrict__ ref, const float *__restrict__ tgt,
                         float *__restrict__ gref, float *__restrict__ gtgt, int C, int H, int W, int m, int G, int cpg,
                         int64_t total) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int j = static_cast<int>(idx % m);
    int64_t t = idx / m;
    const int y = static_cast<int>(t % H);
    t /= H;
    const int c = static_cast<int>(t % C);
    const int64_t b = t / C;
    const int g = c / cpg;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const float inv = 1.0f / static_cast<float>(cpg);
    const float *gp = go + (b * G + g) * (2 * m + 1) * HW + static_cast<int64_t>(y) * W;   // row y of slot 0
    const int64_t frow = (b * C + c) * HW + static_cast<int64_t>(y) * W;
    if (gref) {       // d_ref[x = j] += sum_{k > j} g[slot m-k][x] * tgt[x + W - k]
        float acc = 0.0f;
        for (int k = j + 1; k <= m; ++k)
            acc = fmaf(__ldg(gp + static_cast<int64_t>(m - k) * HW + j), __ldg(tgt + frow + j + W - k), acc);
        gref[frow + j] += acc * inv;
    }
    if (gtgt) {       // d_tgt[x' = W - m + j] += sum_{k >= m - j} g[slot m-k][x' - (W - k)] * ref[x' - (W - k)]
        const int xp = W - m + j;
        float acc = 0.0f;
        for (int k = max(1, m - j); k <= m; ++k) {
            const int xs = xp - (W - k);   // = k - m + j in [0, k)
            acc = fmaf(__ldg(gp + static_cast<int64_t>(m - k) * HW + xs), __ldg(ref + frow + xs), acc);
        }
        gtgt[frow + xp] += acc * inv;
    }
}

// The same one image row per warp (m <= 32, W >= 2 m): lane x keeps the m gradient values g[slot][y][x] in registers (they
// do not depend on the channel), loads ref[c][y][x] and tgt[c][y][W-m+x] per channel, and gets its partners' operands by
// shuffle.  Same term order as corr_negative_bwd_kernel (k ascending), so the sums are bit-identical; that kernel spent
// 71 us at B = 4, 384x1248 on per-thread strided gathers.
__global__ void __launch_bounds__(128)
corr_negative_bwd_rows_kernel(const float *__restrict__ go, const float *__restrict__ ref, const float *__restrict__ tgt,
                              float *__restrict__ gref, float *__restrict__ gtgt, int C, int H, int W, int m, int G, int cpg,
                              int nrows) {
    const int row = blockIdx.x * 4 + threadIdx.x / 32, x = threadIdx.x % 32;
    if (row >= nrows) return;
    const int y = row % H, bg = row / H, g = bg % G, b = bg / G;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const bool live = x < m;
    const int xc = live ? x : 0;
    const float *gp = go + static_cast<int64_t>(bg) * (2 * m + 1) * HW + static_cast<int64_t>(y) * W + xc;   // slot 0, column x
    float gs[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) gs[s] = (s < m && live) ? __ldg(gp + static_cast<int64_t>(s) * HW) : 0.0f;
    const int64_t frow = (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg) * HW + static_cast<int64_t>(y) * W;
    const float inv = 1.0f / static_cast<float>(cpg);
    for (int c0 = 0; c0 < cpg; c0 += 8) {            // loads (operands and the gradients to update) issued before the first use
        float rv8[8], tv8[8], or8[8], ot8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t fr = frow + static_cast<int64_t>(min(c0 + u, cpg - 1)) * HW;
            rv8[u] = __ldg(ref + fr + xc);
            tv8[u] = __ldg(tgt + fr + W - m + xc);
            or8[u] = gref ? gref[fr + xc] : 0.0f;
            ot8[u] = gtgt ? gtgt[fr + W - m + xc] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (c0 + u >= cpg) break;                    // uniform
            const int64_t fr = frow + static_cast<int64_t>(c0 + u) * HW;
            const float rv = rv8[u], tv = tv8[u];
            float ar = 0.0f, at = 0.0f;
#pragma unroll
            for (int s = 31; s >= 0; --s) {                         // slot s <-> k = m - s, k ascending; static register index
                if (s >= m) continue;                                // uniform
                const int k = m - s;
                // d_ref[x] += g[s][x] * tgt[W-k+x]              (x < k)        partner lane m-k+x = s+x
                const float t = __shfl_sync(0xffffffffu, tv, (s + x) & 31);
                if (x < k) ar = fmaf(gs[s], t, ar);
                // d_tgt[W-m+x] += g[s][xs] * ref[xs], xs = x - s   (xs >= 0)     partner lane xs
                const int xs = x - s;
                const float gg = __shfl_sync(0xffffffffu, gs[s], xs & 31), rr = __shfl_sync(0xffffffffu, rv, xs & 31);
                if (xs >= 0) at = fmaf(gg, rr, at);
            }
            if (live) {
                if (gref) gref[fr + x] = or8[u] + ar * inv;
                if (gtgt) gtgt[fr + W - m + x] = ot8[u] + at * inv;
            }
        }
    }
}

template <int CK>
static int launch_gwc_bwd_quad(const float *go, const float *ref, const float *tgt, float *gref, float *gtgt, int B, int C,
                               int HW, int W, int D, int G, int cpg, int Dtot, int dofs, cudaStream_t st) {
    constexpr int SQ = 64, SPAN = SQ * 4;
    const int Dpad = (D + 3) / 4 * 4 + 4;  // the last block of 4 disparities reads up to d0 + 7 floats past a quad
    const size_t smem = sizeof(float) * 2 * CK * (SPAN + Dpad);
    auto kern = gwc_bwd_quad_kernel<CK, SQ>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    dim3 grid((HW + SPAN - 1) / SPAN, G * (cpg / CK), B);
    kern<<<grid, 2 * SQ, smem, st>>>(go, ref, tgt, gref, gtgt, C, HW, W, D, G, cpg, Dpad, Dtot, dofs,
                                     static_cast<int64_t>(B) * G * Dtot * HW);
    return finish_launch();
}

template <int CK, int RS, int NST>
static int launch_gwc_bwd_ring(const float *go, const float *ref, const float *tgt, float *gref, float *gtgt, int B, int C,
                               int HW, int W, int D, int G, int cpg, int Dtot, int dofs, cudaStream_t st) {
    // rows past D are zero-filled in the ring, so D is rounded up to whole blocks; the window reaches DPAD columns back
    if (W % 4 != 0 || (D + 3) / 4 * 4 + 4 > kRingDpad || static_cast<int64_t>(Dtot) * HW > INT32_MAX) return DV_ERR_UNSUPPORTED;
    const int chunks = cpg / CK, nspan = (HW + kRingSpan - 1) / kRingSpan;
    if (static_cast<int64_t>(nspan) * chunks > INT32_MAX) return DV_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * (2 * CK + NST * RS) * kRingPitch;
    auto kern = gwc_bwd_ring_kernel<CK, RS, NST>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    dim3 grid(nspan * chunks, G, B);
    kern<<<grid, 2 * kRingSQ, smem, st>>>(go, ref, tgt, gref, gtgt, C, HW, W, D, G, cpg, chunks, Dtot, dofs);
    return finish_launch();
}

template <int CK>
static int launch_gwc_bwd_best(const float *go, const float *ref, const float *tgt, float *gref, float *gtgt, int B, int C,
                               int HW, int W, int D, int G, int cpg, int Dtot, int dofs, cudaStream_t st) {
    int rc = DV_ERR_UNSUPPORTED;
    switch (DV_TUNE("DV_GWC_BWD_RING", 1)) {
        case 0: break;
        case 2: rc = launch_gwc_bwd_ring<CK, 8, 4>(go, ref, tgt, gref, gtgt, B, C, HW, W, D, G, cpg, Dtot, dofs, st); break;
        case 3: rc = launch_gwc_bwd_ring<CK, 8, 2>(go, ref, tgt, gref, gtgt, B, C, HW, W, D, G, cpg, Dtot, dofs, st); break;
        default: rc = launch_gwc_bwd_ring<CK, 8, 3>(go, ref, tgt, gref, gtgt, B, C, HW, W, D, G, cpg, Dtot, dofs, st); break;
    }
    if (rc != DV_ERR_UNSUPPORTED) return rc;
    return launch_gwc_bwd_quad<CK>(go, ref, tgt, gref, gtgt, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
}

// any channels-per-group: thread = (b, channel, pixel)
__global__ void gwc_bwd_generic_kernel(const float *__restrict__ go, const float *__restrict__ ref,
                                       const float *__restrict__ tgt, float *__restrict__ gref, float *__restrict__ gtgt,
                                       int C, int HW, int W, int D, int G, int cpg, int Dtot, int dofs, int mneg,
                                       int64_t total) {
    const float inv = 1.0f / static_cast<float>(cpg);
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % HW);
        const int64_t bc = idx / HW;
        const int c = static_cast<int>(bc % C);
        const int64_t b = bc / C;
        const int g = c / cpg, x = p % W;
        const float *gp = go + (b * G + g) * Dtot * HW;
        const float *rp = ref + bc * HW, *tp = tgt + bc * HW;
        float ar = 0.0f, at = 0.0f;
        for (int d = 0; d < D; ++d) {
            if (x >= d) ar = fmaf(gp[static_cast<int64_t>(dofs + d) * HW + p], tp[p - d], ar);
            if (x + d < W) at = fmaf(gp[static_cast<int64_t>(dofs + d) * HW + p + d], rp[p + d], at);
        }
        for (int s = 0; s < mneg; ++s) {
            const int kk = mneg - s, off = max(W - kk, 0), xs = x - off;
            if (x < kk) ar = fmaf(gp[static_cast<int64_t>(s) * HW + p], tp[p + off], ar);
            if (xs >= 0 && xs < kk) at = fmaf(gp[static_cast<int64_t>(s) * HW + p - off], rp[p - off], at);
        }
        if (gref) gref[idx] = ar * inv;
        if (gtgt) gtgt[idx] = at * inv;
    }
}

static int gwc_bwd_impl(const float *go, const float *ref, const float *tgt, float *gref, float *gtgt, int64_t B, int64_t C,
                        int64_t H, int64_t W, int64_t D, int64_t G, int64_t Dtot, int64_t dofs, int64_t mneg,
                        cudaStream_t st) {
    if (!go || !ref || !tgt) return DV_ERR_NULL;
    if (!gref && !gtgt) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0 || G <= 0 || C % G != 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || G > 65535 || C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    const int cpg = static_cast<int>(C / G);
    // quad kernel: plain gwc volumes, and two-sided volumes (non-negative slots) when every row holds both m-column bands
    const bool two_sided_ok = mneg > 0 && dofs == mneg && Dtot == 2 * mneg + 1 && W >= 2 * mneg;
    if ((mneg == 0 ? (dofs == 0 && Dtot == D) : two_sided_ok) && HW % 4 == 0 && W >= 4 && D <= 256 && aligned16(go) &&
        aligned16(ref) && aligned16(tgt) && (!gref || aligned16(gref)) && (!gtgt || aligned16(gtgt)) &&
        G * static_cast<int64_t>(cpg) <= 65535 && DV_TUNE("DV_GWC_BWD_QUAD", 1)) {
        const int Bi = static_cast<int>(B), Ci = static_cast<int>(C), HWi = static_cast<int>(HW), Wi = static_cast<int>(W),
                  Di = static_cast<int>(D), Gi = static_cast<int>(G), Dt = static_cast<int>(Dtot), Do = static_cast<int>(dofs);
        int rc = DV_ERR_UNSUPPORTED;
        if (cpg % 8 == 0) rc = launch_gwc_bwd_best<8>(go, ref, tgt, gref, gtgt, Bi, Ci, HWi, Wi, Di, Gi, cpg, Dt, Do, st);
        else if (cpg % 6 == 0) rc = launch_gwc_bwd_best<6>(go, ref, tgt, gref, gtgt, Bi, Ci, HWi, Wi, Di, Gi, cpg, Dt, Do, st);
        else if (cpg % 4 == 0) rc = launch_gwc_bwd_best<4>(go, ref, tgt, gref, gtgt, Bi, Ci, HWi, Wi, Di, Gi, cpg, Dt, Do, st);
        if (rc == DV_OK && mneg > 0) {
            if (mneg <= 32 && B * G * H <= INT32_MAX && DV_TUNE("DV_CORR_NEG_ROWS", 1)) {
                const int nrows = static_cast<int>(B * G * H);
                corr_negative_bwd_rows_kernel<<<(nrows + 3) / 4, 128, 0, st>>>(go, ref, tgt, gref, gtgt, Ci, static_cast<int>(H), Wi,
                                                                               static_cast<int>(mneg), Gi, cpg, nrows);
                return finish_launch();
            }
            const int64_t total = B * C * H * mneg;
            corr_negative_bwd_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, st>>>(
                go, ref, tgt, gref, gtgt, Ci, static_cast<int>(H), Wi, static_cast<int>(mneg), Gi, cpg, total);
            return finish_launch();
        }
        if (rc != DV_ERR_UNSUPPORTED) return rc;
    }
    dim3 grid(static_cast<unsigned>((HW + 127) / 128), static_cast<unsigned>(G), static_cast<unsigned>(B));
#define DV_GB(CPGV)                                                                                                   \
    gwc_bwd_kernel<CPGV><<<grid, 128, 0, st>>>(go, ref, tgt, gref, gtgt, static_cast<int>(C), static_cast<int>(HW),   \
                                               static_cast<int>(W), static_cast<int>(D), static_cast<int>(G),         \
                                               static_cast<int>(Dtot), static_cast<int>(dofs), static_cast<int>(mneg))
    switch (cpg) {
        case 1: DV_GB(1); break;
        case 2: DV_GB(2); break;
        case 4: DV_GB(4); break;
        case 8: DV_GB(8); break;
        case 12: DV_GB(12); break;
        case 16: DV_GB(16); break;
        case 32: DV_GB(32); break;
        default: {
            const int64_t total = B * C * HW;
            const int64_t blocks = (total + 255) / 256;
            const int gsz = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
            gwc_bwd_generic_kernel<<<gsz, 256, 0, st>>>(go, ref, tgt, gref, gtgt, static_cast<int>(C), static_cast<int>(HW),
                                                        static_cast<int>(W), static_cast<int>(D), static_cast<int>(G), cpg,
                                                        static_cast<int>(Dtot), static_cast<int>(dofs),
                                                        static_cast<int>(mneg), total);
        }
    }
#undef DV_GB
    return finish_launch();
}

// concat backward: thread = (b, c, pixel); D independent loads per output
__global__ void __launch_bounds__(256)
concat_bwd_kernel(const float *__restrict__ go, float *__restrict__ gref, float *__restrict__ gtgt, int C, int HW, int W,
                  int D, int mask_left) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int c = blockIdx.y, b = blockIdx.z;      // c in [0, 2C)
    const int x = p % W;
    const float *gp = go + ((static_cast<int64_t>(b) * 2 * C + c) * D) * HW;
    float acc = 0.0f;
    if (c < C) {
        if (!gref) return;
        const int dmax = mask_left ? min(D - 1, x) : D - 1;
#pragma unroll 8
        for (int d = 0; d <= dmax; ++d) acc += ldg_stream_f32(gp + static_cast<int64_t>(d) * HW + p);
        gref[(static_cast<int64_t>(b) * C + c) * HW + p] = acc;
    } else {
        if (!gtgt) return;
        const int dmax = min(D - 1, W - 1 - x);
#pragma unroll 8
        for (int d = 0; d <= dmax; ++d) acc += ldg_stream_f32(gp + static_cast<int64_t>(d) * HW + p + d);
        gtgt[(static_cast<int64_t>(b) * C + (c - C)) * HW + p] = acc;
    }
}

// disparity_regression backward: d_x[b,d,p] = d * g[b,p]
template <int V>
__global__ void __launch_bounds__(256)
regression_bwd_kernel(const float *__restrict__ g, float *__restrict__ gx, int D, int HW) {
    const int b = blockIdx.z;
    const int64_t pv = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * V;
    if (pv >= HW) return;
    float gv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) gv[i] = g[static_cast<int64_t>(b) * HW + pv + i];
    // blockIdx.y walks chunks of 8 disparities so that the grid stays many short CTAs (write-stream locality)
    const int d0 = blockIdx.y * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = d0 + j;
        if (d >= D) break;
        float *o = gx + (static_cast<int64_t>(b) * D + d) * HW + pv;
        if constexpr (V == 4) {
            stg_cs(reinterpret_cast<float4 *>(o), make_float4(gv[0] * d, gv[1] * d, gv[2] * d, gv[3] * d));
        } else {
            o[0] = gv[0] * static_cast<float>(d);
        }
    }
}

__global__ void groupwise_bwd_kernel(const float *__restrict__ go, const float *__restrict__ f1, const float *__restrict__ f2,
                                     float *__restrict__ g1, float *__restrict__ g2, int C, int HW, int G, int cpg,
                                     int64_t total) {
    const float inv = 1.0f / static_cast<float>(cpg);
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t p = idx % HW, bc = idx / HW;
        const int64_t c = bc % C, b = bc / C;
        const float ge = go[(b * G + c / cpg) * HW + p] * inv;
        if (g1) g1[idx] = ge * f2[idx];
        if (g2) g2[idx] = ge * f1[idx];
    }
}

}  // namespace dv

extern "C" int dv_gwc_volume_bwd_f32(const float *grad_out, const float *ref, const float *tgt, float *grad_ref,
                                     float *grad_tgt, int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G,
                                     void *stream) {
    return dv::gwc_bwd_impl(grad_out, ref, tgt, grad_ref, grad_tgt, B, C, H, W, D, G, D, 0, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int dv_corr_volume_2sided_bwd_f32(const float *grad_out, const float *ref, const float *tgt, float *grad_ref,
                                             float *grad_tgt, int64_t B, int64_t C, int64_t H, int64_t W,
                                             int64_t maxdisp, int64_t G, void *stream) {
    if (maxdisp < 0) return DV_ERR_BAD_SHAPE;
    return dv::gwc_bwd_impl(grad_out, ref, tgt, grad_ref, grad_tgt, B, C, H, W, maxdisp + 1, G, 2 * maxdisp + 1, maxdisp,
                            maxdisp, static_cast<cudaStream_t>(stream));
}

extern "C" int dv_groupwise_correlation_bwd_f32(const float *grad_out, const float *fea1, const float *fea2, float *grad1,
                                                float *grad2, int64_t B, int64_t C, int64_t H, int64_t W, int64_t G,
                                                void *stream) {
    using namespace dv;
    if (!grad_out || !fea1 || !fea2 || (!grad1 && !grad2)) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || G <= 0 || C % G != 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W, total = B * C * HW;
    if (HW > INT32_MAX) return DV_ERR_BAD_SHAPE;
    const int64_t blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
    groupwise_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, fea1, fea2, grad1, grad2,
                                                                            static_cast<int>(C), static_cast<int>(HW),
                                                                            static_cast<int>(G), static_cast<int>(C / G), total);
    return finish_launch();
}

extern "C" int dv_concat_volume_bwd_f32(const float *grad_out, float *grad_ref, float *grad_tgt, int64_t B, int64_t C,
                                        int64_t H, int64_t W, int64_t D, int mask_left, void *stream) {
    using namespace dv;
    if (!grad_out || (!grad_ref && !grad_tgt)) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || 2 * C > 65535) return DV_ERR_BAD_SHAPE;
    dim3 grid(static_cast<unsigned>((HW + 255) / 256), static_cast<unsigned>(2 * C), static_cast<unsigned>(B));
    concat_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, grad_ref, grad_tgt, static_cast<int>(C),
                                                                         static_cast<int>(HW), static_cast<int>(W),
                                                                         static_cast<int>(D), mask_left);
    return finish_launch();
}

extern "C" int dv_disparity_regression_bwd_f32(const float *grad_out, float *grad_x, int64_t B, int64_t D, int64_t H,
                                               int64_t W, void *stream) {
    using namespace dv;
    if (!grad_out || !grad_x) return DV_ERR_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || (D + 7) / 8 > 65535) return DV_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((HW % 4 == 0) && aligned16(grad_out) && aligned16(grad_x)) {
        dim3 grid(static_cast<unsigned>((HW / 4 + 255) / 256), static_cast<unsigned>((D + 7) / 8), static_cast<unsigned>(B));
        regression_bwd_kernel<4><<<grid, 256, 0, st>>>(grad_out, grad_x, static_cast<int>(D), static_cast<int>(HW));
    } else {
        dim3 grid(static_cast<unsigned>((HW + 255) / 256), static_cast<unsigned>((D + 7) / 8), static_cast<unsigned>(B));
        regression_bwd_kernel<1><<<grid, 256, 0, st>>>(grad_out, grad_x, static_cast<int>(D), static_cast<int>(HW));
    }
    return finish_launch();
}
