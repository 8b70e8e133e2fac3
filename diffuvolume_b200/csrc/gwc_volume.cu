// gwc_volume.cu — group-wise correlation volume (a1, a2) for sm_100a.
//
// Replaces build_gwc_volume / groupwise_correlation of the reference
// (SceneFlow/models/submodule.py:209-215,228-238; KITTI12/models/submodule.py:100-119;
//  KITTI15/core/submodule.py:151-169), which launches 1 memset + 3*D ATen kernels and
// re-reads both feature maps D times.  Here every feature byte is staged ONCE into shared
// memory by the TMA engine (1-D bulk copies, cp.async.bulk -> SASS UBLKCP) and reused for
// all D shifts; the [B,G,D,H,W] volume is written exactly once with 128-bit streaming
// stores, zeros for x < d included (no memset).
//
// Data layout trick: a channel plane [H,W] is contiguous, so the kernel tiles the FLATTENED
// plane into spans of SQ quads (4 floats).  tgt[y, x-d] is just flat index p-d; whenever
// that index crosses a row start (x < d) the output is defined to be 0 and is selected,
// never multiplied, so whatever the staged window holds there (previous row / previous
// plane) is irrelevant.  This removes all row-tail handling and keeps every bulk copy and
// every store 16-byte aligned for any W with (H*W) % 4 == 0.
//
// Work decomposition: CTA = (span, group g, batch b); thread = (quad q, disparity chunk):
// it keeps a DC x 4 accumulator tile in registers and, per channel k, reads one float4 of
// ref and a (DC+4)-float sliding window of tgt from shared memory (conflict-free: lanes map
// to consecutive quads).  A warp's stores for one d cover 512 contiguous bytes.
#include "common.cuh"

namespace dv {

template <int CPG, int DC, int SQ, int NCH, int MINB, typename OutT>
__global__ void __launch_bounds__(SQ * NCH, MINB)
gwc_volume_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, OutT *__restrict__ out,
                  int C, int HW, int W, int D, int G, int Dpad, int Dtot, int dofs, int tiles_per_cta) {
    // Dtot = planes per (b,g) in `out`, dofs = slot of shift 0 (plain gwc: Dtot = D, dofs = 0;
    // two-sided correlation volume: Dtot = 2m+1, dofs = m).
    // A CTA walks `tiles_per_cta` consecutive spans of one (b, g) plane with a 2-stage pipeline: the bulk
    // copies of span i+1 are in flight while span i is being correlated and stored.
    static_assert(DC % 4 == 0, "DC must be a multiple of 4");
    constexpr int SPAN = SQ * 4;
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar[2];
    const int rpitch = SPAN + Dpad;
    const int stage_floats = CPG * SPAN + CPG * rpitch;  // [CPG][SPAN] ref | [CPG][SPAN + Dpad] tgt window

    const int b = blockIdx.z, g = blockIdx.y;
    const int nspans = (HW + SPAN - 1) / SPAN;
    const int t0 = blockIdx.x * tiles_per_cta;
    const int nt = min(tiles_per_cta, nspans - t0);
    const int64_t plane0 = (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * CPG) * HW;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int i) {  // called by warp 0 only
        float *sL = smem + (i & 1) * stage_floats;
        float *sR = sL + CPG * SPAN;  // element Dpad of a row is flat index p0
        const int p0 = (t0 + i) * SPAN;
        const int len = min(SPAN, HW - p0);  // multiple of 4 (HW % 4 == 0)
        // The tgt window of channel k starts Dpad floats before the span: roff0 + k*HW.  For (b=0, g=0) that lies before
        // the tensor while k*HW + p0 < Dpad (the first plane; several planes when HW < Dpad): each copy is clipped to the
        // tensor (the skipped values are never used: x < d there, and the forward only selects).
        const int64_t roff0 = plane0 + p0 - Dpad;
        auto skip_of = [&](int k) -> int {
            const int64_t o = roff0 + static_cast<int64_t>(k) * HW;
            return o < 0 ? static_cast<int>(o < -(len + Dpad) ? len + Dpad : -o) : 0;
        };
        if (threadIdx.x == 0) {
            int skipped = 0;
            for (int k = 0; k < CPG; ++k) skipped += skip_of(k);
            const uint32_t bytes = static_cast<uint32_t>(CPG) * (2u * len + Dpad) * 4u - 4u * skipped;
            mbar_expect_tx(&bar[i & 1], bytes);
        }
        __syncwarp();
        for (int k = threadIdx.x; k < CPG; k += 32) {
            bulk_g2s(sL + k * SPAN, ref + plane0 + static_cast<int64_t>(k) * HW + p0, 4u * len, &bar[i & 1]);
            const int skip = skip_of(k);
            if (skip < len + Dpad)
                bulk_g2s(sR + k * rpitch + skip, tgt + roff0 + static_cast<int64_t>(k) * HW + skip,
                         4u * (len + Dpad - skip), &bar[i & 1]);
        }
    };

    if (threadIdx.x < 32) issue(0);

    const int q = threadIdx.x % SQ;
    const int ch0 = threadIdx.x / SQ;
    constexpr float inv = 1.0f / CPG;  // mean over the group (torch's mean multiplies by 1/n on CUDA)

    for (int it = 0; it < nt; ++it) {
        // stage (it+1)&1 was last read in iteration it-1; the barrier at the end of that iteration makes
        // it safe to refill now
        if (threadIdx.x < 32 && it + 1 < nt) issue(it + 1);
        mbar_wait(&bar[it & 1], (it >> 1) & 1);

        const float *sL = smem + (it & 1) * stage_floats;
        const float *sR = sL + CPG * SPAN;
        const int p0 = (t0 + it) * SPAN;
        const int p = p0 + 4 * q;
        if (p < HW) {
            int xs[4];
            xs[0] = p % W;
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                xs[i] = xs[i - 1] + 1;
                if (xs[i] >= W) xs[i] -= W;
            }
            for (int ch = ch0; ch * DC < D; ch += NCH) {
                const int d0 = ch * DC;
                float acc[DC][4];
#pragma unroll
                for (int j = 0; j < DC; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;

                const float *lp = sL + 4 * q;
                const float *rp = sR + Dpad + 4 * q - d0 - DC;  // 16-byte aligned: Dpad, d0, DC % 4 == 0
#pragma unroll
                for (int k = 0; k < CPG; ++k) {
                    const float4 l4 = *reinterpret_cast<const float4 *>(lp + k * SPAN);
                    const float l[4] = {l4.x, l4.y, l4.z, l4.w};
                    float rw[DC + 4];
#pragma unroll
                    for (int m = 0; m < DC / 4 + 1; ++m) {
                        const float4 r4 = *reinterpret_cast<const float4 *>(rp + k * rpitch + 4 * m);
                        rw[4 * m + 0] = r4.x;
                        rw[4 * m + 1] = r4.y;
                        rw[4 * m + 2] = r4.z;
                        rw[4 * m + 3] = r4.w;
                    }
                    // out(d0+j, x+i) += ref(x+i) * tgt(x+i-d0-j);  window index = DC + i - j
#pragma unroll
                    for (int j = 0; j < DC; ++j)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(l[i], rw[DC + i - j], acc[j][i]);
                }

                OutT *op = out + ((static_cast<int64_t>(b) * G + g) * Dtot + dofs + d0) * HW + p;
#pragma unroll
                for (int j = 0; j < DC; ++j) {
                    const int d = d0 + j;
                    if (d < D) {
                        float4 v;
                        v.x = xs[0] >= d ? acc[j][0] * inv : 0.0f;
                        v.y = xs[1] >= d ? acc[j][1] * inv : 0.0f;
                        v.z = xs[2] >= d ? acc[j][2] * inv : 0.0f;
                        v.w = xs[3] >= d ? acc[j][3] * inv : 0.0f;
                        store4_cs(op + static_cast<int64_t>(j) * HW, v);
                    }
                }
            }
        }
        if (it + 1 < nt) __syncthreads();
    }
}

// Wide groups (channels per group a multiple of 8, > 8): the same span / sliding-window scheme, but the K dimension
// is streamed through shared memory in chunks of KC channels while the DC x 4 accumulator tile stays in registers, so the
// staging buffer is 17 KB per stage whatever cpg is (the one-shot kernel above needs cpg x 2.2 KB: 140 KB at cpg = 32,
// one 128-thread CTA per SM).  Used by PCWNet's refinement correlation (C = 32, G = 1, KITTI12/models/pwcnet_ddim.py:494).
// Every thread owns ONE disparity chunk (NCH = number of chunks), pipeline stage = (span, k-chunk).
template <int KC, int DC, int SQ, int NCH, int MINB, int STAGES>
__global__ void __launch_bounds__(SQ * NCH, MINB)
gwc_volume_kchunk_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, float *__restrict__ out,
                         int C, int HW, int W, int D, int G, int cpg, int Dpad, int Dtot, int dofs, int tiles_per_cta) {
    // STAGES-deep ring with full / empty mbarriers (no CTA-wide barrier per stage: ncu on the __syncthreads version showed
    // the barrier as the top stall, each stage being only KC x 48 FMA per thread).  Thread 0 is the producer: before it
    // computes stage st it refills the slot of stage st-1 with stage st+STAGES-1, which needs every warp to have released
    // that slot — they are at most one stage behind.
    constexpr int SPAN = SQ * 4;
    constexpr int NWARPS = SQ * NCH / 32;
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    const int rpitch = SPAN + Dpad;
    const int stage_floats = KC * SPAN + KC * rpitch;
    const int b = blockIdx.z, g = blockIdx.y;
    const int nspans = (HW + SPAN - 1) / SPAN;
    const int t0 = blockIdx.x * tiles_per_cta;
    const int nt = min(tiles_per_cta, nspans - t0);
    const int nkc = cpg / KC;
    const int nstages = nt * nkc;
    const int64_t plane0 = (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg) * HW;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], NWARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int st) {  // thread 0 only
        const int slot = st % STAGES;
        float *sL = smem + slot * stage_floats;
        float *sR = sL + KC * SPAN;
        const int p0 = (t0 + st / nkc) * SPAN;
        const int k0 = (st % nkc) * KC;
        const int len = min(SPAN, HW - p0);
        const int64_t roff0 = plane0 + static_cast<int64_t>(k0) * HW + p0 - Dpad;
        // clip every channel's window to the tensor (first planes of (b=0, g=0); several planes when HW < Dpad)
        auto skip_of = [&](int k) -> int {
            const int64_t o = roff0 + static_cast<int64_t>(k) * HW;
            return o < 0 ? static_cast<int>(o < -(len + Dpad) ? len + Dpad : -o) : 0;
        };
        int skipped = 0;
#pragma unroll
        for (int k = 0; k < KC; ++k) skipped += skip_of(k);
        mbar_expect_tx(&full_bar[slot], static_cast<uint32_t>(KC) * (2u * len + Dpad) * 4u - 4u * skipped);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            bulk_g2s(sL + k * SPAN, ref + plane0 + static_cast<int64_t>(k0 + k) * HW + p0, 4u * len, &full_bar[slot]);
            const int skip = skip_of(k);
            if (skip < len + Dpad)
                bulk_g2s(sR + k * rpitch + skip, tgt + roff0 + static_cast<int64_t>(k) * HW + skip, 4u * (len + Dpad - skip),
                         &full_bar[slot]);
        }
    };
    if (threadIdx.x == 0)
        for (int st = 0; st < STAGES - 1 && st < nstages; ++st) issue(st);

    const int q = threadIdx.x % SQ;
    const int d0 = (threadIdx.x / SQ) * DC;
    const float inv = 1.0f / static_cast<float>(cpg);
    float acc[DC][4];
    for (int st = 0; st < nstages; ++st) {
        const int slot = st % STAGES;
        if (threadIdx.x == 0 && st + STAGES - 1 < nstages) {
            const int nxt = st + STAGES - 1;              // goes into the slot stage st-1 used
            if (st >= 1) mbar_wait(&empty_bar[nxt % STAGES], ((st - 1) / STAGES) & 1);
            issue(nxt);
        }
        mbar_wait(&full_bar[slot], (st / STAGES) & 1);
        const int kc = st % nkc;
        const float *sL = smem + slot * stage_floats;
        const float *sR = sL + KC * SPAN;
        const int p = (t0 + st / nkc) * SPAN + 4 * q;
        if (p < HW && d0 < D) {
            if (kc == 0) {
#pragma unroll
                for (int j = 0; j < DC; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j][i] = 0.0f;
            }
            const float *lp = sL + 4 * q;
            const float *rp = sR + Dpad + 4 * q - d0 - DC;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const float4 l4 = *reinterpret_cast<const float4 *>(lp + k * SPAN);
                const float l[4] = {l4.x, l4.y, l4.z, l4.w};
                float rw[DC + 4];
#pragma unroll
                for (int m = 0; m < DC / 4 + 1; ++m) {
                    const float4 r4 = *reinterpret_cast<const float4 *>(rp + k * rpitch + 4 * m);
                    rw[4 * m + 0] = r4.x; rw[4 * m + 1] = r4.y; rw[4 * m + 2] = r4.z; rw[4 * m + 3] = r4.w;
                }
#pragma unroll
                for (int j = 0; j < DC; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(l[i], rw[DC + i - j], acc[j][i]);
            }
        }
        // this warp is done reading the slot
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[slot]);
        if (p < HW && d0 < D && kc == nkc - 1) {
            int xs[4];
            xs[0] = p % W;
#pragma unroll
            for (int i = 1; i < 4; ++i) {
                xs[i] = xs[i - 1] + 1;
                if (xs[i] >= W) xs[i] -= W;
            }
            float *op = out + ((static_cast<int64_t>(b) * G + g) * Dtot + dofs + d0) * HW + p;
#pragma unroll
            for (int j = 0; j < DC; ++j) {
                const int d = d0 + j;
                if (d < D) {
                    float4 v;
                    v.x = xs[0] >= d ? acc[j][0] * inv : 0.0f;
                    v.y = xs[1] >= d ? acc[j][1] * inv : 0.0f;
                    v.z = xs[2] >= d ? acc[j][2] * inv : 0.0f;
                    v.w = xs[3] >= d ? acc[j][3] * inv : 0.0f;
                    stg_cs(reinterpret_cast<float4 *>(op + static_cast<int64_t>(j) * HW), v);
                }
            }
        }
    }
}

// Shape-agnostic kernel for planes with (H*W) % 4 != 0, unaligned pointers or an unusual
// channels-per-group: one thread per output element, coalesced along x.  <1 % of the bytes
// of any reference configuration ever take this path.
__global__ void gwc_volume_generic_kernel(const float *__restrict__ ref, const float *__restrict__ tgt,
                                          float *__restrict__ out, int C, int HW, int W, int D, int G,
                                          int cpg, int Dtot, int dofs, int64_t total) {
    const float inv = 1.0f / static_cast<float>(cpg);
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % HW);
        int64_t t = idx / HW;
        const int d = static_cast<int>(t % D);
        t /= D;
        const int g = static_cast<int>(t % G);
        const int64_t b = t / G;
        const int x = p % W;
        float v = 0.0f;
        if (x >= d) {
            const float *l = ref + (b * C + static_cast<int64_t>(g) * cpg) * HW + p;
            const float *r = tgt + (b * C + static_cast<int64_t>(g) * cpg) * HW + p - d;
            float acc = 0.0f;
            for (int k = 0; k < cpg; ++k) acc = fmaf(l[static_cast<int64_t>(k) * HW], r[static_cast<int64_t>(k) * HW], acc);
            v = acc * inv;
        }
        out[((b * G + g) * Dtot + dofs + d) * HW + p] = v;
    }
}

// Negative shifts of build_corrleation_volume (KITTI12/models/submodule.py:128-131): for
// i = -k the reference writes `volume[..., :-i] = gwc(ref[..., :-i], tgt[..., i:])`, i.e. only
// the FIRST k columns, pairing ref[x] with tgt[W-k+x]; everything else in the plane stays 0.
__global__ void corr_negative_kernel(const float *__restrict__ ref, const float *__restrict__ tgt,
                                     float *__restrict__ out, int C, int HW, int W, int m, int G, int cpg,
                                     int64_t total, int Dtot, int dofs0) {
    const float inv = 1.0f / static_cast<float>(cpg);
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % HW);
        int64_t t = idx / HW;
        const int slot = static_cast<int>(t % m);  // i = slot - m, k = m - slot
        t /= m;
        const int g = static_cast<int>(t % G);
        const int64_t b = t / G;
        const int k = m - slot;
        const int x = p % W;
        float v = 0.0f;
        if (x < k) {  // k > W: both slices degenerate to the whole row (offset 0)
            const float *l = ref + (b * C + static_cast<int64_t>(g) * cpg) * HW + p;
            const float *r = tgt + (b * C + static_cast<int64_t>(g) * cpg) * HW + p + max(W - k, 0);
            float acc = 0.0f;
            for (int c = 0; c < cpg; ++c) acc = fmaf(l[static_cast<int64_t>(c) * HW], r[static_cast<int64_t>(c) * HW], acc);
            v = acc * inv;
        }
        out[((b * G + g) * Dtot + dofs0 + slot) * HW + p] = v;
    }
}

// The same for 16-byte aligned planes with HW % 4 == 0, split in two: >98 % of the negative-shift planes is zeros, so
// (1) a pure 128-bit store stream clears planes [0, m) of every (b, g), then (2) one thread per LIVE element (x < k)
// computes its dot product — every thread of that launch has work, instead of one slow lane per warp of a store loop.
__global__ void __launch_bounds__(256)
zero_planes_kernel(float *__restrict__ out, int64_t plane_stride, int64_t quads_per_bg) {
    float *o = out + static_cast<int64_t>(blockIdx.y) * plane_stride;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t qd = static_cast<int64_t>(blockIdx.x) * 1024 + u * 256 + threadIdx.x;
        if (qd < quads_per_bg) stg_cs(reinterpret_cast<float4 *>(o) + qd, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    }
}
__global__ void __launch_bounds__(128)
corr_negative_live_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, float *__restrict__ out, int C,
                          int H, int W, int m, int G, int cpg, int Dtot, int dofs0) {
    // grid: x = row chunks, y = slot, z = b*G + g; thread = (row, column x < min(k, W))
    const int slot = blockIdx.y, k = m - slot;
    const int kw = min(k, W);
    const int rows_per_cta = max(128 / kw, 1);
    const int y = blockIdx.x * rows_per_cta + threadIdx.x / kw;
    const int x = threadIdx.x % kw;
    if (threadIdx.x >= rows_per_cta * kw || y >= H) return;
    const int g = blockIdx.z % G, b = blockIdx.z / G;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const int64_t p = static_cast<int64_t>(y) * W + x;
    const float *l = ref + (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg) * HW + p;
    const float *r = tgt + (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg) * HW + p + max(W - k, 0);
    float acc = 0.0f;
#pragma unroll 8
    for (int c = 0; c < cpg; ++c) acc = fmaf(__ldg(l + c * HW), __ldg(r + c * HW), acc);
    out[(static_cast<int64_t>(blockIdx.z) * Dtot + dofs0 + slot) * HW + p] = acc * (1.0f / static_cast<float>(cpg));
}

// The live elements one image ROW at a time (m <= 32 <= W): a warp owns row y of (b, g), lane x < m holds ref[c][y][x] and
// tgt[c][y][W-m+x] of the current channel, and slot k's element x < k pairs ref[x] with tgt[W-k+x] = lane (m-k+x)'s value
// (one shuffle).  2 coalesced loads per channel and lane instead of 2 gathers per (slot, channel) — the per-element kernel
// above spends 29 us on 0.5 M dot products at B = 4, 384x1248.  Same summation order over c: bit-identical.
__global__ void __launch_bounds__(128)
corr_negative_rows_kernel(const float *__restrict__ ref, const float *__restrict__ tgt, float *__restrict__ out, int C, int H,
                          int W, int m, int G, int cpg, int Dtot, int dofs0, int nrows) {
    const int row = blockIdx.x * 4 + threadIdx.x / 32, x = threadIdx.x % 32;
    if (row >= nrows) return;
    const int y = row % H, bg = row / H, g = bg % G, b = bg / G;
    const int64_t HW = static_cast<int64_t>(H) * W;
    const bool live = x < m;
    const float *l = ref + (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg) * HW + static_cast<int64_t>(y) * W + (live ? x : 0);
    const float *r = tgt + (static_cast<int64_t>(b) * C + static_cast<int64_t>(g) * cpg) * HW + static_cast<int64_t>(y) * W + W - m + (live ? x : 0);
    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0f;
    for (int c0 = 0; c0 < cpg; c0 += 16) {          // 32 loads in flight per lane before the first use
        float rv[16], tv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int c = min(c0 + u, cpg - 1);
            rv[u] = __ldg(l + c * HW);
            tv[u] = __ldg(r + c * HW);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (c0 + u >= cpg) break;               // uniform
#pragma unroll
            for (int k = 1; k <= 32; ++k) {
                if (k > m) break;                   // uniform
                const float t = __shfl_sync(0xffffffffu, tv[u], (m - k + x) & 31);
                acc[k - 1] = fmaf(rv[u], t, acc[k - 1]);   // meaningful for x < k only
            }
        }
    }
    const float inv = 1.0f / static_cast<float>(cpg);
    float *o = out + (static_cast<int64_t>(bg) * Dtot + dofs0) * HW + static_cast<int64_t>(y) * W + x;
#pragma unroll
    for (int k = 1; k <= 32; ++k)
        if (k <= m && x < k) o[static_cast<int64_t>(m - k) * HW] = acc[k - 1] * inv;
}

template <int KC, int DC, int SQ, int NCH, int MINB>
static int launch_gwc_kchunk(const float *ref, const float *tgt, float *out, int B, int C, int HW, int W, int D, int G,
                             int cpg, int Dtot, int dofs, cudaStream_t st) {
    constexpr int SPAN = SQ * 4;
    const int Dpad = ((D + DC - 1) / DC) * DC;
    const int nspans = (HW + SPAN - 1) / SPAN;
    int tpc = DV_TUNE("DV_GWC_TPC", 8);
    while (tpc > 1 && static_cast<int64_t>((nspans + tpc - 1) / tpc) * G * B < 4LL * num_sms() * MINB) tpc /= 2;
    constexpr int STAGES = 3;
    const size_t smem = sizeof(float) * STAGES * (static_cast<size_t>(KC) * SPAN + static_cast<size_t>(KC) * (SPAN + Dpad));
    auto kern = gwc_volume_kchunk_kernel<KC, DC, SQ, NCH, MINB, STAGES>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return DV_ERR_LAUNCH;
    dim3 grid((nspans + tpc - 1) / tpc, G, B);
    kern<<<grid, SQ * NCH, smem, st>>>(ref, tgt, out, C, HW, W, D, G, cpg, Dpad, Dtot, dofs, tpc);
    return finish_launch();
}

static int dispatch_gwc_kchunk(const float *ref, const float *tgt, float *out, int B, int C, int HW, int W, int D, int G,
                               int cpg, int Dtot, int dofs, cudaStream_t st) {
    constexpr int KC = 8, DC = 12, SQ = 64;
    // 24 < D <= 32 (the +-24 refinement volume: D = 25): one 28-wide (or two 16-wide) disparity chunks instead of three 12-wide
    // ones, the third of which would hold a single plane — the kernel is shared-memory-bandwidth bound (5 LDS.128 per 48 FMA
    // per thread and chunk), and the wide tile needs 9 LDS.128 per 112 FMA: 0.190 -> 0.137 ms at B = 4, 384x1248.
    // (The same widening does nothing for the cpg = 8, D = 48 volume, which is not LDS-bound: measured 0.463 vs 0.456 ms.)
    const int wide = DV_TUNE("DV_KCHUNK_WIDE", 28);
    if (D > 24 && D <= 28 && wide == 28) return launch_gwc_kchunk<KC, 28, SQ, 1, 4>(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
    if (D > 24 && D <= 32 && wide != 0) return launch_gwc_kchunk<KC, 16, SQ, 2, 3>(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
    const int nch = (D + DC - 1) / DC;
    switch (nch) {
        case 1: return launch_gwc_kchunk<KC, DC, SQ, 1, 8>(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
        case 2: return launch_gwc_kchunk<KC, DC, SQ, 2, 4>(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
        case 3: return launch_gwc_kchunk<KC, DC, SQ, 3, 3>(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
        case 4: return launch_gwc_kchunk<KC, DC, SQ, 4, 2>(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
        default: return DV_ERR_UNSUPPORTED;
    }
}

template <int CPG, int DC, int SQ, int NCH, int MINB, typename OutT>
static int launch_gwc(const float *ref, const float *tgt, OutT *out, int B, int C, int HW, int W, int D, int G,
                      int Dtot, int dofs, cudaStream_t st) {
    constexpr int SPAN = SQ * 4;
    const int Dpad = ((D + DC - 1) / DC) * DC;
    const int nspans = (HW + SPAN - 1) / SPAN;
    // spans per CTA (2-stage pipeline inside the CTA); keep >= ~4 CTAs per SM slot for balance
    int tpc = DV_TUNE("DV_GWC_TPC", 8);
    while (tpc > 1 && static_cast<int64_t>((nspans + tpc - 1) / tpc) * G * B < 4LL * num_sms() * MINB) tpc /= 2;
    const int stages = tpc > 1 ? 2 : 1;
    const size_t smem = sizeof(float) * stages * (static_cast<size_t>(CPG) * SPAN + static_cast<size_t>(CPG) * (SPAN + Dpad));
    if (smem > 200 * 1024) return DV_ERR_UNSUPPORTED;
    auto kern = gwc_volume_kernel<CPG, DC, SQ, NCH, MINB, OutT>;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
            return DV_ERR_LAUNCH;
    }
    dim3 grid((nspans + tpc - 1) / tpc, G, B);
    kern<<<grid, SQ * NCH, smem, st>>>(ref, tgt, out, C, HW, W, D, G, Dpad, Dtot, dofs, tpc);
    return finish_launch();
}

template <int CPG, typename OutT>
static int dispatch_gwc(const float *ref, const float *tgt, OutT *out, int B, int C, int HW, int W, int D, int G,
                        int Dtot, int dofs, cudaStream_t st) {
    constexpr int DC = 12, SQ = 64;
    const int nch = (D + DC - 1) / DC;
    if (nch >= 4) {
        if (DV_TUNE("DV_GWC_MINB", 2) == 2)
            return launch_gwc<CPG, DC, SQ, 4, 2, OutT>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st);
        return launch_gwc<CPG, DC, SQ, 4, 3, OutT>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st);
    }
    if (nch >= 2) return launch_gwc<CPG, DC, SQ, 2, 4, OutT>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st);
    return launch_gwc<CPG, DC, SQ, 1, 8, OutT>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st);
}

static int gwc_volume_impl(const float *ref, const float *tgt, float *out, int64_t B, int64_t C, int64_t H, int64_t W,
                           int64_t D, int64_t G, int64_t Dtot, int64_t dofs, cudaStream_t st) {
    if (!ref || !tgt || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0 || G <= 0 || C % G != 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || G > 65535 || C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    const int cpg = static_cast<int>(C / G);
    const bool fast = (HW % 4 == 0) && aligned16(ref) && aligned16(tgt) && aligned16(out) && HW >= 64 && D <= 512 &&
                      W >= 4;
    if (fast) {
        int rc = DV_ERR_UNSUPPORTED;
        if (cpg > 8 && cpg % 8 == 0 && D <= 48 && DV_TUNE("DV_GWC_KCHUNK", 1))
            rc = dispatch_gwc_kchunk(ref, tgt, out, B, C, HW, W, D, G, cpg, Dtot, dofs, st);
        if (rc != DV_ERR_UNSUPPORTED) return rc;
        switch (cpg) {
            case 4: rc = dispatch_gwc<4, float>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st); break;
            case 8: rc = dispatch_gwc<8, float>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st); break;
            case 12: rc = dispatch_gwc<12, float>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st); break;
            case 16: rc = dispatch_gwc<16, float>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st); break;
            case 32: rc = dispatch_gwc<32, float>(ref, tgt, out, B, C, HW, W, D, G, Dtot, dofs, st); break;
            default: break;
        }
        if (rc != DV_ERR_UNSUPPORTED) return rc;
    }
    const int64_t total = B * G * D * HW;
    const int threads = 256;
    const int64_t blocks = (total + threads - 1) / threads;
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
    gwc_volume_generic_kernel<<<grid, threads, 0, st>>>(ref, tgt, out, static_cast<int>(C), static_cast<int>(HW),
                                                        static_cast<int>(W), static_cast<int>(D), static_cast<int>(G), cpg,
                                                        static_cast<int>(Dtot), static_cast<int>(dofs), total);
    return finish_launch();
}

// bf16 volume (fp32 features in, fp32 accumulate, one rounding at the store): the 128-bit-path shapes only
static int gwc_volume_bf16_impl(const float *ref, const float *tgt, __nv_bfloat16 *out, int64_t B, int64_t C, int64_t H,
                                int64_t W, int64_t D, int64_t G, cudaStream_t st) {
    if (!ref || !tgt || !out) return DV_ERR_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || D <= 0 || G <= 0 || C % G != 0) return DV_ERR_BAD_SHAPE;
    const int64_t HW = H * W;
    if (HW > INT32_MAX || B > 65535 || G > 65535 || C > INT32_MAX) return DV_ERR_BAD_SHAPE;
    if (!((HW % 4 == 0) && aligned16(ref) && aligned16(tgt) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0)) return DV_ERR_MISALIGNED;
    if (!(HW >= 64 && D <= 512 && W >= 4)) return DV_ERR_UNSUPPORTED;
    const int Bi = static_cast<int>(B), Ci = static_cast<int>(C), HWi = static_cast<int>(HW), Wi = static_cast<int>(W),
              Di = static_cast<int>(D), Gi = static_cast<int>(G);
    switch (C / G) {
        case 8: return dispatch_gwc<8, __nv_bfloat16>(ref, tgt, out, Bi, Ci, HWi, Wi, Di, Gi, Di, 0, st);
        case 12: return dispatch_gwc<12, __nv_bfloat16>(ref, tgt, out, Bi, Ci, HWi, Wi, Di, Gi, Di, 0, st);
        default: return DV_ERR_UNSUPPORTED;   // reference configurations: cpg = 8 (ACVNet, PCWNet), 12 (IGEV)
    }
}

}  // namespace dv

extern "C" int dv_gwc_volume_bf16(const float *ref, const float *tgt, void *out, int64_t B, int64_t C, int64_t H, int64_t W,
                                  int64_t D, int64_t G, void *stream) {
    return dv::gwc_volume_bf16_impl(ref, tgt, static_cast<__nv_bfloat16 *>(out), B, C, H, W, D, G,
                                    static_cast<cudaStream_t>(stream));
}

extern "C" int dv_gwc_volume_f32(const float *ref, const float *tgt, float *out, int64_t B, int64_t C, int64_t H,
                                 int64_t W, int64_t D, int64_t G, void *stream) {
    return dv::gwc_volume_impl(ref, tgt, out, B, C, H, W, D, G, D, 0, static_cast<cudaStream_t>(stream));
}

// a1: a single shift-0 plane of the volume is exactly groupwise_correlation.
extern "C" int dv_groupwise_correlation_f32(const float *fea1, const float *fea2, float *out, int64_t B, int64_t C,
                                            int64_t H, int64_t W, int64_t G, void *stream) {
    return dv::gwc_volume_impl(fea1, fea2, out, B, C, H, W, 1, G, 1, 0, static_cast<cudaStream_t>(stream));
}

// a5: slots [m, 2m] are a gwc volume with D = m+1 written at plane offset m of a (2m+1)-plane
// output; slots [0, m) are the reference's first-k-columns quirk.  Dtot = planes between consecutive (b, g) blocks of
// `out`, dofs0 = plane offset of slot 0 inside a block: (2m+1, 0) for a contiguous volume; for G == 1 any
// (planes per sample, channel offset) addresses a channel slice of a larger [B, planes, H, W] buffer.
static int corr2_impl(const float *ref, const float *tgt, float *out, int64_t B, int64_t C, int64_t H, int64_t W,
                      int64_t maxdisp, int64_t G, int64_t Dtot, int64_t dofs0, cudaStream_t st) {
    using namespace dv;
    if (maxdisp < 0) return DV_ERR_BAD_SHAPE;
    const int rc = gwc_volume_impl(ref, tgt, out, B, C, H, W, maxdisp + 1, G, Dtot, dofs0 + maxdisp, st);
    if (rc != DV_OK || maxdisp == 0) return rc;
    const int64_t HW = H * W;
    const int64_t total = B * G * maxdisp * HW;
    if (HW % 4 == 0 && aligned16(out) && maxdisp <= 128 && B * G <= 65535 && HW <= INT32_MAX) {
        const int64_t quads_per_bg = maxdisp * HW / 4;
        dim3 zgrid(static_cast<unsigned>((quads_per_bg + 1023) / 1024), static_cast<unsigned>(B * G));
        zero_planes_kernel<<<zgrid, 256, 0, st>>>(out + dofs0 * HW, Dtot * HW, quads_per_bg);
        if (maxdisp <= 32 && W >= maxdisp && B * G * H <= INT32_MAX && DV_TUNE("DV_CORR_NEG_ROWS", 1)) {
            const int nrows = static_cast<int>(B * G * H);
            corr_negative_rows_kernel<<<(nrows + 3) / 4, 128, 0, st>>>(ref, tgt, out, static_cast<int>(C), static_cast<int>(H),
                                                                       static_cast<int>(W), static_cast<int>(maxdisp),
                                                                       static_cast<int>(G), static_cast<int>(C / G),
                                                                       static_cast<int>(Dtot), static_cast<int>(dofs0), nrows);
            return finish_launch(2);
        }
        // rows per CTA for the widest slot (k = m) bound the grid; narrower slots use fewer of their CTAs' threads
        dim3 lgrid(static_cast<unsigned>(H), static_cast<unsigned>(maxdisp), static_cast<unsigned>(B * G));
        corr_negative_live_kernel<<<lgrid, 128, 0, st>>>(ref, tgt, out, static_cast<int>(C), static_cast<int>(H),
                                                         static_cast<int>(W), static_cast<int>(maxdisp), static_cast<int>(G),
                                                         static_cast<int>(C / G), static_cast<int>(Dtot), static_cast<int>(dofs0));
        return finish_launch(2);
    }
    const int64_t blocks = (total + 255) / 256;
    const int grid = static_cast<int>(blocks < static_cast<int64_t>(num_sms()) * 32 ? blocks : static_cast<int64_t>(num_sms()) * 32);
    corr_negative_kernel<<<grid, 256, 0, st>>>(ref, tgt, out, static_cast<int>(C), static_cast<int>(HW),
                                               static_cast<int>(W), static_cast<int>(maxdisp), static_cast<int>(G),
                                               static_cast<int>(C / G), total, static_cast<int>(Dtot), static_cast<int>(dofs0));
    return finish_launch();
}

extern "C" int dv_corr_volume_2sided_f32(const float *ref, const float *tgt, float *out, int64_t B, int64_t C,
                                         int64_t H, int64_t W, int64_t maxdisp, int64_t G, void *stream) {
    return corr2_impl(ref, tgt, out, B, C, H, W, maxdisp, G, 2 * maxdisp + 1, 0, static_cast<cudaStream_t>(stream));
}

// The same volume written into planes [plane_offset, plane_offset + 2*maxdisp + 1) of every sample of a larger
// [B, planes_per_sample, H, W] buffer (one group only: the refinement network's concat buffer, pwcnet_ddim.py:497-499).
extern "C" int dv_corr_volume_2sided_into_f32(const float *ref, const float *tgt, float *buffer, int64_t planes_per_sample,
                                              int64_t plane_offset, int64_t B, int64_t C, int64_t H, int64_t W,
                                              int64_t maxdisp, void *stream) {
    if (planes_per_sample <= 0 || plane_offset < 0 || plane_offset + 2 * maxdisp + 1 > planes_per_sample) return DV_ERR_BAD_SHAPE;
    if (planes_per_sample > INT32_MAX) return DV_ERR_BAD_SHAPE;
    return corr2_impl(ref, tgt, buffer, B, C, H, W, maxdisp, 1, planes_per_sample, plane_offset, static_cast<cudaStream_t>(stream));
}
