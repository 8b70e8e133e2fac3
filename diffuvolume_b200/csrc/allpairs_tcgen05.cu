// allpairs_tcgen05.cu — a14 on the 5th-generation tensor cores: IGEV's all-pairs 1-D correlation
//     out[b,y,x1,x2] = sum_c f1[b,c,y,x1] * f2[b,c,y,x2]            (KITTI15/core/geometry_ddim.py:72-80)
// plus level 1 of the correlation pyramid (avg_pool2d([1,2]), geometry_ddim.py:27-30) from the same accumulators, as a
// persistent, warp-specialised tcgen05 kernel for sm_100a.
//
// Per (b, y) the contraction is [W1 x C] x [C x W2] with C = 96, W = 312: 1.8 GFLOP against 483 MB of compulsory traffic at
// B = 8 — the one op of the path that leans towards compute, and fp32-accurate results are required (1e-4), so it runs as
// 3xTF32: every fp32 operand x is split x = hi + lo (hi = the top 19 bits, lo = x - hi rounded to TF32) and
// hi*hi + lo*hi + hi*lo is accumulated in fp32 in TENSOR MEMORY by tcgen05.mma.kind::tf32.
//
// Layout trick: the features are [B,C,H,W] with W contiguous, i.e. both operands are "MN-major" for the MMA.  A TMA box
// [32 x | 32 channels] with the 128-byte / 32-byte-atom swizzle lands in shared memory exactly in the canonical MN-major
// SWIZZLE_128B_BASE32B operand layout — the only swizzled layout the tensor core accepts for MN-major 32-bit operands —
// (rows of 32 floats = one channel; 4 rows = one K atom of 512 B; boxes 4 KB apart along x), so the raw
// fp32 tile IS the `hi` operand once its low mantissa bits are cleared, and `lo` is an element-wise function of it written
// at the same offsets of a second buffer: the split is a layout-agnostic LDS.128 / STS.128 pass.
//
// Roles (320 threads, one CTA per SM, static round-robin over items = (b*H + y, 128-row tile of x1, 160-column half of x2)):
//   warp 0      TMA producer: per 16-channel chunk 4 + 5 boxes into a ring of R raw slots      (mbarrier full_raw / empty_raw)
//   warps 2-5   splitters: raw -> lo into a ring of L lo slots, fence.proxy.async, arrive      (mbarrier conv_done / empty_lo)
//   warp 1      MMA issuer (one thread): 4 K-steps x 3 tcgen05.mma (M=128, N=160, K=8) per chunk into one of 3 TMEM
//               accumulators (160 columns each, 480 of 512 allocated columns); tcgen05.commit frees the stage / publishes
//               the accumulator                                                                  (mbarrier tmem_full / tmem_empty)
//   warps 6-9   epilogue: tcgen05.ld 32x32b.x32 (thread = one x1 row, 32 x2 columns), pooled pairs, swizzled staging tile,
//               TMA tensor stores (cp.async.bulk.tensor ... bulk_group; rows / columns past W1 / W2 are clipped by the
//               tensor map), double-buffered per warp.
// The epilogue of item i overlaps the loads, splits and MMAs of items i+1 and i+2.
#include "common.cuh"

namespace dv {

namespace ap5 {
constexpr int BM = 128, BN = 160;
constexpr int A_ATOMS = BM / 32, B_ATOMS = BN / 32;           // TMA boxes [32 x | KC channels]
constexpr int EPI_FULL = 32 * 32 * 4, EPI_POOL = 32 * 16 * 4;  // staging tiles per warp and buffer
constexpr int EPI_BUF = EPI_FULL + EPI_POOL;                   // 6 KB
constexpr int THREADS = 320;
// instruction descriptor: D = F32, A = B = TF32, both MN-major, N = 160, M = 128 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((BN >> 3) << 17) | ((BM >> 4) << 24);
// Pipeline shape: KC channels per chunk; R raw slots (TMA landing zone = the `hi` operand) and L lo slots — two rings, so
// the loads run up to R chunks ahead of the tensor core while the split only has to stay L chunks ahead; ACC
// tensor-memory accumulators of BN columns (TMEM columns allocated: the next power of two); EPIB staging buffers per
// epilogue warp; MINB CTAs per SM.
template <int KC_, int R_, int L_, int ACC_, int EPIB_, int MINB_>
struct Cfg {
    static constexpr int KC = KC_, R = R_, L = L_, ACC = ACC_, EPIB = EPIB_, MINB = MINB_;
    static constexpr int ATOM_BYTES = 32 * KC * 4;
    static constexpr int HI_BYTES = (A_ATOMS + B_ATOMS) * ATOM_BYTES;     // A boxes then B boxes
    static constexpr int SMEM_BYTES = (R + L) * HI_BYTES + 4 * EPIB * EPI_BUF + 1024;   // + alignment slack
    static constexpr uint32_t TMEM_COLS = ACC * BN <= 256 ? 256u : 512u;
};
}  // namespace ap5

__device__ __forceinline__ uint64_t ap5_desc(uint32_t smem_addr, uint32_t atom_bytes) {
    // MN-major 32-bit operands have ONE legal swizzled layout: SWIZZLE_128B_BASE32B (layout type 1; 32-byte chunks of a
    // 128-byte row XORed with the row index mod 4 — TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; atom = 32 elements along
    // M/N x 4 K rows = 512 B).  Start address, LBO = the box size (next 32-element block along M/N = next TMA box), SBO = 512 B
    // (next 4-row K atom: rows are contiguous inside a box), descriptor version 1 (Blackwell).
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>(atom_bytes >> 4) << 16) |
           (static_cast<uint64_t>(512u >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void ap5_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(ap5::IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void ap5_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <class CFG>
__global__ void __launch_bounds__(ap5::THREADS, CFG::MINB)
corr1d_allpairs_tcgen05_kernel(const __grid_constant__ CUtensorMap map_f1, const __grid_constant__ CUtensorMap map_f2,
                               const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_pool,
                               int C, int H, int W1, int W2, int mtiles, int nhalves, int nitems, int has_pool) {
    using namespace ap5;
    constexpr int KC = CFG::KC, R = CFG::R, L = CFG::L, ACC = CFG::ACC, EPIB = CFG::EPIB;
    constexpr int ATOM_BYTES = CFG::ATOM_BYTES, HI_BYTES = CFG::HI_BYTES;
    constexpr uint32_t TMEM_COLS = CFG::TMEM_COLS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_raw[R], empty_raw[R], conv_done[L], empty_lo[L], tmem_full[ACC], tmem_empty[ACC];
    __shared__ uint32_t tmem_base_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *lo_base = smem + R * HI_BYTES;
    uint8_t *epi = lo_base + L * HI_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = (C + KC - 1) / KC;

    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) {
            mbar_init(&full_raw[s], 1);
            mbar_init(&empty_raw[s], 1);
        }
        for (int s = 0; s < L; ++s) {
            mbar_init(&conv_done[s], 4);
            mbar_init(&empty_lo[s], 1);
        }
        for (int a = 0; a < ACC; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {   // the MMA warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t n = 0;   // stage uses so far
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int nh = item % nhalves, mt = (item / nhalves) % mtiles, by = item / (nhalves * mtiles);
                const int b = by / H, y = by % H;
                for (int kc = 0; kc < kchunks; ++kc, ++n) {
                    const int s = n % R;
                    mbar_wait(&empty_raw[s], ((n / R) & 1) ^ 1);
                    uint8_t *st = smem + s * HI_BYTES;
                    int boxes = 0;
                    for (int a = 0; a < A_ATOMS; ++a) boxes += (mt * BM + 32 * a < W1) ? 1 : 0;
                    for (int a = 0; a < B_ATOMS; ++a) boxes += (nh * BN + 32 * a < W2) ? 1 : 0;
                    mbar_expect_tx(&full_raw[s], static_cast<uint32_t>(boxes) * ATOM_BYTES);
                    for (int a = 0; a < A_ATOMS; ++a)
                        if (mt * BM + 32 * a < W1)
                            tma_load_4d(st + a * ATOM_BYTES, &map_f1, mt * BM + 32 * a, y, kc * KC, b, &full_raw[s]);
                    for (int a = 0; a < B_ATOMS; ++a)
                        if (nh * BN + 32 * a < W2)
                            tma_load_4d(st + (A_ATOMS + a) * ATOM_BYTES, &map_f2, nh * BN + 32 * a, y, kc * KC, b, &full_raw[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t n = 0, it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                const uint32_t acc = it % ACC;
                mbar_wait(&tmem_empty[acc], ((it / ACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * BN;
                for (int kc = 0; kc < kchunks; ++kc, ++n) {
                    const int s = n % R, l = n % L;
                    mbar_wait(&conv_done[l], (n / L) & 1);           // lo written (which implies the raw tile has landed)
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * HI_BYTES), b_hi = a_hi + A_ATOMS * ATOM_BYTES;
                    const uint32_t a_lo = smem_u32(lo_base + l * HI_BYTES), b_lo = a_lo + A_ATOMS * ATOM_BYTES;
                    const int ksteps = min(KC, C - kc * KC) / 8;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t o = ks * 1024;    // one K atom = 8 channel rows of 128 B
                        ap5_mma(d, ap5_desc(a_lo + o, ATOM_BYTES), ap5_desc(b_hi + o, ATOM_BYTES), (kc | ks) ? 1u : 0u);
                        ap5_mma(d, ap5_desc(a_hi + o, ATOM_BYTES), ap5_desc(b_lo + o, ATOM_BYTES), 1u);
                        ap5_mma(d, ap5_desc(a_hi + o, ATOM_BYTES), ap5_desc(b_hi + o, ATOM_BYTES), 1u);
                    }
                    ap5_commit(&empty_raw[s]);                       // both slots are free once these MMAs have read them
                    ap5_commit(&empty_lo[l]);
                    if (kc == kchunks - 1) ap5_commit(&tmem_full[acc]);   // accumulator complete
                }
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ splitters (128 threads)
        const int t = threadIdx.x - 64;
        uint32_t n = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            for (int kc = 0; kc < kchunks; ++kc, ++n) {
                const int s = n % R, l = n % L;
                mbar_wait(&empty_lo[l], ((n / L) & 1) ^ 1);      // the MMAs that read this lo slot L chunks ago are done
                mbar_wait(&full_raw[s], (n / R) & 1);
                const float4 *hi = reinterpret_cast<const float4 *>(smem + s * HI_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(lo_base + l * HI_BYTES);
                // hi is used as it landed: the tensor core ignores the low 13 mantissa bits of a TF32 operand (truncation),
                // so lo = x - trunc(x) (exact in fp32), rounded to TF32
#pragma unroll 3
                for (int i = t; i < HI_BYTES / 16; i += 128) {
                    const float4 v = hi[i];
                    float4 o;
                    o.x = tf32_rna(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
                    o.y = tf32_rna(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
                    o.z = tf32_rna(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
                    o.w = tf32_rna(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
                    lo[i] = o;
                }
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv_done[l]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 6..9)
        const int q = warp & 3;                                 // TMEM lane quarter this warp may read
        uint8_t *my = epi + (warp - 6) * EPIB * EPI_BUF;
        uint32_t it = 0, nstore = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int nh = item % nhalves, mt = (item / nhalves) % mtiles, by = item / (nhalves * mtiles);
            const uint32_t acc = it % ACC;
            mbar_wait(&tmem_full[acc], (it / ACC) & 1);
            tc_fence_after();
            const int row0 = mt * BM + 32 * q;
            if (row0 < W1) {
                for (int g = 0; g < BN / 32; ++g) {
                    const int col0 = nh * BN + 32 * g;
                    if (col0 >= W2) break;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + acc * BN + 32 * g, v);
                    uint8_t *buf = my + (nstore % EPIB) * EPI_BUF;
                    if (lane == 0) bulk_wait_read<EPIB - 1>();          // the stores that last read this buffer are done with it
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 o = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                                                     __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
                        *reinterpret_cast<float4 *>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) = o;
                    }
                    if (has_pool) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float4 o;
                            o.x = (__uint_as_float(v[8 * c]) + __uint_as_float(v[8 * c + 1])) * 0.5f;
                            o.y = (__uint_as_float(v[8 * c + 2]) + __uint_as_float(v[8 * c + 3])) * 0.5f;
                            o.z = (__uint_as_float(v[8 * c + 4]) + __uint_as_float(v[8 * c + 5])) * 0.5f;
                            o.w = (__uint_as_float(v[8 * c + 6]) + __uint_as_float(v[8 * c + 7])) * 0.5f;
                            *reinterpret_cast<float4 *>(buf + EPI_FULL + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = o;
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&map_out, buf, col0, row0, by);
                        if (has_pool) tma_store_3d(&map_pool, buf + EPI_FULL, col0 / 2, row0, by);
                        bulk_commit();
                    }
                    ++nstore;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (lane == 0) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// Tensor map over an fp32 tensor with explicit byte strides (dims[0] contiguous), given box and swizzle mode.
static bool make_map(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                     const uint32_t *box, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[5], gstride[5];
    cuuint32_t bdim[5], estride[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estride[i] = 1;
        if (i + 1 < rank) {
            if (strides_bytes[i] % 16 != 0) return false;
            gstride[i] = strides_bytes[i];
        }
    }
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<void *>(base), gdim, gstride,
               bdim, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// DV_ERR_UNSUPPORTED when the shape does not fit the kernel's tiling / alignment rules (the caller then takes the
// warp-level MMA kernel).
int launch_allpairs_tcgen05(const float *f1, const float *f2, float *out, float *pooled, int64_t B, int64_t C, int64_t H,
                            int64_t W1, int64_t W2, cudaStream_t st) {
    using namespace ap5;
    if (C % 8 != 0 || W1 % 4 != 0 || W2 % 4 != 0 || (pooled && W2 % 8 != 0)) return DV_ERR_UNSUPPORTED;
    if (!aligned16(f1) || !aligned16(f2) || !aligned16(out) || (pooled && !aligned16(pooled))) return DV_ERR_UNSUPPORTED;
    const int64_t mtiles = (W1 + BM - 1) / BM, nhalves = (W2 + BN - 1) / BN;
    const int64_t nitems = B * H * mtiles * nhalves;
    if (nitems > INT32_MAX || B * H > INT32_MAX) return DV_ERR_UNSUPPORTED;
    CUtensorMap m1, m2, mo, mp;
    const uint64_t d1[4] = {static_cast<uint64_t>(W1), static_cast<uint64_t>(H), static_cast<uint64_t>(C), static_cast<uint64_t>(B)};
    const uint64_t s1[3] = {static_cast<uint64_t>(W1) * 4, static_cast<uint64_t>(H * W1) * 4, static_cast<uint64_t>(C * H * W1) * 4};
    const uint64_t d2[4] = {static_cast<uint64_t>(W2), static_cast<uint64_t>(H), static_cast<uint64_t>(C), static_cast<uint64_t>(B)};
    const uint64_t s2[3] = {static_cast<uint64_t>(W2) * 4, static_cast<uint64_t>(H * W2) * 4, static_cast<uint64_t>(C * H * W2) * 4};
    {
        const uint64_t dout[3] = {static_cast<uint64_t>(W2), static_cast<uint64_t>(W1), static_cast<uint64_t>(B * H)};
        const uint64_t sout[2] = {static_cast<uint64_t>(W2) * 4, static_cast<uint64_t>(W1 * W2) * 4};
        const uint32_t bout[3] = {32u, 32u, 1u};
        if (!make_map(&mo, out, 3, dout, sout, bout, CU_TENSOR_MAP_SWIZZLE_128B)) return DV_ERR_UNSUPPORTED;
        mp = mo;
        if (pooled) {
            const uint64_t dp[3] = {static_cast<uint64_t>(W2 / 2), static_cast<uint64_t>(W1), static_cast<uint64_t>(B * H)};
            const uint64_t sp[2] = {static_cast<uint64_t>(W2 / 2) * 4, static_cast<uint64_t>(W1 * (W2 / 2)) * 4};
            const uint32_t bp[3] = {16u, 32u, 1u};
            if (!make_map(&mp, pooled, 3, dp, sp, bp, CU_TENSOR_MAP_SWIZZLE_64B)) return DV_ERR_UNSUPPORTED;
        }
    }
    const int variant = DV_TUNE("DV_AP5_VARIANT", 0);
#define DV_AP5_LAUNCH(CFG)                                                                                               \
    {                                                                                                                    \
        const uint32_t boxc[4] = {32u, 1u, static_cast<uint32_t>(CFG::KC), 1u};                                          \
        if (!make_map(&m1, f1, 4, d1, s1, boxc, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||                                   \
            !make_map(&m2, f2, 4, d2, s2, boxc, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))                                     \
            return DV_ERR_UNSUPPORTED;                                                                                   \
        auto kern = corr1d_allpairs_tcgen05_kernel<CFG>;                                                                 \
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM_BYTES) != cudaSuccess)     \
            return DV_ERR_LAUNCH;                                                                                        \
        const int64_t slots = static_cast<int64_t>(num_sms()) * CFG::MINB;                                               \
        const int grid = static_cast<int>(nitems < slots ? nitems : slots);                                              \
        kern<<<grid, THREADS, CFG::SMEM_BYTES, st>>>(m1, m2, mo, mp, static_cast<int>(C), static_cast<int>(H),           \
                                                     static_cast<int>(W1), static_cast<int>(W2), static_cast<int>(mtiles), \
                                                     static_cast<int>(nhalves), static_cast<int>(nitems), pooled ? 1 : 0); \
    }
    // Measured at B = 8, C = 96, 96 x 312 (gpurun_out r02: scripts/bench_allpairs.py): <32,3,2,3,1,1> 0.155 ms,
    // <16,8,3,3,1,1> 0.158, <16,6,3,3,2,1> 0.166, <16,3,1,1,1,2> (two CTAs per SM) 0.171, <16,4,2,3,2,1> 0.188;
    // one coupled 2-stage ring (first version) 0.186.
    using CfgA = Cfg<32, 3, 2, 3, 1, 1>;    // 3 raw slots (1.5 items ahead), 2 lo slots, 3 accumulators, one CTA per SM
    using CfgB = Cfg<16, 8, 3, 3, 1, 1>;
    using CfgC = Cfg<16, 6, 3, 3, 2, 1>;
    using CfgE = Cfg<16, 3, 1, 1, 1, 2>;    // two CTAs per SM (256 TMEM columns each)
    switch (variant) {
        case 1: DV_AP5_LAUNCH(CfgB) break;
        case 2: DV_AP5_LAUNCH(CfgC) break;
        case 4: DV_AP5_LAUNCH(CfgE) break;
        default: DV_AP5_LAUNCH(CfgA) break;
    }
#undef DV_AP5_LAUNCH
    return finish_launch();
}

}  // namespace dv
