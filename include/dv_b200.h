/*
 * dv_b200.h — C-ABI of the B200-native DiffuVolume cost-volume hot path.
 *
 * Every entry point is `extern "C" int f(...)`: raw device pointers, int64 sizes, a
 * `cudaStream_t` passed as `void*`; it returns a dv_status (0 = ok) and never throws,
 * allocates, frees or retains memory.  The caller owns every buffer.  All launches go to
 * the stream given (0 = legacy default stream).  The library can be called from one host
 * thread per device concurrently (the reference runs under nn.DataParallel,
 * SceneFlow/main.py:67).  Process-wide state, all of it: a relaxed launch counter
 * (dv_launch_count), the per-device SM count (queried once), tuning environment variables
 * (DV_*, read once per process, unset in production), and — only when the caller passes
 * tile_counters == NULL to the three persistent kernels below — a library-owned pool of
 * 256 device-side tile-counter pairs per kernel family handed out round-robin.
 *
 * `tile_counters` (dv_concat_volume_weighted_f32/_bf16, dv_softmax_regress_f32): the
 * persistent kernels draw their tiles from a {next, done} pair of int32 in device memory.
 * Pass 8 bytes (4-byte aligned) that were zero when first used; the kernel returns them
 * to zero before it exits, so one pair can be reused by consecutive launches on ONE
 * stream.  Launches that may overlap (other streams, other captured graphs) need their
 * own pair.  NULL selects the internal pool: at most 256 overlapping launches per kernel
 * family and device, a captured CUDA graph keeps its pair for life, and a kernel that is
 * aborted mid-run leaves its pair dirty.
 *
 * The reference (iSEE-Laboratory/DiffuVolume) has no FFI of its own: the hot path is a set
 * of module-level Python functions.  Each function below names the
 * reference function (file:line, paths relative to the reference root) it replaces; the
 * Python mirror of those names lives in diffuvolume_b200/{sceneflow,kitti12,kitti15}.py and
 * the binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Layouts: all tensors are contiguous NCHW / NCDHW, fp32 unless a `*_is_f64` flag says
 * otherwise.  "HW" below is H*W.
 */
#ifndef DV_B200_H
#define DV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DV_OK = 0,
    DV_ERR_BAD_SHAPE = 1,   /* a dimension is <= 0, C % G != 0, D out of the supported range ... */
    DV_ERR_BAD_DTYPE = 2,   /* dtype flag not understood */
    DV_ERR_MISALIGNED = 3,  /* a pointer that must be 16-byte aligned is not */
    DV_ERR_LAUNCH = 4,      /* cudaGetLastError() after the launch was not cudaSuccess */
    DV_ERR_NULL = 5,        /* a required pointer is NULL */
    DV_ERR_UNSUPPORTED = 6  /* valid request outside what the kernels implement */
} dv_status;

/* ---- library introspection ------------------------------------------------------------ */
int dv_version(void);                       /* 10000*major + 100*minor + patch */
const char *dv_status_string(int status);   /* static string */
int64_t dv_launch_count(void);              /* kernels launched by this library so far */
int dv_built_for_sm(void);                  /* 100 (sm_100a) */

/* ---- a1: groupwise_correlation  (SceneFlow/models/submodule.py:209-215;
 *          KITTI12/models/submodule.py:100-106; KITTI15/core/submodule.py:151-157)
 * out[b,g,p] = mean_k fea1[b,g*cpg+k,p] * fea2[b,g*cpg+k,p],  p in [0,HW)                  */
int dv_groupwise_correlation_f32(const float *fea1, const float *fea2, float *out,
                                 int64_t B, int64_t C, int64_t H, int64_t W, int64_t G,
                                 void *stream);

/* ---- a2: build_gwc_volume  (SceneFlow/models/submodule.py:228-238;
 *          KITTI12/models/submodule.py:109-119; KITTI15/core/submodule.py:159-169)
 * out[b,g,d,y,x] = mean_k ref[b,g*cpg+k,y,x] * tgt[b,g*cpg+k,y,x-d]  for x >= d, else +0.0
 * out is [B,G,D,H,W]; every element is written (no memset needed).                          */
int dv_gwc_volume_f32(const float *ref, const float *tgt, float *out,
                      int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G,
                      void *stream);

/* ---- a3 (+a4, +a9 fused): build_concat_volume
 *   variant M (mask_left=0): SceneFlow/models/submodule.py:180-191, KITTI15/core/submodule.py:206-217
 *   variant T (mask_left=1): SceneFlow/submodule.py:137-148, KITTI12/models/submodule.py:86-97
 * out[b,c,d,y,x]   = ref[b,c,y,x]      (variant M: all x; variant T: x >= d else 0)
 * out[b,C+c,d,y,x] = tgt[b,c,y,x-d]    (x >= d else 0)
 * Optional fused factors, applied in the reference's rounding order:
 *   att_logits [B,1,D,H,W] != NULL : out = softmax_d(att_logits) * out   (acv_ddim.py:390, acv.py:203)
 *   xt != NULL                     : out = out * float(n),  n = ((clamp(xt + shift[b,d], -s, s)/s)+1)/2
 *                                    (acv_ddim.py:254-260; shift = DynamicHead output, head.py:74-77)
 *     xt is [B,D,H,W] fp32 (xt_is_f64=0) or fp64 (=1); shift is [B,D] fp32 or NULL (= 0).      */
int dv_concat_volume_f32(const float *ref, const float *tgt, float *out,
                         int64_t B, int64_t C, int64_t H, int64_t W, int64_t D,
                         int mask_left,
                         const float *att_logits,
                         const void *xt, int xt_is_f64, const float *shift, double scale,
                         void *stream);

/* ---- a4 / a9 factors as stand-alone fp32 maps, and the volume producer that consumes them.
 * The DDIM loop re-produces the filtered volume T times from the same 1/4-res features; its per-(b,d,pixel)
 * factors are therefore computed once, outside the volume-sized pass:
 *   dv_att_softmax_f32   : weights[b,d,p] = softmax_d(att_logits[b,0,:,p])          (acv_ddim.py:390, acv.py:203)
 *   dv_filter_factor_f32 : n[b,d,p] = float(((clamp(xt + shift[b,d], -s, s)/s)+1)/2)  (acv_ddim.py:256-258)
 *                          (dv_ddim_step can emit the next step's n itself: n_next_out)
 *   dv_concat_volume_weighted_f32 : out = ((concat * att_weights) * n), either factor may be NULL (= 1);
 *                          same values, in the same rounding order, as dv_concat_volume_f32 with att_logits / xt.
 *                          Needs H*W % 4 == 0 and 16-byte aligned pointers (else DV_ERR_MISALIGNED).        */
int dv_att_softmax_f32(const float *att_logits, float *weights, int64_t B, int64_t D, int64_t H, int64_t W,
                       void *stream);
int dv_filter_factor_f32(const void *xt, int xt_is_f64, const float *shift, double scale, float *n_out,
                         int64_t B, int64_t D, int64_t H, int64_t W, void *stream);
/* the same factor in the dtype of xt (n_out_native, what predict_noise_from_start consumes: acv_ddim.py:294,
 * igev_stereo_ddim.py:290) and / or as fp32 (n_out_f32, what multiplies the volume); either may be NULL            */
int dv_filter_factor(const void *xt, int xt_is_f64, const float *shift, double scale, float *n_out_f32,
                     void *n_out_native, int64_t B, int64_t D, int64_t H, int64_t W, void *stream);
int dv_concat_volume_weighted_f32(const float *ref, const float *tgt, float *out,
                                  int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int mask_left,
                                  const float *att_weights, const float *n, void *tile_counters, void *stream);

/* ---- a9 at the op boundary: the DDIM filter multiply on an existing volume
 *          (SceneFlow/models/acv_ddim.py:254-260; KITTI12/models/pwcnet_ddim.py:466-472)
 * out[b,c,d,p] = vol[b,c,d,p] * float(n[b,d,p]);  n as above.  out may alias vol.
 * n_out (optional, same dtype as xt, [B,D,H,W]) receives n itself (the tensor the
 * reference passes on to predict_noise_from_start, acv_ddim.py:294).                        */
int dv_volume_filter_f32(const float *vol, float *out,
                         int64_t B, int64_t C, int64_t D, int64_t H, int64_t W,
                         const void *xt, int xt_is_f64, const float *shift, double scale,
                         void *n_out, void *stream);

/* ---- a5: build_corrleation_volume (sic)  (KITTI12/models/submodule.py:121-135;
 *          SceneFlow/submodule.py:172-186) — two-sided shifts i in [-m, m], slot i+m.
 * i >= 0: out[b,g,i+m,y,x] = mean_k ref[..,x]*tgt[..,x-i]        for x >= i, else 0
 * i <  0 (k=-i): out[b,g,i+m,y,x] = mean_k ref[..,x]*tgt[..,W-k+x] for x <  k, else 0
 *         (the reference's `[..., :-i]` slices select the FIRST k columns — reproduced).   */
int dv_corr_volume_2sided_f32(const float *ref, const float *tgt, float *out,
                              int64_t B, int64_t C, int64_t H, int64_t W, int64_t maxdisp,
                              int64_t G, void *stream);

/* ---- a6 (+a11, +a13 fused): softmax over D followed by disparity_regression
 *          (F.softmax(cost,1) + SceneFlow/models/submodule.py:173-177; call sites
 *          acv_ddim.py:269-270; KITTI12 pwcnet_ddim.py:483-484; KITTI15 igev_stereo_ddim.py:384-385)
 * cost [B,D,H,W] -> disp[b,p] = sum_d d * softmax_d(cost)[b,d,p]
 * Optional outputs (NULL to skip):
 *   prob_out [B,D,H,W]  the softmax itself (PWCNet_ddim returns it, pwcnet_ddim.py:528)
 *   unc_out  [B,H,W]    sum_d |disp - d| * p[d]            (acv_ddim.py:324-329)
 *   vote_out [B,H,W]    1.0f if |disp-used|<thr_dif && unc<thr_unc else 0.0f (acv_ddim.py:322-331);
 *                       needs `used` [B,H,W]
 *   ens_acc  [B,H,W]    ens_acc = (ens_init ? 0 : ens_acc) + ens_coef * disp   (acv_ddim.py:365-369) */
int dv_softmax_regress_f32(const float *cost, int64_t B, int64_t D, int64_t H, int64_t W,
                           float *disp_out, float *prob_out,
                           const float *used, float *unc_out, float *vote_out,
                           float thr_dif, float thr_unc,
                           float *ens_acc, float ens_coef, int ens_init,
                           void *tile_counters, void *stream);

/* ---- f2: F.upsample(..., mode='trilinear') fused into a6 (+a11 +a13)
 *          (SceneFlow/models/acv_ddim.py:267-270 align_corners=False; KITTI12/models/pwcnet_ddim.py:480-484 True)
 * cost_q [B,Dq,h,w] (the squeezed [B,1,Dq,h,w] output of the last 3-D conv) is interpolated to [B,D,H,W] on the fly
 * (ATen upsample_trilinear3d index/lambda rules) and reduced exactly like dv_softmax_regress_f32: the
 * full-resolution logits are never materialised.  Optional outputs as above (no prob_out: that IS the volume).      */
int dv_upsample_softmax_regress_f32(const float *cost_q, int64_t B, int64_t Dq, int64_t h, int64_t w,
                                    int64_t D, int64_t H, int64_t W, int align_corners,
                                    float *disp_out, const float *used, float *unc_out, float *vote_out,
                                    float thr_dif, float thr_unc,
                                    float *ens_acc, float ens_coef, int ens_init, void *stream);

/* ---- a11 alone: uncertainty + renewal vote of a disparity map against a given probability volume
 *          (KITTI12/models/pwcnet_ddim.py:553-570 — the REFINED disparity vs the pre-refinement softmax)
 * unc[b,p] = sum_d |disp[b,p] - d| * prob[b,d,p];  vote = (|disp-used| < thr_dif  [if used]) && unc < thr_unc  */
int dv_uncertainty_vote_f32(const float *prob, const float *disp, const float *used,
                            int64_t B, int64_t D, int64_t H, int64_t W, float thr_dif, float thr_unc,
                            float *unc_out, float *vote_out, void *stream);

/* ---- a11 from the LOGITS: the same uncertainty + vote without a probability volume in HBM
 *          (KITTI12/models/pwcnet_ddim.py:483 F.softmax + :553-570: inside ddim_sample the [B,192,H,W] softmax is only ever
 *          consumed by this reduction, so the sampler keeps the logits and the softmax is recomputed in registers:
 *          the 398 MB/pair write of pred3_volume and its read back are replaced by one read of the logits)
 * unc[b,p] = sum_d |disp[b,p] - d| * softmax_d(cost)[b,d,p];  vote as dv_uncertainty_vote_f32.  tile_counters as above. */
int dv_softmax_uncertainty_vote_f32(const float *cost, const float *disp, const float *used,
                                    int64_t B, int64_t D, int64_t H, int64_t W, float thr_dif, float thr_unc,
                                    float *unc_out, float *vote_out, void *tile_counters, void *stream);

/* ---- a6 alone: disparity_regression on an already-normalised volume
 *          (SceneFlow/models/submodule.py:173-177)  out[b,p] = sum_d d * x[b,d,p]              */
int dv_disparity_regression_f32(const float *x, float *out,
                                int64_t B, int64_t D, int64_t H, int64_t W, void *stream);

/* ---- a7: q_sample  (SceneFlow/models/acv_ddim.py:241-246)
 * out = sqrt_ac * x_start + sqrt_1m_ac * noise, computed and stored in fp64 (buffer
 * promotion); x_start / noise are fp32 or fp64 per flag; n = element count.                  */
int dv_q_sample(const void *x_start, int x_is_f64, const void *noise, int noise_is_f64,
                double sqrt_ac, double sqrt_1m_ac, double *out, int64_t n, void *stream);

/* ---- a8: predict_noise_from_start  (SceneFlow/models/acv_ddim.py:248-252)
 * out = (sqrt_recip * x_t - x0) / sqrt_recipm1   in fp64                                     */
int dv_predict_noise_from_start(const void *x_t, int xt_is_f64, const void *x0, int x0_is_f64,
                                double sqrt_recip, double sqrt_recipm1, double *out,
                                int64_t n, void *stream);

/* ---- a10: disparity map -> 2-tap x_start volume
 *          (SceneFlow/models/acv_ddim.py:272-292, :403-419; KITTI15 igev_stereo_ddim.py:268-288)
 * disp_q [B,h,w] is the quarter-resolution disparity already divided by 4 (and, for IGEV,
 * offset and clamped); r = floor(disp_q); vol[r] = r+1-disp_q; vol[min(r+1,D-1)] = disp_q-r
 * (written second, so it wins when r == D-1 ... then r == D-1 is forced to one-hot(D-1));
 * x0 = clamp(scale*(2*vol-1), -scale, scale).  out [B,D,h,w] fp32.                           */
int dv_xstart_from_disp_f32(const float *disp_q, float *out,
                            int64_t B, int64_t D, int64_t h, int64_t w, double scale,
                            void *stream);

/* ---- bilinear /4 down-sampling of a full-resolution map (F.interpolate(..., size=(H//4, W//4),
 *          mode='bilinear'), align_corners=False; acv_ddim.py:274, :333)
 * out[b,y,x] = post_scale * bilinear(clamp(in, lo, hi))   (clamp skipped when lo > hi)        */
int dv_downsample_bilinear_f32(const float *in, float *out,
                               int64_t B, int64_t H, int64_t W, int64_t h, int64_t w,
                               float lo, float hi, float post_scale, void *stream);

/* ---- a8+a10+a11+a12 fused: one DDIM sampler step on the [B,D,h,w] state
 *          (SceneFlow/models/acv_ddim.py:272-294, :320-362; KITTI12 pwcnet_ddim.py:504-526, :551-593;
 *           KITTI15 igev_stereo_ddim.py:268-290, :315-346)
 * See dv_ddim_step_args below; every tensor is caller-owned.                                  */
typedef struct {
    int64_t B, D, h, w;        /* state shape [B,D,h,w] (D = 48 in the reference)               */
    int64_t H, W;              /* full-resolution shape of `disp` / `vote` ([B,H,W])            */
    /* inputs */
    const float *disp;         /* [B,H,W] regressed disparity of this step                       */
    float disp_clamp_hi;       /* clamp(disp, 0, disp_clamp_hi) before down-sampling (191 / 47)  */
    const float *coords0;      /* IGEV only: [B,h,w] init disparity added after /4, then clamp
                                  to [0, D-1] (igev_stereo_ddim.py:271-273); NULL otherwise      */
    const void *xt;            /* [B,D,h,w] current noisy state (pre time-embedding)             */
    int xt_is_f64;
    const float *shift;        /* [B,D] time-embedding shift or NULL                              */
    double scale;              /* self.scale (1.0)                                               */
    const float *vote;         /* [B,H,W] 0/1 renewal votes at full res (from dv_softmax_regress_f32), or NULL */
    const float *used;         /* when vote == NULL and used != NULL: vote = |disp - used| < vote_thr_dif
                                  computed inline (IGEV: igev_stereo_ddim.py:316-317); both NULL: mask unchanged */
    float vote_thr_dif;
    float *mask;               /* [B,h,w] renewal mask, updated in place: clamp(mask+down4(vote),0,1) */
    /* schedule scalars (host-computed in fp64 from the cosine schedule) */
    double sqrt_recip, sqrt_recipm1;   /* at t                                                   */
    int last_step;             /* time_next < 0: x_next = x0, nothing else                        */
    double sqrt_alpha_next, c, sigma;  /* DDIM update coefficients                                */
    const void *step_noise;    /* [B,D,h,w] randn_like(img) of this step (dtype = xt's)           */
    /* re-noising of un-renewed pixels: x_next = where(mask==0, renoise, x_next)                  */
    int renoise_mode;          /* 0 none; 1 `renoise` given directly (ACV: uniform rand fp64);
                                  2 renoise = sqrt_ac*asd + sqrt_1m_ac*q_noise (PCW / IGEV)        */
    const void *renoise;       /* mode 1: [B,D,h,w] fp64                                         */
    const void *asd;           /* mode 2: [B,D,h,w] x_start of `used` (fp32 or fp64)              */
    int asd_is_f64;
    const void *q_noise;       /* mode 2: [B,D,h,w] randn_like(asd), fp32 or fp64 per flag          */
    int q_noise_is_f64;
    double sqrt_ac, sqrt_1m_ac;
    double *asd_out;           /* mode 2, optional: q_sample result (PCW keeps it cumulatively)    */
    /* outputs */
    float *x0_out;             /* [B,D,h,w] fp32 x_start                                          */
    double *eps_out;           /* [B,D,h,w] fp64 pred_noise, optional                             */
    void *x_next;              /* [B,D,h,w]: fp64 normally; fp32 x0 copy when last_step           */
    /* optional: the NEXT step's filter factor, n_next = float(filter_n(x_next + shift_next[b,d])) */
    const float *shift_next;   /* [B,D] time-embedding shift of the next timestep, or NULL (= 0)   */
    float *n_next_out;         /* [B,D,h,w] fp32, or NULL                                          */
} dv_ddim_step_args;

int dv_ddim_step(const dv_ddim_step_args *args, void *stream);

/* ---- a14: all-pairs 1-D correlation  (KITTI15/core/geometry_ddim.py:72-80)
 * out[b,y,x1,x2] = sum_c fmap1[b,c,y,x1] * fmap2[b,c,y,x2]   ([B,H,W1,W2], no 1/sqrt(C))     */
int dv_corr1d_allpairs_f32(const float *fmap1, const float *fmap2, float *out,
                           int64_t B, int64_t C, int64_t H, int64_t W1, int64_t W2, void *stream);
/* the same, also writing level 1 of the correlation pyramid (geometry_ddim.py:27-30: avg_pool2d([1,2]) of `out`,
 * pooled [B,H,W1,W2/2]) from the accumulators — bit-identical to dv_avgpool_w2_f32 applied to `out`                  */
int dv_corr1d_allpairs_pooled_f32(const float *fmap1, const float *fmap2, float *out, float *pooled,
                                  int64_t B, int64_t C, int64_t H, int64_t W1, int64_t W2, void *stream);

/* ---- a14: pyramid helpers (geometry_ddim.py:18-30)
 * geo [B,C,D,h,w] -> geo_rows [B*h*w, C, D] (permute(0,3,4,1,2));
 * rows [N, L] -> pooled [N, L/2] (avg_pool2d([1,2], stride [1,2]), floor)                     */
int dv_geo_permute_f32(const float *geo, float *rows, int64_t B, int64_t C, int64_t D,
                       int64_t h, int64_t w, void *stream);
int dv_avgpool_w2_f32(const float *rows, float *pooled, int64_t N, int64_t L, void *stream);

/* ---- a15: Combined_Geo_Encoding_Volume.__call__  (KITTI15/core/geometry_ddim.py:33-69,
 *          geometry.py:34-58; bilinear_sampler core/utils/utils.py:59-77)
 * For level i in [0,num_levels): taps dx = -r..r
 *   geo part : bilinear sample (zero padding, align_corners=True) of geo_pyr[i][n, c, :] * noise_i[n, :]
 *              at x = disp[n]/2^i + dx            -> channel  base_i + c*(2r+1) + tap
 *   corr part: bilinear sample of corr_pyr[i][n, :] at x = (coords[n] - disp[n])/2^i + dx
 *                                                 -> channel  base_i + C*(2r+1) + tap
 * out is [B, num_levels*(C+1)*(2r+1), h, w].  noisy may be NULL (geometry.py: no multiply).
 * noisy is the raw [B*h*w, D] reinterpretation of the caller's [B,D,h,w] tensor (the
 * reference reshapes without a permute, geometry_ddim.py:37 — reproduced bit-compatibly);
 * level-i noise is noisy avg-pooled i times.                                                  */
int dv_geo_lookup_f32(const float *const *geo_pyr, const float *const *corr_pyr,
                      const float *noisy, const float *disp, const float *coords, float *out,
                      int64_t B, int64_t C, int64_t D, int64_t h, int64_t w, int64_t W2,
                      int num_levels, int radius, void *stream);

/* ---- a14 + a15, packed pyramid (same reference lines as above).
 * The reference keeps geo as [N, C, D_l] rows (geometry_ddim.py:19); a lookup then touches C separate 40-byte windows
 * per pixel and level.  dv_geo_pack_f32 builds a hypothesis-major pyramid rows_pyr[l] = [N, D >> l, C] for every level
 * l < num_levels in ONE pass over geo [B,C,D,h,w] (permute(0,3,4,1,2) and the avg_pool2d([1,2]) chain fused, same
 * pairwise (a+b)/2 rounding), so the (2r+2)*C floats one lookup needs are one contiguous run.
 * dv_geo_lookup_packed_f32 is dv_geo_lookup_f32 on that layout: identical arithmetic, identical output.             */
int dv_geo_pack_f32(const float *geo, float *const *rows_pyr, int64_t B, int64_t C, int64_t D,
                    int64_t h, int64_t w, int num_levels, void *stream);
int dv_geo_lookup_packed_f32(const float *const *geo_pyr, const float *const *corr_pyr,
                             const float *noisy, const float *disp, const float *coords, float *out,
                             int64_t B, int64_t C, int64_t D, int64_t h, int64_t w, int64_t W2,
                             int num_levels, int radius, void *stream);

/* ---- a9 for IGEV: the DDIM filter on the packed pyramid (KITTI15/core/geometry_ddim.py:37-43,56).
 * rows_out[l][n, j, c] = rows_in[l][n, j, c] * noise_l[n, j]; noise_l = `noisy` (raw [N, D] rows, as in
 * dv_geo_lookup_f32) avg-pooled l times.  The reference redoes this product inside every lookup (64 per pair); the 32
 * GRU iterations of one DDIM step share one noise tensor (igev_stereo_ddim.py:226-240), so the product is taken once per
 * step and the lookups then run on rows_out with noisy = NULL — bit-identical (same single fp32 product per element). */
int dv_geo_filter_packed_f32(const float *const *rows_in, const float *noisy, float *const *rows_out,
                             int64_t N, int64_t C, int64_t D, int num_levels, void *stream);

/* ---- f4: context_upsample  (KITTI15/core/submodule.py:241-253; call sites
 *          igev_stereo_ddim.py:209,462, igev_stereo.py:146,220)
 * out[b,Y,X] = sum_{k=ky*3+kx} disp_low[b, Y/4+ky-1, X/4+kx-1] * up_weights[b,k,Y,X]   (zero padding, taps in order)
 * disp_low [B,1,h,w], up_weights [B,9,4h,4w], out [B,4h,4w]; up_weights/out 16-byte aligned.
 * Backward: grad_low [B,1,h,w] and/or grad_weights [B,9,4h,4w] (either may be NULL).                               */
int dv_context_upsample_f32(const float *disp_low, const float *up_weights, float *out,
                            int64_t B, int64_t h, int64_t w, void *stream);
int dv_context_upsample_bwd_f32(const float *grad_out, const float *disp_low, const float *up_weights,
                                float *grad_low, float *grad_weights, int64_t B, int64_t h, int64_t w, void *stream);

/* ---- f4: ACVNet's depth-wise "patch" convolutions on the gwc volume
 *          (SceneFlow/models/acv_ddim.py:181-188 module definitions, :377-381 call chain; same in acv.py)
 * For channels c in [c0, c1) of in [B,C,D,H,W], per (b,c,d) plane:
 *   mid = conv3x3(in,  w1[c], dilation dil1, zero padding)        w1, w2: [C,9] (= Conv3d weight [C,1,1,3,3] flattened)
 *   out = conv3x3(mid, w2[c], dilation dil2, zero padding)        w2 == NULL: out = mid (a single depth-wise conv)
 * `out` is [B,C,D,H,W] too (the reference's torch.cat of the three dilation classes is written in place); channels
 * outside [c0, c1) are not touched.  in != out.                                                                     */
int dv_depthwise3x3_chain_f32(const float *in, const float *w1, const float *w2, float *out,
                              int64_t B, int64_t C, int64_t D, int64_t H, int64_t W, int64_t c0, int64_t c1,
                              int dil1, int dil2, void *stream);

/* ---- bf16 volumes (north_star: "volume writes ... bf16/fp32"): the two volume PRODUCERS with a bfloat16 output.
 * Features, factor maps, products and accumulation stay fp32; the result is rounded to nearest-even once, at the store
 * (identical to the fp32 entry point followed by torch's .to(torch.bfloat16)), so the volume-sized write — the
 * dominant traffic of a2 / a3+a4 / a9 — is halved.  Stated tolerance: 2^-8 relative per element.  `out` is
 * [B,G,D,H,W] / [B,2C,D,H,W] bfloat16, 8-byte aligned; H*W % 4 == 0; gwc: 8 or 12 channels per group (every reference
 * configuration); else DV_ERR_UNSUPPORTED / DV_ERR_MISALIGNED.                                                          */
int dv_gwc_volume_bf16(const float *ref, const float *tgt, void *out,
                       int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G, void *stream);
int dv_concat_volume_weighted_bf16(const float *ref, const float *tgt, void *out,
                                   int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int mask_left,
                                   const float *att_weights, const float *n, void *tile_counters, void *stream);

/* ---- f1: backward passes of the volume ops — the reference's training scripts differentiate through them
 *          (SceneFlow/main.py:154 -> models/acv_ddim.py:424-482; KITTI12/main.py; KITTI15/train_stereo.py).
 * Gradients of the functions above with respect to their feature inputs; grad_out has the forward output's shape.
 * Either gradient pointer may be NULL (not needed); every other convention as in the forward entry points.
 *   gwc    : grad_ref[c,y,x] = 1/cpg sum_{d<=x}   grad_out[g,d,y,x]   * tgt[c,y,x-d]
 *            grad_tgt[c,y,x] = 1/cpg sum_{x+d<W}  grad_out[g,d,y,x+d] * ref[c,y,x+d]
 *   corr2  : the same over slots [m, 2m] plus the first-k-columns terms of slots [0, m) (KITTI12/models/submodule.py:128-131)
 *   concat : grad_ref[c,y,x] = sum_d grad_out[c,d,y,x] (x >= d when mask_left); grad_tgt[c,y,x] = sum_{x+d<W} grad_out[C+c,d,y,x+d]
 *   regression : grad_x[b,d,y,x] = d * grad_out[b,y,x]                                                          */
int dv_gwc_volume_bwd_f32(const float *grad_out, const float *ref, const float *tgt, float *grad_ref, float *grad_tgt,
                          int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G, void *stream);
int dv_corr_volume_2sided_bwd_f32(const float *grad_out, const float *ref, const float *tgt, float *grad_ref,
                                  float *grad_tgt, int64_t B, int64_t C, int64_t H, int64_t W, int64_t maxdisp,
                                  int64_t G, void *stream);
int dv_groupwise_correlation_bwd_f32(const float *grad_out, const float *fea1, const float *fea2, float *grad1,
                                     float *grad2, int64_t B, int64_t C, int64_t H, int64_t W, int64_t G, void *stream);
int dv_concat_volume_bwd_f32(const float *grad_out, float *grad_ref, float *grad_tgt,
                             int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int mask_left, void *stream);
int dv_disparity_regression_bwd_f32(const float *grad_out, float *grad_x, int64_t B, int64_t D, int64_t H, int64_t W,
                                    void *stream);
/* Backward of the FUSED forward ops — what the training branch of ACVNet_DDIM.forward / ACVNet.forward differentiates
 * between its convolutions (SceneFlow/models/acv_ddim.py:388-390,446-480; acv.py:203,213-236; SceneFlow/main.py:154).
 *   dv_softmax_regress_bwd_f32 : d_cost[b,d,p] = softmax_d(cost)[b,d,p] * (d - disp[b,p]) * grad_disp[b,p]  for
 *        disp = disparity_regression(F.softmax(cost, 1)); the softmax is recomputed, never stored.
 *   dv_acv_volume_bwd_f32 : out = (concat(cl, cr) * att_weights) * n (dv_concat_volume_weighted_f32; either factor may
 *        be NULL = 1).  grad_cl / grad_cr [B,C,H,W] and grad_att_logits [B,D,H,W] (gradient of the LOGITS whose softmax
 *        over D is att_weights; needs cl, cr, att_weights); any output may be NULL.  n carries no gradient in the
 *        reference (torch.tensor(noisy), acv_ddim.py:449).
 * The gradient of the filter multiply vol * n with respect to vol is dv_volume_filter_f32 applied to grad_out.        */
int dv_softmax_regress_bwd_f32(const float *cost, const float *grad_disp, float *grad_cost,
                               int64_t B, int64_t D, int64_t H, int64_t W, void *stream);
int dv_acv_volume_bwd_f32(const float *grad_out, const float *cl, const float *cr, const float *att_weights,
                          const float *n, float *grad_cl, float *grad_cr, float *grad_att_logits,
                          int64_t B, int64_t C, int64_t H, int64_t W, int64_t D, int mask_left, void *stream);

/* ---- f3: warp  (KITTI12/models/submodule.py:137-176; SceneFlow/submodule.py:188-227)
 * out[b,c,y,x] = mask * bilinear(x_in[b,c], ix, iy), zero padding, ix = (x - disp[b,0,y,x]) * W/(W-1) - 0.5,
 * iy = y * H/(H-1) - 0.5 (the reference's grid normalisation + grid_sample's default align_corners=False);
 * mask = 0 where the in-bounds tap weights sum to < 0.999, else 1.                                              */
int dv_warp_f32(const float *x, const float *disp, float *out, int64_t B, int64_t C, int64_t H, int64_t W, void *stream);

/* ---- f1: backward of warp — the autograd of KITTI12/models/submodule.py:169-176 (grid_sample's bilinear
 * backward with zero padding, the reference's grid normalisation, the piecewise-constant validity mask):
 *   grad_x[b,c,tap]   += w_tap * mask * grad_out[b,c,y,x]        (scatter-add, like ATen's grid_sampler_2d_backward)
 *   grad_disp[b,0,y,x] = -(W/2) * 2/(W-1) * sum_c mask * grad_out[b,c,y,x] * d(bilinear)/d(ix)
 * Either output may be NULL (not both); both are fully written (the call clears them first on `stream`).            */
int dv_warp_bwd_f32(const float *grad_out, const float *x, const float *disp, float *grad_x, float *grad_disp,
                    int64_t B, int64_t C, int64_t H, int64_t W, void *stream);

/* ---- f3 + refinement-input assembly  (KITTI12/models/pwcnet_ddim.py:493-499: right_warp = warp(right, pred3);
 *          combine = torch.cat((left - right_warp, left, ..., cost), 1))
 * dv_warp_f32 with two extra outputs written in the same pass, addressed as base + b * batch_stride (floats) + the
 * contiguous [C,H,W] block of the sample — i.e. channel slices of ONE concat buffer, so neither the subtraction nor
 * torch.cat run as separate volume-sized passes:
 *   warp_out[b,c,y,x] = warp(x, disp)            (contiguous [B,C,H,W]; the correlation volume reads it)
 *   diff_out[b,c,y,x] = ref - warp(x, disp)      (optional)
 *   copy_out[b,c,y,x] = ref                      (optional)                                                          */
/* dv_corr_volume_2sided_f32 (G = 1) written into planes [plane_offset, plane_offset + 2*maxdisp + 1) of every sample of
 * a larger contiguous [B, planes_per_sample, H, W] buffer.                                                            */
int dv_corr_volume_2sided_into_f32(const float *ref, const float *tgt, float *buffer, int64_t planes_per_sample,
                                   int64_t plane_offset, int64_t B, int64_t C, int64_t H, int64_t W, int64_t maxdisp,
                                   void *stream);
int dv_warp_assemble_f32(const float *x, const float *disp, const float *ref, float *warp_out,
                         float *diff_out, int64_t diff_batch_stride, float *copy_out, int64_t copy_batch_stride,
                         int64_t B, int64_t C, int64_t H, int64_t W, void *stream);

/* ---- a13: ensemble (acv_ddim.py:365-369): out[p] = sum_i cof[i] * maps[i][p], i < n_maps <= 8
 * `maps` is a HOST array of n_maps device pointers, `cof` a HOST array of n_maps floats.          */
int dv_ensemble_f32(const float *const *maps, const float *cof, int n_maps, float *out,
                    int64_t n, void *stream);

/* ---- IGEV fallback to the initial disparity (igev_stereo_ddim.py:323-325):
 * out[p] = |a[p] - b[p]| < thr ? a[p] : b[p]   (torch.where(torch.abs(disp - used) < 3, disp, used))              */
int dv_select_close_f32(const float *a, const float *b, float thr, float *out, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DV_B200_H */
