"""CPU baseline port of the reference hot path with the reference's own op sequence, on torch
tensors — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.  (Device-agnostic: bench.py times it on the host CPU; the GPU
parity tests also run it on CUDA tensors as a full-size checker, which is what the reference itself does on a GPU box.)

The reference is a PyTorch program; on a host CPU it runs as a chain of ATen kernels (one memset
plus three launches per disparity for a gwc volume, a materialised softmax, broadcast multiplies,
reductions).  `/root/reference` cannot travel to the GPU box, so this file restates that op
sequence (same ATen ops, same temporaries, same number of passes over memory) for
`bench.py`'s `cpu_baseline` and `--impl reference` legs, multi-threaded through ATen's own thread
pool.  It is checked against the numpy oracle and the golden fixtures in tests/test_torch_port.py.
Nothing under diffuvolume_b200/ imports it.

Citations are file:line in the reference.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F


def gwc_volume(ref: torch.Tensor, tgt: torch.Tensor, D: int, G: int) -> torch.Tensor:
    """SceneFlow/models/submodule.py:209-238: zero volume, then per shift slice-mul, view, mean, copy."""
    B, C, H, W = ref.shape
    cpg = C // G
    vol = ref.new_zeros([B, G, D, H, W])
    for d in range(min(D, W)):
        a = ref if d == 0 else ref[..., d:]
        b = tgt if d == 0 else tgt[..., :-d]
        vol[:, :, d, :, d:] = (a * b).view(B, G, cpg, H, W - d).mean(dim=2)
    return vol.contiguous()


def concat_volume(ref: torch.Tensor, tgt: torch.Tensor, D: int, mask_left: bool) -> torch.Tensor:
    """SceneFlow/models/submodule.py:180-191 (mask_left False) / KITTI12/models/submodule.py:86-97 (True)."""
    B, C, H, W = ref.shape
    vol = ref.new_zeros([B, 2 * C, D, H, W])
    for d in range(D):
        if mask_left and d > 0:
            vol[:, :C, d, :, d:] = ref[..., d:]
        else:
            vol[:, :C, d] = ref
        vol[:, C:, d, :, d:] = tgt if d == 0 else tgt[..., :-d]
    return vol.contiguous()


def acv_volume(att_logits: torch.Tensor, concat: torch.Tensor) -> torch.Tensor:
    """SceneFlow/models/acv_ddim.py:390."""
    return F.softmax(att_logits, dim=2) * concat


def filter_volume(volume: torch.Tensor, xt: torch.Tensor, shift: torch.Tensor, scale: float = 1.0):
    """SceneFlow/models/acv_ddim.py:254-260 with DynamicHead's broadcast add (head.py:74-77)."""
    n = xt + shift.view(shift.shape[0], -1, 1, 1)
    n = torch.clamp(n, min=-scale, max=scale)
    n = ((n / scale) + 1) / 2
    return volume * n.unsqueeze(1).float(), n


def softmax_regress(cost: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """acv_ddim.py:269-270 + SceneFlow/models/submodule.py:173-177."""
    D = cost.shape[1]
    prob = F.softmax(cost, dim=1)
    dv = torch.arange(0, D, dtype=prob.dtype, device=prob.device).view(1, D, 1, 1)
    return torch.sum(prob * dv, 1, keepdim=False), prob


def renewal_vote(disp, used, prob, thr_dif=1.0, thr_unc=3.0):
    """acv_ddim.py:320-331."""
    D = prob.shape[1]
    m1 = torch.where(torch.abs(disp - used) < thr_dif, 1, 0)
    dv = torch.arange(0, D, dtype=disp.dtype, device=disp.device).view(1, D, 1, 1)
    unc = torch.sum(torch.abs(disp.unsqueeze(1) - dv) * prob, dim=1)
    m2 = torch.where(unc < thr_unc, 1, 0)
    return (m2 * m1).float()


def xstart_from_pred(pred: torch.Tensor, maxdisp: int = 192, D: int = 48, scale: float = 1.0) -> torch.Tensor:
    """acv_ddim.py:272-292."""
    dn = torch.clamp(pred, 0, maxdisp - 1).unsqueeze(1)
    b, _, H, W = dn.shape
    dn = F.interpolate(dn, size=(H // 4, W // 4), mode="bilinear") / 4
    h, w = dn.shape[-2:]
    real = torch.floor(dn).long()
    coff = real - dn + 1
    vol = torch.zeros([b, D, h * w], dtype=torch.float32, device=pred.device)
    vol.scatter_(1, real.view(b, 1, -1), coff.view(b, 1, -1))
    vol.scatter_(1, torch.clamp(real + 1, 0, D - 1).view(b, 1, -1), (1 - coff).view(b, 1, -1))
    vol = vol.view(b, D, h, w)
    last = torch.zeros([b, D, h, w], dtype=torch.float32, device=pred.device)
    last[:, -1] = 1
    x0 = torch.where((real == D - 1).expand(b, D, h, w), last, vol)
    return torch.clamp(scale * (x0 * 2 - 1.0), min=-scale, max=scale)


def hot_path_pair(feat_l, feat_r, cfeat_l, cfeat_r, att_logits, costs: Sequence[torch.Tensor], used, asd,
                  shifts: Sequence[torch.Tensor], step_noises: Sequence[torch.Tensor],
                  renoises: Sequence[torch.Tensor], sched, cof=(0.5, 0.0, 0.0, 0.0, 0.2, 0.3),
                  D: int = 48, G: int = 40, upsample_to=None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """One "pair" of BASELINE.json's metric with the reference's op sequence: 1x gwc volume, 1x concat + ACV
    multiply, then T x {filter multiply, softmax + regression, uncertainty + vote, x_start, pred_noise,
    DDIM update, re-noise}, and the ensemble.  The conv stack is replaced by nothing:
    `costs[i]` are the synthetic logits standing in for its output at step i (or, with `upsample_to`, the quarter-res
    conv output that the reference upsamples first).  `sched` is an
    oracle.dv_oracle.Schedule (host-side float64 constants)."""
    gwc = gwc_volume(feat_l, feat_r, D, G)
    ac = acv_volume(att_logits, concat_volume(cfeat_l, cfeat_r, D, mask_left=False))
    B, _, _, h, w = ac.shape
    img = asd
    final = [used]
    mask = torch.zeros([B, h, w], dtype=torch.float32, device=ac.device)
    pairs = sched.time_pairs()
    for i, (t, t_next) in enumerate(pairs):
        vol_f, n = filter_volume(ac, img, shifts[i], sched.scale)
        del vol_f  # consumed by the (out-of-scope) 3-D convs
        cost = costs[i]
        if upsample_to is not None:   # acv_ddim.py:267-268: F.upsample(cost_v, [maxdisp, 4h, 4w], mode='trilinear'), squeeze
            cost = torch.squeeze(F.interpolate(cost, list(upsample_to), mode="trilinear"), 1)
        disp, prob = softmax_regress(cost)
        x0 = xstart_from_pred(disp, cost.shape[1], D, sched.scale)
        eps = (float(sched.sqrt_recip_alphas_cumprod[t]) * n.double() - x0) / float(sched.sqrt_recipm1_alphas_cumprod[t])
        final.append(disp)
        vote = renewal_vote(disp, used, prob)
        mask = torch.clamp(mask + F.interpolate(vote.unsqueeze(1), size=(h, w), mode="bilinear").squeeze(1), 0, 1)
        if t_next < 0:
            img = x0
            continue
        san, c, sigma = sched.ddim_coefficients(t, t_next)
        img = x0 * san + c * eps + sigma * step_noises[i]
        img = torch.where(mask.unsqueeze(1) == 0, renoises[i], img)
    stack = torch.stack(final, 0)
    cf = torch.tensor(cof, dtype=torch.float32, device=stack.device).view(-1, 1, 1, 1)
    return torch.sum(stack * cf, dim=0), [gwc, img, mask]


# =====================================================================================================================
# configs[2]: PCWNet + DiffuVolume (KITTI12/models/pwcnet_ddim.py) — the reference's op sequence outside its convolutions
# =====================================================================================================================
def warp(x: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """KITTI12/models/submodule.py:137-176: mesh grid - disp, normalise with (W-1)/(H-1), grid_sample (default
    align_corners=False), validity mask from a warped all-ones tensor (>= 0.999)."""
    B, C, H, W = x.size()
    xx = torch.arange(0, W, device=x.device).view(1, -1).repeat(H, 1)
    yy = torch.arange(0, H, device=x.device).view(-1, 1).repeat(1, W)
    xx = xx.view(1, 1, H, W).repeat(B, 1, 1, 1).float()
    yy = yy.view(1, 1, H, W).repeat(B, 1, 1, 1).float()
    vgrid = torch.cat((xx - disp, yy), 1)
    vgrid[:, 0, :, :] = 2.0 * vgrid[:, 0, :, :].clone() / max(W - 1, 1) - 1.0
    vgrid[:, 1, :, :] = 2.0 * vgrid[:, 1, :, :].clone() / max(H - 1, 1) - 1.0
    vgrid = vgrid.permute(0, 2, 3, 1)
    output = F.grid_sample(x, vgrid)
    mask = F.grid_sample(torch.ones_like(x), vgrid)
    mask[mask < 0.999] = 0
    mask[mask > 0] = 1
    return output * mask


def corr_volume_2sided(ref: torch.Tensor, tgt: torch.Tensor, maxdisp: int, G: int) -> torch.Tensor:
    """KITTI12/models/submodule.py:121-135 (build_corrleation_volume, negative-shift quirk included)."""
    B, C, H, W = ref.shape
    cpg = C // G
    gc = lambda a, b: (a * b).view(B, G, cpg, H, -1).mean(dim=2)
    vol = ref.new_zeros([B, G, 2 * maxdisp + 1, H, W])
    for i in range(-maxdisp, maxdisp + 1):
        if i > 0:
            vol[:, :, i + maxdisp, :, i:] = gc(ref[:, :, :, i:], tgt[:, :, :, :-i])
        elif i < 0:
            vol[:, :, i + maxdisp, :, :-i] = gc(ref[:, :, :, :-i], tgt[:, :, :, i:])
        else:
            vol[:, :, i + maxdisp, :, :] = gc(ref, tgt)
    return vol.contiguous()


def _uncertainty(disp, prob):
    D = prob.shape[1]
    dv = torch.arange(0, D, dtype=disp.dtype, device=disp.device).view(1, D, 1, 1)
    return torch.sum(torch.abs(disp.unsqueeze(1) - dv) * prob, dim=1)


def pcw_hot_path_pair(scales, combine, costs, used, feat_l_full, feat_r_full, start, asd, shifts, step_noises, q_noises, sched,
                      cof=(0.9, 0.0, 0.0, 0.1), maxdisp: int = 192, G: int = 40):
    """One "pair" of configs[2] with the reference's op sequence (pwcnet_ddim.py:604-625 volumes; :530-602 ddim_sample;
    :466-528 model_predictions with dres*/classif3/dispupsample/refinenet3 replaced by nothing: costs[i] stand in for the
    upsampled classif3 output and the regressed disparity for disp_finetune).  Returns (pred, [x_last, mask, prob, corr])."""
    D = maxdisp // 4
    vols = []
    for gl, gr, cl, cr, Ds in scales:
        vols.append(torch.cat((gwc_volume(gl, gr, Ds, G), concat_volume(cl, cr, Ds, mask_left=True)), 1))
    B, _, _, h, w = combine.shape
    img = start
    final = [used]
    mask = torch.zeros([B, h, w], dtype=torch.float32, device=combine.device)
    corr = prob = None
    for i, (t, t_next) in enumerate(sched.time_pairs()):
        vol_f, n = filter_volume(combine, img, shifts[i], sched.scale)
        del vol_f
        disp, prob = softmax_regress(costs[i if len(costs) > 1 else 0])
        warped = warp(feat_r_full, disp.unsqueeze(1))
        corr = torch.squeeze(corr_volume_2sided(feat_l_full, warped, 24, 1), 1)
        x0 = xstart_from_pred(disp, maxdisp, D, sched.scale)
        eps = (float(sched.sqrt_recip_alphas_cumprod[t]) * n.double() - x0) / float(sched.sqrt_recipm1_alphas_cumprod[t])
        final.append(disp)
        dif = torch.abs(disp - used)
        unc = _uncertainty(disp, prob)
        if t_next < 0:
            img = x0
            continue
        vote = (torch.where(dif < 1, 1, 0) * torch.where(unc < 1, 1, 0)).float()
        mask = torch.clamp(mask + F.interpolate(vote.unsqueeze(1), size=(h, w), mode="bilinear").squeeze(1), 0, 1)
        san, c, sigma = sched.ddim_coefficients(t, t_next)
        img = x0 * san + c * eps + sigma * step_noises[i]
        # asd = self.q_sample(asd, t): cumulative (pwcnet_ddim.py:591); float64 from the first application on
        asd = float(sched.sqrt_alphas_cumprod[t]) * asd.double() + float(sched.sqrt_one_minus_alphas_cumprod[t]) * q_noises[i].double()
        img = torch.where(mask.unsqueeze(1) == 0, asd, img)
    stack = torch.stack(final, 0)
    cf = torch.tensor(cof, dtype=torch.float32, device=stack.device).view(-1, 1, 1, 1)
    return torch.sum(stack * cf, dim=0), [img, mask, prob, corr, vols]


# =====================================================================================================================
# configs[3]: IGEV + DiffuVolume (KITTI15/core/igev_stereo_ddim.py, geometry_ddim.py) — op sequence outside convs / GRU
# =====================================================================================================================
class GeoEncodingVolume:
    """KITTI15/core/geometry_ddim.py:6-80 restated: einsum all-pairs correlation, permuted geometry rows, avg_pool2d
    pyramids; __call__ multiplies the whole pyramid by the noise pyramid and samples with grid_sample on EVERY call."""

    def __init__(self, fmap1, fmap2, geo_volume, num_levels=2, radius=4):
        self.num_levels, self.radius = num_levels, radius
        B, Dc, H, W1 = fmap1.shape
        corr = torch.einsum("aijk,aijh->ajkh", fmap1, fmap2).reshape(B, H, W1, 1, -1).contiguous()
        b, c, d, h, w = geo_volume.shape
        geo = geo_volume.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, 1, d)
        corr = corr.reshape(b * h * w, 1, 1, -1)
        self.geo_pyr, self.corr_pyr = [geo], [corr]
        for _ in range(num_levels - 1):
            geo = F.avg_pool2d(geo, [1, 2], stride=[1, 2])
            corr = F.avg_pool2d(corr, [1, 2], stride=[1, 2])
            self.geo_pyr.append(geo)
            self.corr_pyr.append(corr)

    @staticmethod
    def _sample(img, x):
        """core/utils/utils.py:59-77 bilinear_sampler on an [N,C,1,L] image at pixel positions x [N,1,T,1]."""
        L = img.shape[-1]
        grid = torch.cat([2 * x / (L - 1) - 1, torch.zeros_like(x)], dim=-1)
        return F.grid_sample(img, grid, align_corners=True)

    def __call__(self, disp, coords, noisy):
        r = self.radius
        b, _, h, w = disp.shape
        noisy = noisy.reshape(b * h * w, 1, 1, -1)
        noise = [noisy]
        for _ in range(self.num_levels):
            noisy = F.avg_pool2d(noisy, [1, 2], stride=[1, 2])
            noise.append(noisy)
        out = []
        dx = torch.linspace(-r, r, 2 * r + 1).view(1, 1, 2 * r + 1, 1).to(disp.device)
        for i in range(self.num_levels):
            x0 = dx + disp.reshape(b * h * w, 1, 1, 1) / 2 ** i
            out.append(self._sample(self.geo_pyr[i] * noise[i], x0).view(b, h, w, -1))
            x1 = coords.reshape(b * h * w, 1, 1, 1) / 2 ** i - disp.reshape(b * h * w, 1, 1, 1) / 2 ** i + dx
            out.append(self._sample(self.corr_pyr[i], x1).view(b, h, w, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def context_upsample(disp_low, up_weights):
    """KITTI15/core/submodule.py:241-253."""
    b, c, h, w = disp_low.shape
    u = F.unfold(disp_low.reshape(b, c, h, w), 3, 1, 1).reshape(b, -1, h, w)
    u = F.interpolate(u, (h * 4, w * 4), mode="nearest").reshape(b, 9, h * 4, w * 4)
    return (u * up_weights).sum(1)


def igev_hot_path_pair(fmap_l, fmap_r, geo, cost48, up_weights, coords, used, start, asd, shifts, step_noises, q_noises, sched,
                       iters: int = 32, cof=(0.6, 0.1, 0.3), D: int = 48, G: int = 8):
    """One "pair" of configs[3] with the reference's op sequence (igev_stereo_ddim.py:361-427 forward, :294-359 ddim_sample,
    :226-292 model_predictions; the GRU / feature / hourglass networks replaced by nothing: every lookup of a step samples
    at init_disp + 0.125 * iteration and the up-sampled init_disp stands in for the GRU's final disparity)."""
    gwc = gwc_volume(fmap_l, fmap_r, D, G)
    prob = F.softmax(cost48, dim=1)
    dv = torch.arange(0, D, dtype=prob.dtype, device=prob.device).view(1, D, 1, 1)
    disp0 = torch.sum(prob * dv, 1, keepdim=True)
    fn = GeoEncodingVolume(fmap_l.float(), fmap_r.float(), geo.float())
    B, _, h, w = fmap_l.shape
    used_map = used.reshape(B, used.shape[-2], used.shape[-1])
    img = start
    final = [used_map]
    mask = torch.zeros([B, h, w], dtype=torch.float32, device=fmap_l.device)
    look = None
    for i, (t, t_next) in enumerate(sched.time_pairs()):
        n = img + shifts[i].view(B, -1, 1, 1)
        n = torch.clamp(n, min=-sched.scale, max=sched.scale)
        n = ((n / sched.scale) + 1) / 2
        for it in range(iters):
            look = fn(disp0 + 0.125 * it, coords, n.float())
        up = context_upsample(disp0 * 4.0, up_weights)
        # x_start (igev_stereo_ddim.py:268-288): clamp(pred, 0, 47) -> bilinear /4 -> /4 -> + coords0 -> clamp -> 2-tap
        dn = F.interpolate(torch.clamp(up, 0, D - 1).unsqueeze(1), size=(h, w), mode="bilinear") / 4
        tc = torch.clamp(coords.reshape(B, 1, h, w) + dn, 0, D - 1)
        x0 = _xstart_quarter(tc.squeeze(1), D, sched.scale)
        eps = (float(sched.sqrt_recip_alphas_cumprod[t]) * n.double() - x0) / float(sched.sqrt_recipm1_alphas_cumprod[t])
        dif = torch.abs(up - used_map)
        vote = torch.where(dif < 5, 1, 0).float()
        mask = torch.clamp(mask + F.interpolate(vote.unsqueeze(1), size=(h, w), mode="bilinear").squeeze(1), 0, 1)
        final.append(torch.where(dif < 3, up, used_map))
        if t_next < 0:
            img = x0
            continue
        san, c, sigma = sched.ddim_coefficients(t, t_next)
        img = x0 * san + c * eps + sigma * step_noises[i]
        asdd = float(sched.sqrt_alphas_cumprod[t]) * asd.double() + float(sched.sqrt_one_minus_alphas_cumprod[t]) * q_noises[i].double()
        img = torch.where(mask.unsqueeze(1) == 0, asdd, img)
    stack = torch.stack(final, 0)
    cf = torch.tensor(cof, dtype=torch.float32, device=stack.device).view(-1, 1, 1, 1)
    return torch.sum(stack * cf, dim=0), [img, mask, look, gwc]


def _xstart_quarter(dq: torch.Tensor, D: int, scale: float) -> torch.Tensor:
    """The 2-tap scatter of acv_ddim.py:277-292 / igev_stereo_ddim.py:274-288 on an already quarter-res map [B,h,w]."""
    b, h, w = dq.shape
    dn = dq.reshape(b, 1, 1, h, w)
    real = torch.floor(dn).long()
    coff = real - dn + 1
    vol = torch.zeros([b, D, h * w], dtype=torch.float32, device=dq.device)
    vol.scatter_(1, real.view(b, 1, -1), coff.view(b, 1, -1))
    vol.scatter_(1, torch.clamp(real + 1, 0, D - 1).view(b, 1, -1), (1 - coff).view(b, 1, -1))
    vol = vol.view(b, D, h, w)
    last = torch.zeros([b, D, h, w], dtype=torch.float32, device=dq.device)
    last[:, -1] = 1
    x0 = torch.where((real.view(b, 1, h, w) == D - 1).expand(b, D, h, w), last, vol)
    return torch.clamp(scale * (x0 * 2 - 1.0), min=-scale, max=scale)
