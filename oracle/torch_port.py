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
