"""Stand-in for KITTI12/models/pwcnet_ddim.py:PWCNet_ddim — only what its DiffuVolume sampler touches.

tests/golden/make_golden.py binds the REFERENCE's unmodified `q_sample`, `predict_noise_from_start`, `model_predictions`
and `ddim_sample` (pwcnet_ddim.py:453-602) onto this class on the CPU to mint the golden trace (the reference's own warp /
build_corrleation_volume / disparity_regression run underneath); tests/test_gpu_sampler_pcw.py lets
diffuvolume_b200.install bind the product's tier-2 drop-ins (q_sample, predict_noise_from_start, model_predictions,
ddim_sample) onto the same class.  The 3-D hourglasses, `dispupsample` and `refinenet3` are replaced by cheap deterministic modules — they are out of
scope and only have to make the trace sensitive to the filter, the regression, the warp and the +-24
correlation volume.
"""
import numpy as np
import torch
import torch.nn as nn

import synth

PCW_TRACE = dict(B=1, C=32, D=48, h=8, w=16, Cf=32, times=(999, 665, 332))


class ShiftTable(nn.Module):
    """DynamicHead stand-in (KITTI12/models/head.py:74-77): noisy + shift[b, d, 1, 1], one table row per timestep."""

    def __init__(self, table):
        super().__init__()
        self.table = table

    def forward(self, noisy, t):
        return noisy + self.table[int(t.reshape(-1)[0].item())][:, :, None, None]


class MeanClassif(nn.Module):
    def forward(self, x):                      # [B,32,48,h,w] -> [B,1,48,h,w] logits
        return x.mean(1, keepdim=True) * 24.0


class DispFeatures(nn.Module):
    def forward(self, pred):                   # [B,1,H,W] -> [B,4,H,W]
        return pred.repeat(1, 4, 1, 1) * 0.01


class Refine(nn.Module):
    def forward(self, combine, pred):          # the last 49 channels of `combine` are the +-24 correlation volume
        corr = combine[:, -49:]
        return pred + 0.75 * torch.tanh(8.0 * corr[:, 20:29].mean(1, keepdim=True)) + 0.05 * combine[:, :4].mean(1, keepdim=True)


class MockPCW(nn.Module):
    def __init__(self, schedule, shift_table):
        super().__init__()
        self.scale, self.maxdisp = 1.0, 192
        self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta = 1000, 3, 1
        self.renewal, self.use_ensemble = True, True
        for name in ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                     "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            self.register_buffer(name, torch.from_numpy(np.asarray(getattr(schedule, name), dtype=np.float64)))
        self.time_embedding = ShiftTable(shift_table)
        self.dres2, self.dres3, self.dres4 = nn.Identity(), nn.Identity(), nn.Identity()
        self.classif3 = MeanClassif()
        self.dispupsample = DispFeatures()
        self.refinenet3 = Refine()


def pcw_trace_inputs(device="cpu"):
    c = PCW_TRACE
    B, C, D, h, w, Cf = c["B"], c["C"], c["D"], c["h"], c["w"], c["Cf"]
    H, W = 4 * h, 4 * w
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    # a volume with a clear ridge along a smooth disparity surface, so that softmax over D is peaked
    dq = 6.0 + 30.0 * synth.uniform((B, 1, 1, h, w), 501, dtype=np.float32)
    dd = np.arange(D, dtype=np.float32).reshape(1, 1, D, 1, 1)
    volume = (np.exp(-0.5 * ((dd - dq) / 1.5) ** 2) + 0.05 * synth.normal((B, C, D, h, w), 502)).astype(np.float32)
    used = (4.0 * np.repeat(np.repeat(dq[:, 0, 0], 4, axis=1), 4, axis=2) + synth.normal((B, H, W), 503) * np.float32(0.8)).astype(np.float32)
    gt_q = np.clip(dq[:, 0, 0] + synth.normal((B, h, w), 504) * np.float32(1.0), 0, 47).astype(np.float32)
    fl = {"finetune_feature": t(synth.normal((B, Cf, h, w), 505))}
    fr = {"finetune_feature": t(synth.normal((B, Cf, h, w), 506))}
    shifts = {tt: t(synth.normal((B, D), 510 + i) * np.float32(0.1)) for i, tt in enumerate(c["times"])}
    return dict(volume=t(volume), used=t(used), gt_q=t(gt_q), fl=fl, fr=fr, shifts=shifts)
