"""Stand-in for KITTI15/core/igev_stereo_ddim.py:IGEVStereo_ddim — only what its DiffuVolume sampler methods touch.

The real class cannot be constructed here (core/extractor.py needs timm + a pretrained download, core/update.py needs
opt_einsum), but its sampler METHODS (q_sample, predict_noise_from_start, model_predictions, ddim_sample) only use a few
attributes of `self`.  tests/golden/make_golden.py binds the REFERENCE's unmodified methods onto this class (CPU) to mint
the golden trace; tests/test_gpu_sampler_igev.py binds diffuvolume_b200.sampler's drop-ins onto the same class (CUDA).
The GRU update block / convex upsampler are replaced by cheap deterministic functions of the looked-up correlation
features — they are out of scope and only have to make the trace sensitive to the lookup.
"""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import synth

IGEV_TRACE = dict(B=1, Cf=16, Cg=8, D=48, h=8, w=48, iters=3, times=(999, 499))


class IgevShiftTable(nn.Module):
    """DynamicHead stand-in (KITTI15/core/head.py:74-83): noisy + shift[b, d, 1, 1], shift given per timestep."""

    def __init__(self, table):
        super().__init__()
        self.table = table            # {t: tensor [B, D]}

    def forward(self, noisy, t):
        s = self.table[int(t.reshape(-1)[0].item())]
        return noisy + s[:, :, None, None]


class MockIGEV(nn.Module):
    def __init__(self, schedule, shift_table):
        super().__init__()
        self.args = SimpleNamespace(mixed_precision=False, n_gru_layers=1, slow_fast_gru=False)
        self.scale = 1.0
        self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta = 1000, 2, 1
        self.renewal, self.use_ensemble = True, True
        for name in ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                     "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            self.register_buffer(name, torch.from_numpy(np.asarray(getattr(schedule, name), dtype=np.float64)))
        self.time_embedding = IgevShiftTable(shift_table)

    def update_block(self, net_list, inp_list, corr=None, flow=None, iter08=True, iter16=True, iter32=True, update=True):
        c = corr.float()
        delta = 0.6 * torch.tanh(4.0 * c[:, :72].mean(1, keepdim=True) + 2.0 * c[:, 81:].mean(1, keepdim=True)) - 0.02 * flow
        return net_list, torch.ones(1, device=corr.device), delta

    def upsample_disp(self, disp, mask_feat_4, stem_2x):
        return F.interpolate(disp * 4.0, scale_factor=4, mode="bilinear", align_corners=False)


def igev_trace_inputs(device="cpu"):
    """Seeded inputs of the IGEV sampler trace (numpy -> torch on `device`)."""
    c = IGEV_TRACE
    B, h, w, D = c["B"], c["h"], c["w"], c["D"]
    H, W = 4 * h, 4 * w
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    f1, f2 = synth.normal((B, c["Cf"], h, w), 401), synth.normal((B, c["Cf"], h, w), 402)
    geo = synth.normal((B, c["Cg"], D, h, w), 403)
    init_disp = synth.uniform((B, 1, h, w), 404, dtype=np.float32) * np.float32(40.0)
    # `used` (flow_full in the reference's forward) is compared with the up-sampled flow coords1 - coords0, which starts
    # at zero in this variant (ddim_sample(disp, disp, ...), igev_stereo_ddim.py:424): keep it of the same magnitude so
    # that the renewal mask (< 5) and the fallback (< 3) both fire on part of the image
    used = (synth.normal((B, 1, H, W), 405) * np.float32(3.0)).astype(np.float32)
    gt_q = np.clip(init_disp + synth.normal((B, 1, h, w), 406) * np.float32(1.5), 0, 47).astype(np.float32)
    shifts = {tt: t(synth.normal((B, D), 410 + i) * np.float32(0.1)) for i, tt in enumerate(c["times"])}
    return dict(f1=t(f1), f2=t(f2), geo=t(geo), init_disp=t(init_disp), used=t(used), gt_q=t(gt_q), shifts=shifts)
