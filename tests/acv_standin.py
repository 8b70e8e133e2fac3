"""Light stand-ins for the OUT-OF-SCOPE sub-networks of SceneFlow/models/acv_ddim.py:ACVNet_DDIM and acv.py:ACVNet, so that
the reference's `forward` can be replayed on the GPU box, where the reference tree does not exist.

  * tests/golden/make_golden.py (authoring container) builds the REAL reference class, swaps these modules in for its 2-D
    feature extractor, 3-D hourglasses, classifiers and DynamicHead (`graft`), runs the reference's UNMODIFIED `forward`
    (eval and training branch, with its own ddim_sample / model_predictions / q_sample underneath) on the CPU and stores
    the results in tests/golden/tier3.npz;
  * tests/test_gpu_tier3.py builds `AcvStandIn` — the same sub-modules with the same seeded weights, the reference's
    attribute names, NO methods of its own — lets `diffuvolume_b200.install` bind `forward` and the sampler methods onto
    it exactly as it does onto the reference class, and compares.

The patch convolutions keep the reference's exact definition (acv_ddim.py:181-188): they are on the hot path (row f4).
Weights come from synth seeds, not from torch's RNG, so both sides hold bit-identical parameters.
"""
import numpy as np
import torch
import torch.nn as nn

import synth

T3 = dict(B=2, H=32, W=64, maxdisp=192, t_train=437)


def _fill(module, seed, gain=1.0):
    for i, p in enumerate(module.parameters()):
        fan = max(1, int(np.prod(p.shape[1:]))) if p.dim() > 1 else 1
        a = synth.normal(tuple(p.shape), seed + i) * np.float32(gain / np.sqrt(fan))
        with torch.no_grad():
            p.copy_(torch.from_numpy(a))
    return module


class Features(nn.Module):
    """feature_extraction stand-in: {"gwc_feature": [B,320,H/4,W/4]} (submodule.py feature_extraction.forward)."""

    def __init__(self, seed):
        super().__init__()
        self.conv = _fill(nn.Conv2d(3, 320, 4, stride=4, bias=False), seed)

    def forward(self, x):
        return {"gwc_feature": torch.tanh(self.conv(x))}


class Scaled(nn.Module):
    def __init__(self, conv, k):
        super().__init__()
        self.conv, self.k = conv, k

    def forward(self, x):
        return self.conv(x) * self.k


class Residual(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv

    def forward(self, x):
        return x + 0.1 * torch.tanh(self.conv(x))


def _dstar(B, h, w):
    """Per-pixel quarter-res disparity the peaked classifier points at (smooth in [6, 36])."""
    return 6.0 + 30.0 * synth.uniform((B, 1, 1, h, w), 9510, dtype=np.float32)


class Peaked(nn.Module):
    """classif2 stand-in: conv(x) * k plus a strong per-pixel peak over D, so that softmax over D is sharp (low
    uncertainty -> the renewal votes fire where `used` agrees) while the result still depends on the filtered volume."""

    def __init__(self, conv, k):
        super().__init__()
        self.conv, self.k, self._bias = conv, k, {}

    def forward(self, x):
        y = self.conv(x) * self.k
        B, _, D, h, w = y.shape
        key = (B, D, h, w, str(y.device))
        if key not in self._bias:
            dv = np.arange(D, dtype=np.float32).reshape(1, 1, D, 1, 1)
            peak = (-1.0 * (dv - _dstar(B, h, w)) ** 2).astype(np.float32)
            self._bias[key] = torch.from_numpy(peak).to(y.device)
        return y + self._bias[key]


class TimeShift(nn.Module):
    """DynamicHead stand-in (head.py:74-77): noisy + shift(t)[:, :, None, None] with a trainable table."""

    def __init__(self, seed):
        super().__init__()
        self.table = nn.Parameter(torch.from_numpy(synth.normal((1000, 48), seed) * np.float32(0.1)))

    def forward(self, noisy, t):
        return noisy + self.table[t.reshape(-1)[:1]].reshape(1, 48, 1, 1)


def standin_modules(seed=9000):
    """name -> module, in a fixed order (the seeds follow the order)."""
    c3 = lambda i, o, s: _fill(nn.Conv3d(i, o, 3, padding=1, bias=False), s)
    mods = {
        "feature_extraction": Features(seed),
        "concatconv": _fill(nn.Conv2d(320, 32, 1, bias=False), seed + 10),
        "patch": _fill(nn.Conv3d(40, 40, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=40, padding=(0, 1, 1), bias=False), seed + 20),
        "patch_l1": _fill(nn.Conv3d(8, 8, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=8, padding=(0, 1, 1), bias=False), seed + 30),
        "patch_l2": _fill(nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=2, groups=16, padding=(0, 2, 2), bias=False), seed + 40),
        "patch_l3": _fill(nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=3, groups=16, padding=(0, 3, 3), bias=False), seed + 50),
        "dres1_att_": c3(40, 8, seed + 60),
        "dres2_att_": Residual(c3(8, 8, seed + 70)),
        "classif_att_": Scaled(c3(8, 1, seed + 80), 6.0),
        "time_embedding": TimeShift(seed + 90),
        "dres0": nn.Sequential(c3(64, 8, seed + 100), nn.ReLU(inplace=True)),
        "dres1": c3(8, 8, seed + 110),
        "dres2": Residual(c3(8, 8, seed + 120)),
        "dres3": Residual(c3(8, 8, seed + 130)),
        "classif0": Scaled(c3(8, 1, seed + 140), 40.0),
        "classif1": Scaled(c3(8, 1, seed + 150), 40.0),
        "classif2": Peaked(c3(8, 1, seed + 160), 600.0),
    }
    return mods


def graft(net, seed=9000):
    """Swap the stand-ins into a real reference model instance (authoring container)."""
    for name, m in standin_modules(seed).items():
        if hasattr(net, name):
            setattr(net, name, m)
    return net


class AcvStandIn(nn.Module):
    """The attribute surface of ACVNet_DDIM / ACVNet that the hot path touches (acv_ddim.py:120-238) — and nothing else:
    every method (`forward`, `ddim_sample`, `model_predictions`, `q_sample`, `predict_noise_from_start`) is bound by
    diffuvolume_b200.install, as on the reference class."""

    def __init__(self, maxdisp=192, attn_weights_only=False, freeze_attn_weights=False, seed=9000, schedule=None):
        super().__init__()
        self.maxdisp, self.attn_weights_only, self.freeze_attn_weights = maxdisp, attn_weights_only, freeze_attn_weights
        self.num_groups, self.concat_channels = 40, 32
        self.scale = 1.0
        self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta = 1000, 5, 1
        self.renewal, self.use_ensemble = True, True
        if schedule is not None:
            for name in ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                         "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
                self.register_buffer(name, torch.from_numpy(np.asarray(getattr(schedule, name), dtype=np.float64)))
        for name, m in standin_modules(seed).items():
            setattr(self, name, m)


def t3_inputs():
    B, H, W = T3["B"], T3["H"], T3["W"]
    left, right = synth.normal((B, 3, H, W), 9501), synth.normal((B, 3, H, W), 9502)
    # `used` agrees with the peaked classifier's disparity on the left half of the image only: the renewal mask is mixed
    up = np.repeat(np.repeat(_dstar(B, H // 4, W // 4)[:, 0, 0], 4, axis=1), 4, axis=2) * np.float32(4) + np.float32(1.5)
    used = up + (synth.uniform((B, H, W), 9503, dtype=np.float32) - np.float32(0.5)) * np.float32(0.8)
    used[:, :, W // 2:] += np.float32(25)
    used = used.astype(np.float32)
    disp_q = (synth.uniform((B, 1, H // 4, W // 4), 9504, dtype=np.float32) * np.float32(47.75)).astype(np.float32)
    mask_gt = (synth.uniform((B, 1, H // 4, W // 4), 9505, dtype=np.float32) > np.float32(0.25)).astype(np.float32)
    return left, right, used, disp_q, mask_gt


class SeededNoise:
    """torch.randn / randn_like / rand_like / randint replaced by seeded synthetic draws, on both sides:
    `with SeededNoise(device) as rng: ...`; rng.log lists (kind, shape, dtype) of every draw."""

    def __init__(self, device="cpu", base=9600):
        self.device, self.base, self.k, self.log = device, base, 0, []

    def _next(self, kind, shape, dtype):
        seed = self.base + self.k
        self.k += 1
        self.log.append((kind, tuple(shape), str(dtype).replace("torch.", "")))
        a = synth.uniform(tuple(shape), seed, dtype=np.float64) if kind == "rand" else synth.normal(tuple(shape), seed, dtype=np.float64)
        return torch.from_numpy(a).to(dtype).to(self.device)

    def __enter__(self):
        self.saved = (torch.randn, torch.randn_like, torch.rand_like, torch.randint)

        def randn(*shape, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else tuple(shape)
            return self._next("randn", shape, kw.get("dtype") or torch.float32)

        def randn_like(x, **kw):
            return self._next("randn", x.shape, kw.get("dtype") or x.dtype)

        def rand_like(x, **kw):
            return self._next("rand", x.shape, kw.get("dtype") or x.dtype)

        def randint(low, high, size, **kw):
            self.log.append(("randint", tuple(size), f"{low}:{high}"))
            v = T3["t_train"] if high - low > 1 else low
            return torch.full(tuple(size), v, dtype=torch.long, device=self.device)

        torch.randn, torch.randn_like, torch.rand_like, torch.randint = randn, randn_like, rand_like, randint
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like, torch.rand_like, torch.randint = self.saved
