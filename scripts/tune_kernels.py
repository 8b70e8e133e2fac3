"""Time individual hot-path kernels (B=8 pairs, BASELINE config shapes) under the tuning switches of
csrc/common.cuh:tune_variant.  Run on the GPU box; prints one line per (kernel, variant)."""
import os, sys, json
sys.path.insert(0, '.')
import torch
from diffuvolume_b200 import ops

B, H, W, D = 8, 540, 960, 48
h, w = H // 4, W // 4
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
PEAK = 6554.6

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2]

def report(name, variant, ms, gb):
    print(json.dumps({"kernel": name, "variant": variant, "ms": round(ms, 4), "GBs": round(gb/ms*1e3, 1), "frac": round(gb/ms*1e3/PEAK, 4)}), flush=True)

hw = h*w
# ---- softmax_regress
cost = rn(B, 192, H, W) * 4
used = torch.rand(B, H, W, device=dev) * 191
ens = torch.zeros(B, H, W, device=dev)
gb = B*(192*H*W*4 + 5*H*W*4)/1e9
for tma in (1, 0):
    os.environ["DV_SR_TMA"] = str(tma)
    ms = timeit(lambda: ops.softmax_regress(cost, used=used, vote_thresholds=(1.0, 3.0), ens_acc=ens, ens_coef=0.2))
    report("softmax_regress", f"tma{tma}", ms, gb)
    ms = timeit(lambda: ops.softmax_regress(cost))
    report("softmax_regress_disp_only", f"tma{tma}", ms, B*(192*H*W*4 + H*W*4)/1e9)
for v in (() if os.environ.get("DV_TUNE_ALL") is None else (0, 1, 2, 3, 4, 5)):
    os.environ["DV_SR_VARIANT"] = str(v)
    ms = timeit(lambda: ops.softmax_regress(cost, used=used, vote_thresholds=(1.0, 3.0), ens_acc=ens, ens_coef=0.2))
    report("softmax_regress", v, ms, gb)
    ms = timeit(lambda: ops.softmax_regress(cost))
    report("softmax_regress_disp_only", v, ms, B*(192*H*W*4 + H*W*4)/1e9)
os.environ.pop("DV_SR_VARIANT", None); os.environ.pop("DV_SR_TMA", None)
del cost
# ---- gwc
fl, fr = rn(B, 320, h, w), rn(B, 320, h, w)
out = torch.empty(B, 40, D, h, w, device=dev)
gb = B*(2*320*hw*4 + 40*D*hw*4)/1e9
for v in (2, 3):
    for tpc in (1, 2, 4, 8):
        os.environ["DV_GWC_MINB"] = str(v); os.environ["DV_GWC_TPC"] = str(tpc)
        report("gwc_volume", f"minb{v}_tpc{tpc}", timeit(lambda: ops.gwc_volume(fl, fr, D, 40, out=out)), gb)
os.environ.pop("DV_GWC_MINB"); os.environ.pop("DV_GWC_TPC")
del fl, fr, out
# ---- concat variants
cl, cr = rn(B, 32, h, w), rn(B, 32, h, w)
att = rn(B, 1, D, h, w)
xt64 = rn(B, D, h, w, dt=torch.float64)
xt32 = rn(B, D, h, w)
shift = rn(B, D) * 0.1
vol = torch.empty(B, 64, D, h, w, device=dev)
volf = torch.empty_like(vol)
gvol = 64*D*hw*4
for cpc in (8, 16, 32):
    os.environ["DV_CONCAT_CPC"] = str(cpc)
    report("concat_plain", cpc, timeit(lambda: ops.concat_volume(cl, cr, D, mask_left=False, out=vol)), B*(64*hw*4 + gvol)/1e9)
    report("concat_acv", cpc, timeit(lambda: ops.concat_volume(cl, cr, D, mask_left=False, att_logits=att, out=vol)), B*(64*hw*4 + D*hw*4 + gvol)/1e9)
    report("filter_regen_f64", cpc, timeit(lambda: ops.concat_volume(cl, cr, D, mask_left=False, att_logits=att, xt=xt64, shift=shift, out=volf)), B*(64*hw*4 + D*hw*12 + gvol)/1e9)
    report("filter_regen_f32", cpc, timeit(lambda: ops.concat_volume(cl, cr, D, mask_left=False, att_logits=att, xt=xt32, shift=shift, out=volf)), B*(64*hw*4 + D*hw*8 + gvol)/1e9)
os.environ.pop("DV_CONCAT_CPC")
att_w = ops.att_softmax(att)
nf = ops.filter_factor(xt64, shift, 1.0)
report("att_softmax", 0, timeit(lambda: ops.att_softmax(att)), B*2*D*hw*4/1e9)
report("filter_factor_f64", 0, timeit(lambda: ops.filter_factor(xt64, shift, 1.0)), B*D*hw*12/1e9)
for cg, ver, bar in ((8, 31, 0), (8, 31, 1), (16, 31, 0), (16, 31, 1), (8, 21, 0), (8, 21, 1), (8, 22, 0), (8, 22, 1), (4, 22, 1), (4, 31, 1)):
    os.environ["DV_CS_CGT"] = str(cg); os.environ["DV_CS_SHAPE"] = str(ver); os.environ["DV_CS_BAR"] = str(bar); cg = f"cgt{cg}_shape{ver}_bar{bar}"
    report("weighted_plain", cg, timeit(lambda: ops.concat_volume_weighted(cl, cr, D, mask_left=False, out=vol)), B*(64*hw*4 + gvol)/1e9)
    report("weighted_acv", cg, timeit(lambda: ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=att_w, out=vol)), B*(64*hw*4 + D*hw*4 + gvol)/1e9)
    report("weighted_filter", cg, timeit(lambda: ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=att_w, n=nf, out=volf)), B*(64*hw*4 + D*hw*8 + gvol)/1e9)
[os.environ.pop(k, None) for k in ("DV_CONCAT_CG", "DV_CS_CGT", "DV_CS_SHAPE", "DV_CS_BAR")]
report("filter_volume_f64", 0, timeit(lambda: ops.volume_filter(vol, xt64, shift, 1.0, out=volf)), B*(2*gvol + D*hw*8)/1e9)
# plain device copy of the same size for reference
report("torch_copy_3.2GB", 0, timeit(lambda: volf.copy_(vol)), 2*B*gvol/1e9)
report("torch_fill_3.2GB", 0, timeit(lambda: volf.fill_(1.0)), B*gvol/1e9)

# ---- ddim step
from diffuvolume_b200.pipeline import DdimSchedule
sch = DdimSchedule()
disp = torch.rand(B, H, W, device=dev) * 191
vote = (torch.rand(B, H, W, device=dev) > 0.5).float()
mask = torch.zeros(B, h, w, device=dev)
for dt in (torch.float32, torch.float64):
    xt = rn(B, D, h, w, dt=dt); sn = rn(B, D, h, w, dt=dt); rz = torch.rand(B, D, h, w, device=dev, dtype=torch.float64)
    san, c, sg = sch.update_coefficients(999, 799)
    f = lambda: ops.ddim_step(disp=disp, xt=xt, shift=shift, scale=1.0, sqrt_recip=sch.sqrt_recip(999), sqrt_recipm1=sch.sqrt_recipm1(999),
                              last_step=False, vote=vote, mask=mask, sqrt_alpha_next=san, c=c, sigma=sg, step_noise=sn, renoise=rz)
    es = xt.element_size()
    report("ddim_step_" + str(dt).split(".")[-1], 0, timeit(f), B*(8*H*W/16*4 + D*hw*(2*es + 8 + 4 + 8))/1e9)
