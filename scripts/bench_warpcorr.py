import sys, torch
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from diffuvolume_b200 import ops
B, C, H, W = 8, 32, 384, 1248
g = torch.Generator(device="cuda").manual_seed(0)
fl = torch.randn(B, C, H, W, device="cuda", generator=g); fr = torch.randn(B, C, H, W, device="cuda", generator=g)
disp = torch.rand(B, 1, H, W, device="cuda", generator=g) * 190
buf = torch.empty(B, 146, H, W, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
sep = t(lambda: ops.corr_volume_2sided(fl, ops.warp(fr, disp), 24, 1))
full = t(lambda: ops.refine_input_assemble(fl, fr, disp, 24, 1, corr_out=buf[:, 97:], diff_out=buf[:, :32], copy_out=buf[:, 32:64]))
def aten():
    w = ops.warp(fr, disp); c = ops.corr_volume_2sided(fl, w, 24, 1).squeeze(1)
    return torch.cat((fl - w, fl, buf[:, 64:96], disp, c), 1)
at = t(aten)
print(f"B={B}: warp+corr {sep:.4f} ms | warp + diff + copy + corr assembled in the concat buffer {full:.4f} ms | warp, corr, then sub + cat on ATen {at:.4f} ms")
from diffuvolume_b200 import _lib
lib = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
warped = torch.empty_like(fr)
HWs = H * W
t_w = t(lambda: lib.dv_warp_f32(fr.data_ptr(), disp.data_ptr(), warped.data_ptr(), B, C, H, W, st))
t_wa = t(lambda: lib.dv_warp_assemble_f32(fr.data_ptr(), disp.data_ptr(), fl.data_ptr(), warped.data_ptr(), buf[:, :32].data_ptr(), 146 * HWs, buf[:, 32:64].data_ptr(), 146 * HWs, B, C, H, W, st))
t_wd = t(lambda: lib.dv_warp_assemble_f32(fr.data_ptr(), disp.data_ptr(), fl.data_ptr(), warped.data_ptr(), buf[:, :32].data_ptr(), 146 * HWs, None, 0, B, C, H, W, st))
cc = torch.empty(B, 49, H, W, device="cuda")
t_c = t(lambda: lib.dv_corr_volume_2sided_f32(fl.data_ptr(), warped.data_ptr(), cc.data_ptr(), B, C, H, W, 24, 1, st))
t_cs = t(lambda: lib.dv_corr_volume_2sided_into_f32(fl.data_ptr(), warped.data_ptr(), buf[:, 97:].data_ptr(), 146, 0, B, C, H, W, 24, st))
print(f"warp {t_w:.4f} | warp_assemble(diff+copy) {t_wa:.4f} | warp_assemble(diff) {t_wd:.4f} | corr {t_c:.4f} | corr into buffer {t_cs:.4f}")
