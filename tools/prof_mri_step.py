"""A few eager MRI PD-UNet inference steps for an ncu launch list: configs[0] (256^2, 32 -> 256 spokes, one slice) or the
configs[3] shape (320^2, 8 coils, 48 spokes, 2 slices).  python tools/prof_mri_step.py cfg1|cfg4 [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pd_unet_b200 as pdu
from pd_unet_b200 import data
from pd_unet_b200.model import PrimalDualUNetMRI

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = False
which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
if which == "cfg1":
    n, sp_s, sp_f = 256, 32, 256
    m = PrimalDualUNetMRI((n, n), sp_f, 2 * n, coils=1, n_iter=4, n_primal=4, n_dual=4, unet_base=32, unet_depth=3, dual_features=32).to(dev).eval()
    d = data.make_mri_batch((n, n), sp_s, 1, 1, seed=0, device=dev)
    om_full = data.radial_trajectory(sp_f, 2 * n, device=dev)
    dcf_full = pdu.calc_density_compensation_function(om_full, (n, n))
    run = lambda: m(d["kdata"], om_full, None, dcf_full, omega_sparse=d["omega"], dcf_sparse=d["dcf"])
else:
    n, coils, sp = 320, 8, 48
    m = PrimalDualUNetMRI((n, n), sp, 2 * n, coils=coils, n_iter=4, n_primal=4, n_dual=2 * coils, unet_base=32, unet_depth=3, dual_features=32).to(dev).eval()
    d = data.make_mri_batch((n, n), sp, coils, 2, seed=0, device=dev)
    run = lambda: m(d["kdata"], d["omega"], d["smaps"], d["dcf"])
with torch.no_grad():
    for it in range(steps):
        torch.cuda.nvtx.range_push(f"step{it}")
        out = run()
        torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
        print("step", it, tuple(out.shape), float(out.abs().mean()), flush=True)
