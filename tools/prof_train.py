"""Three CT training steps (configs[1] shape, batch 8) for an ncu launch list: where a training step's time goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
from pd_unet_b200 import data
from pd_unet_b200.model import PrimalDualUNetCT, PrimalDualUNetMRI

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = False
KW = dict(n_iter=4, n_primal=4, n_dual=4, unet_base=32, unet_depth=3, dual_features=32)
which = sys.argv[1] if len(sys.argv) > 1 else "ct"
torch.manual_seed(0)
if which == "ct":
    radon = pdu.Radon(256, np.linspace(0, np.pi, 512, endpoint=False))
    b = data.make_ct_batch(radon, 8, 8, seed=0, device=dev)
    model, inputs, target = PrimalDualUNetCT(radon, upsample=8, **KW).to(dev), (b["sino_sparse"],), b["image"]
else:
    b = data.make_mri_batch((320, 320), 48, 8, 2, seed=0, device=dev)
    model = PrimalDualUNetMRI((320, 320), 48, 640, coils=8, n_iter=4, n_primal=4, n_dual=16, unet_base=32, unet_depth=3,
                              dual_features=32).to(dev)
    inputs, target = (b["kdata"], b["omega"], b["smaps"], b["dcf"]), b["image"]
opt = torch.optim.Adam(model.parameters(), 1e-4)
for it in range(3):
    opt.zero_grad(set_to_none=True)
    torch.cuda.nvtx.range_push(f"step{it}")
    loss = (model(*inputs) - target).abs().pow(2).mean()
    loss.backward()
    opt.step()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    print("step", it, float(loss), flush=True)
