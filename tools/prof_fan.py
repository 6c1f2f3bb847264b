"""Fan-beam forward + backprojection at the cfg3 per-GPU share (512^2, 1024 views, batch 8) (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
fan = pdu.RadonFanbeam(512, np.linspace(0, 2 * np.pi, 1024, endpoint=False), 1024.0)
x = torch.rand(8, 512, 512, device="cuda:0")
s = torch.rand(8, 1024, 512, device="cuda:0")
for _ in range(2):
    fan._project(x); fan._backproject(s)
torch.cuda.synchronize()
print("done")
