"""Runs the backprojector variants given on the command line at configs[1] sizes (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu

N, A, B = 256, 512, 16
op = pdu.Radon(N, np.linspace(0, np.pi, A, endpoint=False))
s = torch.rand(B, A, N, device="cuda:0")
for v in [int(a) for a in sys.argv[1:]] or [-1]:
    pdu.set_option("radon_adj_variant", v)
    for _ in range(2):
        op._backproject(s)
    torch.cuda.synchronize()
    print("variant", v, "done", flush=True)
