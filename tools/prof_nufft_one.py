"""One NUFFT forward + adjoint at the cfg4 share (320^2, 8 coils, batch 8, 48 spokes) per variant given (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pd_unet_b200 as pdu
from pd_unet_b200.phantoms import coil_maps
dev = "cuda:0"
n, coils, B, spokes = 320, 8, 8, 48
phi = np.arange(spokes) * (111.246117975 * np.pi / 180.0)
r = (np.arange(2 * n) - n) * (2 * np.pi / (2 * n))
om = torch.from_numpy(np.stack([(r[None] * np.sin(phi)[:, None]).reshape(-1), (r[None] * np.cos(phi)[:, None]).reshape(-1)]).astype(np.float32)).to(dev)
sm = coil_maps(coils, n)[None].to(dev)
img = torch.randn(B, 1, n, n, dtype=torch.complex64, device=dev)
fw, ad = pdu.KbNufft((n, n)), pdu.KbNufftAdjoint((n, n))
for v in [int(a) for a in sys.argv[1:]] or [-1]:
    pdu.set_option("nufft_fwd_variant", v); pdu.set_option("nufft_adj_variant", v)
    for _ in range(2):
        k = fw(img, om, smaps=sm)
        x = ad(k, om, smaps=sm)
    torch.cuda.synchronize()
    print("variant", v, "done", flush=True)
