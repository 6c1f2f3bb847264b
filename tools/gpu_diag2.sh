#!/bin/bash
mkdir -p gpurun_out
for t in 0 1 2 3 4 5 6; do timeout 60 ./tools/diag_tma $t >> gpurun_out/diag_tma.log 2>&1; echo "  rc=$?" >> gpurun_out/diag_tma.log; done
cat gpurun_out/diag_tma.log
cat > /tmp/repro.py <<'PY'
import numpy as np, torch, pd_unet_b200 as pdu
pdu.set_option("radon_fwd_variant", 1)
op = pdu.Radon(64, np.linspace(0, np.pi, 8, endpoint=False))
x = torch.rand(1, 64, 64, device="cuda")
y = op.forward(x)
torch.cuda.synchronize()
pdu.set_option("radon_fwd_variant", 0)
y0 = op.forward(x)
print("max diff", float((y - y0).abs().max()), "rel", float((y - y0).norm() / y0.norm()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/repro.py > gpurun_out/sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer.log
grep -v "^$" gpurun_out/sanitizer.log | head -60
