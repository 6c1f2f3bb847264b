"""Experiment: cuDNN runtime-fused conv3x3 + bias + PReLU (cudnn frontend graph) vs torch conv2d + pdu bias_prelu."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cudnn
import pd_unet_b200 as pdu
from pd_unet_b200 import updates

dev = "cuda:0"
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
N, C, K, H, W = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 32, 32, 256, 256
x = torch.randn(N, C, H, W, device=dev).contiguous(memory_format=torch.channels_last)
w = (torch.randn(K, C, 3, 3, device=dev) * 0.05).contiguous(memory_format=torch.channels_last)
b = torch.randn(1, K, 1, 1, device=dev).contiguous(memory_format=torch.channels_last)
a = (torch.rand(1, K, 1, 1, device=dev) * 0.5).contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(name, fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    print(f"{name:40s} median {statistics.median(ts):8.1f} us  min {min(ts):8.1f} us", flush=True)

def ref():
    y = torch.nn.functional.conv2d(x, w, None, padding=1)
    return torch.nn.functional.prelu(y + b, a.flatten())

def base():
    y = torch.nn.functional.conv2d(x, w, None, padding=1)
    return updates.bias_prelu(y, b.flatten(), a.flatten())

try:
    timed("torch conv2d only", lambda: torch.nn.functional.conv2d(x, w, None, padding=1))
    timed("torch conv2d + pdu bias_prelu", base)
except Exception as ex:
    print("baseline failed:", repr(ex))

handle = cudnn.create_handle()
stream = torch.cuda.current_stream().cuda_stream
cudnn.set_stream(handle=handle, stream=stream)
for comp in (cudnn.data_type.FLOAT,):
    g = cudnn.pygraph(io_data_type=cudnn.data_type.FLOAT, intermediate_data_type=cudnn.data_type.FLOAT,
                      compute_data_type=comp, handle=handle)
    X = g.tensor_like(x); Wt = g.tensor_like(w); B = g.tensor_like(b); A = g.tensor_like(a)
    y = g.conv_fprop(image=X, weight=Wt, padding=[1, 1], stride=[1, 1], dilation=[1, 1])
    yb = g.bias(name="bias", input=y, bias=B)
    p = g.relu(input=yb)
    n = g.relu(input=g.neg(input=yb))
    out = g.sub(a=p, b=g.mul(a=n, b=A))
    out.set_output(True).set_data_type(cudnn.data_type.FLOAT)
    try:
        g.validate(); g.build_operation_graph()
        g.create_execution_plans([cudnn.heur_mode.A, cudnn.heur_mode.FALLBACK])
        g.check_support(); g.build_plans(cudnn.build_plan_policy.ALL)
    except Exception as ex:
        print("fused graph not supported:", repr(ex)[:400]); continue
    nplans = g.get_execution_plan_count()
    print("plans:", nplans)
    o = torch.empty(N, K, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    want = ref()
    for i in range(nplans):
        try:
            ws = torch.empty(max(1, g.get_workspace_size_plan_at_index(i)), dtype=torch.uint8, device=dev)
            run = lambda: g.execute_plan_at_index({X: x, Wt: w, B: b, A: a, out: o}, ws, i, handle=handle)
            run(); torch.cuda.synchronize()
            err = ((o - want).norm() / want.norm()).item()
            timed(f"fused plan {i} {g.get_plan_name_at_index(i)[:24]} err {err:.1e}", run, reps=5)
        except Exception as ex:
            print("plan", i, "failed:", repr(ex)[:200])
