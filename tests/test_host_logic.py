"""Host-side logic of the package on CPU: filter taps and Kaiser-Bessel tables against the oracle,
geometry defaults, batch sharding, and the world_size-2 paths on the gloo backend."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from pd_unet_b200 import parallel
from pd_unet_b200.nufft import kaiser_bessel_scaling, kaiser_bessel_table
from pd_unet_b200.radon import FILTERS, Radon, RadonFanbeam, filter_taps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", FILTERS)
@pytest.mark.parametrize("D,A", [(64, 24), (77, 33), (256, 512)])
def test_filter_taps_equal_the_fft_route(name, D, A):
    """Two constructions: the product's (scikit-image style: np.fft of the constant table, numpy windows, fftshift)
    against the oracle's (closed-form ramp kernel, windows as explicit functions of the frequency index, plain
    cosine sums, direct Toeplitz product) -- they share no code."""
    taps = filter_taps(D, A, name)
    assert taps.shape == (2 * D - 1,)
    assert np.allclose(taps, taps[::-1], atol=1e-15)                 # even => symmetric Toeplitz => self-adjoint
    assert np.allclose(taps, oracle.filter_taps(D, name) * np.pi / (2 * A), atol=1e-15)
    s = torch.randn(2, A, D, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    j = np.arange(D)[:, None]
    i = np.arange(D)[None, :]
    H = torch.from_numpy(taps[(i - j) + D - 1])
    assert torch.allclose(s @ H, oracle.filter_sinogram(s, name), atol=1e-12)


@pytest.mark.parametrize("n,k", [(32, 64), (320, 640), (48, 80)])
def test_kaiser_bessel_tables_equal_the_oracle(n, k):
    """Two constructions: np.i0 on a vector grid + the closed-form transform sinh(w)/w (product) against scipy's
    exponentially scaled i0e entry by entry + Gauss-Legendre quadrature of the kernel's Fourier integral (oracle)."""
    spec = oracle.NufftSpec((n, n), grid_size=(k, k))
    assert np.allclose(kaiser_bessel_table(n, k, 6, 1024, 2.34), oracle.kb_table(spec, 0), atol=1e-14)
    assert np.allclose(kaiser_bessel_scaling(n, k, 6, 2.34), oracle.scaling_coef(spec, 0), rtol=1e-13)


def test_geometry_defaults_follow_torch_radon():
    r = Radon(128, np.linspace(0, np.pi, 10, endpoint=False))
    assert (r.det_count, r.det_spacing, r.geom.geom) == (128, 1.0, 0)
    assert np.allclose(r._internal, -r.angles)
    f = RadonFanbeam(128, np.linspace(0, 2 * np.pi, 10, endpoint=False), 256.0)
    assert (f.det_distance, f.det_spacing, f.det_count) == (256.0, 2.0, 128)
    f2 = RadonFanbeam(128, [0.0], 200.0, det_distance=100.0)
    assert f2.det_spacing == pytest.approx(1.5)
    with pytest.raises(ValueError):
        Radon(64, [])
    with pytest.raises(ValueError):
        RadonFanbeam(64, [0.0], -1.0)


@pytest.mark.parametrize("n,world", [(64, 8), (16, 1), (10, 4), (3, 8), (0, 2)])
def test_shard_range_partitions_the_batch(n, world):
    spans = [parallel.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    # inference: each rank filters its share of a sinogram batch (CPU oracle stands in for the kernels)
    full = torch.randn(5, 12, 16, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    mine = oracle.filter_sinogram(parallel.shard_batch(full, r, w))
    gathered = parallel.gather_batch(mine, full.shape[0])
    if r == 0:
        assert torch.allclose(gathered, oracle.filter_sinogram(full))
    assert parallel.max_over_ranks(1.0 + r) == float(w)
    assert parallel.sum_over_ranks(float(mine.shape[0])) == 5.0
    # training: DDP averages the gradients of the two shards
    from pd_unet_b200.model import DualBlock
    torch.manual_seed(0)
    net = parallel.wrap_ddp(DualBlock(2, 1, 4), r)
    x = torch.randn(4, 2, 8, 8, generator=torch.Generator().manual_seed(1))
    loss = net(parallel.shard_batch(x, r, w)).pow(2).mean()
    loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    torch.save(g, os.path.join(out_dir, f"g{r}.pt"))
    parallel.barrier()
    dist.destroy_process_group()


def test_world_size_2_on_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = (torch.load(tmp_path / f"g{r}.pt") for r in range(2))
    assert torch.equal(g0, g1)
    from pd_unet_b200.model import DualBlock
    torch.manual_seed(0)
    net = DualBlock(2, 1, 4)
    x = torch.randn(4, 2, 8, 8, generator=torch.Generator().manual_seed(1))
    net(x).pow(2).mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(g0, want, atol=1e-6)


def test_graph_capture_switch_sets_nccl_env(monkeypatch):
    """init_distributed(graph_capture=True) must turn NCCL's async error handling off before any process group
    exists (whole-step CUDA-graph capture with DDP), and is a no-op for the communicator at world size 1."""
    from pd_unet_b200 import parallel
    monkeypatch.delenv("TORCH_NCCL_ASYNC_ERROR_HANDLING", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert parallel.init_distributed(graph_capture=True)[:2] == (0, 1)
    import os
    assert os.environ.get("TORCH_NCCL_ASYNC_ERROR_HANDLING") == "0"
    m = torch.nn.Linear(2, 2)
    assert parallel.wrap_ddp(m, 0, graph_capture=True) is m        # world size 1: identity


def test_training_epilogues_refuse_cpu_tensors():
    """No CPU fallback: the differentiable epilogues raise on CPU tensors like every other operator wrapper."""
    from pd_unet_b200 import updates, PduError
    y = torch.zeros(1, 4, 3, 3)
    with pytest.raises(PduError):
        updates.bias_prelu(y, torch.zeros(4), torch.zeros(4))
    with pytest.raises(PduError):
        updates.bias_add(y, torch.zeros(4))
