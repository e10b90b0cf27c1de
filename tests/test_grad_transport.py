"""``--grad-transport bf16`` on the library-collective path (``TorchDistComm``): gradients cross the wire
as bf16, the owner's shard comes back as fp32 within bf16 rounding of the exact sum.  Two gloo processes."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q) -> None:
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spacy_ray_b200.parallel.sync_proxy import TorchDistComm

        cap = 1024
        layout = types.SimpleNamespace(shard_cap=cap)
        g = torch.Generator().manual_seed(100 + rank)
        grad = torch.randn(cap * world, generator=g) * 0.1
        grads = [torch.randn(cap * world, generator=torch.Generator().manual_seed(100 + r)) * 0.1 for r in range(world)]
        exact = sum(grads)[rank * cap:(rank + 1) * cap]
        out = {}
        for mode in ("fp32", "bf16"):
            comm = TorchDistComm(rank, world, grad_transport=mode)
            got = comm.reduce_scatter(grad.clone(), layout)
            assert got.dtype == torch.float32 and got.shape == (cap,)
            out[mode] = float((got - exact).abs().max())
            # the second call reuses the staging buffers
            got2 = comm.reduce_scatter(grad.clone(), layout)
            assert torch.equal(got, got2)
        q.put((rank, out, None))
    except BaseException as e:           # noqa: BLE001 - reported to the parent
        q.put((rank, None, repr(e)))
    finally:
        dist.destroy_process_group()


def test_bf16_gradient_transport_matches_fp32_within_bf16_rounding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, out, err in results:
        assert err is None, (rank, err)
        assert out["fp32"] <= 1e-6
        # each addend is rounded to bf16 (2^-9 relative) and the partial sum is rounded again
        assert 0.0 < out["bf16"] <= 4e-3, out


def test_unknown_transport_is_rejected():
    from spacy_ray_b200.parallel.sync_proxy import TorchDistComm

    with pytest.raises(ValueError):
        TorchDistComm(0, 1, grad_transport="fp8")
