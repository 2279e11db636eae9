"""The N>1 exchange step (csrc/p2p.cu): in-place mean all-reduce of the packed gradients + OR of visibility over
CUDA-IPC peer mappings, checked against the mean computed from the same seeds with torch.

Two ranks are spawned as separate processes; they use two GPUs when the box has them and share cuda:0 otherwise
(IPC mappings and the flag protocol are identical, the kernels just time-slice)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

P, M = 10007, 15


def _fill(rank, n_floats, n_vis, device):
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    flat = torch.randn(n_floats, generator=g)
    rad = (torch.rand(n_vis, generator=g) < 0.3).to(torch.int32)
    return flat.to(device), rad.to(device)


def test_p2p_single_rank_identity():
    from gaussian_lic_b200 import dist as gdist
    dev = torch.device("cuda:0")
    ex = gdist.P2PGradAllReduce(P, M, dev)
    flat, rad = _fill(0, ex.n_floats, P, dev)
    for _ in range(3):
        ex.packed.flat.copy_(flat)
        grads, vis = ex(rad)
        torch.cuda.synchronize()
        assert torch.equal(ex.packed.flat, flat)
        assert torch.equal(vis.bool(), rad > 0)
    assert grads["dL_dsh"].shape == (P, M, 3) and grads["dL_dsh"].data_ptr() >= ex.packed.flat.data_ptr()
    ex.close()


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from gaussian_lic_b200 import dist as gdist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cuda:%d" % (rank % torch.cuda.device_count()))
        torch.cuda.set_device(dev)
        ex = gdist.P2PGradAllReduce(P, M, dev)
        fills = [_fill(r, ex.n_floats, P, dev) for r in range(world)]
        want = sum(f for f, _ in fills) * (1.0 / world)
        want_vis = torch.stack([r_ > 0 for _, r_ in fills]).any(0)
        for it in range(4):
            ex.packed.flat.copy_(fills[rank][0] * (it + 1))
            _, vis = ex(fills[rank][1])
            torch.cuda.synchronize()
            # fp32 sum order differs from torch's by at most one rounding per add
            torch.testing.assert_close(ex.packed.flat, want * (it + 1), rtol=1e-6, atol=1e-6)
            assert torch.equal(vis.bool(), want_vis)
        ex.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001 - reported to the parent
        q.put((rank, repr(e)))


@pytest.mark.timeout(240)
def test_p2p_two_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=200) for _ in procs]
    for p_ in procs:
        p_.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
