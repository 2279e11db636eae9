"""Fused reduce-scatter -> Adam on the local slice -> all-gather of parameters (csrc/p2p.cu, glic_p2p_reduce_adam)
against glic_p2p_allreduce_mean followed by glic_adam_update_packed: bit-identical parameters on every rank."""
import ctypes as C
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

P, M = 10007, 15
LR6 = [1e-3, 1.6e-4, 5e-3, 5e-2, 2.5e-3, 1.25e-4]


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from gaussian_lic_b200 import capi, dist as gdist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cuda:%d" % (rank % torch.cuda.device_count()))
        torch.cuda.set_device(dev)
        lib = capi.lib
        fused = gdist.P2PModelExchange(P, M, dev, LR6)
        plain = gdist.P2PGradAllReduce(P, M, dev)
        n = fused.n_floats
        gen = torch.Generator(device="cpu").manual_seed(7)
        params0 = torch.randn(n, generator=gen).to(dev)                      # identical on every rank
        fused.params.copy_(params0)
        p_ref, m_ref, v_ref = params0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        for it in range(3):
            g = torch.Generator(device="cpu").manual_seed(100 * it + rank)
            grads = (torch.randn(n, generator=g) * 1e-3).to(dev)
            radii = (torch.rand(P, generator=g) < 0.4).to(torch.int32).to(dev)
            fused.packed.flat.copy_(grads)
            plain.packed.flat.copy_(grads)
            fused.step(radii)
            _, vis = plain(radii)
            capi.check(lib.glic_adam_update_packed(capi.ptr(p_ref), capi.ptr(plain.packed.flat), capi.ptr(m_ref), capi.ptr(v_ref),
                                                   capi.ptr(vis), (C.c_float * 6)(*LR6), 0.9, 0.999, 1e-15, P, M, None), "adam")
            torch.cuda.synchronize()
            assert torch.equal(fused.params, p_ref), "iteration %d: fused parameters differ" % it
        # moments: only this rank's slice is maintained by the fused kernel
        sl = (C.c_size_t * 6)()
        lib.glic_p2p_slice(rank, world, n, P, sl)
        lo, hi = int(sl[0]) * 4, min(int(sl[1]) * 4, n)
        assert torch.equal(fused.exp_avg[lo:hi], m_ref[lo:hi]) and torch.equal(fused.exp_avg_sq[lo:hi], v_ref[lo:hi])
        fused.close(); plain.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001
        q.put((rank, repr(e)))


@pytest.mark.timeout(300)
def test_fused_exchange_matches_allreduce_then_adam():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=250) for _ in procs]
    for p_ in procs:
        p_.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
