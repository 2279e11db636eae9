"""World-size-2 gloo test of the view-sharded gradient exchange (host-side logic of SURVEY 8e) -- CPU only."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, P, M, out_dir):
    import torch.distributed as dist
    from gaussian_lic_b200 import dist as gdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ar = gdist.GradAllReduce(P, M, torch.device("cpu"))
    gen = torch.Generator().manual_seed(100 + rank)
    for name in ("dL_drots", "dL_dmeans3D", "dL_dscales", "dL_dopacity", "dL_ddc", "dL_dsh"):
        ar.grads[name].copy_(torch.randn(ar.grads[name].shape, generator=gen))
    radii = (torch.rand(P, generator=gen) < 0.5).to(torch.int32) * 7
    local = {k: v.clone() for k, v in ar.grads.items()}
    grads, vis = ar(radii)
    torch.save({"local": local, "radii": radii, "reduced": {k: v.clone() for k, v in grads.items()}, "vis": vis.clone()},
               os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_grad_allreduce_world2(tmp_path):
    import torch.multiprocessing as mp
    P, M, world = 1001, 15, 2
    mp.spawn(_worker, args=(world, _free_port(), P, M, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, "r%d.pt" % i)) for i in range(world)]
    for name in ("dL_drots", "dL_dmeans3D", "dL_dscales", "dL_dopacity", "dL_ddc", "dL_dsh"):
        mean = (r[0]["local"][name] + r[1]["local"][name]) / 2
        for i in range(world):
            torch.testing.assert_close(r[i]["reduced"][name], mean, rtol=1e-6, atol=1e-7)
        assert torch.equal(r[0]["reduced"][name], r[1]["reduced"][name])          # replicas stay bit-identical
    union = ((r[0]["radii"] > 0) | (r[1]["radii"] > 0)).to(torch.uint8)
    assert torch.equal(r[0]["vis"], union) and torch.equal(r[1]["vis"], union)


def test_packed_layout_and_view_sharding():
    from gaussian_lic_b200 import dist as gdist
    pk = gdist.PackedGrads(10, 15, "cpu")
    assert pk.flat.numel() == 10 * 59 and pk.payload_bytes() == 10 * 59 * 4 + 10
    assert pk.grads["dL_drots"].data_ptr() == pk.flat.data_ptr()                   # float4-aligned block first
    assert pk.grads["dL_dsh"].shape == (10, 15, 3) and pk.grads["dL_ddc"].shape == (10, 1, 3)
    total = sum(pk.grads[k].numel() for k in ("dL_drots", "dL_dmeans3D", "dL_dscales", "dL_dopacity", "dL_ddc", "dL_dsh"))
    assert total == pk.flat.numel()
    pk.grads["dL_dsh"].fill_(1.0)
    assert pk.flat.sum().item() == 10 * 45
    views = [gdist.shard_views(8, r, 4) for r in range(4)]
    assert sorted(sum(views, [])) == list(range(8)) and all(len(v) == 2 for v in views)
    assert np.all([gdist.shard_views(8, 0, 1) == list(range(8))])


def _p2p_fail_worker(rank, world, port, out_dir):
    """Without a GPU the IPC allocation fails locally; every rank must learn it and raise together (no rank may be left
    inside a collective)."""
    import torch.distributed as dist
    from gaussian_lic_b200 import dist as gdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    msg = "no error"
    try:
        gdist.P2PGradAllReduce(64, 15, torch.device("cpu"))
    except RuntimeError as e:
        msg = str(e)
    dist.barrier()                                          # both ranks got out of the constructor
    open(os.path.join(out_dir, "p2p%d.txt" % rank), "w").write(msg)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.is_available(), reason="exercises the no-GPU failure path")
def test_p2p_setup_failure_is_collective(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_p2p_fail_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r_ in range(2):
        msg = open(os.path.join(tmp_path, "p2p%d.txt" % r_)).read()
        assert "P2P gradient exchange unavailable" in msg and "glic_p2p_alloc" in msg, msg
