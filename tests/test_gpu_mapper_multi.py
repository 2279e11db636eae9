"""Render-level multi-GPU parity: an N-rank mapping step must equal the 1-rank step that accumulates the
same views -- union of visibility, mean gradient, identical parameters after Adam -- and the replicas must stay
bit-identical to each other.

Two ranks are spawned as separate processes (two GPUs when the box has them, one shared GPU otherwise: the IPC
mappings, the copy-engine pushes and the flag protocol are the same).  Rank r renders slot r of every batch; the
single-process arm renders both slots itself (views_per_rank = 2)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

W, H, F = 320, 208, 250.0
P = 5000
BATCHES = {2: [[0, 1], [2, 0], [1, 2], [0, 2]],                       # views per iteration -> the batches of the run
           4: [[0, 1, 2, 0], [2, 1, 0, 1], [1, 2, 2, 0]]}


def _scene():
    from gaussian_lic_b200 import synthetic as syn
    from helpers import small_scene
    g, _ = small_scene(P, W, H, 44, 3)
    poses = [syn.orbit_pose(v, radius=2.0) for v in (0, 2, 5)]
    gts = [syn.make_gt_image(W, H, seed=21 + i) for i in range(3)]
    return g, poses, gts


def _run(rank, world, k, handles_exchange):
    from gaussian_lic_b200 import mapper
    batches = BATCHES[world * k]
    g, poses, gts = _scene()
    m = mapper.Mapper(W, H, F, F, W / 2.0, H / 2.0, sh_degree=3, capacity=P, rank=rank, world=world, views_per_rank=k)
    m.initialize(g)
    for (R, t), img in zip(poses, gts):
        m.add_keyframe(R, t, img)
    if world > 1:
        m.connect(handles_exchange(m.export_handle()))
    st = m.optimize([v for b in batches for v in b])
    assert st.iterations == len(batches) and st.overflow_regrows == 0
    out = m.download(moments=True)
    return m, out


def _worker(rank, world, port, q, k):
    try:
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(rank % torch.cuda.device_count())

        def exchange(mine):
            got = [None] * world
            dist.all_gather_object(got, mine)
            return got

        m, out = _run(rank, world, k, exchange)
        dist.barrier()
        m.close()
        dist.destroy_process_group()
        q.put((rank, "ok", {k: v for k, v in out.items()}))
    except Exception as e:          # noqa: BLE001 - reported to the parent
        import traceback
        q.put((rank, "%r\n%s" % (e, traceback.format_exc()), None))


@pytest.mark.timeout(300)
@pytest.mark.parametrize("k", [1, 2])                        # views per rank: 2 ranks x k views == 1 rank x 2k views
def test_two_ranks_equal_one_rank_accumulation(k):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, k)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p_ in procs:
        p_.join(timeout=30)
    assert all(r[1] == "ok" for r in res), [r[1] for r in res]
    a, b = res[0][2], res[1][2]
    for name in a:                                           # replicas: bit-identical parameters AND moments
        assert np.array_equal(a[name], b[name]), "replicas diverged in %s" % name
    m, one = _run(0, 1, 2 * k, None)
    n_batches = len(BATCHES[2 * k])
    m.close()
    lr = dict(rots=0.001, means=1.6e-4, log_scales=0.005, opacity_logits=0.05, dc=2.5e-3, sh=2.5e-3 / 20)
    worst = {}
    for name in lr:
        d = np.abs(a[name].astype(np.float64) - one[name].astype(np.float64))
        worst[name] = d.max() / lr[name]
        # (g0 + g1) * 0.5 is evaluated with the same operations in both arms (two-shot reduce vs local accumulation); the
        # render backward's atomic order is the only difference between two runs => agreement far below one Adam step
        assert np.quantile(d, 0.999) <= 0.02 * lr[name] + 1e-7, (name, np.quantile(d, 0.999) / lr[name])
        assert d.max() <= 6.4 * n_batches * lr[name] + 1e-7, (name, d.max() / lr[name])
    print("2 ranks vs 1 rank x 2 views: max |diff| / lr per group:", {n_: "%.3g" % v for n_, v in worst.items()})
