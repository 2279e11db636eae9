"""View-sharded data parallelism for the mapping iteration (the reference itself is
single-GPU, batch = 1 view).

Every rank holds a full replica of the Gaussians and renders its own training view(s); the one exchange
step is a SUM all-reduce of the per-Gaussian gradients (59 floats / Gaussian at SH degree 3) plus a MAX
all-reduce of the 1-byte visibility mask, after which every replica applies the identical masked Adam
step (identical inputs -> bit-identical replicas).  The gradient tensors handed to glic_backward are
views into ONE planar buffer, so the collective runs in place on exactly the bytes the backward kernels
wrote: no pack / unpack copy.  torch.distributed (NCCL over NVLink on the GPU box, gloo in CPU tests)
is the plumbing.
"""
import torch
import torch.distributed as dist

# (name, floats per Gaussian, shape suffix) -- rotations first: their float4 stores need 16-byte alignment
_LAYOUT = [("dL_drots", 4, (4,)), ("dL_dmeans3D", 3, (3,)), ("dL_dscales", 3, (3,)), ("dL_dopacity", 1, (1,)),
           ("dL_ddc", 3, (1, 3)), ("dL_dsh", None, None)]


class PackedGrads:
    """Planar [59*P] gradient buffer + scratch outputs of glic_backward that are not optimiser inputs."""

    def __init__(self, P, M, device):
        self.P, self.M = int(P), int(M)
        per = 4 + 3 + 3 + 1 + 3 + 3 * self.M
        self.flat = torch.zeros(self.P * per, dtype=torch.float32, device=device)
        self.grads, off = {}, 0
        for name, k, shape in _LAYOUT:
            if name == "dL_dsh":
                k, shape = 3 * self.M, (self.M, 3)
            n = self.P * k
            self.grads[name] = self.flat[off:off + n].view(self.P, *shape)
            off += n
        f32 = dict(dtype=torch.float32, device=device)
        for name, k in (("dL_dmeans2D", 3), ("dL_dconic", 4), ("dL_dcolors", 3), ("dL_dcov3D", 6)):
            self.grads[name] = torch.empty(self.P, k, **f32)
        self.visible = torch.zeros(self.P, dtype=torch.uint8, device=device)

    def payload_bytes(self):
        return self.flat.numel() * 4 + self.visible.numel()


class GradAllReduce:
    """mean over ranks of the packed gradients, union of visibility."""

    def __init__(self, P, M, device, group=None):
        self.packed = PackedGrads(P, M, device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    @property
    def grads(self):
        return self.packed.grads

    def __call__(self, radii):
        pk = self.packed
        torch.gt(radii[:pk.P], 0, out=pk.visible.view(torch.bool))
        if self.world > 1:
            dist.all_reduce(pk.flat, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(pk.visible, op=dist.ReduceOp.MAX, group=self.group)
            pk.flat.mul_(1.0 / self.world)
        return pk.grads, pk.visible


def shard_views(n_views, rank, world):
    """Views handled by `rank` when n_views training views are split over `world` ranks (round robin)."""
    return [v for v in range(n_views) if v % world == rank]
