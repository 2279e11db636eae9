"""View-sharded data parallelism for the mapping iteration (the reference itself is
single-GPU, batch = 1 view).

Every rank holds a full replica of the Gaussians and renders its own training view(s); the one exchange
step is a SUM all-reduce of the per-Gaussian gradients (59 floats / Gaussian at SH degree 3) plus a MAX
all-reduce of the 1-byte visibility mask, after which every replica applies the identical masked Adam
step (identical inputs -> bit-identical replicas).  The gradient tensors handed to glic_backward are
views into ONE planar buffer, so the collective runs in place on exactly the bytes the backward kernels
wrote: no pack / unpack copy.  GradAllReduce runs it through torch.distributed (gloo in the CPU
tests, NCCL as the library baseline); P2PGradAllReduce is the product path on the GPU box: our own kernels
over NVLink peer memory.
"""
import torch
import torch.distributed as dist

# (name, floats per Gaussian, shape suffix) -- rotations first: their float4 stores need 16-byte alignment
_LAYOUT = [("dL_drots", 4, (4,)), ("dL_dmeans3D", 3, (3,)), ("dL_dscales", 3, (3,)), ("dL_dopacity", 1, (1,)),
           ("dL_ddc", 3, (1, 3)), ("dL_dsh", None, None)]


class PackedGrads:
    """Planar [59*P] gradient buffer + scratch outputs of glic_backward that are not optimiser inputs."""

    def __init__(self, P, M, device, flat=None, visible=None):
        self.P, self.M = int(P), int(M)
        per = 4 + 3 + 3 + 1 + 3 + 3 * self.M
        self.flat = torch.zeros(self.P * per, dtype=torch.float32, device=device) if flat is None else flat
        assert self.flat.numel() == self.P * per
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
        self.visible = torch.zeros(self.P, dtype=torch.uint8, device=device) if visible is None else visible

    @staticmethod
    def floats_per_gaussian(M):
        return 4 + 3 + 3 + 1 + 3 + 3 * int(M)

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


class _DevMem:
    """Raw device allocation exposed through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class P2PGradAllReduce:
    """Same contract as GradAllReduce, but the exchange is our own two-shot all-reduce over NVLink peer memory
    (csrc/p2p.cu, glic_p2p_allreduce_mean): every rank's gradient buffer is one cudaMalloc block mapped into all
    peers through CUDA IPC; torch.distributed only carries the 64-byte handles at construction time.  The call
    enqueues three kernels on the current stream, never synchronises the host, and is CUDA-graph capturable."""

    def __init__(self, P, M, device, group=None):
        import ctypes as C
        from . import capi
        self._C, self._capi = C, capi
        self.lib = capi.lib
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise ValueError("P2PGradAllReduce supports up to 8 ranks (one NVSwitch domain)")
        self.n_floats = int(P) * PackedGrads.floats_per_gaussian(M)
        self.n_vis = int(P)
        nbytes = self._nbytes()
        # Every phase that can fail locally (allocation, IPC export, IPC import) is followed by an exchange of the outcome,
        # so that all ranks raise together instead of one rank leaving the others inside a collective.
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        err = None
        try:
            capi.check(self.lib.glic_p2p_alloc(nbytes, C.byref(own), handle), "glic_p2p_alloc")
        except Exception as e:                      # noqa: BLE001 - reported to every rank below
            err = repr(e)
        self._own = own.value
        mine = (err, handle.raw)
        infos = [mine] * self.world
        if self.world > 1:
            dist.all_gather_object(infos, mine, group=group)
        self._peers = (C.c_void_p * self.world)()
        self._opened = []
        bad = [(q, i[0]) for q, i in enumerate(infos) if i[0] is not None]
        if not bad:
            for q in range(self.world):
                if q == self.rank:
                    self._peers[q] = self._own
                    continue
                peer = C.c_void_p()
                try:
                    capi.check(self.lib.glic_p2p_open(infos[q][1], C.byref(peer)), "glic_p2p_open")
                    self._peers[q] = peer.value
                    self._opened.append(peer.value)
                except Exception as e:              # noqa: BLE001
                    err = "rank %d -> %d: %r" % (self.rank, q, e)
                    break
            outcomes = [err] * self.world
            if self.world > 1:
                dist.all_gather_object(outcomes, err, group=group)
            bad = [(q, o) for q, o in enumerate(outcomes) if o is not None]
        if bad:
            self._release()
            raise RuntimeError("P2P gradient exchange unavailable: %s" % (bad,))
        raw = torch.as_tensor(_DevMem(self._own, nbytes), device=device)
        self._raw = raw
        f_bytes = (self.n_floats * 4 + 255) // 256 * 256
        flat = raw[:self.n_floats * 4].view(torch.float32)
        visible = raw[f_bytes:f_bytes + self.n_vis]
        self.packed = PackedGrads(P, M, device, flat=flat, visible=visible)
        if self.world > 1:
            dist.barrier(group=group)            # every peer has mapped every buffer before first use

    def _nbytes(self):
        return int(self.lib.glic_p2p_buffer_bytes(self.n_floats, self.n_vis))

    @property
    def grads(self):
        return self.packed.grads

    def __call__(self, radii):
        pk = self.packed
        torch.gt(radii[:pk.P], 0, out=pk.visible.view(torch.bool))
        stream = torch.cuda.current_stream().cuda_stream
        self._capi.check(self.lib.glic_p2p_allreduce_mean(self.rank, self.world, self._peers, self.n_floats, self.n_vis,
                                                          self._C.c_void_p(stream)), "glic_p2p_allreduce_mean")
        return pk.grads, pk.visible

    def close(self):
        if getattr(self, "_own", None) is None:
            return
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)
        self._release()

    def _release(self):
        for p_ in getattr(self, "_opened", []):
            self.lib.glic_p2p_close(self._C.c_void_p(p_))
        self._opened = []
        if getattr(self, "_own", None):
            self.lib.glic_p2p_free(self._C.c_void_p(self._own))
        self._own = None


class P2PModelExchange(P2PGradAllReduce):
    """The fused form of the exchange on full packed gradients (tests/test_gpu_p2p_adam.py).  The mapped block additionally holds the PARAMETERS (packed layout); `step(radii)` reduce-scatters the gradients,
    applies the visibility-masked Adam to this rank's slice only and all-gathers the updated parameters into every
    replica (csrc/p2p.cu, glic_p2p_reduce_adam).  The moment buffers are touched on the local slice only."""

    def __init__(self, P, M, device, lr6, group=None, betas=(0.9, 0.999), eps=1e-15):
        super().__init__(P, M, device, group=group)
        C = self._C
        f_bytes = (self.n_floats * 4 + 255) // 256 * 256
        v_bytes = (self.n_vis + 255) // 256 * 256
        off = f_bytes + v_bytes + 256
        self.params = self._raw[off:off + self.n_floats * 4].view(torch.float32)
        padded = (self.n_floats + 63) // 64 * 64                      # the kernel works in float4 units
        self.exp_avg = torch.zeros(padded, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(padded, dtype=torch.float32, device=device)
        self.P_, self.M_ = int(P), int(M)
        self.lr6 = (C.c_float * 6)(*[float(x) for x in lr6])
        self.b1, self.b2, self.eps = float(betas[0]), float(betas[1]), float(eps)

    def _nbytes(self):
        return int(self.lib.glic_p2p_model_bytes(self.n_floats, self.n_vis))

    def step(self, radii):
        pk = self.packed
        torch.gt(radii[:pk.P], 0, out=pk.visible.view(torch.bool))
        stream = torch.cuda.current_stream().cuda_stream
        self._capi.check(self.lib.glic_p2p_reduce_adam(self.rank, self.world, self._peers, self.P_, self.M_,
                                                       self._capi.ptr(self.exp_avg), self._capi.ptr(self.exp_avg_sq), self.lr6,
                                                       self.b1, self.b2, self.eps, self._C.c_void_p(stream)), "glic_p2p_reduce_adam")
        return self.params


def shard_views(n_views, rank, world):
    """Views handled by `rank` when n_views training views are split over `world` ranks (round robin)."""
    return [v for v in range(n_views) if v % world == rank]
