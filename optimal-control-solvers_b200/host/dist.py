"""Multi-GPU plumbing for the one-process-per-GPU launch (torchrun): the trajectory batch shards
trivially (independent trajectories) -- contiguous slices, no data-path
collective.  The only exchange is the gather of the small per-trajectory results J and x(T) TO
RANK 0 with torch.distributed (NCCL send/recv over NVLink on the GPU box, gloo in the CPU tests):
non-root ranks send (1+n) x Bt doubles, the root receives world-1 of them -- an eighth of the
bytes of the all-gather round 1 used, and no rank other than the root waits for anybody.
(Inside ONE process the C ABI does the same with ocs_b200_create_multi / ocs_b200_gather_results_dev.)"""
from __future__ import annotations


def shard_range(Bt_total: int, rank: int, world: int):
    """Contiguous slice [b0, b1) of a batch of Bt_total trajectories owned by `rank`;
    the first Bt_total % world ranks take one extra trajectory."""
    base, rem = divmod(int(Bt_total), int(world))
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def gather_results(J, xT, world: int, group=None):
    """All-gathers J [Bt_local] and xT [n, Bt_local] (torch tensors on the rank's device) into
    J_all [world*Bt_local] and xT_all [world, n, Bt_local].  Equal shard sizes (weak scaling)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return J, xT.unsqueeze(0)
    J_all = torch.empty(world * J.numel(), dtype=J.dtype, device=J.device)
    xT_all = torch.empty((world,) + tuple(xT.shape), dtype=xT.dtype, device=xT.device)
    dist.all_gather_into_tensor(J_all, J.contiguous(), group=group)
    dist.all_gather_into_tensor(xT_all.view(-1), xT.contiguous().view(-1), group=group)
    return J_all, xT_all


def gather_ragged(J, world: int, rank: int, Bt_total: int, group=None):
    """All-gather of J when Bt_total does not divide evenly: pads every shard to the largest
    one, gathers, and strips the padding.  Returns the full [Bt_total] tensor on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return J
    sizes = [shard_range(Bt_total, r, world) for r in range(world)]
    big = max(b1 - b0 for b0, b1 in sizes)
    pad = torch.zeros(big, dtype=J.dtype, device=J.device)
    pad[: J.numel()] = J
    out = torch.empty(world * big, dtype=J.dtype, device=J.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * big: r * big + (b1 - b0)] for r, (b0, b1) in enumerate(sizes)])


def gather_to_root(res, out, world: int, rank: int, group=None, root: int = 0):
    """Gathers the packed per-trajectory results res [(1+n) x Bt_local] of every rank into out
    [world, 1+n, Bt_local] on `root` (out is ignored elsewhere).  NCCL implements it as grouped
    point-to-point transfers; non-root ranks only send."""
    import torch.distributed as dist
    if world == 1:
        out[0].copy_(res)
        return
    dist.gather(res, [out[r] for r in range(world)] if rank == root else None, dst=root, group=group)


class PipelinedGather:
    """Gather of the packed per-trajectory results [J; x(T)] ((1+n) x Bt_local per rank) that overlaps with the
    next evaluation: results are double-buffered, the collective runs on its own CUDA stream behind an event of
    the compute stream, and a buffer is only handed out again once its gather has completed.
    mode "root": gather to rank 0 with NCCL send/recv (the north star's "gather the small per-trajectory J and
    final-state results"); mode "all": all-gather (every rank gets everything; 8x the bytes at 8 ranks);
    mode "peer": gather to rank 0 by PEER COPIES -- the root's two result buffers are opened by the other ranks
    over CUDA IPC and every rank pushes its slice with one cudaMemcpyAsync on its own copy engine over NVLink; NCCL
    only carries a 4-byte completion token per rank (gathered to the root) that orders the pushes on the root's stream.
    Measured at 8 GPUs on one box: 4.987 ms per step against 4.961 for mode "root" -- rank 0 is the slow rank in both
    (4.96-4.99 against 4.82-4.92 ms), so it is not the 112 MiB the root receives that costs; with an all-reduce as the
    completion every rank waits for the slowest and the step grows to 5.22 ms (profiles/r2_ab_gather_n8.txt).  Kept as
    an option,
    not the default.  Falls back to "root" (on every rank) when the IPC export / open fails anywhere.

        pg = PipelinedGather(n, Bt, world, device, rank=rank)
        for step in ...:
            res = pg.next_buffer()          # (1+n, Bt) tensor: row 0 = J, rows 1.. = x(T)
            ... sweeps write res ...
            pg.gather(res)                  # asynchronous
        pg.finish()                         # pg.out[k] (world, 1+n, Bt) holds the gathered results of buffer k (on the root)
    """

    def __init__(self, n, Bt, world, device, group=None, rank=0, mode="root"):
        import torch
        self.world, self.group, self.rank, self.mode = world, group, rank, mode
        self.bufs = [torch.empty((1 + n, Bt), dtype=torch.float64, device=device) for _ in range(2)]
        holds_all = mode == "all" or rank == 0
        self.out = [torch.empty((world if holds_all else 1, 1 + n, Bt), dtype=torch.float64, device=device) for _ in range(2)]
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = [None, None]
        self.k = 1
        self.peer = None            # mode "peer", ranks > 0: device pointers of this rank's slice in the root's two buffers
        if mode == "peer":
            if world > 1 and self.cuda:
                self._peer_setup(n, Bt, device)
            else:
                self.mode = "root"

    def _peer_setup(self, n, Bt, device):
        """Exports the root's buffers over CUDA IPC and opens them everywhere else; agrees on the mode."""
        import torch
        import torch.distributed as dist
        ok, exports = 1, None
        try:
            try:
                from cuda.bindings import driver as cu, runtime as rt
            except ImportError:         # older cuda-python layouts
                from cuda import cuda as cu, cudart as rt
            self._rt = rt
            torch.cuda.set_device(device)
            if self.rank == 0:
                exports = []
                for t in self.out:
                    err, base, _size = cu.cuMemGetAddressRange(t.data_ptr())
                    if int(err) != 0:
                        raise RuntimeError(f"cuMemGetAddressRange: {err}")
                    err, h = rt.cudaIpcGetMemHandle(int(base))
                    if int(err) != 0:
                        raise RuntimeError(f"cudaIpcGetMemHandle: {err}")
                    exports.append((bytes(h.reserved), t.data_ptr() - int(base)))
        except Exception:
            ok, exports = 0, None
        box = [exports]
        dist.broadcast_object_list(box, src=0, group=self.group)
        if self.rank != 0 and ok:
            try:
                if box[0] is None:
                    raise RuntimeError("the root exported nothing")
                self.peer = []
                slice_bytes = 8 * (1 + n) * Bt
                for hb, off in box[0]:
                    h = self._rt.cudaIpcMemHandle_t()
                    h.reserved = hb
                    err, ptr = self._rt.cudaIpcOpenMemHandle(h, self._rt.cudaIpcMemLazyEnablePeerAccess)
                    if int(err) != 0:
                        raise RuntimeError(f"cudaIpcOpenMemHandle: {err}")
                    self.peer.append(int(ptr) + off + self.rank * slice_bytes)
            except Exception:
                ok, self.peer = 0, None
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.mode, self.peer = "root", None
        else:
            self.flag = torch.zeros(1, dtype=torch.int32, device=device)
            self.flag_out = torch.zeros((self.world if self.rank == 0 else 1, 1), dtype=torch.int32, device=device)

    def next_buffer(self):
        import torch
        self.k ^= 1
        if self.cuda and self.done[self.k] is not None:
            torch.cuda.current_stream().wait_event(self.done[self.k])   # its previous gather has read it
        return self.bufs[self.k]

    def gather(self, res):
        import torch
        import torch.distributed as dist
        k = self.k
        assert res is self.bufs[k]
        if self.world == 1:
            self.out[k][0].copy_(res)
            return
        def collective():
            if self.mode == "all":
                dist.all_gather_into_tensor(self.out[k].view(-1), res.view(-1), group=self.group)
            elif self.mode == "peer":
                if self.rank == 0:
                    self.out[k][0].copy_(res)
                else:
                    rt = self._rt
                    (err,) = rt.cudaMemcpyAsync(self.peer[k], res.data_ptr(), res.numel() * 8,
                                                rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice, self.stream.cuda_stream)
                    if int(err) != 0:
                        raise RuntimeError(f"peer copy to the root failed: {err}")
                # completion: a 4-byte token gathered to the root -- its stream passes this point once every rank's push
                # (stream-ordered before that rank's send) has finished; no rank waits for any rank but the root
                gather_to_root(self.flag, self.flag_out, self.world, self.rank, group=self.group)
            else:
                gather_to_root(res, self.out[k], self.world, self.rank, group=self.group)
        if not self.cuda:
            collective()
            return
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            collective()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[k] = ev

    def finish(self):
        import torch
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
