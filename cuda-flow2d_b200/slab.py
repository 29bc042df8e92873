"""Wiring helpers for the row-slab decomposition (flow2d_compute_slab_device, include/flow2d.h).

The data path is inside the C library: neighbour ranks write halo rows straight into each other's mailbox (device
memory, peer-mapped) and signal with a flag; a rank's stream waits for the flag in a kernel.  What is left to the
caller is telling every handle where its neighbours' mailboxes are:

  SlabGroup     N handles in ONE process (one per device, or all on one device for tests), one host thread per rank
  connect_ipc   one process per GPU (torchrun): mailboxes are exchanged as CUDA IPC handles through torch.distributed
"""
import threading

from .sharding import slab_rows  # noqa: F401  (re-exported: same split as flow2d_slab_rows)


class SlabGroup:
    """`world` ranks in one process.  devices: one CUDA device index per rank (repeat an index to put several ranks
    on one GPU -- the bit-for-bit tests of the decomposition do that)."""

    def __init__(self, pkg, width, height, devices, constancy=0, min_rows=0):
        self.world = len(devices)
        self.handles = [pkg.Flow2D(width, height, constancy=constancy, device=d) for d in devices]
        boxes = [h.slab_mailbox()[0] for h in self.handles]
        for r, h in enumerate(self.handles):
            h.slab_connect(r, self.world, boxes[r - 1] if r > 0 else None, boxes[r + 1] if r < self.world - 1 else None, min_rows)

    def run(self, fn):
        """fn(rank, handle) on one host thread per rank (a rank's stream waits for kernels of its neighbours, so the
        ranks must be enqueued concurrently); returns the list of results."""
        out, err = [None] * self.world, []

        def work(r):
            try:
                out[r] = fn(r, self.handles[r])
            except Exception as e:  # pragma: no cover
                err.append(e)
        th = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=900)
        if err:
            raise err[0]
        return out

    def destroy(self):
        for h in self.handles:
            h.destroy()


def connect_ipc(handle, dist, rank, world, device, min_rows=0):
    """One process per GPU: all-gather the 64-byte CUDA IPC handles of the mailboxes with torch.distributed (any
    backend), map the two neighbours' mailboxes and connect.  Collective: every rank must call it."""
    import torch
    mine = torch.tensor(list(handle.slab_export()), dtype=torch.uint8)
    on_gpu = dist.get_backend() == "nccl"
    if on_gpu:
        mine = mine.to(device)
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    above = handle.slab_import(bytes(every[rank - 1].cpu().tolist())) if rank > 0 else None
    below = handle.slab_import(bytes(every[rank + 1].cpu().tolist())) if rank < world - 1 else None
    handle.slab_connect(rank, world, above, below, min_rows)
    dist.barrier()  # nobody computes before everybody has reset its flags
