"""Transports for the row-slab decomposition (flow2d_compute_slab_device, include/flow2d.h).

The C library decides WHAT has to move (op 0: ghost rows of du, dv to/from the two neighbour ranks;
op 1: gather the increment of a level on every rank) and calls back; this module moves it:

  NcclExchange      one process per GPU, torch.distributed (NCCL over NVLink) send/recv + broadcast,
                    enqueued in stream order of the handle's stream -- the production path
  ThreadedExchange  N handles in N threads of ONE process on one GPU, device-to-device copies between
                    barriers -- used by the tests to check the decomposition bit for bit without a
                    multi-GPU box
"""
import ctypes as C
import threading

from .sharding import slab_rows

EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                          C.c_size_t, C.c_size_t, C.c_size_t)


class Slab(C.Structure):
    """`flow2d_slab`"""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("exchange", EXCHANGE_FN), ("user", C.c_void_p),
                ("min_rows_per_rank", C.c_size_t)]


class _DevMem:
    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def as_tensor(ptr, n_floats, device):
    import torch
    return torch.as_tensor(_DevMem(ptr, n_floats), device=device)


class NcclExchange:
    def __init__(self, dist, rank, world, device, stream):
        self.dist, self.rank, self.world, self.device, self.stream = dist, rank, world, device, stream
        self.calls = {0: 0, 1: 0}
        self.bytes = {0: 0, 1: 0}
        self.fn = EXCHANGE_FN(self._call)

    def slab(self, min_rows=0):
        return Slab(self.rank, self.world, self.fn, None, min_rows)

    def _call(self, user, op, p_du, p_dv, pitch, width, height, y0, y1, ghost):
        import torch
        try:
            dist, r, n = self.dist, self.rank, self.world
            self.calls[op] += 1
            with torch.cuda.stream(self.stream):
                fields = [as_tensor(p, height * pitch, self.device) for p in (p_du, p_dv)]
                if op == 0:
                    ops = []
                    for t in fields:
                        if r > 0:
                            ops.append(dist.P2POp(dist.isend, t[y0 * pitch:(y0 + ghost) * pitch], r - 1))
                            ops.append(dist.P2POp(dist.irecv, t[(y0 - ghost) * pitch:y0 * pitch], r - 1))
                        if r < n - 1:
                            ops.append(dist.P2POp(dist.isend, t[(y1 - ghost) * pitch:y1 * pitch], r + 1))
                            ops.append(dist.P2POp(dist.irecv, t[y1 * pitch:(y1 + ghost) * pitch], r + 1))
                    self.bytes[0] += sum(o.tensor.numel() * 4 for o in ops) // 2
                    for q in dist.batch_isend_irecv(ops):
                        q.wait()
                else:
                    for src in range(n):
                        a, b = slab_rows(height, src, n)
                        for t in fields:
                            dist.broadcast(t[a * pitch:b * pitch], src=src)
                    self.bytes[1] += 2 * height * pitch * 4
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("NcclExchange failed:", repr(e))
            return 1


class ThreadedExchange:
    """All ranks live in one process (one thread each) and on one GPU."""

    def __init__(self, world, device):
        self.world, self.device = world, device
        self.barrier = threading.Barrier(world)
        self.ptrs = [None] * world
        self.fns = [EXCHANGE_FN(self._make(r)) for r in range(world)]
        self.calls = {0: 0, 1: 0}

    def slab(self, rank, min_rows=0):
        return Slab(rank, self.world, self.fns[rank], None, min_rows)

    def _make(self, r):
        def call(user, op, p_du, p_dv, pitch, width, height, y0, y1, ghost):
            import torch
            try:
                n = self.world
                torch.cuda.synchronize(self.device)      # this rank's passes are done
                self.ptrs[r] = (p_du, p_dv)
                self.barrier.wait()                       # everybody's rows are final and registered
                mine = [as_tensor(p, height * pitch, self.device) for p in (p_du, p_dv)]
                if r == 0:
                    self.calls[op] += 1
                if op == 0:
                    for nb, dst, src in ((r - 1, (y0 - ghost, y0), (y0 - ghost, y0)), (r + 1, (y1, y1 + ghost), (y1, y1 + ghost))):
                        if 0 <= nb < n:
                            for t, p in zip(mine, self.ptrs[nb]):
                                other = as_tensor(p, height * pitch, self.device)
                                t[dst[0] * pitch:dst[1] * pitch].copy_(other[src[0] * pitch:src[1] * pitch])
                else:
                    for src in range(n):
                        if src == r:
                            continue
                        a, b = slab_rows(height, src, n)
                        for t, p in zip(mine, self.ptrs[src]):
                            other = as_tensor(p, height * pitch, self.device)
                            t[a * pitch:b * pitch].copy_(other[a * pitch:b * pitch])
                torch.cuda.synchronize(self.device)
                self.barrier.wait()                       # nobody overwrites rows a neighbour still reads
                return 0
            except Exception as e:
                print("ThreadedExchange failed:", repr(e))
                try:
                    self.barrier.abort()
                except Exception:
                    pass
                return 1
        return call
