"""Sharded pools: one process per GPU (torchrun), the member set range-sharded over the ranks.

`torch.distributed` is plumbing only (rendezvous, the barrier around the descriptor exchange);
all simulation traffic goes GPU-to-GPU over NVLink inside the tick kernel.
Every rank must issue the same Pool calls in the same order; rank 0 executes the host-side part.
"""
from __future__ import annotations

import os
import socket
import struct

from . import _lib
from .pool import GsimError, Pool

_POOL_SEQ = 0


def _exchange_fds(my_fds: list, rank: int, world: int, tag: str) -> dict:
    """Hand this rank's slice descriptors to every peer (SCM_RIGHTS over abstract unix sockets)."""
    import torch.distributed as dist
    name = lambda r: "\0gsim-%s-%d" % (tag, r)  # noqa: E731  (abstract namespace: no files)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(name(rank))
    srv.listen(world)
    dist.barrier()  # everyone is listening
    for peer in range(world):
        if peer == rank:
            continue
        c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        c.connect(name(peer))
        socket.send_fds(c, [struct.pack("I", rank)], list(my_fds))
        c.close()
    got = {}
    for _ in range(world - 1):
        conn, _ = srv.accept()
        msg, fds, _, _ = socket.recv_fds(conn, 4, 250)
        got[struct.unpack("I", msg)[0]] = list(fds)
        conn.close()
    srv.close()
    dist.barrier()
    return got


class ShardedPool(Pool):
    """A Pool whose members live on `world` GPUs.  Needs an initialised process group."""

    def __init__(self, cfg, lib=None):
        import ctypes as C

        import torch.distributed as dist
        global _POOL_SEQ
        if not dist.is_initialized():
            raise RuntimeError("ShardedPool needs torch.distributed (launch with torchrun)")
        rank, world = dist.get_rank(), dist.get_world_size()
        cfg.world_size, cfg.rank = world, rank
        super().__init__(cfg, lib)
        self.rank, self.world = rank, world
        n = C.c_size_t()
        self._ck(self.lib.gsim_shard_export_fds(self.h, None, 0, C.byref(n)))
        mine = (C.c_int * n.value)()
        self._ck(self.lib.gsim_shard_export_fds(self.h, mine, n.value, C.byref(n)))
        tag = "%s-%d" % (os.environ.get("MASTER_PORT", "0"), _POOL_SEQ)
        _POOL_SEQ += 1
        peers = _exchange_fds(list(mine), rank, world, tag)
        for peer, pfds in sorted(peers.items()):
            arr = (C.c_int * len(pfds))(*pfds)
            self._ck(self.lib.gsim_shard_attach(self.h, peer, arr, len(pfds)))
            for f in pfds:
                os.close(f)
        self._ck(self.lib.gsim_shard_ready(self.h))

    def close(self):
        if getattr(self, "h", None):
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()  # nobody unmaps while a peer may still touch the memory
            super().close()
