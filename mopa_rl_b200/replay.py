"""Replicated replay of SMDP transitions across GPUs.

The reference keeps one Python list-of-dicts buffer per MPI rank and never exchanges experience
(rl/dataset.py:7-37, rl/main.py:24-29).  Here every rank (one per GPU) collects transitions from
its own env shard and, once per tick, all ranks exchange the new fixed-width records
(92 fp32 = 368 B each) with one all-gather over NCCL/NVLink, so that every GPU holds the same
replay ring and SAC can sample locally.  Counts differ per rank and tick: counts are gathered
first, then slabs padded to the tick maximum.
"""
from __future__ import annotations

TRANSITION_FLOATS = 92


def pack_counts_and_slab(torch, records, capacity, device):
    """[k, 92] (or None) -> (count tensor [1] int64, slab [capacity, 92]) padded with zeros."""
    k = 0 if records is None else int(records.shape[0])
    if k > capacity:
        raise ValueError("tick emitted %d transitions, slab capacity is %d" % (k, capacity))
    slab = torch.zeros(capacity, TRANSITION_FLOATS, dtype=torch.float32, device=device)
    if k:
        slab[:k] = records
    return torch.tensor([k], dtype=torch.int64, device=device), slab


class ReplicatedReplay:
    def __init__(self, torch, device, capacity=1 << 20, slab_capacity=8192, group=None):
        self.torch, self.device, self.group = torch, device, group
        self.ring = torch.zeros(capacity, TRANSITION_FLOATS, dtype=torch.float32, device=device)
        self.capacity, self.slab_capacity = capacity, slab_capacity
        self.size = 0          # total records ever stored (write pointer = size % capacity)
        self.bytes_exchanged = 0

    def _append(self, rows):
        k = int(rows.shape[0])
        if not k:
            return
        w0 = self.size % self.capacity
        k1 = min(k, self.capacity - w0)
        self.ring[w0:w0 + k1] = rows[:k1]
        if k1 < k:
            self.ring[:k - k1] = rows[k1:]
        self.size += k

    def exchange(self, records):
        """Add this rank's new records and everybody else's.  Collective when world_size > 1."""
        torch = self.torch
        import torch.distributed as dist

        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            if records is not None:
                self._append(records)
            return
        count, slab = pack_counts_and_slab(torch, records, self.slab_capacity, self.device)
        clist = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(world)]
        dist.all_gather(clist, count, group=self.group)
        counts = torch.cat(clist)
        kmax = int(counts.max())
        if kmax == 0:
            return
        glist = [torch.zeros(kmax, TRANSITION_FLOATS, dtype=torch.float32, device=self.device) for _ in range(world)]
        dist.all_gather(glist, slab[:kmax].contiguous(), group=self.group)
        self.bytes_exchanged += world * kmax * TRANSITION_FLOATS * 4
        g = torch.stack(glist)
        valid = torch.arange(kmax, device=self.device)[None, :] < counts[:, None]
        self._append(g[valid])   # rank-major order: identical on every rank

    def exchange_slab(self, slab, flags):
        """Fixed-shape variant for the native runner: ``slab`` [n, 92] holds this tick's records dense by
        environment, ``flags`` [n] (uint8) marks the rows that are records.  No host round trip: the
        all-gather moves whole slabs (NVLink makes the padding irrelevant) and the valid rows are
        appended on the device, rank-major, identically on every rank."""
        torch = self.torch
        import torch.distributed as dist

        world = dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            n = slab.shape[0]
            g = torch.empty(world * n, TRANSITION_FLOATS, dtype=torch.float32, device=self.device)
            f = torch.empty(world * n, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(g, slab.contiguous(), group=self.group)
            dist.all_gather_into_tensor(f, flags.contiguous(), group=self.group)
            self.bytes_exchanged += world * n * (TRANSITION_FLOATS * 4 + 1)
            slab, flags = g, f
        if not hasattr(self, "_size_dev"):
            self._size_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
            self._ring1 = torch.zeros(self.capacity + 1, TRANSITION_FLOATS, dtype=torch.float32, device=self.device)  # last row: dump
            self.ring = self._ring1[:self.capacity]
        live = flags.bool()
        pos = torch.cumsum(live.to(torch.int64), 0) - 1
        idx = torch.where(live, (self._size_dev + pos) % self.capacity, torch.full_like(pos, self.capacity))
        self._ring1.index_copy_(0, idx, slab)
        self._size_dev += live.sum()

    def device_size(self):
        """Records stored through exchange_slab (reads the device counter)."""
        return int(self._size_dev.item()) if hasattr(self, "_size_dev") else self.size

    def sample(self, batch_size, generator=None):
        n = min(self.device_size(), self.capacity)
        idx = self.torch.randint(0, n, (batch_size,), device=self.device, generator=generator)
        return self.ring[idx]
