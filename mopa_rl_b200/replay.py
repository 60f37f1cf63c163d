"""Replicated replay of SMDP transitions across GPUs.

The reference keeps one Python list-of-dicts buffer per MPI rank and never exchanges experience
(rl/dataset.py:7-37, rl/main.py:24-29).  Here every rank (one per GPU) collects transitions from
its own env shard and, once per tick, all ranks exchange the new fixed-width records
(92 fp32 = 368 B each: ob 40, ac 8, rew, done, intra_steps, env id, ob_next 40) so that every GPU
holds the same replay ring and SAC can sample locally.

Per tick (``exchange``):
  1. ``mopa_rollout_pack`` (main stream): the records emitted since the previous tick - main and
     relabelled alike, at most ``slab_capacity`` of them, the rest stays queued in the runner's ring -
     are copied compact behind a one-row header (word 0 = count) into a send block.
  2. one NCCL all-gather of the ``[1 + slab_capacity, 92]`` blocks on a high-priority side stream,
     under the env-step kernel of the same tick (double-buffered: block k is reused two ticks later);
  3. ``mopa_replay_append`` (one kernel, side stream): the gathered blocks are appended rank-major -
     the same order on every rank - to the replicated ring.
There is no host synchronisation and no padding proportional to the number of environments: a burst
(every env finishing a 13-step plan in the same tick emits ~3 n relabelled records) drains over the
following ticks.  Records are never dropped; a queue that outgrows half the runner's ring raises.
"""
from __future__ import annotations

import ctypes as C

TRANSITION_FLOATS = 92


class ReplicatedReplay:
    def __init__(self, torch, device, capacity=1 << 20, slab_capacity=4096, group=None, overlap=True):
        import torch.distributed as dist

        self.torch, self.device, self.group = torch, device, group
        self.capacity, self.slab_capacity = int(capacity), int(slab_capacity)
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.ring = torch.zeros(self.capacity, TRANSITION_FLOATS, dtype=torch.float32, device=device)
        self._size2 = torch.zeros(2, dtype=torch.int64, device=device)   # running record count, ping-pong (see mopa_replay_append)
        self._parity = 0
        rows = 1 + self.slab_capacity
        self._send = [torch.zeros(rows, TRANSITION_FLOATS, dtype=torch.float32, device=device) for _ in range(2)]
        self._recv = [torch.zeros(self.world * rows, TRANSITION_FLOATS, dtype=torch.float32, device=device) for _ in range(2)] if self.world > 1 else None
        self._k = 0
        self.bytes_exchanged = 0          # bytes landed on this rank through the all-gather
        self._init_streams(overlap)

    # ---- device plumbing (overridden by the gloo/CPU test double in tests/test_replay_dist.py)
    def _init_streams(self, overlap):
        torch = self.torch
        if self.device.type != "cuda":
            raise RuntimeError("ReplicatedReplay runs on a CUDA device (libmopa_b200 kernels); there is no CPU path")
        from .capi import lib

        self._L = lib()
        self._L.mopa_replay_append.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
        self._side = torch.cuda.Stream(device=self.device, priority=-1) if overlap else None
        self._done = [torch.cuda.Event(), torch.cuda.Event()]
        self._packed = torch.cuda.Event()

    def _pack(self, runner, send):
        runner.pack(send, self.slab_capacity)

    def _append(self, blocks):
        from .capi import check

        stream = self.torch.cuda.current_stream(self.device)
        check(self._L.mopa_replay_append(blocks.data_ptr(), self.world, self.slab_capacity, self.ring.data_ptr(), self.capacity,
                                         self._size2.data_ptr(), self._parity, C.c_void_p(stream.cuda_stream)))

    def _wait_block(self, b):
        """Main stream: block b (two ticks old) has been gathered and appended."""
        self.torch.cuda.current_stream(self.device).wait_event(self._done[b])

    def _side_stream(self):
        """Context in which the all-gather and the append are enqueued: the side stream, ordered after the pack."""
        torch = self.torch
        main = torch.cuda.current_stream(self.device)
        side = self._side or main
        if side is not main:
            self._packed.record(main)
            side.wait_event(self._packed)
        return torch.cuda.stream(side)

    def _mark_done(self, b):
        self._done[b].record(self.torch.cuda.current_stream(self.device))

    # ---- the per-tick exchange
    def exchange(self, runner):
        """Add this rank's new records and everybody else's.  Collective when world_size > 1; enqueues only."""
        import torch.distributed as dist

        b = self._k & 1
        if self._k >= 2:
            self._wait_block(b)
        send = self._send[b]
        self._pack(runner, send)
        self.last_block = send               # this tick's records of this rank (header row + records), e.g. for a host-side consumer
        with self._side_stream():
            blocks = send
            if self.world > 1:
                blocks = self._recv[b]
                dist.all_gather_into_tensor(blocks, send, group=self.group)
                self.bytes_exchanged += blocks.numel() * 4
            self._append(blocks)
            self._mark_done(b)
        self._parity ^= 1
        self._k += 1

    def sync(self):
        """Make the current stream wait for the exchanges enqueued so far (before sampling / reading the ring)."""
        if self._k:
            self._wait_block((self._k - 1) & 1)

    def device_size(self):
        """Records stored so far (reads the device counter)."""
        self.sync()
        return int(self._size2[self._parity].item())

    def sample(self, batch_size, generator=None):
        n = min(self.device_size(), self.capacity)
        idx = self.torch.randint(0, n, (batch_size,), device=self.device, generator=generator)
        return self.ring[idx]
