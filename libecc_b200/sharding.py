"""Multi-GPU plumbing: the batch is embarrassingly parallel, so each rank takes a contiguous shard
and the only exchange is the gather of the fixed-size results: either a NCCL all-gather (gather_results; gloo in the
CPU tests) or, on the hot path, stores of the normalisation kernel into peer-mapped buffers (PeerGather).
No arithmetic here."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous index range [lo, hi) of rank `rank` among `world` ranks; sizes differ by at most one."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_results(local: torch.Tensor, n_total: int, item: int, group=None) -> torch.Tensor:
    """All-gather per-rank result records (`item` bytes each, contiguous shards from shard_bounds) into the full
    [n_total * item] tensor on every rank.  Even shards use one all_gather_into_tensor; ragged ones are padded to the
    largest shard and trimmed."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    counts = [hi - lo for lo, hi in sizes]
    assert local.numel() == counts[dist.get_rank(group)] * item
    if len(set(counts)) == 1:
        full = torch.empty(n_total * item, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous().view(-1), group=group)
        return full
    mx = max(counts) * item
    padded = torch.zeros(mx, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local.view(-1)
    buf = torch.empty(world * mx, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * mx: r * mx + counts[r] * item] for r in range(world)])


# ------------------------------------------------------------------------------------------------ peer-memory gather
#
# The result gather of the multi-GPU path without a collective kernel: every rank owns one
# IPC-exported region; the normalisation kernel (K4) of a source rank stores its results straight into its slot of
# each destination's region over NVLink and publishes an arrival counter; destinations acknowledge consumption so
# that a slot is never overwritten early.  This module only computes the layout and moves 64-byte handles around
# (torch.distributed); the stores, flags and waits are in libecc_b200.so (eccb200_prj_pt_mul_batch_dev_gather,
# eccb200_flag_wait, eccb200_flag_signal).

FLAG_BYTES = 4096          # [0, 2048): arrive[src] u32 counters; [2048, 4096): ack[dst] u32 counters
ACK_OFFSET = 2048


def gather_layout(world: int, n_items: int, plen: int, nbuf: int = 2) -> dict:
    """Byte layout of one rank's region: flags, then nbuf buffers of `world` slots; a slot holds one source rank's
    [n_items][2*plen] affine points followed by its [n_items] status bytes, padded to 256 bytes."""
    assert 1 <= world <= 512 and nbuf >= 1
    out_bytes = n_items * 2 * plen
    slot = (out_bytes + n_items + 255) // 256 * 256
    return {"flags": 0, "ack": ACK_OFFSET, "data": FLAG_BYTES, "slot_bytes": slot, "status_offset": out_bytes,
            "buf_bytes": world * slot, "total": FLAG_BYTES + nbuf * world * slot, "nbuf": nbuf, "world": world}


def slot_offset(layout: dict, buf: int, src: int) -> int:
    return layout["data"] + buf * layout["buf_bytes"] + src * layout["slot_bytes"]


def gather_destinations(world: int, mode: str):
    """Ranks that receive every rank's results: 'root' -> [0] (the final result gather of SURVEY.md §8e),
    'all' -> every rank (all-gather semantics)."""
    if mode == "root":
        return [0]
    if mode == "all":
        return list(range(world))
    raise ValueError(mode)


class _TorchStreams:
    """Stream / event plumbing of the copy-engine transport (torch.cuda); the CPU tests substitute a recorder."""

    def current(self, device):
        return torch.cuda.current_stream(device)

    def new(self, device):
        return torch.cuda.Stream(device=device)

    def event(self):
        return torch.cuda.Event()

    def empty(self, nbytes, device):
        return torch.empty(nbytes, dtype=torch.uint8, device=device)


class PeerGather:
    """Per-rank state of the peer-memory gather for fixed-size batches of `n_items` results."""

    def __init__(self, eng, rank: int, world: int, n_items: int, mode: str = "root", nbuf: int = 2, group=None,
                 transport: str = "fused", streams=None):
        """transport 'fused': the normalisation kernel stores into the destinations itself (step);
        'ce': the kernel writes locally and a copy-engine transfer on a side stream pushes the slot to the destinations
        while the NEXT batch computes (step_ce / drain)."""
        self.group = group
        self.transport = transport
        self.streams = streams or _TorchStreams()
        self.copy_stream = None
        self.copy_done = {}
        self.local = {}
        self.last_src = 0
        self.eng, self.rank, self.world, self.n, self.nbuf = eng, rank, world, n_items, nbuf
        self.layout = gather_layout(world, n_items, eng.plen, nbuf)
        self.dests = gather_destinations(world, mode)
        if len(self.dests) > 8:
            raise ValueError("at most 8 destinations")
        self.base, handle = eng.ipc_alloc(self.layout["total"])
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle, group=group)
        else:
            handles[0] = handle      # single rank: the gather degenerates to stores into the own region
        self.peer = {}                       # rank -> mapped base pointer of that rank's region
        for r in set(self.dests) | {self.rank}:
            self.peer[r] = self.base if r == rank else eng.ipc_open(handles[r])
        if rank in self.dests:               # a destination acknowledges into every source's region
            for r in range(world):
                if r not in self.peer:
                    self.peer[r] = eng.ipc_open(handles[r])
        self.step_no = 0
        self._dst = {}
        if world > 1:
            dist.barrier(group=group)        # every mapping exists before anybody stores

    def step(self, p_scalars: int, p_points, p_out: int, p_status: int, stream: int):
        """One batch: K1, (wait for the buffer's release), K4 with the fused gather; on a destination rank also the
        wait for every source's arrival and the release acknowledgement.  Returns the buffer index used."""
        c = self.step_no + 1
        b = self.step_no % self.nbuf
        L = self.layout
        if b not in self._dst:       # pointer lists per buffer, computed once
            out = [self.peer[d] + slot_offset(L, b, self.rank) for d in self.dests]
            self._dst[b] = (out, [p + L["status_offset"] for p in out],
                            [self.peer[d] + L["flags"] + 4 * self.rank for d in self.dests])
        dst_out, dst_st, dst_flag = self._dst[b]
        need_ack = c - self.nbuf                     # the step that last used this buffer must have been consumed
        if need_ack >= 1:
            wait_ptr, wait_cnt = self._ack_wait_ptr()
        else:
            wait_ptr, wait_cnt = None, 0
        self.eng.prj_pt_mul_batch_dev_gather(self.n, p_scalars, p_points, p_out, p_status, dst_out, dst_st, dst_flag, c,
                                             wait_ptr, wait_cnt, max(need_ack, 0), stream)
        if self.rank in self.dests:
            self.eng.flag_wait(self.base + L["flags"], self.world, c, stream)        # all sources' step c has landed
            # ... a consumer of buffer b would run here ...
            acks = [self.peer[r] + L["ack"] + 4 * self.rank for r in range(self.world)]
            for i in range(0, len(acks), 8):
                self.eng.flag_signal(acks[i:i + 8], c, stream)
        self.step_no += 1
        return b

    # ---- copy-engine transport ---------------------------------------------------------------------------------
    def step_ce(self, p_scalars: int, p_points, device=None):
        """One batch on torch's current stream C with the push on the side stream X:
          C: [wait until this buffer's previous push has left] K1, K4 -> the rank's own slot (a destination writes
             straight into its region; a pure source into a local buffer of the same layout)
          X: after K4: [wait for the destinations' release of the slot] DMA slot -> every other destination, then
             publish arrive[rank] = c there
          destination, on C: publish the own arrival, wait for every rank's PREVIOUS batch (its push overlapped this
             batch's kernels), release it.  drain() closes the pipeline.  Returns the buffer index."""
        c = self.step_no + 1
        b = self.step_no % self.nbuf
        L = self.layout
        used = self.n * 2 * self.eng.plen + self.n
        C = self.streams.current(device)
        if self.copy_stream is None:
            self.copy_stream = self.streams.new(device)
        X = self.copy_stream
        if self.rank in self.dests:
            src = self.base + slot_offset(L, b, self.rank)
        else:
            if b not in self.local:
                self.local[b] = self.streams.empty(L["slot_bytes"], device)
            src = self.local[b].data_ptr()
        if b in self.copy_done:
            C.wait_event(self.copy_done[b])          # the push that last read this buffer has completed
        self.eng.prj_pt_mul_batch_dev_raw(self.n, p_scalars, p_points, src, src + L["status_offset"], C.cuda_stream)
        others = [d for d in self.dests if d != self.rank]
        if others:
            ev = self.streams.event()
            ev.record(C)
            X.wait_event(ev)
            need_ack = c - self.nbuf
            wait_ptr, wait_cnt = self._ack_wait_ptr() if need_ack >= 1 else (None, 0)
            self.eng.push_results([self.peer[d] + slot_offset(L, b, self.rank) for d in others], src, used,
                                  [self.peer[d] + L["flags"] + 4 * self.rank for d in others], c, wait_ptr, wait_cnt,
                                  max(need_ack, 0), X.cuda_stream)
            done = self.streams.event()
            done.record(X)
            self.copy_done[b] = done
        if self.rank in self.dests:
            self.eng.flag_signal([self.base + L["flags"] + 4 * self.rank], c, C.cuda_stream)   # own slot is written
            if c >= 2:
                self._consume(c - 1, C.cuda_stream)
        self.last_src = src
        self.step_no += 1
        return b

    def _consume(self, upto: int, stream: int):
        """Destination: wait until every rank's batches up to `upto` have landed, then release them."""
        L = self.layout
        self.eng.flag_wait(self.base + L["flags"], self.world, upto, stream)
        # ... a consumer of buffer (upto - 1) % nbuf would run here ...
        acks = [self.peer[r] + L["ack"] + 4 * self.rank for r in range(self.world)]
        for i in range(0, len(acks), 8):
            self.eng.flag_signal(acks[i:i + 8], upto, stream)

    def drain(self, device=None):
        """Closes the pipeline of step_ce on the current stream: a destination waits for (and releases) the last
        batch of every rank; a source waits until its last push has left."""
        C = self.streams.current(device)
        if self.rank in self.dests and self.step_no >= 1:
            self._consume(self.step_no, C.cuda_stream)
        for ev in self.copy_done.values():
            C.wait_event(ev)

    def _ack_wait_ptr(self):
        d = sorted(self.dests)
        if d == list(range(d[0], d[0] + len(d))):   # contiguous ack counters: one wait kernel covers them
            return self.base + self.layout["ack"] + 4 * d[0], len(d)
        raise ValueError("destination ranks must be contiguous")

    def buffer_ptr(self, buf: int, src: int) -> int:
        """Address of source `src`'s slot of buffer `buf` in THIS rank's region (meaningful on destinations)."""
        return self.base + slot_offset(self.layout, buf, src)

    def close(self, group=None):
        """Collective: every rank unmaps the peers' regions, then (after a barrier) frees its own."""
        for r, p in self.peer.items():
            if r != self.rank:
                self.eng.ipc_close(p)
        self.peer = {}
        if self.world > 1:
            dist.barrier(group=group)
        if self.base:
            self.eng.ipc_free(self.base)
            self.base = 0
