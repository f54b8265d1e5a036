"""Host-side logic of the peer-memory result gather (libecc_b200/sharding.py PeerGather) on CPU: the region layout,
and — over two gloo ranks with a recording stand-in for the engine — the per-step protocol: which slot every rank
stores into, the arrival counters it publishes, the acknowledgements it waits for, and what the destination waits for
and acknowledges.  The stores / flags themselves are CUDA (tests/test_gpu_multi.py covers them on a GPU box)."""
import os
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT


def test_layout_slots_do_not_overlap_and_are_aligned():
    from libecc_b200.sharding import FLAG_BYTES, gather_layout, slot_offset
    for world in (1, 2, 8):
        for n in (1, 37, 1 << 20):
            for plen in (24, 32, 48, 66):
                for nbuf in (1, 2, 3):
                    L = gather_layout(world, n, plen, nbuf)
                    spans = []
                    for b in range(nbuf):
                        for r in range(world):
                            o = slot_offset(L, b, r)
                            assert o % 256 == 0 and o >= FLAG_BYTES
                            assert L["status_offset"] == n * 2 * plen
                            spans.append((o, o + n * 2 * plen + n))
                    spans.sort()
                    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
                    assert spans[-1][1] <= L["total"]
                    assert 4 * world <= 2048      # arrival and acknowledgement counters fit their halves


class FakeEngine:
    """Records the calls PeerGather makes; 'device pointers' are rank-tagged integers."""
    plen = 32

    def __init__(self, rank):
        self.rank = rank
        self.calls = []

    @staticmethod
    def base_of(rank):
        return (rank + 1) << 40

    def ipc_alloc(self, nbytes):
        return self.base_of(self.rank), bytes([self.rank]) * 64

    def ipc_open(self, handle):
        return self.base_of(handle[0])

    def ipc_close(self, ptr):
        self.calls.append(("close", ptr))

    def ipc_free(self, ptr):
        self.calls.append(("free", ptr))

    def prj_pt_mul_batch_dev_gather(self, n, p_sc, p_pts, p_out, p_st, dst_out, dst_status, dst_flag, flag_value,
                                    p_wait, wait_count, wait_value, stream):
        self.calls.append(("smul", n, tuple(dst_out), tuple(dst_status), tuple(dst_flag), flag_value, p_wait, wait_count,
                           wait_value))

    def prj_pt_mul_batch_dev_raw(self, n, p_sc, p_pts, p_out, p_st, stream=0):
        self.calls.append(("smul_raw", n, p_out, p_st, stream))

    def push_results(self, dst_ptrs, p_src, nbytes, dst_flags, flag_value, p_wait, wait_count, wait_value, stream=0):
        self.calls.append(("push", tuple(dst_ptrs), p_src, nbytes, tuple(dst_flags), flag_value, p_wait, wait_count,
                           wait_value, stream))

    def flag_wait(self, p_flags, count, value, stream=0):
        self.calls.append(("wait", p_flags, count, value) if stream == 0 else ("wait", p_flags, count, value, stream))

    def flag_signal(self, ptrs, value, stream=0):
        self.calls.append(("signal", tuple(ptrs), value) if stream == 0 else ("signal", tuple(ptrs), value, stream))


class FakeStream:
    def __init__(self, log, handle):
        self.log, self.cuda_stream = log, handle

    def wait_event(self, ev):
        self.log.append(("stream_wait", self.cuda_stream, ev.ident, ev.recorded_on))


class FakeEvent:
    count = 0

    def __init__(self, log):
        FakeEvent.count += 1
        self.log, self.ident, self.recorded_on = log, FakeEvent.count, None

    def record(self, stream):
        self.recorded_on = stream.cuda_stream
        self.log.append(("event_record", self.ident, stream.cuda_stream))


class FakeBuffer:
    def __init__(self, addr):
        self.addr = addr

    def data_ptr(self):
        return self.addr


class FakeStreams:
    """Stand-in for the torch.cuda plumbing of the copy-engine transport: stream 100 = compute, 200 = copy."""

    def __init__(self, log):
        self.log, self.nbuf = log, 0

    def current(self, device):
        return FakeStream(self.log, 100)

    def new(self, device):
        return FakeStream(self.log, 200)

    def event(self):
        return FakeEvent(self.log)

    def empty(self, nbytes, device):
        self.nbuf += 1
        return FakeBuffer((0x77 << 40) + self.nbuf * (1 << 32))


def _worker_ce(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from libecc_b200.sharding import PeerGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FakeEngine(rank)
    pg = PeerGather(eng, rank, world, 1000, mode=mode, nbuf=2, transport="ce", streams=FakeStreams(eng.calls))
    bufs = [pg.step_ce(111, None, None) for _ in range(5)]
    pg.drain(None)
    srcs = pg.last_src
    pg.close()
    q.put((rank, bufs, eng.calls, pg.layout, srcs))
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from libecc_b200.sharding import PeerGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FakeEngine(rank)
    pg = PeerGather(eng, rank, world, 1000, mode=mode, nbuf=2)
    bufs = [pg.step(111, None, 222, 333, 0) for _ in range(5)]
    pg.close()
    q.put((rank, bufs, eng.calls, pg.layout))
    dist.barrier()
    dist.destroy_process_group()


def _run(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, bufs, calls, layout = q.get(timeout=300)
        got[rank] = (bufs, calls, layout)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def _check(mode):
    from libecc_b200.sharding import ACK_OFFSET, slot_offset
    got = _run(mode)
    dests = [0] if mode == "root" else [0, 1]
    for rank, (bufs, calls, L) in got.items():
        assert bufs == [0, 1, 0, 1, 0]
        smul = [c for c in calls if c[0] == "smul"]
        assert len(smul) == 5
        for s, c in enumerate(smul):
            _, n, dst_out, dst_st, dst_flag, flag_value, p_wait, wait_count, wait_value = c
            assert n == 1000 and flag_value == s + 1
            # this rank's slot of buffer s % 2 in every destination's region, arrival counter arrive[rank] there
            assert dst_out == tuple(FakeEngine.base_of(d) + slot_offset(L, s % 2, rank) for d in dests)
            assert dst_st == tuple(p + 1000 * 64 for p in dst_out)
            assert dst_flag == tuple(FakeEngine.base_of(d) + 4 * rank for d in dests)
            # the step that last used the buffer (s + 1 - nbuf) must have been acknowledged by every destination
            if s + 1 - 2 >= 1:
                assert p_wait == FakeEngine.base_of(rank) + ACK_OFFSET and wait_count == len(dests)
                assert wait_value == s + 1 - 2
            else:
                assert wait_count == 0
        waits = [c for c in calls if c[0] == "wait"]
        sigs = [c for c in calls if c[0] == "signal"]
        if rank in dests:
            # a destination waits for all sources' arrival of step c, then acknowledges into every source's region
            assert waits == [("wait", FakeEngine.base_of(rank), 2, s + 1) for s in range(5)]
            assert sigs == [("signal", tuple(FakeEngine.base_of(r) + ACK_OFFSET + 4 * rank for r in range(2)), s + 1)
                            for s in range(5)]
        else:
            assert not waits and not sigs
        # teardown order: peers unmapped before the own region is freed
        kinds = [c[0] for c in calls if c[0] in ("close", "free")]
        assert kinds == ["close"] * (len(kinds) - 1) + ["free"]


def test_protocol_root_gather():
    _check("root")


def test_protocol_all_gather():
    _check("all")


def test_protocol_copy_engine_root_gather():
    """The pipelined transport: sources compute into a local slot and push it on the copy stream behind an event, with the
    acknowledgement wait in front of the push from the third step on; the root computes straight into its own slot,
    publishes its own arrival, consumes every rank's PREVIOUS step and drains the last one."""
    from libecc_b200.sharding import ACK_OFFSET, slot_offset
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_ce, args=(r, 2, port, "root", q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, bufs, calls, layout, last_src = q.get(timeout=300)
        got[rank] = (bufs, calls, layout, last_src)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    used = 1000 * 64 + 1000
    # ---- the source rank
    bufs, calls, L, last_src = got[1]
    assert bufs == [0, 1, 0, 1, 0]
    smul = [c for c in calls if c[0] == "smul_raw"]
    push = [c for c in calls if c[0] == "push"]
    assert len(smul) == 5 and len(push) == 5
    local = [smul[0][2], smul[1][2]]
    assert local[0] != local[1] and [c[2] for c in smul] == [local[0], local[1], local[0], local[1], local[0]]
    assert all(c[3] == c[2] + 1000 * 64 and c[4] == 100 for c in smul)          # status behind the points, compute stream
    for s, c in enumerate(push):
        _, dst, src, nbytes, flags, value, p_wait, wait_count, wait_value, stream = c
        assert dst == (FakeEngine.base_of(0) + slot_offset(L, s % 2, 1),) and src == local[s % 2] and nbytes == used
        assert flags == (FakeEngine.base_of(0) + 4 * 1,) and value == s + 1 and stream == 200
        if s + 1 - 2 >= 1:
            assert p_wait == FakeEngine.base_of(1) + ACK_OFFSET and wait_count == 1 and wait_value == s + 1 - 2
        else:
            assert wait_count == 0
    # every push is ordered behind its step's kernels (event recorded on 100, waited on 200), and a buffer is not
    # recomputed before its previous push has left (event recorded on 200, waited on 100)
    waits = [c for c in calls if c[0] == "stream_wait"]
    assert sum(1 for c in waits if c[1] == 200 and c[3] == 100) == 5
    assert sum(1 for c in waits if c[1] == 100 and c[3] == 200) >= 3
    assert not [c for c in calls if c[0] in ("wait", "signal")]                 # a pure source neither consumes nor acks
    # ---- the root
    bufs, calls, L, last_src = got[0]
    smul = [c for c in calls if c[0] == "smul_raw"]
    assert [c[2] for c in smul] == [FakeEngine.base_of(0) + slot_offset(L, s % 2, 0) for s in range(5)]
    assert not [c for c in calls if c[0] == "push"]
    sig = [c for c in calls if c[0] == "signal"]
    own = [c for c in sig if c[1] == (FakeEngine.base_of(0) + 0,)]
    assert [c[2] for c in own] == [1, 2, 3, 4, 5]                               # its own arrival counter
    acks = [c for c in sig if len(c[1]) == 2]
    assert [c[2] for c in acks] == [1, 2, 3, 4, 5]                              # steps 1-4 consumed one step late, 5 by drain
    assert all(c[1] == tuple(FakeEngine.base_of(r) + ACK_OFFSET for r in range(2)) for c in acks)
    w = [c for c in calls if c[0] == "wait"]
    assert [(c[1], c[2], c[3]) for c in w] == [(FakeEngine.base_of(0), 2, v) for v in (1, 2, 3, 4, 5)]
