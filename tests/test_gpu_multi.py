"""Multi-GPU entry points on whatever the box has (one GPU is enough for the code paths):
  - the result gather fused into the normalisation kernel (eccb200_prj_pt_mul_batch_dev_gather): destination buffers
    in an IPC-exportable allocation, arrival flag, acknowledgement wait — bit-exact against the plain call and the oracle;
  - the single-process multi-device C ABI (eccb200_multi_*) through a plain-C host judged by the unmodified reference
    (tests/dropin/multi_harness.c), over every visible GPU and over two contexts on GPU 0 (ragged shards);
  - with two or more GPUs: two processes, peer-mapped buffers over NVLink (PeerGather) against a NCCL all-gather."""
import os
import subprocess
import sys

import numpy as np
import pytest

from common import ALL_CURVES, ROOT, edge_scalars, oracle_smul, random_scalars

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "oracle", "_ref", "multi_harness")


@pytest.mark.parametrize("curve", ["SECP256R1", "SECP384R1", "SECP521R1"])
def test_fused_gather_on_one_gpu(curve):
    import torch
    import libecc_b200
    _, plen, qlen = ALL_CURVES[curve]
    eng = libecc_b200.Engine(curve, device=0, comb_window=12)
    sc = np.concatenate([random_scalars(curve, 3000, tag=71, below_q=False), edge_scalars(curve)])
    n = sc.shape[0]
    want, wst = oracle_smul(curve, sc)
    dev = torch.device("cuda", 0)
    d_sc = torch.from_numpy(sc).to(dev)
    d_out = torch.zeros(n * 2 * plen, dtype=torch.uint8, device=dev)
    d_st = torch.zeros(n, dtype=torch.int8, device=dev)
    slot = (n * 2 * plen + n + 255) // 256 * 256
    base, handle = eng.ipc_alloc(4096 + 2 * slot)          # flags, then two destination slots
    assert len(handle) == 64 and base % 256 == 0
    dst = [base + 4096, base + 4096 + slot]
    stream = torch.cuda.current_stream().cuda_stream
    for step in (1, 2, 3):
        # acknowledgement counter at base + 2048 is published first, so that the kernel's wait is exercised
        eng.flag_signal([base + 2048], step - 1, stream)
        eng.prj_pt_mul_batch_dev_gather(n, d_sc.data_ptr(), None, d_out.data_ptr(), d_st.data_ptr(), dst,
                                        [p + n * 2 * plen for p in dst], [base, base + 4], step, base + 2048, 1, step - 1,
                                        stream)
        eng.flag_wait(base, 2, step, stream)
        torch.cuda.synchronize()
        flags = eng.copy_to_host(base, 8).view(np.uint32)
        assert list(flags) == [step, step]
        got, gst = d_out.cpu().numpy().reshape(n, 2 * plen), d_st.cpu().numpy()
        assert (gst == wst).all() and (got == want).all()
        for p in dst:
            raw = eng.copy_to_host(p, n * 2 * plen + n)
            assert (raw[: n * 2 * plen].reshape(n, 2 * plen) == want).all()
            assert (raw[n * 2 * plen:].view(np.int8) == wst).all()
    # a misaligned destination is refused, not dereferenced
    if plen % 16 == 0:
        with pytest.raises(libecc_b200.EccB200Error, match="aligned"):
            eng.prj_pt_mul_batch_dev_gather(n, d_sc.data_ptr(), None, d_out.data_ptr(), d_st.data_ptr(), [dst[0] + 4],
                                            [dst[0] + n * 2 * plen], [base], 9, None, 0, 0, stream)
        with pytest.raises(libecc_b200.EccB200Error, match="aligned"):
            eng.prj_pt_mul_batch_dev(d_sc.view(-1)[16 - 4:][: (n - 1) * qlen], None, d_out, d_st, stream)
    eng.ipc_free(base)
    eng.close()


def test_copy_engine_push_on_one_gpu():
    """eccb200_push_results with this GPU as its own destination: DMA of a result slot into two buffers, arrival flags
    published behind the copies, acknowledgement wait in front of them."""
    import torch
    import libecc_b200
    eng = libecc_b200.Engine("SECP256R1", device=0, comb_window=10)
    dev = torch.device("cuda", 0)
    nbytes = 3 * 65 * 1000 + 7
    src = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device=dev)
    base, _ = eng.ipc_alloc(4096 + 2 * ((nbytes + 255) // 256 * 256))
    d0, d1 = base + 4096, base + 4096 + (nbytes + 255) // 256 * 256
    s = torch.cuda.Stream(device=dev)
    eng.flag_signal([base + 2048], 5, s.cuda_stream)                       # the "destination" has released step 5
    eng.push_results([d0, d1], src.data_ptr(), nbytes, [base, base + 4], 6, base + 2048, 1, 5, s.cuda_stream)
    eng.flag_wait(base, 2, 6, s.cuda_stream)
    s.synchronize()
    want = src.cpu().numpy()
    assert (eng.copy_to_host(d0, nbytes) == want).all() and (eng.copy_to_host(d1, nbytes) == want).all()
    assert list(eng.copy_to_host(base, 8).view(np.uint32)) == [6, 6]
    eng.ipc_free(base)
    eng.close()


def test_dev_calls_on_two_streams_share_the_scratch_safely():
    """two *_dev calls on different streams used to race on the context's scratch buffers."""
    import torch
    import libecc_b200
    curve = "SECP256R1"
    eng = libecc_b200.Engine(curve, device=0, comb_window=12)
    dev = torch.device("cuda", 0)
    n = 1 << 15
    sa, sb = random_scalars(curve, n, tag=81, below_q=False), random_scalars(curve, n, tag=82, below_q=False)
    wa, _ = oracle_smul(curve, sa[:200]); wb, _ = oracle_smul(curve, sb[:200])
    da, db = torch.from_numpy(sa).to(dev), torch.from_numpy(sb).to(dev)
    oa = torch.zeros(n * 64, dtype=torch.uint8, device=dev); ob = torch.zeros_like(oa)
    ta = torch.zeros(n, dtype=torch.int8, device=dev); tb = torch.zeros_like(ta)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    for _ in range(4):
        eng.prj_pt_mul_batch_dev(da.view(-1), None, oa, ta, s1.cuda_stream)
        eng.prj_pt_mul_batch_dev(db.view(-1), None, ob, tb, s2.cuda_stream)
    torch.cuda.synchronize()
    assert (oa.cpu().numpy().reshape(n, 64)[:200] == wa).all()
    assert (ob.cpu().numpy().reshape(n, 64)[:200] == wb).all()
    # the whole batches agree with a serial recomputation
    oa2 = torch.zeros_like(oa); eng.prj_pt_mul_batch_dev(da.view(-1), None, oa2, ta, 0); torch.cuda.synchronize()
    ob2 = torch.zeros_like(ob); eng.prj_pt_mul_batch_dev(db.view(-1), None, ob2, tb, 0); torch.cuda.synchronize()
    assert torch.equal(oa, oa2) and torch.equal(ob, ob2)
    eng.close()


def _run_harness(devlist, items):
    if not os.path.exists(HARNESS):
        pytest.fail("oracle/_ref/multi_harness is missing: run `make -C oracle all` where /root/reference exists")
    import libecc_b200
    r = subprocess.run([HARNESS, libecc_b200.LIB_PATH, devlist, str(items)], capture_output=True, text=True, timeout=1500)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HARNESS OK" in r.stdout


def test_multi_device_c_abi_two_contexts_on_one_gpu():
    _run_harness("0,0,0", 100003)     # three shards of different sizes on GPU 0


def test_multi_device_c_abi_all_gpus():
    import torch
    k = torch.cuda.device_count()
    _run_harness(",".join(str(i) for i in range(k)), 400007)


def test_multi_engine_python_matches_single():
    import libecc_b200
    curve = "FRP256V1"
    sc = np.concatenate([random_scalars(curve, 70000, tag=91, below_q=False), edge_scalars(curve)])
    eng = libecc_b200.Engine(curve, device=0, comb_window=12)
    want, wst = eng.prj_pt_mul_batch(sc)
    eng.close()
    me = libecc_b200.MultiEngine(curve, devices=[0, 0], comb_window=10)
    assert me.device_count == 2
    got, gst = me.prj_pt_mul_batch(sc)
    me.close()
    assert (got == want).all() and (gst == wst).all()


_PEER_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, ROOT_); sys.path.insert(0, os.path.join(ROOT_, "tests"))
import torch, torch.distributed as dist
import libecc_b200
from libecc_b200.sharding import PeerGather
from common import oracle_smul, random_scalars
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
curve, n = "SECP256R1", 50000
for mode, transport in (("root", "fused"), ("all", "fused"), ("root", "ce"), ("all", "ce")):
    eng = libecc_b200.Engine(curve, device=rank, comb_window=12)
    pg = PeerGather(eng, rank, world, n, mode=mode, transport=transport)
    d_out = torch.zeros(n * 64, dtype=torch.uint8, device=dev); d_st = torch.zeros(n, dtype=torch.int8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    keep = []
    for step in range(5):
        sc = random_scalars(curve, n, tag=1000 * step + rank, below_q=False)
        d_sc = torch.from_numpy(sc).to(dev)
        keep.append(d_sc)                      # the pipelined transport runs behind the host
        if transport == "ce":
            b = pg.step_ce(d_sc.data_ptr(), None, dev)
        else:
            b = pg.step(d_sc.data_ptr(), None, d_out.data_ptr(), d_st.data_ptr(), stream)
            torch.cuda.synchronize()
    if transport == "ce":
        pg.drain(dev)
        torch.cuda.synchronize()
        mine = torch.from_numpy(eng.copy_to_host(pg.last_src, n * 65)).to(dev)
    else:
        mine = torch.cat([d_out, d_st.view(torch.uint8)])
    ref = torch.empty(world * mine.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(ref, mine)
    torch.cuda.synchronize()
    if rank in pg.dests:
        got = np.concatenate([eng.copy_to_host(pg.buffer_ptr(b, r), mine.numel()) for r in range(world)])
        assert (got == ref.cpu().numpy()).all(), f"{mode}/{transport}: gathered buffer differs from the NCCL all-gather"
        other = (rank + 1) % world
        want, wst = oracle_smul(curve, random_scalars(curve, n, tag=4000 + other, below_q=False)[:64])
        assert (got.reshape(world, -1)[other][: 64 * 64].reshape(64, 64) == want).all()
    pg.close()
    eng.close()
dist.barrier()
dist.destroy_process_group()
print("PEER OK", rank)
'''


def test_peer_gather_two_processes():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (bench.py --gpus N exercises the same path at round end)")
    code = "ROOT_ = %r\n" % ROOT + _PEER_WORKER
    path = os.path.join(ROOT, "gpurun_out", "_peer_worker.py")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").write(code)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29711", path], capture_output=True, text=True,
                       timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and r.stdout.count("PEER OK") == 2
