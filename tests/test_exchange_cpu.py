"""world_size-2 (and 3) gloo tests of the multi-GPU exchange plumbing on CPU: counts all-to-all,
offsets, in-place variable-size payload exchange.  The device kernels on either side of it are
covered by tests/test_gpu_parity.py::test_bucket_exchange_roundtrip."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, words, q):
    sys.path.insert(0, ROOT)
    import sdt_pkg
    sdt_pkg.load()
    from soapdenovo_trans_b200.exchange import exchange_records
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for rnd in range(3):
            rng = np.random.default_rng(100 * rnd + rank)
            cap = 64
            counts = rng.integers(0, cap + 1, size=world)
            if rnd == 1:
                counts[:] = 0                      # an empty round must work
            send = torch.zeros((world, cap, words), dtype=torch.int64)
            for d in range(world):
                # record payload encodes (src, dst, index) so the receiver can verify routing
                for i in range(counts[d]):
                    send[d, i, :] = rank * 1_000_000 + d * 10_000 + i
            recv = torch.full((world * cap, words), -1, dtype=torch.int64)
            total, rc = exchange_records(send, torch.from_numpy(counts), recv)
            # what every source sent to me
            allc = [None] * world
            dist.all_gather_object(allc, counts.tolist())
            want = [allc[src][rank] for src in range(world)]
            assert rc == want and total == sum(want)
            off = 0
            for src in range(world):
                seg = recv[off:off + want[src]]
                exp = torch.arange(want[src], dtype=torch.int64) + src * 1_000_000 + rank * 10_000
                assert torch.equal(seg, exp[:, None].expand(-1, words)), (rank, src)
                off += want[src]
            assert (recv[off:] == -1).all()
        # overflow of a bin is reported, not silently truncated
        try:
            exchange_records(torch.zeros((world, 4, words), dtype=torch.int64), torch.tensor([5] * world), torch.zeros((64, words), dtype=torch.int64))
            q.put((rank, "no overflow error"))
        except OverflowError:
            q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,words", [(2, 2), (3, 5)])
def test_exchange_records_gloo(world, words):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, words, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def _worker_runs(rank, world, port, words, q):
    sys.path.insert(0, ROOT)
    import sdt_pkg
    sdt_pkg.load()
    from soapdenovo_trans_b200.exchange import exchange_runs
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for rnd in range(3):
            rng = np.random.default_rng(7 * rnd + rank)
            counts = rng.integers(0, 50, size=world)
            if rnd == 1:
                counts[:] = 0                      # an epoch without records must work
            # the runs for rank 0, 1, ... back to back, as sdtgpu_skm_stage hands them out; a record encodes (src, dst, index)
            recs = [torch.full((int(counts[d]), words), 0, dtype=torch.int64) + (rank * 1_000_000 + d * 10_000) +
                    torch.arange(int(counts[d]), dtype=torch.int64)[:, None] for d in range(world)]
            send = [r.reshape(-1) for r in recs]
            got = {}

            def make_recv(n):
                got["buf"] = torch.full((n * words,), -1, dtype=torch.int64)
                return got["buf"]
            total, rc = exchange_runs(send, words, make_recv)
            allc = [None] * world
            dist.all_gather_object(allc, counts.tolist())
            want = [allc[src][rank] for src in range(world)]
            assert rc == want and total == sum(want)
            recv = got["buf"].reshape(-1, words)
            off = 0
            for src in range(world):                # received runs arrive in source-rank order
                exp = torch.arange(want[src], dtype=torch.int64) + src * 1_000_000 + rank * 10_000
                assert torch.equal(recv[off:off + want[src]], exp[:, None].expand(-1, words)), (rank, src)
                off += want[src]
            assert off == recv.shape[0]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,words", [(2, 4), (3, 6)])
def test_exchange_runs_gloo(world, words):
    """The super-k-mer exchange (SkmExchange.flush): counts all-to-all, then the runs in place."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_runs, args=(r, world, port, words, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
