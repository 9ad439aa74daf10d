"""Two-GPU parity (skipped on a single-GPU box): each rank chops its shard of the reads, k-mers
are exchanged by owner over NCCL, and the UNION of the ranks' tables must be the oracle's multiset;
the merged nodes, replayed by sdtgpu_build_kmersets, must give the reference's exact (set, slot)
layout — i.e. sharding across GPUs is invisible in the hand-back."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, K, kw, L, d, out_dir, mode):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import sdt_pkg
    pkg = sdt_pkg.load()
    from soapdenovo_trans_b200.exchange import Exchange, ReplicatedReads, SkmExchange
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    synth = pkg.synth
    tr = synth.make_transcriptome(40, 5)
    reads, lens = synth.make_reads(tr, 6000, L, 77, ragged=30)
    n = len(reads)
    lo, hi = rank * n // world, (rank + 1) * n // world          # contiguous shard, global read ordinals kept
    stride = synth.stride_bytes(L)
    d_packed = torch.from_numpy(synth.pack_reads(reads[lo:hi], lens[lo:hi], stride)).to(dev)
    d_lens = torch.from_numpy(lens[lo:hi].astype(np.int32)).to(dev)
    g = pkg.PregraphGPU(K, kw, L, capacity_hint=1_500_000, device=rank, sliced=mode.startswith("skm"))
    if mode.startswith("skm"):      # skm_native: the exchange inside the library (sdtgpu_skm_exchange), NCCL bound by libsdtgpu.so itself
        ex = SkmExchange(pkg, g, world, rank, dev, native=(mode == "skm_native"))
    elif mode == "records":
        ex = Exchange(pkg, g, world, rank, dev, max_round_instances=2048 * (L - K + 1))
    else:
        ex = ReplicatedReads(pkg, g, world, rank, dev, max_round_reads=2048, stride=stride)
    for a in range(0, hi - lo, 2048):                            # several rounds, exercises the double buffering
        b = min(a + 2048, hi - lo)
        ex.round(g, d_packed[a:b], b - a, 0, stride, lo + a, d_lens=d_lens[a:b])
    ex.flush(g)
    g.sync()
    freq, st = g.finalize(d)
    nodes = g.export_nodes(8)
    np.save(os.path.join(out_dir, f"nodes{rank}.npy"), nodes)
    np.save(os.path.join(out_dir, f"freq{rank}.npy"), freq)
    np.save(os.path.join(out_dir, f"stats{rank}.npy"), np.array([st.n_instances, st.n_nodes, st.n_removed, st.n_linear], dtype=np.int64))
    g.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["reads", "records", "skm", "skm_native"])
@pytest.mark.parametrize("K,kw,L,d", [(25, 1, 100, 0), (63, 4, 100, 1), (127, 4, 150, 0)])
def test_two_gpu_union_matches_oracle(pkg, oracle, tmp_path, K, kw, L, d, mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), K, kw, L, d, str(tmp_path), mode), nprocs=world, join=True)
    synth = pkg.synth
    reads, lens = synth.make_reads(synth.make_transcriptome(40, 5), 6000, L, 77, ragged=30)
    ref = oracle.run_hashing(reads, lens, K, kw, 8, d)
    nodes = np.concatenate([np.load(tmp_path / f"nodes{r}.npy") for r in range(world)])
    stats = sum(np.load(tmp_path / f"stats{r}.npy") for r in range(world))
    freq = sum(np.load(tmp_path / f"freq{r}.npy") for r in range(world))
    assert stats.tolist() == [ref.instances, ref.nodes, ref.removed, ref.linear]
    assert np.array_equal(freq, ref.kmerfreq)
    assert np.array_equal(pkg.nodes_to_records(nodes), oracle.sorted_multiset(ref.records))   # owners are disjoint
    sets = (C.POINTER(pkg.pregraph.KmerSet) * 8)()
    nodes = np.ascontiguousarray(nodes)
    assert pkg.library().sdtgpu_build_kmersets(nodes.ctypes.data, len(nodes), kw, 8, None, sets) == 0
    rec, info = pkg.read_kmersets(sets, 8, kw)
    pkg.library().sdtgpu_free_kmersets(sets, 8)
    assert np.array_equal(info, ref.set_info) and np.array_equal(rec, ref.records)
