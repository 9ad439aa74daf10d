"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run
3.5e9 instances in a test): conservation of instances, and an order-independent fingerprint of the
table (sdtgpu_table_checksum) that must not depend on batching, capacity, insert path (single-pass insert, sliced
build with and without a capacity hint) or rounds —
anchored to the oracle by (a) fingerprint equality on small inputs for every key width and (b) a
sub-sample of the full-size device-generated reads checked against the oracle bit for bit."""
import numpy as np
import pytest

from conftest import make_dataset

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,kw,L", [(25, 1, 100), (31, 1, 100), (63, 2, 100), (63, 4, 100), (127, 4, 150)])
def test_fingerprint_matches_oracle_on_small_inputs(pkg, oracle, tiny_transcriptome, K, kw, L):
    reads, lens = make_dataset(pkg, tiny_transcriptome, 3000, L, 3 + K, ragged=20)
    ref = oracle.run_hashing(reads, lens, K, kw, 8, 0)
    stride = pkg.synth.stride_bytes(L)
    with pkg.PregraphGPU(K, kw, L, capacity_hint=500_000) as g:
        g.push_reads(pkg.synth.pack_reads(reads, lens, stride), lens, None, n_reads=len(reads), stride_bytes=stride)
        got = g.table_checksum()
        w = g.stats().device_key_words
    assert np.array_equal(got, oracle.table_checksum(ref.records, w))


def _device_reads(pkg, cfg, n_pairs, dev):
    import torch
    synth = pkg.synth
    tr = synth.make_transcriptome(cfg["n_transcripts"], cfg["seed"], hot=cfg["hot"])
    tr_dev = dict(bases=torch.from_numpy(tr.bases).to(dev), starts=torch.from_numpy(tr.starts.astype(np.int64)).to(dev),
                  lengths=torch.from_numpy(tr.lengths.astype(np.int32)).to(dev),
                  cum=torch.from_numpy(tr.cum.astype(np.int64)).to(dev), n=len(tr.lengths))
    stride = synth.stride_bytes(cfg["read_len"])
    d = torch.empty((2 * n_pairs, stride), dtype=torch.uint8, device=dev)
    pkg.pregraph.synth_reads_device(tr_dev, cfg["seed"], 0, n_pairs, cfg["read_len"], stride, d)
    torch.cuda.synchronize()
    return tr, d, stride


def _insert_all(pkg, g, d_packed, L, stride, batch):
    n = d_packed.shape[0]
    for a in range(0, n, batch):
        b = min(a + batch, n)
        g.push_reads(d_packed[a:b], None, None, n_reads=b - a, uniform_len=L, stride_bytes=stride, first_read_ordinal=a, device=True)
    g.sync()


@pytest.mark.parametrize("name,n_pairs", [("C1", 500_000), ("C2", 25_000_000), ("C5", 5_000_000), ("C3", 50_000_000), ("C4", 50_000_000)])
def test_full_size_properties(pkg, oracle, name, n_pairs):
    import torch
    cfg = dict(pkg.synth.CONFIGS[name])
    K, kw, L = cfg["K"], cfg["key_words"], cfg["read_len"]
    dev = torch.device("cuda", 0)
    free_b, _ = torch.cuda.mem_get_info()
    instances = 2 * n_pairs * (L - K + 1)
    slot = 64 if K > 63 else 32
    # distinct k-mers are dominated by error k-mers: a window is error-free with probability 0.99^K
    est = instances * (1.0 - 0.99 ** K) * 1.03 + 6e7
    cap1 = max(est / 0.85, min(2 * est, (60 << 30) / slot))     # the library's sizing rule (load 0.5 up to 60 GiB)
    if free_b < cap1 * slot + 2 * n_pairs * pkg.synth.stride_bytes(L) + 2e9:
        pytest.skip(f"{name} needs more free HBM than this box has ({free_b / 2**30:.0f} GiB)")
    tr, d_packed, stride = _device_reads(pkg, cfg, n_pairs, dev)
    try:
        g = pkg.PregraphGPU(K, kw, L, capacity_hint=int(est))
    except pkg.SdtGpuError as e:
        pytest.skip(str(e))
    _insert_all(pkg, g, d_packed, L, stride, 1 << 22)
    st = g.stats()
    fp1 = g.table_checksum()
    g.close()
    assert st.n_instances == instances == int(fp1[1])          # conservation: "kmer in reads" == "kmer processed"
    assert st.n_nodes == int(fp1[3])
    # a different capacity and batch size: same multiset
    g = pkg.PregraphGPU(K, kw, L, capacity_hint=int(st.n_nodes * 1.08))
    _insert_all(pkg, g, d_packed, L, stride, (1 << 21) + 4 * 12345)
    fp2 = g.table_checksum()
    g.close()
    assert np.array_equal(fp1, fp2)
    # the sliced build (bench.py's path and the drop-in's default) at full size, WITHOUT a capacity hint and with one:
    # same counters, same fingerprint as the single-pass insert
    for hint in (0, int(st.n_nodes * 1.02) + 1024):
        g = pkg.PregraphGPU(K, kw, L, capacity_hint=hint, sliced=True)
        _insert_all(pkg, g, d_packed, L, stride, 1 << 22)
        st_s = g.stats()
        fp_s = g.table_checksum()
        g.close()
        assert (st_s.n_instances, st_s.n_nodes) == (st.n_instances, st.n_nodes), (name, hint)
        assert np.array_equal(fp1, fp_s), (name, hint)
    # anchor to the oracle: the first 40 000 device-generated reads, bit for bit
    n_sub = 40_000
    sub = d_packed[:n_sub].cpu().numpy()
    reads, lens = pkg.synth.make_reads(tr, n_sub // 2, L, cfg["seed"])
    assert np.array_equal(sub, pkg.synth.pack_reads(reads, lens, stride))      # device generator == numpy generator
    ref = oracle.run_hashing(reads, lens, K, kw, 8, cfg["d"])
    with pkg.PregraphGPU(K, kw, L, capacity_hint=int(ref.nodes) + 1000) as g:
        g.push_reads(d_packed[:n_sub], None, None, n_reads=n_sub, uniform_len=L, stride_bytes=stride, device=True)
        freq, st2 = g.finalize(cfg["d"])
        assert (st2.n_instances, st2.n_nodes, st2.n_removed, st2.n_linear) == (ref.instances, ref.nodes, ref.removed, ref.linear)
        assert np.array_equal(freq, ref.kmerfreq)
        assert np.array_equal(pkg.nodes_to_records(g.export_nodes(8)), oracle.sorted_multiset(ref.records))
