"""GPU parity tests of the sliced build (SDTGPU_F_SLICED, csrc/sdt_skm.cuh): reads -> super-k-mer
records -> one scatter by minimizer slice -> every slice built in shared memory -> compact node store.  Same bar as
test_gpu_parity.py: bit-exact multiset, counters, kmerFreq and the reference's (set, slot) layout
against the oracle, through the C ABI."""
import numpy as np
import pytest

from conftest import make_dataset
from test_gpu_parity import check_against_oracle, run_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,kw,d", [(25, 1, 0), (25, 1, 2), (31, 1, 0), (13, 1, 1), (33, 2, 0), (63, 2, 1), (63, 4, 0),
                                    (25, 4, 2), (65, 4, 0), (95, 4, 0), (97, 4, 1), (127, 4, 0)])
def test_sliced_parity_ragged(pkg, oracle, tiny_transcriptome, K, kw, d):
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 3000, L, 11 + K, ragged=40)
    check_against_oracle(pkg, oracle, reads, lens, K, kw, d=d, batches=3, hint=400_000, sliced=True)


@pytest.mark.parametrize("K,kw", [(31, 1), (63, 2), (127, 4)])
def test_sliced_tiny_slices_retries_and_record_overflow(pkg, oracle, tiny_transcriptome, monkeypatch, K, kw):
    """64-slot slice images: tens of thousands of slices, many of which overflow and are retried split
    by k-mer hash; a record area sized for 1/64 record per window, so the emit kernel runs out of room
    and the whole read log is emitted again."""
    monkeypatch.setenv("SDTGPU_SLICE_SLOTS", "64")
    monkeypatch.setenv("SDTGPU_SLICE_LOAD", "0.7")
    monkeypatch.setenv("SDTGPU_REC_DIV", "64")
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 6000, L, 23, ragged=20)
    ref = oracle.run_hashing(reads, lens, K, kw, 8, 1)
    g, freq, st = run_gpu(pkg, reads, lens, K, kw, d=1, batches=5, hint=int(ref.nodes * 1.05), sliced=True)
    try:
        geo = g.slice_geometry()
        assert geo["retried_items"] > 0 and geo["n_nodes"] == ref.nodes
        assert (st.n_instances, st.n_nodes, st.n_removed, st.n_linear) == (ref.instances, ref.nodes, ref.removed, ref.linear)
        assert np.array_equal(freq, ref.kmerfreq)
        assert np.array_equal(pkg.nodes_to_records(g.export_nodes(8)), oracle.sorted_multiset(ref.records))
        rec, info = g.export_kmersets(8)
        assert np.array_equal(info, ref.set_info) and np.array_equal(rec, ref.records)
    finally:
        g.close()


@pytest.mark.parametrize("K,kw,m", [(13, 1, 0), (15, 1, 0), (21, 1, 9), (31, 1, 5), (33, 2, 15), (127, 4, 11)])
def test_sliced_minimizer_lengths(pkg, oracle, tiny_transcriptome, monkeypatch, K, kw, m):
    """Short K (few m-mers per window) and explicit minimizer lengths; runs longer than a record holds
    (m = 5 at K = 31: 27 m-mers per window, long runs; K = 127 on 150-bp reads)."""
    if m:
        monkeypatch.setenv("SDTGPU_MINIMIZER", str(m))
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 2500, L, 7 + K, ragged=30)
    check_against_oracle(pkg, oracle, reads, lens, K, kw, d=0, batches=2, hint=400_000, sliced=True)


@pytest.mark.parametrize("K,kw", [(31, 1), (63, 2), (127, 4)])
def test_skm_stage_import_one_rank(pkg, oracle, tiny_transcriptome, K, kw):
    """The two halves of the multi-GPU super-k-mer exchange on one GPU: stage (records grouped by slice),
    a device copy standing in for the all-to-all, import (re-base, count, group, build)."""
    import torch
    from soapdenovo_trans_b200.exchange import _wrap
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 3000, L, 5 + K, ragged=30)
    synth = pkg.synth
    stride = synth.stride_bytes(L)
    packed = synth.pack_reads(reads, lens, stride)
    ref = oracle.run_hashing(reads, lens, K, kw, 8, 1)
    dev = torch.device("cuda", 0)
    with pkg.PregraphGPU(K, kw, L, capacity_hint=400_000, sliced=True) as g:
        g.skm_set_world(0, 1)
        g.push_reads(packed, lens, None, n_reads=len(reads), stride_bytes=stride)
        ptr, starts, counts = g.skm_stage()
        n = counts[0]
        rb = g.slice_geometry()["record_bytes"]
        assert starts[0] == 0 and 0 < n <= g.slice_geometry()["n_records"]       # (this rank's copies are merged before they travel)
        src = _wrap(ptr, n * rb, dev).clone()
        torch.cuda.synchronize()
        dst = _wrap(g.skm_import_buffer(n), n * rb, dev)
        dst.copy_(src)
        torch.cuda.synchronize()
        g.skm_import(n)
        freq, st = g.finalize(1)
        assert (st.n_instances, st.n_nodes, st.n_removed, st.n_linear) == (ref.instances, ref.nodes, ref.removed, ref.linear)
        assert np.array_equal(freq, ref.kmerfreq)
        rec, info = g.export_kmersets(8)
        assert np.array_equal(info, ref.set_info) and np.array_equal(rec, ref.records)


def test_sliced_interleaved_stats_and_pushes(pkg, oracle, tiny_transcriptome):
    """stats() between pushes builds the store early; later pushes rebuild it from all records."""
    reads, lens = make_dataset(pkg, tiny_transcriptome, 3000, 100, 41, ragged=10)
    synth = pkg.synth
    stride = synth.stride_bytes(100)
    packed = synth.pack_reads(reads, lens, stride)
    ref = oracle.run_hashing(reads, lens, 31, 1, 8, 0)
    half = len(reads) // 2
    with pkg.PregraphGPU(31, 1, 100, capacity_hint=300_000, sliced=True) as g:
        g.push_reads(packed[:half], lens[:half], None, n_reads=half, stride_bytes=stride, first_read_ordinal=0)
        st = g.stats()
        assert 0 < st.n_nodes < ref.nodes
        g.push_reads(packed[half:], lens[half:], None, n_reads=len(reads) - half, stride_bytes=stride, first_read_ordinal=half)
        freq, st = g.finalize(0)
        assert (st.n_instances, st.n_nodes, st.n_linear) == (ref.instances, ref.nodes, ref.linear)
        rec, info = g.export_kmersets(8)
        assert np.array_equal(rec, ref.records)


def test_sliced_uniform_and_n_kmer(pkg, oracle, tiny_transcriptome):
    reads, lens = make_dataset(pkg, tiny_transcriptome, 4000, 100, 5)
    check_against_oracle(pkg, oracle, reads, lens, 25, 1, d=0, uniform=True, hint=400_000, thrd_num=5, sliced=True)
    for K, kw in [(25, 1), (63, 4), (99, 4)]:
        L = 150 if K > 63 else 100
        reads, lens = make_dataset(pkg, tiny_transcriptome, 1500, L, 21, ragged=20, n_rate=0.004)
        check_against_oracle(pkg, oracle, reads, lens, K, kw, d=0, n_kmer=True, batches=2, hint=300_000, sliced=True)


def test_sliced_edge_cases(pkg, oracle):
    reads = np.zeros((200, 60), dtype=np.uint8)
    reads[100:] = 2
    lens = np.full(200, 60, dtype=np.uint32)
    # poly-A / poly-T: key 0 is an ordinary key; link counters saturate at 63 while count keeps going
    check_against_oracle(pkg, oracle, reads, lens, 25, 1, d=0, hint=1000, sliced=True)
    check_against_oracle(pkg, oracle, reads, lens, 33, 2, d=3, hint=1000, sliced=True)
    reads[:] = 3
    reads[::2] = 1
    check_against_oracle(pkg, oracle, reads, lens, 31, 1, d=0, hint=1000, sliced=True)
    # nothing to insert at all: too-short reads, an empty batch
    lens2 = np.full(200, 25, dtype=np.uint32)
    check_against_oracle(pkg, oracle, reads, lens2, 25, 1, d=0, hint=1000, sliced=True)
    check_against_oracle(pkg, oracle, reads[:0], lens2[:0], 25, 1, d=0, hint=1000, sliced=True)


def test_sliced_hot_kmers(pkg, oracle):
    """Config-5 flavour: thousands of instances of the same k-mers land in the same slice."""
    tr = pkg.synth.make_transcriptome(60, 13, hot=2)
    reads, lens = make_dataset(pkg, tr, 40000, 100, 17)
    check_against_oracle(pkg, oracle, reads, lens, 31, 1, d=2, hint=2_000_000, sliced=True)
    check_against_oracle(pkg, oracle, reads, lens, 63, 2, d=0, hint=2_000_000, layout=False, sliced=True)
    check_against_oracle(pkg, oracle, reads, lens, 99, 4, d=0, hint=2_000_000, layout=False, sliced=True)


def test_sliced_equals_single_pass_fingerprint(pkg, tiny_transcriptome):
    """The order-independent table fingerprint of the sliced build equals the single-pass insert's."""
    reads, lens = make_dataset(pkg, pkg.synth.make_transcriptome(300, 3), 30000, 100, 9)
    sums = []
    for sliced in (False, True):
        g, freq, st = run_gpu(pkg, reads, lens, 31, 1, batches=4, hint=3_000_000, sliced=sliced)
        sums.append((g.table_checksum().tolist(), st.n_nodes, st.n_instances, freq.tolist()))
        g.close()
    assert sums[0] == sums[1]


def test_sliced_reset_and_reuse(pkg, oracle, tiny_transcriptome):
    """bench.py's step: reset, push, sync — repeated on one handle; the table is rebuilt each time."""
    reads, lens = make_dataset(pkg, tiny_transcriptome, 2000, 100, 3)
    synth = pkg.synth
    stride = synth.stride_bytes(100)
    packed = synth.pack_reads(reads, lens, stride)
    ref = oracle.run_hashing(reads, lens, 31, 1, 8, 0)
    with pkg.PregraphGPU(31, 1, 100, capacity_hint=300_000, sliced=True) as g:
        for it in range(3):
            g.reset()
            g.push_reads(packed, lens, None, n_reads=len(reads), stride_bytes=stride)
            g.sync()
            st = g.stats()
            assert (st.n_instances, st.n_nodes) == (ref.instances, ref.nodes)
        assert np.array_equal(pkg.nodes_to_records(g.export_nodes(8)), oracle.sorted_multiset(ref.records))


def test_sliced_without_hint_and_with_a_hint_far_too_small(pkg, oracle, tiny_transcriptome):
    """capacity_hint = 0: the reads are logged, the geometry comes from their number at the end of the epoch.
    A hint far too small: one slice for everything — it overflows, is cut into sub-slices in one pass
    (skm_split_kernel) and the node store grows; nothing is dropped, nothing fails."""
    reads, lens = make_dataset(pkg, tiny_transcriptome, 3000, 100, 1)
    for hint in (0, 1000):
        check_against_oracle(pkg, oracle, reads, lens, 31, 1, d=1, batches=3, hint=hint, sliced=True)
    check_against_oracle(pkg, oracle, reads, lens, 63, 2, d=0, batches=2, hint=0, sliced=True)
    reads, lens = make_dataset(pkg, tiny_transcriptome, 1500, 150, 2)
    check_against_oracle(pkg, oracle, reads, lens, 127, 4, d=0, batches=2, hint=500, sliced=True)


@pytest.mark.parametrize("K,kw", [(31, 1), (63, 2), (127, 4)])
def test_sliced_split_into_sub_slices(pkg, oracle, monkeypatch, K, kw):
    """Config-5 flavour with small images: the hot loci overflow their slices by far and go through the
    one-pass split; with SDTGPU_NO_SPLIT the same input goes through the hash-filtered retries."""
    monkeypatch.setenv("SDTGPU_SLICE_SLOTS", "256")
    L = 150 if K > 63 else 100
    tr = pkg.synth.make_transcriptome(60, 13, hot=2)
    reads, lens = make_dataset(pkg, tr, 20000, L, 17)
    ref = oracle.run_hashing(reads, lens, K, kw, 8, 0)
    for no_split in (False, True):
        if no_split:
            monkeypatch.setenv("SDTGPU_NO_SPLIT", "1")
        g, freq, st = run_gpu(pkg, reads, lens, K, kw, d=0, batches=3, hint=int(ref.nodes * 1.05), sliced=True)
        try:
            assert g.slice_geometry()["retried_items"] > 0
            assert (st.n_instances, st.n_nodes, st.n_linear) == (ref.instances, ref.nodes, ref.linear)
            assert np.array_equal(freq, ref.kmerfreq)
            assert np.array_equal(pkg.nodes_to_records(g.export_nodes(8)), oracle.sorted_multiset(ref.records))
        finally:
            g.close()
