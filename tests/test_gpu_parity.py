"""GPU parity tests: the CUDA path, called through the C ABI (libsdtgpu.so), against the oracle on
the same seeded inputs.  Bit-exact: the sorted multiset of (canonical k-mer, count, l_links,
r_links + linear/deleted/single flags) before and after the -d cutoff, the kmerFreq histogram, the
reference's counters, and the (set, slot) layout of the handed-back KmerSets."""
import numpy as np
import pytest

from conftest import make_dataset

pytestmark = pytest.mark.gpu


def run_gpu(pkg, reads, lens, K, kw, d=0, batches=1, thrd_num=8, hint=0, n_kmer=False, uniform=False, max_read_len=None, sliced=False):
    synth = pkg.synth
    L = reads.shape[1]
    max_read_len = max_read_len or L
    stride = synth.stride_bytes(max_read_len)
    packed = synth.pack_reads(reads, lens, stride)
    nmask = synth.nmask_reads(reads, stride) if n_kmer else None
    g = pkg.PregraphGPU(K, kw, max_read_len, capacity_hint=hint, n_kmer=n_kmer, sliced=sliced)
    n = len(reads)
    step = max((n + batches - 1) // batches, 1)
    for a in range(0, n, step):
        b = min(a + step, n)
        g.push_reads(np.ascontiguousarray(packed[a:b]), None if uniform else np.ascontiguousarray(lens[a:b]),
                     None if nmask is None else np.ascontiguousarray(nmask[a:b]), n_reads=b - a,
                     uniform_len=L if uniform else 0, stride_bytes=stride, first_read_ordinal=a)
    freq, st = g.finalize(d)
    return g, freq, st


def check_against_oracle(pkg, oracle, reads, lens, K, kw, d=0, thrd_num=8, layout=True, **kw_gpu):
    n_kmer = kw_gpu.get("n_kmer", False)
    ref = oracle.run_hashing(reads, lens, K, kw, thrd_num, d, n_kmer=int(n_kmer), max_read_len=kw_gpu.get("max_read_len") or 0)
    g, freq, st = run_gpu(pkg, reads, lens, K, kw, d=d, thrd_num=thrd_num, **kw_gpu)
    try:
        assert st.n_instances == ref.instances
        assert st.n_nodes == ref.nodes
        assert st.n_removed == ref.removed
        assert st.n_linear == ref.linear
        assert np.array_equal(freq, ref.kmerfreq)
        nodes = g.export_nodes(thrd_num)
        assert len(nodes) == ref.nodes
        got, want = pkg.nodes_to_records(nodes), oracle.sorted_multiset(ref.records)
        assert np.array_equal(got, want)
        if layout and ref.nodes:
            # same key -> same set (hash_kmer % thrd_num) and same first-instance ordinal
            rec, info = g.export_kmersets(thrd_num)
            assert np.array_equal(info, ref.set_info)
            assert np.array_equal(rec, ref.records)
        return st
    finally:
        g.close()


@pytest.mark.parametrize("K,kw,d", [(25, 1, 0), (25, 1, 2), (31, 1, 0), (13, 1, 1), (33, 2, 0), (63, 2, 1), (63, 4, 0),
                                    (25, 4, 2), (65, 4, 0), (95, 4, 0), (97, 4, 1), (127, 4, 0)])
def test_table_parity_ragged(pkg, oracle, tiny_transcriptome, K, kw, d):
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 3000, L, 11 + K, ragged=40)
    check_against_oracle(pkg, oracle, reads, lens, K, kw, d=d, batches=3)


@pytest.mark.parametrize("K,kw", [(25, 1), (63, 2), (127, 4)])
def test_table_parity_uniform_device_hint(pkg, oracle, tiny_transcriptome, K, kw):
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 4000, L, 5)
    check_against_oracle(pkg, oracle, reads, lens, K, kw, d=0, uniform=True, hint=400_000, thrd_num=5)


def test_growth_by_device_rehash(pkg, oracle, tiny_transcriptome):
    """capacity_hint = 0: the table starts at 2^20 slots and must re-hash on the device."""
    reads, lens = make_dataset(pkg, pkg.synth.make_transcriptome(400, 3), 30000, 100, 9)
    st = check_against_oracle(pkg, oracle, reads, lens, 31, 1, d=0, batches=6, layout=False)
    assert st.n_grows >= 1


@pytest.mark.parametrize("K,kw", [(25, 1), (63, 4), (99, 4)])
def test_n_kmer_mode(pkg, oracle, tiny_transcriptome, K, kw):
    """-n: windows containing N become key 0 without links (prlHashReads.c:193-196, 242-274)."""
    L = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 1500, L, 21, ragged=20, n_rate=0.004)
    check_against_oracle(pkg, oracle, reads, lens, K, kw, d=0, n_kmer=True, batches=2)


def test_edge_cases(pkg, oracle):
    synth = pkg.synth
    # poly-A / poly-T reads: canonical key 0 (a legal key, not the empty sentinel), link saturation at 63
    reads = np.zeros((200, 60), dtype=np.uint8)
    reads[100:] = 2
    lens = np.full(200, 60, dtype=np.uint32)
    check_against_oracle(pkg, oracle, reads, lens, 25, 1, d=0)
    check_against_oracle(pkg, oracle, reads, lens, 33, 2, d=3)
    # poly-G: every key bit set below 2K (closest legal key to the all-ones sentinel)
    reads[:] = 3
    reads[::2] = 1
    check_against_oracle(pkg, oracle, reads, lens, 31, 1, d=0)
    # reads shorter than K+1 are skipped; a batch may be entirely skipped or empty
    lens2 = np.full(200, 25, dtype=np.uint32)
    check_against_oracle(pkg, oracle, reads, lens2, 25, 1, d=0)
    check_against_oracle(pkg, oracle, reads[:0], lens2[:0], 25, 1, d=0)
    # exactly K+1 bases: two windows per read
    lens3 = np.full(200, 26, dtype=np.uint32)
    check_against_oracle(pkg, oracle, reads, lens3, 25, 1, d=0)


def test_trailing_growth_is_reproduced(pkg, oracle):
    """put_kmerset runs encap_kmerset on every call (newhash.c:415): when a set holds exactly `max`
    keys after its newest one, any later instance grows it once more.  21 reads x 36 + 1 read x 37
    windows = 793 distinct keys = max of the 1031-slot set; the repeated read that follows adds no key
    but the reference ends with 2063 slots.  thrd_num = 1 puts every key in set 0."""
    rng = np.random.default_rng(3)
    K, Wd = 25, 61
    base = rng.integers(0, 4, size=(23, Wd), dtype=np.uint8)
    sizes = []
    for extra in (0, 1):
        reads = np.concatenate([base[:22], base[:extra]])
        lens = np.array([60] * 21 + [61] + [60] * extra, np.uint32)
        ref = oracle.run_hashing(reads, lens, K, 1, 1, 0, max_read_len=Wd)
        assert ref.nodes == 793
        g, freq, st = run_gpu(pkg, reads, lens, K, 1, thrd_num=1, max_read_len=Wd)
        rec, info = g.export_kmersets(1)
        g.close()
        assert np.array_equal(info, ref.set_info) and np.array_equal(rec, ref.records)
        sizes.append(int(info[0, 0]))
    assert sizes == [1031, 2063]


def test_hot_kmer_contention(pkg, oracle, tiny_transcriptome):
    """Config-5 flavour: a handful of transcripts at huge depth; counts far above the 6-bit link
    saturation, then the -d 2 cutoff."""
    tr = pkg.synth.make_transcriptome(60, 13, hot=2)
    reads, lens = make_dataset(pkg, tr, 40000, 100, 17)
    check_against_oracle(pkg, oracle, reads, lens, 31, 1, d=2, hint=2_000_000)
    check_against_oracle(pkg, oracle, reads, lens, 63, 2, d=0, hint=2_000_000, layout=False)
    check_against_oracle(pkg, oracle, reads, lens, 127 if reads.shape[1] > 127 else 99, 4, d=0, hint=2_000_000, layout=False)


def test_bucket_exchange_roundtrip(pkg, oracle, tiny_transcriptome):
    """Send side + receive side of the multi-GPU exchange on one GPU: bucket into 3 owner bins,
    insert every bin back -> identical table."""
    import torch
    synth = pkg.synth
    for K, kw in [(25, 1), (63, 2), (127, 4)]:
        L = 150 if K > 63 else 100
        reads, lens = make_dataset(pkg, tiny_transcriptome, 2500, L, 31, ragged=25)
        ref = oracle.run_hashing(reads, lens, K, kw, 8, 0)
        stride = synth.stride_bytes(L)
        d_packed = torch.from_numpy(synth.pack_reads(reads, lens, stride)).cuda()
        d_lens = torch.from_numpy(lens.astype(np.int32)).cuda()
        g = pkg.PregraphGPU(K, kw, L, capacity_hint=600_000)
        n_ranks, cap = 3, int(ref.instances)
        words = g.record_bytes() // 8
        bins = torch.zeros((n_ranks, cap, words), dtype=torch.int64, device="cuda")
        counts = torch.zeros(n_ranks, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        g.bucket_reads_device(d_packed, d_lens, None, len(reads), 0, stride, 0, n_ranks, bins, cap, counts)
        g.sync()
        c = counts.cpu().numpy()
        assert c.sum() == ref.instances and (c > 0).all()
        for r in range(n_ranks):
            g.insert_records_device(bins[r], int(c[r]))
        freq, st = g.finalize(0)
        assert st.n_nodes == ref.nodes and st.n_instances == ref.instances
        assert np.array_equal(pkg.nodes_to_records(g.export_nodes(8)), oracle.sorted_multiset(ref.records))
        rec, info = g.export_kmersets(8)
        assert np.array_equal(rec, ref.records)
        g.close()


def test_synth_device_matches_numpy(pkg, tiny_transcriptome):
    import torch
    synth, tr = pkg.synth, tiny_transcriptome
    for L in (100, 150):
        reads, lens = synth.make_reads(tr, 3000, L, 77, first_pair=123)
        stride = synth.stride_bytes(L)
        want = synth.pack_reads(reads, lens, stride)
        dev = dict(bases=torch.from_numpy(tr.bases).cuda(), starts=torch.from_numpy(tr.starts.astype(np.int64)).cuda(),
                   lengths=torch.from_numpy(tr.lengths.astype(np.int32)).cuda(),
                   cum=torch.from_numpy(tr.cum.astype(np.int64)).cuda(), n=len(tr.lengths))
        out = torch.zeros((6000, stride), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        pkg.pregraph.synth_reads_device(dev, 77, 123, 3000, L, stride, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want)


def test_state_and_argument_errors(pkg):
    with pytest.raises(pkg.SdtGpuError):
        pkg.PregraphGPU(24, 1, 100)            # even K
    with pytest.raises(pkg.SdtGpuError):
        pkg.PregraphGPU(33, 1, 100)            # K too large for the 31mer build
    g = pkg.PregraphGPU(25, 1, 100, capacity_hint=1000)
    packed = np.zeros((4, 28), dtype=np.uint8)
    with pytest.raises(pkg.SdtGpuError):
        g.push_reads(packed, None, None, n_reads=4, uniform_len=100, stride_bytes=26)   # stride not multiple of 4
    g.push_reads(packed, None, None, n_reads=4, uniform_len=100, stride_bytes=28)
    g.finalize(0)
    with pytest.raises(pkg.SdtGpuError):
        g.push_reads(packed, None, None, n_reads=4, uniform_len=100, stride_bytes=28)   # push after finalize
    g.reset()
    g.push_reads(packed, None, None, n_reads=4, uniform_len=100, stride_bytes=28)
    freq, st = g.finalize(0)
    assert st.n_instances == 4 * 76 and st.n_nodes == 1
    g.close()
