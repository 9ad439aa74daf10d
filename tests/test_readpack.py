"""Parity of the multi-threaded FASTA/FASTQ parser + packer (include/sdtpack.h) with the reference's
parse rules (readseq1by1.c:122-178, 281-340; arrival order prlHashReads.c:493-567): the packed output
must be exactly the packing of the reads that were written, for every container format, thread
count and batch size, including the character-level corner cases.  CPU only."""
import os
import time

import numpy as np
import pytest

from conftest import make_dataset


def _read_all(pkg, p1, p2, fastq, L, max_reads, threads, n_kmer=False, reverse=False):
    rp = pkg.ReadPacker(p1, p2, fastq=fastq, n_threads=threads)
    P, Ln, M = [], [], []
    while True:
        packed, lens, nmask = rp.next(L, max_reads, n_kmer=n_kmer, reverse=reverse)
        if len(lens) == 0:
            break
        P.append(packed.copy()); Ln.append(lens.copy())
        if n_kmer:
            M.append(nmask.copy())
    rp.close()
    stride = pkg.synth.stride_bytes(L)
    if not P:
        return np.zeros((0, stride), np.uint8), np.zeros(0, np.uint32), None
    return np.concatenate(P), np.concatenate(Ln), (np.concatenate(M) if n_kmer else None)


@pytest.mark.parametrize("fastq", [False, True])
@pytest.mark.parametrize("paired", [True, False])
@pytest.mark.parametrize("threads,max_reads", [(1, 1000), (4, 4096), (8, 100000)])
def test_matches_written_reads(pkg, tiny_transcriptome, tmp_path, fastq, paired, threads, max_reads):
    reads, lens = make_dataset(pkg, tiny_transcriptome, 6000, 100, 5, ragged=80)       # some reads end up shorter than K+1
    cfg = pkg.synth.write_library(str(tmp_path), reads, lens, 100, paired=paired, fastq=fastq)
    ext = "fq" if fastq else "fa"
    p1 = str(tmp_path / (f"r1.{ext}" if paired else f"r.{ext}"))
    p2 = str(tmp_path / f"r2.{ext}") if paired else None
    got_p, got_l, _ = _read_all(pkg, p1, p2, fastq, 100, max_reads, threads)
    assert np.array_equal(got_l, lens)
    assert np.array_equal(got_p, pkg.synth.pack_reads(reads, lens))
    assert os.path.exists(cfg)


def test_character_rules_truncation_and_n(pkg, tmp_path):
    """lower case, '.', N with and without -n, junk characters dropped AFTER truncation to max_rd_len,
    CRLF line ends, a second sequence line that must be ignored, an empty record, no final newline."""
    p = tmp_path / "x.fa"
    p.write_bytes(b">a\nacgtACGTnN..xx-ACGT\r\n>b desc\nACGTACGTACGTACGTACGT\nTTTTTTTT\n>c\n\n>d\nGATTACA")
    enc = lambda s: np.array([{"A": 0, "C": 1, "T": 2, "G": 3, "N": 4}[c] for c in s], dtype=np.uint8)   # noqa: E731
    L = 16
    # -n: N -> 4; '.' -> A; 'x' -> base2int('X') = (0x58 & 6) >> 1 = 0; '-' dropped; first 16 raw chars only
    want = ["ACGTACGTNNAAAA", "ACGTACGTACGTACGT", "", "GATTACA"]        # 'xx' -> AA, '-' is the 15th raw char, 16th is 'A'
    want[0] = "ACGTACGTNNAAAA" + "A"                                     # raw[:16] = acgtACGTnN..xx-A -> 15 codes
    packed, lens, nmask = _read_all(pkg, str(p), None, False, L, 100, 2, n_kmer=True)
    assert lens.tolist() == [len(w) for w in want]
    for i, w in enumerate(want):
        codes = enc(w)
        ref = pkg.synth.pack_reads(codes[None, :] if len(codes) else np.zeros((1, 0), np.uint8), np.array([len(codes)]), pkg.synth.stride_bytes(L))
        assert np.array_equal(packed[i], ref[0]), i
        bits = np.unpackbits(nmask[i])[: len(w)]
        assert bits.tolist() == [int(c == "N") for c in w]
    # without -n an N is base2int('N') = 3 = G
    packed2, lens2, _ = _read_all(pkg, str(p), None, False, L, 100, 1)
    assert lens2.tolist() == lens.tolist()
    assert np.array_equal(packed2[0], pkg.synth.pack_reads(enc(want[0].replace("N", "G"))[None, :], np.array([15]), pkg.synth.stride_bytes(L))[0])


def test_reverse_seq_and_unequal_pairs(pkg, tmp_path):
    a, b = tmp_path / "1.fa", tmp_path / "2.fa"
    a.write_bytes(b">1\nAACCGGTT\n>2\nACGT\n>3\nGGGG\n")
    b.write_bytes(b">1\nTTTT\n")
    packed, lens, _ = _read_all(pkg, str(a), str(b), False, 8, 100, 1, reverse=True)
    enc = lambda s: np.array([{"A": 0, "C": 1, "T": 2, "G": 3}[c] for c in s], dtype=np.uint8)            # noqa: E731
    order = ["AACCGGTT", "TTTT", "ACGT", "GGGG"]          # read1, read2, then file 1 drained
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))                                        # noqa: E731
    assert lens.tolist() == [8, 4, 4, 4]
    for i, s in enumerate(order):
        assert np.array_equal(packed[i], pkg.synth.pack_reads(enc(rc(s))[None, :], np.array([len(s)]), 4)[0])


def test_throughput_report(pkg, tiny_transcriptome, tmp_path, capsys):
    """Not a pass/fail bar: prints reads/s so the log shows what the host side delivers."""
    reads, lens = make_dataset(pkg, tiny_transcriptome, 100000, 100, 9)
    pkg.synth.write_library(str(tmp_path), reads, lens, 100, paired=True)
    for threads in (1, 8):
        t0 = time.perf_counter()
        p, l, _ = _read_all(pkg, str(tmp_path / "r1.fa"), str(tmp_path / "r2.fa"), False, 100, 1 << 18, threads)
        dt = time.perf_counter() - t0
        assert len(l) == 200000
        with capsys.disabled():
            print(f"\n[sdtpack] {threads} thread(s): {len(l) / dt / 1e6:.1f} M reads/s")
