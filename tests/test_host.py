"""CPU-side tests of the product's host logic: the C-ABI library loads and exports every symbol
include/sdtgpu.h declares, argument validation happens before any CUDA call, the KmerSet builder
(host/kmerset_builder.cpp) reproduces the reference's (set, slot) layout from unordered nodes, the
product's own hash_kmer agrees with the reference's known answers, and the read generator/packer
are deterministic.  No GPU compute here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT, make_dataset


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "sdtgpu.h")).read()
    names = sorted(set(re.findall(r"\b(sdtgpu_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = pkg.library()
    for n in names:
        assert hasattr(L, n), f"libsdtgpu.so does not export {n}"
    assert L.sdtgpu_version() >= 1


def test_product_hash_kmer_known_answers(pkg):
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        kat = json.load(f)
    for kw in (1, 4):
        for row in kat[str(kw)]:
            fwd = [int(x, 16) for x in row["fwd"]]
            rc = [int(x, 16) for x in row["rc"]]
            assert pkg.hash_kmer(fwd, kw) == row["hash_fwd"]
            assert pkg.hash_kmer(rc, kw) == row["hash_rc"]


def test_create_rejects_bad_arguments_without_a_gpu(pkg):
    for K, kw, mrl in ((24, 1, 100), (11, 1, 100), (33, 1, 100), (65, 2, 100), (129, 4, 200), (25, 3, 100), (25, 1, 25)):
        with pytest.raises(pkg.SdtGpuError) as e:
            pkg.PregraphGPU(K, kw, mrl)
        assert e.value.code == 1


def test_exchange_calls_reject_bad_arguments_without_a_gpu(pkg):
    """sdtgpu_comm_* / sdtgpu_skm_exchange (the super-k-mer exchange behind the C ABI): argument errors are
    reported before NCCL or CUDA are touched."""
    import ctypes as C
    L = pkg.library()
    c = C.c_void_p()
    uid = (C.c_uint8 * 128)()
    assert L.sdtgpu_comm_unique_id(None) == 1                               # SDTGPU_EINVAL
    for rank, world in ((0, 0), (2, 2), (-1, 2), (0, 65)):
        assert L.sdtgpu_comm_create(C.byref(c), 0, uid, rank, world) == 1 and not c.value
        assert b"rank" in L.sdtgpu_comm_last_error(None)
    assert L.sdtgpu_comm_create(None, 0, uid, 0, 1) == 1
    assert L.sdtgpu_skm_exchange(None, None, 0, None, None) == 1
    assert L.sdtgpu_comm_destroy(None) == 0


@pytest.mark.parametrize("K,kw,p", [(25, 1, 8), (31, 1, 3), (45, 2, 8), (63, 4, 5), (127, 4, 8)])
def test_kmerset_builder_reproduces_reference_layout(pkg, oracle, tiny_transcriptome, K, kw, p):
    L_read = 150 if K > 63 else 100
    reads, lens = make_dataset(pkg, tiny_transcriptome, 4000, L_read, 11, ragged=30)
    r = oracle.run_hashing(reads, lens, K, kw, p, 2)
    nodes = np.zeros(r.nodes, dtype=pkg.NODE_DTYPE)
    for f in ("key", "l_links", "rword", "count", "set"):
        nodes[f] = r.records[f]
    nodes["ordinal"] = r.first_ordinals
    nodes = nodes[np.random.default_rng(1).permutation(len(nodes))]
    lib = pkg.library()
    sets = (C.POINTER(pkg.pregraph.KmerSet) * p)()
    assert lib.sdtgpu_build_kmersets(nodes.ctypes.data, len(nodes), kw, p, None, sets) == 0
    rec, info = pkg.read_kmersets(sets, p, kw)
    lib.sdtgpu_free_kmersets(sets, p)
    assert np.array_equal(info, r.set_info)
    assert np.array_equal(rec, r.records)
    # the partition itself: the product's hash_kmer puts every key in the set the reference chose
    for k, s in zip(nodes["key"][:500], nodes["set"][:500]):
        assert pkg.hash_kmer(k, kw) % p == s


def test_kmerset_builder_trailing_growth(pkg, oracle):
    """put_kmerset runs encap_kmerset on every call (newhash.c:415): instances that follow a set's
    last new key can trigger one more growth; the builder honours set_last_ordinal."""
    # 21 reads x 36 windows + 1 read x 37 windows = 793 distinct keys = max of the 1031-slot set; the
    # repeated read that follows adds no key but makes the reference grow the set to 2063 slots
    rng = np.random.default_rng(3)
    K, W = 25, 61
    base = rng.integers(0, 4, size=(23, W), dtype=np.uint8)
    grown = 0
    for extra in (0, 1):
        reads = np.concatenate([base[:22], base[:extra]])
        lens = np.array([60] * 21 + [61] + [60] * extra, np.uint32)
        r = oracle.run_hashing(reads, lens, K, 1, 1, 0, max_read_len=W)
        assert r.nodes == 793
        nodes = np.zeros(r.nodes, dtype=pkg.NODE_DTYPE)
        for f in ("key", "l_links", "rword", "count", "set"):
            nodes[f] = r.records[f]
        nodes["ordinal"] = r.first_ordinals
        last = np.array([(len(reads) - 1) * (W - K + 1) + (35 if extra else 36)], dtype=np.uint64)
        sets = (C.POINTER(pkg.pregraph.KmerSet) * 1)()
        assert pkg.library().sdtgpu_build_kmersets(nodes.ctypes.data, len(nodes), 1, 1, last.ctypes.data, sets) == 0
        rec, info = pkg.read_kmersets(sets, 1, 1)
        pkg.library().sdtgpu_free_kmersets(sets, 1)
        assert np.array_equal(info, r.set_info), extra
        assert np.array_equal(rec, r.records)
        grown += int(info[0, 0]) == 2063
    assert grown == 1


def test_synth_is_deterministic_and_packs_tight_strings(pkg):
    synth = pkg.synth
    tr1, tr2 = synth.make_transcriptome(50, 42), synth.make_transcriptome(50, 42)
    assert np.array_equal(tr1.bases, tr2.bases) and np.array_equal(tr1.cum, tr2.cum)
    a, la = synth.make_reads(tr1, 1000, 100, 9)
    b, _ = synth.make_reads(tr1, 400, 100, 9, first_pair=600)
    assert np.array_equal(a[1200:], b)                       # counter-based: any slice reproducible
    err = (a[0::2] != tr1.bases[:1][0]).mean()
    assert a.max() <= 3 and a.shape == (2000, 100)
    packed = synth.pack_reads(a, la)
    assert packed.shape == (2000, 28)
    # seq.c:49-90 convention: first base in bits 7..6 of byte 0
    assert ((packed[:, 0] >> 6) == a[:, 0]).all() and ((packed[:, 0] & 3) == a[:, 3]).all()
    assert ((packed[:, 24] >> 6) == a[:, 96]).all() and (packed[:, 25:] == 0).all()
    nm = synth.nmask_reads(np.where(np.arange(100)[None, :] == 9, 4, a).astype(np.uint8))
    assert nm.shape == (2000, 14) and (nm[:, 1] == 0x40).all() and nm[:, 0].max() == 0


def test_fasta_writer_avoids_the_32768_multiple_trap(pkg, tmp_path):
    """the reference hangs when a file's size is a multiple of 32768 bytes."""
    synth = pkg.synth
    L, n = 100, 2048          # row = 114 bytes -> 2048 rows = 233472 = 7.125 * 32768; pick a size that collides
    reads = np.zeros((32768, L - 86 + 0), dtype=np.uint8)   # row length 14 + 14 = 28 bytes -> 32768 rows = 28 * 32768
    lens = np.full(len(reads), reads.shape[1], np.uint32)
    p = tmp_path / "x.fa"
    synth.write_fasta(str(p), reads, lens)
    assert os.path.getsize(p) % 32768 != 0
    assert open(p, "rb").read(2) == b">x"
