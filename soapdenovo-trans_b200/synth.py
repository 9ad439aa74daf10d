"""Seeded synthetic transcriptome reads (BASELINE.json configs).

Random transcripts with spliced isoforms (exon skipping), log-normal expression, paired-end
reads with i.i.d. 1 % substitution errors.  Everything that depends on the read index is pure
64-bit integer arithmetic (splitmix-style counter hashing), so that the numpy implementation here
and the CUDA generator in csrc/sdt_synth.cu produce bit-identical reads — the CPU oracle, the
reference binary (via FASTA) and the GPU path can therefore be fed exactly the same input at any
size without moving reads between machines.

Base codes follow the reference: A=0 C=1 T=2 G=3, complement = x ^ 2 (inc/def.h:39-42).
Reads are emitted in the reference's arrival order for paired files: read1, read2 alternately
(prlHashReads.c:493-567), i.e. read r = 2*pair + mate.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

U64 = np.uint64
_M1 = U64(0xBF58476D1CE4E5B9)
_M2 = U64(0x94D049BB133111EB)
_G1 = U64(0x9E3779B97F4A7C15)
_G2 = U64(0xD1B54A32D192ED03)
CUM_BITS = 40
ERR_THRESHOLD = 655  # of 65536 -> 0.9995 % ~ "1 % substitution error"


def mix64(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=U64)
    x = (x ^ (x >> U64(30))) * _M1
    x = (x ^ (x >> U64(27))) * _M2
    return x ^ (x >> U64(31))


def h3(seed: int, a, b) -> np.ndarray:
    """Counter hash h(seed, a, b); identical in csrc/sdt_synth.cu."""
    with np.errstate(over="ignore"):
        a = np.asarray(a, dtype=U64)
        b = np.asarray(b, dtype=U64)
        return mix64(mix64(U64(seed & 0xFFFFFFFFFFFFFFFF) + a * _G1) ^ (b * _G2))


@dataclasses.dataclass
class Transcriptome:
    bases: np.ndarray      # uint8 codes, all transcripts concatenated
    starts: np.ndarray     # uint64 [T] start of each transcript in `bases`
    lengths: np.ndarray    # uint32 [T]
    cum: np.ndarray        # uint64 [T] inclusive cumulative expression weight, cum[-1] == 2**CUM_BITS
    seed: int


def make_transcriptome(n_transcripts: int, seed: int, hot: int = 0, hot_factor: float = 1e5) -> Transcriptome:
    """Genes of 4-9 exons (100-400 bp each), 1-3 isoforms per gene by skipping internal exons
    (first and last exon kept), expression ~ LogNormal(0, 1.5) x length.  `hot` extra weight is
    given to the first `hot` transcripts (config 5: few transcripts at 10^5 x depth)."""
    rng = np.random.default_rng(seed)
    seqs = []
    while len(seqs) < n_transcripts:
        n_exons = int(rng.integers(4, 10))
        exons = [rng.integers(0, 4, size=int(rng.integers(100, 401)), dtype=np.uint8) for _ in range(n_exons)]
        n_iso = int(rng.integers(1, 4))
        seen = set()
        for iso in range(n_iso):
            if iso == 0:
                keep = np.ones(n_exons, dtype=bool)
            else:
                keep = rng.random(n_exons) < 0.6
                keep[0] = keep[-1] = True
            key = keep.tobytes()
            if key in seen:
                continue
            seen.add(key)
            seqs.append(np.concatenate([e for e, k in zip(exons, keep) if k]))
            if len(seqs) == n_transcripts:
                break
    lengths = np.array([len(s) for s in seqs], dtype=np.uint32)
    starts = np.zeros(len(seqs), dtype=np.uint64)
    starts[1:] = np.cumsum(lengths[:-1], dtype=np.uint64)
    w = np.exp(1.5 * rng.standard_normal(len(seqs))) * lengths
    if hot:
        w[:hot] = np.median(w) * hot_factor
    c = np.cumsum(w / w.sum())
    cum = np.minimum(np.floor(c * float(1 << CUM_BITS)), float((1 << CUM_BITS) - 1)).astype(np.uint64)
    cum = np.maximum.accumulate(cum)
    cum[-1] = np.uint64(1 << CUM_BITS)
    return Transcriptome(np.concatenate(seqs), starts, lengths, cum, seed)


def frag_len_from_hash(h: np.ndarray) -> np.ndarray:
    """~N(200, 20): sum of the 8 bytes of h (mean 1020, sd 209) rescaled with integer arithmetic."""
    s = np.zeros(h.shape, dtype=np.int64)
    for k in range(8):
        s += ((h >> U64(8 * k)) & U64(0xFF)).astype(np.int64)
    return (20 * s + 21400) // 209


def make_reads(tr: Transcriptome, n_pairs: int, read_len: int, seed: int, first_pair: int = 0,
               ragged: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """Returns (bases[2*n_pairs, read_len] uint8 codes, lens[2*n_pairs] uint32) in arrival order.
    `ragged` > 0 shortens read r by h(seed^7, r, 0) % ragged bases (test-only raggedness)."""
    p = np.arange(first_pair, first_pair + n_pairs, dtype=np.uint64)
    u = h3(seed, p, 0) >> U64(64 - CUM_BITS)
    t = np.searchsorted(tr.cum, u, side="right")
    tlen = tr.lengths[t].astype(np.int64)
    frag = frag_len_from_hash(h3(seed, p, 1))
    frag = np.minimum(np.maximum(frag, read_len), tlen)
    start = (h3(seed, p, 2) % (tlen - frag + 1).astype(np.uint64)).astype(np.int64)
    base0 = tr.starts[t].astype(np.int64) + start
    idx = np.arange(read_len, dtype=np.int64)
    r1 = tr.bases[base0[:, None] + idx[None, :]]
    r2 = tr.bases[(base0 + frag - 1)[:, None] - idx[None, :]] ^ np.uint8(2)
    reads = np.empty((2 * n_pairs, read_len), dtype=np.uint8)
    reads[0::2] = r1
    reads[1::2] = r2
    r = np.arange(2 * first_pair, 2 * (first_pair + n_pairs), dtype=np.uint64)
    e = h3(seed ^ 0xE, r[:, None], idx[None, :].astype(np.uint64))
    err = (e & U64(0xFFFF)) < U64(ERR_THRESHOLD)
    sub = (((e >> U64(16)) % U64(3)) + U64(1)).astype(np.uint8)
    reads = np.where(err, (reads + sub) & np.uint8(3), reads).astype(np.uint8)
    lens = np.full(2 * n_pairs, read_len, dtype=np.uint32)
    if ragged:
        lens = (read_len - (h3(seed ^ 0x7, r, 0) % U64(ragged)).astype(np.int64)).astype(np.uint32)
    return reads, lens


def stride_bytes(max_read_len: int) -> int:
    """Packed read stride: 4 bases per byte, rounded up to a multiple of 4 bytes (include/sdtgpu.h)."""
    return ((max_read_len + 3) // 4 + 3) // 4 * 4


def pack_reads(reads: np.ndarray, lens: np.ndarray | None = None, stride: int | None = None) -> np.ndarray:
    """2-bit pack in the reference's tight-string convention (seq.c:49-90): 4 bases per byte,
    first base in bits 7..6.  Bases beyond a read's length are packed as 0."""
    n, L = reads.shape
    stride = stride or stride_bytes(L)
    buf = np.zeros((n, stride * 4), dtype=np.uint8)
    buf[:, :L] = reads & 3
    if lens is not None:
        buf[np.arange(stride * 4)[None, :] >= np.asarray(lens)[:, None]] = 0
    q = buf.reshape(n, stride, 4)
    return ((q[:, :, 0] << 6) | (q[:, :, 1] << 4) | (q[:, :, 2] << 2) | q[:, :, 3]).astype(np.uint8)


def nmask_reads(reads: np.ndarray, stride: int | None = None) -> np.ndarray:
    """1 bit per base (bit 7 of byte 0 = base 0) set where the base code is 4 (N with -n);
    stride_bytes/2 bytes per read (include/sdtgpu.h)."""
    n, L = reads.shape
    stride = stride or stride_bytes(L)
    bits = np.zeros((n, stride * 4), dtype=np.uint8)
    bits[:, :L] = reads == 4
    return np.packbits(bits, axis=1, bitorder="big")


_LUT = np.frombuffer(b"ACTGN", dtype=np.uint8)


def write_fasta(path: str, reads: np.ndarray, lens: np.ndarray, tag: str = "r") -> None:
    """One-line FASTA records as the reference's AIO reader needs them:
    one sequence line per record and file size never a multiple of 32768 bytes."""
    n, L = reads.shape
    uniform = bool(np.all(lens == L))
    with open(path, "wb") as f:
        if uniform and n:
            hdr = np.char.zfill(np.arange(n).astype("S10"), 10)
            rows = np.empty((n, 1 + 1 + 10 + 1 + L + 1), dtype=np.uint8)
            rows[:, 0] = ord(">")
            rows[:, 1] = ord(tag[0])
            rows[:, 2:12] = np.frombuffer(hdr.tobytes(), dtype=np.uint8).reshape(n, 10)
            rows[:, 12] = 10
            rows[:, 13:13 + L] = _LUT[reads]
            rows[:, 13 + L] = 10
            size = rows.size
            if size % 32768 == 0:  # lengthen the first header by one byte
                f.write(b">x" + rows[0, 1:].tobytes())
                rows = rows[1:]
            rows.tofile(f)
        else:
            out = []
            for i in range(n):
                out.append(b">%s%010d\n" % (tag.encode(), i))
                out.append(_LUT[reads[i, : lens[i]]].tobytes() + b"\n")
            blob = b"".join(out)
            if len(blob) % 32768 == 0:
                blob = b">x" + blob[1:]
            f.write(blob)


def write_fastq(path: str, reads: np.ndarray, lens: np.ndarray, tag: str = "r") -> None:
    out = []
    for i in range(reads.shape[0]):
        s = _LUT[reads[i, : lens[i]]].tobytes()
        out.append(b"@%s%010d\n%s\n+\n%s\n" % (tag.encode(), i, s, b"I" * len(s)))
    blob = b"".join(out)
    if len(blob) % 32768 == 0:
        blob = b"@x" + blob[1:]
    with open(path, "wb") as f:
        f.write(blob)


def write_library(dirpath: str, reads: np.ndarray, lens: np.ndarray, max_rd_len: int, paired: bool = True,
                  fastq: bool = False) -> str:
    """Writes the reads and a reference config file (lib.c:118-438 keys); returns the config path."""
    os.makedirs(dirpath, exist_ok=True)
    cfg = os.path.join(dirpath, "lib.cfg")
    ext, w = ("fq", write_fastq) if fastq else ("fa", write_fasta)
    k1, k2, k = ("q1", "q2", "q") if fastq else ("f1", "f2", "f")
    with open(cfg, "w") as f:
        f.write(f"max_rd_len={max_rd_len}\n[LIB]\navg_ins=200\nreverse_seq=0\nasm_flags=3\n")
        if paired:
            a, b = os.path.join(dirpath, f"r1.{ext}"), os.path.join(dirpath, f"r2.{ext}")
            w(a, reads[0::2], lens[0::2])
            w(b, reads[1::2], lens[1::2])
            f.write(f"{k1}={a}\n{k2}={b}\n")
        else:
            a = os.path.join(dirpath, f"r.{ext}")
            w(a, reads, lens)
            f.write(f"{k}={a}\n")
    return cfg


# ------------------------------------------------------------------ BASELINE.json configs
CONFIGS = {
    # name: (K, key_words, read_len, n_transcripts, n_pairs, seed, hot, deLowKmer)
    "C1": dict(K=25, key_words=1, read_len=100, n_transcripts=2000, n_pairs=500_000, seed=20261017, hot=0, d=0),
    "C2": dict(K=31, key_words=1, read_len=100, n_transcripts=20000, n_pairs=25_000_000, seed=20261018, hot=0, d=0),
    "C3": dict(K=63, key_words=2, read_len=100, n_transcripts=20000, n_pairs=50_000_000, seed=20261019, hot=0, d=0),
    "C4": dict(K=127, key_words=4, read_len=150, n_transcripts=20000, n_pairs=50_000_000, seed=20261020, hot=0, d=0),
    "C5": dict(K=31, key_words=1, read_len=100, n_transcripts=2000, n_pairs=5_000_000, seed=20261021, hot=5, d=2),
}
