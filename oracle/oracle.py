"""ctypes binding of oracle/liboracle.so and helpers around oracle/_ref — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

RECORD = np.dtype([("set", "<u4"), ("pad0", "<u4"), ("slot", "<u8"), ("key", "<u8", (4,)),
                   ("l_links", "<u4"), ("rword", "<u4"), ("count", "<u4"), ("pad1", "<u4")])
assert RECORD.itemsize == 64

R_LINKS, R_LINEAR, R_DELETED, R_SINGLE = 0x00FFFFFF, 0x01000000, 0x02000000, 0x08000000


class Kmer(C.Structure):
    _fields_ = [("w", C.c_uint64 * 4)]


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    src = [os.path.join(HERE, "sdt_oracle.c"), os.path.join(HERE, "sdt_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return so


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.sdto_base2int.restype = C.c_int
        L.sdto_create_filter.restype = Kmer
        L.sdto_create_filter.argtypes = [C.c_int]
        L.sdto_reverse_complement.restype = Kmer
        L.sdto_reverse_complement.argtypes = [Kmer, C.c_int, C.c_int]
        L.sdto_kmer_smaller.restype = C.c_int
        L.sdto_kmer_smaller.argtypes = [Kmer, Kmer]
        L.sdto_hash_kmer.restype = C.c_uint64
        L.sdto_hash_kmer.argtypes = [Kmer, C.c_int]
        L.sdto_find_next_prime.restype = C.c_uint64
        L.sdto_find_next_prime.argtypes = [C.c_uint64]
        L.sdto_chop_read.restype = C.c_int
        L.sdto_chop_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sdto_run_create.restype = C.c_void_p
        L.sdto_run_create.argtypes = [C.c_int] * 4
        L.sdto_run_push.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.sdto_run_finalize.argtypes = [C.c_void_p, C.c_int]
        for f in ("nodes", "instances", "removed", "linear"):
            getattr(L, "sdto_run_" + f).restype = C.c_uint64
            getattr(L, "sdto_run_" + f).argtypes = [C.c_void_p]
        L.sdto_run_kmerfreq.argtypes = [C.c_void_p, C.c_void_p]
        L.sdto_run_set_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sdto_run_dump.argtypes = [C.c_void_p, C.c_void_p]
        L.sdto_run_destroy.argtypes = [C.c_void_p]
        L.sdto_run_set_max_read_len.argtypes = [C.c_void_p, C.c_int]
        L.sdto_run_first_ordinals.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def encode(seq: str) -> np.ndarray:
    """ASCII -> base codes, inc/def.h:39 (N -> 3 unless the caller substitutes 4)."""
    a = np.frombuffer(seq.upper().encode(), dtype=np.uint8)
    return ((a & 6) >> 1).astype(np.uint8)


def kmer_from_codes(codes) -> Kmer:
    v = 0
    for c in codes:
        v = (v << 2) | int(c)
    k = Kmer()
    for i in range(4):
        k.w[3 - i] = (v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
    return k


def kmer_int(k: Kmer) -> int:
    return (k.w[0] << 192) | (k.w[1] << 128) | (k.w[2] << 64) | k.w[3]


def chop_read(codes: np.ndarray, K: int, key_words: int, n_kmer: int = 0):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n = max(len(codes) - K + 1, 0)
    kmers = np.zeros((max(n, 1), 4), dtype=np.uint64)
    prev = np.zeros(max(n, 1), dtype=np.uint8)
    nxt = np.zeros(max(n, 1), dtype=np.uint8)
    got = lib().sdto_chop_read(codes.ctypes.data, len(codes), K, key_words, n_kmer,
                               kmers.ctypes.data, prev.ctypes.data, nxt.ctypes.data)
    return kmers[:got], prev[:got], nxt[:got]


class OracleResult:
    def __init__(self, records, nodes, instances, removed, linear, freq, set_info, first_ordinals=None):
        self.records = records          # RECORD array in (set, slot) order
        self.first_ordinals = first_ordinals  # uint64 per record: read index * (max_read_len-K+1) + window
        self.nodes, self.instances, self.removed, self.linear = nodes, instances, removed, linear
        self.kmerfreq = freq            # int64[257]
        self.set_info = set_info        # uint64[thrd_num, 3] size, count, max


def run_hashing(reads: np.ndarray, lens: np.ndarray, K: int, key_words: int, thrd_num: int = 8,
                deLowKmer: int = 0, n_kmer: int = 0, batches: int = 1, max_read_len: int = 0) -> OracleResult:
    """The oracle's prlRead2HashTable on in-memory reads (`reads`[n, L] base codes, `lens`[n])."""
    L = lib()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    n, width = reads.shape if reads.ndim == 2 else (0, 0)
    h = L.sdto_run_create(K, key_words, thrd_num, n_kmer)
    L.sdto_run_set_max_read_len(h, max_read_len or width)
    try:
        step = max((n + batches - 1) // batches, 1)
        for a in range(0, n, step):
            b = min(a + step, n)
            offs = (np.arange(b - a, dtype=np.uint64) * np.uint64(width))
            chunk = np.ascontiguousarray(reads[a:b])
            L.sdto_run_push(h, chunk.ctypes.data, offs.ctypes.data, lens[a:b].ctypes.data, b - a)
        L.sdto_run_finalize(h, deLowKmer)
        nodes = L.sdto_run_nodes(h)
        rec = np.zeros(nodes, dtype=RECORD)
        if nodes:
            L.sdto_run_dump(h, rec.ctypes.data)
        freq = np.zeros(257, dtype=np.int64)
        L.sdto_run_kmerfreq(h, freq.ctypes.data)
        info = np.zeros((thrd_num, 3), dtype=np.uint64)
        for t in range(thrd_num):
            L.sdto_run_set_info(h, t, info[t].ctypes.data)
        first = np.zeros(nodes, dtype=np.uint64)
        if nodes:
            L.sdto_run_first_ordinals(h, first.ctypes.data)
        return OracleResult(rec, nodes, L.sdto_run_instances(h), L.sdto_run_removed(h), L.sdto_run_linear(h), freq, info, first)
    finally:
        L.sdto_run_destroy(h)


def sorted_multiset(rec: np.ndarray) -> np.ndarray:
    """Order-free view of a dump: rows (key[4], count, l_links, rword) sorted by key."""
    m = np.zeros(len(rec), dtype=[("key", "<u8", (4,)), ("count", "<u4"), ("l_links", "<u4"), ("rword", "<u4")])
    m["key"], m["count"], m["l_links"], m["rword"] = rec["key"], rec["count"], rec["l_links"], rec["rword"]
    order = np.lexsort((rec["key"][:, 3], rec["key"][:, 2], rec["key"][:, 1], rec["key"][:, 0]))
    return m[order]


def _fmix64(k: np.ndarray) -> np.ndarray:
    k = k.astype(np.uint64)
    k ^= k >> np.uint64(33)
    k *= np.uint64(0xff51afd7ed558ccd)
    k ^= k >> np.uint64(33)
    k *= np.uint64(0xc4ceb9fe1a85ec53)
    k ^= k >> np.uint64(33)
    return k


def table_checksum(rec: np.ndarray, device_key_words: int) -> np.ndarray:
    """CPU twin of sdtgpu_table_checksum over dump records (or exported nodes): order-independent."""
    with np.errstate(over="ignore"):
        x = np.full(len(rec), 0x9E3779B97F4A7C15, dtype=np.uint64)
        for q in range(4 - device_key_words, 4):
            x = _fmix64(x ^ rec["key"][:, q])
        L = rec["l_links"].astype(np.uint64)
        R = (rec["rword"] & R_LINKS).astype(np.uint64)
        x = _fmix64(x ^ ((rec["count"].astype(np.uint64) << np.uint64(32)) | L))
        x = _fmix64(x ^ R)
        lbits = int(np.unpackbits(np.ascontiguousarray(L).view(np.uint8)).sum())
        rbits = int(np.unpackbits(np.ascontiguousarray(R).view(np.uint8)).sum())
        return np.array([x.sum(dtype=np.uint64), rec["count"].astype(np.uint64).sum(dtype=np.uint64),
                         np.uint64(lbits) + (np.uint64(rbits) << np.uint64(32)), len(rec)], dtype=np.uint64)


# ------------------------------------------------------------------ the real reference (oracle/_ref)
def ref_binary(key_words: int, stock: bool = False) -> str | None:
    name = ("SOAPdenovo-Trans-%dmer" if stock else "ref_hash_%d") % (31 if key_words == 1 else 127)
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def read_dump(path: str):
    with open(path, "rb") as f:
        head = f.read(32)
        assert head[:8] == b"SDTDUMP1", head[:8]
        kw, ns = np.frombuffer(head[8:16], dtype="<u4")
        nr, K = np.frombuffer(head[16:32], dtype="<u8")
        info = np.frombuffer(f.read(24 * int(ns)), dtype="<u8").reshape(int(ns), 3)
        rec = np.frombuffer(f.read(), dtype=RECORD)
    assert len(rec) == nr
    return rec, info, int(kw), int(K)


def run_reference(cfg: str, prefix: str, K: int, key_words: int, thrd_num: int = 8, deLowKmer: int = 0,
                  n_kmer: int = 0, dump: bool = True, timeout: int = 3600):
    """Runs the UNMODIFIED reference's prlRead2HashTable through oracle/_ref/ref_hash_*.
    Returns (info dict from the REFJSON line, records or None, set_info or None)."""
    exe = ref_binary(key_words)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    cmd = [exe, "hash", cfg, prefix, "-K", str(K), "-p", str(thrd_num), "-d", str(deLowKmer)]
    if n_kmer:
        cmd.append("-n")
    dpath = prefix + ".dump"
    if dump:
        cmd += ["-o", dpath]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout).stdout
    line = [l for l in out.splitlines() if l.startswith("REFJSON ")][-1]
    info = json.loads(line[len("REFJSON "):])
    info["stdout"] = out
    if not dump:
        return info, None, None
    rec, sinfo, _, _ = read_dump(dpath)
    os.remove(dpath)
    return info, rec, sinfo
