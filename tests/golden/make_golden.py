"""Generates tests/golden/{kat,dumps}.json from the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures pin the oracle (tests/test_oracle.py) wherever oracle/_ref cannot be rebuilt."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import sdt_pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

pkg = sdt_pkg.load()
synth = pkg.synth

DATASETS = {
    # name: transcriptome (n, seed), reads (pairs, L, seed, ragged, n_rate), K, key_words, -p, -d, -n, container
    "k25_31mer_p8_d0": dict(tr=(20, 7), pairs=5000, L=100, seed=11, ragged=0, n_rate=0, K=25, kw=1, p=8, d=0, n=0, fmt="fa2"),
    "k25_31mer_p3_d2": dict(tr=(20, 7), pairs=5000, L=100, seed=11, ragged=0, n_rate=0, K=25, kw=1, p=3, d=2, n=0, fmt="fa2"),
    "k31_31mer_ragged": dict(tr=(20, 7), pairs=3000, L=100, seed=12, ragged=40, n_rate=0, K=31, kw=1, p=8, d=0, n=0, fmt="fa2"),
    "k31_31mer_fastq": dict(tr=(20, 7), pairs=3000, L=100, seed=12, ragged=40, n_rate=0, K=31, kw=1, p=8, d=0, n=0, fmt="fq2"),
    "k31_31mer_single": dict(tr=(20, 7), pairs=3000, L=100, seed=12, ragged=40, n_rate=0, K=31, kw=1, p=8, d=0, n=0, fmt="fa1"),
    "k13_31mer_d1": dict(tr=(20, 7), pairs=2000, L=100, seed=13, ragged=10, n_rate=0, K=13, kw=1, p=4, d=1, n=0, fmt="fa2"),
    "k25_127mer": dict(tr=(20, 7), pairs=3000, L=100, seed=14, ragged=0, n_rate=0, K=25, kw=4, p=8, d=0, n=0, fmt="fa2"),
    "k63_127mer_d1": dict(tr=(20, 7), pairs=5000, L=100, seed=15, ragged=20, n_rate=0, K=63, kw=4, p=8, d=1, n=0, fmt="fa2"),
    "k99_127mer": dict(tr=(20, 7), pairs=3000, L=150, seed=16, ragged=30, n_rate=0, K=99, kw=4, p=5, d=0, n=0, fmt="fa2"),
    "k127_127mer": dict(tr=(20, 7), pairs=3000, L=150, seed=17, ragged=0, n_rate=0, K=127, kw=4, p=8, d=0, n=0, fmt="fa2"),
    "k25_31mer_nkmer": dict(tr=(20, 7), pairs=2000, L=100, seed=18, ragged=20, n_rate=0.004, K=25, kw=1, p=8, d=0, n=1, fmt="fa2"),
    "k63_127mer_nkmer": dict(tr=(20, 7), pairs=2000, L=100, seed=19, ragged=20, n_rate=0.004, K=63, kw=4, p=8, d=0, n=1, fmt="fa2"),
    "k25_31mer_n_as_g": dict(tr=(20, 7), pairs=2000, L=100, seed=18, ragged=20, n_rate=0.004, K=25, kw=1, p=8, d=0, n=0, fmt="fa2"),
}

KAT_SEQS = [
    (25, "ACGTACGTACGTACGTACGTACGTA"), (25, "TTTTTTTTTTTTTTTTTTTTTTTTT"), (25, "AAAAAAAAAAAAAAAAAAAAAAAAA"),
    (31, "GATTACAGATTACAGATTACAGATTACAGAT"), (31, "GGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG"), (13, "ACGTTGCAACGTA"),
    (23, "CCCCCCCCCCCCCCCCCCCCCCC"), (31, "ACGTNACGTNACGTNACGTNACGTNACGTNA"),
]
KAT_SEQS_127 = KAT_SEQS + [
    (33, "ACGTACGTACGTACGTACGTACGTAGGCTTAACC"[:33]), (63, ("GATTACA" * 9)[:63]), (65, ("CAGTTGA" * 10)[:65]),
    (95, ("ACCGTTGAAC" * 10)[:95]), (97, ("TTGACCAGTA" * 10)[:97]), (127, ("ACGGTCATTGCA" * 11)[:127]),
    (127, "G" * 127), (127, "A" * 127),
]


def dataset(spec):
    tr = synth.make_transcriptome(*spec["tr"])
    reads, lens = synth.make_reads(tr, spec["pairs"], spec["L"], spec["seed"], ragged=spec["ragged"])
    if spec["n_rate"]:
        rng = np.random.default_rng(spec["seed"])
        reads = reads.copy()
        reads[rng.random(reads.shape) < spec["n_rate"]] = 4
    return reads, lens


def write_input(spec, reads, lens, d):
    fmt = spec["fmt"]
    return synth.write_library(d, reads, lens, spec["L"], paired=fmt.endswith("2"), fastq=fmt.startswith("fq"))


def digest(rec):
    return hashlib.sha256(np.ascontiguousarray(rec).tobytes()).hexdigest()


def main():
    kat = {}
    for kw, seqs in ((1, KAT_SEQS), (4, KAT_SEQS_127)):
        text = "".join(f"{k} {s}\n" for k, s in seqs)
        out = subprocess.run([O.ref_binary(kw), "kat"], input=text, capture_output=True, text=True, check=True).stdout
        rows = []
        for line in out.splitlines():
            f = line.split()
            rows.append(dict(K=int(f[0]), seq=f[1], fwd=f[2:6], rc=f[6:10], smaller=int(f[10]), hash_fwd=int(f[11]), hash_rc=int(f[12])))
        kat[str(kw)] = rows
    grow = subprocess.run([O.ref_binary(1), "grow", "3000000"], capture_output=True, text=True, check=True).stdout
    kat["grow"] = [[int(x) for x in l.split()] for l in grow.splitlines()]
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    dumps = {}
    for name, spec in DATASETS.items():
        reads, lens = dataset(spec)
        with tempfile.TemporaryDirectory() as d:
            cfg = write_input(spec, reads, lens, d)
            info, rec, sinfo = O.run_reference(cfg, os.path.join(d, "out"), spec["K"], spec["kw"], spec["p"], spec["d"], spec["n"])
            freq = np.loadtxt(os.path.join(d, "out.kmerFreq"), dtype=np.int64)
        dumps[name] = dict(spec=spec, sha256=digest(rec), multiset_sha256=digest(O.sorted_multiset(rec)),
                           nodes=info["nodes"], linear=info["linear"], deleted=info["deleted"], single=info["single"],
                           count_sum=info["count_sum"], set_info=sinfo.tolist(), kmerfreq=freq.tolist())
        print(name, info["nodes"], info["linear"], info["deleted"])
    with open(os.path.join(HERE, "dumps.json"), "w") as f:
        json.dump(dumps, f, indent=1)


if __name__ == "__main__":
    main()
