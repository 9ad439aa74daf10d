"""End-to-end drop-in parity on the GPU: the reference binary with its hashing stage replaced by
host/prlHashReads_gpu.c + libsdtgpu.so (oracle/_ref/SOAPdenovo-Trans-*-gpu) must write the same
`pregraph` outputs as the stock binary, byte for byte: .kmerFreq, .edge.gz (decompressed),
.preArc, .vertex, .preGraphBasic — i.e. the k-mer table was handed back in the layout
node2edge.c / cutTipPreGraph.c / prlRead2path.c consume, including the order-dependent pruning."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import make_dataset

pytestmark = pytest.mark.gpu


def _run(exe, cfg, prefix, K, p, d, extra=(), devices=None, sliced_hint=0, direct=False):
    cmd = [exe, "pregraph", "-s", cfg, "-K", str(K), "-p", str(p), "-d", str(d), "-o", prefix, *extra]
    env = dict(os.environ)
    if devices:
        env["SDTGPU_DEVICES"] = devices
    if sliced_hint:
        env["SDTGPU_CAPACITY_HINT"] = str(sliced_hint)
    if direct:
        env["SDTGPU_DIRECT"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def _outputs(prefix):
    out = {}
    for ext in ("kmerFreq", "preArc", "vertex", "preGraphBasic"):
        with open(f"{prefix}.{ext}", "rb") as f:
            out[ext] = f.read()
    with gzip.open(f"{prefix}.edge.gz", "rb") as f:
        out["edge"] = f.read()
    return out


@pytest.mark.parametrize("direct", [False, True], ids=["sliced", "direct"])
@pytest.mark.parametrize("build,K,p,d,fastq,extra", [
    ("31mer", 25, 8, 0, False, ()),
    ("31mer", 31, 3, 2, True, ()),
    ("127mer", 63, 8, 1, False, ()),
    ("127mer", 127, 4, 0, False, ()),
    ("31mer", 25, 8, 0, False, ("-n",)),
])
def test_pregraph_outputs_identical(pkg, oracle, tmp_path, build, K, p, d, fastq, extra, direct):
    """The drop-in's default is the sliced build without any capacity hint; SDTGPU_DIRECT=1 is the single-pass insert."""
    stock = os.path.join(oracle.REF_DIR, f"SOAPdenovo-Trans-{build}")
    gpu = os.path.join(oracle.REF_DIR, f"SOAPdenovo-Trans-{build}-gpu")
    if not (os.path.exists(stock) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref binaries not present")
    L = 150 if K > 63 else 100
    tr = pkg.synth.make_transcriptome(60, 21)
    reads, lens = make_dataset(pkg, tr, 12000, L, 33 + K, ragged=30, n_rate=0.003 if extra else 0)
    cfg = pkg.synth.write_library(str(tmp_path / "in"), reads, lens, L, paired=True, fastq=fastq)
    a = _run(stock, cfg, str(tmp_path / "ref"), K, p, d, extra)
    b = _run(gpu, cfg, str(tmp_path / "gpu"), K, p, d, extra, direct=direct)
    assert "GPU pregraph hashing" in b and "GPU pregraph hashing" not in a
    ra, rb = _outputs(str(tmp_path / "ref")), _outputs(str(tmp_path / "gpu"))
    for k in ra:
        assert ra[k] == rb[k], f"{k} differs ({len(ra[k])} vs {len(rb[k])} bytes)"
    assert len(ra["edge"]) > 1000 and len(ra["preArc"]) > 0
    # the reference's own conservation/consistency lines agree too
    for key in ("nodes allocated", "linear nodes", "kmer removed"):
        la = [l for l in a.splitlines() if key in l]
        lb = [l for l in b.splitlines() if key in l]
        assert la == lb, (la, lb)


@pytest.mark.parametrize("build,K,p,d", [("31mer", 31, 8, 1), ("127mer", 63, 5, 0)])
def test_pregraph_outputs_identical_sliced_build(pkg, oracle, tmp_path, build, K, p, d):
    """SDTGPU_CAPACITY_HINT: the sliced build with a hint (coarser chains); the hand-back and every pregraph
    output stay byte-identical."""
    stock = os.path.join(oracle.REF_DIR, f"SOAPdenovo-Trans-{build}")
    gpu = os.path.join(oracle.REF_DIR, f"SOAPdenovo-Trans-{build}-gpu")
    if not (os.path.exists(stock) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref binaries not present")
    tr = pkg.synth.make_transcriptome(60, 21)
    reads, lens = make_dataset(pkg, tr, 12000, 100, 5 + K, ragged=30)
    cfg = pkg.synth.write_library(str(tmp_path / "in"), reads, lens, 100, paired=True)
    a = _run(stock, cfg, str(tmp_path / "ref"), K, p, d)
    b = _run(gpu, cfg, str(tmp_path / "gpu"), K, p, d, sliced_hint=4_000_000)
    ra, rb = _outputs(str(tmp_path / "ref")), _outputs(str(tmp_path / "gpu"))
    for k in ra:
        assert ra[k] == rb[k], f"{k} differs ({len(ra[k])} vs {len(rb[k])} bytes)"
    for key in ("nodes allocated", "linear nodes", "kmer removed"):
        pick = lambda t: [l for l in t.splitlines() if key in l and not l.startswith("time spent")]     # (wall-clock lines differ)
        assert pick(a) == pick(b)


@pytest.mark.parametrize("devices", ["0,0,0", "0,1"])
def test_pregraph_outputs_identical_sharded(pkg, oracle, tmp_path, devices):
    """SDTGPU_DEVICES: every table shard receives every batch and keeps the k-mers it owns; the
    shards' nodes are merged at hand-back.  "0,0,0" runs three shards on one GPU (always possible),
    "0,1" two GPUs.  Outputs must still equal the stock binary's byte for byte."""
    import torch
    if devices == "0,1" and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    stock = os.path.join(oracle.REF_DIR, "SOAPdenovo-Trans-31mer")
    gpu = os.path.join(oracle.REF_DIR, "SOAPdenovo-Trans-31mer-gpu")
    if not (os.path.exists(stock) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref binaries not present")
    tr = pkg.synth.make_transcriptome(60, 21)
    reads, lens = make_dataset(pkg, tr, 12000, 100, 71, ragged=30)
    cfg = pkg.synth.write_library(str(tmp_path / "in"), reads, lens, 100, paired=True)
    a = _run(stock, cfg, str(tmp_path / "ref"), 27, 8, 1)
    b = _run(gpu, cfg, str(tmp_path / "gpu"), 27, 8, 1, devices=devices)
    assert f"on {len(devices.split(','))} device(s)" in b
    ra, rb = _outputs(str(tmp_path / "ref")), _outputs(str(tmp_path / "gpu"))
    for k in ra:
        assert ra[k] == rb[k], f"{k} differs"
    for key in ("nodes allocated", "linear nodes", "kmer removed"):
        pick = lambda t: [l for l in t.splitlines() if key in l and not l.startswith("time spent")]     # (wall-clock lines differ)
        assert pick(a) == pick(b)


def _compare(a, b, ra, rb):
    for k in ra:
        assert ra[k] == rb[k], f"{k} differs ({len(ra[k])} vs {len(rb[k])} bytes)"
    for key in ("nodes allocated", "linear nodes", "kmer removed", "kmer in reads"):
        pick = lambda t: [l for l in t.splitlines() if key in l and not l.startswith("time spent")]     # (wall-clock lines differ)
        assert pick(a) == pick(b)


def test_config_c1_as_stated(pkg, oracle, tmp_path):
    """BASELINE.json config 1 as stated: SOAPdenovo-Trans-31mer pregraph K=25 on 1 M synthetic 100 bp PE reads from 2 000
    random transcripts, -p 8 — the stock binary against the GPU drop-in (sliced build, no hint), all five outputs byte
    for byte, and the reference's counters (76 000 000 k-mers in reads)."""
    stock = os.path.join(oracle.REF_DIR, "SOAPdenovo-Trans-31mer")
    gpu = os.path.join(oracle.REF_DIR, "SOAPdenovo-Trans-31mer-gpu")
    if not (os.path.exists(stock) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref binaries not present")
    cfg_d = pkg.synth.CONFIGS["C1"]
    tr = pkg.synth.make_transcriptome(cfg_d["n_transcripts"], cfg_d["seed"])
    parts = [pkg.synth.make_reads(tr, 250_000, 100, cfg_d["seed"], first_pair=a)[0] for a in range(0, cfg_d["n_pairs"], 250_000)]
    reads = np.concatenate(parts)
    lens = np.full(len(reads), 100, np.uint32)
    cfg = pkg.synth.write_library(str(tmp_path / "in"), reads, lens, 100, paired=True)
    a = _run(stock, cfg, str(tmp_path / "ref"), 25, 8, 0)
    b = _run(gpu, cfg, str(tmp_path / "gpu"), 25, 8, 0)
    assert "76000000 kmer in reads" in a.replace(",", "") and "GPU pregraph hashing" in b
    _compare(a, b, _outputs(str(tmp_path / "ref")), _outputs(str(tmp_path / "gpu")))


def test_config_c5_flavour_d2_with_tip_pruning(pkg, oracle, tmp_path):
    """Config 5's shape through the drop-in: a few transcripts at 10^5 x depth (hot k-mers, slices that overflow by far
    and are cut into sub-slices), -d 2, and the order-dependent pruning after the hand-back."""
    stock = os.path.join(oracle.REF_DIR, "SOAPdenovo-Trans-31mer")
    gpu = os.path.join(oracle.REF_DIR, "SOAPdenovo-Trans-31mer-gpu")
    if not (os.path.exists(stock) and os.path.exists(gpu)):
        pytest.skip("oracle/_ref binaries not present")
    tr = pkg.synth.make_transcriptome(200, 20261021, hot=3)
    reads, lens = make_dataset(pkg, tr, 150_000, 100, 77)
    cfg = pkg.synth.write_library(str(tmp_path / "in"), reads, lens, 100, paired=True)
    a = _run(stock, cfg, str(tmp_path / "ref"), 31, 8, 2)
    b = _run(gpu, cfg, str(tmp_path / "gpu"), 31, 8, 2)
    _compare(a, b, _outputs(str(tmp_path / "ref")), _outputs(str(tmp_path / "gpu")))
