"""Config C1 end to end through the reference CLI: stock binary vs the GPU drop-in (wall clock of
`pregraph`, the hashing stage's own stdout timings, byte-identical outputs)."""
import gzip, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdt_pkg
pkg = sdt_pkg.load(); synth = pkg.synth
from oracle import oracle as O
cfg = synth.CONFIGS["C1"]
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n_pairs"]
tr = synth.make_transcriptome(cfg["n_transcripts"], cfg["seed"])
parts = [synth.make_reads(tr, min(250000, n_pairs - a), 100, cfg["seed"], first_pair=a)[0] for a in range(0, n_pairs, 250000)]
import numpy as np
reads = np.concatenate(parts); lens = np.full(len(reads), 100, np.uint32)
d = tempfile.mkdtemp()
c = synth.write_library(os.path.join(d, "in"), reads, lens, 100)
res = {}
for tag, exe in (("stock", "SOAPdenovo-Trans-31mer"), ("gpu", "SOAPdenovo-Trans-31mer-gpu")):
    env = dict(os.environ, SDTGPU_CAPACITY_HINT="17000000")
    t0 = time.time()
    out = subprocess.run([os.path.join(O.REF_DIR, exe), "pregraph", "-s", c, "-K", "25", "-p", str(min(os.cpu_count(), 8)), "-d", "0", "-o", os.path.join(d, tag)],
                         capture_output=True, text=True, env=env).stdout
    res[tag] = time.time() - t0
    print(f"== {tag}: pregraph wall {res[tag]:.2f} s")
    for l in out.splitlines():
        if "time spent" in l or "nodes allocated" in l or "GPU" in l:
            print("   ", l)
same = all(open(f"{d}/stock.{e}", "rb").read() == open(f"{d}/gpu.{e}", "rb").read() for e in ("kmerFreq", "preArc", "vertex", "preGraphBasic"))
same &= gzip.open(f"{d}/stock.edge.gz").read() == gzip.open(f"{d}/gpu.edge.gz").read()
print("outputs byte-identical:", same)
