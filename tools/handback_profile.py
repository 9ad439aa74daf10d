"""Where does the hand-back time go? (config C1: 1 M reads, 15.9 M nodes)"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdt_pkg
pkg = sdt_pkg.load(); synth = pkg.synth
import torch
cfg = synth.CONFIGS["C1"]
tr = synth.make_transcriptome(cfg["n_transcripts"], cfg["seed"])
dev = torch.device("cuda", 0)
tr_dev = dict(bases=torch.from_numpy(tr.bases).to(dev), starts=torch.from_numpy(tr.starts.astype(np.int64)).to(dev),
              lengths=torch.from_numpy(tr.lengths.astype(np.int32)).to(dev), cum=torch.from_numpy(tr.cum.astype(np.int64)).to(dev), n=len(tr.lengths))
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n_pairs"]
d = torch.empty((2 * n_pairs, 28), dtype=torch.uint8, device=dev)
pkg.pregraph.synth_reads_device(tr_dev, cfg["seed"], 0, n_pairs, 100, 28, d); torch.cuda.synchronize()
g = pkg.PregraphGPU(25, 1, 100, capacity_hint=17_000_000 * max(1, n_pairs // 500_000))
t0 = time.perf_counter(); g.push_reads(d, None, None, n_reads=2 * n_pairs, uniform_len=100, stride_bytes=28, device=True); g.sync(); t1 = time.perf_counter()
freq, st = g.finalize(0); t2 = time.perf_counter()
nodes = g.export_nodes(8); t3 = time.perf_counter()
sets = (C.POINTER(pkg.pregraph.KmerSet) * 8)()
rc = pkg.library().sdtgpu_build_kmersets(nodes.ctypes.data, len(nodes), 1, 8, None, sets); t4 = time.perf_counter()
pkg.library().sdtgpu_free_kmersets(sets, 8)
t5 = time.perf_counter(); rec, info = g.export_kmersets(8); t6 = time.perf_counter()
print(f"nodes {st.n_nodes}: insert {t1-t0:.3f} s, finalize {t2-t1:.3f} s, export_nodes (device compaction + D2H) {t3-t2:.3f} s, "
      f"build_kmersets (host replay, 8 threads) {t4-t3:.3f} s, export_kmersets incl. python decode {t6-t5:.3f} s, host cores {os.cpu_count()}")
