"""Shared-memory wavefronts per source line from an ncu report (bank conflicts show as excessive wavefronts):
python tools/ncu_smem_lines.py report.ncu-rep kernel-regex [top-n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
fname, hdr, data = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0]:
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        d["file"] = fname; data.append(d)
f = lambda d, k: float(d.get(k) or 0)
tot = sum(f(d, "L1 Wavefronts Shared") for d in data) or 1
ideal = sum(f(d, "L1 Wavefronts Shared Ideal") for d in data)
print("shared wavefronts %.4g, ideal %.4g (%.2fx)" % (tot, ideal, tot / max(ideal, 1)))
for d in sorted(data, key=lambda d: -f(d, "L1 Wavefronts Shared"))[:n]:
    print("%5.1f%% wavefronts  x%4.1f of ideal | %s:%s %s" % (100 * f(d, "L1 Wavefronts Shared") / tot,
          f(d, "L1 Wavefronts Shared") / max(f(d, "L1 Wavefronts Shared Ideal"), 1), d["file"], d["Line No"], d["Source"].strip()[:100]))
