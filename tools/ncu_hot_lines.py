"""Per-source-line hot spots from an ncu report (needs -lineinfo and --import-source on):
python tools/ncu_hot_lines.py report.ncu-rep kernel-regex [top-n] [launch-skip]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
fname, hdr, data = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0]:      # source-line rows (SASS rows have an empty line number)
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        d["file"] = fname; data.append(d)
f = lambda d, k: float(d.get(k) or 0)
tot = sum(f(d, "Instructions Executed") for d in data) or 1
tots = sum(f(d, "# Samples") for d in data) or 1
print("total warp instructions %.4g, samples %d" % (tot, tots))
for d in sorted(data, key=lambda d: -f(d, sys.argv[5] if len(sys.argv) > 5 else "# Samples"))[:n]:
    ie = f(d, "Instructions Executed") or 1
    print("%5.1f%% smp %5.1f%% inst  thr %4.1f | %s:%s %s" % (100 * f(d, "# Samples") / tots, 100 * f(d, "Instructions Executed") / tot,
          f(d, "Thread Instructions Executed") / ie, d["file"], d["Line No"], d["Source"].strip()[:95]))
