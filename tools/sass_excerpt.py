"""profiles/r2_sass.md: static SASS evidence of the hot kernels (no GPU needed).
python tools/sass_excerpt.py > profiles/r2_sass.md"""
import collections, re, subprocess
LIB = "soapdenovo-trans_b200/libsdtgpu.so"
KERNELS = [("skm_emit_kernelILi1ELb0", "skm_emit_kernel<1, false> (reads -> super-k-mer records appended to their slice's chain)"),
           ("skm_merge_kernelILi1ELb0", "skm_merge_kernel<1, false> (copies merged, chains -> work items)"),
           ("skm_build2_kernelILi1ELb1", "skm_build2_kernel<1, true> (the slice build, 32-bit ordinals)"),
           ("skm_append_kernelILi8", "skm_append_kernel<8> (multi-GPU: received records -> chains)")]
INTEREST = re.compile(r"\b(ATOMG|ATOMS|ATOM|REDG|RED|LDGSTS|LDGDEPBAR|DEPBAR|BAR|MATCH|SHFL|VOTE|VOTEU|REDUX|LDG|STG|LDS|STS|LDL|STL|WARPSYNC|NANOSLEEP|BREV|SHF|IMAD|VIMNMX|VMIN)[.\w]*")
SHOW = re.compile(r"ATOMG|ATOMS\.(CAS|MIN|CAST)|LDGSTS|STG\.E\.ENL2|LDG\.E\.(NA\.)?128|REDG|NANOSLEEP")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line)
print("# Round 2 — SASS of the hot kernels (`cuobjdump -sass soapdenovo-trans_b200/libsdtgpu.so`, sm_100a)\n")
print("Static counts of the instructions that show how each kernel touches memory and synchronises, and the lines that carry the design: "
      "the 64-bit ticket atomic of the chain append (`ATOMG.E.ADD.64`), `cp.async` record staging in the build (`LDGSTS`), native 32-bit "
      "`ATOMS.MIN` for ordinals, `ATOMS.CAST.SPIN.64` for the key claim, 128-bit shared-memory accesses of the swizzled staging area in the "
      "merge (`LDS.128`), 256-bit node stores (`STG.E.ENL2.256`).  Regenerate with `python tools/sass_excerpt.py > profiles/r2_sass.md`.\n")
for key, title in KERNELS:
    names = [n for n in funcs if key in n]
    if not names:
        print(f"## `{title}` — not found\n"); continue
    body = funcs[names[0]]
    cnt = collections.Counter()
    for l in body:
        ins = re.sub(r"/\*[0-9a-f]+\*/", "", l).strip()
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        m = INTEREST.match(ins)
        if m and not m.group(0).startswith(("IMAD", "SHF", "BREV", "VIMNMX")):
            cnt[m.group(0)] += 1
    print(f"## `{title}` — {len(body)} instructions\n")
    print(", ".join(f"`{k}` x{v}" for k, v in sorted(cnt.items())) + "\n")
    print("```")
    shown = 0
    for l in body:
        if SHOW.search(l) and shown < 12:
            m = re.match(r"\s+(/\*[0-9a-f]+\*/)\s+(.*?);", l)
            if m:
                print(m.group(1), "", m.group(2)); shown += 1
    print("```\n")
