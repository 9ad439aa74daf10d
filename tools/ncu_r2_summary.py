"""Round-2 profile summaries.
  python tools/ncu_r2_summary.py traffic gpurun_out/r2_traffic_c2.csv <instances per step>
      per-kernel DRAM bytes / time / instructions from an `ncu --metrics ... --csv` log of ONE step on the real config
      -> profiles/r2_traffic_sliced.json, profiles/r2_traffic_c2.md
  python tools/ncu_r2_summary.py launches gpurun_out/r2_launches.csv -> shares per kernel (stdout)
  python tools/ncu_r2_summary.py full gpurun_out/r2_skm.ncu-rep <instances> -> profiles/r2_sliced_ncu.md"""
import csv, io, json, subprocess, sys
from collections import OrderedDict

SC = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Tbyte': 1e12, 'Kbyte': 1e3, 'byte': 1}
TM = {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3, 'msecond': 1, 'usecond': 1e-3, 'nsecond': 1e-6, 'second': 1e3}


def short(name):
    return name.split('(')[0].replace('void ', '').replace('sdt::', '').replace('(anonymous namespace)::', '')


def read_metric_csv(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {n: i for i, n in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows[1:]:
        k = (r[ix['ID']], short(r[ix['Kernel Name']]))
        launches.setdefault(k, {})[r[ix['Metric Name']]] = (r[ix['Metric Unit']], float(r[ix['Metric Value']].replace(',', '')))
    return launches


mode = sys.argv[1]
if mode in ('traffic', 'launches'):
    L = read_metric_csv(sys.argv[2])
    agg = OrderedDict()
    for (i, name), m in L.items():
        a = agg.setdefault(name, dict(n=0, ms=0.0, rd=0.0, wr=0.0, inst=0.0, issue=[]))
        a['n'] += 1
        u, v = m['gpu__time_duration.sum']
        a['ms'] += v * TM.get(u, 1e-6)
        if 'dram__bytes_read.sum' in m:
            a['rd'] += m['dram__bytes_read.sum'][1] * SC[m['dram__bytes_read.sum'][0]]
            a['wr'] += m['dram__bytes_write.sum'][1] * SC[m['dram__bytes_write.sum'][0]]
            a['inst'] += m['smsp__inst_executed.sum'][1]
            a['issue'].append((m['smsp__issue_active.avg.pct_of_peak_sustained_active'][1], v))
    tot = sum(a['ms'] for a in agg.values())
    if mode == 'launches':
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1]['ms']):
            print(f"{100 * a['ms'] / tot:5.1f} %  {a['ms']:8.2f} ms  {a['n']:4d} launches  {n}")
        sys.exit(0)
    inst = float(sys.argv[3])
    totb = sum(a['rd'] + a['wr'] for a in agg.values())
    out = ["# Round 2 — DRAM traffic and instruction counts of the sliced insert on the REAL C2 config (one step, ncu)", "",
           "Command (under gpurun, one B200): `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,"
           "smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:skm_|chain_|slice_scan python bench.py --steps 1 --warmup 0 "
           "--no-e2e --no-cpu-baseline --no-parity` (C2: 50 M reads, 3.5e9 instances, no capacity hint).  Times under ncu are serialised and cold-cache; "
           "the bench line has the CUDA-event times.", "",
           "| kernel | launches | ms | share | DRAM read | DRAM written | DRAM B / instance | warp instr / instance | issue slots busy |", "|---|---|---|---|---|---|---|---|---|"]
    for n, a in agg.items():
        iss = sum(p * t for p, t in a['issue']) / max(sum(t for p, t in a['issue']), 1e-9) if a['issue'] else 0
        out.append(f"| `{n}` | {a['n']} | {a['ms']:.2f} | {100 * a['ms'] / tot:.1f} % | {a['rd'] / 1e9:.2f} GB | {a['wr'] / 1e9:.2f} GB | {(a['rd'] + a['wr']) / inst:.2f} | {a['inst'] / inst:.2f} | {iss:.0f} % |")
    out += ["", f"**DRAM traffic of the whole insert = {totb / 1e9:.1f} GB per step = {totb / inst:.1f} B per instance** (round 1, with the separate scatter pass: 23.8 B; "
            "the SURVEY §8d figure for a table in HBM: 64 B).  Every byte is sequential or whole-sector; nothing in the pipeline is bound by DRAM bandwidth."]
    open("profiles/r2_traffic_c2.md", "w").write("\n".join(out) + "\n")
    json.dump({"kernels": list(agg), "instances_per_step": inst, "dram_bytes_per_step": totb, "dram_bytes_per_instance": totb / inst, "key_words": 1,
               "source": "ncu dram__bytes_read/write.sum over one step of bench.py on the real C2 config (profiles/r2_traffic_c2.md)"},
              open("profiles/r2_traffic_sliced.json", "w"), indent=1)
    print(totb / inst, tot)
else:
    rep, inst = sys.argv[2], float(sys.argv[3])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
    agg, order = {}, []
    for v in rows[2:]:
        m = dict(zip(hdr, zip(units, v)))
        name = short(m['Kernel Name'][1])
        g = lambda k: float(m[k][1].replace(',', ''))
        a = agg.get(name)
        if a is None:
            a = agg[name] = dict(launches=0, ms=0.0, rd=0.0, wr=0.0, inst=0.0, issue=[], warps=[], regs=m['launch__registers_per_thread'][1],
                                 block=m['launch__block_size'][1], smem=m['launch__shared_mem_per_block_dynamic'], stalls={}, ipc=[])
            order.append(name)
        a['launches'] += 1
        a['ms'] += g('gpu__time_duration.sum') * TM[m['gpu__time_duration.sum'][0]]
        a['rd'] += g('dram__bytes_read.sum') * SC[m['dram__bytes_read.sum'][0]]
        a['wr'] += g('dram__bytes_write.sum') * SC[m['dram__bytes_write.sum'][0]]
        a['inst'] += g('smsp__inst_executed.sum')
        a['issue'].append(g('smsp__issue_active.avg.pct_of_peak_sustained_active'))
        a['warps'].append(g('sm__warps_active.avg.pct_of_peak_sustained_active'))
        for k, x in m.items():
            if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued'):
                a['stalls'][k[33:]] = a['stalls'].get(k[33:], 0) + float(x[1].replace(',', ''))
    tot_ms = sum(a['ms'] for a in agg.values())
    out = ["# Round 2 — ncu `--set full` capture of the sliced insert's three hot kernels", "",
           "Command (under gpurun, one B200): `ncu --set full --clock-control none --import-source on -k regex:skm_emit|skm_merge|skm_build -s 2 -c 5 "
           "python bench.py --pairs 4000000 --transcripts 3200 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity`", "",
           f"Captured: {inst:.3g} k-mer instances per step (C2 reads at C2's coverage on a smaller transcriptome, so that the ~40 replays of every kernel fit "
           "the box's time; the DRAM traffic of the REAL C2 config is in `r2_traffic_c2.md`).  Only some of the step's emit launches are in the capture "
           "(`-s 2 -c 5`), so shares are per captured launch, not per step.", "",
           "| kernel | launches | ms | DRAM read | DRAM written | warp instr (G) | issue slots busy | warps active | regs | block | dyn. smem |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for n in order:
        a = agg[n]
        out.append(f"| `{n}` | {a['launches']} | {a['ms']:.2f} | {a['rd'] / 1e9:.2f} GB | {a['wr'] / 1e9:.2f} GB | {a['inst'] / 1e9:.2f} | "
                   f"{sum(a['issue']) / len(a['issue']):.0f} % | {sum(a['warps']) / len(a['warps']):.0f} % | {a['regs']} | {a['block']} | {a['smem'][1]} {a['smem'][0]} |")
    out += ["", "## Warp stall reasons (pc sampling, share of the kernel's samples)", ""]
    for n in order:
        st = agg[n]['stalls']; t = sum(st.values()) or 1
        out.append(f"* `{n}`: " + ", ".join(f"{k} {100 * x / t:.0f} %" for k, x in sorted(st.items(), key=lambda kv: -kv[1])[:6]))
    open("profiles/r2_sliced_ncu.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out))
