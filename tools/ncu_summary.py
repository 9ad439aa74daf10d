"""Turns an `ncu --set full` capture of the insert kernel into profiles/<round>_insert_reads_ncu.md and
profiles/<round>_traffic.json.   usage: python tools/ncu_summary.py gpurun_out/prof_r1d.ncu-rep r1 4194304 70"""
import csv, io, json, subprocess, sys
rep, rnd, reads, nwin = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
g = lambda k: float(m[k][1])
inst = reads * nwin
kernel = m.get("Kernel Name", ("", "insert_reads_kernel"))[1]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__sectors_read.sum', 'dram__sectors_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_requests_srcunit_tex_op_read.sum',
        'lts__t_requests_srcunit_tex_op_atom_dot_cas.sum', 'lts__t_requests_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum.per_cycle_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Tbyte': 1e12, 'Kbyte': 1e3, 'byte': 1}[m['dram__bytes_read.sum'][0]]
tr = (g('dram__bytes_read.sum') + g('dram__bytes_write.sum')) * scale
out = [f"# Round {rnd[1:]} — ncu `--set full` capture of the dominant kernel", "", f"Kernel: `{kernel[:120]}`", "",
       f"Command (under gpurun, one B200): `ncu --set full --clock-control none --import-source on -k regex:insert_reads -s 2 -c 1 -o {rep} python bench.py --pairs 4000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline`",
       "", f"One launch = {reads} reads x {nwin} windows = {inst} instances (K=31, C2 reads, 17 % of them new keys, table load <= 0.5, 32-byte slots).",
       "Times under ncu are cold-cache and serialised; the bench's own CUDA-event time per launch is in the bench line (`roofline.kernel_ms_per_launch`).", "",
       "| metric | value | per instance |", "|---|---|---|"]
for k in keys:
    if k in m:
        u, v = m[k]; per = ""
        try:
            if k.endswith('.sum') and 'duration' not in k and 'bytes' not in k: per = f"{float(v) / inst:.3f}"
            if 'bytes' in k and k.endswith('.sum'): per = f"{float(v) * scale / inst:.1f} B"
        except ValueError:
            pass
        out.append(f"| `{k}` | {v} {u} | {per} |")
out += ["", f"**DRAM traffic = {tr / 1e9:.2f} GB per launch = {tr / inst:.1f} B per instance** (algorithmic figure: 64 B = one 32-byte sector read + written back).",
        "Reads: every slot load carries `.L2::64B` (SASS `LDG.E.ENL2.LTC64B.256`), so a miss fills 64 bytes; without the qualifier the B200 L2 fills the whole 128-byte line (127 B of DRAM reads per random load, `tools/ldvar_bench.cu`) and the same kernel moved 144 B per instance. Writes: one dirty 32-byte sector per instance.",
        "L2 requests per instance: ~1.04 loads (= probes) + ~1.3 CAS (payload CAS128 + key claim for the 17 % new keys + retries) — the kernel's cost is this request count: B200 completes requests to cold lines of a >L2 table at 36.65 G/s whatever their kind (`profiles/r1_random_access_findings.md`).",
        "", "## Warp stall reasons (pc sampling, share of all samples)", ""]
st = {k: v for k, v in m.items() if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued')}
tot = sum(float(v[1]) for v in st.values())
for k, v in sorted(st.items(), key=lambda kv: -float(kv[1][1]))[:8]:
    out.append(f"* {k.replace('smsp__pcsamp_warps_issue_stalled_', '')}: {100 * float(v[1]) / tot:.1f} %")
rows = list(csv.reader(io.StringIO(src))); h = rows[1]; idx = {x: i for i, x in enumerate(h)}; data = rows[2:]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
out += ["", "## Hottest SASS instructions (source page)", "", "| samples | share | instruction | dominant stall |", "|---|---|---|---|"]
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:12]:
    s = int(r[idx['# Samples']]); stl = {k: int(r[idx[k]] or 0) for k in ['stall_long_sb', 'stall_mio', 'stall_lg', 'stall_barrier', 'stall_short_sb', 'stall_wait', 'stall_branch_resolving']}
    out.append(f"| {s} | {100 * s / tot:.1f} % | `{r[idx['Source']].strip()[:80]}` | {max(stl, key=stl.get)} |")
out += ["", "Reading: warps wait (long scoreboard) right after the 256-bit slot load and after the 128-bit payload CAS, and at the `__syncthreads()` that separate the chop phase from the queue drain; issue slots are mostly idle. The kernel is bound by the rate of L2 requests to cold lines, not by instructions or DRAM bandwidth."]
open(f"profiles/{rnd}_insert_reads_ncu.md", "w").write("\n".join(out) + "\n")
json.dump({"kernel": kernel[:100], "key_words": 1, "instances_per_launch": inst, "dram_bytes_per_launch": tr, "dram_bytes_per_instance": tr / inst,
           "source": f"ncu --set full capture {rep} (profiles/{rnd}_insert_reads_ncu.md)"}, open(f"profiles/{rnd}_traffic.json", "w"), indent=1)
print(kernel[:80], tr / inst, g('gpu__time_duration.sum'))
