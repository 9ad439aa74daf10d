#!/usr/bin/env python
"""bench.py — k-mers inserted/s in pregraph hashing (BASELINE.json metric), 1..8 B200.

One "step" = one pass of the hot path over the whole synthetic workload: empty the table, then
chop + insert every window of every read (sdtgpu_push_reads_device; at N > 1 bucket by owner ->
NCCL all-to-all -> insert).  `value` times the step with the packed reads already resident in HBM;
`e2e` times the same job through the C ABI with HOST buffers (pinned), H2D copies and the D2H read
of the result counters inside the timed region.  `roofline` is the insert kernel's algorithmic
bytes (64 B per instance for K <= 63, 96 B for K <= 127) over its CUDA-event time
against the measured HBM copy bandwidth in MEASURED_PEAKS.json.  `cpu_baseline` times the
unmodified reference's prlRead2HashTable (oracle/_ref) on a bounded sample of the same reads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--pairs P] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_INSTANCE = {1: 64, 2: 64, 4: 96}
# Measured on this pool's B200 with tools/randacc_bench.cu (profiles/r1_randacc_bench.txt): requests to
# cold lines of a table >> L2 complete at 36.65 G/s no matter their kind (load, CAS, RED) or width; an
# upsert needs at least one load and one atomic, so a single-pass insert cannot exceed half of that.
RANDOM_REQUESTS_PER_S = 36.65e9
MIN_REQUESTS_PER_INSTANCE = {1: 2, 2: 2, 4: 3}
DEFAULT_PATH = {1: "sliced", 2: "sliced"}     # key: 1 GPU / more than one GPU (super-k-mer exchange)
METRIC = "k-mers inserted/s in pregraph hashing"
UNIT = "k-mer instances/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def unpack_reads(packed: np.ndarray, read_len: int) -> np.ndarray:
    b = packed[:, : (read_len + 3) // 4]
    out = np.empty((b.shape[0], b.shape[1] * 4), dtype=np.uint8)
    out[:, 0::4], out[:, 1::4], out[:, 2::4], out[:, 3::4] = b >> 6, (b >> 4) & 3, (b >> 2) & 3, b & 3
    return out[:, :read_len]


def reference_run(cfg_d: dict, sample_reads: np.ndarray, threads: int):
    """Unmodified reference prlRead2HashTable (oracle/_ref/ref_hash_*) on `sample_reads` (base codes)."""
    from oracle import oracle as O
    import sdt_pkg
    synth = sdt_pkg.load().synth
    lens = np.full(len(sample_reads), sample_reads.shape[1], dtype=np.uint32)
    with tempfile.TemporaryDirectory() as d:
        cfg = synth.write_library(d, sample_reads, lens, sample_reads.shape[1], paired=True)
        kw = 1 if cfg_d["key_words"] == 1 else 4
        info, _, _ = O.run_reference(cfg, os.path.join(d, "out"), cfg_d["K"], kw, threads, 0, dump=False)
    return info


def host_sample(cfg_d: dict, n_pairs: int) -> np.ndarray:
    """The first n_pairs pairs of the workload, generated on the host (bit-identical to the device generator)."""
    import sdt_pkg
    synth = sdt_pkg.load().synth
    tr = synth.make_transcriptome(cfg_d["n_transcripts"], cfg_d["seed"], hot=cfg_d["hot"])
    out = []
    for a in range(0, n_pairs, 250_000):
        reads, _ = synth.make_reads(tr, min(250_000, n_pairs - a), cfg_d["read_len"], cfg_d["seed"], first_pair=a)
        out.append(reads)
    return np.concatenate(out)


def ref_threads() -> int:
    return max(1, min(os.cpu_count() or 8, 64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--pairs", type=int, default=0, help="override the config's read-pair count (testing)")
    ap.add_argument("--transcripts", type=int, default=0, help="override the config's transcript count (testing: keeps the coverage of the full workload at a smaller size)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--batch-reads", type=int, default=1 << 22)
    ap.add_argument("--exchange", default="reads", choices=["reads", "records", "reads-replicated"],
                    help="multi-GPU sharding: all-gather the packed reads and insert owned k-mers (default) or exchange k-mer records")
    ap.add_argument("--partitioned", action="store_true", help="experimental staged/partitioned insert path")
    ap.add_argument("--path", default="auto", choices=["auto", "direct", "sliced", "partitioned"],
                    help="insert path: single-pass upsert (direct), sliced build (super-k-mer records -> slices built in shared memory), "
                         "or the experimental staged path; auto = the fastest measured one for this GPU count")
    args = ap.parse_args()
    if args.partitioned:
        args.path = "partitioned"
    import sdt_pkg
    pkg = sdt_pkg.load()
    synth = pkg.synth
    cfg_d = dict(synth.CONFIGS[args.config])
    if args.path == "auto":
        # the sliced build is tuned for 1-word keys (K <= 31, the metric's config); at K = 63 / 127 super-k-mers are
        # 2-4x longer, the slices lumpier (11 % of them overflow on C3) and the single-pass insert is faster
        args.path = DEFAULT_PATH[1 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 2] if cfg_d["key_words"] == 1 and cfg_d["K"] <= 31 and not cfg_d["hot"] else "direct"
    args.partitioned = args.path == "partitioned"
    sliced = args.path == "sliced"
    if args.pairs:
        cfg_d["n_pairs"] = args.pairs
    if args.transcripts:
        cfg_d["n_transcripts"] = args.transcripts
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, kw, L = cfg_d["K"], cfg_d["key_words"], cfg_d["read_len"]
    nwin = L - K + 1
    workload = (f"{args.config}: {'31mer' if kw == 1 else '127mer'} build K={K}, {2 * cfg_d['n_pairs']} synthetic "
                f"{L}bp PE reads from {cfg_d['n_transcripts']} transcripts (seed {cfg_d['seed']})")

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        threads = ref_threads()
        sample_pairs = min(args.cpu_sample_pairs, cfg_d["n_pairs"])
        reads = host_sample(cfg_d, sample_pairs)
        vals = []
        for i in range(args.warmup + args.steps):
            info = reference_run(cfg_d, reads, threads)
            if i >= args.warmup:
                vals.append(info["count_sum"] / info["seconds"])
        v = float(np.mean(vals))
        sample = f"first {2 * sample_pairs} reads of the workload ({2 * sample_pairs * nwin} instances) per step, FASTA f1/f2 via the reference's own parser, -p {threads}"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * 2 * sample_pairs * nwin / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    real_stdout = None
    if world > 1:
        # stdout carries ONE JSON line; what libraries write to file descriptor 1 while the job runs (NCCL's
        # version banner at communicator creation) goes to stderr instead
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    hbm_gbs, peak_src = load_peaks()

    # per-rank shard of the workload (weak scaling: the config's reads PER GPU)
    n_pairs = cfg_d["n_pairs"]
    first_pair = rank * n_pairs
    n_reads = 2 * n_pairs
    stride = synth.stride_bytes(L)
    tr = synth.make_transcriptome(cfg_d["n_transcripts"], cfg_d["seed"], hot=cfg_d["hot"])
    tr_dev = dict(bases=torch.from_numpy(tr.bases).to(dev), starts=torch.from_numpy(tr.starts.astype(np.int64)).to(dev),
                  lengths=torch.from_numpy(tr.lengths.astype(np.int32)).to(dev),
                  cum=torch.from_numpy(tr.cum.astype(np.int64)).to(dev), n=len(tr.lengths))
    d_packed = torch.empty((n_reads, stride), dtype=torch.uint8, device=dev)
    pkg.pregraph.synth_reads_device(tr_dev, cfg_d["seed"], first_pair, n_pairs, L, stride, d_packed, device=local_rank)
    torch.cuda.synchronize()
    instances_rank = n_reads * nwin
    batch = args.batch_reads

    # ---- pilot pass with a generous table to learn the distinct count, then size load <= 0.5
    # distinct k-mers are dominated by error k-mers (a window is error-free with probability 0.99^K)
    slot_b = 64 if K > 63 else 32
    est_distinct = instances_rank * (1.0 - 0.99 ** K) * 1.03 + 6e7
    g = pkg.PregraphGPU(K, kw, L, capacity_hint=int(est_distinct) + 1024, device=local_rank, partitioned=args.partitioned, sliced=sliced)
    ext = torch.cuda.ExternalStream(g.stream, device=dev)

    exch, exch_kind = None, None
    if world > 1:
        from soapdenovo_trans_b200.exchange import Exchange, ReplicatedReads, SkmExchange
        if sliced and args.exchange != "reads-replicated":
            exch, exch_kind = SkmExchange(pkg, g, world, rank, dev), "skm"
        elif args.exchange == "records":
            exch = Exchange(pkg, g, world, rank, dev, max_round_instances=min(batch, n_reads) * nwin)
        else:
            exch = ReplicatedReads(pkg, g, world, rank, dev, max_round_reads=min(batch, n_reads), stride=stride)

    def one_step(gg):
        gg.reset()
        for a in range(0, n_reads, batch):
            b = min(a + batch, n_reads)
            if exch is None:
                gg.push_reads(d_packed[a:b], None, None, n_reads=b - a, uniform_len=L, stride_bytes=stride,
                              first_read_ordinal=2 * first_pair + a, device=True)
            else:
                exch.round(gg, d_packed[a:b], b - a, L, stride, 2 * first_pair + a)
        if exch is not None:
            exch.flush(gg)
        gg.sync()      # end of the step: everything staged has been inserted (flushes the epoch)

    one_step(g)
    st = g.stats()
    distinct = st.n_nodes
    assert (exch is not None) or st.n_instances == instances_rank, (st.n_instances, instances_rank)
    if world > 1:       # one geometry on all ranks (the super-k-mer exchange cuts the minimizer space by it)
        dmax = torch.tensor([distinct], dtype=torch.int64, device=dev)
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        distinct = int(dmax.item())
    g.close()
    # the library sizes the table from the hint: load 0.5 up to 60 GiB, denser beyond
    g = pkg.PregraphGPU(K, kw, L, capacity_hint=int(distinct * 1.02) + 1024, device=local_rank, partitioned=args.partitioned, sliced=sliced)
    ext = torch.cuda.ExternalStream(g.stream, device=dev)
    if exch is not None:
        exch.rebind(g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(g)
    barrier()
    g.kernel_time(reset=True)
    pkg.pregraph.debug_prof(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(ext)
    for _ in range(args.steps):
        one_step(g)
    e1.record(ext)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    ms = e0.elapsed_time(e1)
    build_prof = pkg.pregraph.debug_prof(reset=True)
    chk = g.stats()
    assert exch is not None or (chk.n_instances == instances_rank and chk.n_nodes == distinct), (chk.n_instances, chk.n_nodes)
    st = g.stats()
    phases = g.phase_times(reset=False)
    cat_ms, cat_launches = g.kernel_times(reset=False)
    insert_ms, insert_launches, all_launches = g.kernel_time(reset=True)
    t = torch.tensor([ms, float(st.n_instances), float(st.n_nodes)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, total_instances, total_nodes = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        total_instances, total_nodes = float(st.n_instances), float(st.n_nodes)
    ms_per_step = ms / args.steps
    value = total_instances / (ms_per_step * 1e-3)
    bpi = BYTES_PER_INSTANCE[st.device_key_words]
    sliced_info = None
    if sliced:
        # The insert is a pipeline of three streaming kernels (super-k-mer emit, scatter by slice, slice
        # build); the SURVEY figure of 64 (96) algorithmic bytes per instance belongs to the whole insert, so
        # `achieved` is taken over the SUM of their CUDA-event times (all on the handle's stream).  Each
        # phase's own DRAM stream (bytes it must read + write per instance) is reported beside it.
        geo = g.slice_geometry()
        rec_b = geo["record_bytes"] * geo["n_records"] / max(st.n_instances, 1)      # record bytes per instance
        read_b = stride / nwin
        node_b = st.n_nodes * slot_b / max(st.n_instances, 1)
        stream_b = {"emit": read_b + rec_b, "scatter": 2 * rec_b, "dedupe": 2 * rec_b, "build": rec_b + node_b, "scan": 0.0, "retry": 0.0}
        insert_ms = sum(phases[k][0] for k in stream_b)
        insert_launches = max(phases["build"][1] + phases["retry"][1], 1)
        sliced_info = {"geometry": geo, "windows_per_record": st.n_instances / max(geo["n_records"], 1), "phases": {
            k: {"ms_per_step": phases[k][0] / args.steps, "launches_per_step": phases[k][1] / args.steps,
                "stream_bytes_per_instance": stream_b[k],
                "stream_gbs": (st.n_instances * args.steps * stream_b[k] / max(phases[k][0], 1e-9) / 1e6) if stream_b[k] else None}
            for k in stream_b}}
        dominant = max(stream_b, key=lambda k: phases[k][0])
        sliced_info["build_phase_cycles"] = build_prof
        sliced_info["dominant"] = "skm_" + dominant + "_kernel"
        ker_ms = insert_ms / args.steps                      # one pipeline pass = one "launch" of the insert
        inst_per_launch = float(st.n_instances)
    else:
        ker_ms = max(insert_ms, 1e-9) / max(insert_launches, 1)
        inst_per_launch = st.n_instances * args.steps / max(insert_launches, 1)
    achieved = inst_per_launch * bpi / (ker_ms * 1e-3) / 1e9
    traffic = None      # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
    tp = os.path.join(ROOT, "profiles", "r1_traffic_sliced.json" if sliced else "r1_traffic.json")
    if os.path.exists(tp) and world == 1 and args.path in ("direct", "sliced"):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("key_words") == st.device_key_words:
            traffic = tj["dram_bytes_per_instance"] * inst_per_launch

    # ---- e2e: HOST (pinned) buffers, H2D inside the timed region, counters read back (D2H) every step.
    # N = 1: straight through the C ABI's host entry point (sdtgpu_push_reads).  N > 1: every rank
    # copies its round's reads from pinned host memory, then bucket -> exchange -> insert as above.
    e2e = None
    if not args.no_e2e:
        h_packed = torch.empty((n_reads, stride), dtype=torch.uint8).pin_memory()
        h_packed.copy_(d_packed)
        torch.cuda.synchronize()
        d_round = [torch.empty((min(batch, n_reads), stride), dtype=torch.uint8, device=dev) for _ in range(2)] if exch is not None else None

        def e2e_step():
            g.reset()
            for i, a in enumerate(range(0, n_reads, batch)):
                b = min(a + batch, n_reads)
                if exch is None:
                    g.push_reads(h_packed[a:b], None, None, n_reads=b - a, uniform_len=L, stride_bytes=stride,
                                 first_read_ordinal=a)
                else:
                    buf = d_round[i & 1]
                    with torch.cuda.stream(exch.aux):
                        exch.aux.wait_event(exch.inserted[exch.r & 1])      # same buffer parity as the exchange
                        buf[: b - a].copy_(h_packed[a:b], non_blocking=True)
                    exch.round(g, buf, b - a, L, stride, 2 * first_pair + a)
            if exch is not None:
                exch.flush(g)
            g.sync()
            return g.stats()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s2 = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        assert exch is not None or (s2.n_instances == instances_rank and s2.n_nodes == distinct)
        tt = torch.tensor([dt, float(s2.n_instances)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = tt.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = tt.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dt, e2e_inst = float(tm[0]), float(ts[1])
        else:
            e2e_inst = float(s2.n_instances)
        e2e = {"value": e2e_inst / dt, "unit": UNIT, "h2d_bytes_per_step": int(n_reads * stride) * world,
               "d2h_bytes_per_step": 2120 * world, "ms_per_step": dt * 1e3}

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        threads = ref_threads()
        sp = min(args.cpu_sample_pairs, n_pairs)
        sample = unpack_reads(d_packed[: 2 * sp].cpu().numpy(), L)
        info = reference_run(cfg_d, sample, threads)
        cpu = {"value": info["count_sum"] / info["seconds"], "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": f"first {2 * sp} reads ({info['count_sum']} instances, {info['nodes']} nodes) through the unmodified "
                         f"reference prlRead2HashTable (oracle/_ref), FASTA f1/f2, -p {threads}, {info['seconds']:.2f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload + (f" per GPU x {world} GPUs, k-mers sharded by owner ({'super-k-mer records of the sliced build exchanged over NCCL, one all-to-all per step' if exch_kind == 'skm' else ('packed reads all-gathered over NCCL, every rank inserts the k-mers it owns' if args.exchange != 'records' else 'k-mer records exchanged over NCCL')})" if world > 1 else ""),
                       "instances_per_step": total_instances, "distinct_kmers": total_nodes,
                       "table_slots_per_gpu": int(st.capacity), "slot_bytes": 64 if st.device_key_words == 4 else 32,
                       "batch_reads": batch,
                       "insert_path": args.path,
                       "l2": "table (>= 1.6x distinct x slot bytes), k-mer records and reads are far larger than the 126 MB L2; the table is rebuilt from empty every step"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                         "traffic": traffic, "kernel": ("skm_emit + skm_scatter + skm_dedupe + skm_build kernels (the sliced insert pipeline; times summed)" if sliced else
                                    "insert_staged_kernel" if args.partitioned else "insert_reads_kernel") if (world == 1 or args.exchange == "reads") else "insert_records_kernel",
                         "path": args.path, "sliced": sliced_info,
                         "bytes_per_instance": bpi, "kernel_ms_per_launch": ker_ms, "peak_source": peak_src,
                         "random_access": {"cold_line_requests_per_s_measured": RANDOM_REQUESTS_PER_S,
                                           "min_requests_per_instance": MIN_REQUESTS_PER_INSTANCE[st.device_key_words],
                                           "ceiling_instances_per_s_per_gpu": RANDOM_REQUESTS_PER_S / MIN_REQUESTS_PER_INSTANCE[st.device_key_words],
                                           "frac_of_ceiling": (inst_per_launch / (ker_ms * 1e-3)) / (RANDOM_REQUESTS_PER_S / MIN_REQUESTS_PER_INSTANCE[st.device_key_words]),
                                           "source": "tools/randacc_bench.cu, profiles/r1_randacc_bench.txt"},
                         "kernel_ms_per_step": {"insert": cat_ms[0] / args.steps, "partition_count": cat_ms[1] / args.steps,
                                                "partition_scatter": cat_ms[2] / args.steps}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(all_launches),
            "clocks": sampler.summary(),
        }
        sys.stdout.flush()
        if real_stdout is not None:
            os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    # orderly teardown: everything that was enqueued on the handle's streams is done, torch's views of the
    # streams and the library's buffers go first, then the handle (which owns the streams), then NCCL
    torch.cuda.synchronize()
    del ext
    exch = None
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()


if __name__ == "__main__":
    main()
