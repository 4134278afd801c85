#!/usr/bin/env python
"""bench.py — GCUPS / Mbp/s of the triplex-scanning hot path on B200 (contract in the task prompt, §④).

Workload (BASELINE.json configs[3]): synthetic 100 Mbp DNA region (SplitMix64, seed 1001) x 3 kb synthetic lncRNA
(seed 2001), all 48 (rule, strand, orientation) tasks per 5000-bp segment.  One "step" = one pass of the whole hot
path (translate -> scan -> peaks -> windows -> traceback -> triplex records) over the region.  With --gpus N the
region's segments are sharded contiguously over N ranks (strong scaling, no collective on the data path; the ranks
only meet in a barrier and a max/sum reduction of the timings and counters).

  value  : whole-job GCUPS with the DNA already resident in HBM (cells = sum over tasks of m * n_seg, counted once)
  e2e    : the same through the reference-facing C-ABI call with HOST buffers (H2D of the DNA and D2H of all hits
           inside the timed region)
  roofline: the dominant kernel (k_scan) against the measured integer-SIMD peak (profiles/int_simd_peak.json)

`--impl reference` times the unmodified reference binary (oracle/_ref/fasim) on this box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fasim-longtarget_b200"))

REGION_BP = 100_000_000
RNA_NT = 3000
DNA_SEED, RNA_SEED = 1001, 2001
CUT, OVERLAP, TASKS_PER_SEG = 5000, 100, 48


# ------------------------------------------------------------------------------------------------ synthetic data
def splitmix_bases(seed, n, offset=0):
    """SplitMix64 stream, base = 'ACGT'[z >> 62]; `offset` = index of the first base."""
    out = np.empty(n, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    step = 1 << 24
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        idx = np.arange(offset + lo + 1, offset + hi + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        out[lo:hi] = lut[(z >> np.uint64(62)).astype(np.int64)]
    return out


def splitmix_first(seed):
    """First SplitMix64 output of a stream (lncRNA length = 1000 + z % 9001)."""
    M = (1 << 64) - 1
    z = (seed + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def region_cells(n_bases, m, cut=CUT, overlap=OVERLAP, tasks=TASKS_PER_SEG):
    stride = cut - overlap
    total = 0
    pos = 0
    while pos < n_bases:
        total += min(cut, n_bases - pos)
        pos += stride
    return total * m * tasks


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "fasim")
    return p if os.path.exists(p) else None


def run_reference_sample(rna_text, chunks, cores):
    """Runs the unmodified reference binary, one process per host core, each on one single-record, single-line chunk
    file.  Returns wall seconds."""
    d = tempfile.mkdtemp(prefix="fasim_ref_")
    try:
        open(os.path.join(d, "rna.fa"), "w").write(">synRNA3k\n%s\n" % rna_text)
        for k, (start, seq) in enumerate(chunks):
            open(os.path.join(d, "c%03d.fa" % k), "w").write(">syn|chr1|%d-%d\n%s\n" % (start + 1, start + len(seq), seq))
        os.mkdir(os.path.join(d, "out"))
        t0 = time.perf_counter()
        procs = []
        pending = list(range(len(chunks)))
        while pending or procs:
            while pending and len(procs) < cores:
                k = pending.pop(0)
                procs.append(subprocess.Popen([reference_binary(), "-f1", "c%03d.fa" % k, "-f2", "rna.fa", "-O", "out/"], cwd=d,
                                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            procs = [p for p in procs if p.poll() is None]
            time.sleep(0.01)
        return time.perf_counter() - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)


def cpu_baseline(rna_text, chunk_bp, cores, seed_offset=0):
    chunks = []
    for k in range(cores):
        start = (k * 7919 * 4900 + seed_offset) % (REGION_BP - chunk_bp)
        chunks.append((start, splitmix_bases(DNA_SEED, chunk_bp, start).tobytes().decode()))
    secs = run_reference_sample(rna_text, chunks, cores)
    cells = sum(region_cells(len(s), len(rna_text)) for _, s in chunks)
    return cells / secs / 1e9, sum(len(s) for _, s in chunks) / secs / 1e6, secs


def bench_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rna = splitmix_bases(RNA_SEED, RNA_NT).tobytes().decode()
    if reference_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fasim was not built (reference sources absent)"}), file=OUT, flush=True)
        return
    chunk_bp = args.ref_chunk_bp
    for w in range(args.warmup):
        cpu_baseline(rna, 4900 * 2 + 100, cores, w)             # short warm-up passes (page cache, frequency)
    t_all, cells_all, bp_all = 0.0, 0.0, 0.0
    for s in range(args.steps):
        g, mb, secs = cpu_baseline(rna, chunk_bp, cores, 1000 + s)
        t_all += secs; cells_all += g * secs * 1e9; bp_all += mb * secs * 1e6
    gcups = cells_all / t_all / 1e9
    line = {"metric": "GCUPS (scan cells m*n per task, counted once) of the triplex scan, 100 Mbp x 3 kb lncRNA", "value": gcups,
            "unit": "GCUPS", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 saturating SSE2 (reference)", "data": "synthetic",
            "config": {"workload": "synthetic 100 Mbp region x 3 kb lncRNA, 48 tasks/segment (BASELINE.json configs[3])",
                       "sample": "%d chunks of %d bp per step, one reference process per host core" % (cores, chunk_bp)},
            "mbp_per_s": bp_all / t_all / 1e6,
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": "reference",
                             "sample": "%d x %d bp chunks per step x %d steps (linear extrapolation to 100 Mbp)" % (cores, chunk_bp, args.steps)},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def int_simd_peak():
    """Measured integer-SIMD peak of this pool's B200 in GCUPS for the scan recurrence (tools/ubench_simd.cu, 5.5
    packed instructions per cell pair, register-only): profiles/int_simd_peak.json."""
    p = os.path.join(ROOT, "profiles", "int_simd_peak.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["peak_gcups"], j.get("source", "profiles/int_simd_peak.json"), j.get("scan_dram_bytes_per_segment")
    return 8170.0, "fallback: 77.46 thread-instr/clk/SM measured in round 1 -> 8.17 TCUPS at 1.958 GHz", None


def hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md"


def bench_gpu(args, rank, world, local_rank):
    import torch
    import fasim_b200 as fb
    dist = None
    if world > 1:
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG asks for it; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    region = int(args.region_mbp * 1e6)
    first_seg, nseg, lo, nb = fb.shard_segments(region, world, rank, CUT, OVERLAP)     # contiguous run of whole segments
    rna = splitmix_bases(RNA_SEED, RNA_NT).tobytes().decode()
    if args.queries > 0:
        # BASELINE.json configs[4]: lncRNAs of 1000 + z % 9001 nt, seeds 4001.., DNA seed 1002; every
        # step scans all of them against the rank's shard, switching the context's query inside the timed region
        queries = [("synRNA%d" % k, splitmix_bases(4001 + k, 1000 + splitmix_first(4001 + k) % 9001).tobytes().decode()) for k in range(args.queries)]
        dna_seed = 1002
    else:
        queries, dna_seed = [("synRNA3k", rna)], DNA_SEED
    host = torch.from_numpy(splitmix_bases(dna_seed, nb, lo)).pin_memory()
    dev = host.to("cuda", non_blocking=False)
    eng = fb.Engine(local_rank)
    eng.set_query(*queries[0])
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def run(device_resident, steps):
        stats = dict(cells=0, bases=0, rows=0, segs=0, launches=0, scan_ms=0.0, scan_launches=0, win_ms=0.0, win_cells=0, peaks=0, d2h=0, h2d=0,
                     lit_tasks=0, lit_windows=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for qname, qseq in [q for _ in range(steps) for q in queries]:
            if len(queries) > 1:
                eng.set_query(qname, qseq)
            res = C.POINTER(fb.Result)()
            # the reference-facing C-ABI call; device_resident: DNA already in HBM, else HOST buffer (H2D inside the call)
            rc = fb.lib().ltg_scan_shard(eng._h, C.c_void_p(dev.data_ptr() if device_resident else host.data_ptr()),
                                         1 if device_resident else 0, nb, b"chr1", 1, region, first_seg, nseg, C.byref(res))
            if rc != 0:
                raise RuntimeError(fb.lib().ltg_last_error().decode())
            r = res.contents
            stats["cells"] += r.scan_cells; stats["bases"] += min(region, (first_seg + nseg) * (CUT - OVERLAP)) - first_seg * (CUT - OVERLAP)
            stats["rows"] += r.n_triplex; stats["segs"] += r.n_segments
            stats["launches"] += r.gpu_launches; stats["scan_ms"] += r.gpu_ms_scan_kernel; stats["scan_launches"] += r.n_scan_launches
            stats["win_ms"] += r.gpu_ms_window; stats["win_cells"] += r.window_cells; stats["peaks"] += r.n_peaks
            stats["lit_tasks"] += r.n_literal_tasks; stats["lit_windows"] += r.n_literal_windows
            stats["d2h"] += r.d2h_bytes; stats["h2d"] += r.h2d_bytes
            fb.lib().ltg_result_free(res)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
            keys = sorted(k for k in stats if k not in ("scan_ms", "win_ms"))
            v = torch.tensor([float(stats[k]) for k in keys], device="cuda", dtype=torch.float64)
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            for k, x in zip(keys, v.tolist()):
                stats[k] = x
            tm = torch.tensor([stats["scan_ms"], stats["win_ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            stats["scan_ms"], stats["win_ms"] = float(tm[0]), float(tm[1])
        return ms, wall, stats

    # warm-up (>= 3 steps: page-in, clocks, allocator growth), then the timed regions
    run(True, max(args.warmup, 1))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, wall, st = run(True, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    run(False, 1)
    ms_e, wall_e, st_e = run(False, args.steps)

    if rank == 0:
        gcups = st["cells"] / (ms * 1e-3) / 1e9
        gcups_e = st_e["cells"] / (ms_e * 1e-3) / 1e9
        peak, peak_src, dram_per_seg = int_simd_peak()
        # roofline of the dominant kernel: algorithmic cells of one k_scan launch / its CUDA-event duration.  With N ranks
        # each rank runs its own launches concurrently: per-GPU achieved = cells / N / (max-over-ranks scan time).
        scan_gcups = st["cells"] / world / (st["scan_ms"] * 1e-3) / 1e9 if st["scan_ms"] > 0 else 0.0
        dna_bytes = st["bases"]            # 1 B/base read once per item pair-group, hits out are negligible
        # secondary counter: DRAM bandwidth of the dominant kernel against the measured copy peak
        hbm_pk, hbm_src = hbm_peak()
        ms_launch = st["scan_ms"] / max(st["scan_launches"] / world, 1)
        hbm_secondary = None
        if dram_per_seg and ms_launch > 0:
            gbs = dram_per_seg * st["segs"] / max(st["scan_launches"], 1) / (ms_launch * 1e-3) / 1e9
            hbm_secondary = {"achieved": gbs, "peak": hbm_pk, "unit": "GB/s", "frac": gbs / hbm_pk, "peak_source": hbm_src}
        line = {
            "metric": "GCUPS (scan cells m*n per task, counted once) of the triplex scan, " + ("100 Mbp x 3 kb lncRNA" if args.queries <= 0 else "multi-query batch"),
            "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int16x2 (packed SIMD-in-register)", "data": "synthetic",
            "config": {"workload": ("synthetic %g Mbp region x 3 kb lncRNA, 48 tasks per 5000-bp segment, sharded over %d GPU(s) "
                                    "(BASELINE.json configs[3])" % (args.region_mbp, world)) if args.queries <= 0 else
                                   ("%d synthetic lncRNAs (%d..%d nt, %d nt in all) x synthetic %g Mbp DNA, sharded over %d GPU(s) "
                                    "(BASELINE.json configs[4] at reduced size; not the headline workload)"
                                    % (len(queries), min(len(q[1]) for q in queries), max(len(q[1]) for q in queries),
                                       sum(len(q[1]) for q in queries), args.region_mbp, world)),
                       "l2": "inputs (%.0f MB of DNA + per-batch column-max buffers > 126 MB) exceed L2" % (st["bases"] / args.steps / 1e6)},
            "mbp_per_s": st["bases"] / (ms * 1e-3) / 1e6,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "triplex_rows_per_step": st["rows"] / args.steps,
            "peaks_per_step": st["peaks"] / args.steps,
            "window_cells_per_step": st["win_cells"] / args.steps,
            "literal_tasks_per_step": st["lit_tasks"] / args.steps,
            "gpu_launches": int(st["launches"]),
            "stage_ms_per_step": {"scan_kernel": st["scan_ms"] / args.steps, "window": st["win_ms"] / args.steps},
            "e2e": {"value": gcups_e, "unit": "GCUPS", "h2d_bytes_per_step": int(st_e["h2d"] / args.steps),
                    "d2h_bytes_per_step": int(st_e["d2h"] / args.steps), "ms_per_step": ms_e / args.steps,
                    "mbp_per_s": st_e["bases"] / (ms_e * 1e-3) / 1e6},
            "roofline": {"bound": "int_simd", "kernel": "k_scan<32,4> (rows per lane chosen per lncRNA length: 32 here)", "achieved": scan_gcups, "peak": peak, "unit": "GCUPS",
                         "frac": scan_gcups / peak if peak else None, "peak_source": peak_src,
                         "cells_per_launch": st["cells"] / max(st["scan_launches"], 1),
                         "ms_per_launch": st["scan_ms"] / max(st["scan_launches"] / world, 1),
                         "traffic": (dram_per_seg * st["segs"] / max(st["scan_launches"], 1)) if dram_per_seg else None,
                         "traffic_unit": "DRAM bytes per k_scan launch (ncu dram__bytes_read+write of one launch, scaled by segments per launch)",
                         "hbm_secondary": hbm_secondary,
                         "note": "integer-ALU bound (SURVEY.md 8d): HBM traffic is ~1 B per %d cells; hbm_gbs_algorithmic=%.3f"
                                 % (RNA_NT * TASKS_PER_SEG, dna_bytes / max(st["scan_ms"], 1e-9) / 1e6)},
            "clocks": clocks,
        }
        if args.queries > 0:
            line["mbp_per_s_note"] = "DNA bases x lncRNAs scanned per second (each lncRNA is a full pass over the DNA)"
        elif world == 1 and not args.no_cpu_baseline and reference_binary():
            cores = os.cpu_count() or 1
            g, mb, secs = cpu_baseline(rna, args.ref_chunk_bp, cores)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": "reference", "mbp_per_s": mb,
                                    "sample": "%d chunks of %d bp (one unmodified reference process per host core, %.1f s wall)"
                                              % (cores, args.ref_chunk_bp, secs)}
        elif world == 1:
            line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": 0, "kind": "reference", "sample": "reference binary not built"}
        print(json.dumps(line), file=OUT, flush=True)
        if args.debug_stats:
            print(json.dumps({"window_stage_stats": eng.debug_stats()}), file=sys.stderr)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


OUT = sys.stdout


def main():
    # stdout carries exactly ONE JSON line: whatever libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--region-mbp", type=float, default=REGION_BP / 1e6)
    ap.add_argument("--ref-chunk-bp", type=int, default=49100)        # 10 full segments + tail per core and step
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-stats", action="store_true")
    ap.add_argument("--queries", type=int, default=0, help="multi-query workload (configs[4]): this many lncRNAs of 1-10 kb per step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        bench_reference(args, rank, world)
        return
    bench_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
