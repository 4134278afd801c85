#!/usr/bin/env python
"""bench.py — GCUPS / Mbp/s of the triplex-scanning hot path on B200 (contract in the task prompt, §④).

Headline workload (BASELINE.json configs[3], the default): synthetic 100 Mbp DNA region (SplitMix64, seed 1001) x 3 kb
synthetic lncRNA (seed 2001), all 48 (rule, strand, orientation) tasks per 5000-bp segment.  One "step" = one pass of the
whole hot path (translate -> scan -> peaks -> windows -> traceback -> triplex records) over the region.  With --gpus N the
region's segments are sharded contiguously over N ranks (strong scaling, no collective on the data path; the ranks meet in
a barrier, a max/sum reduction of timings and counters, and the gather of every rank's rows on rank 0's host).

  value   : whole-job GCUPS with the DNA already resident in HBM (cells = sum over tasks of m * n_seg, counted once)
  e2e     : the same through the reference-facing C-ABI call with HOST buffers (H2D of the DNA and D2H of all hits inside
            the timed region)
  roofline: the dominant kernel (k_scan) against the integer-SIMD peak, both the 8-slot model and the best-mix bound
  parity_sample: the chunks the reference binary was timed on (cpu_baseline) are scanned by this build as well and the
            -TFOsorted / -TFOclass files compared byte for byte

Other workloads (own lines, for profiles/): --queries N (configs[4]: N synthetic lncRNAs of 1-10 kb x --region-mbp of
DNA seed 1002, (lncRNA, DNA part) jobs pulled from one atomic queue shared by all ranks) and --config demo | meg3 | h19 |
malat1 | neat1 (configs[0]-[2]: the shipped lncRNAs against testDNA / the 532 MEG3 example regions).

`--impl reference` times the unmodified reference binary (oracle/_ref/fasim) on this box's host cores.
"""
import argparse
import ctypes as C
import gzip
import json
import os
import pickle
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "fasim-longtarget_b200"))
DATA = os.path.join(ROOT, "tests", "golden", "data")

REGION_BP = 100_000_000
RNA_NT = 3000
DNA_SEED, RNA_SEED = 1001, 2001
MQ_DNA_SEED, MQ_RNA_SEED0 = 1002, 4001
CUT, OVERLAP, TASKS_PER_SEG = 5000, 100, 48
COMPLEX_FLAGS = ["-i", "70", "-S", "1.0", "-ni", "25", "-pt", "-500", "-ds", "10", "-lg", "60"]       # BASELINE.json configs[2]
COMPLEX_PARAMS = dict(min_identity=70, min_stability=1, nt_min=25, penalty_t=-500, c_distance=10, c_length=60)


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


def synthetic_queries(n):
    return [("synRNA%d" % k, splitmix_bases(MQ_RNA_SEED0 + k, 1000 + splitmix_first(MQ_RNA_SEED0 + k) % 9001).tobytes().decode())
            for k in range(n)]


def region_cells(n_bases, m, cut=CUT, overlap=OVERLAP, tasks=TASKS_PER_SEG):
    stride = cut - overlap
    total = 0
    pos = 0
    while pos < n_bases:
        total += min(cut, n_bases - pos)
        pos += stride
    return total * m * tasks


def read_fasta(path):
    recs, name, parts = [], None, []
    fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    for line in fh:
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if name is not None:
                recs.append((name, "".join(parts)))
            name, parts = line[1:], []
        else:
            parts.append(line)
    if name is not None:
        recs.append((name, "".join(parts)))
    fh.close()
    return recs


def real_config(name):
    """BASELINE.json configs[0]-[2] -> (workload text, lncRNA name, lncRNA, DNA records [(header, seq)], CLI flags, params)."""
    if name == "demo":
        (h, rna), = read_fasta(os.path.join(DATA, "H19.fa"))
        return ("H19 (2812 nt) x testDNA.fa (4366 bp), -lg 40 (BASELINE.json configs[0])", h, rna,
                read_fasta(os.path.join(DATA, "testDNA.fa")), ["-lg", "40"], dict(c_length=40))
    regions = read_fasta(os.path.join(DATA, "MEG3-DNAseq.fa.gz"))
    if name == "meg3":
        (h, rna), = read_fasta(os.path.join(DATA, "MEG3-ENST00000451743.fa"))
        return ("MEG3 (%d nt) x its 532 example regions (1.316 Mbp), default flags -c 5000 -i 60 (BASELINE.json configs[1])" % len(rna),
                h, rna, regions, [], {})
    fn = {"h19": "H19.fa", "malat1": "MALAT1.fa", "neat1": "NEAT1.fa"}[name]
    (h, rna), = read_fasta(os.path.join(DATA, fn))
    return ("%s (%d nt) x the 532 MEG3 example regions (1.316 Mbp; the lncRNA's own example DNA set is absent from the reference "
            "repository), complex flags -i 70 -S 1.0 -ni 25 -pt -500 -ds 10 -lg 60 (BASELINE.json configs[2])" % (h, len(rna)),
            h, rna, regions, COMPLEX_FLAGS, COMPLEX_PARAMS)


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


def run_reference_sample(rna_name, rna_text, chunks, cores, flags=(), keep=False):
    """Runs the unmodified reference binary, one process per host core, each on one single-record, single-line chunk file
    (BASELINE.md §3).  chunks: [(species, chr, start0, seq)].  Returns (wall seconds, {file name: bytes} if keep)."""
    d = tempfile.mkdtemp(prefix="fasim_ref_")
    try:
        open(os.path.join(d, "rna.fa"), "w").write(">%s\n%s\n" % (rna_name, rna_text))
        for k, (sp, ch, start, seq) in enumerate(chunks):
            open(os.path.join(d, "c%03d.fa" % k), "w").write(">%s|%s|%d-%d\n%s\n" % (sp, ch, start + 1, start + len(seq), seq))
        os.mkdir(os.path.join(d, "out"))
        t0 = time.perf_counter()
        procs = []
        pending = list(range(len(chunks)))
        while pending or procs:
            while pending and len(procs) < cores:
                k = pending.pop(0)
                procs.append(subprocess.Popen([reference_binary(), "-f1", "c%03d.fa" % k, "-f2", "rna.fa", "-O", "out/"] + list(flags), cwd=d,
                                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            procs = [p for p in procs if p.poll() is None]
            time.sleep(0.005)
        secs = time.perf_counter() - t0
        outs = {}
        if keep:
            for f in sorted(os.listdir(os.path.join(d, "out"))):
                outs[f] = open(os.path.join(d, "out", f), "rb").read()
        return secs, outs
    finally:
        shutil.rmtree(d, ignore_errors=True)


def syn_chunks(chunk_bp, count, seed=DNA_SEED, region=REGION_BP, seed_offset=0):
    chunks = []
    for k in range(count):
        start = (k * 7919 * 4900 + seed_offset) % max(1, region - chunk_bp)
        chunks.append(("syn", "chr1", start, splitmix_bases(seed, chunk_bp, start).tobytes().decode()))
    return chunks


def cpu_baseline(rna_name, rna_text, chunks, cores, flags=(), keep=False):
    """-> (GCUPS, Mbp/s, wall seconds, outputs)"""
    secs, outs = run_reference_sample(rna_name, rna_text, chunks, cores, flags, keep)
    cells = sum(region_cells(len(c[3]), len(rna_text)) for c in chunks)
    return cells / secs / 1e9, sum(len(c[3]) for c in chunks) / secs / 1e6, secs, outs


def parity_sample(eng, fb, rna_name, rna_text, chunks, ref_outs, params):
    """The chunk files the reference was just timed on, scanned by this build through the C ABI (ltg_scan_record ->
    ltg_cluster -> ltg_write_tfosorted / ltg_write_tfoclass) and compared byte for byte with the reference's files."""
    d = tempfile.mkdtemp(prefix="fasim_par_")
    try:
        eng.set_params(**params)
        eng.set_query(rna_name, rna_text)
        n_files = n_equal = rows = 0
        bad = []
        for k, (sp, ch, start, seq) in enumerate(chunks):
            res = eng.scan_record(seq, ch, start + 1)
            eng.cluster_triplex(res)
            base = "%s-%s-c%03d" % (sp, rna_name, k)
            path = os.path.join(d, base + "-TFOsorted")
            eng.printResult(res, path)
            rc = fb.lib().ltg_write_tfoclass(res, C.byref(eng.params), path.encode(), ch.encode(), start + 1, len(seq), rna_name.encode())
            eng.free(res)
            if rc != 0:
                raise RuntimeError(fb.lib().ltg_last_error().decode())
            for f in sorted(os.listdir(d)):
                if not f.startswith(base + "-"):
                    continue
                got = open(os.path.join(d, f), "rb").read()
                n_files += 1
                if ref_outs.get(f) == got:
                    n_equal += 1
                elif len(bad) < 4:
                    bad.append(f)
                if f.endswith("-TFOsorted"):
                    rows += max(0, got.count(b"\n") - 1)
        missing = [f for f in ref_outs if not os.path.exists(os.path.join(d, f))]
        return {"chunks": len(chunks), "files": n_files, "rows": rows, "equal": bool(n_files > 0 and n_equal == n_files and not missing),
                "mismatching_files": bad + missing[:4],
                "against": "unmodified reference binary (oracle/_ref/fasim) on the same chunk files, byte compare of -TFOsorted and -TFOclass1/2"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def bench_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if reference_binary() is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/fasim was not built (reference sources absent)"}), file=OUT, flush=True)
        return
    flags = []
    if args.config != "syn100":
        workload, rna_name, rna, records, flags, _ = real_config(args.config)
        recs = records[:max(1, min(len(records), args.ref_records))]
        chunk_sets = [[("hg19", "chr%d" % k, 0, s) for k, (_, s) in enumerate(recs)]] * args.steps
        warm = [chunk_sets[0][:min(len(recs), cores)]]
        sample = "%d of %d records per step, one reference process per record, %d at a time" % (len(recs), len(records), cores)
    elif args.queries > 0:
        qs = synthetic_queries(args.queries)
        rna_name, rna = qs[0]
        workload = "%d synthetic lncRNAs x synthetic %g Mbp DNA (BASELINE.json configs[4]); reference timed on lncRNA 0 (%d nt)" % (args.queries, args.region_mbp, len(rna))
        chunk_sets = [syn_chunks(args.ref_chunk_bp, cores, MQ_DNA_SEED, int(args.region_mbp * 1e6), 1000 + s) for s in range(args.steps)]
        warm = [syn_chunks(4900 * 2 + 100, cores, MQ_DNA_SEED, int(args.region_mbp * 1e6), w) for w in range(args.warmup)]
        sample = "%d chunks of %d bp per step, one reference process per host core" % (cores, args.ref_chunk_bp)
    else:
        rna_name, rna = "synRNA3k", splitmix_bases(RNA_SEED, RNA_NT).tobytes().decode()
        workload = "synthetic 100 Mbp region x 3 kb lncRNA, 48 tasks/segment (BASELINE.json configs[3])"
        chunk_sets = [syn_chunks(args.ref_chunk_bp, cores, seed_offset=1000 + s) for s in range(args.steps)]
        warm = [syn_chunks(4900 * 2 + 100, cores, seed_offset=w) for w in range(args.warmup)]     # short warm-up passes (page cache, frequency)
        sample = "%d chunks of %d bp per step, one reference process per host core" % (cores, args.ref_chunk_bp)
    for ch in warm:
        cpu_baseline(rna_name, rna, ch, cores, flags)
    t_all, cells_all, bp_all = 0.0, 0.0, 0.0
    for ch in chunk_sets:
        g, mb, secs, _ = cpu_baseline(rna_name, rna, ch, cores, flags)
        t_all += secs; cells_all += g * secs * 1e9; bp_all += mb * secs * 1e6
    gcups = cells_all / t_all / 1e9
    line = {"metric": "GCUPS (scan cells m*n per task, counted once) of the triplex scan", "value": gcups,
            "unit": "GCUPS", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 saturating SSE2 (reference)", "data": "synthetic" if args.config == "syn100" else "reference example files",
            "config": {"workload": workload, "sample": sample},
            "mbp_per_s": bp_all / t_all / 1e6,
            "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": "reference",
                             "sample": sample + " x %d steps (linear extrapolation to the whole workload)" % args.steps},
            "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def int_simd_peak():
    """Integer-SIMD peaks of this pool's B200 in GCUPS for the scan recurrence: profiles/int_simd_peak.json (issue rates
    measured by tools/ubench_simd.cu / ubench_pairs.cu on the pool's boxes)."""
    p = os.path.join(ROOT, "profiles", "int_simd_peak.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {"peak_gcups": 8170.0, "source": "fallback: 77.46 thread-instr/clk/SM measured in round 1 -> 8.17 TCUPS at 1.958 GHz"}


def hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md"


def live_simd_rates():
    """Runs tools/ubench_simd (register-only issue-rate microbenchmark) in this job when the binary travelled with the repo:
    the roofline denominator is then re-measured on the very box that produced the numerator."""
    exe = os.path.join(ROOT, "tools", "ubench_simd_r2")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception as e:          # a failed microbenchmark must not cost the bench line
        return {"error": str(e)}


class JobQueue:
    """One atomic counter shared by all ranks (per-GPU workers pull from a single queue).  Under torchrun the
    counter lives in torch.distributed's TCP store (`add` is an atomic fetch-and-add); at N = 1 it is a local integer."""

    def __init__(self, dist):
        self.dist, self.local, self.epoch = dist, 0, 0
        self.store = dist.distributed_c10d._get_default_store() if dist else None

    def reset(self):
        self.epoch += 1
        self.local = 0

    def next(self):
        if self.store is None:
            self.local += 1
            return self.local - 1
        return self.store.add("ltg_jobs_%d" % self.epoch, 1) - 1


TRI_BYTES = 104       # sizeof(ltg_triplex), checked against the ctypes mirror in bench_gpu


def result_blob(res):
    """The rows of an ltg_result as two byte strings (ltg_triplex array, text pool) — what a rank ships to rank 0."""
    r = res.contents
    tri = C.string_at(r.triplex, r.n_triplex * TRI_BYTES) if r.n_triplex else b""
    text = C.string_at(r.text, r.text_bytes) if r.text_bytes else b""
    return tri, text


def bench_gpu(args, rank, world, local_rank):
    import torch
    import fasim_b200 as fb
    assert C.sizeof(fb.Triplex) == TRI_BYTES
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")        # host-side gather of the rows (no device collective on the data path)
    torch.cuda.set_device(local_rank)
    region = int(args.region_mbp * 1e6)
    mode = "queries" if args.queries > 0 else ("real" if args.config != "syn100" else "syn100")
    params, flags = {}, []
    if mode == "real":
        workload, rna_name, rna, records, flags, params = real_config(args.config)
        queries = [(rna_name, rna)]
        hdrs = [h.split("|") for h, _ in records]
        rec_meta = [(h[1] if len(h) > 1 else "chr", int(h[2].split("-")[0]) if len(h) > 2 else 1) for h in hdrs]
        lens = [len(s) for _, s in records]
        # this rank's records: contiguous run (whole records; the segments of short records share device batches)
        r_lo, r_hi = (len(records) * rank) // world, (len(records) * (rank + 1)) // world
        cat = np.frombuffer("".join(s for _, s in records[r_lo:r_hi]).encode(), dtype=np.uint8).copy()
        host = torch.from_numpy(cat).pin_memory() if len(cat) else torch.zeros(1, dtype=torch.uint8).pin_memory()
        offs = np.concatenate([[0], np.cumsum(lens[r_lo:r_hi])]).astype(np.int64)
        nrec = r_hi - r_lo
        tags = [rec_meta[r_lo + k][0].encode() for k in range(nrec)]
        a_len = (C.c_int64 * max(nrec, 1))(*lens[r_lo:r_hi])
        a_chr = (C.c_char_p * max(nrec, 1))(*tags)
        a_start = (C.c_int64 * max(nrec, 1))(*[rec_meta[r_lo + k][1] for k in range(nrec)])
        total_bases = sum(lens)
    elif mode == "queries":
        # BASELINE.json configs[4]: lncRNAs of 1000 + z % 9001 nt, seeds 4001.., DNA seed 1002.  Jobs =
        # (lncRNA, DNA part) pairs, longest lncRNA first, pulled from ONE queue by all ranks; every rank keeps the whole DNA.
        queries = synthetic_queries(args.queries)
        order = sorted(range(len(queries)), key=lambda q: -len(queries[q][1]))
        n_parts = max(1, int(round(region / 40e6)))
        parts = [fb.shard_segments(region, n_parts, p, CUT, OVERLAP) for p in range(n_parts)]
        jobs = [(q, p) for q in order for p in range(n_parts)]
        # warm-up passes of this workload: every lncRNA (profile build, query switch, allocator growth) against the first 2 Mbp
        warm_region = min(region, 2_000_000)
        warm_part = fb.shard_segments(warm_region, 1, 0, CUT, OVERLAP)
        host = torch.from_numpy(splitmix_bases(MQ_DNA_SEED, region)).pin_memory()
        total_bases = region * len(queries)
    else:
        first_seg, nseg, lo, nb = fb.shard_segments(region, world, rank, CUT, OVERLAP)     # contiguous run of whole segments
        rna = splitmix_bases(RNA_SEED, RNA_NT).tobytes().decode()
        queries = [("synRNA3k", rna)]
        host = torch.from_numpy(splitmix_bases(DNA_SEED, nb, lo)).pin_memory()
        total_bases = region
    dev = host.to("cuda", non_blocking=False)
    eng = fb.Engine(local_rank, **params)
    eng.set_query(*queries[0])
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    queue = JobQueue(dist)
    lib = fb.lib()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def check(rc):
        if rc != 0:
            raise RuntimeError(lib.ltg_last_error().decode())

    def scan_step(device_resident, stats, warm=False):
        """one pass of this rank's share of the workload; returns the (triplex bytes, text bytes) blobs of its results"""
        blobs = []

        def account(res, bases):
            r = res.contents
            stats["cells"] += r.scan_cells; stats["bases"] += bases
            stats["rows"] += r.n_triplex; stats["segs"] += r.n_segments
            stats["launches"] += r.gpu_launches; stats["scan_ms"] += r.gpu_ms_scan_kernel; stats["scan_launches"] += r.n_scan_launches
            stats["win_ms"] += r.gpu_ms_window; stats["win_cells"] += r.window_cells; stats["peaks"] += r.n_peaks
            stats["lit_tasks"] += r.n_literal_tasks; stats["lit_windows"] += r.n_literal_windows; stats["probed"] += r.n_q4_probed
            stats["d2h"] += r.d2h_bytes; stats["h2d"] += r.h2d_bytes
            blobs.append(result_blob(res))
            lib.ltg_result_free(res)

        base_ptr = dev.data_ptr() if device_resident else host.data_ptr()
        if mode == "real":
            if nrec > 0:
                res = C.POINTER(fb.Result)()
                ptrs = (C.c_void_p * nrec)(*[base_ptr + int(offs[k]) for k in range(nrec)])
                check(lib.ltg_scan_records_at(eng._h, nrec, ptrs, 1 if device_resident else 0, a_len, a_chr, a_start, C.byref(res)))
                account(res, int(offs[-1]))
        elif mode == "queries":
            cur = None
            while True:
                j = queue.next()
                if j >= (len(order) if warm else len(jobs)):
                    break
                q, p = (order[j], 0) if warm else jobs[j]
                if q != cur:
                    eng.set_query(*queries[q])
                    cur = q
                fs, ns, plo, pnb = warm_part if warm else parts[p]
                reg = warm_region if warm else region
                res = C.POINTER(fb.Result)()
                check(lib.ltg_scan_shard(eng._h, C.c_void_p(base_ptr + plo), 1 if device_resident else 0, pnb, b"chr1", 1, reg, fs, ns, C.byref(res)))
                account(res, min(reg, (fs + ns) * (CUT - OVERLAP)) - fs * (CUT - OVERLAP))
                stats["jobs"] += 1
        else:
            res = C.POINTER(fb.Result)()
            # the reference-facing C-ABI call; device_resident: DNA already in HBM, else HOST buffer (H2D inside the call)
            check(lib.ltg_scan_shard(eng._h, C.c_void_p(base_ptr), 1 if device_resident else 0, nb, b"chr1", 1, region, first_seg, nseg, C.byref(res)))
            account(res, min(region, (first_seg + nseg) * (CUT - OVERLAP)) - first_seg * (CUT - OVERLAP))
        return blobs

    def run(device_resident, steps, warm=False):
        stats = dict(cells=0, bases=0, rows=0, segs=0, launches=0, scan_ms=0.0, scan_launches=0, win_ms=0.0, win_cells=0, peaks=0, d2h=0, h2d=0,
                     lit_tasks=0, lit_windows=0, probed=0, jobs=0, gathered_rows=0, gathered_bytes=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        queue.reset()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for step in range(steps):
            blobs = scan_step(device_resident, stats, warm and mode == "queries")
            # hits gathered to the host of rank 0 (north star: "hits gathered to the host"), inside the timed region
            if dist:
                payload = pickle.dumps(blobs, protocol=pickle.HIGHEST_PROTOCOL)
                gathered = [None] * world if rank == 0 else None
                dist.gather_object(payload, gathered, dst=0, group=host_group)
                if rank == 0:
                    for pl in gathered:
                        for tri, text in pickle.loads(pl):
                            stats["gathered_rows"] += len(tri) // TRI_BYTES
                            stats["gathered_bytes"] += len(tri) + len(text)
                if mode == "queries" and step + 1 < steps:
                    dist.barrier(group=host_group)       # every rank has drained this step's queue before the next one opens
                    queue.reset()
            else:
                stats["gathered_rows"] += sum(len(t) // TRI_BYTES for t, _ in blobs)
                stats["gathered_bytes"] += sum(len(t) + len(x) for t, x in blobs)
                if mode == "queries" and step + 1 < steps:
                    queue.reset()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1])
            keys = sorted(k for k in stats if k not in ("scan_ms", "win_ms", "gathered_rows", "gathered_bytes"))
            v = torch.tensor([float(stats[k]) for k in keys], device="cuda", dtype=torch.float64)
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
            for k, x in zip(keys, v.tolist()):
                stats[k] = x
            tm = torch.tensor([stats["scan_ms"], stats["win_ms"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            stats["scan_ms"], stats["win_ms"] = float(tm[0]), float(tm[1])
        return ms, wall, stats

    # warm-up (page-in, clocks, allocator growth), then the timed regions
    # (queries mode: a warm-up pass runs every lncRNA against the first 2 Mbp only — a full pass of configs[4] takes minutes)
    if args.warmup > 0:
        run(True, args.warmup, warm=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if not args.e2e_only:
        ms, wall, st = run(True, args.steps)
        clocks = sampler.stop() if rank == 0 else None
    if args.warmup > 0:
        run(False, 1, warm=True)
    ms_e, wall_e, st_e = run(False, args.steps)
    if args.e2e_only:      # one pass only (full-size configs[4]): the host-buffer pass is the one measured; `value` repeats it and says so
        ms, wall, st = ms_e, wall_e, st_e
        clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        gcups = st["cells"] / (ms * 1e-3) / 1e9
        gcups_e = st_e["cells"] / (ms_e * 1e-3) / 1e9
        pk = int_simd_peak()
        peak, peak_src, dram_per_seg = pk["peak_gcups"], pk.get("source", "profiles/int_simd_peak.json"), pk.get("scan_dram_bytes_per_segment")
        # roofline of the dominant kernel: algorithmic cells of one k_scan launch / its CUDA-event duration.  With N ranks
        # each rank runs its own launches concurrently: per-GPU achieved = cells / N / (max-over-ranks scan time).
        scan_gcups = st["cells"] / world / (st["scan_ms"] * 1e-3) / 1e9 if st["scan_ms"] > 0 else 0.0
        hbm_pk, hbm_src = hbm_peak()
        ms_launch = st["scan_ms"] / max(st["scan_launches"] / world, 1)
        traffic = dram_per_seg * st["segs"] / max(st["scan_launches"], 1) if (dram_per_seg and mode == "syn100") else None
        hbm_secondary = None
        if traffic and ms_launch > 0:
            gbs = traffic / (ms_launch * 1e-3) / 1e9
            hbm_secondary = {"achieved": gbs, "peak": hbm_pk, "unit": "GB/s", "frac": gbs / hbm_pk, "peak_source": hbm_src}
        if mode == "syn100":
            wl = "synthetic %g Mbp region x 3 kb lncRNA, 48 tasks per 5000-bp segment, sharded over %d GPU(s) (BASELINE.json configs[3])" % (args.region_mbp, world)
        elif mode == "queries":
            wl = ("%d synthetic lncRNAs (%d..%d nt, %d nt in all) x synthetic %g Mbp DNA, %d (lncRNA, DNA part) jobs pulled from one atomic "
                  "queue by %d GPU(s) (BASELINE.json configs[4]%s)"
                  % (len(queries), min(len(q[1]) for q in queries), max(len(q[1]) for q in queries), sum(len(q[1]) for q in queries),
                     args.region_mbp, len(jobs), world, "" if (args.queries == 64 and args.region_mbp == 250) else " at reduced size"))
        else:
            wl = workload + "; records sharded contiguously over %d GPU(s)" % world
        best_mix = pk.get("best_mix_gcups")
        line = {
            "metric": "GCUPS (scan cells m*n per task, counted once) of the triplex scan",
            "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int16x2 (packed SIMD-in-register)", "data": "synthetic" if mode != "real" else "reference example files",
            "config": {"workload": wl,
                       "l2": "inputs (%.0f MB of DNA + per-batch column-max buffers > 126 MB) exceed L2" % (st["bases"] / args.steps / 1e6)
                             if mode != "real" else "per-batch column-max buffers (GBs) exceed L2; the DNA itself (1.3 MB) does not"},
            "mbp_per_s": st["bases"] / (ms * 1e-3) / 1e6,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "triplex_rows_per_step": st["rows"] / args.steps,
            "rows_gathered_on_rank0_per_step": st["gathered_rows"] / args.steps,
            "peaks_per_step": st["peaks"] / args.steps,
            "window_cells_per_step": st["win_cells"] / args.steps,
            "literal_tasks_per_step": st["lit_tasks"] / args.steps,
            "q4_probed_pairs_per_step": st["probed"] / args.steps,
            "gpu_launches": int(st["launches"]),
            "stage_ms_per_step": {"scan_kernel": st["scan_ms"] / args.steps, "window": st["win_ms"] / args.steps},
            "e2e": {"value": gcups_e, "unit": "GCUPS", "h2d_bytes_per_step": int(st_e["h2d"] / args.steps),
                    "d2h_bytes_per_step": int(st_e["d2h"] / args.steps), "ms_per_step": ms_e / args.steps,
                    "mbp_per_s": st_e["bases"] / (ms_e * 1e-3) / 1e6,
                    "rows_gathered_on_rank0_per_step": st_e["gathered_rows"] / args.steps,
                    "gathered_bytes_per_step": int(st_e["gathered_bytes"] / args.steps)},
            "roofline": {"bound": "int_simd", "kernel": "k_scan<R,4> (R = 32 or 16 rows per lane, chosen per lncRNA length)", "achieved": scan_gcups, "peak": peak,
                         "unit": "GCUPS", "frac": scan_gcups / peak if peak else None, "peak_source": peak_src,
                         "peak_best_mix": best_mix, "frac_best_mix": (scan_gcups / best_mix) if best_mix else None,
                         "whole_job_frac": gcups / world / peak if peak else None,
                         "whole_job_frac_best_mix": (gcups / world / best_mix) if best_mix else None,
                         "cells_per_launch": st["cells"] / max(st["scan_launches"], 1),
                         "ms_per_launch": ms_launch,
                         "traffic": traffic,
                         "traffic_unit": "DRAM bytes per k_scan launch (ncu dram__bytes_read+write of one launch of this build, %s, scaled by segments per launch)"
                                         % pk.get("traffic_capture", "profiles/"),
                         "hbm_secondary": hbm_secondary,
                         "note": "integer-ALU bound (SURVEY.md 8d): algorithmic HBM traffic is ~1 B of DNA per %d cells"
                                 % (len(queries[0][1]) * TASKS_PER_SEG)},
            "clocks": clocks,
        }
        if args.e2e_only:
            line["value_note"] = "--e2e-only: the DNA-resident pass was skipped; value, stage times and roofline are those of the e2e (host-buffer) pass"
        if mode == "queries":
            line["mbp_per_s_note"] = "DNA bases x lncRNAs scanned per second (each lncRNA is a full pass over the DNA)"
            line["jobs_per_step"] = st["jobs"] / args.steps
        rates = live_simd_rates() if world == 1 else None
        if rates is not None:
            line["roofline"]["live_ubench"] = rates
        # the reference on this box's host cores, and the same chunk files through this build: byte compare
        if world == 1 and not args.no_cpu_baseline and reference_binary():
            cores = os.cpu_count() or 1
            if mode == "real":
                recs = records[:max(1, min(len(records), args.ref_records))]
                chunks = [("hg19", rec_meta[k][0], rec_meta[k][1] - 1, s) for k, (_, s) in enumerate(recs)]
                sample = "%d of %d records, one unmodified reference process per record, %d at a time" % (len(recs), len(records), cores)
                qn, qs = queries[0]
            elif mode == "queries":
                qn, qs = queries[min(3, len(queries) - 1)]
                chunks = syn_chunks(args.ref_chunk_bp, cores, MQ_DNA_SEED, region)
                sample = "%d chunks of %d bp of the DNA x lncRNA %s (%d nt), one unmodified reference process per host core" % (cores, args.ref_chunk_bp, qn, len(qs))
            else:
                qn, qs = queries[0]
                chunks = syn_chunks(args.ref_chunk_bp, cores)
                sample = "%d chunks of %d bp (one unmodified reference process per host core)" % (cores, args.ref_chunk_bp)
            g, mb, secs, outs = cpu_baseline(qn, qs, chunks, cores, flags, keep=True)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": "reference", "mbp_per_s": mb,
                                    "sample": sample + ", %.1f s wall" % secs}
            line["parity_sample"] = parity_sample(eng, fb, qn, qs, chunks, outs, params)
        elif world == 1 and not args.no_cpu_baseline:
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
    ap.add_argument("--ref-records", type=int, default=64, help="--config runs: records of the DNA set the reference is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-stats", action="store_true")
    ap.add_argument("--e2e-only", action="store_true", help="skip the DNA-resident pass (long workloads: time the host-buffer pass only)")
    ap.add_argument("--queries", type=int, default=0, help="multi-query workload (configs[4]): this many lncRNAs of 1-10 kb per step")
    ap.add_argument("--config", default="syn100", choices=["syn100", "demo", "meg3", "h19", "malat1", "neat1"],
                    help="workload: the headline synthetic region (default) or one of the reference's example sets (configs[0]-[2])")
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
