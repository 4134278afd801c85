"""How much would overlapping the stages of consecutive batches buy?  Upper bound: two independent contexts on the same GPU,
each scanning half of the region from its own host thread (their kernels interleave freely), against one context scanning
everything."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fasim-longtarget_b200")); sys.path.insert(0, ROOT)
import torch
import fasim_b200 as fb
from bench import splitmix_bases, DNA_SEED, RNA_SEED, RNA_NT
MBP = float(sys.argv[1]) if len(sys.argv) > 1 else 40
n = int(MBP * 1e6)
rna = splitmix_bases(RNA_SEED, RNA_NT).tobytes().decode()
dev = torch.from_numpy(splitmix_bases(DNA_SEED, n)).cuda()
def scan(eng, first, count, lo, nb):
    res = eng.scan_shard(nb, n, first, count, "chr1", 1, device_ptr=dev.data_ptr() + lo)
    rows = res.contents.n_triplex; eng.free(res); return rows
engs = [fb.Engine(0) for _ in range(2)]
for e in engs: e.set_query("r", rna)
parts = [fb.shard_segments(n, 2, r) for r in range(2)]
whole = fb.shard_segments(n, 1, 0)
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter(); scan(engs[0], *whole); torch.cuda.synchronize(); t1 = time.perf_counter() - t
    torch.cuda.synchronize(); t = time.perf_counter()
    th = [threading.Thread(target=scan, args=(engs[r],) + parts[r]) for r in range(2)]
    [x.start() for x in th]; [x.join() for x in th]
    torch.cuda.synchronize(); t2 = time.perf_counter() - t
    print("rep %d: one context %.1f ms, two concurrent contexts %.1f ms  (%.3fx)" % (rep, t1 * 1e3, t2 * 1e3, t1 / t2))
