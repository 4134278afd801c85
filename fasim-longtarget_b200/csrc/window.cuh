// Window stage: the device counterpart of fastSIM's per-peak loop (fastsim.h:202-272) around
// Aligner::Align -> ssw_align (ssw_cpp.cpp:599, sswNew.cpp:1446): forward Smith-Waterman of the
// lncRNA against a short DNA window ending at the peak, reverse pass to locate the beginning, banded
// traceback to a CIGAR, expansion to the aligned TFO / TTS strings.
//
// Layout is transposed with respect to the scan: lanes own window COLUMNS (R per lane) and the RNA
// streams through them, so a 55-column window keeps ~m/(m+g) of its lanes busy instead of ~60 %.
// Windows are right-aligned in groups of g = 4/8/16/32 lanes (pad columns on the left stay identically
// zero); a warp carries 32/g groups in each 16-bit half.  Scores are computed on the fly
// (XNOR of scaled base codes, then max(~e + 6, -4)), so every group may stream its own RNA slice.
// Result tracking costs one VIMNMX3 per two cells plus a rarely taken slow path: forward keeps the
// maximum with the reference's tie rules (first column, then first row: sswNew.cpp:605-629), reverse keeps
// the first column (then first row) whose value equals the forward score (terminate test :617).
//
// Row pruning (exact, not a heuristic).  The reference sweeps all m RNA rows for every window.  Two facts
// let this stage sweep far fewer rows and still return the identical (score, end column, end row):
//  (1) a local alignment that spans c columns and has a positive score spans fewer than 2.25*c + 1 rows
//      (every inserted row costs >= 4, every column yields <= 5), so the value of a cell only depends on
//      the win_margin(c) rows above it;
//  (2) a window cell can never exceed the same cell of the whole-segment matrix, and the scan stage kept,
//      per granule of kGranRows RNA rows, the column maxima of that matrix (scan.cuh).
// For a guess L of the window's best score, the stream covers the granules whose maximum over the window's
// columns reaches L, the granules between them, and the margin above the first.  Let B be the largest such
// maximum among the granules left out.  Values computed on a sub-range of rows are lower bounds of the true
// values and every left-out cell is <= B, so a result r > B is exact (value, and every tie that matters for the
// tie rules); otherwise the window is swept again with L = r (a proven lower bound), and that second result is
// exact by the same argument.  First guess: L = peak score.
// The reverse pass only needs win_margin(re+1) rows: every cell equal to the forward score belongs to an
// alignment that starts at the forward end cell (tie rules of the forward pass), see DESIGN.md.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace ltg {

#ifndef LTG_WIN_R
#define LTG_WIN_R 10
#endif
constexpr int kWinR = LTG_WIN_R;    // window columns per lane (10: the dominant 67..80-column windows fill 28..32 lanes; 8..12 measured within 1 %)
constexpr int kMaxWindow = 32 * kWinR;
constexpr int kWinRowBuckets = 64;  // stream-length buckets of 64 rows (the last one takes everything longer)
constexpr int kWinKeys = 32 * kWinRowBuckets;

// Work list of one k_win_dp launch.  A window of `len` columns is carried by a group of g = ceil(len / kWinR) lanes
// (class c = g - 1); a warp carries floor(32 / g) groups in each 16-bit half.  Windows are counting-sorted by
// (g, stream length descending), so the windows that share a warp stream about the same number of RNA rows and the
// short bins come last in the queue.
constexpr int kWinCopies = 16;      // the key histogram is kept in this many copies (chosen per block): millions of pieces fall on a few
                                    // dozen keys, and same-address atomics serialise
struct WinSched {
    int hist[kWinCopies][kWinKeys];        // pieces per (copy, key)
    int off[kWinCopies][kWinKeys];         // first list slot of every (copy, key): keys ascending, the copies of a key back to back
    int fill[kWinCopies][kWinKeys];
    int cls_count[32];         // windows per class
    int cls_off[32];           // first list slot of the class
    int bin_start[32];         // first bin (warp work unit) of the class
    int total_bins;
    int bin_counter;
    int n_pieces;              // work items of the launch being planned
    // statistics (LTG_STATS): windows / cells planned per (round, retry); [8] = reverse pass
    unsigned long long st_windows[10];
    unsigned long long st_cells[10];
    unsigned long long st_filter[6];        // LTG_FILTER_STATS: see CompactArgs::filt
};

struct WinState {
    // peak pool
    int n_peaks;
    const int* pk_task;
    const int* pk_pos;
    const int* pk_score;
    // per peak
    int* w_len;        // columns of the window in flight (cut, or re+1 in the reverse pass)
    int* w_ws;         // first column (translated-segment coordinates) of the forward window in flight
    int* w_shift;      // lowercase compat only: columns by which that window starts right of pos - cut + 1 (the start clamp the older
                       // variant applies to the substring but not to the reported coordinates, fastSim.h:204-212); 0 otherwise
    int* w_bound;      // a result >= w_bound is exact (row pruning, see the header comment); 0: all rows were streamed
    int* w_floor;      // cells <= w_floor cannot matter for this sweep (they neither prove exactness nor beat a proven
                       // lower bound), so the per-lane result tracker starts there
    int* w_flight;     // 1: the peak has work items in the launch being planned / run
    int* w_done;       // 1: final alignment chosen
    int* w_next;       // next forward round this peak takes part in (rounds in between are provably no-ops, see k_win_probe)
    int* w_probe;      // 1: the round's result waits for its reverse probe; 2: fin_rb / fin_qb are final already
    int* w_q4;         // k_win_dp<.., Q4CHK>: bit 0 / bit 1 = a forward / reverse sweep of the window in flight saw an F >= 132 enter a row
                       // that starts a stripe of the reference's layout (the only place the Q4 quirk can act); nullptr: not checked, every
                       // window that scores >= 148 goes through the literal emulation
    int* best_sw; int* best_cut; int* best_re; int* best_qe; int* best_ws;
    int* fin_sw; int* fin_cut; int* fin_re; int* fin_qe; int* fin_rb; int* fin_qb; int* fin_ws; int* fin_shift;
    int compat;        // 1: window loop of the older variant (fastSim.h:194-226): no start clamp, accept on equality only, no best candidate
    WinSched* sched;
    // Work items ("pieces") of a launch: a peak's window swept over one range of RNA rows.  First sweeps and reverse
    // sweeps have one piece per peak; a re-planned forward sweep has one piece per run of qualifying granules.
    int* pc_peak; int* pc_lo; int* pc_rows; int* pc_key; int pc_cap;
    int* list;         // [pc_cap] piece indices in key order
    // best cell per peak over its pieces, packed so that atomicMax implements the tie rules:
    // value << 36 | (0xFFF - column) << 24 | (0xFFFFFF - row); low 36 bits 0 = "no cell above the floor, value = plain maximum"
    unsigned long long* res64;
    // decoded result per peak: (best value, column, row, 1 if written by the literal emulation)
    int4* res;
    // geometry
    const uint8_t* codes; const SegDesc* segs; int tasks_per_seg;
    const uint8_t* rna_ssw; int m;
    const uint16_t* rna_sel;  // per lncRNA row: PRMT selector bytes of its code for the low (bits 0..7) and high (8..15) half (k_win_dp TAB)
    const int* cut_table;   // [256][4]  cut length per (peak score, round) — fastsim.h:210 evaluated in float32 on the host
    long long* cell_counter;
    const int* forced_cut;  // probe path: explicit window length per peak (nullptr in the product path)
    // saturated block maxima per granule of the scan stage (scan.cuh; nullptr: no pruning, every window streams the whole lncRNA)
    const uint16_t* gran_blk; int n_gran; int gran_rows; int blk_pitch; int n_pairs; int scan_r;
    // 1: the first sweep of round 0 only tracks cells that reach the peak score (what an accepted window needs, fastsim.h:218);
    // a window that stays below it gets its position from the re-planned sweep (fewer slow-path trips of the tracker)
    int floor_s;
};

// rows spanned by a positive-score local alignment over `cols` columns: < 2.25 * cols + 1
__host__ __device__ inline int win_margin(int cols) { return (9 * cols) / 4 + 2; }
// ... and by one that scores at least `s`: d diagonal steps and `ins` inserted rows give s <= 5d - 4*ins - 12, so
// rows = d + ins <= 2.25 * cols - (s + 12) / 4 (or just d <= cols without insertions)
__host__ __device__ inline int win_margin_for(int cols, int s) { return max(cols, (9 * cols - s - 12) / 4 + 1) + 1; }
__device__ inline int win_key(int len, int rows)
{
    const int g = (len + kWinR - 1) / kWinR;
    const int rb = min(kWinRowBuckets - 1, (rows + 63) >> 6);
    return (g - 1) * kWinRowBuckets + (kWinRowBuckets - 1 - rb);
}

// round >= 0, retry 0: forward plan for round `round` (row range from the bound L = peak score);
// round >= 0, retry 1: windows whose pruned result fell short of the value that proves it exact, re-planned with
//                      L = that result; round < 0: reverse plan over the chosen alignments.
// 8 threads cooperate on one peak (each reads the block maxima of every eighth granule).
__global__ void k_win_plan(const WinState w, int round, int retry)
{
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gt >> 3, sub = gt & 7;
    bool active = i < w.n_peaks;
    int len = 0, rows = 0, bound = 0, proven = 0, accept_floor = 0, ws = 0, shift = 0;
    if (active) {
        if (round >= 0) {
            const int sc = w.pk_score[i], pos = w.pk_pos[i];
            if (round == 0 && !retry && w.floor_s && w.gran_blk != nullptr) accept_floor = sc - 1;
            if (!retry) {
                if (round == 0) { if (sub == 0) { w.w_done[i] = 0; w.best_sw[i] = 0; w.fin_sw[i] = 0; w.w_next[i] = 0; w.w_probe[i] = 0; } }
                else if (w.w_done[i] || w.w_next[i] != round) active = false;
                int cut = w.forced_cut ? w.forced_cut[i] : w.cut_table[min(sc, 255) * 4 + round];
                if (w.compat) {
                    // fastSim.h:204-205: substr(max(pos - cut + 1, 0), cut) — a clamped window keeps its length (cut short only by
                    // the end of the segment) and so reaches beyond the peak column
                    ws = max(pos - cut + 1, 0);
                    shift = ws - (pos - cut + 1);
                    len = min(cut, w.segs[w.pk_task[i] / w.tasks_per_seg].len - ws);
                } else {
                    if (pos - cut + 1 <= 0) cut = pos + 1;             // fastsim.h:211
                    len = cut;
                    ws = pos - cut + 1;
                }
                bound = sc;
                // rounds >= 1 look at a shorter window with the same right end: its best score cannot exceed the previous
                // round's (exact) one, which is therefore the better first guess (any guess is safe, exactness is verified)
                if (round > 0) { const int prev = w.res[i].x; if (prev > 0 && prev < bound) bound = prev; }
            } else {
                const int4 v = w.res[i];
                // not in flight / exact already (value proven and position known; a zero needs no position)
                if (w.w_done[i] || w.w_next[i] != round || (v.x >= w.w_bound[i] && (v.y != 0x7fffffff || v.x <= 0))) active = false;
                len = w.w_len[i];
                ws = w.w_ws[i]; shift = w.w_shift[i];
                bound = max(v.x, 0);
                proven = bound;              // a cell with this value exists: anything below it is irrelevant now
            }
        } else {
            // reverse plans: round -1 = every chosen alignment whose begin is not known yet, round -2 = reverse probes
            if (w.fin_sw[i] <= 0) active = false;
            if (round == -1 ? (w.w_probe[i] == 2) : (w.w_probe[i] != 1 || w.w_done[i])) active = false;
            len = w.fin_re[i] + 1;
        }
    }
    // granule bound scan (forward plans only): which granules can hold a cell >= bound inside the window's columns
    // (first/last klo/khi and the set as a bit mask); `outside` = the largest bound of any granule left out (a result
    // above it is exact)
    int klo = -1, khi = -1, outside = 0;
    unsigned long long qmask = 0ull;
    // Sweeps are split into runs of qualifying granules (far-apart granules are not bridged); gaps of at most `join` mask
    // bits are swept with their neighbours, so their granules do not count as left out.
    const int gq = (w.n_gran + 63) >> 6;                // granules per mask bit (1 up to 8192 rows; long lncRNAs share bits)
    const int bit_rows = gq * w.gran_rows;
    const int join = (win_margin(len) + bit_rows - 1) / bit_rows;           // a gap this short would be covered by the next margin anyway
    if (round >= 0 && w.gran_blk != nullptr) {
        const bool scan = active && bound > 0;
        int task = 0;
        if (scan) task = w.pk_task[i];
        const TaskDef td = c_tasks[task % w.tasks_per_seg];
        const uint16_t* base = w.gran_blk + ((size_t)(task / w.tasks_per_seg) * w.n_pairs + td.pair) * w.n_gran * w.blk_pitch;
        const int lo_col = ws, pos = ws + len - 1;
        const int grp = (threadIdx.x & 31) & ~7;       // first lane of this peak's 8 threads
        int pending = 0;        // largest bound among the non-qualifying granules after the last qualifying one
        // (after round 0 most peaks are finished: a warp whose four peaks have nothing to scan skips the walk altogether)
        const int n_walk = __any_sync(0xffffffffu, scan) ? w.n_gran : 0;
        for (int k0 = 0; k0 < n_walk; k0 += 8) {
            // thread `sub` reads the bound of granule k0 + sub: the maximum of the (skewed) 16-column blocks that cover the window
            int mine_b = 0;
            if (scan && k0 + sub < w.n_gran) {
                const int tl = gran_tail_lane(k0 + sub, w.scan_r);
                const uint16_t* row = base + (size_t)(k0 + sub) * w.blk_pitch;
                for (int b = blk_of(lo_col, tl); b <= blk_of(pos, tl); ++b) { const int x = row[b]; mine_b = max(mine_b, td.half ? (x >> 8) : (x & 0xff)); }
            }
            // ... and all eight walk the granules in order
            for (int q = 0; q < 8 && k0 + q < w.n_gran; ++q) {
                const int k = k0 + q;
                const int b = __shfl_sync(0xffffffffu, mine_b, grp + q);
                if (scan && b >= bound) {
                    // the granules since the previous qualifying one are left out unless the gap is short enough to be joined
                    if (klo < 0) { klo = k; outside = pending; }
                    else if (k / gq - khi / gq - 1 > join) outside = max(outside, pending);
                    pending = 0;
                    khi = k;
                    qmask |= 1ull << (k / gq);
                } else pending = max(pending, b);
            }
        }
        outside = max(outside, pending);
    }
    // From here on one thread per peak works (sub == 0 of an active peak); the others only take part in the warp-wide
    // reservation of piece slots (one atomic per warp instead of one per peak).
    const bool mine = active && sub == 0;
    int floor_v = max(max(proven - 1, 0), accept_floor);
    // only cells above `outside` matter (a result r > outside is exact), and their alignments span at most `margin` rows
    const int margin = win_margin_for(len, outside + 1);
    // walks the runs of qualifying granules (gaps <= join joined, at most 4 runs); f(lo, rows) per run; returns their number
    auto for_runs = [&](auto&& f) -> int {
        const int first = klo / gq;
        int nr = 0, cur_lo = first, cur_hi = first;
        unsigned long long rest = qmask & ~(1ull << first);
        for (;;) {
            int nxt = -1;
            if (rest) { nxt = __ffsll((long long)rest) - 1; rest &= rest - 1; }
            if (nxt >= 0 && (nxt - cur_hi - 1 <= join || nr == 3)) { cur_hi = nxt; continue; }
            const int lo = max(0, cur_lo * bit_rows - margin);
            f(lo, min(w.m - 1, (cur_hi + 1) * bit_rows - 1) - lo + 1);
            ++nr;
            if (nxt < 0) break;
            cur_lo = cur_hi = nxt;
        }
        return nr;
    };
    int np = 0;
    bool pruned = false;
    if (mine) {
        np = 1;
        if (round >= 0 && khi >= 0) {
            long long total = 0;
            const int nr = for_runs([&](int, int rows) { total += rows; });
            if (total < w.m) { pruned = true; np = nr; }
        }
    }
    // reserve np slots per thread: warp-inclusive scan, then one atomicAdd per BLOCK (every warp's total goes through shared memory)
    __shared__ int s_wtot[32], s_wbase[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    int incl = np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_wtot[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < nwarps; ++k) { s_wbase[k] = tot; tot += s_wtot[k]; }
        const int b0 = tot > 0 ? atomicAdd(&w.sched->n_pieces, tot) : 0;      // (pc_cap = 4 * n_peaks: cannot overflow)
        for (int k = 0; k < nwarps; ++k) s_wbase[k] += b0;
    }
    __syncthreads();
    const int base = s_wbase[wid];
    const int copy = blockIdx.x & (kWinCopies - 1);
    unsigned long long cells = 0;
    if (mine) {
        int p = base + incl - np;
        w.w_len[i] = len;
        if (round >= 0) { w.w_ws[i] = ws; w.w_shift[i] = shift; }
        w.w_flight[i] = 1;
        w.res64[i] = 0ull;
        auto emit = [&](int lo, int rows) {
            const int key = win_key(len, rows);
            w.pc_peak[p] = i; w.pc_lo[p] = lo; w.pc_rows[p] = rows; w.pc_key[p] = key | (copy << 16);
            ++p;
            atomicAdd(&w.sched->hist[copy][key], 1);
            cells += (unsigned long long)len * (unsigned long long)rows;
        };
        if (round < 0) emit(0, min(w.fin_qe[i] + 1, win_margin(len)));
        else {
            if (pruned) { for_runs(emit); bound = outside + 1; floor_v = max(floor_v, outside); }
            else { bound = 0; emit(0, w.m); }
            w.w_bound[i] = bound; w.w_floor[i] = floor_v;
        }
    }
    // statistics: one set of global atomics per block
    __shared__ unsigned long long s_cells, s_cnt;
    if (threadIdx.x == 0) { s_cells = 0ull; s_cnt = 0ull; }
    __syncthreads();
    unsigned long long cnt = mine ? 1ull : 0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) { cells += __shfl_xor_sync(0xffffffffu, cells, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if (lane == 0 && cnt) { atomicAdd(&s_cells, cells); atomicAdd(&s_cnt, cnt); }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) {
        const int slot = round >= 0 ? round * 2 + retry : (round == -1 ? 8 : 9);
        atomicAdd(&w.sched->st_windows[slot], s_cnt);
        atomicAdd(&w.sched->st_cells[slot], s_cells);
        if (round >= 0 && w.cell_counter) atomicAdd((unsigned long long*)w.cell_counter, s_cells);
    }
}

// exclusive prefix of the key histogram, per-class counts / list offsets / bin ranges (one block of 1024 threads)
__global__ void __launch_bounds__(1024) k_win_offsets(WinSched* sc)
{
    __shared__ int s_warp[32];
    __shared__ int s_cls[32];
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    int v0 = 0, v1 = 0;
    for (int c = 0; c < kWinCopies; ++c) { v0 += sc->hist[c][2 * tid]; v1 += sc->hist[c][2 * tid + 1]; }
    int x = v0 + v1;
    const int mine = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[wp] = x;
    __syncthreads();
    if (wp == 0) {
        const int t = s_warp[lane];
        int y = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int z = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += z; }
        s_warp[lane] = y - t;
        s_cls[lane] = t;                 // one warp of this block covers exactly one class (64 keys = 32 threads x 2)
    }
    __syncthreads();
    const int excl = s_warp[wp] + x - mine;
    for (int c = 0, a0 = excl, a1 = excl + v0; c < kWinCopies; ++c) {
        sc->off[c][2 * tid] = a0; sc->off[c][2 * tid + 1] = a1;
        a0 += sc->hist[c][2 * tid]; a1 += sc->hist[c][2 * tid + 1];
        sc->fill[c][2 * tid] = 0; sc->fill[c][2 * tid + 1] = 0;
        sc->hist[c][2 * tid] = 0; sc->hist[c][2 * tid + 1] = 0;       // ready for the next plan
    }
    if (wp == 0) {
        const int cnt = s_cls[lane];
        const int g = lane + 1, wpw = 2 * (32 / g);
        const int bins = (cnt + wpw - 1) / wpw;
        int y = bins;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int z = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += z; }
        sc->cls_count[lane] = cnt;
        sc->cls_off[lane] = s_warp[lane];
        sc->bin_start[lane] = y - bins;
        if (lane == 31) { sc->total_bins = y; sc->bin_counter = 0; }
    }
}

__global__ void k_win_place(const WinState w)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= w.sched->n_pieces) return;
    const int key = w.pc_key[p] & 0xFFFF, copy = w.pc_key[p] >> 16;
    w.list[w.sched->off[copy][key] + atomicAdd(&w.sched->fill[copy][key], 1)] = p;
}

// best cell of every peak that had pieces in the launch: unpack res64 into res
__global__ void k_win_combine(const WinState w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks || w.w_flight[i] != 1) return;
    w.w_flight[i] = 0;
    const unsigned long long v = w.res64[i];
    const int val = (int)(v >> 36);
    if ((v & 0xFFFFFFFFFull) == 0) w.res[i] = make_int4(val, 0x7fffffff, 0, 0);       // nothing above the floor: value only
    else w.res[i] = make_int4(val, 0xFFF - (int)((v >> 24) & 0xFFF), 0xFFFFFF - (int)(v & 0xFFFFFF), 0);
}

// signed 16-bit half of a packed register, extracted with one opaque bit-field instruction: written with C casts the
// compiler types the packed value as a pair of shorts and then pays an identity PRMT for every packed-SIMD use of it
__device__ __forceinline__ int half_s16(uint32_t v, int h)
{
    int r;
    if (h) asm("bfe.s32 %0, %1, 16, 16;" : "=r"(r) : "r"(v));
    else asm("bfe.s32 %0, %1, 0, 16;" : "=r"(r) : "r"(v));
    return r;
}

// TAB: score lookup with one PRMT per cell pair (needs an lncRNA made of A/C/G/T/U only); otherwise XNOR + VIADDMNMX.
// Q4CHK: the sweep also watches the vertical gap state (E[] here; the reference's vF) where it crosses into a row that starts a
// stripe of the reference's 16-lane layout of THIS alignment call (forward: the whole lncRNA, stripe = ceil(m / 16) rows; reverse:
// the reversed prefix that ends at the forward end row).  Without a value >= 132 there the reference's signed lazy-F test behaves
// like an unsigned one for every cell the result depends on (they all lie inside the swept rows, whose derivations never leave
// them; cells outside can only be LOWER in the reference, and were below the result already), so the exact result is the
// reference's and the window needs no literal emulation.
template <bool REV, bool TAB, bool Q4CHK = false>
__global__ void __launch_bounds__(128) k_win_dp(const WinState w)
{
    constexpr int R = kWinR;
    const int lane = threadIdx.x & 31;
    const uint32_t kNegOpen = 0xFFF0FFF0u, kNegExt = 0xFFFCFFFCu, kSix = 0x00060006u, kMis = 0xFFFCFFFCu;
    const WinSched* sc = w.sched;
    const int total = sc->total_bins;
    const int my_bin_start = sc->bin_start[lane];
    const bool my_cls_used = sc->cls_count[lane] > 0;

    for (;;) {
        int bin = 0;
        if (lane == 0) bin = atomicAdd(&w.sched->bin_counter, 1);
        bin = __shfl_sync(0xffffffffu, bin, 0);
        if (bin >= total) break;
        // class of the bin: the last class whose first bin is <= bin and that is not empty (empty classes share their start
        // with the next one, so the highest qualifying lane is the right one)
        const int c = 31 - __clz(__ballot_sync(0xffffffffu, my_bin_start <= bin && my_cls_used));
        const int b = bin - __shfl_sync(0xffffffffu, my_bin_start, c);
        const int g = c + 1, gpw = 32 / g;             // lanes per group, groups per half-warp
        const int cls_count = sc->cls_count[c], cls_off = sc->cls_off[c];
        const bool idle = lane >= gpw * g;             // lanes beyond the last whole group carry nothing
        const int grp = idle ? gpw : lane / g, lig = idle ? lane - gpw * g : lane - grp * g;
        const bool leader = (lig == 0);

        // per-half window description for this lane's group
        int wi[2], len[2], slen[2], sbase[2], sdir[2];
        // per owned column: the scaled base codes of the two windows (XNOR path), or (TAB) two 4-entry score tables — byte x of
        // ta[r] / tb[r] is the score of RNA code x against the column of the low / high half's window
        uint32_t dq[TAB ? 1 : R], ta[TAB ? R : 1], tb[TAB ? R : 1];
#pragma unroll
        for (int r = 0; r < (TAB ? 1 : R); ++r) dq[r] = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = (b * 2 + h) * gpw + grp;
            const int piece = (!idle && slot < cls_count) ? w.list[cls_off + slot] : -1;
            wi[h] = piece >= 0 ? w.pc_peak[piece] : -1;
            len[h] = 0; slen[h] = 0; sbase[h] = 0; sdir[h] = 1;
            int colcode[R];
#pragma unroll
            for (int r = 0; r < R; ++r) colcode[r] = 80;
            if (wi[h] >= 0) {
                const int i = wi[h];
                const int task = w.pk_task[i];
                const SegDesc sd = w.segs[task / w.tasks_per_seg];
                const TaskDef td = c_tasks[task % w.tasks_per_seg];
                const int L = w.w_len[i];
                len[h] = L;
                const int ws = REV ? w.fin_ws[i] : w.w_ws[i];              // window start in seq2 coordinates
                if (REV) { slen[h] = min(w.fin_qe[i] + 1, win_margin(L)); sbase[h] = w.fin_qe[i]; sdir[h] = -1; }
                else { slen[h] = w.pc_rows[piece]; sbase[h] = w.pc_lo[piece]; sdir[h] = 1; }
                const int off = g * R - L;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int cc = lig * R + r - off;
                    if (cc >= 0) {
                        const int q = REV ? (ws + w.fin_re[i] - cc) : (ws + cc);
                        const int x = w.codes[sd.start + (td.reversed ? (sd.len - 1 - q) : q)];
                        const int d = td.img[x];
                        colcode[r] = d < 4 ? d * 16 : 80;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (TAB) {
                    // 0xFC = -4 everywhere, 0x05 at the byte of the column's own code (pad / N columns: no match at all)
                    const uint32_t tab = colcode[r] < 64 ? (0xFCFCFCFCu ^ (0xF9u << (colcode[r] >> 1))) : 0xFCFCFCFCu;
                    if (h == 0) ta[r] = tab; else tb[r] = tab;
                } else dq[r] |= (uint32_t)colcode[r] << (16 * h);
            }
        }
        int nsteps = max(slen[0], slen[1]);
#pragma unroll
        for (int o = 16; o; o >>= 1) nsteps = max(nsteps, __shfl_xor_sync(0xffffffffu, nsteps, o));
        nsteps += g - 1;

        // result tracker: starts at the floor below which nothing matters (reverse: forward score - 1; forward: w_floor);
        // `runmax` (one instruction per step) keeps the plain maximum, the proven lower bound a re-plan starts from
        int best[2], bcol[2], brow[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            best[h] = wi[h] >= 0 ? (REV ? w.fin_sw[wi[h]] - 1 : w.w_floor[wi[h]]) : 0;
            bcol[h] = 0x7fffffff; brow[h] = 0;
        }
        uint32_t trigm1 = pack16(max(best[0], 0), max(best[1], 0));
        uint32_t runmax = 0;
        const uint32_t keep = leader ? 0u : 0xffffffffu;

        uint32_t Hd[R], E[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { Hd[r] = 0; E[r] = 0; }
        // The stream symbol travels with the wavefront.  XNOR path: the scaled codes of the two halves' RNA rows (64 = none).
        // TAB path: a ready PRMT selector — for code x of the low half the nibbles (x, x|8) pick the score byte and its sign
        // extension out of ta[r], for the high half (4+x, 4+x|8) out of tb[r]; "none" (past the end of a stream) selects two
        // sign bytes, i.e. a score of -1 or 0, which like -4 can never raise a maximum.
        const uint32_t kNoneLo = TAB ? 0x88u : 64u, kNoneHi = TAB ? 0xCCu : 64u;
        uint32_t hout = 0, fout = 0, xout = kNoneLo | (kNoneHi << (TAB ? 8 : 16)), hdiag = 0;
        // leader prefetch of the stream symbol: only group leaders have a stream (length 0 elsewhere), so the loads are plain
        // predicated instructions — no divergent branch per step
        const int ln0 = leader ? slen[0] : 0, ln1 = leader ? slen[1] : 0;
        auto fetch = [&](int s) -> uint32_t {
            if (TAB) {
                // both selector bytes of a row are precomputed (rna_sel); the low half takes byte 0 of its row's entry, the high
                // half byte 1 of its own row's entry; past the end of a stream the "none" selector stays
                uint32_t v0 = 0xCC88u, v1 = 0xCC88u;
                if (s < ln0) v0 = __ldg(w.rna_sel + (uint32_t)(sbase[0] + sdir[0] * s));       // (unsigned index: one wide multiply-add)
                if (s < ln1) v1 = __ldg(w.rna_sel + (uint32_t)(sbase[1] + sdir[1] * s));
                return __byte_perm(v0, v1, 0x4450);           // byte 0 of v0, byte 1 of v1
            }
            uint32_t c0 = kNoneLo, c1 = kNoneHi;
            if (s < ln0) { const uint32_t q = w.rna_ssw[sbase[0] + sdir[0] * s]; c0 = q < 4 ? q * 16 : 64u; }
            if (s < ln1) { const uint32_t q = w.rna_ssw[sbase[1] + sdir[1] * s]; c1 = q < 4 ? q * 16 : 64u; }
            return c0 | (c1 << 16);
        };
        // Q4CHK: step at which this lane's row of half h is the LAST row of a stripe (E[] after that step enters the stripe start)
        int nq[2] = {0x7fffffff, 0x7fffffff}, ql[2] = {1, 1}, q4seen = 0;
        if (Q4CHK) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (wi[h] < 0) continue;
                const int L = REV ? (w.fin_qe[wi[h]] + 16) / 16 : (w.m + 15) / 16;
                const int b = REV ? 0 : sbase[h];                  // index, in the read the reference aligns, of stream position 0
                nq[h] = lig + (2 * L - 1 - (b % L)) % L;           // first position t with (b + t + 1) % L == 0
                ql[h] = L;
            }
        }
        int nqmin = min(nq[0], nq[1]);
        uint32_t xnext = fetch(0);
#pragma unroll 2
        for (int s = 0; s < nsteps; ++s) {
            const uint32_t myx = xnext;
            xnext = fetch(s + 1);
            uint32_t hin = __shfl_up_sync(0xffffffffu, hout, 1) & keep;
            uint32_t fin = __shfl_up_sync(0xffffffffu, fout, 1) & keep;
            uint32_t xin = __shfl_up_sync(0xffffffffu, xout, 1);
            xin = leader ? myx : xin;
            uint32_t d = hdiag, f = fin, hlast = 0;
            uint32_t t[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t sc2;
                if (TAB) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sc2) : "r"(ta[r]), "r"(tb[r]), "r"(xin));     // +5 / -4 looked up and sign-extended (selector bit 3)
                else sc2 = __viaddmax_s16x2(~(xin ^ dq[r]), kSix, kMis);      // +5 on equal codes, -4 otherwise
                t[r] = __viaddmax_s16x2_relu(d, sc2, E[r]);
                const uint32_t u = __vadd2(t[r], kNegOpen);
                E[r] = __viaddmax_s16x2(E[r], kNegExt, u);
                const uint32_t hh = __vmaxs2(t[r], f);
                f = __viaddmax_s16x2(f, kNegExt, u);
                d = Hd[r];
                Hd[r] = hh;
                hlast = hh;
            }
            uint32_t ms = t[0];
#pragma unroll
            for (int r = 1; r + 1 < R; r += 2) ms = __vimax3_s16x2(ms, t[r], t[r + 1]);
            if ((R & 1) == 0) ms = __vmaxs2(ms, t[R - 1]);
            if (!REV) runmax = __vmaxs2(runmax, ms);
            if (Q4CHK && s == nqmin) {
                uint32_t e = E[0];
#pragma unroll
                for (int r = 1; r < R; ++r) e = __vmaxs2(e, E[r]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (s != nq[h]) continue;
                    if (s - lig + 1 < slen[h] && half_s16(e, h) >= kQ4CarryF) q4seen |= 1 << h;      // (the stripe start itself is streamed)
                    nq[h] += ql[h];
                }
                nqmin = min(nq[0], nq[1]);
            }
            if (__vmaxs2(trigm1, ms) != trigm1) {
                const int row = s - lig;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (wi[h] < 0 || row < 0 || row >= slen[h]) continue;
                    const int off = g * R - len[h];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int cc = lig * R + r - off;
                        const int v = half_s16(t[r], h);
                        if (cc >= 0 && (v > best[h] || (v == best[h] && cc < bcol[h]))) { best[h] = v; bcol[h] = cc; brow[h] = row; }
                    }
                }
                // ties with an ACHIEVED best matter (smaller column wins); an unreached floor only needs strictly larger cells
                if (!REV) trigm1 = pack16(max(best[0] - (bcol[0] != 0x7fffffff), 0), max(best[1] - (bcol[1] != 0x7fffffff), 0));
            }
            hdiag = hin;
            hout = hlast; fout = f; xout = xin;
        }
        if (Q4CHK && q4seen) {
#pragma unroll
            for (int h = 0; h < 2; ++h) if ((q4seen >> h) & 1) atomicOr(&w.w_q4[wi[h]], REV ? 2 : 1);
        }
        // group reduction: highest value, then smallest column (lower lanes own smaller columns)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int rmx = half_s16(runmax, h);
            for (int o = 1; o < g; o <<= 1) {
                const int ob = __shfl_down_sync(0xffffffffu, best[h], o);
                const int oc = __shfl_down_sync(0xffffffffu, bcol[h], o);
                const int orow = __shfl_down_sync(0xffffffffu, brow[h], o);
                const int orm = __shfl_down_sync(0xffffffffu, rmx, o);
                if (lig + o < g) {
                    rmx = max(rmx, orm);
                    if (ob > best[h] || (ob == best[h] && oc < bcol[h])) { best[h] = ob; bcol[h] = oc; brow[h] = orow; }
                }
            }
            // forward: absolute RNA row; reverse: index into the reversed stream (k_win_finish subtracts it from qe).
            // A sweep in which no cell rose above the floor reports its plain maximum (position unknown, never used: such
            // a result is below w_bound, so the window is re-planned from it).  atomicMax over the peak's pieces = highest
            // value, then smallest column, then smallest row.
            if (leader && wi[h] >= 0) {
                unsigned long long key;
                if (bcol[h] == 0x7fffffff) key = (unsigned long long)max(REV ? best[h] : rmx, 0) << 36;
                else key = ((unsigned long long)best[h] << 36) | ((unsigned long long)(0xFFF - bcol[h]) << 24) |
                           (unsigned long long)(0xFFFFFF - (REV ? brow[h] : sbase[h] + brow[h]));
                atomicMax(&w.res64[wi[h]], key);
            }
        }
    }
}

// forward decision of fastSIM's loop (fastsim.h:216-250) for round `round` (0..3)
__global__ void k_win_decide(const WinState w, int round)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks || w.w_done[i] || w.w_next[i] != round) return;
    const int cut = w.w_len[i];
    const int4 v = w.res[i];
    int sw = 0, re = 0, qe = 0;
    if (v.x > 0) { sw = v.x; re = v.y; qe = v.z; }
    const int S = w.pk_score[i];
    const int ws = w.w_ws[i];
    if (w.compat) {
        // fastSim.h:203-210: the loop only stops on equality; whatever the last window gave is converted (an exact 0 stays 0 in
        // every nested window)
        if (sw == S || round == 3 || (sw == 0 && v.w == 0)) {
            w.fin_sw[i] = sw; w.fin_cut[i] = cut; w.fin_re[i] = re; w.fin_qe[i] = qe; w.fin_ws[i] = ws; w.fin_shift[i] = w.w_shift[i]; w.w_done[i] = 1;
        } else w.w_next[i] = round + 1;
        return;
    }
    w.fin_shift[i] = 0;
    if (sw >= S) {
        w.fin_sw[i] = sw; w.fin_cut[i] = cut; w.fin_re[i] = re; w.fin_qe[i] = qe; w.fin_ws[i] = ws; w.w_done[i] = 1;
        return;
    }
    if (sw > w.best_sw[i] && re == cut - 1) {
        w.best_sw[i] = sw; w.best_cut[i] = cut; w.best_re[i] = re; w.best_qe[i] = qe; w.best_ws[i] = ws;
        // The later rounds look at nested, shorter windows with the same right end, so their exact scores cannot exceed
        // this one: they can neither reach S (> sw) nor replace this candidate (needs a strictly larger score), and the
        // loop ends with this candidate (fastsim.h:236-249).  Only when this score came from the literal emulation (which may
        // report less than exact SW, res.w = 1) a later round could still overtake it, so those keep going.
        if (v.w == 0) {
            w.fin_sw[i] = sw; w.fin_cut[i] = cut; w.fin_re[i] = re; w.fin_qe[i] = qe; w.fin_ws[i] = ws; w.w_done[i] = 1;
            return;
        }
    }
    // an exact score of 0 stays 0 in every nested window: nothing will be accepted or become a candidate any more
    const bool dead_end = (sw == 0 && v.w == 0);
    if (round == 3 || dead_end) {
        if (w.best_sw[i] > 0) { w.fin_sw[i] = w.best_sw[i]; w.fin_cut[i] = w.best_cut[i]; w.fin_re[i] = w.best_re[i]; w.fin_qe[i] = w.best_qe[i]; w.fin_ws[i] = w.best_ws[i]; }
        else { w.fin_sw[i] = sw; w.fin_cut[i] = cut; w.fin_re[i] = re; w.fin_qe[i] = qe; w.fin_ws[i] = ws; }
        w.w_done[i] = 1;
        return;
    }
    w.w_next[i] = round + 1;
    // exact result, no candidate so far: park it in fin_* for the reverse probe (k_win_probe) that may skip rounds
    if (v.w == 0 && w.best_sw[i] == 0) {
        w.fin_sw[i] = sw; w.fin_cut[i] = cut; w.fin_re[i] = re; w.fin_qe[i] = qe; w.fin_ws[i] = ws;
        w.w_probe[i] = 1;
    }
}

// Round skipping (exact).  Round k found its best cell e (score s < S, not in the last column) and the reverse probe the
// latest start column rb of an alignment that reaches s at e.  A later round looks at the last cut' columns of the same
// window.  While that window still contains column rb, the alignment fits, so the window's maximum is still s, its
// maximal cells are a subset of round k's and contain e, hence the tie rules pick e again: the round reports the same
// alignment, which still misses S and the last column and so changes nothing (fastsim.h:218-235).  The peak jumps to
// the first round whose window starts after rb; if there is none, the last round's alignment is this one (fastsim.h:253).
__global__ void k_win_probe(const WinState w, int round)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks || w.w_done[i] || w.w_probe[i] != 1) return;
    w.w_probe[i] = 0;
    const int S = w.fin_sw[i];
    const int4 v = w.res[i];
    if (v.x != S) return;                        // (cannot happen for an exact forward score) take every round
    const int rb = w.fin_re[i] - v.y, qb = w.fin_qe[i] - v.z;
    const int pos = w.pk_pos[i], sc = w.pk_score[i];
    const int start_col = w.fin_ws[i] + rb;                  // in translated-segment coordinates
    int k = round + 1;
    for (; k <= 3; ++k) {
        int cut = w.forced_cut ? w.forced_cut[i] : w.cut_table[min(sc, 255) * 4 + k];
        if (pos - cut + 1 <= 0) cut = pos + 1;
        if (pos - cut + 1 > start_col) break;                 // this window no longer holds the alignment
    }
    if (k > 3) { w.fin_rb[i] = rb; w.fin_qb[i] = qb; w.w_probe[i] = 2; w.w_done[i] = 1; }
    else w.w_next[i] = k;
}

// reverse result: beginning of the alignment (sswNew.cpp:1518-1520)
__global__ void k_win_finish(const WinState w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks || w.fin_sw[i] <= 0 || w.w_probe[i] == 2) return;
    const int S = w.fin_sw[i];
    const int4 v = w.res[i];
    int col = -1, row = 0;
    if (v.x == S) { col = v.y; row = v.z; }
    if (col < 0) {
        // cannot happen for an exact forward score; a literal (Q4) forward score is resolved by the literal reverse pass
        if (S < kQ4Guard) w.fin_sw[i] = 0;
        w.fin_rb[i] = 0; w.fin_qb[i] = 0;
        return;
    }
    w.fin_rb[i] = w.fin_re[i] - col;
    w.fin_qb[i] = w.fin_qe[i] - row;
}

// ---------------------------------------------------------------------------------------------
// banded_sw (sswNew.cpp:1071-1259) + getAlignment expansion (fastsim.h:416-560) + the identity / stability
// arithmetic of convertMyTriplex (fastsim.h:323-383), one thread per alignment.  An alignment is described by a
// self-contained TraceJob, so the same kernel serves two passes:
//   pass 1  every chosen alignment of a batch: nt, identity, stability only (what the host needs to de-duplicate,
//           rank and filter — fastsim.h:273-288, Fasim-LongTarget.cpp:589-597);
//   pass 2  the few alignments that survive those filters: the TFO / TTS strings, written at host-assigned offsets.
// Per-thread scratch: three int rows of (2*bw+5), the direction bytes 3*(2*bw+1)*readLen, and the op / string
// staging area.  Alignments whose band outgrows the scratch are flagged (status 2) and re-run by a second launch
// with a large scratch.
struct TraceJob {
    long long seg_start;       // segment offset in the record
    int seg_len, tdef;         // segment length; index into the task table (bits 0..7) | lowercase compat: the coordinate shift << 8
    int ws;                    // window start in seq2 (translated-segment) coordinates
    int rb, re, qb, qe;        // alignment box: window-relative columns, lncRNA rows
    int score;                 // sw_score (<= 0: nothing to do)
    long long out_off;         // pass 2: offset of the string pair in the pool
};

struct TraceOut {
    int status;                // 0 none, 1 ok, 2 needs a larger scratch, 3 traceback left the band (sw_score -> 0),
                               // 4 dead: not traced, can never be reported, takes part in the de-duplication (k_make_trace_jobs)
    int nt;                    // alignment columns (status 4: a lower bound that passes the ntMin gate)
    float identity, tri;       // MeanIdentity(%) / MeanStability, float32 evaluated exactly as fastsim.h:323-383
};

struct TraceArgs {
    const TraceJob* jobs; int n_jobs;
    const uint8_t* codes;          // base codes of the record
    const unsigned char* dna;      // raw record bytes (for the TTS string)
    const uint8_t* rna_ssw;        // lncRNA, SSW codes
    const unsigned char* rna_raw;  // raw lncRNA bytes (for the TFO string)
    unsigned char* scratch; long long scratch_per_thread;
    int skip_dead;                             // pass 1: jobs preset to status 4 by k_make_trace_jobs are left alone
    const int* in_list; const int* in_count;   // work list (nullptr: all jobs)
    int* out_list; int* out_count;             // jobs this launch could not finish (status 2) — input of the next tier
    int nt_min, nt_max, penalty_t, penalty_c;
    TraceOut* out;                 // per job
    char* strpool;                 // pass 2 only (nullptr in pass 1): tfo at out_off, tts at out_off + nt + 1, both NUL-terminated
};

// Alignments that provably cannot be reported skip the traceback ("dead", status 4).  A row is reported only if it passes
// nt >= max(ntMin, cLength), identity >= minIdentity and stability >= minStability (fastsim.h:284-288,
// Fasim-LongTarget.cpp:589-597).  Two necessary conditions need no alignment:
//  (i)  nt = refLen + readLen - #M and #M >= ceil(score / 5), so nt <= refLen + readLen - ceil(score / 5);
//  (ii) every base of ref[rb..re] occupies a column, so two adjacent T's of the source strand always meet as consecutive
//       non-gap source characters and fire the penaltyT rule (fastsim.h:363-367): with penaltyT < 0 the stability sum is
//       at most max(4.5, pc) * (nt - 1) + max(pc, 0) * nt + 2 * penaltyT, i.e. the mean is below
//       max(4.5, pc) + max(pc, 0) + 2 * penaltyT / (refLen + readLen).
// A dead alignment still takes part in its task's de-duplication (which looks at coordinates and score only), provided it
// is certain to pass convertMyTriplex's nt >= ntMin gate (nt >= max(refLen, readLen)); otherwise it is traced normally.
struct DeadRule { int need_nt, nt_min, nt_max, penalty_t, penalty_c; float min_st; const unsigned char* dna; };

__global__ void k_make_trace_jobs(const WinState w, TraceJob* jobs, TraceOut* tout, const DeadRule dr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks) return;
    TraceJob J;
    const int task = w.pk_task[i];
    const SegDesc sd = w.segs[task / w.tasks_per_seg];
    const TaskDef td = c_tasks[task % w.tasks_per_seg];
    J.seg_start = sd.start; J.seg_len = sd.len; J.tdef = task % w.tasks_per_seg;
    J.score = w.fin_sw[i];
    J.ws = 0; J.rb = J.re = J.qb = J.qe = 0; J.out_off = 0;
    TraceOut o; o.status = 0; o.nt = 0; o.identity = 0.0f; o.tri = 0.0f;
    if (J.score > 0) {
        J.ws = w.fin_ws[i];
        J.tdef |= w.fin_shift[i] << 8;
        J.rb = w.fin_rb[i]; J.re = w.fin_re[i]; J.qb = w.fin_qb[i]; J.qe = w.fin_qe[i];
        const int refLen = J.re - J.rb + 1, readLen = J.qe - J.qb + 1;
        if (dr.dna != nullptr && w.fin_shift[i] == 0 && max(refLen, readLen) >= dr.nt_min) {
            bool dead = refLen + readLen - (J.score + 4) / 5 < dr.need_nt;
            if (!dead && dr.penalty_t < 0 && refLen + readLen <= dr.nt_max) {     // (nt > ntMax would zero the stability instead)
                const float pcf = (float)dr.penalty_c;
                const float ub = fmaxf(4.5f, pcf) + fmaxf(pcf, 0.0f) + 2.0f * (float)dr.penalty_t / (float)(refLen + readLen);
                if (ub < dr.min_st - 0.05f) {
                    // adjacent T's of the source strand = adjacent T's (A's for the complemented strands) of the record
                    const int q0 = J.ws + J.rb, q1 = J.ws + J.re;
                    const int g0 = td.reversed ? sd.len - 1 - q1 : q0;
                    const unsigned char want = td.comp_src ? 'A' : 'T';
                    const unsigned char* p = dr.dna + sd.start + g0;
                    for (int k = 0; k + 1 < refLen; ++k) if (p[k] == want && p[k + 1] == want) { dead = true; break; }
                }
            }
            if (dead) { o.status = 4; o.nt = max(refLen, readLen); }
        }
    }
    jobs[i] = J;
    tout[i] = o;
}

__device__ inline int band_u(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x + 1; }
__device__ inline int band_d(int w, int i, int j, int p) { int x = i - w; if (x < 0) x = 0; return (j - x) * 3 + p; }

__device__ inline char comp_char(unsigned char c)
{
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

// Hoogsteen / reverse-Hoogsteen stability of one (DNA base, RNA base) column — sim.h:72-97
__device__ inline float stability_dev(char dna, char rna, int para)
{
    if (para > 0) {
        if (dna == 'A' && rna == 'T') return 3.7f;
        if (dna == 'T' && rna == 'G') return 2.8f;
        if (dna == 'G') { if (rna == 'G') return 2.2f; if (rna == 'T') return 2.4f; if (rna == 'C') return 4.5f; }
        if (dna == 'C') { if (rna == 'T') return 2.6f; if (rna == 'C') return 2.4f; }
    } else {
        if (dna == 'A') { if (rna == 'A') return 3.0f; if (rna == 'T') return 3.5f; if (rna == 'C') return 1.0f; }
        if (dna == 'T' && rna == 'G') return 1.0f;
        if (dna == 'G') { if (rna == 'A') return 1.0f; if (rna == 'G') return 3.0f; if (rna == 'C') return 3.0f; }
        if (dna == 'C') { if (rna == 'T') return 2.0f; if (rna == 'C') return 1.0f; }
    }
    return 0.0f;
}

__global__ void k_traceback(const TraceArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    unsigned char* sc = a.scratch + (size_t)tid * a.scratch_per_thread;
    const int n_todo = a.in_list ? min(*a.in_count, a.n_jobs) : a.n_jobs;
    for (int k0 = tid; k0 < n_todo; k0 += nthreads) {
        const int i = a.in_list ? a.in_list[k0] : k0;
        if (!a.in_list) {
            if (a.skip_dead && a.out[i].status == 4) continue;
            a.out[i].status = 0;
        }
        const TraceJob J = a.jobs[i];
        if (J.score <= 0) continue;

        const TaskDef td = c_tasks[J.tdef & 0xFF];
        const int shift = J.tdef >> 8;      // lowercase compat: the strings are read `shift` columns left of the aligned ones (fastSim.h:211-212)
        const int ws = J.ws, rb = J.rb, re = J.re, qb = J.qb, qe = J.qe;
        const int refLen = re - rb + 1, readLen = qe - qb + 1, score = J.score;
        const uint8_t* gc = a.codes + J.seg_start;
        auto gidx = [&](int q) -> int { return td.reversed ? (J.seg_len - 1 - q) : q; };     // seq2 index -> segment index
        const int ntmax = refLen + readLen;

        int bw = abs(refLen - readLen) + 1, maxv = 0, width_d = 0;
        int8_t* dir = nullptr;
        bool fits = true;
        for (;;) {
            const int width = bw * 2 + 3;
            width_d = bw * 2 + 1;
            const long long need = 3LL * (width + 2) * 4 + 3LL * width_d * readLen + 3LL * (ntmax + 4) + 32;
            if (need > a.scratch_per_thread) { fits = false; break; }
            int* h_b = reinterpret_cast<int*>(sc);
            int* e_b = h_b + width + 2;
            int* h_c = e_b + width + 2;
            dir = reinterpret_cast<int8_t*>(h_c + width + 2);
            for (int j = 0; j < width + 2; ++j) { h_b[j] = 0; e_b[j] = 0; h_c[j] = 0; }
            for (int ii = 0; ii < readLen; ++ii) {
                const int beg = max(0, ii - bw), end = min(refLen - 1, ii + bw);
                const int edge = min(end + 1, width - 1);
                int f = 0, u = 0;
                h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
                int8_t* line = dir + (size_t)width_d * ii * 3;
                const int rc = a.rna_ssw[qb + ii];
                for (int j = beg; j <= end; ++j) {
                    u = band_u(bw, ii, j);
                    const int e = band_u(bw, ii - 1, j), b = band_u(bw, ii, j - 1), dd = band_u(bw, ii - 1, j - 1);
                    const int de = band_d(bw, ii, j, 0), df = band_d(bw, ii, j, 1), dh = band_d(bw, ii, j, 2);
                    int t1 = (ii == 0) ? -kGapOpen : h_b[e] - kGapOpen;
                    int t2 = (ii == 0) ? -kGapExt : e_b[e] - kGapExt;
                    e_b[u] = t1 > t2 ? t1 : t2;
                    line[de] = t1 > t2 ? 3 : 2;
                    t1 = h_c[b] - kGapOpen;
                    t2 = f - kGapExt;
                    f = t1 > t2 ? t1 : t2;
                    line[df] = t1 > t2 ? 5 : 4;
                    const int e1 = e_b[u] > 0 ? e_b[u] : 0, f1 = f > 0 ? f : 0;
                    t1 = e1 > f1 ? e1 : f1;
                    const int rf = td.img[gc[gidx(ws + rb + j)]];
                    t2 = h_b[dd] + ((rf == rc && rf < 4) ? kMatch : kMismatch);
                    h_c[u] = t1 > t2 ? t1 : t2;
                    if (h_c[u] > maxv) maxv = h_c[u];
                    if (t1 <= t2) line[dh] = 1;
                    else line[dh] = e1 > f1 ? line[de] : line[df];
                }
                for (int j = 1; j <= u; ++j) h_b[j] = h_c[j];
            }
            if (maxv >= score) break;
            bw *= 2;
        }
        if (!fits) { a.out[i].status = 2; if (a.out_list) a.out_list[atomicAdd(a.out_count, 1)] = i; continue; }

        // traceback (sswNew.cpp:1159-1238): ops come out end -> start; written backwards into `ops`
        unsigned char* ops = reinterpret_cast<unsigned char*>(dir) + 3LL * width_d * readLen;
        char* tfo = reinterpret_cast<char*>(ops + ntmax + 4);
        char* tts = tfo + ntmax + 4;
        int wp = ntmax + 2;
        int ii = readLen - 1, j = refLen - 1, plane = 2;
        long long line_off = (long long)width_d * (readLen - 1) * 3;
        bool bad = false;
        while (ii > 0) {
            const long long idx = line_off + band_d(bw, ii, j, plane);
            if (j < 0 || idx < 0 || idx >= 3LL * width_d * readLen) { bad = true; break; }
            const int dv = dir[idx];
            if (dv == 1) { --ii; --j; plane = 2; line_off -= (long long)width_d * 3; ops[--wp] = 0; }
            else if (dv == 2) { --ii; plane = 0; line_off -= (long long)width_d * 3; ops[--wp] = 1; }
            else if (dv == 3) { --ii; plane = 2; line_off -= (long long)width_d * 3; ops[--wp] = 1; }
            else if (dv == 4) { --j; plane = 1; ops[--wp] = 2; }
            else if (dv == 5) { --j; plane = 2; ops[--wp] = 2; }
            else { bad = true; break; }
            if (wp <= 1) { bad = true; break; }
        }
        if (bad) { a.out[i].status = 3; continue; }
        ops[--wp] = 0;      // closing rule (:1220-1238): the alignment always starts with one more M column
        const int nt = ntmax + 2 - wp;
        // expansion from the front exactly like getAlignment: q walks the translated DNA from ref_begin, p the RNA
        int q = ws + rb - shift, p = qb, match = 0;
        for (int k = 0; k < nt; ++k) {
            const int op = ops[wp + k];
            char rch = '-', sch = '-', tch = '-';
            if (op != 2) rch = (char)a.rna_raw[p++];
            if (op != 1) {
                if (q >= 0 && q < J.seg_len) {
                    const int gi = gidx(q);
                    const unsigned char raw = a.dna[J.seg_start + gi];
                    sch = td.comp_src ? comp_char(raw) : (char)raw;
                    const int dcode = td.img[gc[gi]];
                    tch = dcode < 4 ? "ACGT"[dcode] : 'N';
                } else { sch = 'N'; tch = 'N'; }          // (shifted outside the segment: the reference reads out of bounds there)
                ++q;
            }
            tfo[k] = rch; tts[k] = sch;
            if (tch == rch) ++match;
        }
        // identity and mean stability in float32, same operation order as convertMyTriplex (fastsim.h:335, 342-383);
        // explicit round-to-nearest intrinsics keep the compiler from contracting or re-associating anything
        const float identity = __fdiv_rn(__int2float_rn(100 * match), __int2float_rn(nt));
        float tri = 0.0f;
        if (nt >= a.nt_min && nt <= a.nt_max) {
            float prev_val = 0.0f;
            char prev_ch = 0;
            const float pt = __int2float_rn(a.penalty_t), pc = __int2float_rn(a.penalty_c);
            for (int k = 0; k < nt; ++k) {
                const char ch = tts[k];
                float val = stability_dev(ch, tfo[k], td.para);
                if (ch == prev_ch && ch == 'T') { tri = __fadd_rn(__fsub_rn(tri, prev_val), pt); val = pt; }
                if (ch == prev_ch && ch == 'C') { tri = __fadd_rn(__fsub_rn(tri, prev_val), pc); val = pc; }
                prev_val = val;
                if (ch != '-') prev_ch = ch;
                tri = __fadd_rn(tri, val);
            }
            tri = __fdiv_rn(tri, __int2float_rn(nt));
        }
                if (a.strpool) {
            char* o = a.strpool + J.out_off;
            for (int k = 0; k < nt; ++k) o[k] = tfo[k];
            o[nt] = 0;
            for (int k = 0; k < nt; ++k) o[nt + 1 + k] = tts[k];
            o[2 * nt + 1] = 0;
        }
        TraceOut o1; o1.status = 1; o1.nt = nt; o1.identity = identity; o1.tri = tri;
        a.out[i] = o1;
    }
}

// ---------------------------------------------------------------------------------------------
// Fast tiers of the traceback: the same banded_sw + expansion, with the whole working set of an alignment in shared
// memory (typical need: ~450 bytes).  Every thread owns BYTES contiguous bytes; BYTES/4 is odd, so threads that
// touch the same offset of their regions (the common case) hit 32 different banks.
// Differences to k_traceback are representational only: int16 score rows, the three direction planes of a band cell
// packed into one byte (bit0: E opened, bit1: F opened, bits 2..4: H code), the window's translated base codes cached
// next to them, no staging of the expanded strings.
// Anything that does not fit, and any traceback step that would leave the band of its row (where the reference reads
// whatever lies next to it in memory), is handed to the next tier through out_list.
template <int TPB, int BYTES>
__global__ void __launch_bounds__(TPB) k_traceback_fast(const TraceArgs a)
{
    static_assert(BYTES % 4 == 0 && (BYTES / 4) % 2 == 1, "per-thread region must be an odd number of words");
    extern __shared__ uint32_t tb_words[];
    unsigned char* const reg = reinterpret_cast<unsigned char*>(tb_words) + (size_t)threadIdx.x * BYTES;
    const int tid = blockIdx.x * TPB + threadIdx.x, nthreads = gridDim.x * TPB;
    const int n_todo = a.in_list ? min(*a.in_count, a.n_jobs) : a.n_jobs;
    for (int k0 = tid; k0 < n_todo; k0 += nthreads) {
        const int i = a.in_list ? a.in_list[k0] : k0;
        if (!a.in_list) {
            if (a.skip_dead && a.out[i].status == 4) continue;
            a.out[i].status = 0;
        }
        const TraceJob J = a.jobs[i];
        if (J.score <= 0) continue;
        const TaskDef td = c_tasks[J.tdef & 0xFF];
        const int shift = J.tdef >> 8;
        const int ws = J.ws, rb = J.rb, qb = J.qb;
        const int refLen = J.re - rb + 1, readLen = J.qe - qb + 1, score = J.score;
        const uint8_t* gc = a.codes + J.seg_start;
        auto gidx = [&](int q) -> int { return td.reversed ? (J.seg_len - 1 - q) : q; };     // seq2 index -> segment index
        const int ntmax = refLen + readLen;

        // Gapless shortcut (most alignments on the bench workload).  If the box is square and its main diagonal scores exactly
        // `score` with every prefix sum positive, banded_sw reproduces that diagonal: no cell of the box can exceed `score`
        // (the forward pass found it as the window maximum), so no path reaches a diagonal cell with more than its prefix
        // sum, every diagonal cell holds exactly its prefix sum, and the direction rule prefers the diagonal on ties
        // (sswNew.cpp:1148).  The band |dlen|+1 = 1 contains the diagonal, so the first pass already reaches `score`.
        bool gapless = false;
        if (refLen == readLen && score < kQ4Guard) {      // (literal-scored alignments may sit below the exact maximum: no proof)
            int sum = 0;
            bool pos = true;
            for (int k = 0; k < refLen; ++k) {
                const int rf = td.img[gc[gidx(ws + rb + k)]], rc = a.rna_ssw[qb + k];
                sum += (rf == rc && rf < 4) ? kMatch : kMismatch;
                pos &= sum > 0;
            }
            gapless = pos && sum == score;
        }
        // region: three int16 score rows | translated base codes of the window | direction nibbles, (2*bw+1) per row,
        // rows padded to whole bytes | ops, 2 bits each
        int bw = abs(refLen - readLen) + 1, maxv = 0, width_d = 0, lineb = 0;
        unsigned char* dir = nullptr;
        unsigned char* ops = nullptr;
        bool fits = true;
        for (; !gapless;) {
            const int width = bw * 2 + 3;
            width_d = bw * 2 + 1;
            lineb = (width_d + 1) >> 1;
            const int rowb = ((width + 2) * 2 + 3) & ~3;
            const int o_ref = 3 * rowb;
            const int o_dir = o_ref + ((refLen + 3) & ~3);
            const long long o_ops = o_dir + (((long long)lineb * readLen + 3) & ~3LL);
            if (o_ops + ((ntmax + 4 + 3) >> 2) > BYTES) { fits = false; break; }
            int16_t* h_b = reinterpret_cast<int16_t*>(reg);
            int16_t* e_b = reinterpret_cast<int16_t*>(reg + rowb);
            int16_t* h_c = reinterpret_cast<int16_t*>(reg + 2 * rowb);
            unsigned char* refc = reg + o_ref;
            dir = reg + o_dir; ops = reg + o_ops;
            for (int j = 0; j < width + 2; ++j) { h_b[j] = 0; e_b[j] = 0; h_c[j] = 0; }
            for (int j = 0; j < refLen; ++j) refc[j] = (unsigned char)td.img[gc[gidx(ws + rb + j)]];
            for (int ii = 0; ii < readLen; ++ii) {
                const int beg = max(0, ii - bw), end = min(refLen - 1, ii + bw);
                const int edge = min(end + 1, width - 1);
                int f = 0, u = 0;
                h_b[0] = 0; e_b[0] = 0; h_b[edge] = 0; e_b[edge] = 0; h_c[0] = 0;
                unsigned char* line = dir + lineb * ii;
                const int rc = a.rna_ssw[qb + ii];
                unsigned acc = 0;
                for (int j = beg; j <= end; ++j) {
                    u = band_u(bw, ii, j);
                    const int e = band_u(bw, ii - 1, j), dd = band_u(bw, ii - 1, j - 1);
                    int t1 = (ii == 0) ? -kGapOpen : h_b[e] - kGapOpen;
                    int t2 = (ii == 0) ? -kGapExt : e_b[e] - kGapExt;
                    const int ev = t1 > t2 ? t1 : t2;
                    const unsigned e_open = t1 > t2;                      // direction code 3 (else 2)
                    e_b[u] = (int16_t)ev;
                    t1 = h_c[u - 1] - kGapOpen;
                    t2 = f - kGapExt;
                    f = t1 > t2 ? t1 : t2;
                    const unsigned f_open = t1 > t2;                      // direction code 5 (else 4)
                    const int e1 = ev > 0 ? ev : 0, f1 = f > 0 ? f : 0;
                    t1 = e1 > f1 ? e1 : f1;
                    const int rf = refc[j];
                    t2 = h_b[dd] + ((rf == rc && rf < 4) ? kMatch : kMismatch);
                    const int hv = t1 > t2 ? t1 : t2;
                    h_c[u] = (int16_t)hv;
                    if (hv > maxv) maxv = hv;
                    const unsigned hsel = (t1 <= t2) ? 0u : (e1 > f1 ? 1u : 2u);      // H came from: diagonal / E / F
                    const unsigned nib = e_open | (f_open << 1) | (hsel << 2);
                    const int c = j - beg;
                    if (c & 1) line[c >> 1] = (unsigned char)(acc | (nib << 4)); else acc = nib;
                }
                if (end >= beg && !((end - beg) & 1)) line[(end - beg) >> 1] = (unsigned char)acc;
                for (int j = 1; j <= u; ++j) h_b[j] = h_c[j];
            }
            if (maxv >= score) break;
            bw *= 2;
        }
        // traceback (sswNew.cpp:1159-1238): ops come out end -> start; written backwards into `ops`
        int wp = ntmax + 2;
        auto put_op = [&](int k, unsigned op) { const unsigned sh = (k & 3) * 2; ops[k >> 2] = (unsigned char)((ops[k >> 2] & ~(3u << sh)) | (op << sh)); };
        if (gapless) wp = ntmax + 2 - (readLen - 1);       // readLen - 1 diagonal steps, all 'M' (ops are not materialised)
        else if (fits) {
            int ii = readLen - 1, j = refLen - 1, plane = 2;
            while (ii > 0) {
                const int c = j - max(0, ii - bw);
                if (j < 0 || c < 0 || c >= width_d || wp <= 2) { fits = false; break; }     // off the band: the generic tier decides
                const unsigned nib = (dir[lineb * ii + (c >> 1)] >> ((c & 1) * 4)) & 15u;
                const unsigned hsel = nib >> 2;
                const int dv = plane == 0 ? 2 + (int)(nib & 1) : (plane == 1 ? 4 + (int)((nib >> 1) & 1)
                               : (hsel == 0 ? 1 : (hsel == 1 ? 2 + (int)(nib & 1) : 4 + (int)((nib >> 1) & 1))));
                --wp;
                if (dv == 1) { --ii; --j; plane = 2; put_op(wp, 0); }
                else if (dv == 2) { --ii; plane = 0; put_op(wp, 1); }
                else if (dv == 3) { --ii; plane = 2; put_op(wp, 1); }
                else if (dv == 4) { --j; plane = 1; put_op(wp, 2); }
                else { --j; plane = 2; put_op(wp, 2); }
            }
        }
        if (!fits) { a.out[i].status = 2; a.out_list[atomicAdd(a.out_count, 1)] = i; continue; }
        if (gapless) --wp; else put_op(--wp, 0);            // closing rule (:1220-1238): the alignment always starts with one more M column
        const int nt = ntmax + 2 - wp;
        // expansion from the front exactly like getAlignment (q walks the translated DNA from ref_begin, p the RNA), fused
        // with the identity count and the stability sum of convertMyTriplex (float32, same operation order)
        const bool with_tri = nt >= a.nt_min && nt <= a.nt_max;
        const float pt = __int2float_rn(a.penalty_t), pc = __int2float_rn(a.penalty_c);
        char* o_tfo = a.strpool ? a.strpool + J.out_off : nullptr;
        char* o_tts = o_tfo ? o_tfo + nt + 1 : nullptr;
        int q = ws + rb - shift, p = qb, match = 0;
        float tri = 0.0f, prev_val = 0.0f;
        char prev_ch = 0;
        for (int k = 0; k < nt; ++k) {
            const int op = gapless ? 0 : (ops[(wp + k) >> 2] >> (((wp + k) & 3) * 2)) & 3;
            char rch = '-', sch = '-', tch = '-';
            if (op != 2) rch = (char)a.rna_raw[p++];
            if (op != 1) {
                if (q >= 0 && q < J.seg_len) {
                    const int gi = gidx(q);
                    const unsigned char raw = a.dna[J.seg_start + gi];
                    sch = td.comp_src ? comp_char(raw) : (char)raw;
                    const int dcode = td.img[gc[gi]];
                    tch = dcode < 4 ? "ACGT"[dcode] : 'N';
                } else { sch = 'N'; tch = 'N'; }
                ++q;
            }
            if (tch == rch) ++match;
            if (with_tri) {
                float val = stability_dev(sch, rch, td.para);
                if (sch == prev_ch && sch == 'T') { tri = __fadd_rn(__fsub_rn(tri, prev_val), pt); val = pt; }
                if (sch == prev_ch && sch == 'C') { tri = __fadd_rn(__fsub_rn(tri, prev_val), pc); val = pc; }
                prev_val = val;
                if (sch != '-') prev_ch = sch;
                tri = __fadd_rn(tri, val);
            }
            if (o_tfo) { o_tfo[k] = rch; o_tts[k] = sch; }
        }
        if (o_tfo) { o_tfo[nt] = 0; o_tts[nt] = 0; }
        if (with_tri) tri = __fdiv_rn(tri, __int2float_rn(nt));
        TraceOut o1;
        o1.status = 1; o1.nt = nt; o1.identity = __fdiv_rn(__int2float_rn(100 * match), __int2float_rn(nt)); o1.tri = tri;
        a.out[i] = o1;
    }
}

// ---------------------------------------------------------------------------------------------
// Candidate compaction.  A row reaches the output only if it passes the per-task filter (fastsim.h:284-288) and the
// record filter (Fasim-LongTarget.cpp:589-597) itself; the sort / unique rounds in between only remove rows.  So only
// the tasks that own at least one such alignment need their records on the host (all of them: the others still take
// part in that task's de-duplication).  One warp per task.
struct CompactArgs {
    const TraceJob* jobs; const TraceOut* tout;
    const int* task_off; int n_tasks; int n_peaks;
    int need_nt; float min_id, min_st;
    int* cand_cnt;      // [task] number of records to ship (0: none)
    int* cand_flag;     // [task] 0 / 1
    const int* cand_poff; const int* cand_toff;        // exclusive scans of the two
    TraceJob* c_jobs; TraceOut* c_tout;                // compacted records
    int* c_task; int* c_poff;                          // per shipped task: task index, first compacted record
    int* n_unfinished;  // alignments no traceback tier could finish (status 2) — reported as an error by the host
    unsigned long long* filt;   // statistics: [0] live alignments, [1] nt upper bound < need_nt, [2] nt ok, [3] identity ok, [4] stability ok, [5] all
};

__global__ void k_task_flag(const CompactArgs a)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= a.n_tasks) return;
    const int b = a.task_off[t], e = (t + 1 < a.n_tasks) ? a.task_off[t + 1] : a.n_peaks;
    bool any = false, bad = false;
    for (int i = b + lane; i < e; i += 32) {
        const TraceOut o = a.tout[i];
        const bool live = a.jobs[i].score > 0;
        bad |= live && o.status == 2;
        const bool ok = live && o.status == 1;
        any |= ok && o.nt >= a.need_nt && o.identity >= a.min_id && o.tri >= a.min_st;
        if (a.filt) {
            const TraceJob J = a.jobs[i];
            const int ub = (J.re - J.rb + 1) + (J.qe - J.qb + 1) - (J.score + 4) / 5;
            const unsigned m0 = __ballot_sync(__activemask(), live), m1 = __ballot_sync(__activemask(), live && ub < a.need_nt);
            const unsigned m2 = __ballot_sync(__activemask(), ok && o.nt >= a.need_nt), m3 = __ballot_sync(__activemask(), ok && o.identity >= a.min_id);
            const unsigned m4 = __ballot_sync(__activemask(), ok && o.tri >= a.min_st);
            const unsigned m5 = __ballot_sync(__activemask(), ok && o.nt >= a.need_nt && o.identity >= a.min_id && o.tri >= a.min_st);
            if ((__activemask() & ((1u << lane) - 1)) == 0) {
                atomicAdd(&a.filt[0], (unsigned long long)__popc(m0)); atomicAdd(&a.filt[1], (unsigned long long)__popc(m1));
                atomicAdd(&a.filt[2], (unsigned long long)__popc(m2)); atomicAdd(&a.filt[3], (unsigned long long)__popc(m3));
                atomicAdd(&a.filt[4], (unsigned long long)__popc(m4)); atomicAdd(&a.filt[5], (unsigned long long)__popc(m5));
            }
        }
    }
    any = __any_sync(0xffffffffu, any);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        a.cand_cnt[t] = any ? e - b : 0;
        a.cand_flag[t] = any ? 1 : 0;
        if (bad) atomicAdd(a.n_unfinished, 1);
    }
}

__global__ void k_task_compact(const CompactArgs a)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= a.n_tasks || !a.cand_flag[t]) return;
    const int b = a.task_off[t], n = a.cand_cnt[t], dst = a.cand_poff[t];
    if (lane == 0) { const int pos = a.cand_toff[t]; a.c_task[pos] = t; a.c_poff[pos] = dst; }
    for (int i = lane; i < n; i += 32) { a.c_jobs[dst + i] = a.jobs[b + i]; a.c_tout[dst + i] = a.tout[b + i]; }
}

}  // namespace ltg
