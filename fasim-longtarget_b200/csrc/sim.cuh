// -F mode on the device: SIM() — sim.h:410-1143 — one warp per (segment, rule, strand, orientation) task.
//
// Phase A, the first pass over the whole lncRNA x segment matrix (sim.h:498-553).  The reference walks the matrix row by row
// and, for every cell whose x10 score exceeds the threshold, updates a 50-entry node list keyed by the alignment's start
// (addnode).  The list is order dependent (a full list evicts its lowest node), so the updates must be replayed in row-major
// order; the cell values are not.  The warp therefore sweeps strips of 32 rows as an anti-diagonal wavefront — lane l owns row
// strip + l, the cell's (C, D) candidates travel to the lane below by shuffle, the strip's last row round-trips through an
// L2-resident buffer — and every lane appends the qualifying cells of its row to a per-row event buffer.  After each strip
// the rows' events are replayed in order by the whole warp: the node list lives in registers (nodes k and k + 32 in lane k),
// a lookup is two compares and a ballot, the "first lowest node" that an insertion into a full list evicts is cached and only
// re-derived (one REDUX) when an update could have changed it.
// A candidate (score, start row, start column) is one 64-bit word (sim_core.cuh): the reference's ORDER macro is a max.
//
// Phase B, the k-best loop (sim.h:554-1142): the best node leaves the list, lane 0 aligns it with the shared Myers-Miller core
// (sim_core.cuh diff(), small: one box of the alignment's size), then the scores the new alignment may have changed are
// recomputed (:853-1140) — hundreds of millions of cells per task on repeat-rich input, so these sweeps are warp-parallel too.
// Each of them walks one LINE (a row of the rectangle, or a column while it grows to the left) whose only sequential part is
// the gap state f carried along the line, f_k = max(f_{k-1} - R, n_{k-1} - Q - R).  That recurrence is a running maximum of
// n_t + t R, so a line is computed 32 positions at a time with a warp prefix maximum on the packed candidates (ties and all:
// adding the same score offset to both sides of a comparison does not change its outcome).  The node-list updates of the
// forward recomputation are replayed in line order like in phase A.  The scalar form of the same loop (sim_core.cuh
// best_alignments) is what the host-side unit tests pin against the reference; tests/test_gpu_parity.py pins this one.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "sim_core.cuh"

namespace ltg {

struct SimHeader { int n_aln, aln_off, error, numnode_first; };     // per task: alignments in the pool at aln_off (int units)

struct SimArgs {
    const int* task_ids; int n_tasks;          // batch task ids (seg * T + task index)
    int tasks_per_seg;
    const uint8_t* codes; const SegDesc* segs;
    const uint8_t* rna_codes; int m;           // lncRNA as A0 C1 G2 T3, anything else 4
    const int* task_thr;                       // [seg * T + task] min_score = (int)(calc_score_once * 0.8)
    unsigned char* scratch; long long scratch_per_warp;
    int max_len;                               // longest segment of the batch
    int* counter;                              // work queue head
    SimHeader* hdr;                            // [n_tasks] in task_ids order
    int* pool; int pool_cap; int* pool_used;   // alignments (8 ints each) followed by their scripts
};

// per-warp scratch for lncRNA length m and segment length n: the regions of k_sim in order, each 16-byte aligned
constexpr int kSimUsedCap = 32768;
__host__ __device__ inline int sim_script_cap(int m, int n) { const int v = 4 * (m + n + 4); return v > 65536 ? v : 65536; }
struct SimLayout {
    long long b, CC, DD, BC, BD, HH, WW, c1, d1, c2, d2, used_head, used_col, used_next, list, ev_k, ev_j, script, alns, total;
    int ev_pitch, script_cap;
};
__host__ __device__ inline SimLayout sim_layout(int m, int n)
{
    SimLayout L;
    const long long N2 = n + 2, M2 = m + 2;
    long long at = 0;
    auto take = [&](long long bytes) { const long long o = at; at += (bytes + 15) & ~15LL; return o; };
    L.b = take(N2);
    L.CC = take(8 * N2); L.DD = take(8 * N2); L.BC = take(8 * N2); L.BD = take(8 * N2);
    L.HH = take(8 * M2); L.WW = take(8 * M2);
    L.c1 = take(4 * N2); L.d1 = take(4 * N2); L.c2 = take(4 * N2); L.d2 = take(4 * N2);
    L.used_head = take(4 * M2); L.used_col = take(4LL * kSimUsedCap); L.used_next = take(4LL * kSimUsedCap);
    L.list = take(32LL * simk::kNodes);
    L.ev_pitch = (n + 7) & ~7;
    L.ev_k = take(32LL * L.ev_pitch * 8); L.ev_j = take(32LL * L.ev_pitch * 2);
    L.script_cap = sim_script_cap(m, n);
    L.script = take(4LL * L.script_cap);
    L.alns = take(32LL * (simk::kNodes + 1));
    L.total = (at + 255) & ~255LL;
    return L;
}
__host__ __device__ inline long long sim_scratch_bytes(int m, int n) { return sim_layout(m, n).total; }

// ---- the 50-node list, spread over the warp's registers: node k in lane k (slot 0), node k + 32 in lane k (slot 1) --------
struct WarpList {
    int score[2], start[2], endi[2], endj[2], top[2], bot[2], left[2], right[2];
    int numnode, low, lowscore;
    bool dirty;                      // the cached "first lowest node" (low, lowscore) must be re-derived before the next eviction
};

__device__ __forceinline__ void wl_init(WarpList& L)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) { L.score[h] = 0; L.start[h] = -1; L.endi[h] = L.endj[h] = L.top[h] = L.bot[h] = L.left[h] = L.right[h] = 0; }
    L.numnode = 0; L.low = 0; L.lowscore = 0; L.dirty = true;
}

// addnode (sim.h:99-148), executed by the whole warp with uniform arguments
__device__ __forceinline__ void wl_add(WarpList& L, int lane, int cs, int st, int ev_i, int ej)
{
    using namespace simk;
    const bool m0 = lane < L.numnode && L.start[0] == st;
    const bool m1 = lane + 32 < L.numnode && L.start[1] == st;
    const unsigned b0 = __ballot_sync(0xffffffffu, m0), b1 = __ballot_sync(0xffffffffu, m1);
    if (b0 | b1) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h ? m1 : m0) {
                if (L.score[h] < cs) { L.score[h] = cs; L.endi[h] = ev_i; L.endj[h] = ej; }
                if (L.top[h] > ev_i) L.top[h] = ev_i;
                if (L.bot[h] < ev_i) L.bot[h] = ev_i;
                if (L.left[h] > ej) L.left[h] = ej;
                if (L.right[h] < ej) L.right[h] = ej;
            }
        }
        const int idx = b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;
        if (idx == L.low && cs > L.lowscore) L.dirty = true;        // the cached lowest node just rose
        return;
    }
    int idx;
    if (L.numnode < kNodes) { idx = L.numnode++; L.dirty = true; }
    else {
        if (L.dirty) {
            // first lowest node: smallest score, then smallest index
            unsigned key = 0xffffffffu;
            if (lane < L.numnode) key = ((unsigned)L.score[0] << 6) | (unsigned)lane;
            if (lane + 32 < L.numnode) key = min(key, ((unsigned)L.score[1] << 6) | (unsigned)(lane + 32));
            key = __reduce_min_sync(0xffffffffu, key);
            L.low = (int)(key & 63u); L.lowscore = (int)(key >> 6);
            L.dirty = false;
        }
        idx = L.low;
        // the new node takes the slot of the lowest one; with a score not above it the slot stays the first lowest
        if (cs <= L.lowscore) L.lowscore = cs; else L.dirty = true;
    }
    if ((idx & 31) == lane) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (h == (idx >> 5)) {
                L.score[h] = cs; L.start[h] = st; L.endi[h] = ev_i; L.endj[h] = ej;
                L.top[h] = L.bot[h] = ev_i; L.left[h] = L.right[h] = ej;
            }
        }
    }
}

__device__ __forceinline__ simk::Node wl_get(const WarpList& L, int idx)
{
    const int h = idx >> 5, src = idx & 31;
    simk::Node n;
    n.score = __shfl_sync(0xffffffffu, h ? L.score[1] : L.score[0], src);
    n.start = __shfl_sync(0xffffffffu, h ? L.start[1] : L.start[0], src);
    n.endi = __shfl_sync(0xffffffffu, h ? L.endi[1] : L.endi[0], src);
    n.endj = __shfl_sync(0xffffffffu, h ? L.endj[1] : L.endj[0], src);
    n.top = __shfl_sync(0xffffffffu, h ? L.top[1] : L.top[0], src);
    n.bot = __shfl_sync(0xffffffffu, h ? L.bot[1] : L.bot[0], src);
    n.left = __shfl_sync(0xffffffffu, h ? L.left[1] : L.left[0], src);
    n.right = __shfl_sync(0xffffffffu, h ? L.right[1] : L.right[0], src);
    return n;
}

__device__ __forceinline__ void wl_set(WarpList& L, int lane, int idx, const simk::Node& n)
{
    if ((idx & 31) != lane) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == (idx >> 5)) {
            L.score[h] = n.score; L.start[h] = n.start; L.endi[h] = n.endi; L.endj[h] = n.endj;
            L.top[h] = n.top; L.bot[h] = n.bot; L.left[h] = n.left; L.right[h] = n.right;
        }
    }
}

// index of the node with the highest score, the first of them (sim.h:557-559)
__device__ __forceinline__ int wl_argmax(const WarpList& L, int lane)
{
    unsigned key = 0;
    if (lane < L.numnode) key = ((unsigned)L.score[0] << 6) | (unsigned)(63 - lane);
    if (lane + 32 < L.numnode) key = max(key, ((unsigned)L.score[1] << 6) | (unsigned)(63 - (lane + 32)));
    key = __reduce_max_sync(0xffffffffu, key);
    return 63 - (int)(key & 63u);
}

// no_cross — sim.h:150-169
__device__ __forceinline__ bool wl_no_cross(const WarpList& L, int lane, int m1, int mm, int n1, int nn, int& rl, int& cl)
{
    using namespace simk;
    bool hit[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int si = start_row(L.start[h]), sj = start_col(L.start[h]);
        hit[h] = lane + 32 * h < L.numnode && si <= mm && sj <= nn && L.bot[h] >= m1 - 1 && L.right[h] >= n1 - 1 && (si < rl || sj < cl);
    }
    const unsigned b0 = __ballot_sync(0xffffffffu, hit[0]), b1 = __ballot_sync(0xffffffffu, hit[1]);
    if (!(b0 | b1)) return true;
    const int idx = b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;
    const int st = __shfl_sync(0xffffffffu, (idx >> 5) ? L.start[1] : L.start[0], idx & 31);
    if (start_row(st) < rl) rl = start_row(st);
    if (start_col(st) < cl) cl = start_col(st);
    return false;
}

// ---- one line of a phase-B sweep, 32 positions at a time (see the header) ----------------------------------------------------
// Position k of the line is the cell (fix, first + dir * k) when ROW (a row of the rectangle, `fix` = its row) or
// (first + dir * k, fix) otherwise (a column, `fix` = its column); LC / LD are indexed by the varying coordinate.
struct LineOut { simk::cand_t c_last, f_last; bool any_pos, any_hit, last_hit; };

template <bool ROW, bool EVENTS>
__device__ __forceinline__ LineOut sweep_line(int lane, simk::cand_t* LC, simk::cand_t* LD, int fix, int first, int dir, int W,
                                              simk::cand_t c0, simk::cand_t f0, simk::cand_t p0, const uint8_t* a, const uint8_t* b,
                                              const int* used_head, const int* used_col, const int* used_next, int rl, int cl,
                                              int floor_min, WarpList& L)
{
    using namespace simk;
    const cand_t kNone = pack(-(1 << 30), 0, 0);
    LineOut o;
    o.c_last = c0; o.f_last = f0; o.any_pos = false; o.any_hit = false; o.last_hit = false;
    cand_t carry = minus(better(minus(f0, kR), minus(c0, kQ + kR)), -kQ);       // f at position 0, lifted by Q (see f_k below)
    cand_t carry_p = p0;
    const int fix_code = ROW ? a[fix - 1] : b[fix - 1];
    const int row_head = ROW ? used_head[fix] : -1;
    for (int k0 = 0; k0 < W; k0 += 32) {
        const int k = k0 + lane;
        const bool ok = k < W;
        const int v = first + dir * k;                              // the varying coordinate
        const int ci = ROW ? fix : v, cj = ROW ? v : fix;
        cand_t lc = kNone, ld = kNone;
        if (ok) { lc = LC[v]; ld = LD[v]; }
        // diagonal predecessor: the old C of the previous position
        cand_t p = __shfl_up_sync(0xffffffffu, lc, 1);
        if (lane == 0) p = carry_p;
        carry_p = __shfl_sync(0xffffffffu, lc, 31);
        cand_t d = kNone, n0 = kNone, g = kNone;
        if (ok) {
            d = better(minus(ld, kR), minus(lc, kQ + kR));
            bool blocked = false;
            for (int e = ROW ? row_head : used_head[v]; e >= 0; e = used_next[e]) if (used_col[e] == cj) { blocked = true; break; }
            const int sub = subst(fix_code, ROW ? b[v - 1] : a[v - 1]);
            int val = 0;
            if (!blocked) val = score_of(p) + sub;
            n0 = val <= 0 ? pack(0, ci, cj) : (cand_t)(((unsigned long long)(unsigned)val << 32) | (unsigned)start_of(p));
            n0 = better(n0, d);
            g = minus(n0, -kR * k);
        }
        // exclusive running maximum of g over the positions before k (carry = everything before this tile, f_0 included)
        cand_t incl = g;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const cand_t y = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl = better(incl, y);
        }
        cand_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
        excl = lane == 0 ? carry : better(excl, carry);
        carry = better(carry, __shfl_sync(0xffffffffu, incl, 31));
        const cand_t f = minus(excl, kQ + kR * k);                  // f_k = max(f_0 - k R, max_{t<k} (n_t + t R) - Q - k R)
        const cand_t n = better(n0, f);
        if (ok) { LC[v] = n; LD[v] = d; }
        const bool pos = ok && score_of(n) > floor_min;
        auto inside = [&](cand_t x) { const int st = start_of(x); return start_row(st) > rl && start_col(st) > cl; };
        const bool hit = ok && (inside(n) || inside(d) || inside(f));
        const unsigned mpos = __ballot_sync(0xffffffffu, pos), mhit = __ballot_sync(0xffffffffu, hit);
        o.any_pos |= mpos != 0; o.any_hit |= mhit != 0;
        if (k0 + 32 >= W) {
            const int lastl = W - 1 - k0;
            o.last_hit = (mhit >> lastl) & 1u;
            o.c_last = __shfl_sync(0xffffffffu, n, lastl);
            o.f_last = __shfl_sync(0xffffffffu, f, lastl);
        }
        if (EVENTS) {
            // node-list updates of this tile, in line order (sim.h:1091-1092)
            unsigned m = mpos;
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const cand_t ek = __shfl_sync(0xffffffffu, n, src);
                wl_add(L, lane, score_of(ek), start_of(ek), fix, first + dir * (k0 + src));
            }
        }
    }
    return o;
}

__global__ void __launch_bounds__(128) k_sim(const SimArgs a)
{
    using namespace simk;
    __shared__ cand_t s_ring[4][2][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * (blockDim.x >> 5) + wib;
    unsigned char* base = a.scratch + (size_t)warp * a.scratch_per_warp;
    const int M = a.m;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= a.n_tasks) break;
        const int task = a.task_ids[t];
        const SegDesc sd = a.segs[task / a.tasks_per_seg];
        const TaskDef td = c_tasks[task % a.tasks_per_seg];
        const int N = sd.len;
        const int min_score = a.task_thr[task];
        // carve the scratch (the stride between warps is sized for the batch's longest segment)
        const SimLayout Lay = sim_layout(M, N);
        uint8_t* b = base + Lay.b;
        cand_t* CC = reinterpret_cast<cand_t*>(base + Lay.CC);
        cand_t* DD = reinterpret_cast<cand_t*>(base + Lay.DD);
        cand_t* BC = reinterpret_cast<cand_t*>(base + Lay.BC);           // strip boundary: C / D of the strip's last row, per column
        cand_t* BD = reinterpret_cast<cand_t*>(base + Lay.BD);
        cand_t* HH = reinterpret_cast<cand_t*>(base + Lay.HH);
        cand_t* WW = reinterpret_cast<cand_t*>(base + Lay.WW);
        int* c1 = reinterpret_cast<int*>(base + Lay.c1);
        int* d1 = reinterpret_cast<int*>(base + Lay.d1);
        int* c2 = reinterpret_cast<int*>(base + Lay.c2);
        int* d2 = reinterpret_cast<int*>(base + Lay.d2);
        int* used_head = reinterpret_cast<int*>(base + Lay.used_head);
        int* used_col = reinterpret_cast<int*>(base + Lay.used_col);
        int* used_next = reinterpret_cast<int*>(base + Lay.used_next);
        cand_t* ev_k = reinterpret_cast<cand_t*>(base + Lay.ev_k);
        const int ev_pitch = Lay.ev_pitch;
        uint16_t* ev_j = reinterpret_cast<uint16_t*>(base + Lay.ev_j);
        int* script = reinterpret_cast<int*>(base + Lay.script);
        const int script_cap = Lay.script_cap;
        Aln* alns = reinterpret_cast<Aln*>(base + Lay.alns);
        const long long M2 = M + 2;

        __syncwarp();
        for (int q = lane; q < N; q += 32) b[q] = (uint8_t)td.img[a.codes[sd.start + (td.reversed ? N - 1 - q : q)]];
        for (int i = lane; i < M2; i += 32) used_head[i] = -1;
        __syncwarp();

        // ---------------- phase A: wavefront first pass + in-order replay of the node-list updates
        WarpList L;
        wl_init(L);
        cand_t* my_evk = ev_k + (size_t)lane * ev_pitch;
        uint16_t* my_evj = ev_j + (size_t)lane * ev_pitch;

        for (int strip0 = 0; strip0 < M; strip0 += 32) {
            const int i = strip0 + lane + 1;                 // this lane's row (1-based)
            const bool row_ok = i <= M;
            const int ai = row_ok ? a.rna_codes[i - 1] : 4;
            cand_t c = pack(0, i, 0), f = pack(-kQ, i, 0), diag = pack(0, i - 1, 0);
            cand_t c_out = 0, d_out = 0;
            int evcnt = 0;
            const bool first = strip0 == 0, last = strip0 + 32 >= M;
            const int steps = N + 31;
            for (int s0 = 0; s0 < steps; s0 += 32) {
                if (!first) {
                    // the row above this strip, columns s0 + 1 .. s0 + 32 (read 31 steps before lane 31 overwrites them)
                    __syncwarp();
                    const int j = s0 + lane + 1;
                    s_ring[wib][0][lane] = j <= N ? BC[j] : 0;
                    s_ring[wib][1][lane] = j <= N ? BD[j] : 0;
                    __syncwarp();
                }
                const int cnt = min(32, steps - s0);
                for (int k = 0; k < cnt; ++k) {
                    const int s = s0 + k;
                    const int j = s - lane + 1;              // this lane's column (1-based) at this step
                    cand_t up_c = __shfl_up_sync(0xffffffffu, c_out, 1);
                    cand_t up_d = __shfl_up_sync(0xffffffffu, d_out, 1);
                    if (lane == 0) {
                        if (first) { up_c = pack(0, 0, j); up_d = pack(-kQ, 0, j); }
                        else { up_c = s_ring[wib][0][k]; up_d = s_ring[wib][1][k]; }
                    }
                    if (row_ok && j >= 1 && j <= N) {
                        cand_t lc = up_c, ld = up_d, pp = diag;
                        cell(c, f, lc, ld, pp, subst(ai, b[j - 1]), false, i, j);
                        diag = up_c;
                        c_out = lc; d_out = ld;
                        if (score_of(lc) > min_score) { my_evk[evcnt] = lc; my_evj[evcnt] = (uint16_t)j; ++evcnt; }
                        if (lane == 31 && !last) { BC[j] = lc; BD[j] = ld; }
                    }
                }
            }
            __syncwarp();
            // ---- replay: rows of the strip in order, each row's events in column order
            for (int r = 0; r < 32; ++r) {
                const int cnt_r = __shfl_sync(0xffffffffu, evcnt, r);
                const int ev_i = strip0 + r + 1;
                const cand_t* rk = ev_k + (size_t)r * ev_pitch;
                const uint16_t* rj = ev_j + (size_t)r * ev_pitch;
                for (int e0 = 0; e0 < cnt_r; e0 += 32) {
                    cand_t myk = 0;
                    int myj = 0;
                    if (e0 + lane < cnt_r) { myk = rk[e0 + lane]; myj = rj[e0 + lane]; }
                    const int nb = min(32, cnt_r - e0);
                    for (int k = 0; k < nb; ++k) {
                        const cand_t ek = __shfl_sync(0xffffffffu, myk, k);
                        const int ej = __shfl_sync(0xffffffffu, myj, k);
                        wl_add(L, lane, score_of(ek), start_of(ek), ev_i, ej);
                    }
                }
            }
            __syncwarp();
        }
        const int numnode_first = L.numnode;

        // ---------------- phase B: the k best alignments (sim.h:554-1142)
        int n_out = 0, script_n = 0, err = 0, floor_min = 0, used_n = 0;
        for (int count = L.numnode - 1; count >= 0 && !err; --count) {
            const int best = wl_argmax(L, lane);
            const Node cur = wl_get(L, best);
            --L.numnode;
            if (best != L.numnode) { const Node lastn = wl_get(L, L.numnode); wl_set(L, lane, best, lastn); }
            L.dirty = true;
            const int score = cur.score;
            const int stari = start_row(cur.start) + 1, starj = start_col(cur.start) + 1, endi = cur.endi, endj = cur.endj;
            int m1 = cur.top, mm = cur.bot, n1 = cur.left, nn = cur.right;
            int rl = endi - stari + 1, cl = endj - starj + 1;
            // the alignment itself: Myers-Miller on lane 0 (marks its cells as used)
            int stop = 0;
            if (lane == 0) {
                Task T;
                T.a = a.rna_codes; T.b = b; T.M = M; T.N = N; T.min_score = min_score;
                T.CC = CC; T.DD = DD; T.HH = HH; T.WW = WW; T.c1 = c1; T.d1 = d1; T.c2 = c2; T.d2 = d2;
                T.used_head = used_head; T.used_col = used_col; T.used_next = used_next; T.used_cap = kSimUsedCap; T.used_n = used_n;
                T.list = nullptr; T.numnode = 0; T.floor_min = 0;
                T.I = stari - 1; T.J = starj - 1; T.last = 0;
                T.script = script; T.script_cap = script_cap; T.script_n = script_n;
                T.out = alns; T.out_cap = kNodes + 1; T.n_out = n_out; T.error = 0;
                diff(T, stari - 1, starj - 1, rl, cl, kQ, kQ);
                if (!T.error) {
                    if (score <= 10 * min_score) stop = 1;                           // score / 10.0 <= min_score (:591)
                    else if (n_out >= kNodes + 1) T.error = 3;
                    else { alns[n_out] = Aln{stari, endi, starj, endj, score, script_n, T.script_n - script_n, 0}; n_out += 1; script_n = T.script_n; }
                }
                used_n = T.used_n; err = T.error;
            }
            __syncwarp();
            stop = __shfl_sync(0xffffffffu, stop, 0);
            err = __shfl_sync(0xffffffffu, err, 0);
            n_out = __shfl_sync(0xffffffffu, n_out, 0);
            script_n = __shfl_sync(0xffffffffu, script_n, 0);
            used_n = __shfl_sync(0xffffffffu, used_n, 0);
            if (stop || err) break;
            if (!count) continue;
            // ---- scores the new alignment may have changed (:853-1140)
            bool flag = false;
            for (int j = n1 + lane; j <= nn; j += 32) { CC[j] = pack(0, mm + 1, j); DD[j] = pack(-kQ, mm + 1, j); }
            __syncwarp();
            for (int i = mm; i >= m1; --i) {
                const LineOut o = sweep_line<true, false>(lane, CC, DD, i, nn, -1, nn - n1 + 1, pack(0, i, nn + 1), pack(-kQ, i, nn + 1), pack(0, i + 1, nn + 1),
                                                          a.rna_codes, b, used_head, used_col, used_next, 0, 0, floor_min, L);
                flag |= o.any_pos;
                if (lane == 0) { HH[i] = o.c_last; WW[i] = o.f_last; }
                __syncwarp();
            }
            for (rl = m1, cl = n1;;) {
                bool rflag = true, cflag = true;
                while ((rflag && m1 > 1) || (cflag && n1 > 1)) {
                    if (rflag && m1 > 1) {                    // one more row on top
                        --m1;
                        const LineOut o = sweep_line<true, false>(lane, CC, DD, m1, nn, -1, nn - n1 + 1, pack(0, m1, nn + 1), pack(-kQ, m1, nn + 1),
                                                                  pack(0, m1 + 1, nn + 1), a.rna_codes, b, used_head, used_col, used_next, rl, cl, floor_min, L);
                        flag |= o.any_pos;
                        rflag = o.any_hit;
                        if (lane == 0) { HH[m1] = o.c_last; WW[m1] = o.f_last; }
                        if (!cflag && o.last_hit) cflag = true;
                        __syncwarp();
                    }
                    if (cflag && n1 > 1) {                    // one more column on the left
                        --n1;
                        const LineOut o = sweep_line<false, false>(lane, HH, WW, n1, mm, -1, mm - m1 + 1, pack(0, mm + 1, n1), pack(-kQ, mm + 1, n1),
                                                                   pack(0, mm + 1, n1 + 1), a.rna_codes, b, used_head, used_col, used_next, rl, cl, floor_min, L);
                        flag |= o.any_pos;
                        cflag = o.any_hit;
                        if (lane == 0) { CC[n1] = o.c_last; DD[n1] = o.f_last; }
                        if (!rflag && o.last_hit) rflag = true;
                        __syncwarp();
                    }
                }
                if ((m1 == 1 && n1 == 1) || wl_no_cross(L, lane, m1, mm, n1, nn, rl, cl)) break;
            }
            --m1; --n1;
            if (flag) {
                for (int j = n1 + 1 + lane; j <= nn; j += 32) { CC[j] = pack(0, m1, j); DD[j] = pack(-kQ, m1, j); }
                __syncwarp();
                for (int i = m1 + 1; i <= mm; ++i) {
                    const LineOut o = sweep_line<true, true>(lane, CC, DD, i, n1 + 1, +1, nn - n1, pack(0, i, n1), pack(-kQ, i, n1), pack(0, i - 1, n1),
                                                             a.rna_codes, b, used_head, used_col, used_next, 0, 0, floor_min, L);
                    if (o.any_pos) floor_min = 1;
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        // ---- results: alignment records, then the scripts, into the batch's pool
        int off = 0;
        const int need = n_out * 8 + script_n;
        if (lane == 0 && need > 0) off = atomicAdd(a.pool_used, need);
        off = __shfl_sync(0xffffffffu, off, 0);
        if (need > 0 && off + need > a.pool_cap) { err = err ? err : 5; }
        else if (need > 0) {
            const int* src_a = reinterpret_cast<const int*>(alns);
            for (int k = lane; k < n_out * 8; k += 32) a.pool[off + k] = src_a[k];
            for (int k = lane; k < script_n; k += 32) a.pool[off + n_out * 8 + k] = script[k];
        }
        if (lane == 0) a.hdr[t] = SimHeader{err ? 0 : n_out, off, err, numnode_first};
        __syncwarp();
    }
}

}  // namespace ltg
