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
// Phase B, the k-best loop (sim.h:554-1142): Myers-Miller alignment of the best node with the cells of earlier alignments
// forbidden, then the rectangle around it recomputed.  Sequential by nature and small (a few 10^4..10^5 cells per alignment):
// lane 0 runs the shared scalar core (sim_core.cuh), which the host-side unit tests pin against the reference.
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
        const SimLayout L = sim_layout(M, N);
        uint8_t* b = base + L.b;
        cand_t* CC = reinterpret_cast<cand_t*>(base + L.CC);
        cand_t* DD = reinterpret_cast<cand_t*>(base + L.DD);
        cand_t* BC = reinterpret_cast<cand_t*>(base + L.BC);           // strip boundary: C / D of the strip's last row, per column
        cand_t* BD = reinterpret_cast<cand_t*>(base + L.BD);
        cand_t* HH = reinterpret_cast<cand_t*>(base + L.HH);
        cand_t* WW = reinterpret_cast<cand_t*>(base + L.WW);
        int* c1 = reinterpret_cast<int*>(base + L.c1);
        int* d1 = reinterpret_cast<int*>(base + L.d1);
        int* c2 = reinterpret_cast<int*>(base + L.c2);
        int* d2 = reinterpret_cast<int*>(base + L.d2);
        int* used_head = reinterpret_cast<int*>(base + L.used_head);
        int* used_col = reinterpret_cast<int*>(base + L.used_col);
        int* used_next = reinterpret_cast<int*>(base + L.used_next);
        Node* list = reinterpret_cast<Node*>(base + L.list);
        cand_t* ev_k = reinterpret_cast<cand_t*>(base + L.ev_k);
        const int ev_pitch = L.ev_pitch;
        uint16_t* ev_j = reinterpret_cast<uint16_t*>(base + L.ev_j);
        int* script = reinterpret_cast<int*>(base + L.script);
        const int script_cap = L.script_cap;
        const long long M2 = M + 2;

        __syncwarp();
        for (int q = lane; q < N; q += 32) b[q] = (uint8_t)td.img[a.codes[sd.start + (td.reversed ? N - 1 - q : q)]];
        for (int i = lane; i < M2; i += 32) used_head[i] = -1;
        __syncwarp();

        // ---------------- phase A: wavefront first pass + in-order replay of the node-list updates
        // node list in registers: node k in lane k (slot 0), node k + 32 in lane k (slot 1)
        int n_score[2] = {0, 0}, n_start[2] = {-1, -1}, n_endi[2] = {0, 0}, n_endj[2] = {0, 0}, n_top[2] = {0, 0}, n_bot[2] = {0, 0},
            n_left[2] = {0, 0}, n_right[2] = {0, 0};
        int numnode = 0, low = 0, lowscore = 0;
        bool dirty = true;
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
                        const int cs = score_of(ek), st = start_of(ek);
                        // addnode (sim.h:99-148) on the register-resident list
                        const bool m0 = lane < numnode && n_start[0] == st;
                        const bool m1 = lane + 32 < numnode && n_start[1] == st;
                        const unsigned b0 = __ballot_sync(0xffffffffu, m0), b1 = __ballot_sync(0xffffffffu, m1);
                        if (b0 | b1) {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (h ? m1 : m0) {
                                    if (n_score[h] < cs) { n_score[h] = cs; n_endi[h] = ev_i; n_endj[h] = ej; }
                                    if (n_top[h] > ev_i) n_top[h] = ev_i;
                                    if (n_bot[h] < ev_i) n_bot[h] = ev_i;
                                    if (n_left[h] > ej) n_left[h] = ej;
                                    if (n_right[h] < ej) n_right[h] = ej;
                                }
                            }
                            const int idx = b0 ? __ffs(b0) - 1 : 32 + __ffs(b1) - 1;
                            if (idx == low && cs > lowscore) dirty = true;        // the cached lowest node just rose
                        } else {
                            int idx;
                            if (numnode < kNodes) { idx = numnode++; dirty = true; }
                            else {
                                if (dirty) {
                                    // first lowest node: smallest score, then smallest index
                                    unsigned key = 0xffffffffu;
                                    if (lane < numnode) key = ((unsigned)n_score[0] << 6) | (unsigned)lane;
                                    if (lane + 32 < numnode) key = min(key, ((unsigned)n_score[1] << 6) | (unsigned)(lane + 32));
                                    key = __reduce_min_sync(0xffffffffu, key);
                                    low = (int)(key & 63u); lowscore = (int)(key >> 6);
                                    dirty = false;
                                }
                                idx = low;
                                // the new node takes the slot of the lowest one; with a score not above it the slot stays the first lowest
                                if (cs <= lowscore) lowscore = cs; else dirty = true;
                            }
                            if ((idx & 31) == lane) {
                                const int h = idx >> 5;
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    if (hh == h) {
                                        n_score[hh] = cs; n_start[hh] = st; n_endi[hh] = ev_i; n_endj[hh] = ej;
                                        n_top[hh] = n_bot[hh] = ev_i; n_left[hh] = n_right[hh] = ej;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        // the list goes to memory for phase B
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int idx = lane + 32 * h;
            if (idx < numnode) list[idx] = Node{n_score[h], n_start[h], n_endi[h], n_endj[h], n_top[h], n_bot[h], n_left[h], n_right[h]};
        }
        __syncwarp();

        // ---------------- phase B: the k best alignments, scalar (lane 0)
        int n_out = 0, script_n = 0, err = 0;
        Aln* alns = reinterpret_cast<Aln*>(base + L.alns);
        if (lane == 0) {
            Task T;
            T.a = a.rna_codes; T.b = b; T.M = M; T.N = N; T.min_score = min_score;
            T.CC = CC; T.DD = DD; T.HH = HH; T.WW = WW; T.c1 = c1; T.d1 = d1; T.c2 = c2; T.d2 = d2;
            T.used_head = used_head; T.used_col = used_col; T.used_next = used_next; T.used_cap = kSimUsedCap; T.used_n = 0;
            T.list = list; T.numnode = numnode; T.floor_min = 0; T.I = T.J = T.last = 0;
            T.script = script; T.script_cap = script_cap; T.script_n = 0;
            T.out = alns; T.out_cap = kNodes + 1; T.n_out = 0; T.error = 0;
            best_alignments(T);
            n_out = T.n_out; script_n = T.script_n; err = T.error;
        }
        __syncwarp();                                  // lane 0's records and scripts are visible to the warp
        n_out = __shfl_sync(0xffffffffu, n_out, 0);
        script_n = __shfl_sync(0xffffffffu, script_n, 0);
        err = __shfl_sync(0xffffffffu, err, 0);
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
        if (lane == 0) a.hdr[t] = SimHeader{err ? 0 : n_out, off, err, numnode};
        __syncwarp();
    }
}

}  // namespace ltg
