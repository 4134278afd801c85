// -F mode: SIM() — sim.h:410-1143 — the sequential core shared by the device kernel (sim.cuh) and the host-side unit test.
//
// SIM is Huang & Miller's "k best non-intersecting local alignments in linear space" as the reference carries it: scores x10
// (match +50, mismatch -40, a gap of k columns costs 120 + 40k), a list of at most 50 candidate nodes keyed by the START of
// their alignment (addnode, sim.h:99-148), the best node aligned by the Myers-Miller divide and conquer (diff, :171-348)
// with the cells of earlier alignments forbidden, then the region the new alignment may have invalidated recomputed
// (:853-1140).  Everything here is a restatement of that control flow over plain arrays:
//   * a candidate (score, start row, start column) is ONE signed 64-bit word, score in the high half and the start packed
//     below it, so the reference's ORDER macro (:483-495: higher score, then larger start row, then larger start column) is
//     a 64-bit maximum;
//   * diff()'s recursion is an explicit stack;
//   * the `row` lists of used cells are linked lists in a pool.
// Letters outside A/C/G/T score as a mismatch (the reference reads an uninitialised matrix entry there, sim.h:468-472).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define LTG_HD __host__ __device__ __forceinline__
#else
#define LTG_HD inline
#endif

namespace ltg {
namespace simk {

constexpr int kNodes = 50;              // sim.h:17 (#define K 50)
constexpr int kColBits = 13;            // columns (segment positions) < 8192
constexpr int kMaxRows = 1 << 18;       // lncRNA rows < 262144
constexpr int kQ = 120, kR = 40;        // -10 * parm_O, -10 * parm_E (Fasim-LongTarget.cpp:424: 5, -4, -12, -4)

typedef long long cand_t;
LTG_HD cand_t pack(int score, int i, int j) { return (cand_t)(((unsigned long long)(unsigned)score << 32) | (unsigned)((i << kColBits) | j)); }
LTG_HD int score_of(cand_t c) { return (int)(c >> 32); }
LTG_HD int start_of(cand_t c) { return (int)(unsigned)c; }
LTG_HD int start_row(int st) { return st >> kColBits; }
LTG_HD int start_col(int st) { return st & ((1 << kColBits) - 1); }
LTG_HD cand_t minus(cand_t c, int d) { return c - ((cand_t)d << 32); }
LTG_HD cand_t better(cand_t a, cand_t b) { return a > b ? a : b; }
LTG_HD int subst(int a, int b) { return (a == b && a < 4) ? 50 : -40; }

struct Node { int score, start, endi, endj, top, bot, left, right; };

// one alignment as the kernel reports it: 1-based inclusive coordinates on the lncRNA (rows) and on the translated segment
// (columns), the x10 score, and its edit script (0 = aligned pair, +k = k DNA-only columns, -k = k RNA-only columns)
struct Aln { int stari, endi, starj, endj, score, script_off, script_len, pad_; };

// working set of one task
struct Task {
    const uint8_t* a;            // lncRNA codes, a[i - 1] = row i (A0 C1 G2 T3 else 4)
    const uint8_t* b;            // translated segment codes, b[j - 1] = column j
    int M, N, min_score;
    cand_t* CC; cand_t* DD;      // [N + 2] per-column C / D candidates (CC,RR,EE / DD,SS,FF of the reference)
    cand_t* HH; cand_t* WW;      // [M + 2] per-row C / F candidates (HH,II,JJ / WW,XX,YY)
    int* c1; int* d1; int* c2; int* d2;   // [N + 2] diff()'s score rows
    int* used_head;              // [M + 2] first pool entry of the row's list (-1: empty)
    int* used_col; int* used_next; int used_cap, used_n;
    Node* list; int numnode;     // LIST
    int floor_min;               // `min` of sim.h:416
    int I, J, last;              // script writer state
    int* script; int script_cap, script_n;      // output pool of this task
    Aln* out; int out_cap, n_out;
    int error;                   // 1 script pool full, 2 used-cell pool full, 3 alignment table full, 4 diff stack overflow
};

LTG_HD bool taken(const Task& T, int i, int j)
{
    for (int e = T.used_head[i]; e >= 0; e = T.used_next[e]) if (T.used_col[e] == j) return true;
    return false;
}
LTG_HD void mark(Task& T, int i, int j)
{
    if (T.used_n >= T.used_cap) { T.error = 2; return; }
    // (appended at the head: membership is all that is ever asked of a row's list)
    T.used_col[T.used_n] = j; T.used_next[T.used_n] = T.used_head[i]; T.used_head[i] = T.used_n++;
}

// addnode — sim.h:99-148
LTG_HD void addnode(Node* list, int& numnode, int c, int start, int i, int j)
{
    int at = -1;
    for (int d = 0; d < numnode; ++d) if (list[d].start == start) { at = d; break; }
    if (at >= 0) {
        Node& n = list[at];
        if (n.score < c) { n.score = c; n.endi = i; n.endj = j; }
        if (n.top > i) n.top = i;
        if (n.bot < i) n.bot = i;
        if (n.left > j) n.left = j;
        if (n.right < j) n.right = j;
        return;
    }
    if (numnode == kNodes) {
        at = 0;
        for (int d = 1; d < numnode; ++d) if (list[d].score < list[at].score) at = d;
    } else at = numnode++;
    Node& n = list[at];
    n.score = c; n.start = start; n.endi = i; n.endj = j; n.top = n.bot = i; n.left = n.right = j;
}

// One cell of a sweep (sim.h:512-547 and its siblings :866-896, :921-956, :977-1012, :1060-1093): c / f run along the sweep
// line, lc / ld are the stored candidates of the neighbouring line at this position, p the diagonal predecessor, (i, j) the
// cell itself (= the start point of an alignment that begins right after it).
LTG_HD void cell(cand_t& c, cand_t& f, cand_t& lc, cand_t& ld, cand_t& p, int sub, bool blocked, int i, int j)
{
    f = better(minus(f, kR), minus(c, kQ + kR));
    const cand_t d = better(minus(ld, kR), minus(lc, kQ + kR));
    int v = 0;
    if (!blocked) v = score_of(p) + sub;
    cand_t n = v <= 0 ? pack(0, i, j) : pack(v, start_row(start_of(p)), start_col(start_of(p)));
    n = better(better(n, d), f);
    p = lc; lc = n; ld = d; c = n;
}

// ---- script writer: DEL / INS / REP of sim.h:176-197 -------------------------------------------------------------------
LTG_HD void push_op(Task& T, int v) { if (T.script_n < T.script_cap) T.script[T.script_n++] = v; else T.error = 1; }
LTG_HD void op_del(Task& T, int k)
{
    T.I += k;
    if (T.last < 0) { T.script[T.script_n - 1] -= k; T.last = T.script[T.script_n - 1]; }
    else { push_op(T, -k); T.last = -k; }
}
LTG_HD void op_ins(Task& T, int k)
{
    T.J += k;
    if (T.last < 0) { T.script[T.script_n - 1] = k; push_op(T, T.last); }
    else { push_op(T, k); T.last = k; }
}
LTG_HD void op_rep(Task& T) { push_op(T, 0); T.last = 0; }
LTG_HD int gap(int k) { return k <= 0 ? 0 : kQ + kR * k; }

// diff — sim.h:171-348: optimal global alignment of rows oa+1..oa+m with columns ob+1..ob+n in linear space (gap-open charge tb
// at the top, te at the bottom boundary; pairs used by earlier alignments are forbidden).  The reference recurses (left part,
// then right part); here the pending parts wait on a stack, the right part below the left one, which visits the sub-problems
// in the same order — that matters, the used-cell test reads the writer's running offsets I / J.
struct DiffFrame { int oa, ob, m, n, tb, te, del2; };
constexpr int kDiffStack = 96;

LTG_HD void diff(Task& T, int oa0, int ob0, int m0, int n0, int tb0, int te0)
{
    DiffFrame st[kDiffStack];
    int sp = 0;
    st[sp++] = DiffFrame{oa0, ob0, m0, n0, tb0, te0, 0};
    int* c1 = T.c1; int* d1 = T.d1; int* c2 = T.c2; int* d2 = T.d2;
    while (sp > 0 && !T.error) {
        const DiffFrame F = st[--sp];
        if (F.del2) op_del(T, 2);                  // the two rows of a type-2 split, between its halves (:341-343)
        const int oa = F.oa, ob = F.ob, m = F.m, n = F.n;
        int tb = F.tb;
        const int te = F.te;
        if (n <= 0) { if (m > 0) op_del(T, m); continue; }
        if (m <= 1) {
            if (m <= 0) { op_ins(T, n); continue; }
            if (tb > te) tb = te;
            int midc = -(tb + kR + gap(n)), midj = 0;
            const int a1 = T.a[oa];
            for (int j = 1; j <= n; ++j) {
                if (taken(T, T.I + 1, j + T.J)) continue;
                const int c = subst(a1, T.b[ob + j - 1]) - (gap(j - 1) + gap(n - j));
                if (c > midc) { midc = c; midj = j; }
            }
            if (midj == 0) { op_ins(T, n); op_del(T, 1); }
            else {
                if (midj > 1) op_ins(T, midj - 1);
                op_rep(T);
                ++T.I; ++T.J;
                mark(T, T.I, T.J);
                if (midj < n) op_ins(T, n - midj);
            }
            continue;
        }
        const int midi = m / 2;
        // forward half: rows 1..midi
        c1[0] = 0;
        int t = -kQ;
        for (int j = 1; j <= n; ++j) { c1[j] = t = t - kR; d1[j] = t - kQ; }
        t = -tb;
        for (int i = 1; i <= midi; ++i) {
            int s = c1[0], c, e, d;
            c1[0] = c = t = t - kR;
            e = t - kQ;
            const int ai = T.a[oa + i - 1];
            for (int j = 1; j <= n; ++j) {
                if ((c = c - kQ - kR) > (e = e - kR)) e = c;
                if ((c = c1[j] - kQ - kR) > (d = d1[j] - kR)) d = c;
                if (!taken(T, i + T.I, j + T.J)) c = s + subst(ai, T.b[ob + j - 1]);
                if (c < d) c = d;
                if (c < e) c = e;
                s = c1[j]; c1[j] = c; d1[j] = d;
            }
        }
        d1[0] = c1[0];
        // reverse half: rows m-1..midi
        c2[n] = 0;
        t = -kQ;
        for (int j = n - 1; j >= 0; --j) { c2[j] = t = t - kR; d2[j] = t - kQ; }
        t = -te;
        for (int i = m - 1; i >= midi; --i) {
            int s = c2[n], c, e, d;
            c2[n] = c = t = t - kR;
            e = t - kQ;
            const int ai = T.a[oa + i];
            for (int j = n - 1; j >= 0; --j) {
                if ((c = c - kQ - kR) > (e = e - kR)) e = c;
                if ((c = c2[j] - kQ - kR) > (d = d2[j] - kR)) d = c;
                if (!taken(T, i + 1 + T.I, j + 1 + T.J)) c = s + subst(ai, T.b[ob + j]);
                if (c < d) c = d;
                if (c < e) c = e;
                s = c2[j]; c2[j] = c; d2[j] = d;
            }
        }
        d2[n] = c2[n];
        // where the halves meet (:319-332)
        int midc = c1[0] + c2[0], midj = 0, type = 1;
        for (int j = 0; j <= n; ++j) {
            const int c = c1[j] + c2[j];
            if (c >= midc && (c > midc || (c1[j] != d1[j] && c2[j] == d2[j]))) { midc = c; midj = j; }
        }
        for (int j = n; j >= 0; --j) {
            const int c = d1[j] + d2[j] + kQ;
            if (c > midc) { midc = c; midj = j; type = 2; }
        }
        if (sp + 2 > kDiffStack) { T.error = 4; return; }
        if (type == 1) {
            st[sp++] = DiffFrame{oa + midi, ob + midj, m - midi, n - midj, kQ, te, 0};
            st[sp++] = DiffFrame{oa, ob, midi, midj, tb, kQ, 0};
        } else {
            st[sp++] = DiffFrame{oa + midi + 1, ob + midj, m - midi - 1, n - midj, 0, te, 1};
            st[sp++] = DiffFrame{oa, ob, midi - 1, midj, tb, 0, 0};
        }
    }
}

// no_cross — sim.h:150-169
LTG_HD bool no_cross(const Node* list, int numnode, int m1, int mm, int n1, int nn, int& rl, int& cl)
{
    for (int k = 0; k < numnode; ++k) {
        const Node& n = list[k];
        const int si = start_row(n.start), sj = start_col(n.start);
        if (si <= mm && sj <= nn && n.bot >= m1 - 1 && n.right >= n1 - 1 && (si < rl || sj < cl)) {
            if (si < rl) rl = si;
            if (sj < cl) cl = sj;
            return false;
        }
    }
    return true;
}

// The first pass over the whole matrix, sequentially (sim.h:498-553).  The device kernel computes the same cells as a
// wavefront over 32 rows and replays the node-list updates in this row-major order (sim.cuh); this form is what the host-side
// unit test runs, and the specification of that kernel.
LTG_HD void first_pass_serial(Task& T)
{
    const int M = T.M, N = T.N;
    for (int j = 1; j <= N; ++j) { T.CC[j] = pack(0, 0, j); T.DD[j] = pack(-kQ, 0, j); }
    for (int i = 1; i <= M; ++i) {
        cand_t c = pack(0, i, 0), f = pack(-kQ, i, 0), p = pack(0, i - 1, 0);
        const int ai = T.a[i - 1];
        for (int j = 1; j <= N; ++j) {
            cell(c, f, T.CC[j], T.DD[j], p, subst(ai, T.b[j - 1]), false, i, j);
            if (score_of(c) > T.min_score) addnode(T.list, T.numnode, score_of(c), start_of(c), i, j);
        }
    }
}

// The k best alignments (sim.h:554-1142), starting from the node list the first pass left.
LTG_HD void best_alignments(Task& T)
{
    const int N = T.N;
    (void)N;
    for (int count = T.numnode - 1; count >= 0 && !T.error; --count) {
        int best = 0;
        for (int k = 1; k < T.numnode; ++k) if (T.list[k].score > T.list[best].score) best = k;
        const Node cur = T.list[best];
        --T.numnode;
        if (best != T.numnode) T.list[best] = T.list[T.numnode];
        const int score = cur.score;
        const int stari = start_row(cur.start) + 1, starj = start_col(cur.start) + 1, endi = cur.endi, endj = cur.endj;
        int m1 = cur.top, mm = cur.bot, n1 = cur.left, nn = cur.right;
        int rl = endi - stari + 1, cl = endj - starj + 1;
        T.I = stari - 1; T.J = starj - 1; T.last = 0;
        const int script_begin = T.script_n;
        diff(T, stari - 1, starj - 1, rl, cl, kQ, kQ);
        if (T.error) return;
        if (score <= 10 * T.min_score) { T.script_n = script_begin; break; }          // score / 10.0 <= min_score (:591)
        if (T.n_out >= T.out_cap) { T.error = 3; return; }
        T.out[T.n_out++] = Aln{stari, endi, starj, endj, score, script_begin, T.script_n - script_begin, 0};
        if (!count) continue;
        // ---- scores the new alignment may have changed (:853-1140)
        bool flag = false;
        for (int j = nn; j >= n1; --j) { T.CC[j] = pack(0, mm + 1, j); T.DD[j] = pack(-kQ, mm + 1, j); }
        for (int i = mm; i >= m1; --i) {
            cand_t c = pack(0, i, nn + 1), f = pack(-kQ, i, nn + 1), p = pack(0, i + 1, nn + 1);
            const int ai = T.a[i - 1];
            for (int j = nn; j >= n1; --j) {
                cell(c, f, T.CC[j], T.DD[j], p, subst(ai, T.b[j - 1]), taken(T, i, j), i, j);
                if (score_of(c) > T.floor_min) flag = true;
            }
            T.HH[i] = T.CC[n1]; T.WW[i] = f;
        }
        for (rl = m1, cl = n1;;) {
            bool rflag = true, cflag = true;
            while ((rflag && m1 > 1) || (cflag && n1 > 1)) {
                if (rflag && m1 > 1) {                    // one more row on top
                    rflag = false;
                    --m1;
                    cand_t c = pack(0, m1, nn + 1), f = pack(-kQ, m1, nn + 1), p = pack(0, m1 + 1, nn + 1);
                    const int ai = T.a[m1 - 1];
                    bool hit = false;
                    for (int j = nn; j >= n1; --j) {
                        cell(c, f, T.CC[j], T.DD[j], p, subst(ai, T.b[j - 1]), taken(T, m1, j), m1, j);
                        if (score_of(c) > T.floor_min) flag = true;
                        const int sc = start_of(c), sd = start_of(T.DD[j]), sf = start_of(f);
                        hit = (start_row(sc) > rl && start_col(sc) > cl) || (start_row(sd) > rl && start_col(sd) > cl) ||
                              (start_row(sf) > rl && start_col(sf) > cl);
                        if (!rflag && hit) rflag = true;
                    }
                    T.HH[m1] = T.CC[n1]; T.WW[m1] = f;
                    if (!cflag && hit) cflag = true;
                }
                if (cflag && n1 > 1) {                    // one more column on the left
                    cflag = false;
                    --n1;
                    cand_t c = pack(0, mm + 1, n1), f = pack(-kQ, mm + 1, n1), p = pack(0, mm + 1, n1 + 1);
                    const int bj = T.b[n1 - 1];
                    bool hit = false;
                    for (int i = mm; i >= m1; --i) {
                        cell(c, f, T.HH[i], T.WW[i], p, subst(bj, T.a[i - 1]), taken(T, i, n1), i, n1);
                        if (score_of(c) > T.floor_min) flag = true;
                        const int sc = start_of(c), sd = start_of(T.WW[i]), sf = start_of(f);
                        hit = (start_row(sc) > rl && start_col(sc) > cl) || (start_row(sd) > rl && start_col(sd) > cl) ||
                              (start_row(sf) > rl && start_col(sf) > cl);
                        if (!cflag && hit) cflag = true;
                    }
                    T.CC[n1] = T.HH[m1]; T.DD[n1] = f;
                    if (!rflag && hit) rflag = true;
                }
            }
            if ((m1 == 1 && n1 == 1) || no_cross(T.list, T.numnode, m1, mm, n1, nn, rl, cl)) break;
        }
        --m1; --n1;
        if (flag) {
            for (int j = n1 + 1; j <= nn; ++j) { T.CC[j] = pack(0, m1, j); T.DD[j] = pack(-kQ, m1, j); }
            for (int i = m1 + 1; i <= mm; ++i) {
                cand_t c = pack(0, i, n1), f = pack(-kQ, i, n1), p = pack(0, i - 1, n1);
                const int ai = T.a[i - 1];
                for (int j = n1 + 1; j <= nn; ++j) {
                    cell(c, f, T.CC[j], T.DD[j], p, subst(ai, T.b[j - 1]), taken(T, i, j), i, j);
                    if (score_of(c) > T.floor_min) { addnode(T.list, T.numnode, score_of(c), start_of(c), i, j); T.floor_min = 1; }
                }
            }
        }
    }
}

}  // namespace simk
}  // namespace ltg
