// Scan stage: translation (rules.h) fused into a packed 16-bit x2 affine-gap Smith-Waterman that
// yields, per DNA column, the maximum score over all (padded) RNA rows — the device counterpart of
// calc_score_once (stats.h:879) + ssw_pre_align / sw_sse2_byte_once (sswNew.cpp:1309, 255) — and the
// threshold-hit compaction + run merge of Aligner::preAlign (ssw_cpp.cpp:442-572).
//
// Work unit ("item") = one DNA segment x one *pair* of tasks that read the segment in the same
// direction; the two tasks live in the two 16-bit halves of every register (same RNA, same DNA,
// different 5-letter rule image), so the halves are independent cells with identical control flow.
//
// One warp owns one item.  Lanes own R consecutive RNA rows each (a strip of 32*R rows); DNA columns
// stream through the lanes as an anti-diagonal wavefront (lane l works on column s-l at step s), the
// hand-off of (H, F, running column max) to the next lane is a warp shuffle.  RNAs longer than one
// strip are strip-mined; the strip boundary row (H, F, column max per column) round-trips through an
// L2-resident buffer.  Query profiles (score of every RNA row against each of the 5 DNA base codes,
// already packed for the two tasks) sit in shared memory; the segment's base codes are staged there too.
//
// Per cell pair (two tasks) the recurrence is 6 instructions:
//   t = VIADDMNMX.RELU(Hdiag, s, E)     t  = max(Hdiag + s, E, 0)
//   u = VIADD.16x2(t, -16)              (FMA pipe)
//   E = VIADDMNMX(E, -4, u)
//   H = VIMNMX(t, F)
//   F = VIADDMNMX(F, -4, u)
//   cm = VIMNMX3(cm, t, t')             (one per two rows)
// Dropping the F->E and E->F openings is exact for these penalties (a gap directly followed by a gap
// of the other kind is always dominated by a diagonal step: 2*16 > 4+16).
#pragma once
#include "common.cuh"

namespace ltg {

struct SegDesc {
    int64_t start;    // offset of the segment in the record
    int32_t len;
    int32_t flags;    // bit0: skip (homopolymer, same_seq Fasim-LongTarget.cpp:873), bit1: contains a non-ACGT byte
};
constexpr int kSegSkip = 1, kSegNonACGT = 2;

struct ScanItem {
    int32_t seg;
    int32_t pair;
};

__constant__ TaskDef c_tasks[kMaxTasks];
__constant__ PairDef c_pairs[kMaxPairs];

// ---------------------------------------------------------------------------------------------
// ASCII -> base codes
__global__ void k_encode(const unsigned char* __restrict__ dna, uint8_t* __restrict__ codes, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) codes[i] = (uint8_t)dna_code(dna[i]);
}

// 2-bit packed DNA (UCSC .2bit coding: 4 bases per byte, the first base in the two high bits, T0 C1 A2 G3) -> ASCII + base codes.
// `first` = index of the record's first base inside `packed` (any value: regions need not start on a byte); the N runs of the
// record come as sorted (start, size) blocks relative to that first base, like the nBlock arrays of a .2bit record.
__global__ void k_unpack_2bit(const uint8_t* __restrict__ packed, int64_t first, int64_t n, const uint32_t* __restrict__ nstart,
                              const uint32_t* __restrict__ nsize, int n_blocks, unsigned char* __restrict__ dna, uint8_t* __restrict__ codes)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t g = first + i;
        const unsigned b = (packed[g >> 2] >> (6 - 2 * (int)(g & 3))) & 3u;
        // last block that starts at or before i (binary search), then the range test
        int lo = 0, hi = n_blocks;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)nstart[mid] <= i) lo = mid + 1; else hi = mid; }
        const bool is_n = lo > 0 && i < (int64_t)nstart[lo - 1] + (int64_t)nsize[lo - 1];
        const unsigned char ch = is_n ? 'N' : (unsigned char)"TCAG"[b];
        dna[i] = ch;
        codes[i] = (uint8_t)dna_code(ch);
    }
}

// cutSequence on the device (fastsim.h:71-90): descriptor k of a shard = segment first_seg + k of a record of record_len bases
// whose bytes start at `base` in the call's DNA buffer; the flags follow from k_seg_flags
__global__ void k_cut_segments(SegDesc* __restrict__ segs, int n_segs, int64_t base, int64_t first_seg, int64_t stride, int cut, int64_t record_len)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_segs) return;
    SegDesc sd;
    sd.start = base + (int64_t)k * stride;
    const int64_t left = record_len - (first_seg + k) * stride;
    sd.len = (int32_t)(left < cut ? left : cut);
    sd.flags = 0;
    segs[k] = sd;
}

// per-segment flags: homopolymer test of same_seq (all bytes equal and one of A C G T U N) and
// "has a byte outside ACGT" (needs the second, N-aware threshold scoring — SURVEY App. B Q3)
__global__ void k_seg_flags(const unsigned char* __restrict__ dna, SegDesc* __restrict__ segs, int n_segs)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_segs) return;
    const SegDesc sd = segs[warp];
    const unsigned char* p = dna + sd.start;
    const unsigned char c0 = p[0];
    bool same = true, clean = true;
    for (int i = lane; i < sd.len; i += 32) {
        const unsigned char c = p[i];
        same &= (c == c0);
        clean &= (c == 'A' || c == 'C' || c == 'G' || c == 'T');
    }
    same = __all_sync(0xffffffffu, same);
    clean = __all_sync(0xffffffffu, clean);
    if (lane == 0) {
        const bool letter = (c0 == 'A' || c0 == 'C' || c0 == 'G' || c0 == 'T' || c0 == 'U' || c0 == 'N');
        segs[warp].flags = ((same && letter) ? kSegSkip : 0) | (clean ? 0 : kSegNonACGT);
    }
}

// ---------------------------------------------------------------------------------------------
// Query profiles.  Layout per (pair, strip): [x = 0..4][k = 0..R/4-1][lane = 0..31][e = 0..3] uint32,
// entry = packed scores of RNA row (strip*32R + lane*R + 4k + e) against base code x under the two
// tasks' rule images.  kind 0: SSW scoring (ssw_cpp.cpp:28-53; pad rows up to 16*ceil(m/16) score 0 as
// in qP_byte sswNew.cpp:195); kind 1: Farrar-side scoring of calc_score_once (stats.h:211-228: N -1, U==T).
template <int R>
__global__ void k_build_profiles(const uint8_t* __restrict__ rna_ssw, const uint8_t* __restrict__ rna_stats, int m,
                                 int n_pairs, int n_strips, int kind, uint32_t* __restrict__ out)
{
    const int per_block = 5 * 32 * R;
    const int64_t total = (int64_t)n_pairs * n_strips * per_block;
    const int m16 = 16 * ((m + 15) / 16);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int rem = (int)(idx % per_block);
        const int64_t blk = idx / per_block;
        const int strip = (int)(blk % n_strips), pair = (int)(blk / n_strips);
        const int x = rem / (32 * R);
        rem -= x * 32 * R;
        const int k = rem / 128, lane = (rem % 128) / 4, e = rem & 3;
        const int row = strip * 32 * R + lane * R + 4 * k + e;
        int sc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const TaskDef& t = c_tasks[c_pairs[pair].task[h]];
            const int d = t.img[x];                 // translated base, SSW code 0..3 / 4 = N
            int s;
            if (row >= m16) s = kGhost;
            else if (row >= m) s = 0;
            else if (kind == 0 || kind == 2) {
                const int q = rna_ssw[row];
                s = (q == d && d < 4) ? kMatch : kMismatch;
                if (kind == 2) s *= 2;              // taint sweep (k_scan TAINT): values are 2 * score - taint bit
            } else {
                const int q = rna_stats[row];       // 0 A,1 C,2 G,3 T,4 U,5 N
                if (q == 5 || d == 4) s = -1;
                else if (q == d || (q == 4 && d == 3)) s = kMatch;
                else s = kMismatch;
            }
            sc[h] = s;
        }
        out[idx] = pack16(sc[0], sc[1]);
    }
}

// ---------------------------------------------------------------------------------------------
struct ScanArgs {
    const uint8_t* codes;       // base codes of the record
    const SegDesc* segs;
    const ScanItem* items;
    int n_items;
    const uint32_t* profiles;   // [pair][strip][5*32*R]
    int n_strips;
    int max_len;                // row pitch of colmax / bnd (>= longest segment)
    uint32_t* colmax_all;       // [item][max_len] packed column maxima of the two tasks over ALL RNA rows (exact, what ssw_pre_align
                                // reports before the Q2 truncation)
    uint16_t* blkmax;           // [item][granule][blk_pitch] per granule of kGranRows RNA rows (n_strips * 32 * R / kGranRows per item) and
                                // per block of kBlkCols wavefront steps: the maxima of the two tasks, saturated to 8 bits each
                                // (task 0 in the low byte) — upper bounds for the window stage's row pruning and the Q4 pre-filter
    int blk_pitch;
    uint2* bnd;                 // [persistent warp][max_len] strip boundary packets (H, F)
    int* counter;               // work queue head
    // shared-profile variant (k_scan<R, W, false, true>): the CTA's warps work on items of ONE task pair and share one copy of
    // that pair's profiles (all strips) in shared memory; work comes in groups of W items of the same pair
    const int* order;           // [group][W] item index, or -1
    const int* group_pair;      // [group]
    int n_groups;
    // Q4 probe variant (k_scan<R, W, true>): no column maxima are written; per item the largest F value carried into a
    // row that starts a stripe of the reference's 16-lane layout (rows k * stripe_len) within the recorded columns.
    // Q4 taint variant (k_scan<R, W, false, false, false, true>, see taint_slow_step): `profiles` hold doubled scores, colmax_all
    // receives 2 * maximum - taint bit per column, probe_out bit 0 / bit 16 = "gave up" for task 0 / 1
    uint32_t* probe_out;        // [item] packed (task 0 | task 1 << 16)
    const int* task_flags;      // taint variant: [seg * T + task] flags of epilogue mode 0 (which tasks of a pair are flagged)
    const int* task_jstar;      // [seg * T + task] first column the reference does not record any more (or n)
    int tasks_per_seg;
    int stripe_len;             // ceil(m / 16)
    // Stripe-start screen recorded by the main sweep (k_scan<R, W, false, false, true>): per stripe start of the reference's
    // layout and per block of kBlkCols wavefront steps, max(fin + 16, largest cell of the lane that holds the stripe start) as two
    // saturated bytes like blkmax.  An F >= 132 can only enter that stripe start in a column where this reaches 148: a pre-filter
    // with the resolution of one lane (R rows) instead of the granules around the stripe start, for 3 instructions per step
    uint16_t* frec;             // [item][kFrecRows][blk_pitch]
};
constexpr int kFrecRows = 15;   // stripe starts k * stripe_len, k = 1..15
// A lane of the scan (R consecutive rows from row0) that holds stripe starts records all of them in the row of the FIRST one:
// index (k - 1) of that stripe start, or -1 when the lane holds none
__host__ __device__ inline int frec_row_of_lane(int row0, int R, int stripe_len)
{
    const int kf = row0 <= 0 ? 1 : (row0 + stripe_len - 1) / stripe_len;
    return (kf <= kFrecRows && kf * stripe_len < row0 + R) ? kf - 1 : -1;
}

// The column maxima are kept PER GRANULE of kGranRows / R lanes (kGranRows RNA rows; the running maximum that travels
// along the lanes restarts at every granule head and the granule's last lane — its "tail" — holds the granule's maximum of
// a column).  Nothing of that is written per cell column any more: the tail lanes park the values in a shared-memory ring;
// once per 32 steps the warp folds the granules of the 32 columns that just became complete into the exact whole-column
// maximum (one coalesced 128-byte row update), and each tail lane keeps a running maximum over blocks of kBlkCols steps
// that it stores as two saturated bytes.  Block b of the granule whose tail is lane t covers the columns
// [kBlkCols * b - t, kBlkCols * b - t + kBlkCols) (the wavefront skew), see blk_of().
constexpr int kGranRows = 128;                   // RNA rows per granule (kGranRows / R lanes; 32 * R / kGranRows granules per strip)
constexpr int kBlkCols = 16;
__host__ __device__ inline int blk_pitch_for(int max_len) { return (max_len + 31 + kBlkCols - 1) / kBlkCols + 1; }
// tail lane of granule g of an item (the lane layout repeats in every strip)
__host__ __device__ inline int gran_tail_lane(int g, int scan_r) { const int gl = kGranRows / scan_r, gps = 32 / gl; return (g % gps) * gl + gl - 1; }
// block that holds column j of a granule with tail lane t
__host__ __device__ inline int blk_of(int j, int tail_lane) { return (j + tail_lane) / kBlkCols; }

constexpr int kCringPitch = 68;                  // words per granule row of the ring: 64 slots + 4 (rows land in different banks, rows stay 16-byte aligned)
template <int R>
__host__ __device__ constexpr int scan_warp_smem_bytes(int max_len)
{
    return 5 * 32 * R * 4 + 32 * 8 + (32 / (kGranRows / R)) * kCringPitch * 4 + ((max_len + 64 + 15) / 16) * 16;
}
// shared-profile variant: per warp only the rings and the base codes; the profiles (n_strips * 5 * 32 * R * 4 bytes) once per CTA
template <int R>
__host__ __device__ constexpr int scan_warp_smem_bytes_shared(int max_len)
{
    return 32 * 8 + (32 / (kGranRows / R)) * kCringPitch * 4 + ((max_len + 64 + 15) / 16) * 16;
}

// ---------------------------------------------------------------------------------------------
// Q4 certification by taint tracking (CPU prototype and soundness harness: tests/test_q4_theory_cpu.py
// `certify(kernel_rule=True)`).  The sweep computes exact Smith-Waterman on values x = 2 * score - tau, tau = 1 meaning "the
// reference's kernel (signed lazy-F test, sswNew.cpp:369) may hold a LOWER value here"; a maximum prefers the untainted side
// on ties, adding even numbers keeps the bit, the zero floor is untainted.  The quirk can only drop contributions of an F chain
// that entered a stripe start of the reference's layout with >= 132, from the row after the chain has passed through [132, 143]
// (its value is <= 139 then): those contributions are tainted.  A chain that is itself tainted and >= 132 at a stripe start is
// beyond the model ("give up").  A task whose recorded column maxima are all untainted has exactly the reference's maxima.
//
// Most steps of a lane cannot start or carry such a chain (screened before the cells, see LTG_SCAN_STEP): they run the plain
// packed recurrence, in which the chain through a stripe start is simply the merged F.  taint_slow_step is the general form of
// one lane-step: the F chain split into the part opened inside the stripe (fs) and the carried chain (fc, with its "started
// >= 132" flag), per 16-bit half; H, E and the scores stay packed.  fcin / fcout: carried chain per half, bit 15 = flag.
template <int R>
__device__ __forceinline__ void taint_slow_step(uint32_t (&Hd)[R], uint32_t (&E)[R], const uint4 (&sc)[R / 4], uint32_t hdiag, uint32_t fin,
                                                uint32_t fcin, uint32_t bmask, uint32_t live, uint32_t& cm, uint32_t& hlast, uint32_t& fout,
                                                uint32_t& fcout, uint32_t& giveup)
{
    const uint32_t kOpen2 = 0xFFE0FFE0u, kExt2 = 0xFFF8FFF8u;
    int fs[2] = {lo16(fin), hi16(fin)};
    int fc[2] = {(int)(fcin & 0x7FFFu), (int)((fcin >> 16) & 0x7FFFu)};
    int og[2] = {(int)((fcin >> 15) & 1u), (int)(fcin >> 31)};
    uint32_t d = hdiag;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (bmask & (1u << r)) {                     // this row starts a stripe: the chain that crosses is the larger of the two
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int fend = fs[h], old = fc[h];
                if (((fend + 1) >> 1) >= ((old + 1) >> 1)) { fc[h] = fend; og[h] = ((fend + 1) >> 1) >= kQ4CarryF; }
                fs[h] = 0;
                if ((fc[h] & 1) && ((fc[h] + 1) >> 1) >= kQ4CarryF && ((live >> h) & 1u)) giveup |= 1u << (16 * h);
            }
        }
        int cb[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool cut = og[h] && ((fc[h] + 1) >> 1) <= 139;
            cb[h] = cut ? ((fc[h] - 1) | 1) : fc[h];
        }
        const uint32_t sv = r % 4 == 0 ? sc[r / 4].x : r % 4 == 1 ? sc[r / 4].y : r % 4 == 2 ? sc[r / 4].z : sc[r / 4].w;
        const uint32_t t = __viaddmax_s16x2_relu(d, sv, E[r]);
        const uint32_t u = __vadd2(t, kOpen2);
        E[r] = __viaddmax_s16x2(E[r], kExt2, u);
        const uint32_t hh = __vmaxs2(__vmaxs2(t, pack16(fs[0], fs[1])), pack16(cb[0], cb[1]));
        fs[0] = max(fs[0] - 8, lo16(u)); fs[1] = max(fs[1] - 8, hi16(u));
        fc[0] = max(fc[0] - 8, 0); fc[1] = max(fc[1] - 8, 0);
        og[0] &= (int)(fc[0] > 0); og[1] &= (int)(fc[1] > 0);
        d = Hd[r]; Hd[r] = hh;
        cm = __vmaxs2(cm, t);
        hlast = hh;
    }
    fout = pack16(fs[0], fs[1]);
    fcout = (uint32_t)fc[0] | ((uint32_t)og[0] << 15) | ((uint32_t)fc[1] << 16) | ((uint32_t)og[1] << 31);
}

// E update: fused VIADDMNMX (2 ALU-pipe slots) or VIADD on the FMA pipe + VIMNMX (1 ALU-pipe slot); the split form
// trades one issue slot for one ALU-pipe slot (measured rates: profiles/int_simd_peak.json)
#ifdef LTG_SCAN_SPLIT_E
// (the empty asm keeps the compiler from fusing the add and the max back into one VIADDMNMX — without it this variant compiles to
//  the very same SASS as the fused form, which is what round 1 unknowingly measured)
#define LTG_E_UPDATE(EV, U) { uint32_t e_ = __vadd2(EV, kNegExt); asm volatile("" : "+r"(e_)); EV = __vmaxs2(e_, (U)); }
#else
#define LTG_E_UPDATE(EV, U) EV = __viaddmax_s16x2(EV, kNegExt, (U))
#endif

// Carried F (PROBE and FREC variants): the F that enters a row which starts a stripe of the reference's layout.  Taking it out
// of the cell loop row by row costs a branch (or 3 instructions) per row; instead every step is SCREENED after its cells with 4
// instructions — an F >= 132 entering any row of this lane needs fin >= 132 or a cell >= 148 in this lane's part of the column,
// and the running column maximum cm covers those cells — and only a step that passes (rare: never on random DNA) rebuilds
// the lane's F chain from the column's H values, which sit in Hd[] after the step: F(r+1) = max(F(r) - 4, H(r) - 16)
// (identical to the chain of the cell loop, F - 16 < F - 4).  ACC: accumulator, MASKED: apply vmask (column range) to a value.
#define LTG_CARRIED_F(ACC, MASKED)                                                                              \
            if (amask) {                                                                                        \
                const uint32_t x_ = __viaddmax_s16x2(fin, 0x00100010u, cm);                                     \
                const bool trig_ = bmask != 0 && __vmaxs2(x_, 0x00930093u) != 0x00930093u;                      \
                if (__any_sync(0xffffffffu, trig_)) {                                                           \
                    uint32_t g_ = fin;                                                                          \
                    _Pragma("unroll") for (int r_ = 0; r_ < R; ++r_) {                                          \
                        if (bmask & (1u << r_)) ACC = __vmaxs2(ACC, (MASKED) ? (g_ & vmask) : g_);              \
                        g_ = __viaddmax_s16x2(g_, kNegExt, __vadd2(Hd[r_], kNegOpen));                          \
                    }                                                                                           \
                }                                                                                               \
            }
#define LTG_CELL(SV, RR)                                          \
    {                                                             \
        const uint32_t t_ = __viaddmax_s16x2_relu(d, (SV), E[RR]); \
        const uint32_t u_ = __vadd2(t_, kNegOpen);                \
        LTG_E_UPDATE(E[RR], u_);                                  \
        const uint32_t h_ = __vmaxs2(t_, f);                      \
        f = __viaddmax_s16x2(f, kNegExt, u_);                     \
        d = Hd[RR];                                               \
        Hd[RR] = h_;                                              \
        tv[(RR) & 1] = t_;                                        \
        if ((RR) & 1) cm = __vimax3_s16x2(cm, tv[0], tv[1]);        \
        hlast = h_;                                               \
    }

template <int R, int WARPS, bool PROBE = false, bool SHARED = false, bool FREC = false, bool TAINT = false>
__global__ void __launch_bounds__(WARPS * 32) k_scan(const ScanArgs a)
{
    static_assert(R % 4 == 0, "R must be a multiple of 4");
    static_assert(!(PROBE && SHARED), "the probe sweep keeps per-warp profiles");
    static_assert(!(FREC && (PROBE || SHARED)), "carried-F recording belongs to the plain main sweep");
    static_assert(!(TAINT && (PROBE || SHARED || FREC)), "the taint sweep is a variant of its own");
    extern __shared__ uint4 smem_u4[];
    __shared__ int s_group;
    __shared__ alignas(8) unsigned long long s_bar[WARPS];      // per warp: mbarrier of the strip-profile bulk copy
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int PLANE = (R / 4) * 32;     // uint4 per base-code plane
    const int warp_bytes = SHARED ? scan_warp_smem_bytes_shared<R>(a.max_len) : scan_warp_smem_bytes<R>(a.max_len);
    // SHARED: [profiles of every strip][per warp: rings, codes]; otherwise per warp: [profile of the strip in flight, rings, codes]
    uint4* const s_prof_all = smem_u4;
    unsigned char* const warp_base = reinterpret_cast<unsigned char*>(smem_u4) + (SHARED ? (size_t)a.n_strips * 5 * PLANE * 16 : 0) + (size_t)wib * warp_bytes;
    uint4* s_prof = reinterpret_cast<uint4*>(warp_base);
    uint2* s_ring = SHARED ? reinterpret_cast<uint2*>(warp_base) : reinterpret_cast<uint2*>(s_prof + 5 * PLANE);
    constexpr int kGranLanes = kGranRows / R, kGranPerStrip = 32 / kGranLanes;
    static_assert(kGranLanes >= 1 && kGranLanes * R == kGranRows, "R must divide the granule");
    uint32_t* s_cring = reinterpret_cast<uint32_t*>(s_ring + 32);          // [granule of the strip][step & 63] granule maxima in flight
    uint8_t* s_codes = reinterpret_cast<uint8_t*>(s_cring + kGranPerStrip * kCringPitch);
    uint2* bnd = a.bnd + (size_t)(blockIdx.x * WARPS + wib) * a.max_len;
    const uint32_t kNegOpen = TAINT ? 0xFFE0FFE0u : 0xFFF0FFF0u, kNegExt = TAINT ? 0xFFF8FFF8u : 0xFFFCFFFCu;      // TAINT: doubled scores
    int cur_pair = -1;
    // Staging of the strip profile (20 KB at R = 32), once per warp and strip: ONE bulk asynchronous copy (cp.async.bulk, the 1-D
    // TMA path: SASS UBLKCP) that completes on the warp's own mbarrier, instead of 40 LDG.128 + STS.128 per lane.  The copy is
    // ~0.1 % of a strip's time, so the two forms measure the same (7086 vs 7072 GCUPS, profiles/README.md; -DLTG_NO_BULK keeps
    // the loop).  The wait loop leaves on a warp vote: with a per-lane exit the compiler treated the warp as possibly diverged
    // for the rest of the kernel (BRA.DIV before every shuffle, +8 registers) and the kernel lost 4.5 %.
    const uint32_t sa_bar = (uint32_t)__cvta_generic_to_shared(&s_bar[wib]);
    uint32_t bar_phase = 0;
    (void)sa_bar; (void)bar_phase;
#ifndef LTG_NO_BULK
    if (!SHARED) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sa_bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
#endif

    for (;;) {
        int item = 0;
        if (SHARED) {
            // one group of WARPS items of the same task pair per round; the pair's profiles are (re)loaded when the pair changes
            if (threadIdx.x == 0) s_group = atomicAdd(a.counter, 1);
            __syncthreads();
            const int g = s_group;
            __syncthreads();                          // everyone holds g (and has left the previous group's profiles) before either changes
            if (g >= a.n_groups) break;
            const int pair = a.group_pair[g];
            if (pair != cur_pair) {
                const uint4* gp = reinterpret_cast<const uint4*>(a.profiles) + (size_t)pair * a.n_strips * (5 * PLANE);
                for (int i = threadIdx.x; i < a.n_strips * 5 * PLANE; i += WARPS * 32) s_prof_all[i] = gp[i];
                cur_pair = pair;
                __syncthreads();
            }
            item = a.order[g * WARPS + wib];
            if (item < 0) continue;                   // a short group: this warp idles until the next round's barrier
        } else {
            if (lane == 0) item = atomicAdd(a.counter, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= a.n_items) break;
        }
        const ScanItem it = a.items[item];
        const SegDesc sd = a.segs[it.seg];
        const int n_full = sd.len;
        int n = n_full;                                            // columns this sweep works on
        const bool rev = c_pairs[it.pair].reversed != 0;
        const uint8_t* gcodes = a.codes + sd.start;
        int jst0 = 0, jst1 = 0;
        if (PROBE || TAINT) {
            const PairDef pd = c_pairs[it.pair];
            // columns the reference processes: up to and including the one where it stops recording (Q2)
            jst0 = min(a.task_jstar[it.seg * a.tasks_per_seg + pd.task[0]] + 1, n_full);
            jst1 = min(a.task_jstar[it.seg * a.tasks_per_seg + pd.task[1]] + 1, n_full);
            if (TAINT) {
                // the verdict only reads the columns a FLAGGED task still records: the sweep ends there (tasks that overflow
                // early, the common case on repeat-rich DNA, need a fraction of the segment)
                const int* tf = a.task_flags + it.seg * a.tasks_per_seg;
                const int need = max((tf[pd.task[0]] & 2 /* kTaskLiteral */) ? jst0 : 0, (tf[pd.task[1]] & 2) ? jst1 : 0);
                n = max(1, min(n_full, need));
            }
        }
        __syncwarp();
        // stage base codes with 32 neutral ('N' plane) columns on both sides
        for (int i = lane; i < n + 64; i += 32) {
            const int j = i - 32;
            s_codes[i] = (j < 0 || j >= n) ? (uint8_t)kBaseOther : gcodes[rev ? (n_full - 1 - j) : j];
        }
        uint32_t* cm_all = a.colmax_all + (size_t)item * a.max_len;
        uint16_t* blk_item = a.blkmax + (size_t)item * a.n_strips * kGranPerStrip * a.blk_pitch;
        const bool gran_head = (lane & (kGranLanes - 1)) == 0, gran_tail = (lane & (kGranLanes - 1)) == kGranLanes - 1;

        uint32_t fb = 0;                                           // PROBE: running maximum of the carried F values
        uint32_t giveup = 0;                                       // TAINT: bit 0 / bit 16 = the model gave up on task 0 / 1

        for (int strip = 0; strip < a.n_strips; ++strip) {
            // PROBE: which of this lane's rows start a stripe of the reference's layout (row = k * stripe_len, k = 1..15)
            uint32_t bmask = 0, vmask = 0, amask = 0;
            uint32_t fr = 0;                                       // FREC: screen value (see LTG_SCAN_STEP) of the block of steps in flight
            uint16_t* frec_row = nullptr;
            if (FREC) {
                const int fk = frec_row_of_lane((strip * 32 + lane) * R, R, a.stripe_len);
                if (fk >= 0) frec_row = a.frec + ((size_t)item * kFrecRows + fk) * a.blk_pitch;
            }
            if (PROBE || FREC || TAINT) {
                const int row0 = (strip * 32 + lane) * R;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int row = row0 + r;
                    if (row > 0 && row < 16 * a.stripe_len && row % a.stripe_len == 0) bmask |= 1u << r;
                }
                amask = __reduce_or_sync(0xffffffffu, bmask);
            }
            const uint32_t bm_all = bmask ? 0xFFFFFFFFu : 0u;     // FREC: only lanes that hold a stripe start record
            if (SHARED) s_prof = s_prof_all + (size_t)strip * (5 * PLANE);
            else {
                const uint4* gp = reinterpret_cast<const uint4*>(a.profiles) + ((size_t)it.pair * a.n_strips + strip) * (5 * PLANE);
                __syncwarp();                                      // every lane is done with the previous strip's profile
#ifdef LTG_NO_BULK
                for (int i = lane; i < 5 * PLANE; i += 32) s_prof[i] = gp[i];
#else
                if (lane == 0) {
                    constexpr uint32_t kBytes = 5 * PLANE * 16;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sa_bar), "r"(kBytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"((uint32_t)__cvta_generic_to_shared(s_prof)), "l"(gp), "r"(kBytes), "r"(sa_bar) : "memory");
                }
                for (;;) {                                         // (the exit is a vote: the warp leaves the loop converged)
                    uint32_t done;
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(sa_bar), "r"(bar_phase) : "memory");
                    if (__all_sync(0xffffffffu, done != 0)) break;
                }
                bar_phase ^= 1;
#endif
                __syncwarp();
            }
            const bool first = (strip == 0), last = (strip == a.n_strips - 1);
            // this lane's granule: its row of block maxima, and its row of the ring
            uint16_t* blk_row = blk_item + ((size_t)strip * kGranPerStrip + (lane / kGranLanes)) * a.blk_pitch;
            uint32_t blk = 0;
            uint32_t Hd[R], E[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { Hd[r] = 0; E[r] = 0; }
            uint32_t hout = 0, fout = 0, cmout = 0, hdiag = 0;
            uint32_t fcout = 0, scr = 0;                          // TAINT: carried chain handed to the next lane; bound of this lane's column
            const int steps = n + 31;
            // Steps run in blocks of 32 (the ring of strip-boundary packets is refilled per block).  Shared memory is
            // addressed with explicit 32-bit shared addresses (one add per profile fetch instead of a generic-pointer
            // conversion); the packed scores of step s+1 are loaded while step s computes and the base code two steps
            // ahead is fetched.  Blocks in which every lane is inside its column range ("interior": all but the first
            // and the last two) need no per-step store predicates: they are loop invariants of the lane.
            const uint32_t sa_prof = (uint32_t)__cvta_generic_to_shared(s_prof) + lane * 16;
            const uint32_t sa_code = (uint32_t)__cvta_generic_to_shared(s_codes) + (32 - lane);
            const uint32_t sa_ring = (uint32_t)__cvta_generic_to_shared(s_ring);
            const uint32_t sa_cring = (uint32_t)__cvta_generic_to_shared(s_cring);
            const uint32_t sa_cmine = sa_cring + (lane / kGranLanes) * (kCringPitch * 4);        // this lane's granule row of the ring
            uint4 sc[R / 4];
            uint32_t xn;
            {
                uint32_t x0;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(x0) : "r"(sa_code));
                const uint32_t ad = sa_prof + x0 * (PLANE * 16);
#pragma unroll
                for (int k = 0; k < R / 4; ++k)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(sc[k].x), "=r"(sc[k].y), "=r"(sc[k].z), "=r"(sc[k].w) : "r"(ad + k * 512));
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(xn) : "r"(sa_code + 1));
            }
            uint2* bnd_ptr = bnd - 31;                            // bnd_ptr[0] is lane 31's slot for step s
            const bool st_bnd = (lane == 31) && !last;

#define LTG_SCAN_STEP(GUARD, S, K)                                                                              \
            {                                                                                                   \
                uint4 scn[R / 4];                                                                               \
                {                                                                                               \
                    const uint32_t ad = sa_prof + xn * (PLANE * 16);                                            \
                    _Pragma("unroll") for (int k = 0; k < R / 4; ++k)                                           \
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"                                 \
                                     : "=r"(scn[k].x), "=r"(scn[k].y), "=r"(scn[k].z), "=r"(scn[k].w) : "r"(ad + k * 512)); \
                    const uint32_t ca = sa_code + (GUARD ? (uint32_t)min((S) + 2, steps) : (uint32_t)((S) + 2)); \
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(xn) : "r"(ca));                                \
                }                                                                                               \
                uint32_t hin = __shfl_up_sync(0xffffffffu, hout, 1);                                            \
                uint32_t fin = __shfl_up_sync(0xffffffffu, fout, 1);                                            \
                uint32_t cmin = __shfl_up_sync(0xffffffffu, cmout, 1);                                          \
                if (gran_head) cmin = 0;                                                                        \
                if (lane == 0) {                                                                                \
                    if (first) { hin = 0; fin = 0; }                                                            \
                    else asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hin), "=r"(fin) : "r"(sa_ring + (K) * 8)); \
                }                                                                                               \
                if (PROBE) { const int j_ = (S) - lane; vmask = (j_ < jst0 ? 0xFFFFu : 0u) | (j_ < jst1 ? 0xFFFF0000u : 0u); } \
                if (FREC && GUARD) vmask = ((S) >= lane && (S) - lane < n) ? 0xFFFFFFFFu : 0u;                  \
                uint32_t d = hdiag, f = fin, cm = FREC ? 0u : cmin, hlast = 0;   /* FREC: the lane's own maximum first */ \
                uint32_t tv[2];                                                                                 \
                /* TAINT: a lane-step that may start or carry a chain of >= 132 through a stripe start takes the general form. \
                   Without a chain coming in, that needs fin >= 132 or a cell >= 148 in this lane's part of the column; a cell  \
                   is at most 5 above the previous column's H values, which scr (previous step: cm, fin, chain) and hdiag bound */ \
                uint32_t fcin = 0, fcnew = 0;                                                                   \
                bool slow_ = false;                                                                             \
                if (TAINT) {                                                                                    \
                    fcin = __shfl_up_sync(0xffffffffu, fcout, 1);                                               \
                    if (lane == 0) fcin = 0;                                                                    \
                    const uint32_t x_ = __viaddmax_s16x2(fin, 0x00160016u, __vimax3_s16x2(scr, hdiag, hdiag));  \
                    slow_ = __any_sync(0xffffffffu, fcin != 0 || (bmask != 0 && __vmaxs2(x_, 0x011C011Cu) != 0x011C011Cu)); \
                }                                                                                               \
                if (TAINT && slow_) {                                                                           \
                    const int j_ = (S) - lane;                                                                  \
                    const uint32_t live_ = (j_ >= 0 && j_ < jst0 ? 1u : 0u) | (j_ >= 0 && j_ < jst1 ? 2u : 0u); \
                    taint_slow_step<R>(Hd, E, sc, hdiag, fin, fcin, bmask, live_, cm, hlast, f, fcnew, giveup); \
                } else {                                                                                        \
                    _Pragma("unroll") for (int k = 0; k < R / 4; ++k) {                                         \
                        LTG_CELL(sc[k].x, 4 * k + 0)                                                            \
                        LTG_CELL(sc[k].y, 4 * k + 1)                                                            \
                        LTG_CELL(sc[k].z, 4 * k + 2)                                                            \
                        LTG_CELL(sc[k].w, 4 * k + 3)                                                            \
                    }                                                                                           \
                }                                                                                               \
                if (TAINT) {                                                                                    \
                    scr = __vimax3_s16x2(cm, fin, fcin & 0x7FFF7FFFu);                                          \
                    fcout = fcnew;                                                                              \
                    if (lane == 31 && !last && (fcnew & 0x7FFF7FFFu) != 0) giveup |= 0x00010001u;   /* a chain would cross into the next strip */ \
                }                                                                                               \
                if (PROBE) LTG_CARRIED_F(fb, true)                                                              \
                if (FREC) {     /* branch-free screen: an F >= 132 entering a row of this lane needs fin >= 132 or a cell >= 148 here */ \
                    const uint32_t x_ = __viaddmax_s16x2(fin, 0x00100010u, cm) & bm_all;                        \
                    fr = __vmaxs2(fr, (GUARD) ? (x_ & vmask) : x_);                                             \
                    cm = __vmaxs2(cm, cmin);                                                                    \
                }                                                                                               \
                _Pragma("unroll") for (int k = 0; k < R / 4; ++k) sc[k] = scn[k];                               \
                hdiag = hin;                                                                                    \
                hout = hlast; fout = f; cmout = cm;                                                             \
                if (GUARD) {                                                                                    \
                    if (!PROBE) {                                                                               \
                        const uint32_t cmv = ((S) >= lane && (S) - lane < n) ? cm : 0u;                         \
                        if (gran_tail) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa_cmine + (((S) & 63) << 2)), "r"(cmv) : "memory"); \
                        blk = __vmaxs2(blk, cmv);                                                               \
                    }                                                                                           \
                    if (st_bnd && (S) >= 31 && (S) - 31 < n) bnd_ptr[(K)] = make_uint2(hout, fout);             \
                } else {                                                                                        \
                    if (!PROBE) {                                                                               \
                        if (gran_tail) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa_cmine + (((S) & 63) << 2)), "r"(cm) : "memory"); \
                        blk = __vmaxs2(blk, cm);                                                                \
                    }                                                                                           \
                    if (st_bnd) bnd_ptr[(K)] = make_uint2(hout, fout);                                          \
                }                                                                                               \
            }
            // block maximum of the last kBlkCols steps: two saturated bytes (task 0 low)
#define LTG_BLK_STORE(S)                                                                                        \
            if (!PROBE && !TAINT) {                                                                             \
                if (gran_tail) blk_row[(S) / kBlkCols] = (uint16_t)__byte_perm(__vminu2(blk, 0x00FF00FFu), 0u, 0x4420); \
                blk = 0;                                                                                        \
                if (FREC) {                                                                                     \
                    if (frec_row) frec_row[(S) / kBlkCols] = (uint16_t)__byte_perm(__vminu2(fr, 0x00FF00FFu), 0u, 0x4420); \
                    fr = 0;                                                                                     \
                }                                                                                               \
            }

            for (int s0 = 0; s0 < steps; s0 += 32) {
                if (!first) {
                    __syncwarp();
                    const int j = s0 + lane;
                    uint2 pk = make_uint2(0, 0);
                    if (j < n) pk = bnd[j];
                    s_ring[lane] = pk;
                    __syncwarp();
                }
                if (!TAINT && s0 >= 32 && s0 + 33 < n) {       // (TAINT branches per step anyway: one copy of the step is enough)
                    static_assert(kBlkCols == 16, "two blocks per 32 steps");
#pragma unroll 4
                    for (int k = 0; k < 16; ++k) LTG_SCAN_STEP(false, s0 + k, k)
                    LTG_BLK_STORE(s0)
#pragma unroll 4
                    for (int k = 16; k < 32; ++k) LTG_SCAN_STEP(false, s0 + k, k)
                    LTG_BLK_STORE(s0 + 16)
                } else {
                    const int cnt = min(32, steps - s0);
                    for (int k = 0; k < cnt; ++k) {
                        LTG_SCAN_STEP(true, s0 + k, k)
                        if ((k & (kBlkCols - 1)) == kBlkCols - 1 || k == cnt - 1) { LTG_BLK_STORE(s0 + k) }
                    }
                }
                bnd_ptr += 32;
                if (!PROBE) {
                    // the 32 columns s0-31 .. s0 are complete in this strip (every granule tail has passed them): fold the
                    // granules into the whole-column maximum; one column per lane, one coalesced row update
                    __syncwarp();
                    const int j = s0 - 31 + lane;
                    uint32_t v = 0;
#pragma unroll
                    for (int q = 0; q < kGranPerStrip; ++q) {
                        uint32_t x;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(sa_cring + q * (kCringPitch * 4) + (((j + q * kGranLanes + kGranLanes - 1) & 63) << 2)));
                        v = __vmaxs2(v, x);
                    }
                    if (j >= 0 && j < n) {
                        if (!first) v = __vmaxs2(v, cm_all[j]);
                        cm_all[j] = v;
                    }
                    __syncwarp();
                }
            }
#undef LTG_SCAN_STEP
#undef LTG_BLK_STORE
        }
        if (PROBE) {
#pragma unroll
            for (int o = 16; o; o >>= 1) fb = __vmaxs2(fb, __shfl_xor_sync(0xffffffffu, fb, o));
            if (lane == 0) a.probe_out[item] = fb;
        }
        if (TAINT) {
            giveup = __reduce_or_sync(0xffffffffu, giveup);
            if (lane == 0) a.probe_out[item] = giveup;
        }
    }
}
#undef LTG_CELL
#undef LTG_CARRIED_F
#undef LTG_E_UPDATE

// ---------------------------------------------------------------------------------------------
// Epilogue: per task the exact maximum, the threshold (int)(max*0.8) (Fasim-LongTarget.cpp:413), the
// 8-bit "stop recording" truncation (Q2), threshold-hit compaction and run merge to peaks
// (ssw_cpp.cpp:446-572).  One warp per item; hits are found with __ballot_sync and consumed in column
// order.  Three passes:
//   mode 0  statistics of every task (maximum, threshold, first overflow column, Q4 pre-filter) + peak COUNT of the tasks
//           that stay on the exact path
//   mode 1  peak count of the tasks re-run by the literal emulation (Q4 guard)
//   (exclusive scan of the counts -> one contiguous, position-ordered slice of the peak pool per task)
//   mode 2  peaks written, 32 at a time, one lane per peak
struct EpiArgs {
    const uint16_t* blkmax;      // [item][granule][blk_pitch] saturated block maxima of the scan (Q4 pre-filter)
    int n_gran;                  // granule rows per item
    int blk_pitch, scan_r;
    const uint32_t* colmax_all;  // [item][max_len] exact column maxima over all rows (written by the scan)
    const uint16_t* lit_colmax;  // [literal row][lit_pitch] column maxima of the literal re-runs
    int lit_pitch;
    const int* task_litrow;      // [task] row in lit_colmax (valid when the task carries kTaskLiteral)
    const ScanItem* items;
    const int* bnd_gran;         // mode 0: granules that hold one of the 28 rows above a stripe start of the reference's layout
    int n_bnd_gran;              //         (nullptr: every task that reaches 148 is flagged)
    const uint16_t* frec;        // mode 0: [item][kFrecRows][blk_pitch] carried-F block maxima recorded by the sweep itself (k_scan FREC), or
    int stripe_len;              //         nullptr; with them the verdict is final (no granule pre-filter, no probe sweep)
    const int* item_orig;        // mode 3: [item] index of the item in the batch's full item list (rows of colmax_all)
    const uint32_t* probe;       // mode 3: [item] packed carried-F maxima of the Q4 probe sweep; mode 4: "gave up" bits of the taint sweep
    const uint32_t* taint_colmax;// mode 4: [item][max_len] 2 * column maximum - taint bit (k_scan TAINT)
    const SegDesc* segs;
    int n_items;
    int max_len;
    int tasks_per_seg;
    const int* stats_max;    // [seg*T + task] exact calc_score_once maxima from the N-aware pass, or nullptr
    int stats_all;           // 1: valid for every task; 0: only for the tasks of segments flagged kSegNonACGT
    int mode;
    int* task_max;           // [seg*T + task]
    int* task_thr;
    int* task_npeaks;
    int* task_flags;         // bit0 overflow(>=251) seen, bit1 literal re-run required (Q4 guard), bit2 int16 range exceeded
    int* task_jstar;         // first column >= 251 (or n): nothing is recorded from there on (Q2)
    const int* task_off;     // mode 2: first peak slot of every task
    int* pk_task;
    int* pk_pos;
    int* pk_score;
};
constexpr int kTaskOverflow = 1, kTaskLiteral = 2, kTaskRange = 4;
constexpr int kTaskSkip = 8;     // literal-only batches: a task that shares a pair with a requested one; it has no peaks in this batch

__device__ inline void epi_flush(const EpiArgs& a, int lane, int& nbuf, int base, int task, int bpos, int bscore)
{
    // lanes [0, nbuf) each hold one buffered peak; `base` = slot of the first of them
    if (lane < nbuf) {
        a.pk_task[base + lane] = task;
        a.pk_pos[base + lane] = bpos;
        a.pk_score[base + lane] = bscore;
    }
    nbuf = 0;
}

__global__ void k_epilogue(const EpiArgs a)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.n_items) return;
    const ScanItem it = a.items[warp];
    const SegDesc sd = a.segs[it.seg];
    const int n = sd.len;
    const int row = (a.mode == 3 || a.mode == 4) ? a.item_orig[warp] : warp;
    const uint16_t* blk = a.blkmax + (size_t)row * a.n_gran * a.blk_pitch;
    const uint32_t* cm_all = a.colmax_all + (size_t)row * a.max_len;
    const PairDef pd = c_pairs[it.pair];
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && pd.task[1] == pd.task[0]) break;
        const int task = it.seg * a.tasks_per_seg + pd.task[h];
        auto exact_at = [&](int j) -> int { const uint32_t v = cm_all[j]; return h ? hi16(v) : lo16(v); };
        int thr, jstar = n;
        const uint16_t* lit = nullptr;
        if (a.mode == 0) {
            int mx = 0;
            for (int j = lane; j < n; j += 32) {
                const int s = exact_at(j);
                mx = max(mx, s);
                if (s >= kOverflowU8) jstar = min(jstar, j);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                jstar = min(jstar, __shfl_xor_sync(0xffffffffu, jstar, o));
            }
            const int score = (a.stats_max && (a.stats_all || (sd.flags & kSegNonACGT))) ? a.stats_max[task] : mx;
            thr = (int)((double)score * 0.8);
            // Q4 can only fire where an F >= 132 enters a stripe start, i.e. below a cell >= 148 at most 26 rows up, in a column
            // the reference still processes: without such a cell in the granules that hold those rows the task stays exact
            bool q4 = mx >= kQ4Guard;
            if (q4 && a.frec) {
                // the sweep recorded, per stripe start and block of 16 steps, max(fin + 16, cells of the lane that holds the stripe
                // start): the quirk needs 148 there in a column the reference still processes (blocks: the last one may reach up
                // to 15 columns further, which can only flag more)
                int carried = 0;
                const int jend = min(jstar + 1, n);
                const uint16_t* fbase = a.frec + (size_t)row * kFrecRows * a.blk_pitch;
                for (int k = 1; k <= kFrecRows; ++k) {
                    const int gl = (k * a.stripe_len) / a.scan_r;                      // lane (counted over all strips) that holds the stripe start
                    if (frec_row_of_lane(gl * a.scan_r, a.scan_r, a.stripe_len) != k - 1) continue;    // recorded with an earlier stripe start of that lane
                    const int tl = gl & 31;
                    const uint16_t* rowp = fbase + (size_t)(k - 1) * a.blk_pitch;
                    for (int b = blk_of(0, tl) + lane; b <= blk_of(jend - 1, tl); b += 32) {
                        const int v = rowp[b];
                        carried = max(carried, h ? (v >> 8) : (v & 0xff));
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) carried = max(carried, __shfl_xor_sync(0xffffffffu, carried, o));
                q4 = carried >= kQ4Guard;
            }
            if (q4 && a.bnd_gran) {
                // (block maxima: upper bounds over 16-column blocks, saturated at 255 — coarser than per column, never smaller)
                int near = 0;
                const int jend = min(jstar + 1, n);
                for (int k = 0; k < a.n_bnd_gran; ++k) {
                    const int g = a.bnd_gran[k], tl = gran_tail_lane(g, a.scan_r);
                    const uint16_t* rowp = blk + (size_t)g * a.blk_pitch;
                    for (int b = blk_of(0, tl) + lane; b <= blk_of(jend - 1, tl); b += 32) {
                        const int v = rowp[b];
                        near = max(near, h ? (v >> 8) : (v & 0xff));
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) near = max(near, __shfl_xor_sync(0xffffffffu, near, o));
                q4 = near >= kQ4Guard;
            }
            const int flags = (jstar < n ? kTaskOverflow : 0) | (q4 ? kTaskLiteral : 0) | (mx >= 32000 ? kTaskRange : 0);
            if (lane == 0) {
                a.task_max[task] = score; a.task_thr[task] = thr; a.task_flags[task] = flags; a.task_npeaks[task] = 0;
                a.task_jstar[task] = jstar;
            }
            if (flags & kTaskLiteral) continue;      // counted in mode 3 (cleared by the probe) or 1 (from the literal re-run)
        } else if (a.mode == 3) {
            // Q4 probe verdict: no F >= 132 was carried into a stripe start within the recorded columns, so the reference's
            // signed lazy-F test behaves like an unsigned one and its column maxima are the exact ones: back to the fast path
            const bool is_lit = (a.task_flags[task] & kTaskLiteral) != 0;
            const uint32_t pv = a.probe[warp];
            __syncwarp();
            if (!is_lit || (h ? hi16(pv) : lo16(pv)) >= kQ4CarryF) continue;
            if (lane == 0) a.task_flags[task] &= ~kTaskLiteral;
            thr = a.task_thr[task];
            jstar = a.task_jstar[task];
        } else if (a.mode == 4) {
            // Q4 taint verdict: every column maximum the reference records is untainted and the model never gave up, so the
            // reference's column maxima are the exact ones (scan.cuh, taint_slow_step): back to the fast path
            const bool is_lit = (a.task_flags[task] & kTaskLiteral) != 0;
            const uint32_t gv = a.probe[warp];
            const int jend = min(a.task_jstar[task] + 1, n);
            const uint32_t* tc = a.taint_colmax + (size_t)warp * a.max_len;
            int bad = 0;
            for (int j = lane; j < jend; j += 32) bad |= (int)((h ? (tc[j] >> 16) : tc[j]) & 1u);
            bad = __any_sync(0xffffffffu, bad != 0) ? 1 : 0;
            // (doubled values must stay inside 16 bits: a task that scores >= 16000 is not judged)
            if (!is_lit || bad || (h ? (gv >> 16) : (gv & 0xffffu)) != 0 || a.task_max[task] >= 16000) continue;
            if (lane == 0) a.task_flags[task] &= ~kTaskLiteral;
            thr = a.task_thr[task];
            jstar = a.task_jstar[task];
        } else {
            const bool is_lit = (a.task_flags[task] & kTaskLiteral) != 0;
            if (a.task_flags[task] & kTaskSkip) continue;
            if (a.mode == 1 && !is_lit) continue;
            thr = a.task_thr[task];
            jstar = is_lit ? n : a.task_jstar[task];     // literal maxima already carry the stop-recording zeros
            if (is_lit && a.lit_colmax == nullptr) continue;      // deferred to a literal-only batch: no peaks here
            if (is_lit) lit = a.lit_colmax + (size_t)a.task_litrow[task] * a.lit_pitch;
        }
        const bool write = (a.mode == 2);
        const int base = write ? a.task_off[task] : 0;
        // hits in column order, run merge (ssw_cpp.cpp:470-572): consecutive hits < 5 apart form a run,
        // a run reports its first maximum
        bool have = false;
        int last_pos = 0, best_pos = 0, best_score = 0, npk = 0;
        int nbuf = 0, bpos = 0, bscore = 0;
        for (int j0 = 0; j0 < jstar; j0 += 32) {
            const int j = j0 + lane;
            int s = 0;
            if (j < jstar) s = lit ? (int)lit[j] : exact_at(j);
            unsigned mask = __ballot_sync(0xffffffffu, j < jstar && s > thr);
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                const int pos = j0 + b;
                const int sc = __shfl_sync(0xffffffffu, s, b);
                if (have && pos - last_pos < 5) {
                    if (sc > best_score) { best_score = sc; best_pos = pos; }
                } else {
                    if (have) {
                        if (lane == nbuf) { bpos = best_pos; bscore = best_score; }
                        ++nbuf; ++npk;
                        if (nbuf == 32) { if (write) epi_flush(a, lane, nbuf, base + npk - 32, task, bpos, bscore); else nbuf = 0; }
                    }
                    have = true; best_score = sc; best_pos = pos;
                }
                last_pos = pos;
            }
        }
        if (have) {
            if (lane == nbuf) { bpos = best_pos; bscore = best_score; }
            ++nbuf; ++npk;
        }
        if (write) { if (nbuf) epi_flush(a, lane, nbuf, base + npk - nbuf, task, bpos, bscore); }
        else if (lane == 0) a.task_npeaks[task] = npk;
    }
}

// literal-only batches: the literal flag of every task is set from a list decided earlier (epilogue mode 0 of such a batch ran
// without any filter, so every listed task carries the flag already); the unlisted ones are marked kTaskSkip: modes 1 and 2 of
// the epilogue pass over them (their rows would be dropped on the host anyway), so they own no slot of the peak pool
__global__ void k_force_literal(int* __restrict__ task_flags, int* __restrict__ task_npeaks, const unsigned char* __restrict__ want, int n_tasks)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    task_flags[t] = (task_flags[t] & ~kTaskLiteral) | (want[t] ? kTaskLiteral : kTaskSkip);
    if (!want[t]) task_npeaks[t] = 0;        // (mode 0 counted the peaks of the unflagged ones)
}

// exclusive prefix sum of the per-task peak counts (one block; n is a few 10^4..10^5)
__global__ void __launch_bounds__(1024) k_exclusive_scan(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ total)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? in[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        if (w == 0) {
            const int t = s_warp[lane];
            int y = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int z = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += z; }
            s_warp[lane] = y - t;            // exclusive offset of every warp
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < n) out[i] = carry + s_warp[w] + x - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[w] + x;
        __syncthreads();
    }
    if (tid == 0) *total = s_carry;
}

// maximum of each half of an item's packed granule maxima (used for the N-aware threshold pass)
__global__ void k_rowmax(const uint32_t* __restrict__ colmax, const ScanItem* __restrict__ items, const SegDesc* __restrict__ segs,
                         int n_items, int n_gran, int max_len, int tasks_per_seg, int* __restrict__ stats_max)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_items) return;
    const ScanItem it = items[warp];
    const int n = segs[it.seg].len;
    const uint32_t* cm = colmax + (size_t)warp * n_gran * max_len;
    uint32_t v = 0;
    for (int k = 0; k < n_gran; ++k)
        for (int j = lane; j < n; j += 32) v = __vmaxs2(v, cm[(size_t)k * max_len + j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) v = __vmaxs2(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) {
        const PairDef pd = c_pairs[it.pair];
        stats_max[it.seg * tasks_per_seg + pd.task[0]] = lo16(v);
        stats_max[it.seg * tasks_per_seg + pd.task[1]] = hi16(v);
    }
}

}  // namespace ltg
