// Literal emulation of the reference's 8-bit striped kernels (sswNew.cpp:255-464 sw_sse2_byte_once and
// :476-672 sw_sse2_byte) — the slow, exact-to-the-quirk path.
//
// The fast kernels compute exact Smith-Waterman.  The reference deviates from exact SW in one narrow,
// constructible case: its lazy-F loop exits on a *signed* byte compare, which can end
// the loop early once an F value >= 132 crosses a stripe boundary.  That needs H >= 148 somewhere, so every
// task / window whose exact maximum is >= 148 is recomputed here with the reference's exact data layout:
// 16 byte lanes striped over the read (row(s,t) = s*L + t), unsigned saturating arithmetic with bias 4,
// the lazy-F loop with the signed compare, the overflow break and the terminate test.  One half-warp
// (16 threads = 16 SSE lanes) per job; the H/E columns live in an L2-resident workspace.
#pragma once
#include <vector>

#include "common.cuh"
#include "scan.cuh"
#include "window.cuh"

namespace ltg {

struct LiteralJob {
    int kind;          // 0 scan column maxima, 1 window forward, 2 window reverse
    int task;          // batch task id (seg * T + task index)
    int seg, tdef;
    int ref_start, ref_len, ref_dir;        // columns: translated-segment indices [ref_start, ref_start+ref_len); dir 1 = high to low
    int read_start, read_len, read_dir;     // rows: rna[read_start + read_dir * k], k in [0, read_len)
    int terminate;
    int peak;          // window jobs: peak index
    long long seg_start; int seg_len;       // explicit segment geometry (side-stream scan jobs: LiteralArgs::segs == nullptr)
};

struct LiteralArgs {
    const LiteralJob* jobs; int n_jobs;
    const int* n_jobs_dev;                              // used when n_jobs < 0: the count was produced on the device
    const uint8_t* codes; const SegDesc* segs; const uint8_t* rna_ssw;
    unsigned char* work; long long work_per_slot;       // global workspace per half-warp slot (used when use_smem == 0)
    int use_smem, slots_per_block, pitch;               // shared-memory workspace: slots per block, elements per (array, lane)
    uint16_t* lit_colmax; int max_len;                  // scan jobs: row `row_base + job index` receives the literal column maxima
    int row_base;
    int* task_litrow;                                   // scan jobs: [task] -> that row (nullptr: the host keeps the mapping)
    WinState w;
};

__device__ inline int sat8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// elements per (array, lane): the stripe length rounded up to 4 * odd, so that a lane's arrays are 8-byte aligned and the
// 16 lanes of a half-warp, which walk their stripes in lock step with 64-bit accesses, cover all 32 banks exactly once
__host__ __device__ inline int literal_pitch(int read_len) { int q = (((read_len + 15) / 16) + 3) / 4; if ((q & 1) == 0) ++q; return 4 * q; }
constexpr int kLitArrays = 9;       // H stored, H previous column, E, H at the best column, 5 profile rows

// One half-warp (16 threads = 16 SSE lanes) per job.  Workspace per job: uint16 arrays [lane][t] — H of the column
// being stored, H of the previous column, E, H at the best column, and the profile (score + bias) of the lane's stripe
// for each of the 5 base codes — in shared memory (global memory only for lncRNAs beyond ~10 knt).  All values are the
// reference's unsigned bytes, held in 16-bit lanes so that the native packed 16x2 instructions apply.
//
// The stripe loop of a column is sequential in the reference (vF runs down the stripe).  Here only
//     F' = max(F - 4, max(hbase - 16, 0)),     hbase = max(subs(adds(Hdiag, profile), bias), E)
// runs step by step (one VIADDMNMX each; it equals the reference's F' = max(subs(F,4), subs(max(hbase,F),16)) because
// F - 16 < F - 4); hbase before it and H = max(hbase, F), E' = max(subs(E,4), subs(H,16)) after it are computed four
// steps at a time.  The lazy-F loop with its signed-byte exit test (Q4) is literal.
__global__ void __launch_bounds__(128) k_literal(const LiteralArgs a)
{
    extern __shared__ uint32_t lit_smem[];
    const int lane = threadIdx.x & 31, s = lane & 15, halfw = lane >> 4;
    const unsigned hmask = halfw ? 0xffff0000u : 0x0000ffffu;
    const int slot_in_block = threadIdx.x >> 4, spb = a.use_smem ? a.slots_per_block : (int)(blockDim.x >> 4);
    if (slot_in_block >= spb) return;
    const int slot = blockIdx.x * spb + slot_in_block;
    const int nslots = gridDim.x * spb;
    const int bias = 4;
    const uint32_t kFF = 0x00FF00FFu, kM4 = 0xFFFCFFFCu, kM16 = 0xFFF0FFF0u;
    const int n_jobs = a.n_jobs >= 0 ? a.n_jobs : *a.n_jobs_dev;
    for (int jb = slot; jb < n_jobs; jb += nslots) {
        const LiteralJob J = a.jobs[jb];
        SegDesc sd;
        if (a.segs) sd = a.segs[J.seg]; else { sd.start = J.seg_start; sd.len = J.seg_len; sd.flags = 0; }
        const TaskDef td = c_tasks[J.tdef];
        const int L = (J.read_len + 15) / 16;
        const int P = a.pitch;                                   // >= literal_pitch(J.read_len)
        uint16_t* base = a.use_smem ? reinterpret_cast<uint16_t*>(lit_smem) + (size_t)slot_in_block * (16 * kLitArrays * (size_t)P)
                                    : reinterpret_cast<uint16_t*>(a.work + (size_t)slot * a.work_per_slot);
        uint16_t* hs = base + (size_t)s * P;                     // this lane's stripes
        uint16_t* hl = hs + 16 * (size_t)P;
        uint16_t* ev = hl + 16 * (size_t)P;
        uint16_t* hm = ev + 16 * (size_t)P;
        uint16_t* prof = hm + 16 * (size_t)P;                    // prof + x * 16 * P: profile row of base code x
        for (int t = 0; t < P; ++t) {
            hs[t] = 0; hl[t] = 0; ev[t] = 0; hm[t] = 0;
            const int row = s * L + t;
            const bool real = t < L && row < J.read_len;
            const int r = real ? a.rna_ssw[J.read_start + J.read_dir * row] : -1;
#pragma unroll
            for (int x = 0; x < 5; ++x)     // qP_byte (sswNew.cpp:176-200): score + bias, pad rows = bias
                prof[(size_t)x * 16 * P + t] = (uint16_t)(real ? (((r == x && x < 4) ? kMatch : kMismatch) + bias) : bias);
        }
        uint16_t* cmrow = nullptr;
        if (J.kind == 0) {
            cmrow = a.lit_colmax + (size_t)(a.row_base + jb) * a.max_len;
            for (int j = s; j < J.ref_len; j += 16) cmrow[j] = 0;
            if (s == 0 && a.task_litrow) a.task_litrow[J.task] = a.row_base + jb;
        }
        __syncwarp(hmask);
        int vMaxScore = 0, vMaxMark = 0, maxv = 0, end_ref = -1;
        bool overflow = false;
        const int begin = J.ref_dir ? J.ref_len - 1 : 0, end = J.ref_dir ? -1 : J.ref_len, step = J.ref_dir ? -1 : 1;
        const int L4 = L & ~3;
        for (int i = begin; i != end; i += step) {
            const int q = J.ref_start + i;
            const int c = td.img[a.codes[sd.start + (td.reversed ? (sd.len - 1 - q) : q)]];
            const uint16_t* pc = prof + (size_t)c * 16 * P;
            int vF = 0;
            uint32_t vH = (uint32_t)__shfl_up_sync(hmask, (int)hs[L - 1], 1, 16);
            if (s == 0) vH = 0;
            { uint16_t* tmp = hl; hl = hs; hs = tmp; }
            uint32_t vmax2 = 0;
            int t0 = 0;
            for (; t0 < L4; t0 += 4) {
                const uint2 hlw = *reinterpret_cast<const uint2*>(hl + t0);
                const uint2 ew = *reinterpret_cast<const uint2*>(ev + t0);
                const uint2 pw = *reinterpret_cast<const uint2*>(pc + t0);
                // H diagonal of steps t0..t0+3 = H of the previous column at steps t0-1..t0+2
                const uint32_t vHa = __byte_perm(vH, hlw.x, 0x5410), vHb = __byte_perm(hlw.x, hlw.y, 0x5432);
                vH = hlw.y >> 16;
                const uint32_t hba = __vmaxs2(__viaddmax_s16x2(__vmins2(__vadd2(vHa, pw.x), kFF), kM4, 0u), ew.x);
                const uint32_t hbb = __vmaxs2(__viaddmax_s16x2(__vmins2(__vadd2(vHb, pw.y), kFF), kM4, 0u), ew.y);
                const uint32_t ga = __viaddmax_s16x2(hba, kM16, 0u), gb = __viaddmax_s16x2(hbb, kM16, 0u);
                // the sequential part: F entering each of the four steps
                const int f0 = vF;
                const int f1 = __viaddmax_s32(f0, -kGapExt, (int)(ga & 0xffffu));
                const int f2 = __viaddmax_s32(f1, -kGapExt, (int)(ga >> 16));
                const int f3 = __viaddmax_s32(f2, -kGapExt, (int)(gb & 0xffffu));
                vF = __viaddmax_s32(f3, -kGapExt, (int)(gb >> 16));
                const uint32_t ha = __vmaxs2(hba, __byte_perm((uint32_t)f0, (uint32_t)f1, 0x5410));
                const uint32_t hb = __vmaxs2(hbb, __byte_perm((uint32_t)f2, (uint32_t)f3, 0x5410));
                vmax2 = __vimax3_s16x2(vmax2, ha, hb);
                *reinterpret_cast<uint2*>(hs + t0) = make_uint2(ha, hb);
                // E' = max(subs(E, 4), subs(H, 16));  subs(H,16) >= 0 makes the floor of the first term redundant
                *reinterpret_cast<uint2*>(ev + t0) = make_uint2(__viaddmax_s16x2(ew.x, kM4, __viaddmax_s16x2(ha, kM16, 0u)),
                                                                 __viaddmax_s16x2(ew.y, kM4, __viaddmax_s16x2(hb, kM16, 0u)));
            }
            int vMaxCol = max((int)(vmax2 & 0xffffu), (int)(vmax2 >> 16));
            for (int t = t0; t < L; ++t) {                       // the last (L mod 4) steps, step by step
                int h = sat8(sat8((int)vH + (int)pc[t]) - bias);
                int e = ev[t];
                h = max(h, e); h = max(h, vF);
                vMaxCol = max(vMaxCol, h);
                hs[t] = (uint16_t)h;
                const int open = sat8(h - kGapOpen);
                ev[t] = (uint16_t)max(sat8(e - kGapExt), open);
                vF = max(sat8(vF - kGapExt), open);
                vH = hl[t];
            }
            bool done = false;
            for (int k = 0; k < 16 && !done; ++k) {
                vF = __shfl_up_sync(hmask, vF, 1, 16);
                if (s == 0) vF = 0;
                for (int t = 0; t < L; ++t) {
                    int h = hs[t];
                    h = max(h, vF);
                    vMaxCol = max(vMaxCol, h);
                    hs[t] = (uint16_t)h;
                    const int open = sat8(h - kGapOpen);
                    vF = sat8(vF - kGapExt);
                    const bool gt = (int)(int8_t)vF > (int)(int8_t)open;        // signed byte compare (Q4)
                    if (!__any_sync(hmask, gt)) { done = true; break; }
                }
            }
            vMaxScore = max(vMaxScore, vMaxCol);
            const bool changed = __any_sync(hmask, vMaxScore != vMaxMark);
            if (changed) {
                vMaxMark = vMaxScore;
                int temp = vMaxScore;
#pragma unroll
                for (int o = 8; o; o >>= 1) temp = max(temp, __shfl_xor_sync(hmask, temp, o, 16));
                if (temp > maxv) {
                    maxv = temp;
                    if (maxv + bias >= 255) { overflow = true; break; }
                    end_ref = i;
                    for (int t = 0; t < P; t += 4) *reinterpret_cast<uint2*>(hm + t) = *reinterpret_cast<const uint2*>(hs + t);
                }
            }
            int cm = vMaxCol;
#pragma unroll
            for (int o = 8; o; o >>= 1) cm = max(cm, __shfl_xor_sync(hmask, cm, o, 16));
            if (cmrow && s == 0) cmrow[i] = (uint16_t)cm;
            if (cm == J.terminate) break;
        }
        __syncwarp(hmask);
        if (J.kind == 0) continue;
        int end_read = J.read_len - 1;
        for (int t = 0; t < L; ++t) if (hm[t] == maxv) end_read = min(end_read, t + s * L);
#pragma unroll
        for (int o = 8; o; o >>= 1) end_read = min(end_read, __shfl_xor_sync(hmask, end_read, o, 16));
        if (s != 0) continue;
        // overflow (score marker 255) sends the reference to its exact 16-bit kernel: keep the exact fast-path result
        if (overflow) continue;
        const int pk = J.peak;
        if (J.kind == 1) {
            a.w.res[pk] = make_int4(maxv, end_ref, end_read, 1);
        } else {
            const int fwd = a.w.fin_sw[pk];
            a.w.fin_sw[pk] = end_ref < 0 ? 0 : min(maxv, fwd);
            a.w.fin_rb[pk] = end_ref;
            a.w.fin_qb[pk] = a.w.fin_qe[pk] - end_read;
        }
    }
}

// ---- column-parallel variant ---------------------------------------------------------------------------------------
// Same emulation, one CTA of 16 x CH threads per job: thread (lane, chunk) owns a chunk of the lane's stripe.  What is
// sequential in the reference's stripe loop is only the F chain  F' = max(subs(F,4), subs(H,16)), and that is max-plus
// linear: each chunk runs it from F = 0, the chunk ends are folded along the stripe, and a carried-in F that is still
// positive is applied to the first offsets of a chunk afterwards (H = max(H, F_in - 4k), E' = max(E', subs(H,16)): the
// values the sequential loop produces).  The lazy-F loop with the signed compare, the maximum bookkeeping, the overflow
// break and the terminate test are run literally by the 16 threads of chunk 0.  Workspace: three BYTE arrays
// [lane][offset] in shared memory — H (updated in place: the diagonal input of an offset is kept in a register, the one
// of a chunk's first offset is read before anyone writes), E, and the one-hot read codes from which the profile value of
// a row is derived (match 9, mismatch 0, pad row 4 = the reference's score + bias).  The reference's copy of the best
// column (pvHmax) is only used to find the smallest row that holds the maximum: that row is taken when the column
// becomes the best one.
constexpr int kLitColMaxChunks = 32;
constexpr int kLitColArrays = 3;

__global__ void __launch_bounds__(512) k_literal_col(const LiteralArgs a)
{
    extern __shared__ uint32_t lit_smem[];
    __shared__ uint8_t s_agg[16 * kLitColMaxChunks];  // [lane][chunk] F at the end of the chunk when it starts from 0
    __shared__ uint8_t s_cmx[16 * kLitColMaxChunks];  // [lane][chunk] maximum of the chunk's H values in this column
    __shared__ int s_ctl[4];                          // 0 new best column, 1 leave the column loop, 2 end_read, 3 max
    const int tid = threadIdx.x, lam = tid & 15, ch = tid >> 4, CH = blockDim.x >> 4;
    const unsigned hmask = 0x0000ffffu;
    const int bias = 4;
    const uint32_t kFF = 0x00FF00FFu, kM4 = 0xFFFCFFFCu, kM16 = 0xFFF0FFF0u;
    const int n_jobs = a.n_jobs >= 0 ? a.n_jobs : *a.n_jobs_dev;
    for (int jb = blockIdx.x; jb < n_jobs; jb += gridDim.x) {
        __syncthreads();                                         // the previous job's readers are done with the workspace
        const LiteralJob J = a.jobs[jb];
        SegDesc sd;
        if (a.segs) sd = a.segs[J.seg]; else { sd.start = J.seg_start; sd.len = J.seg_len; sd.flags = 0; }
        const TaskDef td = c_tasks[J.tdef];
        const int L = (J.read_len + 15) / 16;
        const int P = a.pitch;                                   // bytes per (array, lane), 4 * odd
        uint8_t* hs = reinterpret_cast<uint8_t*>(lit_smem) + (size_t)lam * P;
        uint8_t* ev = hs + 16 * (size_t)P;
        uint8_t* oh = ev + 16 * (size_t)P;
        const int Lc = 4 * ((L + 4 * CH - 1) / (4 * CH));        // chunk length, a multiple of 4
        const int tb = min(L, ch * Lc), te = min(L, tb + Lc);
        const int tv = tb + ((te - tb) & ~3);                    // end of the part done four offsets at a time
        const int klast = (L - 1) / Lc;                          // last chunk that holds an offset (the ones after it are empty)
        const int kfold = 60 / Lc + 2;                           // full chunks that span 60 offsets, plus one, plus the short last one
        for (int t = tb; t < te; ++t) {
            hs[t] = 0; ev[t] = 0;
            const int row = lam * L + t;
            const bool real = row < J.read_len;
            const int r = real ? a.rna_ssw[J.read_start + J.read_dir * row] : 5;
            oh[t] = (uint8_t)(r < 4 ? (1 << r) : (r == 5 ? 0x10 : 0));
        }
        uint16_t* cmrow = nullptr;
        if (J.kind == 0) {
            cmrow = a.lit_colmax + (size_t)(a.row_base + jb) * a.max_len;
            for (int j = tid; j < J.ref_len; j += blockDim.x) cmrow[j] = 0;
            if (tid == 0 && a.task_litrow) a.task_litrow[J.task] = a.row_base + jb;
        }
        // (no best column yet: like the reference's zero-filled pvHmax, row 0 "holds" a maximum of 0)
        if (tid == 0) { s_ctl[0] = 0; s_ctl[1] = 0; s_ctl[2] = 0; s_ctl[3] = 0; }
        __syncthreads();
        int vMaxScore = 0, vMaxMark = 0, maxv = 0, end_ref = -1;          // live in the threads of chunk 0
        bool overflow = false;
        const int begin = J.ref_dir ? J.ref_len - 1 : 0, end = J.ref_dir ? -1 : J.ref_len, step = J.ref_dir ? -1 : 1;
        for (int i = begin; i != end; i += step) {
            const int q = J.ref_start + i;
            const int c = td.img[a.codes[sd.start + (td.reversed ? (sd.len - 1 - q) : q)]];
            const int csh = c < 4 ? c : 7;                       // bit 7 of a one-hot code is never set: no match
            // H diagonal of the chunk's first offset: the previous column's H one offset up — for a stripe start the last
            // offset of the previous lane (the reference's byte shift) — read before this column overwrites anything
            uint32_t vH = 0;
            if (tb > 0) vH = hs[tb - 1];
            else if (lam > 0) vH = (hs - P)[L - 1];
            __syncthreads();
            // ---- sweep 1: the chunk from F = 0
            int vF = 0;
            uint32_t vmax2 = 0;
            // Shared memory is addressed with explicit 32-bit shared addresses; the next four offsets are loaded before the
            // current ones are stored (the arrays carry 16 bytes of slack, so the last, unused prefetch stays inside them).
            {
                const uint32_t sa_h = (uint32_t)__cvta_generic_to_shared(hs), sa_e = (uint32_t)__cvta_generic_to_shared(ev);
                const uint32_t sa_o = (uint32_t)__cvta_generic_to_shared(oh);
                uint32_t nh4, ne4, no4;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(nh4) : "r"(sa_h + tb));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ne4) : "r"(sa_e + tb));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(no4) : "r"(sa_o + tb));
                for (int t0 = tb; t0 < tv; t0 += 4) {
                    const uint32_t h4 = nh4, e4 = ne4, o4 = no4;
                    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(nh4) : "r"(sa_h + t0));
                    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(ne4) : "r"(sa_e + t0));
                    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(no4) : "r"(sa_o + t0));
                    const uint32_t p4 = ((o4 >> csh) & 0x01010101u) * 9u + ((o4 >> 4) & 0x01010101u) * 4u;
                    // H diagonal of offsets t0..t0+3 = H of the previous column at offsets t0-1..t0+2
                    const uint32_t dlo = __byte_perm(vH, h4, 0x1410), dhi = __byte_perm(vH, h4, 0x1615);
                    vH = h4 >> 24;
                    const uint32_t elo = __byte_perm(e4, 0u, 0x4140), ehi = __byte_perm(e4, 0u, 0x4342);
                    const uint32_t plo = __byte_perm(p4, 0u, 0x4140), phi = __byte_perm(p4, 0u, 0x4342);
                    // hbase = max(subs(adds(Hdiag, profile), bias), E): adds saturates at 255, subs at 0 (E >= 0 is the floor)
                    const uint32_t hba = __viaddmax_s16x2_relu(__viaddmin_s16x2(dlo, plo, kFF), kM4, elo);
                    const uint32_t hbb = __viaddmax_s16x2_relu(__viaddmin_s16x2(dhi, phi, kFF), kM4, ehi);
                    const uint32_t ga = __viaddmax_s16x2_relu(hba, kM16, kM16), gb = __viaddmax_s16x2_relu(hbb, kM16, kM16);   // subs(hbase, 16)
                    const int f0 = vF;
                    const int f1 = __viaddmax_s32(f0, -kGapExt, (int)(ga & 0xffffu));
                    const int f2 = __viaddmax_s32(f1, -kGapExt, (int)(ga >> 16));
                    const int f3 = __viaddmax_s32(f2, -kGapExt, (int)(gb & 0xffffu));
                    vF = __viaddmax_s32(f3, -kGapExt, (int)(gb >> 16));
                    const uint32_t ha = __vmaxs2(hba, __byte_perm((uint32_t)f0, (uint32_t)f1, 0x5410));
                    const uint32_t hb = __vmaxs2(hbb, __byte_perm((uint32_t)f2, (uint32_t)f3, 0x5410));
                    vmax2 = __vimax3_s16x2(vmax2, ha, hb);
                    asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa_h + t0), "r"(__byte_perm(ha, hb, 0x6420)) : "memory");
                    // E' = max(subs(E, 4), subs(H, 16))
                    const uint32_t ea = __viaddmax_s16x2_relu(elo, kM4, __vadd2(ha, kM16));
                    const uint32_t eb = __viaddmax_s16x2_relu(ehi, kM4, __vadd2(hb, kM16));
                    asm volatile("st.shared.u32 [%0], %1;" :: "r"(sa_e + t0), "r"(__byte_perm(ea, eb, 0x6420)) : "memory");
                }
            }
            int lmax = max((int)(vmax2 & 0xffffu), (int)(vmax2 >> 16));
            for (int t = tv; t < te; ++t) {                      // the last (length mod 4) offsets of the stripe
                const int o = oh[t];
                const int pv = ((o >> csh) & 1) * 9 + ((o >> 4) & 1) * 4;
                const int hold = hs[t];
                int h = sat8(sat8((int)vH + pv) - bias);
                const int e = ev[t];
                h = max(h, e); h = max(h, vF);
                lmax = max(lmax, h);
                hs[t] = (uint8_t)h;
                const int open = sat8(h - kGapOpen);
                ev[t] = (uint8_t)max(sat8(e - kGapExt), open);
                vF = max(sat8(vF - kGapExt), open);
                vH = hold;
            }
            s_agg[lam * kLitColMaxChunks + ch] = (uint8_t)vF;
            __syncthreads();
            // ---- F carried into this chunk from the chunks before it, applied to the chunk's first offsets
            // (an F is at most 239 and loses 4 per offset: only the chunks within 60 offsets above this one can contribute)
            int fin = 0;
            for (int k = max(0, ch - kfold); k < ch; ++k) {
                const int len = max(0, min(L, (k + 1) * Lc) - min(L, k * Lc));
                fin = max((int)s_agg[lam * kLitColMaxChunks + k], max(fin - kGapExt * len, 0));
            }
            for (int t = tb, fv = fin; t < te && fv > 0; ++t, fv -= kGapExt) {
                if (fv > (int)hs[t]) {
                    hs[t] = (uint8_t)fv;
                    ev[t] = (uint8_t)max((int)ev[t], sat8(fv - kGapOpen));
                    lmax = max(lmax, fv);
                }
            }
            s_cmx[lam * kLitColMaxChunks + ch] = (uint8_t)lmax;
            __syncthreads();
            // ---- lazy-F loop and the per-column bookkeeping: the 16 threads of chunk 0, literally
            if (ch == 0) {
                int vMaxCol = 0;
                for (int k = 0; k < CH; ++k) vMaxCol = max(vMaxCol, (int)s_cmx[lam * kLitColMaxChunks + k]);
                vF = 0;                                          // F after the stripe's last offset
                for (int k = max(0, klast - kfold); k <= klast; ++k) {
                    const int len = max(0, min(L, (k + 1) * Lc) - min(L, k * Lc));
                    vF = max((int)s_agg[lam * kLitColMaxChunks + k], max(vF - kGapExt * len, 0));
                }
                bool done = false;
                for (int k = 0; k < 16 && !done; ++k) {
                    vF = __shfl_up_sync(hmask, vF, 1, 16);
                    if (lam == 0) vF = 0;
                    for (int t = 0; t < L; ++t) {
                        int h = hs[t];
                        h = max(h, vF);
                        vMaxCol = max(vMaxCol, h);
                        hs[t] = (uint8_t)h;
                        const int open = sat8(h - kGapOpen);
                        vF = sat8(vF - kGapExt);
                        const bool gt = (int)(int8_t)vF > (int)(int8_t)open;        // signed byte compare (Q4)
                        if (!__any_sync(hmask, gt)) { done = true; break; }
                    }
                }
                bool best = false, leave = false;
                vMaxScore = max(vMaxScore, vMaxCol);
                const bool changed = __any_sync(hmask, vMaxScore != vMaxMark);
                if (changed) {
                    vMaxMark = vMaxScore;
                    int temp = vMaxScore;
#pragma unroll
                    for (int o = 8; o; o >>= 1) temp = max(temp, __shfl_xor_sync(hmask, temp, o, 16));
                    if (temp > maxv) {
                        maxv = temp;
                        if (maxv + bias >= 255) { overflow = true; leave = true; }
                        else { end_ref = i; best = true; }
                    }
                }
                if (!leave) {
                    int cm = vMaxCol;
#pragma unroll
                    for (int o = 8; o; o >>= 1) cm = max(cm, __shfl_xor_sync(hmask, cm, o, 16));
                    if (cmrow && lam == 0) cmrow[i] = (uint16_t)cm;
                    if (cm == J.terminate) leave = true;
                }
                if (lam == 0) {
                    s_ctl[0] = best ? 1 : 0; s_ctl[1] = leave ? 1 : 0;
                    if (best) { s_ctl[2] = J.read_len - 1; s_ctl[3] = maxv; }
                }
            }
            __syncthreads();
            // a new best column: the smallest row that holds the maximum (what the reference reads off its copy of the column)
            if (s_ctl[0] && J.kind != 0) {
                const int mv = s_ctl[3];
                int er = J.read_len;
                for (int t = tb; t < te; ++t) if ((int)hs[t] == mv) er = min(er, t + lam * L);
                if (er < J.read_len) atomicMin(&s_ctl[2], er);
            }
            if (s_ctl[1]) break;
        }
        __syncthreads();
        if (J.kind == 0 || tid != 0) continue;
        // overflow (score marker 255) sends the reference to its exact 16-bit kernel: keep the exact fast-path result
        if (overflow) continue;
        const int end_read = s_ctl[2];
        const int pk = J.peak;
        if (J.kind == 1) {
            a.w.res[pk] = make_int4(maxv, end_ref, end_read, 1);
        } else {
            const int fwd = a.w.fin_sw[pk];
            a.w.fin_sw[pk] = end_ref < 0 ? 0 : min(maxv, fwd);
            a.w.fin_rb[pk] = end_ref;
            a.w.fin_qb[pk] = a.w.fin_qe[pk] - end_read;
        }
    }
}

// collect the windows whose exact score reaches the Q4 guard
__global__ void k_lit_collect(const WinState w, int reverse, int round, LiteralJob* jobs, int* count, int* count_total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks) return;
    LiteralJob J;
    const int task = w.pk_task[i];
    J.task = task; J.seg = task / w.tasks_per_seg; J.tdef = task % w.tasks_per_seg; J.peak = i;
    if (!reverse) {
        if (w.w_done[i] || w.w_next[i] != round) return;       // only the windows swept in this round
        const int4 v = w.res[i];
        // (no upper bound: with the Q4 quirk the reference's byte kernel can stay below 251 where exact SW reaches it and then
        //  keeps its own, lower result; the literal kernels detect a real overflow themselves and leave the exact result alone)
        if (v.x < kQ4Guard) return;
        if (w.w_q4 && !(w.w_q4[i] & 1)) return;                // no F >= 132 entered a stripe start in any sweep of this window: exact = reference
        const int cut = w.w_len[i];
        J.kind = 1; J.ref_start = w.w_ws[i]; J.ref_len = cut; J.ref_dir = 0;
        J.read_start = 0; J.read_len = w.m; J.read_dir = 1; J.terminate = 255;
    } else {
        const int sw = w.fin_sw[i];
        if (sw < kQ4Guard || sw >= kOverflowU8) return;
        if (w.w_q4 && !(w.w_q4[i] & 2)) return;
        J.kind = 2; J.ref_start = w.fin_ws[i]; J.ref_len = w.fin_re[i] + 1; J.ref_dir = 1;
        J.read_start = w.fin_qe[i]; J.read_len = w.fin_qe[i] + 1; J.read_dir = -1; J.terminate = sw & 0xff;
    }
    jobs[atomicAdd(count, 1)] = J;
    atomicAdd(count_total, 1);
}

}  // namespace ltg
