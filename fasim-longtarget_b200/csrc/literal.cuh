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
};

struct LiteralArgs {
    const LiteralJob* jobs; int n_jobs;
    const int* n_jobs_dev;                              // used when n_jobs < 0: the count was produced on the device
    const uint8_t* codes; const SegDesc* segs; const uint8_t* rna_ssw;
    unsigned char* work; long long work_per_slot;       // 4 * L * 16 bytes per half-warp slot
    uint16_t* lit_colmax; int max_len;                  // scan jobs: row `job index` receives the literal column maxima
    int* task_litrow;                                   // scan jobs: [task] -> that row
    WinState w;
};

__device__ inline int sat8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

__global__ void __launch_bounds__(128) k_literal(const LiteralArgs a)
{
    const int lane = threadIdx.x & 31, s = lane & 15, halfw = lane >> 4;
    const unsigned hmask = halfw ? 0xffff0000u : 0x0000ffffu;
    const int slot = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2 + halfw;
    const int nslots = ((gridDim.x * blockDim.x) >> 5) * 2;
    const int bias = 4;
    const int n_jobs = a.n_jobs >= 0 ? a.n_jobs : *a.n_jobs_dev;
    for (int jb = slot; jb < n_jobs; jb += nslots) {
        const LiteralJob J = a.jobs[jb];
        const SegDesc sd = a.segs[J.seg];
        const TaskDef td = c_tasks[J.tdef];
        const int L = (J.read_len + 15) / 16;
        unsigned char* Hs = a.work + (size_t)slot * a.work_per_slot;
        unsigned char* Hl = Hs + (size_t)L * 16;
        unsigned char* Ev = Hl + (size_t)L * 16;
        unsigned char* Hm = Ev + (size_t)L * 16;
        for (int t = 0; t < L; ++t) { Hs[t * 16 + s] = 0; Hl[t * 16 + s] = 0; Ev[t * 16 + s] = 0; Hm[t * 16 + s] = 0; }
        uint16_t* cmrow = nullptr;
        if (J.kind == 0) {
            cmrow = a.lit_colmax + (size_t)jb * a.max_len;
            for (int j = s; j < J.ref_len; j += 16) cmrow[j] = 0;
            if (s == 0) a.task_litrow[J.task] = jb;
        }
        __syncwarp(hmask);
        int vMaxScore = 0, vMaxMark = 0, maxv = 0, end_ref = -1;
        bool overflow = false;
        const int begin = J.ref_dir ? J.ref_len - 1 : 0, end = J.ref_dir ? -1 : J.ref_len, step = J.ref_dir ? -1 : 1;
        for (int i = begin; i != end; i += step) {
            const int q = J.ref_start + i;
            const int c = td.img[a.codes[sd.start + (td.reversed ? (sd.len - 1 - q) : q)]];
            int vF = 0, vMaxCol = 0;
            int vH = __shfl_up_sync(hmask, (int)Hs[(L - 1) * 16 + s], 1, 16);
            if (s == 0) vH = 0;
            { unsigned char* tmp = Hl; Hl = Hs; Hs = tmp; }
            for (int t = 0; t < L; ++t) {
                const int row = s * L + t;
                int p = bias;
                if (row < J.read_len) { const int r = a.rna_ssw[J.read_start + J.read_dir * row]; p = ((r == c && c < 4) ? kMatch : kMismatch) + bias; }
                int h = sat8(sat8(vH + p) - bias);
                int e = Ev[t * 16 + s];
                h = max(h, e); h = max(h, vF);
                vMaxCol = max(vMaxCol, h);
                Hs[t * 16 + s] = (unsigned char)h;
                const int open = sat8(h - kGapOpen);
                e = max(sat8(e - kGapExt), open);
                Ev[t * 16 + s] = (unsigned char)e;
                vF = max(sat8(vF - kGapExt), open);
                vH = Hl[t * 16 + s];
            }
            bool done = false;
            for (int k = 0; k < 16 && !done; ++k) {
                vF = __shfl_up_sync(hmask, vF, 1, 16);
                if (s == 0) vF = 0;
                for (int t = 0; t < L; ++t) {
                    int h = Hs[t * 16 + s];
                    h = max(h, vF);
                    vMaxCol = max(vMaxCol, h);
                    Hs[t * 16 + s] = (unsigned char)h;
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
                    for (int t = 0; t < L; ++t) Hm[t * 16 + s] = Hs[t * 16 + s];
                }
            }
            int cm = vMaxCol;
#pragma unroll
            for (int o = 8; o; o >>= 1) cm = max(cm, __shfl_xor_sync(hmask, cm, o, 16));
            if (cmrow && s == 0) cmrow[i] = (uint16_t)cm;
            if (cm == J.terminate) break;
        }
        if (J.kind == 0) continue;
        int end_read = J.read_len - 1;
        for (int t = 0; t < L; ++t) if (Hm[t * 16 + s] == maxv) end_read = min(end_read, t + s * L);
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

// collect the windows whose exact score reaches the Q4 guard
__global__ void k_lit_collect(const WinState w, int reverse, LiteralJob* jobs, int* count, int* count_total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_peaks) return;
    LiteralJob J;
    const int task = w.pk_task[i];
    J.task = task; J.seg = task / w.tasks_per_seg; J.tdef = task % w.tasks_per_seg; J.peak = i;
    if (!reverse) {
        if (w.w_done[i]) return;
        const int4 v = w.res[i];
        if (v.x < kQ4Guard || v.x >= kOverflowU8) return;
        const int cut = w.w_len[i];
        J.kind = 1; J.ref_start = w.pk_pos[i] - cut + 1; J.ref_len = cut; J.ref_dir = 0;
        J.read_start = 0; J.read_len = w.m; J.read_dir = 1; J.terminate = 255;
    } else {
        const int sw = w.fin_sw[i];
        if (sw < kQ4Guard || sw >= kOverflowU8) return;
        J.kind = 2; J.ref_start = w.pk_pos[i] - w.fin_cut[i] + 1; J.ref_len = w.fin_re[i] + 1; J.ref_dir = 1;
        J.read_start = w.fin_qe[i]; J.read_len = w.fin_qe[i] + 1; J.read_dir = -1; J.terminate = sw & 0xff;
    }
    jobs[atomicAdd(count, 1)] = J;
    atomicAdd(count_total, 1);
}

}  // namespace ltg
