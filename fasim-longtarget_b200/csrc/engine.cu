// libfasim_b200.so — context, device memory model, pipeline orchestration and the C ABI of
// include/fasim_b200.h.  One context = one GPU = one stream.  No CPU compute fallback: every DP cell is
// computed by the kernels in scan.cuh / window.cuh / literal.cuh.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/fasim_b200.h"
#include "common.cuh"
#include "scan.cuh"
#include "window.cuh"
#include "literal.cuh"
#include "../host/rules_table.hpp"
#include "../host/triplex_host.hpp"

namespace ltg {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return LTG_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        LTG_CUDA_CHECK(cudaMalloc(&p, want));
        cap = want;
        return LTG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kScanR = 16;          // RNA rows per lane in the scan kernel (strip = 512 rows)
constexpr int kScanWarps = 4;       // warps per CTA
constexpr int kScanCtasPerSm = 3;
constexpr int kBatchSegments = 2048;

}  // namespace ltg

using namespace ltg;

struct ltg_context {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ltg_params params;
    // task tables (depend on params.rule / params.strand)
    std::vector<TaskDef> tasks;
    std::vector<PairDef> pairs;
    bool tables_dirty = true;
    // query
    std::string rna_name, rna;
    bool rna_plain = true;          // only ACGT (any case): the SSW-side and Farrar-side scorings coincide
    int m = 0, n_strips = 0;
    bool profiles_dirty = true;
    DevBuf d_rna_raw, d_rna_ssw, d_rna_stats, d_prof_ssw, d_prof_stats, d_cut;
    // record / batch buffers
    DevBuf d_dna, d_codes, d_segs, d_items, d_colmax, d_bnd, d_counters;
    DevBuf d_task_max, d_task_thr, d_task_npk, d_task_flags, d_stats_max;
    DevBuf d_pk_task, d_pk_pos, d_pk_score;
    DevBuf d_w[12], d_cls_list, d_res;
    DevBuf d_al_status, d_al_nt, d_al_match, d_al_stroff, d_strpool, d_scratch, d_scratch_big;
    DevBuf d_lit_colmax, d_lit_work, d_lit_jobs;
    int64_t launches = 0;
};

namespace {

int upload_tables(ltg_context* c)
{
    if (!ltg_host::enumerate_tasks(c->params, c->tasks)) {
        set_error("invalid rule/strand selection (rule=%d strand=%d)", c->params.rule, c->params.strand);
        return LTG_ERR_ARG;
    }
    // pair tasks that read the segment in the same direction (identical control flow, independent cells)
    c->pairs.clear();
    for (int dir = 0; dir < 2; ++dir) {
        std::vector<int> idx;
        for (size_t t = 0; t < c->tasks.size(); ++t) if (c->tasks[t].reversed == dir) idx.push_back((int)t);
        for (size_t k = 0; k < idx.size(); k += 2) {
            PairDef p;
            p.task[0] = (int16_t)idx[k];
            p.task[1] = (int16_t)(k + 1 < idx.size() ? idx[k + 1] : idx[k]);
            p.reversed = (int16_t)dir; p.pad_ = 0;
            c->pairs.push_back(p);
        }
    }
    LTG_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_tasks, c->tasks.data(), sizeof(TaskDef) * c->tasks.size(), 0, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_pairs, c->pairs.data(), sizeof(PairDef) * c->pairs.size(), 0, cudaMemcpyHostToDevice, c->stream));
    // cut-length table: fastsim.h:204-211 evaluated in float32 exactly as written there (Q5)
    std::vector<int> cut(256 * 4);
    for (int s = 0; s < 256; ++s) {
        float Iden = 0.6;
        int k = 0;
        while (Iden <= 1 && k < 4) {
            int cutlength = (int)(s + 24) / (9 * Iden - 4) + 1;
            cut[s * 4 + k] = cutlength;
            Iden += 0.1;
            ++k;
        }
    }
    if (int e = c->d_cut.ensure(cut.size() * sizeof(int))) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_cut.p, cut.data(), cut.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->tables_dirty = false;
    c->profiles_dirty = true;
    return LTG_OK;
}

int build_profiles(ltg_context* c)
{
    if (c->m <= 0) { set_error("no lncRNA loaded (call ltg_set_query first)"); return LTG_ERR_STATE; }
    const int strip_rows = 32 * kScanR;
    const int m16 = 16 * ((c->m + 15) / 16);
    c->n_strips = (m16 + strip_rows - 1) / strip_rows;
    const size_t words = (size_t)c->pairs.size() * c->n_strips * 5 * 32 * kScanR;
    if (int e = c->d_prof_ssw.ensure(words * 4)) return e;
    if (int e = c->d_prof_stats.ensure(words * 4)) return e;
    const int blocks = (int)std::min<size_t>((words + 255) / 256, 148 * 8);
    k_build_profiles<kScanR><<<blocks, 256, 0, c->stream>>>(c->d_rna_ssw.as<uint8_t>(), c->d_rna_stats.as<uint8_t>(), c->m,
                                                             (int)c->pairs.size(), c->n_strips, 0, c->d_prof_ssw.as<uint32_t>());
    k_build_profiles<kScanR><<<blocks, 256, 0, c->stream>>>(c->d_rna_ssw.as<uint8_t>(), c->d_rna_stats.as<uint8_t>(), c->m,
                                                             (int)c->pairs.size(), c->n_strips, 1, c->d_prof_stats.as<uint32_t>());
    c->launches += 2;
    LTG_CUDA_CHECK(cudaGetLastError());
    c->profiles_dirty = false;
    return LTG_OK;
}

int prepare(ltg_context* c)
{
    LTG_CUDA_CHECK(cudaSetDevice(c->device));
    if (c->tables_dirty) if (int e = upload_tables(c)) return e;
    if (c->profiles_dirty) if (int e = build_profiles(c)) return e;
    return LTG_OK;
}

struct HostSeg { int64_t start; int32_t len; int32_t flags; };

// One batch of segments through scan -> peaks -> windows -> traceback.  Everything device-side is
// asynchronous on the context stream; the host syncs where it needs counts.
struct BatchOut {
    std::vector<int> task_max, task_thr, task_npk, task_flags;
    std::vector<int> pk_task, pk_pos, pk_score;
    std::vector<int> fin_sw, fin_cut, fin_rb, fin_re, fin_qb, fin_qe, al_status, al_nt, al_match;
    std::vector<long long> al_stroff;
    std::vector<char> strpool;
    std::vector<uint32_t> colmax;   // only when requested (probe)
    long long window_cells = 0;
    int n_literal_tasks = 0, n_literal_windows = 0;
    float ms_scan = 0, ms_window = 0, ms_scan_kernel = 0;
    int n_scan_launches = 0;
};

int launch_scan(ltg_context* c, int n_items, int max_len, const uint32_t* prof, uint32_t* colmax)
{
    const int blocks = c->num_sms * kScanCtasPerSm;
    const size_t smem = (size_t)kScanWarps * scan_warp_smem_bytes<kScanR>(max_len);
    if (smem > 227 * 1024) { set_error("segment length %d needs %zu bytes of shared memory per CTA", max_len, smem); return LTG_ERR_LIMIT; }
    LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<kScanR, kScanWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int e = c->d_bnd.ensure((size_t)blocks * kScanWarps * max_len * sizeof(uint4))) return e;
    int* counters = c->d_counters.as<int>();
    LTG_CUDA_CHECK(cudaMemsetAsync(counters, 0, sizeof(int), c->stream));
    ScanArgs a;
    a.codes = c->d_codes.as<uint8_t>(); a.segs = c->d_segs.as<SegDesc>(); a.items = c->d_items.as<ScanItem>();
    a.n_items = n_items; a.profiles = prof; a.n_strips = c->n_strips; a.max_len = max_len;
    a.colmax = colmax; a.bnd = c->d_bnd.as<uint4>(); a.counter = counters;
    k_scan<kScanR, kScanWarps><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
    c->launches += 1;
    LTG_CUDA_CHECK(cudaGetLastError());
    return LTG_OK;
}


// ---- literal (Q4) slow path launchers ---------------------------------------------------------
int launch_literal(ltg_context* c, const LiteralJob* d_jobs, int n_jobs, int max_read_len, const WinState* w, int max_len)
{
    const int L = (max_read_len + 15) / 16;
    const long long per_slot = (long long)L * 16 * 4 + 64;
    const int blocks = std::min((n_jobs + 7) / 8, c->num_sms * 8);        // 128 threads = 8 half-warp slots per block
    const int nslots = blocks * 8;
    if (int e = c->d_lit_work.ensure((size_t)nslots * per_slot)) return e;
    LiteralArgs la;
    la.jobs = d_jobs; la.n_jobs = n_jobs; la.codes = c->d_codes.as<uint8_t>(); la.segs = c->d_segs.as<SegDesc>();
    la.rna_ssw = c->d_rna_ssw.as<uint8_t>(); la.work = c->d_lit_work.as<unsigned char>(); la.work_per_slot = per_slot;
    la.colmax16 = c->d_colmax.as<uint16_t>(); la.max_len = max_len;
    if (w) la.w = *w; else memset(&la.w, 0, sizeof la.w);
    k_literal<<<blocks, 128, 0, c->stream>>>(la);
    c->launches += 1;
    LTG_CUDA_CHECK(cudaGetLastError());
    return LTG_OK;
}

int run_literal_scan(ltg_context* c, const std::vector<LiteralJob>& jobs, int max_len)
{
    if (int e = c->d_lit_jobs.ensure(sizeof(LiteralJob) * jobs.size())) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_lit_jobs.p, jobs.data(), sizeof(LiteralJob) * jobs.size(), cudaMemcpyHostToDevice, c->stream));
    if (int e = launch_literal(c, c->d_lit_jobs.as<LiteralJob>(), (int)jobs.size(), c->m, nullptr, max_len)) return e;
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));      // `jobs` is host memory owned by the caller
    return LTG_OK;
}

int literal_windows(ltg_context* c, const WinState& w, bool reverse, int* total)
{
    if (int e = c->d_lit_jobs.ensure(sizeof(LiteralJob) * (size_t)w.n_peaks)) return e;
    int* cnt = c->d_counters.as<int>() + 24;
    LTG_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
    k_lit_collect<<<(w.n_peaks + 255) / 256, 256, 0, c->stream>>>(w, reverse ? 1 : 0, c->d_lit_jobs.as<LiteralJob>(), cnt);
    c->launches += 1;
    int n = 0;
    LTG_CUDA_CHECK(cudaMemcpyAsync(&n, cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (n == 0) return LTG_OK;
    *total += n;
    return launch_literal(c, c->d_lit_jobs.as<LiteralJob>(), n, c->m, &w, 0);
}

int run_windows(ltg_context* c, int n_peaks, int T, const int* forced_cut, BatchOut& out);

int run_batch(ltg_context* c, const std::vector<HostSeg>& segs, bool want_colmax, bool want_alignments, BatchOut& out)
{
    const int T = (int)c->tasks.size(), P = (int)c->pairs.size();
    const int S = (int)segs.size();
    int max_len = 1;
    bool any_stats = !c->rna_plain;
    for (const HostSeg& s : segs) { max_len = std::max(max_len, s.len); if (s.flags & kSegNonACGT) any_stats = true; }
    max_len = (max_len + 3) & ~3;
    const int n_items = S * P, n_tasks = S * T;

    if (int e = c->d_segs.ensure(sizeof(SegDesc) * S)) return e;
    if (int e = c->d_items.ensure(sizeof(ScanItem) * n_items)) return e;
    if (int e = c->d_colmax.ensure((size_t)n_items * max_len * 4)) return e;
    if (int e = c->d_counters.ensure(256)) return e;
    for (DevBuf* b : {&c->d_task_max, &c->d_task_thr, &c->d_task_npk, &c->d_task_flags, &c->d_stats_max})
        if (int e = b->ensure(sizeof(int) * n_tasks)) return e;

    std::vector<SegDesc> hs(S);
    std::vector<ScanItem> items(n_items);
    for (int s = 0; s < S; ++s) {
        hs[s].start = segs[s].start; hs[s].len = segs[s].len; hs[s].flags = segs[s].flags;
        for (int p = 0; p < P; ++p) { items[(size_t)s * P + p].seg = s; items[(size_t)s * P + p].pair = p; }
    }
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_segs.p, hs.data(), sizeof(SegDesc) * S, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_items.p, items.data(), sizeof(ScanItem) * n_items, cudaMemcpyHostToDevice, c->stream));

    LTG_CUDA_CHECK(cudaEventRecord(c->ev[0], c->stream));
    // optional N/U-aware threshold pass (Q3): exact maxima under the Farrar-side scoring
    const int* stats_max = nullptr;
    if (any_stats) {
        if (int e = launch_scan(c, n_items, max_len, c->d_prof_stats.as<uint32_t>(), c->d_colmax.as<uint32_t>())) return e;
        k_rowmax<<<(n_items * 32 + 127) / 128, 128, 0, c->stream>>>(c->d_colmax.as<uint32_t>(), c->d_items.as<ScanItem>(), c->d_segs.as<SegDesc>(),
                                                                   n_items, max_len, T, c->d_stats_max.as<int>());
        c->launches += 1;
        stats_max = c->d_stats_max.as<int>();
    }
    LTG_CUDA_CHECK(cudaEventRecord(c->ev[4], c->stream));
    if (int e = launch_scan(c, n_items, max_len, c->d_prof_ssw.as<uint32_t>(), c->d_colmax.as<uint32_t>())) return e;
    LTG_CUDA_CHECK(cudaEventRecord(c->ev[5], c->stream));
    out.n_scan_launches = 1;

    // peaks
    int pk_cap = std::max(4096, n_tasks * 48);
    int n_peaks = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (DevBuf* b : {&c->d_pk_task, &c->d_pk_pos, &c->d_pk_score}) if (int e = b->ensure(sizeof(int) * (size_t)pk_cap)) return e;
        int* counters = c->d_counters.as<int>();
        LTG_CUDA_CHECK(cudaMemsetAsync(counters + 1, 0, sizeof(int), c->stream));
        EpiArgs ea;
        ea.colmax = c->d_colmax.as<uint32_t>(); ea.items = c->d_items.as<ScanItem>(); ea.segs = c->d_segs.as<SegDesc>();
        ea.n_items = n_items; ea.max_len = max_len; ea.tasks_per_seg = T; ea.stats_max = stats_max; ea.mode = 0;
        ea.task_max = c->d_task_max.as<int>(); ea.task_thr = c->d_task_thr.as<int>(); ea.task_npeaks = c->d_task_npk.as<int>();
        ea.task_flags = c->d_task_flags.as<int>(); ea.pk_count = counters + 1; ea.pk_cap = pk_cap;
        ea.pk_task = c->d_pk_task.as<int>(); ea.pk_pos = c->d_pk_pos.as<int>(); ea.pk_score = c->d_pk_score.as<int>();
        k_epilogue<<<(n_items * 32 + 127) / 128, 128, 0, c->stream>>>(ea);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());

        // Q4 guard: tasks whose exact maximum could carry F >= 132 across a stripe boundary are re-run through
        // the literal striped emulation, their peaks come from the literal column maxima
        out.task_flags.resize(n_tasks);
        LTG_CUDA_CHECK(cudaMemcpyAsync(out.task_flags.data(), c->d_task_flags.p, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::vector<LiteralJob> jobs;
        for (int t = 0; t < n_tasks; ++t) {
            if (out.task_flags[t] & kTaskRange) { set_error("alignment score exceeds the 16-bit range (segment too long for this build)"); return LTG_ERR_LIMIT; }
            if (out.task_flags[t] & kTaskLiteral) {
                LiteralJob j; j.kind = 0; j.task = t; j.seg = t / T; j.tdef = t % T; j.ref_start = 0; j.ref_len = segs[t / T].len;
                j.read_start = 0; j.read_len = c->m; j.read_dir = 1; j.ref_dir = 0; j.terminate = 255; j.peak = -1;
                j.out_item = 0; j.out_half = 0;
                for (int p = 0; p < P; ++p) for (int h = 1; h >= 0; --h) if (c->pairs[p].task[h] == j.tdef) { j.out_item = j.seg * P + p; j.out_half = h; }
                jobs.push_back(j);
            }
        }
        out.n_literal_tasks = (int)jobs.size();
        if (!jobs.empty()) {
            if (int e = run_literal_scan(c, jobs, max_len)) return e;
            ea.mode = 1;
            k_epilogue<<<(n_items * 32 + 127) / 128, 128, 0, c->stream>>>(ea);
            c->launches += 1;
            LTG_CUDA_CHECK(cudaGetLastError());
        }
        LTG_CUDA_CHECK(cudaMemcpyAsync(&n_peaks, counters + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (n_peaks <= pk_cap) break;
        pk_cap = n_peaks + 1024;      // pool too small: grow and redo the (cheap) epilogue
        if (attempt == 1) { set_error("peak pool overflow"); return LTG_ERR_LIMIT; }
    }
    LTG_CUDA_CHECK(cudaEventRecord(c->ev[1], c->stream));

    out.task_max.resize(n_tasks); out.task_thr.resize(n_tasks); out.task_npk.resize(n_tasks);
    LTG_CUDA_CHECK(cudaMemcpyAsync(out.task_max.data(), c->d_task_max.p, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(out.task_thr.data(), c->d_task_thr.p, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(out.task_npk.data(), c->d_task_npk.p, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(out.task_flags.data(), c->d_task_flags.p, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
    out.pk_task.resize(n_peaks); out.pk_pos.resize(n_peaks); out.pk_score.resize(n_peaks);
    if (n_peaks) {
        LTG_CUDA_CHECK(cudaMemcpyAsync(out.pk_task.data(), c->d_pk_task.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(out.pk_pos.data(), c->d_pk_pos.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(out.pk_score.data(), c->d_pk_score.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
    }
    if (want_colmax) {
        out.colmax.resize((size_t)n_items * max_len);
        LTG_CUDA_CHECK(cudaMemcpyAsync(out.colmax.data(), c->d_colmax.p, out.colmax.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    LTG_CUDA_CHECK(cudaEventElapsedTime(&out.ms_scan, c->ev[0], c->ev[1]));
    LTG_CUDA_CHECK(cudaEventElapsedTime(&out.ms_scan_kernel, c->ev[4], c->ev[5]));
    if (!want_alignments || n_peaks == 0) {
        out.fin_sw.assign(n_peaks, 0);
        return LTG_OK;
    }
    return run_windows(c, n_peaks, T, nullptr, out);
}

// ---------------- window stage: peaks (device-resident pool) -> chosen alignments + strings ----------------
int run_windows(ltg_context* c, int n_peaks, int T, const int* forced_cut, BatchOut& out)
{
    LTG_CUDA_CHECK(cudaEventRecord(c->ev[2], c->stream));
    for (int k = 0; k < 12; ++k) if (int e = c->d_w[k].ensure(sizeof(int) * (size_t)n_peaks)) return e;
    if (int e = c->d_cls_list.ensure(sizeof(int) * (size_t)n_peaks * kWinClasses)) return e;
    if (int e = c->d_res.ensure(sizeof(int4) * (size_t)n_peaks)) return e;
    int* counters = c->d_counters.as<int>();
    WinState w;
    w.n_peaks = n_peaks;
    w.pk_task = c->d_pk_task.as<int>(); w.pk_pos = c->d_pk_pos.as<int>(); w.pk_score = c->d_pk_score.as<int>();
    w.w_len = c->d_w[0].as<int>(); w.w_done = c->d_w[1].as<int>();
    w.best_sw = c->d_w[2].as<int>(); w.best_cut = c->d_w[3].as<int>(); w.best_re = c->d_w[4].as<int>(); w.best_qe = c->d_w[5].as<int>();
    w.fin_sw = c->d_w[6].as<int>(); w.fin_cut = c->d_w[7].as<int>(); w.fin_re = c->d_w[8].as<int>(); w.fin_qe = c->d_w[9].as<int>();
    w.fin_rb = c->d_w[10].as<int>(); w.fin_qb = c->d_w[11].as<int>();
    w.cls_count = counters + 8; w.cls_list = c->d_cls_list.as<int>(); w.cap = n_peaks;
    w.res = c->d_res.as<int4>(); w.bin_counter = counters + 2;
    w.codes = c->d_codes.as<uint8_t>(); w.segs = c->d_segs.as<SegDesc>(); w.tasks_per_seg = T;
    w.rna_ssw = c->d_rna_ssw.as<uint8_t>(); w.m = c->m; w.cut_table = c->d_cut.as<int>();
    w.cell_counter = reinterpret_cast<long long*>(counters + 16);
    w.forced_cut = forced_cut;
    LTG_CUDA_CHECK(cudaMemsetAsync(counters + 16, 0, 8, c->stream));
    const int pb = (n_peaks + 255) / 256;
    const int dp_blocks = c->num_sms * 4;
    std::vector<int> h_sw(n_peaks), h_done(n_peaks);
    for (int round = 0; round < 4; ++round) {
        LTG_CUDA_CHECK(cudaMemsetAsync(counters + 2, 0, sizeof(int), c->stream));
        LTG_CUDA_CHECK(cudaMemsetAsync(counters + 8, 0, sizeof(int) * kWinClasses, c->stream));
        k_win_plan<<<pb, 256, 0, c->stream>>>(w, round);
        k_win_dp<false><<<dp_blocks, 128, 0, c->stream>>>(w);
        c->launches += 2;
        // Q4 guard for windows: exact forward scores >= 148 are recomputed by the literal emulation
        if (int e = literal_windows(c, w, /*reverse=*/false, &out.n_literal_windows)) return e;
        k_win_decide<<<pb, 256, 0, c->stream>>>(w, round);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
    }
    // reverse pass over the chosen alignments
    LTG_CUDA_CHECK(cudaMemsetAsync(counters + 2, 0, sizeof(int), c->stream));
    LTG_CUDA_CHECK(cudaMemsetAsync(counters + 8, 0, sizeof(int) * kWinClasses, c->stream));
    k_win_plan<<<pb, 256, 0, c->stream>>>(w, -1);
    k_win_dp<true><<<dp_blocks, 128, 0, c->stream>>>(w);
    k_win_finish<<<pb, 256, 0, c->stream>>>(w);
    c->launches += 3;
    if (int e = literal_windows(c, w, /*reverse=*/true, &out.n_literal_windows)) return e;
    LTG_CUDA_CHECK(cudaGetLastError());

    // traceback + string expansion
    const int tb_threads = c->num_sms * 256;
    const long long scratch_small = 24 * 1024, scratch_big = 8LL << 20;
    if (int e = c->d_scratch.ensure((size_t)tb_threads * scratch_small)) return e;
    for (DevBuf* b : {&c->d_al_status, &c->d_al_nt, &c->d_al_match}) if (int e = b->ensure(sizeof(int) * (size_t)n_peaks)) return e;
    if (int e = c->d_al_stroff.ensure(sizeof(long long) * (size_t)n_peaks)) return e;
    long long strcap = std::max<long long>(1 << 20, (long long)n_peaks * 512);
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (int e = c->d_strpool.ensure((size_t)strcap)) return e;
        LTG_CUDA_CHECK(cudaMemsetAsync(counters + 20, 0, 8, c->stream));
        TraceArgs ta;
        ta.w = w; ta.dna = c->d_dna.as<unsigned char>(); ta.rna_raw = c->d_rna_raw.as<unsigned char>();
        ta.scratch = c->d_scratch.as<unsigned char>(); ta.scratch_per_thread = scratch_small; ta.only_overflow = 0;
        ta.al_status = c->d_al_status.as<int>(); ta.al_nt = c->d_al_nt.as<int>(); ta.al_match = c->d_al_match.as<int>();
        ta.al_stroff = c->d_al_stroff.as<long long>(); ta.strpool = c->d_strpool.as<char>(); ta.strcap = strcap;
        ta.str_count = reinterpret_cast<long long*>(counters + 20);
        k_traceback<<<tb_threads / 64, 64, 0, c->stream>>>(ta);
        c->launches += 1;
        out.al_status.resize(n_peaks);
        long long used = 0;
        LTG_CUDA_CHECK(cudaMemcpyAsync(out.al_status.data(), c->d_al_status.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(&used, counters + 20, 8, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (used > strcap) { strcap = used + (1 << 20); continue; }      // string pool too small: redo
        int n_over = 0;
        for (int s : out.al_status) n_over += (s == 2);
        if (n_over) {
            const int big_threads = 256;
            if (int e = c->d_scratch_big.ensure((size_t)big_threads * scratch_big)) return e;
            ta.scratch = c->d_scratch_big.as<unsigned char>(); ta.scratch_per_thread = scratch_big; ta.only_overflow = 1;
            k_traceback<<<big_threads / 64, 64, 0, c->stream>>>(ta);
            c->launches += 1;
            LTG_CUDA_CHECK(cudaMemcpyAsync(out.al_status.data(), c->d_al_status.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
            LTG_CUDA_CHECK(cudaMemcpyAsync(&used, counters + 20, 8, cudaMemcpyDeviceToHost, c->stream));
            LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            if (used > strcap) { strcap = used + (1 << 20); continue; }
            for (int s : out.al_status) if (s == 2) { set_error("traceback band exceeds the device scratch"); return LTG_ERR_LIMIT; }
        }
        out.strpool.resize((size_t)used);
        if (used) LTG_CUDA_CHECK(cudaMemcpyAsync(out.strpool.data(), c->d_strpool.p, (size_t)used, cudaMemcpyDeviceToHost, c->stream));
        break;
    }
    LTG_CUDA_CHECK(cudaEventRecord(c->ev[3], c->stream));
    auto fetch = [&](std::vector<int>& v, const int* d) -> int {
        v.resize(n_peaks);
        LTG_CUDA_CHECK(cudaMemcpyAsync(v.data(), d, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
        return LTG_OK;
    };
    if (int e = fetch(out.fin_sw, w.fin_sw)) return e;
    if (int e = fetch(out.fin_cut, w.fin_cut)) return e;
    if (int e = fetch(out.fin_rb, w.fin_rb)) return e;
    if (int e = fetch(out.fin_re, w.fin_re)) return e;
    if (int e = fetch(out.fin_qb, w.fin_qb)) return e;
    if (int e = fetch(out.fin_qe, w.fin_qe)) return e;
    if (int e = fetch(out.al_nt, c->d_al_nt.as<int>())) return e;
    if (int e = fetch(out.al_match, c->d_al_match.as<int>())) return e;
    out.al_stroff.resize(n_peaks);
    LTG_CUDA_CHECK(cudaMemcpyAsync(out.al_stroff.data(), c->d_al_stroff.p, sizeof(long long) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(&out.window_cells, counters + 16, 8, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    LTG_CUDA_CHECK(cudaEventElapsedTime(&out.ms_window, c->ev[2], c->ev[3]));
    return LTG_OK;
}

// cutSequence — fastsim.h:71-90
int cut_segments(int64_t len, const ltg_params& P, std::vector<HostSeg>& segs)
{
    if (P.cut_length <= 0 || P.cut_length - P.overlap <= 0) { set_error("cut length must exceed the overlap"); return LTG_ERR_ARG; }
    segs.clear();
    for (int64_t pos = 0; pos < len; pos += P.cut_length - P.overlap) {
        HostSeg s; s.start = pos; s.len = (int32_t)std::min<int64_t>(P.cut_length, len - pos); s.flags = 0;
        segs.push_back(s);
    }
    return LTG_OK;
}

struct ResultBuilder {
    std::vector<ltg_triplex> tri;
    std::string text;
    int64_t chr_off = -1;
    void add(const ltg_host::Triplex& t, const char* chr, int64_t record_start, int record)
    {
        if (chr_off < 0) { chr_off = (int64_t)text.size(); text += (chr ? chr : ""); text += '\0'; }
        ltg_triplex o;
        memset(&o, 0, sizeof o);
        o.stari = t.stari; o.endi = t.endi; o.starj = t.starj; o.endj = t.endj; o.reverse = t.reverse; o.strand = t.strand;
        o.rule = t.rule; o.nt = t.nt; o.score = t.score; o.identity = t.identity; o.tri_score = t.tri_score;
        o.genomestart = t.starj + record_start - 1;         // Fasim-LongTarget.cpp:146-147
        o.genomeend = t.endj + record_start - 1;
        o.tfo_off = (int64_t)text.size(); text += t.tfo; text += '\0';
        o.tts_off = (int64_t)text.size(); text += t.tts; text += '\0';
        o.chr_off = chr_off;
        o.record = record;
        tri.push_back(o);
    }
};

ltg_result* finish_result(ResultBuilder& rb)
{
    ltg_result* r = (ltg_result*)calloc(1, sizeof(ltg_result));
    r->n_triplex = (int64_t)rb.tri.size();
    r->triplex = (ltg_triplex*)malloc(sizeof(ltg_triplex) * std::max<size_t>(1, rb.tri.size()));
    if (!rb.tri.empty()) memcpy(r->triplex, rb.tri.data(), sizeof(ltg_triplex) * rb.tri.size());
    r->text_bytes = (int64_t)rb.text.size();
    r->text = (char*)malloc(std::max<size_t>(1, rb.text.size()));
    if (!rb.text.empty()) memcpy(r->text, rb.text.data(), rb.text.size());
    return r;
}

int scan_device_impl(ltg_context* c, const unsigned char* d_dna_user, const char* h_dna, int64_t len, const char* chr,
                     int64_t record_start, ltg_result** out)
{
    (void)chr;
    if (!c || !out) { set_error("null argument"); return LTG_ERR_ARG; }
    if (int e = prepare(c)) return e;
    const int64_t launches0 = c->launches;
    std::vector<HostSeg> segs;
    if (int e = cut_segments(len, c->params, segs)) return e;
    ResultBuilder rb;
    ltg_result stats;
    memset(&stats, 0, sizeof stats);
    if (len > 0) {
        // the record lives in d_dna (either copied from the host or device-to-device from the user's buffer)
        if (int e = c->d_dna.ensure((size_t)len)) return e;
        if (int e = c->d_codes.ensure((size_t)len)) return e;
        if (h_dna) LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_dna.p, h_dna, (size_t)len, cudaMemcpyHostToDevice, c->stream));
        else LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_dna.p, d_dna_user, (size_t)len, cudaMemcpyDeviceToDevice, c->stream));
        k_encode<<<std::min<int64_t>((len + 255) / 256, 148 * 16), 256, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_codes.as<uint8_t>(), len);
        c->launches += 1;
        // segment flags (same_seq + non-ACGT) for the whole record
        const int NS = (int)segs.size();
        if (int e = c->d_segs.ensure(sizeof(SegDesc) * NS)) return e;
        std::vector<SegDesc> hs(NS);
        for (int s = 0; s < NS; ++s) { hs[s].start = segs[s].start; hs[s].len = segs[s].len; hs[s].flags = 0; }
        LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_segs.p, hs.data(), sizeof(SegDesc) * NS, cudaMemcpyHostToDevice, c->stream));
        k_seg_flags<<<(NS * 32 + 127) / 128, 128, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_segs.as<SegDesc>(), NS);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaMemcpyAsync(hs.data(), c->d_segs.p, sizeof(SegDesc) * NS, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::vector<HostSeg> active;
        for (int s = 0; s < NS; ++s) { segs[s].flags = hs[s].flags; if (!(hs[s].flags & kSegSkip)) active.push_back(segs[s]); }

        const int T = (int)c->tasks.size();
        std::vector<ltg_host::Triplex> record_list;
        for (size_t b0 = 0; b0 < active.size(); b0 += kBatchSegments) {
            std::vector<HostSeg> batch(active.begin() + b0, active.begin() + std::min(active.size(), b0 + (size_t)kBatchSegments));
            BatchOut bo;
            if (int e = run_batch(c, batch, false, true, bo)) return e;
            stats.gpu_ms_scan += bo.ms_scan; stats.gpu_ms_window += bo.ms_window;
            stats.gpu_ms_scan_kernel += bo.ms_scan_kernel; stats.n_scan_launches += bo.n_scan_launches;
            stats.n_peaks += (int64_t)bo.pk_task.size(); stats.window_cells += bo.window_cells;
            stats.n_literal_tasks += bo.n_literal_tasks; stats.n_literal_windows += bo.n_literal_windows;
            // order peaks by (task, position): the reference walks tasks in order and peaks by ascending column
            const int np = (int)bo.pk_task.size();
            std::vector<int> order(np);
            std::iota(order.begin(), order.end(), 0);
            std::sort(order.begin(), order.end(), [&](int x, int y) {
                if (bo.pk_task[x] != bo.pk_task[y]) return bo.pk_task[x] < bo.pk_task[y];
                return bo.pk_pos[x] < bo.pk_pos[y];
            });
            size_t k = 0;
            std::vector<ltg_host::Triplex> mine;
            for (int task = 0; task < (int)batch.size() * T; ++task) {
                mine.clear();
                const HostSeg& sg = batch[task / T];
                const TaskDef& td = c->tasks[task % T];
                while (k < order.size() && bo.pk_task[order[k]] == task) {
                    const int i = order[k++];
                    if (bo.fin_sw[i] <= 0 || bo.al_status[i] != 1) continue;     // fastsim.h:253 (sw_score == 0 -> skipped)
                    ltg_host::DeviceAlignment al;
                    const int ws = bo.pk_pos[i] - bo.fin_cut[i] + 1;
                    al.sw_score = bo.fin_sw[i]; al.ref_begin = ws + bo.fin_rb[i]; al.ref_end = ws + bo.fin_re[i];
                    al.query_begin = bo.fin_qb[i]; al.query_end = bo.fin_qe[i];
                    al.nt = bo.al_nt[i]; al.match = bo.al_match[i];
                    al.tfo = bo.strpool.data() + bo.al_stroff[i];
                    al.tts = al.tfo + al.nt + 1;
                    ltg_host::make_triplex(al, sg.len, (long)sg.start, td.para, td.strand, td.rule, c->params, mine);
                }
                if (!mine.empty()) ltg_host::finish_task(mine, c->params, record_list);
            }
            stats.n_segments += (int64_t)batch.size();
            stats.n_tasks += (int64_t)batch.size() * T;
            for (const HostSeg& sg : batch) stats.scan_cells += (int64_t)sg.len * c->m * T;
        }
        for (const ltg_host::Triplex& t : record_list)
            if (ltg_host::passes_record_filter(t, c->params)) rb.add(t, chr, record_start, 0);
    }
    ltg_result* r = finish_result(rb);
    r->n_segments = stats.n_segments; r->n_tasks = stats.n_tasks; r->scan_cells = stats.scan_cells; r->dna_bases = len;
    r->n_peaks = stats.n_peaks; r->window_cells = stats.window_cells; r->n_literal_tasks = stats.n_literal_tasks;
    r->n_literal_windows = stats.n_literal_windows; r->gpu_ms_scan = stats.gpu_ms_scan; r->gpu_ms_window = stats.gpu_ms_window;
    r->gpu_launches = c->launches - launches0;
    r->gpu_ms_scan_kernel = stats.gpu_ms_scan_kernel; r->n_scan_launches = stats.n_scan_launches;
    *out = r;
    return LTG_OK;
}

}  // namespace

// =========================================== C ABI ===========================================
extern "C" {

const char* ltg_last_error(void) { return g_err; }

int ltg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void ltg_default_params(ltg_params* p)
{
    if (!p) return;
    p->rule = 0; p->cut_length = 5000; p->strand = 0; p->overlap = 100; p->nt_min = 20; p->nt_max = 100000;
    p->min_identity = 60; p->min_stability = 1; p->penalty_t = -1000; p->penalty_c = 0; p->c_distance = 15; p->c_length = 50;
}

int ltg_create(int device, ltg_context** out)
{
    if (!out) { set_error("null argument"); return LTG_ERR_ARG; }
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) { set_error("no CUDA device available (%s) — this library has no CPU fallback", cudaGetErrorString(ce)); return LTG_ERR_CUDA; }
    if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return LTG_ERR_ARG; }
    LTG_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    LTG_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_error("device %s is sm_%d%d; this build contains sm_100a code only", prop.name, prop.major, prop.minor); return LTG_ERR_CUDA; }
    ltg_context* c = new ltg_context();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    ltg_default_params(&c->params);
    LTG_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 6; ++i) LTG_CUDA_CHECK(cudaEventCreate(&c->ev[i]));
    *out = c;
    return LTG_OK;
}

void ltg_destroy(ltg_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->d_rna_raw, &c->d_rna_ssw, &c->d_rna_stats, &c->d_prof_ssw, &c->d_prof_stats, &c->d_cut, &c->d_dna, &c->d_codes,
                      &c->d_segs, &c->d_items, &c->d_colmax, &c->d_bnd, &c->d_counters, &c->d_task_max, &c->d_task_thr, &c->d_task_npk,
                      &c->d_task_flags, &c->d_stats_max, &c->d_pk_task, &c->d_pk_pos, &c->d_pk_score, &c->d_cls_list, &c->d_res,
                      &c->d_al_status, &c->d_al_nt, &c->d_al_match, &c->d_al_stroff, &c->d_strpool, &c->d_scratch, &c->d_scratch_big,
                      &c->d_lit_colmax, &c->d_lit_work, &c->d_lit_jobs})
        b->release();
    for (int k = 0; k < 12; ++k) c->d_w[k].release();
    for (int i = 0; i < 6; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* ltg_stream(ltg_context* c) { return c ? (void*)c->stream : nullptr; }

int ltg_set_params(ltg_context* c, const ltg_params* p)
{
    if (!c || !p) { set_error("null argument"); return LTG_ERR_ARG; }
    if (p->cut_length <= 0 || p->cut_length - p->overlap <= 0) { set_error("cut length (%d) must be positive and exceed the overlap (%d)", p->cut_length, p->overlap); return LTG_ERR_ARG; }
    if (p->cut_length > 6500) { set_error("cut length %d exceeds the 16-bit score envelope of this build (max 6500)", p->cut_length); return LTG_ERR_LIMIT; }
    std::vector<TaskDef> probe;
    if (!ltg_host::enumerate_tasks(*p, probe)) { set_error("invalid rule/strand selection (rule=%d strand=%d)", p->rule, p->strand); return LTG_ERR_ARG; }
    if (p->rule != c->params.rule || p->strand != c->params.strand) c->tables_dirty = true;
    c->params = *p;
    return LTG_OK;
}

int ltg_set_query(ltg_context* c, const char* name, const char* rna, int64_t len)
{
    if (!c || !rna || len <= 0) { set_error("empty lncRNA"); return LTG_ERR_ARG; }
    if (len > (1 << 24)) { set_error("lncRNA longer than 16 Mnt is not supported"); return LTG_ERR_LIMIT; }
    LTG_CUDA_CHECK(cudaSetDevice(c->device));
    c->rna_name = name ? name : "";
    c->rna.assign(rna, (size_t)len);
    c->m = (int)len;
    std::vector<uint8_t> q1(len), q2(len);
    c->rna_plain = true;
    for (int64_t i = 0; i < len; ++i) {
        const unsigned char ch = (unsigned char)rna[i];
        q1[i] = (uint8_t)ssw_code(ch); q2[i] = (uint8_t)stats_code(ch);
        if (q1[i] == 4 || q2[i] >= 4) c->rna_plain = false;      // U or any non-ACGT letter: scorings differ (Q3)
    }
    for (DevBuf* b : {&c->d_rna_raw, &c->d_rna_ssw, &c->d_rna_stats}) if (int e = b->ensure((size_t)len)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_raw.p, rna, (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_ssw.p, q1.data(), (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_stats.p, q2.data(), (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->profiles_dirty = true;
    return prepare(c);
}

int ltg_scan_record(ltg_context* c, const char* dna, int64_t len, const char* chr, int64_t record_start, ltg_result** out)
{
    if (!dna && len > 0) { set_error("null DNA"); return LTG_ERR_ARG; }
    return scan_device_impl(c, nullptr, dna, len, chr, record_start, out);
}

int ltg_scan_device(ltg_context* c, const void* d_dna, int64_t len, const char* chr, int64_t record_start, ltg_result** out)
{
    if (!d_dna && len > 0) { set_error("null DNA"); return LTG_ERR_ARG; }
    return scan_device_impl(c, (const unsigned char*)d_dna, nullptr, len, chr, record_start, out);
}

int ltg_result_new(ltg_result** out)
{
    if (!out) return LTG_ERR_ARG;
    ResultBuilder rb;
    *out = finish_result(rb);
    return LTG_OK;
}

void ltg_result_free(ltg_result* r)
{
    if (!r) return;
    free(r->triplex);
    free(r->text);
    free(r);
}

int ltg_result_append(ltg_result* dst, const ltg_result* src)
{
    if (!dst || !src) { set_error("null argument"); return LTG_ERR_ARG; }
    const int64_t n0 = dst->n_triplex, t0 = dst->text_bytes;
    dst->triplex = (ltg_triplex*)realloc(dst->triplex, sizeof(ltg_triplex) * (size_t)std::max<int64_t>(1, n0 + src->n_triplex));
    dst->text = (char*)realloc(dst->text, (size_t)std::max<int64_t>(1, t0 + src->text_bytes));
    memcpy(dst->text + t0, src->text, (size_t)src->text_bytes);
    for (int64_t i = 0; i < src->n_triplex; ++i) {
        ltg_triplex t = src->triplex[i];
        t.tfo_off += t0; t.tts_off += t0; t.chr_off += t0;
        dst->triplex[n0 + i] = t;
    }
    dst->n_triplex += src->n_triplex; dst->text_bytes += src->text_bytes;
    dst->n_segments += src->n_segments; dst->n_tasks += src->n_tasks; dst->scan_cells += src->scan_cells; dst->dna_bases += src->dna_bases;
    dst->n_peaks += src->n_peaks; dst->window_cells += src->window_cells; dst->n_literal_tasks += src->n_literal_tasks;
    dst->n_literal_windows += src->n_literal_windows; dst->gpu_ms_scan += src->gpu_ms_scan; dst->gpu_ms_window += src->gpu_ms_window;
    dst->gpu_launches += src->gpu_launches;
    dst->gpu_ms_scan_kernel += src->gpu_ms_scan_kernel; dst->n_scan_launches += src->n_scan_launches;
    return LTG_OK;
}

int ltg_cluster(ltg_result* r, const ltg_params* p)
{
    if (!r || !p) { set_error("null argument"); return LTG_ERR_ARG; }
    std::vector<ltg_triplex> v(r->triplex, r->triplex + r->n_triplex);
    ltg_host::cluster(v, p->c_distance, p->c_length, nullptr);
    std::sort(v.begin(), v.end(), [](const ltg_triplex& a, const ltg_triplex& b) { return a.motif < b.motif; });   // :813, :847
    if (!v.empty()) memcpy(r->triplex, v.data(), sizeof(ltg_triplex) * v.size());
    return LTG_OK;
}

int ltg_probe_segment(ltg_context* c, const char* seg, int32_t seg_len, ltg_task_probe* tasks, int32_t n_tasks, int32_t* colmax,
                      int32_t* peak_score, int32_t* peak_pos, int32_t peak_cap)
{
    if (!c || !seg || seg_len <= 0 || !tasks) { set_error("bad argument"); return LTG_ERR_ARG; }
    if (seg_len > c->params.cut_length) { set_error("segment longer than the cut length"); return LTG_ERR_ARG; }
    if (int e = prepare(c)) return e;
    if (int e = c->d_dna.ensure((size_t)seg_len)) return e;
    if (int e = c->d_codes.ensure((size_t)seg_len)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_dna.p, seg, (size_t)seg_len, cudaMemcpyHostToDevice, c->stream));
    k_encode<<<(seg_len + 255) / 256, 256, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_codes.as<uint8_t>(), seg_len);
    c->launches += 1;
    std::vector<HostSeg> segs(1);
    segs[0].start = 0; segs[0].len = seg_len; segs[0].flags = 0;
    for (int i = 0; i < seg_len; ++i) { const char ch = seg[i]; if (!(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T')) segs[0].flags |= kSegNonACGT; }
    BatchOut bo;
    if (int e = run_batch(c, segs, colmax != nullptr, false, bo)) return e;
    const int T = (int)c->tasks.size(), P = (int)c->pairs.size();
    const int max_len = (seg_len + 3) & ~3;
    for (int q = 0; q < n_tasks; ++q) {
        int t = -1;
        for (int k = 0; k < T; ++k) if (c->tasks[k].para == tasks[q].para && c->tasks[k].strand == tasks[q].strand && c->tasks[k].rule == tasks[q].rule) t = k;
        if (t < 0) { set_error("task (%d,%d,%d) is not part of the configured rule/strand selection", tasks[q].para, tasks[q].strand, tasks[q].rule); return LTG_ERR_ARG; }
        tasks[q].max_score = bo.task_max[t]; tasks[q].threshold = bo.task_thr[t]; tasks[q].n_peaks = bo.task_npk[t];
        tasks[q].literal = (bo.task_flags[t] & kTaskLiteral) ? 1 : 0;
        if (colmax) {
            int item = -1, half = 0;
            for (int p = 0; p < P; ++p) for (int h = 0; h < 2; ++h) if (c->pairs[p].task[h] == t && item < 0) { item = p; half = h; }
            const uint32_t* row = bo.colmax.data() + (size_t)item * max_len;
            bool cut = false;
            for (int j = 0; j < seg_len; ++j) {
                int v = half ? hi16(row[j]) : lo16(row[j]);
                if (v >= kOverflowU8) cut = true;          // Q2: nothing is recorded from the first >= 251 column on
                colmax[(size_t)q * seg_len + j] = cut ? 0 : v;
            }
        }
        if (peak_score && peak_pos) {
            std::vector<std::pair<int, int>> pk;
            for (size_t i = 0; i < bo.pk_task.size(); ++i) if (bo.pk_task[i] == t) pk.push_back({bo.pk_pos[i], bo.pk_score[i]});
            std::sort(pk.begin(), pk.end());
            for (size_t i = 0; i < pk.size() && (int)i < peak_cap; ++i) { peak_pos[(size_t)q * peak_cap + i] = pk[i].first; peak_score[(size_t)q * peak_cap + i] = pk[i].second; }
        }
    }
    return LTG_OK;
}


// Aligner::Align for explicit windows: every window becomes a one-task "segment" under an identity rule image,
// with one forced peak at its last column, and runs through the very same window kernels as the product path.
int ltg_probe_align(ltg_context* c, const char* const* windows, const int32_t* window_len, int32_t n, int32_t* out6, uint32_t* cigar,
                    int32_t cigar_cap)
{
    if (!c || !windows || !window_len || n <= 0 || !out6) { set_error("bad argument"); return LTG_ERR_ARG; }
    if (int e = prepare(c)) return e;
    // identity task table (restored afterwards by marking the tables dirty)
    TaskDef id; memset(&id, 0, sizeof id);
    id.para = 1; id.strand = 0; id.rule = 1; id.reversed = 0; id.comp_src = 0;
    for (int k = 0; k < 5; ++k) id.img[k] = (int8_t)k;
    PairDef pd; pd.task[0] = pd.task[1] = 0; pd.reversed = 0; pd.pad_ = 0;
    LTG_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_tasks, &id, sizeof id, 0, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_pairs, &pd, sizeof pd, 0, cudaMemcpyHostToDevice, c->stream));
    c->tables_dirty = true;
    std::string cat;
    std::vector<SegDesc> segs(n);
    std::vector<int> pk_task(n), pk_pos(n), pk_score(n, 0), fcut(n);
    for (int i = 0; i < n; ++i) {
        if (window_len[i] <= 0 || window_len[i] > kMaxWindow) { set_error("window %d: length %d outside 1..%d", i, window_len[i], kMaxWindow); return LTG_ERR_LIMIT; }
        segs[i].start = (int64_t)cat.size(); segs[i].len = window_len[i]; segs[i].flags = 0;
        cat.append(windows[i], (size_t)window_len[i]);
        pk_task[i] = i; pk_pos[i] = window_len[i] - 1; fcut[i] = window_len[i];
    }
    if (int e = c->d_dna.ensure(cat.size())) return e;
    if (int e = c->d_codes.ensure(cat.size())) return e;
    if (int e = c->d_segs.ensure(sizeof(SegDesc) * n)) return e;
    if (int e = c->d_counters.ensure(256)) return e;
    for (DevBuf* b : {&c->d_pk_task, &c->d_pk_pos, &c->d_pk_score, &c->d_stats_max}) if (int e = b->ensure(sizeof(int) * (size_t)n)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_dna.p, cat.data(), cat.size(), cudaMemcpyHostToDevice, c->stream));
    // translated text -> codes: the same A0 C1 G2 T3 else 4 coding (the windows are already translated DNA)
    k_encode<<<(int)((cat.size() + 255) / 256), 256, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_codes.as<uint8_t>(), (int64_t)cat.size());
    c->launches += 1;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_segs.p, segs.data(), sizeof(SegDesc) * n, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_pk_task.p, pk_task.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_pk_pos.p, pk_pos.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_pk_score.p, pk_score.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_stats_max.p, fcut.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    BatchOut bo;
    bo.pk_task = pk_task; bo.pk_pos = pk_pos; bo.pk_score = pk_score;
    if (int e = run_windows(c, n, 1, c->d_stats_max.as<int>(), bo)) return e;
    for (int i = 0; i < n; ++i) {
        int32_t* o = out6 + (size_t)i * 6;
        if (bo.fin_sw[i] <= 0 || bo.al_status[i] != 1) { for (int k = 0; k < 6; ++k) o[k] = 0; continue; }
        o[0] = bo.fin_sw[i]; o[1] = bo.fin_rb[i]; o[2] = bo.fin_re[i]; o[3] = bo.fin_qb[i]; o[4] = bo.fin_qe[i];
        // run-length CIGAR from the aligned strings (M: both present, I: gap on the DNA side, D: gap on the RNA side)
        const char* tfo = bo.strpool.data() + bo.al_stroff[i];
        const char* tts = tfo + bo.al_nt[i] + 1;
        int nc = 0, run = 0, prev = -1;
        for (int k = 0; k <= bo.al_nt[i]; ++k) {
            const int op = k == bo.al_nt[i] ? -2 : (tts[k] == '-' ? 1 : (tfo[k] == '-' ? 2 : 0));
            if (op == prev) { ++run; continue; }
            if (prev >= 0) { if (cigar && nc < cigar_cap) cigar[(size_t)i * cigar_cap + nc] = ((uint32_t)run << 4) | (uint32_t)prev; ++nc; }
            prev = op; run = 1;
        }
        o[5] = nc;
    }
    return LTG_OK;
}

}  // extern "C"

#include "../host/driver.inl"
