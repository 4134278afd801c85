// libfasim_b200.so — context, device memory model, pipeline orchestration and the C ABI of
// include/fasim_b200.h.  One context = one GPU = one stream.  No CPU compute fallback: every DP cell is
// computed by the kernels in scan.cuh / window.cuh / literal.cuh.
//
// Pipeline per DNA record:  H2D -> k_encode / k_seg_flags -> batches of segments, each
//   device phase  k_scan (strip maxima) -> k_epilogue x3 (+ literal re-runs) -> window rounds -> traceback pass 1
//                 -> asynchronous D2H of one plain record per peak into a pinned batch slot
//   host phase    (worker threads, overlapped with the device phase of the next batch) per-task
//                 de-duplication / ranking / filters of fastSIM's tail (fastsim.h:273-288)
// -> record filter -> traceback pass 2 (strings of the surviving rows only) -> result.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fasim_b200.h"
#include "common.cuh"
#include "scan.cuh"
#include "window.cuh"
#include "literal.cuh"
#include "sim.cuh"
#include "../host/rules_table.hpp"
#include "../host/triplex_host.hpp"
#include "../host/sim_host.hpp"

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

// page-locked host memory (asynchronous D2H at full PCIe rate)
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return LTG_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        LTG_CUDA_CHECK(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
        return LTG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Scan tiling: R RNA rows per lane (strip = 32 * R rows).  R = 32 has the lowest per-cell overhead; R = 16 halves the
// strip and is chosen for lncRNAs that would mostly pad a 1024-row strip.  LTG_SCAN_R forces one of them (tuning builds).
#ifndef LTG_SCAN_WARPS
#define LTG_SCAN_WARPS 4
#endif
constexpr int kScanWarps = LTG_SCAN_WARPS;  // warps per CTA
inline int scan_ctas_per_sm(int R) { return R >= 32 ? 2 : 3; }
inline int choose_scan_r(int m)
{
#ifdef LTG_SCAN_R
    (void)m;
    return LTG_SCAN_R;
#else
    const long long m16 = 16LL * ((m + 15) / 16);
    const long long pad32 = ((m16 + 1023) / 1024) * 1024, pad16 = ((m16 + 511) / 512) * 512;
    return pad32 * 100 <= pad16 * 104 ? 32 : 16;      // R = 32 is ~4.5 % faster per row
#endif
}
constexpr int kMaxCutLength = 24000;       // shared memory: the segment's base codes sit next to the strip profile (scan.cuh)
#ifndef LTG_BATCH_SEGMENTS
#define LTG_BATCH_SEGMENTS 2048
#endif
constexpr int kBatchSegments = LTG_BATCH_SEGMENTS;
constexpr size_t kStripBytesPerBatch = 8ull << 30;      // cap of the column-maxima buffers (whole-column rows + block maxima) of one batch

struct HostSeg {
    int64_t start;      // offset in the device DNA buffer of this call (records are laid out back to back)
    int32_t len;
    int32_t flags;
    int64_t coord;      // offset of the segment in its own record (what StartInSeq / EndInSeq count from)
    int32_t record;     // index of the record within the call
    int32_t pad_;
};

// counters of a batch's device phase as they land in pinned memory (each member is the target of one D2H copy)
struct BatchScalars {
    long long window_cells;                 // counters[kCntCells..]
    int cells_pad[6];
    int lit[2]; int lit_total; int lit_pad[5];          // counters[kCntLit], [kCntLit+1], [kCntLitTotal]
    unsigned long long win_stats[20];       // WinSched::st_windows, st_cells
    int trace_handed[4];                    // counters[kCntOvf..]
    int cand[4];                            // shipped records, shipped tasks, unfinished alignments
    unsigned long long filter_stats[6];     // WinSched::st_filter
};

// what one batch leaves on the host: pinned copies filled by asynchronous D2H, consumed by the host phase
struct HostBatch {
    std::vector<HostSeg> segs;
    int n_tasks = 0, n_peaks = 0;
    PinBuf task_off, jobs, tout, task_info, scalars, c_task, c_poff;
    bool keep_only_literal = false;                 // kLitOnly batches: the host phase keeps the literal tasks' rows only
    DevBuf d_cjobs, d_ctout, d_ctask, d_cpoff;      // compacted records of this batch (kept until the slot is retired)
    int n_ctasks = 0, n_cpeaks = 0;
    int64_t d2h_late = 0;                           // bytes the worker copied
    cudaEvent_t ready = nullptr, ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool timed = false, timed_windows = false, has_stats_pass = false;
    int n_literal_tasks = 0;
    std::vector<ltg_host::Triplex> rows;        // output of the host phase, task order
    std::string error;
    std::thread worker;
};

}  // namespace ltg

using namespace ltg;

struct ltg_context {
    int device = 0;
    int num_sms = 0;
    int host_threads = 1;
    bool scan_shared = false;           // LTG_SCAN_SHARED=1: k_scan with one profile copy per CTA and 3 warps per scheduler — measured
                                        // SLOWER than one copy per warp at 2 warps per scheduler (5581 vs 7089 GCUPS, profiles/README.md)
    bool sim_mode = false;              // -F: SIM() instead of fastSIM() per task (ltg_set_sim_mode)
    bool compat = false;                // window loop / per-task tail of the older variant (ltg_set_compat)
    bool prune = true, dead_rule = true, skip_rounds = true, q4_probe = true, lit_col = true, floor_s = false;
    int batch_segments = kBatchSegments;                  // segments per device batch (LTG_BATCH_SEGMENTS)
    int lit_rows_per_chunk = 24, lit_min_chunks = 0;      // tuning of the column-parallel literal kernel (LTG_LIT_ROWS / LTG_LIT_CH)
    int64_t n_probe_items = 0;          // pairs swept a second time by the Q4 probe (diagnostics)
    // Stripe-start screen in the main sweep (k_scan FREC, 3 instructions per step, measured 4.9 % of the sweep): a pre-filter with
    // the resolution of one lane next to the granule pre-filter.  On the example lncRNAs it spares only 12-15 % of the pairs the
    // second (taint / probe) sweep (their high-scoring regions span many lanes), which does not pay for it: off unless LTG_FREC=1
    // (-1 = switch a query to it once a batch had to sweep more than 4 % of its pairs again).
    int frec_mode = 0;                  // -1 auto, 0 off, 1 on
    bool q4_taint = true;               // the flagged pairs are swept by the taint variant (certifies tasks) instead of the probe variant
                                        // (LTG_Q4_TAINT=0: probe)
    // Window sweeps that also watch for an F >= 132 entering a stripe start (k_win_dp Q4CHK: ~3 % more instructions): only the
    // windows that saw one go through the literal emulation.  Switched on for a query once a batch sent more than 64 windows
    // there (LTG_WIN_Q4CHK=0 never, 1 always); without it every window that scores >= 148 does.
    int win_q4_mode = -1;               // -1 auto, 0 off, 1 on
    bool win_q4_on = false;
    bool frec_on = false;               // current query: main sweeps record the carried F
    cudaStream_t stream = nullptr, copy_stream = nullptr, lit_stream = nullptr;
    cudaEvent_t lit_event = nullptr;
    ltg_params params;
    // task tables (depend on params.rule / params.strand)
    std::vector<TaskDef> tasks;
    std::vector<PairDef> pairs;
    bool tables_dirty = true;
    bool params_changed = true;     // rule / strand changed since the profiles were built
    // query
    std::string rna_name, rna;
    bool rna_plain = true;          // only ACGT (any case): the SSW-side and Farrar-side scorings coincide
    bool rna_acgt = true;           // every SSW code of the lncRNA is 0..3 (A/C/G/T, U counts as A): table lookup scoring applies
    int m = 0, n_strips = 0, scan_r = 32;
    bool profiles_dirty = true;
    DevBuf d_rna_raw, d_rna_ssw, d_rna_stats, d_rna_sel, d_prof_ssw, d_prof_stats, d_prof_ssw2, d_cut;
    DevBuf d_packed, d_nblocks;         // 2-bit packed input staged for the device-side expansion
    DevBuf d_rna_sim, d_sim_scratch, d_sim_hdr, d_sim_pool, d_sim_tasks;      // -F mode (sim.cuh)
    // record / batch buffers
    DevBuf d_dna, d_codes, d_segs, d_items, d_items_stats, d_blkmax, d_bnd, d_counters;
    DevBuf d_probe_items, d_probe_orig, d_probe_out, d_bnd_gran, d_scan_order, d_frec, d_taint_colmax;
    int n_bnd_gran = 0;                 // granules that hold the rows just above a stripe start (Q4 pre-filter)
    DevBuf d_task_info, d_task_off, d_stats_max, d_task_litrow, d_cand;
    DevBuf d_pk_task, d_pk_pos, d_pk_score;
    DevBuf d_w[24], d_pc[4], d_res64, d_win_list, d_win_sched, d_res, d_colmax_all, d_ovf_list;
    DevBuf d_jobs, d_tout, d_strpool, d_scratch, d_scratch_big;
    DevBuf d_lit_colmax, d_lit_work, d_lit_jobs;
    DevBuf d_side_jobs, d_side_colmax;       // literal scan jobs running on the side stream while the main batches compute
    int side_rows = 0, side_pitch = 0;
    HostBatch hb[2];
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    unsigned long long win_stats[30] = {0};     // windows / cells planned per (round, retry) + reverse; [20..22] traceback tier hand-overs
};

namespace {

// The task / pair tables are __constant__ symbols, i.e. one copy per DEVICE, while rule / strand selections belong to a
// context.  Contexts that share a device (`--devices 0,0`) may therefore only coexist with identical tables; the registry
// below (per device: content of the uploaded tables and the contexts that rely on it) refuses anything else loudly
// instead of letting one context's kernels run with another one's rule images.
struct DeviceTables { std::vector<unsigned char> bytes; std::vector<const ltg_context*> users; };
std::mutex g_tables_mu;
DeviceTables g_tables[64];

std::vector<unsigned char> table_bytes(const std::vector<TaskDef>& t, const std::vector<PairDef>& p)
{
    std::vector<unsigned char> b(t.size() * sizeof(TaskDef) + p.size() * sizeof(PairDef));
    if (!t.empty()) memcpy(b.data(), t.data(), t.size() * sizeof(TaskDef));
    if (!p.empty()) memcpy(b.data() + t.size() * sizeof(TaskDef), p.data(), p.size() * sizeof(PairDef));
    return b;
}
// claims the device's tables for `c` with content `bytes`; false if another live context relies on different content
bool claim_tables(const ltg_context* c, int device, const std::vector<unsigned char>& bytes)
{
    std::lock_guard<std::mutex> lk(g_tables_mu);
    DeviceTables& d = g_tables[device & 63];
    bool others = false;
    for (const ltg_context* u : d.users) if (u != c) others = true;
    if (others && d.bytes != bytes) return false;
    d.bytes = bytes;
    if (std::find(d.users.begin(), d.users.end(), c) == d.users.end()) d.users.push_back(c);
    return true;
}
void release_tables(const ltg_context* c, int device)
{
    std::lock_guard<std::mutex> lk(g_tables_mu);
    DeviceTables& d = g_tables[device & 63];
    d.users.erase(std::remove(d.users.begin(), d.users.end(), c), d.users.end());
    if (d.users.empty()) d.bytes.clear();
}
bool tables_current(const ltg_context* c, int device, const std::vector<unsigned char>& bytes)
{
    std::lock_guard<std::mutex> lk(g_tables_mu);
    const DeviceTables& d = g_tables[device & 63];
    return d.bytes == bytes && std::find(d.users.begin(), d.users.end(), c) != d.users.end();
}

// No exception may cross the C ABI: every extern "C" entry point runs its body through this.
template <class F> int guarded(F&& body)
{
    try { return body(); }
    catch (const std::bad_alloc&) { set_error("out of host memory"); return LTG_ERR_LIMIT; }
    catch (const std::exception& e) { set_error("internal error: %s", e.what()); return LTG_ERR_STATE; }
    catch (...) { set_error("internal error (unknown exception)"); return LTG_ERR_STATE; }
}

// per-task scalars of the scan stage, one row of 5 ints per array: [max | thr | npeaks | flags | jstar][n_tasks]
struct TaskInfo {
    int* max; int* thr; int* npk; int* flags; int* jstar;
    TaskInfo(int* base, int n) : max(base), thr(base + n), npk(base + 2 * (size_t)n), flags(base + 3 * (size_t)n), jstar(base + 4 * (size_t)n) {}
};

enum { kCntScan = 0, kCntPeaks = 1, kCntOvf = 2 /* .. 5 */, kCntCand = 6 /* peaks, tasks, unfinished */, kCntCells = 16, kCntLit = 24, kCntLitTotal = 26, kCntTotal = 32 };

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
            for (int h = 1; h >= 0; --h) { c->tasks[p.task[h]].pair = (int8_t)c->pairs.size(); c->tasks[p.task[h]].half = (int8_t)h; }
            c->pairs.push_back(p);
        }
    }
    if (!claim_tables(c, c->device, table_bytes(c->tasks, c->pairs))) {
        set_error("another context on device %d uses a different rule/strand selection: contexts that share a GPU must share it", c->device);
        return LTG_ERR_STATE;
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
    c->scan_r = choose_scan_r(c->m);
    const int strip_rows = 32 * c->scan_r;
    const int m16 = 16 * ((c->m + 15) / 16);
    c->n_strips = (m16 + strip_rows - 1) / strip_rows;
    const size_t words = (size_t)c->pairs.size() * c->n_strips * 5 * 32 * c->scan_r;
    if (int e = c->d_prof_ssw.ensure(words * 4)) return e;
    if (int e = c->d_prof_stats.ensure(words * 4)) return e;
    if (int e = c->d_prof_ssw2.ensure(words * 4)) return e;
    const int blocks = (int)std::min<size_t>((words + 255) / 256, 148 * 8);
    for (int kind = 0; kind < 3; ++kind) {
        uint32_t* dst = kind == 1 ? c->d_prof_stats.as<uint32_t>() : kind == 2 ? c->d_prof_ssw2.as<uint32_t>() : c->d_prof_ssw.as<uint32_t>();
        if (c->scan_r == 32)
            k_build_profiles<32><<<blocks, 256, 0, c->stream>>>(c->d_rna_ssw.as<uint8_t>(), c->d_rna_stats.as<uint8_t>(), c->m, (int)c->pairs.size(), c->n_strips, kind, dst);
        else
            k_build_profiles<16><<<blocks, 256, 0, c->stream>>>(c->d_rna_ssw.as<uint8_t>(), c->d_rna_stats.as<uint8_t>(), c->m, (int)c->pairs.size(), c->n_strips, kind, dst);
    }
    c->launches += 3;
    LTG_CUDA_CHECK(cudaGetLastError());
    // Q4 pre-filter: the granules that hold one of the 28 rows above a stripe start of the reference's 16-lane layout (an F
    // that enters such a row with >= 132 was opened below a cell >= 148 at most 26 rows up)
    {
        const int stripe = (c->m + 15) / 16, n_gran = c->n_strips * (32 * c->scan_r / kGranRows);
        std::vector<char> mark((size_t)n_gran, 0);
        for (int k = 1; k < 16; ++k) {
            const int b = k * stripe;
            if (b >= m16) break;
            for (int g = std::max(0, b - 28) / kGranRows; g <= (b - 1) / kGranRows && g < n_gran; ++g) mark[g] = 1;
        }
        std::vector<int> list;
        for (int g = 0; g < n_gran; ++g) if (mark[g]) list.push_back(g);
        c->n_bnd_gran = (int)list.size();
        if (int e = c->d_bnd_gran.ensure(sizeof(int) * std::max<size_t>(1, list.size()))) return e;
        if (!list.empty()) LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_bnd_gran.p, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));      // `list` is host memory of this scope
    }
    {
        c->frec_on = (c->frec_mode == 1);
        c->win_q4_on = (c->win_q4_mode == 1);
    }
    c->profiles_dirty = false;
    return LTG_OK;
}

int prepare(ltg_context* c)
{
    LTG_CUDA_CHECK(cudaSetDevice(c->device));
    if (!c->tables_dirty && !tables_current(c, c->device, table_bytes(c->tasks, c->pairs))) c->tables_dirty = true;
    if (c->tables_dirty) { const bool keep_profiles = !c->profiles_dirty && !c->params_changed; if (int e = upload_tables(c)) return e; if (keep_profiles) c->profiles_dirty = false; c->params_changed = false; }
    if (c->profiles_dirty) if (int e = build_profiles(c)) return e;
    if (int e = c->d_counters.ensure(sizeof(int) * kCntTotal)) return e;
    // 16-bit cells: a score never exceeds 5 * min(rows, columns).  Cut lengths beyond 6500 are therefore fine for lncRNAs up to
    // 6399 nt (the reference accepts any -c); a long lncRNA AND a long cut would overflow and is refused here, before any work
    if (c->m > 0 && 5LL * std::min<long long>(16LL * ((c->m + 15) / 16), c->params.cut_length) >= 32000) {
        set_error("cut length %d with a lncRNA of %d nt exceeds the 16-bit score envelope of this build (min(lncRNA, cut) must stay below 6400)",
                  c->params.cut_length, c->m);
        return LTG_ERR_LIMIT;
    }
    return LTG_OK;
}

// what the function-level probes want back from a batch (everything host-side, synchronous)
struct ProbeOut {
    std::vector<int> task_max, task_thr, task_npk, task_flags, task_off;
    std::vector<int> pk_pos, pk_score;
    std::vector<uint32_t> colmax;           // [item][max_len] maximum over the granules
    std::vector<uint16_t> lit_colmax;       // [literal row][max_len]
    std::vector<int> task_litrow;
    int max_len = 0;
};

// probe_out != nullptr: the Q4 probe variant (no column maxima; per item the largest F carried into a stripe start)
constexpr int kScanSharedWarps = 6;         // warps per CTA of the shared-profile scan (2 CTAs per SM: 3 warps per scheduler)
// probe_out != nullptr: the Q4 probe variant (no column maxima; per item the largest F carried into a stripe start).
// h_items (host copy of the item list) enables the shared-profile variant: the items are grouped by task pair, kScanSharedWarps to a
// group, and a CTA keeps ONE copy of the pair's profiles for its six warps — what lifts the scan from 2 to 3 warps per scheduler.
int launch_scan(ltg_context* c, const ScanItem* d_items, int n_items, int max_len, const uint32_t* prof, uint32_t* colmax_all, uint16_t* blkmax,
                uint32_t* probe_out = nullptr, const int* task_jstar = nullptr, const std::vector<ScanItem>* h_items = nullptr, uint16_t* frec = nullptr,
                bool taint = false, const int* task_flags = nullptr)
{
    const int R = c->scan_r;
    int* counters = c->d_counters.as<int>();
    LTG_CUDA_CHECK(cudaMemsetAsync(counters + kCntScan, 0, sizeof(int), c->stream));
    ScanArgs a;
    a.codes = c->d_codes.as<uint8_t>(); a.segs = c->d_segs.as<SegDesc>(); a.items = d_items;
    a.n_items = n_items; a.profiles = prof; a.n_strips = c->n_strips; a.max_len = max_len;
    a.colmax_all = colmax_all; a.blkmax = blkmax; a.blk_pitch = blk_pitch_for(max_len);
    a.counter = counters + kCntScan;
    a.order = nullptr; a.group_pair = nullptr; a.n_groups = 0;
    a.probe_out = probe_out; a.task_jstar = task_jstar; a.tasks_per_seg = (int)c->tasks.size(); a.stripe_len = (c->m + 15) / 16;
    a.frec = frec; a.task_flags = task_flags;
    // shared-profile variant when the pair's profiles (all strips) and six warps' private parts fit twice into an SM
    const size_t prof_bytes = (size_t)c->n_strips * 5 * 32 * R * 4;
    const size_t smem_shared = prof_bytes + (size_t)kScanSharedWarps * (R == 32 ? scan_warp_smem_bytes_shared<32>(max_len) : scan_warp_smem_bytes_shared<16>(max_len));
    if (c->scan_shared && !probe_out && !frec && !taint && h_items && n_items >= 4 * kScanSharedWarps && smem_shared <= 112 * 1024) {
        const int W = kScanSharedWarps, P = (int)c->pairs.size();
        std::vector<std::vector<int> > by_pair((size_t)P);
        for (int i = 0; i < n_items; ++i) by_pair[(*h_items)[i].pair].push_back(i);
        std::vector<int> order, gpair;
        for (int p = 0; p < P; ++p)
            for (size_t k = 0; k < by_pair[p].size(); k += W) {
                gpair.push_back(p);
                for (int w = 0; w < W; ++w) order.push_back(k + w < by_pair[p].size() ? by_pair[p][k + w] : -1);
            }
        const int n_groups = (int)gpair.size();
        if (int e = c->d_scan_order.ensure(sizeof(int) * (order.size() + gpair.size()))) return e;
        int* d_order = c->d_scan_order.as<int>();
        LTG_CUDA_CHECK(cudaMemcpyAsync(d_order, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(d_order + order.size(), gpair.data(), sizeof(int) * gpair.size(), cudaMemcpyHostToDevice, c->stream));
        c->h2d_bytes += (int64_t)sizeof(int) * (order.size() + gpair.size());
        a.order = d_order; a.group_pair = d_order + order.size(); a.n_groups = n_groups;
        const int blocks = std::min(n_groups, c->num_sms * 2);
        if (int e = c->d_bnd.ensure((size_t)blocks * W * max_len * sizeof(uint2))) return e;
        a.bnd = c->d_bnd.as<uint2>();
        if (R == 32) {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<32, kScanSharedWarps, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_shared));
            k_scan<32, kScanSharedWarps, false, true><<<blocks, W * 32, smem_shared, c->stream>>>(a);
        } else {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<16, kScanSharedWarps, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_shared));
            k_scan<16, kScanSharedWarps, false, true><<<blocks, W * 32, smem_shared, c->stream>>>(a);
        }
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));          // `order` / `gpair` are host memory of this scope
        return LTG_OK;
    }
    const int blocks = c->num_sms * scan_ctas_per_sm(R);
    const size_t smem = (size_t)kScanWarps * (R == 32 ? scan_warp_smem_bytes<32>(max_len) : scan_warp_smem_bytes<16>(max_len));
    if (smem > 227 * 1024) { set_error("segment length %d needs %zu bytes of shared memory per CTA", max_len, smem); return LTG_ERR_LIMIT; }
    if (int e = c->d_bnd.ensure((size_t)blocks * kScanWarps * max_len * sizeof(uint2))) return e;
    a.bnd = c->d_bnd.as<uint2>();
    if (taint) {
        if (R == 32) {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<32, kScanWarps, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_scan<32, kScanWarps, false, false, false, true><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
        } else {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<16, kScanWarps, false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_scan<16, kScanWarps, false, false, false, true><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
        }
    } else if (probe_out) {
        if (R == 32) {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<32, kScanWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_scan<32, kScanWarps, true><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
        } else {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<16, kScanWarps, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_scan<16, kScanWarps, true><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
        }
    } else if (frec) {
        if (R == 32) {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<32, kScanWarps, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_scan<32, kScanWarps, false, false, true><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
        } else {
            LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<16, kScanWarps, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_scan<16, kScanWarps, false, false, true><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
        }
    } else if (R == 32) {
        LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<32, kScanWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_scan<32, kScanWarps><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
    } else {
        LTG_CUDA_CHECK(cudaFuncSetAttribute(k_scan<16, kScanWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_scan<16, kScanWarps><<<blocks, kScanWarps * 32, smem, c->stream>>>(a);
    }
    c->launches += 1;
    LTG_CUDA_CHECK(cudaGetLastError());
    return LTG_OK;
}

// ---- literal (Q4) slow path launchers ---------------------------------------------------------
// n_jobs < 0: the job count lives on the device (n_jobs_dev); the grid is sized for the worst case and idles when
// the list is empty, so no host round trip is needed to decide whether to launch.
int launch_literal(ltg_context* c, const LiteralJob* d_jobs, int n_jobs, const int* n_jobs_dev, int max_read_len, const WinState* w, int max_len,
                   bool side = false, int row_base = 0)
{
    const int pitch = literal_pitch(max_read_len);
    // column-parallel kernel (one CTA per job, byte workspace in shared memory) for all but short reads
    const long long col_smem = (long long)kLitColArrays * 16 * pitch + 16;     // + slack for the kernel's one-ahead loads
    if (c->lit_col && max_read_len >= 256 && col_smem <= 220 * 1024) {
        const int L = (max_read_len + 15) / 16;
        int CH = 8;
        while (CH < kLitColMaxChunks && L > c->lit_rows_per_chunk * CH) CH *= 2;
        if (c->lit_min_chunks > 0) CH = std::max(2, std::min(kLitColMaxChunks, c->lit_min_chunks & ~1));      // forced (tuning)
        const int threads = 16 * CH;
        const int per_sm = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(16, 2048 / threads), (220 * 1024) / col_smem));
        const int blocks = n_jobs >= 0 ? std::max(1, std::min(n_jobs, c->num_sms * per_sm)) : c->num_sms * per_sm;
        LiteralArgs la;
        la.jobs = d_jobs; la.n_jobs = n_jobs; la.n_jobs_dev = n_jobs_dev; la.codes = c->d_codes.as<uint8_t>(); la.segs = c->d_segs.as<SegDesc>();
        la.rna_ssw = c->d_rna_ssw.as<uint8_t>(); la.use_smem = 1; la.slots_per_block = 1; la.pitch = pitch;
        la.work = nullptr; la.work_per_slot = 0;
        la.lit_colmax = c->d_lit_colmax.as<uint16_t>(); la.max_len = max_len; la.task_litrow = c->d_task_litrow.as<int>(); la.row_base = row_base;
        if (side) { la.segs = nullptr; la.lit_colmax = c->d_side_colmax.as<uint16_t>(); la.task_litrow = nullptr; }
        if (w) la.w = *w; else memset(&la.w, 0, sizeof la.w);
        LTG_CUDA_CHECK(cudaFuncSetAttribute(k_literal_col, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)col_smem));
        k_literal_col<<<blocks, threads, (size_t)col_smem, side ? c->lit_stream : c->stream>>>(la);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
        return LTG_OK;
    }
    // workspace per half-warp slot: kLitArrays uint16 arrays of 16 lanes x pitch; in shared memory when at least one slot fits
    const long long per_slot = 2LL * 16 * kLitArrays * pitch;
    const long long smem_budget = 200 * 1024;
    int spb = (int)std::min<long long>(8, smem_budget / per_slot);          // slots per block
    const bool use_smem = spb >= 1;
    if (!use_smem) spb = 8;
    const int threads = std::max(32, ((spb * 16 + 31) / 32) * 32);
    const int want_blocks = n_jobs >= 0 ? std::max(1, (n_jobs + spb - 1) / spb) : c->num_sms * 2;
    const int blocks = std::min(want_blocks, c->num_sms * 8);
    LiteralArgs la;
    la.jobs = d_jobs; la.n_jobs = n_jobs; la.n_jobs_dev = n_jobs_dev; la.codes = c->d_codes.as<uint8_t>(); la.segs = c->d_segs.as<SegDesc>();
    la.rna_ssw = c->d_rna_ssw.as<uint8_t>(); la.use_smem = use_smem ? 1 : 0; la.slots_per_block = spb; la.pitch = pitch;
    la.work = nullptr; la.work_per_slot = per_slot;
    size_t smem = 0;
    if (use_smem) {
        smem = (size_t)spb * per_slot;
        LTG_CUDA_CHECK(cudaFuncSetAttribute(k_literal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        if (int e = c->d_lit_work.ensure((size_t)blocks * spb * per_slot)) return e;
        la.work = c->d_lit_work.as<unsigned char>();
    }
    la.lit_colmax = c->d_lit_colmax.as<uint16_t>(); la.max_len = max_len; la.task_litrow = c->d_task_litrow.as<int>(); la.row_base = row_base;
    if (side) { la.segs = nullptr; la.lit_colmax = c->d_side_colmax.as<uint16_t>(); la.task_litrow = nullptr; }      // jobs carry their geometry
    if (w) la.w = *w; else memset(&la.w, 0, sizeof la.w);
    k_literal<<<blocks, threads, smem, side ? c->lit_stream : c->stream>>>(la);
    c->launches += 1;
    LTG_CUDA_CHECK(cudaGetLastError());
    return LTG_OK;
}

int literal_windows(ltg_context* c, const WinState& w, bool reverse, int round)
{
    if (int e = c->d_lit_jobs.ensure(sizeof(LiteralJob) * (size_t)w.n_peaks)) return e;
    int* cnt = c->d_counters.as<int>() + kCntLit + (reverse ? 1 : 0);
    LTG_CUDA_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), c->stream));
    k_lit_collect<<<(w.n_peaks + 255) / 256, 256, 0, c->stream>>>(w, reverse ? 1 : 0, round, c->d_lit_jobs.as<LiteralJob>(), cnt,
                                                                 c->d_counters.as<int>() + kCntLitTotal);
    c->launches += 1;
    return launch_literal(c, c->d_lit_jobs.as<LiteralJob>(), -1, cnt, c->m, &w, 0);
}

// ---------------- window stage: peaks (device-resident pool) -> chosen alignments -> traceback pass 1 ----------------
int run_windows(ltg_context* c, int n_peaks, int T, const int* forced_cut, const uint16_t* gran_blk, int max_len, HostBatch* hb, bool dead_rule)
{
    for (int k = 0; k < 24; ++k) if (int e = c->d_w[k].ensure(sizeof(int) * (size_t)n_peaks)) return e;
    const int pc_cap = 4 * n_peaks;          // kMaxRuns pieces per peak at most
    if (int e = c->d_win_list.ensure(sizeof(int) * (size_t)pc_cap)) return e;
    for (int k = 0; k < 4; ++k) if (int e = c->d_pc[k].ensure(sizeof(int) * (size_t)pc_cap)) return e;
    if (int e = c->d_res64.ensure(sizeof(unsigned long long) * (size_t)n_peaks)) return e;
    if (int e = c->d_win_sched.ensure(sizeof(WinSched))) return e;
    if (int e = c->d_res.ensure(sizeof(int4) * (size_t)n_peaks)) return e;
    int* counters = c->d_counters.as<int>();
    WinState w;
    w.n_peaks = n_peaks;
    w.pk_task = c->d_pk_task.as<int>(); w.pk_pos = c->d_pk_pos.as<int>(); w.pk_score = c->d_pk_score.as<int>();
    w.w_len = c->d_w[0].as<int>(); w.w_done = c->d_w[1].as<int>();
    w.best_sw = c->d_w[2].as<int>(); w.best_cut = c->d_w[3].as<int>(); w.best_re = c->d_w[4].as<int>(); w.best_qe = c->d_w[5].as<int>();
    w.fin_sw = c->d_w[6].as<int>(); w.fin_cut = c->d_w[7].as<int>(); w.fin_re = c->d_w[8].as<int>(); w.fin_qe = c->d_w[9].as<int>();
    w.fin_rb = c->d_w[10].as<int>(); w.fin_qb = c->d_w[11].as<int>();
    w.w_bound = c->d_w[14].as<int>(); w.w_flight = c->d_w[15].as<int>(); w.w_floor = c->d_w[16].as<int>();
    w.pc_peak = c->d_pc[0].as<int>(); w.pc_lo = c->d_pc[1].as<int>(); w.pc_rows = c->d_pc[2].as<int>(); w.pc_key = c->d_pc[3].as<int>(); w.pc_cap = pc_cap;
    w.res64 = c->d_res64.as<unsigned long long>(); w.w_next = c->d_w[17].as<int>(); w.w_probe = c->d_w[18].as<int>();
    w.w_ws = c->d_w[12].as<int>(); w.fin_ws = c->d_w[13].as<int>(); w.w_shift = c->d_w[19].as<int>(); w.best_ws = c->d_w[20].as<int>();
    w.fin_shift = c->d_w[21].as<int>();
    const bool q4chk = c->win_q4_on;
    w.w_q4 = q4chk ? c->d_w[22].as<int>() : nullptr;
    w.compat = (c->compat && forced_cut == nullptr) ? 1 : 0;
    w.sched = c->d_win_sched.as<WinSched>(); w.list = c->d_win_list.as<int>();
    w.res = c->d_res.as<int4>();
    w.codes = c->d_codes.as<uint8_t>(); w.segs = c->d_segs.as<SegDesc>(); w.tasks_per_seg = T;
    w.rna_ssw = c->d_rna_ssw.as<uint8_t>(); w.rna_sel = c->d_rna_sel.as<uint16_t>(); w.m = c->m; w.cut_table = c->d_cut.as<int>();
    w.cell_counter = reinterpret_cast<long long*>(counters + kCntCells);
    w.forced_cut = forced_cut;
    w.gran_blk = c->prune ? gran_blk : nullptr; w.n_gran = c->n_strips * (32 * c->scan_r / kGranRows); w.gran_rows = kGranRows;
    w.blk_pitch = blk_pitch_for(max_len); w.scan_r = c->scan_r; w.floor_s = c->floor_s ? 1 : 0;
    w.n_pairs = (int)c->pairs.size();
    LTG_CUDA_CHECK(cudaMemsetAsync(counters + kCntCells, 0, 8, c->stream));
    LTG_CUDA_CHECK(cudaMemsetAsync(counters + kCntLitTotal, 0, sizeof(int), c->stream));
    LTG_CUDA_CHECK(cudaMemsetAsync(w.sched, 0, sizeof(WinSched), c->stream));
    LTG_CUDA_CHECK(cudaMemsetAsync(w.w_flight, 0, sizeof(int) * (size_t)n_peaks, c->stream));
    const int pb = (n_peaks + 255) / 256, plan_blocks = (n_peaks * 8 + 255) / 256;
    const int dp_blocks = c->num_sms * 4;
    // plan -> key offsets -> placement: the sorted work list of the next k_win_dp launch
    const int place_blocks = (pc_cap + 255) / 256;
    auto schedule = [&](int round, int retry) {
        cudaMemsetAsync(&w.sched->n_pieces, 0, sizeof(int), c->stream);
        k_win_plan<<<plan_blocks, 256, 0, c->stream>>>(w, round, retry);
        k_win_offsets<<<1, 1024, 0, c->stream>>>(w.sched);
        k_win_place<<<place_blocks, 256, 0, c->stream>>>(w);
        c->launches += 3;
    };
    // the sweep itself, then the per-peak combination of its pieces
    auto sweep = [&](bool rev) {
        if (q4chk) {
            if (rev) { if (c->rna_acgt) k_win_dp<true, true, true><<<dp_blocks, 128, 0, c->stream>>>(w); else k_win_dp<true, false, true><<<dp_blocks, 128, 0, c->stream>>>(w); }
            else { if (c->rna_acgt) k_win_dp<false, true, true><<<dp_blocks, 128, 0, c->stream>>>(w); else k_win_dp<false, false, true><<<dp_blocks, 128, 0, c->stream>>>(w); }
        }
        else if (rev) { if (c->rna_acgt) k_win_dp<true, true><<<dp_blocks, 128, 0, c->stream>>>(w); else k_win_dp<true, false><<<dp_blocks, 128, 0, c->stream>>>(w); }
        else { if (c->rna_acgt) k_win_dp<false, true><<<dp_blocks, 128, 0, c->stream>>>(w); else k_win_dp<false, false><<<dp_blocks, 128, 0, c->stream>>>(w); }
        k_win_combine<<<pb, 256, 0, c->stream>>>(w);
        c->launches += 2;
    };
    for (int round = 0; round < 4; ++round) {
        if (q4chk) LTG_CUDA_CHECK(cudaMemsetAsync(w.w_q4, 0, sizeof(int) * (size_t)n_peaks, c->stream));      // flags of this round's windows
        for (int retry = 0; retry < (w.gran_blk ? 2 : 1); ++retry) {
            schedule(round, retry);
            sweep(false);
        }
        // Q4 guard for windows: exact forward scores >= 148 are recomputed by the literal emulation
        if (int e = literal_windows(c, w, /*reverse=*/false, round)) return e;
        k_win_decide<<<pb, 256, 0, c->stream>>>(w, round);
        c->launches += 1;
        if (round < 3 && c->skip_rounds && !w.compat) {
            // reverse probe of the rounds that failed without a candidate: may skip later rounds or finish the peak
            schedule(-2, 0);
            sweep(true);
            k_win_probe<<<pb, 256, 0, c->stream>>>(w, round);
            c->launches += 1;
        }
        LTG_CUDA_CHECK(cudaGetLastError());
    }
    // reverse pass over the chosen alignments
    if (q4chk) LTG_CUDA_CHECK(cudaMemsetAsync(w.w_q4, 0, sizeof(int) * (size_t)n_peaks, c->stream));
    schedule(-1, 0);
    sweep(true);
    k_win_finish<<<pb, 256, 0, c->stream>>>(w);
    c->launches += 1;
    if (int e = literal_windows(c, w, /*reverse=*/true, -1)) return e;
    LTG_CUDA_CHECK(cudaGetLastError());

    // traceback pass 1: nt / identity / stability of every chosen alignment
    if (int e = c->d_jobs.ensure(sizeof(TraceJob) * (size_t)n_peaks)) return e;
    if (int e = c->d_tout.ensure(sizeof(TraceOut) * (size_t)n_peaks)) return e;
    DeadRule dr;
    dr.need_nt = std::max(c->params.nt_min, c->params.c_length); dr.nt_min = c->params.nt_min; dr.nt_max = c->params.nt_max;
    dr.penalty_t = c->params.penalty_t; dr.penalty_c = c->params.penalty_c; dr.min_st = (float)c->params.min_stability;
    dr.dna = (dead_rule && c->dead_rule) ? c->d_dna.as<unsigned char>() : nullptr;     // nullptr: trace everything
    k_make_trace_jobs<<<pb, 256, 0, c->stream>>>(w, c->d_jobs.as<TraceJob>(), c->d_tout.as<TraceOut>(), dr);
    c->launches += 1;
    (void)hb;
    return LTG_OK;
}

// The tiers of the traceback over a job list: three shared-memory tiers (small region / many threads for the typical
// alignment, larger regions / fewer threads for long alignments and wide bands), then the generic kernel with a 24 KB
// global scratch, then with an 8 MB scratch for the rare huge bands.  Each tier hands what it cannot finish to the next through a device-side list.
constexpr int kTrace1Tpb = 128, kTrace1Bytes = 452, kTrace2Tpb = 64, kTrace2Bytes = 1796, kTrace3Tpb = 32, kTrace3Bytes = 7172;
int run_traceback(ltg_context* c, const TraceJob* d_jobs, int n_jobs, TraceOut* d_out, char* strpool, bool skip_dead = false)
{
    const int tb_threads = c->num_sms * 64;
    const long long scratch_small = 24 * 1024, scratch_big = 8LL << 20;
    const int big_threads = 128;
    if (int e = c->d_scratch.ensure((size_t)tb_threads * scratch_small)) return e;
    if (int e = c->d_scratch_big.ensure((size_t)big_threads * scratch_big)) return e;
    if (int e = c->d_ovf_list.ensure(sizeof(int) * 4 * (size_t)n_jobs)) return e;
    int* cnt = c->d_counters.as<int>() + kCntOvf;
    int* list1 = c->d_ovf_list.as<int>(), * list2 = list1 + n_jobs, * list3 = list2 + n_jobs, * list4 = list3 + n_jobs;
    LTG_CUDA_CHECK(cudaMemsetAsync(cnt, 0, 4 * sizeof(int), c->stream));
    TraceArgs ta;
    ta.jobs = d_jobs; ta.n_jobs = n_jobs; ta.codes = c->d_codes.as<uint8_t>(); ta.dna = c->d_dna.as<unsigned char>();
    ta.rna_ssw = c->d_rna_ssw.as<uint8_t>(); ta.rna_raw = c->d_rna_raw.as<unsigned char>();
    ta.nt_min = c->params.nt_min; ta.nt_max = c->params.nt_max; ta.penalty_t = c->params.penalty_t; ta.penalty_c = c->params.penalty_c;
    ta.out = d_out; ta.strpool = strpool; ta.skip_dead = skip_dead ? 1 : 0;
    ta.scratch = nullptr; ta.scratch_per_thread = 0;
    // tier 1
    ta.in_list = nullptr; ta.in_count = nullptr; ta.out_list = list1; ta.out_count = cnt;
    const int smem1 = kTrace1Tpb * kTrace1Bytes, smem2 = kTrace2Tpb * kTrace2Bytes, smem3 = kTrace3Tpb * kTrace3Bytes;
    LTG_CUDA_CHECK(cudaFuncSetAttribute(k_traceback_fast<kTrace1Tpb, kTrace1Bytes>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    LTG_CUDA_CHECK(cudaFuncSetAttribute(k_traceback_fast<kTrace2Tpb, kTrace2Bytes>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    LTG_CUDA_CHECK(cudaFuncSetAttribute(k_traceback_fast<kTrace3Tpb, kTrace3Bytes>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    const int blocks1 = std::max(1, std::min(c->num_sms * 4, (n_jobs + kTrace1Tpb - 1) / kTrace1Tpb));
    k_traceback_fast<kTrace1Tpb, kTrace1Bytes><<<blocks1, kTrace1Tpb, smem1, c->stream>>>(ta);
    // tier 2
    ta.in_list = list1; ta.in_count = cnt; ta.out_list = list2; ta.out_count = cnt + 1;
    const int blocks2 = std::max(1, std::min(c->num_sms * 2, (n_jobs + kTrace2Tpb - 1) / kTrace2Tpb));
    k_traceback_fast<kTrace2Tpb, kTrace2Bytes><<<blocks2, kTrace2Tpb, smem2, c->stream>>>(ta);
    // tier 3
    ta.in_list = list2; ta.in_count = cnt + 1; ta.out_list = list3; ta.out_count = cnt + 2;
    k_traceback_fast<kTrace3Tpb, kTrace3Bytes><<<c->num_sms, kTrace3Tpb, smem3, c->stream>>>(ta);
    // tier 4
    ta.scratch = c->d_scratch.as<unsigned char>(); ta.scratch_per_thread = scratch_small;
    ta.in_list = list3; ta.in_count = cnt + 2; ta.out_list = list4; ta.out_count = cnt + 3;
    k_traceback<<<tb_threads / 64, 64, 0, c->stream>>>(ta);
    // tier 5
    ta.scratch = c->d_scratch_big.as<unsigned char>(); ta.scratch_per_thread = scratch_big;
    ta.in_list = list4; ta.in_count = cnt + 3; ta.out_list = nullptr; ta.out_count = nullptr;
    k_traceback<<<big_threads / 64, 64, 0, c->stream>>>(ta);
    c->launches += 5;
    LTG_CUDA_CHECK(cudaGetLastError());
    return LTG_OK;
}

// Device phase of one batch of segments: scan -> peaks -> windows -> traceback pass 1 -> async D2H into `hb`.
// The host only synchronises where it needs a count (literal tasks, number of peaks).
// How a batch treats the tasks that need the literal (Q4) re-run, which is slow per task (one half-warp sweeps the whole
// matrix) but embarrassingly parallel across tasks:
//   kLitInline  re-run them inside the batch (function-level probes)
//   kLitDefer   leave them out (no peaks) and report them in `deferred`; the record collects them ...
//   kLitOnly    ... and runs them together in batches of this kind: only the pairs that carry a task of `only_tasks`
//               are scanned, and the host phase keeps the rows of exactly those tasks
enum LitMode { kLitInline, kLitDefer, kLitOnly, kLitSkip /* thresholds only (-F mode): the Q4 flags are ignored */ };

constexpr int kSideCap = 4096;      // literal scan jobs a call may park on the side stream

int run_batch_device(ltg_context* c, const std::vector<HostSeg>& segs, HostBatch& hb, bool want_alignments, ProbeOut* probe,
                     LitMode lit_mode = kLitInline, std::vector<std::pair<int, int>>* deferred = nullptr,
                     const std::vector<int>* only_tasks = nullptr, const std::vector<int>* only_rows = nullptr)
{
    const int T = (int)c->tasks.size(), P = (int)c->pairs.size();
    const int S = (int)segs.size();
    int max_len = 1;
    bool any_stats = !c->rna_plain;
    for (const HostSeg& s : segs) { max_len = std::max(max_len, s.len); if (s.flags & kSegNonACGT) any_stats = true; }
    max_len = (max_len + 3) & ~3;
    // scan items: every pair of every segment, or (kLitOnly) just the pairs that carry a requested task
    std::vector<ScanItem> items;
    if (lit_mode == kLitOnly) {
        std::vector<char> want((size_t)S * P, 0);
        for (int t : *only_tasks) want[(size_t)(t / T) * P + c->tasks[t % T].pair] = 1;
        for (int s = 0; s < S; ++s) for (int p = 0; p < P; ++p) if (want[(size_t)s * P + p]) { ScanItem it; it.seg = s; it.pair = p; items.push_back(it); }
    } else {
        items.resize((size_t)S * P);
        for (int s = 0; s < S; ++s) for (int p = 0; p < P; ++p) { items[(size_t)s * P + p].seg = s; items[(size_t)s * P + p].pair = p; }
    }
    const int n_items = (int)items.size(), n_tasks = S * T;
    hb.segs = segs; hb.n_tasks = n_tasks; hb.n_peaks = 0; hb.rows.clear(); hb.error.clear();
    hb.keep_only_literal = (lit_mode == kLitOnly);
    hb.timed = true; hb.timed_windows = false; hb.has_stats_pass = any_stats; hb.n_literal_tasks = 0;
    hb.n_ctasks = 0; hb.n_cpeaks = 0; hb.d2h_late = 0;

    if (int e = c->d_segs.ensure(sizeof(SegDesc) * S)) return e;
    if (int e = c->d_items.ensure(sizeof(ScanItem) * n_items)) return e;
    const int n_gran = c->n_strips * (32 * c->scan_r / kGranRows);
    const int blk_pitch = blk_pitch_for(max_len);
    if (int e = c->d_blkmax.ensure((size_t)n_items * n_gran * blk_pitch * sizeof(uint16_t))) return e;
    if (int e = c->d_colmax_all.ensure((size_t)n_items * max_len * 4)) return e;
    if (int e = c->d_task_info.ensure(sizeof(int) * 5 * (size_t)n_tasks)) return e;
    if (int e = c->d_task_off.ensure(sizeof(int) * ((size_t)n_tasks + 1))) return e;
    if (int e = c->d_stats_max.ensure(sizeof(int) * (size_t)n_tasks)) return e;
    if (int e = c->d_task_litrow.ensure(sizeof(int) * (size_t)n_tasks)) return e;
    if (int e = hb.task_info.ensure(sizeof(int) * 5 * (size_t)n_tasks)) return e;
    if (int e = hb.task_off.ensure(sizeof(int) * ((size_t)n_tasks + 1))) return e;
    if (int e = hb.scalars.ensure(sizeof(BatchScalars))) return e;
    BatchScalars* bs = hb.scalars.as<BatchScalars>();
    TaskInfo ti(c->d_task_info.as<int>(), n_tasks), hti(hb.task_info.as<int>(), n_tasks);
    int* counters = c->d_counters.as<int>();

    std::vector<SegDesc> hs(S);
    for (int s = 0; s < S; ++s) { hs[s].start = segs[s].start; hs[s].len = segs[s].len; hs[s].flags = segs[s].flags; }
    // tasks without an item in this batch (kLitOnly) must read as "no peaks, no flags"
    if (lit_mode == kLitOnly) LTG_CUDA_CHECK(cudaMemsetAsync(c->d_task_info.p, 0, sizeof(int) * 5 * (size_t)n_tasks, c->stream));
    // (pageable sources: these copies return once the data is staged, so the vectors may die at scope exit)
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_segs.p, hs.data(), sizeof(SegDesc) * S, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_items.p, items.data(), sizeof(ScanItem) * n_items, cudaMemcpyHostToDevice, c->stream));

    LTG_CUDA_CHECK(cudaEventRecord(hb.ev[0], c->stream));
    // optional N/U-aware threshold pass (Q3): exact maxima under the Farrar-side scoring.  With a plain (ACGT) lncRNA only
    // the segments that contain a byte outside ACGT need it; with U / N in the lncRNA every segment does.
    const int* stats_max = nullptr;
    if (any_stats) {
        std::vector<ScanItem> sitems;
        if (c->rna_plain) { for (const ScanItem& it : items) if (segs[it.seg].flags & kSegNonACGT) sitems.push_back(it); }
        else sitems = items;
        const int ns = (int)sitems.size();
        if (int e = c->d_items_stats.ensure(sizeof(ScanItem) * (size_t)std::max(ns, 1))) return e;
        LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_items_stats.p, sitems.data(), sizeof(ScanItem) * (size_t)ns, cudaMemcpyHostToDevice, c->stream));
        if (int e = launch_scan(c, c->d_items_stats.as<ScanItem>(), ns, max_len, c->d_prof_stats.as<uint32_t>(), c->d_colmax_all.as<uint32_t>(),
                                c->d_blkmax.as<uint16_t>())) return e;
        k_rowmax<<<(ns * 32 + 127) / 128, 128, 0, c->stream>>>(c->d_colmax_all.as<uint32_t>(), c->d_items_stats.as<ScanItem>(), c->d_segs.as<SegDesc>(),
                                                              ns, 1, max_len, T, c->d_stats_max.as<int>());
        c->launches += 1;
        stats_max = c->d_stats_max.as<int>();
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));          // `sitems` is host memory of this scope
    }
    // fused carried-F recording (see ltg_context::frec_mode): this batch's verdict on the Q4 quirk comes from the main sweep
    const bool lit_forced = (lit_mode == kLitOnly);     // literal-only batch: the literal tasks are the requested ones
    const bool use_frec = c->q4_probe && c->frec_on && lit_mode != kLitSkip && lit_mode != kLitOnly;
    uint16_t* frec = nullptr;
    if (use_frec) {
        if (int e = c->d_frec.ensure((size_t)n_items * kFrecRows * blk_pitch * sizeof(uint16_t))) return e;
        frec = c->d_frec.as<uint16_t>();
    }
    LTG_CUDA_CHECK(cudaEventRecord(hb.ev[4], c->stream));
    if (int e = launch_scan(c, c->d_items.as<ScanItem>(), n_items, max_len, c->d_prof_ssw.as<uint32_t>(), c->d_colmax_all.as<uint32_t>(),
                            c->d_blkmax.as<uint16_t>(), nullptr, nullptr, &items, frec)) return e;
    LTG_CUDA_CHECK(cudaEventRecord(hb.ev[5], c->stream));

    // peaks: statistics + count, literal re-runs, count of those, exclusive scan, write
    EpiArgs ea;
    ea.blkmax = c->d_blkmax.as<uint16_t>(); ea.n_gran = n_gran; ea.blk_pitch = blk_pitch; ea.scan_r = c->scan_r;
    ea.colmax_all = c->d_colmax_all.as<uint32_t>(); ea.lit_colmax = nullptr; ea.task_litrow = c->d_task_litrow.as<int>();
    ea.items = c->d_items.as<ScanItem>(); ea.segs = c->d_segs.as<SegDesc>();
    ea.item_orig = nullptr; ea.probe = nullptr;
    ea.bnd_gran = (c->q4_probe && c->n_bnd_gran > 0 && !lit_forced) ? c->d_bnd_gran.as<int>() : nullptr; ea.n_bnd_gran = c->n_bnd_gran;
    ea.frec = frec; ea.stripe_len = (c->m + 15) / 16; ea.taint_colmax = nullptr;
    ea.n_items = n_items; ea.max_len = max_len; ea.tasks_per_seg = T; ea.stats_max = stats_max; ea.stats_all = c->rna_plain ? 0 : 1; ea.mode = 0;
    ea.task_max = ti.max; ea.task_thr = ti.thr; ea.task_npeaks = ti.npk; ea.task_flags = ti.flags; ea.task_jstar = ti.jstar;
    ea.task_off = c->d_task_off.as<int>(); ea.pk_task = nullptr; ea.pk_pos = nullptr; ea.pk_score = nullptr;
    const int epi_blocks = (n_items * 32 + 127) / 128;
    k_epilogue<<<epi_blocks, 128, 0, c->stream>>>(ea);
    c->launches += 1;
    LTG_CUDA_CHECK(cudaGetLastError());
    if (lit_forced) {
        // the verdict on these tasks was reached in their main batch (by the probe sweep or by the recording sweep, whose
        // block granularity may flag a few tasks more): it is not taken again, the requested tasks ARE the literal ones
        std::vector<unsigned char> want((size_t)n_tasks, 0);
        for (int t : *only_tasks) want[(size_t)t] = 1;
        if (int e = c->d_probe_orig.ensure((size_t)n_tasks)) return e;
        LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_probe_orig.p, want.data(), (size_t)n_tasks, cudaMemcpyHostToDevice, c->stream));
        k_force_literal<<<(n_tasks + 255) / 256, 256, 0, c->stream>>>(ti.flags, ti.npk, c->d_probe_orig.as<unsigned char>(), n_tasks);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));          // `want` is host memory of this scope
        c->h2d_bytes += n_tasks;
    }

    // Q4 guard: tasks whose exact maximum could carry F >= 132 across a stripe boundary are re-run through
    // the literal striped emulation, their peaks come from the literal column maxima
    LTG_CUDA_CHECK(cudaMemcpyAsync(hti.flags, ti.flags, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += sizeof(int) * (int64_t)n_tasks;
    // Q4 probe: a second sweep of the pairs that carry a flagged task finds the largest F carried into a stripe start of the
    // reference's layout; below 132 the signed compare cannot misfire and the task returns to the exact path (mode 3)
    if (c->q4_probe && lit_mode != kLitSkip && !lit_forced && (c->q4_taint || !use_frec)) {
        std::vector<ScanItem> pitems;
        std::vector<int> porig;
        for (int i = 0; i < n_items; ++i) {
            const PairDef& pd = c->pairs[items[i].pair];
            const int base = items[i].seg * T;
            if ((hti.flags[base + pd.task[0]] | hti.flags[base + pd.task[1]]) & kTaskLiteral) { pitems.push_back(items[i]); porig.push_back(i); }
        }
        if (!pitems.empty()) {
            const int np = (int)pitems.size();
            if (int e = c->d_probe_items.ensure(sizeof(ScanItem) * (size_t)np)) return e;
            if (int e = c->d_probe_orig.ensure(sizeof(int) * (size_t)np)) return e;
            if (int e = c->d_probe_out.ensure(sizeof(uint32_t) * (size_t)np)) return e;
            LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_probe_items.p, pitems.data(), sizeof(ScanItem) * (size_t)np, cudaMemcpyHostToDevice, c->stream));
            LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_probe_orig.p, porig.data(), sizeof(int) * (size_t)np, cudaMemcpyHostToDevice, c->stream));
            EpiArgs pa = ea;
            pa.items = c->d_probe_items.as<ScanItem>(); pa.item_orig = c->d_probe_orig.as<int>(); pa.probe = c->d_probe_out.as<uint32_t>();
            pa.n_items = np;
            if (c->q4_taint) {
                // taint sweep: exact Smith-Waterman once more over these pairs, carrying one bit per value that says whether the
                // quirk could have lowered it; a task whose recorded column maxima are all untainted returns to the exact path
                if (int e = c->d_taint_colmax.ensure((size_t)np * max_len * 4)) return e;
                if (int e = launch_scan(c, c->d_probe_items.as<ScanItem>(), np, max_len, c->d_prof_ssw2.as<uint32_t>(), c->d_taint_colmax.as<uint32_t>(),
                                        nullptr, c->d_probe_out.as<uint32_t>(), ti.jstar, nullptr, nullptr, true, ti.flags)) return e;
                pa.taint_colmax = c->d_taint_colmax.as<uint32_t>();
                pa.mode = 4;
            } else {
                if (int e = launch_scan(c, c->d_probe_items.as<ScanItem>(), np, max_len, c->d_prof_ssw.as<uint32_t>(), nullptr, nullptr,
                                        c->d_probe_out.as<uint32_t>(), ti.jstar)) return e;
                pa.mode = 3;
            }
            k_epilogue<<<(np * 32 + 127) / 128, 128, 0, c->stream>>>(pa);
            c->launches += 1;
            LTG_CUDA_CHECK(cudaGetLastError());
            LTG_CUDA_CHECK(cudaMemcpyAsync(hti.flags, ti.flags, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
            LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));      // (also: `pitems` / `porig` are host memory of this scope)
            c->d2h_bytes += sizeof(int) * (int64_t)n_tasks;
            c->h2d_bytes += (int64_t)(sizeof(ScanItem) + sizeof(int)) * np;
            c->n_probe_items += np;
            // np of n_items pairs were swept again: when that is more than a few percent, later batches of this query record the
            // stripe-start screen in the main sweep (kLitOnly batches and tiny ones say little about the query: at least 64 items)
            if (c->frec_mode < 0 && !use_frec && n_items >= 64 && (double)np > 0.04 * (double)n_items) c->frec_on = true;
        }
    }
    // kLitOnly: were all requested tasks already swept by the side stream (their literal column maxima wait in d_side_colmax)?
    bool from_side = (lit_mode == kLitOnly && only_rows && !only_rows->empty());
    if (from_side) for (int r : *only_rows) if (r < 0) from_side = false;
    std::vector<LiteralJob> jobs, side_jobs;
    for (int t = 0; t < n_tasks; ++t) {
        if (hti.flags[t] & kTaskRange) { set_error("alignment score exceeds the 16-bit range (segment too long for this build)"); return LTG_ERR_LIMIT; }
        if ((hti.flags[t] & kTaskLiteral) && lit_mode != kLitSkip) {
            LiteralJob j; memset(&j, 0, sizeof j);
            j.kind = 0; j.task = t; j.seg = t / T; j.tdef = t % T; j.ref_start = 0; j.ref_len = segs[t / T].len;
            j.read_start = 0; j.read_len = c->m; j.read_dir = 1; j.ref_dir = 0; j.terminate = 255; j.peak = -1;
            j.seg_start = segs[t / T].start; j.seg_len = segs[t / T].len;
            if (lit_mode == kLitDefer) {
                // start the (slow, single half-warp) literal sweep of this task right away on the side stream; the call
                // collects its column maxima at the end
                int row = -1;
                if (c->side_pitch > 0 && c->side_rows + (int)side_jobs.size() < kSideCap) { row = c->side_rows + (int)side_jobs.size(); side_jobs.push_back(j); }
                deferred->push_back(std::make_pair(t, row));
                continue;
            }
            jobs.push_back(j);
        }
    }
    if (!side_jobs.empty()) {
        if (int e = c->d_side_jobs.ensure(sizeof(LiteralJob) * kSideCap)) return e;
        if (int e = c->d_side_colmax.ensure(sizeof(uint16_t) * (size_t)kSideCap * c->side_pitch)) return e;
        LiteralJob* dj = c->d_side_jobs.as<LiteralJob>() + c->side_rows;
        LTG_CUDA_CHECK(cudaMemcpyAsync(dj, side_jobs.data(), sizeof(LiteralJob) * side_jobs.size(), cudaMemcpyHostToDevice, c->lit_stream));
        if (int e = launch_literal(c, dj, (int)side_jobs.size(), nullptr, c->m, nullptr, c->side_pitch, /*side=*/true, c->side_rows)) return e;
        LTG_CUDA_CHECK(cudaEventRecord(c->lit_event, c->lit_stream));
        c->side_rows += (int)side_jobs.size();
        c->h2d_bytes += (int64_t)(sizeof(LiteralJob) * side_jobs.size());
    }
    hb.n_literal_tasks = (int)jobs.size();
    ea.lit_pitch = max_len;
    if (from_side) {
        // the literal sweeps ran on the side stream: point the tasks at their rows and wait for that stream
        std::vector<int> litrow(n_tasks, -1);
        for (size_t k = 0; k < only_tasks->size(); ++k) litrow[(*only_tasks)[k]] = (*only_rows)[k];
        for (const LiteralJob& j : jobs) if (litrow[j.task] < 0) { set_error("literal task without a side-stream row"); return LTG_ERR_STATE; }
        LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_task_litrow.p, litrow.data(), sizeof(int) * (size_t)n_tasks, cudaMemcpyHostToDevice, c->stream));
        LTG_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->lit_event, 0));
        ea.lit_colmax = c->d_side_colmax.as<uint16_t>(); ea.lit_pitch = c->side_pitch;
        ea.mode = 1;
        k_epilogue<<<epi_blocks, 128, 0, c->stream>>>(ea);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));      // `litrow` is host memory of this scope
    } else if (!jobs.empty()) {
        if (int e = c->d_lit_jobs.ensure(sizeof(LiteralJob) * jobs.size())) return e;
        if (int e = c->d_lit_colmax.ensure(sizeof(uint16_t) * jobs.size() * (size_t)max_len)) return e;
        LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_lit_jobs.p, jobs.data(), sizeof(LiteralJob) * jobs.size(), cudaMemcpyHostToDevice, c->stream));
        if (int e = launch_literal(c, c->d_lit_jobs.as<LiteralJob>(), (int)jobs.size(), nullptr, c->m, nullptr, max_len)) return e;
        ea.lit_colmax = c->d_lit_colmax.as<uint16_t>();
        ea.mode = 1;
        k_epilogue<<<epi_blocks, 128, 0, c->stream>>>(ea);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
    }
    k_exclusive_scan<<<1, 1024, 0, c->stream>>>(ti.npk, c->d_task_off.as<int>(), n_tasks, counters + kCntPeaks);
    c->launches += 1;
    int n_peaks = 0;
    LTG_CUDA_CHECK(cudaMemcpyAsync(&n_peaks, counters + kCntPeaks, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));      // (also: `jobs` is host memory of this scope)
    hb.n_peaks = n_peaks;
    if (n_peaks > 0) {
        for (DevBuf* b : {&c->d_pk_task, &c->d_pk_pos, &c->d_pk_score}) if (int e = b->ensure(sizeof(int) * (size_t)n_peaks)) return e;
        ea.pk_task = c->d_pk_task.as<int>(); ea.pk_pos = c->d_pk_pos.as<int>(); ea.pk_score = c->d_pk_score.as<int>();
        ea.mode = 2;
        k_epilogue<<<epi_blocks, 128, 0, c->stream>>>(ea);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
    }
    LTG_CUDA_CHECK(cudaEventRecord(hb.ev[1], c->stream));

    LTG_CUDA_CHECK(cudaMemcpyAsync(hb.task_info.p, c->d_task_info.p, sizeof(int) * 5 * (size_t)n_tasks, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(hb.task_off.p, c->d_task_off.p, sizeof(int) * (size_t)n_tasks, cudaMemcpyDeviceToHost, c->stream));
    c->d2h_bytes += sizeof(int) * 6 * (int64_t)n_tasks;
    if (probe) {
        probe->max_len = max_len;
        probe->pk_pos.resize(n_peaks); probe->pk_score.resize(n_peaks);
        if (n_peaks) {
            LTG_CUDA_CHECK(cudaMemcpyAsync(probe->pk_pos.data(), c->d_pk_pos.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
            LTG_CUDA_CHECK(cudaMemcpyAsync(probe->pk_score.data(), c->d_pk_score.p, sizeof(int) * n_peaks, cudaMemcpyDeviceToHost, c->stream));
        }
        probe->colmax.resize((size_t)n_items * max_len);
        LTG_CUDA_CHECK(cudaMemcpyAsync(probe->colmax.data(), c->d_colmax_all.p, probe->colmax.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        probe->lit_colmax.resize(jobs.size() * (size_t)max_len);
        probe->task_litrow.assign(n_tasks, -1);
        if (!jobs.empty()) {
            LTG_CUDA_CHECK(cudaMemcpyAsync(probe->lit_colmax.data(), c->d_lit_colmax.p, probe->lit_colmax.size() * 2, cudaMemcpyDeviceToHost, c->stream));
            LTG_CUDA_CHECK(cudaMemcpyAsync(probe->task_litrow.data(), c->d_task_litrow.p, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, c->stream));
        }
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        probe->task_max.assign(hti.max, hti.max + n_tasks); probe->task_thr.assign(hti.thr, hti.thr + n_tasks);
        probe->task_npk.assign(hti.npk, hti.npk + n_tasks); probe->task_flags.assign(hti.flags, hti.flags + n_tasks);
        probe->task_off.assign(hb.task_off.as<int>(), hb.task_off.as<int>() + n_tasks);
    }
    if (want_alignments && n_peaks > 0) {
        LTG_CUDA_CHECK(cudaEventRecord(hb.ev[2], c->stream));
        if (int e = run_windows(c, n_peaks, T, nullptr, lit_mode == kLitOnly ? nullptr : c->d_blkmax.as<uint16_t>(), max_len, &hb, true)) return e;
        if (int e = run_traceback(c, c->d_jobs.as<TraceJob>(), n_peaks, c->d_tout.as<TraceOut>(), nullptr, true)) return e;
        LTG_CUDA_CHECK(cudaEventRecord(hb.ev[3], c->stream));
        hb.timed_windows = true;
        // ship only the tasks that own a candidate row (k_task_flag): flags -> two exclusive scans -> compaction into this
        // slot's device buffers; the worker thread copies the (few) records on the copy stream once the counts are known
        if (int e = hb.jobs.ensure(sizeof(TraceJob) * (size_t)n_peaks)) return e;
        if (int e = hb.tout.ensure(sizeof(TraceOut) * (size_t)n_peaks)) return e;
        if (int e = hb.c_task.ensure(sizeof(int) * (size_t)n_tasks)) return e;
        if (int e = hb.c_poff.ensure(sizeof(int) * (size_t)n_tasks)) return e;
        if (int e = hb.d_cjobs.ensure(sizeof(TraceJob) * (size_t)n_peaks)) return e;
        if (int e = hb.d_ctout.ensure(sizeof(TraceOut) * (size_t)n_peaks)) return e;
        if (int e = hb.d_ctask.ensure(sizeof(int) * (size_t)n_tasks)) return e;
        if (int e = hb.d_cpoff.ensure(sizeof(int) * (size_t)n_tasks)) return e;
        if (int e = c->d_cand.ensure(sizeof(int) * 4 * (size_t)n_tasks)) return e;
        CompactArgs ca;
        ca.jobs = c->d_jobs.as<TraceJob>(); ca.tout = c->d_tout.as<TraceOut>(); ca.task_off = c->d_task_off.as<int>();
        ca.n_tasks = n_tasks; ca.n_peaks = n_peaks;
        ca.need_nt = std::max(c->params.nt_min, c->params.c_length);
        ca.min_id = (float)c->params.min_identity; ca.min_st = (float)c->params.min_stability;
        int* cand = c->d_cand.as<int>();
        ca.cand_cnt = cand; ca.cand_flag = cand + n_tasks; ca.cand_poff = cand + 2 * (size_t)n_tasks; ca.cand_toff = cand + 3 * (size_t)n_tasks;
        ca.c_jobs = hb.d_cjobs.as<TraceJob>(); ca.c_tout = hb.d_ctout.as<TraceOut>(); ca.c_task = hb.d_ctask.as<int>(); ca.c_poff = hb.d_cpoff.as<int>();
        ca.n_unfinished = counters + kCntCand + 2;
        ca.filt = getenv("LTG_FILTER_STATS") ? &c->d_win_sched.as<WinSched>()->st_filter[0] : nullptr;
        LTG_CUDA_CHECK(cudaMemsetAsync(counters + kCntCand, 0, 3 * sizeof(int), c->stream));
        const int task_blocks = (n_tasks * 32 + 255) / 256;
        k_task_flag<<<task_blocks, 256, 0, c->stream>>>(ca);
        k_exclusive_scan<<<1, 1024, 0, c->stream>>>(ca.cand_cnt, cand + 2 * (size_t)n_tasks, n_tasks, counters + kCntCand);
        k_exclusive_scan<<<1, 1024, 0, c->stream>>>(ca.cand_flag, cand + 3 * (size_t)n_tasks, n_tasks, counters + kCntCand + 1);
        k_task_compact<<<task_blocks, 256, 0, c->stream>>>(ca);
        c->launches += 4;
        LTG_CUDA_CHECK(cudaGetLastError());
        LTG_CUDA_CHECK(cudaMemcpyAsync(hb.c_task.p, hb.d_ctask.p, sizeof(int) * (size_t)n_tasks, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(hb.c_poff.p, hb.d_cpoff.p, sizeof(int) * (size_t)n_tasks, cudaMemcpyDeviceToHost, c->stream));
        static_assert(kCntLit - kCntCells == 8 && kCntLitTotal - kCntLit == 2, "BatchScalars mirrors counters[kCntCells .. kCntCells + 16)");
        LTG_CUDA_CHECK(cudaMemcpyAsync(&bs->window_cells, counters + kCntCells, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(bs->win_stats, &c->d_win_sched.as<WinSched>()->st_windows[0], 20 * sizeof(unsigned long long),
                                       cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(bs->trace_handed, counters + kCntOvf, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(bs->cand, counters + kCntCand, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(bs->filter_stats, &c->d_win_sched.as<WinSched>()->st_filter[0], 6 * sizeof(unsigned long long),
                                       cudaMemcpyDeviceToHost, c->stream));
        c->d2h_bytes += (int64_t)sizeof(int) * 2 * n_tasks + 272;
    } else {
        hb.n_peaks = want_alignments ? n_peaks : 0;
        memset(bs, 0, sizeof(BatchScalars));
    }
    LTG_CUDA_CHECK(cudaEventRecord(hb.ready, c->stream));
    return LTG_OK;
}

// Host phase of one batch (fastSIM's tail, fastsim.h:253-288, per task in the reference's task order) over the tasks the
// device shipped (those owning at least one candidate row): runs on `threads` worker threads over contiguous ranges of
// shipped tasks; the per-thread outputs are concatenated in order.
void host_phase(const ltg_context* c, HostBatch& hb, int threads)
{
    const int T = (int)c->tasks.size();
    const int n_ct = hb.n_ctasks, n_cp = hb.n_cpeaks;
    hb.rows.clear();
    if (n_ct == 0) return;
    const int* ctask = hb.c_task.as<int>();
    const int* cpoff = hb.c_poff.as<int>();
    const TraceJob* jobs = hb.jobs.as<TraceJob>();
    const TraceOut* tout = hb.tout.as<TraceOut>();
    const int* task_flags = TaskInfo(hb.task_info.as<int>(), hb.n_tasks).flags;
    threads = std::max(1, std::min(threads, n_ct / 64 + 1));
    std::vector<std::vector<ltg_host::Triplex>> part(threads);
    auto work = [&](int k) {
        const int i0 = (int)((long long)n_ct * k / threads), i1 = (int)((long long)n_ct * (k + 1) / threads);
        std::vector<ltg_host::Triplex> mine;
        for (int idx = i0; idx < i1; ++idx) {
            const int task = ctask[idx];
            if (hb.keep_only_literal && !(task_flags[task] & kTaskLiteral)) continue;
            const int b = cpoff[idx], e = (idx + 1 < n_ct) ? cpoff[idx + 1] : n_cp;
            mine.clear();
            const HostSeg& sg = hb.segs[task / T];
            const TaskDef& td = c->tasks[task % T];
            for (int i = b; i < e; ++i) {                       // peaks of a task come in ascending column order
                const TraceJob& J = jobs[i];
                if (J.score <= 0) continue;                     // fastsim.h:253 (sw_score == 0 -> skipped)
                if (tout[i].status != 1 && tout[i].status != 4) continue;       // banded_sw failed -> sw_score 0 (ssw_cpp.cpp:627-633)
                ltg_host::DeviceAlignment al;
                al.sw_score = J.score; al.ws = J.ws; al.shift = J.tdef >> 8; al.rb = J.rb; al.re = J.re; al.query_begin = J.qb; al.query_end = J.qe;
                al.nt = tout[i].nt; al.identity = tout[i].identity; al.tri_score = tout[i].tri;
                // dead alignment (window.cuh DeadRule): de-duplicates like any other, fails every output filter
                if (tout[i].status == 4) { al.identity = -INFINITY; al.tri_score = -INFINITY; }
                ltg_host::make_triplex(al, task % T, sg.len, (long)sg.start, (long)sg.coord, sg.record, td.para, td.strand, td.rule, c->params, mine);
            }
            if (!mine.empty()) ltg_host::finish_task(mine, c->params, part[k], c->compat);
        }
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int k = 1; k < threads; ++k) pool.emplace_back(work, k);
        work(0);
        for (std::thread& t : pool) t.join();
    }
    size_t total = 0;
    for (auto& v : part) total += v.size();
    hb.rows.reserve(total);
    for (int k = 0; k < threads; ++k) hb.rows.insert(hb.rows.end(), part[k].begin(), part[k].end());
}

// Worker of a batch slot: waits for the device phase, fetches the shipped records on the copy stream (concurrent with
// the next batch's kernels), runs the host phase.
void batch_worker(ltg_context* c, HostBatch* hb)
{
    cudaSetDevice(c->device);
    if (cudaEventSynchronize(hb->ready) != cudaSuccess) { hb->error = "device phase failed"; return; }
    if (hb->n_peaks == 0) return;
    const int* cand = hb->scalars.as<BatchScalars>()->cand;
    hb->n_cpeaks = cand[0]; hb->n_ctasks = cand[1];
    if (cand[2] > 0) { hb->error = "traceback band exceeds the device scratch"; return; }
    if (hb->n_cpeaks > 0) {
        cudaError_t e1 = cudaMemcpyAsync(hb->jobs.p, hb->d_cjobs.p, sizeof(TraceJob) * (size_t)hb->n_cpeaks, cudaMemcpyDeviceToHost, c->copy_stream);
        cudaError_t e2 = cudaMemcpyAsync(hb->tout.p, hb->d_ctout.p, sizeof(TraceOut) * (size_t)hb->n_cpeaks, cudaMemcpyDeviceToHost, c->copy_stream);
        cudaError_t e3 = cudaStreamSynchronize(c->copy_stream);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { hb->error = "copy of the batch records failed"; return; }
        hb->d2h_late = (int64_t)(sizeof(TraceJob) + sizeof(TraceOut)) * hb->n_cpeaks;
    }
    host_phase(c, *hb, c->host_threads);
}

// cutSequence — fastsim.h:71-90.  A shard is the run of segments [first_seg, first_seg + n_seg) of a record of `record_len`
// bases whose bytes start at the first segment's start; segment starts are relative to the shard's buffer, lengths follow
// the RECORD's geometry (so the union of shards is exactly the record's segment list — no extra tail segments).
int cut_segments(int64_t buf_len, int64_t record_len, int64_t first_seg, int64_t n_seg, const ltg_params& P, int64_t base, int record,
                 std::vector<HostSeg>& segs)
{
    if (P.cut_length <= 0 || P.cut_length - P.overlap <= 0) { set_error("cut length must exceed the overlap"); return LTG_ERR_ARG; }
    const int64_t stride = P.cut_length - P.overlap;
    for (int64_t k = first_seg; (n_seg < 0 || k < first_seg + n_seg) && k * stride < record_len; ++k) {
        HostSeg s; s.start = base + (k - first_seg) * stride; s.len = (int32_t)std::min<int64_t>(P.cut_length, record_len - k * stride); s.flags = 0;
        s.coord = k * stride; s.record = record; s.pad_ = 0;
        if (s.start - base + s.len > buf_len) { set_error("shard buffer (%lld bytes) does not cover segment %lld", (long long)buf_len, (long long)k); return LTG_ERR_ARG; }
        segs.push_back(s);
    }
    return LTG_OK;
}

struct ResultBuilder {
    std::vector<ltg_triplex> tri;
    std::string text;
    std::vector<int64_t> chr_off;       // per record: offset of its chromosome tag in `text` (-1: not stored yet)
    // text_len: columns of the alignment strings (= t.nt for fastSIM rows; SIM rows report the lncRNA span as Nt(bp), sim.h:589)
    void add(const ltg_host::Triplex& t, const char* tfo, const char* tts, const char* chr, int64_t record_start, int record, int64_t text_len = -1)
    {
        if (text_len < 0) text_len = t.nt;
        if ((size_t)record >= chr_off.size()) chr_off.resize((size_t)record + 1, -1);
        if (chr_off[record] < 0) { chr_off[record] = (int64_t)text.size(); text += (chr ? chr : ""); text += '\0'; }
        ltg_triplex o;
        memset(&o, 0, sizeof o);
        o.stari = t.stari; o.endi = t.endi; o.starj = t.starj; o.endj = t.endj;
        o.reverse = t.reverse; o.strand = t.strand;
        o.rule = t.rule; o.nt = t.nt; o.score = t.score; o.identity = t.identity; o.tri_score = t.tri_score;
        o.genomestart = o.starj + record_start - 1;         // Fasim-LongTarget.cpp:146-147
        o.genomeend = o.endj + record_start - 1;
        o.tfo_off = (int64_t)text.size(); text.append(tfo, (size_t)text_len); text += '\0';
        o.tts_off = (int64_t)text.size(); text.append(tts, (size_t)text_len); text += '\0';
        o.chr_off = chr_off[record];
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

// traceback pass 2: TFO / TTS strings of `rows` (alignments of the record resident in d_dna / d_codes)
int fetch_strings(ltg_context* c, const std::vector<ltg_host::Triplex>& rows, std::vector<char>& pool, std::vector<int64_t>& offs)
{
    const size_t n = rows.size();
    offs.resize(n);
    pool.clear();
    if (n == 0) return LTG_OK;
    std::vector<TraceJob> jobs(n);
    int64_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        const ltg_host::Triplex& t = rows[i];
        TraceJob& J = jobs[i];
        J.seg_start = t.seg_start; J.seg_len = t.seg_len; J.tdef = t.tdef | (t.shift << 8); J.ws = t.ws; J.rb = t.rb; J.re = t.re; J.qb = t.qb; J.qe = t.qe;
        J.score = (int)t.score; J.out_off = total;
        offs[i] = total;
        total += 2 * ((int64_t)t.nt + 1);
    }
    if (int e = c->d_jobs.ensure(sizeof(TraceJob) * n)) return e;
    if (int e = c->d_tout.ensure(sizeof(TraceOut) * n)) return e;
    if (int e = c->d_strpool.ensure((size_t)total)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_jobs.p, jobs.data(), sizeof(TraceJob) * n, cudaMemcpyHostToDevice, c->stream));
    if (int e = run_traceback(c, c->d_jobs.as<TraceJob>(), (int)n, c->d_tout.as<TraceOut>(), c->d_strpool.as<char>())) return e;
    pool.resize((size_t)total);
    std::vector<TraceOut> tout(n);
    LTG_CUDA_CHECK(cudaMemcpyAsync(pool.data(), c->d_strpool.p, (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(tout.data(), c->d_tout.p, sizeof(TraceOut) * n, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->h2d_bytes += (int64_t)(sizeof(TraceJob) * n);
    c->d2h_bytes += total + (int64_t)(sizeof(TraceOut) * n);
    for (size_t i = 0; i < n; ++i)
        if (tout[i].status != 1 || tout[i].nt != rows[i].nt) { set_error("string pass disagrees with the first traceback pass (row %zu)", i); return LTG_ERR_STATE; }
    return LTG_OK;
}

struct RecordStats {
    int64_t n_segments = 0, n_tasks = 0, scan_cells = 0, n_peaks = 0, window_cells = 0, n_literal_tasks = 0, n_literal_windows = 0;
    double ms_scan = 0, ms_window = 0, ms_scan_kernel = 0;
    int64_t n_scan_launches = 0;
};

// closes a batch slot: waits for its host phase, collects timings / counters / rows
int retire_batch(ltg_context* c, HostBatch& hb, RecordStats& st, std::vector<ltg_host::Triplex>& record_list)
{
    if (hb.worker.joinable()) hb.worker.join();
    if (!hb.timed) return LTG_OK;
    hb.timed = false;
    if (!hb.error.empty()) { set_error("%s", hb.error.c_str()); return LTG_ERR_LIMIT; }
    float ms = 0;
    LTG_CUDA_CHECK(cudaEventSynchronize(hb.ready));
    float ms01 = 0, ms45 = 0, ms23 = 0, ms12 = 0;
    LTG_CUDA_CHECK(cudaEventElapsedTime(&ms01, hb.ev[0], hb.ev[1])); st.ms_scan += ms01;
    LTG_CUDA_CHECK(cudaEventElapsedTime(&ms45, hb.ev[4], hb.ev[5])); st.ms_scan_kernel += ms45;
    if (hb.timed_windows) {
        LTG_CUDA_CHECK(cudaEventElapsedTime(&ms23, hb.ev[2], hb.ev[3])); st.ms_window += ms23;
        LTG_CUDA_CHECK(cudaEventElapsedTime(&ms12, hb.ev[1], hb.ev[2]));
    }
    if (getenv("LTG_TIMING"))
        fprintf(stderr, "[ltg batch] %zu segs, %d peaks: scan+peaks %.1f ms (k_scan %.1f), gap %.1f, windows+traceback %.1f, shipped %d tasks / %d records\n",
                hb.segs.size(), hb.n_peaks, ms01, ms45, ms12, ms23, hb.n_ctasks, hb.n_cpeaks);
    (void)ms;
    st.n_scan_launches += 1;
    const int T = (int)c->tasks.size();
    st.n_segments += (int64_t)hb.segs.size();
    st.n_tasks += (int64_t)hb.segs.size() * T;
    for (const HostSeg& sg : hb.segs) st.scan_cells += (int64_t)sg.len * c->m * T;
    st.n_peaks += hb.n_peaks;
    c->d2h_bytes += hb.d2h_late;
    st.n_literal_tasks += hb.n_literal_tasks;
    const BatchScalars* bs = hb.scalars.as<BatchScalars>();
    st.window_cells += bs->window_cells;
    st.n_literal_windows += bs->lit_total;
    if (c->win_q4_mode < 0 && bs->lit_total > 64) c->win_q4_on = true;        // later batches of this query check their windows
    for (int k = 0; k < 20; ++k) c->win_stats[k] += bs->win_stats[k];
    for (int k = 0; k < 4; ++k) c->win_stats[20 + k] += (unsigned long long)bs->trace_handed[k];
    for (int k = 0; k < 6; ++k) c->win_stats[24 + k] += bs->filter_stats[k];
    record_list.insert(record_list.end(), hb.rows.begin(), hb.rows.end());
    hb.rows.clear();
    return LTG_OK;
}

// one DNA record (or shard of a record) of a scan call
struct RecordIn {
    const char* h_dna = nullptr;               // host bytes, or ...
    const unsigned char* d_dna = nullptr;      // ... device bytes
    int64_t len = 0;                           // readable bytes
    const char* chr = nullptr;
    int64_t record_start = 0;
    int64_t record_len = -1, first_seg = 0, n_seg = -1;     // shard geometry (ltg_scan_shard); defaults: the whole record
    // 2-bit packed input (ltg_scan_packed): `len` bases starting at base index packed_first of the packed bytes
    const unsigned char* packed = nullptr; int packed_on_device = 0; int64_t packed_first = 0;
    const uint32_t* n_start = nullptr; const uint32_t* n_size = nullptr; int n_blocks = 0;
};

// -F mode: every task goes through SIM() (sim.h:410) instead of fastSIM().  Per batch of segments: the scan stage yields the
// thresholds (calc_score_once * 0.8, Fasim-LongTarget.cpp:421), k_sim the alignments of every task (coordinates + edit scripts),
// and the host turns them into rows in the reference's order (segment, task, alignment) — sim_host.hpp.
constexpr int kSimBatchSegments = 64;
constexpr long long kSimPoolInts = 32LL << 20;           // 128 MB of alignment records + scripts per batch

int run_sim_batches(ltg_context* c, const std::vector<HostSeg>& active, const RecordIn* recs, ResultBuilder& rb, RecordStats& st)
{
    if (c->m >= simk::kMaxRows) { set_error("-F (SIM) mode supports lncRNAs up to %d nt", simk::kMaxRows - 1); return LTG_ERR_LIMIT; }
    if (c->params.cut_length >= (1 << simk::kColBits) - 2) { set_error("-F (SIM) mode supports cut lengths below %d", (1 << simk::kColBits) - 2); return LTG_ERR_LIMIT; }
    const int T = (int)c->tasks.size();
    int* counters = c->d_counters.as<int>();
    for (size_t b0 = 0; b0 < active.size(); b0 += kSimBatchSegments) {
        std::vector<HostSeg> batch(active.begin() + b0, active.begin() + std::min(active.size(), b0 + (size_t)kSimBatchSegments));
        HostBatch& hb = c->hb[0];
        const int rc = run_batch_device(c, batch, hb, false, nullptr, kLitSkip);
        hb.timed = false;
        if (rc != LTG_OK) return rc;
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const int S = (int)batch.size(), n_tasks = S * T;
        int max_len = 1;
        for (const HostSeg& sg : batch) max_len = std::max(max_len, sg.len);
        std::vector<int> ids(n_tasks);
        for (int t = 0; t < n_tasks; ++t) ids[t] = t;
        // one warp per task; few tasks (a single record of the demo's size) spread one warp per block over the SMs
        const int wpb = n_tasks <= c->num_sms * 8 ? 1 : 4;
        const int blocks = std::max(1, std::min(c->num_sms * (wpb == 1 ? 8 : 4), (n_tasks + wpb - 1) / wpb));
        const long long per_warp = sim_scratch_bytes(c->m, max_len);
        if (int e = c->d_sim_scratch.ensure((size_t)blocks * wpb * (size_t)per_warp)) return e;
        if (int e = c->d_sim_hdr.ensure(sizeof(SimHeader) * (size_t)n_tasks)) return e;
        if (int e = c->d_sim_pool.ensure(sizeof(int) * (size_t)kSimPoolInts)) return e;
        if (int e = c->d_sim_tasks.ensure(sizeof(int) * (size_t)n_tasks)) return e;
        LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_sim_tasks.p, ids.data(), sizeof(int) * (size_t)n_tasks, cudaMemcpyHostToDevice, c->stream));
        LTG_CUDA_CHECK(cudaMemsetAsync(counters + kCntScan, 0, 2 * sizeof(int), c->stream));        // queue head, pool fill
        SimArgs sa;
        sa.task_ids = c->d_sim_tasks.as<int>(); sa.n_tasks = n_tasks; sa.tasks_per_seg = T;
        sa.codes = c->d_codes.as<uint8_t>(); sa.segs = c->d_segs.as<SegDesc>();
        sa.rna_codes = c->d_rna_sim.as<uint8_t>(); sa.m = c->m;
        sa.task_thr = TaskInfo(c->d_task_info.as<int>(), n_tasks).thr;
        sa.scratch = c->d_sim_scratch.as<unsigned char>(); sa.scratch_per_warp = per_warp; sa.max_len = max_len;
        sa.counter = counters + kCntScan; sa.hdr = c->d_sim_hdr.as<SimHeader>();
        sa.pool = c->d_sim_pool.as<int>(); sa.pool_cap = (int)kSimPoolInts; sa.pool_used = counters + kCntPeaks;
        k_sim<<<blocks, 32 * wpb, 0, c->stream>>>(sa);
        c->launches += 1;
        LTG_CUDA_CHECK(cudaGetLastError());
        std::vector<SimHeader> hdr(n_tasks);
        int used = 0;
        LTG_CUDA_CHECK(cudaMemcpyAsync(hdr.data(), c->d_sim_hdr.p, sizeof(SimHeader) * (size_t)n_tasks, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(&used, counters + kCntPeaks, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        for (int t = 0; t < n_tasks; ++t)
            if (hdr[t].error) { set_error("SIM kernel: task %d failed with code %d (1 script pool, 2 used-cell pool, 3 alignment table, 4 diff stack, 5 batch pool)", t, hdr[t].error); return LTG_ERR_LIMIT; }
        used = std::min<long long>(used, kSimPoolInts);
        std::vector<int> pool((size_t)std::max(used, 1));
        // the batch's DNA bytes for the TTS strings (host copy of the device buffer: the input may be device resident)
        const int64_t lo = batch.front().start, hi = batch.back().start + batch.back().len;
        std::vector<char> dna((size_t)(hi - lo));
        if (used > 0) LTG_CUDA_CHECK(cudaMemcpyAsync(pool.data(), c->d_sim_pool.p, sizeof(int) * (size_t)used, cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaMemcpyAsync(dna.data(), c->d_dna.as<unsigned char>() + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, c->stream));
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->d2h_bytes += (int64_t)sizeof(SimHeader) * n_tasks + (int64_t)sizeof(int) * used + (hi - lo);
        c->h2d_bytes += (int64_t)sizeof(int) * n_tasks;
        std::vector<ltg_host::SimRow> rows;
        for (int t = 0; t < n_tasks; ++t) {
            const HostSeg& sg = batch[t / T];
            const TaskDef& td = c->tasks[t % T];
            rows.clear();
            for (int k = 0; k < hdr[t].n_aln; ++k) {
                simk::Aln al;
                memcpy(&al, pool.data() + hdr[t].aln_off + 8 * k, sizeof al);
                const int* script = pool.data() + hdr[t].aln_off + 8 * hdr[t].n_aln + al.script_off;
                ltg_host::sim_convert(al, script, c->rna.c_str(), td, dna.data() + (sg.start - lo), sg.len, (long)sg.coord, c->params, rows);
            }
            for (ltg_host::SimRow& r : rows) {
                if (!ltg_host::passes_record_filter(r.t, c->params)) continue;            // Fasim-LongTarget.cpp:589-597
                r.t.record = sg.record;
                rb.add(r.t, r.tfo.c_str(), r.tts.c_str(), recs[sg.record].chr, recs[sg.record].record_start, sg.record, (int64_t)r.tfo.size());
            }
        }
        st.n_segments += S; st.n_tasks += n_tasks;
        for (const HostSeg& sg : batch) st.scan_cells += (int64_t)sg.len * c->m * T;
    }
    return LTG_OK;
}


int scan_impl(ltg_context* c, const RecordIn* recs, int64_t n_recs, ltg_result** out)
{
    if (!c || !out || (n_recs > 0 && !recs)) { set_error("null argument"); return LTG_ERR_ARG; }
    if (int e = prepare(c)) return e;
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->lit_stream));
    const int64_t launches0 = c->launches, h2d0 = c->h2d_bytes, d2h0 = c->d2h_bytes, probed0 = c->n_probe_items;
    const bool trace_time = getenv("LTG_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_prep = 0, t_loop = 0, t_retire = 0, t_filter = 0, t_strings = 0;
    // records back to back in one device buffer; every record is cut on its own (cutSequence never spans records)
    std::vector<HostSeg> segs;
    std::vector<int64_t> base((size_t)n_recs + 1, 0);
    for (int64_t r = 0; r < n_recs; ++r) {
        const RecordIn& R = recs[r];
        if (R.first_seg < 0 || R.len < 0 || (R.len > 0 && !R.h_dna && !R.d_dna && !R.packed)) { set_error("bad record %lld", (long long)r); return LTG_ERR_ARG; }
        if (int e = cut_segments(R.len, R.record_len < 0 ? R.len : R.record_len, R.first_seg, R.n_seg, c->params, base[r], (int)r, segs)) return e;
        base[r + 1] = base[r] + R.len;
    }
    const int64_t len = base[n_recs];
    ResultBuilder rb;
    RecordStats st;
    if (len > 0) {
        if (int e = c->d_dna.ensure((size_t)len)) return e;
        if (int e = c->d_codes.ensure((size_t)len)) return e;
        bool any_packed = false;
        for (int64_t r = 0; r < n_recs; ++r) {
            const RecordIn& R = recs[r];
            if (R.len == 0) continue;
            unsigned char* dst = c->d_dna.as<unsigned char>() + base[r];
            if (R.packed) {
                // 2-bit packed record: 0.25 B/base over PCIe (or none: the packed store may already sit in HBM), expanded on the device
                const int64_t b0 = R.packed_first >> 2, b1 = (R.packed_first + R.len + 3) >> 2;
                const unsigned char* dpk = R.packed + b0;
                if (!R.packed_on_device) {
                    if (int e = c->d_packed.ensure((size_t)(b1 - b0))) return e;
                    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_packed.p, R.packed + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, c->stream));
                    c->h2d_bytes += b1 - b0;
                    dpk = c->d_packed.as<unsigned char>();
                }
                if (int e = c->d_nblocks.ensure(sizeof(uint32_t) * 2 * (size_t)std::max(R.n_blocks, 1))) return e;
                uint32_t* dn = c->d_nblocks.as<uint32_t>();
                if (R.n_blocks > 0) {
                    LTG_CUDA_CHECK(cudaMemcpyAsync(dn, R.n_start, sizeof(uint32_t) * (size_t)R.n_blocks, cudaMemcpyHostToDevice, c->stream));
                    LTG_CUDA_CHECK(cudaMemcpyAsync(dn + R.n_blocks, R.n_size, sizeof(uint32_t) * (size_t)R.n_blocks, cudaMemcpyHostToDevice, c->stream));
                    c->h2d_bytes += (int64_t)sizeof(uint32_t) * 2 * R.n_blocks;
                }
                k_unpack_2bit<<<(int)std::min<int64_t>((R.len + 255) / 256, 148 * 16), 256, 0, c->stream>>>(
                    dpk, R.packed_first - (b0 << 2), R.len, dn, dn + R.n_blocks, R.n_blocks, dst, c->d_codes.as<uint8_t>() + base[r]);
                c->launches += 1;
                // (the staging buffers are reused by the next packed record of the call)
                LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
                any_packed = true;
            } else if (R.h_dna) { LTG_CUDA_CHECK(cudaMemcpyAsync(dst, R.h_dna, (size_t)R.len, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += R.len; }
            else LTG_CUDA_CHECK(cudaMemcpyAsync(dst, R.d_dna, (size_t)R.len, cudaMemcpyDeviceToDevice, c->stream));
        }
        k_encode<<<(int)std::min<int64_t>((len + 255) / 256, 148 * 16), 256, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_codes.as<uint8_t>(), len);
        c->launches += 1;
        // segment flags (same_seq + non-ACGT) for every segment of the call
        const int NS = (int)segs.size();
        if (int e = c->d_segs.ensure(sizeof(SegDesc) * std::max(NS, 1))) return e;
        std::vector<SegDesc> hs(NS);
        for (int s = 0; s < NS; ++s) { hs[s].start = segs[s].start; hs[s].len = segs[s].len; hs[s].flags = 0; }
        std::vector<HostSeg> active;
        if (NS > 0) {
            if (n_recs <= 16) {
                // cutSequence on the device (k_cut_segments): the descriptors of a record's segments are arithmetic in (record length,
                // cut length, overlap), so nothing but those numbers crosses PCIe; calls with many short records upload the list instead
                int at = 0;
                for (int64_t r = 0; r < n_recs; ++r) {
                    int cnt = 0;
                    while (at + cnt < NS && segs[at + cnt].record == (int)r) ++cnt;
                    if (cnt > 0)
                        k_cut_segments<<<(cnt + 255) / 256, 256, 0, c->stream>>>(c->d_segs.as<SegDesc>() + at, cnt, base[r], recs[r].first_seg,
                                                                                  c->params.cut_length - c->params.overlap, c->params.cut_length,
                                                                                  recs[r].record_len < 0 ? recs[r].len : recs[r].record_len);
                    c->launches += cnt > 0 ? 1 : 0;
                    at += cnt;
                }
            } else {
                LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_segs.p, hs.data(), sizeof(SegDesc) * NS, cudaMemcpyHostToDevice, c->stream));
                c->h2d_bytes += (int64_t)sizeof(SegDesc) * NS;
            }
            k_seg_flags<<<(NS * 32 + 127) / 128, 128, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_segs.as<SegDesc>(), NS);
            c->launches += 1;
            LTG_CUDA_CHECK(cudaMemcpyAsync(hs.data(), c->d_segs.p, sizeof(SegDesc) * NS, cudaMemcpyDeviceToHost, c->stream));
            LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            c->d2h_bytes += (int64_t)sizeof(SegDesc) * NS;
            for (int s = 0; s < NS; ++s) {
                if (hs[s].start != segs[s].start || hs[s].len != segs[s].len) { set_error("device segmenter disagrees with the host geometry (segment %d)", s); return LTG_ERR_STATE; }
                segs[s].flags = hs[s].flags;
                if (!(hs[s].flags & kSegSkip)) active.push_back(segs[s]);
            }
        }

        if (c->sim_mode) {
            t_prep = now();
            if (int e = run_sim_batches(c, active, recs, rb, st)) return e;
            ltg_result* r = finish_result(rb);
            r->n_segments = st.n_segments; r->n_tasks = st.n_tasks; r->scan_cells = st.scan_cells; r->dna_bases = len;
            r->gpu_launches = c->launches - launches0; r->h2d_bytes = c->h2d_bytes - h2d0; r->d2h_bytes = c->d2h_bytes - d2h0;
            if (trace_time) fprintf(stderr, "[ltg timing] -F mode: prep %.1f ms, SIM batches %.1f ms\n", t_prep - t_begin, now() - t_prep);
            *out = r;
            return LTG_OK;
        }
        // batch size: bounded by the strip-maxima buffer; at least two batches when there is enough work so that the
        // host phase of one overlaps the device phase of the next
        const int cutp = (c->params.cut_length + 3) & ~3;
        const size_t per_seg = (size_t)c->pairs.size() * ((size_t)c->n_strips * (32 * c->scan_r / kGranRows) * blk_pitch_for(cutp) * sizeof(uint16_t) + (size_t)cutp * 4);
        size_t bs = std::max<size_t>(16, std::min<size_t>((size_t)c->batch_segments, kStripBytesPerBatch / std::max<size_t>(1, per_seg)));
        // equal-sized batches (a short last batch would expose its drain and host tail): 2551 segments -> 3 x 851, not 1024 + 1024 + 503
        if (!active.empty()) {
            size_t nbatch = (active.size() + bs - 1) / bs;
            if (active.size() > 256 && nbatch < 2) nbatch = 2;
            bs = (active.size() + nbatch - 1) / nbatch;
        }
        std::vector<ltg_host::Triplex> record_list;
        int rc = LTG_OK;
        size_t nb = 0;
        const int T = (int)c->tasks.size();
        t_prep = now();
        struct Deferred { HostSeg seg; int tdef; int side_row; };
        std::vector<Deferred> deferred;                 // literal (Q4-guard) tasks of the record, in task order
        std::vector<std::pair<int, int>> batch_deferred;
        // side-stream literal sweeps need their workspace in shared memory (the global fallback workspace is shared with the main stream)
        c->side_rows = 0;
        c->side_pitch = (2LL * 16 * kLitArrays * literal_pitch(c->m) <= 200 * 1024) ? ((c->params.cut_length + 3) & ~3) : 0;
        for (size_t b0 = 0; b0 < active.size() && rc == LTG_OK; b0 += bs, ++nb) {
            HostBatch& hb = c->hb[nb & 1];
            if ((rc = retire_batch(c, hb, st, record_list)) != LTG_OK) break;       // batch nb-2: slot free again
            std::vector<HostSeg> batch(active.begin() + b0, active.begin() + std::min(active.size(), b0 + bs));
            batch_deferred.clear();
            if ((rc = run_batch_device(c, batch, hb, true, nullptr, kLitDefer, &batch_deferred)) != LTG_OK) { hb.timed = false; break; }
            for (const std::pair<int, int>& tr : batch_deferred) { Deferred d; d.seg = batch[tr.first / T]; d.tdef = tr.first % T; d.side_row = tr.second; deferred.push_back(d); }
            hb.worker = std::thread(batch_worker, c, &hb);
        }
        t_loop = now();
        // retire in batch order: the older slot first
        for (size_t k = 0; k < 2; ++k) {
            HostBatch& hb = c->hb[(nb + k) & 1];
            const int e = retire_batch(c, hb, st, record_list);
            if (rc == LTG_OK) rc = e;
        }
        // the deferred literal tasks, all together (they run concurrently on the device); their rows are merged back into
        // the record's (segment, task) order
        if (rc == LTG_OK && !deferred.empty()) {
            std::vector<ltg_host::Triplex> lit_rows;
            RecordStats lit_st;
            for (size_t d0 = 0; d0 < deferred.size() && rc == LTG_OK;) {
                std::vector<HostSeg> lsegs;
                std::vector<int> ltasks, lrows;
                size_t d1 = d0;
                for (; d1 < deferred.size(); ++d1) {
                    if (lsegs.empty() || lsegs.back().start != deferred[d1].seg.start) {
                        if (lsegs.size() >= bs) break;
                        lsegs.push_back(deferred[d1].seg);
                    }
                    ltasks.push_back((int)(lsegs.size() - 1) * T + deferred[d1].tdef);
                    lrows.push_back(deferred[d1].side_row);
                }
                HostBatch& hb = c->hb[0];
                if ((rc = run_batch_device(c, lsegs, hb, true, nullptr, kLitOnly, nullptr, &ltasks, &lrows)) != LTG_OK) { hb.timed = false; break; }
                hb.worker = std::thread(batch_worker, c, &hb);
                rc = retire_batch(c, hb, lit_st, lit_rows);
                d0 = d1;
            }
            st.n_literal_tasks += (int64_t)deferred.size();
            st.n_peaks += lit_st.n_peaks; st.window_cells += lit_st.window_cells; st.n_literal_windows += lit_st.n_literal_windows;
            st.ms_scan += lit_st.ms_scan; st.ms_window += lit_st.ms_window;
            if (rc == LTG_OK && !lit_rows.empty()) {
                auto key_less = [](const ltg_host::Triplex& x, const ltg_host::Triplex& y) {
                    return x.seg_start != y.seg_start ? x.seg_start < y.seg_start : x.tdef < y.tdef;
                };
                std::vector<ltg_host::Triplex> merged(record_list.size() + lit_rows.size());
                std::merge(record_list.begin(), record_list.end(), lit_rows.begin(), lit_rows.end(), merged.begin(), key_less);
                record_list.swap(merged);
            }
        }
        if (rc != LTG_OK) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->lit_stream); return rc; }
        LTG_CUDA_CHECK(cudaStreamSynchronize(c->lit_stream));      // (normally long finished: the literal-only batch waited for it)
        t_retire = now();
        std::vector<ltg_host::Triplex> keep;
        for (const ltg_host::Triplex& t : record_list)
            if (ltg_host::passes_record_filter(t, c->params)) keep.push_back(t);
        std::vector<char> pool;
        std::vector<int64_t> offs;
        t_filter = now();
        if (int e = fetch_strings(c, keep, pool, offs)) return e;
        t_strings = now();
        for (size_t i = 0; i < keep.size(); ++i)
            rb.add(keep[i], pool.data() + offs[i], pool.data() + offs[i] + keep[i].nt + 1, recs[keep[i].record].chr, recs[keep[i].record].record_start, keep[i].record);
    }
    ltg_result* r = finish_result(rb);
    r->n_segments = st.n_segments; r->n_tasks = st.n_tasks; r->scan_cells = st.scan_cells; r->dna_bases = len;
    r->n_peaks = st.n_peaks; r->window_cells = st.window_cells; r->n_literal_tasks = st.n_literal_tasks;
    r->n_literal_windows = st.n_literal_windows; r->gpu_ms_scan = st.ms_scan; r->gpu_ms_window = st.ms_window;
    r->gpu_launches = c->launches - launches0;
    r->gpu_ms_scan_kernel = st.ms_scan_kernel; r->n_scan_launches = st.n_scan_launches;
    r->h2d_bytes = c->h2d_bytes - h2d0; r->d2h_bytes = c->d2h_bytes - d2h0;
    r->n_q4_probed = c->n_probe_items - probed0;
    if (trace_time && len > 0)
        fprintf(stderr, "[ltg timing] prep %.1f ms, batches %.1f, last retire %.1f, record filter %.1f, strings %.1f, result %.1f (total %.1f)\n",
                t_prep - t_begin, t_loop - t_prep, t_retire - t_loop, t_filter - t_retire, t_strings - t_filter, now() - t_strings, now() - t_begin);
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
    return guarded([&]() -> int {
    if (!out) { set_error("null argument"); return LTG_ERR_ARG; }
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) { set_error("no CUDA device available (%s) — this library has no CPU fallback", cudaGetErrorString(ce)); return LTG_ERR_CUDA; }
    if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return LTG_ERR_ARG; }
    LTG_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    LTG_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_error("device %s is sm_%d%d; this build contains sm_100a code only", prop.name, prop.major, prop.minor); return LTG_ERR_CUDA; }
    *out = nullptr;
    ltg_context* c = new ltg_context();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    ltg_default_params(&c->params);
    // streams and events; on any failure the half-built context is torn down again (ltg_destroy copes with null handles)
    auto make = [&]() -> int {
        LTG_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        LTG_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        LTG_CUDA_CHECK(cudaStreamCreateWithFlags(&c->lit_stream, cudaStreamNonBlocking));
        LTG_CUDA_CHECK(cudaEventCreateWithFlags(&c->lit_event, cudaEventDisableTiming));
        for (HostBatch& hb : c->hb) {
            LTG_CUDA_CHECK(cudaEventCreateWithFlags(&hb.ready, cudaEventDisableTiming));
            for (int i = 0; i < 6; ++i) LTG_CUDA_CHECK(cudaEventCreate(&hb.ev[i]));
        }
        return LTG_OK;
    };
    if (int e = make()) { ltg_destroy(c); return e; }
    // host-phase worker threads: LTG_HOST_THREADS, else the host cores divided among the visible GPUs (one process
    // per GPU is the deployment model), capped at 16
    int threads = 0;
    if (const char* e = getenv("LTG_HOST_THREADS")) threads = atoi(e);
    if (threads <= 0) threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency() / (unsigned)std::max(1, n)));
    c->host_threads = threads;
    // LTG_NO_PRUNE=1 makes every window sweep the whole lncRNA (the reference's amount of work); results are identical
    if (const char* e = getenv("LTG_NO_PRUNE")) c->prune = atoi(e) == 0;
    // LTG_NO_DEAD=1 traces every alignment, also those that provably cannot be reported (window.cuh DeadRule)
    if (const char* e = getenv("LTG_NO_DEAD")) c->dead_rule = atoi(e) == 0;
    // LTG_NO_SKIP=1 runs every window round of fastSIM's loop even when it provably repeats the previous result
    if (const char* e = getenv("LTG_NO_SKIP")) c->skip_rounds = atoi(e) == 0;
    if (const char* e = getenv("LTG_NO_Q4PROBE")) c->q4_probe = atoi(e) == 0;
    if (const char* e = getenv("LTG_FREC")) c->frec_mode = atoi(e) < 0 ? -1 : (atoi(e) != 0 ? 1 : 0);
    if (const char* e = getenv("LTG_Q4_TAINT")) c->q4_taint = atoi(e) != 0;
    if (const char* e = getenv("LTG_WIN_Q4CHK")) c->win_q4_mode = atoi(e) != 0 ? 1 : 0;
    if (const char* e = getenv("LTG_SCAN_SHARED")) c->scan_shared = atoi(e) != 0;
    // LTG_FLOORS=1: the first window sweep only tracks cells that reach the peak score (fewer slow-path trips of the tracker, more
    // re-planned sweeps; measured neutral on the headline workload, profiles/README.md)
    if (const char* e = getenv("LTG_FLOORS")) c->floor_s = atoi(e) != 0;
    if (const char* e = getenv("LTG_BATCH_SEGMENTS")) c->batch_segments = std::max(16, atoi(e));
    if (const char* e = getenv("LTG_LIT_OLD")) c->lit_col = atoi(e) == 0;
    if (const char* e = getenv("LTG_LIT_ROWS")) c->lit_rows_per_chunk = std::max(4, atoi(e));
    if (const char* e = getenv("LTG_LIT_CH")) c->lit_min_chunks = atoi(e);
    *out = c;
    return LTG_OK;
    });
}

void ltg_destroy(ltg_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    release_tables(c, c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (HostBatch& hb : c->hb) {
        if (hb.worker.joinable()) hb.worker.join();
        for (PinBuf* b : {&hb.task_off, &hb.jobs, &hb.tout, &hb.task_info, &hb.scalars, &hb.c_task, &hb.c_poff}) b->release();
        for (DevBuf* b : {&hb.d_cjobs, &hb.d_ctout, &hb.d_ctask, &hb.d_cpoff}) b->release();
        if (hb.ready) cudaEventDestroy(hb.ready);
        for (int i = 0; i < 6; ++i) if (hb.ev[i]) cudaEventDestroy(hb.ev[i]);
    }
    for (DevBuf* b : {&c->d_rna_raw, &c->d_rna_ssw, &c->d_rna_stats, &c->d_rna_sel, &c->d_prof_ssw, &c->d_prof_stats, &c->d_cut, &c->d_dna, &c->d_codes,
                      &c->d_segs, &c->d_items, &c->d_items_stats, &c->d_blkmax, &c->d_bnd, &c->d_counters, &c->d_task_info, &c->d_task_off,
                      &c->d_scan_order, &c->d_stats_max, &c->d_task_litrow, &c->d_cand, &c->d_pk_task, &c->d_pk_pos, &c->d_pk_score, &c->d_win_list, &c->d_win_sched, &c->d_res, &c->d_colmax_all, &c->d_ovf_list,
                      &c->d_jobs, &c->d_tout, &c->d_strpool, &c->d_scratch, &c->d_scratch_big,
                      &c->d_lit_colmax, &c->d_lit_work, &c->d_lit_jobs, &c->d_side_jobs, &c->d_side_colmax})
        b->release();
    for (DevBuf* b : {&c->d_packed, &c->d_nblocks, &c->d_rna_sim, &c->d_sim_scratch, &c->d_sim_hdr, &c->d_sim_pool, &c->d_sim_tasks}) b->release();
    for (int k = 0; k < 24; ++k) c->d_w[k].release();
    for (int k = 0; k < 4; ++k) c->d_pc[k].release();
    c->d_res64.release();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->lit_stream) { cudaStreamSynchronize(c->lit_stream); cudaStreamDestroy(c->lit_stream); }
    if (c->lit_event) cudaEventDestroy(c->lit_event);
    delete c;
}

void* ltg_stream(ltg_context* c) { return c ? (void*)c->stream : nullptr; }

void ltg_debug_stats(ltg_context* c, int64_t* out30, int reset)
{
    if (!c || !out30) return;
    for (int k = 0; k < 30; ++k) { out30[k] = (int64_t)c->win_stats[k]; if (reset) c->win_stats[k] = 0; }
}

int ltg_set_params(ltg_context* c, const ltg_params* p)
{
    return guarded([&]() -> int {
    if (!c || !p) { set_error("null argument"); return LTG_ERR_ARG; }
    if (p->cut_length <= 0 || p->cut_length - p->overlap <= 0) { set_error("cut length (%d) must be positive and exceed the overlap (%d)", p->cut_length, p->overlap); return LTG_ERR_ARG; }
    if (p->cut_length > kMaxCutLength) { set_error("cut length %d exceeds this build's limit of %d", p->cut_length, kMaxCutLength); return LTG_ERR_LIMIT; }
    std::vector<TaskDef> probe;
    if (!ltg_host::enumerate_tasks(*p, probe)) { set_error("invalid rule/strand selection (rule=%d strand=%d)", p->rule, p->strand); return LTG_ERR_ARG; }
    if (p->rule != c->params.rule || p->strand != c->params.strand) { c->tables_dirty = true; c->params_changed = true; }
    c->params = *p;
    return LTG_OK;
    });
}

int ltg_set_compat(ltg_context* c, int lowercase_variant)
{
    if (!c) { set_error("null argument"); return LTG_ERR_ARG; }
    c->compat = lowercase_variant != 0;
    return LTG_OK;
}

int ltg_set_sim_mode(ltg_context* c, int on)
{
    if (!c) { set_error("null argument"); return LTG_ERR_ARG; }
    c->sim_mode = on != 0;
    return LTG_OK;
}

int ltg_set_query(ltg_context* c, const char* name, const char* rna, int64_t len)
{
    return guarded([&]() -> int {
    if (!c || !rna || len <= 0) { set_error("empty lncRNA"); return LTG_ERR_ARG; }
    if (len > (1 << 24)) { set_error("lncRNA longer than 16 Mnt is not supported"); return LTG_ERR_LIMIT; }
    LTG_CUDA_CHECK(cudaSetDevice(c->device));
    c->rna_name = name ? name : "";
    c->rna.assign(rna, (size_t)len);
    c->m = (int)len;
    std::vector<uint8_t> q1(len), q2(len);
    c->rna_plain = true; c->rna_acgt = true;
    for (int64_t i = 0; i < len; ++i) {
        const unsigned char ch = (unsigned char)rna[i];
        q1[i] = (uint8_t)ssw_code(ch); q2[i] = (uint8_t)stats_code(ch);
        if (q1[i] == 4 || q2[i] >= 4) c->rna_plain = false;      // U or any non-ACGT letter: scorings differ (Q3)
        if (q1[i] == 4) c->rna_acgt = false;
    }
    for (DevBuf* b : {&c->d_rna_raw, &c->d_rna_ssw, &c->d_rna_stats}) if (int e = b->ensure((size_t)len)) return e;
    // PRMT selectors of every row for the window DP's table-lookup scoring (window.cuh, k_win_dp TAB)
    std::vector<uint16_t> sel(len);
    for (int64_t i = 0; i < len; ++i) {
        const unsigned q = q1[i] & 3u;
        sel[i] = (uint16_t)((q | ((q | 8u) << 4)) | (((4u + q) | ((12u + q) << 4)) << 8));
    }
    // -F mode: plain letter codes (A0 C1 G2 T3, anything else 4 — SIM's substitution matrix knows no U, sim.h:468-472)
    std::vector<uint8_t> q3(len);
    for (int64_t i = 0; i < len; ++i) q3[i] = (uint8_t)dna_code((unsigned char)rna[i]);
    if (int e = c->d_rna_sim.ensure((size_t)len)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_sim.p, q3.data(), (size_t)len, cudaMemcpyHostToDevice, c->stream));
    if (int e = c->d_rna_sel.ensure(sizeof(uint16_t) * (size_t)len)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_sel.p, sel.data(), sizeof(uint16_t) * (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_raw.p, rna, (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_ssw.p, q1.data(), (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_rna_stats.p, q2.data(), (size_t)len, cudaMemcpyHostToDevice, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->profiles_dirty = true;
    return prepare(c);
    });
}

int ltg_scan_record(ltg_context* c, const char* dna, int64_t len, const char* chr, int64_t record_start, ltg_result** out)
{
    return guarded([&]() -> int {
    if (!dna && len > 0) { set_error("null DNA"); return LTG_ERR_ARG; }
    RecordIn R; R.h_dna = dna; R.len = len; R.chr = chr; R.record_start = record_start;
    return scan_impl(c, &R, 1, out);
    });
}

int ltg_scan_records(ltg_context* c, int64_t n_records, const char* const* dna, const int64_t* len, const char* const* chr,
                     const int64_t* record_start, ltg_result** out)
{
    return guarded([&]() -> int {
    if (n_records < 0 || (n_records > 0 && (!dna || !len))) { set_error("null argument"); return LTG_ERR_ARG; }
    std::vector<RecordIn> recs((size_t)n_records);
    for (int64_t r = 0; r < n_records; ++r) {
        if (!dna[r] && len[r] > 0) { set_error("null DNA in record %lld", (long long)r); return LTG_ERR_ARG; }
        recs[r].h_dna = dna[r]; recs[r].len = len[r]; recs[r].chr = chr ? chr[r] : nullptr; recs[r].record_start = record_start ? record_start[r] : 0;
    }
    return scan_impl(c, recs.data(), n_records, out);
    });
}

int ltg_scan_records_at(ltg_context* c, int64_t n_records, const void* const* dna, int dna_on_device, const int64_t* len, const char* const* chr,
                        const int64_t* record_start, ltg_result** out)
{
    return guarded([&]() -> int {
        if (n_records < 0 || (n_records > 0 && (!dna || !len))) { set_error("null argument"); return LTG_ERR_ARG; }
        std::vector<RecordIn> recs((size_t)n_records);
        for (int64_t r = 0; r < n_records; ++r) {
            if (!dna[r] && len[r] > 0) { set_error("null DNA in record %lld", (long long)r); return LTG_ERR_ARG; }
            if (dna_on_device) recs[r].d_dna = (const unsigned char*)dna[r]; else recs[r].h_dna = (const char*)dna[r];
            recs[r].len = len[r]; recs[r].chr = chr ? chr[r] : nullptr; recs[r].record_start = record_start ? record_start[r] : 0;
        }
        return scan_impl(c, recs.data(), n_records, out);
    });
}

int ltg_scan_device(ltg_context* c, const void* d_dna, int64_t len, const char* chr, int64_t record_start, ltg_result** out)
{
    return guarded([&]() -> int {
    if (!d_dna && len > 0) { set_error("null DNA"); return LTG_ERR_ARG; }
    RecordIn R; R.d_dna = (const unsigned char*)d_dna; R.len = len; R.chr = chr; R.record_start = record_start;
    return scan_impl(c, &R, 1, out);
    });
}

int ltg_scan_shard(ltg_context* c, const void* dna, int dna_on_device, int64_t len, const char* chr, int64_t record_start,
                   int64_t record_len, int64_t first_segment, int64_t n_segments, ltg_result** out)
{
    return guarded([&]() -> int {
    if (!dna && len > 0) { set_error("null DNA"); return LTG_ERR_ARG; }
    RecordIn R;
    if (dna_on_device) R.d_dna = (const unsigned char*)dna; else R.h_dna = (const char*)dna;
    R.len = len; R.chr = chr; R.record_start = record_start; R.record_len = record_len; R.first_seg = first_segment; R.n_seg = n_segments;
    return scan_impl(c, &R, 1, out);
    });
}

int ltg_scan_packed(ltg_context* c, const void* packed, int packed_on_device, int64_t first_base, int64_t len, const uint32_t* n_start,
                    const uint32_t* n_size, int32_t n_blocks, const char* chr, int64_t record_start, int64_t record_len, int64_t first_segment,
                    int64_t n_segments, ltg_result** out)
{
    return guarded([&]() -> int {
        if ((!packed && len > 0) || first_base < 0 || n_blocks < 0 || (n_blocks > 0 && (!n_start || !n_size))) { set_error("bad argument"); return LTG_ERR_ARG; }
        RecordIn R;
        R.packed = (const unsigned char*)packed; R.packed_on_device = packed_on_device; R.packed_first = first_base; R.len = len;
        R.n_start = n_start; R.n_size = n_size; R.n_blocks = n_blocks;
        R.chr = chr; R.record_start = record_start; R.record_len = record_len; R.first_seg = first_segment; R.n_seg = n_segments;
        return scan_impl(c, &R, 1, out);
    });
}

int ltg_result_new(ltg_result** out)
{
    return guarded([&]() -> int {
    if (!out) return LTG_ERR_ARG;
    ResultBuilder rb;
    *out = finish_result(rb);
    return LTG_OK;
    });
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
    return guarded([&]() -> int {
    if (!dst || !src) { set_error("null argument"); return LTG_ERR_ARG; }
    const int64_t n0 = dst->n_triplex, t0 = dst->text_bytes;
    // grow both blocks before touching anything: on failure dst stays as it was (a successful first realloc only moved a block)
    ltg_triplex* nt = (ltg_triplex*)realloc(dst->triplex, sizeof(ltg_triplex) * (size_t)std::max<int64_t>(1, n0 + src->n_triplex));
    if (!nt) { set_error("out of memory appending %lld triplexes", (long long)src->n_triplex); return LTG_ERR_LIMIT; }
    dst->triplex = nt;
    char* tx = (char*)realloc(dst->text, (size_t)std::max<int64_t>(1, t0 + src->text_bytes));
    if (!tx) { set_error("out of memory appending %lld text bytes", (long long)src->text_bytes); return LTG_ERR_LIMIT; }
    dst->text = tx;
    if (src->text_bytes > 0) memcpy(dst->text + t0, src->text, (size_t)src->text_bytes);
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
    dst->h2d_bytes += src->h2d_bytes; dst->d2h_bytes += src->d2h_bytes; dst->n_q4_probed += src->n_q4_probed;
    return LTG_OK;
    });
}

int ltg_cluster(ltg_result* r, const ltg_params* p)
{
    return guarded([&]() -> int {
    if (!r || !p) { set_error("null argument"); return LTG_ERR_ARG; }
    std::vector<ltg_triplex> v(r->triplex, r->triplex + r->n_triplex);
    ltg_host::cluster(v, p->c_distance, p->c_length, nullptr);
    std::sort(v.begin(), v.end(), [](const ltg_triplex& a, const ltg_triplex& b) { return a.motif < b.motif; });   // :813, :847
    if (!v.empty()) memcpy(r->triplex, v.data(), sizeof(ltg_triplex) * v.size());
    return LTG_OK;
    });
}

int ltg_probe_segment(ltg_context* c, const char* seg, int32_t seg_len, ltg_task_probe* tasks, int32_t n_tasks, int32_t* colmax,
                      int32_t* peak_score, int32_t* peak_pos, int32_t peak_cap)
{
    return guarded([&]() -> int {
    if (!c || !seg || seg_len <= 0 || !tasks) { set_error("bad argument"); return LTG_ERR_ARG; }
    if (seg_len > c->params.cut_length) { set_error("segment longer than the cut length"); return LTG_ERR_ARG; }
    if (int e = prepare(c)) return e;
    if (int e = c->d_dna.ensure((size_t)seg_len)) return e;
    if (int e = c->d_codes.ensure((size_t)seg_len)) return e;
    LTG_CUDA_CHECK(cudaMemcpyAsync(c->d_dna.p, seg, (size_t)seg_len, cudaMemcpyHostToDevice, c->stream));
    k_encode<<<(seg_len + 255) / 256, 256, 0, c->stream>>>(c->d_dna.as<unsigned char>(), c->d_codes.as<uint8_t>(), seg_len);
    c->launches += 1;
    std::vector<HostSeg> segs(1);
    segs[0].start = 0; segs[0].len = seg_len; segs[0].flags = 0; segs[0].coord = 0; segs[0].record = 0; segs[0].pad_ = 0;
    for (int i = 0; i < seg_len; ++i) { const char ch = seg[i]; if (!(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T')) segs[0].flags |= kSegNonACGT; }
    ProbeOut po;
    HostBatch& hb = c->hb[0];
    const int rc = run_batch_device(c, segs, hb, false, &po);
    hb.timed = false;
    if (rc != LTG_OK) return rc;
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const int T = (int)c->tasks.size();
    const int max_len = po.max_len;
    for (int q = 0; q < n_tasks; ++q) {
        int t = -1;
        for (int k = 0; k < T; ++k) if (c->tasks[k].para == tasks[q].para && c->tasks[k].strand == tasks[q].strand && c->tasks[k].rule == tasks[q].rule) t = k;
        if (t < 0) { set_error("task (%d,%d,%d) is not part of the configured rule/strand selection", tasks[q].para, tasks[q].strand, tasks[q].rule); return LTG_ERR_ARG; }
        tasks[q].max_score = po.task_max[t]; tasks[q].threshold = po.task_thr[t]; tasks[q].n_peaks = po.task_npk[t];
        const bool literal = (po.task_flags[t] & kTaskLiteral) != 0;
        tasks[q].literal = literal ? 1 : 0;
        if (colmax) {
            if (literal) {
                // the literal emulation reports exactly what the reference's 8-bit kernel records
                const uint16_t* row = po.lit_colmax.data() + (size_t)po.task_litrow[t] * max_len;
                for (int j = 0; j < seg_len; ++j) colmax[(size_t)q * seg_len + j] = row[j];
            } else {
                const int item = c->tasks[t].pair, half = c->tasks[t].half;
                const uint32_t* row = po.colmax.data() + (size_t)item * max_len;
                bool cut = false;
                for (int j = 0; j < seg_len; ++j) {
                    const int v = half ? hi16(row[j]) : lo16(row[j]);
                    if (v >= kOverflowU8) cut = true;          // Q2: nothing is recorded from the first >= 251 column on
                    colmax[(size_t)q * seg_len + j] = cut ? 0 : v;
                }
            }
        }
        if (peak_score && peak_pos) {
            const int b = po.task_off[t], e = b + po.task_npk[t];
            for (int i = b; i < e && i - b < peak_cap; ++i) { peak_pos[(size_t)q * peak_cap + (i - b)] = po.pk_pos[i]; peak_score[(size_t)q * peak_cap + (i - b)] = po.pk_score[i]; }
        }
    }
    return LTG_OK;
    });
}


// Aligner::Align for explicit windows: every window becomes a one-task "segment" under an identity rule image,
// with one forced peak at its last column, and runs through the very same window kernels as the product path.
int ltg_probe_align(ltg_context* c, const char* const* windows, const int32_t* window_len, int32_t n, int32_t* out6, uint32_t* cigar,
                    int32_t cigar_cap)
{
    return guarded([&]() -> int {
    if (!c || !windows || !window_len || n <= 0 || !out6) { set_error("bad argument"); return LTG_ERR_ARG; }
    if (int e = prepare(c)) return e;
    // identity task table (restored afterwards by marking the tables dirty)
    TaskDef id; memset(&id, 0, sizeof id);
    id.para = 1; id.strand = 0; id.rule = 1; id.reversed = 0; id.comp_src = 0;
    for (int k = 0; k < 5; ++k) id.img[k] = (int8_t)k;
    PairDef pd; pd.task[0] = pd.task[1] = 0; pd.reversed = 0; pd.pad_ = 0;
    id.pair = 0; id.half = 0;
    if (!claim_tables(c, c->device, table_bytes(std::vector<TaskDef>(1, id), std::vector<PairDef>(1, pd)))) {
        set_error("ltg_probe_align replaces the device's task tables: not available while another context uses device %d", c->device);
        return LTG_ERR_STATE;
    }
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
    if (int e = run_windows(c, n, 1, c->d_stats_max.as<int>(), nullptr, 0, nullptr, false)) return e;
    if (int e = run_traceback(c, c->d_jobs.as<TraceJob>(), n, c->d_tout.as<TraceOut>(), nullptr)) return e;
    std::vector<TraceJob> jobs(n);
    std::vector<TraceOut> tout(n);
    LTG_CUDA_CHECK(cudaMemcpyAsync(jobs.data(), c->d_jobs.p, sizeof(TraceJob) * n, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaMemcpyAsync(tout.data(), c->d_tout.p, sizeof(TraceOut) * n, cudaMemcpyDeviceToHost, c->stream));
    LTG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    // strings of every successful window through the second traceback pass (the product's own string path)
    std::vector<ltg_host::Triplex> rows;
    std::vector<int> row_of(n, -1);
    for (int i = 0; i < n; ++i) {
        if (jobs[i].score <= 0 || tout[i].status != 1) continue;
        ltg_host::Triplex t;
        t.nt = tout[i].nt; t.score = (float)jobs[i].score; t.tdef = 0; t.seg_len = jobs[i].seg_len; t.seg_start = (long)jobs[i].seg_start;
        t.ws = jobs[i].ws; t.rb = jobs[i].rb; t.re = jobs[i].re; t.qb = jobs[i].qb; t.qe = jobs[i].qe;
        row_of[i] = (int)rows.size();
        rows.push_back(t);
    }
    std::vector<char> pool;
    std::vector<int64_t> offs;
    if (int e = fetch_strings(c, rows, pool, offs)) return e;
    for (int i = 0; i < n; ++i) {
        int32_t* o = out6 + (size_t)i * 6;
        if (row_of[i] < 0) { for (int k = 0; k < 6; ++k) o[k] = 0; continue; }
        o[0] = jobs[i].score; o[1] = jobs[i].rb; o[2] = jobs[i].re; o[3] = jobs[i].qb; o[4] = jobs[i].qe;
        // run-length CIGAR from the aligned strings (M: both present, I: gap on the DNA side, D: gap on the RNA side)
        const int nt = tout[i].nt;
        const char* tfo = pool.data() + offs[row_of[i]];
        const char* tts = tfo + nt + 1;
        int nc = 0, run = 0, prev = -1;
        for (int k = 0; k <= nt; ++k) {
            const int op = k == nt ? -2 : (tts[k] == '-' ? 1 : (tfo[k] == '-' ? 2 : 0));
            if (op == prev) { ++run; continue; }
            if (prev >= 0) { if (cigar && nc < cigar_cap) cigar[(size_t)i * cigar_cap + nc] = ((uint32_t)run << 4) | (uint32_t)prev; ++nc; }
            prev = op; run = 1;
        }
        o[5] = nc;
    }
    return LTG_OK;
    });
}

}  // extern "C"

#include "../host/driver.inl"
