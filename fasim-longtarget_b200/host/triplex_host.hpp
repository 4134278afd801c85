// Host side of the hot path that stays scalar C++: turning device alignments
// into triplex records, the per-task de-duplication (std::sort / std::unique with the reference's comparators,
// whose outcome depends on libstdc++'s permutation — SURVEY App. B Q8), clustering and the -TFOsorted /
// -TFOclass writers.  Everything here is cheap list work on plain records; all DP, the traceback, the identity
// and stability arithmetic and the string expansion run on the GPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/fasim_b200.h"

namespace ltg_host {

// One candidate triplex while it travels through the per-task de-duplication and the filters.  Plain data: the
// TFO / TTS strings are only produced (by the second traceback pass on the GPU) for the rows that survive.
struct Triplex {
    int stari = 0, endi = 0, starj = 0, endj = 0, reverse = 0, strand = 0, rule = 0, nt = 0;
    float score = 0, identity = 0, tri_score = 0;
    // where the alignment lives: input of the string pass (window.cuh TraceJob)
    int tdef = 0, seg_len = 0, ws = 0, rb = 0, re = 0, qb = 0, qe = 0;
    int shift = 0;          // lowercase compat: reported coordinates and strings lie `shift` columns left of the aligned ones
    long seg_start = 0;     // offset of the segment in the device DNA buffer of the call
    int record = 0;         // record index within the call
};

struct DeviceAlignment {        // one chosen alignment as it comes back from the GPU (pass 1 of the traceback)
    int sw_score, ws, rb, re, query_begin, query_end;   // ws + rb / ws + re = ref_begin / ref_end in translated-segment coordinates
    int shift = 0;              // ... minus this in the older variant's reporting (fastSim.h:211-212), 0 otherwise
    int nt;
    float identity, tri_score;  // evaluated on the device exactly as fastsim.h:323-383 does (float32, same order)
};

// tail of convertMyTriplex — fastsim.h:385-405: orientation-dependent coordinates (:389-396), nt >= ntMin gate (:397)
inline void make_triplex(const DeviceAlignment& al, int tdef, int seg_len, long seg_start, long seg_coord, int record, int para, int strand,
                         int rule, const ltg_params& P, std::vector<Triplex>& out)
{
    if (al.nt < P.nt_min) return;
    const int ref_begin = al.ws + al.rb - al.shift, ref_end = al.ws + al.re - al.shift;
    int a, b;
    if ((para > 0 && strand == 1) || (para < 0 && strand == 0)) { a = seg_len - ref_end - 1; b = seg_len - ref_begin - 1; }
    else { a = ref_begin + 1; b = ref_end + 1; }
    Triplex t;
    t.stari = al.query_begin + 1; t.endi = al.query_end + 1;
    t.starj = (int)(a + seg_coord); t.endj = (int)(b + seg_coord);        // dnaStartPos of fastSIM = the segment's offset in its record
    t.strand = strand; t.reverse = para; t.rule = rule; t.nt = al.nt;
    t.score = (float)al.sw_score; t.identity = al.identity; t.tri_score = al.tri_score;
    t.tdef = tdef; t.seg_len = seg_len; t.seg_start = seg_start; t.record = record;
    t.ws = al.ws; t.shift = al.shift; t.rb = al.rb; t.re = al.re; t.qb = al.query_begin; t.qe = al.query_end;
    out.push_back(t);
}

// comparators of fastsim.h:92-156 (not strict weak orders — kept verbatim in meaning, Q8)
inline bool by_start(const Triplex& a, const Triplex& b)
{
    if (a.stari == b.stari && a.starj == b.starj) return a.score > b.score;
    return a.starj > b.starj;
}
inline bool by_end(const Triplex& a, const Triplex& b)
{
    if (a.endi == b.endi && a.starj == b.starj) return a.score > b.score;
    return a.starj < b.starj;
}
inline bool by_score(const Triplex& a, const Triplex& b) { return a.score > b.score; }
inline bool redundant(const Triplex& a, const Triplex& b)
{
    const bool same = a.stari == b.stari && a.starj == b.starj && a.endi == b.endi && a.endj == b.endj && a.score == b.score;
    const bool inside = b.stari >= a.stari && b.starj >= a.starj && b.endi <= a.endi && b.endj <= a.endj && b.score < a.score;
    return same || inside;
}

// tail of fastSIM — fastsim.h:273-288: two sort/unique rounds, sort by score, first 50, per-task filter
// (compat: the older variant keeps the first 50 without the identity / stability / nt filter, fastSim.h:311-313)
inline void finish_task(std::vector<Triplex>& mine, const ltg_params& P, std::vector<Triplex>& out, bool compat = false)
{
    std::sort(mine.begin(), mine.end(), by_start);
    mine.erase(std::unique(mine.begin(), mine.end(), redundant), mine.end());
    std::sort(mine.begin(), mine.end(), by_end);
    mine.erase(std::unique(mine.begin(), mine.end(), redundant), mine.end());
    std::sort(mine.begin(), mine.end(), by_score);
    const size_t lim = std::min<size_t>(mine.size(), 50);
    const float min_id = (float)P.min_identity, min_st = (float)P.min_stability;
    for (size_t i = 0; i < lim; ++i)
        if (compat || (mine[i].identity >= min_id && mine[i].tri_score >= min_st && mine[i].nt >= P.nt_min)) out.push_back(mine[i]);
}

// final filter of LongTarget — Fasim-LongTarget.cpp:589-597
inline bool passes_record_filter(const Triplex& t, const ltg_params& P)
{
    return t.score >= 0.0f && t.identity >= (float)P.min_identity && t.tri_score >= (float)P.min_stability && t.nt >= P.c_length;
}

// cluster_triplex — Fasim-LongTarget.cpp:600-691.  class_cov (optional) receives, for classes 1..5, the
// per-position coverage map used by the -TFOclass writer (class1[] of the reference).
inline void cluster(std::vector<ltg_triplex>& v, int dd, int length, std::map<size_t, size_t>* class_cov)
{
    std::map<long, long> weight;
    long top = 0, center = 0;
    bool found = false;
    for (ltg_triplex& t : v) {
        if (t.nt <= length) continue;
        const int mid = (t.stari + t.endi) / 2;
        t.middle = mid;
        t.motif = 0;
        weight[mid];
        for (int k = -dd; k <= dd; ++k) {
            long& wk = weight[mid + k];
            if (k != 0) wk += dd - (k < 0 ? -k : k);
            if (wk > top) { top = wk; center = mid + k; found = true; }
        }
    }
    // The reference walks the whole list for every position of every class window (quadratic in the number of rows).
    // Same assignment order from an index: rows grouped by MidPoint, each group in list order.  (Rows that were not
    // counted above keep MidPoint 0 and can be claimed by a window that covers position 0, Q9.)
    std::map<long, std::vector<size_t>> by_mid;
    for (size_t i = 0; i < v.size(); ++i) by_mid[v[i].middle].push_back(i);
    int cls = 1;
    while (found) {
        for (long p = center - dd; p <= center + dd; ++p) {
            auto it = by_mid.find(p);
            if (it != by_mid.end()) {
                for (size_t i : it->second) {
                    ltg_triplex& t = v[i];
                    if (t.motif != 0) continue;
                    t.motif = cls;
                    t.center = (int)center;
                    if (class_cov && cls <= 5) {
                        if (t.endj > t.starj) for (int j = t.starj; j < t.endj; ++j) class_cov[cls][(size_t)j]++;
                        else for (int j = t.endj; j < t.starj; ++j) class_cov[cls][(size_t)j]++;
                    }
                }
                by_mid.erase(it);           // every row of the group is assigned now
            }
            weight.erase(p);
        }
        top = 0;
        found = false;
        // the reference rescans keys 0,1,2,... (creating empty ones); only existing positive weights can win,
        // in ascending key order with a strict comparison
        for (const auto& kv : weight) if (kv.second > top) { top = kv.second; center = kv.first; found = true; }
        ++cls;
    }
}

inline const char* strand_name(int reverse, int strand)     // Fasim-LongTarget.cpp:851-871
{
    if (reverse == 1) return strand == 0 ? "ParaPlus" : (strand == 1 ? "ParaMinus" : "");
    if (reverse == -1) return strand == 1 ? "AntiMinus" : (strand == 0 ? "AntiPlus" : "");
    return "";
}

}  // namespace ltg_host
