// Host side of the hot path that stays scalar C++: turning device alignments
// into triplex records, the per-task de-duplication, clustering and the -TFOsorted / -TFOclass writers.
// Everything here is cheap string / list work; all DP runs on the GPU.
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

struct Triplex {
    int stari = 0, endi = 0, starj = 0, endj = 0, reverse = 0, strand = 0, rule = 0, nt = 0;
    float score = 0, identity = 0, tri_score = 0;
    std::string tfo, tts;               // stri_align / strj_align of sim.h:36-37
    int middle = 0, center = 0, motif = 0;
    long genomestart = 0, genomeend = 0;
    int record = 0;
};

// Hoogsteen / reverse-Hoogsteen stability table — sim.h:72-97
inline float stability(char dna, char rna, int para)
{
    if (para > 0) {
        switch (dna) {
        case 'A': if (rna == 'T') return 3.7; break;
        case 'T': if (rna == 'G') return 2.8; break;
        case 'G': if (rna == 'G') return 2.2; if (rna == 'T') return 2.4; if (rna == 'C') return 4.5; break;
        case 'C': if (rna == 'T') return 2.6; if (rna == 'C') return 2.4; break;
        default: break;
        }
    } else {
        switch (dna) {
        case 'A': if (rna == 'A') return 3.0; if (rna == 'T') return 3.5; if (rna == 'C') return 1.0; break;
        case 'T': if (rna == 'G') return 1.0; break;
        case 'G': if (rna == 'A') return 1.0; if (rna == 'G') return 3.0; if (rna == 'C') return 3.0; break;
        case 'C': if (rna == 'T') return 2.0; if (rna == 'C') return 1.0; break;
        default: break;
        }
    }
    return 0;
}

struct DeviceAlignment {        // one chosen alignment as it comes back from the GPU
    int sw_score, ref_begin, ref_end, query_begin, query_end;   // ref_* in translated-segment coordinates
    int nt, match;
    const char* tfo;            // lncRNA side, '-' in D columns
    const char* tts;            // DNA source-strand side, '-' in I columns
};

// convertMyTriplex — fastsim.h:291-414 (identity :323-335, stability with TT/CC penalties :342-383,
// orientation-dependent coordinates :389-396).  float32 throughout, same evaluation order.
inline void make_triplex(const DeviceAlignment& al, int seg_len, long seg_start, int para, int strand, int rule,
                         const ltg_params& P, std::vector<Triplex>& out)
{
    const int nt = al.nt;
    const float identity = (float)(100 * al.match) / (float)(nt);
    float sum = 0.0f, prev_val = 0.0f, val = 0.0f;
    char prev_ch = 0, ch = 0;
    if (nt >= P.nt_min && nt <= P.nt_max) {
        for (int i = 0; i < nt; ++i) {
            ch = al.tts[i];
            val = stability(ch, al.tfo[i], para);
            if (ch == prev_ch && ch == 'T') { sum = sum - prev_val + P.penalty_t; val = P.penalty_t; }
            if (ch == prev_ch && ch == 'C') { sum = sum - prev_val + P.penalty_c; val = P.penalty_c; }
            prev_val = val;
            if (ch != '-') prev_ch = ch;
            sum += val;
        }
        sum = sum / nt;
    }
    int a, b;
    if ((para > 0 && strand == 1) || (para < 0 && strand == 0)) { a = seg_len - al.ref_end - 1; b = seg_len - al.ref_begin - 1; }
    else { a = al.ref_begin + 1; b = al.ref_end + 1; }
    if (nt < P.nt_min) return;
    Triplex t;
    t.stari = al.query_begin + 1; t.endi = al.query_end + 1;
    t.starj = (int)(a + seg_start); t.endj = (int)(b + seg_start);
    t.strand = strand; t.reverse = para; t.rule = rule; t.nt = nt;
    t.score = (float)al.sw_score; t.identity = identity; t.tri_score = sum;
    t.tfo.assign(al.tfo, nt); t.tts.assign(al.tts, nt);
    out.push_back(std::move(t));
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
inline void finish_task(std::vector<Triplex>& mine, const ltg_params& P, std::vector<Triplex>& out)
{
    std::sort(mine.begin(), mine.end(), by_start);
    mine.erase(std::unique(mine.begin(), mine.end(), redundant), mine.end());
    std::sort(mine.begin(), mine.end(), by_end);
    mine.erase(std::unique(mine.begin(), mine.end(), redundant), mine.end());
    std::sort(mine.begin(), mine.end(), by_score);
    const size_t lim = std::min<size_t>(mine.size(), 50);
    const float min_id = (float)P.min_identity, min_st = (float)P.min_stability;
    for (size_t i = 0; i < lim; ++i)
        if (mine[i].identity >= min_id && mine[i].tri_score >= min_st && mine[i].nt >= P.nt_min) out.push_back(mine[i]);
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
    int cls = 1;
    while (found) {
        for (long p = center - dd; p <= center + dd; ++p) {
            for (ltg_triplex& t : v) {
                if (t.middle != p || t.motif != 0) continue;
                t.motif = cls;
                t.center = (int)center;
                if (class_cov && cls <= 5) {
                    if (t.endj > t.starj) for (int j = t.starj; j < t.endj; ++j) class_cov[cls][(size_t)j]++;
                    else for (int j = t.endj; j < t.starj; ++j) class_cov[cls][(size_t)j]++;
                }
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
