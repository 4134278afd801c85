// -F mode, host side: an alignment reported by the SIM kernel (coordinates + edit script) becomes a triplex record — the
// tail of SIM()'s loop body, sim.h:595-743 (display :350-388 for the gapped strings and the identity, the stability loop
// :683-710, the orientation-dependent coordinates :714-727).  Plain float32 arithmetic in the reference's order (this file
// is compiled without FMA contraction, like the rest of the host code).
#pragma once
#include <string>
#include <vector>

#include "../csrc/common.cuh"
#include "../csrc/sim_core.cuh"
#include "triplex_host.hpp"

namespace ltg_host {

struct SimRow { Triplex t; std::string tfo, tts; };

inline float sim_stability(char c1, char c2, int para)      // triplex_score, sim.h:72-97
{
    if (para > 0) {
        if (c1 == 'A' && c2 == 'T') return 3.7;
        if (c1 == 'T' && c2 == 'G') return 2.8;
        if (c1 == 'G' && c2 == 'G') return 2.2;
        if (c1 == 'G' && c2 == 'T') return 2.4;
        if (c1 == 'G' && c2 == 'C') return 4.5;
        if (c1 == 'C' && c2 == 'T') return 2.6;
        if (c1 == 'C' && c2 == 'C') return 2.4;
    } else {
        if (c1 == 'A' && c2 == 'A') return 3.0;
        if (c1 == 'A' && c2 == 'T') return 3.5;
        if (c1 == 'A' && c2 == 'C') return 1.0;
        if (c1 == 'T' && c2 == 'G') return 1.0;
        if (c1 == 'G' && c2 == 'A') return 1.0;
        if (c1 == 'G' && c2 == 'G') return 3.0;
        if (c1 == 'G' && c2 == 'C') return 3.0;
        if (c1 == 'C' && c2 == 'T') return 2.0;
        if (c1 == 'C' && c2 == 'C') return 1.0;
    }
    return 0;
}

inline char sim_comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; } }

// `seg` = the raw segment (seg_len bases), `seg_coord` = its offset in the record (dnaStartPos of SIM)
inline void sim_convert(const ltg::simk::Aln& al, const int* script, const char* rna, const ltg::TaskDef& td, const char* seg, int seg_len,
                        long seg_coord, const ltg_params& P, std::vector<SimRow>& out)
{
    const int stari = al.stari, starj = al.starj, endi = al.endi, endj = al.endj;
    const int m = endi - stari + 1, n = endj - starj + 1;
    const int nt = m;                                                     // sim.h:589: the lncRNA bases spanned
    auto seg_index = [&](int col) { const int q = col - 1; return td.reversed ? seg_len - 1 - q : q; };     // column (1-based) -> segment index
    auto translated = [&](int col) { const int d = td.img[ltg::dna_code((unsigned char)seg[seg_index(col)])]; return d < 4 ? "ACGT"[d] : 'N'; };
    auto source = [&](int col) { const char raw = seg[seg_index(col)]; return td.comp_src ? sim_comp(raw) : raw; };
    // display (:350-388)
    std::string sa, sb;
    long match = 0, mis = 0;
    {
        int i = 0, j = 0, at = 0;
        while (i < m || j < n) {
            while (i < m && j < n && at < al.script_len && script[at] == 0) {
                ++i; ++j;
                const char x = rna[stari - 1 + i - 1], y = translated(starj - 1 + j);
                if (x == y) ++match; else ++mis;
                sa += x; sb += y;
                ++at;
            }
            if (i < m || j < n) {
                const int op = at < al.script_len ? script[at] : 0;
                ++at;
                if (op > 0) for (int f = 0; f < op; ++f) { sa += '-'; sb += translated(starj - 1 + (++j)); ++mis; }
                else for (int f = 0; f < -op; ++f) { sb += '-'; sa += rna[stari - 1 + (++i) - 1]; ++mis; }
                if (op == 0) break;         // (cannot happen: the script covers the box)
            }
        }
    }
    const float identity = (float)(100 * match) / (float)(match + mis);
    if (!(nt >= P.nt_min && nt <= P.nt_max)) return;
    // stability (:683-710)
    float tri_score = 0.0f, hashvalue = 0, prescore = 0;
    char prechar = 0, curchar = 0;
    std::string tts;
    int j = 0;
    for (size_t i = 0; i < sb.size(); ++i) {
        if (sb[i] == '-') { curchar = '-'; hashvalue = sim_stability(curchar, sa[i], td.para); tts += '-'; }
        else {
            curchar = source(starj + j);
            hashvalue = sim_stability(curchar, sa[i], td.para);
            tts += curchar;
            ++j;
        }
        if (curchar == prechar && curchar == 'T') { tri_score = tri_score - prescore + P.penalty_t; hashvalue = P.penalty_t; }
        if (curchar == prechar && curchar == 'C') { tri_score = tri_score - prescore + P.penalty_c; hashvalue = P.penalty_c; }
        prescore = hashvalue;
        if (sb[i] != '-') prechar = curchar;
        tri_score += hashvalue;
    }
    tri_score /= nt;
    const long N = seg_len;
    int refStart, refEnd;
    if (td.para < 0 && td.strand == 0) { refStart = (int)(N - endj + 1); refEnd = (int)(N - starj + 1); }
    else if (td.para > 0 && td.strand == 1) { refStart = (int)(N - endj - 1); refEnd = (int)(N - starj - 1); }      // (sic: off by two, :724-727)
    else { refStart = starj; refEnd = endj; }
    SimRow r;
    r.t.stari = stari; r.t.endi = endi; r.t.starj = (int)(refStart + seg_coord); r.t.endj = (int)(refEnd + seg_coord);
    r.t.strand = td.strand; r.t.reverse = td.para; r.t.rule = td.rule; r.t.nt = nt;
    r.t.score = (float)(long)(al.score / 10); r.t.identity = identity; r.t.tri_score = tri_score;
    r.tfo = sa; r.tts = tts;
    out.push_back(r);
}

}  // namespace ltg_host
