// Hoogsteen / reverse-Hoogsteen base-pairing rule sets (rules.h:6-53) as data: for every
// (orientation, strand, rule) the images of the DNA bases, and the reference's task enumeration order
// (Fasim-LongTarget.cpp:404-585).
#pragma once
#include <vector>

#include "../../include/fasim_b200.h"
#include "../csrc/common.cuh"

namespace ltg_host {

// images listed for A, C, G, T (N always maps to N; any other byte is treated as N — rules.h:308-311)
struct RuleImage { char a, c, g, t; };

// parallel (Hoogsteen) rules 1..6; [0] = purine strand as given (strand 0), [1] = opposite strand (strand 1)
static const RuleImage kParallel[2][6] = {
    {{'T','T','G','G'}, {'T','T','C','G'}, {'T','T','T','G'}, {'T','C','G','G'}, {'T','C','C','G'}, {'T','C','T','G'}},
    {{'G','G','T','T'}, {'G','C','T','T'}, {'G','T','T','T'}, {'G','G','C','T'}, {'G','C','C','T'}, {'G','T','C','T'}},
};
// anti-parallel (reverse Hoogsteen) rules 1..18; [0] = strand 0 ("REV" strings of rules.h), [1] = strand 1
static const RuleImage kAntiParallel[2][18] = {
    {{'T','T','G','G'}, {'T','T','C','G'}, {'T','T','A','G'}, {'T','C','G','G'}, {'T','C','C','G'}, {'T','C','A','G'},
     {'A','T','G','G'}, {'A','T','C','G'}, {'A','T','A','G'}, {'A','C','G','G'}, {'A','C','C','G'}, {'A','C','A','G'},
     {'C','T','G','G'}, {'C','T','C','G'}, {'C','T','A','G'}, {'C','C','G','G'}, {'C','C','C','G'}, {'C','C','A','G'}},
    {{'G','G','T','T'}, {'G','C','T','T'}, {'G','A','T','T'}, {'G','G','C','T'}, {'G','C','C','T'}, {'G','A','C','T'},
     {'G','G','T','A'}, {'G','C','T','A'}, {'G','A','T','A'}, {'G','G','C','A'}, {'G','C','C','A'}, {'G','A','C','A'},
     {'G','G','T','C'}, {'G','C','T','C'}, {'G','A','T','C'}, {'G','G','C','C'}, {'G','C','C','C'}, {'G','A','C','C'}},
};

inline int letter_code(char ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; }

inline bool make_task(int para, int strand, int rule, ltg::TaskDef& t)
{
    const RuleImage* im = nullptr;
    if (para > 0) { if (rule < 1 || rule > 6) return false; im = &kParallel[strand ? 1 : 0][rule - 1]; }
    else { if (rule < 1 || rule > 18) return false; im = &kAntiParallel[strand ? 1 : 0][rule - 1]; }
    t.para = (int8_t)para; t.strand = (int8_t)strand; t.rule = (int8_t)rule;
    // ParaMinus (+1,1) and AntiPlus (-1,0) read the segment reversed; ParaMinus and AntiMinus (-1,1) report the
    // complementary strand as the TTS (Fasim-LongTarget.cpp:427-431, 499-501, 519-522)
    t.reversed = (int8_t)((para > 0 && strand == 1) || (para < 0 && strand == 0));
    t.comp_src = (int8_t)(strand == 1);
    t.img[0] = (int8_t)letter_code(im->a); t.img[1] = (int8_t)letter_code(im->c);
    t.img[2] = (int8_t)letter_code(im->g); t.img[3] = (int8_t)letter_code(im->t); t.img[4] = 4;
    t.pair = 0; t.half = 0;
    return true;
}

// task enumeration of LongTarget() for the -r / -t selection (Fasim-LongTarget.cpp:404-585)
inline bool enumerate_tasks(const ltg_params& P, std::vector<ltg::TaskDef>& out)
{
    out.clear();
    ltg::TaskDef t;
    auto add = [&](int para, int strand, int rule) { if (!make_task(para, strand, rule, t)) return false; out.push_back(t); return true; };
    if (P.strand >= 0) {
        if (P.rule == 0) { for (int r = 1; r <= 6; ++r) { add(1, 0, r); add(1, 1, r); } }
        else if (P.rule > 0 && P.rule < 7) { add(1, 0, P.rule); add(1, 1, P.rule); }
    }
    if (P.strand <= 0) {
        if (P.rule == 0) { for (int r = 1; r <= 18; ++r) { add(-1, 1, r); add(-1, 0, r); } }
        else { if (!add(-1, 1, P.rule)) return false; add(-1, 0, P.rule); }     // reference exit(1)s on an invalid rule
    }
    return !out.empty();
}

}  // namespace ltg_host
