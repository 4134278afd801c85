// TEST INFRASTRUCTURE ONLY — never linked into the product.
//
// Thin C-ABI shim around the *unmodified* reference sources, compiled where they lie
// (-I/root/reference) by oracle/Makefile into oracle/_ref/libref_shim.so.  It pulls the whole
// reference driver into this TU (its `main` renamed) so that every function on the hot path
// can be called directly from the parity tests:
//   rules.h:94 transferString, :59 complement, :88 reverseSeq
//   stats.h:879 calc_score_once
//   sswNew.cpp:1274 ssw_init, :1309 ssw_pre_align (per-column maxima)
//   ssw_cpp.cpp:388 Aligner::preAlign (peaks), :599 Aligner::Align (window alignment)
//   fastsim.h:158 fastSIM (one task), Fasim-LongTarget.cpp:379 LongTarget (one record),
//   Fasim-LongTarget.cpp:600 cluster_triplex
// Results cross the boundary as plain ints / a tab-separated text buffer (floats as raw hex
// bits so that tests can compare bit patterns).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define main ref_main_unused
#include "Fasim-LongTarget.cpp"
#undef main

extern "C" int* ssw_pre_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO,
                              const uint8_t weight_gapE, const uint8_t flag, const uint16_t filters,
                              const int32_t filterd, const int32_t maskLen, int threshold);

namespace {

// same table as ssw_cpp.cpp:13-26 (A0 C1 G2 T3 U0 else 4, case-insensitive) — needed because
// Aligner::TranslateBase is private; only used to feed ssw_init/ssw_pre_align directly.
int8_t ssw_code(char ch)
{
    switch (ch) {
    case 'A': case 'a': case 'U': case 'u': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

void default_matrix(int8_t* mat)   // ssw_cpp.cpp:28-53 with match 5, mismatch 4
{
    int id = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) mat[id++] = (i == j) ? 5 : -4;
        mat[id++] = -4;
    }
    for (int i = 0; i < 5; ++i) mat[id++] = -4;
}

unsigned fbits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

void append_triplex(std::string& out, const triplex& t)
{
    char buf[256];
    snprintf(buf, sizeof buf, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%08x\t%08x\t%08x\t%d\t%d\t%d\t%d\t%ld\t%ld\t",
             t.stari, t.endi, t.starj, t.endj, t.strand, t.reverse, t.rule, t.nt, fbits(t.score), fbits(t.identity),
             fbits(t.tri_score), t.middle, t.center, t.motif, t.neartriplex, t.genomestart, t.genomeend);
    out += buf;
    out += t.stri_align; out += '\t'; out += t.strj_align; out += '\t'; out += t.chr; out += '\n';
}

int emit(const std::string& s, char* out, long cap)
{
    if ((long)s.size() + 1 > cap) return -(int)(s.size() + 1);
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

para make_para(const int* p)
{
    // p: rule, cutLength, strand, overlap, ntMin, ntMax, minIdentity, minStability, penaltyT, penaltyC, cDistance, cLength
    para q;
    q.file1path = q.file2path = q.outpath = "./";
    q.corenum = 1; q.rule = p[0]; q.cutLength = p[1]; q.strand = p[2]; q.overlapLength = p[3];
    q.minScore = 0; q.detailOutput = false; q.doFastSim = true; q.ntMin = p[4]; q.ntMax = p[5];
    q.scoreMin = 0.0; q.minIdentity = p[6]; q.minStability = p[7]; q.penaltyT = p[8]; q.penaltyC = p[9];
    q.cDistance = p[10]; q.cLength = p[11];
    return q;
}

}  // namespace

extern "C" {

int ref_transfer(const char* seg, int strand, int para_, int rule, char* out)
{
    std::string s = transferString(std::string(seg), strand, para_, rule);
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// seq2 / src strings of one task exactly as LongTarget() builds them (Fasim-LongTarget.cpp:410-529).
// para_ = +1/-1, strand = 0/1.  Returns length of seq2; src may be shorter (complement drops chars).
int ref_task_strings(const char* seg, int para_, int strand, int rule, char* seq2_out, char* src_out)
{
    std::string seq1(seg), seq2, src;
    if (para_ > 0 && strand == 0) { seq2 = transferString(seq1, 0, 1, rule); src = seq1; }
    else if (para_ > 0 && strand == 1) { seq2 = transferString(seq1, 1, 1, rule); reverseSeq(seq2); src = seq1; complement(src); reverseSeq(src); }
    else if (para_ < 0 && strand == 1) { seq2 = transferString(seq1, 1, -1, rule); src = seq1; complement(src); }
    else { seq2 = transferString(seq1, 0, -1, rule); reverseSeq(seq2); src = seq1; reverseSeq(src); }
    memcpy(seq2_out, seq2.c_str(), seq2.size() + 1);
    memcpy(src_out, src.c_str(), src.size() + 1);
    return (int)seq2.size();
}

int ref_calc_score_once(const char* rna, const char* seq2)
{
    std::string a(rna), b(seq2);
    return calc_score_once(a, b, 0, 0);
}

// per-column maxima exactly as Aligner::preAlign obtains them (ssw_cpp.cpp:397-415)
int ref_colmax(const char* rna, const char* seq2, int n, int* out)
{
    int m = (int)strlen(rna);
    std::vector<int8_t> q(m), r(n);
    for (int i = 0; i < m; ++i) q[i] = ssw_code(rna[i]);
    for (int i = 0; i < n; ++i) r[i] = ssw_code(seq2[i]);
    int8_t mat[25]; default_matrix(mat);
    s_profile* p = ssw_init(q.data(), m, mat, 5, 2);
    int* cm = ssw_pre_align(p, r.data(), n, 16, 4, 0x0f, 0, 32767, 15, 0);
    for (int i = 0; i < n; ++i) out[i] = cm[i];
    free(cm);
    init_destroy(p);
    return n;
}

int ref_prealign(const char* rna, const char* seq2, int n, int threshold, int* scores, int* positions, int cap)
{
    StripedSmithWaterman::Aligner aligner;
    StripedSmithWaterman::Filter filter;
    StripedSmithWaterman::Alignment al;
    std::vector<StripedSmithWaterman::scoreInfo> peaks;
    aligner.preAlign(rna, seq2, n, filter, &al, 15, threshold, peaks, 5, -4);
    int k = 0;
    for (size_t i = 0; i < peaks.size() && k < cap; ++i, ++k) { scores[k] = peaks[i].score; positions[k] = peaks[i].position; }
    return (int)peaks.size();
}

// out6: sw_score, ref_begin, ref_end, query_begin, query_end, n_cigar
int ref_align(const char* rna, const char* win, int wlen, int* out6, unsigned* cigar, int cap)
{
    StripedSmithWaterman::Aligner aligner;
    StripedSmithWaterman::Filter filter;
    StripedSmithWaterman::Alignment al;
    aligner.Align(rna, win, wlen, filter, &al, 15);
    out6[0] = al.sw_score; out6[1] = al.ref_begin; out6[2] = al.ref_end; out6[3] = al.query_begin; out6[4] = al.query_end;
    out6[5] = (int)al.cigar.size();
    for (size_t i = 0; i < al.cigar.size() && (int)i < cap; ++i) cigar[i] = al.cigar[i];
    return 0;
}

// One task through the reference's calc_score_once + fastSIM (Fasim-LongTarget.cpp:413-417).
// Returns text length (negative = needed capacity); *minscore_out receives the threshold.
int ref_task(const char* rna, const char* seg, long dna_start, int para_, int strand, int rule, const int* params,
             int* minscore_out, char* out, long cap)
{
    std::vector<char> s2(strlen(seg) + 1), sr(strlen(seg) + 1);
    ref_task_strings(seg, para_, strand, rule, s2.data(), sr.data());
    std::string A(rna), B(s2.data()), S(sr.data());
    para pl = make_para(params);
    int minscore = calc_score_once(A, B, dna_start, pl.rule) * 0.8;
    if (minscore_out) *minscore_out = minscore;
    std::vector<triplex> list;
    fastSIM(A, B, S, dna_start, minscore, 5, -4, -12, -4, list, strand, para_, rule, pl.ntMin, pl.ntMax, pl.penaltyT,
            pl.penaltyC, pl);
    std::string txt;
    for (size_t i = 0; i < list.size(); ++i) append_triplex(txt, list[i]);
    return emit(txt, out, cap);
}

// One DNA record through LongTarget() (all segments, all tasks, final filter) — Fasim-LongTarget.cpp:379-598.
int ref_longtarget(const char* rna, const char* dna, const int* params, char* out, long cap)
{
    para pl = make_para(params);
    std::vector<triplex> list;
    FILE* keep = stdout;  (void)keep;
    std::streambuf* old = std::cout.rdbuf(nullptr);      // silence "dnaPos = ..." chatter
    LongTarget(pl, std::string(rna), std::string(dna), list);
    std::cout.rdbuf(old);
    std::string txt;
    for (size_t i = 0; i < list.size(); ++i) append_triplex(txt, list[i]);
    return emit(txt, out, cap);
}

// One task through calc_score_once + SIM (the -F path, Fasim-LongTarget.cpp:419-426; sim.h:410).
int ref_sim_task(const char* rna, const char* seg, long dna_start, int para_, int strand, int rule, const int* params,
                 int* minscore_out, char* out, long cap)
{
    std::vector<char> s2(strlen(seg) + 1), sr(strlen(seg) + 1);
    ref_task_strings(seg, para_, strand, rule, s2.data(), sr.data());
    std::string A(rna), B(s2.data()), S(sr.data());
    para pl = make_para(params);
    int minscore = calc_score_once(A, B, dna_start, pl.rule) * 0.8;
    if (minscore_out) *minscore_out = minscore;
    std::vector<triplex> list;
    SIM(A, B, S, dna_start, minscore, 5, -4, -12, -4, list, strand, para_, rule, pl.ntMin, pl.ntMax, pl.penaltyT, pl.penaltyC);
    std::string txt;
    for (size_t i = 0; i < list.size(); ++i) append_triplex(txt, list[i]);
    return emit(txt, out, cap);
}

// One DNA record through LongTarget() with doFastSim = false (-F).
int ref_sim_longtarget(const char* rna, const char* dna, const int* params, char* out, long cap)
{
    para pl = make_para(params);
    pl.doFastSim = false;
    std::vector<triplex> list;
    std::streambuf* old = std::cout.rdbuf(nullptr);
    LongTarget(pl, std::string(rna), std::string(dna), list);
    std::cout.rdbuf(old);
    std::string txt;
    for (size_t i = 0; i < list.size(); ++i) append_triplex(txt, list[i]);
    return emit(txt, out, cap);
}

// cluster_triplex on (stari, endi, nt) triples; returns middle/center/motif per element.
int ref_cluster(int n, const int* stari, const int* endi, const int* nt, int dd, int length, int* middle, int* center,
                int* motif)
{
    std::vector<triplex> v(n);
    for (int i = 0; i < n; ++i) {
        v[i] = triplex(stari[i], endi[i], 1, 2, 0, 1, 1, nt[i], 0.f, 0.f, 0.f, "", "", 0, 0, 0, 0, 0, 0, "");
    }
    std::map<size_t, size_t> c1[6], c1a[6], c1b[6];
    cluster_triplex(dd, length, v, c1, c1a, c1b, 5);
    for (int i = 0; i < n; ++i) { middle[i] = v[i].middle; center[i] = v[i].center; motif[i] = v[i].motif; }
    return 0;
}

}  // extern "C"
