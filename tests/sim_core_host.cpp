// TEST INFRASTRUCTURE ONLY: compiles the product's sequential SIM core (csrc/sim_core.cuh) and its host-side row conversion
// (host/sim_host.hpp) for the CPU so that tests/test_sim_cpu.py can pin them against the oracle / the reference shim without
// a GPU.  On the device the first pass is a warp wavefront (csrc/sim.cuh); here it is first_pass_serial, its specification.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../fasim-longtarget_b200/host/rules_table.hpp"
#include "../fasim-longtarget_b200/host/sim_host.hpp"

using namespace ltg;

extern "C" int simcore_task(const char* rna, const char* seg, long dna_start, int para, int strand, int rule, int min_score, const int* params,
                            char* out, long cap)
{
    ltg_params P;
    P.rule = params[0]; P.cut_length = params[1]; P.strand = params[2]; P.overlap = params[3]; P.nt_min = params[4]; P.nt_max = params[5];
    P.min_identity = params[6]; P.min_stability = params[7]; P.penalty_t = params[8]; P.penalty_c = params[9]; P.c_distance = params[10]; P.c_length = params[11];
    TaskDef td;
    if (!ltg_host::make_task(para, strand, rule, td)) return -1;
    const int M = (int)strlen(rna), N = (int)strlen(seg);
    std::vector<uint8_t> a(M), b(N);
    for (int i = 0; i < M; ++i) a[i] = (uint8_t)dna_code((unsigned char)rna[i]);
    for (int q = 0; q < N; ++q) b[q] = (uint8_t)td.img[dna_code((unsigned char)seg[td.reversed ? N - 1 - q : q])];
    std::vector<simk::cand_t> CC(N + 2), DD(N + 2), HH(M + 2), WW(M + 2);
    std::vector<int> c1(N + 2), d1(N + 2), c2(N + 2), d2(N + 2), used_head(M + 2, -1), used_col(1 << 16), used_next(1 << 16), script(1 << 18);
    std::vector<simk::Node> list(simk::kNodes);
    std::vector<simk::Aln> alns(simk::kNodes + 1);
    simk::Task T;
    memset(&T, 0, sizeof T);
    T.a = a.data(); T.b = b.data(); T.M = M; T.N = N; T.min_score = min_score;
    T.CC = CC.data(); T.DD = DD.data(); T.HH = HH.data(); T.WW = WW.data();
    T.c1 = c1.data(); T.d1 = d1.data(); T.c2 = c2.data(); T.d2 = d2.data();
    T.used_head = used_head.data(); T.used_col = used_col.data(); T.used_next = used_next.data(); T.used_cap = (int)used_col.size();
    T.list = list.data(); T.script = script.data(); T.script_cap = (int)script.size(); T.out = alns.data(); T.out_cap = (int)alns.size();
    if (M > 0 && N > 0) {
        simk::first_pass_serial(T);
        simk::best_alignments(T);
    }
    if (T.error) return -100 - T.error;
    std::vector<ltg_host::SimRow> rows;
    for (int k = 0; k < T.n_out; ++k) ltg_host::sim_convert(alns[k], script.data() + alns[k].script_off, rna, td, seg, N, dna_start, P, rows);
    std::string txt;
    for (const ltg_host::SimRow& r : rows) {
        char buf[256];
        unsigned s, id, tr;
        memcpy(&s, &r.t.score, 4); memcpy(&id, &r.t.identity, 4); memcpy(&tr, &r.t.tri_score, 4);
        snprintf(buf, sizeof buf, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%08x\t%08x\t%08x\t0\t0\t0\t0\t0\t0\t", r.t.stari, r.t.endi, r.t.starj, r.t.endj,
                 r.t.strand, r.t.reverse, r.t.rule, r.t.nt, s, id, tr);
        txt += buf; txt += r.tfo; txt += '\t'; txt += r.tts; txt += "\t\n";
    }
    if ((long)txt.size() + 1 > cap) return -2;
    memcpy(out, txt.c_str(), txt.size() + 1);
    return (int)txt.size();
}
