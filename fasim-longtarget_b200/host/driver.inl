// Drop-in process surface of the reference `fasim` (Fasim-LongTarget.cpp:78-172 main, :269-377 initEnv,
// :174-267 FASTA readers, :797-845 printResult, :694-795 print_cluster) on top of the C ABI above.
// Included at the end of engine.cu (same translation unit: it uses the anonymous-namespace helpers).
#include <atomic>
#include <fstream>
#include <mutex>
#include <getopt.h>
#include <time.h>

#include "dna_input.hpp"

namespace {

using ltg_host::FastaRecord;
using ltg_host::read_dna_fasta;

// readRna — Fasim-LongTarget.cpp:174-200: first line is the name (every '>' removed), all other lines are
// concatenated.
bool read_rna_fasta(const std::string& path, std::string& name, std::string& seq)
{
    std::ifstream in(path.c_str());
    if (!in) return false;
    std::string line;
    name.clear(); seq.clear();
    if (!std::getline(in, line)) return true;
    for (char ch : line) if (ch != '>' && ch != '\r' && ch != '\n') name += ch;
    while (std::getline(in, line)) {
        for (char ch : line) if (ch != '\r' && ch != '\n') seq += ch;
    }
    return true;
}

// --queries: a multi-record -f2 file, one lncRNA per '>' record (new surface, SURVEY 8d config 5: the reference reads one
// lncRNA per run and would splice a second header line into the sequence).  Names are cleaned like readRna does.
bool read_rna_fasta_multi(const std::string& path, std::vector<std::pair<std::string, std::string> >& out)
{
    std::ifstream in(path.c_str());
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '>') {
            out.emplace_back();
            for (char ch : line) if (ch != '>' && ch != '\r' && ch != '\n') out.back().first += ch;
        } else if (!out.empty()) {
            for (char ch : line) if (ch != '\r' && ch != '\n') out.back().second += ch;
        }
    }
    return true;
}

void usage()
{
    printf("fasim (B200 build) — genome-wide lncRNA:DNA triplex scan\n"
           "  -f1 <dna.fa>   DNA FASTA, header >species|chr|start-end\n"
           "  -f2 <rna.fa>   lncRNA FASTA\n"
           "  -O  <dir/>     output directory (must exist)\n"
           "  -r N  rule (0 = all)          -t N  strand (0 both, >0 parallel, <0 anti-parallel)\n"
           "  -c N  segment length (5000)   -o N  segment overlap (100)\n"
           "  -i N  min identity (60)       -S N  min stability (1)\n"
           "  -ni N min triplex nt (20)     -na N max triplex nt (100000)\n"
           "  -pc N penalty C (0)           -pt N penalty T (-1000)\n"
           "  -ds N cluster distance (15)   -lg N min length for clustering (50)\n"
           "  --device N  CUDA device (default 0)\n"
           "  --devices a,b,..|all   several GPUs in one process: one context and one host thread per GPU pull chunks of the\n"
           "                         file from a shared queue (results are identical to a single-GPU run)\n"
           "  -f1 also takes a gzip/bgzip-compressed FASTA and a UCSC .2bit file; for .2bit:\n"
           "  --seq name[:start-end][,...]   sequences / 1-based inclusive regions to scan (default: every sequence, whole)\n"
           "  --species S                    species field of the output file name (default: the .2bit file's base name)\n"
           "  --softmask upper|n|error       lower-case (soft-masked) bases of -f1: upper-case them (default; a notice is printed),\n"
           "                                 score them as N (what the reference's scan does), or refuse the input\n"
           "  --compat lowercase   behave like the older fasim-LongTarget.cpp / fastSim.h variant (window loop without start clamp, no\n"
           "                       per-task filter) and write <species>-<lncRNA>-fastSim-TFOsorted only\n"
           "  --list-records   print the DNA records -f1 and the lncRNAs -f2 yield (name, start, length, CRC-32) and exit (no GPU needed)\n"
           "  --queries   -f2 holds several lncRNAs (one per '>' record): every lncRNA is scanned against -f1 and gets its own\n"
           "              output files; the (lncRNA, chunk) pairs go through the same queue\n");
}

}  // namespace

extern "C" {

int ltg_write_tfosorted(const ltg_result* r, const char* path)
{
    return guarded([&]() -> int {
    if (!r || !path) { set_error("null argument"); return LTG_ERR_ARG; }
    FILE* f = fopen(path, "w");
    if (!f) { set_error("cannot write %s", path); return LTG_ERR_IO; }
    fputs("QueryStart\tQueryEnd\tStartInSeq\tEndInSeq\tDirection\tChr\tStartInGenome\tEndInGenome\tMeanStability\t"
          "MeanIdentity(%)\tStrand\tRule\tScore\tNt(bp)\tClass\tMidPoint\tCenter\tTFO sequence\tTTS sequence\n", f);
    const char* chr = r->text + (r->n_triplex ? 0 : 0);
    (void)chr;
    for (int64_t i = 0; i < r->n_triplex; ++i) {
        const ltg_triplex& t = r->triplex[i];
        if (t.motif == 0) continue;                             // Fasim-LongTarget.cpp:819-822
        const char* chrname = r->text + t.chr_off;
        fprintf(f, "%d\t%d\t%d\t%d\t%s\t%s\t%ld\t%ld\t%g\t%g\t%s\t%d\t%g\t%d\t%d\t%d\t%d\t%s\t%s\n", t.stari, t.endi, t.starj, t.endj,
                t.starj < t.endj ? "R" : "L", chrname, (long)t.genomestart, (long)t.genomeend, (double)t.tri_score, (double)t.identity,
                ltg_host::strand_name(t.reverse, t.strand), t.rule, (double)t.score, t.nt, t.motif, t.middle, t.center,
                r->text + t.tfo_off, r->text + t.tts_off);
    }
    fclose(f);
    return LTG_OK;
    });
}

// print_cluster — Fasim-LongTarget.cpp:694-795 for class levels 1 and 2 (called from printResult :831-836
// with start_genome - 1).  Restated from the coverage map semantics; see the line comments.
int ltg_write_tfoclass(const ltg_result* r, const ltg_params* p, const char* sorted_path, const char* chr, int64_t record_start,
                       int64_t dna_size, const char* rna_name)
{
    return guarded([&]() -> int {
    if (!r || !p || !sorted_path) { set_error("null argument"); return LTG_ERR_ARG; }
    // rebuild the per-class coverage maps exactly as cluster_triplex fills class1[] (:661-672)
    std::map<size_t, size_t> cov[6];
    for (int64_t i = 0; i < r->n_triplex; ++i) {
        const ltg_triplex& t = r->triplex[i];
        if (t.motif < 1 || t.motif > 5) continue;
        if (t.endj > t.starj) for (int j = t.starj; j < t.endj; ++j) cov[t.motif][(size_t)j]++;
        else for (int j = t.endj; j < t.starj; ++j) cov[t.motif][(size_t)j]++;
    }
    const std::string base(sorted_path);
    const long start_genome = (long)record_start - 1;
    for (int level = 1; level <= 2; ++level) {
        char name[64];
        snprintf(name, sizeof name, "-TFOclass%d-%d-%d", level, p->c_distance, p->c_length);
        const std::string path = base.substr(0, base.size() >= 10 ? base.size() - 10 : 0) + name;     // strips "-TFOsorted" (:706)
        FILE* f = fopen(path.c_str(), "w");
        if (!f) { set_error("cannot write %s", path.c_str()); return LTG_ERR_IO; }
        fprintf(f, "browser position %s:%ld-%ld\n", chr, start_genome, start_genome + (long)dna_size);
        fputs("browser hide all\nbrowser pack refGene encodeRegions\nbrowser full altGraph\n"
              "# 300 base wide bar graph, ausoScale is on by default == graphing\n"
              "# limits will dynamically change to always show full range of data\n"
              "# in viewing window, priority = 20 position this as the second graph\n"
              "# Note, zero-relative, half-open coordinate system in use for bedGraph format\n", f);
        fprintf(f, "track type=bedGraph name='%s TTS (%d)' description='%d-%d' visibility=full color=200,100,0 altColor=0,100,200 priority=20\n",
                rna_name, level, p->c_distance, p->c_length);
        const std::map<size_t, size_t>& m = cov[level];
        struct Row { long a, b, v; };
        std::vector<Row> rows;
        long final_genome = 0;
        for (auto& kv : m) final_genome = (long)kv.first + start_genome;                              // :723-726
        int count = 0;
        for (auto it = m.begin(); it != m.end();) {                                                   // :727-766
            const long first0 = (long)it->first;
            long tmp1 = (long)it->first, tmp2 = (long)it->second;
            if (tmp1 + start_genome == final_genome) { rows.push_back({first0 + start_genome - 1, tmp1 + start_genome, tmp2}); break; }
            ++it;
            while (it != m.end() && labs((long)it->first - tmp1) == 1 && (long)it->second == tmp2) {
                if ((long)it->first + start_genome == final_genome) break;
                tmp1 = (long)it->first; tmp2 = (long)it->second;
                ++it;
            }
            rows.push_back({first0 + start_genome - (count == 0 ? 2 : 1), tmp1 + start_genome, tmp2});
            ++count;
            if (it != m.end() && labs((long)it->first - tmp1) != 1) rows.push_back({tmp1 + start_genome, (long)it->first + start_genome - 1, 0});
        }
        for (const Row& rw : rows) fprintf(f, "%s\t%ld\t%ld\t%ld\n", chr, rw.a, rw.b, rw.v);
        fclose(f);
    }
    return LTG_OK;
    });
}

int ltg_main(int argc, char* const* argv)
{
    return guarded([&]() -> int {
    ltg_params P;
    ltg_default_params(&P);
    std::string f1 = "./", f2 = "./", outdir = "./", devices_arg;
    int device = 0;
    // same option table as initEnv (Fasim-LongTarget.cpp:271-283); -m, -d, -cn, -F are accepted and ignored
    // (-F selects the SIM path: ltg_set_sim_mode)
    const char* optstring = "f:s:r:O:c:m:t:i:S:z:Y:Z:h:C:D:E:o:y:Fd";
    static const struct option long_options[] = {
        {"f1", required_argument, nullptr, 'f'}, {"f2", required_argument, nullptr, 's'}, {"ni", required_argument, nullptr, 'y'},
        {"na", required_argument, nullptr, 'z'}, {"pc", required_argument, nullptr, 'Y'}, {"pt", required_argument, nullptr, 'Z'},
        {"cn", required_argument, nullptr, 'C'}, {"ds", required_argument, nullptr, 'D'}, {"lg", required_argument, nullptr, 'E'},
        {"device", required_argument, nullptr, 1000}, {"devices", required_argument, nullptr, 1001}, {"queries", no_argument, nullptr, 1002},
        {"seq", required_argument, nullptr, 1003}, {"species", required_argument, nullptr, 1004}, {"list-records", no_argument, nullptr, 1005},
        {"softmask", required_argument, nullptr, 1006}, {"compat", required_argument, nullptr, 1007},
        {nullptr, 0, nullptr, 0}};
    if (argc <= 1) { usage(); return 1; }
    optind = 1;
    int opt;
    bool want_sim = false, multi_query = false, list_records = false, compat_lc = false;
    std::string seq_arg, species_arg;
    int softmask = ltg_host::kSoftUpper;
    while ((opt = getopt_long_only(argc, argv, optstring, long_options, nullptr)) != -1) {
        switch (opt) {
        case 'f': f1 = optarg; break;
        case 's': f2 = optarg; break;
        case 'r': P.rule = atoi(optarg); break;
        case 'O': outdir = optarg; break;
        case 'c': P.cut_length = atoi(optarg); break;
        case 't': P.strand = atoi(optarg); break;
        case 'i': P.min_identity = atoi(optarg); break;
        case 'S': P.min_stability = atoi(optarg); break;
        case 'y': P.nt_min = atoi(optarg); break;
        case 'z': P.nt_max = atoi(optarg); break;
        case 'Y': P.penalty_c = atoi(optarg); break;
        case 'Z': P.penalty_t = atoi(optarg); break;
        case 'o': P.overlap = atoi(optarg); break;
        case 'D': P.c_distance = atoi(optarg); break;
        case 'E': P.c_length = atoi(optarg); break;
        case 'F': want_sim = true; break;
        case 'h': usage(); return 1;
        case 1000: device = atoi(optarg); break;
        case 1001: devices_arg = optarg; break;
        case 1002: multi_query = true; break;
        case 1003: seq_arg = seq_arg.empty() ? std::string(optarg) : seq_arg + "," + optarg; break;
        case 1004: species_arg = optarg; break;
        case 1005: list_records = true; break;
        case 1007:
            if (!strcmp(optarg, "lowercase")) compat_lc = true;
            else if (strcmp(optarg, "canonical")) { fprintf(stderr, "fasim: --compat takes canonical or lowercase\n"); return 2; }
            break;
        case 1006:
            if (!strcmp(optarg, "upper")) softmask = ltg_host::kSoftUpper;
            else if (!strcmp(optarg, "n") || !strcmp(optarg, "N")) softmask = ltg_host::kSoftAsN;
            else if (!strcmp(optarg, "error")) softmask = ltg_host::kSoftError;
            else { fprintf(stderr, "fasim: --softmask takes upper, n or error\n"); return 2; }
            break;
        default: break;
        }
    }
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    const bool timing = getenv("LTG_TIMING") != nullptr;
    auto lap = [&](const char* what) {
        if (!timing) return;
        struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[fasim timing] %-10s at %.3f s\n", what, (t.tv_sec - t0.tv_sec) + 1e-9 * (t.tv_nsec - t0.tv_nsec));
    };
    printf(want_sim ? "Searching triplexes using Sim\n" : "Searching triplexes using Fasim\n");        // Fasim-LongTarget.cpp:105-108
    // base name of -f1: the reference drops the last 3 characters (".fa", Fasim-LongTarget.cpp:800); the new input kinds
    // drop their own extension (".2bit"; ".gz" and then 3 more)
    std::string base = f1;
    {
        const size_t slash = base.find_last_of('/');
        if (slash != std::string::npos) base = base.substr(slash + 1);
    }
    auto ends_with = [](const std::string& s, const char* suf) { const size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; };
    std::vector<FastaRecord> recs;
    int64_t n_lower = 0;                      // soft-masked (lower-case) bases met in -f1
    if (ltg_host::TwoBitFile::is_twobit(f1)) {
        if (ends_with(base, ".2bit")) base = base.substr(0, base.size() - 5);
        std::string err;
        if (!ltg_host::read_dna_twobit(f1, seq_arg, species_arg.empty() ? base : species_arg, recs, err, softmask, &n_lower, /*keep_packed=*/true)) { fprintf(stderr, "fasim: %s\n", err.c_str()); return 2; }
    } else {
        if (!seq_arg.empty()) { fprintf(stderr, "fasim: --seq needs a .2bit file as -f1\n"); return 2; }
        if (!read_dna_fasta(f1, recs)) recs.clear();
        if (ends_with(base, ".gz")) base = base.substr(0, base.size() - 3);
        base = base.substr(0, base.size() >= 3 ? base.size() - 3 : 0);
        if (!species_arg.empty()) for (FastaRecord& r : recs) r.species = species_arg;
        for (FastaRecord& r : recs) n_lower += ltg_host::apply_softmask(r.seq, softmask);
        if (n_lower > 0 && softmask == ltg_host::kSoftError) {
            fprintf(stderr, "fasim: %s holds %lld lower-case (soft-masked) bases; choose --softmask upper or --softmask n\n", f1.c_str(), (long long)n_lower);
            return 2;
        }
    }
    if (n_lower > 0 && softmask == ltg_host::kSoftUpper)
        fprintf(stderr, "fasim: note: %lld lower-case (soft-masked) bases of %s were upper-cased (--softmask upper, the default); the reference "
                        "scores them as N (--softmask n)\n", (long long)n_lower, f1.c_str());
    else if (n_lower > 0 && softmask == ltg_host::kSoftAsN)
        fprintf(stderr, "fasim: note: %lld lower-case (soft-masked) bases of %s are scored as N (--softmask n)\n", (long long)n_lower, f1.c_str());
    if (recs.empty()) { fprintf(stderr, "fasim: cannot read DNA file %s\n", f1.c_str()); return 2; }
    if (list_records) {
        for (const FastaRecord& r : recs)
        {
            const std::string text = r.is_packed() ? r.expand(0, r.length()) : std::string();
            const std::string& sq = r.is_packed() ? text : r.seq;
            printf("record\t%s\t%s\t%ld\t%zu\t%08lx\n", r.species.c_str(), r.chr.c_str(), r.start, sq.size(),
                   (unsigned long)crc32(0L, (const Bytef*)sq.data(), (uInt)sq.size()));
        }
    }
    std::vector<std::pair<std::string, std::string> > queries;            // (name, sequence); one entry unless --queries
    if (multi_query) {
        if (!read_rna_fasta_multi(f2, queries)) queries.clear();
        for (size_t q = 0; q < queries.size();) { if (queries[q].second.empty()) queries.erase(queries.begin() + q); else ++q; }
    } else {
        std::string lnc_name, lnc;
        if (read_rna_fasta(f2, lnc_name, lnc) && !lnc.empty()) queries.emplace_back(lnc_name, lnc);
    }
    if (queries.empty()) { fprintf(stderr, "fasim: cannot read RNA file %s\n", f2.c_str()); return 2; }
    for (const auto& q : queries) printf("%s\n", q.first.c_str());
    if (list_records) {                                             // ... and the lncRNAs -f2 yields; nothing is scanned
        for (const auto& q : queries)
            printf("query\t%s\t%zu\t%08lx\n", q.first.c_str(), q.second.size(), (unsigned long)crc32(0L, (const Bytef*)q.second.data(), (uInt)q.second.size()));
        return 0;
    }
    lap("read");

    // devices: --device N, or --devices a,b,.. / all (one context + one host thread per entry; an entry may repeat)
    std::vector<int> devs;
    if (devices_arg == "all") { for (int d = 0; d < ltg_device_count(); ++d) devs.push_back(d); }
    else if (!devices_arg.empty()) {
        size_t at = 0;
        while (at <= devices_arg.size()) {
            const size_t comma = devices_arg.find(',', at);
            const std::string tok = devices_arg.substr(at, comma == std::string::npos ? std::string::npos : comma - at);
            if (!tok.empty()) devs.push_back(atoi(tok.c_str()));
            if (comma == std::string::npos) break;
            at = comma + 1;
        }
    }
    if (devs.empty()) devs.push_back(device);

    // Work units in file order: runs of whole records (short records share device batches), or shards of
    // kUnitSegments segments of a long record.  GPUs pull units from one atomic queue; the results are appended in unit order,
    // which is the order a single ltg_scan_records call over the file would produce.
    struct Unit { size_t r0, r1; int64_t first_seg, n_seg; };
    std::vector<Unit> units;
    {
        const int64_t stride = P.cut_length - P.overlap;
        // Unit size: large enough that a call's pipeline fill / drain (~10 ms) is noise, small enough that every GPU gets
        // at least ~8 jobs, so the tail of the queue (the last job of the slowest GPU) stays a small fraction of the run:
        // 1024 .. 16384 segments (5 .. 80 Mbp), chosen from the total work of the run.
        int64_t total_segs = 0;
        if (stride > 0) for (const FastaRecord& r : recs) total_segs += (r.length() + stride - 1) / stride;
        const int64_t want = total_segs * (int64_t)queries.size() / ((int64_t)devs.size() * 8);
        const int64_t kUnitSegments = devs.size() > 1 ? std::max<int64_t>(1024, std::min<int64_t>(16384, want)) : 2048;
        const int64_t unit_bases = kUnitSegments * (stride > 0 ? stride : 1);
        if (stride <= 0) { fprintf(stderr, "fasim: cut length (%d) must exceed the overlap (%d)\n", P.cut_length, P.overlap); return 2; }
        size_t i = 0;
        while (i < recs.size()) {
            const int64_t n = recs[i].length();
            if (n > unit_bases + stride && devs.size() > 1) {                 // a long record: shards of whole segments
                const int64_t n_seg = (n + stride - 1) / stride;
                for (int64_t s0 = 0; s0 < n_seg; s0 += kUnitSegments) units.push_back(Unit{i, i + 1, s0, std::min(kUnitSegments, n_seg - s0)});
                ++i;
                continue;
            }
            size_t j = i;
            int64_t bytes = 0;
            const int64_t cap = devs.size() > 1 ? unit_bases : (256ll << 20);
            for (; j < recs.size() && (j == i || bytes + recs[j].length() <= cap); ++j) {
                if (j > i && recs[j].length() > unit_bases + stride && devs.size() > 1) break;
                bytes += recs[j].length();
            }
            units.push_back(Unit{i, j, 0, -1});
            i = j;
        }
    }
    // Jobs = (lncRNA, unit) pairs, lncRNA-major, so the GPUs work on the same lncRNA most of the time and a context changes
    // its query only when it pulls a job of the next one.  The worker that completes the last unit of a lncRNA merges,
    // clusters and writes that lncRNA's files, so results do not pile up over a many-query run.
    const size_t n_units = units.size(), n_jobs = n_units * queries.size();
    std::vector<ltg_result*> results(n_jobs, nullptr);
    std::vector<std::atomic<size_t> > remaining(queries.size());
    for (auto& r : remaining) r.store(n_units);
    std::vector<size_t> q_order(queries.size());              // longest lncRNA first: the queue's tail is made of short jobs
    for (size_t q = 0; q < q_order.size(); ++q) q_order[q] = q;
    std::stable_sort(q_order.begin(), q_order.end(), [&](size_t a, size_t b) { return queries[a].second.size() > queries[b].second.size(); });
    std::vector<std::thread> finishers;                       // merge / cluster / write of a finished lncRNA, off the GPU threads
    std::mutex finishers_mu;
    std::vector<std::string> finish_errors;
    std::atomic<size_t> next_job(0);
    std::atomic<int> failed(0);
    std::vector<std::string> errors(devs.size());
    // <O>/<species>-<lncName>-<f1 minus last 3 chars>-TFOsorted (Fasim-LongTarget.cpp:123, 800-802); the directory part of
    // -f1 is dropped (the reference embeds it and then silently fails to open the file, Q14)
    auto finish_query = [&](size_t q) -> int {
        ltg_result* all = nullptr;
        int rc = ltg_result_new(&all);
        for (size_t u = 0; rc == LTG_OK && u < n_units; ++u) {
            ltg_result* part = results[q * n_units + u];
            if (!part) { set_error("internal: a work unit of %s has no result", queries[q].first.c_str()); rc = LTG_ERR_STATE; break; }
            for (int64_t k = 0; k < part->n_triplex; ++k) part->triplex[k].record += (int32_t)units[u].r0;
            rc = ltg_result_append(all, part);
        }
        for (size_t u = 0; u < n_units; ++u) if (results[q * n_units + u]) { ltg_result_free(results[q * n_units + u]); results[q * n_units + u] = nullptr; }
        if (rc == LTG_OK) rc = ltg_cluster(all, &P);
        const std::string& lnc_name = queries[q].first;
        // (the older variant: <species>-<lncName>-fastSim-TFOsorted and no -TFOclass files, fasim-LongTarget.cpp:883)
        const std::string out_path = compat_lc ? outdir + "/" + recs[0].species + "-" + lnc_name + "-fastSim-TFOsorted"
                                               : outdir + "/" + recs[0].species + "-" + lnc_name + "-" + base + "-TFOsorted";
        if (rc == LTG_OK) rc = ltg_write_tfosorted(all, out_path.c_str());
        if (rc == LTG_OK && !compat_lc) rc = ltg_write_tfoclass(all, &P, out_path.c_str(), recs[0].chr.c_str(), recs[0].start, recs[0].length(), lnc_name.c_str());
        if (rc == LTG_OK && all->scan_cells > 0)
            printf("[b200] %s: segments=%ld tasks=%ld peaks=%ld scan_cells=%.3e gpu_scan_ms=%.2f gpu_window_ms=%.2f literal_tasks=%ld literal_windows=%ld\n",
                   lnc_name.c_str(), (long)all->n_segments, (long)all->n_tasks, (long)all->n_peaks, (double)all->scan_cells, all->gpu_ms_scan,
                   all->gpu_ms_window, (long)all->n_literal_tasks, (long)all->n_literal_windows);
        if (all) ltg_result_free(all);
        return rc;
    };
    // bases [lo, hi) of a packed record as one ltg_scan_packed call (segments first_seg .. of the whole record)
    auto scan_packed_range = [&](ltg_context* ctx, const FastaRecord& R, int64_t lo, int64_t hi, int64_t first_seg, int64_t n_seg, ltg_result** out) -> int {
        std::vector<uint32_t> ns, nz;
        for (size_t k = 0; k < R.n_start.size(); ++k) {
            const int64_t a = std::max<int64_t>(R.n_start[k], lo), e = std::min<int64_t>((int64_t)R.n_start[k] + R.n_size[k], hi);
            if (e > a) { ns.push_back((uint32_t)(a - lo)); nz.push_back((uint32_t)(e - a)); }
        }
        return ltg_scan_packed(ctx, R.packed.data(), 0, R.packed_first + lo, hi - lo, ns.data(), nz.data(), (int32_t)ns.size(), R.chr.c_str(), R.start,
                               R.length(), first_seg, n_seg, out);
    };
    auto gpu_worker = [&](size_t w) {
        ltg_context* ctx = nullptr;
        auto tnow = []() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
        double t_create = tnow(), t_query = 0, t_scan = 0;
        int rc = ltg_create(devs[w], &ctx);
        t_create = tnow() - t_create;
        if (rc == LTG_OK) rc = ltg_set_params(ctx, &P);
        if (rc == LTG_OK && want_sim) rc = ltg_set_sim_mode(ctx, 1);
        if (rc == LTG_OK && compat_lc) rc = ltg_set_compat(ctx, 1);
        size_t cur_q = (size_t)-1;
        while (rc == LTG_OK && !failed.load()) {
            const size_t job = next_job.fetch_add(1);
            if (job >= n_jobs) break;
            const size_t q = q_order[job / n_units], u = job % n_units;
            if (q != cur_q) {
                const double t0q = tnow();
                rc = ltg_set_query(ctx, queries[q].first.c_str(), queries[q].second.c_str(), (int64_t)queries[q].second.size());
                t_query += tnow() - t0q;
                if (rc != LTG_OK) break;
                cur_q = q;
            }
            const double t0s = tnow();
            const Unit& U = units[u];
            ltg_result** out = &results[q * n_units + u];
            if (U.n_seg >= 0) {                                  // shard of one long record
                const FastaRecord& R = recs[U.r0];
                const int64_t stride = P.cut_length - P.overlap, lo = U.first_seg * stride;
                const int64_t hi = std::min<int64_t>(R.length(), (U.first_seg + U.n_seg - 1) * stride + P.cut_length);
                if (R.is_packed()) rc = scan_packed_range(ctx, R, lo, hi, U.first_seg, U.n_seg, out);
                else rc = ltg_scan_shard(ctx, R.seq.data() + lo, 0, hi - lo, R.chr.c_str(), R.start, R.length(), U.first_seg, U.n_seg, out);
            } else {
                bool any_packed = false;
                for (size_t r = U.r0; r < U.r1; ++r) any_packed |= recs[r].is_packed();
                if (any_packed) {
                    // .2bit records: one ltg_scan_packed call per record, results appended in record order
                    rc = ltg_result_new(out);
                    for (size_t r = U.r0; rc == LTG_OK && r < U.r1; ++r) {
                        ltg_result* one = nullptr;
                        rc = scan_packed_range(ctx, recs[r], 0, recs[r].length(), 0, -1, &one);
                        if (rc == LTG_OK) { for (int64_t k = 0; k < one->n_triplex; ++k) one->triplex[k].record = (int32_t)(r - U.r0); rc = ltg_result_append(*out, one); }
                        if (one) ltg_result_free(one);
                    }
                } else {
                    std::vector<const char*> dna, chr;
                    std::vector<int64_t> len, start;
                    for (size_t r = U.r0; r < U.r1; ++r) {
                        dna.push_back(recs[r].seq.data()); len.push_back((int64_t)recs[r].seq.size());
                        chr.push_back(recs[r].chr.c_str()); start.push_back(recs[r].start);
                    }
                    rc = ltg_scan_records(ctx, (int64_t)(U.r1 - U.r0), dna.data(), len.data(), chr.data(), start.data(), out);
                }
            }
            t_scan += tnow() - t0s;
            if (rc == LTG_OK && remaining[q].fetch_sub(1) == 1) {
                if (queries.size() == 1) rc = finish_query(q);
                else {
                    std::lock_guard<std::mutex> lk(finishers_mu);
                    finishers.emplace_back([&, q]() { if (finish_query(q) != LTG_OK) { std::lock_guard<std::mutex> lk2(finishers_mu); finish_errors.push_back(ltg_last_error()); failed.store(1); } });
                }
            }
        }
        if (rc != LTG_OK) { errors[w] = ltg_last_error(); failed.store(1); }
        const double t0d = tnow();
        if (ctx) ltg_destroy(ctx);
        if (timing) fprintf(stderr, "[fasim timing] device %d: create %.3f s, set_query %.3f s, scan calls %.3f s, destroy %.3f s\n",
                            devs[w], t_create, t_query, t_scan, tnow() - t0d);
    };
    {
        std::vector<std::thread> pool;
        for (size_t w = 1; w < devs.size(); ++w) pool.emplace_back(gpu_worker, w);
        gpu_worker(0);
        for (std::thread& t : pool) t.join();
    }
    for (std::thread& t : finishers) t.join();
    for (const std::string& e : finish_errors) errors.push_back(e);
    lap("scan+write");
    for (ltg_result* r : results) if (r) ltg_result_free(r);
    if (failed.load()) {
        for (const std::string& e : errors) if (!e.empty()) fprintf(stderr, "fasim: %s\n", e.c_str());
        return 3;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    printf("finished normally\nRunning time is %g\n", secs);
    return 0;
    });
}

}  // extern "C"
