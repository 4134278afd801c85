// DNA input of the `fasim` command line: FASTA as the reference reads it (readDna, Fasim-LongTarget.cpp:202-267), the same
// FASTA gzip/bgzip-compressed, and UCSC .2bit files (chromosome-scale inputs).  Host-only code.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace ltg_host {

// A DNA record of -f1: its bases as text (FASTA), or — for .2bit input — still 2-bit packed, the way the file stores them: the
// expansion then happens on the GPU (ltg_scan_packed) and PCIe carries 0.25 B/base.
struct FastaRecord {
    std::string species, chr; long start = 0; std::string header, seq;
    std::vector<unsigned char> packed;          // .2bit coding, covering bases [packed_first, packed_first + n_packed) of the bytes
    int64_t packed_first = 0, n_packed = 0;
    std::vector<uint32_t> n_start, n_size;      // N runs, sorted, relative to the record's first base
    bool is_packed() const { return n_packed > 0; }
    int64_t length() const { return is_packed() ? n_packed : (int64_t)seq.size(); }
    // text of bases [lo, hi) (packed records only; used by --list-records)
    std::string expand(int64_t lo, int64_t hi) const
    {
        static const char kBase[4] = {'T', 'C', 'A', 'G'};
        std::string out((size_t)(hi - lo), 'N');
        for (int64_t i = lo; i < hi; ++i) { const int64_t g = packed_first + i; out[(size_t)(i - lo)] = kBase[(packed[(size_t)(g >> 2)] >> (6 - 2 * (int)(g & 3))) & 3]; }
        for (size_t k = 0; k < n_start.size(); ++k) {
            const int64_t a = std::max<int64_t>(n_start[k], lo), e = std::min<int64_t>((int64_t)n_start[k] + n_size[k], hi);
            for (int64_t i = a; i < e; ++i) out[(size_t)(i - lo)] = 'N';
        }
        return out;
    }
};

// Lower-case (soft-masked) bases.  The reference's transferString scores every byte outside "ATGCN" as N while its
// complement() silently DROPS such bytes (rules.h:82-83, 308-311), so a soft-masked genome loses every
// repeat and gets shifted TTS strings.  This build makes the choice explicit (--softmask): upper-case them (default, with
// a notice on stderr), score them as N like the reference's scan does, or refuse the input.
enum SoftMask { kSoftUpper = 0, kSoftAsN = 1, kSoftError = 2 };

// applies the policy to one record; returns the number of lower-case letters found
inline int64_t apply_softmask(std::string& seq, int policy)
{
    int64_t n = 0;
    for (char& ch : seq) {
        if (ch >= 'a' && ch <= 'z') {
            ++n;
            if (policy == kSoftUpper) ch = (char)(ch - 'a' + 'A');
            else if (policy == kSoftAsN) ch = 'N';
        }
    }
    return n;
}

// Header ">species|chr|start-end" (Fasim-LongTarget.cpp:211-248): the first two '|' close species and chr, the first '-'
// after them closes the start (atoi); everything else is ignored.
inline void parse_dna_header(const std::string& line, FastaRecord& cur)
{
    cur.header = line.substr(1);
    std::string field, startstr;
    int bars = 0;
    for (size_t i = 1; i < line.size(); ++i) {
        const char ch = line[i];
        if (ch == '|' && bars == 0) { cur.species = field; field.clear(); ++bars; continue; }
        if (ch == '|' && bars == 1) { cur.chr = field; field.clear(); ++bars; continue; }
        if (ch == '-' && bars == 2) { startstr = field; field.clear(); continue; }
        field += ch;
    }
    cur.start = atoi(startstr.c_str());
}

// FASTA, plain or gzip/bgzip (zlib reads both, and plain files transparently).  Unlike the canonical readDna (which never
// resets its accumulator) every record is parsed on its own, like fasim-LongTarget.cpp:215-263 does.
inline bool read_dna_fasta(const std::string& path, std::vector<FastaRecord>& out)
{
    gzFile in = gzopen(path.c_str(), "rb");
    if (!in) return false;
    gzbuffer(in, 1 << 20);
    std::vector<char> buf(1 << 20);
    std::string line;
    FastaRecord cur;
    bool have = false;
    auto end_line = [&]() {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        if (!line.empty() && line[0] == '>') {
            if (have) out.push_back(std::move(cur));
            cur = FastaRecord();
            have = true;
            parse_dna_header(line, cur);
        } else if (have) {
            cur.seq += line;
        }
        line.clear();
    };
    for (;;) {
        const int got = gzread(in, buf.data(), (unsigned)buf.size());
        if (got < 0) { gzclose(in); return false; }
        if (got == 0) break;
        const char* p = buf.data();
        const char* const e = p + got;
        while (p < e) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            if (!nl) { line.append(p, e); break; }
            line.append(p, nl);
            end_line();
            p = nl + 1;
        }
    }
    if (!line.empty()) end_line();
    if (have) out.push_back(std::move(cur));
    gzclose(in);
    return true;
}

// ---- UCSC .2bit ------------------------------------------------------------------------------------------------------
// Layout (version 0): signature 0x1A412743, version, sequenceCount, reserved; index of (nameSize u8, name, offset u32);
// per sequence: dnaSize, nBlockCount, nBlockStarts[], nBlockSizes[], maskBlockCount, maskBlockStarts[], maskBlockSizes[],
// reserved, packed DNA (4 bases per byte, first base in the two high bits, T=0 C=1 A=2 G=3).  N blocks become 'N'; the
// soft-mask blocks (lower case in FASTA) are ignored: the scan works on upper-case letters.
struct TwoBitSeq { std::string name; uint64_t offset; };

class TwoBitFile {
public:
    ~TwoBitFile() { if (f_) fclose(f_); }
    static bool is_twobit(const std::string& path)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        uint32_t sig = 0;
        const bool ok = fread(&sig, 4, 1, f) == 1 && (sig == 0x1A412743u || sig == 0x4327411Au);
        fclose(f);
        return ok;
    }
    bool open(const std::string& path, std::string& err)
    {
        f_ = fopen(path.c_str(), "rb");
        if (!f_) { err = "cannot open " + path; return false; }
        if (fseeko(f_, 0, SEEK_END) != 0) { err = "cannot seek in " + path; return false; }
        file_size_ = (int64_t)ftello(f_);
        rewind(f_);
        uint32_t sig = 0, version = 0, count = 0, reserved = 0;
        if (fread(&sig, 4, 1, f_) != 1) { err = "truncated .2bit header"; return false; }
        if (sig == 0x4327411Au) swap_ = true;
        else if (sig != 0x1A412743u) { err = "not a .2bit file"; return false; }
        if (!u32(version) || !u32(count) || !u32(reserved)) { err = "truncated .2bit header"; return false; }
        if (version != 0) { err = ".2bit version " + std::to_string(version) + " is not supported (only version 0)"; return false; }
        // every index entry takes at least 5 bytes: a count the file cannot hold is a corrupt header, not an allocation size
        if ((int64_t)count * 5 > file_size_) { err = "corrupt .2bit header (sequence count exceeds the file size)"; return false; }
        for (uint32_t i = 0; i < count; ++i) {
            unsigned char len = 0;
            if (fread(&len, 1, 1, f_) != 1) { err = "truncated .2bit index"; return false; }
            std::string name(len, '\0');
            uint32_t off = 0;
            if ((len && fread(&name[0], 1, len, f_) != len) || !u32(off)) { err = "truncated .2bit index"; return false; }
            if ((int64_t)off + 16 > file_size_) { err = "corrupt .2bit index (offset of '" + name + "' is outside the file)"; return false; }
            seqs.push_back(TwoBitSeq{name, off});
        }
        return true;
    }
    // bases [lo, hi) of sequence `idx` (0-based, half open; hi < 0 or beyond the end = to the end)
    bool fetch(size_t idx, int64_t lo, int64_t hi, std::string& out, int64_t& dna_size, std::string& err, int softmask = kSoftUpper,
               int64_t* n_masked = nullptr)
    {
        if (fseeko(f_, (off_t)seqs[idx].offset, SEEK_SET) != 0) { err = "bad .2bit offset"; return false; }
        uint32_t size = 0, nb = 0, mb = 0, reserved = 0;
        if (!u32(size) || !u32(nb)) { err = "truncated .2bit record"; return false; }
        // counts come from the file: check them against what the file can hold before they size an allocation or a seek
        const int64_t left = file_size_ - (int64_t)ftello(f_);
        if ((int64_t)nb * 8 > left) { err = "corrupt .2bit record (N-block count exceeds the file size)"; return false; }
        std::vector<uint32_t> nstart(nb), nsize(nb);
        for (uint32_t& v : nstart) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        for (uint32_t& v : nsize) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        if (!u32(mb)) { err = "truncated .2bit record"; return false; }
        if ((int64_t)mb * 8 > file_size_ - (int64_t)ftello(f_)) { err = "corrupt .2bit record (mask-block count exceeds the file size)"; return false; }
        std::vector<uint32_t> mstart(mb), msize(mb);
        for (uint32_t& v : mstart) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        for (uint32_t& v : msize) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        if (!u32(reserved)) { err = "truncated .2bit record"; return false; }
        if (((int64_t)size + 3) / 4 > file_size_ - (int64_t)ftello(f_)) { err = "corrupt .2bit record (sequence size exceeds the file size)"; return false; }
        dna_size = size;
        if (lo < 0) lo = 0;
        if (hi < 0 || hi > (int64_t)size) hi = size;
        if (lo > hi) lo = hi;
        const int64_t first_byte = lo / 4, last_byte = (hi + 3) / 4;
        std::vector<unsigned char> packed((size_t)(last_byte - first_byte));
        if (fseeko(f_, (off_t)first_byte, SEEK_CUR) != 0 || (!packed.empty() && fread(packed.data(), 1, packed.size(), f_) != packed.size())) {
            err = "truncated .2bit sequence data"; return false;
        }
        out.resize((size_t)(hi - lo));
        static const char kBase[4] = {'T', 'C', 'A', 'G'};
        for (int64_t i = lo; i < hi; ++i) {
            const unsigned char b = packed[(size_t)(i / 4 - first_byte)];
            out[(size_t)(i - lo)] = kBase[(b >> (6 - 2 * (i & 3))) & 3];
        }
        for (uint32_t k = 0; k < nb; ++k) {
            const int64_t a = std::max<int64_t>(nstart[k], lo), e = std::min<int64_t>((int64_t)nstart[k] + nsize[k], hi);
            for (int64_t i = a; i < e; ++i) out[(size_t)(i - lo)] = 'N';
        }
        // soft-mask blocks = the lower-case stretches of the FASTA view
        int64_t masked = 0;
        for (uint32_t k = 0; k < mb; ++k) {
            const int64_t a = std::max<int64_t>(mstart[k], lo), e = std::min<int64_t>((int64_t)mstart[k] + msize[k], hi);
            if (e <= a) continue;
            masked += e - a;
            if (softmask == kSoftAsN) for (int64_t i = a; i < e; ++i) out[(size_t)(i - lo)] = 'N';
        }
        if (n_masked) *n_masked += masked;
        if (masked > 0 && softmask == kSoftError) { err = "sequence '" + seqs[idx].name + "' holds soft-masked (lower-case) bases; choose --softmask upper or --softmask n"; return false; }
        return true;
    }
    // the same region kept packed: the file's own bytes plus the N blocks (and, with --softmask n, the mask blocks) clipped to it
    bool fetch_packed(size_t idx, int64_t lo, int64_t hi, FastaRecord& rec, int64_t& dna_size, std::string& err, int softmask = kSoftUpper,
                      int64_t* n_masked = nullptr)
    {
        if (fseeko(f_, (off_t)seqs[idx].offset, SEEK_SET) != 0) { err = "bad .2bit offset"; return false; }
        uint32_t size = 0, nb = 0, mb = 0, reserved = 0;
        if (!u32(size) || !u32(nb)) { err = "truncated .2bit record"; return false; }
        if ((int64_t)nb * 8 > file_size_ - (int64_t)ftello(f_)) { err = "corrupt .2bit record (N-block count exceeds the file size)"; return false; }
        std::vector<uint32_t> nstart(nb), nsize(nb);
        for (uint32_t& v : nstart) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        for (uint32_t& v : nsize) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        if (!u32(mb)) { err = "truncated .2bit record"; return false; }
        if ((int64_t)mb * 8 > file_size_ - (int64_t)ftello(f_)) { err = "corrupt .2bit record (mask-block count exceeds the file size)"; return false; }
        std::vector<uint32_t> mstart(mb), msize(mb);
        for (uint32_t& v : mstart) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        for (uint32_t& v : msize) if (!u32(v)) { err = "truncated .2bit record"; return false; }
        if (!u32(reserved)) { err = "truncated .2bit record"; return false; }
        if (((int64_t)size + 3) / 4 > file_size_ - (int64_t)ftello(f_)) { err = "corrupt .2bit record (sequence size exceeds the file size)"; return false; }
        dna_size = size;
        if (lo < 0) lo = 0;
        if (hi < 0 || hi > (int64_t)size) hi = size;
        if (lo > hi) lo = hi;
        const int64_t first_byte = lo / 4, last_byte = (hi + 3) / 4;
        rec.packed.resize((size_t)(last_byte - first_byte));
        if (fseeko(f_, (off_t)first_byte, SEEK_CUR) != 0 || (!rec.packed.empty() && fread(rec.packed.data(), 1, rec.packed.size(), f_) != rec.packed.size())) {
            err = "truncated .2bit sequence data"; return false;
        }
        rec.packed_first = lo - first_byte * 4;
        rec.n_packed = hi - lo;
        // N blocks (and mask blocks scored as N) clipped to [lo, hi), relative to lo, merged into one sorted list
        std::vector<std::pair<int64_t, int64_t> > blocks;
        for (uint32_t k = 0; k < nb; ++k) {
            const int64_t a = std::max<int64_t>(nstart[k], lo), e = std::min<int64_t>((int64_t)nstart[k] + nsize[k], hi);
            if (e > a) blocks.emplace_back(a - lo, e - lo);
        }
        int64_t masked = 0;
        for (uint32_t k = 0; k < mb; ++k) {
            const int64_t a = std::max<int64_t>(mstart[k], lo), e = std::min<int64_t>((int64_t)mstart[k] + msize[k], hi);
            if (e <= a) continue;
            masked += e - a;
            if (softmask == kSoftAsN) blocks.emplace_back(a - lo, e - lo);
        }
        if (n_masked) *n_masked += masked;
        if (masked > 0 && softmask == kSoftError) { err = "sequence '" + seqs[idx].name + "' holds soft-masked (lower-case) bases; choose --softmask upper or --softmask n"; return false; }
        std::sort(blocks.begin(), blocks.end());
        rec.n_start.clear(); rec.n_size.clear();
        for (const auto& b : blocks) {
            if (!rec.n_start.empty() && b.first <= (int64_t)rec.n_start.back() + rec.n_size.back()) {
                const int64_t end = std::max<int64_t>((int64_t)rec.n_start.back() + rec.n_size.back(), b.second);
                rec.n_size.back() = (uint32_t)(end - rec.n_start.back());
            } else { rec.n_start.push_back((uint32_t)b.first); rec.n_size.push_back((uint32_t)(b.second - b.first)); }
        }
        return true;
    }
    std::vector<TwoBitSeq> seqs;

private:
    bool u32(uint32_t& v)
    {
        if (fread(&v, 4, 1, f_) != 1) return false;
        if (swap_) v = (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
        return true;
    }
    FILE* f_ = nullptr;
    bool swap_ = false;
    int64_t file_size_ = 0;
};

// `.2bit` records for the scan.  `regions` = "", or a comma-separated list of  name | name:start-end  (1-based, inclusive,
// the convention of the FASTA header's start field).  No list = every sequence of the file, whole.
inline bool read_dna_twobit(const std::string& path, const std::string& regions, const std::string& species,
                            std::vector<FastaRecord>& out, std::string& err, int softmask = kSoftUpper, int64_t* n_masked = nullptr,
                            bool keep_packed = false)
{
    TwoBitFile tb;
    if (!tb.open(path, err)) return false;
    struct Want { size_t idx; int64_t lo, hi; };
    std::vector<Want> wants;
    if (regions.empty()) {
        for (size_t i = 0; i < tb.seqs.size(); ++i) wants.push_back(Want{i, 0, -1});
    } else {
        size_t at = 0;
        while (at <= regions.size()) {
            const size_t comma = regions.find(',', at);
            const std::string tok = regions.substr(at, comma == std::string::npos ? std::string::npos : comma - at);
            if (!tok.empty()) {
                std::string name = tok;
                int64_t lo = 0, hi = -1;
                const size_t colon = tok.rfind(':');
                if (colon != std::string::npos) {
                    const std::string range = tok.substr(colon + 1);
                    const size_t dash = range.find('-');
                    char* endp = nullptr;
                    const long long a = strtoll(range.c_str(), &endp, 10);
                    if (dash != std::string::npos && endp == range.c_str() + dash && a >= 1) {
                        const long long b = strtoll(range.c_str() + dash + 1, &endp, 10);
                        if (*endp == '\0' && b >= a) { name = tok.substr(0, colon); lo = a - 1; hi = b; }
                    }
                }
                size_t idx = tb.seqs.size();
                for (size_t i = 0; i < tb.seqs.size(); ++i) if (tb.seqs[i].name == name) { idx = i; break; }
                if (idx == tb.seqs.size()) { err = "sequence '" + name + "' is not in " + path; return false; }
                wants.push_back(Want{idx, lo, hi});
            }
            if (comma == std::string::npos) break;
            at = comma + 1;
        }
    }
    for (const Want& w : wants) {
        FastaRecord r;
        int64_t size = 0;
        if (keep_packed) { if (!tb.fetch_packed(w.idx, w.lo, w.hi, r, size, err, softmask, n_masked)) return false; }
        else if (!tb.fetch(w.idx, w.lo, w.hi, r.seq, size, err, softmask, n_masked)) return false;
        r.species = species;
        r.chr = tb.seqs[w.idx].name;
        r.start = (long)w.lo + 1;
        r.header = species + "|" + r.chr + "|" + std::to_string(r.start) + "-" + std::to_string((long)(w.lo + r.length()));
        out.push_back(std::move(r));
    }
    return true;
}

}  // namespace ltg_host
