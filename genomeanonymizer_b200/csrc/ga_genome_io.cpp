// ga_genome_io.cpp - host-side readers behind include/ga_genome_io.h: BGZF/BAM -> the
// structure-of-arrays read batch of ga_b200.h, FASTA -> contig bases.  zlib only; every stage that touches all
// the bytes (inflate + CRC, record packing) runs on all host threads.
//
// Formats restated from the SAM/BAM specification (SAMv1 4.1 BGZF, 4.2 BAM): a BGZF file is a series of gzip
// members of at most 64 KiB with a 'BC' extra subfield holding the member size; the inflated stream is
// "BAM\1", header text, the reference dictionary, then alignment records (block_size, refID, pos, l_read_name,
// mapq, bin, n_cigar_op, flag, l_seq, next_refID, next_pos, tlen, name, CIGAR words, 4-bit bases HIGH nibble
// first, qualities, tags).  The reference reads the same files through pysam / htslib
// (short_read_tumor_normal_anonymizer.py:661-664, pileup_io.pyx:12-17, 138-139).
#include "../../include/ga_genome_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

int n_workers(int requested) {
    if (requested > 0) return requested;
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

template <class F> void parallel_for(int64_t n, int threads, int64_t grain, F&& body) {
    if (n <= 0) return;
    const int64_t n_chunks = (n + grain - 1) / grain;
    const int t = (int)std::min<int64_t>(threads, n_chunks);
    if (t <= 1) { body(0, n); return; }
    std::atomic<int64_t> next(0);
    std::vector<std::thread> pool;
    pool.reserve(t);
    for (int k = 0; k < t; ++k)
        pool.emplace_back([&] {
            for (;;) {
                const int64_t c = next.fetch_add(1);
                if (c >= n_chunks) return;
                body(c * grain, std::min(n, (c + 1) * grain));
            }
        });
    for (auto& th : pool) th.join();
}

// Read-only view of a whole file (mmap: the compressed BAM is consumed straight from the page cache).
struct FileView {
    const uint8_t* p = nullptr; size_t n = 0; int fd = -1;
    bool open(const char* path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) return false;
        n = (size_t)st.st_size;
        if (n == 0) return true;
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = static_cast<const uint8_t*>(m);
        return true;
    }
    const uint8_t* data() const { return p; }
    size_t size() const { return n; }
    ~FileView() { if (p) munmap(const_cast<uint8_t*>(p), n); if (fd >= 0) ::close(fd); }
};

bool read_file(const char* path, std::vector<uint8_t>& out) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out.resize((size_t)n);
    const size_t got = n ? std::fread(out.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

struct Block { uint64_t in_off; uint32_t in_size; uint64_t out_off; uint32_t out_size; };

// Walks the gzip members of a BGZF file; false when the container is malformed.
bool index_bgzf(const FileView& file, std::vector<Block>& blocks, uint64_t* total, std::string* why) {
    uint64_t off = 0, out = 0;
    const uint64_t n = file.size();
    while (off < n) {
        if (n - off < 18) { *why = "truncated BGZF block header"; return false; }
        const uint8_t* h = file.data() + off;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { *why = "not a BGZF block (gzip magic / FEXTRA missing)"; return false; }
        const uint32_t xlen = le16(h + 10);
        if (n - off < 12 + xlen) { *why = "truncated BGZF extra field"; return false; }
        int64_t bsize = -1;
        for (uint32_t x = 0; x + 4 <= xlen;) {
            const uint8_t* sf = h + 12 + x;
            const uint32_t slen = le16(sf + 2);
            if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (int64_t)le16(sf + 4) + 1;
            x += 4 + slen;
        }
        if (bsize < (int64_t)(12 + xlen + 8) || (uint64_t)bsize > n - off) { *why = "BGZF block size field missing or beyond the file"; return false; }
        const uint32_t isize = le32(h + bsize - 4);
        if (isize > 65536u) { *why = "BGZF block inflates to more than 64 KiB"; return false; }
        blocks.push_back({off, (uint32_t)bsize, out, isize});
        off += (uint64_t)bsize;
        out += isize;
    }
    *total = out;
    return true;
}

bool inflate_block(const uint8_t* src, uint32_t src_size, uint8_t* dst, uint32_t dst_size) {
    const uint32_t xlen = le16(src + 10);
    const uint8_t* cdata = src + 12 + xlen;
    const uint32_t clen = src_size - 12 - xlen - 8;
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef*>(cdata); zs.avail_in = clen;
    zs.next_out = dst; zs.avail_out = dst_size;
    const int rc = inflate(&zs, Z_FINISH);
    const bool ok = (rc == Z_STREAM_END) && zs.total_out == dst_size;
    inflateEnd(&zs);
    if (!ok) return false;
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, dst_size);
    return crc == le32(src + src_size - 8);
}

inline int ref_span_of(const uint8_t* cigar, uint32_t n_ops) {
    int span = 0;
    for (uint32_t k = 0; k < n_ops; ++k) {
        const uint32_t w = le32(cigar + 4 * k), op = w & 15u;
        if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += (int)(w >> 4);
    }
    return span;
}

}  // namespace

struct ga_bam {
    // the inflated BAM stream; not zero-initialised - its pages are first touched by the inflating threads
    struct Bytes {
        std::unique_ptr<uint8_t[]> p; uint64_t n = 0;
        void resize(uint64_t k) { p.reset(new uint8_t[k ? k : 1]); n = k; }
        uint8_t* data() { return p.get(); }
        const uint8_t* data() const { return p.get(); }
        uint64_t size() const { return n; }
    } data;
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::vector<std::vector<uint64_t>> by_ref;    // per reference: offsets (of refID, i.e. past block_size) into data, in file order
    int64_t n_records = 0;
    // Large files whose records are grouped by reference (every sorted BAM) are not kept inflated: open() only walks
    // them once, window by window, to find each reference's span of the uncompressed stream, and the records of ONE
    // reference at a time are inflated when they are asked for (memory = one contig, not the file).
    bool lazy = false;
    FileView file;                                // stays mapped in lazy mode
    std::vector<Block> blocks;
    uint64_t total = 0;                           // size of the uncompressed stream
    int threads = 0;
    struct Span { uint64_t u_begin = 0, u_end = 0; int64_t n = 0; };
    std::vector<Span> spans;                      // per reference, lazy mode
    int loaded_ref = -1;
};

namespace {

// Inflates the blocks that hold bytes [u_lo, u_hi) of the uncompressed stream; *base = stream offset of out[0].
bool inflate_range(const FileView& file, const std::vector<Block>& blocks, uint64_t u_lo, uint64_t u_hi, int threads, ga_bam::Bytes& out, uint64_t* base) {
    if (u_hi <= u_lo || blocks.empty()) { out.resize(0); *base = u_lo; return true; }
    size_t lo = (size_t)(std::upper_bound(blocks.begin(), blocks.end(), u_lo, [](uint64_t v, const Block& b) { return v < b.out_off; }) - blocks.begin());
    lo = lo ? lo - 1 : 0;
    size_t hi = (size_t)(std::lower_bound(blocks.begin(), blocks.end(), u_hi, [](const Block& b, uint64_t v) { return b.out_off < v; }) - blocks.begin());
    *base = blocks[lo].out_off;
    const uint64_t end = hi < blocks.size() ? blocks[hi].out_off : blocks.back().out_off + blocks.back().out_size;
    out.resize(end - *base);
    std::atomic<int> bad(0);
    parallel_for((int64_t)(hi - lo), n_workers(threads), 64, [&](int64_t a, int64_t z) {
        for (int64_t k = a; k < z; ++k) {
            const Block& bl = blocks[lo + (size_t)k];
            if (bl.out_size == 0) continue;                               // the EOF marker block
            if (!inflate_block(file.data() + bl.in_off, bl.in_size, out.data() + (bl.out_off - *base), bl.out_size)) bad.store(1);
        }
    });
    return !bad.load();
}

// Parses "BAM\1", the header text and the reference dictionary from the head of the stream; returns the offset of the
// first alignment record, 0 when the buffer does not hold the whole header yet, -1 when it is not a BAM header.
int64_t parse_header(const uint8_t* d, uint64_t n, std::vector<std::string>& names, std::vector<int64_t>& lens) {
    names.clear(); lens.clear();
    if (n < 12) return 0;
    if (std::memcmp(d, "BAM\1", 4) != 0) return -1;
    uint64_t off = 4;
    const uint32_t l_text = le32(d + off); off += 4;
    if (off + l_text + 4 > n) return 0;
    off += l_text;
    const uint32_t n_ref = le32(d + off); off += 4;
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (off + 4 > n) return 0;
        const uint32_t l_name = le32(d + off); off += 4;
        if (l_name == 0) return -1;
        if (off + l_name + 4 > n) return 0;
        names.emplace_back(reinterpret_cast<const char*>(d + off), l_name - 1);
        off += l_name;
        lens.push_back((int64_t)le32(d + off)); off += 4;
    }
    return (int64_t)off;
}

// SAM text (SAM v1.6 section 1) -> the uncompressed BAM stream the rest of the reader works on (section 4.2): header
// dictionary from the @SQ lines, one record per alignment line.  Optional fields are dropped (nothing on the path reads
// them); a '*' sequence is a record without stored bases, '*' qualities are 0xff as in BAM.
bool sam_to_stream(const uint8_t* text, size_t n, std::vector<uint8_t>& out, std::string* why) {
    auto put32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) out.push_back((uint8_t)(v >> (8 * k))); };
    auto put16 = [&](uint32_t v) { out.push_back((uint8_t)v); out.push_back((uint8_t)(v >> 8)); };
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    size_t p = 0;
    while (p < n && text[p] == '@') {                                  // header lines
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(text + p, '\n', n - p));
        const size_t e = nl ? (size_t)(nl - text) : n;
        if (e - p >= 3 && text[p + 1] == 'S' && text[p + 2] == 'Q') {
            std::string sn; long long ln = -1;
            size_t f = p;
            while (f < e) {
                size_t g = f;
                while (g < e && text[g] != '\t') ++g;
                if (g - f > 3 && text[f + 2] == ':') {
                    if (text[f] == 'S' && text[f + 1] == 'N') { sn.assign(reinterpret_cast<const char*>(text + f + 3), g - f - 3); }
                    else if (text[f] == 'L' && text[f + 1] == 'N') ln = std::strtoll(std::string(reinterpret_cast<const char*>(text + f + 3), g - f - 3).c_str(), nullptr, 10);
                }
                f = g + 1;
            }
            while (!sn.empty() && sn.back() == '\r') sn.pop_back();
            if (sn.empty() || ln < 0 || ln > 0x7fffffffLL) { *why = "@SQ line without SN / LN"; return false; }
            names.push_back(sn); lens.push_back((uint32_t)ln);
        }
        p = nl ? e + 1 : n;
    }
    out.clear();
    out.insert(out.end(), {'B', 'A', 'M', 1});
    put32(0);                                                          // l_text: the text is not kept
    put32((uint32_t)names.size());
    for (size_t k = 0; k < names.size(); ++k) {
        put32((uint32_t)names[k].size() + 1);
        out.insert(out.end(), names[k].begin(), names[k].end());
        out.push_back(0);
        put32(lens[k]);
    }
    auto ref_of = [&](const uint8_t* s, size_t len, int same) -> int {
        if (len == 1 && s[0] == '*') return -1;
        if (len == 1 && s[0] == '=' && same > -2) return same;
        for (size_t k = 0; k < names.size(); ++k) if (names[k].size() == len && std::memcmp(names[k].data(), s, len) == 0) return (int)k;
        return -2;
    };
    uint8_t code_of[256];
    std::memset(code_of, 15, sizeof code_of);                          // anything else is N, as in htslib's table
    { const char* t = "=ACMGRSVTWYHKDBN"; for (int k = 0; k < 16; ++k) { code_of[(uint8_t)t[k]] = (uint8_t)k; code_of[(uint8_t)std::tolower(t[k])] = (uint8_t)k; } }
    while (p < n) {                                                    // alignment lines
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(text + p, '\n', n - p));
        size_t e = nl ? (size_t)(nl - text) : n;
        const size_t next = nl ? e + 1 : n;
        if (e > p && text[e - 1] == '\r') --e;
        if (e == p) { p = next; continue; }
        const uint8_t* f[11]; size_t fl[11];
        size_t q = p; int nf = 0;
        while (nf < 11) {
            size_t g = q;
            while (g < e && text[g] != '\t') ++g;
            f[nf] = text + q; fl[nf] = g - q; ++nf;
            if (g >= e) break;
            q = g + 1;
        }
        if (nf < 11) { *why = "alignment line with fewer than 11 fields"; return false; }
        auto num = [&](int k) { return std::strtoll(std::string(reinterpret_cast<const char*>(f[k]), fl[k]).c_str(), nullptr, 10); };
        const int ref = ref_of(f[2], fl[2], -2);
        if (ref == -2) { *why = "alignment line names a reference without an @SQ line"; return false; }
        int nref = ref_of(f[6], fl[6], ref);
        if (nref == -2) nref = -1;
        if (fl[0] == 0 || fl[0] > 254) { *why = "read name longer than 254 characters"; return false; }
        std::vector<uint32_t> ops;
        if (!(fl[5] == 1 && f[5][0] == '*')) {
            uint64_t len = 0; bool have = false;
            for (size_t k = 0; k < fl[5]; ++k) {
                const uint8_t ch = f[5][k];
                if (ch >= '0' && ch <= '9') { len = len * 10 + (ch - '0'); have = true; if (len >= (1u << 28)) { *why = "CIGAR operation longer than 2^28"; return false; } continue; }
                const char* t = "MIDNSHP=X";
                const char* at = std::strchr(t, (char)ch);
                if (!at || !have || ch == 0) { *why = "malformed CIGAR"; return false; }
                ops.push_back((uint32_t)(len << 4) | (uint32_t)(at - t));
                len = 0; have = false;
            }
            if (have) { *why = "malformed CIGAR"; return false; }
        }
        if (ops.size() > 65535) { *why = "more than 65535 CIGAR operations: not supported"; return false; }
        const bool no_seq = fl[9] == 1 && f[9][0] == '*';
        const uint32_t l_seq = no_seq ? 0u : (uint32_t)fl[9];
        const bool no_qual = fl[10] == 1 && f[10][0] == '*';               // also for a one-base read, as htslib reads it
        if (!no_qual && !no_seq && fl[10] != fl[9]) { *why = "SEQ and QUAL differ in length"; return false; }
        const uint32_t l_name = (uint32_t)fl[0] + 1;
        put32(32 + l_name + 4 * (uint32_t)ops.size() + (l_seq + 1) / 2 + l_seq);
        put32((uint32_t)ref);
        put32((uint32_t)(num(3) - 1));
        out.push_back((uint8_t)l_name);
        out.push_back((uint8_t)num(4));
        put16(0);                                                      // bin: nothing on the path reads it
        put16((uint32_t)ops.size());
        put16((uint32_t)num(1));
        put32(l_seq);
        put32((uint32_t)nref);
        put32((uint32_t)(num(7) - 1));
        put32((uint32_t)num(8));
        out.insert(out.end(), f[0], f[0] + fl[0]);
        out.push_back(0);
        for (const uint32_t w : ops) put32(w);
        for (uint32_t k = 0; k < l_seq; k += 2) out.push_back((uint8_t)((code_of[f[9][k]] << 4) | (k + 1 < l_seq ? code_of[f[9][k + 1]] : 0)));
        for (uint32_t k = 0; k < l_seq; ++k) out.push_back(no_qual ? (uint8_t)0xff : (uint8_t)(f[10][k] - 33));
        p = next;
    }
    return true;
}

uint64_t env_u64(const char* name, uint64_t dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::strtoull(v, nullptr, 10) : dflt;
}

}  // namespace

struct ga_fasta {
    std::vector<std::string> names;
    std::vector<std::string> seqs;                // newline-free bases, case kept
};

extern "C" {

const char* ga_io_last_error(void) { return g_err.c_str(); }
int ga_io_set_error(int code, const char* msg) { return fail(code, msg ? msg : ""); }

// Indexes the records in data[from, data.size()) per reference (one hop per record); data_base = stream offset of data[0].
static int index_records(ga_bam* b, uint64_t from, const std::string& path) {
    const uint8_t* d = b->data.data();
    const uint64_t n = b->data.size();
    const uint32_t n_ref = (uint32_t)b->ref_names.size();
    uint64_t off = from;
    while (off < n) {
        if (off + 4 > n) return fail(GA_IO_ERR_FORMAT, path + ": truncated alignment record");
        const uint32_t block_size = le32(d + off);
        if (block_size < 32 || off + 4 + block_size > n) return fail(GA_IO_ERR_FORMAT, path + ": alignment record runs past the end of the stream");
        const int32_t ref_id = (int32_t)le32(d + off + 4);
        if (ref_id >= 0 && (uint32_t)ref_id < n_ref) b->by_ref[ref_id].push_back(off + 4);
        else if (ref_id != -1) return fail(GA_IO_ERR_FORMAT, path + ": alignment record names a reference that is not in the header");
        b->n_records++;                                                   // records without a reference count too
        off += 4 + (uint64_t)block_size;
    }
    return GA_IO_OK;
}

// Lazy mode: one walk over the file, window by window, that finds the header and each reference's span of the
// uncompressed stream.  *grouped = false when a reference's records come in more than one run (unsorted file).
static int scan_spans(ga_bam* b, const std::string& path, bool* grouped) {
    uint64_t window = std::max<uint64_t>(env_u64("GA_BAM_WINDOW_BYTES", 256ull << 20), 1u << 16);
    ga_bam::Bytes buf;
    uint64_t base = 0;
    int64_t hdr = 0;
    for (uint64_t want = window;; want *= 2) {                             // the header: grow until it is whole
        if (!inflate_range(b->file, b->blocks, 0, std::min(b->total, want), b->threads, buf, &base))
            return fail(GA_IO_ERR_FORMAT, path + ": a BGZF block failed to inflate or its CRC32 does not match");
        hdr = parse_header(buf.data(), buf.size(), b->ref_names, b->ref_lens);
        if (hdr < 0) return fail(GA_IO_ERR_FORMAT, path + ": BAM magic missing");
        if (hdr > 0) break;
        if (want >= b->total) return fail(GA_IO_ERR_FORMAT, path + ": truncated header");
    }
    const int32_t n_ref = (int32_t)b->ref_names.size();
    b->spans.assign((size_t)n_ref, ga_bam::Span());
    std::vector<uint8_t> seen((size_t)n_ref, 0);
    *grouped = true;
    int32_t last = -2;
    uint64_t cur = (uint64_t)hdr;
    while (cur < b->total) {
        if (!inflate_range(b->file, b->blocks, cur, std::min(b->total, cur + window), b->threads, buf, &base))
            return fail(GA_IO_ERR_FORMAT, path + ": a BGZF block failed to inflate or its CRC32 does not match");
        const uint8_t* d = buf.data();
        const uint64_t n = buf.size();
        uint64_t p = cur - base;
        bool progressed = false;
        while (p + 4 <= n) {
            const uint32_t bs = le32(d + p);
            if (bs < 32) return fail(GA_IO_ERR_FORMAT, path + ": alignment record shorter than its fixed fields");
            if (p + 4 + bs > n) break;                                   // continues in the next window
            const int32_t ref = (int32_t)le32(d + p + 4);
            if (ref < -1 || ref >= n_ref) return fail(GA_IO_ERR_FORMAT, path + ": alignment record names a reference that is not in the header");
            if (ref != last) {
                if (last >= 0) b->spans[last].u_end = base + p;
                if (ref >= 0) {
                    if (seen[ref]) *grouped = false;
                    seen[ref] = 1;
                    b->spans[ref].u_begin = base + p;
                }
                last = ref;
            }
            if (ref >= 0) b->spans[ref].n++;
            b->n_records++;
            p += 4 + (uint64_t)bs;
            progressed = true;
        }
        if (!progressed) {
            if (base + n >= b->total) return fail(GA_IO_ERR_FORMAT, path + ": alignment record runs past the end of the stream");
            window *= 2;                                                 // a record longer than the window
            continue;
        }
        cur = base + p;
    }
    if (last >= 0) b->spans[last].u_end = cur;
    return GA_IO_OK;
}

// Lazy mode: makes the records of reference ref_id the loaded ones.
static int ensure_loaded(const ga_bam* cb, int ref_id) {
    ga_bam* b = const_cast<ga_bam*>(cb);                                 // the handle caches one contig; single-threaded use
    if (!b->lazy || b->loaded_ref == ref_id) return GA_IO_OK;
    for (auto& v : b->by_ref) { v.clear(); v.shrink_to_fit(); }
    b->loaded_ref = -1;
    const ga_bam::Span& sp = b->spans[(size_t)ref_id];
    uint64_t base = 0;
    if (!inflate_range(b->file, b->blocks, sp.u_begin, sp.u_end, b->threads, b->data, &base))
        return fail(GA_IO_ERR_FORMAT, "a BGZF block failed to inflate or its CRC32 does not match");
    if (sp.n) {
        // only this reference's records: the span ends where the next reference begins
        const uint64_t from = sp.u_begin - base, to = sp.u_end - base;
        const uint8_t* d = b->data.data();
        b->by_ref[(size_t)ref_id].reserve((size_t)sp.n);
        for (uint64_t off = from; off < to;) {
            const uint32_t bs = le32(d + off);
            b->by_ref[(size_t)ref_id].push_back(off + 4);
            off += 4 + (uint64_t)bs;
        }
    }
    b->loaded_ref = ref_id;
    return GA_IO_OK;
}

int ga_bam_open(const char* path, int n_threads, ga_bam** out) {
    if (!path || !out) return fail(GA_IO_ERR_ARGUMENT, "ga_bam_open: NULL argument");
    *out = nullptr;
    std::unique_ptr<ga_bam> b(new ga_bam());
    b->threads = n_threads;
    if (!b->file.open(path)) return fail(GA_IO_ERR_OPEN, std::string("cannot read ") + path);
    std::string why;
    const bool is_text = b->file.size() > 0 && b->file.data()[0] != 31;    // no gzip magic: SAM text (pysam.AlignmentFile reads it too)
    if (is_text) {
        std::vector<uint8_t> stream;
        if (!sam_to_stream(b->file.data(), b->file.size(), stream, &why)) return fail(GA_IO_ERR_FORMAT, std::string(path) + ": SAM: " + why);
        b->total = stream.size();
        b->data.resize(stream.size());
        std::memcpy(b->data.data(), stream.data(), stream.size());
    } else if (!index_bgzf(b->file, b->blocks, &b->total, &why)) return fail(GA_IO_ERR_FORMAT, std::string(path) + ": " + why);
    // Files above the threshold are opened lazily when their records are grouped by reference
    if (!is_text && b->total > env_u64("GA_BAM_EAGER_BYTES", 2ull << 30)) {
        bool grouped = false;
        const int rc = scan_spans(b.get(), path, &grouped);
        if (rc != GA_IO_OK) return rc;
        if (grouped) {
            b->lazy = true;
            b->by_ref.resize(b->ref_names.size());
            *out = b.release();
            return GA_IO_OK;
        }
        b->n_records = 0; b->spans.clear();                                // interleaved references: keep the whole stream
    }
    uint64_t base = 0;
    if (!is_text && !inflate_range(b->file, b->blocks, 0, b->total, n_threads, b->data, &base))
        return fail(GA_IO_ERR_FORMAT, std::string(path) + ": a BGZF block failed to inflate or its CRC32 does not match");
    const int64_t hdr = parse_header(b->data.data(), b->data.size(), b->ref_names, b->ref_lens);
    if (hdr < 0) return fail(GA_IO_ERR_FORMAT, std::string(path) + ": BAM magic missing");
    if (hdr == 0) return fail(GA_IO_ERR_FORMAT, std::string(path) + ": truncated header");
    b->by_ref.resize(b->ref_names.size());
    const int rc = index_records(b.get(), (uint64_t)hdr, path);
    if (rc != GA_IO_OK) return rc;
    *out = b.release();
    return GA_IO_OK;
}

void ga_bam_close(ga_bam* b) { delete b; }
int ga_bam_n_references(const ga_bam* b) { return b ? (int)b->ref_names.size() : 0; }
const char* ga_bam_reference_name(const ga_bam* b, int ref_id) {
    return (b && ref_id >= 0 && ref_id < (int)b->ref_names.size()) ? b->ref_names[ref_id].c_str() : nullptr;
}
int64_t ga_bam_reference_length(const ga_bam* b, int ref_id) {
    return (b && ref_id >= 0 && ref_id < (int)b->ref_lens.size()) ? b->ref_lens[ref_id] : -1;
}
int64_t ga_bam_n_records(const ga_bam* b) { return b ? b->n_records : 0; }
int64_t ga_bam_inflated_bytes(const ga_bam* b) { return b ? (int64_t)b->total : 0; }

// Fixed part of an alignment record, p = address of refID.
struct RecView {
    int32_t pos; uint32_t l_name, n_cigar, flag, l_seq;
    const uint8_t *name, *cigar, *seq, *qual;
};
static inline RecView view_of(const uint8_t* p) {
    RecView v;
    v.pos = (int32_t)le32(p + 4);
    v.l_name = p[8];
    v.n_cigar = le16(p + 12);
    v.flag = le16(p + 14);
    v.l_seq = le32(p + 16);
    v.name = p + 32;
    v.cigar = v.name + v.l_name;
    v.seq = v.cigar + 4ull * v.n_cigar;
    v.qual = v.seq + (v.l_seq + 1) / 2;
    return v;
}
static inline uint32_t units_of(uint32_t l_seq) { const uint32_t u = (l_seq + 31u) / 32u; return u ? u : 1u; }
// A CIGAR of more than 65,535 ops lives in the CG:B,I tag; its place holds "<l_seq>S<reference span>N" (SAM v1.6 4.2.2).
static inline bool cigar_in_cg_tag(const RecView& v) {
    if (v.n_cigar != 2 || v.l_seq == 0) return false;
    const uint32_t c0 = le32(v.cigar), c1 = le32(v.cigar + 4);
    return (c0 & 15u) == 4u && (c0 >> 4) == v.l_seq && (c1 & 15u) == 3u;
}

// What one slice of a contig's records adds to the totals (both passes over the records run slice by slice on all
// host threads: one record is one hop through the inflated stream, i.e. one cache miss).
struct SliceSums {
    int64_t n_reads = 0, units = 0, n_cigar = 0, name_bytes = 0;
    int32_t max_ref_span = 0, first_pos = 0, last_pos = 0;
    bool sorted = true;
    int err = GA_IO_OK;
    const char* msg = nullptr;
};
constexpr int64_t kSliceRecords = 8192;

static SliceSums sum_slice(const ga_bam* b, const std::vector<uint64_t>& offs, int64_t lo, int64_t hi, uint32_t flag_exclude) {
    SliceSums s;
    int32_t prev = INT32_MIN;
    for (int64_t k = lo; k < hi; ++k) {
        const uint8_t* p = b->data.data() + offs[(size_t)k];
        const RecView v = view_of(p);
        if (v.flag & flag_exclude) continue;
        if (v.l_seq > 0xffffu) { s.err = GA_IO_ERR_UNSUPPORTED; s.msg = "reads longer than 65535 bases are not supported"; return s; }
        const uint32_t block_size = le32(p - 4);
        if (32ull + v.l_name + 4ull * v.n_cigar <= block_size && cigar_in_cg_tag(v)) {
            s.err = GA_IO_ERR_UNSUPPORTED; s.msg = "a record keeps its CIGAR in the CG tag (more than 65535 ops): not supported"; return s;
        }
        if (32ull + v.l_name + 4ull * v.n_cigar + (v.l_seq + 1) / 2 + v.l_seq > block_size) {
            s.err = GA_IO_ERR_FORMAT; s.msg = "alignment record fields exceed its block size"; return s;
        }
        if (!s.n_reads) s.first_pos = v.pos;
        s.n_reads++;
        s.units += units_of(v.l_seq);
        s.n_cigar += v.n_cigar;
        s.name_bytes += v.l_name ? v.l_name - 1 : 0;
        s.max_ref_span = std::max(s.max_ref_span, (int32_t)ref_span_of(v.cigar, v.n_cigar));
        if (v.pos < prev) s.sorted = false;
        prev = v.pos;
        s.last_pos = v.pos;
    }
    return s;
}

static int sum_slices(const ga_bam* b, int ref_id, uint32_t flag_exclude, int n_threads, std::vector<SliceSums>& slices) {
    const std::vector<uint64_t>& offs = b->by_ref[ref_id];
    const int64_t n = (int64_t)offs.size();
    slices.assign((size_t)((n + kSliceRecords - 1) / kSliceRecords), SliceSums());
    parallel_for(n, n_workers(n_threads), kSliceRecords, [&](int64_t lo, int64_t hi) {
        slices[(size_t)(lo / kSliceRecords)] = sum_slice(b, offs, lo, hi, flag_exclude);
    });
    for (const SliceSums& s : slices) if (s.err != GA_IO_OK) return fail(s.err, s.msg);
    return GA_IO_OK;
}

int ga_bam_contig_sizes(const ga_bam* b, int ref_id, uint32_t flag_exclude, ga_bam_sizes* out) {
    if (!b || !out || ref_id < 0 || ref_id >= (int)b->by_ref.size()) return fail(GA_IO_ERR_ARGUMENT, "ga_bam_contig_sizes: bad argument");
    { const int rc = ensure_loaded(b, ref_id); if (rc != GA_IO_OK) return rc; }
    std::vector<SliceSums> slices;
    { const int rc = sum_slices(b, ref_id, flag_exclude, b->threads, slices); if (rc != GA_IO_OK) return rc; }
    ga_bam_sizes s;
    std::memset(&s, 0, sizeof(s));
    s.sorted = 1;
    int32_t prev = INT32_MIN;
    for (const SliceSums& c : slices) {
        if (!c.n_reads) continue;
        s.n_reads += c.n_reads; s.seq16_units += c.units; s.n_cigar += c.n_cigar; s.name_bytes += c.name_bytes;
        s.max_ref_span = std::max(s.max_ref_span, c.max_ref_span);
        if (!c.sorted || c.first_pos < prev) s.sorted = 0;
        prev = c.last_pos;
    }
    *out = s;
    return GA_IO_OK;
}

int ga_bam_pack_contig(const ga_bam* b, int ref_id, uint32_t flag_exclude, const ga_bam_dest* dst, int n_threads) {
    if (!b || !dst || ref_id < 0 || ref_id >= (int)b->by_ref.size()) return fail(GA_IO_ERR_ARGUMENT, "ga_bam_pack_contig: bad argument");
    if (!dst->pos || !dst->len_flag || !dst->seq_off16 || !dst->cigar_off || !dst->seq4 || !dst->cigar)
        return fail(GA_IO_ERR_ARGUMENT, "ga_bam_pack_contig: NULL destination array");
    { const int rc = ensure_loaded(b, ref_id); if (rc != GA_IO_OK) return rc; }
    // ---- pass 1 (parallel, one hop per record): what every slice of records holds; a scan over the slices says where each starts
    std::vector<SliceSums> slices;
    { const int rc = sum_slices(b, ref_id, flag_exclude, n_threads, slices); if (rc != GA_IO_OK) return rc; }
    struct Start { int64_t read; uint64_t unit, cigar, name; };
    std::vector<Start> starts(slices.size() + 1);
    starts[0] = Start{0, dst->seq16_base, dst->cigar_base, dst->name_base};
    for (size_t c = 0; c < slices.size(); ++c)
        starts[c + 1] = Start{starts[c].read + slices[c].n_reads, starts[c].unit + (uint64_t)slices[c].units, starts[c].cigar + (uint64_t)slices[c].n_cigar,
                              starts[c].name + (uint64_t)slices[c].name_bytes};
    const Start& total = starts.back();
    if (total.unit > 0xffffffffull || total.cigar > 0xffffffffull) return fail(GA_IO_ERR_UNSUPPORTED, "batch exceeds the 32-bit record offsets: pack the contig in chunks");
    const std::vector<uint64_t>& offs = b->by_ref[ref_id];
    const int64_t n_all = (int64_t)offs.size();
    dst->cigar_off[total.read] = (uint32_t)total.cigar;
    if (dst->name_off) dst->name_off[total.read] = total.name;
    // ---- pass 2 (parallel): offsets and record bytes of every slice
    parallel_for(n_all, n_workers(n_threads), kSliceRecords, [&](int64_t lo, int64_t hi) {
        Start at = starts[(size_t)(lo / kSliceRecords)];
        for (int64_t r = lo; r < hi; ++r) {
            const RecView v = view_of(b->data.data() + offs[(size_t)r]);
            if (v.flag & flag_exclude) continue;
            const int64_t k = at.read++;
            dst->seq_off16[k] = (uint32_t)at.unit;
            dst->cigar_off[k] = (uint32_t)at.cigar;
            if (dst->name_off) dst->name_off[k] = at.name;
            dst->pos[k] = v.pos;
            dst->len_flag[k] = (v.flag << 16) | v.l_seq;
            if (dst->ref_end) dst->ref_end[k] = v.pos + ref_span_of(v.cigar, v.n_cigar);
            std::memcpy(dst->cigar + at.cigar, v.cigar, 4ull * v.n_cigar);                // BAM words are little endian, as is the host
            const uint32_t cap = units_of(v.l_seq) * 16u, nb = (v.l_seq + 1) / 2;
            uint8_t* s = dst->seq4 + 16ull * at.unit;
            for (uint32_t j = 0; j < nb; ++j) s[j] = (uint8_t)((v.seq[j] >> 4) | (v.seq[j] << 4));   // HIGH-nibble-first -> LOW-nibble-first
            std::memset(s + nb, 0, cap - nb);
            if (dst->qual) {
                uint8_t* q = dst->qual + 32ull * at.unit;
                std::memcpy(q, v.qual, v.l_seq);
                std::memset(q + v.l_seq, 0, 2ull * cap - v.l_seq);
            }
            if (dst->names && dst->name_off && v.l_name) std::memcpy(dst->names + at.name, v.name, v.l_name - 1);
            at.unit += units_of(v.l_seq); at.cigar += v.n_cigar; at.name += v.l_name ? v.l_name - 1 : 0;
        }
    });
    return GA_IO_OK;
}

// ------------------------------------------------------------------------------------------------ FASTA
int ga_fasta_open(const char* path, ga_fasta** out) {
    if (!path || !out) return fail(GA_IO_ERR_ARGUMENT, "ga_fasta_open: NULL argument");
    *out = nullptr;
    std::vector<uint8_t> file;
    if (!read_file(path, file)) return fail(GA_IO_ERR_OPEN, std::string("cannot read ") + path);
    if (file.size() >= 2 && file[0] == 31 && file[1] == 139) {
        // gzip or bgzip (what pysam.FastaFile reads besides plain text): a sequence of gzip members, inflated one after the
        // other; zlib checks every member's CRC32 and length
        std::vector<uint8_t> text;
        text.reserve(file.size() * 4);
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, 15 + 16) != Z_OK) return fail(GA_IO_ERR_FORMAT, std::string(path) + ": zlib refused to start");
        zs.next_in = file.data();
        uint64_t left = file.size();
        std::vector<uint8_t> chunk(1u << 22);
        bool ok = true, in_member = false;
        for (;;) {
            if (zs.avail_in == 0 && left) { const uint64_t n = std::min<uint64_t>(left, 1u << 30); zs.avail_in = (uInt)n; left -= n; }
            if (zs.avail_in == 0 && !in_member) break;                   // the end of the file, between members
            zs.next_out = chunk.data(); zs.avail_out = (uInt)chunk.size();
            in_member = true;
            const int rc = inflate(&zs, Z_NO_FLUSH);
            text.insert(text.end(), chunk.data(), chunk.data() + (chunk.size() - zs.avail_out));
            if (rc == Z_STREAM_END) {                                    // one member done; another may follow
                in_member = false;
                if (inflateReset(&zs) != Z_OK) { ok = false; break; }
            } else if (rc != Z_OK || (zs.avail_in == 0 && left == 0 && zs.avail_out != 0)) { ok = false; break; }   // corrupt or truncated
        }
        inflateEnd(&zs);
        if (!ok) return fail(GA_IO_ERR_FORMAT, std::string(path) + ": the gzip stream is corrupt or truncated");
        file.swap(text);
    }
    ga_fasta* f = new ga_fasta();
    const uint8_t* p = file.data();
    const uint8_t* end = p + file.size();
    while (p < end) {
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(p, '\n', (size_t)(end - p)));
        const uint8_t* le = nl ? nl : end;
        const uint8_t* stop = le;
        if (stop > p && stop[-1] == '\r') --stop;
        if (p < stop && *p == '>') {
            const uint8_t* q = p + 1;
            while (q < stop && *q != ' ' && *q != '\t') ++q;                 // the name ends at the first blank
            f->names.emplace_back(reinterpret_cast<const char*>(p + 1), (size_t)(q - p - 1));
            f->seqs.emplace_back();
        } else if (p < stop) {
            if (f->seqs.empty()) { delete f; return fail(GA_IO_ERR_FORMAT, std::string(path) + ": sequence data before the first '>' line"); }
            f->seqs.back().append(reinterpret_cast<const char*>(p), (size_t)(stop - p));
        }
        p = nl ? nl + 1 : end;
    }
    if (f->names.empty()) { delete f; return fail(GA_IO_ERR_FORMAT, std::string(path) + ": no FASTA record"); }
    *out = f;
    return GA_IO_OK;
}

void ga_fasta_close(ga_fasta* f) { delete f; }
int ga_fasta_n_references(const ga_fasta* f) { return f ? (int)f->names.size() : 0; }
const char* ga_fasta_reference_name(const ga_fasta* f, int idx) {
    return (f && idx >= 0 && idx < (int)f->names.size()) ? f->names[idx].c_str() : nullptr;
}
int64_t ga_fasta_reference_length(const ga_fasta* f, int idx) {
    return (f && idx >= 0 && idx < (int)f->seqs.size()) ? (int64_t)f->seqs[idx].size() : -1;
}
int64_t ga_fasta_fetch(const ga_fasta* f, int idx, int64_t start, int64_t end, uint8_t* out) {
    if (!f || !out || idx < 0 || idx >= (int)f->seqs.size()) return fail(GA_IO_ERR_ARGUMENT, "ga_fasta_fetch: bad argument");
    const int64_t n = (int64_t)f->seqs[idx].size();
    if (start < 0) start = 0;
    if (end > n) end = n;
    if (end <= start) return 0;
    std::memcpy(out, f->seqs[idx].data() + start, (size_t)(end - start));
    return end - start;
}

}  // extern "C"
