// ga_plan.cpp - the host-side plan of one contig of a tumor-normal sample behind include/ga_plan.h
// (SURVEY.md 8(f) N2 + N3): sections, island sessions, yield order inside a session, mate pairing, first write wins.
// Same algorithm as genomeanonymizer_b200/driver.py: plan_sample (which stays as its checker); the reference code each
// step restates is cited there and below.
#include "../../include/ga_plan.h"
#include "../../include/ga_genome_io.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>

extern "C" int ga_io_set_error(int code, const char* msg);   // ga_genome_io.cpp

struct ga_plan {
    std::vector<int32_t> s_first, s_last, s_window;
    std::vector<int32_t> pairs;      // 5 per row
    std::vector<int32_t> singles;    // 4 per row: dataset, read, version, pairs written before it was stored
};

namespace {

constexpr int64_t kNone = INT64_MIN;                       // an empty mate slot
constexpr int32_t kReapply = 1 << 30;                      // flag on a version: the read's left-over indels are applied twice (GA_PLAN_REAPPLY)
inline int64_t slot_of(int32_t read, int32_t version) { return ((int64_t)read << 32) | (uint32_t)version; }
inline int32_t slot_read(int64_t v) { return (int32_t)(v >> 32); }
inline int32_t slot_version(int64_t v) { return (int32_t)(uint32_t)v; }

struct Planner {
    int64_t n, n_tumor;
    const int32_t *pos, *end;
    const uint32_t* lf;
    std::vector<int32_t> name_id;                          // dense id of every read's name
    int32_t n_names = 0;
    int span = 0;
    ga_plan* plan;

    // to_pair_anonymized_reads: insertion-ordered map name -> [mate 1, mate 2] (a popped name re-enters at the end)
    struct Entry { int32_t name; int64_t slot[2]; bool alive; int64_t ins_at; };
    // ins_at: number of pairs that were (or, for a region's deferred records, will be) in the file before a pair completed by
    // this entry's read would be written - where a mate carried over from an earlier contig joins it
    std::vector<Entry> entries;
    std::vector<int32_t> entry_of;                         // name id -> index into entries, -1 when absent
    std::vector<uint8_t> written;                          // written_read_ids

    // per-session scratch, indexed by name id through a stamp
    std::vector<int32_t> stamp, local_of;
    int32_t cur_stamp = 0;

    int dataset(int32_t i) const { return i < n_tumor ? 0 : 1; }
    int mate(int32_t i) const { return ((lf[i] >> 16) & 0x40u) ? 0 : 1; }
    bool unmapped(int32_t i) const { return ((lf[i] >> 16) & 0x4u) != 0u; }   // a mate placed at its partner's position

    // Read selection of one dataset: pileup() takes mapped reads with pos < stop and end > start (htslib drops unmapped
    // records); fetch() also returns unmapped reads, as intervals of length one at their position.
    void overlapping(int ds, int64_t start, int64_t stop, std::vector<int32_t>& out, bool fetch = false) const {
        out.clear();
        const int32_t* b = pos + (ds ? n_tumor : 0);
        const int32_t* e = pos + (ds ? n : n_tumor);
        const int32_t* lo = std::lower_bound(b, e, start - std::max(span, 1), [](int32_t p, int64_t v) { return (int64_t)p < v; });
        const int32_t* hi = std::lower_bound(b, e, stop, [](int32_t p, int64_t v) { return (int64_t)p < v; });
        for (const int32_t* p = lo; p < hi; ++p) {
            const int32_t i = (int32_t)(p - pos);
            if (unmapped(i)) { if (fetch && (int64_t)pos[i] + 1 > start) out.push_back(i); }
            else if ((int64_t)end[i] > start) out.push_back(i);
        }
    }

    int64_t* store(int32_t name, int m, int64_t value, int64_t ins_at, int32_t* entry = nullptr) {   // add_*_to_collection: an occupied slot keeps its read
        int32_t k = entry_of[name];
        if (k < 0) { k = (int32_t)entries.size(); entries.push_back({name, {kNone, kNone}, true, ins_at}); entry_of[name] = k; }
        if (entries[k].slot[m] == kNone) entries[k].slot[m] = value;
        if (entry) *entry = k;
        return entries[k].slot;
    }
    void pop(int32_t name) {
        const int32_t k = entry_of[name];
        if (k >= 0) { entries[k].alive = false; entry_of[name] = -1; }
    }
    void write_pair(int32_t name, int64_t s1, int64_t s2, std::vector<int32_t>& sink) {   // write_pair, SR.py:134-165
        if (written[name]) return;
        written[name] = 1;
        const int32_t row[5] = {dataset(slot_read(s1)), slot_read(s1), slot_version(s1), slot_read(s2), slot_version(s2)};
        sink.insert(sink.end(), row, row + 5);
    }

    // Pairs of one session in the order CompleteGermlineAnonymizer.anonymize yields them (AM.py:472-476, 489-512,
    // 521-532): a pair whose two mates are in the session leaves at the first normal pileup column right of both
    // mates, pairs in first-appearance order; everything else at the end, in registry order.
    struct Yield { int32_t name, m1, m2; };
    void session_yield_order(const std::vector<int32_t>& t_idx, const std::vector<int32_t>& n_idx, std::vector<Yield>& out) {
        out.clear();
        ++cur_stamp;
        struct Local { int32_t name, slot[2]; int32_t max_end; };
        std::vector<Local> loc;
        size_t a = 0, b = 0;
        while (a < t_idx.size() || b < n_idx.size()) {      // reads appear at their own start, tumor column before normal column
            int32_t i;
            if (b >= n_idx.size() || (a < t_idx.size() && pos[t_idx[a]] <= pos[n_idx[b]])) i = t_idx[a++]; else i = n_idx[b++];
            const int32_t nm = name_id[i];
            if (stamp[nm] != cur_stamp) { stamp[nm] = cur_stamp; local_of[nm] = (int32_t)loc.size(); loc.push_back({nm, {-1, -1}, -1}); }
            Local& L = loc[local_of[nm]];
            const int m = mate(i);
            if (L.slot[m] < 0) L.slot[m] = i;                // the first alignment of a (name, mate) is the read
            L.max_end = std::max(L.max_end, end[i]);
        }
        // normal pileup columns = positions covered by a normal read of the session
        std::vector<std::pair<int32_t, int32_t>> merged;
        for (const int32_t i : n_idx) {
            if (!merged.empty() && pos[i] <= merged.back().second) merged.back().second = std::max(merged.back().second, end[i]);
            else merged.push_back({pos[i], end[i]});
        }
        auto first_normal_column_after = [&](int32_t p, int32_t* col) -> bool {   // smallest covered position > p
            int64_t k = (int64_t)(std::upper_bound(merged.begin(), merged.end(), p + 1,
                                                   [](int32_t v, const std::pair<int32_t, int32_t>& m) { return v < m.first; }) - merged.begin()) - 1;
            if (k >= 0 && merged[k].second > p + 1) { *col = p + 1; return true; }
            ++k;
            if (k < (int64_t)merged.size()) { *col = merged[k].first; return true; }
            return false;
        };
        struct Early { int32_t col, order; };
        std::vector<Early> early;
        std::vector<int32_t> late;
        for (int32_t o = 0; o < (int32_t)loc.size(); ++o) {
            int32_t col;
            if (loc[o].slot[0] >= 0 && loc[o].slot[1] >= 0 && first_normal_column_after(loc[o].max_end, &col)) early.push_back({col, o});
            else late.push_back(o);
        }
        std::sort(early.begin(), early.end(), [](const Early& x, const Early& y) { return x.col != y.col ? x.col < y.col : x.order < y.order; });
        for (const Early& e : early) out.push_back({loc[e.order].name, loc[e.order].slot[0], loc[e.order].slot[1]});
        for (const int32_t o : late) out.push_back({loc[o].name, loc[o].slot[0], loc[o].slot[1]});
    }

    void run_session(int32_t first, int32_t last, int32_t window) {
        const int32_t s = (int32_t)plan->s_first.size();
        plan->s_first.push_back(first); plan->s_last.push_back(last); plan->s_window.push_back(window);
        std::vector<int32_t> t_idx, n_idx;
        overlapping(0, first, last, t_idx);
        overlapping(1, first, last, n_idx);
        std::vector<Yield> ys;
        session_yield_order(t_idx, n_idx, ys);
        for (const Yield& y : ys) {
            if (y.m1 >= 0 && y.m2 >= 0) { write_pair(y.name, slot_of(y.m1, s), slot_of(y.m2, s), plan->pairs); continue; }   // SR.py:310-312
            int64_t* slot = nullptr;
            const int64_t at = (int64_t)plan->pairs.size() / 5;
            // quirk Q12: a read that waits unpaired with the masking of an earlier session and is met again has its left-over
            // indels switched on again (anonymizer_methods.py:281-287) - whoever writes it applies them a second time
            auto meet = [&](int m, int32_t read) {
                slot = store(y.name, m, slot_of(read, s), at);
                if (slot[m] != slot_of(read, s) && slot_version(slot[m]) >= 0) slot[m] = slot_of(slot_read(slot[m]), slot_version(slot[m]) | kReapply);
            };
            if (y.m1 >= 0) meet(0, y.m1);                                                                                    // SR.py:320-333
            if (y.m2 >= 0) meet(1, y.m2);
            if (slot && slot[0] != kNone && slot[1] != kNone) {                                                              // SR.py:348-359
                const int64_t s0 = slot[0], s1 = slot[1];
                write_pair(y.name, s0, s1, plan->pairs);
                pop(y.name);
            }
        }
    }

    // Chains of reads in which every read overlaps (or touches, or ends with) the one before it
    // (collect_intersecting_reads, pileup_io.pyx:78-106 with compare, :44-59).
    struct Island { size_t begin, end_; int32_t first_pos, max_end; };   // [begin, end_) into the index list
    // idx: the fetched reads without the unmapped ones that iter_fetch_pair sets aside (only the first fetched read can
    // be unmapped here: it seeds an island as a read of length zero, pileup_io.pyx:72-73); max_end counts mapped reads
    // only and is 0 without any (get_righmost_pos, :109-121).
    void islands(const std::vector<int32_t>& idx, std::vector<Island>& out) const {
        out.clear();
        for (size_t k = 0; k < idx.size(); ++k) {
            const int32_t i = idx[k];
            if (!out.empty()) {
                const int32_t l = idx[k - 1];
                const int32_t l_end = unmapped(l) ? pos[l] : end[l];
                if ((pos[i] <= l_end && end[i] >= pos[l]) || end[i] == l_end) {
                    out.back().end_ = k + 1; out.back().max_end = std::max(out.back().max_end, end[i]);
                    continue;
                }
            }
            out.push_back({k, k + 1, pos[i], unmapped(i) ? 0 : end[i]});
        }
    }
    // the unmapped reads behind the first fetched read leave the list (pileup_io.pyx:93-96)
    void set_unmapped_aside(std::vector<int32_t>& idx, std::vector<int32_t>& aside) const {
        aside.clear();
        size_t w = 0;
        for (size_t k = 0; k < idx.size(); ++k) {
            if (k > 0 && unmapped(idx[k])) aside.push_back(idx[k]);
            else idx[w++] = idx[k];
        }
        idx.resize(w);
    }
    static int cmp_islands(const Island& a, const Island& b) {   // compare() of pileup_io.pyx:44-59
        const int32_t f1 = a.first_pos, l1 = a.max_end, f2 = b.first_pos, l2 = b.max_end;
        const bool overlap = f2 <= l1 && l2 >= f1;
        if (l1 < l2) return overlap ? -1 : -2;
        if (l2 < l1) return overlap ? 1 : 2;
        return f1 < f2 ? -1 : (f2 < f1 ? 1 : 0);
    }

    // pair_unmapped_or_non_pileup_pairs_and_write (SR.py:375-406) for one island that is yielded singly
    std::vector<int32_t> region_entries;                   // entries first stored by the current region: ins_at still relative
    void pass_through(const std::vector<int32_t>& idx, const Island* isl, std::vector<int32_t>& deferred) {
        if (!isl) return;
        pass_reads(idx, isl->begin, isl->end_, deferred);
    }
    void pass_reads(const std::vector<int32_t>& idx, size_t begin, size_t end_, std::vector<int32_t>& deferred) {
        for (size_t k = begin; k < end_; ++k) {
            const int32_t i = idx[k];
            const size_t n_before = entries.size();
            int32_t ent = -1;
            int64_t* slot = store(name_id[i], mate(i), slot_of(i, -1), (int64_t)deferred.size() / 5, &ent);
            if (entries.size() != n_before) region_entries.push_back(ent);
            if (slot[0] != kNone && slot[1] != kNone) write_pair(name_id[i], slot[0], slot[1], deferred);   // stays in the collection until the end (SR.py:737-741)
        }
    }

    // What iter_fetch_pair (pileup_io.pyx:124-298) yields for the fetched reads of one inter-window region.
    void region(int64_t start, int64_t stop) {
        std::vector<int32_t> t_idx, n_idx, deferred, t_um, n_um;
        overlapping(0, start, stop, t_idx, true);
        overlapping(1, start, stop, n_idx, true);
        if (t_idx.empty() && n_idx.empty()) return;
        set_unmapped_aside(t_idx, t_um);
        set_unmapped_aside(n_idx, n_um);
        region_entries.clear();
        std::vector<Island> ti, ni;
        islands(t_idx, ti);
        islands(n_idx, ni);
        size_t a = 0, b = 0;
        for (;;) {
            const bool more_t = a + 1 < ti.size(), more_n = b + 1 < ni.size();
            if (!more_t && !more_n) {                          // the last island of each dataset is always yielded singly
                pass_through(t_idx, a < ti.size() ? &ti[a] : nullptr, deferred);
                pass_through(n_idx, b < ni.size() ? &ni[b] : nullptr, deferred);
                break;
            }
            if (more_t && more_n) {
                const int c = cmp_islands(ti[a], ni[b]);
                if (c < -1) { pass_through(t_idx, &ti[a], deferred); ++a; }
                else if (c > 1) { pass_through(n_idx, &ni[b], deferred); ++b; }
                else {                                         // an island session has no variant to keep (SR.py:523-534)
                    run_session(std::min(ti[a].first_pos, ni[b].first_pos), std::max(ti[a].max_end, ni[b].max_end), -1);
                    ++a; ++b;
                }
            } else {
                if (more_t) { pass_through(t_idx, &ti[a], deferred); ++a; }
                if (more_n) { pass_through(n_idx, &ni[b], deferred); ++b; }
            }
        }
        // the unmapped reads that were set aside come last (pileup_io.pyx:298 -> SR.py:536-545), tumor then normal
        pass_reads(t_um, 0, t_um.size(), deferred);
        pass_reads(n_um, 0, n_um.size(), deferred);
        // the region's own records reach the files behind those of its island sessions (stream buffering)
        const int64_t base = (int64_t)plan->pairs.size() / 5;
        for (const int32_t k : region_entries) entries[k].ins_at += base;
        plan->pairs.insert(plan->pairs.end(), deferred.begin(), deferred.end());
    }
};

}  // namespace

extern "C" {

int ga_plan_sample(int64_t n_reads, int64_t n_tumor, const int32_t* pos, const int32_t* ref_end, const uint32_t* len_flag,
                   const int64_t* name_off, const uint8_t* names, int32_t n_windows, const int32_t* win_first,
                   const int32_t* win_last, int64_t contig_len, ga_plan** out) {
    if (!out || n_reads < 0 || n_tumor < 0 || n_tumor > n_reads || n_windows < 0 ||
        (n_reads && (!pos || !ref_end || !len_flag || !name_off || !names)) || (n_windows && (!win_first || !win_last)))
        return ga_io_set_error(GA_IO_ERR_ARGUMENT, "ga_plan_sample: bad argument");
    *out = nullptr;
    for (int64_t i = 1; i < n_reads; ++i)
        if (i != n_tumor && pos[i] < pos[i - 1]) return ga_io_set_error(GA_IO_ERR_ARGUMENT, "reads of a dataset must be in coordinate order");
    Planner P;
    P.n = n_reads; P.n_tumor = n_tumor; P.pos = pos; P.end = ref_end; P.lf = len_flag;
    P.name_id.resize((size_t)n_reads);
    {
        // dense name ids through a flat open-addressing table keyed by a 64-bit hash of the name (one probe sequence per
        // read, no node allocation); equal hashes are confirmed by comparing the bytes
        size_t cap = 16;
        while (cap < (size_t)n_reads * 2) cap <<= 1;
        struct Cell { uint64_t hash; int32_t id, read; };
        std::vector<Cell> table(cap, Cell{0, -1, -1});
        const char* base = reinterpret_cast<const char*>(names);
        auto hash_of = [](const char* p, size_t len) {
            uint64_t h = 0x9E3779B97F4A7C15ull ^ (len * 0xFF51AFD7ED558CCDull);
            while (len >= 8) { uint64_t w; std::memcpy(&w, p, 8); h = (h ^ w) * 0x9FB21C651E98DF25ull; h ^= h >> 29; p += 8; len -= 8; }
            uint64_t w = 0;
            std::memcpy(&w, p, len);
            h = (h ^ w) * 0x9FB21C651E98DF25ull;
            return h ^ (h >> 32);
        };
        std::vector<uint64_t> hashes((size_t)n_reads);
        for (int64_t i = 0; i < n_reads; ++i) hashes[(size_t)i] = hash_of(base + name_off[i], (size_t)(name_off[i + 1] - name_off[i]));
        constexpr int64_t kAhead = 24;                          // the table is far larger than the caches: cells are fetched ahead of their probe
        for (int64_t i = 0; i < n_reads; ++i) {
            if (i + kAhead < n_reads) __builtin_prefetch(&table[(size_t)hashes[(size_t)(i + kAhead)] & (cap - 1)]);
            const char* nm = base + name_off[i];
            const size_t len = (size_t)(name_off[i + 1] - name_off[i]);
            const uint64_t h = hashes[(size_t)i];
            size_t k = (size_t)h & (cap - 1);
            for (;; k = (k + 1) & (cap - 1)) {
                Cell& c = table[k];
                if (c.id < 0) { c = Cell{h, P.n_names++, (int32_t)i}; break; }
                if (c.hash == h && (size_t)(name_off[c.read + 1] - name_off[c.read]) == len && std::memcmp(base + name_off[c.read], nm, len) == 0) break;
            }
            P.name_id[(size_t)i] = table[k].id;
            P.span = std::max(P.span, ref_end[i] - pos[i]);
        }
    }
    P.entry_of.assign((size_t)P.n_names, -1);
    P.written.assign((size_t)P.n_names, 0);
    P.stamp.assign((size_t)P.n_names, 0);
    P.local_of.assign((size_t)P.n_names, 0);
    ga_plan* plan = new ga_plan();
    P.plan = plan;
    // get_genome_sections (SR.py:245-276): window k, the regions between windows, stably sorted by (first, last)
    struct Section { int64_t first, last; int32_t window; };
    std::vector<Section> secs;
    if (n_windows == 0) secs.push_back({0, 0, -1});
    else {
        int64_t nxt = 1;
        for (int32_t k = 0; k < n_windows; ++k) {
            secs.push_back({nxt, (int64_t)win_first[k] - 1, -1});
            secs.push_back({win_first[k], win_last[k], k});
            nxt = (int64_t)win_last[k] + 1;
        }
        secs.push_back({nxt, contig_len - 1, -1});
        std::stable_sort(secs.begin(), secs.end(), [](const Section& x, const Section& y) { return x.first != y.first ? x.first < y.first : x.last < y.last; });
    }
    for (const Section& sc : secs) {
        if (sc.window >= 0) { P.run_session((int32_t)sc.first, (int32_t)sc.last, sc.window); continue; }
        int64_t start, stop;
        if (sc.first + sc.last == 0) { start = 0; stop = contig_len; }
        else {
            if (sc.first < 0 || sc.first > sc.last) {          // pysam rejects these coordinates
                delete plan;
                return ga_io_set_error(GA_IO_ERR_ARGUMENT, ("inter-window region (" + std::to_string(sc.first) + ", " + std::to_string(sc.last) +
                                                            ") is not fetchable: variants closer than a window").c_str());
            }
            start = sc.first; stop = sc.last;
        }
        P.region(start, stop);
    }
    // pair_unmapped_mates (SR.py:561-600, called at :725-732 when the collection is not empty): every window is fetched
    // again (start = first - 1), tumor then normal, and an unmapped read whose name waits in the collection joins it
    bool waiting = false;
    for (const Planner::Entry& e : P.entries) waiting |= e.alive;
    if (waiting) {
        std::vector<int32_t> idx;
        for (int32_t k = 0; k < n_windows; ++k)
            for (int ds = 0; ds < 2; ++ds) {
                P.overlapping(ds, std::max<int64_t>((int64_t)win_first[k] - 1, 0), win_last[k], idx, true);
                for (const int32_t i : idx) {
                    if (!P.unmapped(i) || P.entry_of[P.name_id[i]] < 0) continue;
                    int64_t* slot = P.store(P.name_id[i], P.mate(i), slot_of(i, -1), (int64_t)plan->pairs.size() / 5);
                    if (slot[0] != kNone && slot[1] != kNone) P.write_pair(P.name_id[i], slot[0], slot[1], plan->pairs);
                }
            }
    }
    for (const Planner::Entry& e : P.entries) {                // write_single_end_reads (SR.py:603-622), insertion order
        if (!e.alive || P.written[e.name]) continue;
        const int64_t v = e.slot[0] != kNone ? e.slot[0] : e.slot[1];
        const int32_t row[4] = {P.dataset(slot_read(v)), slot_read(v), slot_version(v), (int32_t)e.ins_at};
        plan->singles.insert(plan->singles.end(), row, row + 4);
    }
    *out = plan;
    return GA_IO_OK;
}

void ga_plan_free(ga_plan* p) { delete p; }
int64_t ga_plan_n_sessions(const ga_plan* p) { return p ? (int64_t)p->s_first.size() : 0; }
int64_t ga_plan_n_pairs(const ga_plan* p) { return p ? (int64_t)p->pairs.size() / 5 : 0; }
int64_t ga_plan_n_singles(const ga_plan* p) { return p ? (int64_t)p->singles.size() / 4 : 0; }
void ga_plan_sessions(const ga_plan* p, int32_t* first, int32_t* last, int32_t* window) {
    if (!p) return;
    const size_t n = p->s_first.size();
    if (n) { std::memcpy(first, p->s_first.data(), 4 * n); std::memcpy(last, p->s_last.data(), 4 * n); std::memcpy(window, p->s_window.data(), 4 * n); }
}
void ga_plan_pairs(const ga_plan* p, int32_t* rows) { if (p && !p->pairs.empty()) std::memcpy(rows, p->pairs.data(), 4 * p->pairs.size()); }
void ga_plan_singles(const ga_plan* p, int32_t* rows) { if (p && !p->singles.empty()) std::memcpy(rows, p->singles.data(), 4 * p->singles.size()); }

}  // extern "C"
