// Ingest (the first row after the hot path): text files of
// offsetbasedgraph intervals -> the reads structure of arrays, on the host.
//
// An `.intervalcollection` text file holds one JSON object per line, as written
// by Interval.to_file_line / IntervalCollection.to_file(text_file=True) -- the
// format the reference reads its alignments from after vg JSON conversion
// (callpeaks_interface.py:64-66) and writes its peaks in (reporter.py:25-33):
//
//   {"start": 3, "end": 17, "region_paths": [12, 13, 15], "direction": 1}
//
// Keys come in any order, other keys (peaks carry scores, ids, ...) are skipped.
// A read's start / end node is the first / last region path.  The Python json
// module parses ~10^5 such lines per second; this scanner does a few 10^7 per
// second and core, and splits the buffer over host threads at line ends.
// Host code only: no kernel, all pointers are host pointers.
#include <stdlib.h>

#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace gpc {
namespace {

struct Cursor {
    const char* p;
    const char* end;
    bool eof() const { return p >= end; }
    void skip_ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; }
};

bool parse_int(Cursor& c, long long& out) {
    c.skip_ws();
    bool neg = false;
    if (!c.eof() && (*c.p == '-' || *c.p == '+')) { neg = *c.p == '-'; ++c.p; }
    if (c.eof() || *c.p < '0' || *c.p > '9') return false;
    long long v = 0;
    while (!c.eof() && *c.p >= '0' && *c.p <= '9') { v = v * 10 + (*c.p - '0'); ++c.p; }
    // "12.0" is accepted (json.dumps of a float that is an integer), a fraction is not
    if (!c.eof() && *c.p == '.') {
        ++c.p;
        while (!c.eof() && *c.p >= '0' && *c.p <= '9') { if (*c.p != '0') return false; ++c.p; }
    }
    out = neg ? -v : v;
    return true;
}

bool skip_string(Cursor& c) {               // c.p at the opening quote
    ++c.p;
    while (!c.eof() && *c.p != '"') { if (*c.p == '\\') ++c.p; ++c.p; }
    if (c.eof()) return false;
    ++c.p;
    return true;
}

bool skip_value(Cursor& c) {
    c.skip_ws();
    if (c.eof()) return false;
    if (*c.p == '"') return skip_string(c);
    if (*c.p == '[' || *c.p == '{') {
        int depth = 0;
        while (!c.eof()) {
            if (*c.p == '"') { if (!skip_string(c)) return false; continue; }
            if (*c.p == '[' || *c.p == '{') ++depth;
            if (*c.p == ']' || *c.p == '}') { --depth; if (depth == 0) { ++c.p; return true; } }
            ++c.p;
        }
        return false;
    }
    while (!c.eof() && *c.p != ',' && *c.p != '}') ++c.p;   // number / true / false / null
    return true;
}

// One line.  `path` may be NULL (counting pass); returns the number of region
// paths, or -1 on malformed input.
long long parse_line(const char* b, const char* e, long long& start, long long& stop,
                     int32_t* path, long long& first, long long& last) {
    Cursor c{b, e};
    c.skip_ws();
    if (c.eof() || *c.p != '{') return -1;
    ++c.p;
    bool have_start = false, have_end = false, have_path = false;
    long long n_path = 0;
    while (true) {
        c.skip_ws();
        if (c.eof()) return -1;
        if (*c.p == '}') break;
        if (*c.p == ',') { ++c.p; continue; }
        if (*c.p != '"') return -1;
        const char* k0 = c.p + 1;
        if (!skip_string(c)) return -1;
        const size_t klen = (size_t)(c.p - 1 - k0);
        c.skip_ws();
        if (c.eof() || *c.p != ':') return -1;
        ++c.p;
        if (klen == 5 && !memcmp(k0, "start", 5)) {
            if (!parse_int(c, start)) return -1;
            have_start = true;
        } else if (klen == 3 && !memcmp(k0, "end", 3)) {
            if (!parse_int(c, stop)) return -1;
            have_end = true;
        } else if (klen == 12 && !memcmp(k0, "region_paths", 12)) {
            c.skip_ws();
            if (c.eof() || *c.p != '[') return -1;
            ++c.p;
            while (true) {
                c.skip_ws();
                if (c.eof()) return -1;
                if (*c.p == ']') { ++c.p; break; }
                if (*c.p == ',') { ++c.p; continue; }
                long long v;
                if (!parse_int(c, v)) return -1;
                if (v > 2147483647ll || v < -2147483647ll) return -1;
                if (n_path == 0) first = v;
                last = v;
                if (path) path[n_path] = (int32_t)v;
                ++n_path;
            }
            have_path = true;
        } else if (!skip_value(c)) {
            return -1;
        }
    }
    if (!have_start || !have_end || !have_path || n_path == 0) return -1;
    if (start < 0 || stop < 0 || start > 2147483647ll || stop > 2147483647ll) return -1;
    return n_path;
}

bool blank(const char* b, const char* e) {
    for (; b < e; ++b) if (*b != ' ' && *b != '\t' && *b != '\r') return false;
    return true;
}

struct Chunk {
    const char* b;
    const char* e;
    unsigned long long reads = 0, nodes = 0;
    long long bad_line = -1;           // index of the first malformed line inside the chunk
};

// Applies fn(line begin, line end) to every non-blank line of the chunk.
template <typename Fn>
void for_lines(const Chunk& ch, Fn&& fn) {
    const char* p = ch.b;
    while (p < ch.e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(ch.e - p));
        const char* le = nl ? nl : ch.e;
        if (!blank(p, le)) fn(p, le);
        p = nl ? nl + 1 : ch.e;
    }
}

}  // namespace
}  // namespace gpc

using namespace gpc;

extern "C" int gpc_parse_intervals_host(const char* text, uint64_t n_bytes, int32_t n_threads,
                                        int32_t* start_node, int32_t* start_off,
                                        int32_t* end_node, int32_t* end_off, int64_t* path_ptr,
                                        int32_t* path, uint64_t* n_reads,
                                        uint64_t* n_path_nodes) {
    const bool fill = start_node != nullptr;
    const uint64_t want_reads = *n_reads, want_nodes = *n_path_nodes;
    *n_reads = 0;
    *n_path_nodes = 0;
    if (fill && (!start_off || !end_node || !end_off || !path_ptr || !path)) {
        set_error("gpc_parse_intervals_host: all output arrays or none");
        return GPC_ERR_ARG;
    }
    unsigned threads = n_threads > 0 ? (unsigned)n_threads : std::thread::hardware_concurrency();
    if (threads == 0) threads = 1;
    if (threads > 64) threads = 64;
    if (n_bytes < (1u << 20)) threads = 1;
    // chunks end at line ends
    std::vector<Chunk> chunks;
    const char* end = text + n_bytes;
    const char* p = text;
    for (unsigned t = 0; t < threads && p < end; ++t) {
        const char* stop = (t + 1 == threads) ? end : text + (n_bytes / threads) * (t + 1);
        if (stop < p) stop = p;
        if (stop < end) {
            const char* nl = (const char*)memchr(stop, '\n', (size_t)(end - stop));
            stop = nl ? nl + 1 : end;
        }
        Chunk ch;
        ch.b = p;
        ch.e = stop;
        chunks.push_back(ch);
        p = stop;
    }
    // pass 1: reads and path nodes per chunk
    {
        std::vector<std::thread> pool;
        for (auto& ch : chunks)
            pool.emplace_back([&ch]() {
                long long line = 0;
                for_lines(ch, [&](const char* b, const char* e) {
                    long long s, t, f, l;
                    long long k = parse_line(b, e, s, t, nullptr, f, l);
                    if (k < 0) { if (ch.bad_line < 0) ch.bad_line = line; }
                    else { ch.reads += 1; ch.nodes += (unsigned long long)k; }
                    ++line;
                });
            });
        for (auto& th : pool) th.join();
    }
    unsigned long long reads = 0, nodes = 0, lines_before = 0;
    for (auto& ch : chunks) {
        if (ch.bad_line >= 0) {
            set_error("interval file: malformed line (non-blank line %llu of its chunk, ~line %llu)",
                      (unsigned long long)ch.bad_line + 1, lines_before + ch.bad_line + 1);
            return GPC_ERR_ARG;
        }
        lines_before += ch.reads;
        reads += ch.reads;
        nodes += ch.nodes;
    }
    *n_reads = reads;
    *n_path_nodes = nodes;
    if (!fill) return GPC_OK;
    if (want_reads < reads || want_nodes < nodes) return GPC_ERR_CAPACITY;
    // pass 2: every chunk writes its own range
    {
        std::vector<std::thread> pool;
        unsigned long long r0 = 0, p0 = 0;
        for (auto& ch : chunks) {
            pool.emplace_back([&ch, r0, p0, start_node, start_off, end_node, end_off, path_ptr,
                               path]() {
                unsigned long long r = r0, q = p0;
                for_lines(ch, [&](const char* b, const char* e) {
                    long long s = 0, t = 0, f = 0, l = 0;
                    long long k = parse_line(b, e, s, t, path + q, f, l);
                    start_node[r] = (int32_t)f;
                    start_off[r] = (int32_t)s;
                    end_node[r] = (int32_t)l;
                    end_off[r] = (int32_t)t;
                    path_ptr[r] = (int64_t)q;
                    q += (unsigned long long)k;
                    ++r;
                });
            });
            r0 += ch.reads;
            p0 += ch.nodes;
        }
        for (auto& th : pool) th.join();
    }
    path_ptr[reads] = (int64_t)nodes;
    return GPC_OK;
}
