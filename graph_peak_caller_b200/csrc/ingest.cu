// Ingest (the first row after the hot path): text files of
// offsetbasedgraph intervals and of vg alignments -> the reads structure of arrays,
// on the host.
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
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <mutex>
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
    while (true) {
        const char* q = c.eof() ? nullptr : (const char*)memchr(c.p, '"', (size_t)(c.end - c.p));
        if (!q) { c.p = c.end; return false; }
        // the quote closes the string unless an odd number of backslashes precedes it
        const char* b = q;
        while (b > c.p && b[-1] == '\\') --b;
        c.p = q + 1;
        if (((q - b) & 1) == 0) return true;
    }
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

// One line.  Region paths are appended to `path`; returns their number, or -1 on
// malformed input (what was appended is then dropped by the caller).
long long parse_line(const char* b, const char* e, long long& start, long long& stop,
                     std::vector<int32_t>& path, long long& first, long long& last) {
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
                path.push_back((int32_t)v);
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


// ---- vg alignments (`vg view -aj`): one JSON Alignment per line -----------------------
// What pyvg.conversion.vg_json_file_to_intervals makes of a line (pyvg is not part of
// the reference tree; restated from its published classes, anchored on the reference's
// tests/examples.py:5-31 which builds Position(node_id, offset, is_reverse),
// Edit(to_length, from_length, sequence), Mapping(position, edits) and expects
// Interval(10, 15, [0]) from one mapping at offset 10 with edits of from_length 4 + 1):
//   region_paths = [mapping.position.node_id, negated when position.is_reverse]
//   start        = offset of the first mapping's position (absent = 0)
//   end          = offset of the last mapping's position + sum of its edits' from_length
// protobuf's JSON form writes 64-bit integers as strings ("node_id": "8235641"), so
// quoted integers are accepted (the reference's own splitter allows for them too,
// preprocess_interface.py:109).  Alignments without a path or with no mappings
// (unaligned reads) produce no interval.

bool parse_json_int(Cursor& c, long long& out) {
    c.skip_ws();
    const bool quoted = !c.eof() && *c.p == '"';
    if (quoted) ++c.p;
    if (!parse_int(c, out)) return false;
    if (quoted) {
        if (c.eof() || *c.p != '"') return false;
        ++c.p;
    }
    return true;
}

bool parse_json_bool(Cursor& c, bool& out) {
    c.skip_ws();
    if (c.end - c.p >= 4 && !memcmp(c.p, "true", 4)) { out = true; c.p += 4; return true; }
    if (c.end - c.p >= 5 && !memcmp(c.p, "false", 5)) { out = false; c.p += 5; return true; }
    return false;
}

// Iterates the members of a JSON object (c.p at '{'): on_key(key, length) must consume the
// value and return false on malformed input.
template <typename Fn>
bool for_members(Cursor& c, Fn&& on_key) {
    c.skip_ws();
    if (c.eof() || *c.p != '{') return false;
    ++c.p;
    while (true) {
        c.skip_ws();
        if (c.eof()) return false;
        if (*c.p == '}') { ++c.p; return true; }
        if (*c.p == ',') { ++c.p; continue; }
        if (*c.p != '"') return false;
        const char* k0 = c.p + 1;
        if (!skip_string(c)) return false;
        const size_t klen = (size_t)(c.p - 1 - k0);
        c.skip_ws();
        if (c.eof() || *c.p != ':') return false;
        ++c.p;
        if (!on_key(k0, klen)) return false;
    }
}

// Iterates the elements of a JSON array (c.p at '['): on_item() consumes one element.
template <typename Fn>
bool for_items(Cursor& c, Fn&& on_item) {
    c.skip_ws();
    if (c.eof() || *c.p != '[') return false;
    ++c.p;
    while (true) {
        c.skip_ws();
        if (c.eof()) return false;
        if (*c.p == ']') { ++c.p; return true; }
        if (*c.p == ',') { ++c.p; continue; }
        if (!on_item()) return false;
    }
}

template <size_t N>
inline bool key_is(const char* k, size_t n, const char (&name)[N]) {
    return n == N - 1 && !memcmp(k, name, N - 1);
}

struct VgMapping {
    long long node = 0, offset = 0, from_length = 0;
    bool reverse = false, have_node = false;
};

bool parse_vg_mapping(Cursor& c, VgMapping& m) {
    return for_members(c, [&](const char* k, size_t n) {
        if (key_is(k, n, "position"))
            return for_members(c, [&](const char* k2, size_t n2) {
                if (key_is(k2, n2, "node_id")) { m.have_node = true; return parse_json_int(c, m.node); }
                if (key_is(k2, n2, "offset")) return parse_json_int(c, m.offset);
                if (key_is(k2, n2, "is_reverse")) return parse_json_bool(c, m.reverse);
                return skip_value(c);
            });
        if (key_is(k, n, "edit"))
            return for_items(c, [&]() {
                return for_members(c, [&](const char* k2, size_t n2) {
                    if (key_is(k2, n2, "from_length")) {
                        long long v;
                        if (!parse_json_int(c, v)) return false;
                        m.from_length += v;
                        return true;
                    }
                    return skip_value(c);
                });
            });
        return skip_value(c);
    });
}

// One alignment line.  Returns the number of mappings (0: no interval), -1 on malformed
// input, -2 when a node lies outside [min_node, max_node] (checked when max_node > 0).
long long parse_vg_line(const char* b, const char* e, long long min_node, long long max_node,
                        long long& start, long long& stop, std::vector<int32_t>& path,
                        long long& first, long long& last) {
    Cursor c{b, e};
    long long n_path = 0;
    bool outside = false;
    VgMapping m;
    const bool ok = for_members(c, [&](const char* k, size_t n) {
        if (!key_is(k, n, "path")) return skip_value(c);
        return for_members(c, [&](const char* k2, size_t n2) {
            if (!key_is(k2, n2, "mapping")) return skip_value(c);
            return for_items(c, [&]() {
                m = VgMapping();
                if (!parse_vg_mapping(c, m) || !m.have_node) return false;
                if (m.node < 0 || (m.node == 0 && m.reverse) || m.node > 2147483647ll ||
                    m.offset < 0 || m.offset > 2147483647ll || m.from_length < 0)
                    return false;
                if (max_node > 0 && (m.node < min_node || m.node > max_node)) outside = true;
                const long long id = m.reverse ? -m.node : m.node;
                if (n_path == 0) { first = id; start = m.offset; }
                last = id;
                path.push_back((int32_t)id);
                ++n_path;
                return true;
            });
        });
    });
    if (!ok) return -1;
    if (n_path == 0) return 0;
    stop = m.offset + m.from_length;
    if (stop > 2147483647ll) return -1;
    return outside ? -2 : n_path;
}

// The driver both text formats share.  The buffer is cut into chunks of whole lines,
// one host thread per chunk parses its lines ONCE into chunk-local arrays.  The
// counting call (no output arrays) keeps that parse; the filling call for the same
// buffer copies it out instead of parsing again (a filling call without a matching
// counting call simply parses).  line_fn(begin, end, start, stop, path vector, first,
// last) returns the number of path nodes it appended, 0 for a line that holds no read,
// -1 for a malformed line, -2 for a read outside the graph.
struct ParsedChunk {
    std::vector<int32_t> start_node, start_off, end_node, end_off, n_path, path;
    unsigned long long empty = 0, lines = 0;
    long long bad_line = -1;           // index of the first bad line inside the chunk
    int bad_kind = 0;                  // -1 malformed, -2 outside the graph
};

struct ParseCache {
    const char* text = nullptr;
    uint64_t n_bytes = 0;
    int kind = 0;
    long long lo = 0, hi = 0;
    std::vector<ParsedChunk> chunks;
};
std::mutex g_parse_mutex;
ParseCache g_parse_cache;              // at most one pending parse

// New threads of a short job tend to stay on the CPU that created them until the
// scheduler's periodic balancing moves them (measured in a VM: eight 15 ms chunks ran
// one after the other on one core).  Each worker is therefore placed on its own CPU
// of the set this process may use; a placement that fails is simply not made.
void spread_over_cpus(std::vector<std::thread>& pool) {
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return;
    std::vector<int> cpus;
    for (int c = 0; c < CPU_SETSIZE; ++c)
        if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
    if (cpus.size() < 2 || pool.size() < 2) return;
    // processes of one node (one per GPU) start at different CPUs; after the move the
    // thread may again run anywhere this process may
    const size_t first = (size_t)getpid() * 7;
    for (size_t i = 0; i < pool.size(); ++i) {
        cpu_set_t one;
        CPU_ZERO(&one);
        CPU_SET(cpus[(first + i) % cpus.size()], &one);
        if (pthread_setaffinity_np(pool[i].native_handle(), sizeof(one), &one) == 0)
            pthread_setaffinity_np(pool[i].native_handle(), sizeof(allowed), &allowed);
    }
}

template <typename LineFn>
void parse_chunks(const char* text, uint64_t n_bytes, int32_t n_threads, LineFn& line_fn,
                  std::vector<ParsedChunk>& out) {
    unsigned threads = n_threads > 0 ? (unsigned)n_threads : 0;
    if (threads == 0) {                      // GPC_INGEST_THREADS, else every CPU this process may use
        const char* env = getenv("GPC_INGEST_THREADS");
        if (env && atoi(env) > 0) threads = (unsigned)atoi(env);
    }
    if (threads == 0) {
        cpu_set_t allowed;
        CPU_ZERO(&allowed);
        if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) threads = (unsigned)CPU_COUNT(&allowed);
    }
    if (threads == 0) threads = std::thread::hardware_concurrency();
    if (threads == 0) threads = 1;
    if (threads > 64) threads = 64;
    if (n_bytes < (1u << 20)) threads = 1;
    std::vector<Chunk> chunks;         // chunks end at line ends
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
    out.clear();
    out.resize(chunks.size());
    std::vector<std::thread> pool;
    for (size_t i = 0; i < chunks.size(); ++i)
        pool.emplace_back([&chunks, &out, &line_fn, i]() {
            ParsedChunk pc;            // filled locally, moved out at the end
            const Chunk& ch = chunks[i];
            const size_t guess = (size_t)(ch.e - ch.b) / 256 + 16;
            pc.start_node.reserve(guess);
            pc.start_off.reserve(guess);
            pc.end_node.reserve(guess);
            pc.end_off.reserve(guess);
            pc.n_path.reserve(guess);
            pc.path.reserve(3 * guess);
            long long line = 0;
            for_lines(ch, [&](const char* b, const char* e) {
                long long s = 0, t = 0, f = 0, l = 0;
                const size_t before = pc.path.size();
                const long long k = line_fn(b, e, s, t, pc.path, f, l);
                if (k <= 0) {
                    pc.path.resize(before);
                    if (k == 0) pc.empty += 1;
                    else if (pc.bad_line < 0) { pc.bad_line = line; pc.bad_kind = (int)k; }
                } else {
                    pc.start_node.push_back((int32_t)f);
                    pc.start_off.push_back((int32_t)s);
                    pc.end_node.push_back((int32_t)l);
                    pc.end_off.push_back((int32_t)t);
                    pc.n_path.push_back((int32_t)k);
                }
                ++line;
            });
            pc.lines = (unsigned long long)line;
            out[i] = std::move(pc);
        });
    spread_over_cpus(pool);
    for (auto& th : pool) th.join();
}

template <typename LineFn>
int parse_reads_text(const char* what, int kind, long long lo, long long hi, const char* text,
                     uint64_t n_bytes, int32_t n_threads, int32_t* start_node,
                     int32_t* start_off, int32_t* end_node, int32_t* end_off, int64_t* path_ptr,
                     int32_t* path, uint64_t* n_reads, uint64_t* n_path_nodes, uint64_t* n_empty,
                     LineFn line_fn) {
    const bool fill = start_node != nullptr;
    const uint64_t want_reads = *n_reads, want_nodes = *n_path_nodes;
    *n_reads = 0;
    *n_path_nodes = 0;
    if (n_empty) *n_empty = 0;
    if (fill && (!start_off || !end_node || !end_off || !path_ptr || !path)) {
        set_error("%s: all output arrays or none", what);
        return GPC_ERR_ARG;
    }
    std::vector<ParsedChunk> chunks;
    bool have = false;
    if (fill) {
        std::lock_guard<std::mutex> lock(g_parse_mutex);
        ParseCache& c = g_parse_cache;
        if (c.text == text && c.n_bytes == n_bytes && c.kind == kind && c.lo == lo && c.hi == hi &&
            text != nullptr) {
            chunks.swap(c.chunks);
            have = true;
        }
        c = ParseCache();
    }
    if (!have) parse_chunks(text, n_bytes, n_threads, line_fn, chunks);
    unsigned long long reads = 0, nodes = 0, empty = 0, lines_before = 0;
    for (auto& ch : chunks) {
        if (ch.bad_line >= 0) {
            if (ch.bad_kind == -2) {
                set_error("%s: interval not in graph (non-blank line %llu)", what,
                          lines_before + ch.bad_line + 1);
                return GPC_ERR_INVALID_INTERVAL;
            }
            set_error("%s: malformed line (non-blank line %llu)", what,
                      lines_before + ch.bad_line + 1);
            return GPC_ERR_ARG;
        }
        lines_before += ch.lines;
        reads += ch.start_node.size();
        nodes += ch.path.size();
        empty += ch.empty;
    }
    *n_reads = reads;
    *n_path_nodes = nodes;
    if (n_empty) *n_empty = empty;
    if (!fill) {
        std::lock_guard<std::mutex> lock(g_parse_mutex);
        ParseCache& c = g_parse_cache;
        c.text = text;
        c.n_bytes = n_bytes;
        c.kind = kind;
        c.lo = lo;
        c.hi = hi;
        c.chunks.swap(chunks);
        return GPC_OK;
    }
    if (want_reads < reads || want_nodes < nodes) return GPC_ERR_CAPACITY;
    // every chunk copies into its own range
    {
        std::vector<std::thread> pool;
        unsigned long long r0 = 0, p0 = 0;
        for (auto& ch : chunks) {
            pool.emplace_back([&ch, r0, p0, start_node, start_off, end_node, end_off, path_ptr,
                               path]() {
                const size_t r = ch.start_node.size();
                if (r) {
                    memcpy(start_node + r0, ch.start_node.data(), r * sizeof(int32_t));
                    memcpy(start_off + r0, ch.start_off.data(), r * sizeof(int32_t));
                    memcpy(end_node + r0, ch.end_node.data(), r * sizeof(int32_t));
                    memcpy(end_off + r0, ch.end_off.data(), r * sizeof(int32_t));
                }
                if (!ch.path.empty())
                    memcpy(path + p0, ch.path.data(), ch.path.size() * sizeof(int32_t));
                unsigned long long q = p0;
                for (size_t i = 0; i < r; ++i) {
                    path_ptr[r0 + i] = (int64_t)q;
                    q += (unsigned long long)ch.n_path[i];
                }
            });
            r0 += ch.start_node.size();
            p0 += ch.path.size();
        }
        spread_over_cpus(pool);
        for (auto& th : pool) th.join();
    }
    path_ptr[reads] = (int64_t)nodes;
    return GPC_OK;
}

// ---- vg graphs (`vg view -Vj`): {"node": [{"id", "sequence"}], "edge": [{"from", "to",
// "from_start", "to_end"}], "path": [...]} -- one object per line, any number of lines.
// What pyvg.conversion.json_file_to_obg_numpy_graph reads of it (the reference's
// create_ob_graph, preprocess_interface.py:51-57): a node's length is the length of its
// sequence; an edge leaves -from when from_start is set and enters -to when to_end is
// set.  Paths, names and everything else are skipped.

// Length of a JSON string's content in characters of the decoded text (c.p at the
// opening quote).  Sequences are plain ASCII; an escape counts as one character.
bool string_length(Cursor& c, long long& len) {
    ++c.p;
    const char* p = c.p;
    const char* q = (const char*)memchr(p, '"', (size_t)(c.end - p));
    if (!q) return false;
    if (!memchr(p, '\\', (size_t)(q - p))) {           // the usual case: no escapes
        len = q - p;
        c.p = q + 1;
        return true;
    }
    len = 0;
    while (!c.eof() && *c.p != '"') {
        if (*c.p == '\\') {
            ++c.p;
            if (!c.eof() && *c.p == 'u') c.p += 4;
        }
        ++c.p;
        ++len;
    }
    if (c.eof()) return false;
    ++c.p;
    return true;
}

struct VgGraphSink {
    int32_t* node_id;
    int32_t* node_len;
    int32_t* edge_from;
    int32_t* edge_to;
    unsigned long long node_room = 0, edge_room = 0;   // entries the arrays can take
    unsigned long long nodes = 0, edges = 0;
};

bool parse_vg_graph_object(Cursor& c, VgGraphSink& out) {
    return for_members(c, [&](const char* k, size_t n) {
        if (key_is(k, n, "node"))
            return for_items(c, [&]() {
                long long id = 0, len = 0;
                bool have_id = false;
                if (!for_members(c, [&](const char* k2, size_t n2) {
                        if (key_is(k2, n2, "id")) { have_id = true; return parse_json_int(c, id); }
                        if (key_is(k2, n2, "sequence")) {
                            c.skip_ws();
                            if (c.eof() || *c.p != '"') return false;
                            return string_length(c, len);
                        }
                        return skip_value(c);
                    }))
                    return false;
                if (!have_id || id <= 0 || id > 2147483647ll || len > 2147483647ll) return false;
                if (out.node_id && out.nodes < out.node_room) {
                    out.node_id[out.nodes] = (int32_t)id;
                    out.node_len[out.nodes] = (int32_t)len;
                }
                ++out.nodes;
                return true;
            });
        if (key_is(k, n, "edge"))
            return for_items(c, [&]() {
                long long from = 0, to = 0;
                bool from_start = false, to_end = false, have_from = false, have_to = false;
                if (!for_members(c, [&](const char* k2, size_t n2) {
                        if (key_is(k2, n2, "from")) { have_from = true; return parse_json_int(c, from); }
                        if (key_is(k2, n2, "to")) { have_to = true; return parse_json_int(c, to); }
                        if (key_is(k2, n2, "from_start")) return parse_json_bool(c, from_start);
                        if (key_is(k2, n2, "to_end")) return parse_json_bool(c, to_end);
                        return skip_value(c);
                    }))
                    return false;
                if (!have_from || !have_to || from <= 0 || to <= 0 || from > 2147483647ll ||
                    to > 2147483647ll)
                    return false;
                if (out.edge_from && out.edges < out.edge_room) {
                    out.edge_from[out.edges] = (int32_t)(from_start ? -from : from);
                    out.edge_to[out.edges] = (int32_t)(to_end ? -to : to);
                }
                ++out.edges;
                return true;
            });
        return skip_value(c);
    });
}

}  // namespace
}  // namespace gpc

using namespace gpc;

extern "C" int gpc_parse_intervals_host(const char* text, uint64_t n_bytes, int32_t n_threads,
                                        int32_t* start_node, int32_t* start_off,
                                        int32_t* end_node, int32_t* end_off, int64_t* path_ptr,
                                        int32_t* path, uint64_t* n_reads,
                                        uint64_t* n_path_nodes) {
    return parse_reads_text("interval file", 1, 0, 0, text, n_bytes, n_threads, start_node,
                            start_off, end_node, end_off, path_ptr, path, n_reads, n_path_nodes,
                            nullptr,
                            [](const char* b, const char* e, long long& s, long long& t,
                               std::vector<int32_t>& path_out, long long& f, long long& l) {
                                return parse_line(b, e, s, t, path_out, f, l);
                            });
}

extern "C" int gpc_parse_vg_alignments_host(const char* text, uint64_t n_bytes,
                                            int32_t n_threads, int32_t min_node,
                                            int32_t max_node, int32_t* start_node,
                                            int32_t* start_off, int32_t* end_node,
                                            int32_t* end_off, int64_t* path_ptr, int32_t* path,
                                            uint64_t* n_reads, uint64_t* n_path_nodes,
                                            uint64_t* n_without_path) {
    const long long lo = min_node, hi = max_node;
    return parse_reads_text("vg alignment file", 2, lo, hi, text, n_bytes, n_threads, start_node,
                            start_off, end_node, end_off, path_ptr, path, n_reads, n_path_nodes,
                            n_without_path,
                            [lo, hi](const char* b, const char* e, long long& s, long long& t,
                                     std::vector<int32_t>& path_out, long long& f, long long& l) {
                                return parse_vg_line(b, e, lo, hi, s, t, path_out, f, l);
                            });
}

extern "C" int gpc_parse_vg_graph_host(const char* text, uint64_t n_bytes, int32_t* node_id,
                                       int32_t* node_len, uint64_t* n_nodes,
                                       int32_t* edge_from, int32_t* edge_to,
                                       uint64_t* n_edges) {
    const bool fill = node_id != nullptr;
    const uint64_t want_nodes = *n_nodes, want_edges = *n_edges;
    if (fill && (!node_len || !edge_from || !edge_to)) {
        set_error("gpc_parse_vg_graph_host: all output arrays or none");
        return GPC_ERR_ARG;
    }
    // one pass: entries beyond the room given are counted, not stored
    VgGraphSink sink{node_id, node_len, edge_from, edge_to};
    sink.node_room = fill ? want_nodes : 0;
    sink.edge_room = fill ? want_edges : 0;
    Chunk whole;
    whole.b = text;
    whole.e = text + n_bytes;
    long long line = 0, bad = -1;
    for_lines(whole, [&](const char* b, const char* e) {
        ++line;
        Cursor c{b, e};
        if (bad < 0 && !parse_vg_graph_object(c, sink)) bad = line;
    });
    if (bad >= 0) {
        set_error("vg graph file: malformed JSON (non-blank line %lld)", bad);
        return GPC_ERR_ARG;
    }
    *n_nodes = sink.nodes;
    *n_edges = sink.edges;
    if (fill && (want_nodes < sink.nodes || want_edges < sink.edges)) return GPC_ERR_CAPACITY;
    return GPC_OK;
}

// Node sequences of a vg graph in JSON form (the same files as above): the bases of
// every node, concatenated in file order, for the sequence graph that turns peaks
// into FASTA records (reference: SequenceGraph.set_sequences_using_vg_json_graph,
// called at preprocess_interface.py:60-62,65-66).  Sequences must be plain text
// (no escapes).
extern "C" int gpc_parse_vg_sequences_host(const char* text, uint64_t n_bytes, int32_t* node_id,
                                           int64_t* seq_ptr, char* seq, uint64_t* n_nodes,
                                           uint64_t* n_bases) {
    const bool fill = node_id != nullptr;
    const uint64_t want_nodes = *n_nodes, want_bases = *n_bases;
    if (fill && (!seq_ptr || !seq)) {
        set_error("gpc_parse_vg_sequences_host: all output arrays or none");
        return GPC_ERR_ARG;
    }
    // one pass: what does not fit the room given is counted, not stored
    unsigned long long nodes = 0, bases = 0;
    long long line = 0, bad = -1;
    Chunk whole;
    whole.b = text;
    whole.e = text + n_bytes;
    for_lines(whole, [&](const char* b, const char* e) {
        ++line;
        if (bad >= 0) return;
        Cursor c{b, e};
        const bool ok = for_members(c, [&](const char* k, size_t n) {
            if (!key_is(k, n, "node")) return skip_value(c);
            return for_items(c, [&]() {
                long long id = 0;
                bool have_id = false;
                const char* s0 = nullptr;
                long long len = 0;
                if (!for_members(c, [&](const char* k2, size_t n2) {
                        if (key_is(k2, n2, "id")) { have_id = true; return parse_json_int(c, id); }
                        if (key_is(k2, n2, "sequence")) {
                            c.skip_ws();
                            if (c.eof() || *c.p != '"') return false;
                            s0 = c.p + 1;
                            if (!string_length(c, len)) return false;
                            return (c.p - 1 - s0) == len;          // no escapes inside
                        }
                        return skip_value(c);
                    }))
                    return false;
                if (!have_id || id <= 0 || id > 2147483647ll) return false;
                if (fill && nodes < want_nodes && bases + (unsigned long long)len <= want_bases) {
                    node_id[nodes] = (int32_t)id;
                    seq_ptr[nodes] = (int64_t)bases;
                    if (len) memcpy(seq + bases, s0, (size_t)len);
                }
                ++nodes;
                bases += (unsigned long long)len;
                return true;
            });
        });
        if (!ok) bad = line;
    });
    if (bad >= 0) {
        set_error("vg graph file: malformed JSON or escaped sequence (non-blank line %lld)", bad);
        return GPC_ERR_ARG;
    }
    *n_nodes = nodes;
    *n_bases = bases;
    if (fill) {
        if (want_nodes < nodes || want_bases < bases) return GPC_ERR_CAPACITY;
        seq_ptr[nodes] = (int64_t)bases;
    }
    return GPC_OK;
}

// Shards an alignment file by chromosome (reference
// preprocess_interface.py:83-145): every line goes, byte for byte, to the output of
// the first node-id range that holds the first node id found in the line -- the
// first place where `node_id":` is followed by at most one white-space character, an
// optional quote and digits (the reference's regular expression, :109) -- or to the
// last output ("unmapped") when there is no such place or no range holds the id.
// fds: n_ranges + 1 open file descriptors; n_lines[n_ranges + 1] counts the lines
// each one received.  One pass at memchr speed with a 4 MB buffer per output, where
// the reference runs a Python regular expression per line.

namespace gpc {
namespace {

struct OutBuffer {
    int fd = -1;
    std::vector<char> data;
    bool failed = false;
    void flush() {
        size_t done = 0;
        while (done < data.size() && !failed) {
            const ssize_t k = ::write(fd, data.data() + done, data.size() - done);
            if (k <= 0) failed = true; else done += (size_t)k;
        }
        data.clear();
    }
    void add(const char* b, const char* e) {
        if (data.size() + (size_t)(e - b) > (4u << 20)) flush();
        if ((size_t)(e - b) > (4u << 20)) {          // a line larger than the buffer: straight out
            data.assign(b, e);
            flush();
            return;
        }
        data.insert(data.end(), b, e);
    }
};

// first node id of a line, or false
bool first_node_id(const char* b, const char* e, unsigned long long& id, bool& huge) {
    static const char key[] = "node_id\":";
    const size_t klen = sizeof(key) - 1;
    const char* p = b;
    while (p < e) {
        const char* hit = (const char*)memmem(p, (size_t)(e - p), key, klen);
        if (!hit) return false;
        const char* q = hit + klen;
        const char* digits = nullptr;
        // \s? "? [0-9]+ with the regular expression's backtracking: the optional parts may
        // each be taken or left
        for (int ws = 1; ws >= 0 && !digits; --ws) {
            const char* r = q;
            if (ws) {
                if (r < e && (*r == ' ' || *r == '\t' || *r == '\n' || *r == '\r' || *r == '\f' || *r == '\v')) ++r;
                else continue;
            }
            for (int quote = 1; quote >= 0 && !digits; --quote) {
                const char* s = r;
                if (quote) {
                    if (s < e && *s == '"') ++s; else continue;
                }
                if (s < e && *s >= '0' && *s <= '9') digits = s;
            }
        }
        if (digits) {
            id = 0;
            huge = false;
            for (const char* s = digits; s < e && *s >= '0' && *s <= '9'; ++s) {
                if (id > 900000000000000000ull) huge = true; else id = id * 10 + (unsigned)(*s - '0');
            }
            return true;
        }
        p = hit + 1;
    }
    return false;
}

}  // namespace
}  // namespace gpc

extern "C" int gpc_split_alignments_host(const char* text, uint64_t n_bytes, const int64_t* range_lo,
                                         const int64_t* range_hi, uint64_t n_ranges,
                                         const int32_t* fds, uint64_t* n_lines) {
    if ((!text && n_bytes) || !fds || !n_lines || (n_ranges && (!range_lo || !range_hi))) {
        set_error("gpc_split_alignments_host: bad arguments");
        return GPC_ERR_ARG;
    }
    std::vector<OutBuffer> outs(n_ranges + 1);
    for (uint64_t i = 0; i <= n_ranges; ++i) {
        outs[i].fd = fds[i];
        outs[i].data.reserve(4u << 20);
        n_lines[i] = 0;
    }
    const char* p = text;
    const char* end = text + n_bytes;
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* e = nl ? nl + 1 : end;
        uint64_t where = n_ranges;
        unsigned long long id = 0;
        bool huge = false;
        if (first_node_id(p, e, id, huge) && !huge) {
            for (uint64_t i = 0; i < n_ranges; ++i)
                if ((long long)id >= range_lo[i] && (long long)id <= range_hi[i]) { where = i; break; }
        }
        outs[where].add(p, e);
        n_lines[where] += 1;
        p = e;
    }
    for (auto& o : outs) {
        o.flush();
        if (o.failed) { set_error("gpc_split_alignments_host: write failed"); return GPC_ERR_INTERNAL; }
    }
    return GPC_OK;
}
