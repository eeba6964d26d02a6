// idelucs_b200 — host-side FASTA ingest: the record loop of kmersFasta
// (idelucs/utils.py:229-261), run ONCE per file instead of once per pass, straight from the file
// image into one flat (ideally pinned) byte buffer that idl_pack consumes on the device.
//
// Semantics kept from the reference loop (Python iterates the binary file line by line, lines end at '\n'):
//   * a line starting with '#' is skipped (:230);
//   * a line starting with '>' closes the running record only while the running id is non-empty (:233-252;
//     an empty id keeps accumulating into the same record), then id = line[1:-1] (:253);
//   * any other line contributes line.strip() (:257; bytes.strip() = ASCII whitespace " \t\n\r\v\f");
//   * the last record is always closed (:259-261), even with an empty id.
// Alphabet handling (check_sequence, :26-51) is not done here: idl_pack does it on the device.
//
// Parallel form: the stripped sequence bytes land in the output in FILE ORDER whatever the record logic decides
// (a record that is not closed simply keeps growing), so the file is cut at line boundaries into one chunk per
// thread; every chunk is walked independently (header lines + stripped bytes before each of them), a serial pass
// over the header lines alone applies the flush rules, and the chunks copy their pieces to known offsets.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"

namespace {

inline bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

struct Header { int64_t off, len, bytes_before; };   // id span in the file; stripped sequence bytes of the chunk before this line

struct Chunk {
    int64_t begin, end;            // [begin, end) of the file image, whole lines
    int64_t seq_bytes;             // stripped sequence bytes in the chunk
    std::vector<Header> headers;
};

// walks the lines of [begin, end): header(off, len) per '>' line, piece(ptr, len) per non-empty stripped sequence line
template <class Piece, class Hdr>
void walk_lines(const uint8_t* buf, int64_t begin, int64_t end, Piece piece, Hdr header) {
    int64_t pos = begin;
    while (pos < end) {
        const void* nl = memchr(buf + pos, '\n', (size_t)(end - pos));
        const int64_t e = nl ? (const uint8_t*)nl - buf + 1 : end;
        const uint8_t c = buf[pos];
        if (c == '#') {
        } else if (c == '>') {
            header(pos + 1, e - pos >= 2 ? e - pos - 2 : 0);
        } else {
            int64_t a = pos, b = e;
            while (a < b && is_space(buf[a])) ++a;
            while (b > a && is_space(buf[b - 1])) --b;
            if (b > a) piece(buf + a, b - a);
        }
        pos = e;
    }
}

int n_threads_for(int64_t nbytes) {
    int64_t per = 4 << 20;                                   // at least 4 MB of file per thread
    if (const char* e = getenv("IDL_FASTA_CHUNK")) { const long long v = atoll(e); if (v > 0) per = v; }   // tests: force many chunks
    int64_t t = nbytes / per;
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    if (hw > 32) hw = 32;
    if (const char* e = getenv("IDL_FASTA_THREADS")) { const int v = atoi(e); if (v > 0) hw = v; }
    if (t > hw) t = hw;
    return t < 1 ? 1 : (int)t;
}

template <class F>
void run_parallel(int n, F f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (int i = 1; i < n; ++i) th.emplace_back([&f, i] { f(i); });
    f(0);
    for (auto& t : th) t.join();
}

// cut at line boundaries and index every chunk (pass A)
std::vector<Chunk> index_chunks(const uint8_t* buf, int64_t n) {
    const int T = n_threads_for(n);
    std::vector<Chunk> ch(T);
    int64_t prev = 0;
    for (int t = 0; t < T; ++t) {
        int64_t cut = n;
        if (t + 1 < T) {
            cut = n / T * (t + 1);
            if (cut < prev) cut = prev;
            const void* nl = cut < n ? memchr(buf + cut, '\n', (size_t)(n - cut)) : nullptr;
            cut = nl ? (const uint8_t*)nl - buf + 1 : n;
        }
        ch[t].begin = prev;
        ch[t].end = cut;
        prev = cut;
    }
    run_parallel(T, [&](int t) {
        Chunk& c = ch[t];
        c.seq_bytes = 0;
        walk_lines(buf, c.begin, c.end, [&](const uint8_t*, int64_t len) { c.seq_bytes += len; },
                   [&](int64_t off, int64_t len) { c.headers.push_back(Header{off, len, c.seq_bytes}); });
    });
    return ch;
}

// the flush rules over the header lines alone (pass B): rec(hdr_off, hdr_len, end_of_record_in_output)
template <class Rec>
int64_t stitch(const std::vector<Chunk>& ch, Rec rec) {
    int64_t hdr_off = 0, hdr_len = 0, base = 0;
    for (const Chunk& c : ch) {
        for (const Header& h : c.headers) {
            if (hdr_len > 0) rec(hdr_off, hdr_len, base + h.bytes_before);
            hdr_off = h.off;
            hdr_len = h.len;
        }
        base += c.seq_bytes;
    }
    rec(hdr_off, hdr_len, base);
    return base;
}

// the index idl_fasta_scan built, handed to the idl_fasta_extract call that follows on the same thread for the same image
// (consumed by it: one scan, one extract; anything else re-indexes)
struct IndexCache { const uint8_t* buf = nullptr; int64_t n = -1; std::vector<Chunk> ch; };
thread_local IndexCache g_index;

}  // namespace

extern "C" {

int idl_fasta_scan(const uint8_t* buf, int64_t nbytes, int64_t* n_records, int64_t* n_seq_bytes) {
    if ((!buf && nbytes > 0) || nbytes < 0 || !n_records || !n_seq_bytes) return idl::set_error(IDL_EINVAL, "idl_fasta_scan: bad argument%s", "");
    g_index.ch = index_chunks(buf, nbytes);
    g_index.buf = buf;
    g_index.n = nbytes;
    int64_t nr = 0;
    *n_seq_bytes = stitch(g_index.ch, [&](int64_t, int64_t, int64_t) { ++nr; });
    *n_records = nr;
    return IDL_OK;
}

int idl_fasta_extract(const uint8_t* buf, int64_t nbytes, int64_t n_records, uint8_t* seq_out, int64_t seq_cap,
                      int64_t* byte_off, int64_t* hdr_off, int64_t* hdr_len) {
    if ((!buf && nbytes > 0) || nbytes < 0 || n_records < 1 || (!seq_out && seq_cap > 0) || !byte_off || !hdr_off || !hdr_len)
        return idl::set_error(IDL_EINVAL, "idl_fasta_extract: bad argument%s", "");
    std::vector<Chunk> ch;
    if (g_index.buf == buf && g_index.n == nbytes) ch.swap(g_index.ch);
    else ch = index_chunks(buf, nbytes);
    g_index.buf = nullptr;
    g_index.n = -1;
    g_index.ch.clear();
    int64_t r = 0;
    byte_off[0] = 0;
    const int64_t total = stitch(ch, [&](int64_t ho, int64_t hl, int64_t end) {
        if (r < n_records) { hdr_off[r] = ho; hdr_len[r] = hl; byte_off[r + 1] = end; }
        ++r;
    });
    if (total > seq_cap || r != n_records)
        return idl::set_error(IDL_EINVAL, "idl_fasta_extract: buffers do not match idl_fasta_scan%s (records found: %lld)", "", (long long)r);
    std::vector<int64_t> base(ch.size() + 1, 0);
    for (size_t t = 0; t < ch.size(); ++t) base[t + 1] = base[t] + ch[t].seq_bytes;
    run_parallel((int)ch.size(), [&](int t) {   // pass C: every chunk copies its pieces to its own range of the output
        uint8_t* w = seq_out + base[t];
        walk_lines(buf, ch[t].begin, ch[t].end, [&](const uint8_t* p, int64_t len) { memcpy(w, p, (size_t)len); w += len; },
                   [](int64_t, int64_t) {});
    });
    return IDL_OK;
}

}  // extern "C"
