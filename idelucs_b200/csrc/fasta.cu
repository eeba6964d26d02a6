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
#include <cstdint>
#include <cstring>

#include "common.h"

namespace {

inline bool is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// calls rec(hdr_off, hdr_len, seq_begin, seq_end) per closed record; piece(ptr, len) per stripped sequence line
template <class Piece, class Rec>
void fasta_walk(const uint8_t* buf, int64_t n, Piece piece, Rec rec) {
    int64_t hdr_off = 0, hdr_len = 0;
    int64_t pos = 0;
    while (pos < n) {
        const void* nl = memchr(buf + pos, '\n', (size_t)(n - pos));
        const int64_t end = nl ? (const uint8_t*)nl - buf + 1 : n;
        const uint8_t c = buf[pos];
        if (c == '#') {
        } else if (c == '>') {
            if (hdr_len > 0) rec(hdr_off, hdr_len);
            hdr_off = pos + 1;
            hdr_len = end - pos >= 2 ? end - pos - 2 : 0;
        } else {
            int64_t a = pos, b = end;
            while (a < b && is_space(buf[a])) ++a;
            while (b > a && is_space(buf[b - 1])) --b;
            if (b > a) piece(buf + a, b - a);
        }
        pos = end;
    }
    rec(hdr_off, hdr_len);
}

}  // namespace

extern "C" {

int idl_fasta_scan(const uint8_t* buf, int64_t nbytes, int64_t* n_records, int64_t* n_seq_bytes) {
    if ((!buf && nbytes > 0) || nbytes < 0 || !n_records || !n_seq_bytes) return idl::set_error(IDL_EINVAL, "idl_fasta_scan: bad argument%s", "");
    int64_t nr = 0, nb = 0;
    fasta_walk(buf, nbytes, [&](const uint8_t*, int64_t len) { nb += len; }, [&](int64_t, int64_t) { ++nr; });
    *n_records = nr;
    *n_seq_bytes = nb;
    return IDL_OK;
}

int idl_fasta_extract(const uint8_t* buf, int64_t nbytes, int64_t n_records, uint8_t* seq_out, int64_t seq_cap,
                      int64_t* byte_off, int64_t* hdr_off, int64_t* hdr_len) {
    if ((!buf && nbytes > 0) || nbytes < 0 || n_records < 1 || (!seq_out && seq_cap > 0) || !byte_off || !hdr_off || !hdr_len)
        return idl::set_error(IDL_EINVAL, "idl_fasta_extract: bad argument%s", "");
    int64_t r = 0, w = 0;
    bool overflow = false;
    byte_off[0] = 0;
    fasta_walk(buf, nbytes,
               [&](const uint8_t* p, int64_t len) {
                   if (w + len > seq_cap) { overflow = true; return; }
                   memcpy(seq_out + w, p, (size_t)len);
                   w += len;
               },
               [&](int64_t ho, int64_t hl) {
                   if (r < n_records) { hdr_off[r] = ho; hdr_len[r] = hl; byte_off[r + 1] = w; }
                   ++r;
               });
    if (overflow || r != n_records) return idl::set_error(IDL_EINVAL, "idl_fasta_extract: buffers do not match idl_fasta_scan%s (records found: %lld)", "", (long long)r);
    return IDL_OK;
}

}  // extern "C"
