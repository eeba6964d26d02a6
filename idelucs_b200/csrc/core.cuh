// idelucs_b200 — per-thread building blocks of the k-mer / mimic kernels.
//
// Everything in this header is plain integer / IEEE arithmetic on one thread's data and is
// compiled BOTH by nvcc (device code of kernels.cu) and by g++ (tests/host_emul.cpp, which
// replays the same per-thread routines sequentially on the CPU so the index logic can be
// checked in a container without a GPU).  No shared memory, shuffles or atomics in here.
//
// Data layout:
//   codes : uint32 words, 16 bases per word, BIG-endian inside the word (base j of a word at
//           bits [30-2j, 31-2j]); A=0 C=1 G=2 T=3 (idelucs/kmers.pyx:19-34 numbering), so a
//           run of k consecutive bases read as a bit-field IS the k-mer index of
//           kmers.pyx:45 (first base most significant).
//   nmask : uint32 words, 32 bases per word, base j at bit 31-j; 1 = "resets the window"
//           (N / IUPAC / gap after check_sequence, any non-ACGT byte in strict mode, and the
//           padding bases behind the end of a sequence).
//   every sequence starts on a 64-base chunk (4 code words = 16 B, 2 mask words = 8 B).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IDL_HD __host__ __device__ __forceinline__
#else
#define IDL_HD inline
#include <math.h>
#endif

namespace idl {

constexpr int CHUNK_BASES = 64;
constexpr int RNG_BLOCK = 64;        // bases per Bernoulli generation block (= one chunk)
constexpr uint32_t VAL_N = 4;        // edit value "set N"

enum Kind : int { KIND_CLEAN = 0, KIND_TRANSITION = 1, KIND_TRANSVERSION = 2, KIND_BOTH = 3,
                  KIND_RANDOM_N = 4, KIND_EXPLICIT = 5 };
enum Stream : uint32_t { STREAM_TRANSITION = 0, STREAM_TRANSVERSION = 1, STREAM_RANDOM_N = 2 };

// ---------------------------------------------------------------------------------------
// alphabet (idelucs/utils.py:42-46 check_sequence; idelucs/kmers.pyx:19-34 strict LUT)
// class: 0..3 = A C G T, 4 = N-class (window reset), 5 = deleted byte (' ' \t \n \r),
//        6 = invalid byte (check_sequence raises ValueError, utils.py:46-50)
// ---------------------------------------------------------------------------------------
IDL_HD uint32_t base_class(uint32_t b, int strict) {
    if (strict) {  // kmers.pyx LUT: only uppercase ACGT are bases, everything else resets
        return b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : 4u;
    }
    switch (b) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        case 'N': case 'n': case '-':
        case 'S': case 's': case 'W': case 'w': case 'K': case 'k': case 'M': case 'm':
        case 'Y': case 'y': case 'R': case 'r': case 'B': case 'b': case 'D': case 'd':
        case 'H': case 'h': case 'V': case 'v': return 4;
        case ' ': case '\t': case '\n': case '\r': return 5;
        default: return 6;
    }
}

// ---------------------------------------------------------------------------------------
// bit helpers
// ---------------------------------------------------------------------------------------
IDL_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {  // low 32 bits of (hi:lo) >> sh, sh in [0,31]
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? ((lo >> sh) | (hi << (32u - sh))) : lo;
#endif
}
IDL_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
IDL_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
IDL_HD uint32_t code_at(const uint32_t* codes, int pos) {
    return (codes[pos >> 4] >> (30 - 2 * (pos & 15))) & 3u;
}
IDL_HD uint32_t nflag_at(const uint32_t* nmask, int pos) {
    return (nmask[pos >> 5] >> (31 - (pos & 31))) & 1u;
}

// ---------------------------------------------------------------------------------------
// K1: pack 16 ASCII bytes -> one code word + 16 mask bits (bit 15-j for base j).
// `avail` = number of real bases in this word (0..16); the rest is padding (mask 1).
// Returns the index (0..15) of the first invalid/deleted byte in *bad (16 if none);
// *bad_class tells which (5 deleted, 6 invalid).
// ---------------------------------------------------------------------------------------
IDL_HD void pack16(const uint8_t* src, int avail, int strict, uint32_t* code_word, uint32_t* mask16,
                   int* bad, int* bad_class) {
    uint32_t cw = 0, mw = 0;
    int first_bad = 16, cls_bad = 0;
    for (int j = 0; j < 16; ++j) {
        uint32_t cls = 4;
        if (j < avail) {
            cls = base_class(src[j], strict);
            if (cls >= 5) {
                if (first_bad == 16) { first_bad = j; cls_bad = (int)cls; }
                cls = 4;
            }
        }
        cw = (cw << 2) | (cls & 3u & (cls < 4 ? 3u : 0u));
        mw = (mw << 1) | (cls >= 4 ? 1u : 0u);
    }
    *code_word = cw; *mask16 = mw; *bad = first_bad; *bad_class = cls_bad;
}

// ---------------------------------------------------------------------------------------
// K2: count the 32 window ends of half-chunk `h` (restates idelucs/kmers.pyx:38-50; the
// rolling k_mer/countdown state of the reference is replaced by reading the k-1 preceding
// bases from the previous word, so half-chunks are independent).  w0,w1 = code words 2h,
// 2h+1.  upd(kmer_index) is called once per counted window; returns the number counted.
// ---------------------------------------------------------------------------------------
template <int K, class Upd>
IDL_HD int count_half(const uint32_t* codes, const uint32_t* nmask, int h, uint32_t w0, uint32_t w1, Upd upd) {
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1u;
    constexpr uint32_t PM = (1u << (K - 1)) - 1u;
    const uint32_t prev = h > 0 ? codes[2 * h - 1] : 0u;
    const uint32_t m = nmask[h];
    // P: bit d-1 = N flag of position -d relative to the half-chunk (d = 1..K-1); the sequence
    // start behaves like kmers.pyx:14 countdown = k-1, i.e. as if preceded by resets.
    const uint32_t P = h > 0 ? (nmask[h - 1] & PM) : PM;
    uint32_t inv = m;
#pragma unroll
    for (int d = 1; d < K; ++d) inv |= (m >> d) | (P << (32 - d));
    const uint32_t w[3] = {prev, w0, w1};
#pragma unroll
    for (int wi = 0; wi < 2; ++wi) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (!((inv >> (31 - 16 * wi - j)) & 1u)) upd(funnel_r(w[wi + 1], w[wi], 30 - 2 * j) & KMASK);
        }
    }
#if defined(__CUDA_ARCH__)
    return __popc(~inv);
#else
    return __builtin_popcount(~inv);
#endif
}

// ---------------------------------------------------------------------------------------
// window of W = 2K-1 consecutive bases starting at q0 (may be negative: positions < 0 read
// as N).  base t of the window (position q0+t) sits at bits [2(W-1-t), +2) of `bases` and
// its N flag at bit W-1-t of `nbits`.  Reads one word past the window's last word (the
// packed buffers carry a slack chunk, and positions >= L are never used by callers).
// ---------------------------------------------------------------------------------------
template <int K>
struct Window { uint32_t bases; uint32_t nbits; };

template <int K>
IDL_HD Window<K> load_window(const uint32_t* codes, const uint32_t* nmask, int q0) {
    constexpr int W = 2 * K - 1;
    const int d = q0 < 0 ? -q0 : 0;
    const int q = q0 + d;
    const uint64_t cv = ((uint64_t)codes[q >> 4] << 32) | (uint64_t)codes[(q >> 4) + 1];
    const uint64_t nv = ((uint64_t)nmask[q >> 5] << 32) | (uint64_t)nmask[(q >> 5) + 1];
    uint32_t b = (uint32_t)(cv >> (64 - 2 * (q & 15) - 2 * W)) & ((1u << (2 * W)) - 1u);
    uint32_t n = (uint32_t)(nv >> (64 - (q & 31) - W)) & ((1u << W) - 1u);
    b >>= 2 * d;
    n = (n >> d) | (((1u << d) - 1u) << (W - d));
    Window<K> w; w.bases = b; w.nbits = n;
    return w;
}

// ---------------------------------------------------------------------------------------
// K3 delta: entry i of a position-sorted edit list (entry = pos<<3 | val, val 0..3 = set
// base, 4 = set N).  Entry i owns the window ends e in [p_i, min(p_i+K-1, p_{i+1}-1, L-1)];
// for each it removes the clean k-mer (if the clean window was counted) and adds the
// mutated one (if the mutated window is countable), calling upd(kmer, -1 / +1).  Mutated
// window content = clean bases + this edit + the preceding edits within K-1 positions
// (later edits inside an owned window cannot exist by construction).  Returns the change
// of the number of counted windows.  Equivalent to re-running kmers.pyx:38-50 on the
// mutated sequence, restricted to the windows the edit touches.
// ---------------------------------------------------------------------------------------
// `list` is anything indexable by int (a plain array, or a view that stitches per-block slots together).
template <int K, class List, class Upd>
IDL_HD int apply_entry(const uint32_t* codes, const uint32_t* nmask, int L, const List& list, int n, int i,
                       Upd upd) {
    constexpr int W = 2 * K - 1;
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1u;
    constexpr uint32_t NMASKK = (1u << K) - 1u;
    const uint32_t ent = list[i];
    const int p = (int)(ent >> 3);
    const int pn = (i + 1 < n) ? (int)(list[i + 1] >> 3) : 0x7fffffff;
    int e_hi = p + K - 1;
    if (pn - 1 < e_hi) e_hi = pn - 1;
    if (L - 1 < e_hi) e_hi = L - 1;
    if (e_hi < p) return 0;
    const int q0 = p - (K - 1);
    const Window<K> cw = load_window<K>(codes, nmask, q0);
    uint32_t mb = cw.bases, mn = cw.nbits;
    for (int j = i; j >= 0; --j) {
        const uint32_t ej = list[j];
        const int pj = (int)(ej >> 3);
        if (pj < q0) break;
        const int sh = W - 1 - (pj - q0);
        const uint32_t v = ej & 7u;
        if (v < 4u) { mb = (mb & ~(3u << (2 * sh))) | (v << (2 * sh)); mn &= ~(1u << sh); }
        else        { mn |= (1u << sh); }
    }
    int dtot = 0;
    for (int e = p; e <= e_hi; ++e) {
        const int sh = K - 1 - (e - p);
        const bool okc = ((cw.nbits >> sh) & NMASKK) == 0u;
        const bool okm = ((mn >> sh) & NMASKK) == 0u;
        const uint32_t kc = (cw.bases >> (2 * sh)) & KMASK;
        const uint32_t km = (mb >> (2 * sh)) & KMASK;
        if (okc && okm && kc == km) continue;
        if (okc) { upd(kc, -1); --dtot; }
        if (okm) { upd(km, +1); ++dtot; }
    }
    return dtot;
}

// apply_entry with the deltas returned in registers instead of through a callback (statically indexed: one word per owned
// window end): d[t] = removed k-mer | 0x1000 when the clean window ending at p+t is uncounted by the edit, (added k-mer << 16)
// | 0x10000000 when a mutated window is counted there.  Returns the number of +-1 updates (0..2K); *dtot = change of the
// number of counted windows.  One evaluation serves both the reservation of list space and the writes.
template <int K, class List>
IDL_HD int entry_deltas(const uint32_t* codes, const uint32_t* nmask, int L, const List& list, int n, int i, uint32_t (&d)[K], int* dtot) {
    constexpr int W = 2 * K - 1;
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1u;
    constexpr uint32_t NMASKK = (1u << K) - 1u;
    const uint32_t ent = list[i];
    const int p = (int)(ent >> 3);
    const int pn = (i + 1 < n) ? (int)(list[i + 1] >> 3) : 0x7fffffff;
    int e_hi = p + K - 1;
    if (pn - 1 < e_hi) e_hi = pn - 1;
    if (L - 1 < e_hi) e_hi = L - 1;
    const int q0 = p - (K - 1);
    const Window<K> cw = load_window<K>(codes, nmask, q0);
    uint32_t mb = cw.bases, mn = cw.nbits;
    for (int j = i; j >= 0; --j) {
        const uint32_t ej = list[j];
        const int pj = (int)(ej >> 3);
        if (pj < q0) break;
        const int sh = W - 1 - (pj - q0);
        const uint32_t v = ej & 7u;
        if (v < 4u) { mb = (mb & ~(3u << (2 * sh))) | (v << (2 * sh)); mn &= ~(1u << sh); }
        else        { mn |= (1u << sh); }
    }
    int cnt = 0, dt = 0;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        uint32_t v = 0u;
        if (p + t <= e_hi) {
            const int sh = K - 1 - t;
            const bool okc = ((cw.nbits >> sh) & NMASKK) == 0u;
            const bool okm = ((mn >> sh) & NMASKK) == 0u;
            const uint32_t kc = (cw.bases >> (2 * sh)) & KMASK;
            const uint32_t km = (mb >> (2 * sh)) & KMASK;
            if (!(okc && okm && kc == km)) {
                if (okc) { v |= kc | 0x1000u; ++cnt; --dt; }
                if (okm) { v |= (km << 16) | 0x10000000u; ++cnt; ++dt; }
            }
        }
        d[t] = v;
    }
    *dtot = dt;
    return cnt;
}

// ---------------------------------------------------------------------------------------
// Random_N specialisation of the delta rule (idelucs/utils.py:89-95 sets bases to N, so
// windows are only ever REMOVED): draw i of an UNSORTED list of n draws (entry = pos<<3|4)
// removes the clean windows ending in [p_i, min(p_i+K-1, p_next-1, L-1)], p_next = smallest
// drawn position > p_i; among duplicate draws of one position only the first acts.
// emit(kmer) is called once per removed window.  Same result as apply_entry on the sorted
// list, without sorting or rebuilding the mutated window.
// ---------------------------------------------------------------------------------------
template <int K, class Emit>
IDL_HD void random_n_removals(const uint32_t* codes, const uint32_t* nmask, int L, const uint32_t* ent, int n, int i,
                              Emit emit) {
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1u;
    constexpr uint32_t NMASKK = (1u << K) - 1u;
    const int p = (int)(ent[i] >> 3);
    int pn = 0x7fffffff;
    for (int j = 0; j < n; ++j) {
        const int pj = (int)(ent[j] >> 3);
        if (pj == p && j < i) return;
        if (pj > p && pj < pn) pn = pj;
    }
    int e_hi = p + K - 1;
    if (pn - 1 < e_hi) e_hi = pn - 1;
    if (L - 1 < e_hi) e_hi = L - 1;
    const Window<K> cw = load_window<K>(codes, nmask, p - (K - 1));
    for (int e = p; e <= e_hi; ++e) {
        const int sh = K - 1 - (e - p);
        if (((cw.nbits >> sh) & NMASKK) == 0u) emit((cw.bases >> (2 * sh)) & KMASK);
    }
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter (c0,c1,c2,c3), key (k0,k1)
// ---------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };

IDL_HD U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// T[g-1] = floor((1-(1-p)^g) * 2^32), g = 1..RNG_BLOCK (host only; IEEE double, no pow()).
inline void geometric_table(double p, uint32_t* T) {
    volatile double q = 1.0;
    const double omp = 1.0 - p;
    for (int g = 0; g < RNG_BLOCK; ++g) {
        q = q * omp;
        double v = (1.0 - q) * 4294967296.0;
        T[g] = v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v;
    }
}
// 1/log2(1-p): slope of the gap estimate used by the device-side search (any value is
// safe: the estimate is always corrected against the integer table)
inline float gap_slope(double p) {
    if (!(p > 0.0)) return 0.0f;
    if (!(p < 1.0)) return 0.0f;
    return (float)(1.0 / log2(1.0 - p));
}

// smallest g in 1..RNG_BLOCK with u < T[g-1]; 0 when u >= T[RNG_BLOCK-1].  The result is
// DEFINED by the integer table; the device starts the search from a log estimate, the host
// uses a binary search — both return the same g.
IDL_HD int gap_of(uint32_t u, const uint32_t* T, float slope) {
    if (u >= T[RNG_BLOCK - 1]) return 0;
#if defined(__CUDA_ARCH__)
    const float x = (float)(~u) * 2.3283064365386963e-10f;  // ~ 1 - u/2^32
    int g = (int)(__log2f(x) * slope) + 1;
    g = g < 1 ? 1 : (g > RNG_BLOCK ? RNG_BLOCK : g);
    while (g > 1 && u < T[g - 2]) --g;
    while (u >= T[g - 1]) ++g;  // terminates: u < T[RNG_BLOCK-1]
    return g;
#else
    (void)slope;
    int lo = 0, hi = RNG_BLOCK - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u < T[mid]) hi = mid; else lo = mid + 1;
    }
    return lo + 1;
#endif
}

// Bernoulli(p) hit positions of one 64-base block by geometric gap skipping.  Word stream
// = Philox outputs for counter (j, block, seq_id, variant*4+stream), j = 0,1,...; the
// transition stream uses one word per step (gap), the transversion stream two (gap, choice
// = top bit of the second word; idelucs/utils.py:118 random.choice of two).
struct BernGen {
    U4 r; uint32_t j, wi;
    uint32_t c1, c2, c3, k0, k1;
    int pos, end;
    bool with_choice, done;
    const uint32_t* T;
    float slope;

    IDL_HD void init(uint64_t seed, uint32_t seq_id, uint32_t variant, uint32_t stream, int block, int L,
                     const uint32_t* table, float slp, bool choice) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        c1 = (uint32_t)block; c2 = seq_id; c3 = (variant << 2) | stream;
        j = 0; wi = 4; T = table; slope = slp; with_choice = choice;
        pos = block * RNG_BLOCK - 1;
        end = (block + 1) * RNG_BLOCK; if (L < end) end = L;
        done = (block * RNG_BLOCK >= L);
    }
    IDL_HD uint32_t word() {
        if (wi == 4) { r = philox4x32_10(j, c1, c2, c3, k0, k1); ++j; wi = 0; }
        const uint32_t v = wi == 0 ? r.x : wi == 1 ? r.y : wi == 2 ? r.z : r.w;
        ++wi;
        return v;
    }
    // next hit: returns false when the block is exhausted
    IDL_HD bool next(int* p, uint32_t* choice) {
        if (done) return false;
        const int g = gap_of(word(), T, slope);
        const uint32_t ch = with_choice ? (word() >> 31) : 0u;
        if (g == 0) { done = true; return false; }
        pos += g;
        if (pos >= end) { done = true; return false; }
        *p = pos; *choice = ch;
        return true;
    }
};

// Merged, position-ordered edits of one block for kinds TRANSITION / TRANSVERSION / BOTH
// (idelucs/utils.py:65-76, 108-118, 132-135 on codes: transition = code ^ 2; transversion:
// purine (even code) -> [T,C][choice] = 3,1; pyrimidine (odd) -> [A,G][choice] = 0,2; N stays
// N -> no entry; when both streams hit a base the transversion decides, because it only
// depends on the purine/pyrimidine class that a transition preserves).  emit(entry).
template <class Emit>
IDL_HD int block_edits(int kind, uint64_t seed, uint32_t seq_id, uint32_t variant, int block, int L,
                       const uint32_t* codes, const uint32_t* nmask, const uint32_t* T1, float slope1,
                       const uint32_t* T2, float slope2, Emit emit) {
    BernGen ga, gb;
    int pa = 0x7fffffff, pb = 0x7fffffff;
    uint32_t ca = 0, cb = 0;
    bool va = false, vb = false;
    if (kind == KIND_TRANSITION || kind == KIND_BOTH) {
        ga.init(seed, seq_id, variant, STREAM_TRANSITION, block, L, T1, slope1, false);
        va = ga.next(&pa, &ca);
    }
    if (kind == KIND_TRANSVERSION || kind == KIND_BOTH) {
        gb.init(seed, seq_id, variant, STREAM_TRANSVERSION, block, L, T2, slope2, true);
        vb = gb.next(&pb, &cb);
    }
    int n = 0;
    while (va || vb) {
        const bool take_b = vb && (!va || pb <= pa);
        const int p = take_b ? pb : pa;
        uint32_t val;
        if (take_b) {
            if (va && pa == pb) va = ga.next(&pa, &ca);
            const uint32_t c = code_at(codes, p);
            val = (c & 1u) ? (cb << 1) : (1u | ((1u - cb) << 1));
            vb = gb.next(&pb, &cb);
        } else {
            val = code_at(codes, p) ^ 2u;
            va = ga.next(&pa, &ca);
        }
        if (!nflag_at(nmask, p)) { emit(((uint32_t)p << 3) | val); ++n; }
    }
    return n;
}

// ---------------------------------------------------------------------------------------
// MASK block generator: the same edits as block_edits(), with the hits of each stream kept as a 64-bit
// mask (bit i = block-relative position i).  Merging the two streams, dropping hits on N and ordering
// the edits by position are then a handful of logic instructions instead of list merges, and a block
// may hold any number of hits (the gap loop is data dependent; at the reference's rates a warp leaves
// it after one Philox call per stream).  Everything stays in registers.
// ---------------------------------------------------------------------------------------
IDL_HD uint32_t u4_sel(const U4& r, int i) { return i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w; }

IDL_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
IDL_HD int ffs64(uint64_t x) {   // index of the lowest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

// hits of one stream in one block (same word consumption as BernGen: one word per step, or gap + choice)
template <bool CHOICE>
IDL_HD uint64_t stream_mask(uint64_t seed, uint32_t seq_id, uint32_t variant, uint32_t stream, int block, int L,
                            const uint32_t* T, float slope, uint64_t& choice) {
    int end = L - block * RNG_BLOCK;
    if (end > RNG_BLOCK) end = RNG_BLOCK;
    uint64_t m = 0;
    choice = 0;
    if (end <= 0) return 0;
    int cur = -1;
    for (uint32_t j = 0;; ++j) {
        const U4 r = philox4x32_10(j, (uint32_t)block, seq_id, (variant << 2) | stream, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
        for (int s = 0; s < (CHOICE ? 2 : 4); ++s) {
            const int g = gap_of(CHOICE ? (s == 0 ? r.x : r.z) : u4_sel(r, s), T, slope);
            if (g == 0 || cur + g >= end) return m;
            cur += g;
            m |= 1ull << cur;
            if (CHOICE) choice |= (uint64_t)((s == 0 ? r.y : r.w) >> 31) << cur;
        }
    }
}

struct BlockMasks {
    uint64_t a;    // bases that take a transition (A<->G, C<->T)
    uint64_t b;    // bases that take a transversion ...
    uint64_t ch;   // ... and its choice bit
};
IDL_HD int block_masks_count(const BlockMasks& m) { return popc64(m.a | m.b); }

IDL_HD BlockMasks block_masks(int kind, uint64_t seed, uint32_t seq_id, uint32_t variant, int block, int L, const uint32_t* nmask,
                              const uint32_t* T1, float slope1, const uint32_t* T2, float slope2) {
    BlockMasks m;
    m.a = m.b = m.ch = 0;
    uint64_t dummy;
    if (kind == KIND_TRANSITION || kind == KIND_BOTH) m.a = stream_mask<false>(seed, seq_id, variant, STREAM_TRANSITION, block, L, T1, slope1, dummy);
    if (kind == KIND_TRANSVERSION || kind == KIND_BOTH) m.b = stream_mask<true>(seed, seq_id, variant, STREAM_TRANSVERSION, block, L, T2, slope2, m.ch);
    // reset flags of the block: nmask bit 31-j of word w = base 32w + j
    const uint64_t nm = ((uint64_t)brev32(nmask[2 * block + 1]) << 32) | (uint64_t)brev32(nmask[2 * block]);
    m.b &= ~nm;             // N stays N
    m.a &= ~(nm | m.b);     // where both streams hit, the transversion decides (it only depends on the purine / pyrimidine class)
    return m;
}

// the edit at block-relative position i (a set bit of m.a | m.b)
IDL_HD uint32_t block_masks_entry(const BlockMasks& m, int block, const uint32_t* codes, int i) {
    const int pos = block * RNG_BLOCK + i;
    const uint32_t c = code_at(codes, pos);
    uint32_t val;
    if ((m.b >> i) & 1ull) {
        const uint32_t cb = (uint32_t)((m.ch >> i) & 1ull);
        val = (c & 1u) ? (cb << 1) : (1u | ((1u - cb) << 1));
    } else {
        val = c ^ 2u;
    }
    return ((uint32_t)pos << 3) | val;
}

// write the block's edits, position-sorted, to dst[0 .. block_masks_count(m))
IDL_HD void block_masks_write(const BlockMasks& m, int block, const uint32_t* codes, uint32_t* dst) {
    uint64_t u = m.a | m.b;
    while (u) {
        const int i = ffs64(u);
        u &= u - 1;
        *dst++ = block_masks_entry(m, block, codes, i);
    }
}

// Random_N (idelucs/utils.py:89-95): draw i of the variant -> position floor(w_i * L / 2^32)
IDL_HD U4 random_n_words(uint64_t seed, uint32_t seq_id, uint32_t variant, uint32_t call) {
    return philox4x32_10(call, 0u, seq_id, (variant << 2) | STREAM_RANDOM_N, (uint32_t)seed, (uint32_t)(seed >> 32));
}
IDL_HD uint32_t random_n_entry(uint32_t word, int L) { return (mulhi32(word, (uint32_t)L) << 3) | VAL_N; }

// ---------------------------------------------------------------------------------------
// correctly-rounded float division a/b given y = RN(1/b) (Markstein's theorem: with the
// correctly rounded reciprocal, one multiply and one fused residual correction give
// RN(a/b)); checked against IEEE division in tests/test_host_emul.py (0 mismatches in
// 6e8 trials over the operand ranges the kernels produce: integers < 2^24 for the
// frequency, normal floats for the standardisation).
// ---------------------------------------------------------------------------------------
IDL_HD float div_rn(float a, float b, float y) {
    const float q = a * y;
    const float r = fmaf(-q, b, a);
    return fmaf(r, y, q);
}

}  // namespace idl
