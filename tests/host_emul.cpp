// CPU replay of the per-thread routines of idelucs_b200/csrc/core.cuh (the exact code the
// CUDA kernels run) so that the index logic — packing, chunk counting, window extraction,
// delta ownership, RNG streams — can be checked against the oracle without a GPU.
// Built by tests/test_host_emul.py with g++; not part of the product library.
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

#include "../idelucs_b200/csrc/core.cuh"

using namespace idl;

extern "C" {

// pack one sequence; buffers sized (nchunks+1)*4 / *2 words. returns first bad (pos<<3|cls) or -1
long long emul_pack(const uint8_t* ascii, long long L, int strict, uint32_t* codes, uint32_t* nmask) {
    const long long nchunks = (L + 63) / 64;
    long long first_bad = -1;
    for (long long h = 0; h < nchunks * 2; ++h) {
        uint32_t mw = 0;
        for (int t = 0; t < 2; ++t) {
            const long long base = h * 32 + t * 16;
            long long avail = L - base;
            avail = avail < 0 ? 0 : (avail > 16 ? 16 : avail);
            uint8_t buf[16];
            for (int j = 0; j < 16; ++j) buf[j] = j < avail ? ascii[base + j] : (uint8_t)'N';
            uint32_t cw, m16;
            int bj, bc;
            pack16(buf, (int)avail, strict, &cw, &m16, &bj, &bc);
            codes[h * 2 + t] = cw;
            mw = (mw << 16) | m16;
            if (bj < 16 && first_bad < 0) first_bad = ((base + bj) << 3) | bc;
        }
        nmask[h] = mw;
    }
    for (int i = 0; i < 4; ++i) codes[nchunks * 4 + i] = 0;
    for (int i = 0; i < 2; ++i) nmask[nchunks * 2 + i] = 0xFFFFFFFFu;
    return first_bad;
}

}  // extern "C"

template <int K>
static int counts_k(const uint32_t* codes, const uint32_t* nmask, int L, int* counts) {
    int nv = 0;
    const int nhalf = ((L + 63) / 64) * 2;
    for (int h = 0; h < nhalf; ++h)
        nv += count_half<K>(codes, nmask, h, codes[2 * h], codes[2 * h + 1], [&](uint32_t kmer) { counts[kmer] += 1; });
    return nv;
}

extern "C" {

int emul_counts(const uint32_t* codes, const uint32_t* nmask, int L, int K, int* counts) {
    switch (K) {
        case 1: return counts_k<1>(codes, nmask, L, counts);
        case 2: return counts_k<2>(codes, nmask, L, counts);
        case 3: return counts_k<3>(codes, nmask, L, counts);
        case 4: return counts_k<4>(codes, nmask, L, counts);
        case 5: return counts_k<5>(codes, nmask, L, counts);
        case 6: return counts_k<6>(codes, nmask, L, counts);
    }
    return -1;
}

}  // extern "C"

static long long g_entry_deltas_bad = 0;   // entry_deltas (register form) vs apply_entry (callback form), every entry ever applied

template <int K>
static int apply_k(const uint32_t* codes, const uint32_t* nmask, int L, const std::vector<uint32_t>& list, int* counts) {
    int d = 0;
    for (int i = 0; i < (int)list.size(); ++i) {
        std::vector<std::pair<uint32_t, int>> a, b;
        const int di = apply_entry<K>(codes, nmask, L, list.data(), (int)list.size(), i, [&](uint32_t kmer, int dd) { counts[kmer] += dd; a.emplace_back(kmer, dd); });
        d += di;
        uint32_t reg[K];
        int dt = 0;
        const int cnt = entry_deltas<K>(codes, nmask, L, list.data(), (int)list.size(), i, reg, &dt);
        for (int t = 0; t < K; ++t) {
            if (reg[t] & 0x1000u) b.emplace_back(reg[t] & 0xFFFu, -1);
            if (reg[t] & 0x10000000u) b.emplace_back((reg[t] >> 16) & 0xFFFu, +1);
        }
        if (a != b || cnt != (int)a.size() || dt != di) ++g_entry_deltas_bad;
    }
    return d;
}

template <int K>
static int randn_k(const uint32_t* codes, const uint32_t* nmask, int L, const std::vector<uint32_t>& ent, int* counts) {
    int d = 0;
    for (int i = 0; i < (int)ent.size(); ++i)
        random_n_removals<K>(codes, nmask, L, ent.data(), (int)ent.size(), i, [&](uint32_t kmer) { counts[kmer] -= 1; --d; });
    return d;
}

extern "C" {

// Random_N through the kernels' specialised path (unsorted draws -> removed windows)
int emul_random_n(const uint32_t* codes, const uint32_t* nmask, int L, int K, unsigned long long seed, unsigned seq_id,
                  unsigned rng_id, int n_bp, int* counts) {
    std::vector<uint32_t> ent;
    if (L > 0)
        for (int call = 0; call * 4 < n_bp; ++call) {
            const U4 r = random_n_words(seed, seq_id, rng_id, (uint32_t)call);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
            for (int t = 0; t < 4; ++t)
                if (call * 4 + t < n_bp) ent.push_back(random_n_entry(w[t], L));
        }
    switch (K) {
        case 1: return randn_k<1>(codes, nmask, L, ent, counts);
        case 2: return randn_k<2>(codes, nmask, L, ent, counts);
        case 3: return randn_k<3>(codes, nmask, L, ent, counts);
        case 4: return randn_k<4>(codes, nmask, L, ent, counts);
        case 5: return randn_k<5>(codes, nmask, L, ent, counts);
        case 6: return randn_k<6>(codes, nmask, L, ent, counts);
    }
    return 0;
}

// build the variant's edit list exactly like the kernel and patch `counts` (clean histogram)
// in place; returns the change of the counted-window total.  n_list_out / list_out optional.
int emul_variant(const uint32_t* codes, const uint32_t* nmask, int L, int K, unsigned long long seed, unsigned seq_id,
                 unsigned rng_id, int kind, double p1, double p2, int n_bp, const uint32_t* explicit_list, int n_explicit,
                 int* counts, uint32_t* list_out, int* n_list_out) {
    std::vector<uint32_t> list;
    if (kind == KIND_EXPLICIT) {
        list.assign(explicit_list, explicit_list + n_explicit);
    } else if (kind == KIND_RANDOM_N) {
        if (L > 0) {
            for (int call = 0; call * 4 < n_bp; ++call) {
                const U4 r = random_n_words(seed, seq_id, rng_id, (uint32_t)call);
                const uint32_t w[4] = {r.x, r.y, r.z, r.w};
                for (int t = 0; t < 4; ++t)
                    if (call * 4 + t < n_bp) list.push_back(random_n_entry(w[t], L));
            }
            std::stable_sort(list.begin(), list.end());
        }
    } else if (kind != KIND_CLEAN) {
        uint32_t T1[RNG_BLOCK], T2[RNG_BLOCK];
        geometric_table(p1, T1);
        geometric_table(p2, T2);
        const int nblocks = (L + RNG_BLOCK - 1) / RNG_BLOCK;
        for (int b = 0; b < nblocks; ++b)
            block_edits(kind, seed, seq_id, rng_id, b, L, codes, nmask, T1, gap_slope(p1), T2, gap_slope(p2),
                        [&](uint32_t e) { list.push_back(e); });
    }
    if (n_list_out) {
        *n_list_out = (int)list.size();
        if (list_out) std::memcpy(list_out, list.data(), list.size() * sizeof(uint32_t));
    }
    switch (K) {
        case 1: return apply_k<1>(codes, nmask, L, list, counts);
        case 2: return apply_k<2>(codes, nmask, L, list, counts);
        case 3: return apply_k<3>(codes, nmask, L, list, counts);
        case 4: return apply_k<4>(codes, nmask, L, list, counts);
        case 5: return apply_k<5>(codes, nmask, L, list, counts);
        case 6: return apply_k<6>(codes, nmask, L, list, counts);
    }
    return 0;
}

// mask block generator vs the stream-walking one (block_edits): returns the number of blocks where they disagree
int emul_masks_vs_slow(const uint32_t* codes, const uint32_t* nmask, int L, unsigned long long seed, unsigned seq_id,
                       unsigned rng_id, int kind, double p1, double p2, int* max_per_block) {
    uint32_t T1[RNG_BLOCK], T2[RNG_BLOCK];
    geometric_table(p1, T1);
    geometric_table(p2, T2);
    int bad = 0;
    *max_per_block = 0;
    const int nblocks = (L + RNG_BLOCK - 1) / RNG_BLOCK;
    for (int b = 0; b < nblocks; ++b) {
        std::vector<uint32_t> slow;
        block_edits(kind, seed, seq_id, rng_id, b, L, codes, nmask, T1, gap_slope(p1), T2, gap_slope(p2),
                    [&](uint32_t e) { slow.push_back(e); });
        const BlockMasks m = block_masks(kind, seed, seq_id, rng_id, b, L, nmask, T1, gap_slope(p1), T2, gap_slope(p2));
        uint32_t out[RNG_BLOCK];
        const int cnt = block_masks_count(m);
        block_masks_write(m, b, codes, out);
        if (cnt > *max_per_block) *max_per_block = cnt;
        if (cnt != (int)slow.size() || !std::equal(slow.begin(), slow.end(), out)) ++bad;
    }
    return bad;
}

void emul_geometric_table(double p, uint32_t* out) { geometric_table(p, out); }
long long emul_entry_deltas_mismatches() { return g_entry_deltas_bad; }

void emul_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    const U4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// number of mismatches of div_rn(a, b, RN(1/b)) against IEEE a/b
long long emul_div_check(const float* a, const float* b, long long n) {
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        const float y = 1.0f / b[i];
        const float q = div_rn(a[i], b[i], y), ref = a[i] / b[i];
        if (std::memcmp(&q, &ref, 4) != 0) ++bad;
    }
    return bad;
}

}  // extern "C"
