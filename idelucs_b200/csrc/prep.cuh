// idelucs_b200 — "prepare" pass of the whole-schedule featurisation (k = 6): everything the AugmentFasta
// schedule (idelucs/utils.py:330-351) needs from the Bernoulli mimics (transition / transversion / combined),
// computed ONCE per sequence by a high-occupancy kernel and handed to the consumers through a caller-owned
// scratch buffer:
//
//   * the +-1 histogram deltas of every Bernoulli ("dense") slot, as a packed list
//     (kmer | ordinal << 12 | (add ? 0x8000 : 0)) — the producer/consumer kernel (profiles_pc.cuh) TMA-loads the
//     list instead of generating the mutations itself (its producers were its bottleneck: 26 k of 41 k cycles
//     per sequence went into the Bernoulli slots);
//   * the uint16 histogram of slot 0 (t_norm, the 'true' side of every pair) and its window total — the
//     StandardScaler statistics of AugmentFasta (utils.py:354-359) are column sums over these rows
//     (colstats16_kernel: HBM-bound read of 8 KB per sequence, float64 shifted-data sums in registers),
//     so the slot-0 mutations are generated once for the statistics AND the profile pass.
//
// One CTA of 256 threads per sequence at a time, 4-6 CTAs per SM (35 KB of shared memory each): count the clean
// histogram into packed uint16 counters (shared atomics), mutation masks of every (slot, 64-base block) by
// the register-only mask generator (core.cuh), CTA scan -> position-sorted edit lists -> +-1 deltas
// (apply_entry), slot 0's deltas applied to the histogram.  Sequences that do not fit (longer than
// 20 480 bases, list overflows at rates far above the reference's) are flagged (meta.flags, d_status bit 1)
// and taken by the generic kernel.  Included by kernels.cu.
#pragma once

namespace idl {

constexpr int PR_NT = 256;
constexpr int PR_LIST = 2048;              // edits of all dense slots of one sequence

struct PrSmem {
    alignas(16) uint32_t hist[PC_F / 2];   // packed uint16 pairs: clean histogram, then slot 0's
    alignas(16) uint32_t sseq[PC_SSEQ_W];  // staged sequence: codes | mask
    alignas(16) uint32_t list[PR_LIST + 8];
    alignas(16) uint16_t delta[PC_DELTA];
    uint32_t gtabs[STABS][RNG_BLOCK];
    VarDesc dvar[PC_DENSE];
    int seg_off[PC_DENSE + 1];
    int scan[PR_NT / 32 + 2];
    int dtot[PC_DENSE];
    int nvalid, n_delta;
    int nd, slot0_dense;
    long long next_item;
};

template <int K>
__global__ void __launch_bounds__(PR_NT, 4) prep_kernel(const ProfParams p, const __grid_constant__ Plan plan) {
    static_assert(K == PC_K, "the prepare pass is specialised for k = 6");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PrSmem& sm = *reinterpret_cast<PrSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const long long n = p.n_items;
    unsigned char* prep = reinterpret_cast<unsigned char*>(p.prep);
    PrepMeta* meta = reinterpret_cast<PrepMeta*>(prep + prep_meta_off());
    uint4* hist_out = reinterpret_cast<uint4*>(prep + prep_hist_off(n));
    uint4* delta_out = reinterpret_cast<uint4*>(prep + prep_delta_off(n));
    const VarDesc* gvars = p.inline_plan ? plan.vars : p.vars;

    if (tid == 0) {
        int nd = 0;
        for (int s = 0; s < p.n_vars; ++s) {
            const VarDesc vd = gvars[s];
            if (vd.kind == KIND_TRANSITION || vd.kind == KIND_TRANSVERSION || vd.kind == KIND_BOTH) {
                if (nd < PC_DENSE) sm.dvar[nd] = vd;
                ++nd;
            }
        }
        sm.nd = nd <= PC_DENSE ? nd : PC_DENSE;   // (the host refuses more dense slots than PC_DENSE)
        const int k0 = gvars[0].kind;
        sm.slot0_dense = (k0 == KIND_TRANSITION || k0 == KIND_TRANSVERSION || k0 == KIND_BOTH) ? 1 : 0;
        if (blockIdx.x == 0) {
            PrepHeader h;
            h.stamp = p.prep_stamp; h.n_items = n; h.n_dense = sm.nd; h.slot0_dense = sm.slot0_dense;
            for (int i = 0; i < 10; ++i) h.pad[i] = 0;
            *reinterpret_cast<PrepHeader*>(prep) = h;
        }
        sm.next_item = (long long)atomicAdd(p.work_counter, 1ull);
    }
    for (int i = tid; i < STABS * RNG_BLOCK; i += PR_NT)
        (&sm.gtabs[0][0])[i] = i < p.n_tabs * RNG_BLOCK ? (p.inline_plan ? (&plan.gtab[0][0])[i] : p.gtab[i]) : 0u;
    auto table = [&](int t) -> const uint32_t* { return t < STABS ? sm.gtabs[t] : p.gtab + t * RNG_BLOCK; };

    for (;;) {
        __syncthreads();   // the previous item is finished; thread 0's next_item is visible
        const long long item = sm.next_item;
        const int nd = sm.nd;
        const bool slot0_dense = sm.slot0_dense != 0;
        __syncthreads();   // everybody has read it
        if (item >= n) break;
        if (tid == 0) {
            sm.next_item = (long long)atomicAdd(p.work_counter, 1ull);   // in flight while this item is processed
            sm.nvalid = 0; sm.n_delta = 0;
            for (int j = 0; j < PC_DENSE; ++j) sm.dtot[j] = 0;
        }
        const long long seq = p.sidx ? (long long)__ldg(p.sidx + item) : item;
        const int L = __ldg(p.len + seq);
        const long long c0 = __ldg(reinterpret_cast<const long long*>(p.chunk_off) + seq);
        const uint32_t seq_id = (uint32_t)(p.seq_id0 + seq);
        const int nchunks = (L + CHUNK_BASES - 1) / CHUNK_BASES;
        bool defer = nchunks > SSEQ_CHUNKS;
        if (!defer) {
            // ---- stage the packed sequence, clear the histogram ----
            uint32_t* codes = sm.sseq;
            uint32_t* nmask = sm.sseq + SSEQ_CW;
            const uint4* gc = reinterpret_cast<const uint4*>(p.codes + c0 * 4);
            const uint2* gm = reinterpret_cast<const uint2*>(p.nmask + c0 * 2);
            for (int c = tid; c < nchunks; c += PR_NT) {
                reinterpret_cast<uint4*>(codes)[c] = __ldg(gc + c);
                reinterpret_cast<uint2*>(nmask)[c] = __ldg(gm + c);
            }
            if (tid < 4) codes[nchunks * 4 + tid] = 0u;           // slack chunk (window reads run one word past the end)
            if (tid < 2) nmask[nchunks * 2 + tid] = 0xFFFFFFFFu;
            for (int i = tid; i < PC_F / 8; i += PR_NT) reinterpret_cast<uint4*>(sm.hist)[i] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
            // ---- clean histogram (idelucs/kmers.pyx:38-50), two uint16 counters per word ----
            {
                int nv = 0;
                for (int h = tid; h < nchunks * 2; h += PR_NT) {
                    const uint2 w = reinterpret_cast<const uint2*>(codes)[h];
                    nv += count_half<K>(codes, nmask, h, w.x, w.y, [&](uint32_t kmer) { atomicAdd(&sm.hist[kmer >> 1], 1u << ((kmer & 1u) * 16u)); });
                }
                nv = warp_sum(nv);
                if (lane == 0 && nv) atomicAdd(&sm.nvalid, nv);
            }
            // ---- mutation masks of every (dense slot, block), position-sorted edit lists (slot-major) ----
            const int nblocks = nchunks, W = nd * nblocks;
            int base = 0;
            for (int w0 = 0; w0 < W; w0 += PR_NT) {   // uniform trip count (CTA scan inside)
                const int w = w0 + tid;
                const bool active = w < W;
                const int j = active ? w / nblocks : 0, b = active ? w - j * nblocks : 0;
                BlockMasks m;
                m.a = m.b = m.ch = 0;
                if (active) {
                    const VarDesc vd = sm.dvar[j];
                    m = block_masks(vd.kind, p.seed, seq_id, (uint32_t)vd.rng_id, b, L, nmask, table(vd.tab1), vd.slope1, table(vd.tab2), vd.slope2);
                }
                int total;
                const int off = block_exscan<PR_NT>(block_masks_count(m), sm.scan, &total);
                if (base + total > PR_LIST) { defer = true; break; }   // uniform
                if (active && b == 0) sm.seg_off[j] = base + off;
                block_masks_write(m, b, codes, sm.list + base + off);
                base += total;
            }
            if (tid == 0) sm.seg_off[nd] = base;
            __syncthreads();   // lists, seg_off and the clean histogram are complete
            if (!defer) {
                // ---- edits -> +-1 deltas (thread <-> edit); slot 0's deltas also go into the histogram ----
                for (int i0 = 0; i0 < base; i0 += PR_NT) {   // uniform trip count (warp collectives inside)
                    const int i = i0 + tid;
                    int jj = 0, so = 0, cnt = 0;
                    if (i < base) {
                        while (i >= sm.seg_off[jj + 1]) ++jj;
                        so = sm.seg_off[jj];
                        apply_entry<K>(codes, nmask, L, sm.list + so, sm.seg_off[jj + 1] - so, i - so, [&](uint32_t, int) { ++cnt; });
                    }
                    int inc = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += t;
                    }
                    const int wtot = __shfl_sync(0xffffffffu, inc, 31);
                    int wbase = 0;
                    if (lane == 31 && wtot) wbase = atomicAdd(&sm.n_delta, wtot);
                    wbase = __shfl_sync(0xffffffffu, wbase, 31);
                    int slot = wbase + inc - cnt;
                    if (cnt && wbase + wtot <= PC_DELTA) {
                        const uint32_t tag = (uint32_t)jj << 12;
                        const bool to_hist = slot0_dense && jj == 0;
                        const int d = apply_entry<K>(codes, nmask, L, sm.list + so, sm.seg_off[jj + 1] - so, i - so, [&](uint32_t kmer, int dd) {
                            sm.delta[slot++] = (uint16_t)(kmer | tag | (dd > 0 ? 0x8000u : 0u));
                            if (to_hist) upd16(sm.hist, kmer, dd);
                        });
                        if (d) atomicAdd(&sm.dtot[jj], d);
                    }
                }
                __syncthreads();
                if (sm.n_delta > PC_DELTA) defer = true;   // uniform
            }
        }
        if (defer) {
            if (tid == 0) {
                PrepMeta mt;
                mt.n_delta = 0; mt.total0 = 0; mt.base_total = 0; mt.flags = 1; mt.pad = 0;
                for (int j = 0; j < PC_DENSE; ++j) mt.dtot[j] = 0;
                meta[item] = mt;
                if (p.status) atomicOr(p.status + item, 2);
                atomicAdd(p.work_counter + 1, 1ull);
            }
            continue;
        }
        // ---- hand over: slot 0's histogram, the delta list, the totals ----
        const int n_delta = sm.n_delta;
        for (int i = tid; i < PC_F / 8; i += PR_NT) hist_out[(size_t)item * (PC_F / 8) + i] = reinterpret_cast<const uint4*>(sm.hist)[i];
        const int nq = (n_delta * 2 + 15) >> 4;
        for (int i = tid; i < nq; i += PR_NT) delta_out[(size_t)item * (PC_DELTA / 8) + i] = reinterpret_cast<const uint4*>(sm.delta)[i];
        if (tid == 0) {
            PrepMeta mt;
            mt.n_delta = n_delta;
            mt.base_total = PC_F * p.pseudocount + sm.nvalid;
            mt.total0 = mt.base_total + (slot0_dense ? sm.dtot[0] : 0);
            mt.flags = 0; mt.pad = 0;
            for (int j = 0; j < PC_DENSE; ++j) mt.dtot[j] = sm.dtot[j];
            meta[item] = mt;
        }
    }
}

// StandardScaler statistics (idelucs/utils.py:354-359) of the prepared slot-0 rows: float32(count / total) per bin
// (the same correctly rounded division as the profile kernels), float64 shifted-data sums in registers.  Thread <-> 4
// bins, CTA <-> 1024 bins x one block of rows; the rows of a part are visited in ascending order and the parts merged in
// index order (idl_scaler_finalize), so the result is run-to-run identical.  HBM-bound: 8 KB read per sequence.
constexpr int CS16_NT = 256;

__global__ void __launch_bounds__(CS16_NT) colstats16_kernel(const unsigned char* __restrict__ prep, long long n, int rows_per_part, int pseudocount,
                                                              double* __restrict__ partials, double* __restrict__ part_n) {
    const int vec = blockIdx.x * CS16_NT + threadIdx.x;   // 4-bin granule
    const long long r0 = (long long)blockIdx.y * rows_per_part;
    const long long r1 = r0 + rows_per_part < n ? r0 + rows_per_part : n;
    const int4* meta = reinterpret_cast<const int4*>(prep + prep_meta_off());   // first half of PrepMeta: n_delta, total0, base_total, flags
    const uint2* hist = reinterpret_cast<const uint2*>(prep + prep_hist_off(n));
    double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
    float shift[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
    constexpr int U = 4;   // rows in flight per thread
    for (long long r = r0; r < r1; r += U) {
        uint2 pk[U];
        int4 mt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u < r1 ? r + u : r1 - 1;
            mt[u] = __ldg(meta + rr * 2);
            pk[u] = __ldg(hist + rr * (PC_F / 4) + vec);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u >= r1 || mt[u].w != 0) continue;   // uniform over the CTA
            const float ftot = (float)mt[u].y;
            const float y = 1.0f / ftot;
            const float c[4] = {(float)((int)(pk[u].x & 0xFFFFu) + pseudocount), (float)((int)(pk[u].x >> 16) + pseudocount),
                                (float)((int)(pk[u].y & 0xFFFFu) + pseudocount), (float)((int)(pk[u].y >> 16) + pseudocount)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float q = div_rn(c[e], ftot, y);
                if (cnt == 0) shift[e] = q;
                const double dd = (double)q - (double)shift[e];
                a1[e] += dd;
                a2[e] = fma(dd, dd, a2[e]);
            }
            ++cnt;
        }
    }
    if (vec == 0) part_n[blockIdx.y] = (double)cnt;
    const double m = cnt > 0 ? (double)cnt : 1.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        partials[((size_t)blockIdx.y * 2 + 0) * PC_F + vec * 4 + e] = (double)shift[e] + a1[e] / m;
        partials[((size_t)blockIdx.y * 2 + 1) * PC_F + vec * 4 + e] = a2[e] - a1[e] * a1[e] / m;
    }
}

}  // namespace idl
