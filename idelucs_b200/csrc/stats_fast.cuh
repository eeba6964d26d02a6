// idelucs_b200 — scaler statistics of slot 0 (t_norm = transition_transversion(1e-2, 0.5e-2) in
// AugmentFasta, idelucs/utils.py:330, 354-359; clean profiles in SequenceDataset, utils.py:401-405), k = 6.
//
// Column (rows, mean, M2) of float32(count/total) over the sequences; nothing is written per row.
// One persistent CTA of 1024 threads per SM.  Thread t owns bins 4t..4t+3 and keeps their float64
// shifted-data sums in 20 registers (the generic kernel's 2 CTAs x 512 threads x 8 bins do not fit
// 64 registers and spill the accumulators).  Four sequences are in flight per CTA, one per stage,
// and the stages of one iteration are separated from the next iteration by ONE barrier:
//
//   S0 (item j+3)  count: thread <-> 16-base code word, operands prefetched into registers one
//                  iteration earlier (lengths / offsets two iterations earlier: no global-memory
//                  latency on the critical path); predicated shared atomics into histogram (j+3)&3;
//                  the words are also staged in shared memory for S1/S2
//   S1 (items j+2 .. j+7, every 6th iteration)  Bernoulli edits of SIX sequences at once: thread <-> 64-base
//                  block, register-only generator (fast_block), written position-sorted into the block's
//                  12-entry slot of an 8-item ring (no CTA-wide scan).  One sequence at a time this stage is
//                  5 busy warps and pure latency; six together fill the SM
//   S2 (item j+1)  thread <-> EDIT: warps 16..31 each list the edits of their 16 blocks (warp scan of the
//                  per-block counts) and apply them as +-1 shared atomics (apply_entry on a view that stitches
//                  the previous block's slot, this slot and the next block's first entry)
//   S3 (item j)    fold: 4 bins per thread -> float32 frequency -> float64 sums; bins zeroed
//
// so the LSU/atomic work, the integer RNG work and the FP32/FP64 work of different sequences overlap.
// Items are assigned statically (item = blockIdx.x + j*gridDim.x): the sums are reproducible.
// Items this kernel cannot take (longer than SF_MAXL bases, a block with more than FAST_CAP hits
// per stream) are flagged in d_status (bit 1), left out of the sums and folded by the generic kernel
// into further parts (static assignment as well).  Included by kernels.cu (uses its helpers).
#pragma once

namespace idl {

constexpr int SF_NT = 1024;
constexpr int SF_MAXL = 16320;            // 255 blocks: 1020 code words + 4 slack words <= 1024 threads
constexpr int SF_BLOCKS = 256;
constexpr int SF_SLOT = 2 * FAST_CAP;     // edits per 64-base block the fast generator can emit
constexpr int SF_BURST = 6;               // items whose edits are generated together (one block per thread, every SF_BURST-th iteration)
constexpr int SF_RING = 8;                // edit-slot ring (items)

struct SfSmem {
    int hist[4][4096];
    uint32_t codes[4][1024 + 8];
    uint32_t mask[4][512 + 8];
    uint32_t slots[SF_RING][SF_BLOCKS * SF_SLOT];
    int cnt[SF_RING][SF_BLOCKS];   // edits per block; -1: the fast generator overflowed (item deferred)
    unsigned short wl[SF_BLOCKS / 16][16 * SF_SLOT];   // S2: (block << 4 | entry) of a warp's 16 blocks
    int nvalid[8];
    int defer[8];
    uint32_t gtab[2][RNG_BLOCK];
};

struct SfUnit { uint32_t w, prev, m, mprev; };

// unit u = the u-th 16-base code word of the sequence at (codes, nmask)
__device__ __forceinline__ SfUnit sf_load_unit(const uint32_t* __restrict__ codes, const uint32_t* __restrict__ nmask, int u) {
    SfUnit x;
    x.w = __ldg(codes + u);
    x.prev = u > 0 ? __ldg(codes + u - 1) : 0u;
    x.m = __ldg(nmask + (u >> 1));
    x.mprev = u > 1 ? __ldg(nmask + (u >> 1) - 1) : 0xFFFFFFFFu;   // sequence start = preceded by resets (kmers.pyx:14)
    return x;
}

// counts the windows ending in unit u (kmers.pyx:36-47); returns how many were counted
template <int K>
__device__ __forceinline__ int sf_count_unit(const SfUnit& x, int u, int* hist) {
    constexpr uint32_t KMASK = (1u << (2 * K)) - 1u;
    const uint64_t flags = ((uint64_t)x.mprev << 32) | (uint64_t)x.m;   // bit 31-j of m = reset flag of base j of the half-chunk
    uint64_t inv64 = flags;
#pragma unroll
    for (int d = 1; d < K; ++d) inv64 |= flags >> d;                    // a window is void if any of its K bases is a reset
    const uint32_t inv = (uint32_t)(inv64 >> ((u & 1) ? 0 : 16)) & 0xFFFFu;   // bit 15-j = window ending at base j of this word
#pragma unroll
    for (int j = 0; j < 16; ++j)
        if (!((inv >> (15 - j)) & 1u)) atomicAdd(&hist[funnel_r(x.w, x.prev, 30 - 2 * j) & KMASK], 1);
    return __popc(~inv & 0xFFFFu);
}

// position-sorted edit list seen from block b: the previous block's slot, this block's slot, the
// first entry of the next block (edits further away cannot share a window with this block's)
struct SfList {
    const uint32_t* a; int na;
    const uint32_t* b; int nb;
    const uint32_t* c;
    __device__ __forceinline__ uint32_t operator[](int j) const { return j < na ? a[j] : (j < na + nb ? b[j - na] : c[0]); }
};

struct SfMeta { int L; long long c0; uint32_t seq; };

template <int K>
__global__ void __launch_bounds__(SF_NT, 1) stats_fast_kernel(const ProfParams p) {
    constexpr int F = 1 << (2 * K);
    static_assert(F == 4 * SF_NT, "one int4 of bins per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SfSmem& sm = *reinterpret_cast<SfSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const long long G = gridDim.x;
    const long long n_items = p.n_items;
    const VarDesc vd = p.vars[0];
    const bool bern = (vd.kind == KIND_TRANSITION || vd.kind == KIND_TRANSVERSION || vd.kind == KIND_BOTH) && !(p.dbg & 0x100);   // development switches (IDL_PC_DBG bits 8..11): stages off for timing, results are then wrong
    const bool dbg_s2 = !(p.dbg & 0x800);

#pragma unroll
    for (int r = 0; r < 4; ++r) reinterpret_cast<int4*>(sm.hist[r])[tid] = make_int4(0, 0, 0, 0);
    if (tid < 8) { sm.nvalid[tid] = 0; sm.defer[tid] = 0; }
    if (tid < 2 * RNG_BLOCK) sm.gtab[tid / RNG_BLOCK][tid % RNG_BLOCK] = bern ? p.gtab[(tid < RNG_BLOCK ? vd.tab1 : vd.tab2) * RNG_BLOCK + tid % RNG_BLOCK] : 0u;

    auto item_meta = [&](long long j) {   // j-th item of this CTA
        SfMeta m; m.L = -1; m.c0 = 0; m.seq = 0u;   // L < 0: no such item
        const long long it = (long long)blockIdx.x + j * G;
        if (it < n_items) {
            const long long seq = p.sidx ? (long long)__ldg(p.sidx + it) : it;
            m.L = __ldg(p.len + seq);
            m.c0 = __ldg(reinterpret_cast<const long long*>(p.chunk_off) + seq);
            m.seq = (uint32_t)(p.seq_id0 + seq);
        }
        return m;
    };
    auto units_of = [](int L) { return (L > 0 && L <= SF_MAXL) ? ((L + 63) >> 6) << 2 : 0; };   // chunk-padded code words; 0: nothing to count / deferred
    auto prefetch = [&](const SfMeta& m, SfUnit& u) {
        if (tid < units_of(m.L)) u = sf_load_unit(p.codes + m.c0 * 4, p.nmask + m.c0 * 2, tid);
    };

    // metadata pipeline: m0 = item of S3 (fold) ... m3 = item of S0 (count), ma / mb = the two after it
    SfMeta m0, m1, m2, m3, ma, mb;
    m0.L = m1.L = m2.L = -1; m0.c0 = m1.c0 = m2.c0 = 0; m0.seq = m1.seq = m2.seq = 0u;
    m3 = item_meta(0);
    ma = item_meta(1);
    mb = item_meta(2);
    SfUnit ua;
    ua.w = ua.prev = ua.m = 0u; ua.mprev = 0xFFFFFFFFu;
    prefetch(m3, ua);
    __syncthreads();

    double acc1[4] = {0.0, 0.0, 0.0, 0.0}, acc2[4] = {0.0, 0.0, 0.0, 0.0};
    float shiftK[4] = {0.f, 0.f, 0.f, 0.f};
    int n_acc = 0;
    const int pc = p.pseudocount;
    // iteration n folds the CTA's n-th item; n = -3, -2, -1 only fill the pipeline
    for (long long n = -3; m0.L >= 0 || m1.L >= 0 || m2.L >= 0 || m3.L >= 0; ++n) {
        // ---- S0: count item n+3 from the prefetched registers, stage its words ----
        {
            const int r = (int)(n + 3) & 3;
            const int nu = units_of(m3.L);
            if (nu > 0) {
                int nv = 0;
                if (tid < nu) {
                    if (!(p.dbg & 0x200)) nv = sf_count_unit<K>(ua, tid, sm.hist[r]);
                    sm.codes[r][tid] = ua.w;
                    if (!(tid & 1)) sm.mask[r][tid >> 1] = ua.m;
                } else if (tid < nu + 4) {   // slack behind the sequence (window reads run one word past the end)
                    sm.codes[r][tid] = 0u;
                    if (tid < nu + 2) sm.mask[r][(nu >> 1) + tid - nu] = 0xFFFFFFFFu;
                }
                nv = warp_sum(nv);
                if (lane == 0 && nv) atomicAdd(&sm.nvalid[(int)(n + 3) & 7], nv);
            }
            if (tid == 0) { sm.nvalid[(int)(n + 4) & 7] = 0; sm.defer[(int)(n + 4) & 7] = 0; }
        }
        // ---- prefetch item n+4, metadata of item n+6 ----
        const SfMeta m4 = ma;
        prefetch(m4, ua);
        ma = mb;
        mb = item_meta(n + 6);
        // ---- S1 (every SF_BURST-th iteration): Bernoulli edits of items n+2 .. n+2+SF_BURST-1, one thread per 64-base
        // block.  A block is a ~1000-instruction dependent chain: done for one item at a time it is pure latency
        // (5 warps busy, 8 k cycles per item); six items' blocks together fill the SM.  N flags and bases of the hit
        // positions come straight from global memory (the sequences of these items are not staged yet).
        if (bern && (n + 3) % SF_BURST == 0) {
            int pre[SF_BURST + 1];
            pre[0] = 0;
#pragma unroll
            for (int i = 0; i < SF_BURST; ++i) {
                const long long it = (long long)blockIdx.x + (n + 2 + i) * G;
                int nblk = 0;
                if (n + 2 + i >= 0 && it < n_items) {
                    const int L = __ldg(p.len + (p.sidx ? (long long)__ldg(p.sidx + it) : it));
                    if (L > 0 && L <= SF_MAXL) nblk = (L + RNG_BLOCK - 1) / RNG_BLOCK;
                }
                pre[i + 1] = pre[i] + nblk;
            }
            for (int t = SF_NT - 1 - tid; t < pre[SF_BURST]; t += SF_NT) {
                int i = 0;
#pragma unroll
                for (int q = 1; q < SF_BURST; ++q) i += (t >= pre[q]) ? 1 : 0;
                int b = t;
#pragma unroll
                for (int q = 1; q < SF_BURST; ++q) b = (i == q) ? t - pre[q] : b;
                const long long it = (long long)blockIdx.x + (n + 2 + i) * G;
                const long long seq = p.sidx ? (long long)__ldg(p.sidx + it) : it;
                const int L = __ldg(p.len + seq);
                const long long c0 = __ldg(reinterpret_cast<const long long*>(p.chunk_off) + seq);
                const int e = (int)(n + 2 + i) & (SF_RING - 1);
                const FastBlock f = fast_block(vd.kind, p.seed, (uint32_t)(p.seq_id0 + seq), (uint32_t)vd.rng_id, b, L, p.nmask + c0 * 2,
                                               sm.gtab[0], vd.slope1, sm.gtab[1], vd.slope2);
                if (f.ok && f.cnt) fast_block_write(f, b, p.codes + c0 * 4, sm.slots[e] + b * SF_SLOT);
                sm.cnt[e][b] = f.ok ? f.cnt : -1;
            }
        }
        // ---- S2: histogram deltas of item n+1, one thread per EDIT: warps 16..31 each own 16 blocks, list their edits
        // (warp scan of the per-block counts, no CTA barrier) and share them out lane by lane (~15 edits per warp at
        // the reference's rates: one round) ----
        if (bern && dbg_s2 && m1.L > 0 && m1.L <= SF_MAXL && tid >= SF_NT / 2) {
            const int r = (int)(n + 1) & 3, e = (int)(n + 1) & (SF_RING - 1);
            const int nblk = (m1.L + RNG_BLOCK - 1) / RNG_BLOCK;
            const int wi = (tid >> 5) - 16;
            const int b = wi * 16 + lane;
            int c = (lane < 16 && b < nblk) ? sm.cnt[e][b] : 0;
            if (c < 0) { sm.defer[(int)(n + 1) & 7] = 1; c = 0; }
            int incl = c;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int E = __shfl_sync(0xffffffffu, incl, 15);
            unsigned short* wl = sm.wl[wi];
            for (int i = 0; i < c; ++i) wl[incl - c + i] = (unsigned short)((b << 4) | i);
            __syncwarp();
            int* hist = sm.hist[r];
            int d = 0;
            for (int t = lane; t < E; t += 32) {
                const int x = wl[t];
                const int bb = x >> 4, i = x & 15;
                SfList lst;
                lst.na = bb > 0 ? max(sm.cnt[e][bb - 1], 0) : 0;
                lst.a = sm.slots[e] + (bb - 1) * SF_SLOT;
                lst.nb = sm.cnt[e][bb];
                lst.b = sm.slots[e] + bb * SF_SLOT;
                lst.c = sm.slots[e] + (bb + 1) * SF_SLOT;
                const int nn = lst.na + lst.nb + ((bb + 1 < nblk && sm.cnt[e][bb + 1] > 0) ? 1 : 0);
                d += apply_entry<K>(sm.codes[r], sm.mask[r], m1.L, lst, nn, lst.na + i, [&](uint32_t kmer, int dd) { atomicAdd(&hist[kmer], dd); });
            }
            if (d) atomicAdd(&sm.nvalid[(int)(n + 1) & 7], d);
        }
        // ---- S3: fold item n ----
        if (m0.L >= 0) {
            const int r = (int)n & 3;
            const int4 h = reinterpret_cast<const int4*>(sm.hist[r])[tid];
            reinterpret_cast<int4*>(sm.hist[r])[tid] = make_int4(0, 0, 0, 0);
            const bool deferred = m0.L > SF_MAXL || sm.defer[(int)n & 7] != 0;
            if (deferred) {
                if (tid == 0) {
                    atomicOr(p.status + ((long long)blockIdx.x + n * G), 2);
                    atomicAdd(p.work_counter + 1, 1ull);
                }
            } else if (!(p.dbg & 0x400)) {
                const int total = F * pc + sm.nvalid[(int)n & 7];
                const float ftot = (float)total;
                const float y = 1.0f / ftot;
                const bool big = total >= (1 << 24);
                const int ci[4] = {h.x + pc, h.y + pc, h.z + pc, h.w + pc};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float q = big ? (float)((double)ci[e] / (double)total) : div_rn((float)ci[e], ftot, y);
                    if (n_acc == 0) shiftK[e] = q;
                    const double dd = (double)q - (double)shiftK[e];
                    acc1[e] += dd;
                    acc2[e] = fma(dd, dd, acc2[e]);
                }
                ++n_acc;
            }
        }
        m0 = m1; m1 = m2; m2 = m3; m3 = m4;
        __syncthreads();
    }
    // this CTA's part: (rows, mean, M2) per column, merged by scaler_finalize_kernel
    if (tid == 0) p.stats_n[blockIdx.x] = (double)n_acc;
    const double m = n_acc > 0 ? (double)n_acc : 1.0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        p.stats_partials[((size_t)blockIdx.x * 2 + 0) * F + tid * 4 + e] = (double)shiftK[e] + acc1[e] / m;
        p.stats_partials[((size_t)blockIdx.x * 2 + 1) * F + tid * 4 + e] = acc2[e] - acc1[e] * acc1[e] / m;
    }
}

}  // namespace idl
