// idelucs_b200 — producer/consumer variant of the profiles kernel (k = 6, float outputs).
//
// One CTA of 1024 threads per SM.  Warps 16..31 (PRODUCERS) prepare sequence i+1 — count the
// clean histogram, pack it to uint16, precompute the Random_N removal lists of every slot and
// the +-1 delta list of the Bernoulli slots — into one of two shared-memory contexts while
// warps 0..15 (CONSUMERS) stream the 51 profiles of sequence i from the other context:
// 8 private uint16 copies in two sets of 4; the consumers subtract the removals of the next
// half-group with fire-and-forget shared atomics, stream the current half-group (restoring
// its copies on the way) and meet at a consumer-only named barrier once per 4 variants.
// Producers and consumers meet at ONE CTA-wide barrier per sequence.  The counting, RNG and
// delta work therefore overlaps with the HBM-bound streaming instead of alternating with it.
//
// Items this kernel cannot take (longer than 20 480 bases, > FAST_CAP hits in a 64-base block,
// delta list overflow) are flagged in d_status (bit 1) and redone by the generic kernel.
// Included by kernels.cu (uses its helpers).
#pragma once

namespace idl {

constexpr int PC_NT = 1024, PC_HALF = 512;
constexpr int PC_K = 6, PC_F = 4096, PC_VEC = 1024, PC_VPT = 2, PC_PRIVW = 2048;
constexpr int PC_G = 8, PC_HG = 4;
constexpr int PC_MAXS = 64;        // variant slots per sequence
constexpr int PC_REM = 6144;       // removed k-mers of all Random_N slots of one sequence (51 x 20 x 6 = 6120)
constexpr int PC_DELTA = 6144;     // Bernoulli +-1 deltas of one sequence
constexpr int PC_MAXB = 8;         // Bernoulli slots per sequence

struct alignas(16) PcCtx {
    uint2 clean16[PC_VEC];         // packed clean histogram
    uint16_t rem[PC_REM];          // Random_N: removed k-mers, slot s at [rem_off[s]*K, +nbp*K), 0xFFFF = unused
    uint16_t delta[PC_DELTA];      // Bernoulli: kmer | ordinal << 12 | (add ? 0x8000 : 0)
    int dtot[PC_MAXS];             // change of the counted-window total per slot
    int n_delta;
    int defer;                     // item must be redone by the generic kernel
    int base_total;
    long long item;                // -1: no more work
};

struct PcSmem {
    PcCtx ctx[2];
    // producer scratch
    alignas(16) int hist[PC_F];
    alignas(16) uint32_t sseq[SSEQ_CW + SSEQ_MW];
    alignas(16) uint32_t list[LIST_CAP + 8];
    uint32_t gtabs[STABS][RNG_BLOCK];
    int scan[PC_HALF / 32 + 2];
    int seg_off[PC_MAXB + 1];
    int nvalid;
    int any_over;
    long long next_item;
    // launch-uniform slot tables (built once)
    VarDesc svars[PC_MAXS];
    long long sout_off[PC_MAXS];
    int kind_class[PC_MAXS];       // 0 none, 1 Random_N, 2 Bernoulli
    int rem_off[PC_MAXS];          // first draw of the slot (x K = first rem entry)
    int nbp[PC_MAXS];
    int bern_ord[PC_MAXS];         // ordinal of a Bernoulli slot
    int bern_slot[PC_MAXB];        // slot of ordinal
    int n_bern, n_ent;
    // consumer side
    alignas(16) uint32_t priv[PC_G * PC_PRIVW];
    float2 gy[PC_MAXS];
    long long grow[PC_MAXS];
};

__device__ __forceinline__ void bar_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// exclusive scan over the 512 producer threads (named barrier 1)
__device__ __forceinline__ int pc_exscan(int v, int* scratch, int* total, int ptid) {
    const int lane = ptid & 31, wid = ptid >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scratch[wid] = inc;
    bar_named(1, PC_HALF);
    if (wid == 0) {
        int w = lane < PC_HALF / 32 ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < PC_HALF / 32) scratch[lane] = winc - w;
        if (lane == 31) scratch[PC_HALF / 32] = winc;
    }
    bar_named(1, PC_HALF);
    *total = scratch[PC_HALF / 32];
    return scratch[wid] + inc - v;
}

template <int OUT>
__device__ __noinline__ void pc_producer(PcSmem& sm, const ProfParams& p) {
    constexpr int K = PC_K;
    const int ptid = threadIdx.x - PC_HALF, lane = ptid & 31;
    int b = 0;
    long long t_prev = p.phase_prof ? clock64() : 0;
    auto tick = [&](int id) {
        if (p.phase_prof && ptid == 0) { const long long t = clock64(); atomicAdd(p.phase_prof + id, (unsigned long long)(t - t_prev)); t_prev = t; }
    };
    for (;;) {
        PcCtx& cx = sm.ctx[b];
        if (ptid == 0) sm.next_item = (long long)atomicAdd(p.work_counter, 1ull);
        bar_named(1, PC_HALF);
        const long long item = sm.next_item;
        if (item >= p.n_items) {
            if (ptid == 0) cx.item = -1;
            __syncthreads();
            return;
        }
        const long long seq = p.sidx ? (long long)p.sidx[item] : item;
        const int L = p.len[seq];
        const long long c0 = p.chunk_off[seq];
        const uint32_t* gcodes = p.codes + c0 * 4;
        const uint32_t* gnmask = p.nmask + c0 * 2;
        const uint32_t seq_id = (uint32_t)(p.seq_id0 + seq);
        const int nhalf = ((L + CHUNK_BASES - 1) / CHUNK_BASES) * 2;
        const bool fits = nhalf <= 2 * SSEQ_CHUNKS;
        if (ptid == 0) { cx.item = item; cx.defer = fits ? 0 : 1; cx.n_delta = 0; sm.nvalid = 0; sm.any_over = 0; }
        for (int i = ptid; i < PC_MAXS; i += PC_HALF) cx.dtot[i] = 0;
        if (fits) {
            // ---- count + stage the sequence ----
            for (int i = ptid; i < PC_VEC; i += PC_HALF) reinterpret_cast<int4*>(sm.hist)[i] = make_int4(0, 0, 0, 0);
            bar_named(1, PC_HALF);
            int nv = 0;
            for (int h = ptid; h < nhalf; h += PC_HALF) {
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(gcodes) + h);
                reinterpret_cast<uint2*>(sm.sseq)[h] = w;
                sm.sseq[SSEQ_CW + h] = gnmask[h];
                nv += count_half<K>(gcodes, gnmask, h, w.x, w.y, [&](uint32_t kmer) { atomicAdd(&sm.hist[kmer], 1); });
            }
            if (ptid < 4) sm.sseq[nhalf * 2 + ptid] = 0u;
            if (ptid < 2) sm.sseq[SSEQ_CW + nhalf + ptid] = 0xFFFFFFFFu;
            nv = warp_sum(nv);
            if (lane == 0 && nv) atomicAdd(&sm.nvalid, nv);
            bar_named(1, PC_HALF);
            const uint32_t* codes = sm.sseq;
            const uint32_t* nmask = sm.sseq + SSEQ_CW;
            tick(2);
            if (ptid == 0) cx.base_total = PC_F * p.pseudocount + sm.nvalid;
            for (int vec = ptid; vec < PC_VEC; vec += PC_HALF) {
                const int4 h = reinterpret_cast<const int4*>(sm.hist)[vec];
                cx.clean16[vec] = make_uint2((uint32_t)h.x | ((uint32_t)h.y << 16), (uint32_t)h.z | ((uint32_t)h.w << 16));
            }
            // ---- Random_N slots: draws (thread <-> slot, philox call) in rounds of LIST_CAP draws ----
            const int n_ent = sm.n_ent;   // <= LIST_CAP (host-checked)
            {
                constexpr int e0 = 0;
                for (int q = ptid; q < p.S * 8; q += PC_HALF) {
                    const int c = q >> 3, j = q & 7;
                    const int nb = sm.kind_class[c] == 1 ? sm.nbp[c] : 0;
                    const int off = sm.rem_off[c] - e0;
                    if (4 * j < nb && off >= 0 && off + nb <= LIST_CAP) {
                        const U4 r = random_n_words(p.seed, seq_id, (uint32_t)sm.svars[c].rng_id, (uint32_t)j);
                        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (4 * j + t < nb) sm.list[off + 4 * j + t] = random_n_entry(w[t], L) | ((uint32_t)c << 25);
                    }
                }
                bar_named(1, PC_HALF);
                constexpr uint32_t KMASK = (1u << (2 * K)) - 1u, NMASKK = (1u << K) - 1u;
                const int n_here = n_ent - e0 < LIST_CAP ? n_ent - e0 : LIST_CAP;
                for (int q = ptid; q < n_here; q += PC_HALF) {
                    const uint32_t me = sm.list[q];
                    const int c = (int)(me >> 25);
                    const int nb = sm.nbp[c], off = sm.rem_off[c] - e0, i = q - off;
                    const uint32_t pme = me >> 3;
                    uint32_t pnx = 0xFFFFFFFFu;
                    bool dup = false;
                    const uint32_t* e = sm.list + off;
                    int j = 0;
                    if ((off & 3) == 0) {
                        for (; j + 4 <= nb; j += 4) {
                            const uint4 v = *reinterpret_cast<const uint4*>(e + j);
                            const uint32_t pv[4] = {v.x >> 3, v.y >> 3, v.z >> 3, v.w >> 3};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                dup = dup || (pv[t] == pme && j + t < i);
                                if (pv[t] > pme && pv[t] < pnx) pnx = pv[t];
                            }
                        }
                    }
                    for (; j < nb; ++j) {
                        const uint32_t pj = e[j] >> 3;
                        dup = dup || (pj == pme && j < i);
                        if (pj > pme && pj < pnx) pnx = pj;
                    }
                    uint16_t* dst = cx.rem + (sm.rem_off[c] + i) * K;
                    int cnt = 0;
                    if (!dup) {
                        const int pos = (int)(pme & 0x3FFFFFu);
                        int e_hi = pos + K - 1;
                        if (pnx != 0xFFFFFFFFu && (int)(pnx & 0x3FFFFFu) - 1 < e_hi) e_hi = (int)(pnx & 0x3FFFFFu) - 1;
                        if (L - 1 < e_hi) e_hi = L - 1;
                        const Window<K> cw = load_window<K>(codes, nmask, pos - (K - 1));
                        for (int ee = pos; ee <= e_hi; ++ee) {
                            const int sh = K - 1 - (ee - pos);
                            if (((cw.nbits >> sh) & NMASKK) == 0u) dst[cnt++] = (uint16_t)((cw.bases >> (2 * sh)) & KMASK);
                        }
                    }
                    for (int r = cnt; r < K; ++r) dst[r] = 0xFFFFu;
                    if (cnt) atomicSub(&cx.dtot[c], cnt);
                }
                bar_named(1, PC_HALF);
            }
            tick(3);
            // ---- Bernoulli slots jointly: thread <-> (slot, 64-base block) ----
            const int nbs = sm.n_bern;
            const int nblocks = (L + RNG_BLOCK - 1) / RNG_BLOCK;
            if (nbs > 0 && nblocks > 0) {
                if (nbs * nblocks > PC_HALF) { if (ptid == 0) cx.defer = 1; }
                else {
                    const int j = ptid / nblocks, blk = ptid - j * nblocks;
                    const bool active = ptid < nbs * nblocks;
                    const int c = active ? sm.bern_slot[j] : 0;
                    const VarDesc vd = sm.svars[c];
                    auto table = [&](int t) -> const uint32_t* { return t < STABS ? sm.gtabs[t] : p.gtab + t * RNG_BLOCK; };
                    FastBlock f;
                    f.cnt = 0; f.ok = true;
                    if (active)
                        f = fast_block(vd.kind, p.seed, seq_id, (uint32_t)vd.rng_id, blk, L, nmask, table(vd.tab1), vd.slope1,
                                       table(vd.tab2), vd.slope2);
                    if (active && !f.ok) sm.any_over = 1;
                    int total;
                    const int off = pc_exscan(f.cnt, sm.scan, &total, ptid);
                    if (sm.any_over || total > LIST_CAP) { if (ptid == 0) cx.defer = 1; }
                    else {
                        if (active && blk == 0) sm.seg_off[j] = off;
                        if (ptid == 0) sm.seg_off[nbs] = total;
                        if (active && f.cnt) fast_block_write(f, blk, codes, sm.list + off);
                        bar_named(1, PC_HALF);
                        for (int i0 = 0; i0 < total; i0 += PC_HALF) {   // uniform trip count: warp collectives inside
                            const int i = i0 + ptid;
                            int jj = 0, so = 0, cnt = 0;
                            if (i < total) {
                                while (i >= sm.seg_off[jj + 1]) ++jj;
                                so = sm.seg_off[jj];
                                // pass 1: how many +-1 deltas does this edit produce
                                apply_entry<K>(codes, nmask, L, sm.list + so, sm.seg_off[jj + 1] - so, i - so, [&](uint32_t, int) { ++cnt; });
                            }
                            // one reservation per warp in the delta list
                            int inc = cnt;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                                if (lane >= o) inc += t;
                            }
                            const int wtot = __shfl_sync(0xffffffffu, inc, 31);
                            int wbase = 0;
                            if (lane == 31 && wtot) wbase = atomicAdd(&cx.n_delta, wtot);
                            wbase = __shfl_sync(0xffffffffu, wbase, 31);
                            int slot = wbase + inc - cnt;
                            if (cnt && wbase + wtot <= PC_DELTA) {
                                const uint32_t tag = (uint32_t)jj << 12;
                                const int d = apply_entry<K>(codes, nmask, L, sm.list + so, sm.seg_off[jj + 1] - so, i - so, [&](uint32_t kmer, int dd) {
                                    cx.delta[slot++] = (uint16_t)(kmer | tag | (dd > 0 ? 0x8000u : 0u));
                                });
                                if (d) atomicAdd(&cx.dtot[sm.bern_slot[jj]], d);
                            }
                        }
                    }
                }
            }
            bar_named(1, PC_HALF);
            if (ptid == 0 && cx.n_delta > PC_DELTA) cx.defer = 1;
        }
        if (ptid == 0 && cx.defer) { atomicOr(p.status + item, 2); atomicAdd(p.work_counter + 1, 1ull); }
        tick(4);
        __syncthreads();   // hand the context over to the consumers
        tick(5);
        b ^= 1;
    }
}

template <int OUT>
__device__ __forceinline__ void pc_consumer(PcSmem& sm, const ProfParams& p) {
    constexpr int K = PC_K, VEC = PC_VEC, VPT = PC_VPT, NT = PC_HALF, HG = PC_HG, PRIVW = PC_PRIVW;
    constexpr int ESZ = 4;
    constexpr int TPH = NT / HG;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float magic = 8388608.0f - (float)p.pseudocount;
    uint32_t* priv = sm.priv;
    // scaler statistics of this thread's two granules: fixed for the whole launch, kept in registers
    Stats2 st[VPT];
#pragma unroll
    for (int vv = 0; vv < VPT; ++vv) st[vv] = load_stats2(p.mean, p.scale, p.rscale, tid + vv * NT, OUT == IDL_OUT_STD_F32);
    int b = 0;
    long long t_prev = p.phase_prof ? clock64() : 0;
    auto tick = [&](int id) {
        if (p.phase_prof && tid == 0) { const long long t = clock64(); atomicAdd(p.phase_prof + id, (unsigned long long)(t - t_prev)); t_prev = t; }
    };
    for (;;) {
        tick(1);
        __syncthreads();   // context b is ready
        tick(0);
        PcCtx& cx = sm.ctx[b];
        b ^= 1;
        const long long item = cx.item;
        if (item < 0) return;
        if (cx.defer) continue;
        const int S = p.S;
        const int base_total = cx.base_total;
        // per-slot totals / output rows, private copies
        for (int s = tid; s < S; s += NT) {
            const float ft2 = (float)(base_total + cx.dtot[s]);
            sm.gy[s] = make_float2(ft2, 1.0f / ft2);
            sm.grow[s] = (long long)ESZ * (sm.sout_off[s] + item * p.out_stride);
        }
#pragma unroll
        for (int vv = 0; vv < VPT; ++vv) {
            const int vec = tid + vv * NT;
            const uint2 pk = cx.clean16[vec];
#pragma unroll
            for (int c = 0; c < PC_G; ++c) reinterpret_cast<uint2*>(priv + c * PRIVW)[vec] = pk;
        }
        bar_named(2, NT);
        const int n_delta = cx.n_delta;
        auto patch_half = [&](int h0, int hs, int set) {
            const int c = tid / TPH;
            if (c < hs && sm.kind_class[h0 + c] == 1) {
                const int n = sm.nbp[h0 + c] * K;
                const uint16_t* rl = cx.rem + sm.rem_off[h0 + c] * K;
                uint32_t* privc = priv + (set * HG + c) * PRIVW;
                for (int r = tid - c * TPH; r < n; r += TPH) {
                    const uint32_t km = rl[r];
                    if (km != 0xFFFFu) upd16(privc, km, -1);
                }
            }
            // Bernoulli deltas of slots that live in this half-group
            bool any = false;
            for (int c2 = 0; c2 < hs; ++c2) any = any || sm.kind_class[h0 + c2] == 2;
            if (any) {
                for (int r = tid; r < n_delta; r += NT) {
                    const uint32_t e = cx.delta[r];
                    const int slot = sm.bern_slot[(e >> 12) & 7u];
                    if (slot >= h0 && slot < h0 + hs) upd16(priv + (set * HG + slot - h0) * PRIVW, e & 0xFFFu, (e & 0x8000u) ? 1 : -1);
                }
            }
        };
        const int nh = (S + HG - 1) / HG;
        patch_half(0, S < HG ? S : HG, 0);
        bar_named(2, NT);
        for (int h = 0; h < nh; ++h) {
            const int cur = h & 1, h0 = h * HG;
            const int hs = S - h0 < HG ? S - h0 : HG;
            if (h + 1 < nh) patch_half(h0 + HG, S - h0 - HG < HG ? S - h0 - HG : HG, cur ^ 1);
            const uint2 clean0 = cx.clean16[tid], clean1 = cx.clean16[tid + NT];
#pragma unroll 2
            for (int c = 0; c < hs; ++c) {
                const float2 fy = sm.gy[h0 + c];
                unsigned char* row = reinterpret_cast<unsigned char*>(p.out) + sm.grow[h0 + c];
                uint2* src = reinterpret_cast<uint2*>(priv + (cur * HG + c) * PRIVW);
                const uint2 pk0 = src[tid], pk1 = src[tid + NT];
                src[tid] = clean0;
                src[tid + NT] = clean1;
                emit_granule_u16x2<OUT == IDL_OUT_STD_F32>(row, tid, pk0, magic, fy, st[0]);
                emit_granule_u16x2<OUT == IDL_OUT_STD_F32>(row, tid + NT, pk1, magic, fy, st[1]);
            }
            bar_named(2, NT);
        }
        (void)lane; (void)wid;
    }
}

template <int OUT>
__global__ void __launch_bounds__(PC_NT, 1) profiles_pc_kernel(const ProfParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PcSmem& sm = *reinterpret_cast<PcSmem*>(smem_raw);
    const int tid = threadIdx.x;
    // launch-uniform slot tables
    for (int i = tid; i < p.n_vars; i += PC_NT) sm.svars[i] = p.vars[i];
    for (int i = tid; i < p.S; i += PC_NT) sm.sout_off[i] = p.out_off[i];
    for (int i = tid; i < STABS * RNG_BLOCK; i += PC_NT) (&sm.gtabs[0][0])[i] = i < p.n_tabs * RNG_BLOCK ? p.gtab[i] : 0u;
    __syncthreads();
    if (tid == 0) {
        int ent = 0, nb = 0;
        for (int s = 0; s < p.S; ++s) {
            const VarDesc vd = sm.svars[s];
            int kc = 0;
            if (vd.kind == KIND_RANDOM_N) kc = vd.n_bp > 0 ? 1 : 0;
            else if (vd.kind != KIND_CLEAN) kc = 2;
            sm.kind_class[s] = kc;
            sm.nbp[s] = kc == 1 ? vd.n_bp : 0;
            sm.rem_off[s] = ent;
            sm.bern_ord[s] = kc == 2 ? nb : 0;
            if (kc == 1) ent += vd.n_bp;
            if (kc == 2) { sm.bern_slot[nb] = s; ++nb; }
        }
        sm.n_ent = ent;
        sm.n_bern = nb;
    }
    __syncthreads();
    if (tid >= PC_HALF) pc_producer<OUT>(sm, p);
    else pc_consumer<OUT>(sm, p);
}

}  // namespace idl
