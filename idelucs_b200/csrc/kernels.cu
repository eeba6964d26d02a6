// idelucs_b200 — sm_100a kernels K1 (pack), K2/K3 (k-mer profile + mimic variants), K4 (scaler)
// and the C ABI of include/idelucs_b200.h for them.  K5 (IIC loss) lives in iid_loss.cu.
//
// Design: one CTA owns one sequence at a time.  It builds the CLEAN 4^k
// histogram once in shared memory (rolling k-mer index from 2-bit packed bases read with
// 128-bit loads, shared-memory atomics, no global atomics).  Every mimic variant is then
// produced by patching that histogram with the +-1 deltas of the k windows around each
// mutated base (mutations come from a counter-based Philox RNG in registers, or from an
// explicit edit list), streaming the normalised / standardised profile to HBM with 128-bit
// streaming stores, and un-patching.  Mutated sequences never exist in memory; the path is
// bound by the HBM write of the profiles.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/idelucs_b200.h"
#include "core.cuh"
#include "common.h"

namespace idl {

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, const char* a, long long b) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

constexpr int LIST_CAP = 4096;   // on-chip edit list entries per CTA
constexpr int MAX_GROUP = 256;   // variant slots per Random_N group

struct VarDesc {
    int32_t kind, rng_id, n_bp, explicit_idx, tab1, tab2;
};
struct GroupDesc {
    int32_t first_slot, n_slots, kind;
};

struct ProfParams {
    const uint32_t* codes;
    const uint32_t* nmask;
    const int64_t* chunk_off;
    const int32_t* len;
    const int32_t* sidx;
    const int32_t* sel;
    long long n_items, seq_id0, n_seqs_total;
    int S, n_groups;
    const VarDesc* vars;
    const GroupDesc* groups;
    unsigned long long seed;
    const uint32_t* gtab;
    const int64_t* edit_off;
    const uint32_t* edits;
    void* out;
    const int64_t* out_off;
    long long out_stride;
    int pseudocount, accumulate;
    const float* mean;
    const float* scale;
    int32_t* status;
    unsigned long long* work_counter;
};

// ---------------------------------------------------------------------------------------
// K1 pack
// ---------------------------------------------------------------------------------------
constexpr int PACK_NT = 256;

__global__ void __launch_bounds__(PACK_NT) pack_kernel(const uint8_t* __restrict__ ascii,
                                                        const int64_t* __restrict__ byte_off,
                                                        const int64_t* __restrict__ chunk_off, long long n, int strict,
                                                        uint32_t* __restrict__ codes, uint32_t* __restrict__ nmask,
                                                        int32_t* __restrict__ len, unsigned long long* __restrict__ bad) {
    for (long long s = blockIdx.x; s <= n; s += gridDim.x) {
        if (s == n) {  // slack chunk behind the last sequence: all padding
            if (blockIdx.y == 0 && threadIdx.x < 4) codes[chunk_off[n] * 4 + threadIdx.x] = 0u;
            if (blockIdx.y == 0 && threadIdx.x < 2) nmask[chunk_off[n] * 2 + threadIdx.x] = 0xFFFFFFFFu;
            continue;
        }
        const long long b0 = byte_off[s];
        const long long L = byte_off[s + 1] - b0;
        const long long c0 = chunk_off[s];
        const long long nhalf = (chunk_off[s + 1] - c0) * 2;  // 32-base half chunks
        if (blockIdx.y == 0 && threadIdx.x == 0) len[s] = (int32_t)L;
        unsigned long long first_bad = ~0ull;
        for (long long h = (long long)blockIdx.y * PACK_NT + threadIdx.x; h < nhalf; h += (long long)PACK_NT * gridDim.y) {
            uint32_t mw = 0;
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const long long base = h * 32 + t * 16;
                long long avail = L - base;
                avail = avail < 0 ? 0 : (avail > 16 ? 16 : avail);
                uint8_t buf[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) buf[j] = j < avail ? __ldg(ascii + b0 + base + j) : (uint8_t)'N';
                uint32_t cw, m16;
                int bj, bc;
                pack16(buf, (int)avail, strict, &cw, &m16, &bj, &bc);
                codes[c0 * 4 + h * 2 + t] = cw;
                mw = (mw << 16) | m16;
                if (bj < 16) {
                    const unsigned long long v = ((unsigned long long)(base + bj) << 3) | (unsigned long long)bc;
                    first_bad = v < first_bad ? v : first_bad;
                }
            }
            nmask[c0 * 2 + h] = mw;
        }
        if (first_bad != ~0ull) atomicMin(bad + s, first_bad);
    }
}

// ---------------------------------------------------------------------------------------
// block helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exclusive scan of one int per thread over the CTA; returns the prefix, *total = CTA sum.
// scratch: NT/32 + 1 ints of shared memory.  Contains two __syncthreads.
template <int NT>
__device__ __forceinline__ int block_exscan(int v, int* scratch, int* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < NT / 32 ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < NT / 32) scratch[lane] = winc - w;
        if (lane == 31) scratch[NT / 32] = winc;
    }
    __syncthreads();
    *total = scratch[NT / 32];
    return scratch[wid] + inc - v;
}

// ---------------------------------------------------------------------------------------
// K2/K3 profiles kernel
// ---------------------------------------------------------------------------------------
template <int K, int NT>
struct ProfSmem {
    static constexpr int F = 1 << (2 * K);
    int hist[F];
    uint32_t list[LIST_CAP + 8];
    uint32_t tmp[LIST_CAP];  // unsorted Random_N draws
    uint32_t gtab[2][RNG_BLOCK];
    int dtot[MAX_GROUP];
    int scan[NT / 32 + 2];
    int nvalid;
    long long item;
};

template <int K, int NT, int OUT>
__device__ __forceinline__ void stream_slot(const int* hist, int total, int pc, int accumulate, void* out_row,
                                            const float (&mean)[((1 << (2 * K)) / 4 + NT - 1) / NT][4],
                                            const float (&scale)[((1 << (2 * K)) / 4 + NT - 1) / NT][4],
                                            const float (&rscale)[((1 << (2 * K)) / 4 + NT - 1) / NT][4]) {
    constexpr int F = 1 << (2 * K);
    constexpr int VEC = F / 4;
    constexpr int VPT = (VEC + NT - 1) / NT;
    const float ft = (float)total;
    const float y = 1.0f / ft;  // IEEE (correctly rounded) reciprocal
    const bool big = total >= (1 << 24);  // float no longer exact: divide in double
#pragma unroll
    for (int vv = 0; vv < VPT; ++vv) {
        const int vec = threadIdx.x + vv * NT;
        if (VEC % NT != 0 && vec >= VEC) break;
        const int4 c = reinterpret_cast<const int4*>(hist)[vec];
        const int cc[4] = {c.x + pc, c.y + pc, c.z + pc, c.w + pc};
        if (OUT == IDL_OUT_COUNTS_I32) {
            int4* dst = reinterpret_cast<int4*>(out_row) + vec;
            int4 o = c;
            if (accumulate) { const int4 prev = *dst; o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w; }
            *dst = o;
        } else if (OUT == IDL_OUT_FREQ_F64) {
            double2* dst = reinterpret_cast<double2*>(out_row) + 2 * vec;
            const double dt = (double)total;
            __stcs(dst, make_double2((double)cc[0] / dt, (double)cc[1] / dt));
            __stcs(dst + 1, make_double2((double)cc[2] / dt, (double)cc[3] / dt));
        } else {
            float q[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // float32(count/total): for count,total < 2^24 the float division is exactly the
                // reference's float64 division followed by astype(float32)
                q[e] = big ? (float)((double)cc[e] / (double)total) : div_rn((float)cc[e], ft, y);
                if (OUT == IDL_OUT_STD_F32) q[e] = div_rn(q[e] - mean[vv][e], scale[vv][e], rscale[vv][e]);
            }
            __stcs(reinterpret_cast<float4*>(out_row) + vec, make_float4(q[0], q[1], q[2], q[3]));
        }
    }
}

template <int K, int NT, int OUT>
__global__ void __launch_bounds__(NT, (K == 6 ? 2 : K == 5 ? 4 : 8)) profiles_kernel(const ProfParams p) {
    constexpr int F = 1 << (2 * K);
    constexpr int VEC = F / 4;
    constexpr int VPT = (VEC + NT - 1) / NT;
    constexpr int ESZ = OUT == IDL_OUT_FREQ_F64 ? 8 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ProfSmem<K, NT>& sm = *reinterpret_cast<ProfSmem<K, NT>*>(smem_raw);
    const int tid = threadIdx.x;

    // per-thread scaler statistics of the bins this thread streams (fixed for the whole launch)
    float mean[VPT][4], scale[VPT][4], rscale[VPT][4];
#pragma unroll
    for (int vv = 0; vv < VPT; ++vv) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { mean[vv][e] = 0.f; scale[vv][e] = 1.f; rscale[vv][e] = 1.f; }
        const int vec = tid + vv * NT;
        if (OUT == IDL_OUT_STD_F32 && vec < VEC) {
            const float4 m = reinterpret_cast<const float4*>(p.mean)[vec];
            const float4 s = reinterpret_cast<const float4*>(p.scale)[vec];
            mean[vv][0] = m.x; mean[vv][1] = m.y; mean[vv][2] = m.z; mean[vv][3] = m.w;
            scale[vv][0] = s.x; scale[vv][1] = s.y; scale[vv][2] = s.z; scale[vv][3] = s.w;
#pragma unroll
            for (int e = 0; e < 4; ++e) rscale[vv][e] = 1.0f / scale[vv][e];
        }
    }
    for (int i = tid; i < 2 * RNG_BLOCK; i += NT) (&sm.gtab[0][0])[i] = 0u;

    for (;;) {
        __syncthreads();  // previous item fully done (also protects sm.item)
        if (tid == 0) sm.item = (long long)atomicAdd(p.work_counter, 1ull);
        __syncthreads();
        const long long item = sm.item;
        if (item >= p.n_items) break;
        const long long seq = p.sidx ? (long long)p.sidx[item] : item;
        const int L = p.len[seq];
        const long long c0 = p.chunk_off[seq];
        const uint32_t* __restrict__ codes = p.codes + c0 * 4;
        const uint32_t* __restrict__ nmask = p.nmask + c0 * 2;
        const int nchunks = (L + CHUNK_BASES - 1) / CHUNK_BASES;
        const uint32_t seq_id = (uint32_t)(p.seq_id0 + seq);

        // ---- clean histogram -------------------------------------------------------------
        for (int i = tid; i < VEC; i += NT) reinterpret_cast<int4*>(sm.hist)[i] = make_int4(0, 0, 0, 0);
        if (tid == 0) sm.nvalid = 0;
        __syncthreads();
        {
            int nv = 0;
            for (int c = tid; c < nchunks; c += NT) {
                const uint4 w = __ldg(reinterpret_cast<const uint4*>(codes) + c);
                nv += count_chunk<K>(codes, nmask, c, w.x, w.y, w.z, w.w, [&](uint32_t kmer) { atomicAdd(&sm.hist[kmer], 1); });
            }
            nv = warp_sum(nv);
            if ((tid & 31) == 0 && nv) atomicAdd(&sm.nvalid, nv);
        }
        __syncthreads();
        const int base_total = F * p.pseudocount + sm.nvalid;

        // ---- variant groups --------------------------------------------------------------
        const int n_groups = p.sel ? p.S : p.n_groups;
        for (int g = 0; g < n_groups; ++g) {
            int s0, ns, kind, v0;
            if (p.sel) { s0 = g; ns = 1; v0 = p.sel[item * p.S + g]; kind = p.vars[v0].kind; }
            else { const GroupDesc gd = p.groups[g]; s0 = gd.first_slot; ns = gd.n_slots; kind = gd.kind; v0 = s0; }
            const VarDesc vd = p.vars[v0];
            auto out_row = [&](int slot) -> void* {
                return reinterpret_cast<unsigned char*>(p.out) + (size_t)ESZ * (size_t)(p.out_off[slot] + item * p.out_stride);
            };

            if (kind == KIND_CLEAN || (kind == KIND_RANDOM_N && (L == 0 || vd.n_bp <= 0))) {
                for (int sl = 0; sl < ns; ++sl)
                    stream_slot<K, NT, OUT>(sm.hist, base_total, p.pseudocount, p.accumulate, out_row(s0 + sl), mean, scale, rscale);
                __syncthreads();
            } else if (kind == KIND_RANDOM_N) {
                // all slots of the group: draw -> rank-sort per slot -> lists in shared memory
                const int n_bp = vd.n_bp;
                const int cps = (n_bp + 3) / 4;  // philox calls per slot
                for (int idx = tid; idx < ns * cps; idx += NT) {
                    const int sl = idx / cps, call = idx - sl * cps;
                    const int rng_id = p.sel ? vd.rng_id : p.vars[s0 + sl].rng_id;
                    const U4 r = random_n_words(p.seed, seq_id, (uint32_t)rng_id, (uint32_t)call);
                    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (call * 4 + t < n_bp) sm.tmp[sl * n_bp + call * 4 + t] = random_n_entry(w[t], L);
                }
                for (int i = tid; i < ns; i += NT) sm.dtot[i] = 0;
                __syncthreads();
                for (int idx = tid; idx < ns * n_bp; idx += NT) {
                    const int sl = idx / n_bp, i = idx - sl * n_bp;
                    const uint32_t* t = sm.tmp + sl * n_bp;
                    const uint32_t e = t[i];
                    int rank = 0;
                    for (int j = 0; j < n_bp; ++j) { const uint32_t ej = t[j]; rank += (ej < e || (ej == e && j < i)) ? 1 : 0; }
                    sm.list[sl * n_bp + rank] = e;
                }
                __syncthreads();
                const int n_pad = (n_bp + 31) & ~31;
                for (int sl = 0; sl <= ns; ++sl) {
                    // un-apply slot sl-1 and apply slot sl (independent +-1 updates, different warps)
                    for (int idx = tid; idx < 2 * n_pad; idx += NT) {
                        const int role = idx / n_pad, i = idx - role * n_pad;
                        const int tsl = role == 0 ? sl : sl - 1;
                        if (i < n_bp && tsl >= 0 && tsl < ns) {
                            const int sgn = role == 0 ? 1 : -1;
                            const int d = apply_entry<K>(codes, nmask, L, sm.list + tsl * n_bp, n_bp, i,
                                                         [&](uint32_t kmer, int dd) { atomicAdd(&sm.hist[kmer], sgn * dd); });
                            if (role == 0 && d) atomicAdd(&sm.dtot[tsl], d);
                        }
                    }
                    __syncthreads();
                    if (sl < ns) {
                        stream_slot<K, NT, OUT>(sm.hist, base_total + sm.dtot[sl], p.pseudocount, p.accumulate, out_row(s0 + sl), mean, scale, rscale);
                        __syncthreads();
                    }
                }
            } else if (kind == KIND_EXPLICIT) {
                const long long li = (long long)vd.explicit_idx * p.n_seqs_total + seq;
                const uint32_t* glist = p.edits + p.edit_off[li];
                const int n = (int)(p.edit_off[li + 1] - p.edit_off[li]);
                if (tid == 0) sm.dtot[0] = 0;
                __syncthreads();
                for (int sgn = 1; sgn >= -1; sgn -= 2) {
                    int d = 0;
                    for (int i = tid; i < n; i += NT)
                        d += apply_entry<K>(codes, nmask, L, glist, n, i, [&](uint32_t kmer, int dd) { atomicAdd(&sm.hist[kmer], sgn * dd); });
                    if (sgn > 0) { d = warp_sum(d); if ((tid & 31) == 0 && d) atomicAdd(&sm.dtot[0], d); }
                    __syncthreads();
                    if (sgn > 0) {
                        stream_slot<K, NT, OUT>(sm.hist, base_total + sm.dtot[0], p.pseudocount, p.accumulate, out_row(s0), mean, scale, rscale);
                        __syncthreads();
                    }
                }
            } else {
                // TRANSITION / TRANSVERSION / BOTH: Bernoulli hits generated per 128-base block
                if (tid < RNG_BLOCK) {
                    sm.gtab[0][tid] = p.gtab[vd.tab1 * RNG_BLOCK + tid];
                    sm.gtab[1][tid] = p.gtab[vd.tab2 * RNG_BLOCK + tid];
                }
                if (tid == 0) sm.dtot[0] = 0;
                __syncthreads();
                constexpr int TB = NT - 2;  // blocks per tile; one context block on each side
                const int nblocks = (L + RNG_BLOCK - 1) / RNG_BLOCK;
                for (int sgn = 1; sgn >= -1; sgn -= 2) {
                    for (int tb0 = 0; tb0 < nblocks; tb0 += TB) {
                        const int b = tb0 - 1 + tid;
                        const bool active = b >= 0 && b < nblocks && b <= tb0 + TB;
                        int cnt = 0;
                        if (active)
                            cnt = block_edits(kind, p.seed, seq_id, (uint32_t)vd.rng_id, b, L, codes, nmask, sm.gtab[0], sm.gtab[1], [](uint32_t) {});
                        int total;
                        int off = block_exscan<NT>(cnt, sm.scan, &total);
                        if (total > LIST_CAP) {
                            if (tid == 0 && p.status) atomicOr(p.status + item, 1);
                            continue;  // uniform: total is the same for every thread
                        }
                        if (active && cnt)
                            block_edits(kind, p.seed, seq_id, (uint32_t)vd.rng_id, b, L, codes, nmask, sm.gtab[0], sm.gtab[1], [&](uint32_t e) { sm.list[off++] = e; });
                        __syncthreads();
                        const int lo = tb0 * RNG_BLOCK;
                        const long long hi = (long long)(tb0 + TB) * RNG_BLOCK;
                        int d = 0;
                        for (int i = tid; i < total; i += NT) {
                            const int pos = (int)(sm.list[i] >> 3);
                            if (pos >= lo && pos < hi)
                                d += apply_entry<K>(codes, nmask, L, sm.list, total, i, [&](uint32_t kmer, int dd) { atomicAdd(&sm.hist[kmer], sgn * dd); });
                        }
                        if (sgn > 0) { d = warp_sum(d); if ((tid & 31) == 0 && d) atomicAdd(&sm.dtot[0], d); }
                        __syncthreads();
                    }
                    if (sgn > 0) {
                        stream_slot<K, NT, OUT>(sm.hist, base_total + sm.dtot[0], p.pseudocount, p.accumulate, out_row(s0), mean, scale, rscale);
                        __syncthreads();
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// K4: column statistics / scaler / standardise
// ---------------------------------------------------------------------------------------
constexpr int CS_NT = 128;
constexpr int CS_ROWS = 2048;  // rows per partial

template <typename T>
__global__ void __launch_bounds__(CS_NT) colstats_kernel(const T* __restrict__ x, long long n, int F,
                                                          double* __restrict__ partials, double* __restrict__ part_n) {
    const int col = blockIdx.x * CS_NT + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * CS_ROWS;
    const long long r1 = r0 + CS_ROWS < n ? r0 + CS_ROWS : n;
    if (col == 0) part_n[blockIdx.y] = (double)(r1 - r0);
    if (col >= F) return;
    // shifted-data sums (shift = first row of this part): exact zeros for a constant column
    const double shift = (double)x[r0 * F + col];
    double s1 = 0.0, s2 = 0.0;
    for (long long r = r0; r < r1; ++r) {
        const double d = (double)x[r * F + col] - shift;
        s1 += d;
        s2 = fma(d, d, s2);
    }
    const double m = (double)(r1 - r0);
    partials[((long long)blockIdx.y * 2 + 0) * F + col] = shift + s1 / m;  // mean of the part
    partials[((long long)blockIdx.y * 2 + 1) * F + col] = s2 - s1 * s1 / m;  // M2 of the part
}

__global__ void scaler_finalize_kernel(const double* __restrict__ partials, const double* __restrict__ part_n, int n_parts,
                                       int F, double* mean64, double* var64, double* scale64, float* mean32, float* scale32) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= F) return;
    double na = 0.0, ma = 0.0, M2 = 0.0;
    for (int g = 0; g < n_parts; ++g) {  // Chan et al. pairwise merge, fixed order
        const double nb = part_n[g];
        if (nb <= 0.0) continue;
        const double mb = partials[((long long)g * 2 + 0) * F + col];
        const double Mb = partials[((long long)g * 2 + 1) * F + col];
        const double nt = na + nb;
        const double delta = mb - ma;
        ma = ma + delta * (nb / nt);
        M2 = M2 + Mb + delta * delta * (na * nb / nt);
        na = nt;
    }
    double var = na > 0.0 ? M2 / na : 0.0;
    if (var < 0.0) var = 0.0;
    // sklearn/preprocessing/_data.py::_is_constant_feature (float64 eps) -> scale 1
    const double eps = 2.220446049250313e-16;
    const double bound = na * eps * var + (na * ma * eps) * (na * ma * eps);
    double sc = sqrt(var);
    if (var <= bound) sc = 1.0;
    if (mean64) mean64[col] = ma;
    if (var64) var64[col] = var;
    if (scale64) scale64[col] = sc;
    if (mean32) mean32[col] = (float)ma;
    if (scale32) scale32[col] = (float)sc;
}

__global__ void standardize_f32_kernel(const float* __restrict__ x, float* __restrict__ out, long long n4, int F4,
                                       const float* __restrict__ mean, const float* __restrict__ scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % F4);
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c);
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale) + c);
        float4 o;
        o.x = (v.x - m.x) / s.x; o.y = (v.y - m.y) / s.y; o.z = (v.z - m.z) / s.z; o.w = (v.w - m.w) / s.w;
        reinterpret_cast<float4*>(out)[i] = o;
    }
}

__global__ void standardize_f64_kernel(const double* __restrict__ x, double* __restrict__ out64, float* __restrict__ out32,
                                       long long total, int F, const double* __restrict__ mean, const double* __restrict__ scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % F);
        const double o = (x[i] - mean[c]) / scale[c];
        if (out64) out64[i] = o;
        if (out32) out32[i] = (float)o;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_sm_count = 148;
    }
    return g_sm_count;
}

// workspace layout: [0,8) work counter | vars | groups | out_off | gtab
constexpr size_t WS_VARS = 64;
constexpr int MAX_VARIANTS = 4096;
constexpr int MAX_TABS = 64;
constexpr size_t WS_GROUPS = WS_VARS + sizeof(VarDesc) * MAX_VARIANTS;
constexpr size_t WS_OUTOFF = WS_GROUPS + sizeof(GroupDesc) * MAX_VARIANTS;
constexpr size_t WS_GTAB = WS_OUTOFF + sizeof(int64_t) * MAX_VARIANTS;
constexpr size_t WS_TOTAL = WS_GTAB + sizeof(uint32_t) * RNG_BLOCK * MAX_TABS;

template <int K, int NT, int OUT>
static int launch_profiles(const ProfParams& p, cudaStream_t st) {
    const size_t smem = sizeof(ProfSmem<K, NT>);
    auto kern = profiles_kernel<K, NT, OUT>;
    static bool configured = false;
    if (!configured) {
        IDL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int per_sm = 0;
    IDL_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) return set_error(IDL_ECUDA, "profiles kernel does not fit on an SM%s", "");
    long long grid = (long long)sm_count() * per_sm;
    if (grid > p.n_items) grid = p.n_items;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, NT, smem, st>>>(p);
    IDL_CUDA_CHECK(cudaGetLastError());
    return IDL_OK;
}

template <int K, int NT>
static int dispatch_out(const ProfParams& p, int out_kind, cudaStream_t st) {
    switch (out_kind) {
        case IDL_OUT_COUNTS_I32: return launch_profiles<K, NT, IDL_OUT_COUNTS_I32>(p, st);
        case IDL_OUT_FREQ_F32: return launch_profiles<K, NT, IDL_OUT_FREQ_F32>(p, st);
        case IDL_OUT_STD_F32: return launch_profiles<K, NT, IDL_OUT_STD_F32>(p, st);
        case IDL_OUT_FREQ_F64: return launch_profiles<K, NT, IDL_OUT_FREQ_F64>(p, st);
    }
    return set_error(IDL_EINVAL, "unknown out_kind%s %lld", "", out_kind);
}

}  // namespace idl

using namespace idl;

extern "C" {

int idl_abi_version(void) { return IDL_ABI_VERSION; }
const char* idl_last_error(void) { return g_err; }

int idl_geometric_table(double p, uint32_t* out128) {
    if (!out128 || !(p >= 0.0) || !(p <= 1.0)) return set_error(IDL_EINVAL, "idl_geometric_table: bad argument%s", "");
    geometric_table(p, out128);
    return IDL_OK;
}

int idl_pack(const uint8_t* d_ascii, const int64_t* d_byte_off, int64_t n, int alphabet, int64_t max_len,
             const int64_t* d_chunk_off, uint32_t* d_codes, uint32_t* d_nmask, int32_t* d_len,
             unsigned long long* d_bad, void* stream) {
    if (n < 0 || !d_byte_off || !d_chunk_off || !d_codes || !d_nmask || !d_len || !d_bad)
        return set_error(IDL_EINVAL, "idl_pack: null pointer or negative n%s", "");
    if (max_len > 0x7fffffffLL) return set_error(IDL_EUNSUPPORTED, "idl_pack: sequences longer than 2^31-1 bases%s", "");
    long long gx = n + 1;
    const long long cap = (long long)sm_count() * 32;
    if (gx > cap) gx = cap;
    long long gy = (max_len + (32LL * PACK_NT * 8) - 1) / (32LL * PACK_NT * 8);
    if (gy < 1) gy = 1;
    if (gy > 1024) gy = 1024;
    if (gx * gy > cap * 4 && gx > 1) { gx = cap * 4 / gy; if (gx < 1) gx = 1; }
    pack_kernel<<<dim3((unsigned)gx, (unsigned)gy), PACK_NT, 0, (cudaStream_t)stream>>>(
        d_ascii, d_byte_off, d_chunk_off, n, alphabet, d_codes, d_nmask, d_len, d_bad);
    IDL_CUDA_CHECK(cudaGetLastError());
    return IDL_OK;
}

size_t idl_profiles_workspace_bytes(void) { return WS_TOTAL; }

int idl_profiles(const uint32_t* d_codes, const uint32_t* d_nmask, const int64_t* d_chunk_off,
                 const int32_t* d_len, int64_t n_seqs_total, const int32_t* d_sidx, int64_t n_items,
                 int64_t seq_id0, int k, const idl_variant* variants, int n_variants, const int32_t* d_sel,
                 int S, uint64_t seed, const int64_t* d_edit_off, const uint32_t* d_edits, int out_kind,
                 void* d_out, const int64_t* out_off, int64_t out_stride, int pseudocount, int accumulate,
                 const float* d_mean, const float* d_scale, int32_t* d_status, void* d_workspace,
                 size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!d_codes || !d_nmask || !d_chunk_off || !d_len || !variants || !d_out || !out_off || !d_workspace)
        return set_error(IDL_EINVAL, "idl_profiles: null pointer%s", "");
    if (workspace_bytes < WS_TOTAL) return set_error(IDL_EINVAL, "idl_profiles: workspace too small%s (need %lld bytes)", "", (long long)WS_TOTAL);
    if (k < 1 || k > 6) return set_error(IDL_EUNSUPPORTED, "idl_profiles: k must be in 1..6%s (got %lld)", "", k);
    if (n_variants < 1 || n_variants > MAX_VARIANTS || S < 1 || S > MAX_VARIANTS)
        return set_error(IDL_EINVAL, "idl_profiles: n_variants / S out of range%s", "");
    if (!d_sel && S != n_variants) return set_error(IDL_EINVAL, "idl_profiles: S must equal n_variants without d_sel%s", "");
    if (out_kind == IDL_OUT_STD_F32 && (!d_mean || !d_scale)) return set_error(IDL_EINVAL, "idl_profiles: mean/scale required%s", "");
    if (accumulate && out_kind != IDL_OUT_COUNTS_I32) return set_error(IDL_EINVAL, "idl_profiles: accumulate only for counts%s", "");
    if (n_items <= 0) return IDL_OK;
    const int F = 1 << (2 * k);
    if (k >= 1 && F >= 4 && (out_stride % 4 != 0)) return set_error(IDL_EINVAL, "idl_profiles: out_stride must be a multiple of 4%s", "");

    // ---- host-side plan: variant descriptors, gap tables, slot groups ----
    static thread_local VarDesc h_vars[MAX_VARIANTS];
    static thread_local GroupDesc h_groups[MAX_VARIANTS];
    static thread_local uint32_t h_gtab[MAX_TABS * RNG_BLOCK];
    double tab_p[MAX_TABS];
    int n_tabs = 0;
    auto table_of = [&](double pr) -> int {
        for (int t = 0; t < n_tabs; ++t)
            if (tab_p[t] == pr) return t;
        if (n_tabs == MAX_TABS) return -1;
        tab_p[n_tabs] = pr;
        geometric_table(pr, h_gtab + n_tabs * RNG_BLOCK);
        return n_tabs++;
    };
    for (int v = 0; v < n_variants; ++v) {
        const idl_variant& iv = variants[v];
        VarDesc& d = h_vars[v];
        d.kind = iv.kind; d.rng_id = iv.rng_id; d.n_bp = iv.n_bp; d.explicit_idx = iv.explicit_idx; d.tab1 = 0; d.tab2 = 0;
        if (iv.kind < IDL_KIND_CLEAN || iv.kind > IDL_KIND_EXPLICIT) return set_error(IDL_EINVAL, "idl_profiles: bad variant kind%s %lld", "", iv.kind);
        if (iv.kind == IDL_KIND_TRANSITION || iv.kind == IDL_KIND_BOTH) {
            if (!(iv.p1 >= 0.0 && iv.p1 <= 1.0)) return set_error(IDL_EINVAL, "idl_profiles: p1 out of range%s", "");
            d.tab1 = table_of(iv.p1);
        }
        if (iv.kind == IDL_KIND_TRANSVERSION || iv.kind == IDL_KIND_BOTH) {
            if (!(iv.p2 >= 0.0 && iv.p2 <= 1.0)) return set_error(IDL_EINVAL, "idl_profiles: p2 out of range%s", "");
            d.tab2 = table_of(iv.p2);
        }
        if (d.tab1 < 0 || d.tab2 < 0) return set_error(IDL_EUNSUPPORTED, "idl_profiles: too many distinct mutation rates%s", "");
        if (iv.kind == IDL_KIND_RANDOM_N && (iv.n_bp < 0 || iv.n_bp > LIST_CAP)) return set_error(IDL_EUNSUPPORTED, "idl_profiles: Random_N n_bp must be <= 4096%s", "");
        if (iv.kind == IDL_KIND_EXPLICIT && (!d_edit_off || !d_edits)) return set_error(IDL_EINVAL, "idl_profiles: explicit variant without edit lists%s", "");
    }
    int n_groups = 0;
    if (!d_sel) {
        for (int s = 0; s < S;) {
            GroupDesc& g = h_groups[n_groups++];
            g.first_slot = s; g.kind = h_vars[s].kind; g.n_slots = 1;
            if (g.kind == IDL_KIND_RANDOM_N || g.kind == IDL_KIND_CLEAN) {
                const int n_bp = h_vars[s].n_bp > 0 ? h_vars[s].n_bp : 1;
                int cap = LIST_CAP / n_bp;
                if (cap > MAX_GROUP) cap = MAX_GROUP;
                while (s + g.n_slots < S && g.n_slots < cap && h_vars[s + g.n_slots].kind == g.kind &&
                       (g.kind == IDL_KIND_CLEAN || h_vars[s + g.n_slots].n_bp == h_vars[s].n_bp))
                    ++g.n_slots;
            }
            s += g.n_slots;
        }
    }
    unsigned char* ws = reinterpret_cast<unsigned char*>(d_workspace);
    IDL_CUDA_CHECK(cudaMemsetAsync(ws, 0, 8, st));
    IDL_CUDA_CHECK(cudaMemcpyAsync(ws + WS_VARS, h_vars, sizeof(VarDesc) * n_variants, cudaMemcpyHostToDevice, st));
    if (n_groups) IDL_CUDA_CHECK(cudaMemcpyAsync(ws + WS_GROUPS, h_groups, sizeof(GroupDesc) * n_groups, cudaMemcpyHostToDevice, st));
    IDL_CUDA_CHECK(cudaMemcpyAsync(ws + WS_OUTOFF, out_off, sizeof(int64_t) * S, cudaMemcpyHostToDevice, st));
    if (n_tabs) IDL_CUDA_CHECK(cudaMemcpyAsync(ws + WS_GTAB, h_gtab, sizeof(uint32_t) * RNG_BLOCK * n_tabs, cudaMemcpyHostToDevice, st));
    // the staging arrays are thread_local statics reused by the next call: make sure the
    // async copies have read them before returning (pageable memcpy is staged synchronously
    // by the runtime with respect to the host buffer, so no stream sync is required)

    ProfParams p;
    p.codes = d_codes; p.nmask = d_nmask; p.chunk_off = d_chunk_off; p.len = d_len; p.sidx = d_sidx; p.sel = d_sel;
    p.n_items = n_items; p.seq_id0 = seq_id0; p.n_seqs_total = n_seqs_total; p.S = S; p.n_groups = n_groups;
    p.vars = reinterpret_cast<const VarDesc*>(ws + WS_VARS);
    p.groups = reinterpret_cast<const GroupDesc*>(ws + WS_GROUPS);
    p.seed = seed; p.gtab = reinterpret_cast<const uint32_t*>(ws + WS_GTAB);
    p.edit_off = d_edit_off; p.edits = d_edits; p.out = d_out;
    p.out_off = reinterpret_cast<const int64_t*>(ws + WS_OUTOFF);
    p.out_stride = out_stride; p.pseudocount = pseudocount; p.accumulate = accumulate;
    p.mean = d_mean; p.scale = d_scale; p.status = d_status;
    p.work_counter = reinterpret_cast<unsigned long long*>(ws);
    switch (k) {
        case 1: return dispatch_out<1, 64>(p, out_kind, st);
        case 2: return dispatch_out<2, 64>(p, out_kind, st);
        case 3: return dispatch_out<3, 64>(p, out_kind, st);
        case 4: return dispatch_out<4, 128>(p, out_kind, st);
        case 5: return dispatch_out<5, 256>(p, out_kind, st);
        case 6: return dispatch_out<6, 512>(p, out_kind, st);
    }
    return set_error(IDL_EUNSUPPORTED, "idl_profiles: unsupported k%s", "");
}

int idl_kmer_counts(const uint32_t* d_codes, const uint32_t* d_nmask, const int64_t* d_chunk_off,
                    const int32_t* d_len, int64_t n, int k, int32_t* d_counts, int accumulate,
                    void* d_workspace, size_t workspace_bytes, void* stream) {
    idl_variant v;
    memset(&v, 0, sizeof(v));
    v.kind = IDL_KIND_CLEAN;
    const int64_t off0 = 0;
    if (k < 1 || k > 6) return set_error(IDL_EUNSUPPORTED, "idl_kmer_counts: k must be in 1..6%s (got %lld)", "", k);
    return idl_profiles(d_codes, d_nmask, d_chunk_off, d_len, n, nullptr, n, 0, k, &v, 1, nullptr, 1, 0, nullptr, nullptr,
                        IDL_OUT_COUNTS_I32, d_counts, &off0, (int64_t)1 << (2 * k), 0, accumulate, nullptr, nullptr, nullptr,
                        d_workspace, workspace_bytes, stream);
}

int idl_colstats_parts(int64_t n) { return (int)((n + CS_ROWS - 1) / CS_ROWS); }

int idl_colstats(const void* d_x, int is_f64, int64_t n, int F, double* d_partials, double* d_part_n, void* stream) {
    if (!d_x || !d_partials || !d_part_n || n <= 0 || F <= 0) return set_error(IDL_EINVAL, "idl_colstats: bad argument%s", "");
    const dim3 grid((F + CS_NT - 1) / CS_NT, (unsigned)idl_colstats_parts(n));
    if (is_f64) colstats_kernel<double><<<grid, CS_NT, 0, (cudaStream_t)stream>>>((const double*)d_x, n, F, d_partials, d_part_n);
    else colstats_kernel<float><<<grid, CS_NT, 0, (cudaStream_t)stream>>>((const float*)d_x, n, F, d_partials, d_part_n);
    IDL_CUDA_CHECK(cudaGetLastError());
    return IDL_OK;
}

int idl_scaler_finalize(const double* d_partials, const double* d_part_n, int n_parts, int F, double* d_mean64,
                        double* d_var64, double* d_scale64, float* d_mean32, float* d_scale32, void* stream) {
    if (!d_partials || !d_part_n || n_parts <= 0 || F <= 0) return set_error(IDL_EINVAL, "idl_scaler_finalize: bad argument%s", "");
    scaler_finalize_kernel<<<(F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_partials, d_part_n, n_parts, F, d_mean64, d_var64,
                                                                              d_scale64, d_mean32, d_scale32);
    IDL_CUDA_CHECK(cudaGetLastError());
    return IDL_OK;
}

int idl_standardize_f32(const float* d_x, float* d_out, int64_t n, int F, const float* d_mean32, const float* d_scale32,
                        void* stream) {
    if (!d_x || !d_out || !d_mean32 || !d_scale32 || n < 0 || F <= 0 || F % 4) return set_error(IDL_EINVAL, "idl_standardize_f32: bad argument%s", "");
    if (n == 0) return IDL_OK;
    const long long n4 = n * (F / 4);
    long long grid = (n4 + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (grid > cap) grid = cap;
    standardize_f32_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_x, d_out, n4, F / 4, d_mean32, d_scale32);
    IDL_CUDA_CHECK(cudaGetLastError());
    return IDL_OK;
}

int idl_standardize_f64(const double* d_x, double* d_out64, float* d_out32, int64_t n, int F, const double* d_mean64,
                        const double* d_scale64, void* stream) {
    if (!d_x || (!d_out64 && !d_out32) || !d_mean64 || !d_scale64 || n < 0 || F <= 0) return set_error(IDL_EINVAL, "idl_standardize_f64: bad argument%s", "");
    if (n == 0) return IDL_OK;
    const long long total = n * (long long)F;
    long long grid = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    if (grid > cap) grid = cap;
    standardize_f64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_x, d_out64, d_out32, total, F, d_mean64, d_scale64);
    IDL_CUDA_CHECK(cudaGetLastError());
    return IDL_OK;
}

}  // extern "C"
