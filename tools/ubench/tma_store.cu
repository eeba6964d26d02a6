// Micro-benchmark: what is the write ceiling of the TMA bulk-store path (cp.async.bulk.global.shared::cta) that the
// dominant kernel (profiles_pc_kernel) streams its rows through, against plain STG.128 fills of the same bytes?
// One persistent CTA per SM, `nbuf` shared-memory row buffers of `row_bytes`, `nissue` issuing threads (one per warp),
// rows written variant-major like the product ([V, N, 4096] float32: consecutive rows of a CTA are N * 16 KB apart)
// or back to back.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store tma_store.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) fill_kernel(float4* out, size_t n16) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

__global__ void __launch_bounds__(256) fill_cs_kernel(float4* out, size_t n16) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, v);
}

// CTA c takes "items" c, c + G, ...; an item is V consecutive rows (variants) of one sequence.  Issuing thread w (lane 0 of warp w)
// owns NBW buffers and writes variants w, w + nissue, ...
template <int NBW>
__global__ void __launch_bounds__(128) tma_store_kernel(unsigned char* out, long long n_items, int V, int row_bytes, int nissue, int variant_major,
                                                         int hint) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    for (int i = tid; i < NBW * nissue * (row_bytes / 4); i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int w = tid >> 5;
    if ((tid & 31) != 0 || w >= nissue) return;
    uint64_t policy = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(policy));
    long long cnt = 0;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (int v = w; v < V; v += nissue, ++cnt) {
            unsigned char* src = smem + (size_t)(w * NBW + (int)(cnt % NBW)) * row_bytes;
            const size_t row = variant_major ? (size_t)v * n_items + item : (size_t)item * V + v;
            unsigned char* dst = out + row * (size_t)row_bytes;
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBW - 1) : "memory");   // this buffer's previous copy has been read
            if (hint)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)), "r"(row_bytes), "l"(policy) : "memory");
            else
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(row_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NBW>
static float run_tma(unsigned char* out, long long n_items, int V, int row_bytes, int nissue, int vm, int hint, int grid, int reps) {
    const int smem = NBW * nissue * row_bytes;
    CK(cudaFuncSetAttribute(tma_store_kernel<NBW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    tma_store_kernel<NBW><<<grid, 128, smem>>>(out, n_items, V, row_bytes, nissue, vm, hint);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        tma_store_kernel<NBW><<<grid, 128, smem>>>(out, n_items, V, row_bytes, nissue, vm, hint);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    const long long n_items = argc > 1 ? atoll(argv[1]) : 50000;   // sequences
    const int V = 51;
    const size_t bytes = (size_t)n_items * V * 16384;
    unsigned char* out;
    CK(cudaMalloc(&out, bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d, %.2f GB per pass\n", sms, bytes / 1e9);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int which = 0; which < 2; ++which) {
        for (int mult : {4, 8, 16, 32}) {
            float best = 1e30f;
            for (int r = 0; r < 4; ++r) {
                CK(cudaEventRecord(e0));
                if (which == 0) fill_kernel<<<sms * mult, 256>>>(reinterpret_cast<float4*>(out), bytes / 16);
                else fill_cs_kernel<<<sms * mult, 256>>>(reinterpret_cast<float4*>(out), bytes / 16);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r > 0 && ms < best) best = ms;
            }
            printf("%-10s grid %4d x 256: %.3f ms = %.0f GB/s\n", which ? "STG.cs" : "STG", sms * mult, best, bytes / best / 1e6);
        }
    }
    CK(cudaMemsetAsync(out, 0, bytes));
    {
        CK(cudaEventRecord(e0));
        CK(cudaMemsetAsync(out, 0, bytes));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("cudaMemset: %.3f ms = %.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    struct Cfg { int nbw, row, nissue, vm, hint, gmul; };   // nbw = buffers per issuing thread
    const Cfg cfgs[] = {
        {4, 16384, 1, 1, 0, 1}, {8, 16384, 1, 1, 0, 1}, {12, 16384, 1, 1, 0, 1}, {2, 16384, 1, 1, 0, 1}, {1, 16384, 1, 1, 0, 1},
        {4, 16384, 1, 0, 0, 1}, {8, 16384, 1, 0, 0, 1},
        {4, 16384, 1, 1, 1, 1}, {4, 16384, 1, 1, 2, 1}, {8, 16384, 1, 1, 1, 1},
        {4, 16384, 2, 1, 0, 1}, {2, 16384, 4, 1, 0, 1}, {2, 16384, 2, 1, 0, 1}, {1, 16384, 4, 1, 0, 1},
        {8, 8192, 1, 1, 0, 1}, {12, 8192, 1, 1, 0, 1}, {4, 8192, 1, 1, 0, 1},
        {4, 32768, 1, 1, 0, 1},
        {4, 16384, 1, 1, 0, 2}, {2, 16384, 1, 1, 0, 2}, {4, 16384, 1, 1, 0, 3},
        {1, 4096, 1, 1, 0, 1}, {1, 8192, 1, 1, 0, 1}, {1, 32768, 1, 1, 0, 1}, {1, 65536, 1, 1, 0, 1}, {1, 16384, 1, 0, 0, 1}, {1, 16384, 1, 1, 1, 1},
        {1, 16384, 2, 1, 0, 1}, {1, 16384, 1, 1, 0, 2}, {1, 16384, 1, 1, 0, 4}, {1, 8192, 2, 1, 0, 1}, {1, 8192, 1, 1, 0, 2}, {1, 16384, 1, 1, 0, 1},
    };
    for (const Cfg& c : cfgs) {
        float ms = 0.f;
        const long long items = n_items * 16384 / c.row;
        const int grid = sms * c.gmul;
        switch (c.nbw) {
            case 1: ms = run_tma<1>(out, items, V, c.row, c.nissue, c.vm, c.hint, grid, 3); break;
            case 2: ms = run_tma<2>(out, items, V, c.row, c.nissue, c.vm, c.hint, grid, 3); break;
            case 4: ms = run_tma<4>(out, items, V, c.row, c.nissue, c.vm, c.hint, grid, 3); break;
            case 8: ms = run_tma<8>(out, items, V, c.row, c.nissue, c.vm, c.hint, grid, 3); break;
            case 12: ms = run_tma<12>(out, items, V, c.row, c.nissue, c.vm, c.hint, grid, 3); break;
        }
        printf("TMA bufs/thread %2d row %5d issuing threads %d %s hint %d ctas/SM %d: %.3f ms = %.0f GB/s\n", c.nbw, c.row, c.nissue,
               c.vm ? "variant-major" : "contiguous   ", c.hint, c.gmul, ms, (double)items * V * c.row / ms / 1e6);
    }
    return 0;
}
