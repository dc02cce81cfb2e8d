// scripts/microbench.cu -- measures the denominators the insert/query kernels are judged against
// on the box they run on (SURVEY.md section 8d "R_rand"):
//   * independent random 32-bit RED.OR / ATOM.OR / CAS over a footprint F  (G ops/s)
//   * independent random 32 B-sector loads over a footprint F             (G loads/s)
//   * shared-memory atomicOr rate                                          (G ops/s)
//   * streaming read+write copy                                            (GB/s)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench scripts/microbench.cu
// Output: one JSON line per measurement (timed with CUDA events, 3 warm-ups, best of 5).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {  // splitmix64
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// MODE 0: RED.OR (no return), 1: ATOM.OR (return consumed), 2: CAS loop byte increment, 3: 32-bit load
template <int MODE, int PER>
__global__ void __launch_bounds__(256) k_rand(uint32_t* tbl, uint64_t n_words, uint64_t n_ops, uint64_t seed, unsigned long long* sink) {
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = tid; i < n_ops; i += stride * PER) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            uint64_t r = mix(seed + i + k * stride);
            uint64_t w = ((r >> 8) * (unsigned __int128)n_words) >> 56;  // uniform in [0, n_words)
            uint32_t mask = 1u << (r & 31);
            if (MODE == 0) atomicOr(tbl + w, mask);
            else if (MODE == 1) acc += atomicOr(tbl + w, mask) & mask ? 1 : 0;
            else if (MODE == 2) {
                unsigned sh = (r & 3) * 8;
                uint32_t old = __ldcg(tbl + w);
                while (true) {
                    if (((old >> sh) & 255u) == 255u) break;
                    uint32_t prev = atomicCAS(tbl + w, old, old + (1u << sh));
                    if (prev == old) break;
                    old = prev;
                }
            } else acc += __ldg(tbl + w) & mask ? 1 : 0;
        }
    }
    if (MODE == 1 || MODE == 3) { if (acc == 0xffffffffu) atomicAdd(sink, 1ull); }
}

__global__ void __launch_bounds__(256) k_smem_atomic(uint64_t n_ops_per_thread, uint64_t seed, unsigned long long* sink) {
    extern __shared__ uint32_t sm[];
    const int words = 32768;  // 128 KB
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint64_t x = seed + blockIdx.x * 1315423911ull + threadIdx.x;
    for (uint64_t i = 0; i < n_ops_per_thread; ++i) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        atomicOr(&sm[(x >> 40) & (words - 1)], 1u << ((x >> 33) & 31));
    }
    __syncthreads();
    if (sm[threadIdx.x] == 0xdeadbeef) atomicAdd(sink, 1ull);
}

__global__ void __launch_bounds__(256) k_copy(const uint4* __restrict__ a, uint4* __restrict__ b, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F>
static float best_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d}\n", p.name, sms, p.l2CacheSize);
    unsigned long long* sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    const uint64_t max_bytes = 4ull << 30;
    uint32_t* tbl; CK(cudaMalloc(&tbl, max_bytes)); CK(cudaMemset(tbl, 0, max_bytes));
    const uint64_t n_ops = 1ull << 31;
    int grid = sms * 8;
    const uint64_t foots[] = {1ull << 20, 16ull << 20, 32ull << 20, 64ull << 20, 128ull << 20, 512ull << 20, 1ull << 30, 4ull << 30};
    const char* names[] = {"red_or", "atom_or", "cas_byte", "load32"};
    for (uint64_t F : foots) {
        uint64_t nw = F / 4;
        for (int mode = 0; mode < 4; ++mode) {
            uint64_t seed = 12345 + mode;
            float ms;
            if (mode == 0) ms = best_ms([&] { k_rand<0, 8><<<grid, 256>>>(tbl, nw, n_ops, seed, sink); });
            else if (mode == 1) ms = best_ms([&] { k_rand<1, 8><<<grid, 256>>>(tbl, nw, n_ops, seed, sink); });
            else if (mode == 2) { CK(cudaMemset(tbl, 0, F)); ms = best_ms([&] { k_rand<2, 8><<<grid, 256>>>(tbl, nw, n_ops, seed, sink); }); }
            else ms = best_ms([&] { k_rand<3, 8><<<grid, 256>>>(tbl, nw, n_ops, seed, sink); });
            CK(cudaGetLastError());
            printf("{\"bench\": \"%s\", \"footprint_mb\": %llu, \"gops\": %.2f, \"ms\": %.3f}\n", names[mode],
                   (unsigned long long)(F >> 20), n_ops / (ms * 1e6), ms);
            fflush(stdout);
        }
    }
    {
        CK(cudaFuncSetAttribute(k_smem_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        uint64_t per = 1 << 14;
        float ms = best_ms([&] { k_smem_atomic<<<sms, 256, 131072>>>(per, 7, sink); });
        printf("{\"bench\": \"smem_atomic_or_256thr\", \"gops\": %.2f, \"ms\": %.3f}\n", (double)sms * 256 * per / (ms * 1e6), ms);
        ms = best_ms([&] { k_smem_atomic<<<sms, 1024, 131072>>>(per / 4, 7, sink); });
        printf("{\"bench\": \"smem_atomic_or_1024thr\", \"gops\": %.2f, \"ms\": %.3f}\n", (double)sms * 1024 * (per / 4) / (ms * 1e6), ms);
    }
    {
        uint64_t n = (2ull << 30) / 16;
        uint4* a = (uint4*)tbl; uint4* b = a + n;
        float ms = best_ms([&] { k_copy<<<sms * 16, 256>>>(a, b, n); });
        printf("{\"bench\": \"copy_rw\", \"gbs\": %.1f, \"ms\": %.3f}\n", 2.0 * n * 16 / (ms * 1e6), ms);
    }
    return 0;
}
