// membench.cu -- read-only HBM streaming microbenchmark (sm_100a): what can one SM / the chip pull?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// (1) plain LDG.128, U independent loads per thread in flight
template <int U>
__global__ void k_ldg(const uint4* __restrict__ p, size_t n16, unsigned* out) {
    size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (size_t i = tid; i + (U - 1) * nth < n16; i += U * nth) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(p + i + u * nth);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) *out = acc;
}

// (2) cp.async ring: each CTA streams contiguous CHUNK-byte blocks (block b -> CTA b % grid) into S stages
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int S, int CHUNK, int THREADS>
__global__ void __launch_bounds__(THREADS) k_cpasync(const unsigned char* __restrict__ p, size_t nblocks, unsigned* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    constexpr int PER = CHUNK / 16 / THREADS;
    size_t b = blockIdx.x;
    int issued = 0;
    // prologue
    for (int s = 0; s < S - 1; ++s, b += gridDim.x) {
        if (b < nblocks) {
#pragma unroll
            for (int j = 0; j < PER; ++j) cp16(sbase + s * CHUNK + (j * THREADS + threadIdx.x) * 16, p + b * CHUNK + (size_t)(j * THREADS + threadIdx.x) * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++issued;
    }
    int stage = S - 1;
    for (; b < nblocks + (size_t)(S - 1) * gridDim.x; b += gridDim.x) {
        if (b < nblocks) {
#pragma unroll
            for (int j = 0; j < PER; ++j) cp16(sbase + stage * CHUNK + (j * THREADS + threadIdx.x) * 16, p + b * CHUNK + (size_t)(j * THREADS + threadIdx.x) * 16);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
        stage = (stage + 1 == S) ? 0 : stage + 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (smem[threadIdx.x] == 0x5a && issued == -1) *out = 1;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t bytes = (size_t)1536 << 20;
    unsigned char* d; unsigned* out;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(d, 1, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, buffer %.0f MB\n", sms, bytes / 1e6);
    auto rep = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    const size_t n16 = bytes / 16;
    for (int bps : {1, 2, 4}) {
        char nm[96];
        snprintf(nm, 96, "ldg U=4  grid=%dxSM block=512", bps);
        rep(nm, time_ms([&] { k_ldg<4><<<sms * bps, 512>>>((const uint4*)d, n16, out); }, 20));
        snprintf(nm, 96, "ldg U=8  grid=%dxSM block=512", bps);
        rep(nm, time_ms([&] { k_ldg<8><<<sms * bps, 512>>>((const uint4*)d, n16, out); }, 20));
    }
    rep("ldg U=8 grid=37 block=512", time_ms([&] { k_ldg<8><<<37, 512>>>((const uint4*)d, n16, out); }, 5));
    rep("ldg U=8 grid=37 block=1024", time_ms([&] { k_ldg<8><<<37, 1024>>>((const uint4*)d, n16, out); }, 5));
    {
        constexpr int CH = 16384;
        const size_t nb = bytes / CH;
        auto run = [&](auto kern, int S, int threads, int grid, const char* nm) {
            size_t smem = (size_t)S * CH;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rep(nm, time_ms([&] { kern<<<grid, threads, smem>>>(d, nb, out); }, 20));
            CK(cudaGetLastError());
        };
        run(k_cpasync<4, CH, 128>, 4, 128, sms, "cp.async S=4  x16KB 128thr grid=SM");
        run(k_cpasync<8, CH, 128>, 8, 128, sms, "cp.async S=8  x16KB 128thr grid=SM");
        run(k_cpasync<12, CH, 128>, 12, 128, sms, "cp.async S=12 x16KB 128thr grid=SM");
        run(k_cpasync<4, CH, 128>, 4, 128, sms * 2, "cp.async S=4  x16KB 128thr grid=2xSM");
        run(k_cpasync<6, CH, 128>, 6, 128, sms * 2, "cp.async S=6  x16KB 128thr grid=2xSM");
        run(k_cpasync<3, CH, 128>, 3, 128, sms * 4, "cp.async S=3  x16KB 128thr grid=4xSM");
        run(k_cpasync<8, CH, 512>, 8, 512, sms, "cp.async S=8  x16KB 512thr grid=SM");
        run(k_cpasync<8, CH, 128>, 8, 128, 37, "cp.async S=8  x16KB 128thr grid=37");
    }
    printf("done\n");
    return 0;
}
