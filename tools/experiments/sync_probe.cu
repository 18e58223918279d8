// Probe: cost of a grid-wide barrier inside a persistent kernel vs. a dependent-launch edge in a CUDA graph (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_probe sync_probe.cu && ./sync_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

struct Bar { unsigned int count; unsigned int gen; };

__device__ __forceinline__ void grid_barrier(Bar* b, unsigned int n) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&b->gen) : "memory");
        __threadfence();
        if (atomicAdd(&b->count, 1u) == n - 1) {
            b->count = 0;
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&b->gen), "r"(gen + 1) : "memory");
        } else {
            unsigned int g2;
            long long t0 = clock64();
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g2) : "l"(&b->gen) : "memory");
                if (clock64() - t0 > 2000000000LL) __trap();
            } while (g2 == gen);
        }
        __threadfence();
    }
    __syncthreads();
}

// monotonic-counter variant: one atomic, everyone polls the counter (no reset, no second store)
__device__ __forceinline__ void grid_barrier_mono(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1ull);
        unsigned long long v;
        long long t0 = clock64();
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
            if (clock64() - t0 > 2000000000LL) __trap();
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 1) k_bar(Bar* b, int iters, float* data, int check) {
    extern __shared__ char smem[];
    for (int i = 0; i < iters; ++i) {
        if (check) {
            // every CTA writes its slot, then after the barrier reads its neighbour's slot
            if (threadIdx.x == 0) data[blockIdx.x] = (float)(i + 1);
        }
        grid_barrier(b, gridDim.x);
        if (check) {
            if (threadIdx.x == 0) {
                float v = data[(blockIdx.x + 37) % gridDim.x];
                if (v != (float)(i + 1)) { printf("stale read cta %d it %d got %f\n", blockIdx.x, i, v); __trap(); }
            }
            grid_barrier(b, gridDim.x);
        }
    }
}
__global__ void __launch_bounds__(256, 1) k_bar_mono(unsigned long long* ctr, unsigned long long base, int iters) {
    extern __shared__ char smem[];
    for (int i = 0; i < iters; ++i) grid_barrier_mono(ctr, base + (unsigned long long)(i + 1) * gridDim.x);
}
__global__ void __launch_bounds__(256, 1) k_empty(float* p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.f;
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, sms);
    Bar* b; CK(cudaMalloc(&b, sizeof(Bar))); CK(cudaMemset(b, 0, sizeof(Bar)));
    unsigned long long* ctr; CK(cudaMalloc(&ctr, 8)); CK(cudaMemset(ctr, 0, 8));
    float* data; CK(cudaMalloc(&data, 4096)); CK(cudaMemset(data, 0, 4096));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t smem = 200 * 1024;
    CK(cudaFuncSetAttribute(k_bar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_bar_mono, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int grid : {sms, sms / 2, 16}) {
        for (int check = 0; check < 2; ++check) {
            int iters = 2000;
            void* args[] = {&b, &iters, &data, &check};
            CK(cudaLaunchCooperativeKernel((void*)k_bar, dim3(grid), dim3(256), args, smem, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(e0, st));
            CK(cudaLaunchCooperativeKernel((void*)k_bar, dim3(grid), dim3(256), args, smem, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("grid barrier (count+gen), %3d CTAs, check=%d: %.3f us per barrier\n", grid, check, 1000.f * ms / (iters * (check ? 2 : 1)));
        }
        {
            int iters = 2000;
            unsigned long long base = 0;
            CK(cudaMemset(ctr, 0, 8));
            void* args[] = {&ctr, &base, &iters};
            CK(cudaLaunchCooperativeKernel((void*)k_bar_mono, dim3(grid), dim3(256), args, smem, st));
            CK(cudaStreamSynchronize(st));
            base = (unsigned long long)iters * grid;
            CK(cudaEventRecord(e0, st));
            CK(cudaLaunchCooperativeKernel((void*)k_bar_mono, dim3(grid), dim3(256), args, smem, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("grid barrier (monotonic),  %3d CTAs: %.3f us per barrier\n", grid, 1000.f * ms / iters);
        }
    }
    // dependent-launch chain in a graph
    for (int pdl = 0; pdl < 2; ++pdl) {
        for (int grid : {sms, 64}) {
            const int n = 1000;
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int i = 0; i < n; ++i) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = pdl;
                CK(cudaLaunchKernelEx(&cfg, k_empty, data));
            }
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(e0, st));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("graph chain of empty 200KB-smem kernels, %3d CTAs, pdl=%d: %.3f us per launch\n", grid, pdl, 1000.f * ms / n);
            cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
        }
    }
    return 0;
}
