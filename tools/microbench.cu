// microbench.cu — measured pipe rates on the B200 that the fused-kernel design
// depends on (DESIGN.md "Measured pipe rates"): FP32 FMA peak (the roofline's
// FP32 denominator, SURVEY.md §6.2 "not measured"), packed f32x2 math, shared
// memory, shuffles, MUFU.  Prints one JSON object per line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__global__ void k_ffma(float *out, float a, float b, long long *cyc) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fadd(float *out, float a, float b, long long *cyc) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = x[i] + a;
    long long t1 = clock64();
    float s = b; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ffma2(float *out, float a, float b, long long *cyc) {
    float2 x[ILP];
    const float2 aa = make_float2(a, a * 1.01f), bb = make_float2(b, b * 0.99f);
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fadd2(float *out, float a, float b, long long *cyc) {
    float2 x[ILP];
    const float2 aa = make_float2(a, a * 1.01f);
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __fadd2_rn(x[i], aa);
    long long t1 = clock64();
    float s = b; for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// half FFMA2, half FADD2 (the butterfly mix)
__global__ void k_mix2(float *out, float a, float b, long long *cyc) {
    float2 x[ILP];
    const float2 aa = make_float2(a, a * 1.01f), bb = make_float2(b, b * 0.99f);
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; i += 2) { x[i] = __ffma2_rn(x[i], aa, bb); x[i + 1] = __fadd2_rn(x[i + 1], aa); }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// scalar FFMA interleaved with conflict-free LDS.32 (1 LDS per `ratio` FFMA)
template <int RATIO>
__global__ void k_ffma_lds(float *out, float a, float b, long long *cyc) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 1e-4f;
    __syncthreads();
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    int idx = threadIdx.x;
    float acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            x[i] = fmaf(x[i], a, b);
            if (i % RATIO == 0) { acc += sm[idx]; idx = (idx + 256) & 2047; }
        }
    }
    long long t1 = clock64();
    float s = acc; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// packed FFMA2 interleaved with LDS.64
template <int RATIO>
__global__ void k_ffma2_lds(float *out, float a, float b, long long *cyc) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(i * 1e-4f, i);
    __syncthreads();
    float2 x[ILP];
    const float2 aa = make_float2(a, a * 1.01f), bb = make_float2(b, b * 0.99f);
    for (int i = 0; i < ILP; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    int idx = threadIdx.x;
    float2 acc = make_float2(0, 0);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            x[i] = __ffma2_rn(x[i], aa, bb);
            if (i % RATIO == 0) { acc = __fadd2_rn(acc, sm[idx]); idx = (idx + 256) & 2047; }
        }
    }
    long long t1 = clock64();
    float s = acc.x + acc.y; for (int i = 0; i < ILP; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <typename T>
__global__ void k_lds(float *out, long long *cyc) {
    __shared__ T sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { T v; memset(&v, 0, sizeof(T)); sm[i] = v; }
    __syncthreads();
    int idx = threadIdx.x;
    float acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            T v = sm[idx];
            acc += *reinterpret_cast<float *>(&v);
            idx = (idx + 256) & 1023;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(float *out, long long *cyc) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (i & 15));
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lg2(float *out, long long *cyc) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i + 2.0f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __log2f(x[i]) + 3.0f;
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_i2f(float *out, int seed, long long *cyc) {
    int x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 7 + i + seed;
    float acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) { acc += static_cast<float>(static_cast<short>(x[i])); x[i] += 13; }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
int run(const char *name, F launch, double lane_ops_per_thread, int sms, int ctas_per_sm, int threads, const char *unit) {
    float *out; long long *cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * threads));
    CK(cudaMalloc(&cyc, sizeof(long long)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch(out, cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f; long long c = 0;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(out, cyc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) { best = ms; CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost)); }
    }
    CK(cudaGetLastError());
    const double total = lane_ops_per_thread * threads * ctas_per_sm * sms;
    // (the clock64 pair inside the kernels is not ordered against the arithmetic, so it is not reported: the rates come
    // from event time; the FP32 peak used by bench.py is measured by tools/fp32_peak.cu over >= 100 ms)
    (void)c;
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"%s_per_s\": %.4e, \"ctas_per_sm\": %d, \"threads\": %d}\n",
           name, best, unit, total / (best * 1e-3), ctas_per_sm, threads);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", prop.name, sms, prop.major, prop.minor, prop.clockRate);
    const int T = 256, C = 4;  // 4 CTAs x 256 threads = 32 warps / SM
    const double n = (double)ITERS * ILP;
    auto L = [&](auto kern, auto... args) { return [=](float *o, long long *c) { kern<<<sms * C, T>>>(o, args..., c); }; };
    run("ffma_scalar (lane-FMAs)", L(k_ffma, 1.0001f, 0.5f), n, sms, C, T, "fma");
    run("fadd_scalar (lane-adds)", L(k_fadd, 1.0001f, 0.5f), n, sms, C, T, "add");
    run("ffma2_packed (lane-FMAs, 2 per instr)", L(k_ffma2, 1.0001f, 0.5f), 2 * n, sms, C, T, "fma");
    run("fadd2_packed (lane-adds, 2 per instr)", L(k_fadd2, 1.0001f, 0.5f), 2 * n, sms, C, T, "add");
    run("mix_ffma2_fadd2 (lane-ops)", L(k_mix2, 1.0001f, 0.5f), 2 * n, sms, C, T, "op");
    run("ffma + LDS.32 every 1 (lane-FMAs)", L(k_ffma_lds<1>, 1.0001f, 0.5f), n, sms, C, T, "fma");
    run("ffma + LDS.32 every 2 (lane-FMAs)", L(k_ffma_lds<2>, 1.0001f, 0.5f), n, sms, C, T, "fma");
    run("ffma + LDS.32 every 4 (lane-FMAs)", L(k_ffma_lds<4>, 1.0001f, 0.5f), n, sms, C, T, "fma");
    run("ffma2 + LDS.64 every 1 (lane-FMAs)", L(k_ffma2_lds<1>, 1.0001f, 0.5f), 2 * n, sms, C, T, "fma");
    run("ffma2 + LDS.64 every 2 (lane-FMAs)", L(k_ffma2_lds<2>, 1.0001f, 0.5f), 2 * n, sms, C, T, "fma");
    run("ffma2 + LDS.64 every 4 (lane-FMAs)", L(k_ffma2_lds<4>, 1.0001f, 0.5f), 2 * n, sms, C, T, "fma");
    run("lds32 (lane-loads)", L(k_lds<float>), n, sms, C, T, "ld");
    run("lds64 (lane-loads)", L(k_lds<float2>), n, sms, C, T, "ld");
    run("lds128 (lane-loads)", L(k_lds<float4>), n, sms, C, T, "ld");
    run("shfl_xor (lane-shuffles)", L(k_shfl), n, sms, C, T, "shfl");
    run("mufu_lg2 (lane-ops)", L(k_lg2), n, sms, C, T, "op");
    run("i2f_s16 (lane-converts)", L(k_i2f, 3), n, sms, C, T, "cvt");
    return 0;
}
