// microbench2.cu — second set of measured rates for the fused-kernel design:
// shared-memory bandwidth (volatile asm so nothing is hoisted), uniform-address
// (broadcast) LDS.128, register-indexed constant-bank loads (LDC.64), and the
// instruction-cache cost of warp-role-specialised straight-line code.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;

__device__ __forceinline__ float lds32(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ float2 lds64(unsigned a) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds128(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned a, float2 v) { asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(a), "f"(v.x), "f"(v.y)); }
__device__ __forceinline__ void sts32(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v)); }

// MODE 0: LDS.32 lane-contiguous; 1: LDS.64; 2: LDS.128; 3: LDS.128 uniform address; 4: STS.64; 5: STS.32
template <int MODE>
__global__ void k_smem(float *out, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(sm));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned row = ((it + i * 3 + warp) & 15);
            if (MODE == 0) acc += lds32(base + (row * 32 + lane) * 4);
            if (MODE == 1) { float2 v = lds64(base + (row * 32 + lane) * 8); acc += v.x + v.y; }
            if (MODE == 2) { float4 v = lds128(base + (row * 32 + lane) * 16); acc += v.x + v.w; }
            if (MODE == 3) { float4 v = lds128(base + (row * 32) * 16); acc += v.x + v.w; }
            if (MODE == 4) sts64(base + (row * 32 + lane) * 8 + warp * 4096, make_float2(acc, (float)it));
            if (MODE == 5) sts32(base + (row * 32 + lane) * 4 + warp * 2048, acc);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

struct Tab { float2 w[1024]; };
// register-indexed constant loads: every lane of a warp uses the same index
template <int FMA_PER_LDC>
__global__ void k_ldc(float *out, long long *cyc, const __grid_constant__ Tab tab) {
    const int warp = threadIdx.x >> 5;
    float acc0 = threadIdx.x, acc1 = 1.0f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 w = tab.w[(it * 8 + i + warp * 37) & 1023];
            acc0 = fmaf(acc0, w.x, w.y);
#pragma unroll
            for (int j = 1; j < FMA_PER_LDC; ++j) acc1 = fmaf(acc1, 1.0001f, 0.5f + j);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// I-cache: ROLES different straight-line bodies of BODY FFMAs each, one per warp (warp % ROLES).
template <int ROLE, int BODY>
__device__ __forceinline__ void body(float (&x)[8]) {
#pragma unroll
    for (int i = 0; i < BODY; ++i) {
        // distinct immediates per (ROLE, i) so that bodies cannot be merged
        x[i & 7] = fmaf(x[i & 7], 0.5f + (ROLE * 4096 + i) * 1e-7f, x[(i + 3) & 7]);
    }
}
template <int ROLES, int BODY>
__global__ void __launch_bounds__(256) k_icache(float *out, long long *cyc, int iters) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    const int role = (threadIdx.x >> 5) % ROLES;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        switch (role) {
            case 0: body<0, BODY>(x); break;
            case 1: if (ROLES > 1) body<1, BODY>(x); break;
            case 2: if (ROLES > 2) body<2, BODY>(x); break;
            case 3: if (ROLES > 3) body<3, BODY>(x); break;
            case 4: if (ROLES > 4) body<4, BODY>(x); break;
            case 5: if (ROLES > 5) body<5, BODY>(x); break;
            case 6: if (ROLES > 6) body<6, BODY>(x); break;
            default: if (ROLES > 7) body<7, BODY>(x); break;
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
int run(const char *name, F launch, double lane_ops_per_thread, int sms, int ctas_per_sm, int threads, double bytes_per_op) {
    float *out; long long *cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * threads));
    CK(cudaMalloc(&cyc, sizeof(long long)));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) launch(out, cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(out, cyc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double total = lane_ops_per_thread * threads * ctas_per_sm * sms;
    const double per_clk_sm = total / (best * 1e-3) / (1.965e9 * sms);
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"lane_ops_per_s\": %.4e, \"lane_ops_per_clk_per_sm@1965\": %.2f, \"bytes_per_clk_per_sm\": %.1f}\n",
           name, best, total / (best * 1e-3), per_clk_sm, per_clk_sm * bytes_per_op);
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int T = 256, C = 2;
    const double n = (double)ITERS * 8;
    CK(cudaFuncSetAttribute(k_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    auto S = [&](auto kern, int smem) { return [=](float *o, long long *c) { kern<<<sms * C, T, smem>>>(o, c); }; };
    run("lds32 contiguous", S(k_smem<0>, 32768), n, sms, C, T, 4);
    run("lds64 contiguous", S(k_smem<1>, 32768), n, sms, C, T, 8);
    run("lds128 contiguous", S(k_smem<2>, 32768), n, sms, C, T, 16);
    run("lds128 uniform address (broadcast)", S(k_smem<3>, 32768), n, sms, C, T, 16);
    run("sts64 contiguous", S(k_smem<4>, 65536), n, sms, C, T, 8);
    run("sts32 contiguous", S(k_smem<5>, 32768), n, sms, C, T, 4);
    Tab *tab = new Tab;
    for (int i = 0; i < 1024; ++i) tab->w[i] = make_float2(1.0f + i * 1e-6f, 0.001f * i);
    Tab tv = *tab;
    run("ldc64 indexed + 1 fma", [=](float *o, long long *c) { k_ldc<1><<<sms * C, T>>>(o, c, tv); }, n, sms, C, T, 8);
    run("ldc64 indexed + 4 fma (lane-ldc)", [=](float *o, long long *c) { k_ldc<4><<<sms * C, T>>>(o, c, tv); }, n, sms, C, T, 8);
    run("ldc64 indexed + 8 fma (lane-ldc)", [=](float *o, long long *c) { k_ldc<8><<<sms * C, T>>>(o, c, tv); }, n, sms, C, T, 8);
    // I-cache: total FFMAs per thread constant (= BODY * iters)
    const double f = 1200.0 * 64;
    run("icache 1 role x 1200 ffma (19 KB)", [=](float *o, long long *c) { k_icache<1, 1200><<<sms * C, T>>>(o, c, 64); }, f, sms, C, T, 0);
    run("icache 2 roles x 1200 ffma (38 KB)", [=](float *o, long long *c) { k_icache<2, 1200><<<sms * C, T>>>(o, c, 64); }, f, sms, C, T, 0);
    run("icache 4 roles x 1200 ffma (77 KB)", [=](float *o, long long *c) { k_icache<4, 1200><<<sms * C, T>>>(o, c, 64); }, f, sms, C, T, 0);
    run("icache 8 roles x 1200 ffma (154 KB)", [=](float *o, long long *c) { k_icache<8, 1200><<<sms * C, T>>>(o, c, 64); }, f, sms, C, T, 0);
    run("icache 8 roles x 600 ffma (77 KB)", [=](float *o, long long *c) { k_icache<8, 600><<<sms * C, T>>>(o, c, 128); }, f, sms, C, T, 0);
    run("icache 8 roles x 300 ffma (38 KB)", [=](float *o, long long *c) { k_icache<8, 300><<<sms * C, T>>>(o, c, 256); }, f, sms, C, T, 0);
    return 0;
}
