// fp32_peak.cu — the FP32 FMA peak of this GPU, the denominator of bench.py's `roofline` (the round-1
// figure came from a 0.14 ms kernel with launch and ramp inside the timed region, and a broken cycle counter).
//
// One kernel of pure FFMA (8 independent chains per thread, 32 warps per SM, every SM), sized so that ONE launch runs
// >= 5 ms; 3 warm-up launches, then 20 launches back to back between two events: >= 100 ms of steady load, launch gaps
// < 0.1 %.  Prints one JSON line: TFLOP/s (2 FLOP per FMA), lane-FMAs per clock per SM against the clock the driver
// reports as the maximum AND (with --sm-mhz X, given by bench.py from its nvidia-smi samples under load) against the
// sampled clock, so that a reader sees how far the measured peak sits below SMs x 128 x 2 x clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_peak tools/fp32_peak.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int ILP = 8;

__global__ void __launch_bounds__(256) ffma_kernel(float *out, float a, float b, int iters)
{
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char **argv)
{
    double sm_mhz = 0.0;
    for (int i = 1; i + 1 < argc; ++i)
        if (std::strcmp(argv[i], "--sm-mhz") == 0) sm_mhz = std::atof(argv[i + 1]);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, threads = 256, ctas_per_sm = 4, launches = 20;
    float *out = nullptr;
    CK(cudaMalloc(&out, sizeof(float) * sms * ctas_per_sm * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    // size one launch to >= 5 ms: calibrate on a short launch first
    int iters = 2048;
    float ms = 0.0f;
    for (int pass = 0; pass < 2; ++pass) {
        ffma_kernel<<<sms * ctas_per_sm, threads>>>(out, 1.0001f, 0.5f, iters);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        ffma_kernel<<<sms * ctas_per_sm, threads>>>(out, 1.0001f, 0.5f, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (pass == 0) iters = static_cast<int>(iters * (6.0f / ms)) + 1;
    }
    for (int w = 0; w < 3; ++w) ffma_kernel<<<sms * ctas_per_sm, threads>>>(out, 1.0001f, 0.5f, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int l = 0; l < launches; ++l) ffma_kernel<<<sms * ctas_per_sm, threads>>>(out, 1.0001f, 0.5f, iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    const double fma = static_cast<double>(iters) * 8 * ILP * threads * ctas_per_sm * sms * launches;
    const double fma_per_s = fma / (ms * 1e-3);
    const double max_mhz = prop.clockRate / 1e3;
    const double derived_max = sms * 128.0 * 2.0 * max_mhz * 1e6 / 1e12;
    printf("{\"bench\": \"fp32_peak\", \"device\": \"%s\", \"sms\": %d, \"launches\": %d, \"ms_per_launch\": %.3f, \"tflops\": %.3f, "
           "\"max_sm_mhz\": %.0f, \"derived_tflops_at_max_clock\": %.2f, \"frac_of_derived_at_max_clock\": %.4f, "
           "\"lane_fma_per_clk_per_sm_at_max_clock\": %.2f",
           prop.name, sms, launches, ms / launches, 2.0 * fma_per_s / 1e12, max_mhz, derived_max,
           2.0 * fma_per_s / 1e12 / derived_max, fma_per_s / (sms * max_mhz * 1e6));
    if (sm_mhz > 0.0)
        printf(", \"sampled_sm_mhz\": %.0f, \"frac_of_derived_at_sampled_clock\": %.4f, \"lane_fma_per_clk_per_sm_at_sampled_clock\": %.2f",
               sm_mhz, 2.0 * fma_per_s / 1e12 / (sms * 128.0 * 2.0 * sm_mhz * 1e6 / 1e12), fma_per_s / (sms * sm_mhz * 1e6));
    printf("}\n");
    cudaFree(out);
    return 0;
}
