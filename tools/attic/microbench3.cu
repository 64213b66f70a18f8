// microbench3.cu — third set of measured rates behind the fused-kernel design:
//   * TMEM as a per-lane scratchpad: tcgen05.st / tcgen05.ld 32x32b.x32 throughput,
//     alone and next to FFMA2 / LDS traffic (is it a separate pipe?);
//   * issue-slot sharing: FFMA2 (2 pipe-clocks, 1 issue slot) interleaved with ALU-pipe
//     integer instructions and with LDS.64;
//   * instruction-cache behaviour of ONE long straight-line body shared by every warp;
//   * shared-memory bandwidth per access width, long runs.
// Prints one JSON object per line.  All rates are derived from CUDA-event times.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench3 tools/microbench3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %d\"}\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// ------------------------------------------------------------------ TMEM helpers
#define R32(v) "%" #v
#define OUT8(a, o) "=r"(a[o+0]), "=r"(a[o+1]), "=r"(a[o+2]), "=r"(a[o+3]), "=r"(a[o+4]), "=r"(a[o+5]), "=r"(a[o+6]), "=r"(a[o+7])
#define IN8(a, o) "r"(a[o+0]), "r"(a[o+1]), "r"(a[o+2]), "r"(a[o+3]), "r"(a[o+4]), "r"(a[o+5]), "r"(a[o+6]), "r"(a[o+7])

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : OUT8(r, 0), OUT8(r, 8), OUT8(r, 16), OUT8(r, 24)
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
        :: IN8(r, 0), IN8(r, 8), IN8(r, 16), IN8(r, 24), "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t *slot)
{
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                     :: "r"(static_cast<uint32_t>(__cvta_generic_to_shared(slot))));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t base)
{
    __syncthreads();
    if ((threadIdx.x >> 5) == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base));
}

// MODE 0: st only, 1: ld only, 2: st+ld round trip, 3: ld + 32 FFMA2 per ld, 4: ld + 8 LDS.64 per ld,
// 5: the FFMA2 work of mode 3 without TMEM, 6: the LDS work of mode 4 without TMEM
template <int MODE>
__global__ void __launch_bounds__(512) k_tmem(float *out, int iters, int check)
{
    __shared__ uint32_t slot;
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
    const uint32_t base = tmem_alloc_all(&slot);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    const int per = 512 / ((nw + 3) / 4);                 // columns owned by this warp
    const uint32_t t0 = base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>((warp >> 2) * per);
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 64 + i;
    float2 x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    const float2 aa = make_float2(1.0001f, 0.9999f), bb = make_float2(0.5f, 0.25f);
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(sm));
    float2 acc = make_float2(0.f, 0.f);
    // seed the columns so that loads see defined data
    for (int c = 0; c < per; c += 32) tmem_st32(t0 + c, r);
    tmem_wait_st();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = t0 + ((it * 32) % per);
        if (MODE == 0 || MODE == 2) { r[0] += 1; tmem_st32(col, r); }
        if (MODE == 2) tmem_wait_st();
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 4) tmem_ld32(col, r);
        if (MODE == 3 || MODE == 5) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], aa, bb);
        }
        if (MODE == 4 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y)
                             : "r"(sbase + ((((it + i * 5) & 15) * 32 + lane) * 8)));
                acc.x += v.x; acc.y += v.y;
            }
        }
        if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 4) {
            tmem_wait_ld();
            acc.x += __uint_as_float(r[0] ^ r[31]) * 1e-30f;
        }
    }
    tmem_wait_st();
    if (check && MODE == 2) {
        // round trip must return what was stored
        uint32_t q[32];
        tmem_ld32(t0, q);
        tmem_wait_ld();
        // column block 0 was last written at the last `it` with (it*32)%per == 0
        if (q[1] != r[1] && lane == 0 && blockIdx.x == 0) printf("{\"tmem_roundtrip\": \"MISMATCH\"}\n");
    }
    float s = acc.x + acc.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += __uint_as_float(r[i]) * 1e-30f;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    tmem_free_all(base);
}

// ------------------------------------------------------------------ issue sharing
// MODE 0: 8 FFMA2; 1: 8 FFMA2 + 8 ALU (LOP3/IADD3); 2: 8 ALU only; 3: 8 FFMA2 + 8 scalar FFMA;
// 4: 8 FFMA2 + 4 LDS.64; 5: 8 FFMA2 + 8 ALU + 4 LDS.64; 6: 16 scalar FFMA + 8 ALU
template <int MODE>
__global__ void __launch_bounds__(256) k_issue(float *out, int iters, int seed)
{
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(sm));
    float2 x[8];
    float f[16];
    unsigned y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x * 1e-3f + i, i); y[i] = threadIdx.x * 7 + i + seed; }
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x + i;
    const float2 aa = make_float2(1.0001f, 0.9999f), bb = make_float2(0.5f, 0.25f);
    float2 acc = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 1 || MODE == 3 || MODE == 4 || MODE == 5) x[i] = __ffma2_rn(x[i], aa, bb);
            if (MODE == 1 || MODE == 2 || MODE == 5 || MODE == 6)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 1) & 7]), "r"(seed));
            if (MODE == 3) f[i] = fmaf(f[i], 1.0001f, 0.5f);
            if (MODE == 6) { f[i] = fmaf(f[i], 1.0001f, 0.5f); f[i + 8] = fmaf(f[i + 8], 1.0001f, 0.5f); }
            if ((MODE == 4 || MODE == 5) && (i & 1) == 0) {
                float2 v;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y)
                             : "r"(sbase + ((((it + i * 5) & 15) * 32 + lane) * 8)));
                acc.x += v.x; acc.y += v.y;
            }
        }
    }
    float s = acc.x + acc.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + __uint_as_float(y[i] & 0x3fffffffu);
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// ------------------------------------------------------------------ issue sharing, second series
// A: packed op  0 none, 1 FADD2, 2 FMUL2, 3 FFMA2 (3 register operands)
// B: companion  0 none, 1 LOP3, 2 IADD3, 3 FADD, 4 FFMA, 5 LDS.32, 6 LDS.64, 7 MOV-ish (PRMT), 8 second packed chain
template <int A, int B>
__global__ void __launch_bounds__(256) k_issue2(float *out, int iters, int seed)
{
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(sm)) + lane * 8;
    float2 x[8], z[8];
    float f[8];
    unsigned y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = make_float2(threadIdx.x * 1e-3f + i, i); z[i] = make_float2(i, threadIdx.x);
        y[i] = threadIdx.x * 7 + i + seed; f[i] = threadIdx.x + i;
    }
    float2 aa = make_float2(1.0001f + seed, 0.9999f + seed), bb = make_float2(0.5f + seed, 0.25f + seed);
    float fa = 1.0f + seed * 1e-6f, fb = seed;
    float accf = 0.f;
    for (int it = 0; it < iters; ++it) {
        const unsigned bs = sbase + ((it & 1) << 12);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (A == 1) x[i] = __fadd2_rn(x[i], aa);
            if (A == 2) x[i] = __fmul2_rn(x[i], aa);
            if (A == 3) x[i] = __ffma2_rn(x[i], aa, bb);
            if (A == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i].x) : "f"(fa));
            if (A == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i].x) : "f"(fa), "f"(fb));
            if (B == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[(i + 1) & 7]), "r"(seed));
            if (B == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(y[(i + 1) & 7]));
            if (B == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
            if (B == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            if (B == 5) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(bs + i * 256) : "memory"); f[i] += v; }
            if (B == 6) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(bs + i * 256) : "memory"); f[i] += v.x; }
            if (B == 7) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(y[i]) : "r"(y[(i + 1) & 7]));
            if (B == 8) z[i] = __fadd2_rn(z[i], bb);
        }
    }
    float s = accf;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y + z[i].x + z[i].y + f[i] + __uint_as_float(y[i] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------ one long body, every warp
template <int BODY>
__global__ void __launch_bounds__(256) k_long(float *out, int iters)
{
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < BODY; ++i) x[i & 7] = fmaf(x[i & 7], 0.5f + i * 1e-7f, x[(i + 3) & 7]);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ------------------------------------------------------------------ shared memory, long runs
// MODE 0: LDS.32, 1: LDS.64, 2: LDS.128, 3: STS.64, 4: STS.128, 5: LDS.64 uniform address, 6: LDS.128 uniform
template <int MODE>
__global__ void __launch_bounds__(256) k_smem(float *out, int iters)
{
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(sm));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned row = ((it + i * 3 + warp) & 15);
            if (MODE == 0) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (row * 32 + lane) * 4)); acc += v; }
            if (MODE == 1) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + (row * 32 + lane) * 8)); acc += v.x + v.y; }
            if (MODE == 2) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + (row * 32 + lane) * 16)); acc += v.x + v.w; }
            if (MODE == 3) asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(base + (row * 32 + lane) * 8 + (warp & 3) * 4096), "f"(acc), "f"((float)it));
            if (MODE == 4) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%1,%2};" :: "r"(base + (row * 32 + lane) * 16 + (warp & 3) * 8192), "f"(acc), "f"((float)it));
            if (MODE == 5) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + (row * 32) * 8)); acc += v.x + v.y; }
            if (MODE == 6) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + (row * 32) * 16)); acc += v.x + v.w; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static int g_sms = 148;
static double g_clk = 1.965e9;

template <typename F>
int run(const char *name, F launch, double ops_per_sm, const char *what, double bytes_per_op = 0)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double per_clk = ops_per_sm / (best * 1e-3 * g_clk);
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"%s_per_clk_per_sm\": %.3f, \"bytes_per_clk_per_sm\": %.1f}\n",
           name, best, what, per_clk, per_clk * bytes_per_op);
    fflush(stdout);
    return 0;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    g_clk = prop.clockRate * 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, g_sms, prop.clockRate);
    float *out; CK(cudaMalloc(&out, sizeof(float) * g_sms * 4 * 1024));
    const int sms = g_sms;

    // ---- TMEM: warp-instructions of 32 lanes x 32 columns x 4 B = 4 KB each
    {
        const int iters = 20000;
        for (int nw : {4, 8, 16}) {
            char nm[128];
            const double instr = (double)iters * nw;
            auto go = [&](const char *label, auto kern, double mult) {
                snprintf(nm, sizeof nm, "tmem %s, %d warps/SM", label, nw);
                run(nm, [=]() { kern<<<sms, nw * 32, 4096 * 4>>>(out, iters, 1); }, instr * mult, "warp_instr", 4096);
            };
            go("st.x32", k_tmem<0>, 1);
            go("ld.x32 (wait each)", k_tmem<1>, 1);
            go("st+wait+ld+wait round trip (instr pairs)", k_tmem<2>, 1);
            go("ld.x32 + 32 FFMA2 (ld instr)", k_tmem<3>, 1);
            go("32 FFMA2 alone (groups)", k_tmem<5>, 1);
            go("ld.x32 + 8 LDS.64 (ld instr)", k_tmem<4>, 1);
            go("8 LDS.64 alone (groups)", k_tmem<6>, 1);
        }
    }
    // ---- issue sharing, 16 warps/SM (2 CTAs x 256)
    {
        const int iters = 40000;
        const double groups = (double)iters * 16;   // per SM: one "group" = one trip of the 8-wide body per warp
        auto go = [&](const char *label, auto kern) {
            run(label, [=]() { kern<<<sms * 2, 256, 4096 * 4>>>(out, iters, 3); }, groups, "body");
        };
        go("issue: 8 FFMA2", k_issue<0>);
        go("issue: 8 FFMA2 + 8 LOP3", k_issue<1>);
        go("issue: 8 LOP3", k_issue<2>);
        go("issue: 8 FFMA2 + 8 FFMA", k_issue<3>);
        go("issue: 8 FFMA2 + 4 LDS.64", k_issue<4>);
        go("issue: 8 FFMA2 + 8 LOP3 + 4 LDS.64", k_issue<5>);
        go("issue: 16 FFMA + 8 LOP3", k_issue<6>);
    }

    // ---- issue sharing, second series: clocks per body of 8 (A) + 8 (B) per SMSP, at 4 / 2 / 1 warps per SMSP
    {
        const int iters = 20000;
        auto go = [&](const char *label, auto kern) {
            for (int wps : {4, 2, 1}) {
                char nm[160];
                snprintf(nm, sizeof nm, "issue2: %s, %d warps/SMSP", label, wps);
                // ops_per_sm = bodies per SMSP so that the printed value is bodies/clk/SMSP
                run(nm, [=]() { kern<<<sms, wps * 128, 4096 * 4>>>(out, iters, 3); }, (double)iters * wps, "body_smsp");
            }
        };
        go("8 FADD2", k_issue2<1, 0>);
        go("8 FMUL2", k_issue2<2, 0>);
        go("8 FFMA2", k_issue2<3, 0>);
        go("8 FADD (scalar)", k_issue2<0, 3>);
        go("8 FFMA (scalar)", k_issue2<0, 4>);
        go("8 LOP3", k_issue2<0, 1>);
        go("8 IADD", k_issue2<0, 2>);
        go("8 LDS.32", k_issue2<0, 5>);
        go("8 LDS.64", k_issue2<0, 6>);
        go("8 FADD2 + 8 LOP3", k_issue2<1, 1>);
        go("8 FADD2 + 8 IADD", k_issue2<1, 2>);
        go("8 FADD2 + 8 FADD", k_issue2<1, 3>);
        go("8 FADD2 + 8 FFMA", k_issue2<1, 4>);
        go("8 FADD2 + 8 LDS.32", k_issue2<1, 5>);
        go("8 FADD2 + 8 LDS.64", k_issue2<1, 6>);
        go("8 FADD2 + 8 PRMT", k_issue2<1, 7>);
        go("8 FADD2 + 8 FADD2", k_issue2<1, 8>);
        go("8 FFMA2 + 8 LOP3", k_issue2<3, 1>);
        go("8 FFMA2 + 8 FADD", k_issue2<3, 3>);
        go("8 FFMA2 + 8 LDS.64", k_issue2<3, 6>);
        go("8 FFMA2 + 8 FADD2", k_issue2<3, 8>);
        go("8 FADD + 8 LOP3 (scalar pair)", k_issue2<4, 1>);
        go("8 FFMA + 8 LOP3 (scalar pair)", k_issue2<5, 1>);
        go("8 FFMA + 8 FADD (scalar pair)", k_issue2<5, 3>);
        go("8 FFMA + 8 LDS.64", k_issue2<5, 6>);
    }
    // ---- one long straight-line body for every warp (16 warps/SM); FFMAs per clk per SM (warp instr)
    {
        auto go = [&](const char *label, auto kern, int body, int iters) {
            run(label, [=]() { kern<<<sms * 2, 256>>>(out, iters); }, (double)body * iters * 16, "warp_ffma");
        };
        go("long body 1200 FFMA (19 KB)", k_long<1200>, 1200, 512);
        go("long body 2400 FFMA (38 KB)", k_long<2400>, 2400, 256);
        go("long body 4800 FFMA (77 KB)", k_long<4800>, 4800, 128);
        go("long body 9600 FFMA (154 KB)", k_long<9600>, 9600, 64);
        go("long body 19200 FFMA (307 KB)", k_long<19200>, 19200, 32);
    }
    // ---- shared memory per width, 16 warps/SM
    {
        const int iters = 40000;
        const double instr = (double)iters * 8 * 16;
        CK(cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        auto go = [&](const char *label, auto kern, double bytes) {
            run(label, [=]() { kern<<<sms * 2, 256, 65536>>>(out, iters); }, instr, "warp_instr", bytes);
        };
        CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_smem<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_smem<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        go("smem LDS.32", k_smem<0>, 128);
        go("smem LDS.64", k_smem<1>, 256);
        go("smem LDS.128", k_smem<2>, 512);
        go("smem STS.64", k_smem<3>, 256);
        go("smem STS.128", k_smem<4>, 512);
        go("smem LDS.64 uniform", k_smem<5>, 256);
        go("smem LDS.128 uniform", k_smem<6>, 512);
    }
    return 0;
}
