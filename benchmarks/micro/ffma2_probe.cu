// ffma2_probe.cu -- micro-benchmark (B200): issue/pipe cost of packed FFMA2 against scalar FFMA, and candidate inner loops
// of the tile-accumulate kernel (cycles per evaluated pixel per warp).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fsqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- raw pipes ----
template <int MODE>
__global__ void __launch_bounds__(256) raw_kernel(float *out, int iters, float seed)
{
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
    const float m = 0.999f + seed * 1e-6f, c = 1e-3f * seed;
    if (MODE == 0) {                       // 16 scalar FFMA per iteration
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], m, c);
    } else if (MODE == 1) {                // 8 FFMA2 per iteration = 16 lane-fma
        u64 p[8];
        const u64 M = pk(m, m), C = pk(c, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pk(v[2 * i], v[2 * i + 1]);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], M, C);
#pragma unroll
        for (int i = 0; i < 8; ++i) upk(p[i], v[2 * i], v[2 * i + 1]);
    } else if (MODE == 2) {                // 8 FFMA2 + 8 alu-pipe FMNMX
        u64 p[4];
        const u64 M = pk(m, m), C = pk(c, c);
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = pk(v[2 * i], v[2 * i + 1]);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], M, C); p[i] = fma2(p[i], M, C); }
#pragma unroll
            for (int i = 8; i < 16; ++i) v[i] = fminf(v[i], c + (float)it);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) upk(p[i], v[2 * i], v[2 * i + 1]);
    } else if (MODE == 3) {                // 4 MUFU.SQRT + 12 FFMA
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = fsqrt(v[i]) + 1.0f;
#pragma unroll
            for (int i = 4; i < 16; ++i) v[i] = fmaf(v[i], m, c);
        }
    } else if (MODE == 4) {                // 16 MUFU.SQRT
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fsqrt(v[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- candidate inner loops: E staged entries, 2x2 pixels per thread, 2 weight fields ----
struct Ent { float4 P; float2 C; };
template <int V>
__global__ void __launch_bounds__(256) loop_kernel(float *out, int E, int reps, float seed)
{
    __shared__ float4 sP[8][32];
    __shared__ float2 sC[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sP[warp][lane] = make_float4(0.3f + 0.01f * lane + seed, 0.2f + 0.02f * lane, 0.05f, 0.05f);
    sC[warp][lane] = make_float2(1.0f + lane, 2.0f);
    __syncwarp();
    const float xf[2] = {(float)(2 * (lane >> 3)), (float)(2 * (lane >> 3) + 1)};
    const float yf[2] = {(float)(2 * (lane & 7)), (float)(2 * (lane & 7) + 1)};
    float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    u64 acc2[2][2] = {{0, 0}, {0, 0}};
    const u64 yf2 = pk(yf[0], yf[1]);
    for (int r = 0; r < reps; ++r) {
        for (int e = 0; e < E; ++e) {
            const float4 q = sP[warp][e & 31];
            const float2 c = sC[warp][e & 31];
            if (V == 0) {                  // current full path
                float ax2[2], by2[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) { float t = fmaf(-xf[i], q.z, q.x); ax2[i] = t * t; }
#pragma unroll
                for (int i = 0; i < 2; ++i) { float t = fmaf(-yf[i], q.w, q.y); by2[i] = t * t; }
#pragma unroll
                for (int ix = 0; ix < 2; ++ix)
#pragma unroll
                    for (int iy = 0; iy < 2; ++iy) {
                        const float qq = fsqrt(ax2[ix] + by2[iy]);
                        const float a = __saturatef(fmaf(qq, -0.5f, 1.0f)), b = __saturatef(1.0f - qq);
                        const float f = fmaf(2.0f, a * a * a, -(b * b * b));
                        acc[0][ix * 2 + iy] = fmaf(c.x, f, acc[0][ix * 2 + iy]);
                        acc[1][ix * 2 + iy] = fmaf(c.y, f, acc[1][ix * 2 + iy]);
                    }
            } else if (V == 1) {           // scalar min trick: f/2 = min(0.5 + s(0.375 q - 0.75), sat(1-q/2)^3)
                float ax2[2], by2[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) { float t = fmaf(-xf[i], q.z, q.x); ax2[i] = t * t; }
#pragma unroll
                for (int i = 0; i < 2; ++i) { float t = fmaf(-yf[i], q.w, q.y); by2[i] = t * t; }
#pragma unroll
                for (int ix = 0; ix < 2; ++ix)
#pragma unroll
                    for (int iy = 0; iy < 2; ++iy) {
                        const float s = ax2[ix] + by2[iy];
                        const float qq = fsqrt(s);
                        const float p = fmaf(s, fmaf(0.375f, qq, -0.75f), 0.5f);
                        const float a = __saturatef(fmaf(qq, -0.5f, 1.0f));
                        const float f = fminf(p, a * a * a);
                        acc[0][ix * 2 + iy] = fmaf(c.x, f, acc[0][ix * 2 + iy]);
                        acc[1][ix * 2 + iy] = fmaf(c.y, f, acc[1][ix * 2 + iy]);
                    }
            } else if (V == 2) {           // packed pairs along y, min trick, clamp s at 4
                const u64 qy2 = pk(q.y, q.y), qw2 = pk(-q.w, -q.w);
                u64 ty = fma2(yf2, qw2, qy2);
                const u64 by2 = mul2(ty, ty);
                const u64 cx2 = pk(c.x, c.x), cy2 = pk(c.y, c.y);
#pragma unroll
                for (int ix = 0; ix < 2; ++ix) {
                    float t = fmaf(-xf[ix], q.z, q.x);
                    t = t * t;
                    u64 s2 = add2(pk(t, t), by2);
                    float s0, s1;
                    upk(s2, s0, s1);
                    s0 = fminf(s0, 4.0f); s1 = fminf(s1, 4.0f);
                    const float q0 = fsqrt(s0), q1 = fsqrt(s1);
                    const u64 qq = pk(q0, q1);
                    s2 = pk(s0, s1);
                    const u64 tt = fma2(qq, pk(0.375f, 0.375f), pk(-0.75f, -0.75f));
                    const u64 p = fma2(s2, tt, pk(0.5f, 0.5f));
                    const u64 a = fma2(qq, pk(-0.5f, -0.5f), pk(1.0f, 1.0f));
                    const u64 a3 = mul2(mul2(a, a), a);
                    float p0, p1, a0, a1;
                    upk(p, p0, p1); upk(a3, a0, a1);
                    const u64 f = pk(fminf(p0, a0), fminf(p1, a1));
                    acc2[0][ix] = fma2(cx2, f, acc2[0][ix]);
                    acc2[1][ix] = fma2(cy2, f, acc2[1][ix]);
                }
            } else if (V == 3) {           // current outer-annulus path (scalar)
                float ax2[2], by2[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) { float t = fmaf(-xf[i], q.z, q.x); ax2[i] = t * t; }
#pragma unroll
                for (int i = 0; i < 2; ++i) { float t = fmaf(-yf[i], q.w, q.y); by2[i] = t * t; }
#pragma unroll
                for (int ix = 0; ix < 2; ++ix)
#pragma unroll
                    for (int iy = 0; iy < 2; ++iy) {
                        const float a1 = __saturatef(fmaf(fsqrt(ax2[ix] + by2[iy]), -0.5f, 1.0f));
                        const float f = a1 * a1 * a1;
                        acc[0][ix * 2 + iy] = fmaf(c.x, f, acc[0][ix * 2 + iy]);
                        acc[1][ix * 2 + iy] = fmaf(c.y, f, acc[1][ix * 2 + iy]);
                    }
            } else if (V == 4) {           // packed outer-annulus path
                const u64 qy2 = pk(q.y, q.y), qw2 = pk(-q.w, -q.w);
                u64 ty = fma2(yf2, qw2, qy2);
                const u64 by2 = mul2(ty, ty);
                const u64 cx2 = pk(c.x, c.x), cy2 = pk(c.y, c.y);
#pragma unroll
                for (int ix = 0; ix < 2; ++ix) {
                    float t = fmaf(-xf[ix], q.z, q.x);
                    t = t * t;
                    const u64 s2 = add2(pk(t, t), by2);
                    float s0, s1;
                    upk(s2, s0, s1);
                    const float q0 = fsqrt(fminf(s0, 4.0f)), q1 = fsqrt(fminf(s1, 4.0f));
                    const u64 a = fma2(pk(q0, q1), pk(-0.5f, -0.5f), pk(1.0f, 1.0f));
                    const u64 f = mul2(mul2(a, a), a);
                    acc2[0][ix] = fma2(cx2, f, acc2[0][ix]);
                    acc2[1][ix] = fma2(cy2, f, acc2[1][ix]);
                }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[k][j];
#pragma unroll
        for (int j = 0; j < 2; ++j) { float lo, hi; upk(acc2[k][j], lo, hi); s += lo + hi; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount, grid = sms * 8;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out;
    cudaMalloc(&out, sizeof(float) * grid * 256);
    printf("device %s, %d SMs, %d kHz\n", pr.name, sms, khz);
    const int iters = 20000;
    const char *names[] = {"16 FFMA", "8 FFMA2", "8 FFMA2 + 8 FMNMX", "4 MUFU.SQRT + 4 FADD + 12 FFMA", "16 MUFU.SQRT"};
    for (int m = 0; m < 5; ++m) {
        float ms = 0;
        switch (m) {
        case 0: ms = time_ms([&] { raw_kernel<0><<<grid, 256>>>(out, iters, 1.f); }); break;
        case 1: ms = time_ms([&] { raw_kernel<1><<<grid, 256>>>(out, iters, 1.f); }); break;
        case 2: ms = time_ms([&] { raw_kernel<2><<<grid, 256>>>(out, iters, 1.f); }); break;
        case 3: ms = time_ms([&] { raw_kernel<3><<<grid, 256>>>(out, iters, 1.f); }); break;
        case 4: ms = time_ms([&] { raw_kernel<4><<<grid, 256>>>(out, iters, 1.f); }); break;
        }
        // cycles per iteration per SMSP-warp: 8 CTAs x 8 warps per SM = 16 warps per SMSP
        const double cyc = ms * 1e-3 * khz * 1e3 / iters / 16.0;
        printf("raw %-34s %8.3f ms  %6.2f SMSP-cycles per warp-iteration\n", names[m], ms, cyc);
    }
    const int E = 4096, reps = 8;
    const char *ln[] = {"full (current)", "full min-trick scalar", "full min-trick packed", "outer (current)", "outer packed"};
    for (int v = 0; v < 5; ++v) {
        float ms = 0;
        switch (v) {
        case 0: ms = time_ms([&] { loop_kernel<0><<<grid, 256>>>(out, E, reps, 0.f); }); break;
        case 1: ms = time_ms([&] { loop_kernel<1><<<grid, 256>>>(out, E, reps, 0.f); }); break;
        case 2: ms = time_ms([&] { loop_kernel<2><<<grid, 256>>>(out, E, reps, 0.f); }); break;
        case 3: ms = time_ms([&] { loop_kernel<3><<<grid, 256>>>(out, E, reps, 0.f); }); break;
        case 4: ms = time_ms([&] { loop_kernel<4><<<grid, 256>>>(out, E, reps, 0.f); }); break;
        }
        const double cyc = ms * 1e-3 * khz * 1e3 / ((double)E * reps) / 16.0 / 4.0;
        printf("loop %-26s %8.3f ms  %6.2f SMSP-cycles per evaluated pixel (warp)\n", ln[v], ms, cyc);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
