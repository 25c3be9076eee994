// Issue-rate microbenchmark for the edge-solve inner loop on sm_100a: scalar vs packed FP32, FMNMX, MUFU.RCP and
// mixes of them.  Every test runs ILP independent chains per thread, 1024 threads x 2 CTAs per SM x all SMs, and
// reports warp-instructions per clock per SM sub-partition (SMSP).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

#define PACK2(op, d, a, b) asm volatile("{\n\t.reg .b64 x,y,z;\n\tmov.b64 x,{%2,%3};\n\tmov.b64 y,{%4,%5};\n\t" op " z,x,y;\n\tmov.b64 {%0,%1},z;\n\t}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y))
#define FMA2(d, a, b, c) asm volatile("{\n\t.reg .b64 x,y,w,z;\n\tmov.b64 x,{%2,%3};\n\tmov.b64 y,{%4,%5};\n\tmov.b64 w,{%6,%7};\n\tfma.rn.f32x2 z,x,y,w;\n\tmov.b64 {%0,%1},z;\n\t}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y))

constexpr int ILP = 8, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024) bench(float* out, float seed, long long* clk) {
    float2 a[ILP];
    float2 b = make_float2(seed, seed * 0.5f), c = make_float2(0.25f, 0.75f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = make_float2(seed + i, seed - i);
    __shared__ float sm[2048];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 1024] = seed;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(b.x), "f"(c.x)); }           // FFMA
            if (MODE == 1) { FMA2(a[i], a[i], b, c); }                                                                    // FFMA2
            if (MODE == 2) { PACK2("add.rn.f32x2", a[i], a[i], b); }                                                      // FADD2
            if (MODE == 3) { asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(b.x)); }                              // FMNMX
            if (MODE == 4) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i].x)); }                                  // MUFU.RCP
            if (MODE == 5) { FMA2(a[i], a[i], b, c); asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(b.x)); }      // FFMA2 + FMNMX
            if (MODE == 6) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(b.x), "f"(c.x));
                             asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].y) : "f"(b.y)); }                              // FFMA + FMNMX
            if (MODE == 7) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(b.x)); }                           // FADD
            if (MODE == 8) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + i + it) & 2047))));
                             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(b.x), "f"(v)); }               // LDS + FFMA
            if (MODE == 9) { FMA2(a[i], a[i], b, c); FMA2(a[i], a[i], b, c);
                             asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(b.x));
                             asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i].y)); }                                  // 2 FFMA2 + FMNMX + MUFU
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, float* out, long long* clk, int sms) {
    const int grid = sms * 2;
    bench<MODE><<<grid, 1024>>>(out, 1.0001f, clk);
    cudaDeviceSynchronize();
    bench<MODE><<<grid, 1024>>>(out, 1.0001f, clk);
    cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += (double)h[i];
    avg /= grid;
    // warp-instructions per SM: 2 CTAs x 32 warps x ITERS x ILP x per_iter, spread over 4 SMSPs
    const double winstr = 2.0 * 32 * ITERS * ILP * per_iter / 4.0;
    printf("%-28s %8.3f warp-instr/clk/SMSP  (%.0f clk)\n", name, winstr / avg, avg);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* clk;
    cudaMalloc(&out, sizeof(float) * sms * 2 * 1024);
    cudaMalloc(&clk, sizeof(long long) * 1024);
    printf("SMs %d\n", sms);
    run<0>("FFMA", 1, out, clk, sms);
    run<7>("FADD", 1, out, clk, sms);
    run<1>("FFMA2 (packed)", 1, out, clk, sms);
    run<2>("FADD2 (packed)", 1, out, clk, sms);
    run<3>("FMNMX", 1, out, clk, sms);
    run<4>("MUFU.RCP", 1, out, clk, sms);
    run<5>("FFMA2 + FMNMX", 2, out, clk, sms);
    run<6>("FFMA + FMNMX", 2, out, clk, sms);
    run<8>("LDS.32 + FFMA", 2, out, clk, sms);
    run<9>("2 FFMA2 + FMNMX + MUFU.RCP", 4, out, clk, sms);
    return 0;
}
