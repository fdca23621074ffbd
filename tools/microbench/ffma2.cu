// Microbenchmark: scalar FFMA vs packed FFMA2 / FADD2 issue and pipe throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(512) k(float2* o, float2 a, float2 b) {
    float2 x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f);
    u64 A = *reinterpret_cast<u64*>(&a), B = *reinterpret_cast<u64*>(&b);
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) {  // 2 scalar FFMA
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j].x) : "f"(a.x), "f"(b.x));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j].y) : "f"(a.y), "f"(b.y));
            } else if (MODE == 1) {  // 1 FFMA2
                u64& X = *reinterpret_cast<u64*>(&x[j]);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X) : "l"(A), "l"(B));
            } else if (MODE == 2) {  // 1 FADD2
                u64& X = *reinterpret_cast<u64*>(&x[j]);
                asm volatile("add.f32x2 %0, %0, %1;" : "+l"(X) : "l"(B));
            } else if (MODE == 3) {  // 2 scalar FADD
                asm volatile("add.f32 %0, %0, %1;" : "+f"(x[j].x) : "f"(b.x));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(x[j].y) : "f"(b.y));
            } else if (MODE == 4) {  // FFMA2 + independent integer op (issue sharing)
                u64& X = *reinterpret_cast<u64*>(&x[j]);
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(X) : "l"(A), "l"(B));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*reinterpret_cast<unsigned*>(&a.x)) : "r"(i), "r"(j));
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s.x += x[j].x; s.y += x[j].y; }
    s.x += a.x;
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float2* o) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 4;
    k<MODE><<<grid, 512>>>(o, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f));
    cudaEventRecord(e0);
    k<MODE><<<grid, 512>>>(o, make_float2(1.0001f, 0.9999f), make_float2(1e-3f, -1e-3f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = double(grid) * 512 * ITERS * 8 * 2;  // scalar-equivalent FP ops
    printf("%-28s %8.3f ms  %8.2f T scalar-op/s  (%.1f ops/clk/SM @1.965GHz)\n", name, ms, lane_ops / ms / 1e9,
           lane_ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float2* o; cudaMalloc(&o, 148 * 4 * 512 * sizeof(float2));
    run<0>("2x FFMA scalar", o); run<1>("FFMA2", o); run<2>("FADD2", o); run<3>("2x FADD scalar", o); run<4>("FFMA2 + LOP3", o);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
