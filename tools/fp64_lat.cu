// Microbenchmark: dependent-chain latency and multi-warp throughput of the FP64 pipe on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, long long *cyc, double a, double b, int iters) {
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) x = fma(x, y, b);                  // DFMA
            else if (OP == 1) x = x + y;                    // DADD
            else if (OP == 2) x = x * y;                    // DMUL
            else if (OP == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }  // MUFU.RCP64H
            else if (OP == 4) x = (double)(float)x;         // F2F pair
            else if (OP == 5) { unsigned long long q = __double_as_longlong(x); q = (q + 0x10000000ull) & 0xFFFFFFFFE0000000ull; x = __longlong_as_double(q); }
            else if (OP == 6) { float f = __double2float_rn(x); f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// throughput: ILP independent chains per thread
template <int ILP>
__global__ void tput(double *out, long long *cyc, double a, double b, int iters) {
    double x[ILP];
    for (int j = 0; j < ILP; ++j) x[j] = a + j + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], b, a);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const char *names[] = {"DFMA", "DADD", "DMUL", "RCP64H", "F2F f64->f32->f64", "int round (3 ops)", "F2F+FFMA+F2F"};
    const int iters = 2000;
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-22s dependent latency: %.2f cycles/op\n", names[OP], (double)h / (iters * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    for (int warps = 1; warps <= 16; warps *= 2) {
        tput<1><<<1, 32 * warps * 4>>>(out, cyc, 1.0, 0.9999999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double c1 = (double)h / (iters * 8.0);
        tput<4><<<1, 32 * warps * 4>>>(out, cyc, 1.0, 0.9999999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double c4 = (double)h / (iters * 8.0 * 4);
        printf("DFMA %2d warps/SMSP: ILP1 %.2f cyc per warp-DFMA-slot, ILP4 %.2f (cycles per DFMA per warp)\n", warps, c1, c4);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
