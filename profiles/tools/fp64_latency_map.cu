// fp64_latency_map.cu -- FP64 pipe utilisation of sm_100a as a function of resident warps per scheduler and independent
// dependency chains per thread (profiles/r2_notes.md).  The generated-problem sweeps run at 2-3 warps per scheduler with
// a handful of independent rows per thread; this map says what the pipe can deliver there, and the dependent-issue
// latency of DFMA / DADD / DMUL.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/bin/fp64_latency_map profiles/tools/fp64_latency_map.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C, int MIX>
__global__ void chains(double *out, long long *clk, int iters, double a, double b)
{
    double x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = threadIdx.x * 1e-3 + c;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (MIX == 0) x[c] = fma(x[c], a, b);
                else if (MIX == 1) x[c] = x[c] + b;
                else if (MIX == 2) x[c] = x[c] * a;
                else x[c] = (k % 3 == 0) ? fma(x[c], a, b) : ((k % 3 == 1) ? x[c] + b : x[c] * a);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// Three distinct, changing register operands per instruction (what real right-hand sides look like: no operand of the
// previous instruction is reused, nothing sits in a constant bank): every chain rotates through four registers,
// w = fma(x, y, z); (x, y, z) <- (y, z, w).
template <int C>
__global__ void chains3(double *out, long long *clk, int iters)
{
    double x[C], y[C], z[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        x[c] = 1.0 + threadIdx.x * 1e-9 + c * 1e-7;
        y[c] = 1.0 - threadIdx.x * 1e-9;
        z[c] = 1e-9 * c;
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double w = fma(x[c], y[c], z[c]);
                x[c] = y[c];
                y[c] = z[c];
                z[c] = w;
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c] + y[c] + z[c];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// Fixed operand registers per chain (the allocator is free to place them in different banks): FORM 0: x = fma(x, a_c, b_c);
// 1: x = fma(a_c, b_c, x); 2: x = x + a_c; 3: x = x * a_c; 4: x = fma(x, const-bank, b_c); 5: x = fma(x, a_c, const-bank)
__constant__ double c_k[2];
template <int C, int FORM>
__global__ void chainsf(double *out, long long *clk, int iters, const double *in)
{
    double x[C], a[C], b[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        x[c] = 1.0 + threadIdx.x * 1e-9 + c * 1e-7;
        a[c] = in[threadIdx.x * 32 + c];            // per-thread values: vector registers, not uniform ones
        b[c] = in[threadIdx.x * 32 + 16 + c];
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                if (FORM == 0) x[c] = fma(x[c], a[c], b[c]);
                else if (FORM == 1) x[c] = fma(a[c], b[c], x[c]);
                else if (FORM == 2) x[c] = x[c] + a[c];
                else if (FORM == 3) x[c] = x[c] * a[c];
                else if (FORM == 4) x[c] = fma(x[c], c_k[0], b[c]);
                else x[c] = fma(x[c], a[c], c_k[1]);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += x[c];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int C, int FORM>
void runf(int sms, int warps_per_sched, double *out, long long *clk, const double *in)
{
    const int threads = 32 * 4 * warps_per_sched, iters = 2048;
    chainsf<C, FORM><<<sms, threads>>>(out, clk, iters, in);
    cudaDeviceSynchronize();
    chainsf<C, FORM><<<sms, threads>>>(out, clk, iters, in);
    cudaDeviceSynchronize();
    long long h[1024], mx = 0;
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    for (int k = 0; k < sms; ++k) mx = h[k] > mx ? h[k] : mx;
    const double util = (double)iters * 12 * C * warps_per_sched * 2.0 / (double)mx;
    static const char *names[] = {"x=fma(x,a_c,b_c)", "x=fma(a_c,b_c,x)", "x=x+a_c", "x=x*a_c", "x=fma(x,const,b_c)", "x=fma(x,a_c,const)"};
    printf("FORM %s warps/sched %d chains %d: pipe %.3f (cycles per warp instruction %.2f)\n", names[FORM], warps_per_sched, C, util,
           2.0 / util);
}

template <int C>
void run3(int sms, int warps_per_sched, double *out, long long *clk)
{
    const int threads = 32 * 4 * warps_per_sched, iters = 2048;
    chains3<C><<<sms, threads>>>(out, clk, iters);
    cudaDeviceSynchronize();
    chains3<C><<<sms, threads>>>(out, clk, iters);
    cudaDeviceSynchronize();
    long long h[1024], mx = 0;
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    for (int k = 0; k < sms; ++k) mx = h[k] > mx ? h[k] : mx;
    const double util = (double)iters * 12 * C * warps_per_sched * 2.0 / (double)mx;
    printf("DFMA3 (3 distinct register operands) warps/sched %d chains %d: pipe %.3f  cycles per chain step %.2f\n",
           warps_per_sched, C, util, (double)mx / ((double)iters * 12));
}

template <int C, int MIX>
void run(int sms, int warps_per_sched, double *out, long long *clk)
{
    const int threads = 32 * 4 * warps_per_sched, iters = 2048;   // one block per SM, warps spread over the 4 schedulers
    chains<C, MIX><<<sms, threads>>>(out, clk, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    chains<C, MIX><<<sms, threads>>>(out, clk, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h[1024], mx = 0;
    cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    for (int k = 0; k < sms; ++k) mx = h[k] > mx ? h[k] : mx;
    const double inst_per_sched = (double)iters * 12 * C * warps_per_sched;     // warp instructions per scheduler
    const double util = inst_per_sched * 2.0 / (double)mx;                      // 16 FP64 lanes per scheduler: 2 cycles per warp instruction
    const double lat = (double)mx / ((double)iters * 12);                       // cycles per dependent step of one chain set
    printf("%s warps/sched %d chains %d: pipe %.3f  cycles per chain step %.2f\n",
           MIX == 0 ? "DFMA" : MIX == 1 ? "DADD" : MIX == 2 ? "DMUL" : "MIX ", warps_per_sched, C, util, lat);
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double *out;
    long long *clk;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&clk, sizeof(long long) * sms);
    for (int w = 1; w <= 4; ++w) {
        run<1, 0>(sms, w, out, clk); run<2, 0>(sms, w, out, clk); run<3, 0>(sms, w, out, clk); run<4, 0>(sms, w, out, clk);
        run<6, 0>(sms, w, out, clk); run<8, 0>(sms, w, out, clk);
        run<1, 3>(sms, w, out, clk); run<2, 3>(sms, w, out, clk); run<4, 3>(sms, w, out, clk); run<6, 3>(sms, w, out, clk);
    }
    run<1, 1>(sms, 1, out, clk); run<1, 2>(sms, 1, out, clk); run<8, 1>(sms, 4, out, clk); run<8, 2>(sms, 4, out, clk);
    run<8, 3>(sms, 8, out, clk); run<8, 0>(sms, 8, out, clk);
    {
        static double hin[1024 * 32];
        double kk[2] = {1.0000001, 1e-9};
        for (int k = 0; k < 1024 * 32; ++k) hin[k] = (k % 32) < 16 ? 1.0 + 1e-8 * (k % 32) : 1e-9 * (k % 32);
        double *in;
        cudaMalloc(&in, sizeof hin);
        cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
        cudaMemcpyToSymbol(c_k, kk, sizeof kk);
        for (int w = 2; w <= 4; w += 2) {
            runf<8, 0>(sms, w, out, clk, in); runf<8, 1>(sms, w, out, clk, in); runf<8, 2>(sms, w, out, clk, in);
            runf<8, 3>(sms, w, out, clk, in); runf<8, 4>(sms, w, out, clk, in); runf<8, 5>(sms, w, out, clk, in);
            runf<6, 0>(sms, w, out, clk, in); runf<6, 1>(sms, w, out, clk, in);
        }
    }
    for (int w = 1; w <= 4; ++w) {
        run3<1>(sms, w, out, clk); run3<2>(sms, w, out, clk); run3<4>(sms, w, out, clk); run3<6>(sms, w, out, clk); run3<8>(sms, w, out, clk);
    }
    return 0;
}
