// fp64_uniform_operand.cu -- does it pay to feed block-uniform operands (grid-point tables, step sizes) of the sweeps
// through the constant bank / uniform registers instead of shared memory?  A DFMA reading three vector registers takes
// its scheduler's FP64 pipe for 3 cycles, one with a constant/uniform operand for 2 (fp64_latency_map.cu).  The loop
// below has the shape of the test problem's forward RK4 stage: per trajectory t = m - y; f = fma(y, t, -u);
// y = fma(h, f, y) with (u, h) block-uniform per step, four trajectories per thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/bin/fp64_uniform_operand profiles/tools/fp64_uniform_operand.cu
#include <cstdio>
#include <cuda_runtime.h>

#define STEPS 2048
__constant__ double c_rec[2 * STEPS];

template <int SRC>   // 0: shared memory tile, 1: constant bank with a loop-uniform index, 2: REDUX of the shared value
__global__ void __launch_bounds__(128) stage_loop(double *out, long long *clk, const double *rec, int reps, double m)
{
    extern __shared__ double sm[];
    for (int k = threadIdx.x; k < 2 * STEPS; k += blockDim.x) sm[k] = rec[k];
    __syncthreads();
    double y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) y[q] = 0.5 + 1e-3 * threadIdx.x + 0.1 * q;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
        for (int s = 0; s < STEPS; ++s) {
            double u, h;
            if (SRC == 0) {
                u = sm[2 * s];
                h = sm[2 * s + 1];
            } else if (SRC == 1) {
                u = c_rec[2 * s];
                h = c_rec[2 * s + 1];
            } else {
                const double us = sm[2 * s], hs = sm[2 * s + 1];
                const unsigned ul = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(us));
                const unsigned uh = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(us));
                const unsigned hl = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(hs));
                const unsigned hh = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(hs));
                u = __hiloint2double((int)uh, (int)ul);
                h = __hiloint2double((int)hh, (int)hl);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double t = m - y[q];
                const double f = fma(y[q], t, -u);
                y[q] = fma(h, f, y[q]);
            }
        }
    }
    const long long t1 = clock64();
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (y[0] + y[1]) + (y[2] + y[3]);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, reps = 8;
    static double h[2 * STEPS];
    for (int s = 0; s < STEPS; ++s) {
        h[2 * s] = 0.3 + 1e-4 * (s % 7);
        h[2 * s + 1] = 1e-3;
    }
    double *rec, *out;
    long long *clk;
    cudaMalloc(&rec, sizeof h);
    cudaMemcpy(rec, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_rec, h, sizeof h);
    cudaMalloc(&out, sizeof(double) * sms * 8 * 128);
    cudaMalloc(&clk, sizeof(long long) * sms * 8);
    const char *names[] = {"shared-memory operands", "constant-bank operands", "REDUX to uniform registers"};
    for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
        for (int src = 0; src < 3; ++src) {
            for (int rep = 0; rep < 2; ++rep) {
                if (src == 0) stage_loop<0><<<sms * blocks_per_sm, 128, sizeof h>>>(out, clk, rec, reps, 3.0);
                else if (src == 1) stage_loop<1><<<sms * blocks_per_sm, 128, sizeof h>>>(out, clk, rec, reps, 3.0);
                else stage_loop<2><<<sms * blocks_per_sm, 128, sizeof h>>>(out, clk, rec, reps, 3.0);
                cudaDeviceSynchronize();
            }
            long long hc[2048], mx = 0;
            cudaMemcpy(hc, clk, sizeof(long long) * sms * blocks_per_sm, cudaMemcpyDeviceToHost);
            for (int k = 0; k < sms * blocks_per_sm; ++k) mx = hc[k] > mx ? hc[k] : mx;
            // per scheduler: blocks_per_sm warps (128-thread blocks = one warp per scheduler each), 12 FP64 instructions per step
            const double cyc_per_step = (double)mx / ((double)reps * STEPS);
            printf("%-28s warps/sched %d: %.2f cycles per step per scheduler, %.2f per FP64 warp instruction (%s)\n", names[src],
                   blocks_per_sm, cyc_per_step, cyc_per_step / (12.0 * blocks_per_sm), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
