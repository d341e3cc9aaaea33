// microbench.cu -- roofline denominators measured on the device the sweeps run on.
// Not on the product path: bench.py calls these to obtain the FP64 FMA peak (absent from
// MEASURED_PEAKS.json) and the HBM bandwidth of the sweeps' own access pattern.
#include <cuda_runtime.h>

#include <vector>

#include "../../include/ocs_b200.h"

namespace {

// 8 independent dependent-FMA chains per thread: enough ILP to cover the DFMA latency
// even at one warp per scheduler.
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// The same chains with the multiplier read from the constant bank (DFMA R, R, c[..], R): one register operand fewer
// per instruction, in case register-bank conflicts of the three-register form cost issue slots.  clk[block] receives the
// SM cycles the block ran (every block of the grid is resident at once, so this is the kernel's length in SM cycles).
__constant__ double c_mul[1];
template <int CONSTOP>
__global__ void __launch_bounds__(256) dfma_clock_kernel(double *out, long long *clk, int iters, double a, double b)
{
    const long long t0 = clock64();
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (CONSTOP) {
                x0 = fma(x0, c_mul[0], b); x1 = fma(x1, c_mul[0], b); x2 = fma(x2, c_mul[0], b); x3 = fma(x3, c_mul[0], b);
                x4 = fma(x4, c_mul[0], b); x5 = fma(x5, c_mul[0], b); x6 = fma(x6, c_mul[0], b); x7 = fma(x7, c_mul[0], b);
            } else {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
                x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
            }
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

__global__ void __launch_bounds__(256) hbm_write_kernel(double2 *dst, size_t n2, double v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) dst[k] = make_double2(v, v);
}
__global__ void __launch_bounds__(256) hbm_read_kernel(const double2 *__restrict__ src, size_t n2, double *out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n2; k += stride) {
        const double2 x = src[k];
        s += x.x + x.y;
    }
    if (s == 123.456) out[0] = s;   // keep the loads alive
}

}  // namespace

extern "C" {

int ocs_b200_measure_fp64_peak(int device, double *tflops)
{
    if (!tflops) return OCS_B200_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return OCS_B200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OCS_B200_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return OCS_B200_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return OCS_B200_ERR_CUDA;
    *tflops = best;
    return 0;
}

// FP64 FMA peak with the SM clock the device actually ran at while it was measured: longer runs than
// ocs_b200_measure_fp64_peak (~40 ms each, so that a power- or thermally-limited clock shows), both operand forms,
// and sm_mhz = SM cycles of the kernel (clock64) / its CUDA-event duration.
int ocs_b200_measure_fp64_peak_ex(int device, double *tflops, double *sm_mhz)
{
    if (!tflops) return OCS_B200_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return OCS_B200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OCS_B200_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 16384;
    double *out = nullptr;
    long long *clk = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return OCS_B200_ERR_NOMEM;
    if (cudaMalloc(&clk, sizeof(long long) * blocks) != cudaSuccess) {
        cudaFree(out);
        return OCS_B200_ERR_NOMEM;
    }
    const double mul = 1.0000001;
    cudaMemcpyToSymbol(c_mul, &mul, sizeof mul);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0, best_mhz = 0;
    std::vector<long long> hclk(blocks);
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (rep & 1) dfma_clock_kernel<1><<<blocks, threads>>>(out, clk, iters, mul, 1e-9);
        else dfma_clock_kernel<0><<<blocks, threads>>>(out, clk, iters, mul, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(hclk.data(), clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        long long cyc = 0;
        for (long long c : hclk) cyc = c > cyc ? c : cyc;
        const double flop = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 1 && tf > best) {
            best = tf;
            best_mhz = (double)cyc / (ms * 1e3);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    cudaFree(clk);
    if (cudaGetLastError() != cudaSuccess) return OCS_B200_ERR_CUDA;
    *tflops = best;
    if (sm_mhz) *sm_mhz = best_mhz;
    return 0;
}

int ocs_b200_measure_hbm(int device, size_t bytes, double *write_gbs, double *read_gbs)
{
    if (cudaSetDevice(device) != cudaSuccess) return OCS_B200_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OCS_B200_ERR_CUDA;
    const size_t n2 = bytes / sizeof(double2);
    double2 *buf = nullptr;
    double *out = nullptr;
    if (cudaMalloc(&buf, n2 * sizeof(double2)) != cudaSuccess) return OCS_B200_ERR_NOMEM;
    cudaMalloc(&out, sizeof(double));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double bw = 0, br = 0;
    for (int rep = 0; rep < 5; ++rep) {
        float ms = 0;
        cudaEventRecord(e0);
        hbm_write_kernel<<<blocks, threads>>>(buf, n2, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double g = n2 * sizeof(double2) / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > bw) bw = g;
        cudaEventRecord(e0);
        hbm_read_kernel<<<blocks, threads>>>(buf, n2, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        g = n2 * sizeof(double2) / (ms * 1e-3) / 1e9;
        if (rep > 0 && g > br) br = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return OCS_B200_ERR_CUDA;
    if (write_gbs) *write_gbs = bw;
    if (read_gbs) *read_gbs = br;
    return 0;
}

}  // extern "C"
