// Measures the integer-pipe issue peaks of the GPU this runs on: LOP3, POPC and the 14:1:1
// LOP3:POPC:IADD mix of the match-count kernel.  These are the roofline denominators for the INT-bound
// tile kernel (MEASURED_PEAKS.json only has HBM and bf16 peaks).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu && ./int_peak [seconds]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int CHAINS = 16;
constexpr int UNROLL = 8;

enum Mode { M_LOP3 = 0, M_POPC = 1, M_IADD3 = 2, M_MIX = 3 };

template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(uint32_t *out, uint32_t a, uint32_t b, int iters,
                                                   unsigned long long *cycles) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 2654435761u + c * 40503u + a;
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < (MODE == M_MIX ? 2 : UNROLL); ++u) {
            if (MODE == M_LOP3) {
#pragma unroll
                for (int c = 0; c < CHAINS; ++c)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b));
            } else if (MODE == M_POPC) {
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
            } else if (MODE == M_IADD3) {
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(a));
            } else {
                // 14 LOP3 (acc |= a ^ b style) then one POPC and one add, per chain, like the tile kernel
#pragma unroll
                for (int p = 0; p < 14; ++p)
#pragma unroll
                    for (int c = 0; c < CHAINS; ++c)
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0xF6;" : "+r"(x[c]) : "r"(a + p), "r"(b + c));
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    uint32_t pc;
                    asm volatile("popc.b32 %0, %1;" : "=r"(pc) : "r"(x[c]));
                    acc += pc;
                }
            }
        }
    }
    long long t1 = clock64();
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

struct Result {
    double ops_per_s, ops_per_clk_sm, sm_mhz, ms;
};

template <int MODE>
Result run(int sms, int ctas_per_sm, double seconds, uint32_t *out, unsigned long long *cyc) {
    const int grid = sms * ctas_per_sm, block = 256;
    const double ops_per_iter = (MODE == M_MIX ? 14.0 * 2 : 1.0 * UNROLL) * CHAINS;  // counted op: LOP3 in the mix
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 2000;
    float ms = 0;
    for (int attempt = 0; attempt < 6; ++attempt) {  // warm up and scale to the wanted duration
        cudaEventRecord(e0);
        peak_kernel<MODE><<<grid, block>>>(out, 0x12345u, 0x9e3779b9u, iters, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms > seconds * 1000 * 0.8) break;
        double scale = seconds * 1000 / (ms > 0.01 ? ms : 0.01);
        iters = (int)(iters * (scale > 50 ? 50 : scale));
    }
    std::vector<unsigned long long> h(grid);
    cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double cyc_mean = 0;
    for (auto c : h) cyc_mean += (double)c;
    cyc_mean /= grid;
    double total = ops_per_iter * (double)iters * block * grid;
    Result r;
    r.ms = ms;
    r.ops_per_s = total / (ms * 1e-3);
    // every CTA is resident for the whole kernel: per-SM rate = ops of its CTAs / cycles it ran
    r.ops_per_clk_sm = ops_per_iter * (double)iters * block * ctas_per_sm / cyc_mean;
    r.sm_mhz = cyc_mean / (ms * 1e3);
    return r;
}

int main(int argc, char **argv) {
    double seconds = argc > 1 ? atof(argv[1]) : 1.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int cps = argc > 2 ? atoi(argv[2]) : 4;  // CTAs of 256 threads per SM (4 => 32 warps per SM)
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, (size_t)sms * cps * 256 * 4);
    cudaMalloc(&cyc, (size_t)sms * cps * 8);
    Result l = run<M_LOP3>(sms, cps, seconds, out, cyc);
    Result p = run<M_POPC>(sms, cps, seconds, out, cyc);
    Result m = run<M_MIX>(sms, cps, seconds, out, cyc);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"seconds_each\": %.2f, \"warps_per_sm\": %d,\n", prop.name, sms, seconds, cps * 8);
    printf(" \"lop3_per_s\": %.4e, \"lop3_per_clk_sm\": %.2f, \"lop3_sm_mhz\": %.0f,\n", l.ops_per_s, l.ops_per_clk_sm, l.sm_mhz);
    printf(" \"popc_per_s\": %.4e, \"popc_per_clk_sm\": %.2f, \"popc_sm_mhz\": %.0f,\n", p.ops_per_s, p.ops_per_clk_sm, p.sm_mhz);
    printf(" \"mix_lop3_per_s\": %.4e, \"mix_lop3_per_clk_sm\": %.2f, \"mix_sm_mhz\": %.0f,\n", m.ops_per_s, m.ops_per_clk_sm, m.sm_mhz);
    printf(" \"how\": \"16 independent chains per thread, 32 warps per SM, all SMs; volatile inline PTX; LOP3 of the 14:1:1 LOP3:POPC:IADD mix counted for mix_*\"}\n");
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
