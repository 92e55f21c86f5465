// Measures the integer-pipe issue peaks of the GPU this runs on: LOP3, POPC and the 14:1:1
// LOP3:POPC:IADD mix of the match-count kernel.  These are the roofline denominators for the INT-bound
// tile kernel (MEASURED_PEAKS.json only has HBM and bf16 peaks).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu && ./int_peak [seconds]
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

// SM clock from NVML (libnvidia-ml.so.1, loaded at run time), sampled by the host while a kernel runs: clock64()
// does not tick at the SM clock on this GPU (round 1 derived 1230 MHz from it while nvidia-smi showed 1965 MHz).
struct Nvml {
    void *lib = nullptr, *dev = nullptr;
    int (*clock_info)(void *, int, unsigned *) = nullptr;
    bool ok = false;
    Nvml() {
        lib = dlopen("libnvidia-ml.so.1", RTLD_NOW);
        if (!lib) return;
        auto init = (int (*)())dlsym(lib, "nvmlInit_v2");
        auto by_pci = (int (*)(const char *, void **))dlsym(lib, "nvmlDeviceGetHandleByPciBusId_v2");
        clock_info = (int (*)(void *, int, unsigned *))dlsym(lib, "nvmlDeviceGetClockInfo");
        if (!init || !by_pci || !clock_info || init() != 0) return;
        char bus[32];
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev_id) != cudaSuccess) return;
        ok = by_pci(bus, &dev) == 0;
    }
    double sm_mhz() const {
        unsigned mhz = 0;
        if (!ok || clock_info(dev, 1 /* NVML_CLOCK_SM */, &mhz) != 0) return 0.0;
        return (double)mhz;
    }
};

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
    int clock_samples;
};

template <int MODE>
Result run(const Nvml &nvml, int sms, int ctas_per_sm, double seconds, uint32_t *out, unsigned long long *cyc) {
    const int grid = sms * ctas_per_sm, block = 256;
    const double ops_per_iter = (MODE == M_MIX ? 14.0 * 2 : 1.0 * UNROLL) * CHAINS;  // counted op: LOP3 in the mix
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int iters = 2000;
    float ms = 0;
    std::vector<double> clocks;
    for (int attempt = 0; attempt < 6; ++attempt) {  // warm up and scale to the wanted duration
        clocks.clear();
        cudaEventRecord(e0);
        peak_kernel<MODE><<<grid, block>>>(out, 0x12345u, 0x9e3779b9u, iters, cyc);
        cudaEventRecord(e1);
        while (cudaEventQuery(e1) == cudaErrorNotReady) {  // sample the SM clock while the kernel runs
            const double mhz = nvml.sm_mhz();
            if (mhz > 0) clocks.push_back(mhz);
            usleep(10000);
        }
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms > seconds * 1000 * 0.8) break;
        double scale = seconds * 1000 / (ms > 0.01 ? ms : 0.01);
        iters = (int)(iters * (scale > 50 ? 50 : scale));
    }
    double total = ops_per_iter * (double)iters * block * grid;
    Result r;
    r.ms = ms;
    r.ops_per_s = total / (ms * 1e-3);
    // median of the NVML samples taken during the final (long) launch; 0 if NVML is unavailable
    std::sort(clocks.begin(), clocks.end());
    r.sm_mhz = clocks.empty() ? 0.0 : clocks[clocks.size() / 2];
    r.clock_samples = (int)clocks.size();
    r.ops_per_clk_sm = r.sm_mhz > 0 ? r.ops_per_s / (sms * r.sm_mhz * 1e6) : 0.0;
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
    Nvml nvml;
    Result l = run<M_LOP3>(nvml, sms, cps, seconds, out, cyc);
    Result p = run<M_POPC>(nvml, sms, cps, seconds, out, cyc);
    Result m = run<M_MIX>(nvml, sms, cps, seconds, out, cyc);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"seconds_each\": %.2f, \"warps_per_sm\": %d,\n", prop.name, sms, seconds, cps * 8);
    printf(" \"lop3_per_s\": %.4e, \"lop3_per_clk_sm\": %.2f, \"lop3_sm_mhz\": %.0f,\n", l.ops_per_s, l.ops_per_clk_sm, l.sm_mhz);
    printf(" \"popc_per_s\": %.4e, \"popc_per_clk_sm\": %.2f, \"popc_sm_mhz\": %.0f,\n", p.ops_per_s, p.ops_per_clk_sm, p.sm_mhz);
    printf(" \"mix_lop3_per_s\": %.4e, \"mix_lop3_per_clk_sm\": %.2f, \"mix_sm_mhz\": %.0f,\n", m.ops_per_s, m.ops_per_clk_sm, m.sm_mhz);
    printf(" \"sm_mhz\": %.0f, \"clock_source\": \"%s\", \"clock_samples\": %d,\n", l.sm_mhz,
           nvml.ok ? "NVML nvmlDeviceGetClockInfo(SM), median of samples taken while the LOP3 kernel ran" : "unavailable",
           l.clock_samples);
    printf(" \"how\": \"16 independent chains per thread, 32 warps per SM, all SMs; volatile inline PTX; LOP3 of the 14:1:1 LOP3:POPC:IADD mix counted for mix_*\"}\n");
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
