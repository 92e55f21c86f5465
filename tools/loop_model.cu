// Micro-model of the tile kernel's main loop, used to choose the thread tile and operand-load width.
// Per iteration (one 64-bin word) a thread holding a TI x TJ pair tile issues 28*TI*TJ LOP3 (dependent chains per
// pair), 2*TI*TJ POPC + IMAD, and loads its operands from a 3-stage shared-memory ring either as LDS.64 (plane
// pairs) or LDS.128 (plane quads); the CTA synchronises every SYNC_EVERY iterations (0 = never).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int STAGE = 2 * 8 * 64;  // uint2 per operand stage, padded to 8 plane pairs per half

template <int TI, int TJ, int VEC, int SYNC_EVERY>
__global__ void __launch_bounds__(256) loop_kernel(uint32_t *out, int iters, uint32_t one) {
    extern __shared__ uint2 sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < 3 * 2 * STAGE; e += blockDim.x) sm[e] = make_uint2(e * 2654435761u, e * 40503u + one);
    __syncthreads();
    const int arow0 = lane >> 3, bcol0 = lane & 7;  // row operands broadcast over 8 lanes, column operands over 4
    uint32_t mism[TI * TJ];
#pragma unroll
    for (int c = 0; c < TI * TJ; ++c) mism[c] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (SYNC_EVERY > 0 && it % SYNC_EVERY == 0) __syncthreads();
        const uint2 *sa = sm + (it % 3) * (2 * STAGE), *sb = sa + STAGE;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t d[TI * TJ];
            if (VEC == 2) {
#pragma unroll
                for (int pp = 0; pp < 7; ++pp) {
                    uint2 av[TI], bv[TJ];
#pragma unroll
                    for (int i = 0; i < TI; ++i) av[i] = sa[(half * 8 + pp) * 64 + (arow0 + 4 * i)];
#pragma unroll
                    for (int j = 0; j < TJ; ++j) bv[j] = sb[(half * 8 + pp) * 64 + (bcol0 + 8 * j)];
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j) {
                            if (pp == 0) d[i * TJ + j] = av[i].x ^ bv[j].x; else d[i * TJ + j] |= av[i].x ^ bv[j].x;
                            d[i * TJ + j] |= av[i].y ^ bv[j].y;
                        }
                }
            } else {
                const uint4 *sa4 = reinterpret_cast<const uint4 *>(sa), *sb4 = reinterpret_cast<const uint4 *>(sb);
#pragma unroll
                for (int pq = 0; pq < 4; ++pq) {
                    uint4 av[TI], bv[TJ];
#pragma unroll
                    for (int i = 0; i < TI; ++i) av[i] = sa4[(half * 4 + pq) * 64 + (arow0 + 4 * i)];
#pragma unroll
                    for (int j = 0; j < TJ; ++j) bv[j] = sb4[(half * 4 + pq) * 64 + (bcol0 + 8 * j)];
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j) {
                            if (pq == 0) d[i * TJ + j] = av[i].x ^ bv[j].x; else d[i * TJ + j] |= av[i].x ^ bv[j].x;
                            d[i * TJ + j] |= av[i].y ^ bv[j].y;
                            if (pq < 3) {  // planes 12, 13 live in the first half of the last quad
                                d[i * TJ + j] |= av[i].z ^ bv[j].z;
                                d[i * TJ + j] |= av[i].w ^ bv[j].w;
                            }
                        }
                }
            }
#pragma unroll
            for (int c = 0; c < TI * TJ; ++c) mism[c] = __popc(d[c]) * one + mism[c];
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < TI * TJ; ++c) acc ^= mism[c];
    out[blockIdx.x * blockDim.x + tid] = acc;
}

template <int TI, int TJ, int VEC, int SE>
void run(int sms, int ctas_per_sm, int threads, uint32_t *out) {
    const int grid = sms * ctas_per_sm;
    const size_t smem = 3 * 2 * STAGE * sizeof(uint2);
    cudaFuncSetAttribute(loop_kernel<TI, TJ, VEC, SE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, loop_kernel<TI, TJ, VEC, SE>);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 16000 * 16 / (TI * TJ);
    loop_kernel<TI, TJ, VEC, SE><<<grid, threads, smem>>>(out, 1000, 1);
    cudaEventRecord(e0);
    loop_kernel<TI, TJ, VEC, SE><<<grid, threads, smem>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, loop_kernel<TI, TJ, VEC, SE>, threads, smem);
    printf("%dx%d  LDS.%-3d %7d %7d(%d) %5d %6d   %.2f\n", TI, TJ, VEC * 32, threads / 32, ctas_per_sm, occ, SE, fa.numRegs,
           28.0 * TI * TJ * iters * threads * grid / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t *out; cudaMalloc(&out, (size_t)sms * 8 * 1024 * 4);
    printf("tile  load   warps/CTA CTAs/SM(max) sync regs   TLOP3/s\n");
    run<4, 4, 2, 0>(sms, 2, 256, out); run<4, 4, 2, 1>(sms, 2, 256, out);
    run<4, 4, 4, 0>(sms, 2, 256, out); run<4, 4, 4, 1>(sms, 2, 256, out);
    run<4, 8, 2, 0>(sms, 2, 256, out); run<4, 8, 2, 1>(sms, 2, 256, out);
    run<4, 8, 4, 0>(sms, 2, 256, out); run<4, 8, 4, 1>(sms, 2, 256, out);
    run<8, 4, 2, 1>(sms, 2, 256, out); run<8, 4, 4, 1>(sms, 2, 256, out);
    run<4, 8, 2, 1>(sms, 1, 256, out); run<4, 8, 4, 1>(sms, 1, 256, out);
    run<4, 8, 4, 1>(sms, 1, 512, out); run<8, 8, 4, 1>(sms, 1, 256, out);
    run<4, 4, 4, 1>(sms, 3, 256, out); run<4, 4, 4, 1>(sms, 4, 256, out);
    run<2, 4, 4, 1>(sms, 4, 256, out); run<4, 2, 4, 1>(sms, 4, 256, out);
    return 0;
}
