// Which warps of two co-resident CTAs share a scheduler on sm_100a?  296 CTAs x 4 warps, 110 KB of shared memory each
// (two per SM, like the sequential solver kernels); in every CTA one warp runs an issue-bound FADD loop, the others
// exit.  The worker warp is `wa` in CTAs whose warp 0 sits in an even group of four hardware warp slots, `wb` in the
// others.  ~1.1 cycles / instruction = the two workers of an SM have a scheduler each, ~2.05 = they share one.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP16(x) x x x x x x x x x x x x x x x x

__global__ void probe(float* out, long long* cycles, int* info, int wa, int wb, int rule) {
    extern __shared__ unsigned char smem[];
    __shared__ int sType;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t slot, smid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(slot));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) sType = rule == 0 ? (int)((slot >> 2) & 1u) : (int)(blockIdx.x & 1);
    if (lane == 0) { info[(blockIdx.x * 4 + warp) * 2] = (int)smid; info[(blockIdx.x * 4 + warp) * 2 + 1] = (int)slot; }
    __syncthreads();
    const int worker = sType ? wb : wa;
    if (warp != worker) return;
    float a0 = out[lane], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float k = out[32 + lane];
    smem[threadIdx.x] = 1;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 2048; ++it) {
        REP16(asm volatile("add.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; add.rn.f32 %3, %3, %8;"
                           "add.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; add.rn.f32 %6, %6, %8; add.rn.f32 %7, %7, %8;"
                           : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(k));)
    }
    long long t1 = clock64();
    out[128 + lane] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc; int* info;
    const int ctas = 296;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, ctas * 8); cudaMalloc(&info, ctas * 8 * 4);
    cudaMemset(out, 0, 4096 * 4);
    const int smemBytes = 110 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
    static long long h[ctas]; static int hi[ctas * 8];
    for (int rule = 0; rule < 2; ++rule) {
        printf("CTA type = %s\n", rule == 0 ? "(%warpid of warp 0 >> 2) & 1" : "blockIdx & 1");
        for (int wa = 0; wa < 4; ++wa) {
            for (int wb = 0; wb < 4; ++wb) {
                probe<<<ctas, 128, smemBytes>>>(out, cyc, info, wa, wb, rule);
                cudaDeviceSynchronize();
                cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                double worst = 0, sum = 0;
                for (int i = 0; i < ctas; ++i) { double c = h[i] / (2048.0 * 128); worst = c > worst ? c : worst; sum += c; }
                printf("  wa=%d wb=%d  cycles/instr mean %.2f max %.2f\n", wa, wb, sum / ctas, worst);
            }
        }
    }
    cudaMemcpy(hi, info, sizeof(hi), cudaMemcpyDeviceToHost);
    printf("warp slots (%%warpid) of the CTAs on SM 0 and SM 1:\n");
    for (int i = 0; i < ctas; ++i)
        if (hi[i * 8] < 2) printf("  cta %3d on sm %d: slots %d %d %d %d\n", i, hi[i * 8], hi[i * 8 + 1], hi[i * 8 + 3], hi[i * 8 + 5], hi[i * 8 + 7]);
    // three-warp CTAs, as the round-1 kernels were launched
    return 0;
}
