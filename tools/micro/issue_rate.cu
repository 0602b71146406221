// Single-warp issue-rate probe for sm_100a: cycles per warp instruction of independent FP32 / integer streams.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rate issue_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP16(x) x x x x x x x x x x x x x x x x
#define N_OUTER 256

template <int MODE>
__global__ void probe(float* out, long long* cycles, int activeLanes, int warpsPerBlock) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a0 = out[lane], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float k = out[32 + lane], m = out[64 + lane];
    int i0 = lane, i1 = lane + 1, i2 = lane + 2, i3 = lane + 3;
    unsigned long long p0 = 0, p1 = 1, p2 = 2, p3 = 3;
    if (lane >= activeLanes) return;
    __syncwarp(__activemask());
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_OUTER; ++it) {
        if (MODE == 0) {  // FADD, 8 independent chains
            REP16(asm volatile("add.rn.f32 %0, %0, %8; add.rn.f32 %1, %1, %8; add.rn.f32 %2, %2, %8; add.rn.f32 %3, %3, %8;"
                               "add.rn.f32 %4, %4, %8; add.rn.f32 %5, %5, %8; add.rn.f32 %6, %6, %8; add.rn.f32 %7, %7, %8;"
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(k));)
        } else if (MODE == 1) {  // FMUL
            REP16(asm volatile("mul.rn.f32 %0, %0, %8; mul.rn.f32 %1, %1, %8; mul.rn.f32 %2, %2, %8; mul.rn.f32 %3, %3, %8;"
                               "mul.rn.f32 %4, %4, %8; mul.rn.f32 %5, %5, %8; mul.rn.f32 %6, %6, %8; mul.rn.f32 %7, %7, %8;"
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(k));)
        } else if (MODE == 2) {  // FADD / FMUL alternating
            REP16(asm volatile("add.rn.f32 %0, %0, %8; mul.rn.f32 %1, %1, %9; add.rn.f32 %2, %2, %8; mul.rn.f32 %3, %3, %9;"
                               "add.rn.f32 %4, %4, %8; mul.rn.f32 %5, %5, %9; add.rn.f32 %6, %6, %8; mul.rn.f32 %7, %7, %9;"
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(k), "f"(m));)
        } else if (MODE == 3) {  // FADD / IADD alternating
            REP16(asm volatile("add.rn.f32 %0, %0, %8; add.s32 %4, %4, 3; add.rn.f32 %1, %1, %8; add.s32 %5, %5, 5;"
                               "add.rn.f32 %2, %2, %8; add.s32 %6, %6, 7; add.rn.f32 %3, %3, %8; add.s32 %7, %7, 9;"
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "f"(k));)
        } else if (MODE == 4) {  // FFMA (3 register operands)
            REP16(asm volatile("fma.rn.f32 %0, %0, %8, %9; fma.rn.f32 %1, %1, %8, %9; fma.rn.f32 %2, %2, %8, %9; fma.rn.f32 %3, %3, %8, %9;"
                               "fma.rn.f32 %4, %4, %8, %9; fma.rn.f32 %5, %5, %8, %9; fma.rn.f32 %6, %6, %8, %9; fma.rn.f32 %7, %7, %8, %9;"
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(k), "f"(m));)
        } else if (MODE == 5) {  // integer adds only
            REP16(asm volatile("add.s32 %0, %0, 3; add.s32 %1, %1, 5; add.s32 %2, %2, 7; add.s32 %3, %3, 9;"
                               "add.s32 %0, %0, 3; add.s32 %1, %1, 5; add.s32 %2, %2, 7; add.s32 %3, %3, 9;"
                               : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3));)
        } else if (MODE == 6) {  // packed add.f32x2 (8 instructions = 16 lane-adds)
            REP16(asm volatile("add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                               "add.rn.f32x2 %0, %0, %4; add.rn.f32x2 %1, %1, %4; add.rn.f32x2 %2, %2, %4; add.rn.f32x2 %3, %3, %4;"
                               : "+l"(p0), "+l"(p1), "+l"(p2), "+l"(p3) : "l"(p3 + 1));)
        } else if (MODE == 7) {  // dependent FADD chain (latency)
            REP16(asm volatile("add.rn.f32 %0, %0, %1; add.rn.f32 %0, %0, %1; add.rn.f32 %0, %0, %1; add.rn.f32 %0, %0, %1;"
                               "add.rn.f32 %0, %0, %1; add.rn.f32 %0, %0, %1; add.rn.f32 %0, %0, %1; add.rn.f32 %0, %0, %1;"
                               : "+f"(a0) : "f"(k));)
        } else if (MODE == 8) {  // two dependent chains interleaved (ILP 2)
            REP16(asm volatile("add.rn.f32 %0, %0, %2; add.rn.f32 %1, %1, %2; add.rn.f32 %0, %0, %2; add.rn.f32 %1, %1, %2;"
                               "add.rn.f32 %0, %0, %2; add.rn.f32 %1, %1, %2; add.rn.f32 %0, %0, %2; add.rn.f32 %1, %1, %2;"
                               : "+f"(a0), "+f"(a1) : "f"(k));)
        } else if (MODE >= 10) {  // loop-body size probe: FADD x4 (ILP 4) + IADD x4 blocks, body = 8 * 16 * 2^(MODE-10) instructions
#define BLK REP16(asm volatile("add.rn.f32 %0, %0, %8; add.s32 %4, %4, 3; add.rn.f32 %1, %1, %8; add.s32 %5, %5, %6;" \
                               "add.rn.f32 %2, %2, %8; add.s32 %6, %6, %7; add.rn.f32 %3, %3, %8; add.s32 %7, %7, %4;" \
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "f"(k));)
            BLK
            if (MODE >= 11) { BLK }
            if (MODE >= 12) { BLK BLK }
            if (MODE >= 13) { BLK BLK BLK BLK }
            if (MODE >= 14) { BLK BLK BLK BLK BLK BLK BLK BLK }
            if (MODE >= 15) { BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK BLK }
        } else if (MODE == 9) {  // FSEL-like: selp on a fixed predicate
            REP16(asm volatile("{.reg .pred p; setp.gt.f32 p, %8, 0f00000000;"
                               "selp.f32 %0, %1, %0, p; selp.f32 %1, %2, %1, p; selp.f32 %2, %3, %2, p; selp.f32 %3, %4, %3, p;"
                               "selp.f32 %4, %5, %4, p; selp.f32 %5, %6, %5, p; selp.f32 %6, %7, %6, p; selp.f32 %7, %0, %7, p;}"
                               : "+f"(a0), "+f"(a1), "+f"(a2), "+f"(a3), "+f"(a4), "+f"(a5), "+f"(a6), "+f"(a7) : "f"(k));)
        }
    }
    long long t1 = clock64();
    out[128 + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + (float)(i0 + i1 + i2 + i3) + (float)(p0 + p1 + p2 + p3);
    if (lane == 0) cycles[blockIdx.x * warpsPerBlock + warp] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instrPerRep, int warps, int lanes, float* out, long long* cyc) {
    probe<MODE><<<1, warps * 32>>>(out, cyc, lanes, warps);
    cudaDeviceSynchronize();
    probe<MODE><<<1, warps * 32>>>(out, cyc, lanes, warps);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    const double n = (double)N_OUTER * 16 * instrPerRep;
    printf("%-34s warps=%d lanes=%2d  cycles/instr:", name, warps, lanes);
    for (int w = 0; w < warps; ++w) printf(" %.2f", h[w] / n);
    printf("\n");
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 64 * 8);
    cudaMemset(out, 0, 4096 * 4);
    for (int lanes : {32, 16}) {
        run<0>("FADD x8 independent", 8, 1, lanes, out, cyc);
        run<1>("FMUL x8 independent", 8, 1, lanes, out, cyc);
        run<2>("FADD/FMUL alternating", 8, 1, lanes, out, cyc);
        run<3>("FADD/IADD alternating", 8, 1, lanes, out, cyc);
        run<4>("FFMA 3-reg", 8, 1, lanes, out, cyc);
        run<5>("IADD", 8, 1, lanes, out, cyc);
        run<6>("add.f32x2 (per instruction)", 8, 1, lanes, out, cyc);
        run<7>("FADD dependent chain (latency)", 8, 1, lanes, out, cyc);
        run<8>("FADD two chains (ILP 2)", 8, 1, lanes, out, cyc);
        run<9>("SELP.f32 x8", 8, 1, lanes, out, cyc);
    }
    run<10>("body 128 instr (2 KB)", 8, 1, 32, out, cyc);
    run<11>("body 256 instr (4 KB)", 16, 1, 32, out, cyc);
    run<12>("body 512 instr (8 KB)", 32, 1, 32, out, cyc);
    run<13>("body 1024 instr (16 KB)", 64, 1, 32, out, cyc);
    run<14>("body 2048 instr (32 KB)", 128, 1, 32, out, cyc);
    run<15>("body 4096 instr (64 KB)", 256, 1, 32, out, cyc);
    run<13>("body 1024 instr (16 KB) 4 warps", 64, 4, 32, out, cyc);
    run<14>("body 2048 instr (32 KB) 4 warps", 128, 4, 32, out, cyc);
    // several warps in one block: which share a scheduler?  (warp w and w + 4 should, if SMSP = wid % 4)
    for (int warps : {2, 4, 5, 8}) {
        run<0>("FADD x8, several warps", 8, warps, 32, out, cyc);
        run<2>("FADD/FMUL, several warps", 8, warps, 32, out, cyc);
    }
    return 0;
}
