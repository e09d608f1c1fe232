// Micro-probe: issue rate (warp-instr/clk/SMSP) of the Blackwell packed-fp32 instructions (FFMA2 / FADD2 / FMUL2) next
// to their scalar forms, alone and mixed with the integer multiply Philox uses, at 2 and 8 warps per SMSP.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/probes/f32x2_probe.cu -o scripts/probes/f32x2_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int KIND>
__global__ void k(float* out, int iters) {
  float a[8]; u64 p[8]; uint32_t q[8]; u64 w[8];
  for (int i = 0; i < 8; ++i) {
    a[i] = threadIdx.x * 1e-3f + i;
    float2 f = make_float2(a[i], a[i] + 0.5f);
    p[i] = *reinterpret_cast<u64*>(&f);
    q[i] = threadIdx.x * 2654435761u + i; w[i] = q[i];
  }
  const float2 c2 = make_float2(0.999f, 1.001f);
  const u64 cc = *reinterpret_cast<const u64*>(&c2);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (KIND == 0) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(a[i]));
        if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
        if (KIND == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        if (KIND == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        if (KIND == 4) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
          uint32_t lo = (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
          asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"(lo));
        }
        if (KIND == 5) {
          asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(a[i]));
          uint32_t lo = (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
          asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(w[i]) : "r"(lo));
        }
        if (KIND == 6) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(q[i]));
        if (KIND == 7) {  // 4 scalar FFMA per MUFU
          asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, 0f3A83126F;" : "+f"(a[i]));
          if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 1) & 7]));
        }
        if (KIND == 8) {  // 4 FFMA2 per MUFU
          asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(cc));
          if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 1) & 7]));
        }
        if (KIND == 9) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) { float2 f = *reinterpret_cast<float2*>(&p[i]); s += a[i] + f.x + f.y + (float)q[i] + (float)(uint32_t)w[i] + (float)(uint32_t)(w[i] >> 32); }
  if (s == 0.1234567f) out[0] = s;
}
template <int KIND>
void run(const char* name, double instr_per_slot) {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
  for (int wps : {1, 2, 8}) {
    const int blocks = 148 * wps, threads = 128, iters = 4096;
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(t0); k<KIND><<<blocks, threads>>>(d, iters); cudaEventRecord(t1); cudaEventSynchronize(t1);
      float ms; cudaEventElapsedTime(&ms, t0, t1); if (rep && ms < best) best = ms;
    }
    double winstr = (double)blocks * threads / 32 * iters * 64.0 * instr_per_slot;
    printf("%-34s warps/SMSP %d: %.3f warp-instr/clk/SMSP (1.965 GHz assumed)\n", name, wps, winstr / (best * 1e-3) / (148 * 4) / 1.965e9);
  }
}
int main() {
  run<0>("FFMA", 1);
  run<1>("FFMA2", 1);
  run<2>("FADD2", 1);
  run<3>("FMUL2", 1);
  run<4>("FFMA2 + IMAD.WIDE + LOP3", 3);
  run<5>("FFMA + IMAD.WIDE + LOP3", 3);
  run<6>("MUFU.EX2.F16x2", 1);
  run<7>("4 FFMA : 1 MUFU.EX2", 1.25);
  run<8>("4 FFMA2 : 1 MUFU.EX2", 1.25);
  run<9>("MUFU.EX2", 1);
  return 0;
}
