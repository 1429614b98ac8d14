// Issue model of the FP64 CUDA-core pipe on sm_100a: how many issue cycles does a warp-level FP64 instruction cost when
// it is interleaved with full-rate integer / select instructions?  Every thread runs 8 independent DADD chains and,
// per DADD, N independent 32-bit ALU instructions (N = 0..6) on its own registers; all SMs filled, 16 warps per SM
// (the occupancy of k_sweep).  Prints cycles per (DADD + N ALU) group per scheduler.  If FP64 instructions only
// occupied their own half-rate pipe, the cost would be max(2, 1 + N); measured it is 2 + N: an FP64 instruction holds
// the scheduler's issue port for two cycles.  The max-plus candidate of k_sweep (2 DADD + DSETP + 2 FSEL + 1 predicated
// integer op) therefore costs 9 issue cycles for 3 FP64 operations: ceiling 6/9 of the DADD rate before any load,
// store or address instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model issue_model.cu && ./issue_model
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int FSEL>
__global__ void __launch_bounds__(512, 1) k_mix(double *out, int iters, double seed, unsigned useed) {
  double a[8];
  unsigned r[8][N > 0 ? N : 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int n = 0; n < (N > 0 ? N : 1); ++n) r[i][n] = useed + threadIdx.x + 8 * n + i;
  }
  const double b = seed * 1e-6 + 1.0000001;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("add.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b));
#pragma unroll
      for (int n = 0; n < N; ++n) {
        if (FSEL) asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n selp.b32 %0, %0, %1, p;\n}" : "+r"(r[i][n]) : "r"(useed));   // the compare is uniform and hoisted: one SEL
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i][n]) : "r"(useed), "r"(it));
      }
    }
  }
  double s = 0;
  unsigned u = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s += a[i];
#pragma unroll
    for (int n = 0; n < (N > 0 ? N : 1); ++n) u ^= r[i][n];
  }
  if (s == 12345.678 || u == 0x12345u) out[0] = s + u;
}

// the max-plus candidate exactly as k_sweep issues it: c0 = R0 + T0, c1 = R1 + T1, p = c1 > c0, O = p ? c1 : c0 (two
// 32-bit selects), @p bits += mask
__global__ void __launch_bounds__(512, 1) k_candidate(double *out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
  const double t0 = seed * 1e-6 - 0.1, t1 = seed * 1e-6 - 0.2;
  unsigned bits = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double c0 = a[i] + t0, c1 = a[(i + 1) & 7] + t1;
      asm volatile("{\n .reg .pred p;\n setp.gt.f64 p, %2, %3;\n selp.f64 %0, %2, %3, p;\n @p or.b32 %1, %1, %4;\n}"
                   : "=d"(a[i]), "+r"(bits) : "d"(c1), "d"(c0), "n"(1 << 3));
    }
  }
  double s = bits;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount, iters = 1 << 13, threads = 512;
  double *d; cudaMalloc(&d, 64);
  const double groups_per_sched = (double)(threads / 32 / 4) * iters * 8.0;   // per launch with one CTA per SM
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"note\": \"cycles per group per scheduler at the nominal clock; 4 warps per scheduler\"}\n", prop.name, sms, khz / 1000);
#define RUN(N, FS)                                                                                        \
  { float ms = time_ms([&] { k_mix<N, FS><<<sms, threads>>>(d, iters, 1.0, 7u); });                       \
    printf("{\"mix\": \"1 DADD + %d %s\", \"cycles_per_group\": %.3f, \"pipe_only_model\": %d, \"issue_port_model\": %d}\n", N, FS ? "SEL" : "LOP3", \
           ms * 1e-3 * khz * 1e3 / groups_per_sched, 1 + N > 2 ? 1 + N : 2, 2 + N); }
  RUN(0, 0) RUN(1, 0) RUN(2, 0) RUN(3, 0) RUN(4, 0) RUN(6, 0) RUN(1, 1) RUN(2, 1)
  { float ms = time_ms([&] { k_candidate<<<sms, threads>>>(d, iters, 1.0); });
    const double cyc = ms * 1e-3 * khz * 1e3 / groups_per_sched;
    printf("{\"mix\": \"max-plus candidate pair: 2 DADD + DSETP + 2 FSEL + predicated OR\", \"cycles_per_group\": %.3f, \"pipe_only_model\": 6, \"issue_port_model\": 9, \"fp64_pipe_frac\": %.3f}\n", cyc, 6.0 / cyc); }
  cudaFree(d);
  return 0;
}
