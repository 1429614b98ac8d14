// Pipe model of an sm_100a scheduler (SMSP) for the instruction mix of k_sweep.  Every thread runs 8 independent DADD
// chains and, per DADD, N independent 32-bit instructions of one kind on its own registers; all SMs filled, 16 warps
// per SM (the occupancy of k_sweep).  Prints cycles per (DADD + N x) group per scheduler.  Measured (B200): the FP64 pipe,
// the ALU pipe (LOP3 / SEL / FSEL / IADD3 / VIADD / SHF / ISETP) and the FMA pipe (IMAD, FFMA) each take a warp
// instruction every TWO cycles and run side by side; the scheduler issues one instruction per cycle.  A group of
// 1 DADD + N LOP3 therefore costs max(2, 2 N) cycles: the ALU pipe, not the FP64 pipe, bounds the max-plus candidate as the
// compiler emits it (2 DADD + DSETP on the FP64 pipe; 2 FSEL + the predicated back-pointer OR on the ALU pipe, next to
// every address LOP3 of the kernel).  Moving the two selects to the FMA pipe (predicated IMAD by an opaque 1) balances
// the three pipes; the variants below measure that.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_model issue_model.cu && ./issue_model
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int KIND>
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
        if (KIND == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i][n]) : "r"(useed), "r"(it));   // opaque multiplier: a real IMAD
        else if (KIND == 1) asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n selp.b32 %0, %0, %1, p;\n}" : "+r"(r[i][n]) : "r"(useed));   // the compare is uniform and hoisted: one SEL
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

// the max-plus candidate: c0 = R0 + T0, c1 = R1 + T1, p = c1 > c0, O = p ? c1 : c0, @p bits += mask.
//   V = 0: as k_sweep issued it until round 2 (selp.f64 = two FSEL, predicated OR: three ALU-pipe instructions)
//   V = 1: the winner is moved by two predicated IMADs (x * one + 0, `one` an opaque kernel argument: FMA pipe), OR on the ALU pipe
//   V = 2: both selects and the back-pointer as predicated IMADs
//   V = 3: one FSEL (high word) + one predicated IMAD (low word), OR on the ALU pipe
template <int V>
__global__ void __launch_bounds__(512, 1) k_candidate(double *out, int iters, double seed, unsigned one, double done = 1.0) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 1e-9 + i;
  const double t0 = seed * 1e-6 - 0.1, t1 = seed * 1e-6 - 0.2;
  unsigned bits = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double c0 = a[i] + t0, c1 = a[(i + 1) & 7] + t1;
      if (V == 0)
        asm volatile("{\n .reg .pred p;\n setp.gt.f64 p, %2, %3;\n selp.f64 %0, %2, %3, p;\n @p or.b32 %1, %1, %4;\n}"
                     : "=d"(a[i]), "+r"(bits) : "d"(c1), "d"(c0), "n"(1 << 3));
      else if (V == 1)
        asm volatile("{\n .reg .pred p;\n .reg .b32 lo, hi, l1, h1;\n setp.gt.f64 p, %2, %3;\n mov.b64 {lo, hi}, %3;\n mov.b64 {l1, h1}, %2;\n"
                     " @p mad.lo.u32 lo, l1, %5, 0;\n @p mad.lo.u32 hi, h1, %5, 0;\n mov.b64 %0, {lo, hi};\n @p or.b32 %1, %1, %4;\n}"
                     : "=d"(a[i]), "+r"(bits) : "d"(c1), "d"(c0), "n"(1 << 3), "r"(one));
      else if (V == 2)
        asm volatile("{\n .reg .pred p;\n .reg .b32 lo, hi, l1, h1;\n setp.gt.f64 p, %2, %3;\n mov.b64 {lo, hi}, %3;\n mov.b64 {l1, h1}, %2;\n"
                     " @p mad.lo.u32 lo, l1, %5, 0;\n @p mad.lo.u32 hi, h1, %5, 0;\n mov.b64 %0, {lo, hi};\n @p mad.lo.u32 %1, %5, %4, %1;\n}"
                     : "=d"(a[i]), "+r"(bits) : "d"(c1), "d"(c0), "n"(1 << 3), "r"(one));
      else if (V == 4) {
        // RECOMPUTE: c0 lands in the output register, the winner is recomputed by a predicated DFMA (x * 1.0 + t, `1.0` opaque:
        // bit-identical to the add, and ptxas cannot merge it with the first evaluation)
        double o;
        asm volatile("{\n .reg .pred p;\n .reg .f64 c1;\n add.f64 %0, %2, %3;\n add.f64 c1, %4, %5;\n setp.gt.f64 p, c1, %0;\n"
                     " @p fma.rn.f64 %0, %4, %7, %5;\n @p or.b32 %1, %1, %6;\n}"
                     : "=&d"(o), "+r"(bits) : "d"(a[i]), "d"(t0), "d"(a[(i + 1) & 7]), "d"(t1), "n"(1 << 3), "d"(done));
        a[i] = o;
      } else
        asm volatile("{\n .reg .pred p;\n .reg .b32 lo, hi, l1, h1;\n setp.gt.f64 p, %2, %3;\n mov.b64 {lo, hi}, %3;\n mov.b64 {l1, h1}, %2;\n"
                     " @p mad.lo.u32 lo, l1, %5, 0;\n selp.b32 hi, h1, hi, p;\n mov.b64 %0, {lo, hi};\n @p or.b32 %1, %1, %4;\n}"
                     : "=d"(a[i]), "+r"(bits) : "d"(c1), "d"(c0), "n"(1 << 3), "r"(one));
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
#define RUN(N, KIND)                                                                                      \
  { float ms = time_ms([&] { k_mix<N, KIND><<<sms, threads>>>(d, iters, 1.0, 7u); });                     \
    printf("{\"mix\": \"1 DADD + %d %s\", \"cycles_per_group\": %.3f, \"three_pipe_model\": %d}\n", N, KIND == 2 ? "IMAD" : KIND ? "SEL" : "LOP3", \
           ms * 1e-3 * khz * 1e3 / groups_per_sched, 2 * N > 2 ? 2 * N : 2); }
  RUN(0, 0) RUN(1, 0) RUN(2, 0) RUN(3, 0) RUN(4, 0) RUN(6, 0) RUN(1, 1) RUN(2, 1) RUN(1, 2) RUN(2, 2) RUN(3, 2)
#define CAND(V, WHAT, MODEL)                                                                              \
  { float ms = time_ms([&] { k_candidate<V><<<sms, threads>>>(d, iters, 1.0, 1u, 1.0); });                     \
    const double cyc = ms * 1e-3 * khz * 1e3 / groups_per_sched;                                         \
    printf("{\"mix\": \"max-plus candidate pair: 2 DADD + DSETP + %s\", \"cycles_per_group\": %.3f, \"three_pipe_model\": %d, \"fp64_pipe_frac\": %.3f}\n", WHAT, cyc, MODEL, 6.0 / cyc); }
  CAND(0, "2 FSEL + predicated OR (ALU pipe 3)", 6)
  CAND(1, "2 predicated IMAD + predicated OR (FMA pipe 2, ALU pipe 1)", 6)
  CAND(2, "3 predicated IMAD (FMA pipe 3)", 6)
  CAND(4, "predicated DFMA recompute + predicated OR (FP64 pipe 4, ALU pipe 1)", 8)
  CAND(3, "1 FSEL + 1 predicated IMAD + predicated OR (ALU 2, FMA 1)", 6)
  cudaFree(d);
  return 0;
}
