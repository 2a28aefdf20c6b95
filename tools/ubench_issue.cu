// ubench_issue.cu — two questions the replica engine's design rests on, asked of the B200 itself:
//   1. does the packed FP32 instruction (fma.rn.f32x2, SASS FFMA2) retire two FMAs per issue slot?
//      (the pair loops are issue-bound, not pipe-bound: DESIGN.md §4)
//   2. where do the CTAs of 2-CTA clusters land when two CTAs fit on one SM (bid -> smid), and how
//      many clusters are resident at once?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_issue tools/ubench_issue.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void __launch_bounds__(1024, 1) k_ffma(float *out, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3f + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__global__ void __launch_bounds__(1024, 1) k_ffma2(float *out, int iters, float a, float b) {
  unsigned long long x[8];
  const unsigned long long a2 = pack2(a, a), b2 = pack2(b, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = pack2(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(a2), "l"(b2));
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { float lo, hi; unpack2(x[k], lo, hi); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// a stand-in for the Debye-Hueckel pair body: 1 LDS.128, ~14 FP32, 2 MUFU per pair (scalar) ...
__global__ void __launch_bounds__(1024, 1) k_pair_scalar(float *out, int iters, float nk, float rc2) {
  __shared__ float4 pos[1024];
  pos[threadIdx.x] = make_float4(threadIdx.x * 0.37f, threadIdx.x * 0.11f, threadIdx.x * 0.23f, 1.0f);
  __syncthreads();
  const float4 pi = pos[threadIdx.x];
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll 8
    for (int j = 0; j < 64; ++j) {
      const float4 pj = pos[(threadIdx.x + 1 + j + i) & 1023];
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
      const float rinv = rsqrtf(r2), r = r2 * rinv, u = nk * r;
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u));
      const float w = e * ((rinv * pj.w) * (rinv * rinv));
      float fr = fmaf(w * u, -0.693147180559945f, w);
      fr = r2 < rc2 ? fr : 0.f;
      fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = fx + fy + fz;
}

// ... and the same with two partners per lane in packed FP32 (MUFU stays scalar)
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  // a - b = fma(b, -1, a)
  const unsigned long long m1 = 0xBF800000BF800000ull;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(b), "l"(m1), "l"(a));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__global__ void __launch_bounds__(1024, 1) k_pair_packed(float *out, int iters, float nk, float rc2) {
  __shared__ float4 pos[1024];
  pos[threadIdx.x] = make_float4(threadIdx.x * 0.37f, threadIdx.x * 0.11f, threadIdx.x * 0.23f, 1.0f);
  __syncthreads();
  const float4 pi = pos[threadIdx.x];
  const unsigned long long pix = pack2(pi.x, pi.x), piy = pack2(pi.y, pi.y), piz = pack2(pi.z, pi.z);
  const unsigned long long nk2 = pack2(nk, nk), ln2 = pack2(-0.693147180559945f, -0.693147180559945f);
  unsigned long long fx = 0ull, fy = 0ull, fz = 0ull;
  for (int i = 0; i < iters; ++i) {
#pragma unroll 4
    for (int j = 0; j < 64; j += 2) {
      const float4 pa = pos[(threadIdx.x + 1 + j + i) & 1023], pb = pos[(threadIdx.x + 2 + j + i) & 1023];
      const unsigned long long dx = sub2(pix, pack2(pa.x, pb.x)), dy = sub2(piy, pack2(pa.y, pb.y)), dz = sub2(piz, pack2(pa.z, pb.z));
      const unsigned long long r2 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
      float r2a, r2b;
      unpack2(r2, r2a, r2b);
      const float ria = rsqrtf(r2a), rib = rsqrtf(r2b);
      const unsigned long long rinv = pack2(ria, rib);
      const unsigned long long r = mul2(r2, rinv), u = mul2(nk2, r);
      float ua, ub, ea, eb;
      unpack2(u, ua, ub);
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(ua));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(ub));
      const unsigned long long w = mul2(pack2(ea, eb), mul2(mul2(rinv, pack2(pa.w, pb.w)), mul2(rinv, rinv)));
      unsigned long long fr = fma2(mul2(w, u), ln2, w);
      float fa, fb;
      unpack2(fr, fa, fb);
      fa = r2a < rc2 ? fa : 0.f; fb = r2b < rc2 ? fb : 0.f;
      fr = pack2(fa, fb);
      fx = fma2(fr, dx, fx); fy = fma2(fr, dy, fy); fz = fma2(fr, dz, fz);
    }
  }
  float a, b, s = 0.f;
  unpack2(fx, a, b); s += a + b; unpack2(fy, a, b); s += a + b; unpack2(fz, a, b); s += a + b;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MUFU throughput: 8 independent chains of rsqrt / ex2 per thread
__global__ void __launch_bounds__(1024, 1) k_mufu(float *out, int iters, int which) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = 1.0f + threadIdx.x * 1e-3f + k;
  if (which == 0) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
    }
  } else {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// pair body with exp(-kappa r) as a degree-6 polynomial on the FMA pipe (one MUFU per pair: the rsqrt)
#define EXP_POLY(t) fmaf(fmaf(fmaf(fmaf(fmaf(fmaf(0.00083676545f, (t), -0.0076200562f), (t), 0.041187618f), (t), -0.16649818f), (t), 0.49997148f), (t), -0.99999815f), (t), 1.0f)
__global__ void __launch_bounds__(1024, 1) k_pair_scalar_poly(float *out, int iters, float kappa, float rc2) {
  __shared__ float4 pos[1024];
  pos[threadIdx.x] = make_float4(threadIdx.x * 0.37f, threadIdx.x * 0.11f, threadIdx.x * 0.23f, 1.0f);
  __syncthreads();
  const float4 pi = pos[threadIdx.x];
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll 8
    for (int j = 0; j < 64; ++j) {
      const float4 pj = pos[(threadIdx.x + 1 + j + i) & 1023];
      const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      const float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
      const float rinv = rsqrtf(r2), t = kappa * (r2 * rinv);
      const float e = EXP_POLY(t);
      const float w = e * ((rinv * pj.w) * (rinv * rinv));
      float fr = fmaf(w, t, w);
      fr = r2 < rc2 ? fr : 0.f;
      fx = fmaf(fr, dx, fx); fy = fmaf(fr, dy, fy); fz = fmaf(fr, dz, fz);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = fx + fy + fz;
}

__global__ void __launch_bounds__(1024, 1) k_pair_packed_poly(float *out, int iters, float kappa, float rc2) {
  __shared__ float4 pos[1024];
  pos[threadIdx.x] = make_float4(threadIdx.x * 0.37f, threadIdx.x * 0.11f, threadIdx.x * 0.23f, 1.0f);
  __syncthreads();
  const float4 pi = pos[threadIdx.x];
  const unsigned long long pix = pack2(pi.x, pi.x), piy = pack2(pi.y, pi.y), piz = pack2(pi.z, pi.z);
  const unsigned long long k2 = pack2(kappa, kappa);
  const unsigned long long c6 = pack2(0.00083676545f, 0.00083676545f), c5 = pack2(-0.0076200562f, -0.0076200562f),
                           c4 = pack2(0.041187618f, 0.041187618f), c3 = pack2(-0.16649818f, -0.16649818f),
                           c2 = pack2(0.49997148f, 0.49997148f), c1 = pack2(-0.99999815f, -0.99999815f), c0 = pack2(1.0f, 1.0f);
  unsigned long long fx = 0ull, fy = 0ull, fz = 0ull;
  for (int i = 0; i < iters; ++i) {
#pragma unroll 4
    for (int j = 0; j < 64; j += 2) {
      const float4 pa = pos[(threadIdx.x + 1 + j + i) & 1023], pb = pos[(threadIdx.x + 2 + j + i) & 1023];
      const unsigned long long dx = sub2(pix, pack2(pa.x, pb.x)), dy = sub2(piy, pack2(pa.y, pb.y)), dz = sub2(piz, pack2(pa.z, pb.z));
      const unsigned long long r2 = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
      float r2a, r2b;
      unpack2(r2, r2a, r2b);
      const unsigned long long rinv = pack2(rsqrtf(r2a), rsqrtf(r2b));
      const unsigned long long t = mul2(k2, mul2(r2, rinv));
      const unsigned long long e = fma2(fma2(fma2(fma2(fma2(fma2(c6, t, c5), t, c4), t, c3), t, c2), t, c1), t, c0);
      const unsigned long long w = mul2(e, mul2(mul2(rinv, pack2(pa.w, pb.w)), mul2(rinv, rinv)));
      unsigned long long fr = fma2(w, t, w);
      float fa, fb;
      unpack2(fr, fa, fb);
      fa = r2a < rc2 ? fa : 0.f; fb = r2b < rc2 ? fb : 0.f;
      fr = pack2(fa, fb);
      fx = fma2(fr, dx, fx); fy = fma2(fr, dy, fy); fz = fma2(fr, dz, fz);
    }
  }
  float a, b, s = 0.f;
  unpack2(fx, a, b); s += a + b; unpack2(fy, a, b); s += a + b; unpack2(fz, a, b); s += a + b;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// placement of 2-CTA clusters: every CTA records its SM, cluster rank and the time it started
__global__ void __launch_bounds__(512, 2) k_where(int *smid, long long *t0, int spin) {
  extern __shared__ unsigned char dummy[];
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) {
    unsigned s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    smid[blockIdx.x] = (int)s | ((int)cl.block_rank() << 16);
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    t0[blockIdx.x] = t;
  }
  dummy[threadIdx.x] = (unsigned char)threadIdx.x;
  cl.sync();
  long long start = clock64();
  while (clock64() - start < spin) { }
  cl.sync();
}

template <typename K>
static double time_kernel(K launch, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, clk_khz);
  float *out;
  CK(cudaMalloc(&out, sizeof(float) * sms * 4 * 1024));
  const int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    const double ms1 = time_kernel([&] { k_ffma<<<sms, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
    const double ms2 = time_kernel([&] { k_ffma2<<<sms, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
    const double fma1 = (double)iters * 8 * threads, fma2n = fma1 * 2;
    printf("threads %4d: FFMA  %.3f ms -> %.1f FMA/clk/SM | FFMA2 %.3f ms -> %.1f FMA/clk/SM (at %d kHz nominal)\n", threads,
           ms1, fma1 / (ms1 * 1e-3 * clk_khz * 1e3), ms2, fma2n / (ms2 * 1e-3 * clk_khz * 1e3), clk_khz);
  }
  {
    const int it = 400;
    const double ms1 = time_kernel([&] { k_pair_scalar<<<sms, 1024>>>(out, it, -0.05f, 700.f); }, 5);
    const double ms2 = time_kernel([&] { k_pair_packed<<<sms, 1024>>>(out, it, -0.05f, 700.f); }, 5);
    const double pairs = (double)it * 64 * 1024;
    printf("pair body: scalar %.3f ms -> %.2f pairs/clk/SM | packed %.3f ms -> %.2f pairs/clk/SM\n", ms1,
           pairs / (ms1 * 1e-3 * clk_khz * 1e3), ms2, pairs / (ms2 * 1e-3 * clk_khz * 1e3));
  }
  {
    const int it = 400;
    const double ms3 = time_kernel([&] { k_pair_scalar_poly<<<sms, 1024>>>(out, it, 0.0362f, 700.f); }, 5);
    const double ms4 = time_kernel([&] { k_pair_packed_poly<<<sms, 1024>>>(out, it, 0.0362f, 700.f); }, 5);
    const double ms5 = time_kernel([&] { k_pair_packed_poly<<<sms, 512>>>(out, it, 0.0362f, 700.f); }, 5);
    const double pairs = (double)it * 64 * 1024;
    printf("pair body, exp as a degree-6 polynomial: scalar %.3f ms -> %.2f pairs/clk/SM | packed %.3f ms -> %.2f pairs/clk/SM | packed, 512 threads %.3f ms -> %.2f pairs/clk/SM\n",
           ms3, pairs / (ms3 * 1e-3 * clk_khz * 1e3), ms4, pairs / (ms4 * 1e-3 * clk_khz * 1e3), ms5, 0.5 * pairs / (ms5 * 1e-3 * clk_khz * 1e3));
    for (int which = 0; which < 2; ++which) {
      const int itm = 4000;
      const double ms = time_kernel([&] { k_mufu<<<sms, 1024>>>(out, itm, which); }, 5);
      printf("MUFU.%s: %.3f ms -> %.2f op/clk/SM\n", which == 0 ? "RSQ" : "EX2", ms, (double)itm * 8 * 1024 / (ms * 1e-3 * clk_khz * 1e3));
    }
  }
  // cluster placement
  for (size_t smem : {(size_t)60 * 1024, (size_t)100 * 1024}) {
    CK(cudaFuncSetAttribute(k_where, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_cta = 256;
    int *d_smid; long long *d_t0;
    CK(cudaMalloc(&d_smid, sizeof(int) * n_cta)); CK(cudaMalloc(&d_t0, sizeof(long long) * n_cta));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_cta); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int max_clusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, k_where, &cfg));
    CK(cudaLaunchKernelEx(&cfg, k_where, d_smid, d_t0, 2000000));
    CK(cudaDeviceSynchronize());
    std::vector<int> smid(n_cta); std::vector<long long> t0(n_cta);
    CK(cudaMemcpy(smid.data(), d_smid, sizeof(int) * n_cta, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(t0.data(), d_t0, sizeof(long long) * n_cta, cudaMemcpyDeviceToHost));
    long long tmin = t0[0];
    for (auto t : t0) tmin = t < tmin ? t : tmin;
    std::vector<int> per_sm(sms + 16, 0);
    int late = 0;
    for (int b = 0; b < n_cta; ++b) { per_sm[smid[b] & 0xFFFF]++; if (t0[b] - tmin > 500000) ++late; }
    int hist[4] = {0, 0, 0, 0};
    for (int s = 0; s < sms; ++s) hist[per_sm[s] > 3 ? 3 : per_sm[s]]++;
    printf("clusters of 2 x 512 threads, %zu KB smem: max active clusters %d; 256 CTAs: SMs with 0/1/2/3+ CTAs = %d/%d/%d/%d; CTAs starting late: %d\n",
           smem / 1024, max_clusters, hist[0], hist[1], hist[2], hist[3], late);
    printf("  bid->smid (first 40):");
    for (int b = 0; b < 40; ++b) printf(" %d", smid[b] & 0xFFFF);
    printf("\n  bid->smid (128..167):");
    for (int b = 128; b < 168; ++b) printf(" %d", smid[b] & 0xFFFF);
    printf("\n");
    cudaFree(d_smid); cudaFree(d_t0);
  }
  return 0;
}
