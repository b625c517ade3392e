// Developer microbenchmark (not part of the product): where does the tier-1 streaming kernel's time go?
// Builds up the per-read work in stages over config-2-sized synthetic arrays and times each variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int kBlock = 256;

template <int STAGE, int UNROLL>
__global__ void __launch_bounds__(kBlock) stream_kernel(const int4* __restrict__ f1, const int4* __restrict__ f2,
                                                        const unsigned* __restrict__ lens, double* __restrict__ values,
                                                        const int4* __restrict__ slots1, const int4* __restrict__ slots2,
                                                        const double* __restrict__ pw, const double* __restrict__ ins,
                                                        const double* __restrict__ thr, int n, double rcp, double* out) {
  double sum = 0;
  const int stride = gridDim.x * blockDim.x;
  for (int r0 = blockIdx.x * blockDim.x + threadIdx.x; r0 < n; r0 += UNROLL * stride) {
    int4 a[UNROLL], b[UNROLL];
    unsigned l[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int r = r0 + u * stride;
      if (r < n) { a[u] = __ldg(f1 + r); b[u] = __ldg(f2 + r); l[u] = __ldg(lens + r); }
      else { a[u] = make_int4(0, 0, 0, 0); b[u] = a[u]; l[u] = 0; }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const int r = r0 + u * stride;
      if (r >= n) continue;
      double acc = (double)(a[u].y ^ b[u].y) * 1e-9;
      if (STAGE >= 1) {   // slot gathers
        const int4 o1 = __ldg(slots1 + a[u].x), o2 = __ldg(slots2 + b[u].x);
        acc += (double)(o1.z + o2.z) * 1e-12;
        if (STAGE >= 2) {   // pow + ins + thr tables
          const int e1 = a[u].z & 3, e2 = b[u].z & 3;
          const double p = pw[e1] * pw[100 - e1] * pw[e2] * pw[100 - e2];
          const int d = 200 + ((a[u].y + o1.z) & 255);
          acc = p * ins[d];
          if (STAGE >= 3) {   // division + log
            double v = acc * rcp;
            const double t = thr[(l[u] & 0xffff) + (l[u] >> 16)];
            if (v < t) v = t;
            acc = log(v);
          }
        }
      }
      if (STAGE >= 0) values[r] = acc;
      sum += acc;
    }
  }
  if (sum == 1.2345e-300) out[0] = sum;
}

// ---- same work, the three streamed arrays staged through shared memory by bulk async copies (TMA engine) ----------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}

template <int STAGE, int NST>
__global__ void __launch_bounds__(kBlock) bulk_kernel(const int4* __restrict__ f1, const int4* __restrict__ f2,
                                                      const unsigned* __restrict__ lens, double* __restrict__ values,
                                                      const int4* __restrict__ slots1, const int4* __restrict__ slots2,
                                                      const double* __restrict__ pw, const double* __restrict__ ins,
                                                      const double* __restrict__ thr, int n_tiles, double rcp, double* out) {
  constexpr int kTile = 2 * kBlock;
  constexpr unsigned kStageBytes = kTile * 36;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[NST];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < NST; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int t) {
    const unsigned bar = smem_u32(&bars[s]);
    const unsigned base = smem_u32(smem + (size_t)s * kStageBytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kStageBytes) : "memory");
    bulk_load(base, f1 + (size_t)t * kTile, kTile * 16, bar);
    bulk_load(base + kTile * 16, f2 + (size_t)t * kTile, kTile * 16, bar);
    bulk_load(base + kTile * 32, lens + (size_t)t * kTile, kTile * 4, bar);
  };
  if (tid == 0)
    for (int j = 0; j < NST; j++) {
      const int t = blockIdx.x + j * gridDim.x;
      if (t < n_tiles) issue(j, t);
    }
  double sum = 0;
  for (int j = 0;; j++) {
    const int t = blockIdx.x + j * gridDim.x;
    if (t >= n_tiles) break;
    const int s = j % NST;
    bar_wait(smem_u32(&bars[s]), (unsigned)((j / NST) & 1));
    const unsigned char* st = smem + (size_t)s * kStageBytes;
    int4 a[2], b[2];
    unsigned l[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      a[u] = reinterpret_cast<const int4*>(st)[tid + u * kBlock];
      b[u] = reinterpret_cast<const int4*>(st + kTile * 16)[tid + u * kBlock];
      l[u] = reinterpret_cast<const unsigned*>(st + kTile * 32)[tid + u * kBlock];
    }
    __syncthreads();   // the stage has been read by everyone: refill it
    if (tid == 0) {
      const int tn = t + NST * gridDim.x;
      if (tn < n_tiles) issue(s, tn);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = t * kTile + tid + u * kBlock;
      double acc = (double)(a[u].y ^ b[u].y) * 1e-9;
      if (STAGE >= 1) {
        const int4 o1 = __ldg(slots1 + a[u].x), o2 = __ldg(slots2 + b[u].x);
        acc += (double)(o1.z + o2.z) * 1e-12;
        if (STAGE >= 2) {
          const int e1 = a[u].z & 3, e2 = b[u].z & 3;
          const double p = pw[e1] * pw[100 - e1] * pw[e2] * pw[100 - e2];
          const int d = 200 + ((a[u].y + o1.z) & 255);
          acc = p * ins[d];
          if (STAGE >= 3) {
            double v = acc * rcp;
            const double tt = thr[(l[u] & 0xffff) + (l[u] >> 16)];
            if (v < tt) v = tt;
            acc = log(v);
          }
        }
      }
      values[r] = acc;
      sum += acc;
    }
  }
  if (sum == 1.2345e-300) out[0] = sum;
}

template <int STAGE, int NST>
float run_bulk(int grid, const int4* f1, const int4* f2, const unsigned* lens, double* values, const int4* s1, const int4* s2,
               const double* pw, const double* ins, const double* thr, int n, double* out, char* flush, size_t flush_bytes) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t smem = (size_t)NST * 512 * 36;
  CK(cudaFuncSetAttribute(bulk_kernel<STAGE, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float best = 1e9, tot = 0;
  for (int it = 0; it < 8; it++) {
    CK(cudaMemset(flush, it, flush_bytes));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    bulk_kernel<STAGE, NST><<<grid, kBlock, smem>>>(f1, f2, lens, values, s1, s2, pw, ins, thr, n / 512, 1.0 / 9192674.0, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 2) { best = fminf(best, ms); tot += ms; }
  }
  printf("BULK stage %d nst %d grid %5d: best %.1f us, mean %.1f us\n", STAGE, NST, grid, best * 1e3, tot / 6 * 1e3);
  return best;
}

template <int STAGE, int UNROLL>
float run(int grid, const int4* f1, const int4* f2, const unsigned* lens, double* values, const int4* s1, const int4* s2,
          const double* pw, const double* ins, const double* thr, int n, double* out, char* flush, size_t flush_bytes) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9, tot = 0;
  for (int it = 0; it < 8; it++) {
    CK(cudaMemset(flush, it, flush_bytes));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    stream_kernel<STAGE, UNROLL><<<grid, kBlock>>>(f1, f2, lens, values, s1, s2, pw, ins, thr, n, 1.0 / 9192674.0, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 2) { best = fminf(best, ms); tot += ms; }
  }
  printf("stage %d unroll %d grid %5d: best %.1f us, mean %.1f us\n", STAGE, UNROLL, grid, best * 1e3, tot / 6 * 1e3);
  return best;
}

int main() {
  const int n = 2000000, nkeys = 469;
  std::vector<int4> h1(n), h2(n);
  std::vector<unsigned> hl(n, 100u | (100u << 16));
  srand(1);
  for (int i = 0; i < n; i++) {
    h1[i] = make_int4(rand() % nkeys, rand() % 10000, (rand() % 3) | (1 << 16), i);
    h2[i] = make_int4(rand() % nkeys, rand() % 10000, (rand() % 3) | (1 << 16), i);
  }
  std::vector<int4> hs(nkeys);
  for (int i = 0; i < nkeys; i++) hs[i] = make_int4(7, i, i * 10000, 0);
  std::vector<double> hp(128), hi(2048), ht(512);
  for (int i = 0; i < 128; i++) hp[i] = pow(0.96, i);
  for (int i = 0; i < 2048; i++) hi[i] = exp(-0.5 * (i - 300.0) * (i - 300.0) / 900.0) / 75.0;
  for (int i = 0; i < 512; i++) ht[i] = exp(-10 - 0.7 * i);
  int4 *f1, *f2, *s1, *s2; unsigned* lens; double *values, *pw, *ins, *thr, *out; char* flush;
  const size_t fb = 256u << 20;
  CK(cudaMalloc(&f1, n * 16)); CK(cudaMalloc(&f2, n * 16)); CK(cudaMalloc(&lens, n * 4)); CK(cudaMalloc(&values, n * 8));
  CK(cudaMalloc(&s1, nkeys * 16)); CK(cudaMalloc(&s2, nkeys * 16)); CK(cudaMalloc(&pw, 128 * 8)); CK(cudaMalloc(&ins, 2048 * 8));
  CK(cudaMalloc(&thr, 512 * 8)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&flush, fb));
  CK(cudaMemcpy(f1, h1.data(), n * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(f2, h2.data(), n * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(lens, hl.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s1, hs.data(), nkeys * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(s2, hs.data(), nkeys * 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pw, hp.data(), 128 * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ins, hi.data(), 2048 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(thr, ht.data(), 512 * 8, cudaMemcpyHostToDevice));
  printf("bytes per launch: %.1f MB (2 x 16 + 4 + 8 per read)\n", n * 44.0 / 1e6);
  for (int grid : {148 * 2, 148 * 3, 148 * 4, 148 * 5, 148 * 6}) {
    run_bulk<0, 2>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run_bulk<3, 2>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run_bulk<0, 3>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run_bulk<3, 3>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run_bulk<3, 4>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
  }
  for (int grid : {148 * 4, 148 * 5, 148 * 8}) {
    run<0, 1>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run<0, 2>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run<1, 2>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run<2, 2>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run<3, 2>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
    run<3, 4>(grid, f1, f2, lens, values, s1, s2, pw, ins, thr, n, out, flush, fb);
  }
  return 0;
}
