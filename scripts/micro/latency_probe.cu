// Micro-probe of the fixed costs inside a tiny round kernel on B200 (back-to-back launches on one stream, CUDA events):
// empty kernel, publication to host-mapped memory with and without the system fence, a dependent global-load chain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu ; run under gpurun.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_empty() {}
__global__ void k_publish(uint4* host_vals, volatile unsigned int* host_seq, unsigned int v, int fence) {
  if (threadIdx.x == 0) {
    host_vals[0] = make_uint4(v, v, v, v); host_vals[1] = make_uint4(v, v, v, v);
    if (fence) __threadfence_system();
    *host_seq = v;
  }
}
__global__ void k_chain(const unsigned int* __restrict__ idx, unsigned int* out, int depth) {
  unsigned int i = threadIdx.x;
  for (int d = 0; d < depth; d++) i = idx[i];       // dependent global loads
  out[threadIdx.x] = i;
}
__global__ void k_atomic_tail(unsigned int* counter, unsigned int* out) {
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) out[0] = 1;
}
__global__ void k_mail(const volatile unsigned int* mail, unsigned int seq, uint4* host_vals, volatile unsigned int* host_seq, int relay_blocks,
                       volatile unsigned int* dev_flag) {
  // block 0 polls the host entry and relays through device memory; the LAST block to see it publishes
  __shared__ unsigned int s;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) { while (mail[4] != seq) {} dev_flag[0] = seq; }
    else { while (dev_flag[0] != seq) {} }
    s = mail[0];
  }
  __syncthreads();
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    host_vals[0] = make_uint4(s, s, s, s); host_vals[1] = make_uint4(s, s, s, s);
    __threadfence_system();
    *host_seq = seq;
  }
}
template <class F> float timeit(F f, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 20; i++) f();
  cudaEventRecord(a);
  for (int i = 0; i < iters; i++) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms * 1e3f / iters;
}
int main() {
  uint4* hv; unsigned int* hs; cudaHostAlloc(&hv, 4096, cudaHostAllocMapped); hs = (unsigned int*)(hv + 64);
  uint4* dv; cudaHostGetDevicePointer(&dv, hv, 0);
  unsigned int *idx, *out, *ctr; cudaMalloc(&idx, 1 << 20); cudaMalloc(&out, 4096); cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
  unsigned int h[1024]; for (int i = 0; i < 1024; i++) h[i] = (i * 37 + 11) & 1023;
  cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
  const int N = 2000;
  printf("empty kernel                      %.2f us/launch\n", timeit([&] { k_empty<<<1, 256>>>(); }, N));
  printf("empty kernel, 16 blocks           %.2f us/launch\n", timeit([&] { k_empty<<<16, 256>>>(); }, N));
  unsigned int v = 1;
  printf("publish to mapped host + fence    %.2f us/launch\n", timeit([&] { k_publish<<<1, 256>>>(dv, (unsigned int*)(dv + 64), v++, 1); }, N));
  printf("publish to mapped host, no fence  %.2f us/launch\n", timeit([&] { k_publish<<<1, 256>>>(dv, (unsigned int*)(dv + 64), v++, 0); }, N));
  for (int d : {1, 2, 4, 8})
    printf("dependent global loads x%d          %.2f us/launch\n", d, timeit([&] { k_chain<<<1, 256>>>(idx, out, d); }, N));
  printf("16 blocks + fence/atomic tail     %.2f us/launch\n", timeit([&] { k_atomic_tail<<<16, 256>>>(ctr, out); }, N));
  // host round trip: launch, spin on the mapped flag
  {
    double tot = 0; const int R = 500;
    for (int i = 0; i < R + 20; i++) {
      cudaEvent_t e; (void)e;
      const unsigned int want = v++;
      timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
      k_publish<<<1, 256>>>(dv, (unsigned int*)(dv + 64), want, 1);
      while (*(volatile unsigned int*)hs != want) {}
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (i >= 20) tot += (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3;
    }
    printf("host: launch -> mapped flag seen   %.2f us\n", tot / R);
  }
  {
    unsigned int* hm; cudaHostAlloc(&hm, 64, cudaHostAllocMapped); unsigned int* dm; cudaHostGetDevicePointer(&dm, hm, 0);
    unsigned int* dflag; cudaMalloc(&dflag, 64); cudaMemset(dflag, 0, 64);
    for (int blocks : {1, 16, 128}) {
      double tot = 0; const int R = 300;
      for (int i = 0; i < R + 20; i++) {
        const unsigned int want = v++;
        k_mail<<<blocks, 256>>>(dm, want, dv, (unsigned int*)(dv + 64), 1, dflag);
        timespec w0, w1; clock_gettime(CLOCK_MONOTONIC, &w0);
        do { clock_gettime(CLOCK_MONOTONIC, &w1); } while ((w1.tv_sec - w0.tv_sec) * 1e6 + (w1.tv_nsec - w0.tv_nsec) * 1e-3 < 30.0);   // kernel is resident and polling
        timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
        hm[0] = want; __atomic_thread_fence(__ATOMIC_RELEASE); ((volatile unsigned int*)hm)[4] = want;
        while (*(volatile unsigned int*)hs != want) {}
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if (i >= 20) tot += (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3;
      }
      printf("mailbox: post -> kernel (%3d blocks) -> mapped flag seen   %.2f us\n", blocks, tot / R);
    }
  }
  return 0;
}
