// Micro-probe for the persistent sumcheck kernel (round 2): the fixed costs of ONE Fiat-Shamir round when the round loop
// lives on the device.  Measured with %globaltimer inside the kernels (ns), averaged over many iterations:
//   blake2b    one Blake2b-256 compression on ONE thread (12 rounds fully unrolled), dependent chain of compressions
//   montmul    dependent chain of Montgomery products on one thread (per-product latency)
//   gridsync   arrive (atomic) -> control block sees all -> flag -> workers see flag, cooperative grid of N blocks
//   hostrtt    device publishes a 48-byte tagged vector to mapped host memory, host thread answers through a mapped
//              mailbox word, device polls it: one full PCIe round trip as the persistent kernel would pay per round
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I jolt_atlas_b200/csrc -o /tmp/persist_probe scripts/micro/persist_probe.cu
#include <cstdio>
#include <cstdint>
#include <thread>
#include <atomic>
#include <cuda_runtime.h>
#include "fp.cuh"
using namespace ja;

__device__ __forceinline__ uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
#define G(a,b,c,d,x,y) v[a]+=v[b]+(x); v[d]=ror64(v[d]^v[a],32); v[c]+=v[d]; v[b]=ror64(v[b]^v[c],24); v[a]+=v[b]+(y); v[d]=ror64(v[d]^v[a],16); v[c]+=v[d]; v[b]=ror64(v[b]^v[c],63);
__device__ constexpr uint8_t S[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void k_blake(uint64_t* io, int iters, unsigned long long* ns) {
  const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                          0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  uint64_t st[4]; for (int i = 0; i < 4; i++) st[i] = io[i];
  const uint64_t p0 = io[8], p1 = io[9], p2 = io[10], p3 = io[11];
  const unsigned long long t0 = gtime();
  for (int it = 0; it < iters; it++) {
    uint64_t m[16] = {st[0], st[1], st[2], st[3], 0, 0, 0, (uint64_t)it << 32, p0, p1, p2, p3, 0, 0, 0, 0};
    uint64_t v[16];
    for (int i = 0; i < 8; i++) { v[i] = IV[i]; v[8 + i] = IV[i]; }
    v[0] ^= 0x01010020ull; v[12] ^= 96; v[14] = ~v[14];
#pragma unroll
    for (int r = 0; r < 12; r++) {
      G(0,4,8,12,m[S[r][0]],m[S[r][1]]) G(1,5,9,13,m[S[r][2]],m[S[r][3]]) G(2,6,10,14,m[S[r][4]],m[S[r][5]]) G(3,7,11,15,m[S[r][6]],m[S[r][7]])
      G(0,5,10,15,m[S[r][8]],m[S[r][9]]) G(1,6,11,12,m[S[r][10]],m[S[r][11]]) G(2,7,8,13,m[S[r][12]],m[S[r][13]]) G(3,4,9,14,m[S[r][14]],m[S[r][15]])
    }
    for (int i = 0; i < 4; i++) st[i] = IV[i] ^ v[i] ^ v[8 + i] ^ (i == 0 ? 0x01010020ull : 0);
  }
  const unsigned long long t1 = gtime();
  for (int i = 0; i < 4; i++) io[i] = st[i];
  ns[0] = t1 - t0;
}

__global__ void k_montmul(Fr* io, int iters, unsigned long long* ns) {
  Fr a = io[0], b = io[1];
  const unsigned long long t0 = gtime();
  for (int it = 0; it < iters; it++) a = fp_mul<FrParams>(a, b);
  const unsigned long long t1 = gtime();
  io[0] = a; ns[0] = t1 - t0;
  Challenge c; c.c[0] = b.l[0]; c.c[1] = b.l[1]; c.c[2] = b.l[2]; c.c[3] = b.l[3] & 0x1fffffff;
  const unsigned long long t2 = gtime();
  for (int it = 0; it < iters; it++) a = fp_mul_challenge<FrParams>(a, c);
  const unsigned long long t3 = gtime();
  io[2] = a; ns[1] = t3 - t2;
}

// control = block 0; workers = blocks 1..N-1: per round every worker arrives on a counter, control waits for all of them,
// publishes the round flag, workers wait for it
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) { unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned int* p, unsigned int v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned int* p, unsigned int v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__global__ void k_gridsync(unsigned int* counters, unsigned int* flag, int rounds, unsigned long long* ns) {
  const unsigned int nw = gridDim.x - 1;
  unsigned long long t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) t0 = gtime();
  for (int r = 0; r < rounds; r++) {
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        while (ld_acquire(counters + r) < nw) {}
        st_release(flag, (unsigned int)r + 1);
      }
    } else {
      __syncthreads();
      if (threadIdx.x == 0) {
        red_release(counters + r, 1u);
        while (ld_acquire(flag) < (unsigned int)r + 1) {}
      }
      __syncthreads();
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) ns[0] = gtime() - t0;
}

// host round trip: publish tagged vector -> host answers in mailbox -> device sees it
__global__ void k_hostrtt(uint4* host_out, const volatile uint4* mail, int rounds, unsigned long long* ns) {
  if (threadIdx.x != 0) return;
  const unsigned long long t0 = gtime();
  for (int r = 1; r <= rounds; r++) {
    host_out[0] = make_uint4(r, r, r, r); host_out[1] = make_uint4(r, r, r, r); host_out[2] = make_uint4(r, r, r, r);
    uint4 v;
    do { v.x = mail->x; v.y = mail->y; v.z = mail->z; v.w = mail->w; } while (v.w != (unsigned int)r);
  }
  ns[0] = gtime() - t0;
}

int main() {
  unsigned long long* ns; cudaMallocManaged(&ns, 64);
  uint64_t* io; cudaMallocManaged(&io, 256);
  for (int i = 0; i < 16; i++) io[i] = 0x0123456789abcdefull * (i + 3);
  const int IT = 2000;
  k_blake<<<1, 1>>>(io, 10, ns); cudaDeviceSynchronize();
  k_blake<<<1, 1>>>(io, IT, ns); cudaDeviceSynchronize();
  printf("blake2b compress, one thread:      %.1f ns each (chain of %d)\n", (double)ns[0] / IT, IT);
  k_blake<<<1, 32>>>(io, IT, ns); cudaDeviceSynchronize();
  printf("blake2b compress, full warp same:  %.1f ns each\n", (double)ns[0] / IT);
  Fr* f; cudaMallocManaged(&f, 4 * sizeof(Fr));
  for (int i = 0; i < 8; i++) { f[0].l[i] = 0x1234567u * (i + 1); f[1].l[i] = 0x7654321u * (i + 2); }
  f[0].l[7] &= 0x0fffffff; f[1].l[7] &= 0x0fffffff;
  k_montmul<<<1, 1>>>(f, 10, ns); cudaDeviceSynchronize();
  k_montmul<<<1, 1>>>(f, IT, ns); cudaDeviceSynchronize();
  printf("montgomery product, dependent:     %.1f ns full, %.1f ns challenge (one thread)\n", (double)ns[0] / IT, (double)ns[1] / IT);
  k_montmul<<<1, 32>>>(f, IT, ns); cudaDeviceSynchronize();
  printf("montgomery product, dependent:     %.1f ns full, %.1f ns challenge (full warp)\n", (double)ns[0] / IT, (double)ns[1] / IT);
  unsigned int *ctr, *flag; cudaMalloc(&ctr, 4 * 4096); cudaMalloc(&flag, 4);
  for (int nb : {2, 9, 33, 74, 148}) {
    cudaMemset(ctr, 0, 4 * 4096); cudaMemset(flag, 0, 4);
    int rounds = 1000;
    void* args[] = {&ctr, &flag, &rounds, &ns};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_gridsync, dim3(nb), dim3(256), args, 0, 0);
    cudaDeviceSynchronize();
    printf("grid sync control+%3d workers:     %.1f ns per round (%s)\n", nb - 1, (double)ns[0] / rounds, cudaGetErrorString(e));
  }
  uint4* hv; cudaHostAlloc(&hv, 4096, cudaHostAllocMapped); uint4* dv; cudaHostGetDevicePointer(&dv, hv, 0);
  volatile uint4* hmail = hv + 16; uint4* dmail = dv + 16;
  memset((void*)hv, 0, 4096);
  const int R = 20000;
  std::atomic<bool> stop{false};
  std::thread host([&] {
    volatile unsigned int* q = reinterpret_cast<volatile unsigned int*>(hv);
    volatile unsigned int* m = reinterpret_cast<volatile unsigned int*>(hmail);
    for (int r = 1; r <= R && !stop; r++) {
      while (!(q[3] == (unsigned)r && q[7] == (unsigned)r && q[11] == (unsigned)r)) { if (stop) return; }
      m[0] = r; m[1] = r; m[2] = r; __atomic_thread_fence(__ATOMIC_RELEASE); m[3] = r;
    }
  });
  k_hostrtt<<<1, 32>>>(dv, dmail, R, ns); cudaDeviceSynchronize();
  stop = true; host.join();
  printf("host round trip (publish 48 B -> host answers -> device polls mailbox): %.1f ns per round\n", (double)ns[0] / R);
  return 0;
}
