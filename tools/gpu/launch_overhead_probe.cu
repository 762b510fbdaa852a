// Micro-probe: what a launch costs around an (almost) empty 148 x 512 kernel as a function of the parameter block size,
// and what publishing a result through mapped pinned host memory adds.  CUDA-event time per launch and host-to-host time
// (launch -> spin on the flag).   nvcc -arch=sm_100a -O3 -o launch_overhead_probe launch_overhead_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
template <int WORDS> struct P { unsigned int* flag; unsigned int seq; int publish; float pad[WORDS]; };
template <int WORDS>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ P<WORDS> p) {
  __shared__ float s[64];
  if (threadIdx.x < 64) s[threadIdx.x] = p.pad[threadIdx.x % WORDS];
  __syncthreads();
  if (p.publish && blockIdx.x == 0 && threadIdx.x == 0) {
    if (s[3] == 123456.f) p.flag[1] = 1;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(p.flag) = p.seq;
  }
}
template <int WORDS>
void run(const char* name, unsigned int* hflag, cudaStream_t st, int publish) {
  static P<WORDS> p;
  p.flag = hflag; p.publish = publish;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ev, hh;
  unsigned int seq = 0;
  for (int i = 0; i < 1200; ++i) {
    p.seq = ++seq;
    auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0, st);
    k<WORDS><<<148, 512, 16384, st>>>(p);
    cudaEventRecord(e1, st);
    if (publish) { while (*reinterpret_cast<volatile unsigned int*>(hflag) != seq) {} }
    else cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (i >= 200) { ev.push_back(ms * 1e3f); hh.push_back(std::chrono::duration<float, std::micro>(t1 - t0).count()); }
  }
  std::sort(ev.begin(), ev.end()); std::sort(hh.begin(), hh.end());
  printf("{\"what\": \"%s\", \"param_bytes\": %zu, \"publish_to_host\": %d, \"events_us_p50\": %.2f, \"host_to_host_us_p50\": %.2f}\n",
         name, sizeof(P<WORDS>), publish, ev[ev.size() / 2], hh[hh.size() / 2]);
}
int main() {
  unsigned int* hflag;
  cudaHostAlloc(&hflag, 64, cudaHostAllocMapped);
  hflag[0] = 0;
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  for (int pub = 0; pub < 2; ++pub) {
    run<16>("params 80 B", hflag, st, pub);
    run<768>("params 3 KB", hflag, st, pub);
    run<1000>("params 4 KB", hflag, st, pub);
    run<2048>("params 8 KB", hflag, st, pub);
  }
  return 0;
}
