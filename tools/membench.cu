// membench.cu — what HBM bandwidth do multi-stream fp64 SoA kernels reach on this B200?
// (ceiling for the node/element passes; compare with the 2-stream copy in MEASURED_PEAKS.json)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

template <int R, int W>
__global__ void streams(const double *__restrict__ in, double *__restrict__ out, long long n, long long pitch) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[R];
#pragma unroll
  for (int r = 0; r < R; r++) v[r] = in[r * pitch + i];
  double s = 0;
#pragma unroll
  for (int r = 0; r < R; r++) s += v[r];
#pragma unroll
  for (int w = 0; w < W; w++) out[w * pitch + i] = s + w;
}

// structured hexa-like gather: 8 indices per element (SoA), 3 planes gathered, 1 output
__global__ void gather8(const int *__restrict__ idx, const double *__restrict__ tab, double *__restrict__ out,
                        long long ne, long long ep, long long np) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int id[8];
#pragma unroll
  for (int k = 0; k < 8; k++) id[k] = idx[k * ep + e];
  double s = 0;
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int k = 0; k < 8; k++) s += tab[c * np + id[k]];
  out[e] = s;
}

template <class F>
float timeit(F f, int reps = 20) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; i++) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

template <int R, int W>
void run_streams(double *in, double *out, long long n, long long pitch, int tpb) {
  float ms = timeit([&] { streams<R, W><<<(unsigned)((n + tpb - 1) / tpb), tpb>>>(in, out, n, pitch); });
  printf("streams R=%2d W=%2d tpb=%3d : %7.3f ms  %7.1f GB/s\n", R, W, tpb, ms, (R + W) * 8.0 * n / ms / 1e6);
}

int main() {
  const long long n = 10077696, pitch = (n + 31) / 32 * 32; // nodes of the 215^3 hexa cube
  double *in, *out;
  cudaMalloc(&in, 32 * pitch * 8); cudaMalloc(&out, 32 * pitch * 8);
  cudaMemset(in, 0, 32 * pitch * 8);
  for (int tpb : {128, 256, 512}) {
    run_streams<1, 1>(in, out, n, pitch, tpb);
    run_streams<2, 1>(in, out, n, pitch, tpb);
    run_streams<8, 8>(in, out, n, pitch, tpb);
    run_streams<15, 15>(in, out, n, pitch, tpb);
    run_streams<30, 1>(in, out, n, pitch, tpb);
    run_streams<24, 6>(in, out, n, pitch, tpb);
    run_streams<6, 24>(in, out, n, pitch, tpb);
    run_streams<1, 24>(in, out, n, pitch, tpb);
  }
  // gather like E1
  const int nx = 215;
  const long long ne = (long long)nx * nx * nx, ep = (ne + 31) / 32 * 32;
  std::vector<int> h(8 * ep, 0);
  const int n1 = nx + 1;
  for (long long e = 0; e < ne; e++) {
    int ez = e / (nx * nx), ey = (e / nx) % nx, ex = e % nx;
    long long nb1 = (long long)n1 * n1 * ez + n1 * ey + ex, nb2 = nb1 + n1, nz = (long long)n1 * n1;
    long long nh[8] = {nb1, nb1 + 1, nb2 + 1, nb2, nb1 + nz, nb1 + nz + 1, nb2 + nz + 1, nb2 + nz};
    for (int k = 0; k < 8; k++) h[k * ep + e] = (int)nh[k];
  }
  int *idx;
  cudaMalloc(&idx, 8 * ep * 4);
  cudaMemcpy(idx, h.data(), 8 * ep * 4, cudaMemcpyHostToDevice);
  for (int tpb : {128, 256}) {
    float ms = timeit([&] { gather8<<<(unsigned)((ne + tpb - 1) / tpb), tpb>>>(idx, in, out, ne, ep, pitch); });
    printf("gather8 (E1-like) tpb=%d: %7.3f ms  alg %7.1f GB/s (32B idx + 24B x + 8B out per element)\n", tpb, ms,
           64.0 * ne / ms / 1e6);
  }
  return 0;
}
