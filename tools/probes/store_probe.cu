// Store-throughput probe: how fast can all SMs write an output tile the way the GEMM epilogue does
// (warp instruction = 4 rows x 128 contiguous bytes, 16 B per lane), by cache operator, footprint and row pitch?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_probe store_probe.cu && ./store_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ void st16(float* p, float4 v) {
  if (MODE == 0) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 1) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 2) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 3) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 4) asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// rows of `row_floats` floats; each CTA owns `rows_per_cta` consecutive rows and writes `cols` (multiple of 32) floats of each
template <int MODE>
__global__ void __launch_bounds__(256) store_kernel(float* out, long long pitch, int rows_per_cta, int cols, long long wrap_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rsub = lane >> 3, cq = lane & 7;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  long long row0 = ((long long)blockIdx.x * rows_per_cta) % wrap_rows;
  for (int c = 0; c < cols; c += 32)
    for (int r = warp * 4 + rsub; r < rows_per_cta; r += 32) st16<MODE>(out + (row0 + r) * pitch + c + cq * 4, v);
}

template <int MODE>
float run(float* buf, long long pitch, int ctas, int rows_per_cta, int cols, long long wrap_rows, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  store_kernel<MODE><<<ctas, 256>>>(buf, pitch, rows_per_cta, cols, wrap_rows);
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) store_kernel<MODE><<<ctas, 256>>>(buf, pitch, rows_per_cta, cols, wrap_rows);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const long long bytes = 2048ll << 20;
  float* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  const char* names[5] = {"st", "st.cg", "st.cs", "st.wt", "red.add"};
  // footprints: 19 MB (one 28x28 layer output, L2-resident), 66 MB, 600 MB (streaming)
  struct Case { const char* what; int ctas, rows, cols; long long pitch, wrap_rows; } cases[] = {
      {"147 CTAs x 128 rows x 256 cols, pitch 256 (19 MB, one wave)", 147, 128, 256, 256, 1ll << 40},
      {"1008 CTAs x 128 rows x 160 cols, pitch 320 (83 MB)", 1008, 128, 160, 320, 1ll << 40},
      {"4736 CTAs x 128 rows x 256 cols, pitch 256 (620 MB, streaming)", 4736, 128, 256, 256, 1ll << 40},
      {"4736 CTAs x 128 rows x 256 cols wrapped onto 19 MB (L2-resident)", 4736, 128, 256, 256, 147 * 128},
  };
  for (auto& c : cases) {
    printf("%s\n", c.what);
    const double mb = (double)c.ctas * c.rows * c.cols * 4 / 1e6;
    for (int m = 0; m < 5; ++m) {
      float ms = 0;
      if (m == 0) ms = run<0>(buf, c.pitch, c.ctas, c.rows, c.cols, c.wrap_rows, 20);
      if (m == 1) ms = run<1>(buf, c.pitch, c.ctas, c.rows, c.cols, c.wrap_rows, 20);
      if (m == 2) ms = run<2>(buf, c.pitch, c.ctas, c.rows, c.cols, c.wrap_rows, 20);
      if (m == 3) ms = run<3>(buf, c.pitch, c.ctas, c.rows, c.cols, c.wrap_rows, 20);
      if (m == 4) ms = run<4>(buf, c.pitch, c.ctas, c.rows, c.cols, c.wrap_rows, 20);
      printf("  %-8s %8.2f us  %7.2f TB/s  (%.1f B/clk/SM at 1.9 GHz x 148)\n", names[m], ms * 1e3, mb / ms / 1e3,
             mb * 1e6 / (ms * 1e-3) / 1.9e9 / 148);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
