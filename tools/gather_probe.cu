// Prints which texel of the 2x2 footprint each component of tex2Dgather returns (documents K2's assumption:
// x = (x0, y1), y = (x1, y1), z = (x1, y0), w = (x0, y0)) and that a gather at a texel corner with wrap addressing
// selects the four texels around that corner.  nvcc -arch=sm_100a -o tools/gather_probe tools/gather_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(cudaTextureObject_t obj, int w, int h, float* out) {
  // corner between texels (1,2) and (2,3) of an 8x4 texture whose red channel holds 10 * y + x
  float4 a = tex2Dgather<float4>(obj, 2.0f / w, 3.0f / h, 0);
  out[0] = a.x * 255.0f; out[1] = a.y * 255.0f; out[2] = a.z * 255.0f; out[3] = a.w * 255.0f;
  float4 b = tex2Dgather<float4>(obj, 0.0f, 0.0f, 0);          // corner (0, 0): wraps to texels (w-1, h-1) .. (0, 0)
  out[4] = b.x * 255.0f; out[5] = b.y * 255.0f; out[6] = b.z * 255.0f; out[7] = b.w * 255.0f;
  float4 c = tex2Dgather<float4>(obj, 1.0f, 1.0f, 0);          // corner (w, h): the same four texels
  out[8] = c.x * 255.0f; out[9] = c.y * 255.0f; out[10] = c.z * 255.0f; out[11] = c.w * 255.0f;
}
int main() {
  const int w = 8, h = 4;
  uchar4 tex[w * h];
  for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) tex[y * w + x] = make_uchar4(10 * y + x, 0, 0, 255);
  cudaChannelFormatDesc f = cudaCreateChannelDesc<uchar4>();
  cudaArray_t arr; cudaMallocArray(&arr, &f, w, h, cudaArrayTextureGather);
  cudaMemcpy2DToArray(arr, 0, 0, tex, w * 4, w * 4, h, cudaMemcpyHostToDevice);
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
  cudaTextureObject_t obj; cudaCreateTextureObject(&obj, &rd, &td, nullptr);
  float* d; cudaMalloc(&d, 12 * sizeof(float)); probe<<<1, 1>>>(obj, w, h, d);
  float o[12]; cudaMemcpy(o, d, sizeof(o), cudaMemcpyDeviceToHost);
  printf("corner(2,3): x=%g y=%g z=%g w=%g   expect x=(1,3)=31 y=(2,3)=32 z=(2,2)=22 w=(1,2)=21\n", o[0], o[1], o[2], o[3]);
  printf("corner(0,0): x=%g y=%g z=%g w=%g   expect x=(7,0)=7 y=(0,0)=0 z=(0,3)=30 w=(7,3)=37\n", o[4], o[5], o[6], o[7]);
  printf("corner(8,4): x=%g y=%g z=%g w=%g   expect the same\n", o[8], o[9], o[10], o[11]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
