// texprobe.cu -- which texel does the B200 texture unit pick for normalised coordinates at / near
// texel boundaries (point filter, clamp)?  Output: gpurun_out/texprobe.bin = records (float u, int W, int idx).
#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>

__global__ void probe(cudaTextureObject_t tex, const float *u, int *idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int)tex2D<float>(tex, u[i], 0.25f);
}

int main() {
  const int Ws[] = {1200, 800, 300, 1024, 37};
  FILE *f = fopen("gpurun_out/texprobe.bin", "wb");
  for (int W : Ws) {
    const int H = 2;
    std::vector<float> h((size_t)W * H);
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) h[(size_t)y * W + x] = (float)x;
    cudaArray_t arr;
    cudaChannelFormatDesc d = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    cudaMallocArray(&arr, &d, W, H);
    cudaMemcpy2DToArray(arr, 0, 0, h.data(), W * 4, W * 4, H, cudaMemcpyHostToDevice);
    cudaResourceDesc res{}; res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
    cudaTextureObject_t tex;
    cudaCreateTextureObject(&tex, &res, &td, nullptr);
    std::vector<float> us;
    for (int k = 0; k <= W; k++) {
      float c = (float)((double)k / W);
      float v = c;
      for (int s = 0; s < 6; s++) v = nextafterf(v, -1.0f);
      for (int s = 0; s < 13; s++) { us.push_back(v); v = nextafterf(v, 2.0f); }
    }
    for (int k = 0; k < 4000; k++) us.push_back((float)((k * 2654435761u) % 1000003u) / 1000003.0f);
    float *du; int *di;
    cudaMalloc(&du, us.size() * 4); cudaMalloc(&di, us.size() * 4);
    cudaMemcpy(du, us.data(), us.size() * 4, cudaMemcpyHostToDevice);
    probe<<<(unsigned)((us.size() + 255) / 256), 256>>>(tex, du, di, (int)us.size());
    std::vector<int> idx(us.size());
    cudaMemcpy(idx.data(), di, us.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < us.size(); i++) { fwrite(&us[i], 4, 1, f); fwrite(&W, 4, 1, f); fwrite(&idx[i], 4, 1, f); }
    printf("W=%d: %zu probes, err=%s\n", W, us.size(), cudaGetErrorString(cudaGetLastError()));
  }
  fclose(f);
  return 0;
}
