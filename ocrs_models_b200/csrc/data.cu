// Input pipeline on the device for the recognition path: batch assembly of variable-width line images.
// Replaces, for CUDA training, the per-sample F.pad + default_collate stack of reference
// ocrs_models/train_rec.py:286-304 (collate_samples) fused with the uint8 -> [-0.5, 0.5] normalisation of
// ocrs_models/datasets/util.py:27-35 (transform_image): the host ships ONE packed buffer of raw pixels, one kernel
// writes the padded fp32 batch.
#include "common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
collate_lines_kernel(const T* __restrict__ packed, const long long* __restrict__ offsets,
                     const int* __restrict__ widths, int H, int Wpad, float* __restrict__ out) {
  const int n = blockIdx.z, y = blockIdx.y;
  const int w = widths[n];
  const T* src = packed + offsets[n] + (size_t)y * w;
  float* dst = out + ((size_t)n * H + y) * Wpad;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < Wpad; x += gridDim.x * blockDim.x) {
    float v = 0.f;  // image_pad_value = 0.0 ("grey"), train_rec.py:295
    if (x < w) {
      if (sizeof(T) == 1) v = __fsub_rn(__fdiv_rn((float)src[x], 255.0f), 0.5f);  // img.float() / 255.0 - 0.5, bit for bit
      else v = (float)src[x];
    }
    dst[x] = v;
  }
}

}  // namespace

extern "C" {

// Assemble a padded batch [N][1][H][Wpad] (fp32) from N row-major images [H][widths[n]] stored back to back in
// `packed` at element offsets `offsets[n]` (device pointers). is_u8 != 0: raw 8-bit pixels, normalised to
// [-0.5, 0.5] exactly like transform_image; else fp32 pixels copied. Columns >= widths[n] are zero.
int ocrs_collate_lines(const void* packed, int is_u8, const long long* offsets, const int* widths, int N, int H, int Wpad,
                       float* out, void* stream) {
  OCRS_CHECK_ARG(N > 0 && H > 0 && Wpad > 0, "collate_lines: bad dims");
  dim3 grid(ocrs_cdiv(Wpad, 256) < 4 ? ocrs_cdiv(Wpad, 256) : 4, H, N);
  if (is_u8)
    collate_lines_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)packed, offsets, widths, H, Wpad, out);
  else
    collate_lines_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)packed, offsets, widths, H, Wpad, out);
  OCRS_CHECK_LAUNCH("collate_lines_kernel");
  return 0;
}

}  // extern "C"
