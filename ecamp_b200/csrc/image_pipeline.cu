// Image half of the reference loader on the GPU (pretrain_datasets.py:47-52 applied at :27-31,113):
//   RandomResizedCrop(448, scale=(0.2, 1.0), BICUBIC) -> RandomHorizontalFlip -> Grayscale(3) -> ToTensor -> Normalize
// The random parameters (crop box, flip) are drawn on the host with torch's generator in torchvision's order
// (ecamp_b200/image_pipeline.py), the pixels are processed here: the crop box of a decoded 8-bit grayscale frame is
// resampled to 448 x 448 exactly as Pillow does it for the PIL image the reference transform sees -
//   * two passes, horizontal then vertical, with an 8-bit intermediate image (ImagingResample, Resample.c);
//   * bicubic kernel with a = -0.5, support 2 x max(scale, 1) (i.e. antialiased when shrinking), window rounded as in
//     precompute_coeffs, weights normalised in double precision, then quantised to 22-bit fixed point
//     (normalize_coeffs_8bpc); accumulation in int32 from 1 << 21, result clipped to [0, 255] after >> 22 -
// so the output bytes are bit-identical with torchvision.transforms.functional.resized_crop on the PIL image (checked
// against Pillow itself by tests/test_image_pipeline_cpu.py through the host entry point below, and kernel == host on
// the GPU).  A grayscale frame converted to RGB has three equal channels, Pillow resamples them identically and
// Grayscale(3) maps (v, v, v) back to v, so one channel is processed.  Flip is applied on the write of the second pass.
// ToTensor + Normalize (+ the 3-channel expansion) is the existing ecamp_image_u8_normalize / the uint8 input of ECAMP.
#include "kernels.cuh"

#include <vector>

namespace ecamp {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

// IEEE double arithmetic without FMA contraction on either side (Pillow is compiled C: a * b + c is two roundings)
#ifdef __CUDA_ARCH__
#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#else
#define DMUL(a, b) ((a) * (b))
#define DADD(a, b) ((a) + (b))
#define DSUB(a, b) ((a) - (b))
#define DDIV(a, b) ((a) / (b))
#endif

__host__ __device__ inline double bicubic_filter(double x) {  // Pillow Resample.c: bicubic_filter, a = -0.5
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return DADD(DMUL(DMUL(DSUB(DMUL(a + 2.0, x), a + 3.0), x), x), 1.0);
  if (x < 2.0) return DMUL(DSUB(DMUL(DADD(DMUL(DSUB(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

// window and fixed-point weights of output index xx when `in_size` input samples are resampled to `out_size`
// (Pillow precompute_coeffs with in0 = 0, in1 = in_size, + normalize_coeffs_8bpc).  k must hold ksize ints.
__host__ __device__ inline void resample_coeffs(int in_size, int out_size, int xx, int ksize, int* xmin_out, int* xcnt_out,
                                                int* k) {
  const double scale = DDIV((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = DMUL(2.0, filterscale);
  const double center = DMUL(DADD((double)xx, 0.5), scale);
  const double ss = DDIV(1.0, filterscale);
  int xmin = (int)DADD(DSUB(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)DADD(DADD(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) ww = DADD(ww, bicubic_filter(DMUL(DADD(DSUB((double)(x + xmin), center), 0.5), ss)));
  for (int x = 0; x < ksize; ++x) {
    int q = 0;
    if (x < xmax) {
      double w = bicubic_filter(DMUL(DADD(DSUB((double)(x + xmin), center), 0.5), ss));
      if (ww != 0.0) w = DDIV(w, ww);
      const double s = DMUL(w, (double)(1 << kPrecisionBits));
      q = w < 0 ? (int)DADD(-0.5, s) : (int)DADD(0.5, s);
    }
    k[x] = q;
  }
  *xmin_out = xmin;
  *xcnt_out = xmax;
}
__host__ __device__ inline int resample_ksize(int in_size, int out_size) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 2.0 * filterscale;
  return (int)ceil(support) * 2 + 1;
}
__host__ __device__ inline uint8_t clip8(int v) {
  v >>= kPrecisionBits;  // arithmetic shift, as Pillow's lookup of in >> PRECISION_BITS
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct CropDesc {         // one image of the batch
  long long src_off;      // bytes into the packed crop buffer (row-major [h, w] of the crop box)
  long long tmp_off;      // bytes into the intermediate buffer (row-major [h, out])
  int h, w, flip, pad;
};

// coefficient tables: per image [2][out][kmax] ints + [2][out][2] bounds (pass 0 = horizontal, 1 = vertical)
__global__ void coeff_kernel(const CropDesc* __restrict__ desc, int out, int kmax, int* __restrict__ coef, int* __restrict__ bounds) {
  const int b = blockIdx.y, pass = blockIdx.z;
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out) return;
  const int in_size = pass == 0 ? desc[b].w : desc[b].h;
  const int ks = resample_ksize(in_size, out);
  int* k = coef + (((size_t)b * 2 + pass) * out + xx) * kmax;
  int xmin, xcnt;
  resample_coeffs(in_size, out, xx, ks < kmax ? ks : kmax, &xmin, &xcnt, k);
  for (int x = ks; x < kmax; ++x) k[x] = 0;
  int* bd = bounds + (((size_t)b * 2 + pass) * out + xx) * 2;
  bd[0] = xmin;
  bd[1] = xcnt;
}

// horizontal pass: tmp[y, xx] = clip8(2^21 + sum_x src[y, xmin + x] * k[x]); one thread per output, 8 rows per block row-loop
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ src, const CropDesc* __restrict__ desc, int out,
                                                         int kmax, const int* __restrict__ coef, const int* __restrict__ bounds,
                                                         uint8_t* __restrict__ tmp) {
  const int b = blockIdx.z;
  const CropDesc d = desc[b];
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out) return;
  const int* k = coef + (((size_t)b * 2 + 0) * out + xx) * kmax;
  const int xmin = bounds[(((size_t)b * 2 + 0) * out + xx) * 2], xcnt = bounds[(((size_t)b * 2 + 0) * out + xx) * 2 + 1];
  for (int y = blockIdx.y; y < d.h; y += gridDim.y) {
    const uint8_t* row = src + d.src_off + (size_t)y * d.w + xmin;
    int ss = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < xcnt; ++x) ss += (int)row[x] * k[x];
    tmp[d.tmp_off + (size_t)y * out + xx] = clip8(ss);
  }
}

// vertical pass + flip: dst[b, yy, xx'] = clip8(2^21 + sum_y tmp[ymin + y, xx] * k[y]), xx' = flip ? out - 1 - xx : xx
__global__ void __launch_bounds__(256) resample_v_kernel(const uint8_t* __restrict__ tmp, const CropDesc* __restrict__ desc, int out,
                                                         int kmax, const int* __restrict__ coef, const int* __restrict__ bounds,
                                                         uint8_t* __restrict__ dst) {
  const int b = blockIdx.z, yy = blockIdx.y;
  const CropDesc d = desc[b];
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out) return;
  const int* k = coef + (((size_t)b * 2 + 1) * out + yy) * kmax;
  const int ymin = bounds[(((size_t)b * 2 + 1) * out + yy) * 2], ycnt = bounds[(((size_t)b * 2 + 1) * out + yy) * 2 + 1];
  const uint8_t* col = tmp + d.tmp_off + (size_t)ymin * out + xx;
  int ss = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < ycnt; ++y) ss += (int)col[(size_t)y * out] * k[y];
  const int xo = d.flip ? out - 1 - xx : xx;
  dst[((size_t)b * out + yy) * out + xo] = clip8(ss);
}

}  // namespace

// ---- host reference of the same arithmetic (CPU test against Pillow; never on the product path) -------------------------
int image_resized_crop_host(const uint8_t* crop, int h, int w, int flip, int out, uint8_t* dst) {
  ECAMP_REQUIRE(crop && dst && h > 0 && w > 0 && out > 0, "image_resized_crop_host: bad argument");
  const int kx = resample_ksize(w, out), ky = resample_ksize(h, out);
  std::vector<uint8_t> tmp((size_t)h * out);
  std::vector<int> k((size_t)(kx > ky ? kx : ky));
  for (int xx = 0; xx < out; ++xx) {
    int xmin, xcnt;
    resample_coeffs(w, out, xx, kx, &xmin, &xcnt, k.data());
    for (int y = 0; y < h; ++y) {
      int ss = 1 << (kPrecisionBits - 1);
      for (int x = 0; x < xcnt; ++x) ss += (int)crop[(size_t)y * w + xmin + x] * k[x];
      tmp[(size_t)y * out + xx] = clip8(ss);
    }
  }
  for (int yy = 0; yy < out; ++yy) {
    int ymin, ycnt;
    resample_coeffs(h, out, yy, ky, &ymin, &ycnt, k.data());
    for (int xx = 0; xx < out; ++xx) {
      int ss = 1 << (kPrecisionBits - 1);
      for (int y = 0; y < ycnt; ++y) ss += (int)tmp[(size_t)(ymin + y) * out + xx] * k[y];
      dst[(size_t)yy * out + (flip ? out - 1 - xx : xx)] = clip8(ss);
    }
  }
  return 0;
}

// bytes of device scratch for a batch: descriptors are passed separately
size_t image_resized_crop_ws_bytes(int B, int out, int kmax, long long tmp_bytes) {
  return (size_t)B * 2 * out * kmax * sizeof(int) + (size_t)B * 2 * out * 2 * sizeof(int) + (size_t)tmp_bytes + 512;
}
int image_resample_kmax(int in_size, int out) { return resample_ksize(in_size, out); }

// crops: packed crop boxes (device); desc: B descriptors (device, layout of CropDesc = ecamp_crop_desc); dst: [B, out, out] u8
int image_resized_crop(const uint8_t* crops, const void* desc_dev, int B, int hmax, int out, int kmax, void* ws, size_t ws_bytes,
                       long long tmp_bytes, uint8_t* dst, cudaStream_t st) {
  ECAMP_REQUIRE(crops && desc_dev && ws && dst && B > 0 && out > 0 && kmax > 0 && hmax > 0, "image_resized_crop: bad argument");
  ECAMP_REQUIRE(ws_bytes >= image_resized_crop_ws_bytes(B, out, kmax, tmp_bytes), "image_resized_crop: workspace too small");
  const CropDesc* desc = static_cast<const CropDesc*>(desc_dev);
  int* coef = static_cast<int*>(ws);
  int* bounds = coef + (size_t)B * 2 * out * kmax;
  uint8_t* tmp = reinterpret_cast<uint8_t*>(bounds + (size_t)B * 2 * out * 2);
  const int tx = 128;
  coeff_kernel<<<dim3((out + tx - 1) / tx, B, 2), tx, 0, st>>>(desc, out, kmax, coef, bounds);
  ECAMP_LAUNCHED();
  const int rows = hmax < 64 ? hmax : 64;
  resample_h_kernel<<<dim3((out + 255) / 256, rows, B), 256, 0, st>>>(crops, desc, out, kmax, coef, bounds, tmp);
  ECAMP_LAUNCHED();
  resample_v_kernel<<<dim3((out + 255) / 256, out, B), 256, 0, st>>>(tmp, desc, out, kmax, coef, bounds, dst);
  ECAMP_LAUNCHED();
  return 0;
}

}  // namespace ecamp
