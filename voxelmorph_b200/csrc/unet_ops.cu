// MaxPool(2), nearest Upsample(2) + channel concat, and the flat-buffer Adam step:
// reference voxelmorph/torch/networks.py:83-85,130,137-138 and scripts/torch/train.py:161,220.
// fp32 NCDHW; a 2-D problem has D == 1 and pools / upsamples H and W only.
#include "common.cuh"

namespace vxm {

__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          uint8_t* __restrict__ idx, int BC, int Do, int Ho, int Wo,
                                                          int fd) {
  size_t n = (size_t)BC * Do * Ho * Wo;
  int Di = Do * fd, Hi = Ho * 2, Wi = Wo * 2;
  (void)Di;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int ow = (int)(i % Wo);
    int oh = (int)((i / Wo) % Ho);
    int od = (int)((i / ((size_t)Wo * Ho)) % Do);
    size_t bc = i / ((size_t)Wo * Ho * Do);
    const float* xb = x + bc * (size_t)(Do * fd) * Hi * Wi;
    float best = -INFINITY;
    int bi = 0;
    // window scan order d, h, w; first maximum wins (NaN propagates like ATen: x > best || isnan)
    for (int kd = 0; kd < fd; ++kd)
      for (int kh = 0; kh < 2; ++kh)
        for (int kw = 0; kw < 2; ++kw) {
          float v = __ldg(xb + ((size_t)(od * fd + kd) * Hi + (oh * 2 + kh)) * Wi + (ow * 2 + kw));
          int code = kd * 4 + kh * 2 + kw;
          if (v > best || v != v) { best = v; bi = code; }
        }
    y[i] = best;
    if (idx) idx[i] = (uint8_t)bi;
  }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ gy, const uint8_t* __restrict__ idx,
                                                          float* __restrict__ gx, int BC, int Do, int Ho, int Wo, int fd) {
  int Hi = Ho * 2, Wi = Wo * 2, Di = Do * fd;
  size_t n = (size_t)BC * Di * Hi * Wi;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int w = (int)(i % Wi);
    int h = (int)((i / Wi) % Hi);
    int d = (int)((i / ((size_t)Wi * Hi)) % Di);
    size_t bc = i / ((size_t)Wi * Hi * Di);
    int od = d / fd, oh = h >> 1, ow = w >> 1;
    int code = (d - od * fd) * 4 + (h & 1) * 2 + (w & 1);
    size_t o = ((bc * Do + od) * Ho + oh) * (size_t)Wo + ow;
    gx[i] = (__ldg(idx + o) == code) ? __ldg(gy + o) : 0.f;
  }
}

__global__ void __launch_bounds__(256) upcat_fwd_kernel(const float* __restrict__ a, const float* __restrict__ skip,
                                                        float* __restrict__ out, int B, int Ca, int Cb, int D, int H, int W,
                                                        int fd) {
  int Do = D * fd, Ho = H * 2, Wo = W * 2, C = Ca + Cb;
  size_t n = (size_t)B * C * Do * Ho * Wo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int w = (int)(i % Wo);
    int h = (int)((i / Wo) % Ho);
    int d = (int)((i / ((size_t)Wo * Ho)) % Do);
    int c = (int)((i / ((size_t)Wo * Ho * Do)) % C);
    int b = (int)(i / ((size_t)Wo * Ho * Do * C));
    float v;
    if (c < Ca) v = __ldg(a + ((((size_t)b * Ca + c) * D + d / fd) * H + (h >> 1)) * W + (w >> 1));
    else v = __ldg(skip + ((((size_t)b * Cb + (c - Ca)) * Do + d) * Ho + h) * Wo + w);
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) upcat_bwd_a_kernel(const float* __restrict__ go, float* __restrict__ ga, int B, int Ca,
                                                          int Cb, int D, int H, int W, int fd) {
  int Do = D * fd, Ho = H * 2, Wo = W * 2, C = Ca + Cb;
  size_t n = (size_t)B * Ca * D * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    int h = (int)((i / W) % H);
    int d = (int)((i / ((size_t)W * H)) % D);
    int c = (int)((i / ((size_t)W * H * D)) % Ca);
    int b = (int)(i / ((size_t)W * H * D * Ca));
    const float* gb = go + ((size_t)b * C + c) * Do * Ho * Wo;
    float acc = 0.f;
    for (int kd = 0; kd < fd; ++kd)
      for (int kh = 0; kh < 2; ++kh) {
        const float* r = gb + ((size_t)(d * fd + kd) * Ho + (h * 2 + kh)) * Wo + w * 2;
        acc += __ldg(r) + __ldg(r + 1);
      }
    ga[i] = acc;
  }
}

__global__ void __launch_bounds__(256) upcat_bwd_skip_kernel(const float* __restrict__ go, float* __restrict__ gs, int B, int Ca,
                                                             int Cb, size_t Vo) {
  size_t n = (size_t)B * Cb * Vo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    size_t v = i % Vo;
    int c = (int)((i / Vo) % Cb);
    int b = (int)(i / (Vo * Cb));
    gs[i] = __ldg(go + ((size_t)b * (Ca + Cb) + Ca + c) * Vo + v);
  }
}

// torch.optim.Adam (no amsgrad, L2 weight decay folded into the gradient):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t n, float lr_c, float inv_sqrt_bc2, float b1,
                                                   float b2, float eps, float wd, float gscale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale, pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - lr_c * (mi / denom);
  }
}

// Graph-friendly variant: the step count lives on the device (the bias corrections cannot be baked into a captured graph).
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, size_t n, const int* __restrict__ step_ptr, float lr, float b1,
                                                       float b2, float eps, float wd, float gscale) {
  const int step = *step_ptr;
  const float lr_c = (float)((double)lr / (1.0 - pow((double)b1, (double)step)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)step)));
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale, pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - lr_c * (mi / denom);
  }
}
__global__ void increment_kernel(int* c) { *c += 1; }

static int ew_grid(size_t n) {
  size_t b = (n + 1023) / 1024;
  size_t cap = (size_t)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vxm

using namespace vxm;

static int pool_check(int B, int C, int D, int H, int W, int nd) {
  VXM_REQUIRE(B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "pool/upsample: non-positive dimension");
  VXM_REQUIRE(nd == 2 || nd == 3, "pool/upsample: nd must be 2 or 3");
  VXM_REQUIRE(nd == 3 || D == 1, "pool/upsample: a 2-D problem must be passed with D == 1");
  return VXM_OK;
}

extern "C" int vxm_maxpool2_fwd(const float* x, float* y, uint8_t* idx, int B, int C, int D, int H, int W, int nd,
                                void* stream) {
  int rc = pool_check(B, C, D, H, W, nd);
  if (rc) return rc;
  VXM_REQUIRE(x && y, "maxpool_fwd: null pointer");
  int fd = nd == 3 ? 2 : 1;
  int Do = D / fd, Ho = H / 2, Wo = W / 2;
  VXM_REQUIRE(Do > 0 && Ho > 0 && Wo > 0, "maxpool_fwd: input smaller than the pooling window");
  size_t n = (size_t)B * C * Do * Ho * Wo;
  // floor semantics like nn.MaxPool: trailing odd planes are ignored -> index with the INPUT strides
  VXM_REQUIRE(D == Do * fd && H == Ho * 2 && W == Wo * 2, "maxpool_fwd: odd sizes are not supported (U-Net needs /16 shapes)");
  maxpool_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, y, idx, B * C, Do, Ho, Wo, fd);
  return check_launch("maxpool_fwd");
}

extern "C" int vxm_maxpool2_bwd(const float* grad_y, const uint8_t* idx, float* grad_x, int B, int C, int D, int H, int W,
                                int nd, void* stream) {
  int rc = pool_check(B, C, D, H, W, nd);
  if (rc) return rc;
  VXM_REQUIRE(grad_y && idx && grad_x, "maxpool_bwd: null pointer");
  int fd = nd == 3 ? 2 : 1;
  int Do = D / fd, Ho = H / 2, Wo = W / 2;
  VXM_REQUIRE(D == Do * fd && H == Ho * 2 && W == Wo * 2, "maxpool_bwd: odd sizes are not supported");
  size_t n = (size_t)B * C * D * H * W;
  maxpool_bwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(grad_y, idx, grad_x, B * C, Do, Ho, Wo, fd);
  return check_launch("maxpool_bwd");
}

extern "C" int vxm_upcat_fwd(const float* a, const float* skip, float* out, int B, int Ca, int Cb, int D, int H, int W,
                             int nd, void* stream) {
  int rc = pool_check(B, Ca, D, H, W, nd);
  if (rc) return rc;
  VXM_REQUIRE(a && out && (Cb == 0 || skip) && Cb >= 0, "upcat_fwd: bad argument");
  int fd = nd == 3 ? 2 : 1;
  size_t n = (size_t)B * (Ca + Cb) * D * fd * H * 2 * W * 2;
  upcat_fwd_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(a, skip, out, B, Ca, Cb, D, H, W, fd);
  return check_launch("upcat_fwd");
}

extern "C" int vxm_upcat_bwd(const float* grad_out, float* grad_a, float* grad_skip, int B, int Ca, int Cb, int D, int H,
                             int W, int nd, void* stream) {
  int rc = pool_check(B, Ca, D, H, W, nd);
  if (rc) return rc;
  VXM_REQUIRE(grad_out && grad_a && Cb >= 0, "upcat_bwd: bad argument");
  int fd = nd == 3 ? 2 : 1;
  size_t n = (size_t)B * Ca * D * H * W;
  upcat_bwd_a_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(grad_out, grad_a, B, Ca, Cb, D, H, W, fd);
  rc = check_launch("upcat_bwd_a");
  if (rc) return rc;
  if (grad_skip && Cb > 0) {
    size_t Vo = (size_t)D * fd * H * 2 * W * 2;
    upcat_bwd_skip_kernel<<<ew_grid((size_t)B * Cb * Vo), 256, 0, as_stream(stream)>>>(grad_out, grad_skip, B, Ca, Cb, Vo);
    rc = check_launch("upcat_bwd_skip");
  }
  return rc;
}

extern "C" int vxm_adam_step(float* p, const float* g, float* m, float* v, size_t n, int step, float lr, float beta1,
                             float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  VXM_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adam_step: bad argument");
  double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  float lr_c = (float)((double)lr / bc1);
  float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr_c, inv_sqrt_bc2, beta1, beta2, eps,
                                                         weight_decay, grad_scale);
  return check_launch("adam_step");
}

extern "C" int vxm_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, int* step_counter, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  VXM_REQUIRE(p && g && m && v && n > 0 && step_counter, "adam_step_dev: bad argument");
  increment_kernel<<<1, 1, 0, as_stream(stream)>>>(step_counter);
  int rc = check_launch("adam_increment");
  if (rc) return rc;
  adam_dev_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, step_counter, lr, beta1, beta2, eps, weight_decay, grad_scale);
  return check_launch("adam_step_dev");
}
