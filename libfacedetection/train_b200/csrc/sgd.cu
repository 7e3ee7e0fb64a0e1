// Fused SGD step over the flat parameter bucket: DDP gradient mean (grad_scale = 1/world),
// weight decay, momentum and the update in one pass (torch.optim.SGD semantics as configured in
// configs/yunet_n.py:1: lr 0.01, momentum 0.9, weight_decay 5e-4, no dampening/nesterov; the first
// step's momentum buffer equals the gradient, which a zero-initialised buffer reproduces).
#include "kernels.h"

namespace yunet {

namespace {
__global__ void sgd_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ v,
                           long long n, float lr, float momentum, float wd, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float wi = w[i];
  const float gi = fmaf(wd, wi, g[i] * gscale);
  const float vi = fmaf(momentum, v[i], gi);
  v[i] = vi;
  w[i] = wi - lr * vi;
}
// same update with the learning rate read from device memory: the launch parameters stay constant, so
// the whole training step can be replayed as a CUDA graph while the schedule changes lr per iteration
__global__ void sgd_kernel_dev(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ v,
                               long long n, const float* __restrict__ lr_dev, float momentum, float wd,
                               float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr = __ldg(lr_dev);
  const float wi = w[i];
  const float gi = fmaf(wd, wi, g[i] * gscale);
  const float vi = fmaf(momentum, v[i], gi);
  v[i] = vi;
  w[i] = wi - lr * vi;
}
}  // namespace

cudaError_t launch_sgd_dev(float* params, const float* grad, float* mom, long long n, const float* lr_dev,
                           float momentum, float wd, float grad_scale, cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256);
  sgd_kernel_dev<<<blocks, 256, 0, s>>>(params, grad, mom, n, lr_dev, momentum, wd, grad_scale);
  return cudaGetLastError();
}

cudaError_t launch_sgd(float* params, const float* grad, float* mom, long long n, float lr,
                       float momentum, float wd, float grad_scale, cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256);
  sgd_kernel<<<blocks, 256, 0, s>>>(params, grad, mom, n, lr, momentum, wd, grad_scale);
  return cudaGetLastError();
}

}  // namespace yunet
