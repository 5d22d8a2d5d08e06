// Flat-arena optimizer step for data-parallel training: global gradient norm,
// 1/world averaging, norm clipping and AdamW in two launches over the whole model.
//
// Replaces what the reference gets from Lightning around its training_step
// (analysisgnn/train/train_analysisgnn.py:246-255: gradient_clip_val=1.0, i.e.
// torch.nn.utils.clip_grad_norm_, then torch.optim.AdamW built at
// analysisgnn/models/analysis.py:1380) -- hundreds of small foreach launches over
// ~200 parameter tensors -- with
//   sumsq:  per-block partial sums of g^2 over the gradient arena (fixed order => deterministic)
//   adamw:  every block re-reduces the partials (<= 1184 floats), derives the clip factor and
//           applies the update; parameters are addressed through a chunk table so the
//           modules keep their own tensors (state_dict and cuDNN weight layout untouched).
// Roofline: HBM, 28 B / parameter (p, g, m, v read; p, m, v written).
#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 4096;  // elements per block in the update kernel (arena offsets are chunk aligned)

__device__ __forceinline__ float block_sum(float v, float* smem) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < kThreads / 32 ? smem[threadIdx.x] : 0.f;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(kThreads) sumsq_kernel(const float* __restrict__ g, int64_t n4,
                                                          float* __restrict__ partials, int* __restrict__ step_counter) {
  // the step counter lives on the device so that a CUDA-graph replay advances it (bias correction)
  if (step_counter && blockIdx.x == 0 && threadIdx.x == 0) *step_counter += 1;
  __shared__ float smem[kThreads / 32];
  float acc = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    const float4 t = __ldg(g4 + i);
    acc = fmaf(t.x, t.x, acc);
    acc = fmaf(t.y, t.y, acc);
    acc = fmaf(t.z, t.z, acc);
    acc = fmaf(t.w, t.w, acc);
  }
  const float tot = block_sum(acc, smem);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

struct AdamParams {
  const agnn_param_chunk_t* chunks;  // device table, one entry per block
  const float* grad;                 // arena
  float* m;
  float* v;
  const float* partials;
  int n_partials;
  float lr, beta1, beta2, eps, weight_decay, grad_scale, max_norm, bias1, bias2_sqrt;
  const int* step_dev;               // optional: step number on the device (overrides bias1 / bias2_sqrt)
  const float* lr_dev;               // optional: learning rate on the device (overrides lr) -- a scheduler keeps
                                     // working when the launch is replayed from a CUDA graph
  float* norm_out;
};

__global__ void __launch_bounds__(kThreads) adamw_kernel(const __grid_constant__ AdamParams p) {
  __shared__ float smem[kThreads / 32];
  __shared__ float s_clip;
  float acc = 0.f;
  for (int i = threadIdx.x; i < p.n_partials; i += kThreads) acc += p.partials[i];
  const float tot = block_sum(acc, smem);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(tot) * p.grad_scale;  // norm of the averaged gradient
    float clip = 1.f;
    if (p.max_norm > 0.f) clip = fminf(1.f, p.max_norm / (norm + 1e-6f));  // torch clip_grad_norm_
    s_clip = clip * p.grad_scale;
    if (blockIdx.x == 0 && p.norm_out) *p.norm_out = norm;
  }
  __syncthreads();
  const float gs = s_clip;
  const agnn_param_chunk_t c = p.chunks[blockIdx.x];
  float* const w = static_cast<float*>(c.param);
  const float lr = p.lr_dev ? __ldg(p.lr_dev) : p.lr;
  const float decay = 1.f - lr * p.weight_decay;
  float bias1 = p.bias1, bias2_sqrt = p.bias2_sqrt;
  if (p.step_dev) {
    const float t = (float)__ldg(p.step_dev);
    bias1 = 1.f - powf(p.beta1, t);
    bias2_sqrt = sqrtf(1.f - powf(p.beta2, t));
  }
  const float step_size = lr / bias1;
  for (int i = threadIdx.x * 4; i < c.count; i += kThreads * 4) {
    const int64_t a = c.arena_off + i;
    float gv[4], mv[4], vv[4], wv[4];
    const int rem = c.count - i;
    if (rem >= 4) {
      const float4 g4 = *reinterpret_cast<const float4*>(p.grad + a);
      const float4 m4 = *reinterpret_cast<const float4*>(p.m + a);
      const float4 v4 = *reinterpret_cast<const float4*>(p.v + a);
      gv[0] = g4.x; gv[1] = g4.y; gv[2] = g4.z; gv[3] = g4.w;
      mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
      vv[0] = v4.x; vv[1] = v4.y; vv[2] = v4.z; vv[3] = v4.w;
    } else {
      for (int k = 0; k < 4; ++k) {
        gv[k] = k < rem ? p.grad[a + k] : 0.f;
        mv[k] = k < rem ? p.m[a + k] : 0.f;
        vv[k] = k < rem ? p.v[a + k] : 0.f;
      }
    }
    const bool wvec = rem >= 4 && c.param_aligned;
    if (wvec) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + c.param_off + i);
      wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
    } else {
      for (int k = 0; k < 4; ++k) wv[k] = k < rem ? w[c.param_off + i + k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = gv[k] * gs;
      wv[k] *= decay;                                       // param.mul_(1 - lr * wd)
      mv[k] = mv[k] + (g - mv[k]) * (1.f - p.beta1);        // exp_avg.lerp_(grad, 1 - beta1)
      vv[k] = vv[k] * p.beta2 + (1.f - p.beta2) * g * g;    // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
      const float denom = sqrtf(vv[k]) / bias2_sqrt + p.eps;
      wv[k] -= step_size * (mv[k] / denom);                 // param.addcdiv_(exp_avg, denom, -step_size)
    }
    if (rem >= 4) {
      *reinterpret_cast<float4*>(p.m + a) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(p.v + a) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
      for (int k = 0; k < rem; ++k) { p.m[a + k] = mv[k]; p.v[a + k] = vv[k]; }
    }
    if (wvec) {
      *reinterpret_cast<float4*>(w + c.param_off + i) = make_float4(wv[0], wv[1], wv[2], wv[3]);
    } else {
      for (int k = 0; k < rem && k < 4; ++k) w[c.param_off + i + k] = wv[k];
    }
  }
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_optim_chunk_elems(void) { return kChunk; }

extern "C" int agnn_sumsq_blocks(int64_t n) {
  int64_t b = ceil_div(ceil_div(n, 4), kThreads * 4);
  if (b > kNumSM * 8) b = kNumSM * 8;
  return b < 1 ? 1 : (int)b;
}

extern "C" int agnn_sumsq_partials(const float* grad, int64_t n, float* partials, int* step_counter,
                                   agnn_stream_t stream) {
  if (n < 0 || (n % 4) || !partials || (n > 0 && !grad) || !aligned16(grad))
    return fail(AGNN_ERR_ARG, "sumsq_partials: the arena must be 16-byte aligned and a multiple of 4 elements");
  sumsq_kernel<<<agnn_sumsq_blocks(n), kThreads, 0, (cudaStream_t)stream>>>(grad, n / 4, partials, step_counter);
  return check_launch("sumsq_partials");
}

extern "C" int agnn_adamw_clip_step(const agnn_param_chunk_t* chunks, int n_chunks, const float* grad, float* m,
                                    float* v, float lr, float beta1, float beta2, float eps, float weight_decay,
                                    int step, const int* step_dev, const float* lr_dev, float grad_scale,
                                    float max_norm, const float* partials, int n_partials, float* norm_out,
                                    agnn_stream_t stream) {
  if (n_chunks < 0 || (step < 1 && !step_dev) || !chunks || !grad || !m || !v || !partials || n_partials < 1)
    return fail(AGNN_ERR_ARG, "adamw_clip_step: bad arguments (n_chunks=%d step=%d)", n_chunks, step);
  if (!aligned16(grad) || !aligned16(m) || !aligned16(v))
    return fail(AGNN_ERR_ARG, "adamw_clip_step: arenas must be 16-byte aligned");
  if (n_chunks == 0) return AGNN_OK;
  AdamParams p;
  p.chunks = chunks; p.grad = grad; p.m = m; p.v = v; p.partials = partials; p.n_partials = n_partials;
  p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.weight_decay = weight_decay;
  p.grad_scale = grad_scale; p.max_norm = max_norm;
  p.bias1 = 1.f - powf(beta1, (float)(step < 1 ? 1 : step));
  p.bias2_sqrt = sqrtf(1.f - powf(beta2, (float)(step < 1 ? 1 : step)));
  p.step_dev = step_dev;
  p.lr_dev = lr_dev;
  p.norm_out = norm_out;
  adamw_kernel<<<n_chunks, kThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("adamw_clip_step");
}
