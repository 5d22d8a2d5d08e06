// Recurrent part of the bidirectional GRU sequence branches.
//
// The reference runs nn.GRU over every score's padded note (or beat / measure) sequence: the hybrid
// branch of the encoder (analysisgnn/models/cadence.py:249-260, 276-285; analysisgnn/models/analysis.py:
// 527-537) and MetricalConvLayer.seq (analysisgnn/models/core/gnn.py:498, 523).  cuDNN's fp32 GRU (TF32
// is off for parity) became the critical path of the step: 11 of 17.5 ms.  Here the GRU is split the
// way its arithmetic allows:
//   * everything that is a big GEMM -- the input projections X W_ih^T of ALL time steps and, in the
//     backward, dX, dW_ih, dW_hh -- runs on agnn_gemm (tensor cores, 3xTF32);
//   * only the truly sequential part stays in these two kernels: per time step the H x 3H recurrent
//     mat-vec, the gate nonlinearities and (backward) the carried dh.
// Sequences are independent, so there is no inter-CTA synchronisation at all: a CTA owns kSeq sequences
// of one direction for their whole length, 3H threads, and every thread keeps its row (forward) or its
// column block (backward) of W_hh in REGISTERS for all T steps -- the weights are read from memory once.
// The per-step inputs are prefetched kPrefetch steps ahead to hide the global-load latency of the
// dependent chain.  PyTorch gate order (r, z, n):
//   r = s(gi_r + gh_r)  z = s(gi_z + gh_z)  n = tanh(gi_n + r * gh_n)  h' = (1 - z) n + z h,  gh = W_hh h + b_hh
#include <cstring>

#include "common.cuh"

namespace agnn {
namespace {

constexpr int kSeq = 2;       // sequences per CTA (reuse of the register-resident weights)
constexpr int kPrefetch = 4;  // time steps of input prefetched ahead

struct GruParams {
  int batch, steps, n_dir;
  // per direction d: gi [B, T, 3H] (input projections + b_ih), w_hh [3H, H], b_hh [3H]
  const float* gi[2];
  const float* w_hh[2];
  const float* b_hh[2];
  float* out;          // [B, T, n_dir * H]
  int64_t ld_out;      // n_dir * H
  float* gates[2];     // [B, T, 4H]: r, z, n, gh_n  (saved for the backward)
  // backward
  const float* dout;   // [B, T, n_dir * H]
  float* dgi[2];       // [B, T, 3H]
  float* dgh[2];       // [B, T, 3H]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int H>
__global__ void __launch_bounds__(3 * H, 1) gru_fwd_kernel(const __grid_constant__ GruParams p) {
  __shared__ __align__(16) float h_s[kSeq][H];
  __shared__ float gh_s[kSeq][3 * H];
  const int j = threadIdx.x;                     // gate row 0 .. 3H-1
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * kSeq;
  const int T = p.steps;
  float w[H];
  {
    const float* wr = p.w_hh[dir] + (int64_t)j * H;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(wr + k));
      w[k] = t.x; w[k + 1] = t.y; w[k + 2] = t.z; w[k + 3] = t.w;
    }
  }
  const float bias = __ldg(p.b_hh[dir] + j);
  for (int s = 0; s < kSeq; ++s)
    if (j < H) h_s[s][j] = 0.f;
  // threads j < H own hidden unit j of every sequence of the CTA: prefetch ring of their gi inputs
  float gi_ring[kPrefetch][kSeq][3];
  auto time_of = [&](int step) { return dir == 0 ? step : T - 1 - step; };
  auto load_gi = [&](int step, float (&dst)[kSeq][3]) {
    if (j < H && step < T) {
      const int t = time_of(step);
#pragma unroll
      for (int s = 0; s < kSeq; ++s) {
        const int b = b0 + s;
        if (b < p.batch) {
          const float* g = p.gi[dir] + ((int64_t)b * T + t) * (3 * H);
          dst[s][0] = __ldg(g + j); dst[s][1] = __ldg(g + H + j); dst[s][2] = __ldg(g + 2 * H + j);
        }
      }
    }
  };
#pragma unroll
  for (int q = 0; q < kPrefetch; ++q) load_gi(q, gi_ring[q]);
  __syncthreads();
  for (int step0 = 0; step0 < T; step0 += kPrefetch) {
#pragma unroll
    for (int q = 0; q < kPrefetch; ++q) {
      const int step = step0 + q;
      if (step >= T) break;
      // 1. recurrent mat-vec: gh[s][j] = b_hh[j] + W_hh[j, :] . h[s]
      float acc[kSeq];
#pragma unroll
      for (int s = 0; s < kSeq; ++s) acc[s] = bias;
#pragma unroll
      for (int k = 0; k < H; k += 4) {
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          const float4 hv = *reinterpret_cast<const float4*>(&h_s[s][k]);
          acc[s] = fmaf(w[k], hv.x, acc[s]);
          acc[s] = fmaf(w[k + 1], hv.y, acc[s]);
          acc[s] = fmaf(w[k + 2], hv.z, acc[s]);
          acc[s] = fmaf(w[k + 3], hv.w, acc[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < kSeq; ++s) gh_s[s][j] = acc[s];
      __syncthreads();
      // 2. gates and the new hidden state (threads j < H)
      if (j < H) {
        const int t = time_of(step);
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          const int b = b0 + s;
          if (b < p.batch) {
            const float ghn = gh_s[s][2 * H + j];
            const float r = sigmoidf_(gi_ring[q][s][0] + gh_s[s][j]);
            const float z = sigmoidf_(gi_ring[q][s][1] + gh_s[s][H + j]);
            const float n = tanhf(gi_ring[q][s][2] + r * ghn);
            const float hn = (1.f - z) * n + z * h_s[s][j];
            h_s[s][j] = hn;
            const int64_t row = (int64_t)b * T + t;
            p.out[row * p.ld_out + dir * H + j] = hn;
            if (p.gates[dir]) {
              float* g = p.gates[dir] + row * (4 * H);
              g[j] = r; g[H + j] = z; g[2 * H + j] = n; g[3 * H + j] = ghn;
            }
          }
        }
      }
      load_gi(step + kPrefetch, gi_ring[q]);      // refill this ring slot for kPrefetch steps ahead
      __syncthreads();
    }
  }
}

template <int H>
__global__ void __launch_bounds__(3 * H, 1) gru_bwd_kernel(const __grid_constant__ GruParams p) {
  __shared__ float dgh_s[kSeq][3 * H];
  __shared__ float part_s[kSeq][3][H];
  const int j = threadIdx.x;
  const int gate = j / H, k = j - gate * H;      // this thread: column k of gate block `gate`
  const int dir = blockIdx.y;
  const int b0 = blockIdx.x * kSeq;
  const int T = p.steps;
  constexpr int kPf = H >= 128 ? 2 : kPrefetch;  // 128 weight registers leave room for a shorter ring
  float w[H];                                    // w[i] = W_hh[gate * H + i, k]
#pragma unroll
  for (int i = 0; i < H; ++i) w[i] = __ldg(p.w_hh[dir] + (int64_t)(gate * H + i) * H + k);
  float carry[kSeq];                             // dL/dh carried to the previous step (threads j < H)
#pragma unroll
  for (int s = 0; s < kSeq; ++s) carry[s] = 0.f;
  // processing order is the reverse of the forward's: forward-direction t = T-1 .. 0, reverse t = 0 .. T-1
  auto time_of = [&](int step) { return dir == 0 ? T - 1 - step : step; };
  struct In { float g, r, z, n, ghn, hp; };
  In ring[kPf][kSeq];
  auto load_in = [&](int step, In (&dst)[kSeq]) {
    if (j < H && step < T) {
      const int t = time_of(step);
      const int tp = dir == 0 ? t - 1 : t + 1;   // time of h_{prev} in the forward recurrence
#pragma unroll
      for (int s = 0; s < kSeq; ++s) {
        const int b = b0 + s;
        if (b < p.batch) {
          const int64_t row = (int64_t)b * T + t;
          const float* g = p.gates[dir] + row * (4 * H);
          dst[s].g = __ldg(p.dout + row * p.ld_out + dir * H + j);
          dst[s].r = __ldg(g + j); dst[s].z = __ldg(g + H + j); dst[s].n = __ldg(g + 2 * H + j);
          dst[s].ghn = __ldg(g + 3 * H + j);
          dst[s].hp = (tp >= 0 && tp < T) ? __ldg(p.out + ((int64_t)b * T + tp) * p.ld_out + dir * H + j) : 0.f;
        }
      }
    }
  };
#pragma unroll
  for (int q = 0; q < kPf; ++q) load_in(q, ring[q]);
  for (int step0 = 0; step0 < T; step0 += kPf) {
#pragma unroll
    for (int q = 0; q < kPf; ++q) {
      const int step = step0 + q;
      if (step >= T) break;
      float keep[kSeq];                          // dh_total * z, the direct path to h_prev
      if (j < H) {
        const int t = time_of(step);
#pragma unroll
        for (int s = 0; s < kSeq; ++s) {
          const int b = b0 + s;
          keep[s] = 0.f;
          if (b < p.batch) {
            const In& in = ring[q][s];
            const float dh = in.g + carry[s];
            const float dn_pre = dh * (1.f - in.z) * (1.f - in.n * in.n);
            const float dz_pre = dh * (in.hp - in.n) * in.z * (1.f - in.z);
            const float dr_pre = dn_pre * in.ghn * in.r * (1.f - in.r);
            keep[s] = dh * in.z;
            const int64_t row = ((int64_t)b * T + t) * (3 * H);
            p.dgi[dir][row + j] = dr_pre; p.dgi[dir][row + H + j] = dz_pre; p.dgi[dir][row + 2 * H + j] = dn_pre;
            const float dghn = dn_pre * in.r;
            p.dgh[dir][row + j] = dr_pre; p.dgh[dir][row + H + j] = dz_pre; p.dgh[dir][row + 2 * H + j] = dghn;
            dgh_s[s][j] = dr_pre; dgh_s[s][H + j] = dz_pre; dgh_s[s][2 * H + j] = dghn;
          } else {
            dgh_s[s][j] = 0.f; dgh_s[s][H + j] = 0.f; dgh_s[s][2 * H + j] = 0.f;
          }
        }
      }
      __syncthreads();                           // (A)
      // dh_prev[k] += sum_i dgh[gate*H + i] * W_hh[gate*H + i, k]   (three gate blocks, summed below)
      float acc[kSeq];
#pragma unroll
      for (int s = 0; s < kSeq; ++s) acc[s] = 0.f;
#pragma unroll
      for (int i = 0; i < H; ++i) {
#pragma unroll
        for (int s = 0; s < kSeq; ++s) acc[s] = fmaf(w[i], dgh_s[s][gate * H + i], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < kSeq; ++s) part_s[s][gate][k] = acc[s];
      __syncthreads();                           // (B)
      if (j < H) {
#pragma unroll
        for (int s = 0; s < kSeq; ++s) carry[s] = keep[s] + (part_s[s][0][j] + part_s[s][1][j]) + part_s[s][2][j];
      }
      load_in(step + kPf, ring[q]);
      // no third barrier: dgh_s is rewritten only after barrier (B), which every thread passes after its
      // mat-vec reads; part_s is rewritten only after the next barrier (A), which follows these reads
    }
  }
}

int check_gru(const char* what, int batch, int steps, int hidden, int n_dir) {
  if (batch < 0 || steps < 0 || (n_dir != 1 && n_dir != 2))
    return fail(AGNN_ERR_ARG, "%s: bad sizes (batch=%d steps=%d n_dir=%d)", what, batch, steps, n_dir);
  if (hidden != 32 && hidden != 64 && hidden != 128)
    return fail(AGNN_ERR_UNSUPPORTED, "%s: hidden size %d (register-resident W_hh supports 32, 64, 128)", what, hidden);
  return AGNN_OK;
}

template <int H>
int launch_fwd(const GruParams& p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.batch, kSeq), (unsigned)p.n_dir);
  gru_fwd_kernel<H><<<grid, 3 * H, 0, st>>>(p);
  return check_launch("gru_fwd");
}
template <int H>
int launch_bwd(const GruParams& p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.batch, kSeq), (unsigned)p.n_dir);
  gru_bwd_kernel<H><<<grid, 3 * H, 0, st>>>(p);
  return check_launch("gru_bwd");
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_gru_supported(int hidden) { return hidden == 32 || hidden == 64 || hidden == 128; }

extern "C" int agnn_gru_fwd(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* gi,
                            const float* const* w_hh, const float* const* b_hh, float* out, float* const* gates,
                            agnn_stream_t stream) {
  int rc = check_gru("gru_fwd", batch, steps, hidden, n_dir);
  if (rc) return rc;
  if (!gi || !w_hh || !b_hh || !out) return fail(AGNN_ERR_ARG, "gru_fwd: null pointer");
  if (batch == 0 || steps == 0) return AGNN_OK;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.batch = batch; p.steps = steps; p.n_dir = n_dir; p.out = out; p.ld_out = (int64_t)n_dir * hidden;
  for (int d = 0; d < n_dir; ++d) {
    if (!gi[d] || !w_hh[d] || !b_hh[d] || !aligned16(w_hh[d])) return fail(AGNN_ERR_ARG, "gru_fwd: null / unaligned operand");
    p.gi[d] = gi[d]; p.w_hh[d] = w_hh[d]; p.b_hh[d] = b_hh[d];
    p.gates[d] = gates ? gates[d] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (hidden == 128) return launch_fwd<128>(p, st);
  if (hidden == 64) return launch_fwd<64>(p, st);
  return launch_fwd<32>(p, st);
}

extern "C" int agnn_gru_bwd(int32_t batch, int32_t steps, int32_t hidden, int32_t n_dir, const float* const* w_hh,
                            const float* out, const float* const* gates, const float* dout, float* const* dgi,
                            float* const* dgh, agnn_stream_t stream) {
  int rc = check_gru("gru_bwd", batch, steps, hidden, n_dir);
  if (rc) return rc;
  if (!w_hh || !out || !gates || !dout || !dgi || !dgh) return fail(AGNN_ERR_ARG, "gru_bwd: null pointer");
  if (batch == 0 || steps == 0) return AGNN_OK;
  GruParams p;
  memset(&p, 0, sizeof(p));
  p.batch = batch; p.steps = steps; p.n_dir = n_dir; p.out = const_cast<float*>(out);
  p.ld_out = (int64_t)n_dir * hidden; p.dout = dout;
  for (int d = 0; d < n_dir; ++d) {
    if (!w_hh[d] || !gates[d] || !dgi[d] || !dgh[d]) return fail(AGNN_ERR_ARG, "gru_bwd: null operand");
    p.w_hh[d] = w_hh[d]; p.gates[d] = const_cast<float*>(gates[d]); p.dgi[d] = dgi[d]; p.dgh[d] = dgh[d];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (hidden == 128) return launch_bwd<128>(p, st);
  if (hidden == 64) return launch_bwd<64>(p, st);
  return launch_bwd<32>(p, st);
}
