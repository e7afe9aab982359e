// HBM-bound helpers of the sampling path: layout conversion on the API edge, resampling, GEGLU, row softmax,
// time embedding, GEMV-class linears and the fused scheduler updates.  All are single-pass, coalesced and
// (where the layout allows) 128-bit vectorised; none of them has data reuse worth staging in shared memory
// except the two transposes.
#include "common.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// NC[D]HW fp32 <-> NDHWC bf16: 32x32 shared-memory tile transpose (coalesced on both sides).
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, long long spatial,
                                    h16* __restrict__ y, int pitch) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xb = x + (long long)n * C * spatial;
  h16* yb = y + (long long)n * spatial * pitch;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long s = s0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && s < spatial) ? xb[(long long)c * spatial + s] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long s = s0 + j;
    const int c = c0 + threadIdx.x;
    if (s < spatial && c < pitch) yb[s * pitch + c] = f2h(tile[threadIdx.x][j]);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int C, long long spatial, int pitch,
                                    float* __restrict__ y) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const T* xb = x + (long long)n * spatial * pitch;
  float* yb = y + (long long)n * C * spatial;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long s = s0 + j;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (s < spatial && c < C) {
      if constexpr (sizeof(T) == 2) v = h2f(xb[s * pitch + c]);
      else v = xb[s * pitch + c];
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long s = s0 + threadIdx.x;
    if (c < C && s < spatial) yb[(long long)c * spatial + s] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------------------------------------
// nearest x2 upsample / 2x average pool on channels-last bf16; one thread per (output voxel, 8 channels)
// ------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int N, int D, int H, int W, int pv, int dims,
                                  uint4* __restrict__ y) {
  pdl_entry();
  const int OD = dims == 3 ? 2 * D : D, OH = 2 * H, OW = 2 * W;
  const long long total = (long long)N * OD * OH * OW * pv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int v = (int)(t % pv); t /= pv;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD); t /= OD;
    const int n = (int)t;
    const int id = dims == 3 ? od >> 1 : od;
    y[idx] = __ldg(x + ((((long long)n * D + id) * H + (oh >> 1)) * W + (ow >> 1)) * pv + v);
  }
}

__global__ void avgpool2_kernel(const uint4* __restrict__ x, int N, int D, int H, int W, int pv, int dims,
                                uint4* __restrict__ y) {
  pdl_entry();
  const int OD = dims == 3 ? D / 2 : D, OH = H / 2, OW = W / 2;
  const int kd = dims == 3 ? 2 : 1;
  const float inv = 1.0f / (float)(kd * 4);
  const long long total = (long long)N * OD * OH * OW * pv;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int v = (int)(t % pv); t /= pv;
    const int ow = (int)(t % OW); t /= OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD); t /= OD;
    const int n = (int)t;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < kd; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) {
          const int id = dims == 3 ? od * 2 + a : od;
          uint4 q = __ldg(x + ((((long long)n * D + id) * H + (oh * 2 + b)) * W + (ow * 2 + c)) * pv + v);
          float f[8];
          unpack8(q, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    y[idx] = pack8(acc);
  }
}

__global__ void axpy_h16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, float alpha,
                                 uint4* __restrict__ y, long long nvec) {
  pdl_entry();
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nvec;
       idx += (long long)gridDim.x * blockDim.x) {
    float fa[8], fb[8];
    unpack8(__ldg(a + idx), fa);
    unpack8(__ldg(b + idx), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] = fmaf(alpha, fb[j], fa[j]);
    y[idx] = pack8(fa);
  }
}

__global__ void copy_channels_kernel(const h16* __restrict__ src, int C, int src_pitch,
                                     h16* __restrict__ dst, int dst_pitch, int dst_off, long long rows, int vec) {
  pdl_entry();
  const int per_row = C / vec;
  const long long total = rows * per_row;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / per_row;
    const int c = (int)(idx % per_row) * vec;
    if (vec == 8)
      *reinterpret_cast<uint4*>(dst + r * dst_pitch + dst_off + c) = __ldg(reinterpret_cast<const uint4*>(src + r * src_pitch + c));
    else
      dst[r * dst_pitch + dst_off + c] = src[r * src_pitch + c];
  }
}

// ------------------------------------------------------------------------------------------------
// GEGLU: y = x[:, :H] * gelu_erf(x[:, H:])
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__global__ void geglu_kernel(const h16* __restrict__ x, long long M, int H, int x_pitch,
                             h16* __restrict__ y, int y_pitch) {
  pdl_entry();
  const int HV = H / 8;
  const long long total = M * HV;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / HV;
    const int h = (int)(idx % HV) * 8;
    float a[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + m * x_pitch + h)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + m * x_pitch + H + h)), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= gelu_erf(g[j]);
    *reinterpret_cast<uint4*>(y + m * y_pitch + h) = pack8(a);
  }
}

// ------------------------------------------------------------------------------------------------
// Row softmax: fp32 scores -> bf16 probabilities. TPR threads cooperate on one row.
// ------------------------------------------------------------------------------------------------
template <int TPR>
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long M, int S, long long s_pitch,
                                    h16* __restrict__ p, long long p_pitch) {
  pdl_entry();
  constexpr int RPB = 256 / TPR;
  const long long row = (long long)blockIdx.x * RPB + threadIdx.x / TPR;
  const int t = threadIdx.x % TPR;
  __shared__ float red[8];
  const bool live = row < M;
  const float* sr = s + (live ? row : 0) * s_pitch;
  float mx = -INFINITY;
  if (live)
    for (int c = t; c < S; c += TPR) mx = fmaxf(mx, sr[c]);
  mx = warp_max(mx);
  if constexpr (TPR > 32) {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
  }
  float sum = 0.f;
  if (live)
    for (int c = t; c < S; c += TPR) sum += __expf(sr[c] - mx);
  sum = warp_sum(sum);
  if constexpr (TPR > 32) {
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += red[i];
  }
  if (!live) return;
  const float inv = 1.0f / sum;
  h16* pr = p + row * p_pitch;
  for (int c = t; c < S; c += TPR) pr[c] = f2h(__expf(sr[c] - mx) * inv);
  for (long long c = S + t; c < p_pitch; c += TPR) pr[c] = f2h(0.f);
}

// One-pass variant: the score GEMM's epilogue already left (max, sum exp) per 256-column tile of every row, so the
// row maximum and denominator come from a few hundred partials and the scores are read exactly once.
__global__ void softmax_rows_partials_kernel(const float* __restrict__ s, int S, long long s_pitch,
                                             const float2* __restrict__ part, int n_tiles,
                                             h16* __restrict__ p, long long p_pitch) {
  pdl_entry();
  const long long row = blockIdx.x;
  const int t = threadIdx.x;
  __shared__ float red[8];
  const float2* pr = part + row * n_tiles;
  float mx = -INFINITY;
  for (int i = t; i < n_tiles; i += 256) mx = fmaxf(mx, pr[i].x);
  mx = warp_max(mx);
  if ((t & 31) == 0) red[t >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = t; i < n_tiles; i += 256) {
    const float2 q = pr[i];
    if (q.x > -INFINITY) sum += q.y * __expf(q.x - mx);
  }
  sum = warp_sum(sum);
  if ((t & 31) == 0) red[t >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  const float* sr = s + row * s_pitch;
  h16* po = p + row * p_pitch;
  const bool vec = (s_pitch % 4 == 0) && (p_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(s) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p) & 7) == 0);
  if (vec) {
    const int S4 = S / 4;
    for (int c = t; c < S4; c += 256) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(sr) + c);
      h162 lo = f2h2(__expf(v.x - mx) * inv, __expf(v.y - mx) * inv);
      h162 hi = f2h2(__expf(v.z - mx) * inv, __expf(v.w - mx) * inv);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&lo);
      o.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(po + 4 * (long long)c) = o;
    }
    for (int c = S4 * 4 + t; c < S; c += 256) po[c] = f2h(__expf(sr[c] - mx) * inv);
  } else {
    for (int c = t; c < S; c += 256) po[c] = f2h(__expf(sr[c] - mx) * inv);
  }
  for (long long c = S + t; c < p_pitch; c += 256) po[c] = f2h(0.f);
}

// ------------------------------------------------------------------------------------------------
// time embedding + GEMV-class linear
// ------------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int N, int dim, float max_period,
                                          float* __restrict__ emb) {
  pdl_entry();
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * dim) return;
  const int n = idx / dim, i = idx % dim;
  float v = 0.f;
  if (i < 2 * half) {
    const int k = i < half ? i : i - half;
    // exponent = -ln(max_period) * k, freq = exp(exponent / half): same operation order as the reference
    const float freq = expf((-logf(max_period) * (float)k) / (float)half);
    const float arg = t[n] * freq;
    v = i < half ? cosf(arg) : sinf(arg);
  }
  emb[idx] = v;
}

// one warp per output feature; loops over the (few) rows.  VEC: K % 128 == 0 and 16-byte aligned rows — every lane
// issues all its float4 weight loads before the first FMA (a time-embedding projection is one 4 KB row per warp: the
// scalar form spent ~15 us per call waiting on 32 dependent-latency loads).
template <bool VEC>
__global__ void small_linear_kernel(const float* __restrict__ x, int M, int K, const float* __restrict__ W,
                                    const float* __restrict__ b, int O, int act_in, int act_out,
                                    float* __restrict__ y) {
  pdl_entry();
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= O) return;
  const int lane = threadIdx.x & 31;
  const float* w = W + (long long)o * K;
  for (int m = 0; m < M; ++m) {
    const float* xr = x + (long long)m * K;
    float acc = 0.f;
    if constexpr (VEC) {
      for (int k0 = 0; k0 < K; k0 += 1024) {
        float4 wv[8], xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = k0 + i * 128 + lane * 4;
          if (k < K) {
            wv[i] = __ldg(reinterpret_cast<const float4*>(w + k));
            xv[i] = *reinterpret_cast<const float4*>(xr + k);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (k0 + i * 128 + lane * 4 < K) {
            acc = fmaf(apply_act(xv[i].x, act_in), wv[i].x, acc);
            acc = fmaf(apply_act(xv[i].y, act_in), wv[i].y, acc);
            acc = fmaf(apply_act(xv[i].z, act_in), wv[i].z, acc);
            acc = fmaf(apply_act(xv[i].w, act_in), wv[i].w, acc);
          }
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) acc = fmaf(apply_act(xr[k], act_in), __ldg(w + k), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) y[(long long)m * O + o] = apply_act(acc + (b ? b[o] : 0.f), act_out);
  }
}

// ------------------------------------------------------------------------------------------------
// scheduler steps
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ddim_one(const b200_ddim_coef& c, float m, float s, float nz, bool has_noise, float& p,
                                         float& x0) {
  float eps;
  if (c.prediction_type == B200_PRED_EPSILON) {
    x0 = (s - c.sqrt_beta_prod_t * m) / c.sqrt_alpha_prod_t;
    eps = m;
  } else if (c.prediction_type == B200_PRED_SAMPLE) {
    x0 = m;
    eps = (s - c.sqrt_alpha_prod_t * x0) / c.sqrt_beta_prod_t;
  } else {
    x0 = c.sqrt_alpha_prod_t * s - c.sqrt_beta_prod_t * m;
    eps = c.sqrt_alpha_prod_t * m + c.sqrt_beta_prod_t * s;
  }
  if (c.clip) x0 = fminf(fmaxf(x0, c.clip_min), c.clip_max);
  p = c.sqrt_alpha_prod_prev * x0 + c.dir_coef * eps;
  if (has_noise) p += c.sigma * nz;
}

// 128-bit loads / stores, two independent vectors per thread and iteration (the update is 2 reads + 2 writes of 4 B per
// element: pure HBM streaming, so what matters is bytes in flight per SM); scalar tail for n % 4.
// VEC = 0: pointers not 16-byte aligned -> scalar path.
template <int VEC>
__global__ void __launch_bounds__(256) ddim_step_kernel(const float* __restrict__ eps_in, const float* __restrict__ x,
                                                        const float* __restrict__ noise, b200_ddim_coef c,
                                                        float* __restrict__ prev, float* __restrict__ x0_out, long long n) {
  pdl_entry();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  const bool has_noise = noise != nullptr;
  long long done = 0;
  if (VEC) {
    const long long nv = n >> 2;
    const float4* e4 = reinterpret_cast<const float4*>(eps_in);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* n4 = reinterpret_cast<const float4*>(noise);
    float4* p4 = reinterpret_cast<float4*>(prev);
    float4* o4 = reinterpret_cast<float4*>(x0_out);
    for (long long i = tid; i < nv; i += 2 * nthr) {
      const long long i2 = i + nthr;
      const bool two = i2 < nv;
      const float4 ma = __ldg(e4 + i), sa = __ldg(x4 + i);
      const float4 mb = two ? __ldg(e4 + i2) : ma, sb = two ? __ldg(x4 + i2) : sa;
      float4 za = make_float4(0.f, 0.f, 0.f, 0.f), zb = za;
      if (has_noise) { za = __ldg(n4 + i); if (two) zb = __ldg(n4 + i2); }
      float4 pa, oa, pb, ob;
      ddim_one(c, ma.x, sa.x, za.x, has_noise, pa.x, oa.x); ddim_one(c, ma.y, sa.y, za.y, has_noise, pa.y, oa.y);
      ddim_one(c, ma.z, sa.z, za.z, has_noise, pa.z, oa.z); ddim_one(c, ma.w, sa.w, za.w, has_noise, pa.w, oa.w);
      ddim_one(c, mb.x, sb.x, zb.x, has_noise, pb.x, ob.x); ddim_one(c, mb.y, sb.y, zb.y, has_noise, pb.y, ob.y);
      ddim_one(c, mb.z, sb.z, zb.z, has_noise, pb.z, ob.z); ddim_one(c, mb.w, sb.w, zb.w, has_noise, pb.w, ob.w);
      p4[i] = pa;
      if (x0_out) o4[i] = oa;
      if (two) { p4[i2] = pb; if (x0_out) o4[i2] = ob; }
    }
    done = nv << 2;
  }
  for (long long i = done + tid; i < n; i += nthr) {
    float p, x0;
    ddim_one(c, eps_in[i], x[i], has_noise ? noise[i] : 0.f, has_noise, p, x0);
    prev[i] = p;
    if (x0_out) x0_out[i] = x0;
  }
}

__global__ void ddpm_step_kernel(const float* __restrict__ eps_in, const float* __restrict__ x,
                                 const float* __restrict__ noise, const float* __restrict__ pred_var,
                                 b200_ddpm_coef c, float* __restrict__ prev, float* __restrict__ x0_out,
                                 long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float m = eps_in[i], s = x[i];
    float x0;
    if (c.prediction_type == B200_PRED_EPSILON) x0 = (s - c.sqrt_beta_prod_t * m) / c.sqrt_alpha_prod_t;
    else if (c.prediction_type == B200_PRED_SAMPLE) x0 = m;
    else x0 = c.sqrt_alpha_prod_t * s - c.sqrt_beta_prod_t * m;
    if (c.clip) x0 = fminf(fmaxf(x0, c.clip_min), c.clip_max);
    float p = c.coef_x0 * x0 + c.coef_xt * s;
    if (noise) {
      float sig = c.sigma;
      if (c.var_mode == 1) sig = sqrtf(pred_var[i]);
      else if (c.var_mode == 2) {
        const float frac = (pred_var[i] + 1.0f) / 2.0f;
        sig = sqrtf(frac * c.max_log + (1.0f - frac) * c.min_log);
      }
      p += sig * noise[i];
    }
    prev[i] = p;
    if (x0_out) x0_out[i] = x0;
  }
}

// tanh approximation of the standard normal CDF used by the reference (inferer.py:279-283)
__device__ __forceinline__ float approx_normal_cdf(float x) {
  return 0.5f * (1.0f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
}

__global__ void ddpm_kl_kernel(const float* __restrict__ x0, const float* __restrict__ xt,
                               const float* __restrict__ mo, b200_kl_coef c, float* __restrict__ kl_out,
                               double* __restrict__ sample_sum, long long per_sample) {
  pdl_entry();
  const int n = blockIdx.y;
  const long long base = (long long)n * per_sample;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x) {
    const float a = x0[base + i], s = xt[base + i], m = mo[base + i];
    float p0;
    if (c.prediction_type == B200_PRED_EPSILON) p0 = (s - c.sqrt_beta_prod_t * m) / c.sqrt_alpha_prod_t;
    else if (c.prediction_type == B200_PRED_SAMPLE) p0 = m;
    else p0 = c.sqrt_alpha_prod_t * s - c.sqrt_beta_prod_t * m;
    if (c.clip) p0 = fminf(fmaxf(p0, -1.0f), 1.0f);
    const float pred_mean = c.coef_x0 * p0 + c.coef_xt * s;
    float kl;
    if (c.is_t0) {
      // -log p(x_0 | x_1): discretised Gaussian (inferer.py:285-321)
      const float centered = a - pred_mean;
      const float inv_stdv = expf(-0.5f * c.log_pred_var);
      const float cdf_plus = approx_normal_cdf(inv_stdv * (centered + c.bin_width / 2));
      const float cdf_min = approx_normal_cdf(inv_stdv * (centered - c.bin_width / 2));
      float lp;
      if (a < -0.999f) lp = logf(fmaxf(cdf_plus, 1e-12f));
      else if (a > 0.999f) lp = logf(fmaxf(1.0f - cdf_min, 1e-12f));
      else lp = logf(fmaxf(cdf_plus - cdf_min, 1e-12f));
      kl = -lp;
    } else {
      const float post_mean = c.coef_x0 * a + c.coef_xt * s;
      const float d = post_mean - pred_mean;
      kl = 0.5f * (-1.0f + c.log_pred_var - c.log_post_var + expf(c.log_post_var - c.log_pred_var) +
                   d * d * expf(-c.log_pred_var));
    }
    if (kl_out) kl_out[base + i] = kl;
    acc += kl;
  }
  __shared__ double red[8];
  double d = (double)acc;
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    atomicAdd(sample_sum + n, t);
  }
}

struct PndmPtrs { const float* h[4]; };

__global__ void pndm_step_kernel(PndmPtrs hp, const float* __restrict__ x, b200_pndm_coef c,
                                 float* __restrict__ prev, float* __restrict__ eps_out, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float e = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < c.n_hist) e = fmaf(c.w[k], hp.h[k][i], e);
    if (eps_out) eps_out[i] = e;
    if (prev) {
      const float s = x[i];
      if (c.prediction_type == B200_PRED_V) e = c.v_alpha * e + c.v_beta * s;
      prev[i] = c.sample_coeff * s - c.eps_coeff * e;
    }
  }
}

__global__ void add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                 const float* __restrict__ ca, const float* __restrict__ cb, float sign_b,
                                 long long per_sample, float* __restrict__ out) {
  pdl_entry();
  const int n = blockIdx.y;
  const float a = ca[n], b = cb[n] * sign_b;
  const long long base = (long long)n * per_sample;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample;
       i += (long long)gridDim.x * blockDim.x)
    out[base + i] = a * x0[base + i] + b * noise[base + i];
}

__global__ void exp_half_clamped_kernel(const float* __restrict__ x, float lo, float hi, float* __restrict__ y, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = expf(fminf(fmaxf(x[i], lo), hi) / 2.0f);
}

__global__ void fma_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                               float* __restrict__ y, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = a[i] + b[i] * c[i];
}

__global__ void scale_f32_kernel(const float* __restrict__ x, float mul, float div, float* __restrict__ y, long long n) {
  pdl_entry();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] * mul / div;
}

static unsigned grid_for(long long work, int threads = 256, int waves = 8) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)waves * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}


// ------------------------------------------------------------------------------------------------
// Tap reformulations of the two degenerate convolutions of the UNet (conv_in: 1 -> C, out: C -> 1): both are
// HBM-bound, but as 27-tap implicit GEMMs they pad K (or N) 2-60x.  tap_gather builds the tiny im2col matrix
// X2[v][tap * Cin + c] so that conv_in becomes ONE 64-wide K chunk; tap_sum evaluates
// out[v][co] = b[co] + sum_tap Y[v + off(tap)][tap * Cout + co] after a 1x1x1 GEMM Y = x W2^T that reads x once.
// ------------------------------------------------------------------------------------------------
struct TapGeom {
  int N, D, H, W, OD, OH, OW, kd, kh, kw, sd, sh, sw, pd, ph, pw;
};

__global__ void tap_gather_kernel(const h16* __restrict__ x, int C, int x_pitch, TapGeom g,
                                  h16* __restrict__ out, int out_pitch) {
  pdl_entry();
  // one thread per (output voxel, 8-column vector): the voxel coordinates are decoded once, the row is written with
  // 16-byte stores (out_pitch is a multiple of 8: it is the K pitch of the GEMM that follows)
  const int taps = g.kd * g.kh * g.kw;
  const int vecs = out_pitch >> 3;
  const long long total = (long long)g.N * g.OD * g.OH * g.OW * vecs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v8 = (int)(i % vecs);
    long long v = i / vecs;
    const int ow = (int)(v % g.OW); long long t = v / g.OW;
    const int oh = (int)(t % g.OH); t /= g.OH;
    const int od = (int)(t % g.OD); const int n = (int)(t / g.OD);
    const h16* xn = x + (long long)n * g.D * g.H * g.W * x_pitch;
    __align__(16) h16 vals[8];
    int col = v8 * 8;
    int tap = col / C, c = col - tap * C;
#pragma unroll
    for (int e = 0; e < 8; ++e, ++col) {
      h16 val = f2h(0.f);
      if (tap < taps) {
        const int cw = tap % g.kw, bh = (tap / g.kw) % g.kh, ad = tap / (g.kw * g.kh);
        const int iw = ow * g.sw + cw - g.pw, ih = oh * g.sh + bh - g.ph, id = od * g.sd + ad - g.pd;
        if (iw >= 0 && iw < g.W && ih >= 0 && ih < g.H && id >= 0 && id < g.D)
          val = xn[(((long long)id * g.H + ih) * g.W + iw) * x_pitch + c];
      }
      vals[e] = val;
      if (++c == C) { c = 0; ++tap; }
    }
    *reinterpret_cast<uint4*>(out + v * out_pitch + v8 * 8) = *reinterpret_cast<const uint4*>(vals);
  }
}

template <int COUT>
__global__ void tap_sum_kernel(const float* __restrict__ y, int y_pitch, TapGeom g, const float* __restrict__ bias,
                               void* __restrict__ out, int out_pitch, int out_dtype) {
  pdl_entry();
  const long long total = (long long)g.N * g.OD * g.OH * g.OW;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(v % g.OW); long long t = v / g.OW;
    const int oh = (int)(t % g.OH); t /= g.OH;
    const int od = (int)(t % g.OD); const int n = (int)(t / g.OD);
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = bias ? bias[co] : 0.f;
    for (int a = 0; a < g.kd; ++a) {
      const int id = od + a - g.pd;
      if (id < 0 || id >= g.D) continue;
      for (int b = 0; b < g.kh; ++b) {
        const int ih = oh + b - g.ph;
        if (ih < 0 || ih >= g.H) continue;
        const float* row = y + (((long long)n * g.D + id) * g.H + ih) * g.W * y_pitch + ((a * g.kh + b) * g.kw) * COUT;
#pragma unroll 3
        for (int c = 0; c < g.kw; ++c) {
          const int iw = ow + c - g.pw;
          if (iw < 0 || iw >= g.W) continue;
          const float* src = row + (long long)iw * y_pitch + c * COUT;
#pragma unroll
          for (int co = 0; co < COUT; ++co) acc[co] += __ldg(src + co);
        }
      }
    }
    if (out_dtype == B200_DT_H16) {
      h16* o = reinterpret_cast<h16*>(out) + v * out_pitch;
      for (int co = 0; co < out_pitch; ++co) o[co] = f2h(co < COUT ? acc[co < COUT ? co : 0] : 0.f);
    } else {
      float* o = reinterpret_cast<float*>(out) + v * out_pitch;
      for (int co = 0; co < out_pitch; ++co) o[co] = co < COUT ? acc[co < COUT ? co : 0] : 0.f;
    }
  }
}


// token + absolute position embedding rows -> bf16 (nets/transformer.py:97-99)
__global__ void embed_tokens_kernel(const long long* __restrict__ tokens, long long M, int seq_len, int pos0,
                                    const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, int C,
                                    h16* __restrict__ out, int pitch, const int* __restrict__ pos_dev) {
  pdl_entry();
  if (pos_dev) pos0 = *pos_dev;
  const long long total = M * pitch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % pitch);
    const long long m = i / pitch;
    float v = 0.f;
    if (c < C) v = tok_emb[tokens[m] * C + c] + pos_emb[(long long)(pos0 + (int)(m % seq_len)) * C + c];
    out[i] = f2h(v);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_nchw_to_nhwc(const float* x, int32_t N, int32_t C, int64_t spatial, void* y, int32_t pitch,
                                 void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && N >= 1 && C >= 1 && spatial >= 1 && pitch >= C, "nchw_to_nhwc: bad arguments");
  const long long bx = (spatial + 31) / 32;
  B200_CHECK_ARG(bx < (1ll << 31) && N <= 65535, "nchw_to_nhwc: extent too large");
  dim3 grid((unsigned)bx, (pitch + 31) / 32, N);
  B200_CUDA(b200::launch_pdl(nchw_to_nhwc_kernel, grid, dim3(32, 8), 0, stream, x, C, spatial, reinterpret_cast<h16*>(y), pitch));
  B200_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return B200_OK;
}

extern "C" int b200_nhwc_to_nchw(const void* x, int32_t x_dtype, int32_t N, int32_t C, int64_t spatial,
                                 int32_t pitch, float* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && N >= 1 && C >= 1 && spatial >= 1 && pitch >= C, "nhwc_to_nchw: bad arguments");
  const long long bx = (spatial + 31) / 32;
  B200_CHECK_ARG(bx < (1ll << 31) && N <= 65535, "nhwc_to_nchw: extent too large");
  dim3 grid((unsigned)bx, (C + 31) / 32, N);
  if (x_dtype == B200_DT_H16)
    B200_CUDA(b200::launch_pdl(nhwc_to_nchw_kernel<h16>, grid, dim3(32, 8), 0, stream, reinterpret_cast<const h16*>(x), C, spatial, pitch, y));
  else
    B200_CUDA(b200::launch_pdl(nhwc_to_nchw_kernel<float>, grid, dim3(32, 8), 0, stream, reinterpret_cast<const float*>(x), C, spatial, pitch, y));
  B200_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return B200_OK;
}

extern "C" int b200_upsample_nearest2x(const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t pitch,
                                       int32_t dims, void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && pitch % 8 == 0 && (dims == 2 || dims == 3), "upsample2x: bad arguments");
  const long long total = (long long)N * (dims == 3 ? 2 * D : D) * 2 * H * 2 * W * (pitch / 8);
  B200_CUDA(b200::launch_pdl(upsample2x_kernel, grid_for(total), 256, 0, stream, reinterpret_cast<const uint4*>(x), N, D, H, W, pitch / 8, dims,
                                                        reinterpret_cast<uint4*>(y)));
  B200_LAUNCH_CHECK("upsample2x_kernel");
  return B200_OK;
}

extern "C" int b200_avgpool2(const void* x, int32_t N, int32_t D, int32_t H, int32_t W, int32_t pitch, int32_t dims,
                             void* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && pitch % 8 == 0 && (dims == 2 || dims == 3), "avgpool2: bad arguments");
  const long long total = (long long)N * (dims == 3 ? D / 2 : D) * (H / 2) * (W / 2) * (pitch / 8);
  if (total == 0) return B200_OK;
  B200_CUDA(b200::launch_pdl(avgpool2_kernel, grid_for(total), 256, 0, stream, reinterpret_cast<const uint4*>(x), N, D, H, W, pitch / 8, dims,
                                                      reinterpret_cast<uint4*>(y)));
  B200_LAUNCH_CHECK("avgpool2_kernel");
  return B200_OK;
}

extern "C" int b200_axpy_h16(const void* a, const void* b, float alpha, void* y, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(a && b && y && n % 8 == 0, "axpy_h16: element count must be a multiple of 8");
  if (n == 0) return B200_OK;
  B200_CUDA(b200::launch_pdl(axpy_h16_kernel, grid_for(n / 8), 256, 0, stream, reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
                                                       alpha, reinterpret_cast<uint4*>(y), n / 8));
  B200_LAUNCH_CHECK("axpy_h16_kernel");
  return B200_OK;
}


static bool tap_geom_ok(const b200::TapGeom& g) {
  return g.N >= 1 && g.D >= 1 && g.H >= 1 && g.W >= 1 && g.OD >= 1 && g.OH >= 1 && g.OW >= 1 && g.kd >= 1 &&
         g.kh >= 1 && g.kw >= 1 && g.sd >= 1 && g.sh >= 1 && g.sw >= 1 && g.pd >= 0 && g.ph >= 0 && g.pw >= 0;
}

extern "C" int b200_tap_gather(const void* x, int32_t C, int32_t x_pitch, const int32_t* geom, void* out,
                               int32_t out_pitch, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && out && geom && C >= 1 && x_pitch >= C, "tap_gather: bad arguments");
  b200::TapGeom g;
  memcpy(&g, geom, sizeof(g));
  B200_CHECK_ARG(tap_geom_ok(g) && out_pitch >= g.kd * g.kh * g.kw * C && out_pitch % 8 == 0 &&
                 ((uintptr_t)out % 16) == 0, "tap_gather: bad geometry (out_pitch must be a multiple of 8)");
  const long long total = (long long)g.N * g.OD * g.OH * g.OW * (out_pitch / 8);
  B200_CUDA(b200::launch_pdl(tap_gather_kernel, grid_for(total), 256, 0, stream, reinterpret_cast<const h16*>(x), C, x_pitch, g,
                                                        reinterpret_cast<h16*>(out), out_pitch));
  B200_LAUNCH_CHECK("tap_gather_kernel");
  return B200_OK;
}

extern "C" int b200_tap_sum(const float* y, int32_t y_pitch, const int32_t* geom, int32_t cout, const float* bias,
                            void* out, int32_t out_pitch, int32_t out_dtype, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(y && out && geom && cout >= 1 && cout <= 4 && out_pitch >= cout, "tap_sum: 1 <= cout <= 4");
  b200::TapGeom g;
  memcpy(&g, geom, sizeof(g));
  B200_CHECK_ARG(tap_geom_ok(g) && g.sd == 1 && g.sh == 1 && g.sw == 1 && y_pitch >= g.kd * g.kh * g.kw * cout,
                 "tap_sum: bad geometry (stride must be 1)");
  const long long total = (long long)g.N * g.OD * g.OH * g.OW;
  const unsigned grid = grid_for(total);
  switch (cout) {
    case 1: B200_CUDA(b200::launch_pdl(tap_sum_kernel<1>, grid, 256, 0, stream, y, y_pitch, g, bias, out, out_pitch, out_dtype)); break;
    case 2: B200_CUDA(b200::launch_pdl(tap_sum_kernel<2>, grid, 256, 0, stream, y, y_pitch, g, bias, out, out_pitch, out_dtype)); break;
    case 3: B200_CUDA(b200::launch_pdl(tap_sum_kernel<3>, grid, 256, 0, stream, y, y_pitch, g, bias, out, out_pitch, out_dtype)); break;
    default: B200_CUDA(b200::launch_pdl(tap_sum_kernel<4>, grid, 256, 0, stream, y, y_pitch, g, bias, out, out_pitch, out_dtype)); break;
  }
  B200_LAUNCH_CHECK("tap_sum_kernel");
  return B200_OK;
}


// rows of T new tokens per sequence appended to a [B, L, pitch] key/value cache at the device-side position
__global__ void cache_append_kernel(const h16* __restrict__ src, h16* __restrict__ cache, int B,
                                    int T, int L, int pitch, const int* __restrict__ pos_dev) {
  pdl_entry();
  const int pos = *pos_dev;
  const long long total = (long long)B * T * pitch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % pitch);
    const long long r = i / pitch;
    const int t = (int)(r % T), b = (int)(r / T);
    if (pos + t < L) cache[((long long)b * L + pos + t) * pitch + c] = src[i];
  }
}
__global__ void advance_i32_kernel(int* p, int delta) {
  pdl_entry(); *p += delta; }

extern "C" int b200_cache_append(const void* src, void* cache, int32_t B, int32_t T, int32_t L, int32_t pitch,
                                 const int32_t* pos_dev, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(src && cache && pos_dev && B >= 1 && T >= 1 && L >= T && pitch >= 1, "cache_append: bad arguments");
  B200_CUDA(b200::launch_pdl(cache_append_kernel, grid_for((long long)B * T * pitch), 256, 0, stream, 
      reinterpret_cast<const h16*>(src), reinterpret_cast<h16*>(cache), B, T, L, pitch, pos_dev));
  B200_LAUNCH_CHECK("cache_append_kernel");
  return B200_OK;
}

extern "C" int b200_advance_i32(int32_t* p, int32_t delta, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(p != nullptr, "advance_i32: null pointer");
  B200_CUDA(b200::launch_pdl(advance_i32_kernel, 1, 1, 0, stream, p, delta));
  B200_LAUNCH_CHECK("advance_i32_kernel");
  return B200_OK;
}

extern "C" int b200_embed_tokens(const int64_t* tokens, int64_t M, int32_t seq_len, int32_t pos0, const float* tok_emb,
                                 const float* pos_emb, int32_t C, void* out, int32_t pitch, const int32_t* pos_dev,
                                 void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(tokens && tok_emb && pos_emb && out && M >= 1 && seq_len >= 1 && pos0 >= 0 && C >= 1 && pitch >= C,
                 "embed_tokens: bad arguments");
  B200_CUDA(b200::launch_pdl(embed_tokens_kernel, grid_for(M * pitch), 256, 0, stream, reinterpret_cast<const long long*>(tokens), M, seq_len,
                                                              pos0, tok_emb, pos_emb, C,
                                                              reinterpret_cast<h16*>(out), pitch, pos_dev));
  B200_LAUNCH_CHECK("embed_tokens_kernel");
  return B200_OK;
}

extern "C" int b200_copy_channels(const void* src, int32_t C, int32_t src_pitch, void* dst, int32_t dst_pitch,
                                  int32_t dst_off, int64_t rows, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(src && dst && C >= 1 && src_pitch >= C && dst_pitch >= dst_off + C && rows >= 1, "copy_channels: bad arguments");
  const int vec = (C % 8 == 0 && src_pitch % 8 == 0 && dst_pitch % 8 == 0 && dst_off % 8 == 0 &&
                   ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0)) ? 8 : 1;
  B200_CUDA(b200::launch_pdl(copy_channels_kernel, grid_for(rows * (C / vec)), 256, 0, stream, reinterpret_cast<const h16*>(src), C, src_pitch,
                                                                     reinterpret_cast<h16*>(dst), dst_pitch, dst_off, rows, vec));
  B200_LAUNCH_CHECK("copy_channels_kernel");
  return B200_OK;
}

extern "C" int b200_geglu(const void* x, int64_t M, int32_t H, int32_t x_pitch, void* y, int32_t y_pitch,
                          void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && y && H % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_pitch >= 2 * H && y_pitch >= H,
                 "geglu: bad arguments");
  B200_CUDA(b200::launch_pdl(geglu_kernel, grid_for(M * (H / 8)), 256, 0, stream, reinterpret_cast<const h16*>(x), M, H, x_pitch,
                                                         reinterpret_cast<h16*>(y), y_pitch));
  B200_LAUNCH_CHECK("geglu_kernel");
  return B200_OK;
}

extern "C" int b200_softmax_rows(const float* s, int64_t M, int32_t S, int64_t s_pitch, void* p, int64_t p_pitch,
                                 void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(s && p && M >= 1 && S >= 1 && s_pitch >= S && p_pitch >= S, "softmax_rows: bad arguments");
  h16* pp = reinterpret_cast<h16*>(p);
  if (S <= 1024) {
    const long long blocks = (M + 7) / 8;
    B200_CHECK_ARG(blocks < (1ll << 31), "softmax_rows: too many rows");
    B200_CUDA(b200::launch_pdl(softmax_rows_kernel<32>, (unsigned)blocks, 256, 0, stream, s, M, S, s_pitch, pp, p_pitch));
  } else {
    B200_CHECK_ARG(M < (1ll << 31), "softmax_rows: too many rows");
    B200_CUDA(b200::launch_pdl(softmax_rows_kernel<256>, (unsigned)M, 256, 0, stream, s, M, S, s_pitch, pp, p_pitch));
  }
  B200_LAUNCH_CHECK("softmax_rows_kernel");
  return B200_OK;
}

extern "C" int b200_softmax_rows_partials(const float* s, int64_t M, int32_t S, int64_t s_pitch, const float* partials,
                                          int32_t n_tiles, void* p, int64_t p_pitch, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(s && p && partials && M >= 1 && M < (1ll << 31) && S >= 1 && s_pitch >= S && p_pitch >= S && n_tiles >= 1,
                 "softmax_rows_partials: bad arguments");
  B200_CUDA(b200::launch_pdl(softmax_rows_partials_kernel, (unsigned)M, 256, 0, stream, s, S, s_pitch, reinterpret_cast<const float2*>(partials),
                                                               n_tiles, reinterpret_cast<h16*>(p), p_pitch));
  B200_LAUNCH_CHECK("softmax_rows_partials_kernel");
  return B200_OK;
}

extern "C" int b200_timestep_embedding(const float* t, int32_t N, int32_t dim, float max_period, float* emb,
                                       void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(t && emb && N >= 1 && dim >= 1, "timestep_embedding: bad arguments");
  B200_CUDA(b200::launch_pdl(timestep_embedding_kernel, (N * dim + 255) / 256, 256, 0, stream, t, N, dim, max_period, emb));
  B200_LAUNCH_CHECK("timestep_embedding_kernel");
  return B200_OK;
}

extern "C" int b200_small_linear(const float* x, int32_t M, int32_t K, const float* W, const float* b, int32_t O,
                                 int32_t act_in, int32_t act_out, float* y, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && W && y && M >= 1 && M <= 4096 && K >= 1 && O >= 1, "small_linear: bad arguments");
  const bool vec = (K % 128 == 0) && (((uintptr_t)x | (uintptr_t)W) & 15) == 0;
  if (vec) B200_CUDA(b200::launch_pdl(small_linear_kernel<true>, (O + 7) / 8, 256, 0, stream, x, M, K, W, b, O, act_in, act_out, y));
  else B200_CUDA(b200::launch_pdl(small_linear_kernel<false>, (O + 7) / 8, 256, 0, stream, x, M, K, W, b, O, act_in, act_out, y));
  B200_LAUNCH_CHECK("small_linear_kernel");
  return B200_OK;
}

extern "C" int b200_ddim_step(const float* model_out, const float* sample, const float* noise, const b200_ddim_coef* c,
                              float* prev_sample, float* pred_x0, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(model_out && sample && c && prev_sample && n >= 1, "ddim_step: bad arguments");
  const bool vec = (((uintptr_t)model_out | (uintptr_t)sample | (uintptr_t)noise | (uintptr_t)prev_sample | (uintptr_t)pred_x0) & 15) == 0;
  if (vec) B200_CUDA(b200::launch_pdl(ddim_step_kernel<1>, grid_for((n + 7) / 8, 256, 16), 256, 0, stream, model_out, sample, noise, *c, prev_sample, pred_x0, (long long)n));
  else B200_CUDA(b200::launch_pdl(ddim_step_kernel<0>, grid_for(n), 256, 0, stream, model_out, sample, noise, *c, prev_sample, pred_x0, (long long)n));
  B200_LAUNCH_CHECK("ddim_step_kernel");
  return B200_OK;
}

extern "C" int b200_ddpm_step(const float* model_out, const float* sample, const float* noise, const float* pred_var,
                              const b200_ddpm_coef* c, float* prev_sample, float* pred_x0, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(model_out && sample && c && prev_sample && n >= 1, "ddpm_step: bad arguments");
  B200_CHECK_ARG(c->var_mode == 0 || pred_var, "ddpm_step: learned variance needs pred_var");
  B200_CUDA(b200::launch_pdl(ddpm_step_kernel, grid_for(n), 256, 0, stream, model_out, sample, noise, pred_var, *c, prev_sample, pred_x0, n));
  B200_LAUNCH_CHECK("ddpm_step_kernel");
  return B200_OK;
}

extern "C" int b200_pndm_step(const float* const* hist, const float* sample, const b200_pndm_coef* c, float* prev_sample,
                              float* eps_out, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(hist && c && c->n_hist >= 1 && c->n_hist <= 4 && n >= 1, "pndm_step: bad arguments");
  B200_CHECK_ARG(prev_sample == nullptr || sample != nullptr, "pndm_step: prev_sample needs sample");
  PndmPtrs hp;
  for (int k = 0; k < 4; ++k) hp.h[k] = k < c->n_hist ? hist[k] : nullptr;
  for (int k = 0; k < c->n_hist; ++k) B200_CHECK_ARG(hp.h[k], "pndm_step: null history tensor %d", k);
  B200_CUDA(b200::launch_pdl(pndm_step_kernel, grid_for(n), 256, 0, stream, hp, sample, *c, prev_sample, eps_out, n));
  B200_LAUNCH_CHECK("pndm_step_kernel");
  return B200_OK;
}

extern "C" int b200_add_noise(const float* x0, const float* noise, const float* ca, const float* cb, float sign_b,
                              int32_t N, int64_t per_sample, float* out, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x0 && noise && ca && cb && out && N >= 1 && N <= 65535 && per_sample >= 1, "add_noise: bad arguments");
  dim3 grid(grid_for(per_sample, 256, 4), N);
  B200_CUDA(b200::launch_pdl(add_noise_kernel, grid, 256, 0, stream, x0, noise, ca, cb, sign_b, per_sample, out));
  B200_LAUNCH_CHECK("add_noise_kernel");
  return B200_OK;
}

extern "C" int b200_exp_half_clamped(const float* log_var, float lo, float hi, float* sigma, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(log_var && sigma && n >= 1, "exp_half_clamped: bad arguments");
  B200_CUDA(b200::launch_pdl(exp_half_clamped_kernel, grid_for(n), 256, 0, stream, log_var, lo, hi, sigma, n));
  B200_LAUNCH_CHECK("exp_half_clamped_kernel");
  return B200_OK;
}

extern "C" int b200_fma_f32(const float* a, const float* b, const float* c, float* out, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(a && b && c && out && n >= 1, "fma_f32: bad arguments");
  B200_CUDA(b200::launch_pdl(fma_f32_kernel, grid_for(n), 256, 0, stream, a, b, c, out, n));
  B200_LAUNCH_CHECK("fma_f32_kernel");
  return B200_OK;
}

extern "C" int b200_scale_f32(const float* x, float mul, float div, float* out, int64_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x && out && n >= 1 && div != 0.f, "scale_f32: bad arguments");
  B200_CUDA(b200::launch_pdl(scale_f32_kernel, grid_for(n), 256, 0, stream, x, mul, div, out, n));
  B200_LAUNCH_CHECK("scale_f32_kernel");
  return B200_OK;
}

extern "C" int b200_ddpm_kl(const float* x0, const float* xt, const float* model_out, const b200_kl_coef* c, float* kl_out,
                            double* sample_sum, int32_t N, int64_t per_sample, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  B200_CHECK_ARG(x0 && xt && model_out && c && sample_sum && N >= 1 && N <= 65535 && per_sample >= 1, "ddpm_kl: bad arguments");
  dim3 grid(grid_for(per_sample, 256, 4), N);
  B200_CUDA(b200::launch_pdl(ddpm_kl_kernel, grid, 256, 0, stream, x0, xt, model_out, *c, kl_out, sample_sum, per_sample));
  B200_LAUNCH_CHECK("ddpm_kl_kernel");
  return B200_OK;
}
