// Fused gradient-norm / clip / AdamW over one flat fp32 parameter buffer.
//
// Reference: rnn_trainer.py:259-292 (AdamW, three param groups, eps 0.1), :550-558
// (clip_grad_norm_ max_norm, error_if_nonfinite; optimizer.step).  Parameters whose gradient is
// None in the reference (day layers not sampled in this batch, rnn_model.py:95-96) are skipped
// entirely -- no moment decay, no step increment -- which is reproduced with a per-segment
// `active` flag.  The kernel also refreshes the bf16 shadow copy the tensor-core kernels read.
#pragma once
#include "sm100.cuh"

namespace b2t {

struct Segment {
  long long offset;     // element offset into the flat buffers (multiple of 64)
  long long size;
  int group;            // 0 bias, 1 day, 2 other (rnn_trainer.py:267-269)
  int day;              // day index for day params, -1 otherwise
};

constexpr int OPT_CHUNK = 4096;   // elements per block

struct ChunkRef { int seg; int first; };   // chunk -> segment, first element inside the segment

// sum of squares of the flat gradient buffer (n elements) -> out[0] (atomicAdd, zeroed before)
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
  float s = 0.f;
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += g[i] * g[i];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}

struct AdamParams {
  float* p;                  // flat params
  const float* g;            // flat grads
  float* m;                  // exp_avg
  float* v;                  // exp_avg_sq
  __nv_bfloat16* shadow;     // bf16 mirror of p
  const Segment* segs;
  const ChunkRef* chunks;
  int* step;                 // per-segment step counters (incremented here when active)
  const float* day_touched;  // [n_days] > 0 => day params active (nullable => all active)
  const float* sumsq;        // [1] global gradient sum of squares
  float* stats;              // [0] = total norm, [1] = clip coefficient   (written by block 0)
  float max_norm;            // <= 0 disables clipping
  float lr[3], wd[3];
  float beta1, beta2, eps;
};

__global__ void __launch_bounds__(256)
clip_adamw_kernel(const AdamParams a) {
  const ChunkRef cr = a.chunks[blockIdx.x];
  const Segment sg = a.segs[cr.seg];
  const float total = sqrtf(*a.sumsq);
  float coef = 1.0f;
  if (a.max_norm > 0.f) coef = fminf(1.0f, a.max_norm / (total + 1e-6f));
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.stats[0] = total; a.stats[1] = coef; }
  // clip_grad_norm_(error_if_nonfinite=True) raises before optimizer.step(): a non-finite norm leaves parameters, moments and
  // step counters untouched here, and the host raises on stats[0] (rnn_trainer mirror)
  if (a.max_norm > 0.f && !isfinite(total)) return;
  const bool active = !(sg.day >= 0 && a.day_touched && !(a.day_touched[sg.day] > 0.f));
  if (!active) return;
  const int stp = a.step[cr.seg] + 1;              // every chunk of the segment reads the pre-increment value
  const float lr = a.lr[sg.group], wd = a.wd[sg.group];
  const float bc1 = 1.0f - powf(a.beta1, (float)stp);
  const float bc2s = sqrtf(1.0f - powf(a.beta2, (float)stp));
  const float step_size = lr / bc1;
  const long long base = sg.offset + cr.first;
  const int n = (int)((sg.size - cr.first) < OPT_CHUNK ? (sg.size - cr.first) : OPT_CHUNK);
  const float decay = 1.0f - lr * wd, omb1 = 1.0f - a.beta1, omb2 = 1.0f - a.beta2;
  auto upd = [&](float g, float& p, float& m, float& v) {
    g *= coef;
    p *= decay;
    m = a.beta1 * m + omb1 * g;
    v = a.beta2 * v + omb2 * g * g;
    const float denom = sqrtf(v) / bc2s + a.eps;
    p -= step_size * (m / denom);
  };
  // 16-byte accesses (segment offsets are multiples of 64 elements, chunks of 4096): 4 loads + 3 stores of 16 B and one of 8 B per thread and pass
  const int n4 = n >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const long long k = base + 4ll * i;
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(a.g + k));          // gradients are dead after this pass
    float4 p4 = *reinterpret_cast<const float4*>(a.p + k), m4 = *reinterpret_cast<const float4*>(a.m + k), v4 = *reinterpret_cast<const float4*>(a.v + k);
    upd(g4.x, p4.x, m4.x, v4.x); upd(g4.y, p4.y, m4.y, v4.y); upd(g4.z, p4.z, m4.z, v4.z); upd(g4.w, p4.w, m4.w, v4.w);
    *reinterpret_cast<float4*>(a.p + k) = p4;
    *reinterpret_cast<float4*>(a.m + k) = m4;
    *reinterpret_cast<float4*>(a.v + k) = v4;
    st_bf16x4(a.shadow + k, p4.x, p4.y, p4.z, p4.w);
  }
  for (int i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
    const long long k = base + i;
    float p = a.p[k], m = a.m[k], v = a.v[k];
    upd(a.g[k], p, m, v);
    a.p[k] = p; a.m[k] = m; a.v[k] = v;
    a.shadow[k] = __float2bfloat16_rn(p);
  }
}

// second tiny kernel: bump the step counters of active segments (kept separate so that all chunks of a
// segment observe the same pre-increment value above)
__global__ void bump_steps_kernel(const Segment* segs, int nseg, const float* day_touched, int* step, const float* sumsq, float max_norm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  if (max_norm > 0.f && !isfinite(sqrtf(*sumsq))) return;
  const Segment sg = segs[s];
  const bool active = !(sg.day >= 0 && day_touched && !(day_touched[sg.day] > 0.f));
  if (active) step[s] += 1;
}

// day bookkeeping: touched[day_idx[b]] = 1; zero the grads of touched day segments + all non-day small segments
__global__ void mark_days_kernel(const int* day_idx, int B, float* touched) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) touched[day_idx[b]] = 1.0f;
}

}  // namespace b2t
