// sampling.cuh — Sampler.forward (layers/sampler.py:14-36) and verify() (utils/verify.py:5-181)
// as single launches: vocabulary-parallel partial reductions with warp shuffles, a
// device-wide arrive/ticket to finish, no host control flow and no fp32 [B,K+1,V] temporaries.
//
// RNG: the reference consumes torch's global Philox stream (exponential_ in the sampler,
// rand_like + 2x multinomial in verify).  A fused kernel cannot replay that stream, so the
// draws are keyed explicitly — Philox4x32-10, key = seed, counter = (element index, row,
// call id, stream tag) — and oracle/philox.py reproduces them bit-for-bit.  multinomial(·,1)
// is the exponential race argmax(p_i / E_i) (what torch itself does for one sample), evaluated
// in the log domain: argmax(l_i/T - log E_i); the softmax normaliser is a common positive
// factor and drops out.
#pragma once
#include "common.cuh"

namespace ssdk {

enum { TAG_SAMPLE = 1, TAG_ACCEPT = 2, TAG_RECOVER = 3 };

SSDK_DEVINL uint4 philox_draw(uint32_t idx, uint32_t row, uint64_t call_id, uint32_t tag, uint64_t seed) {
  uint4 ctr = make_uint4(idx, row, (uint32_t)call_id, ((uint32_t)(call_id >> 32) & 0x00FFFFFFu) | (tag << 24));
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return philox4x32_10(ctr, key);
}
SSDK_DEVINL uint32_t u4_word(const uint4& v, int w) { return w == 0 ? v.x : (w == 1 ? v.y : (w == 2 ? v.z : v.w)); }
// uniform in [0,1) like torch.rand
SSDK_DEVINL float u32_to_unit_half_open(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// block-wide argmax (blockDim.x <= 1024); result valid in thread 0
SSDK_DEVINL ArgMax block_argmax(ArgMax a, ArgMax* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_argmax(a);
  __syncthreads();
  if (lane == 0) red[wid] = a;
  __syncthreads();
  if (wid == 0) {
    ArgMax b = (lane < nw) ? red[lane] : ArgMax{-INFINITY, 0x7fffffff};
    a = warp_argmax(b);
  }
  return a;
}

// ----------------------------------------------------------------------------------
// sampler.  grid = (n_chunks, B), 256 threads.
// ----------------------------------------------------------------------------------
struct SampleParams {
  const __nv_bfloat16* logits;
  int64_t ld;
  const float* temps;
  int V;
  uint64_t seed, call_id;
  int64_t* out;
  int out_stride;
  ArgMax* partial;     // [B, n_chunks]
  unsigned* counters;  // [B], zero on entry, zero again on exit
  const uint64_t* dyn; // optional device {seed, step}: lets a static CUDA graph draw fresh numbers
  int sub;             // call_id = step * 16 + sub when dyn != nullptr
};

__global__ void __launch_bounds__(256) sample_kernel(SampleParams p) {
  SSDK_STATIC_SMEM(ArgMax, red, 32);
  SSDK_SHARED_VAR(bool, is_last);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) trace_mark(TR_SAMPLE);
  if (p.dyn) {
    p.seed = p.dyn[0];
    p.call_id = p.dyn[1] * 16ull + (uint64_t)p.sub;
  }
  const int b = blockIdx.y, c = blockIdx.x, nch = gridDim.x;
  const int cs = (((p.V + nch - 1) / nch) + 7) & ~7;
  const int beg = c * cs, end = min(p.V, beg + cs);
  const float T = p.temps[b];
  const bool greedy = (T == 0.f);
  const float invT = greedy ? 1.f : 1.f / T;
  const __nv_bfloat16* row = p.logits + (size_t)b * p.ld;
  ArgMax best{-INFINITY, 0x7fffffff};
  for (int i = beg + threadIdx.x * 4; i < end; i += blockDim.x * 4) {
    uint4 rnd = make_uint4(0, 0, 0, 0);
    if (!greedy) rnd = philox_draw((uint32_t)(i >> 2), (uint32_t)b, p.call_id, TAG_SAMPLE, p.seed);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = i + j;
      if (idx < end) {
        const float l = bf2f(row[idx]);
        float sc = l;
        if (!greedy) {
          // scores = softmax(l/T) / (E + 1e-10)   (sampler.py:27-34), log domain
          const float e = u32_to_exp1(u4_word(rnd, j)) + 1e-10f;
          sc = l * invT - __logf(e);
        }
        best = argmax_better(best, ArgMax{sc, idx});
      }
    }
  }
  best = block_argmax(best, red);
  if (threadIdx.x == 0) {
    p.partial[(size_t)b * nch + c] = best;
    __threadfence();
    const unsigned t = atomicAdd(&p.counters[b], 1u);
    is_last = (t == (unsigned)nch - 1u);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    ArgMax a{-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < nch; i += blockDim.x) {
      ArgMax q;
      q.v = __ldcg(&p.partial[(size_t)b * nch + i].v);
      q.i = __ldcg(&p.partial[(size_t)b * nch + i].i);
      a = argmax_better(a, q);
    }
    a = block_argmax(a, red);
    if (threadIdx.x == 0) {
      p.out[(size_t)b * p.out_stride] = a.i;
      p.counters[b] = 0u;
    }
  }
}

// ----------------------------------------------------------------------------------
// verify.  grid = n_ctas (all co-resident: n_ctas <= #SMs), 128 threads.
//   phase 1  per (row, vocab slice): argmax over bf16 logits (lowest index wins) and, for
//            temp>0 rows, online (max, sum exp) of l/T          -> partials
//   barrier
//   phase 2  every CTA redundantly finishes the row statistics and walks the K draft tokens:
//            greedy prefix (verify.py:29-48) or ratio acceptance u <= min(1, p/(q+1e-10))
//            (verify.py:107-124)
//   phase 3  (rows with target temp>0) recovery draw from max(0,p-q) renormalised, falling
//            back to p (verify.py:137-164)                      -> partials
//   ticket   the last CTA reduces phase-3 partials, writes n_accept / recovery, resets counters.
// ----------------------------------------------------------------------------------
struct RowPart {
  float best_v;
  int best_i;
  float m;
  float s;
};
struct RecPart {
  float adj_v;
  int adj_i;
  float p_v;
  int p_i;
};
struct VerifyParams {
  const __nv_bfloat16* lp;  // [B, K+1, V]
  const __nv_bfloat16* lq;  // [B, K, V]
  const int64_t* spec;      // [B, K+1]
  const float* temps_t;
  const float* temps_q;
  const int32_t* cache_hits;  // may be null
  int jit, B, K, V;
  uint64_t seed, call_id;
  int32_t* n_accept;   // [B]
  int64_t* recovery;   // [B]
  RowPart* row_part;   // [B*(2K+1), n_ctas]
  RecPart* rec_part;   // [B, n_ctas]
  unsigned* counters;  // [2]: barrier, ticket (zero on entry/exit)
  const uint64_t* dyn; // optional device {seed, step}
  int sub;
};

constexpr int kVerifyThreads = 128;
constexpr int kVerifyMaxBatch = 32;                    // sequences per verify launch
constexpr int kVerifyMaxRows = kVerifyMaxBatch * 15;  // B*(2K+1) bound for the smem row table (K <= 7)

SSDK_DEVINL void online_merge(float& m, float& s, float m2, float s2) {
  if (m2 == -INFINITY) return;
  if (m == -INFINITY) {
    m = m2;
    s = s2;
    return;
  }
  const float mm = fmaxf(m, m2);
  s = s * __expf(m - mm) + s2 * __expf(m2 - mm);
  m = mm;
}

__global__ void __launch_bounds__(kVerifyThreads) verify_kernel(VerifyParams p) {
  SSDK_STATIC_SMEM(ArgMax, red, 32);
  SSDK_STATIC_SMEM(float, redm, 4);
  SSDK_STATIC_SMEM(float, reds, 4);
  SSDK_STATIC_SMEM(int, row_arg, kVerifyMaxRows);
  SSDK_STATIC_SMEM(float, row_m, kVerifyMaxRows);
  SSDK_STATIC_SMEM(float, row_z, kVerifyMaxRows);
  SSDK_STATIC_SMEM(int, s_n, kVerifyMaxBatch);      // accepted count
  SSDK_STATIC_SMEM(int, s_flags, kVerifyMaxBatch);  // bit0 = needs recovery draw, bit1 = adjust
  SSDK_STATIC_SMEM(long long, s_rec_greedy, kVerifyMaxBatch);
  SSDK_SHARED_VAR(bool, is_last);
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) trace_mark(TR_VERIFY);
  if (p.dyn) {
    p.seed = p.dyn[0];
    p.call_id = p.dyn[1] * 16ull + (uint64_t)p.sub;
  }

  const int B = p.B, K = p.K, V = p.V, nct = gridDim.x, c = blockIdx.x;
  const int n_prow = B * (K + 1), n_rows = n_prow + B * K;
  const int cs = (((V + nct - 1) / nct) + 7) & ~7;
  const int beg = min(V, c * cs), end = min(V, beg + cs);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  // ---------------- phase 1 ----------------
  for (int r = 0; r < n_rows; ++r) {
    const bool is_p = r < n_prow;
    const int b = is_p ? r / (K + 1) : (r - n_prow) / K;
    const __nv_bfloat16* row = is_p ? p.lp + (size_t)r * V : p.lq + (size_t)(r - n_prow) * V;
    const float T = is_p ? p.temps_t[b] : p.temps_q[b];
    const bool soft = T > 0.f;
    const float invT = soft ? 1.f / fmaxf(T, 1e-8f) : 1.f;
    ArgMax best{-INFINITY, 0x7fffffff};
    float m = -INFINITY, s = 0.f;
    for (int i = beg + threadIdx.x * 8; i < end; i += kVerifyThreads * 8) {
      float f[8];
      if (i + 8 <= end && ((reinterpret_cast<uintptr_t>(row + i) & 15) == 0)) {
        unpack_bf16x8(*reinterpret_cast<const uint4*>(row + i), f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (i + j < end) ? bf2f(row[i + j]) : -INFINITY;
      }
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        best = argmax_better(best, ArgMax{f[j], i + j});
        cm = fmaxf(cm, f[j]);
      }
      if (soft && cm != -INFINITY) {
        const float cms = cm * invT;
        float cs8 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) cs8 += __expf(f[j] * invT - cms);
        online_merge(m, s, cms, cs8);
      }
    }
    // block reduce
    best = block_argmax(best, red);
    if (soft) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        online_merge(m, s, m2, s2);
      }
      if (lane == 0) {
        redm[wid] = m;
        reds[wid] = s;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        m = redm[0];
        s = reds[0];
        for (int w = 1; w < kVerifyThreads / 32; ++w) online_merge(m, s, redm[w], reds[w]);
      }
    }
    if (threadIdx.x == 0) {
      RowPart rp;
      rp.best_v = best.v;
      rp.best_i = best.i;
      rp.m = m;
      rp.s = s;
      p.row_part[(size_t)r * nct + c] = rp;
    }
    __syncthreads();
  }

  // ---------------- grid barrier ----------------
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&p.counters[0], 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(&p.counters[0]) < (unsigned)nct) {
      if (clock64() - t0 > 8000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();

  // ---------------- phase 2: finish row statistics (warp per row) ----------------
  for (int r = wid; r < n_rows; r += kVerifyThreads / 32) {
    ArgMax best{-INFINITY, 0x7fffffff};
    float m = -INFINITY, s = 0.f;
    for (int i = lane; i < nct; i += 32) {
      const RowPart* rp = &p.row_part[(size_t)r * nct + i];
      best = argmax_better(best, ArgMax{__ldcg(&rp->best_v), __ldcg(&rp->best_i)});
      online_merge(m, s, __ldcg(&rp->m), __ldcg(&rp->s));
    }
    best = warp_argmax(best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      online_merge(m, s, m2, s2);
    }
    if (lane == 0) {
      row_arg[r] = best.i;
      row_m[r] = m;
      row_z[r] = s;
    }
  }
  __syncthreads();

  if (threadIdx.x < B) {
    const int b = threadIdx.x;
    const float Tt = p.temps_t[b], Tq = p.temps_q[b];
    const bool hit = p.jit || (p.cache_hits && p.cache_hits[b] != 0);
    const bool ratio = ((Tt > 0.f) || (Tq > 0.f)) && hit;
    const int64_t* sp = p.spec + (size_t)b * (K + 1);
    int ng = K;
    for (int j = 0; j < K; ++j)
      if (sp[j + 1] != (int64_t)row_arg[b * (K + 1) + j]) {
        ng = j;
        break;
      }
    int n = ng;
    if (ratio) {
      n = K;
      for (int j = 0; j < K; ++j) {
        const int x = (int)sp[j + 1];
        const int rp = b * (K + 1) + j, rq = n_prow + b * K + j;
        float pv, qv;
        if (Tt > 0.f) pv = __expf(bf2f(p.lp[(size_t)rp * V + x]) / fmaxf(Tt, 1e-8f) - row_m[rp]) / row_z[rp];
        else pv = (x == row_arg[rp]) ? 1.f : 0.f;
        if (Tq > 0.f) qv = __expf(bf2f(p.lq[(size_t)(rq - n_prow) * V + x]) / fmaxf(Tq, 1e-8f) - row_m[rq]) / row_z[rq];
        else qv = (x == row_arg[rq]) ? 1.f : 0.f;
        const float a = fminf(pv / (qv + 1e-10f), 1.f);
        const float u = u32_to_unit_half_open(philox_draw((uint32_t)j, (uint32_t)b, p.call_id, TAG_ACCEPT, p.seed).x);
        if (!(u <= a)) {
          n = j;
          break;
        }
      }
    }
    s_n[b] = n;
    s_rec_greedy[b] = row_arg[b * (K + 1) + ng];  // verify.py:48 uses the GREEDY count here
    int fl = 0;
    if (Tt > 0.f) fl |= 1;
    if (Tt > 0.f && ratio && n < K) fl |= 2;
    s_flags[b] = fl;
  }
  __syncthreads();

  // ---------------- phase 3: recovery draw for target-temp>0 rows ----------------
  for (int b = 0; b < B; ++b) {
    if (!(s_flags[b] & 1)) continue;
    const bool adjust = (s_flags[b] & 2) != 0;
    const int n = s_n[b];
    const int rp = b * (K + 1) + n;
    const int jq = min(n, K - 1);
    const int rq = n_prow + b * K + jq;
    const float Tt = p.temps_t[b], Tq = p.temps_q[b];
    const float invTt = 1.f / fmaxf(Tt, 1e-8f), invTq = (Tq > 0.f) ? 1.f / fmaxf(Tq, 1e-8f) : 0.f;
    const __nv_bfloat16* prow = p.lp + (size_t)rp * V;
    const __nv_bfloat16* qrow = p.lq + (size_t)(b * K + jq) * V;
    const float mp = row_m[rp], zp = row_z[rp], mq = row_m[rq], zq = row_z[rq];
    const int aq = row_arg[rq];
    ArgMax badj{-INFINITY, 0x7fffffff}, bp{-INFINITY, 0x7fffffff};
    for (int i = beg + threadIdx.x; i < end; i += kVerifyThreads) {
      const uint4 rnd = philox_draw((uint32_t)i, (uint32_t)b, p.call_id, TAG_RECOVER, p.seed);
      const float lt = bf2f(prow[i]) * invTt;
      // fallback: multinomial(p)  (verify.py:139-141,159)
      bp = argmax_better(bp, ArgMax{lt - __logf(u32_to_exp1(rnd.y)), i});
      if (adjust) {
        const float pi = __expf(lt - mp) / zp;
        const float qi = (Tq > 0.f) ? __expf(bf2f(qrow[i]) * invTq - mq) / zq : (i == aq ? 1.f : 0.f);
        const float adj = pi - qi;  // clamp(min=0): non-positive entries can never be drawn
        if (adj > 0.f) badj = argmax_better(badj, ArgMax{__logf(adj) - __logf(u32_to_exp1(rnd.x)), i});
      }
    }
    badj = block_argmax(badj, red);
    __syncthreads();
    bp = block_argmax(bp, red);
    if (threadIdx.x == 0) {
      RecPart r;
      r.adj_v = badj.v;
      r.adj_i = badj.i;
      r.p_v = bp.v;
      r.p_i = bp.i;
      p.rec_part[(size_t)b * nct + c] = r;
    }
    __syncthreads();
  }

  // ---------------- ticket: last CTA finalises ----------------
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(&p.counters[1], 1u);
    is_last = (t == (unsigned)nct - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int b = wid; b < B; b += kVerifyThreads / 32) {
    long long rec = s_rec_greedy[b];
    if (s_flags[b] & 1) {
      ArgMax badj{-INFINITY, 0x7fffffff}, bp{-INFINITY, 0x7fffffff};
      for (int i = lane; i < nct; i += 32) {
        const RecPart* r = &p.rec_part[(size_t)b * nct + i];
        badj = argmax_better(badj, ArgMax{__ldcg(&r->adj_v), __ldcg(&r->adj_i)});
        bp = argmax_better(bp, ArgMax{__ldcg(&r->p_v), __ldcg(&r->p_i)});
      }
      badj = warp_argmax(badj);
      bp = warp_argmax(bp);
      // sums > 0 ? adj/sums : fallbackDist   (verify.py:155)
      rec = ((s_flags[b] & 2) && badj.v != -INFINITY) ? badj.i : bp.i;
    }
    if (lane == 0) {
      p.n_accept[b] = s_n[b];
      p.recovery[b] = rec;
    }
  }
  if (threadIdx.x == 0) {
    p.counters[0] = 0u;
    p.counters[1] = 0u;
  }
}

}  // namespace ssdk
