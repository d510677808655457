// draft_persistent.cuh — EXPERIMENTAL (compiles, NOT yet validated on hardware; off unless SSDK_DRAFT_PERSISTENT=1).
//
// One persistent kernel per draft decode forward (batch 1, one token): the nine kernels per layer of the regular path
//  cost ~2.5 us of boundary each, which is what bounds the 1B draft (0.79 ms per forward against 0.38 ms of
// weight streaming, profiles/r01_small_kernels.md).  Here one CTA per SM stays resident for the whole forward and the
// layer is five phases separated by a device-wide barrier:
//
//   A  [residual add + input RMSNorm, recomputed by every CTA] -> q|k|v GEMV           (LlamaDecoderLayer.forward,
//   B  q/k head norm + RoPE + KV store + split-KV attention for (kv head, split) units   models/llama3.py:248-273;
//   C  [merge of the split partials, every CTA] -> o-proj GEMV                           qwen3.py:252-262)
//   D  [residual add + post-attention RMSNorm, every CTA] -> gate|up GEMV + SiLU*mul
//   E  down-proj GEMV
//
// and the final norm + lm_head GEMV after the last layer.  At one token the projections are GEMVs: CUDA cores stream the
// weights at the HBM roofline (1 FMA per bf16 weight element), so no tensor-core pipeline is needed; each warp owns
// whole output rows (two in flight, 16-byte loads, 512 B per warp instruction) and rows are dealt round-robin over the
// CTAs so every SM streams the same number of bytes.  The small vectors between phases live in L2 (read with ld.cg —
// L1 is not coherent across the barrier); the element-wise work is folded into the consumer phase's prologue and simply
// recomputed by every CTA, which removes the norm / RoPE / combine kernels and their boundaries altogether.
// Rounding points are the reference's: every linear output, the residual, the norm output, q/k after RoPE and the
// attention output are rounded to bf16 (SURVEY §8a checklist 1-4); accumulation is fp32.
#pragma once
#include "common.cuh"

namespace ssdk {

constexpr int kDpThreads = 256;
constexpr int kDpWarps = kDpThreads / 32;
constexpr int kDpMaxLayers = 32;
constexpr int kDpSplits = 8;  // KV splits per kv head in phase B

struct DpLayer {
  const __nv_bfloat16 *qkv, *o, *gate_up, *down, *in_norm, *post_norm, *q_norm, *k_norm;
};

struct DpParams {
  int d, L, H, KV, ffn, vocab, qk_norm;
  float eps, scale_log2;
  const __nv_bfloat16 *embed, *final_norm, *lm_head;
  const float* rope;  // [max_pos, hd]: cos | sin
  __nv_bfloat16 *k_cache, *v_cache;
  long long cache_layer_stride;  // elements between layers
  int block_size, max_blocks;
  const int64_t* token;          // input token id
  const int32_t* ctx0;           // tokens in the cache before this step's first forward
  int pos_offset;                // index of this forward inside the step
  const int32_t* block_table;    // [max_blocks]
  __nv_bfloat16 *vec_qkv, *vec_o, *vec_act, *vec_down, *resid0, *resid1;
  float* attn_part;              // [H][kDpSplits][hd + 2]: o | m | l
  __nv_bfloat16* logits;         // [vocab]; nullptr: no lm_head (the step's last draft forward only writes KV)
  unsigned *bar_counter, *launch_count;
  DpLayer layers[kDpMaxLayers];
};

// ---------------------------------------------------------------------------------------------
// device-wide barrier: monotonically increasing arrival counter, never reset (signed distance copes with wrap-around);
// every launch performs exactly 5 * L barriers, so a launch's first target follows from the launch counter.
// ---------------------------------------------------------------------------------------------
struct DpGridBar {
  unsigned* counter;
  unsigned target;
  __device__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      const long long t0 = clock64();
      do {
        v = ld_acquire_u32(counter);
        if (clock64() - t0 > 4000000000LL) __trap();  // a CTA never arrived: fail loudly instead of hanging the GPU
      } while ((int)(v - target) < 0);
    }
    __syncthreads();
  }
};

SSDK_DEVINL uint4 dp_ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
SSDK_DEVINL float2 dp_bf2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

// xs[i] = bf16r( r_i * rsqrt(mean r^2 + eps) * w_i ),  r = a (+ b) in fp32;  resid_out = bf16(r) (written by CTA 0 only).
// a / b are L2-resident vectors produced by earlier phases.  d is a multiple of 8.
SSDK_DEVINL void dp_norm_prologue(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* resid_out,
                                  const __nv_bfloat16* w, float eps, int d, float* xs, float* red) {
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += kDpThreads * 8) {
    float x[8];
    unpack_bf16x8(dp_ldcg16(a + i), x);
    if (b) {
      float y[8];
      unpack_bf16x8(dp_ldcg16(b + i), y);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += y[j];
    }
    if (resid_out && blockIdx.x == 0) *reinterpret_cast<uint4*>(resid_out + i) = pack_bf16x8(x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xs[i + j] = x[j];
      ss += x[j] * x[j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / (float)d + eps);
  for (int i = threadIdx.x * 8; i < d; i += kDpThreads * 8) {
    float wv[8];
    unpack_bf16x8(*reinterpret_cast<const uint4*>(w + i), wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) xs[i + j] = bf16_round(xs[i + j] * rstd * wv[j]);
  }
  __syncthreads();
}

// one row: partial dot over this lane's 16-byte columns (K multiple of 256), two accumulators to shorten the chain
SSDK_DEVINL void dp_fma8(float& a0, float& a1, const uint4& w, const float4& xa, const float4& xb) {
  float2 f = dp_bf2(w.x);
  a0 = fmaf(f.x, xa.x, a0); a1 = fmaf(f.y, xa.y, a1);
  f = dp_bf2(w.y);
  a0 = fmaf(f.x, xa.z, a0); a1 = fmaf(f.y, xa.w, a1);
  f = dp_bf2(w.z);
  a0 = fmaf(f.x, xb.x, a0); a1 = fmaf(f.y, xb.y, a1);
  f = dp_bf2(w.w);
  a0 = fmaf(f.x, xb.z, a0); a1 = fmaf(f.y, xb.w, a1);
}

// dot(W[r0], x) and dot(W[r1], x) with both rows' loads in flight (r1 < 0: only r0).  xs: fp32 x in shared memory.
SSDK_DEVINL void dp_dot2(const __nv_bfloat16* W, int K, int r0, int r1, const float* xs, int lane, float& y0, float& y1) {
  const uint4* p0 = reinterpret_cast<const uint4*>(W + (size_t)r0 * K) + lane;
  const uint4* p1 = reinterpret_cast<const uint4*>(W + (size_t)(r1 >= 0 ? r1 : r0) * K) + lane;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  const int steps = K >> 8;  // 256 elements (32 lanes x 8) per step
  for (int j0 = 0; j0 < steps; j0 += 8) {  // 16 x 16 B per lane in flight: 8 KB per warp, 64 KB per SM
    uint4 w0[8], w1[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool in = j0 + u < steps;
      w0[u] = in ? ld_nc_v4_evict_first(p0 + (size_t)(j0 + u) * 32) : make_uint4(0, 0, 0, 0);
      w1[u] = (in && r1 >= 0) ? ld_nc_v4_evict_first(p1 + (size_t)(j0 + u) * 32) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (j0 + u < steps) {
        const float4 xa = *reinterpret_cast<const float4*>(xs + (j0 + u) * 256 + lane * 8);
        const float4 xb = *reinterpret_cast<const float4*>(xs + (j0 + u) * 256 + lane * 8 + 4);
        dp_fma8(a0, a1, w0[u], xa, xb);
        dp_fma8(b0, b1, w1[u], xa, xb);
      }
    }
  }
  y0 = warp_sum(a0 + a1);
  y1 = warp_sum(b0 + b1);
}

// The weights do not depend on the activations, so HBM need not idle while the CTAs meet at a barrier or run a
// phase prologue: at the START of a phase every warp asks L2 (cp.async.bulk.prefetch.L2 — fire and forget, no data
// returns to the SM) for all the weight rows it will consume in a LATER phase.  The GEMV of that phase then streams from
// L2 while HBM is already fetching the phase after it; 126 MB of L2 hold more than a whole 1B layer (121 MB).
// Units = rows, or gate|up pairs (i, i + pair_offset), dealt like dp_gemv_rows / dp_gemv_gate_up deal them:
// unit = warp * #CTAs + CTA + j * #warps.
SSDK_DEVINL void dp_bulk_prefetch_l2(const void* p, uint32_t bytes) {
#ifndef SSDK_HOST_EMU
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#else
  (void)p; (void)bytes;
#endif
}
SSDK_DEVINL void dp_prefetch_next(const __nv_bfloat16* W, int K, int n_units, int pair_offset, int max_units_per_warp = 32) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0) return;  // the instruction takes warp-uniform operands: one lane issues the warp's whole list
  const int nw = kDpWarps * (int)gridDim.x;
  int u = warp * (int)gridDim.x + (int)blockIdx.x;
  for (int j = 0; j < max_units_per_warp && u < n_units; ++j, u += nw) {
    dp_bulk_prefetch_l2(W + (size_t)u * K, (uint32_t)K * 2u);
    if (pair_offset) dp_bulk_prefetch_l2(W + (size_t)(u + pair_offset) * K, (uint32_t)K * 2u);
  }
}

// y[r] = bf16(W[r] . x) for r < n_rows; rows dealt warp-major over the CTAs so every SM streams the same bytes
SSDK_DEVINL void dp_gemv_rows(const __nv_bfloat16* W, int K, int n_rows, const float* xs, __nv_bfloat16* y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = kDpWarps * gridDim.x;
  for (int r0 = warp * gridDim.x + blockIdx.x; r0 < n_rows; r0 += 2 * nw) {
    const int r1 = (r0 + nw < n_rows) ? r0 + nw : -1;
    float y0, y1;
    dp_dot2(W, K, r0, r1, xs, lane, y0, y1);
    if (lane == 0) {
      y[r0] = f2bf(y0);
      if (r1 >= 0) y[r1] = f2bf(y1);
    }
  }
}

// act[i] = bf16( silu(bf16(Wg[i].x)) * bf16(Wu[i].x) ),  gate rows [0, ffn), up rows [ffn, 2 ffn)
SSDK_DEVINL void dp_gemv_gate_up(const __nv_bfloat16* W, int K, int ffn, const float* xs, __nv_bfloat16* act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = kDpWarps * gridDim.x;
  for (int i = warp * gridDim.x + blockIdx.x; i < ffn; i += nw) {
    float g, u;
    dp_dot2(W, K, i, ffn + i, xs, lane, g, u);
    if (lane == 0) {
      g = bf16_round(g);
      u = bf16_round(u);
      act[i] = f2bf((g / (1.0f + __expf(-g))) * u);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// phase B unit: kv head h, split s.  Rebuilds the rotated q rows of the head group and the new token's k / v from the
// q|k|v vector, stores k / v into the page slot (split 0 only), runs the online-softmax over its token range (the new
// token comes from shared memory, never from the cache) and writes (o, m, l) per query head.
// ---------------------------------------------------------------------------------------------
template <int HD, int GMAX>
SSDK_DEVINL void dp_attention_unit(const DpParams& p, int layer, int h, int s, int ctx, float* sm) {
  constexpr int HALF = HD / 2;
  constexpr int EPL = HD / 32;  // elements per lane in the dot layout (dims lane*EPL ..)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = p.H / p.KV;
  const int pos = ctx - 1;
  float* sq = sm;                      // [G][HD] rotated q (bf16-rounded values)
  float* sk = sq + GMAX * HD;          // [HD] new k
  float* sv = sk + HD;                 // [HD] new v
  float* sred = sv + HD;               // [kDpWarps][G][HD + 2] per-warp partials

  // ---- q rows (warps 0..G-1), k (warp G... may wrap), v: one warp per row, rotate-half pairs (i, i + HALF) ----
  const float* cs = p.rope + (size_t)pos * HD;
  for (int row = warp; row < G + 2; row += kDpWarps) {
    const int kind = row < G ? 0 : (row == G ? 1 : 2);
    const int col0 = (kind == 0 ? (h * G + row) : (kind == 1 ? p.H + h : p.H + p.KV + h)) * HD;
    float x1[(HALF + 31) / 32], x2[(HALF + 31) / 32];
    float ss = 0.f;
    // plain L2 loads (the vector was written by other SMs in phase A)
#pragma unroll
    for (int t = 0; t < (HALF + 31) / 32; ++t) {
      const int i = lane + 32 * t;
      x1[t] = x2[t] = 0.f;
      if (i < HALF) {
        const unsigned short a = __ldcg(reinterpret_cast<const unsigned short*>(p.vec_qkv + col0 + i));
        const unsigned short b = __ldcg(reinterpret_cast<const unsigned short*>(p.vec_qkv + col0 + HALF + i));
        x1[t] = __bfloat162float(__ushort_as_bfloat16(a));
        x2[t] = __bfloat162float(__ushort_as_bfloat16(b));
        ss += x1[t] * x1[t] + x2[t] * x2[t];
      }
    }
    float* dst = kind == 0 ? sq + row * HD : (kind == 1 ? sk : sv);
    if (kind == 2) {
#pragma unroll
      for (int t = 0; t < (HALF + 31) / 32; ++t) {
        const int i = lane + 32 * t;
        if (i < HALF) {
          dst[i] = x1[t];
          dst[HALF + i] = x2[t];
        }
      }
      continue;
    }
    const __nv_bfloat16* nw = p.qk_norm ? (kind == 0 ? p.layers[layer].q_norm : p.layers[layer].k_norm) : nullptr;
    if (nw) {
      ss = warp_sum(ss);
      const float rstd = rsqrtf(ss / (float)HD + p.eps);
#pragma unroll
      for (int t = 0; t < (HALF + 31) / 32; ++t) {
        const int i = lane + 32 * t;
        if (i < HALF) {
          x1[t] = bf16_round(x1[t] * rstd * bf2f(nw[i]));
          x2[t] = bf16_round(x2[t] * rstd * bf2f(nw[HALF + i]));
        }
      }
    }
#pragma unroll
    for (int t = 0; t < (HALF + 31) / 32; ++t) {
      const int i = lane + 32 * t;
      if (i < HALF) {
        const float c = cs[i], sn = cs[HALF + i];
        dst[i] = bf16_round(x1[t] * c - x2[t] * sn);
        dst[HALF + i] = bf16_round(x2[t] * c + x1[t] * sn);
      }
    }
  }
  __syncthreads();

  // ---- KV store of the new token (one unit per kv head) ----
  const int blk_new = p.block_table[pos / p.block_size];
  if (s == 0 && blk_new >= 0) {
    const size_t slot = (size_t)blk_new * p.block_size + pos % p.block_size;
    __nv_bfloat16* kc = p.k_cache + (size_t)layer * p.cache_layer_stride + (slot * p.KV + h) * HD;
    __nv_bfloat16* vc = p.v_cache + (size_t)layer * p.cache_layer_stride + (slot * p.KV + h) * HD;
    for (int i = threadIdx.x; i < HD; i += kDpThreads) {
      kc[i] = f2bf(sk[i]);
      vc[i] = f2bf(sv[i]);
    }
  }

  // ---- token range of this split ----
  const int per = (ctx + kDpSplits - 1) / kDpSplits;
  const int t0 = s * per, t1 = min(ctx, t0 + per);

  // dot layout: lane owns dims [lane * EPL, lane * EPL + EPL)
  float qreg[GMAX][EPL];
#pragma unroll
  for (int g = 0; g < GMAX; ++g)
#pragma unroll
    for (int e = 0; e < EPL; ++e) qreg[g][e] = (g < G) ? sq[g * HD + lane * EPL + e] : 0.f;
  float m[GMAX], l[GMAX], acc[GMAX][EPL];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    m[g] = -INFINITY;
    l[g] = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[g][e] = 0.f;
  }
  const __nv_bfloat16* kbase = p.k_cache + (size_t)layer * p.cache_layer_stride;
  const __nv_bfloat16* vbase = p.v_cache + (size_t)layer * p.cache_layer_stride;
  for (int t = t0 + warp; t < t1; t += kDpWarps) {
    float kv[EPL], vv[EPL];
    if (t == pos) {
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        kv[e] = sk[lane * EPL + e];
        vv[e] = sv[lane * EPL + e];
      }
    } else {
      const int blk = p.block_table[t / p.block_size];
      const size_t off = (((size_t)blk * p.block_size + t % p.block_size) * p.KV + h) * HD + lane * EPL;
      if constexpr (EPL == 2) {
        const float2 a = dp_bf2(__ldcg(reinterpret_cast<const uint32_t*>(kbase + off)));
        const float2 b = dp_bf2(__ldcg(reinterpret_cast<const uint32_t*>(vbase + off)));
        kv[0] = a.x; kv[1] = a.y; vv[0] = b.x; vv[1] = b.y;
      } else {
        static_assert(EPL == 2 || EPL == 4, "head_dim 64 or 128");
        const uint2 a = __ldcg(reinterpret_cast<const uint2*>(kbase + off));
        const uint2 b = __ldcg(reinterpret_cast<const uint2*>(vbase + off));
        float2 f = dp_bf2(a.x); kv[0] = f.x; kv[1] = f.y;
        f = dp_bf2(a.y); kv[EPL - 2] = f.x; kv[EPL - 1] = f.y;
        f = dp_bf2(b.x); vv[0] = f.x; vv[1] = f.y;
        f = dp_bf2(b.y); vv[EPL - 2] = f.x; vv[EPL - 1] = f.y;
      }
    }
#pragma unroll
    for (int g = 0; g < GMAX; ++g) {
      if (g < G) {
        float sc = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) sc = fmaf(qreg[g][e], kv[e], sc);
        sc = warp_sum(sc) * p.scale_log2;
        const float mn = fmaxf(m[g], sc);
        const float corr = exp2f(m[g] - mn);  // m = -inf -> 0
        const float pr = exp2f(sc - mn);
        l[g] = l[g] * corr + pr;
#pragma unroll
        for (int e = 0; e < EPL; ++e) acc[g][e] = acc[g][e] * corr + pr * vv[e];
        m[g] = mn;
      }
    }
  }
  // ---- merge the warps through shared memory, one thread per (head, dim) ----
  constexpr int LDR = HD + 2;
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    if (g < G) {
      float* w = sred + ((size_t)warp * G + g) * LDR;
#pragma unroll
      for (int e = 0; e < EPL; ++e) w[lane * EPL + e] = acc[g][e];
      if (lane == 0) {
        w[HD] = m[g];
        w[HD + 1] = l[g];
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * HD; idx += kDpThreads) {
    const int g = idx / HD, dim = idx - g * HD;
    float mx = -INFINITY;
    for (int w = 0; w < kDpWarps; ++w) mx = fmaxf(mx, sred[((size_t)w * G + g) * LDR + HD]);
    float o = 0.f, ll = 0.f;
    if (mx != -INFINITY) {
      for (int w = 0; w < kDpWarps; ++w) {
        const float* r = sred + ((size_t)w * G + g) * LDR;
        const float wt = exp2f(r[HD] - mx);
        o += r[dim] * wt;
        ll += r[HD + 1] * wt;
      }
    }
    float* out = p.attn_part + ((size_t)(h * G + g) * kDpSplits + s) * LDR;
    out[dim] = o;  // un-normalised: sum_t 2^(s_t - mx) v_t
    if (dim == 0) {
      out[HD] = mx;
      out[HD + 1] = ll;
    }
  }
  __syncthreads();
}

// phase C prologue: attention output of every head from the split partials -> xs (bf16-rounded), recomputed per CTA
template <int HD>
SSDK_DEVINL void dp_combine_prologue(const DpParams& p, float* xs) {
  constexpr int LDR = HD + 2;
  for (int idx = threadIdx.x; idx < p.H * HD; idx += kDpThreads) {
    const int head = idx / HD, dim = idx - head * HD;
    const float* base = p.attn_part + (size_t)head * kDpSplits * LDR;
    float ms[kDpSplits], ls[kDpSplits], os[kDpSplits];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < kDpSplits; ++s) {
      ms[s] = __ldcg(base + s * LDR + HD);
      ls[s] = __ldcg(base + s * LDR + HD + 1);
      os[s] = __ldcg(base + s * LDR + dim);
      mx = fmaxf(mx, ms[s]);
    }
    float o = 0.f, l = 0.f;
#pragma unroll
    for (int s = 0; s < kDpSplits; ++s) {
      const float wt = (ms[s] == -INFINITY) ? 0.f : exp2f(ms[s] - mx);
      o += os[s] * wt;
      l += ls[s] * wt;
    }
    xs[idx] = bf16_round(l > 0.f ? o / l : 0.f);
  }
  __syncthreads();
}

template <int HD, int GMAX>
__global__ void __launch_bounds__(kDpThreads, 1) draft_forward_persistent_kernel(const __grid_constant__ DpParams p) {
  SSDK_DYN_SMEM(float, dp_smem);
  float* xs = dp_smem;                         // [max(d, ffn, H * HD)] fp32 operand vector of the running phase
  float* scratch = xs + max(max(p.d, p.ffn), p.H * HD);  // attention scratch: q, k, v, per-warp partials
  SSDK_STATIC_SMEM(float, red, 32);
  if (threadIdx.x == 0) trace_mark(TR_MISC);

  DpGridBar bar;
  bar.counter = p.bar_counter;
  bar.target = __ldcg(p.launch_count) * (unsigned)(5 * p.L) * gridDim.x;

  const int ctx = p.ctx0[0] + p.pos_offset + 1;  // tokens visible to this forward, the new one included
  const long long tok = p.token[0];
  const __nv_bfloat16* emb = p.embed + (size_t)tok * p.d;
  __nv_bfloat16* resid[2] = {p.resid0, p.resid1};
  int cur = 0;  // resid[cur] holds the residual entering the layer (layer 0: the embedding row itself)

  const int qkv_rows = (p.H + 2 * p.KV) * HD;
  dp_prefetch_next(p.layers[0].qkv, p.d, qkv_rows, 0);
  dp_prefetch_next(p.layers[0].o, p.H * HD, p.d, 0);
  for (int l = 0; l < p.L; ++l) {
    const DpLayer& lw = p.layers[l];
    // ---- A: (add +) input norm -> q|k|v ----      (HBM meanwhile: this layer's gate|up, 55 % of the layer's bytes)
    dp_prefetch_next(lw.gate_up, p.d, p.ffn, p.ffn);
    if (l == 0) {
      // first layer: hidden = norm(embed), residual = embed (models/llama3.py:192-193)
      dp_norm_prologue(emb, nullptr, resid[cur ^ 1], lw.in_norm, p.eps, p.d, xs, red);
    } else {
      dp_norm_prologue(p.vec_down, resid[cur], resid[cur ^ 1], lw.in_norm, p.eps, p.d, xs, red);
    }
    cur ^= 1;
    dp_gemv_rows(lw.qkv, p.d, qkv_rows, xs, p.vec_qkv);
    bar.sync();
    // ---- B: RoPE + KV store + attention units ----
    for (int u = blockIdx.x; u < p.KV * kDpSplits; u += gridDim.x)
      dp_attention_unit<HD, GMAX>(p, l, u / kDpSplits, u % kDpSplits, ctx, scratch);
    bar.sync();
    // ---- C: merge splits -> o-proj ----            (HBM meanwhile: this layer's down-proj)
    dp_prefetch_next(lw.down, p.ffn, p.d, 0);
    dp_combine_prologue<HD>(p, xs);
    dp_gemv_rows(lw.o, p.H * HD, p.d, xs, p.vec_o);
    bar.sync();
    // ---- D: add + post-attention norm -> gate|up with SiLU*mul ----   (HBM meanwhile: next layer's q|k|v and o)
    if (l + 1 < p.L) {
      dp_prefetch_next(p.layers[l + 1].qkv, p.d, qkv_rows, 0);
      dp_prefetch_next(p.layers[l + 1].o, p.H * HD, p.d, 0);
    } else if (p.logits) {
      dp_prefetch_next(p.lm_head, p.d, p.vocab, 0, 8);  // the first ~40 MB of the lm_head stream
    }
    dp_norm_prologue(p.vec_o, resid[cur], resid[cur ^ 1], lw.post_norm, p.eps, p.d, xs, red);
    cur ^= 1;
    dp_gemv_gate_up(lw.gate_up, p.d, p.ffn, xs, p.vec_act);
    bar.sync();
    // ---- E: down-proj ----
    for (int i = threadIdx.x * 8; i < p.ffn; i += kDpThreads * 8) {
      float x[8];
      unpack_bf16x8(dp_ldcg16(p.vec_act + i), x);
#pragma unroll
      for (int j = 0; j < 8; ++j) xs[i + j] = x[j];
    }
    __syncthreads();
    dp_gemv_rows(lw.down, p.ffn, p.d, xs, p.vec_down);
    bar.sync();
  }
  if (p.logits) {
    // final add + norm (models/llama3.py:198) -> lm_head; logits rounded to bf16 like every linear output
    dp_norm_prologue(p.vec_down, resid[cur], nullptr, p.final_norm, p.eps, p.d, xs, red);
    dp_gemv_rows(p.lm_head, p.d, p.vocab, xs, p.logits);
  }
  // every CTA read launch_count before its first barrier arrival, so the bump cannot be seen early
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.launch_count = __ldcg(p.launch_count) + 1u;
}

}  // namespace ssdk
