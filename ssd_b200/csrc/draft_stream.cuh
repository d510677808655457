// draft_stream.cuh — the whole speculate phase of a sync-SD step (SpeculatorSync.speculate, engine/speculator_sync.py:25-69:
// K+1 single-token draft forwards + K samplings) as ONE persistent kernel, batch 1.
//
// Why: the 1B draft streams 2.47 GB per forward (0.376 ms at the measured HBM peak) but the kernel-per-op path needs
// 0.79 ms — nine kernel boundaries per layer, during each of which HBM idles (profiles/r01_small_kernels.md).  A first
// persistent kernel (round 1, five device-wide barriers per layer, weights read straight from global memory) measured
// 0.85 ms on hardware: a barrier costs as much as a boundary, and HBM still idles at every one of them
// (profiles/r02_draft_persistent.md).  The weights, however, do not depend on the activations.  So here every CTA
// (one per SM) streams ITS share of every matrix, in program order and without ever waiting for a phase, through a ring
// of 32 KB shared-memory stages fed by bulk async copies (cp.async.bulk global->shared, mbarrier completion, L2
// evict-first): while the CTAs meet at a barrier or recompute a norm, the ring (148 x 160 KB = 23 MB on chip) fills
// with the NEXT phases' weights, optionally backed by an L2 prefetch window beyond it.  Consumption is plain CUDA-core
// GEMV from shared memory (one token: 1 FMA per weight element; 0.4 us of shared-memory time per 0.73 us of HBM time).
//
// Program order per forward (same five phases per layer as the reference's decoder layer, models/llama3.py:185-199):
//   A  [residual add + input RMSNorm, recomputed by every CTA] -> q|k|v rows
//   B  q/k head norm + RoPE + KV store + split-KV attention for (kv head, split) units
//   C  [merge of the split partials, every CTA] -> o-proj rows
//   D  [residual add + post-attention RMSNorm, every CTA] -> gate|up row pairs + SiLU*mul
//   E  down-proj rows
// then final norm -> lm_head rows -> in-kernel sampling (greedy argmax over the bf16 logits, lowest index wins, or the
// Philox exponential race of layers/sampler.py:27-34 — the SAME scores sample_kernel computes, so the tokens are identical)
// -> the next forward starts inside the same launch.  The last forward of a step only writes KV (speculator_sync.py:52-56).
// Rounding points are the reference's (SURVEY §8a checklist 1-4): every linear output, the residual, the norm output, q/k
// after RoPE and the attention output are rounded to bf16; accumulation is fp32.
//
// Matrix rows are dealt to CTAs in stages of R consecutive rows (R * K * 2 bytes <= 32 KB): stage s of a matrix belongs to
// CTA s mod #CTAs, and a stage is ONE contiguous bulk copy (two for gate|up: R/2 gate rows + the R/2 matching up rows).
// Inside a stage the 8 warps split (row, K-segment) units: K <= 2048: 8 rows x 1 segment, K <= 4096: 4 x 2, K <= 8192: 2 x 4.
#pragma once
#include "common.cuh"
#include "sampling.cuh"

namespace ssdk {

constexpr int kDsThreads = 256;
constexpr int kDsWarps = kDsThreads / 32;
constexpr int kDsMaxLayers = 32;
constexpr int kDsSplits = 8;          // KV splits per kv head in phase B
constexpr int kDsStageBytes = 32768;  // one ring stage
constexpr int kDsMaxStages = 6;

struct DsLayer {
  const __nv_bfloat16 *qkv, *o, *gate_up, *down, *in_norm, *post_norm, *q_norm, *k_norm;
};

struct DsParams {
  int d, L, H, KV, ffn, vocab, qk_norm;
  float eps, scale_log2;
  const __nv_bfloat16 *embed, *final_norm, *lm_head;
  const float* rope;  // [max_pos, hd]: cos | sin
  __nv_bfloat16 *k_cache, *v_cache;
  long long cache_layer_stride;  // elements between layers
  int block_size, max_blocks;
  int64_t* tok_buf;              // [n_fwd (+1)]: tok_buf[0] = first input token; the kernel writes tok_buf[f + 1]
  int n_fwd;                     // forwards in this launch
  int skip_last_head;            // 1: the last forward runs without lm_head / sampling (it only writes KV)
  const int32_t* ctx0;           // tokens in the cache before the first forward
  const int32_t* block_table;    // [max_blocks]
  __nv_bfloat16 *vec_qkv, *vec_o, *vec_act, *vec_down, *resid0, *resid1;
  float* attn_part;              // [H][kDsSplits][hd + 2]: o | m | l
  __nv_bfloat16* logits;         // row f at logits + f * logits_ld (may be null: no logits kept)
  long long logits_ld;
  const float* temp;             // draft temperature [1] (device)
  const uint64_t* dyn;           // optional device {seed, step}: call_id = step * 16 + f
  uint64_t seed, call_base;      // used when dyn == nullptr: call_id = call_base + f
  ArgMax* samp_partial;          // [#CTAs]
  unsigned* bar_state;           // [0] arrivals, [1] generation (both zero before the first launch ever)
  int n_stages;                  // ring depth (3 .. kDsMaxStages)
  int l2_ahead;                  // stages requested into L2 beyond the ring (0 = off)
  DsLayer layers[kDsMaxLayers];
};

// ---------------------------------------------------------------------------------------------
// device-wide barrier, generation based (no per-launch bookkeeping): the last arriver resets the arrival count and bumps
// the generation, everybody else polls the generation.  Every CTA reads the generation before its first arrival, i.e.
// before the first barrier of the launch can complete.
// ---------------------------------------------------------------------------------------------
struct DsGridBar {
  unsigned* state;
  unsigned gen;
  __device__ void init() {
    if (threadIdx.x == 0) gen = ld_acquire_u32(state + 1);
  }
  __device__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned old = atomicAdd(state, 1u);
      if (old == gridDim.x - 1) {
        atomicAdd(state, 0u - gridDim.x);  // back to zero (atomic: nobody else touches it until the generation moves)
        __threadfence();
        atomicAdd(state + 1, 1u);
      } else {
        const long long t0 = clock64();
        while (ld_acquire_u32(state + 1) == gen) {
          if (clock64() - t0 > 4000000000LL) __trap();  // a CTA never arrived: fail loudly instead of hanging the GPU
        }
      }
      gen += 1u;
      __threadfence();
    }
    __syncthreads();
  }
};

SSDK_DEVINL uint4 ds_ldcg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
SSDK_DEVINL float2 ds_bf2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

// ---------------------------------------------------------------------------------------------
// matrix geometry and the per-CTA job sequence
// ---------------------------------------------------------------------------------------------
enum { DS_QKV = 0, DS_O = 1, DS_GU = 2, DS_DOWN = 3, DS_HEAD = 4 };
struct DsMat {
  const __nv_bfloat16* w;
  int K;      // row length
  int rows;   // output rows (gate|up: pairs)
  int R;      // rows per stage (gate|up: R / 2 pairs)
  int segs;   // K segments per row: R * segs == 8 warps
  int pair;   // 1: gate|up
};
__host__ SSDK_DEVINL void ds_geometry(int K, int* R, int* segs) {
  *R = K <= 2048 ? 8 : (K <= 4096 ? 4 : 2);
  *segs = kDsWarps / *R;
}
template <int HD>
SSDK_DEVINL DsMat ds_mat(const DsParams& p, int l, int m) {
  DsMat t;
  t.pair = 0;
  const int lc = l < p.L ? l : 0;
  switch (m) {
    case DS_QKV: t.w = p.layers[lc].qkv; t.K = p.d; t.rows = (p.H + 2 * p.KV) * HD; break;
    case DS_O: t.w = p.layers[lc].o; t.K = p.H * HD; t.rows = p.d; break;
    case DS_GU: t.w = p.layers[lc].gate_up; t.K = p.d; t.rows = p.ffn; t.pair = 1; break;
    case DS_DOWN: t.w = p.layers[lc].down; t.K = p.ffn; t.rows = p.d; break;
    default: t.w = p.lm_head; t.K = p.d; t.rows = p.vocab; break;
  }
  ds_geometry(t.K, &t.R, &t.segs);
  return t;
}
SSDK_DEVINL int ds_rows_per_stage(const DsMat& t) { return t.pair ? t.R / 2 : t.R; }
SSDK_DEVINL int ds_num_stages(const DsMat& t) {
  const int r = ds_rows_per_stage(t);
  return (t.rows + r - 1) / r;
}
SSDK_DEVINL bool ds_has_head(const DsParams& p, int f) { return !(p.skip_last_head && f == p.n_fwd - 1); }

// position in the program: forward f, layer l (l == L: the lm_head), matrix m, stage s (s = CTA, CTA + #CTAs, ...)
struct DsCursor {
  int f, l, m, s;
  bool valid;
};
template <int HD>
SSDK_DEVINL void ds_cursor_settle(const DsParams& p, DsCursor& c) {  // skip matrices in which this CTA has no stage left
  while (c.valid && c.s >= ds_num_stages(ds_mat<HD>(p, c.l, c.m))) {
    c.s = (int)blockIdx.x;
    if (c.m == DS_HEAD) {
      c.f++; c.l = 0; c.m = DS_QKV;
    } else if (c.m == DS_DOWN) {
      c.l++; c.m = DS_QKV;
      if (c.l == p.L) {
        if (ds_has_head(p, c.f)) c.m = DS_HEAD;
        else { c.f++; c.l = 0; }
      }
    } else {
      c.m++;
    }
    if (c.f >= p.n_fwd) c.valid = false;
  }
}
template <int HD>
SSDK_DEVINL void ds_cursor_init(const DsParams& p, DsCursor& c) {
  c.f = 0; c.l = 0; c.m = DS_QKV; c.s = (int)blockIdx.x; c.valid = p.n_fwd > 0;
  ds_cursor_settle<HD>(p, c);
}
template <int HD>
SSDK_DEVINL void ds_cursor_advance(const DsParams& p, DsCursor& c) {
  c.s += (int)gridDim.x;
  ds_cursor_settle<HD>(p, c);
}
// (source, bytes) of a job: one contiguous run of rows, two for gate|up (gate rows, then the matching up rows)
struct DsJob {
  const __nv_bfloat16 *src0, *src1;
  unsigned bytes0, bytes1, off1;
};
template <int HD>
SSDK_DEVINL DsJob ds_job(const DsParams& p, const DsCursor& c) {
  const DsMat t = ds_mat<HD>(p, c.l, c.m);
  const int r = ds_rows_per_stage(t);
  const int n = min(r, t.rows - c.s * r);
  DsJob j;
  j.src0 = t.w + (size_t)c.s * r * t.K;
  j.bytes0 = (unsigned)n * (unsigned)t.K * 2u;
  j.src1 = nullptr; j.bytes1 = 0; j.off1 = 0;
  if (t.pair) {
    j.src1 = t.w + ((size_t)t.rows + (size_t)c.s * r) * t.K;
    j.bytes1 = j.bytes0;
    j.off1 = (unsigned)r * (unsigned)t.K * 2u;
  }
  return j;
}

// ---------------------------------------------------------------------------------------------
// the ring: stage `slot` is refilled by thread 0 right after the block has finished reading it
// ---------------------------------------------------------------------------------------------
struct DsRing {
  uint8_t* base;
  uint64_t* full;     // [n_stages] mbarriers (count 1 + transaction bytes)
  int n_stages;
  unsigned consumed;  // jobs consumed so far (uniform over the block)
  DsCursor issue;     // thread 0 only: next job to request into shared memory
  DsCursor ahead;     // thread 0 only: next job to request into L2
};
template <int HD>
SSDK_DEVINL void ds_issue(const DsParams& p, DsRing& r, int slot) {  // thread 0
  if (!r.issue.valid) return;
  const DsJob j = ds_job<HD>(p, r.issue);
  uint8_t* dst = r.base + (size_t)slot * kDsStageBytes;
  mbar_arrive_expect_tx(&r.full[slot], j.bytes0 + j.bytes1);
  bulk_load_g2s(dst, j.src0, j.bytes0, &r.full[slot]);
  if (j.bytes1) bulk_load_g2s(dst + j.off1, j.src1, j.bytes1, &r.full[slot]);
  ds_cursor_advance<HD>(p, r.issue);
  if (p.l2_ahead > 0 && r.ahead.valid) {
    const DsJob a = ds_job<HD>(p, r.ahead);
    bulk_prefetch_l2(a.src0, a.bytes0);
    if (a.bytes1) bulk_prefetch_l2(a.src1, a.bytes1);
    ds_cursor_advance<HD>(p, r.ahead);
  }
}

// partial dot of one row segment held in shared memory with the matching slice of x (fp32, shared memory)
SSDK_DEVINL float ds_dot_seg(const uint8_t* wseg, const float* xseg, int len, int lane) {
  const uint4* wp = reinterpret_cast<const uint4*>(wseg) + lane;
  const float* xp = xseg + lane * 8;
  float a0 = 0.f, a1 = 0.f;
  const int steps = len >> 8;  // 256 elements per step
#pragma unroll 4
  for (int j = 0; j < steps; ++j) {
    const uint4 w = wp[j * 32];
    const float4 xa = *reinterpret_cast<const float4*>(xp + j * 256);
    const float4 xb = *reinterpret_cast<const float4*>(xp + j * 256 + 4);
    float2 f = ds_bf2(w.x);
    a0 = fmaf(f.x, xa.x, a0); a1 = fmaf(f.y, xa.y, a1);
    f = ds_bf2(w.y);
    a0 = fmaf(f.x, xa.z, a0); a1 = fmaf(f.y, xa.w, a1);
    f = ds_bf2(w.z);
    a0 = fmaf(f.x, xb.x, a0); a1 = fmaf(f.y, xb.y, a1);
    f = ds_bf2(w.w);
    a0 = fmaf(f.x, xb.z, a0); a1 = fmaf(f.y, xb.w, a1);
  }
  return warp_sum(a0 + a1);
}

// sampling state of the lm_head phase (threads 0 .. R-1 of warp 0 each follow their own rows)
struct DsSample {
  bool greedy;
  float invT;
  uint64_t seed, call_id;
  ArgMax best;
};
SSDK_DEVINL float ds_score(const DsSample& s, float logit, int idx) {
  if (s.greedy) return logit;
  // scores = softmax(l/T) / (E + 1e-10) in the log domain, the very expression of sample_kernel (sampling.cuh)
  const uint4 rnd = philox_draw((uint32_t)(idx >> 2), 0u, s.call_id, TAG_SAMPLE, s.seed);
  const float e = u32_to_exp1(u4_word(rnd, idx & 3)) + 1e-10f;
  return logit * s.invT - __logf(e);
}

// consume this CTA's stages of matrix (l, m).  MODE 0: y[row] = bf16(W[row] . x); MODE 1: gate|up pairs ->
// act[i] = bf16(silu(bf16 g) * bf16 u) (layers/activation.py:11-14); MODE 2: lm_head rows -> logits + running argmax.
template <int HD, int MODE>
SSDK_DEVINL void ds_consume(const DsParams& p, DsRing& ring, int l, int m, const float* xs, float* res,
                            __nv_bfloat16* y, DsSample* smp) {
  const DsMat t = ds_mat<HD>(p, l, m);
  const int ns = ds_num_stages(t);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_in_stage = warp / t.segs, seg = warp - row_in_stage * t.segs;
  const int seg_len = t.K / t.segs;
  const int rps = ds_rows_per_stage(t);
  for (int s = (int)blockIdx.x; s < ns; s += (int)gridDim.x) {
    const int slot = (int)(ring.consumed % (unsigned)ring.n_stages);
    const uint32_t parity = (ring.consumed / (unsigned)ring.n_stages) & 1u;
    mbar_wait(&ring.full[slot], parity);
    const uint8_t* st = ring.base + (size_t)slot * kDsStageBytes;
    const float acc = ds_dot_seg(st + ((size_t)row_in_stage * t.K + (size_t)seg * seg_len) * 2, xs + seg * seg_len, seg_len, lane);
    float* rb = res + (ring.consumed & 1u) * kDsWarps;
    if (lane == 0) rb[warp] = acc;
    __syncthreads();  // the stage has been read by every warp; the partial sums are visible
    if (threadIdx.x == 0) ds_issue<HD>(p, ring, slot);
    if ((int)threadIdx.x < rps) {
      const int row = s * rps + (int)threadIdx.x;
      if (row < t.rows) {
        if (MODE == 1) {
          float g = 0.f, u = 0.f;
          for (int q = 0; q < t.segs; ++q) {
            g += rb[(int)threadIdx.x * t.segs + q];
            u += rb[(rps + (int)threadIdx.x) * t.segs + q];
          }
          g = bf16_round(g);
          u = bf16_round(u);
          y[row] = f2bf((g / (1.0f + __expf(-g))) * u);
        } else {
          float v = 0.f;
          for (int q = 0; q < t.segs; ++q) v += rb[(int)threadIdx.x * t.segs + q];
          const __nv_bfloat16 o = f2bf(v);
          if (MODE == 0) {
            y[row] = o;
          } else {
            if (y) y[row] = o;
            smp->best = argmax_better(smp->best, ArgMax{ds_score(*smp, bf2f(o), row), row});
          }
        }
      }
    }
    ring.consumed += 1u;
  }
}

// xs[i] = bf16r( r_i * rsqrt(mean r^2 + eps) * w_i ),  r = a (+ b) in fp32;  resid_out = bf16(r) (written by CTA 0 only).
// a / b are L2-resident vectors produced by earlier phases.  d is a multiple of 8.
SSDK_DEVINL void ds_norm_prologue(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* resid_out,
                                  const __nv_bfloat16* w, float eps, int d, float* xs, float* red) {
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < d; i += kDsThreads * 8) {
    float x[8];
    unpack_bf16x8(ds_ldcg16(a + i), x);
    if (b) {
      float y[8];
      unpack_bf16x8(ds_ldcg16(b + i), y);
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
  for (int i = threadIdx.x * 8; i < d; i += kDsThreads * 8) {
    float wv[8];
    unpack_bf16x8(*reinterpret_cast<const uint4*>(w + i), wv);
#pragma unroll
    for (int j = 0; j < 8; ++j) xs[i + j] = bf16_round(xs[i + j] * rstd * wv[j]);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// phase B unit: kv head h, split s.  Rebuilds the rotated q rows of the head group and the new token's k / v from the
// q|k|v vector, stores k / v into the page slot (split 0 only), runs the online-softmax over its token range (the new
// token comes from shared memory, never from the cache) and writes (o, m, l) per query head.
// ---------------------------------------------------------------------------------------------
template <int HD, int GMAX>
SSDK_DEVINL void ds_attention_unit(const DsParams& p, int layer, int h, int s, int ctx, float* sm) {
  constexpr int HALF = HD / 2;
  constexpr int EPL = HD / 32;  // elements per lane in the dot layout (dims lane*EPL ..)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = p.H / p.KV;
  const int pos = ctx - 1;
  float* sq = sm;                      // [G][HD] rotated q (bf16-rounded values)
  float* sk = sq + GMAX * HD;          // [HD] new k
  float* sv = sk + HD;                 // [HD] new v
  float* sred = sv + HD;               // [kDsWarps][G][HD + 2] per-warp partials

  // ---- q rows, k, v: one warp per row, rotate-half pairs (i, i + HALF) ----
  const float* cs = p.rope + (size_t)pos * HD;
  for (int row = warp; row < G + 2; row += kDsWarps) {
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
    for (int i = threadIdx.x; i < HD; i += kDsThreads) {
      kc[i] = f2bf(sk[i]);
      vc[i] = f2bf(sv[i]);
    }
  }

  // ---- token range of this split ----
  const int per = (ctx + kDsSplits - 1) / kDsSplits;
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
  for (int t = t0 + warp; t < t1; t += kDsWarps) {
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
        const float2 a = ds_bf2(__ldcg(reinterpret_cast<const uint32_t*>(kbase + off)));
        const float2 b = ds_bf2(__ldcg(reinterpret_cast<const uint32_t*>(vbase + off)));
        kv[0] = a.x; kv[1] = a.y; vv[0] = b.x; vv[1] = b.y;
      } else {
        static_assert(EPL == 2 || EPL == 4, "head_dim 64 or 128");
        const uint2 a = __ldcg(reinterpret_cast<const uint2*>(kbase + off));
        const uint2 b = __ldcg(reinterpret_cast<const uint2*>(vbase + off));
        float2 f = ds_bf2(a.x); kv[0] = f.x; kv[1] = f.y;
        f = ds_bf2(a.y); kv[EPL - 2] = f.x; kv[EPL - 1] = f.y;
        f = ds_bf2(b.x); vv[0] = f.x; vv[1] = f.y;
        f = ds_bf2(b.y); vv[EPL - 2] = f.x; vv[EPL - 1] = f.y;
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
  for (int idx = threadIdx.x; idx < G * HD; idx += kDsThreads) {
    const int g = idx / HD, dim = idx - g * HD;
    float mx = -INFINITY;
    for (int w = 0; w < kDsWarps; ++w) mx = fmaxf(mx, sred[((size_t)w * G + g) * LDR + HD]);
    float o = 0.f, ll = 0.f;
    if (mx != -INFINITY) {
      for (int w = 0; w < kDsWarps; ++w) {
        const float* r = sred + ((size_t)w * G + g) * LDR;
        const float wt = exp2f(r[HD] - mx);
        o += r[dim] * wt;
        ll += r[HD + 1] * wt;
      }
    }
    float* out = p.attn_part + ((size_t)(h * G + g) * kDsSplits + s) * LDR;
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
SSDK_DEVINL void ds_combine_prologue(const DsParams& p, float* xs) {
  constexpr int LDR = HD + 2;
  for (int idx = threadIdx.x; idx < p.H * HD; idx += kDsThreads) {
    const int head = idx / HD, dim = idx - head * HD;
    const float* base = p.attn_part + (size_t)head * kDsSplits * LDR;
    float ms[kDsSplits], ls[kDsSplits], os[kDsSplits];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < kDsSplits; ++s) {
      ms[s] = __ldcg(base + s * LDR + HD);
      ls[s] = __ldcg(base + s * LDR + HD + 1);
      os[s] = __ldcg(base + s * LDR + dim);
      mx = fmaxf(mx, ms[s]);
    }
    float o = 0.f, l = 0.f;
#pragma unroll
    for (int s = 0; s < kDsSplits; ++s) {
      const float wt = (ms[s] == -INFINITY) ? 0.f : exp2f(ms[s] - mx);
      o += os[s] * wt;
      l += ls[s] * wt;
    }
    xs[idx] = bf16_round(l > 0.f ? o / l : 0.f);
  }
  __syncthreads();
}

template <int HD, int GMAX>
__global__ void __launch_bounds__(kDsThreads, 1) draft_stream_kernel(const __grid_constant__ DsParams p) {
  SSDK_DYN_SMEM(uint8_t, ds_smem);
  SSDK_STATIC_SMEM(uint64_t, full_bar, kDsMaxStages);
  SSDK_STATIC_SMEM(float, red, 32);
  SSDK_STATIC_SMEM(float, res, 2 * kDsWarps);
  SSDK_STATIC_SMEM(ArgMax, ared, 32);
  SSDK_SHARED_VAR(int, tok_s);
  // dynamic shared memory: [ring: n_stages x 32 KB][xs: max(d, ffn, H*HD) floats][attention scratch]
  DsRing ring;
  ring.base = ds_smem;
  ring.full = full_bar;
  ring.n_stages = p.n_stages;
  ring.consumed = 0u;
  float* xs = reinterpret_cast<float*>(ds_smem + (size_t)p.n_stages * kDsStageBytes);
  float* scratch = xs + max(max(p.d, p.ffn), p.H * HD);
  if (threadIdx.x == 0) {
    trace_mark(TR_MISC);
    for (int s = 0; s < p.n_stages; ++s) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
    // the weights do not depend on anything this launch computes: fill the ring and request the L2 window at once
    ds_cursor_init<HD>(p, ring.issue);
    for (int s = 0; s < p.n_stages && ring.issue.valid; ++s) {
      const DsJob j = ds_job<HD>(p, ring.issue);
      uint8_t* dst = ring.base + (size_t)s * kDsStageBytes;
      mbar_arrive_expect_tx(&ring.full[s], j.bytes0 + j.bytes1);
      bulk_load_g2s(dst, j.src0, j.bytes0, &ring.full[s]);
      if (j.bytes1) bulk_load_g2s(dst + j.off1, j.src1, j.bytes1, &ring.full[s]);
      ds_cursor_advance<HD>(p, ring.issue);
    }
    ring.ahead = ring.issue;  // from here on `ahead` stays l2_ahead jobs in front of `issue` (ds_issue moves both)
    for (int s = 0; s < p.l2_ahead && ring.ahead.valid; ++s) {
      const DsJob a = ds_job<HD>(p, ring.ahead);
      bulk_prefetch_l2(a.src0, a.bytes0);
      if (a.bytes1) bulk_prefetch_l2(a.src1, a.bytes1);
      ds_cursor_advance<HD>(p, ring.ahead);
    }
  }
  __syncthreads();  // mbarriers initialised before anybody waits on them

  DsGridBar bar;
  bar.state = p.bar_state;
  bar.gen = 0;
  bar.init();

  const int ctx_base = p.ctx0[0];
  const float T = p.temp ? p.temp[0] : 0.f;
  const uint64_t seed = p.dyn ? p.dyn[0] : p.seed;
  const uint64_t call0 = p.dyn ? p.dyn[1] * 16ull : p.call_base;
  __nv_bfloat16* resid[2] = {p.resid0, p.resid1};
  long long tok = p.tok_buf[0];

  for (int f = 0; f < p.n_fwd; ++f) {
    if (threadIdx.x == 0 && f > 0) trace_mark(TR_MISC);
    const int ctx = ctx_base + f + 1;  // tokens visible to this forward, the new one included
    const __nv_bfloat16* emb = p.embed + (size_t)tok * p.d;
    int cur = 0;  // resid[cur] holds the residual entering the layer (layer 0: the embedding row itself)
    for (int l = 0; l < p.L; ++l) {
      const DsLayer& lw = p.layers[l];
      // ---- A: (add +) input norm -> q|k|v ----
      if (l == 0) {
        // first layer: hidden = norm(embed), residual = embed (models/llama3.py:192-193)
        ds_norm_prologue(emb, nullptr, resid[cur ^ 1], lw.in_norm, p.eps, p.d, xs, red);
      } else {
        ds_norm_prologue(p.vec_down, resid[cur], resid[cur ^ 1], lw.in_norm, p.eps, p.d, xs, red);
      }
      cur ^= 1;
      ds_consume<HD, 0>(p, ring, l, DS_QKV, xs, res, p.vec_qkv, nullptr);
      bar.sync();
      // ---- B: RoPE + KV store + attention units ----
      for (int u = blockIdx.x; u < p.KV * kDsSplits; u += gridDim.x)
        ds_attention_unit<HD, GMAX>(p, l, u / kDsSplits, u % kDsSplits, ctx, scratch);
      bar.sync();
      // ---- C: merge splits -> o-proj ----
      ds_combine_prologue<HD>(p, xs);
      ds_consume<HD, 0>(p, ring, l, DS_O, xs, res, p.vec_o, nullptr);
      bar.sync();
      // ---- D: add + post-attention norm -> gate|up with SiLU*mul ----
      ds_norm_prologue(p.vec_o, resid[cur], resid[cur ^ 1], lw.post_norm, p.eps, p.d, xs, red);
      cur ^= 1;
      ds_consume<HD, 1>(p, ring, l, DS_GU, xs, res, p.vec_act, nullptr);
      bar.sync();
      // ---- E: down-proj ----
      for (int i = threadIdx.x * 8; i < p.ffn; i += kDsThreads * 8) {
        float x[8];
        unpack_bf16x8(ds_ldcg16(p.vec_act + i), x);
#pragma unroll
        for (int j = 0; j < 8; ++j) xs[i + j] = x[j];
      }
      __syncthreads();
      ds_consume<HD, 0>(p, ring, l, DS_DOWN, xs, res, p.vec_down, nullptr);
      bar.sync();
    }
    if (!ds_has_head(p, f)) break;
    // ---- final add + norm (models/llama3.py:198) -> lm_head; logits rounded to bf16 like every linear output ----
    ds_norm_prologue(p.vec_down, resid[cur], nullptr, p.final_norm, p.eps, p.d, xs, red);
    DsSample smp;
    smp.greedy = (T == 0.f);
    smp.invT = smp.greedy ? 1.f : 1.f / T;
    smp.seed = seed;
    smp.call_id = call0 + (uint64_t)f;
    smp.best = ArgMax{-INFINITY, 0x7fffffff};
    ds_consume<HD, 2>(p, ring, p.L, DS_HEAD, xs, res, p.logits ? p.logits + (size_t)f * p.logits_ld : nullptr, &smp);
    // ---- sampling: per-CTA best -> device-wide reduction (every CTA learns the token) ----
    if (threadIdx.x < 32) {
      const ArgMax b = warp_argmax(smp.best);  // threads 0 .. R-1 hold candidates, the others the neutral element
      if (threadIdx.x == 0) p.samp_partial[blockIdx.x] = b;
    }
    bar.sync();
    {
      ArgMax a{-INFINITY, 0x7fffffff};
      for (int i = threadIdx.x; i < (int)gridDim.x; i += kDsThreads) {
        ArgMax q;
        q.v = __ldcg(&p.samp_partial[i].v);
        q.i = __ldcg(&p.samp_partial[i].i);
        a = argmax_better(a, q);
      }
      a = block_argmax(a, ared);
      if (threadIdx.x == 0) tok_s = a.i;
      __syncthreads();
      tok = tok_s;
      if (blockIdx.x == 0 && threadIdx.x == 0) p.tok_buf[f + 1] = tok;
    }
  }
}

}  // namespace ssdk
