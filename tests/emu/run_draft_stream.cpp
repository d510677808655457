// Runs the SOURCE of csrc/draft_stream.cuh on host threads (cuda_emu.h).  TEST INFRASTRUCTURE.
//   run_draft_persistent <input blob> <output blob>
// blob layout: see tests/test_draft_persistent_emu_cpu.py (the writer).
#include "cuda_emu.h"
#define SSDK_HOST_EMU 1
#include "../../ssd_b200/csrc/draft_stream.cuh"

#include <fstream>
#include <iostream>

using bf16 = __nv_bfloat16;

struct Reader {
  std::ifstream f;
  explicit Reader(const char* p) : f(p, std::ios::binary) {
    if (!f) {
      std::cerr << "cannot open " << p << "\n";
      std::exit(2);
    }
  }
  template <typename T>
  std::vector<T> vec(size_t n) {
    std::vector<T> v(n);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
    if (!f) {
      std::cerr << "short read\n";
      std::exit(2);
    }
    return v;
  }
  int i32() { return vec<int32_t>(1)[0]; }
  float f32() { return vec<float>(1)[0]; }
};

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  Reader r(argv[1]);
  const int d = r.i32(), L = r.i32(), H = r.i32(), KV = r.i32(), hd = r.i32(), ffn = r.i32(), vocab = r.i32();
  const int qk_norm = r.i32(), block_size = r.i32(), max_blocks = r.i32(), nslots = r.i32(), ctx0 = r.i32();
  const int n_fwd = r.i32(), grid = r.i32(), max_pos = r.i32(), n_stages = r.i32(), skip_last = r.i32();
  const float eps = r.f32();
  float temp = r.f32();
  auto rng = r.vec<uint64_t>(2);  // seed, call_base
  auto tokens = r.vec<int64_t>(1);  // first input token
  auto block_table = r.vec<int32_t>(max_blocks);
  auto embed = r.vec<bf16>((size_t)vocab * d), final_norm = r.vec<bf16>(d), lm_head = r.vec<bf16>((size_t)vocab * d);
  auto rope = r.vec<float>((size_t)max_pos * hd);
  const int qkv_dim = (H + 2 * KV) * hd;
  struct LW {
    std::vector<bf16> qkv, o, gate_up, down, in_norm, post_norm, q_norm, k_norm;
  };
  std::vector<LW> lw(L);
  for (auto& w : lw) {
    w.qkv = r.vec<bf16>((size_t)qkv_dim * d);
    w.o = r.vec<bf16>((size_t)d * H * hd);
    w.gate_up = r.vec<bf16>((size_t)2 * ffn * d);
    w.down = r.vec<bf16>((size_t)d * ffn);
    w.in_norm = r.vec<bf16>(d);
    w.post_norm = r.vec<bf16>(d);
    w.q_norm = r.vec<bf16>(hd);
    w.k_norm = r.vec<bf16>(hd);
  }
  const size_t cache_layer = (size_t)nslots * KV * hd;
  auto kc = r.vec<bf16>(cache_layer * L), vc = r.vec<bf16>(cache_layer * L);

  std::vector<bf16> vecs((size_t)qkv_dim + 4 * d + ffn + H * hd + 64), logits((size_t)n_fwd * vocab);
  std::vector<float> attn((size_t)H * ssdk::kDsSplits * (hd + 2));
  std::vector<ssdk::ArgMax> partial(grid);
  std::vector<int64_t> tok_buf(n_fwd + 1, -1);
  tok_buf[0] = tokens[0];
  alignas(8) unsigned sync[64] = {0};
  int32_t ctx0_dev = ctx0;

  ssdk::DsParams p;
  std::memset(&p, 0, sizeof(p));
  p.d = d; p.L = L; p.H = H; p.KV = KV; p.ffn = ffn; p.vocab = vocab; p.qk_norm = qk_norm;
  p.eps = eps;
  p.scale_log2 = (1.0f / std::sqrt((float)hd)) * 1.4426950408889634f;
  p.embed = embed.data(); p.final_norm = final_norm.data(); p.lm_head = lm_head.data(); p.rope = rope.data();
  p.k_cache = kc.data(); p.v_cache = vc.data();
  p.cache_layer_stride = (long long)cache_layer;
  p.block_size = block_size; p.max_blocks = max_blocks;
  p.tok_buf = tok_buf.data(); p.n_fwd = n_fwd; p.skip_last_head = skip_last;
  p.ctx0 = &ctx0_dev; p.block_table = block_table.data();
  bf16* v = vecs.data();
  p.vec_qkv = v; v += (qkv_dim + 7) / 8 * 8;
  p.vec_attn = v; v += H * hd;
  p.vec_o = v; v += d;
  p.vec_down = v; v += d;
  p.resid0 = v; v += d;
  p.resid1 = v; v += d;
  p.vec_act = v;
  p.attn_part = attn.data();
  p.logits = logits.data(); p.logits_ld = vocab;
  p.temp = &temp; p.dyn = nullptr; p.seed = rng[0]; p.call_base = rng[1];
  p.samp_partial = partial.data();
  p.bar_state = sync;
  p.attn_ticket = sync + 8;
  p.n_slots = n_stages;
  for (int l = 0; l < L; ++l)
    p.layers[l] = ssdk::DsLayer{lw[l].qkv.data(), lw[l].o.data(), lw[l].gate_up.data(), lw[l].down.data(),
                                lw[l].in_norm.data(), lw[l].post_norm.data(), lw[l].q_norm.data(), lw[l].k_norm.data()};
  const int G = H / KV, gmax = G <= 4 ? 4 : 8;
  const size_t xs = (size_t)std::max(std::max(d, ffn), H * hd);
  const size_t scratch = (size_t)gmax * hd + 2 * hd + (size_t)ssdk::kDsWarps * gmax * (hd + 2);
  const size_t smem = (xs + scratch) * 4 + 256 + (size_t)n_stages * ssdk::kDsSlotBytes;
  // two launches on the same barrier state: the generation-based barrier must carry over
  for (int rep = 0; rep < 2; ++rep) {
    if (rep == 1) {  // second launch: same inputs again (KV rows are simply rewritten with the same values)
      std::fill(tok_buf.begin() + 1, tok_buf.end(), -1);
    }
    if (hd == 64 && gmax == 4) emu::launch(ssdk::draft_stream_kernel<64, 4>, p, grid, ssdk::kDsThreads, smem);
    else if (hd == 64) emu::launch(ssdk::draft_stream_kernel<64, 8>, p, grid, ssdk::kDsThreads, smem);
    else if (gmax == 4) emu::launch(ssdk::draft_stream_kernel<128, 4>, p, grid, ssdk::kDsThreads, smem);
    else emu::launch(ssdk::draft_stream_kernel<128, 8>, p, grid, ssdk::kDsThreads, smem);
    for (int i = 2; i < 64; ++i)
      if (sync[i] != 0) {
        std::cerr << "attention ticket " << i << " not back to zero\n";
        return 3;
      }
    unsigned long long arrivals;
    std::memcpy(&arrivals, sync, 8);
    if (arrivals % (unsigned long long)grid != 0) {
      std::cerr << "barrier arrival counter is not a whole number of barriers\n";
      return 3;
    }
  }
  std::ofstream o(argv[2], std::ios::binary);
  o.write(reinterpret_cast<const char*>(logits.data()), (std::streamsize)(logits.size() * 2));
  o.write(reinterpret_cast<const char*>(kc.data()), (std::streamsize)(kc.size() * 2));
  o.write(reinterpret_cast<const char*>(vc.data()), (std::streamsize)(vc.size() * 2));
  o.write(reinterpret_cast<const char*>(tok_buf.data()), (std::streamsize)(tok_buf.size() * 8));
  return 0;
}
