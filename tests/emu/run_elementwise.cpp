// Runs the SOURCE of add_rmsnorm_kernel (split-K / embedding input), rope_store_kernel and silu_mul_kernel
// (csrc/elementwise.cuh) on host threads.  TEST INFRASTRUCTURE.   run_elementwise <norm|rope|silu> <in> <out>
#include "cuda_emu.h"
#define SSDK_HOST_EMU 1
#include "../../ssd_b200/csrc/elementwise.cuh"

#include <fstream>
#include <string>

using bf16 = __nv_bfloat16;
template <typename T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
  if (!f && n) std::exit(2);
  return v;
}
template <typename T>
static void wr(std::ofstream& o, const std::vector<T>& v) {
  o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  const std::string mode = argv[1];
  std::ifstream f(argv[2], std::ios::binary);
  std::ofstream o(argv[3], std::ios::binary);
  if (mode == "norm") {
    // header: M d S threads use_ids vocab has_residual ; eps ; w[d] ; residual[M*d] ; (S>0: partials[S*M*d] | S==0: dense[M*d])
    // ; ids[M] int64 ; embed[vocab*d]
    auto h = rd<int32_t>(f, 7);
    const int M = h[0], d = h[1], S = h[2], threads = h[3], use_ids = h[4], vocab = h[5], has_res = h[6];
    const float eps = rd<float>(f, 1)[0];
    auto w = rd<bf16>(f, d);
    auto residual = rd<bf16>(f, (size_t)M * d);
    auto partial = rd<float>(f, S > 0 ? (size_t)S * M * d : 0);
    auto dense = rd<bf16>(f, S == 0 ? (size_t)M * d : 0);
    auto ids = rd<int64_t>(f, M);
    auto embed = rd<bf16>(f, use_ids ? (size_t)vocab * d : 0);
    std::vector<bf16> y((size_t)M * d), res_out((size_t)M * d);
    ssdk::NormParams p;
    std::memset(&p, 0, sizeof(p));
    p.x.dense = S == 0 ? dense.data() : nullptr; p.x.partial = S > 0 ? partial.data() : nullptr; p.x.S = S; p.x.M = M; p.x.N = d;
    if (use_ids) { p.ids = ids.data(); p.ids_stride = 1; p.embed = embed.data(); p.vocab_start = 0; p.vocab_rows = vocab; }
    p.residual_in = has_res ? residual.data() : nullptr; p.w = w.data(); p.eps = eps; p.y = y.data();
    p.residual_out = res_out.data(); p.d = d;
    const int slices = (d + 8 * threads - 1) / (8 * threads);
    if (slices == 1) emu::launch(ssdk::add_rmsnorm_kernel<1>, p, M, threads, 0);
    else if (slices == 2) emu::launch(ssdk::add_rmsnorm_kernel<2>, p, M, threads, 0);
    else emu::launch(ssdk::add_rmsnorm_kernel<0>, p, M, threads, (size_t)d * 4);
    wr(o, y); wr(o, res_out);
    return 0;
  }
  if (mode == "rope") {
    // header: M H KV hd S max_pos nslots qk_norm ; eps ; positions[M] i64 ; slots[M] i32 ; table[max_pos*hd] f32 ;
    // qn[hd] kn[hd] bf16 ; partials[S*M*qkv_dim]
    auto h = rd<int32_t>(f, 8);
    const int M = h[0], H = h[1], KV = h[2], hd = h[3], S = h[4], max_pos = h[5], nslots = h[6], qk_norm = h[7];
    const float eps = rd<float>(f, 1)[0];
    auto pos = rd<int64_t>(f, M);
    auto slots = rd<int32_t>(f, M);
    auto table = rd<float>(f, (size_t)max_pos * hd);
    auto qn = rd<bf16>(f, hd), kn = rd<bf16>(f, hd);
    const int qkv_dim = (H + 2 * KV) * hd;
    auto partial = rd<float>(f, (size_t)S * M * qkv_dim);
    std::vector<bf16> q((size_t)M * H * hd), kc((size_t)nslots * KV * hd), vc((size_t)nslots * KV * hd);
    ssdk::RopeParams p;
    std::memset(&p, 0, sizeof(p));
    p.qkv.partial = partial.data(); p.qkv.S = S; p.qkv.M = M; p.qkv.N = qkv_dim;
    p.positions = pos.data(); p.slot_mapping = slots.data(); p.rope_table = table.data();
    p.q_norm_w = qk_norm ? qn.data() : nullptr; p.k_norm_w = qk_norm ? kn.data() : nullptr; p.norm_eps = eps;
    p.q_out = q.data(); p.k_cache = kc.data(); p.v_cache = vc.data(); p.heads = H; p.kv_heads = KV; p.head_dim = hd;
    dim3 grid;
    grid.x = (unsigned)M;
    grid.y = (unsigned)((H + 2 * KV + 3) / 4);
    // CTAs exit early for head slots beyond H + 2 KV (whole warps), never between barriers: waves are fine
    if (hd == 64) emu::launch(ssdk::rope_store_kernel<64>, p, grid, 128, 0, 8);
    else if (hd == 128) emu::launch(ssdk::rope_store_kernel<128>, p, grid, 128, 0, 8);
    else emu::launch(ssdk::rope_store_kernel<256>, p, grid, 128, 0, 8);
    wr(o, q); wr(o, kc); wr(o, vc);
    return 0;
  }
  if (mode == "silu") {
    auto h = rd<int32_t>(f, 3);
    const int M = h[0], ffn = h[1], S = h[2];
    auto partial = rd<float>(f, (size_t)S * M * 2 * ffn);
    std::vector<bf16> out((size_t)M * ffn);
    ssdk::GemmOut x;
    std::memset(&x, 0, sizeof(x));
    x.partial = partial.data(); x.S = S; x.M = M; x.N = 2 * ffn;
    struct P { ssdk::GemmOut x; bf16* out; int M, ffn; } pp{x, out.data(), M, ffn};
    emu::launch(+[](P q) { ssdk::silu_mul_kernel(q.x, q.out, q.M, q.ffn); }, pp, (M * ffn / 8 + 255) / 256, 256, 0);
    wr(o, out);
    return 0;
  }
  return 2;
}
