// Runs the SOURCE of csrc/sampling.cuh (sample_kernel / verify_kernel) on host threads (cuda_emu.h).  TEST INFRASTRUCTURE.
//   run_sampling sample <in> <out>    |    run_sampling verify <in> <out>
// blob layouts: see tests/test_sampling_emu_cpu.py (the writer).
#include "cuda_emu.h"
#define SSDK_HOST_EMU 1
#include "../../ssd_b200/csrc/sampling.cuh"

#include <fstream>
#include <iostream>
#include <string>

using bf16 = __nv_bfloat16;

struct Reader {
  std::ifstream f;
  explicit Reader(const char* p) : f(p, std::ios::binary) {
    if (!f) std::exit(2);
  }
  template <typename T>
  std::vector<T> vec(size_t n) {
    std::vector<T> v(n);
    f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
    if (!f && n) std::exit(2);
    return v;
  }
  int64_t i64() { return vec<int64_t>(1)[0]; }
};

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  const std::string mode = argv[1];
  Reader r(argv[2]);
  std::ofstream o(argv[3], std::ios::binary);
  if (mode == "sample") {
    const int B = (int)r.i64(), V = (int)r.i64(), nch = (int)r.i64();
    const uint64_t seed = (uint64_t)r.i64(), call_id = (uint64_t)r.i64();
    auto temps = r.vec<float>(B);
    auto logits = r.vec<bf16>((size_t)B * V);
    std::vector<int64_t> out(B, -1);
    std::vector<ssdk::ArgMax> partial((size_t)B * nch);
    std::vector<unsigned> counters(B, 0);
    ssdk::SampleParams p;
    std::memset(&p, 0, sizeof(p));
    p.logits = logits.data(); p.ld = V; p.temps = temps.data(); p.V = V; p.seed = seed; p.call_id = call_id;
    p.out = out.data(); p.out_stride = 1; p.partial = partial.data(); p.counters = counters.data();
    dim3 grid;
    grid.x = (unsigned)nch;
    grid.y = (unsigned)B;
    emu::launch(ssdk::sample_kernel, p, grid, 256, 0, /*wave=*/4);  // last-block pattern: CTAs never wait for each other
    for (unsigned c : counters)
      if (c != 0) return 3;  // the ticket must be reset for the next launch
    o.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)(out.size() * 8));
    return 0;
  }
  if (mode == "verify") {
    const int B = (int)r.i64(), K = (int)r.i64(), V = (int)r.i64(), nct = (int)r.i64(), jit = (int)r.i64();
    const int has_hits = (int)r.i64();
    const uint64_t seed = (uint64_t)r.i64(), call_id = (uint64_t)r.i64();
    auto tt = r.vec<float>(B), tq = r.vec<float>(B);
    auto hits = r.vec<int32_t>(has_hits ? B : 0);
    auto spec = r.vec<int64_t>((size_t)B * (K + 1));
    auto lp = r.vec<bf16>((size_t)B * (K + 1) * V), lq = r.vec<bf16>((size_t)B * K * V);
    std::vector<int32_t> nacc(B, -1);
    std::vector<int64_t> rec(B, -1);
    std::vector<ssdk::RowPart> rows((size_t)B * (2 * K + 1) * nct);
    std::vector<ssdk::RecPart> recs((size_t)B * nct);
    unsigned counters[2] = {0, 0};
    ssdk::VerifyParams p;
    std::memset(&p, 0, sizeof(p));
    p.lp = lp.data(); p.lq = lq.data(); p.spec = spec.data(); p.temps_t = tt.data(); p.temps_q = tq.data();
    p.cache_hits = has_hits ? hits.data() : nullptr;
    p.jit = jit; p.B = B; p.K = K; p.V = V; p.seed = seed; p.call_id = call_id;
    p.n_accept = nacc.data(); p.recovery = rec.data(); p.row_part = rows.data(); p.rec_part = recs.data();
    p.counters = counters;
    emu::launch(ssdk::verify_kernel, p, nct, ssdk::kVerifyThreads, 0);  // device-wide barrier: all CTAs co-resident
    if (counters[0] != 0 || counters[1] != 0) return 3;
    o.write(reinterpret_cast<const char*>(nacc.data()), (std::streamsize)(nacc.size() * 4));
    o.write(reinterpret_cast<const char*>(rec.data()), (std::streamsize)(rec.size() * 8));
    return 0;
  }
  return 2;
}
