// Runs the SOURCE of the one-shot all-reduce (ar_publish_kernel + add_rmsnorm_kernel with symmetric input,
// csrc/elementwise.cuh) for R emulated ranks that share host memory the way NVLink peer mappings share HBM.
// TEST INFRASTRUCTURE.   run_allreduce <in> <out>     blob layout: tests/test_allreduce_emu_cpu.py (the writer).
#include "cuda_emu.h"
#define SSDK_HOST_EMU 1
#include "../../ssd_b200/csrc/elementwise.cuh"

#include <fstream>
#include <iostream>
#include <random>

using bf16 = __nv_bfloat16;

template <typename T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
  if (!f && n) std::exit(2);
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  auto hdr = rd<int32_t>(f, 7);
  // n_calls all-reduces per forward (odd, like the 2L+1 of a real forward), n_fwd forwards back to back with NO other
  // cross-rank synchronisation in between; rank `slow` dawdles in the last consumer of every forward
  const int R = hdr[0], M = hdr[1], d = hdr[2], S = hdr[3], per_fwd = hdr[4], threads = hdr[5], n_fwd = hdr[6];
  const int n_calls = per_fwd * n_fwd, slow = R - 1;
  const float eps = rd<float>(f, 1)[0];
  auto w = rd<bf16>(f, d);
  auto resid0 = rd<bf16>(f, (size_t)M * d);
  std::vector<std::vector<float>> partials(R);  // [call][S][M][d] per rank
  for (int r = 0; r < R; ++r) partials[r] = rd<float>(f, (size_t)n_calls * S * M * d);

  const unsigned slot_bytes = (unsigned)((size_t)M * d * 4);
  std::vector<std::vector<uint64_t>> symm(R, std::vector<uint64_t>((size_t)2 * ssdk::kSymmMaxRanks * slot_bytes / 8, 0));
  std::vector<std::vector<bf16>> resid(R, resid0), y(R, std::vector<bf16>((size_t)n_calls * M * d));
  std::vector<unsigned> fwd_seq(R, 1);
  const int slices = (d + 8 * threads - 1) / (8 * threads);

  std::vector<std::thread> ranks;
  for (int rank = 0; rank < R; ++rank)
    ranks.emplace_back([&, rank] {
      std::mt19937 rng(1234 + rank);
      for (int call = 0; call < n_calls; ++call) {
        const int idx = call % per_fwd;
        if (idx == 0 && call > 0) fwd_seq[rank] += 1;  // prep_kernel bumps the forward sequence number
        std::this_thread::sleep_for(std::chrono::microseconds(rng() % 3000));  // skew the ranks against each other
        ssdk::ArPublishParams ap;
        std::memset(&ap, 0, sizeof(ap));
        ap.x.partial = partials[rank].data() + (size_t)call * S * M * d;
        ap.x.S = S; ap.x.M = M; ap.x.N = d;
        ap.M = M; ap.d = d; ap.n_ranks = R; ap.rank = rank;
        for (int r = 0; r < R; ++r) ap.peer[r] = reinterpret_cast<uint8_t*>(symm[r].data());
        ap.slot_bytes = slot_bytes; ap.fwd_seq = &fwd_seq[rank]; ap.call_idx = idx; ap.n_calls = per_fwd;
        emu::launch(ssdk::ar_publish_kernel, ap, std::max(1, std::min((M * d / 8 + 255) / 256, 4)), 256, 0);
        std::this_thread::sleep_for(std::chrono::microseconds(rng() % 2000));
        if (rank == slow && idx == per_fwd - 1) std::this_thread::sleep_for(std::chrono::milliseconds(30));
        ssdk::NormParams np;
        std::memset(&np, 0, sizeof(np));
        np.symm.base = reinterpret_cast<const uint8_t*>(symm[rank].data());
        np.symm.fwd_seq = &fwd_seq[rank]; np.symm.no_dep_wait = 1; np.symm.call_idx = idx; np.symm.n_calls = per_fwd; np.symm.n_ranks = R;
        np.symm.slot_bytes = slot_bytes;
        np.residual_in = resid[rank].data(); np.residual_out = resid[rank].data(); np.w = w.data(); np.eps = eps;
        np.y = y[rank].data() + (size_t)call * M * d; np.d = d;
        if (slices == 1) emu::launch(ssdk::add_rmsnorm_kernel<1>, np, M, threads, 0);
        else if (slices == 2) emu::launch(ssdk::add_rmsnorm_kernel<2>, np, M, threads, 0);
        else emu::launch(ssdk::add_rmsnorm_kernel<0>, np, M, threads, (size_t)d * 4);
      }
    });
  for (auto& t : ranks) t.join();
  std::ofstream o(argv[2], std::ios::binary);
  for (int r = 0; r < R; ++r) {
    o.write(reinterpret_cast<const char*>(y[r].data()), (std::streamsize)(y[r].size() * 2));
    o.write(reinterpret_cast<const char*>(resid[r].data()), (std::streamsize)(resid[r].size() * 2));
  }
  return 0;
}
