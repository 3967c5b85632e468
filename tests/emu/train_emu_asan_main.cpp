// Sanitizer run of the training-step kernel chain on the CPU (TEST INFRASTRUCTURE ONLY): every workspace slice is followed
// by a poisoned red zone, so any out-of-range read or write of any kernel aborts with an AddressSanitizer report.
// Build + run:  g++ -O1 -g -std=c++17 -fsanitize=address -DCATRE_HOST_EMU -DCATRE_EMU_ASAN train_emu_asan_main.cpp -o /tmp/train_asan && /tmp/train_asan
// (values are random: this checks addressing, not arithmetic -- tests/test_train_emu.py does that)
#include <stdio.h>

#include <random>
#include <utility>

#include "train_emu.cpp"

int main(int argc, char** argv) {
  // (objects, points per set): ragged tiles, one / many GroupNorm chunks; with any argument also the 4096-row shape that
  // switches the weight-gradient GEMMs to split-K (about a minute under the sanitizer)
  std::vector<std::pair<int, int>> shapes = {{1, 100}, {3, 64}, {2, 136}};
  if (argc > 1) shapes.push_back({2, 1024});
  std::mt19937 rng(5);
  std::normal_distribution<float> nd(0.0f, 0.05f);
  for (auto& sh : shapes) {
    const int B = sh.first, N = sh.second;
    std::vector<std::vector<float>> w(W_COUNT), g(W_COUNT);
    std::vector<const float*> wp(W_COUNT);
    std::vector<float*> gp(W_COUNT);
    for (int i = 0; i < W_COUNT; ++i) {
      w[i].resize(weight_numel(i, N));
      g[i].resize(weight_numel(i, N));
      for (auto& v : w[i]) v = nd(rng);
      wp[i] = w[i].data();
      gp[i] = g[i].data();
    }
    auto vec = [&](size_t n, float scale, float shift) { std::vector<float> v(n); for (auto& x : v) x = nd(rng) * scale + shift; return v; };
    auto pcl = vec((size_t)B * N * 3, 2.0f, 0.0f), kps = vec((size_t)B * N * 3, 4.0f, 0.0f), scale = vec((size_t)B * 3, 0.5f, 0.2f);
    auto gt_scale = vec((size_t)B * 3, 0.5f, 0.2f);
    std::vector<float> pose((size_t)B * 12, 0.0f), gt_pose((size_t)B * 12, 0.0f), K((size_t)B * 9, 0.0f);
    for (int b = 0; b < B; ++b) {
      for (int d = 0; d < 3; ++d) { pose[b * 12 + d * 4 + d] = 1.0f; gt_pose[b * 12 + d * 4 + d] = 1.0f; }
      pose[b * 12 + 11] = 1.0f; gt_pose[b * 12 + 11] = 1.1f;
      K[b * 9 + 0] = 591.0f; K[b * 9 + 4] = 590.0f; K[b * 9 + 8] = 1.0f;
    }
    std::vector<unsigned char> is_sym(B);
    for (int b = 0; b < B; ++b) is_sym[b] = b % 2 == 0;
    std::vector<float> rots(5 * 9, 0.0f);
    for (int r = 0; r < 5; ++r) { rots[r * 9] = rots[r * 9 + 4] = rots[r * 9 + 8] = 1.0f; }
    std::vector<float> pose_out((size_t)B * 12), scale_out((size_t)B * 3), losses(6);
    long launches = 0;
    double macs = 0;
    for (int reposed = 0; reposed < 2; ++reposed) {
      int rc = emu_train_step(wp.data(), B, N, pcl.data(), kps.data(), pose.data(), scale.data(), K.data(), gt_pose.data(), gt_scale.data(),
                              is_sym.data(), rots.data(), 5, pose_out.data(), scale_out.data(), losses.data(), gp.data(), &launches,
                              reposed ? pcl.data() : nullptr, reposed ? kps.data() : nullptr, &macs, nullptr);
      printf("B=%d N=%d reposed=%d: rc=%d launches=%ld loss[0]=%g\n", B, N, reposed, rc, launches, losses[0]);
      if (rc) return 1;
    }
  }
  printf("no out-of-range access\n");
  return 0;
}
