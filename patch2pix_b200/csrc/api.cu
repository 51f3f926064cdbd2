// C ABI of libp2p_b200.so (declared in include/p2p_b200.h): handle, weight packing, stage drivers.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../../include/p2p_b200.h"
#include "kernels.h"
#include "umma_gemm.h"

namespace p2p {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
long long g_launch_count = 0;

int ensure_dyn_smem(const void* kernel, int bytes) {
  struct Key { const void* k; int dev; };
  static thread_local std::vector<std::pair<Key, int>> granted;
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& g : granted)
    if (g.first.k == kernel && g.first.dev == dev) {
      if (g.second >= bytes) return 0;
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
        set_last_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ") + cudaGetErrorString(cudaGetLastError()));
        return -2;
      }
      g.second = bytes;
      return 0;
    }
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
    set_last_error(std::string("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: ") + cudaGetErrorString(cudaGetLastError()));
    return -2;
  }
  granted.push_back({Key{kernel, dev}, bytes});
  return 0;
}

int Arena::reserve(size_t bytes) {
  off = 0;
  if (bytes <= cap) return 0;
  if (base != nullptr) {
    cudaDeviceSynchronize();
    cudaFree(base);
    base = nullptr;
    cap = 0;
  }
  bytes = align_up(bytes + (bytes >> 3), 1 << 20);
  if (cudaMalloc(&base, bytes) != cudaSuccess) {
    cudaGetLastError();
    set_last_error("out of device memory reserving " + std::to_string(bytes >> 20) + " MiB of scratch");
    return -3;
  }
  cap = bytes;
  return 0;
}
void Arena::release() {
  if (base != nullptr) cudaFree(base);
  base = nullptr;
  cap = off = 0;
}

struct Regressor {
  bool set = false;
  __half *w1_hi = nullptr, *w1_lo = nullptr;  // [512][73*64]
  __half *w2_hi = nullptr, *w2_lo = nullptr;  // [512][72*64]
  float *scale1 = nullptr, *bias1 = nullptr, *scale2 = nullptr, *bias2 = nullptr;
  float y_scale = 1.f;
  FcWeights fc = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // tensor-core FC path: K-major fp16 hi/lo weights [out][in] with per-row pow2 scale, 1/(act*w scale), bias
  __half *f1_hi = nullptr, *f1_lo = nullptr, *f2_hi = nullptr, *f2_lo = nullptr;
  float *fa1 = nullptr, *fa2 = nullptr;
  KStep steps1[kConv1Steps];
  KStep steps2[kConv2Steps];
  KStep *d_steps1 = nullptr, *d_steps2 = nullptr;
  char* blob = nullptr;  // one allocation backing all of the above
};

}  // namespace p2p

using namespace p2p;

struct p2p_handle_s {
  int device = 0;
  int num_sms = 148;
  int opt_mid_passes = 3, opt_fine_passes = 1, opt_corr_passes = 3, opt_seg_len = 3, opt_gemm_impl = 0, opt_num_sms = 0;
  int opt_mid_band = 26;  // thousandths of a pixel (2x the largest 1-pass/3-pass mid difference over 125k distinct coordinates,
                          // profiles/r02_band_stats.json); 0 = pure 3-pass mid stage
  int opt_gemm_pair = 35;   // bitmask of launches that use the CTA-pair (cta_group::2) GEMM kernel:
                            // 1: 1-pass convs, 2: 3-pass convs, 4: FC, 8: correlation, 16: p2p_test_gemm,
                            // 32: fused-gather conv1
  int opt_fc_impl = 1;      // 1: the two big Linear layers on the tensor cores (3-pass); 0: CUDA-core FC kernel
  int opt_fuse_gather = 3;  // conv1 A operand of the 1-pass launches: 3 (default): strided TMA boxes of a per-image window map
                            // (umma_conv1_tma_kernel); 1: gathered by producer warps (128x512 tiles, lookup tables); 2: first-
                            // generation fused kernel (128x256 tiles); 0: separate gather kernel + TMA of the patch tensor
  const int* last_band_count = nullptr;  // device counter of the last risk-band subset
  unsigned long long* band_totals = nullptr;   // device: {band rows, rows} summed over mid-stage calls
  bool nc_set = false;
  float *nc_w1p = nullptr, *nc_b1p = nullptr, *nc_w2p = nullptr;
  float nc_b2 = 0.f;
  NcUmmaWeights ncw;            // tensor-core NC operand images
  void* dbg_nc[4] = {nullptr, nullptr, nullptr, nullptr};   // scratch of the last p2p_neigh_consensus call (tools/nc_debug.py)
  int opt_unique_impl = 1;      // 1: rank sort over the whole GPU for lists <= 8192 rows; 0: single-block bitonic network
  int* uniq_rank = nullptr;     // zeroed scratch of the rank-sort path
  int opt_nc_l2_mode = 0;       // NC layer 2 block layout: 0 auto, 1 one haloed block per tile, 2 one block per column tap
  int opt_nc_impl = 1;          // 1: NeighConsensus on the tensor cores (nc_umma.cu); 0: fp32 CUDA-core kernels (shape-capped)
  Regressor reg[2];
  Arena coarse, refine, feat, misc, uniq, pre;
  std::vector<PreprocessCoefs> pre_coefs;   // cached resampling tables, one per image geometry
  PairFeatures pf[2];
  bool prepared = false;
  // optional per-kernel CUDA-event profile (p2p_set_option("profile", 1))
  int opt_profile = 0;
  struct ProfEntry { cudaEvent_t a, b; int kind; };
  std::vector<ProfEntry> prof;
  std::vector<cudaEvent_t> event_pool;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define P2P_ENTER(h)                                                  \
  P2P_REQUIRE((h) != nullptr, "null handle");                         \
  DeviceGuard _guard((h)->device);                                    \
  if (!_guard.ok) {                                                   \
    set_last_error("cannot select CUDA device of the handle");        \
    return -2;                                                        \
  }

float pow2_floor_scale(float maxabs, float target_hi) {
  // power of two s such that maxabs * s lies in [target_hi/2, target_hi)
  if (!(maxabs > 0.f) || !isfinite(maxabs)) return 1.f;
  int e;
  frexpf(maxabs, &e);  // maxabs = m * 2^e, m in [0.5,1)
  int et;
  frexpf(target_hi, &et);
  return ldexpf(1.f, et - 1 - e);
}

void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

template <typename T>
T* carve(char*& p, size_t count) {
  T* r = reinterpret_cast<T*>(p);
  p += align_up(count * sizeof(T), 256);
  return r;
}

void fold_bn(const p2p_bn_t& bn, int n, float eps, std::vector<float>& g, std::vector<float>& b) {
  g.resize(n);
  b.resize(n);
  for (int i = 0; i < n; ++i) {
    g[i] = bn.weight[i] / sqrtf(bn.running_var[i] + eps);
    b[i] = bn.bias[i] - bn.running_mean[i] * g[i];
  }
}

int tap_plane(int t, int& start) {  // conv1 tap -> (parity plane bit, box start)
  if (t == 0) { start = -1; return 0; }
  if (t == 1) { start = 0; return 1; }
  start = 0;
  return 0;
}

int pack_regressor(p2p_handle_s* h, Regressor& R, const p2p_regressor_weights_t& w) {
  const int K1 = kConv1Steps * 64, K2 = kConv2Steps * 64;
  std::vector<float> g1, b1, g2, b2, gf1, bf1, gf2, bf2;
  fold_bn(w.conv1_bn, 512, w.bn_eps, g1, b1);
  fold_bn(w.conv3_bn, 512, w.bn_eps, g2, b2);
  fold_bn(w.fc1_bn, 512, w.bn_eps, gf1, bf1);
  fold_bn(w.fc4_bn, 256, w.bn_eps, gf2, bf2);

  std::vector<__half> w1h((size_t)512 * K1), w1l((size_t)512 * K1), w2h((size_t)512 * K2), w2l((size_t)512 * K2);
  std::vector<float> sc1(512), sc2(512), bi1(b1), bi2(b2);
  float max_bound = 0.f;
  std::vector<float> sw1(512), sw2(512);
  for (int o = 0; o < 512; ++o) {
    const float* wo = w.conv0_weight + (size_t)o * 518 * 9;
    float m = 0.f;
    double tapn[9] = {0};
    for (int c = 0; c < 518; ++c)
      for (int t = 0; t < 9; ++t) {
        const float v = wo[c * 9 + t] * g1[o];
        m = fmaxf(m, fabsf(v));
        tapn[t] += (double)v * v;
      }
    float bound = fabsf(b1[o]);
    for (int t = 0; t < 9; ++t) bound += 1.41421357f * (float)sqrt(tapn[t]);
    max_bound = fmaxf(max_bound, bound);
    sw1[o] = pow2_floor_scale(m, 1024.f);
    sc1[o] = 1.f / (kActScale * sw1[o]);
    __half* dh = w1h.data() + (size_t)o * K1;
    __half* dl = w1l.data() + (size_t)o * K1;
    for (int s = 0; s < 72; ++s) {
      const int tap = s / 8, chunk = s % 8;
      for (int kk = 0; kk < 64; ++kk) {
        const int c512 = chunk * 64 + kk;
        const int orig = (c512 / 256) * 259 + 3 + (c512 % 256);
        split_half(wo[orig * 9 + tap] * g1[o] * sw1[o], dh[s * 64 + kk], dl[s * 64 + kk]);
      }
    }
    for (int kk = 0; kk < 64; ++kk) {
      float v = 0.f;
      if (kk < 54) {
        const int tap = kk / 6, r = kk % 6;
        const int orig = (r / 3) * 259 + (r % 3);
        v = wo[orig * 9 + tap] * g1[o] * sw1[o];
      }
      split_half(v, dh[72 * 64 + kk], dl[72 * 64 + kk]);
    }
  }
  R.y_scale = pow2_floor_scale(max_bound, 32768.f);
  for (int o = 0; o < 512; ++o) {
    const float* wo = w.conv2_weight + (size_t)o * 512 * 9;
    float m = 0.f;
    for (int i = 0; i < 512 * 9; ++i) m = fmaxf(m, fabsf(wo[i] * g2[o]));
    sw2[o] = pow2_floor_scale(m, 1024.f);
    sc2[o] = 1.f / (R.y_scale * sw2[o]);
    __half* dh = w2h.data() + (size_t)o * K2;
    __half* dl = w2l.data() + (size_t)o * K2;
    for (int s = 0; s < 72; ++s) {
      const int tap = s / 8, chunk = s % 8;
      for (int kk = 0; kk < 64; ++kk)
        split_half(wo[(chunk * 64 + kk) * 9 + tap] * g2[o] * sw2[o], dh[s * 64 + kk], dl[s * 64 + kk]);
    }
  }
  // k-step plans
  for (int s = 0; s < 72; ++s) {
    const int tap = s / 8, chunk = s % 8, ty = tap / 3, tx = tap % 3;
    int sx, sy;
    const int px = tap_plane(tx, sx), py = tap_plane(ty, sy);
    R.steps1[s] = KStep{(short)(chunk * 64), (signed char)sx, (signed char)sy, (signed char)(py * 2 + px), 0, 0, s * 64};
    R.steps2[s] = KStep{(short)(chunk * 64), (signed char)(tx - 1), (signed char)(ty - 1), 0, 0, 0, s * 64};
  }
  R.steps1[72] = KStep{0, 0, 0, 0, 1, 0, 72 * 64};
  // FC (BN folded, transposed)
  std::vector<float> f1t((size_t)512 * 512), f1b(512), f2t((size_t)512 * 256), f2b(256), f3t(256 * 5), f3b(5);
  for (int o = 0; o < 512; ++o) {
    for (int k = 0; k < 512; ++k) f1t[(size_t)k * 512 + o] = w.fc0_weight[(size_t)o * 512 + k] * gf1[o];
    f1b[o] = w.fc0_bias[o] * gf1[o] + bf1[o];
  }
  for (int o = 0; o < 256; ++o) {
    for (int k = 0; k < 512; ++k) f2t[(size_t)k * 256 + o] = w.fc3_weight[(size_t)o * 512 + k] * gf2[o];
    f2b[o] = w.fc3_bias[o] * gf2[o] + bf2[o];
  }
  for (int o = 0; o < 5; ++o) {
    for (int k = 0; k < 256; ++k) f3t[k * 5 + o] = w.fc6_weight[o * 256 + k];
    f3b[o] = w.fc6_bias[o];
  }
  // tensor-core FC operands
  std::vector<__half> f1h((size_t)512 * 512), f1l((size_t)512 * 512), f2h((size_t)256 * 512), f2l((size_t)256 * 512);
  std::vector<float> fa1(512), fa2(256);
  for (int o = 0; o < 512; ++o) {
    float m = 0.f;
    for (int k = 0; k < 512; ++k) m = fmaxf(m, fabsf(w.fc0_weight[(size_t)o * 512 + k] * gf1[o]));
    const float sw = pow2_floor_scale(m, 1024.f);
    fa1[o] = 1.f / (kFcActScale * sw);
    for (int k = 0; k < 512; ++k)
      split_half(w.fc0_weight[(size_t)o * 512 + k] * gf1[o] * sw, f1h[(size_t)o * 512 + k], f1l[(size_t)o * 512 + k]);
  }
  for (int o = 0; o < 256; ++o) {
    float m = 0.f;
    for (int k = 0; k < 512; ++k) m = fmaxf(m, fabsf(w.fc3_weight[(size_t)o * 512 + k] * gf2[o]));
    const float sw = pow2_floor_scale(m, 1024.f);
    fa2[o] = 1.f / (kFcActScale * sw);
    for (int k = 0; k < 512; ++k)
      split_half(w.fc3_weight[(size_t)o * 512 + k] * gf2[o] * sw, f2h[(size_t)o * 512 + k], f2l[(size_t)o * 512 + k]);
  }
  // device blob
  const size_t total = 2 * 1024 * 1024 + 65536 + 4 * align_up((size_t)512 * K1 * 2, 256) + 8 * 4096 + align_up(f1t.size() * 4, 256) +
                       align_up(f2t.size() * 4, 256) + 8 * 8192 + 65536;
  if (R.blob == nullptr) {
    if (cudaMalloc(&R.blob, total) != cudaSuccess) {
      cudaGetLastError();
      set_last_error("out of device memory packing regressor weights");
      return -3;
    }
  }
  char* p = R.blob;
  R.w1_hi = carve<__half>(p, (size_t)512 * K1);
  R.w1_lo = carve<__half>(p, (size_t)512 * K1);
  R.w2_hi = carve<__half>(p, (size_t)512 * K2);
  R.w2_lo = carve<__half>(p, (size_t)512 * K2);
  R.scale1 = carve<float>(p, 512);
  R.bias1 = carve<float>(p, 512);
  R.scale2 = carve<float>(p, 512);
  R.bias2 = carve<float>(p, 512);
  R.fc.w1t = carve<float>(p, f1t.size());
  R.fc.b1 = carve<float>(p, 512);
  R.fc.w2t = carve<float>(p, f2t.size());
  R.fc.b2 = carve<float>(p, 256);
  R.fc.w3t = carve<float>(p, f3t.size());
  R.fc.b3 = carve<float>(p, 5);
  R.d_steps1 = carve<KStep>(p, kConv1Steps);
  R.d_steps2 = carve<KStep>(p, kConv2Steps);
  R.f1_hi = carve<__half>(p, f1h.size());
  R.f1_lo = carve<__half>(p, f1l.size());
  R.f2_hi = carve<__half>(p, f2h.size());
  R.f2_lo = carve<__half>(p, f2l.size());
  R.fa1 = carve<float>(p, 512);
  R.fa2 = carve<float>(p, 256);
#define UP(dst, src, bytes) P2P_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice))
  UP(R.w1_hi, w1h.data(), w1h.size() * 2);
  UP(R.w1_lo, w1l.data(), w1l.size() * 2);
  UP(R.w2_hi, w2h.data(), w2h.size() * 2);
  UP(R.w2_lo, w2l.data(), w2l.size() * 2);
  UP(R.scale1, sc1.data(), 2048);
  UP(R.bias1, bi1.data(), 2048);
  UP(R.scale2, sc2.data(), 2048);
  UP(R.bias2, bi2.data(), 2048);
  UP(R.fc.w1t, f1t.data(), f1t.size() * 4);
  UP(R.fc.b1, f1b.data(), 2048);
  UP(R.fc.w2t, f2t.data(), f2t.size() * 4);
  UP(R.fc.b2, f2b.data(), 1024);
  UP(R.fc.w3t, f3t.data(), f3t.size() * 4);
  UP(R.fc.b3, f3b.data(), 20);
  UP(R.d_steps1, R.steps1, sizeof(R.steps1));
  UP(R.d_steps2, R.steps2, sizeof(R.steps2));
  UP(R.f1_hi, f1h.data(), f1h.size() * 2);
  UP(R.f1_lo, f1l.data(), f1l.size() * 2);
  UP(R.f2_hi, f2h.data(), f2h.size() * 2);
  UP(R.f2_lo, f2l.data(), f2l.size() * 2);
  UP(R.fa1, fa1.data(), 2048);
  UP(R.fa2, fa2.data(), 1024);
#undef UP
  R.set = true;
  (void)h;
  return 0;
}

int sms(const p2p_handle_s* h) { return h->opt_num_sms > 0 ? h->opt_num_sms : h->num_sms; }

// Brackets a group of launches with CUDA events on the launching stream when profiling is on.
struct ProfScope {
  p2p_handle_s* h;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  int kind;
  static cudaEvent_t get(p2p_handle_s* h) {
    if (!h->event_pool.empty()) {
      cudaEvent_t e = h->event_pool.back();
      h->event_pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  ProfScope(p2p_handle_s* h_, int kind_, cudaStream_t st_) : h(h_), st(st_), kind(kind_) {
    if (h->opt_profile) {
      a = get(h);
      b = get(h);
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope() {
    if (a != nullptr) {
      cudaEventRecord(b, st);
      h->prof.push_back({a, b, kind});
    }
  }
};

}  // namespace

extern "C" {

const char* p2p_last_error(void) { return g_last_error.c_str(); }
int p2p_version(void) { return 100; }

int p2p_create(int device, p2p_handle_t* out) {
  P2P_REQUIRE(out != nullptr, "out is null");
  int ndev = 0;
  P2P_CUDA_OK(cudaGetDeviceCount(&ndev));
  P2P_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  cudaDeviceProp prop;
  P2P_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_last_error(std::string("libp2p_b200 targets sm_100a (B200) only; device is sm_") + std::to_string(prop.major) +
                   std::to_string(prop.minor));
    return -2;
  }
  p2p_handle_s* h = new p2p_handle_s();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  {
    DeviceGuard g(device);
    if (cudaMalloc(&h->band_totals, 2 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(h->band_totals, 0, 2 * sizeof(unsigned long long));
    else
      h->band_totals = nullptr;
  }
  *out = h;
  return 0;
}

int p2p_destroy(p2p_handle_t h) {
  if (h == nullptr) return 0;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  h->coarse.release();
  h->refine.release();
  h->feat.release();
  h->misc.release();
  h->uniq.release();
  h->pre.release();
  for (auto& c : h->pre_coefs)
    if (c.d) cudaFree(c.d);
  for (auto& e : h->prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  if (h->nc_w1p) cudaFree(h->nc_w1p);
  if (h->ncw.blob) cudaFree(h->ncw.blob);
  if (h->band_totals) cudaFree(h->band_totals);
  if (h->uniq_rank) cudaFree(h->uniq_rank);
  for (int i = 0; i < 2; ++i)
    if (h->reg[i].blob) cudaFree(h->reg[i].blob);
  delete h;
  return 0;
}

int p2p_set_ncn_weights(p2p_handle_t h, const float* w1, const float* b1, const float* w2, const float* b2) {
  P2P_ENTER(h);
  P2P_REQUIRE(w1 && b1 && w2 && b2, "null weight pointer");
  // reference layout [k1][Cout][Cin][k2][k3][k4]; logical W[o][c][a][b][d][e] = weight[a][o][c][b][d][e]
  std::vector<float> w1p(81 * 32), b1p(32), w2p(81 * 32);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
          const int tap = ((a * 3 + b) * 3 + d) * 3 + e;
          for (int c = 0; c < 16; ++c) {
            // net 0: plain weights; net 1: (a,b) <-> (d,e) swapped (the "transposed" pass)
            w1p[tap * 32 + c] = w1[((a * 16 + c) * 1 + 0) * 27 + (b * 3 + d) * 3 + e];
            w1p[tap * 32 + 16 + c] = w1[((d * 16 + c) * 1 + 0) * 27 + (e * 3 + a) * 3 + b];
            w2p[tap * 32 + c] = w2[((a * 1 + 0) * 16 + c) * 27 + (b * 3 + d) * 3 + e];
            w2p[tap * 32 + 16 + c] = w2[((d * 1 + 0) * 16 + c) * 27 + (e * 3 + a) * 3 + b];
          }
        }
  for (int c = 0; c < 16; ++c) b1p[c] = b1p[16 + c] = b1[c];
  if (h->nc_w1p == nullptr) P2P_CUDA_OK(cudaMalloc(&h->nc_w1p, sizeof(float) * (81 * 32 * 2 + 32)));
  h->nc_w2p = h->nc_w1p + 81 * 32;
  h->nc_b1p = h->nc_w2p + 81 * 32;
  P2P_CUDA_OK(cudaMemcpy(h->nc_w1p, w1p.data(), sizeof(float) * 81 * 32, cudaMemcpyHostToDevice));
  P2P_CUDA_OK(cudaMemcpy(h->nc_w2p, w2p.data(), sizeof(float) * 81 * 32, cudaMemcpyHostToDevice));
  P2P_CUDA_OK(cudaMemcpy(h->nc_b1p, b1p.data(), sizeof(float) * 32, cudaMemcpyHostToDevice));
  h->nc_b2 = b2[0];
  {
    int rc = nc_umma_pack(w1p.data(), b1p.data(), w2p.data(), h->ncw);
    if (rc) return rc;
  }
  h->nc_set = true;
  return 0;
}

int p2p_set_regressor_weights(p2p_handle_t h, int which, const p2p_regressor_weights_t* w) {
  P2P_ENTER(h);
  P2P_REQUIRE(which == 0 || which == 1, "which must be 0 (mid) or 1 (fine)");
  P2P_REQUIRE(w != nullptr && w->conv0_weight && w->conv2_weight && w->fc0_weight && w->fc3_weight && w->fc6_weight,
              "null weight pointer");
  return pack_regressor(h, h->reg[which], *w);
}

static int* option_slot(p2p_handle_t h, const char* key) {
  if (!strcmp(key, "mid_passes")) return &h->opt_mid_passes;
  if (!strcmp(key, "fine_passes")) return &h->opt_fine_passes;
  if (!strcmp(key, "corr_passes")) return &h->opt_corr_passes;
  if (!strcmp(key, "seg_len")) return &h->opt_seg_len;
  if (!strcmp(key, "gemm_impl")) return &h->opt_gemm_impl;
  if (!strcmp(key, "num_sms")) return &h->opt_num_sms;
  if (!strcmp(key, "profile")) return &h->opt_profile;
  if (!strcmp(key, "mid_band")) return &h->opt_mid_band;
  if (!strcmp(key, "fuse_gather")) return &h->opt_fuse_gather;
  if (!strcmp(key, "fc_impl")) return &h->opt_fc_impl;
  if (!strcmp(key, "gemm_pair")) return &h->opt_gemm_pair;
  if (!strcmp(key, "nc_impl")) return &h->opt_nc_impl;
  if (!strcmp(key, "nc_l2_mode")) return &h->opt_nc_l2_mode;
  if (!strcmp(key, "unique_impl")) return &h->opt_unique_impl;
  return nullptr;
}

int p2p_set_option(p2p_handle_t h, const char* key, int value) {
  P2P_REQUIRE(h != nullptr && key != nullptr, "null argument");
  int* s = option_slot(h, key);
  P2P_REQUIRE(s != nullptr, std::string("unknown option ") + key);
  if (s == &h->opt_mid_passes || s == &h->opt_fine_passes) P2P_REQUIRE(value == 1 || value == 3, "passes must be 1 or 3");
  if (s == &h->opt_corr_passes) P2P_REQUIRE(value == 0 || value == 1 || value == 3, "corr_passes must be 0, 1 or 3");
  P2P_REQUIRE(value >= 0, "option values are non-negative");
  *s = value;
  return 0;
}

int p2p_get_option(p2p_handle_t h, const char* key, int* value) {
  P2P_REQUIRE(h != nullptr && key != nullptr && value != nullptr, "null argument");
  if (!strcmp(key, "band_rows")) {  // rows re-computed 3-pass by the last mid-stage call (synchronises)
    *value = 0;
    if (h->last_band_count != nullptr) {
      DeviceGuard g(h->device);
      P2P_CUDA_OK(cudaDeviceSynchronize());
      P2P_CUDA_OK(cudaMemcpy(value, h->last_band_count, sizeof(int), cudaMemcpyDeviceToHost));
    }
    return 0;
  }
  if (!strcmp(key, "band_rows_total") || !strcmp(key, "band_calls_rows_total")) {
    // running totals since the last read of "band_calls_rows_total" (synchronises): band rows / all mid-stage rows
    *value = 0;
    if (h->band_totals != nullptr) {
      DeviceGuard g(h->device);
      P2P_CUDA_OK(cudaDeviceSynchronize());
      unsigned long long t[2] = {0, 0};
      P2P_CUDA_OK(cudaMemcpy(t, h->band_totals, sizeof(t), cudaMemcpyDeviceToHost));
      const bool rows = !strcmp(key, "band_calls_rows_total");
      *value = (int)(rows ? t[1] : t[0]);
      if (rows) P2P_CUDA_OK(cudaMemset(h->band_totals, 0, sizeof(t)));
    }
    return 0;
  }
  int* s = option_slot(h, key);
  P2P_REQUIRE(s != nullptr, std::string("unknown option ") + key);
  *value = *s;
  return 0;
}

int p2p_launch_count(p2p_handle_t h, long long* count) {
  P2P_REQUIRE(h != nullptr && count != nullptr, "null argument");
  *count = g_launch_count;
  return 0;
}

int p2p_profile_read(p2p_handle_t h, float* ms_by_kind, int* count_by_kind, int nkinds) {
  P2P_ENTER(h);
  P2P_REQUIRE(ms_by_kind && count_by_kind && nkinds >= P2P_PROF_KINDS, "need room for P2P_PROF_KINDS entries");
  for (int i = 0; i < nkinds; ++i) {
    ms_by_kind[i] = 0.f;
    count_by_kind[i] = 0;
  }
  for (auto& e : h->prof) {
    P2P_CUDA_OK(cudaEventSynchronize(e.b));
    float ms = 0.f;
    P2P_CUDA_OK(cudaEventElapsedTime(&ms, e.a, e.b));
    ms_by_kind[e.kind] += ms;
    count_by_kind[e.kind] += 1;
    h->event_pool.push_back(e.a);
    h->event_pool.push_back(e.b);
  }
  h->prof.clear();
  return 0;
}

// -------------------------------------------------------------------------------------------------
// coarse
// -------------------------------------------------------------------------------------------------
static int corr_umma(p2p_handle_s* h, const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo,
                     int C, int n1, int n2, int n1pad, int n2pad, int ksize, float* out, uint8_t* code,
                     cudaStream_t st) {
  UmmaGemmParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t ad[5] = {(uint64_t)C, 1, 1, 1, (uint64_t)n1pad};
  const uint64_t as[4] = {(uint64_t)C * 2, (uint64_t)C * 2, (uint64_t)C * 2, (uint64_t)C * 2};
  const uint32_t ab[5] = {64, 1, 1, 1, 128};
  const uint64_t bd[2] = {(uint64_t)C, (uint64_t)n2pad};
  const uint64_t bs[1] = {(uint64_t)C * 2};
  p.pair = (h->opt_gemm_pair & 8) ? 1 : 0;
  const uint32_t bb[2] = {64, p.pair ? 128u : 256u};
  int rc;
  if ((rc = make_tmap_fp16(&p.a_main_hi, a_hi, 5, ad, as, ab))) return rc;
  if ((rc = make_tmap_fp16(&p.b_hi, b_hi, 2, bd, bs, bb))) return rc;
  if (h->opt_corr_passes == 3) {
    if ((rc = make_tmap_fp16(&p.a_main_lo, a_lo, 5, ad, as, ab))) return rc;
    if ((rc = make_tmap_fp16(&p.b_lo, b_lo, 2, bd, bs, bb))) return rc;
  }
  p.a_rgb_hi = p.a_main_hi;
  p.a_rgb_lo = p.a_main_hi;
  p.nsteps = C / 64;
  for (int s = 0; s < p.nsteps; ++s) p.steps[s] = KStep{(short)(s * 64), 0, 0, 0, 0, 0, s * 64};
  p.m_tiles = n1pad / 128;
  p.n_tiles = n2pad / 256;
  p.a_units_per_tile = 128;
  p.seg_len = h->opt_seg_len > 0 ? 1 : 0;
  p.epi.c = out;
  p.epi.alpha = 1.f / (kActScale * kActScale);
  if (ksize == 2) {
    p.epi.code = code;
    p.epi.np1 = n1 / 4;
    p.epi.np2 = n2 / 4;
    return launch_umma_gemm(p, EPI_CORR, h->opt_corr_passes, sms(h), st);
  }
  p.epi.ldc = n2;
  p.epi.m_rows = n1;
  p.epi.n_cols = n2;
  return launch_umma_gemm(p, EPI_PLAIN, h->opt_corr_passes, sms(h), st);
}

static int coarse_impl(p2p_handle_t h, const float* feat1, const float* feat2, int fmt, int c, int h1, int w1, int h2, int w2,
                       int ksize, float* corr4d_out, uint8_t* delta_code_out, float* pooled_out, float* ncn_out,
                       void* stream) {
  P2P_ENTER(h);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  P2P_REQUIRE(h->nc_set, "p2p_set_ncn_weights has not been called");
  P2P_REQUIRE(feat1 && feat2 && corr4d_out, "null tensor pointer");
  P2P_REQUIRE(ksize == 1 || ksize == 2, "ksize must be 1 or 2");
  P2P_REQUIRE(c > 0 && h1 > 0 && w1 > 0 && h2 > 0 && w2 > 0, "empty feature map");
  if (ksize == 2) {
    P2P_REQUIRE(h1 % 2 == 0 && w1 % 2 == 0 && h2 % 2 == 0 && w2 % 2 == 0, "ksize 2 needs even feature sizes");
    P2P_REQUIRE(delta_code_out != nullptr, "delta_code_out is required for ksize 2");
  }
  const int n1 = h1 * w1, n2 = h2 * w2;
  const int hA = h1 / ksize, wA = w1 / ksize, hB = h2 / ksize, wB = w2 / ksize;
  const int nA = hA * wA, nB = hB * wB;
  const size_t V = (size_t)nA * nB;
  const bool tc = h->opt_corr_passes > 0;
  if (tc) P2P_REQUIRE(c % 64 == 0 && c / 64 <= kMaxKSteps, "tensor-core correlation needs C % 64 == 0");
  P2P_REQUIRE(fmt == 0 || tc, "channels-last fp16 features need the tensor-core correlation (corr_passes 1 or 3)");
  const int n1pad = (int)align_up(n1, 128), n2pad = (int)align_up(n2, 256);
  size_t need = 4 * V * 4 + nc_umma_scratch_bytes(V) + nc_umma_xp_bytes(hA, wA, hB, wB) + (size_t)(nA + nB) * 8 + (1 << 16);
  need += tc ? (size_t)(n1pad + n2pad) * c * 4 : (size_t)(n1 + n2) * c * 4;
  int rc = h->coarse.reserve(need);
  if (rc) return rc;
  Arena& A = h->coarse;
  float* pooled = pooled_out ? pooled_out : (float*)A.take(V * 4);
  float* m1 = (float*)A.take(V * 4);
  float* nc = ncn_out ? ncn_out : (float*)A.take(V * 4);
  float* hidden = (float*)A.take(V * 128);               // fp32 [nA][32][nB] (nc_impl 0) or fp16 hi/lo [V][64] (nc_impl 1)
  float* partial = (float*)A.take(18 * V * 4);
  uint32_t* xp = (uint32_t*)A.take(nc_umma_xp_bytes(hA, wA, hB, wB));
  float* rowmax = (float*)A.take((size_t)nA * 4);
  unsigned int* colmax = (unsigned int*)A.take((size_t)nB * 4 + 16);
  unsigned int* xmax = colmax != nullptr ? colmax + nB : nullptr;
  if (tc) {
    __half* a_hi = (__half*)A.take((size_t)n1pad * c * 2);
    __half* a_lo = (__half*)A.take((size_t)n1pad * c * 2);
    __half* b_hi = (__half*)A.take((size_t)n2pad * c * 2);
    __half* b_lo = (__half*)A.take((size_t)n2pad * c * 2);
    P2P_REQUIRE(a_hi && a_lo && b_hi && b_lo && hidden && colmax, "scratch carve failed");
    const bool lo = h->opt_corr_passes == 3;
    {
      ProfScope ps(h, P2P_PROF_L2NORM, st);
      if (fmt == 1)
        rc = launch_l2norm_perm_kmajor_pair_nhwc16(reinterpret_cast<const __half*>(feat1), reinterpret_cast<const __half*>(feat2),
                                                   a_hi, lo ? a_lo : nullptr, b_hi, lo ? b_lo : nullptr, c, h1, w1, h2, w2, ksize, st);
      else
        rc = launch_l2norm_perm_kmajor_pair(feat1, feat2, a_hi, lo ? a_lo : nullptr, b_hi, lo ? b_lo : nullptr, c, h1, w1, h2,
                                            w2, ksize, st);
      if (rc) return rc;
    }
    ProfScope ps(h, P2P_PROF_CORR, st);
    if ((rc = corr_umma(h, a_hi, a_lo, b_hi, b_lo, c, n1, n2, n1pad, n2pad, ksize, pooled, delta_code_out, st)))
      return rc;
  } else {
    float* fa = (float*)A.take((size_t)n1 * c * 4);
    float* fb = (float*)A.take((size_t)n2 * c * 4);
    P2P_REQUIRE(fa && fb && hidden && colmax, "scratch carve failed");
    {
      ProfScope ps(h, P2P_PROF_L2NORM, st);
      if ((rc = launch_l2norm_perm(feat1, fa, c, h1, w1, ksize, st))) return rc;
      if ((rc = launch_l2norm_perm(feat2, fb, c, h2, w2, ksize, st))) return rc;
    }
    ProfScope ps(h, P2P_PROF_CORR, st);
    if ((rc = launch_corr_pool_simt(fa, fb, c, n1, n2, ksize, pooled, delta_code_out, st))) return rc;
  }
  P2P_REQUIRE(partial != nullptr && xmax != nullptr && xp != nullptr, "scratch carve failed");
  if (h->opt_nc_impl == 1) {
    {
      ProfScope ps(h, P2P_PROF_MUTUAL, st);
      if ((rc = launch_mutual_matching(pooled, nA, nB, rowmax, colmax, m1, xmax, st))) return rc;
    }
    {
      // layer 1, layer 2 (tensor cores) and the combine pass, which also yields the maxima of the second MutualMatching
      ProfScope ps(h, P2P_PROF_NC, st);
      if ((rc = launch_neigh_consensus_umma(m1, hA, wA, hB, wB, h->ncw, h->nc_b1p, h->nc_b2, xmax, xp, (__half*)hidden, partial,
                                            nc, rowmax, colmax, h->opt_nc_l2_mode, sms(h), st)))
        return rc;
    }
    ProfScope ps(h, P2P_PROF_MUTUAL, st);
    return launch_mutual_apply(nc, nA, nB, rowmax, colmax, corr4d_out, nullptr, st);
  }
  {
    ProfScope ps(h, P2P_PROF_MUTUAL, st);
    if ((rc = launch_mutual_matching(pooled, nA, nB, rowmax, colmax, m1, nullptr, st))) return rc;
  }
  {
    ProfScope ps(h, P2P_PROF_NC, st);
    if ((rc = launch_neigh_consensus(m1, hA, wA, hB, wB, h->nc_w1p, h->nc_b1p, h->nc_w2p, h->nc_b2, hidden, nc, st)))
      return rc;
  }
  ProfScope ps(h, P2P_PROF_MUTUAL, st);
  if ((rc = launch_mutual_matching(nc, nA, nB, rowmax, colmax, corr4d_out, nullptr, st))) return rc;
  return 0;
}

int p2p_coarse(p2p_handle_t h, const float* feat1, const float* feat2, int c, int h1, int w1, int h2, int w2,
               int ksize, float* corr4d_out, uint8_t* delta_code_out, float* pooled_out, float* ncn_out,
               void* stream) {
  return coarse_impl(h, feat1, feat2, 0, c, h1, w1, h2, w2, ksize, corr4d_out, delta_code_out, pooled_out, ncn_out, stream);
}

int p2p_coarse_nhwc16(p2p_handle_t h, const void* feat1_nhwc16, const void* feat2_nhwc16, int c, int h1, int w1, int h2, int w2,
                      int ksize, float* corr4d_out, uint8_t* delta_code_out, float* pooled_out, float* ncn_out,
                      void* stream) {
  return coarse_impl(h, reinterpret_cast<const float*>(feat1_nhwc16), reinterpret_cast<const float*>(feat2_nhwc16), 1, c, h1, w1,
                     h2, w2, ksize, corr4d_out, delta_code_out, pooled_out, ncn_out, stream);
}

int p2p_delta_unpack(p2p_handle_t h, const uint8_t* code, long long n, int ksize, int64_t* di, int64_t* dj,
                     int64_t* dk, int64_t* dl, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(code && di && dj && dk && dl && n >= 0 && ksize >= 1 && ksize <= 3, "bad argument");
  if (n == 0) return 0;
  return launch_delta_unpack(code, (size_t)n, ksize, (long long*)di, (long long*)dj, (long long*)dk, (long long*)dl,
                             reinterpret_cast<cudaStream_t>(stream));
}

int p2p_delta_pack(p2p_handle_t h, const int64_t* di, const int64_t* dj, const int64_t* dk, const int64_t* dl,
                   long long n, int ksize, uint8_t* code, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(code && di && dj && dk && dl && n >= 0 && ksize >= 1 && ksize <= 3, "bad argument");
  if (n == 0) return 0;
  return launch_delta_pack((const long long*)di, (const long long*)dj, (const long long*)dk, (const long long*)dl,
                           (size_t)n, ksize, code, reinterpret_cast<cudaStream_t>(stream));
}

int p2p_mutual_matching(p2p_handle_t h, const float* in, int nA, int nB, float* out, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(in && out && nA > 0 && nB > 0, "bad argument");
  int rc = h->misc.reserve((size_t)(nA + nB) * 4 + 4096);
  if (rc) return rc;
  float* rowmax = (float*)h->misc.take((size_t)nA * 4);
  unsigned int* colmax = (unsigned int*)h->misc.take((size_t)nB * 4);
  return launch_mutual_matching(in, nA, nB, rowmax, colmax, out, nullptr, reinterpret_cast<cudaStream_t>(stream));
}

int p2p_neigh_consensus(p2p_handle_t h, const float* in, int hA, int wA, int hB, int wB, float* out, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(h->nc_set, "p2p_set_ncn_weights has not been called");
  P2P_REQUIRE(in && out && hA > 0 && wA > 0 && hB > 0 && wB > 0, "bad argument");
  const size_t V = (size_t)hA * wA * hB * wB;
  int rc = h->misc.reserve(nc_umma_scratch_bytes(V) + nc_umma_xp_bytes(hA, wA, hB, wB) + 8192);
  if (rc) return rc;
  float* hidden = (float*)h->misc.take(V * 128);
  float* partial = (float*)h->misc.take(18 * V * 4);
  uint32_t* xp = (uint32_t*)h->misc.take(nc_umma_xp_bytes(hA, wA, hB, wB));
  unsigned int* xmax = (unsigned int*)h->misc.take(16);
  P2P_REQUIRE(hidden && partial && xmax && xp, "scratch carve failed");
  h->dbg_nc[0] = hidden; h->dbg_nc[1] = partial; h->dbg_nc[2] = xp; h->dbg_nc[3] = xmax;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (h->opt_nc_impl == 1) {
    if ((rc = launch_absmax(in, V, xmax, st))) return rc;
    return launch_neigh_consensus_umma(in, hA, wA, hB, wB, h->ncw, h->nc_b1p, h->nc_b2, xmax, xp, (__half*)hidden, partial, out,
                                       nullptr, nullptr, h->opt_nc_l2_mode, sms(h), st);
  }
  return launch_neigh_consensus(in, hA, wA, hB, wB, h->nc_w1p, h->nc_b1p, h->nc_w2p, h->nc_b2, hidden, out, st);
}

// Development hook (tools/nc_debug.py; not part of include/p2p_b200.h): copies `bytes` of an intermediate of the last
// p2p_neigh_consensus call to the host -- which = 0 hidden [V][64] fp16, 1 partial [18][V] f32, 2 xp (padded hi|lo
// words), 3 xmax word.
P2P_API int p2p_debug_nc_scratch(p2p_handle_t h, int which, void* host_dst, size_t bytes) {
  P2P_ENTER(h);
  P2P_REQUIRE(which >= 0 && which < 4 && host_dst != nullptr && h->dbg_nc[which] != nullptr, "bad argument");
  P2P_CUDA_OK(cudaDeviceSynchronize());
  P2P_CUDA_OK(cudaMemcpy(host_dst, h->dbg_nc[which], bytes, cudaMemcpyDeviceToHost));
  return 0;
}

int p2p_proposals(p2p_handle_t h, const float* corr4d, const uint8_t* delta_code, int hA, int wA, int hB, int wB,
                  int ksize, int upsample, int center, int do_softmax, int64_t* matches_out, float* scores_out,
                  void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(corr4d && matches_out && scores_out, "null tensor pointer");
  P2P_REQUIRE(hA > 0 && wA > 0 && hB > 0 && wB > 0 && ksize >= 1 && ksize <= 3 && upsample > 0, "bad dims");
  P2P_REQUIRE(ksize == 1 || delta_code != nullptr, "delta_code is required for ksize > 1");
  ProfScope ps(h, P2P_PROF_PROPOSALS, reinterpret_cast<cudaStream_t>(stream));
  return launch_proposals(corr4d, ksize > 1 ? delta_code : nullptr, hA, wA, hB, wB, ksize, upsample, center,
                          do_softmax, (long long*)matches_out, scores_out, reinterpret_cast<cudaStream_t>(stream));
}

int p2p_unique_rows(p2p_handle_t h, const int64_t* rows, int n, int mutual, const float* scores, float thres,
                    int32_t* ids_out, int32_t* count_out, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(rows && ids_out && count_out, "null tensor pointer");
  unsigned char* scratch = nullptr;
  const size_t sb = unique_rows_scratch_bytes(n);
  if (sb > 0) {   // candidate lists beyond 16384 rows sort in global scratch (same kernel, same result)
    int rc = h->uniq.reserve(sb + 4096);
    if (rc) return rc;
    scratch = (unsigned char*)h->uniq.take(sb);
  }
  if (h->uniq_rank == nullptr && h->opt_unique_impl == 1) {     // zeroed once; the kernel leaves it zeroed
    P2P_CUDA_OK(cudaMalloc(&h->uniq_rank, unique_rank_scratch_bytes()));
    P2P_CUDA_OK(cudaMemset(h->uniq_rank, 0, unique_rank_scratch_bytes()));
  }
  ProfScope ps(h, P2P_PROF_PROPOSALS, reinterpret_cast<cudaStream_t>(stream));
  return launch_unique_rows((const long long*)rows, n, mutual, scores, thres, ids_out, count_out, scratch,
                            h->opt_unique_impl == 1 ? h->uniq_rank : nullptr, reinterpret_cast<cudaStream_t>(stream));
}

int p2p_select_anchor(p2p_handle_t h, const int64_t* rows, const float* scores, const int32_t* ids, const int32_t* sel,
                      int m, int panc, int pshift, int64_t* matches_out, float* scores_out, int64_t* anchors_out,
                      void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(m >= 0 && rows != nullptr, "bad argument");
  P2P_REQUIRE(panc == 1 || panc == 8, "panc must be 1 or 8 (networks/patch2pix.py:377-402)");
  P2P_REQUIRE(scores_out == nullptr || scores != nullptr, "scores_out needs scores");
  P2P_REQUIRE(panc == 1 || anchors_out != nullptr, "anchors_out is required for panc 8");
  ProfScope ps(h, P2P_PROF_PROPOSALS, reinterpret_cast<cudaStream_t>(stream));
  return launch_select_anchor((const long long*)rows, scores, ids, sel, m, panc, pshift, (long long*)matches_out,
                              scores_out, (long long*)anchors_out, reinterpret_cast<cudaStream_t>(stream));
}

// -------------------------------------------------------------------------------------------------
// refine
// -------------------------------------------------------------------------------------------------
static int refine_prepare_impl(p2p_handle_t h, const float* const* feats1, const float* const* feats2, int fmt, int H1, int W1,
                               int H2, int W2, void* stream) {
  P2P_ENTER(h);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  P2P_REQUIRE(feats1 && feats2, "null feature list");
  for (int l = 0; l < 4; ++l) P2P_REQUIRE(feats1[l] && feats2[l], "null feature level");
  P2P_REQUIRE(H1 >= 8 && W1 >= 8 && H2 >= 8 && W2 >= 8 && H1 % 8 == 0 && W1 % 8 == 0 && H2 % 8 == 0 && W2 % 8 == 0,
              "image sizes must be positive multiples of 8");
  const int Hs[2] = {H1, H2}, Ws[2] = {W1, W2};
  size_t need = 1 << 16;
  for (int s = 0; s < 2; ++s) {
    const size_t px = (size_t)Hs[s] * Ws[s];
    need += (px / 4 * 64 + px / 16 * 64 + px / 64 * 128) * 6 + (px + px / 4 + px / 16 + px / 64) * 4 + 32768;
  }
  const bool want_map = h->opt_fuse_gather == 3;
  if (want_map)
    for (int s = 0; s < 2; ++s) need += (size_t)(Hs[s] + 2 * kMapPad) * (Ws[s] + 2 * kMapPad) * (512 + 8) + 4096;
  int rc = h->feat.reserve(need);
  if (rc) return rc;
  const int chans[3] = {64, 64, 128};
  for (int s = 0; s < 2; ++s) {
    PairFeatures& pf = h->pf[s];
    for (int l = 0; l < 4; ++l) {
      const int ds = 1 << l;
      pf.nsq[l] = (float*)h->feat.take((size_t)(Hs[s] / ds) * (Ws[s] / ds) * 4);
      P2P_REQUIRE(pf.nsq[l] != nullptr, "scratch carve failed");
    }
    for (int l = 0; l < 3; ++l) {
      const int ds = 2 << l;
      pf.nhwc[l] = (float*)h->feat.take((size_t)(Hs[s] / ds) * (Ws[s] / ds) * chans[l] * 4);
      pf.nhwc16[l] = (__half*)h->feat.take((size_t)(Hs[s] / ds) * (Ws[s] / ds) * chans[l] * 2);
      P2P_REQUIRE(pf.nhwc[l] != nullptr && pf.nhwc16[l] != nullptr, "scratch carve failed");
    }
    pf.wmap = pf.rgbn = nullptr;
    if (want_map) {
      const size_t pxp = (size_t)(Hs[s] + 2 * kMapPad) * (Ws[s] + 2 * kMapPad);
      pf.wmap = (__half*)h->feat.take(pxp * 512);
      pf.rgbn = (__half*)h->feat.take(pxp * 8);
      P2P_REQUIRE(pf.wmap != nullptr && pf.rgbn != nullptr, "scratch carve failed");
    }
  }
  {
    ProfScope ps(h, P2P_PROF_PREP, st);
    if ((rc = launch_feature_prep_pair(feats1, feats2, Hs, Ws, h->pf, fmt, st))) return rc;
    if (want_map && (rc = launch_window_map(h->pf, st))) return rc;
  }
  h->prepared = true;
  return 0;
}

int p2p_refine_prepare(p2p_handle_t h, const float* const* feats1, const float* const* feats2, int H1, int W1,
                       int H2, int W2, void* stream) {
  return refine_prepare_impl(h, feats1, feats2, 0, H1, W1, H2, W2, stream);
}

int p2p_refine_prepare_nhwc16(p2p_handle_t h, const void* const* feats1, const void* const* feats2, int H1, int W1,
                              int H2, int W2, void* stream) {
  return refine_prepare_impl(h, reinterpret_cast<const float* const*>(feats1), reinterpret_cast<const float* const*>(feats2), 1,
                             H1, W1, H2, W2, stream);
}

namespace {

struct RefineBuffers {
  __half *p_hi, *p_lo, *r_hi, *r_lo, *y_hi, *y_lo;
  __half *q_hi, *q_lo, *h1_hi, *h1_lo, *h2_hi, *h2_lo;   // tensor-core FC operands, rows padded to 128
  float *pooled, *raw;
  int *rowmap, *d_count;
  int npad;
};

// gather -> conv1 -> conv2 -> fc/parse for the rows selected by (rowmap, d_count) [all rows if null]
int run_regressor(p2p_handle_s* h, Regressor& R, int which, int passes, const RefineBuffers& B, const void* matches_in,
                  int is_float, int n, const int* rowmap, const int* d_count, float* matches_out, float* probs_out,
                  float* raw_out, cudaStream_t st) {
  const bool lo = passes == 3;
  const int kb = rowmap != nullptr ? P2P_PROF_GATHER_BAND : (which == 0 ? P2P_PROF_GATHER_MID : P2P_PROF_GATHER_FINE);
  int rc;
  const bool mapped = passes == 1 && rowmap == nullptr && h->opt_fuse_gather == 3 && h->opt_gemm_impl == 0;
  const bool fused = passes == 1 && rowmap == nullptr && h->opt_fuse_gather && h->opt_gemm_impl == 0 && !mapped;
  if (mapped) P2P_REQUIRE(h->pf[0].wmap != nullptr && h->pf[1].wmap != nullptr,
                          "fuse_gather = 3 needs p2p_refine_prepare to have run with the same option");
  if (!fused && !mapped) {
    ProfScope ps(h, kb, st);
    if ((rc = launch_patch_gather(h->pf[0], h->pf[1], matches_in, is_float, n, B.p_hi, lo ? B.p_lo : nullptr, B.r_hi,
                                  lo ? B.r_lo : nullptr, rowmap, d_count, st)))
      return rc;
  }
  if (h->opt_gemm_impl == 1) {
    P2P_REQUIRE(rowmap == nullptr, "the CUDA-core checker GEMM does not support row subsets");
    GemmOperands g1 = {B.p_hi, B.p_lo, B.r_hi, B.r_lo, R.w1_hi, R.w1_lo, 4, kConv1Steps * 64, n, passes, R.d_steps1, kConv1Steps};
    ConvEpilogue e1 = {R.scale1, R.bias1, 0, R.y_scale, B.y_hi, lo ? B.y_lo : nullptr, nullptr};
    {
      ProfScope ps(h, kb + 1, st);
      if ((rc = launch_conv_gemm_simt(g1, e1, st))) return rc;
    }
    ProfScope ps(h, kb + 2, st);
    GemmOperands g2 = {B.y_hi, B.y_lo, nullptr, nullptr, R.w2_hi, R.w2_lo, 1, kConv2Steps * 64, n, passes, R.d_steps2, kConv2Steps};
    ConvEpilogue e2 = {R.scale2, R.bias2, 1, 1.f, nullptr, nullptr, B.pooled};
    if ((rc = launch_conv_gemm_simt(g2, e2, st))) return rc;
  } else {
    UmmaGemmParams p;
    memset(&p, 0, sizeof(p));
    const uint32_t abox[5] = {64, 8, 8, 1, 2};
    const int pair = fused ? ((h->opt_gemm_pair & 32) && h->opt_fuse_gather == 1 ? 1 : 0)      // 32: fused conv1
                           : ((h->opt_gemm_pair & (lo ? 2 : 1)) ? 1 : 0);
    const int pair2 = (h->opt_gemm_pair & (lo ? 2 : 1)) ? 1 : 0;       // conv2 never runs fused
    uint32_t bbox[2] = {64, pair ? 128u : 256u};
    const uint64_t npad = (uint64_t)B.npad;
    {  // conv1
      const uint64_t ad[5] = {512, 8, 8, 4, npad};
      const uint64_t as[4] = {1024, 8192, 65536, 262144};
      const uint64_t rd[5] = {64, 8, 8, 1, npad};
      const uint64_t rs[4] = {128, 1024, 8192, 8192};
      const uint64_t bd[2] = {(uint64_t)kConv1Steps * 64, 512};
      const uint64_t bs[1] = {(uint64_t)kConv1Steps * 64 * 2};
      if ((rc = make_tmap_fp16(&p.a_main_hi, B.p_hi, 5, ad, as, abox))) return rc;
      if ((rc = make_tmap_fp16(&p.a_rgb_hi, B.r_hi, 5, rd, rs, abox))) return rc;
      if ((rc = make_tmap_fp16(&p.b_hi, R.w1_hi, 2, bd, bs, bbox))) return rc;
      if ((rc = make_tmap_fp16(&p.a_main_lo, lo ? B.p_lo : B.p_hi, 5, ad, as, abox))) return rc;
      if ((rc = make_tmap_fp16(&p.a_rgb_lo, lo ? B.r_lo : B.r_hi, 5, rd, rs, abox))) return rc;
      if ((rc = make_tmap_fp16(&p.b_lo, R.w1_lo, 2, bd, bs, bbox))) return rc;
      p.pair = pair;
      p.nsteps = kConv1Steps;
      memcpy(p.steps, R.steps1, sizeof(R.steps1));
      p.m_tiles = (n + 1) / 2;
      p.n_tiles = 2;
      p.a_units_per_tile = 2;
      p.seg_len = lo ? h->opt_seg_len : 0;
      p.d_units = d_count;
      p.epi.scale = R.scale1;
      p.epi.bias = R.bias1;
      p.epi.y_scale = R.y_scale;
      p.epi.y_hi = B.y_hi;
      p.epi.y_lo = lo ? B.y_lo : nullptr;
      p.epi.pooled = B.pooled;       // conv1's epilogue zeroes the max-pool accumulator conv2 merges into
      p.epi.n_patches = n;
      if (fused) {
        for (int s2 = 0; s2 < 2; ++s2) {
          p.fg.img[s2] = h->pf[s2].img;
          for (int l = 0; l < 3; ++l) p.fg.nhwc16[s2][l] = h->pf[s2].nhwc16[l];
          for (int l = 0; l < 4; ++l) p.fg.nsq[s2][l] = h->pf[s2].nsq[l];
          p.fg.H[s2] = h->pf[s2].H;
          p.fg.W[s2] = h->pf[s2].W;
        }
        p.fg.matches = matches_in;
        p.fg.is_float = is_float;
        p.fg.generation = h->opt_fuse_gather == 2 ? 1 : 2;
      }
      ProfScope ps(h, kb + 1, st);
      if (mapped) {
        Conv1TmaParams q;
        memset(&q, 0, sizeof(q));
        const uint32_t wbox[2] = {64, 128};
        if ((rc = make_tmap_fp16(&q.b_hi, R.w1_hi, 2, bd, bs, wbox))) return rc;
        for (int s2 = 0; s2 < 2; ++s2) {
          const uint64_t Wp = (uint64_t)h->pf[s2].W + 2 * kMapPad, Hp = (uint64_t)h->pf[s2].H + 2 * kMapPad;
          const uint64_t md[3] = {256, Wp, Hp};
          const uint64_t ms[2] = {512, Wp * 512};
          const uint32_t mb[3] = {64, 16, 16};       // 8 elements at traversal stride 2 (box = N * stride)
          const uint32_t me[3] = {1, 2, 2};
          if ((rc = make_tmap_fp16(&q.wm.map[s2], h->pf[s2].wmap, 3, md, ms, mb, me))) return rc;
          q.wm.rgbn[s2] = h->pf[s2].rgbn;
          q.wm.H[s2] = h->pf[s2].H;
          q.wm.W[s2] = h->pf[s2].W;
        }
        q.wm.matches = matches_in;
        q.wm.is_float = is_float;
        q.nsteps = kConv1Steps;
        memcpy(q.steps, R.steps1, sizeof(R.steps1));
        q.m_tiles = (n + 1) / 2;
        q.epi = p.epi;
        if ((rc = launch_conv1_tma(q, sms(h), st))) return rc;
      } else if ((rc = launch_umma_gemm(p, EPI_CONV1, passes, sms(h), st, fused))) {
        return rc;
      }
    }
    {  // conv2
      const uint64_t ad[5] = {512, 8, 8, 1, npad};
      const uint64_t as[4] = {1024, 8192, 65536, 65536};
      const uint64_t bd[2] = {(uint64_t)kConv2Steps * 64, 512};
      const uint64_t bs[1] = {(uint64_t)kConv2Steps * 64 * 2};
      p.pair = pair2;
      bbox[1] = pair2 ? 128u : 256u;
      if ((rc = make_tmap_fp16(&p.a_main_hi, B.y_hi, 5, ad, as, abox))) return rc;
      if ((rc = make_tmap_fp16(&p.a_main_lo, lo ? B.y_lo : B.y_hi, 5, ad, as, abox))) return rc;
      if ((rc = make_tmap_fp16(&p.b_hi, R.w2_hi, 2, bd, bs, bbox))) return rc;
      if ((rc = make_tmap_fp16(&p.b_lo, R.w2_lo, 2, bd, bs, bbox))) return rc;
      p.a_rgb_hi = p.a_main_hi;
      p.a_rgb_lo = p.a_main_lo;
      p.nsteps = kConv2Steps;
      memcpy(p.steps, R.steps2, sizeof(R.steps2));
      p.epi.scale = R.scale2;
      p.epi.bias = R.bias2;
      p.epi.pooled = B.pooled;
      ProfScope ps(h, kb + 2, st);
      if ((rc = launch_umma_gemm(p, EPI_CONV2, passes, sms(h), st))) return rc;
    }
  }
  ProfScope ps(h, kb + 3, st);
  if (h->opt_fc_impl == 0 || h->opt_gemm_impl == 1)
    return launch_fc_parse(B.pooled, R.fc, matches_in, is_float, n, h->pf[0].W, h->pf[0].H, h->pf[1].W, h->pf[1].H,
                           matches_out, probs_out, raw_out, rowmap, d_count, st);
  // tensor-core FC: split -> Linear(512,512)+BN+ReLU -> Linear(512,256)+BN+ReLU (3-pass, segmented) -> Linear(256,5)+parse
  if ((rc = launch_pooled_split(B.pooled, n, B.q_hi, B.q_lo, d_count, st))) return rc;
  const uint64_t m128 = align_up(n, 128);
  const uint32_t abx[5] = {64, 1, 1, 1, 128};
  const uint32_t bbx[2] = {64, (h->opt_gemm_pair & 4) ? 128u : 256u};
  for (int layer = 0; layer < 2; ++layer) {
    UmmaGemmParams p;
    memset(&p, 0, sizeof(p));
    p.pair = (h->opt_gemm_pair & 4) ? 1 : 0;
    const int nout = layer == 0 ? 512 : 256;
    const uint64_t ad[5] = {512, 1, 1, 1, m128};
    const uint64_t as[4] = {1024, 1024, 1024, 1024};
    const uint64_t bd[2] = {512, (uint64_t)nout};
    const uint64_t bs[1] = {1024};
    const __half* a_hi = layer == 0 ? B.q_hi : B.h1_hi;
    const __half* a_lo = layer == 0 ? B.q_lo : B.h1_lo;
    if ((rc = make_tmap_fp16(&p.a_main_hi, a_hi, 5, ad, as, abx))) return rc;
    if ((rc = make_tmap_fp16(&p.a_main_lo, a_lo, 5, ad, as, abx))) return rc;
    if ((rc = make_tmap_fp16(&p.b_hi, layer == 0 ? R.f1_hi : R.f2_hi, 2, bd, bs, bbx))) return rc;
    if ((rc = make_tmap_fp16(&p.b_lo, layer == 0 ? R.f1_lo : R.f2_lo, 2, bd, bs, bbx))) return rc;
    p.a_rgb_hi = p.a_main_hi;
    p.a_rgb_lo = p.a_main_lo;
    p.nsteps = 8;
    for (int s2 = 0; s2 < 8; ++s2) p.steps[s2] = KStep{(short)(s2 * 64), 0, 0, 0, 0, 0, s2 * 64};
    p.m_tiles = (int)(m128 / 128);
    p.n_tiles = nout / 256;
    p.a_units_per_tile = 128;
    p.seg_len = 2;
    p.d_units = d_count;
    p.epi.scale = layer == 0 ? R.fa1 : R.fa2;
    p.epi.bias = layer == 0 ? R.fc.b1 : R.fc.b2;
    p.epi.y_scale = kFcActScale;
    p.epi.y_hi = layer == 0 ? B.h1_hi : B.h2_hi;
    p.epi.y_lo = layer == 0 ? B.h1_lo : B.h2_lo;
    p.epi.ldc = nout;
    p.epi.n_patches = n;
    if ((rc = launch_umma_gemm(p, EPI_FC, 3, sms(h), st))) return rc;
  }
  return launch_fc3_parse(B.h2_hi, B.h2_lo, R.fc.w3t, R.fc.b3, matches_in, is_float, n, h->pf[0].W, h->pf[0].H,
                          h->pf[1].W, h->pf[1].H, matches_out, probs_out, raw_out, rowmap, d_count, st);
}

}  // namespace

int p2p_refine(p2p_handle_t h, int which, const void* matches_in, int is_float, int n, float* matches_out,
               float* probs_out, void* stream) {
  P2P_ENTER(h);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  P2P_REQUIRE(which == 0 || which == 1, "which must be 0 (mid) or 1 (fine)");
  P2P_REQUIRE(h->reg[which].set, "p2p_set_regressor_weights has not been called for this regressor");
  P2P_REQUIRE(h->prepared, "p2p_refine_prepare has not been called");
  P2P_REQUIRE(n >= 0, "negative match count");
  if (n == 0) return 0;
  P2P_REQUIRE(matches_in && matches_out && probs_out, "null tensor pointer");
  Regressor& R = h->reg[which];
  const int passes = which == 0 ? h->opt_mid_passes : h->opt_fine_passes;
  // Risk band (mid stage only): 1-pass for every row, fp32-grade 3-pass re-computation only for the
  // rows whose coordinates sit within mid_band/1000 px of an integer (trunc() must match the reference).
  const bool band = which == 0 && passes == 3 && h->opt_mid_band > 0 && h->opt_gemm_impl == 0;
  const int npad = (int)align_up(n, 2);
  const size_t pbytes = (size_t)npad * kPatchPos * kMainCh * 2, rbytes = (size_t)npad * 4096 * 2,
               ybytes = (size_t)npad * 64 * 512 * 2, qbytes = (size_t)npad * 512 * 4;
  const size_t n128 = align_up(n, 128);
  int rc = h->refine.reserve(2 * (pbytes + rbytes + ybytes) + qbytes + (size_t)n * 24 + n128 * (512 + 512 + 256) * 4 +
                             (1 << 16));
  if (rc) return rc;
  Arena& A = h->refine;
  RefineBuffers B;
  B.npad = npad;
  B.p_hi = (__half*)A.take(pbytes);
  B.p_lo = (__half*)A.take(pbytes);
  B.r_hi = (__half*)A.take(rbytes);
  B.r_lo = (__half*)A.take(rbytes);
  B.y_hi = (__half*)A.take(ybytes);
  B.y_lo = (__half*)A.take(ybytes);
  B.pooled = (float*)A.take(qbytes);
  B.q_hi = (__half*)A.take(n128 * 512 * 2);
  B.q_lo = (__half*)A.take(n128 * 512 * 2);
  B.h1_hi = (__half*)A.take(n128 * 512 * 2);
  B.h1_lo = (__half*)A.take(n128 * 512 * 2);
  B.h2_hi = (__half*)A.take(n128 * 256 * 2);
  B.h2_lo = (__half*)A.take(n128 * 256 * 2);
  B.raw = (float*)A.take((size_t)n * 5 * 4);
  B.rowmap = (int*)A.take((size_t)n * 4 + 16);
  P2P_REQUIRE(B.p_hi && B.p_lo && B.r_hi && B.r_lo && B.y_hi && B.y_lo && B.pooled && B.raw && B.rowmap && B.q_hi &&
                  B.q_lo && B.h1_hi && B.h1_lo && B.h2_hi && B.h2_lo,
              "scratch carve failed");
  B.d_count = B.rowmap + n;
  if (which == 0) h->last_band_count = band ? B.d_count : nullptr;
  if (!band)
    return run_regressor(h, R, which, passes, B, matches_in, is_float, n, nullptr, nullptr, matches_out, probs_out,
                         nullptr, st);
  if ((rc = run_regressor(h, R, which, 1, B, matches_in, is_float, n, nullptr, nullptr, matches_out, probs_out, B.raw,
                          st)))
    return rc;
  {
    ProfScope ps(h, P2P_PROF_FLAG, st);
    if ((rc = launch_flag_risky(matches_in, is_float, B.raw, n, h->opt_mid_band * 1e-3f, 0.02f, h->pf[0].W, h->pf[0].H,
                                h->pf[1].W, h->pf[1].H, B.rowmap, B.d_count, h->band_totals, st)))
      return rc;
  }
  return run_regressor(h, R, which, 3, B, matches_in, is_float, n, B.rowmap, B.d_count, matches_out, probs_out,
                       nullptr, st);
}

int p2p_finalize_matches(p2p_handle_t h, const float* fine, const float* scores, const int64_t* coarse, int n, float io_thres,
                         const double* upscale4, double* packed_out, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(scores && coarse && upscale4 && packed_out && n >= 0, "bad argument");
  return launch_finalize_matches(fine, scores, (const long long*)coarse, n, io_thres, upscale4, packed_out,
                                 reinterpret_cast<cudaStream_t>(stream));
}

int p2p_preprocess_image(p2p_handle_t h, const uint8_t* rgb_hwc, int ho, int wo, int ht, int wt, float* out_chw,
                         uint8_t* resized_hwc_out, void* stream) {
  P2P_ENTER(h);
  P2P_REQUIRE(rgb_hwc && out_chw && ho > 0 && wo > 0 && ht > 0 && wt > 0, "bad argument");
  P2P_REQUIRE((long long)ho * wo < (1ll << 28) && (long long)ht * wt < (1ll << 28), "image too large");
  const PreprocessCoefs* C = nullptr;
  for (const auto& c : h->pre_coefs)
    if (c.ho == ho && c.wo == wo && c.ht == ht && c.wt == wt) C = &c;
  if (C == nullptr) {
    if (h->pre_coefs.size() >= 32) {            // drop the oldest table (nothing may still be reading it)
      P2P_CUDA_OK(cudaDeviceSynchronize());
      if (h->pre_coefs.front().d) cudaFree(h->pre_coefs.front().d);
      h->pre_coefs.erase(h->pre_coefs.begin());
    }
    PreprocessCoefs c;
    int rc = preprocess_build_coefs(ho, wo, ht, wt, c);
    if (rc) return rc;
    h->pre_coefs.push_back(c);
    C = &h->pre_coefs.back();
  }
  int rc = h->pre.reserve((size_t)ho * wt * 3 + 4096);
  if (rc) return rc;
  uint8_t* tmp = (uint8_t*)h->pre.take((size_t)ho * wt * 3);
  P2P_REQUIRE(tmp != nullptr, "scratch carve failed");
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};   // ImageNet, preprocess.py:93
  return launch_preprocess(rgb_hwc, *C, mean, stdv, out_chw, resized_hwc_out, tmp, reinterpret_cast<cudaStream_t>(stream));
}

int p2p_test_gemm(p2p_handle_t h, const float* a, const float* b, float* c, int M, int N, int K, int passes,
                  int seg_len, float in_scale, void* stream) {
  P2P_ENTER(h);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  P2P_REQUIRE(a && b && c && M > 0 && N > 0 && K > 0, "bad argument");
  P2P_REQUIRE(K % 64 == 0 && K / 64 <= kMaxKSteps, "K must be a multiple of 64 and at most 6144");
  P2P_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
  const int mpad = (int)align_up(M, 128), npad = (int)align_up(N, 256);
  const size_t ab = (size_t)mpad * K * 2, bb = (size_t)npad * K * 2;
  int rc = h->misc.reserve(2 * (ab + bb) + (1 << 16));
  if (rc) return rc;
  __half* a_hi = (__half*)h->misc.take(ab);
  __half* a_lo = (__half*)h->misc.take(ab);
  __half* b_hi = (__half*)h->misc.take(bb);
  __half* b_lo = (__half*)h->misc.take(bb);
  P2P_CUDA_OK(cudaMemsetAsync(a_hi, 0, 2 * ab, st));
  P2P_CUDA_OK(cudaMemsetAsync(b_hi, 0, 2 * bb, st));
  if ((rc = launch_split_rows(a, a_hi, a_lo, (size_t)M * K, in_scale, st))) return rc;
  if ((rc = launch_split_rows(b, b_hi, b_lo, (size_t)N * K, in_scale, st))) return rc;
  UmmaGemmParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t ad[5] = {(uint64_t)K, 1, 1, 1, (uint64_t)mpad};
  const uint64_t as[4] = {(uint64_t)K * 2, (uint64_t)K * 2, (uint64_t)K * 2, (uint64_t)K * 2};
  const uint32_t abx[5] = {64, 1, 1, 1, 128};
  const uint64_t bd[2] = {(uint64_t)K, (uint64_t)npad};
  const uint64_t bs[1] = {(uint64_t)K * 2};
  p.pair = (h->opt_gemm_pair & 16) ? 1 : 0;
  const uint32_t bbx[2] = {64, p.pair ? 128u : 256u};
  if ((rc = make_tmap_fp16(&p.a_main_hi, a_hi, 5, ad, as, abx))) return rc;
  if ((rc = make_tmap_fp16(&p.a_main_lo, a_lo, 5, ad, as, abx))) return rc;
  if ((rc = make_tmap_fp16(&p.b_hi, b_hi, 2, bd, bs, bbx))) return rc;
  if ((rc = make_tmap_fp16(&p.b_lo, b_lo, 2, bd, bs, bbx))) return rc;
  p.a_rgb_hi = p.a_main_hi;
  p.a_rgb_lo = p.a_main_lo;
  p.nsteps = K / 64;
  for (int s = 0; s < p.nsteps; ++s) p.steps[s] = KStep{(short)(s * 64), 0, 0, 0, 0, 0, s * 64};
  p.m_tiles = mpad / 128;
  p.n_tiles = npad / 256;
  p.a_units_per_tile = 128;
  p.seg_len = seg_len;
  p.epi.c = c;
  p.epi.ldc = N;
  p.epi.m_rows = M;
  p.epi.n_cols = N;
  p.epi.alpha = 1.f / (in_scale * in_scale);
  return launch_umma_gemm(p, EPI_PLAIN, passes, sms(h), st);
}

}  // extern "C"
