// xemo_net.cu -- graph-level entry points of libxemo.so (include/xemo.h, section C): whole networks behind the C ABI.
//
// What dagnn.DagNN.eval and cnn_train_dag do in the reference --
//   dag.eval({'data', faces})                      /root/reference/emoVoxCeleb/fetch_emovoxceleb_imdb.m:129,
//                                                  /root/reference/external/compute_visual_feats.m:90   (teacher, test mode)
//   dag.eval({'data', spec})                       /root/reference/external/compute_audio_feats.m:126    (student, test mode)
//   cnn_train_dag(net, imdb, getBatch, ...)        /root/reference/emoVoxCeleb/run_distillation.m:170-182 (student step:
//                                                  forward, loss emoVoxZoo.m:137-157, backward, accumulateGradients, with
//                                                  the gradient sum over the labs of 'gpus', opts.gpus :179-181)
// -- as single calls: the library owns the device-resident network (fp16 KRSC filters, fp32 master / momentum /
// gradient buffers, NHWC fp16 activations), sequences the xemo_op_* kernels, captures the sequence in CUDA graphs and
// replays them.  A host in any language (the MEX shim mex/xemo_dagnn_mex.c, the ctypes wrapper net.py) only moves
// parameters in, inputs in and logits / metrics out.  Data-parallel gradient exchange is ncclAllReduce on a communicator
// the library creates (NCCL is resolved at run time with dlsym: no link-time dependency), issued INSIDE the captured
// step on a forked stream so that the fc6..fc8 bucket travels while conv5..conv1 are differentiated.
#include "xemo_internal.h"

#include <cuda_fp16.h>
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

using namespace xemo;

// ------------------------------------------------------------------------------------------------ NCCL at run time
namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
struct NcclApi {
  bool ok = false;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    void* h = RTLD_DEFAULT;
    if (!dlsym(h, "ncclAllReduce")) {   // not in the process yet (a torch host has it loaded): try the shared object
      h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
      if (!h) return a;
    }
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce;
    return a;
  }();
  return api;
}
constexpr int kNcclFloat = 7, kNcclSum = 0;   // ncclFloat32, ncclSum (nccl.h)

}  // namespace

struct xemo_comm {
  xemo_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t side = nullptr;   // the forked stream the bucket all-reduces run on
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

extern "C" int xemo_comm_unique_id(void* id128) {
  if (!id128 || !nccl().ok) return XEMO_ERR_INVALID;
  NcclId id;
  if (nccl().GetUniqueId(&id) != 0) return XEMO_ERR_CUDA;
  memcpy(id128, &id, sizeof(id));
  return XEMO_OK;
}

extern "C" int xemo_comm_create(xemo_ctx* ctx, const void* id128, int rank, int world, xemo_comm** out) {
  XEMO_REQUIRE(ctx, ctx && id128 && out && world >= 1 && rank >= 0 && rank < world, "comm_create: bad arguments");
  XEMO_REQUIRE(ctx, nccl().ok, "comm_create: NCCL (libnccl.so.2) could not be resolved at run time");
  *out = nullptr;
  XEMO_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  xemo_comm* c = new xemo_comm();
  c->ctx = ctx; c->rank = rank; c->world = world;
  const int rc = nccl().CommInitRank(&c->comm, world, id, rank);
  if (rc != 0) {
    delete c;
    return fail(ctx, XEMO_ERR_CUDA, "ncclCommInitRank failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  }
  // highest priority: the all-reduce kernels must get SMs as soon as CTAs of the (persistent, all-SM) convolution kernels
  // retire, or the exchange trails the backward pass instead of hiding behind it; captured nodes inherit the priority
  int prio_lo = 0, prio_hi = 0;
  XEMO_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  XEMO_CUDA(ctx, cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi));
  XEMO_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  XEMO_CUDA(ctx, cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  *out = c;
  return XEMO_OK;
}

extern "C" void xemo_comm_destroy(xemo_comm* c) {
  if (!c) return;
  if (c->comm) nccl().CommDestroy(c->comm);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
}

extern "C" int xemo_comm_allreduce_f32(xemo_comm* c, float* buf, size_t n) {
  if (!c) return XEMO_ERR_INVALID;
  if (c->world == 1) return XEMO_OK;
  const int rc = nccl().AllReduce(buf, buf, n, kNcclFloat, kNcclSum, c->comm, c->ctx->stream);
  if (rc != 0) return fail(c->ctx, XEMO_ERR_CUDA, "ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
  return XEMO_OK;
}

// ------------------------------------------------------------------------------------------------ the network object
namespace {

constexpr float kBnEps = 1e-5f;   // dagnn.BatchNorm default
const int kTeacherStages[4][4] = {{3, 64, 256, 1}, {4, 128, 512, 2}, {6, 256, 1024, 2}, {3, 512, 2048, 2}};
const float kAverageImage[3] = {131.0912f, 103.8827f, 91.4953f};

struct ParamSpec { int64_t d[4]; };   // MatConvNet dims (FH x FW x FC x K; vectors K x 1; moments C x 2)

struct ConvLayer {   // student
  std::string name, bn;
  int fh, fw, cin, cout, kp, cp, sh, sw, pad[4], h, w, oh, ow;
  bool has_bn;
  bool full_height = false;  // a filter as tall as its input, one column wide, unpadded, stride 1 (fc6): dgrad as a GEMM
  int pool_method = -1;  // -1 none, 0 max, 1 avg
  int pwh = 0, pww = 0, psh = 1, psw = 1, ph = 0, pw = 0, poh = 0, pow_ = 0;
};

struct Block {   // teacher bottleneck
  std::string pre;
  int cin, mid, cout, stride;
  bool proj;
};

inline int pad16(int v) { return (v + 15) / 16 * 16; }
inline int out_dim(int h, int pt, int pb, int f, int s) { return (h + pt + pb - f) / s + 1; }

}  // namespace

struct xemo_net {
  xemo_ctx* ctx = nullptr;
  int kind = 0, N = 0, W = 0, input_mode = 0, face_size = 48, K = 8;
  bool finalized = false;
  std::map<std::string, ParamSpec> spec;
  std::vector<std::string> order;                       // parameter names in graph order
  std::map<std::string, std::vector<float>> host;       // values handed in before finalize (column-major MatConvNet)
  std::vector<void*> owned;
  std::map<std::string, void*> buf;                     // named device buffers (activations, folded vectors, filters)
  // teacher
  std::vector<Block> blocks;
  // SE blocks by linearity (the squeeze is linear in the 3x3 output: s = a3 (W3 mean_hw t2) + b3, so the gate is known before
  // the expand convolution runs and the excite folds into its epilogue; u is never written or re-read) on the stages whose
  // feature map is at least this wide -- measured on B200: a gain at 56 x 56 and 28 x 28, a loss at 14 x 14 and 7 x 7
  // (profiles/r02_ab_experimental_options.json).  XEMO_SE_LIN_MIN_HW overrides (0 disables).
  int se_lin_min_hw = 28;
  bool se_lin(int hw) const { return se_lin_min_hw > 0 && hw >= se_lin_min_hw; }
  // student
  std::vector<ConvLayer> layers;
  std::map<std::string, size_t> seg;                    // offset (floats) of each parameter inside the flat buffers
  size_t nparam = 0;
  float *master = nullptr, *momentum = nullptr, *grad = nullptr, *hyper = nullptr;
  __half* w16 = nullptr;
  int* guard = nullptr;
  int s2d_hp = 0, s2d_ow = 0, pool1_ld = 0;
  bool stem_pairs = true;
  float grad_scale = 1024.f;
  int loss_type = 0;
  float temperature = 2.f;
  size_t split_offset = 0;                              // first float of the fc6..fc8 gradient bucket
  int split_layer = 5;
  // graphs
  xemo_graph *g_fwd = nullptr, *g_train = nullptr, *g_update = nullptr, *g_step = nullptr;
  xemo_comm* g_train_comm = nullptr;
  xemo_comm* g_step_comm = nullptr;
  xemo_net* g_step_teacher = nullptr;
  const int *g_step_start = nullptr, *g_step_end = nullptr;
  int *win_start = nullptr, *win_end = nullptr;         // coupling windows set by xemo_distill_set_windows
  // concurrency inside the captured step (xemo_net_set_overlap): kernels of a small per-GPU batch leave most SMs idle, so
  // independent branches run side by side -- the teacher forward beside the student forward (they meet at the loss) and the
  // filter gradients beside the data-gradient chain
  float* bm_flat = nullptr;                             // batch moments of all BN layers, contiguous
  float* mom_flat = nullptr;                            // the moments parameters at the same offsets
  size_t bm_elems = 0;
  float bm_scale = 1.f;                                 // 1 / ranks once the batch moments are summed across ranks
  int overlap = -1;                                     // -1 auto (on for batch <= 64), 0 off, 1 on
  cudaStream_t side_a = nullptr, side_b = nullptr;
  bool use_overlap() const { return overlap < 0 ? N <= 64 : overlap != 0; }
  bool packs_ahead = false;                             // the dgrad filter packs were recorded beside the forward pass
  int g_step_mean = 0;

  template <typename T>
  T* alloc(const std::string& name, size_t n) {
    void* p = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    cudaMemsetAsync(p, 0, bytes, ctx->stream);
    owned.push_back(p);
    if (!name.empty()) buf[name] = p;
    return static_cast<T*>(p);
  }
  template <typename T>
  T* get(const std::string& name) { auto it = buf.find(name); return it == buf.end() ? nullptr : static_cast<T*>(it->second); }
};

namespace {

#define NET_OP(call)        \
  do {                      \
    const int rc_ = (call); \
    if (rc_) return rc_;    \
  } while (0)

void add_param(xemo_net* n, const std::string& name, int64_t a, int64_t b, int64_t c, int64_t d) {
  n->spec[name] = ParamSpec{{a, b, c, d}};
  n->order.push_back(name);
}
size_t numel_of(const ParamSpec& s) { return size_t(s.d[0]) * s.d[1] * s.d[2] * s.d[3]; }

// upload a host fp32 vector to a fresh named device buffer
float* upload_f32(xemo_net* n, const std::string& name, const std::vector<float>& v) {
  float* d = n->alloc<float>(name, v.size());
  if (d) cudaMemcpyAsync(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice, n->ctx->stream);
  return d;
}
__half* upload_f16(xemo_net* n, const std::string& name, const std::vector<float>& v) {
  std::vector<__half> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = __float2half_rn(v[i]);
  __half* d = n->alloc<__half>(name, v.size());
  if (d) {
    cudaMemcpyAsync(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice, n->ctx->stream);
    cudaStreamSynchronize(n->ctx->stream);   // `h` goes out of scope
  }
  return d;
}

// FH x FW x FC x K column-major -> [Kp][FH][FW][Cp] row-major, zero padded (the device layout "KRSC")
std::vector<float> krsc_host(const std::vector<float>& f, int FH, int FW, int FC, int K, int Kp, int Cp) {
  std::vector<float> o(size_t(Kp) * FH * FW * Cp, 0.f);
  for (int k = 0; k < K; ++k)
    for (int c = 0; c < FC; ++c)
      for (int s = 0; s < FW; ++s)
        for (int r = 0; r < FH; ++r)
          o[((size_t(k) * FH + r) * FW + s) * Cp + c] = f[r + size_t(FH) * (s + size_t(FW) * (c + size_t(FC) * k))];
  return o;
}
void unkrsc_host(const float* w, int FH, int FW, int FC, int K, int Cp, float* f) {
  for (int k = 0; k < K; ++k)
    for (int c = 0; c < FC; ++c)
      for (int s = 0; s < FW; ++s)
        for (int r = 0; r < FH; ++r)
          f[r + size_t(FH) * (s + size_t(FW) * (c + size_t(FC) * k))] = w[((size_t(k) * FH + r) * FW + s) * Cp + c];
}
// student conv1: 7 x 7 x 1 x K stride-2 filter <-> [K][4][1][16] filter over the space-to-depth input,
// G[k][j][0][dr*8 + s] = F[2j + dr, s, 0, k]
std::vector<float> conv1_to_s2d(const std::vector<float>& f, int K) {
  std::vector<float> g(size_t(K) * 64, 0.f);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < 4; ++j)
      for (int dr = 0; dr < 2; ++dr) {
        const int r = 2 * j + dr;
        if (r >= 7) continue;
        for (int s = 0; s < 7; ++s) g[(size_t(k) * 4 + j) * 16 + dr * 8 + s] = f[r + 7 * (s + 7 * size_t(k))];
      }
  return g;
}
void conv1_from_s2d(const float* g, int K, float* f) {
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < 4; ++j)
      for (int dr = 0; dr < 2; ++dr) {
        const int r = 2 * j + dr;
        if (r >= 7) continue;
        for (int s = 0; s < 7; ++s) f[r + 7 * (s + 7 * size_t(k))] = g[(size_t(k) * 4 + j) * 16 + dr * 8 + s];
      }
}

int conv(xemo_net* n, const void* x, int N, int h, int w, int cin, const void* wt, int kout, int r, int s, int sh, int sw,
         const int pad[4], const float* scale, const float* shift, const void* residual, int relu, void* out16, float* out32 = nullptr,
         int ldc = 0) {
  return xemo_op_conv_fwd(n->ctx, x, N, h, w, cin, wt, kout, r, s, sh, sw, pad[0], pad[1], pad[2], pad[3], scale, shift, residual, relu,
                          out16, out32, ldc);
}
const int kPad0[4] = {0, 0, 0, 0};
const int kPad1[4] = {1, 1, 1, 1};

// ================================================================================================ teacher
void teacher_describe(xemo_net* n) {
  add_param(n, "conv1f", 7, 7, 3, 64);
  add_param(n, "bn1m", 64, 1, 1, 1); add_param(n, "bn1b", 64, 1, 1, 1); add_param(n, "bn1x", 64, 2, 1, 1);
  int cin = 64;
  for (int si = 0; si < 4; ++si)
    for (int bi = 0; bi < kTeacherStages[si][0]; ++bi) {
      const int mid = kTeacherStages[si][1], cout = kTeacherStages[si][2], stride = bi == 0 ? kTeacherStages[si][3] : 1;
      char pre[32];
      snprintf(pre, sizeof(pre), "s%db%d_", si + 2, bi + 1);
      const std::string p(pre);
      auto conv_bn = [&](const char* c, const char* bn, int fh, int ci, int co) {
        add_param(n, p + c + "f", fh, fh, ci, co);
        add_param(n, p + bn + "m", co, 1, 1, 1); add_param(n, p + bn + "b", co, 1, 1, 1); add_param(n, p + bn + "x", co, 2, 1, 1);
      };
      conv_bn("c1", "bn1", 1, cin, mid);
      conv_bn("c2", "bn2", 3, mid, mid);
      conv_bn("c3", "bn3", 1, mid, cout);
      if (bi == 0) conv_bn("proj", "bnp", 1, cin, cout);
      if (n->kind == XEMO_NET_SENET50) {
        add_param(n, p + "se1f", 1, 1, cout, cout / 16); add_param(n, p + "se1b", cout / 16, 1, 1, 1);
        add_param(n, p + "se2f", 1, 1, cout / 16, cout); add_param(n, p + "se2b", cout, 1, 1, 1);
      }
      n->blocks.push_back(Block{p, cin, mid, cout, stride, bi == 0});
      cin = cout;
    }
  add_param(n, "classifierf", 1, 1, 2048, n->K);
  add_param(n, "classifierb", n->K, 1, 1, 1);
}

// test-mode BN as y = a x + b: a = g / sigma, b = beta - a mu (moments = [mu | sigma], column-major C x 2)
void fold_bn(xemo_net* n, const std::string& bn, int C, int reps, const std::string& key) {
  const std::vector<float>&g = n->host[bn + "m"], &beta = n->host[bn + "b"], &mom = n->host[bn + "x"];
  std::vector<float> a(size_t(C) * reps), b(size_t(C) * reps);
  for (int c = 0; c < C; ++c) {
    const double av = double(g[c]) / double(mom[C + c]);
    const double bv = double(beta[c]) - av * double(mom[c]);
    for (int r = 0; r < reps; ++r) { a[size_t(r) * C + c] = float(av); b[size_t(r) * C + c] = float(bv); }
  }
  upload_f32(n, key + ":a", a);
  upload_f32(n, key + ":b", b);
}

int teacher_finalize(xemo_net* n) {
  xemo_ctx* ctx = n->ctx;
  const int N = n->N;
  if (const char* e = getenv("XEMO_SE_LIN_MIN_HW")) n->se_lin_min_hw = atoi(e);
  // stem: row-im2col filter G[k][r][0][s*4 + c] = F[r, s, c, k] in pixel-pair (block-diagonal) form [128][7][1][64]
  {
    const std::vector<float>& f = n->host["conv1f"];
    std::vector<float> g2(size_t(128) * 7 * 64, 0.f);
    for (int k = 0; k < 64; ++k)
      for (int r = 0; r < 7; ++r)
        for (int s = 0; s < 7; ++s)
          for (int c = 0; c < 3; ++c) {
            const float v = f[r + 7 * (s + 7 * (c + 3 * size_t(k)))];
            for (int e = 0; e < 2; ++e) g2[((size_t(e) * 64 + k) * 7 + r) * 64 + e * 32 + s * 4 + c] = v;
          }
    upload_f16(n, "conv1:w", g2);
    fold_bn(n, "bn1", 64, 2, "conv1");
  }
  for (const Block& b : n->blocks) {
    auto conv_w = [&](const char* c, const char* bn, int fh, int ci, int co) {
      upload_f16(n, b.pre + c + ":w", krsc_host(n->host[b.pre + c + "f"], fh, fh, ci, co, pad16(co), pad16(ci)));
      fold_bn(n, b.pre + bn, co, 1, b.pre + c);
    };
    conv_w("c1", "bn1", 1, b.cin, b.mid);
    conv_w("c2", "bn2", 3, b.mid, b.mid);
    conv_w("c3", "bn3", 1, b.mid, b.cout);
    if (b.proj) conv_w("proj", "bnp", 1, b.cin, b.cout);
    if (n->kind == XEMO_NET_SENET50) {
      const int C = b.cout, Cr = C / 16;
      upload_f32(n, b.pre + "se1:w", n->host[b.pre + "se1f"]);   // 1 x 1 x C x Cr column-major == [Cr][C]
      upload_f32(n, b.pre + "se1:b", n->host[b.pre + "se1b"]);
      const std::vector<float>& f2 = n->host[b.pre + "se2f"];    // 1 x 1 x Cr x C column-major -> transposed [Cr][C]
      std::vector<float> t(size_t(Cr) * C);
      for (int c = 0; c < C; ++c)
        for (int j = 0; j < Cr; ++j) t[size_t(j) * C + c] = f2[j + size_t(Cr) * c];
      upload_f32(n, b.pre + "se2:w", t);
      upload_f32(n, b.pre + "se2:b", n->host[b.pre + "se2b"]);
    }
  }
  const int Kp = pad16(n->K);
  upload_f16(n, "classifier:w", krsc_host(n->host["classifierf"], 1, 1, 2048, n->K, Kp, 2048));
  std::vector<float> cb(Kp, 0.f);
  for (int k = 0; k < n->K; ++k) cb[k] = n->host["classifierb"][k];
  upload_f32(n, "classifier:b", cb);
  upload_f32(n, "mean3", std::vector<float>(kAverageImage, kAverageImage + 3));
  // activations
  if (n->input_mode == XEMO_INPUT_U8) n->alloc<uint8_t>("faces", size_t(N) * n->face_size * n->face_size);
  else n->alloc<float>("faces", size_t(N) * 3 * 224 * 224);
  n->alloc<__half>("rows", size_t(N) * 224 * 112 * 32);
  n->alloc<__half>("c1", size_t(N) * 112 * 112 * 64);
  n->alloc<__half>("p1", size_t(N) * 56 * 56 * 64);
  int hw = 56;
  for (const Block& b : n->blocks) {
    const int o = hw / b.stride;
    const size_t px = size_t(N) * o * o;
    n->alloc<__half>(b.pre + "t1", px * b.mid);
    n->alloc<__half>(b.pre + "t2", px * b.mid);
    if (b.proj) n->alloc<__half>(b.pre + "sc", px * b.cout);
    if (n->kind == XEMO_NET_SENET50 && n->se_lin(o)) {
      n->alloc<float>(b.pre + "m2", size_t(N) * b.mid);
      n->alloc<float>(b.pre + "gs", size_t(N) * b.cout);
      n->alloc<float>(b.pre + "gh", size_t(N) * b.cout);
    } else if (n->kind == XEMO_NET_SENET50) {
      n->alloc<__half>(b.pre + "u", px * b.cout);
      n->alloc<float>(b.pre + "s", size_t(N) * b.cout);
      n->alloc<float>(b.pre + "g", size_t(N) * b.cout);
    }
    n->alloc<__half>(b.pre + "y", px * b.cout);
    hw = o;
  }
  n->alloc<__half>("pool5", size_t(N) * 2048);
  n->alloc<float>("logits", size_t(N) * Kp);
  for (const auto& kv : n->buf)
    if (!kv.second) return fail(ctx, XEMO_ERR_NOMEM, "net: device allocation of %s failed", kv.first.c_str());
  return XEMO_OK;
}

int teacher_record(xemo_net* n) {
  xemo_ctx* ctx = n->ctx;
  const int N = n->N;
  auto H = [&](const std::string& k) { return n->get<__half>(k); };
  auto F = [&](const std::string& k) { return n->get<float>(k); };
  if (n->input_mode == XEMO_INPUT_U8)
    NET_OP(xemo_op_face_u8_rows_im2col(ctx, n->get<uint8_t>("faces"), n->face_size, n->face_size, N, 224, 224, F("mean3"), 7, 2, 3, 112, H("rows")));
  else
    NET_OP(xemo_op_face_rows_im2col(ctx, F("faces"), 224, 224, 3, N, 7, 2, 3, 112, H("rows")));
  const int pad_stem[4] = {3, 3, 0, 0}, stride1[2] = {1, 1};
  (void)stride1;
  // the stem in pixel-pair form: [N][224][56][64] view of the row-im2col tensor, output = the [N][112][56][128] view of c1
  NET_OP(conv(n, H("rows"), N, 224, 56, 64, H("conv1:w"), 128, 7, 1, 2, 1, pad_stem, F("conv1:a"), F("conv1:b"), nullptr, 1, H("c1")));
  NET_OP(xemo_op_maxpool_fwd(ctx, H("c1"), N, 112, 112, 64, 3, 3, 2, 2, 0, 1, 0, 1, nullptr, nullptr, H("p1"), nullptr));
  const __half* cur = H("p1");
  int hw = 56;
  const bool se = n->kind == XEMO_NET_SENET50;
  for (const Block& b : n->blocks) {
    const int o = hw / b.stride;
    const std::string& p = b.pre;
    NET_OP(conv(n, cur, N, hw, hw, b.cin, H(p + "c1:w"), b.mid, 1, 1, b.stride, b.stride, kPad0, F(p + "c1:a"), F(p + "c1:b"), nullptr, 1, H(p + "t1")));
    NET_OP(conv(n, H(p + "t1"), N, o, o, b.mid, H(p + "c2:w"), b.mid, 3, 3, 1, 1, kPad1, F(p + "c2:a"), F(p + "c2:b"), nullptr, 1, H(p + "t2")));
    const __half* sc = cur;
    if (b.proj) {
      NET_OP(conv(n, cur, N, hw, hw, b.cin, H(p + "proj:w"), b.cout, 1, 1, b.stride, b.stride, kPad0, F(p + "proj:a"), F(p + "proj:b"), nullptr, 0, H(p + "sc")));
      sc = H(p + "sc");
    }
    if (se && n->se_lin(o)) {
      NET_OP(xemo_op_se_squeeze(ctx, H(p + "t2"), N, o * o, b.mid, F(p + "m2")));
      NET_OP(xemo_op_se_gate_lin(ctx, F(p + "m2"), N, b.cout, b.mid, b.cout / 16, H(p + "c3:w"), F(p + "c3:a"), F(p + "c3:b"), F(p + "se1:w"), F(p + "se1:b"),
                                 F(p + "se2:w"), F(p + "se2:b"), F(p + "gs"), F(p + "gh")));
      NET_OP(xemo_op_conv_fwd_nc(ctx, H(p + "t2"), N, o, o, b.mid, H(p + "c3:w"), b.cout, 1, 1, 1, 1, 0, 0, 0, 0, F(p + "gs"), F(p + "gh"), sc, 1, H(p + "y")));
    } else if (se) {
      NET_OP(conv(n, H(p + "t2"), N, o, o, b.mid, H(p + "c3:w"), b.cout, 1, 1, 1, 1, kPad0, F(p + "c3:a"), F(p + "c3:b"), nullptr, 0, H(p + "u")));
      NET_OP(xemo_op_se_squeeze(ctx, H(p + "u"), N, o * o, b.cout, F(p + "s")));
      NET_OP(xemo_op_se_gate(ctx, F(p + "s"), N, b.cout, b.cout / 16, F(p + "se1:w"), F(p + "se1:b"), F(p + "se2:w"), F(p + "se2:b"), F(p + "g")));
      NET_OP(xemo_op_se_excite(ctx, H(p + "u"), F(p + "g"), sc, N, o * o, b.cout, 1, H(p + "y")));
    } else {
      NET_OP(conv(n, H(p + "t2"), N, o, o, b.mid, H(p + "c3:w"), b.cout, 1, 1, 1, 1, kPad0, F(p + "c3:a"), F(p + "c3:b"), sc, 1, H(p + "y")));
    }
    cur = H(p + "y");
    hw = o;
  }
  NET_OP(xemo_op_avgpool_fwd(ctx, cur, N, 7, 7, 2048, 7, 7, 1, 1, 0, 0, 0, 0, H("pool5")));
  const int Kp = pad16(n->K);
  NET_OP(conv(n, H("pool5"), N, 1, 1, 2048, H("classifier:w"), Kp, 1, 1, 1, 1, kPad0, nullptr, F("classifier:b"), nullptr, 0, nullptr, F("logits"), Kp));
  return XEMO_OK;
}

// ================================================================================================ student
struct StudentConvDef { const char* name; int fh, fw, cin, cout, sh, sw, pad; bool bn; };
const StudentConvDef kStudentConvs[8] = {
    {"conv1", 7, 7, 1, 96, 2, 2, 1, true},   {"conv2", 5, 5, 96, 256, 2, 2, 1, true}, {"conv3", 3, 3, 256, 384, 1, 1, 1, true},
    {"conv4", 3, 3, 384, 256, 1, 1, 1, true}, {"conv5", 3, 3, 256, 256, 1, 1, 1, true}, {"fc6", 9, 1, 256, 4096, 1, 1, 0, true},
    {"fc7", 1, 1, 4096, 1024, 1, 1, 0, true}, {"fc8", 1, 1, 1024, 8, 1, 1, 0, false}};

int student_describe(xemo_net* n) {
  int h = 512, w = n->W;
  for (const StudentConvDef& d : kStudentConvs) {
    ConvLayer L;
    L.name = d.name;
    L.bn = d.bn ? std::string("bn") + L.name.back() : "";
    L.fh = d.fh; L.fw = d.fw; L.cin = d.cin; L.cout = L.name == "fc8" ? n->K : d.cout;
    L.kp = pad16(L.cout); L.cp = pad16(L.cin);
    L.sh = d.sh; L.sw = d.sw;
    for (int i = 0; i < 4; ++i) L.pad[i] = d.pad;
    L.has_bn = d.bn;
    L.h = h; L.w = w;
    L.oh = out_dim(h, d.pad, d.pad, d.fh, d.sh); L.ow = out_dim(w, d.pad, d.pad, d.fw, d.sw);
    if (L.oh <= 0 || L.ow <= 0) return fail(n->ctx, XEMO_ERR_INVALID, "net: a %d-column spectrogram is too narrow for %s", n->W, d.name);
    {
      const char* e = getenv("XEMO_DGRAD_FULLHEIGHT");
      L.full_height = d.fh == h && d.fw == 1 && d.fh > 1 && d.sh == 1 && d.sw == 1 && d.pad == 0 && L.cp <= 256 && !(e && e[0] == '0');
    }
    h = L.oh; w = L.ow;
    if (L.name == "conv1" || L.name == "conv2") { L.pool_method = 0; L.pwh = 3; L.pww = 3; L.psh = 2; L.psw = 2; }
    if (L.name == "conv5") { L.pool_method = 0; L.pwh = 5; L.pww = 3; L.psh = 3; L.psw = 2; }
    if (L.name == "fc6") { L.pool_method = 1; L.pwh = 1; L.pww = w; L.psh = 1; L.psw = 1; }   // pool6 averages the whole remaining width
    if (L.pool_method >= 0) {
      L.ph = h; L.pw = w;
      L.poh = out_dim(h, 0, 0, L.pwh, L.psh); L.pow_ = out_dim(w, 0, 0, L.pww, L.psw);
      if (L.poh <= 0 || L.pow_ <= 0) return fail(n->ctx, XEMO_ERR_INVALID, "net: a %d-column spectrogram is too narrow for the pooling after %s", n->W, d.name);
      h = L.poh; w = L.pow_;
    }
    n->layers.push_back(L);
    add_param(n, L.name + "f", L.fh, L.fw, L.cin, L.cout);
    add_param(n, L.name + "b", L.cout, 1, 1, 1);
    if (L.has_bn) {
      add_param(n, L.bn + "m", L.cout, 1, 1, 1); add_param(n, L.bn + "b", L.cout, 1, 1, 1); add_param(n, L.bn + "x", L.cout, 2, 1, 1);
    }
  }
  if (h != 1 || w != 1) return fail(n->ctx, XEMO_ERR_INVALID, "net: the student graph must reduce to 1 x 1 (got %d x %d for width %d)", h, w, n->W);
  ConvLayer& L1 = n->layers[0];
  n->s2d_hp = L1.oh + 3; n->s2d_ow = L1.ow;
  n->stem_pairs = L1.ow % 2 == 0;
  // conv2 reads pool1's output with a 128-channel pitch (96 real + 32 zero channels): 128-byte TMA rows
  n->layers[1].cp = (n->layers[1].cin + 63) / 64 * 64;
  n->pool1_ld = n->layers[1].cp;
  return XEMO_OK;
}

size_t seg_len(const ConvLayer& L, char what) {   // f: device-layout filter, b: padded bias, m / B: BN mult / bias
  if (what == 'f') return L.name == "conv1" ? size_t(L.kp) * 64 : size_t(L.kp) * L.fh * L.fw * L.cp;
  if (what == 'b') return size_t(L.kp);
  return size_t(L.cout);
}

int student_finalize(xemo_net* n) {
  xemo_ctx* ctx = n->ctx;
  const int N = n->N;
  // one flat fp32 master / momentum / gradient buffer (the single all-reduce payload) + an fp16 mirror at equal offsets
  size_t off = 0;
  auto seg = [&](const std::string& name, size_t len) { n->seg[name] = off; off += (len + 63) / 64 * 64; };
  for (const ConvLayer& L : n->layers) {
    seg(L.name + "f", seg_len(L, 'f'));
    seg(L.name + "b", seg_len(L, 'b'));
    if (L.has_bn) { seg(L.bn + "m", L.cout); seg(L.bn + "b", L.cout); }
  }
  n->nparam = off;
  n->split_layer = 5;   // fc6
  n->split_offset = n->seg["fc6f"];
  std::vector<float> flat(off, 0.f);
  for (const ConvLayer& L : n->layers) {
    const std::vector<float>& f = n->host[L.name + "f"];
    const std::vector<float> dev = L.name == "conv1" ? conv1_to_s2d(f, L.cout) : krsc_host(f, L.fh, L.fw, L.cin, L.cout, L.kp, L.cp);
    memcpy(&flat[n->seg[L.name + "f"]], dev.data(), dev.size() * 4);
    memcpy(&flat[n->seg[L.name + "b"]], n->host[L.name + "b"].data(), size_t(L.cout) * 4);
    if (L.has_bn) {
      memcpy(&flat[n->seg[L.bn + "m"]], n->host[L.bn + "m"].data(), size_t(L.cout) * 4);
      memcpy(&flat[n->seg[L.bn + "b"]], n->host[L.bn + "b"].data(), size_t(L.cout) * 4);
    }
  }
  // the batch moments of all layers in one buffer: under data parallelism they are summed across the ranks with one small
  // all-reduce (the parameter server of cnn_train_dag sums the labs' moments like any other derivative)
  n->bm_elems = 0;
  for (const ConvLayer& L : n->layers) if (L.has_bn) n->bm_elems += size_t(2) * L.cout;
  n->bm_flat = n->alloc<float>("batch_moments", n->bm_elems);
  {
    // ... and the moments parameters ([mu | sigma] per layer) at the same offsets of a second buffer: one moving-average launch
    std::vector<float> mom(n->bm_elems);
    size_t o = 0;
    for (const ConvLayer& L : n->layers)
      if (L.has_bn) { memcpy(&mom[o], n->host[L.bn + "x"].data(), size_t(2) * L.cout * 4); o += size_t(2) * L.cout; }
    n->mom_flat = upload_f32(n, "moments", mom);
    o = 0;
    for (const ConvLayer& L : n->layers)
      if (L.has_bn) {
        n->buf[L.bn + ":batch_moments"] = n->bm_flat + o;
        n->buf[L.bn + ":moments"] = n->mom_flat + o;
        o += size_t(2) * L.cout;
      }
  }
  n->master = upload_f32(n, "master", flat);
  n->momentum = n->alloc<float>("momentum", off);
  n->grad = n->alloc<float>("grad", off);
  n->w16 = upload_f16(n, "w16", flat);
  const std::vector<float> hy = {1e-4f, 0.9f, 5e-4f, 1.f / N};
  n->hyper = upload_f32(n, "hyper", hy);
  n->guard = n->alloc<int>("guard", 3);
  // activations
  n->alloc<float>("spec", size_t(N) * 512 * n->W);
  n->alloc<__half>("s2d", size_t(N) * n->s2d_hp * n->s2d_ow * 16);
  const int c1 = n->layers[0].kp;
  if (n->stem_pairs) {
    n->alloc<__half>("stem:w2", size_t(2) * c1 * 4 * 32);
    n->alloc<float>("stem:shift2", 2 * c1);
    n->alloc<float>("stem:scale2", 2 * c1);
  }
  n->alloc<double>("stem:ws", xemo_stem_ws_doubles());
  n->alloc<float>("target", size_t(N) * n->K);
  upload_f32(n, "weights", std::vector<float>(N, 1.f));
  for (const ConvLayer& L : n->layers) {
    const std::string& s = L.name;
    const size_t px = size_t(N) * L.oh * L.ow;
    n->alloc<__half>(s + ":raw", px * L.kp);
    n->alloc<__half>(s + ":draw", px * L.kp);
    if (L.has_bn) {
      n->alloc<float>(s + ":a", L.cout);
      n->alloc<float>(s + ":b", L.cout);
      n->alloc<double>(s + ":ws", size_t(2) * L.cout);
    }
    if (L.pool_method >= 0) {
      const int pc = s == "conv1" ? n->pool1_ld : L.cout;   // (padding channels stay zero: never written)
      const size_t pp = size_t(N) * L.poh * L.pow_ * pc;
      n->alloc<__half>(s + ":out", pp);
      n->alloc<__half>(s + ":dout", pp);
      if (L.pool_method == 0) {
        n->alloc<uint8_t>(s + ":arg", pp);
        if (s == "conv1") n->alloc<__half>(s + ":xwin", pp);
      } else {
        n->alloc<__half>(s + ":act", px * L.cout);
        n->alloc<__half>(s + ":dact", px * L.cout);
      }
    } else if (L.has_bn) {
      n->alloc<__half>(s + ":out", px * L.cout);
      n->alloc<__half>(s + ":dout", px * L.cout);
    }
    if (s != "conv1") n->alloc<__half>(s + ":packed", xemo_dgrad_pack_elems(L.cp, L.kp, L.fh, L.fw, L.sh, L.sw));
  }
  XEMO_CUDA(ctx, cudaStreamCreateWithFlags(&n->side_a, cudaStreamNonBlocking));
  XEMO_CUDA(ctx, cudaStreamCreateWithFlags(&n->side_b, cudaStreamNonBlocking));
  n->alloc<float>("pred32", size_t(N) * n->layers.back().kp);
  n->alloc<float>("scalars", 2);
  n->alloc<float>("class_stats", size_t(2) * n->K);
  n->alloc<int>("max_label", N);
  for (const auto& kv : n->buf)
    if (!kv.second) return fail(ctx, XEMO_ERR_NOMEM, "net: device allocation of %s failed", kv.first.c_str());
  return XEMO_OK;
}

// conv1 as a 4 x 1 convolution over the space-to-depth tensor, in pixel-pair form when the output width is even
int student_stem_conv(xemo_net* n, const __half* wt, const float* scale, const float* shift, int relu, __half* dst) {
  xemo_ctx* ctx = n->ctx;
  const ConvLayer& L = n->layers[0];
  const __half* x = n->get<__half>("s2d");
  if (!n->stem_pairs) return conv(n, x, n->N, n->s2d_hp, n->s2d_ow, 16, wt, L.kp, 4, 1, 1, 1, kPad0, scale, shift, nullptr, relu, dst);
  NET_OP(xemo_op_stem_pair_filter(ctx, wt, L.kp, n->get<__half>("stem:w2")));
  NET_OP(xemo_op_tile_f32(ctx, shift, L.kp, 2, 0.f, n->get<float>("stem:shift2")));
  if (scale) NET_OP(xemo_op_tile_f32(ctx, scale, L.kp, 2, 1.f, n->get<float>("stem:scale2")));
  return conv(n, x, n->N, n->s2d_hp, n->s2d_ow / 2, 32, n->get<__half>("stem:w2"), 2 * L.kp, 4, 1, 1, 1, kPad0,
              scale ? n->get<float>("stem:scale2") : nullptr, n->get<float>("stem:shift2"), nullptr, relu, dst);
}

int student_record_packs(xemo_net* n);

int student_record_forward_train(xemo_net* n) {
  xemo_ctx* ctx = n->ctx;
  const int N = n->N;
  auto H = [&](const std::string& k) { return n->get<__half>(k); };
  auto F = [&](const std::string& k) { return n->get<float>(k); };
  // with overlap: the data-gradient filter packs (weights only) run on a forked stream beside the forward pass
  n->packs_ahead = n->use_overlap() && n->side_b && ctx->stream == ctx->primary;
  if (n->packs_ahead) {
    NET_OP(xemo_stream_wait(ctx, n->side_b, nullptr));
    NET_OP(xemo_set_stream(ctx, n->side_b));
    const int rc = student_record_packs(n);
    NET_OP(xemo_set_stream(ctx, nullptr));
    if (rc) return rc;
  }
  NET_OP(xemo_op_spec_s2d(ctx, F("spec"), 512, n->W, N, 1, 1, n->s2d_hp, n->s2d_ow, H("s2d")));
  NET_OP(xemo_op_stem_autocorr(ctx, H("s2d"), N, n->s2d_hp, n->s2d_ow, n->layers[0].oh, n->get<double>("stem:ws")));
  const __half* cur = H("s2d");
  for (const ConvLayer& L : n->layers) {
    const std::string& s = L.name;
    const __half* wt = n->w16 + n->seg[s + "f"];
    const float* bias = n->master + n->seg[s + "b"];
    const bool last = s == "fc8", stem = s == "conv1";
    if (stem) NET_OP(student_stem_conv(n, wt, nullptr, bias, 0, H(s + ":raw")));
    else NET_OP(conv(n, cur, N, L.h, L.w, L.cp, wt, L.kp, L.fh, L.fw, L.sh, L.sw, L.pad, nullptr, bias, nullptr, 0, H(s + ":raw"),
                     last ? F("pred32") : nullptr, L.kp));
    cur = H(s + ":raw");
    if (!L.has_bn) continue;
    const float *g = n->master + n->seg[L.bn + "m"], *beta = n->master + n->seg[L.bn + "b"];
    const size_t rows = size_t(N) * L.oh * L.ow;
    if (stem)   // batch statistics of w.patch + b from the patch autocorrelation: no pass over the activation
      NET_OP(xemo_op_stem_bn_train(ctx, n->get<double>("stem:ws"), wt, bias, rows, L.cout, g, beta, kBnEps, F(L.bn + ":batch_moments"), F(s + ":a"), F(s + ":b")));
    else
      NET_OP(xemo_op_bn_train(ctx, cur, rows, L.cout, g, beta, kBnEps, n->get<double>(s + ":ws"), F(L.bn + ":batch_moments"), F(s + ":a"), F(s + ":b")));
    if (L.pool_method == 0 && stem)
      NET_OP(xemo_op_maxpool_fwd_win(ctx, cur, N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, F(s + ":a"), F(s + ":b"), H(s + ":out"),
                                     n->get<uint8_t>(s + ":arg"), H(s + ":xwin"), n->pool1_ld));
    else if (L.pool_method == 0)
      NET_OP(xemo_op_maxpool_fwd(ctx, cur, N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, F(s + ":a"), F(s + ":b"), H(s + ":out"),
                                 n->get<uint8_t>(s + ":arg")));
    else if (L.pool_method == 1) {
      NET_OP(xemo_op_affine_act(ctx, cur, rows, L.cout, F(s + ":a"), F(s + ":b"), 1, H(s + ":act")));
      NET_OP(xemo_op_avgpool_fwd(ctx, H(s + ":act"), N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, H(s + ":out")));
    } else
      NET_OP(xemo_op_affine_act(ctx, cur, rows, L.cout, F(s + ":a"), F(s + ":b"), 1, H(s + ":out")));
    cur = H(s + ":out");
  }
  if (n->packs_ahead) NET_OP(xemo_stream_wait(ctx, nullptr, n->side_b));   // join before the backward pass
  return XEMO_OK;
}

// dag.mode = 'test' (external/compute_audio_feats.m:106): BN uses the stored moments, so it folds -- with the conv bias --
// into the convolution's scale / shift epilogue
int student_record_forward_test(xemo_net* n) {
  xemo_ctx* ctx = n->ctx;
  const int N = n->N;
  auto H = [&](const std::string& k) { return n->get<__half>(k); };
  auto F = [&](const std::string& k) { return n->get<float>(k); };
  NET_OP(xemo_op_spec_s2d(ctx, F("spec"), 512, n->W, N, 1, 1, n->s2d_hp, n->s2d_ow, H("s2d")));
  const __half* cur = H("s2d");
  for (const ConvLayer& L : n->layers) {
    const std::string& s = L.name;
    const __half* wt = n->w16 + n->seg[s + "f"];
    const float* bias = n->master + n->seg[s + "b"];
    const float *scale = nullptr, *shift = bias;
    float* out32 = nullptr;
    int relu = 0;
    if (L.has_bn) {
      NET_OP(xemo_op_bn_test(ctx, F(L.bn + ":moments"), L.cout, n->master + n->seg[L.bn + "m"], n->master + n->seg[L.bn + "b"], bias, F(s + ":a"), F(s + ":b")));
      scale = F(s + ":a"); shift = F(s + ":b"); relu = 1;
    } else
      out32 = F("pred32");
    __half* dst = (L.pool_method >= 0 || !L.has_bn) ? H(s + ":raw") : H(s + ":out");
    if (s == "conv1") NET_OP(student_stem_conv(n, wt, scale, shift, relu, dst));
    else NET_OP(conv(n, cur, N, L.h, L.w, L.cp, wt, L.kp, L.fh, L.fw, L.sh, L.sw, L.pad, scale, shift, nullptr, relu, dst, out32, L.kp));
    if (L.pool_method == 0 && s == "conv1" && n->pool1_ld != L.cout) {
      NET_OP(xemo_op_maxpool_fwd_win(ctx, dst, N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, nullptr, nullptr, H(s + ":out"), nullptr, nullptr, n->pool1_ld));
      dst = H(s + ":out");
    } else if (L.pool_method == 0) {
      NET_OP(xemo_op_maxpool_fwd(ctx, dst, N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, nullptr, nullptr, H(s + ":out"), nullptr));
      dst = H(s + ":out");
    } else if (L.pool_method == 1) {
      NET_OP(xemo_op_avgpool_fwd(ctx, dst, N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, H(s + ":out")));
      dst = H(s + ":out");
    }
    cur = dst;
  }
  return XEMO_OK;
}

// the parity-decomposed (flipped / transposed) filter copies of the data-gradient convolutions: they depend on the weights only
int student_record_packs(xemo_net* n) {
  for (size_t i = 1; i < n->layers.size(); ++i) {
    const ConvLayer& L = n->layers[i];
    if (L.full_height)
      NET_OP(xemo_op_pack_dgrad_filters_fullheight(n->ctx, n->w16 + n->seg[L.name + "f"], L.kp, L.fh, L.cp, n->get<__half>(L.name + ":packed")));
    else
      NET_OP(xemo_op_pack_dgrad_filters(n->ctx, n->w16 + n->seg[L.name + "f"], L.kp, L.fh, L.fw, L.cp, L.sh, L.sw, L.pad[0], L.pad[2],
                                        n->get<__half>(L.name + ":packed")));
  }
  return XEMO_OK;
}

// loss (when `loss`) and the backward sweep over layers [lo, hi) in reverse order.  With overlap the filter gradients of a
// layer run on a forked stream (nothing downstream needs them before the exchange / update) while the primary stream
// continues with the data gradient; the fork is joined before returning.
int student_record_backward(xemo_net* n, int lo, int hi, bool loss) {
  xemo_ctx* ctx = n->ctx;
  const bool fork = n->use_overlap() && n->side_b;
  const int N = n->N;
  const float gs = n->grad_scale, inv = 1.f / gs;
  auto H = [&](const std::string& k) { return n->get<__half>(k); };
  auto F = [&](const std::string& k) { return n->get<float>(k); };
  const ConvLayer& last = n->layers.back();
  if (loss) {
    NET_OP(xemo_memset(ctx, n->grad, 0, n->nparam * 4));
    NET_OP(xemo_memset(ctx, H("fc8:draw"), 0, size_t(N) * last.kp * 2));
    NET_OP(xemo_memset(ctx, F("scalars"), 0, 8));   // objective / classerror of THIS batch (class_stats keep accumulating)
    const bool soft = n->loss_type == XEMO_LOSS_SOFTMAXCE;
    const int kernel_type = n->loss_type == XEMO_LOSS_EUCLIDEAN ? 1 : n->loss_type == XEMO_LOSS_HUBER ? 2 : 0;
    NET_OP(xemo_op_loss(ctx, F("pred32"), 1, last.kp, F("target"), n->K, kernel_type ? F("weights") : nullptr, N, n->K, kernel_type,
                        soft ? n->temperature : 1.f, soft ? 1 : 0, 1.f, gs, H("fc8:draw"), 0, last.kp, F("scalars"), F("class_stats"),
                        n->get<int>("max_label")));
  }
  for (int i = hi - 1; i >= lo; --i) {
    const ConvLayer& L = n->layers[i];
    const std::string& s = L.name;
    const size_t rows = size_t(N) * L.oh * L.ow;
    bool fused_bias = false;
    const bool stem = s == "conv1";
    if (stem) {
      // ReLU mask + the two BN reductions at the pooled resolution, then the (masked) gradient w.r.t. the never-materialised
      // ReLU output at the conv resolution: dz, which the filter gradient consumes directly
      const size_t prow = size_t(N) * L.poh * L.pow_;
      NET_OP(xemo_op_stem_pool_bn_reduce(ctx, H(s + ":xwin"), H(s + ":dout"), prow, L.cout, n->pool1_ld, F(L.bn + ":batch_moments"), F(s + ":a"), F(s + ":b"),
                                         n->get<double>(s + ":ws")));
      NET_OP(xemo_op_maxpool_bwd_ld(ctx, H(s + ":dout"), n->get<uint8_t>(s + ":arg"), N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, H(s + ":draw"),
                                    n->pool1_ld));
      fused_bias = true;
    } else if (L.has_bn) {
      const __half* dcur = H(s + ":dout");
      fused_bias = L.kp == L.cout;
      if (L.pool_method == 0) {
        NET_OP(xemo_op_maxpool_bwd(ctx, dcur, n->get<uint8_t>(s + ":arg"), N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, H(s + ":draw")));
        dcur = H(s + ":draw");
      } else if (L.pool_method == 1) {
        NET_OP(xemo_op_avgpool_bwd(ctx, dcur, N, L.oh, L.ow, L.cout, L.pwh, L.pww, L.psh, L.psw, 0, 0, 0, 0, H(s + ":dact")));
        dcur = H(s + ":dact");
      }
      NET_OP(xemo_op_bn_bwd(ctx, H(s + ":raw"), dcur, rows, L.cout, F(L.bn + ":batch_moments"), F(s + ":a"), F(s + ":b"), 1, 0, n->get<double>(s + ":ws"), H(s + ":draw"),
                            n->grad + n->seg[L.bn + "m"], n->grad + n->seg[L.bn + "b"], fused_bias ? n->grad + n->seg[s + "b"] : nullptr, inv));
    }
    const __half* dy = H(s + ":draw");
    const __half* x = i == 0 ? H("s2d") : H(n->layers[i - 1].name + ":out");
    float* gf = n->grad + n->seg[s + "f"];
    if (fork) {
      NET_OP(xemo_stream_wait(ctx, n->side_b, nullptr));
      NET_OP(xemo_set_stream(ctx, n->side_b));
    }
    if (stem) {
      NET_OP(xemo_op_conv_wgrad(ctx, x, N, n->s2d_hp, n->s2d_ow, 16, dy, L.kp, L.kp, 4, 1, 1, 1, 0, 0, 0, 0, gf, inv));
      NET_OP(xemo_op_stem_wgrad_finalize(ctx, n->get<double>("stem:ws"), n->w16 + n->seg[s + "f"], n->master + n->seg[s + "b"], n->get<double>(s + ":ws"), rows,
                                         L.cout, F(L.bn + ":batch_moments"), F(s + ":a"), inv, gf, n->grad + n->seg[s + "b"], n->grad + n->seg["bn1m"],
                                         n->grad + n->seg["bn1b"], nullptr));
    } else {
      NET_OP(xemo_op_conv_wgrad(ctx, x, N, L.h, L.w, L.cp, dy, L.kp, L.kp, L.fh, L.fw, L.sh, L.sw, L.pad[0], L.pad[1], L.pad[2], L.pad[3], gf, inv));
    }
    if (!fused_bias) NET_OP(xemo_op_colsum(ctx, dy, rows, L.kp, L.kp, inv, n->grad + n->seg[s + "b"]));
    if (fork) NET_OP(xemo_set_stream(ctx, nullptr));
    if (i > 0) {
      if (!n->packs_ahead) {
        if (L.full_height) NET_OP(xemo_op_pack_dgrad_filters_fullheight(ctx, n->w16 + n->seg[s + "f"], L.kp, L.fh, L.cp, H(s + ":packed")));
        else NET_OP(xemo_op_pack_dgrad_filters(ctx, n->w16 + n->seg[s + "f"], L.kp, L.fh, L.fw, L.cp, L.sh, L.sw, L.pad[0], L.pad[2], H(s + ":packed")));
      }
      // fc6 (a 9 x 1 filter over a 9 x W map, one output row): the data gradient as a plain GEMM over (n, w) rows
      if (L.full_height) NET_OP(xemo_op_conv_dgrad_fullheight(ctx, dy, N, L.h, L.w, L.cp, H(s + ":packed"), L.kp, H(n->layers[i - 1].name + ":dout")));
      else NET_OP(xemo_op_conv_dgrad(ctx, dy, N, L.h, L.w, L.cp, H(s + ":packed"), L.kp, L.fh, L.fw, L.sh, L.sw, L.pad[0], L.pad[1], L.pad[2], L.pad[3],
                                     H(n->layers[i - 1].name + ":dout")));
    }
  }
  if (fork) NET_OP(xemo_stream_wait(ctx, nullptr, n->side_b));   // join
  return XEMO_OK;
}

// cnn_train_dag's accumulateGradients: one launch over the flat master buffer (also refreshes the fp16 mirror), guarded
// against a non-finite gradient; BN moments moving average
int student_record_update(xemo_net* n) {
  xemo_ctx* ctx = n->ctx;
  NET_OP(xemo_op_grad_guard(ctx, n->grad, n->nparam, n->guard));
  NET_OP(xemo_op_sgd_momentum_guarded(ctx, n->master, n->momentum, n->grad, n->nparam, n->hyper, 1.f, 1.f, 1.f, n->w16, n->guard));
  return xemo_op_moments_average_guarded(ctx, n->mom_flat, n->bm_flat, int(n->bm_elems), 0.1f, n->bm_scale, n->guard);
}

// run `record` eagerly once (kernel attributes must be set outside a capture), then capture it
template <typename Fn>
int capture(xemo_net* n, xemo_graph** g, Fn record, bool warm = true) {
  if (warm) NET_OP(record());
  NET_OP(xemo_capture_begin(n->ctx));
  const int rc = record();
  xemo_graph* out = nullptr;
  const int rc2 = xemo_capture_end(n->ctx, &out);
  if (rc) { if (out) xemo_graph_destroy(out); return rc; }
  if (rc2) return rc2;
  *g = out;
  return XEMO_OK;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// [N][ld] device rows -> caller's K x N column-major (== [N][K]) array, host or device
int copy_out_rows(xemo_ctx* ctx, const float* src, int N, int ld, int K, float* dst) {
  XEMO_CUDA(ctx, cudaMemcpy2DAsync(dst, size_t(K) * 4, src, size_t(ld) * 4, size_t(K) * 4, N, cudaMemcpyDefault, ctx->stream));
  if (!is_device_ptr(dst)) XEMO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return XEMO_OK;
}

}  // namespace

// ================================================================================================ C entry points
extern "C" int xemo_net_create(xemo_ctx* ctx, int kind, int batch, int size, int input_mode, int num_outputs, xemo_net** out) {
  // (ctx may be NULL: a description-only network -- parameter names / dims can be enumerated without a device, nothing else)
  XEMO_REQUIRE(ctx, out, "net_create: null pointer");
  *out = nullptr;
  XEMO_REQUIRE(ctx, kind == XEMO_NET_RESNET50 || kind == XEMO_NET_SENET50 || kind == XEMO_NET_VGGVOX, "net_create: unknown network kind %d", kind);
  XEMO_REQUIRE(ctx, batch >= 1 && num_outputs >= 1 && num_outputs <= 16, "net_create: batch >= 1 and 1 <= num_outputs <= 16 required");
  xemo_net* n = new xemo_net();
  n->ctx = ctx; n->kind = kind; n->N = batch; n->K = num_outputs; n->input_mode = input_mode;
  int rc = XEMO_OK;
  if (kind == XEMO_NET_VGGVOX) {
    n->W = size;
    if (input_mode != XEMO_INPUT_F32) rc = fail(ctx, XEMO_ERR_INVALID, "net_create: the student takes single-precision spectrograms");
    else rc = student_describe(n);
  } else {
    n->face_size = input_mode == XEMO_INPUT_U8 ? size : 224;
    if (input_mode != XEMO_INPUT_U8 && input_mode != XEMO_INPUT_F32) rc = fail(ctx, XEMO_ERR_INVALID, "net_create: input_mode must be XEMO_INPUT_F32 or XEMO_INPUT_U8");
    else if (input_mode == XEMO_INPUT_U8 && size < 2) rc = fail(ctx, XEMO_ERR_INVALID, "net_create: face size must be >= 2");
    else teacher_describe(n);
  }
  if (rc) { delete n; return rc; }
  *out = n;
  return XEMO_OK;
}

extern "C" void xemo_net_destroy(xemo_net* n) {
  if (!n) return;
  if (n->ctx) cudaStreamSynchronize(n->ctx->primary);
  for (xemo_graph* g : {n->g_fwd, n->g_train, n->g_update, n->g_step}) xemo_graph_destroy(g);
  for (void* p : n->owned) cudaFree(p);
  if (n->side_a) cudaStreamDestroy(n->side_a);
  if (n->side_b) cudaStreamDestroy(n->side_b);
  delete n;
}

extern "C" int xemo_net_batch(xemo_net* n) { return n ? n->N : 0; }
extern "C" int xemo_net_num_params(xemo_net* n) { return n ? int(n->order.size()) : 0; }
extern "C" const char* xemo_net_param_name(xemo_net* n, int i) { return (n && i >= 0 && i < int(n->order.size())) ? n->order[i].c_str() : nullptr; }
extern "C" int xemo_net_param_dims(xemo_net* n, const char* name, int64_t dims[4]) {
  if (!n || !name || !dims) return XEMO_ERR_INVALID;
  auto it = n->spec.find(name);
  XEMO_REQUIRE(n->ctx, it != n->spec.end(), "net: no parameter named %s", name);
  for (int i = 0; i < 4; ++i) dims[i] = it->second.d[i];
  return XEMO_OK;
}

static int student_write_param(xemo_net* n, const std::string& name, const float* data);

extern "C" int xemo_net_set_param(xemo_net* n, const char* name, const float* data, size_t numel) {
  if (!n || !name || !data) return XEMO_ERR_INVALID;
  auto it = n->spec.find(name);
  XEMO_REQUIRE(n->ctx, it != n->spec.end(), "net: no parameter named %s", name);
  XEMO_REQUIRE(n->ctx, numel == numel_of(it->second), "net: %s has %zu elements, got %zu", name, numel_of(it->second), numel);
  XEMO_REQUIRE(n->ctx, !is_device_ptr(data), "net: parameters are handed over from host memory");
  if (n->finalized) {
    XEMO_REQUIRE(n->ctx, n->kind == XEMO_NET_VGGVOX, "net: the teacher's parameters are folded at finalize; set them before");
    return student_write_param(n, name, data);
  }
  n->host[name].assign(data, data + numel);
  return XEMO_OK;
}

extern "C" int xemo_net_finalize(xemo_net* n) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->ctx, "net: a description-only network (NULL context) cannot be finalized");
  XEMO_REQUIRE(n->ctx, !n->finalized, "net: already finalized");
  for (const std::string& name : n->order)
    XEMO_REQUIRE(n->ctx, n->host.count(name), "net: parameter %s was never set", name.c_str());
  XEMO_CUDA(n->ctx, cudaSetDevice(n->ctx->device));
  NET_OP(n->kind == XEMO_NET_VGGVOX ? student_finalize(n) : teacher_finalize(n));
  XEMO_CUDA(n->ctx, cudaStreamSynchronize(n->ctx->stream));
  n->host.clear();
  n->finalized = true;
  return XEMO_OK;
}

// ---- student parameter access in MatConvNet layouts (master / momentum / gradient / moments)
static int student_locate(xemo_net* n, const std::string& name, const ConvLayer** L, char* what) {
  for (const ConvLayer& l : n->layers) {
    if (name == l.name + "f") { *L = &l; *what = 'f'; return XEMO_OK; }
    if (name == l.name + "b") { *L = &l; *what = 'b'; return XEMO_OK; }
    if (l.has_bn && name == l.bn + "m") { *L = &l; *what = 'm'; return XEMO_OK; }
    if (l.has_bn && name == l.bn + "b") { *L = &l; *what = 'B'; return XEMO_OK; }
    if (l.has_bn && name == l.bn + "x") { *L = &l; *what = 'x'; return XEMO_OK; }
  }
  return fail(n->ctx, XEMO_ERR_INVALID, "net: no parameter named %s", name.c_str());
}

static int student_read(xemo_net* n, const float* flat, const char* which_moments, const std::string& name, float* out) {
  const ConvLayer* L; char what;
  NET_OP(student_locate(n, name, &L, &what));
  XEMO_CUDA(n->ctx, cudaStreamSynchronize(n->ctx->primary));
  if (what == 'x') {
    XEMO_CUDA(n->ctx, cudaMemcpy(out, n->get<float>(L->bn + which_moments), size_t(2) * L->cout * 4, cudaMemcpyDeviceToHost));
    return XEMO_OK;
  }
  const size_t len = what == 'f' ? seg_len(*L, 'f') : size_t(L->cout);
  std::vector<float> tmp(len);
  XEMO_CUDA(n->ctx, cudaMemcpy(tmp.data(), flat + n->seg[name], len * 4, cudaMemcpyDeviceToHost));
  if (what != 'f') memcpy(out, tmp.data(), len * 4);
  else if (L->name == "conv1") conv1_from_s2d(tmp.data(), L->cout, out);
  else unkrsc_host(tmp.data(), L->fh, L->fw, L->cin, L->cout, L->cp, out);
  return XEMO_OK;
}

static int student_write_param(xemo_net* n, const std::string& name, const float* data) {
  const ConvLayer* L; char what;
  NET_OP(student_locate(n, name, &L, &what));
  XEMO_CUDA(n->ctx, cudaStreamSynchronize(n->ctx->primary));
  if (what == 'x') {
    XEMO_CUDA(n->ctx, cudaMemcpy(n->get<float>(L->bn + ":moments"), data, size_t(2) * L->cout * 4, cudaMemcpyHostToDevice));
    return XEMO_OK;
  }
  std::vector<float> dev;
  if (what != 'f') dev.assign(data, data + L->cout);
  else {
    const std::vector<float> f(data, data + size_t(L->fh) * L->fw * L->cin * L->cout);
    dev = L->name == "conv1" ? conv1_to_s2d(f, L->cout) : krsc_host(f, L->fh, L->fw, L->cin, L->cout, L->kp, L->cp);
  }
  XEMO_CUDA(n->ctx, cudaMemcpy(n->master + n->seg[name], dev.data(), dev.size() * 4, cudaMemcpyHostToDevice));
  std::vector<__half> h(dev.size());
  for (size_t i = 0; i < dev.size(); ++i) h[i] = __float2half_rn(dev[i]);
  XEMO_CUDA(n->ctx, cudaMemcpy(n->w16 + n->seg[name], h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  return XEMO_OK;
}

extern "C" int xemo_net_get_tensor(xemo_net* n, int which, const char* name, float* out, size_t numel) {
  if (!n || !name || !out) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "net_get_tensor: a finalized student network is required");
  auto it = n->spec.find(name);
  XEMO_REQUIRE(n->ctx, it != n->spec.end(), "net: no parameter named %s", name);
  XEMO_REQUIRE(n->ctx, numel == numel_of(it->second), "net: %s has %zu elements, got %zu", name, numel_of(it->second), numel);
  switch (which) {
    case XEMO_TENSOR_PARAM: return student_read(n, n->master, ":moments", name, out);
    case XEMO_TENSOR_GRAD: return student_read(n, n->grad, ":batch_moments", name, out);
    case XEMO_TENSOR_MOMENTUM: return student_read(n, n->momentum, ":moments", name, out);
    default: return fail(n->ctx, XEMO_ERR_INVALID, "net_get_tensor: which must be XEMO_TENSOR_PARAM / _GRAD / _MOMENTUM");
  }
}

extern "C" int xemo_net_set_momentum(xemo_net* n, const char* name, const float* data, size_t numel) {
  if (!n || !name || !data) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "net_set_momentum: a finalized student network is required");
  const ConvLayer* L; char what;
  NET_OP(student_locate(n, name, &L, &what));
  XEMO_REQUIRE(n->ctx, what != 'x' && numel == numel_of(n->spec[name]), "net_set_momentum: %s is not an optimised parameter of that size", name);
  std::vector<float> dev;
  if (what != 'f') dev.assign(data, data + L->cout);
  else {
    const std::vector<float> f(data, data + numel);
    dev = L->name == "conv1" ? conv1_to_s2d(f, L->cout) : krsc_host(f, L->fh, L->fw, L->cin, L->cout, L->kp, L->cp);
  }
  XEMO_CUDA(n->ctx, cudaStreamSynchronize(n->ctx->primary));
  XEMO_CUDA(n->ctx, cudaMemcpy(n->momentum + n->seg[name], dev.data(), dev.size() * 4, cudaMemcpyHostToDevice));
  return XEMO_OK;
}

// ---- inputs
extern "C" int xemo_net_input_bytes(xemo_net* n, size_t* bytes) {
  if (!n || !bytes) return XEMO_ERR_INVALID;
  if (n->kind == XEMO_NET_VGGVOX) *bytes = size_t(n->N) * 512 * n->W * 4;
  else *bytes = n->input_mode == XEMO_INPUT_U8 ? size_t(n->N) * n->face_size * n->face_size : size_t(n->N) * 3 * 224 * 224 * 4;
  return XEMO_OK;
}

extern "C" int xemo_net_set_input(xemo_net* n, const void* data, size_t bytes) {
  if (!n || !data) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized, "net: finalize first");
  size_t want = 0;
  xemo_net_input_bytes(n, &want);
  XEMO_REQUIRE(n->ctx, bytes == want, "net_set_input: the input is %zu bytes, got %zu", want, bytes);
  void* dst = n->kind == XEMO_NET_VGGVOX ? n->buf["spec"] : n->buf["faces"];
  XEMO_CUDA(n->ctx, cudaMemcpyAsync(dst, data, bytes, cudaMemcpyDefault, n->ctx->stream));
  return XEMO_OK;
}

extern "C" void* xemo_net_buffer(xemo_net* n, const char* name) {
  if (!n || !name) return nullptr;
  if (!strcmp(name, "grad")) return n->grad;
  if (!strcmp(name, "master")) return n->master;
  auto it = n->buf.find(name);
  return it == n->buf.end() ? nullptr : it->second;
}
extern "C" size_t xemo_net_grad_elems(xemo_net* n) { return n ? n->nparam : 0; }

extern "C" int xemo_net_set_target(xemo_net* n, const float* target, const float* weights) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "net_set_target: a finalized student network is required");
  if (target) XEMO_CUDA(n->ctx, cudaMemcpyAsync(n->buf["target"], target, size_t(n->N) * n->K * 4, cudaMemcpyDefault, n->ctx->stream));
  if (weights) XEMO_CUDA(n->ctx, cudaMemcpyAsync(n->buf["weights"], weights, size_t(n->N) * 4, cudaMemcpyDefault, n->ctx->stream));
  return XEMO_OK;
}

extern "C" int xemo_net_set_loss(xemo_net* n, int loss_type, float temperature, float grad_scale) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->kind == XEMO_NET_VGGVOX, "net_set_loss: student networks only");
  XEMO_REQUIRE(n->ctx, loss_type >= XEMO_LOSS_SOFTMAXCE && loss_type <= XEMO_LOSS_HUBER, "unrecognised regression loss: %d", loss_type);
  XEMO_REQUIRE(n->ctx, temperature > 0.f && grad_scale > 0.f, "net_set_loss: temperature and grad_scale must be positive");
  if (n->loss_type != loss_type || n->temperature != temperature || n->grad_scale != grad_scale) {   // baked into the captured step
    xemo_graph_destroy(n->g_train);
    xemo_graph_destroy(n->g_step);
    n->g_train = n->g_step = nullptr;
  }
  n->loss_type = loss_type; n->temperature = temperature; n->grad_scale = grad_scale;
  return XEMO_OK;
}

// ---- forward passes
extern "C" int xemo_teacher_forward(xemo_net* n, float* logits_out) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind != XEMO_NET_VGGVOX, "teacher_forward: a finalized teacher network is required");
  if (n->ctx->capturing) NET_OP(teacher_record(n));   // inside a caller's capture: record the kernels themselves
  else {
    if (!n->g_fwd) NET_OP(capture(n, &n->g_fwd, [&] { return teacher_record(n); }));
    NET_OP(xemo_graph_launch(n->ctx, n->g_fwd));
  }
  if (logits_out) NET_OP(copy_out_rows(n->ctx, n->get<float>("logits"), n->N, pad16(n->K), n->K, logits_out));
  return XEMO_OK;
}

extern "C" int xemo_student_forward(xemo_net* n, int train_mode, float* pred_out) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "student_forward: a finalized student network is required");
  if (train_mode || n->ctx->capturing) NET_OP(train_mode ? student_record_forward_train(n) : student_record_forward_test(n));
  else {
    if (!n->g_fwd) NET_OP(capture(n, &n->g_fwd, [&] { return student_record_forward_test(n); }));
    NET_OP(xemo_graph_launch(n->ctx, n->g_fwd));
  }
  if (pred_out) NET_OP(copy_out_rows(n->ctx, n->get<float>("pred32"), n->N, n->layers.back().kp, n->K, pred_out));
  return XEMO_OK;
}

// gradient exchange of the data-parallel step inside the recorded sequence: the backward pass is cut after fc6; the
// fc6..fc8 bucket (82 % of the bytes) is all-reduced on the communicator's forked stream while conv5..conv1 are
// differentiated, the head bucket after them on the same stream (one communicator: its operations stay ordered)
static int record_backward_with_exchange(xemo_net* n, xemo_comm* comm) {
  xemo_ctx* ctx = n->ctx;
  n->bm_scale = (comm && comm->world > 1) ? 1.f / float(comm->world) : 1.f;
  if (!comm || comm->world == 1) return student_record_backward(n, 0, int(n->layers.size()), true);
  NET_OP(student_record_backward(n, n->split_layer, int(n->layers.size()), true));
  NET_OP(xemo_stream_wait(ctx, comm->side, nullptr));          // fork
  NET_OP(xemo_set_stream(ctx, comm->side));
  int rc = xemo_comm_allreduce_f32(comm, n->grad + n->split_offset, n->nparam - n->split_offset);
  NET_OP(xemo_set_stream(ctx, nullptr));
  if (rc) return rc;
  NET_OP(student_record_backward(n, 0, n->split_layer, false));
  NET_OP(xemo_stream_wait(ctx, comm->side, nullptr));          // the head bucket needs conv5..conv1's gradients
  NET_OP(xemo_set_stream(ctx, comm->side));
  rc = xemo_comm_allreduce_f32(comm, n->grad, n->split_offset);
  if (!rc) rc = xemo_comm_allreduce_f32(comm, n->bm_flat, n->bm_elems);   // BN batch moments: sum over ranks, averaged in the update
  NET_OP(xemo_set_stream(ctx, nullptr));
  if (rc) return rc;
  return xemo_stream_wait(ctx, nullptr, comm->side);           // join
}

extern "C" int xemo_student_train_step(xemo_net* n, xemo_comm* comm) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "student_train_step: a finalized student network is required");
  auto record = [&] {
    NET_OP(student_record_forward_train(n));
    return record_backward_with_exchange(n, comm);
  };
  if (n->ctx->capturing) return record();
  if (!n->g_train || n->g_train_comm != comm) {
    xemo_graph_destroy(n->g_train);
    xemo_graph_destroy(n->g_update);      // (the moments-average scale 1 / ranks is baked into the captured update)
    n->g_train = n->g_update = nullptr;
    NET_OP(capture(n, &n->g_train, record));
    n->g_train_comm = comm;
    // the warm-up pass already ran the step once on these inputs (and accumulated its class counters): undo that part
    NET_OP(xemo_memset(n->ctx, n->buf["class_stats"], 0, size_t(2) * n->K * 4));
  }
  return xemo_graph_launch(n->ctx, n->g_train);
}

// hyper-parameters of the update live in device memory, so that the captured update follows the learning-rate schedule
extern "C" int xemo_net_set_hyper(xemo_net* n, float lr, float momentum, float weight_decay, int batch_size) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX && batch_size >= 1 && !n->ctx->capturing, "net_set_hyper: bad arguments");
  const float hy[4] = {lr, momentum, weight_decay, 1.f / float(batch_size)};
  // (a 16-byte pageable source is staged by the runtime before the call returns; the copy itself is stream-ordered)
  XEMO_CUDA(n->ctx, cudaMemcpyAsync(n->hyper, hy, sizeof(hy), cudaMemcpyHostToDevice, n->ctx->stream));
  return XEMO_OK;
}

extern "C" int xemo_sgd_step(xemo_net* n, float lr, float momentum, float weight_decay, int batch_size) {
  if (!n) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "sgd_step: a finalized student network is required");
  XEMO_REQUIRE(n->ctx, batch_size >= 1, "sgd_step: batch_size is the GLOBAL batch the summed gradient is divided by");
  if (n->ctx->capturing) return student_record_update(n);    // (hyper-parameters live in device memory: set them outside the capture)
  NET_OP(xemo_net_set_hyper(n, lr, momentum, weight_decay, batch_size));
  if (!n->g_update) NET_OP(capture(n, &n->g_update, [&] { return student_record_update(n); }, false));
  return xemo_graph_launch(n->ctx, n->g_update);
}

extern "C" int xemo_allreduce_grads(xemo_net* n, xemo_comm* comm) {
  if (!n || !comm) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX, "allreduce_grads: a finalized student network is required");
  return xemo_comm_allreduce_f32(comm, n->grad, n->nparam);
}

// teacher -> student coupling (emoVoxCeleb/getBatchEmoVoxCeleb.m:133-159,179-188): rows [start[i], end[i]) of the teacher's
// frame logits -> max / mean -> the student's target; start / end are device int arrays of the student's batch size
extern "C" int xemo_distill_set_windows(xemo_net* student, const int* start, const int* end) {
  if (!student || !start || !end) return XEMO_ERR_INVALID;
  xemo_net* n = student;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX && !is_device_ptr(start) && !is_device_ptr(end),
               "distill_set_windows: a finalized student network and host arrays are required");
  if (!n->win_start) { n->win_start = n->alloc<int>("win:start", n->N); n->win_end = n->alloc<int>("win:end", n->N); }
  XEMO_REQUIRE(n->ctx, n->win_start && n->win_end, "distill_set_windows: allocation failed");
  XEMO_CUDA(n->ctx, cudaStreamSynchronize(n->ctx->primary));
  XEMO_CUDA(n->ctx, cudaMemcpy(n->win_start, start, size_t(n->N) * 4, cudaMemcpyHostToDevice));
  XEMO_CUDA(n->ctx, cudaMemcpy(n->win_end, end, size_t(n->N) * 4, cudaMemcpyHostToDevice));
  return XEMO_OK;
}

extern "C" int xemo_distill_couple(xemo_net* teacher, xemo_net* student, const int* start, const int* end, int use_mean) {
  if (!teacher || !student) return XEMO_ERR_INVALID;
  if (!start && !end) { start = student->win_start; end = student->win_end; }
  XEMO_REQUIRE(student->ctx, start && end, "distill_couple: no frame windows (pass device arrays or call xemo_distill_set_windows)");
  XEMO_REQUIRE(student->ctx, teacher->finalized && student->finalized && teacher->kind != XEMO_NET_VGGVOX && student->kind == XEMO_NET_VGGVOX,
               "distill_couple: (teacher, student) networks required");
  XEMO_REQUIRE(student->ctx, student->K <= teacher->K, "distill_couple: numPredEmotions exceeds the teacher's outputs");
  return xemo_op_logit_aggregate(student->ctx, teacher->get<float>("logits"), pad16(teacher->K), start, end, student->N, student->K, use_mean,
                                 student->get<float>("target"));
}

// The whole distillation step as ONE captured graph: teacher forward -> coupling -> student forward (train-mode BN) ->
// loss -> backward (with the bucketed gradient exchange when `comm` spans several ranks) -> guarded SGD-momentum update.
// Inputs are whatever xemo_net_set_input put into the two networks.
extern "C" int xemo_distill_step(xemo_net* teacher, xemo_net* student, xemo_comm* comm, const int* start, const int* end, int use_mean,
                                 float lr, float momentum, float weight_decay, int batch_size) {
  if (!teacher || !student) return XEMO_ERR_INVALID;
  xemo_net* n = student;
  XEMO_REQUIRE(n->ctx, teacher->ctx == student->ctx, "distill_step: both networks must live in one context");
  XEMO_REQUIRE(n->ctx, !n->ctx->capturing, "distill_step: cannot be nested in a capture");
  if (!start && !end) { start = student->win_start; end = student->win_end; }
  XEMO_REQUIRE(n->ctx, start && end, "distill_step: no frame windows (pass device arrays or call xemo_distill_set_windows)");
  NET_OP(xemo_net_set_hyper(student, lr, momentum, weight_decay, batch_size));
  // teacher forward + coupling, then (or, with overlap, beside) the student's train-mode forward: they meet at the loss
  auto forward_both = [&] {
    const bool fork = student->use_overlap() && student->side_a;
    if (fork) {
      NET_OP(xemo_stream_wait(n->ctx, student->side_a, nullptr));
      NET_OP(xemo_set_stream(n->ctx, student->side_a));
    }
    int rc = teacher_record(teacher);
    if (!rc) rc = xemo_distill_couple(teacher, student, start, end, use_mean);
    if (fork) NET_OP(xemo_set_stream(n->ctx, nullptr));
    if (rc) return rc;
    NET_OP(student_record_forward_train(student));
    if (fork) NET_OP(xemo_stream_wait(n->ctx, nullptr, student->side_a));
    return int(XEMO_OK);
  };
  auto record = [&] {
    NET_OP(forward_both());
    NET_OP(record_backward_with_exchange(student, comm));
    return student_record_update(student);
  };
  if (!n->g_step || n->g_step_comm != comm || n->g_step_teacher != teacher || n->g_step_start != start || n->g_step_end != end ||
      n->g_step_mean != use_mean) {
    xemo_graph_destroy(n->g_step);
    n->g_step = nullptr;
    // eager pass without the update (kernel attributes), then the capture; the eager pass's class counters are undone
    NET_OP(forward_both());
    NET_OP(record_backward_with_exchange(student, comm));
    NET_OP(xemo_memset(n->ctx, n->buf["class_stats"], 0, size_t(2) * n->K * 4));
    NET_OP(capture(n, &n->g_step, record, false));
    n->g_step_comm = comm; n->g_step_teacher = teacher; n->g_step_start = start; n->g_step_end = end; n->g_step_mean = use_mean;
  }
  return xemo_graph_launch(n->ctx, n->g_step);
}

extern "C" int xemo_net_set_overlap(xemo_net* n, int mode) {
  if (!n || n->kind != XEMO_NET_VGGVOX) return XEMO_ERR_INVALID;
  if (mode != n->overlap) {     // baked into the captured sequences
    xemo_graph_destroy(n->g_train); xemo_graph_destroy(n->g_step);
    n->g_train = n->g_step = nullptr;
  }
  n->overlap = mode;
  return XEMO_OK;
}

extern "C" int xemo_net_num_kernels(xemo_net* n) {
  if (!n) return 0;
  int k = 0;
  for (xemo_graph* g : {n->g_fwd, n->g_train, n->g_update, n->g_step}) k += xemo_graph_num_kernels(g);
  return k;
}

extern "C" int xemo_net_reset_metrics(xemo_net* n) {
  if (!n || n->kind != XEMO_NET_VGGVOX || !n->finalized) return XEMO_ERR_INVALID;
  NET_OP(xemo_memset(n->ctx, n->buf["scalars"], 0, 8));
  return xemo_memset(n->ctx, n->buf["class_stats"], 0, size_t(2) * n->K * 4);
}

// out[0] objective, out[1] classerror (last step); out[2 .. 2+K) correct per class, out[2+K .. 2+2K) count per class
// (accumulated since reset); out[2+2K] non-finite gradient in the last update, out[3+2K] updates skipped so far
extern "C" int xemo_net_metrics(xemo_net* n, float* out, int n_out) {
  if (!n || !out) return XEMO_ERR_INVALID;
  XEMO_REQUIRE(n->ctx, n->finalized && n->kind == XEMO_NET_VGGVOX && n_out >= 4 + 2 * n->K, "net_metrics: student network and 4 + 2K floats required");
  XEMO_CUDA(n->ctx, cudaStreamSynchronize(n->ctx->primary));
  int g[3];
  XEMO_CUDA(n->ctx, cudaMemcpy(out, n->buf["scalars"], 8, cudaMemcpyDeviceToHost));
  XEMO_CUDA(n->ctx, cudaMemcpy(out + 2, n->buf["class_stats"], size_t(2) * n->K * 4, cudaMemcpyDeviceToHost));
  XEMO_CUDA(n->ctx, cudaMemcpy(g, n->guard, sizeof(g), cudaMemcpyDeviceToHost));
  out[2 + 2 * n->K] = float(g[0]);
  out[3 + 2 * n->K] = float(g[1]);
  return XEMO_OK;
}
