// ypb200 — engine (graph builder, planner, executor) and C ABI.  See include/ypb200.h.
//
// The engine is the native replacement of UPSTREAM ultralytics `AutoBackend.forward` +
// `SegmentationPredictor.postprocess` behind `model.predict(...)` (reference yolo_seg/app.py:91,
// yolo_seg/yolo_with_deva.py:51).  Topologies restate cfg/models/v8/yolov8-seg.yaml and
// cfg/models/v10/yolov10n.yaml; weight names follow the upstream state_dict (A.6).
//
// Data layout in HBM: every activation is NHWC bf16 inside ONE caller-allocated workspace arena;
// Concat / C2f-chunk / SPPF-cat never materialise: producers write channel slices of the consumer's
// buffer and consumers read channel slices through TMA coordinates.  The three head branches of a
// level share one fused first 3x3 conv.  The last 1x1 convs of the head write fp32 rows
// [64 box logits | nc class logits | 32 mask coefs] of a (B, A, no) buffer that decode/NMS consume.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ypb200.h"
#include "conv_plan.cuh"
#include "head_kernels.cuh"
#include "mask_kernels.cuh"
#include "misc_kernels.cuh"
#include "v10_kernels.cuh"
#if YPB_DIAG
#include "../../include/ypb200_diag.h"
#include "tma_bench.cuh"
#endif

using namespace ypb;

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(YPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace {

struct WeightEntry {
  std::string name;
  std::vector<int64_t> shape;
  std::vector<float> data;
  bool loaded = false, used = true;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

struct BufDesc {
  std::string name;
  int lvl = 0, C = 0, dtype = 0;  // dtype 0 bf16, 1 fp32
  int H = 0, W = 0;
  size_t offset = 0, bytes = 0;
};
struct View { int buf = -1, c_off = 0, C = 0; };

enum OpKind { OP_STEM, OP_CONV, OP_UPSAMPLE, OP_SPPF, OP_DW, OP_ATTN };
enum SrcKind { SRC_CONV_BN = 0, SRC_CONV_BIAS = 1, SRC_CONVT = 2 };
// slot > cout: the source's output channels are padded with zero weights / zero bias up to `slot` channels of the op's
// output (SiLU(0) = 0, so the padded channels of the activation stay zero): lets a layer with fewer than 16 channels
// (yolo11n's 8-channel bottleneck) run on the tensor-core kernels, whose channel counts are multiples of 16
struct ConvSrc { std::string mod; int kind; int cout; int slot = 0; int width() const { return slot > cout ? slot : cout; } };

struct Op {
  OpKind kind = OP_CONV;
  std::string name;
  std::vector<ConvSrc> srcs;
  int cin = 0, cout = 0, k = 1, s = 1, act = 1, out_mode = OUT_BF16;
  int cin_real = 0;  // > 0: the module's real input channels (< cin: the rest of the input slice is zero padding)
  View in, out, res;
  int head_lvl = -1, head_coff = 0;
  size_t w_off = 0, b_off = 0;  // byte offsets in the weight arena
  ConvLaunch L;    // launch of batch half 0 (the whole batch when the plan is not split)
  ConvLaunch L1;   // launch of batch half 1
  double flops = 0;  // algorithmic FLOPs of the op over the whole batch
  // OP_ATTN: in = qkv buffer view, res = pe view, out = output view
  int heads = 0;
};

struct NamedView { std::string name; View v; };

}  // namespace

struct ypb_engine {
  std::string spec;
  int nc = 80, nm = 0;
  bool end2end = false;
  std::vector<WeightEntry> weights;
  std::map<std::string, int> windex;
  std::vector<BufDesc> bufs;
  std::vector<Op> ops;
  std::vector<NamedView> views;
  View feat[3];  // P3,P4,P5 (for anchors)
  int proto_buf = -1;
  // weights on device
  int device = -1;
  void* w_arena = nullptr;
  size_t w_bytes = 0;
  bool finalized = false;
  // plan
  int B = 0, H = 0, W = 0, A = 0, no = 0;
  HeadGeom hg{};
  size_t ws_bytes = 0;
  size_t off_head = 0, off_dbox = 0, off_dcls = 0, off_keys = 0, off_count = 0, off_moff = 0;
  bool planned = false, bound = false;
  // Batch halves: with B >= 4 the plan holds TWO launch sets over images [0, sB[0]) and [sB[0], B).  In the captured
  // graph they are independent chains, so the ~10 us a layer spends in launch latency, prologue, first operand fetch,
  // epilogue drain and wave quantisation is covered by the other half's tensor work instead of idling the GPU.
  int n_split = 1, sB[2] = {0, 0}, sb0[2] = {0, 0};
  uint8_t* ws = nullptr;
  int conv_impl = 0;
  int launches = 0;
  double flops = 0;
  // CUDA graph of one ypb_infer(): captured once per distinct argument set, then replayed
  struct GraphKey {
    const void *frames, *xform, *det, *det_lb, *keep, *coef, *count, *cmask;
    float conf, iou;
    int max_det, agnostic, impl;
    bool operator==(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) == 0; }
  };
  bool use_graph = true;
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec; unsigned long long stamp; };
  std::vector<GraphEntry> graphs;  // small LRU cache: callers alternate between a few input buffers
  unsigned long long graph_clock = 0;
  cudaStream_t cap_stream = nullptr;
  // Branch-parallel capture: the layer ops form a DAG (proto branch, the three head levels and the rest of the neck
  // only meet at decode); each op is pinned to one of a few capture streams so that independent chains become parallel
  // branches of the CUDA graph and fill each other's tail waves / launch gaps.  Built once per topology.
  struct OpSched { int stream = 0; std::vector<int> waits; bool record = false; };
  std::vector<OpSched> sched;
  int n_streams = 1;
  bool branch_parallel = true;
  cudaStream_t side_streams[16] = {};
  std::vector<cudaEvent_t> op_events;
  void build_schedule(int max_streams);
  void drop_graph() {
    for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }

  // ------------------------------------------------------------------ builder helpers
  int add_weight(const std::string& name, std::vector<int64_t> shape, bool used = true) {
    WeightEntry w;
    w.name = name; w.shape = std::move(shape); w.used = used;
    weights.push_back(std::move(w));
    windex[name] = (int)weights.size() - 1;
    return (int)weights.size() - 1;
  }
  void add_conv_weights(const ConvSrc& s, int cin, int k, bool used = true) {
    if (s.kind == SRC_CONV_BN) {
      add_weight(s.mod + ".conv.weight", {s.cout, cin, k, k}, used);
      for (const char* leaf : {"weight", "bias", "running_mean", "running_var"})
        add_weight(s.mod + ".bn." + leaf, {s.cout}, used);
    } else if (s.kind == SRC_CONV_BIAS) {
      add_weight(s.mod + ".weight", {s.cout, cin, k, k}, used);
      add_weight(s.mod + ".bias", {s.cout}, used);
    } else {
      add_weight(s.mod + ".weight", {cin, s.cout, 2, 2}, used);
      add_weight(s.mod + ".bias", {s.cout}, used);
    }
  }
  int new_buf(const std::string& name, int lvl, int C, int dtype = 0) {
    BufDesc b;
    b.name = name; b.lvl = lvl; b.C = C; b.dtype = dtype;
    bufs.push_back(b);
    return (int)bufs.size() - 1;
  }
  View whole(int buf) const { return View{buf, 0, bufs[buf].C}; }
  View slice(int buf, int off, int C) const { return View{buf, off, C}; }
  void name_view(const std::string& n, View v) { views.push_back({n, v}); }

  // Conv(+BN)+SiLU, possibly several modules fused along Cout (they must share input, k, s).
  void conv(const std::vector<std::string>& mods, const std::vector<int>& couts, View in, View out, int k, int s,
            bool act = true, View res = View(), int cin_real = 0, int slot = 0) {
    Op op;
    op.kind = OP_CONV; op.name = mods[0];
    op.cin = in.C; op.k = k; op.s = s; op.act = act ? 1 : 0; op.in = in; op.out = out; op.res = res;
    op.cin_real = cin_real;
    int tot = 0;
    for (size_t i = 0; i < mods.size(); ++i) {
      ConvSrc src{mods[i], SRC_CONV_BN, couts[i]};
      src.slot = slot;
      add_conv_weights(src, cin_real > 0 ? cin_real : in.C, k);
      op.srcs.push_back(src);
      tot += src.width();
    }
    op.cout = tot;
    ops.push_back(op);
  }
  void conv1(const std::string& mod, View in, View out, int k, int s, bool act = true, View res = View()) {
    conv({mod}, {out.C}, in, out, k, s, act, res);
  }
  // plain nn.Conv2d(k=1, bias) writing fp32 head rows
  void head_out(const std::string& mod, View in, int cout, int lvl, int coff) {
    Op op;
    op.kind = OP_CONV; op.name = mod;
    op.cin = in.C; op.cout = cout; op.k = 1; op.s = 1; op.act = 0; op.in = in; op.out_mode = OUT_F32;
    op.head_lvl = lvl; op.head_coff = coff;
    ConvSrc src{mod, SRC_CONV_BIAS, cout};
    add_conv_weights(src, in.C, 1);
    op.srcs.push_back(src);
    ops.push_back(op);
  }

  // UPSTREAM block.py::C2f
  void c2f(const std::string& mod, View in, View out, int n, bool shortcut, int lvl) {
    const int c = out.C / 2;
    const int Y = new_buf(mod + ".cat", lvl, (2 + n) * c);
    conv1(mod + ".cv1", in, slice(Y, 0, 2 * c), 1, 1);
    for (int j = 0; j < n; ++j) {
      const int t = new_buf(mod + ".m" + std::to_string(j) + ".t", lvl, c);
      const View x = slice(Y, (1 + j) * c, c);
      conv1(mod + ".m." + std::to_string(j) + ".cv1", x, whole(t), 3, 1);
      conv1(mod + ".m." + std::to_string(j) + ".cv2", whole(t), slice(Y, (2 + j) * c, c), 3, 1, true,
            shortcut ? x : View());
    }
    conv1(mod + ".cv2", whole(Y), out, 1, 1);
  }
  // UPSTREAM block.py::SPPF
  void sppf(const std::string& mod, View in, View out, int lvl) {
    const int c_ = in.C / 2;
    const int S = new_buf(mod + ".cat", lvl, 4 * c_);
    conv1(mod + ".cv1", in, slice(S, 0, c_), 1, 1);
    Op op;
    op.kind = OP_SPPF; op.name = mod + ".pool"; op.in = slice(S, 0, c_); op.out = whole(S);
    ops.push_back(op);
    conv1(mod + ".cv2", whole(S), out, 1, 1);
  }
  // depthwise Conv(+BN)(+SiLU); several modules may be summed into one kernel (RepVGGDW: 7x7 + 3x3)
  void dwconv(const std::vector<std::string>& mods, const std::vector<int>& ks, View in, View out, int k, int s,
              bool act, View res = View()) {
    Op op;
    op.kind = OP_DW; op.name = mods[0];
    op.cin = in.C; op.cout = in.C; op.k = k; op.s = s; op.act = act ? 1 : 0; op.in = in; op.out = out; op.res = res;
    for (size_t i = 0; i < mods.size(); ++i) {
      ConvSrc src{mods[i], SRC_CONV_BN, in.C};
      add_conv_weights(src, 1, ks[i]);
      op.srcs.push_back(src);
    }
    ops.push_back(op);
  }
  // plain Conv(+BN), no activation, optional residual (Attention.proj / PSA.ffn[1])
  void conv_noact(const std::string& mod, View in, View out, View res = View()) { conv1(mod, in, out, 1, 1, false, res); }
  // UPSTREAM block.py::SCDown
  void scdown(const std::string& mod, View in, View out, int lvl_in) {
    const int t = new_buf(mod + ".t", lvl_in, out.C);
    conv1(mod + ".cv1", in, whole(t), 1, 1);
    dwconv({mod + ".cv2"}, {3}, whole(t), out, 3, 2, false);
  }
  // UPSTREAM block.py::PSA (+ Attention), heads = c/64, key_dim 32, head_dim 64
  void psa(const std::string& mod, View in, View out, int lvl) {
    const int c = in.C / 2;
    const int P = new_buf(mod + ".ab", lvl, 2 * c);
    conv1(mod + ".cv1", in, whole(P), 1, 1);
    psa_block(mod, slice(P, c, c), lvl);
    conv1(mod + ".cv2", whole(P), out, 1, 1);
  }
  // UPSTREAM block.py::C2PSA: PSA with n PSABlocks named m.{j}
  void c2psa(const std::string& mod, View in, View out, int n, int lvl) {
    const int c = in.C / 2;
    const int P = new_buf(mod + ".ab", lvl, 2 * c);
    conv1(mod + ".cv1", in, whole(P), 1, 1);
    for (int j = 0; j < n; ++j) psa_block(mod + ".m." + std::to_string(j), slice(P, c, c), lvl);
    conv1(mod + ".cv2", whole(P), out, 1, 1);
  }
  // b = b + attn(b); b = b + ffn(b), in place on the channel slice b (PSA body / PSABlock)
  void psa_block(const std::string& mod, View b, int lvl) {
    const int c = b.C, nh = c / 64;
    const int Q = new_buf(mod + ".qkv", lvl, 2 * c), PE = new_buf(mod + ".pe", lvl, c);
    const int AO = new_buf(mod + ".ao", lvl, c), F = new_buf(mod + ".ffn", lvl, 2 * c);
    conv_noact(mod + ".attn.qkv", b, whole(Q));
    {  // pe: depthwise 3x3 over v (the v channels of each head are a strided subset of the qkv buffer)
      Op op;
      op.kind = OP_DW; op.name = mod + ".attn.pe";
      op.cin = c; op.cout = c; op.k = 3; op.s = 1; op.act = 0; op.in = whole(Q); op.out = whole(PE); op.heads = nh;
      ConvSrc src{mod + ".attn.pe", SRC_CONV_BN, c};
      add_conv_weights(src, 1, 3);
      op.srcs.push_back(src);
      ops.push_back(op);
    }
    {
      Op op;
      op.kind = OP_ATTN; op.name = mod + ".attn"; op.in = whole(Q); op.res = whole(PE); op.out = whole(AO); op.heads = nh;
      ops.push_back(op);
    }
    conv_noact(mod + ".attn.proj", whole(AO), b, b);          // b = b + attn(b), in place
    conv1(mod + ".ffn.0", b, whole(F), 1, 1);
    conv_noact(mod + ".ffn.1", whole(F), b, b);               // b = b + ffn(b), in place
  }
  static int pad16(int c) { return (c + 15) & ~15; }
  // UPSTREAM block.py::Bottleneck(c, c, shortcut, k=(3,3), e): t = cv1(x) (c_ = c*e channels), out = [x +] cv2(t)
  void bottleneck(const std::string& mod, View x, View out, bool shortcut, int c_mid, int lvl) {
    const int cp = pad16(c_mid);
    const int t = new_buf(mod + ".t", lvl, cp);
    conv({mod + ".cv1"}, {c_mid}, x, whole(t), 3, 1, true, View(), 0, cp);
    conv({mod + ".cv2"}, {out.C}, whole(t), out, 3, 1, true, shortcut ? x : View(), cp != c_mid ? c_mid : 0, 0);
  }
  // UPSTREAM block.py::C3k(c, c, n, shortcut, e=0.5): cv3(cat(m(cv1(x)), cv2(x))), m = n x Bottleneck(c_, c_, e=1.0).
  // cv1 and cv2 read the same input: one GEMM writes [a | b]; the bottleneck chain updates a in place.
  void c3k(const std::string& mod, View x, View out, int n, bool shortcut, int lvl) {
    const int c_ = x.C / 2;
    const int Z = new_buf(mod + ".ab", lvl, 2 * c_);
    conv({mod + ".cv1", mod + ".cv2"}, {c_, c_}, x, whole(Z), 1, 1);
    const View a = slice(Z, 0, c_);
    for (int j = 0; j < n; ++j) bottleneck(mod + ".m." + std::to_string(j), a, a, shortcut, c_, lvl);
    conv1(mod + ".cv3", whole(Z), out, 1, 1);
  }
  // UPSTREAM block.py::C3k2(c1, c2, n, c3k, e, shortcut=True): C2f whose blocks are C3k(c, c, 2) or Bottleneck(c, c, e=0.5)
  void c3k2(const std::string& mod, View in, View out, int n, bool use_c3k, double e, int lvl) {
    const int c = (int)(out.C * e);
    const int Y = new_buf(mod + ".cat", lvl, (2 + n) * c);
    conv1(mod + ".cv1", in, slice(Y, 0, 2 * c), 1, 1);
    for (int j = 0; j < n; ++j) {
      const View x = slice(Y, (1 + j) * c, c), y = slice(Y, (2 + j) * c, c);
      const std::string mj = mod + ".m." + std::to_string(j);
      if (use_c3k) c3k(mj, x, y, 2, true, lvl);
      else bottleneck(mj, x, y, true, c / 2, lvl);
    }
    conv1(mod + ".cv2", whole(Y), out, 1, 1);
  }
  // UPSTREAM block.py::C2fCIB with CIB(c, c, shortcut, e=1.0, lk)
  void c2fcib(const std::string& mod, View in, View out, int n, bool shortcut, bool lk, int lvl) {
    const int c = out.C / 2;
    const int Y = new_buf(mod + ".cat", lvl, (2 + n) * c);
    conv1(mod + ".cv1", in, slice(Y, 0, 2 * c), 1, 1);
    for (int j = 0; j < n; ++j) {
      const std::string cm = mod + ".m." + std::to_string(j) + ".cv1.";
      const View x = slice(Y, (1 + j) * c, c);
      const int t0 = new_buf(cm + "t0", lvl, c), t1 = new_buf(cm + "t1", lvl, 2 * c), t2 = new_buf(cm + "t2", lvl, 2 * c);
      const int t3 = new_buf(cm + "t3", lvl, c);
      dwconv({cm + "0"}, {3}, x, whole(t0), 3, 1, true);
      conv1(cm + "1", whole(t0), whole(t1), 1, 1);
      if (lk) dwconv({cm + "2.conv", cm + "2.conv1"}, {7, 3}, whole(t1), whole(t2), 7, 1, true);
      else dwconv({cm + "2"}, {3}, whole(t1), whole(t2), 3, 1, true);
      conv1(cm + "3", whole(t2), whole(t3), 1, 1);
      dwconv({cm + "4"}, {3}, whole(t3), slice(Y, (2 + j) * c, c), 3, 1, true, shortcut ? x : View());
    }
    conv1(mod + ".cv2", whole(Y), out, 1, 1);
  }
  void upsample(View in, View out) {
    Op op;
    op.kind = OP_UPSAMPLE; op.name = "upsample"; op.in = in; op.out = out;
    ops.push_back(op);
  }
};

// ------------------------------------------------------------------------------------------------
// topologies
// ------------------------------------------------------------------------------------------------
static int make_div8(double x) { return (int)(std::ceil(x / 8.0) * 8); }

// Zero channels to append to a fused conv's N so that it splits into equal tiles of a multiple of 16 channels, each
// <= 256 wide and at least 128 wide (see conv_plan_geometry's split rule).  0 when N already splits well.
static int fused_n_pad(int n) {
  for (int pad = 0; pad <= 64; pad += 16) {
    const int t = n + pad;
    if (t <= 256) return pad;
    const int splits = (t + 255) / 256;
    if (t % (16 * splits) == 0) return pad;
  }
  return 0;
}

static bool build_v8seg(ypb_engine& e, char scale) {
  double d, w; int mc;
  switch (scale) {
    case 'n': d = 0.33; w = 0.25; mc = 1024; break;
    case 's': d = 0.33; w = 0.50; mc = 1024; break;
    case 'm': d = 0.67; w = 0.75; mc = 768; break;
    case 'l': d = 1.00; w = 1.00; mc = 512; break;
    case 'x': d = 1.00; w = 1.25; mc = 512; break;
    default: return false;
  }
  auto ch = [&](int c) { return make_div8(std::min(c, mc) * w); };
  auto rep = [&](int n) { return std::max((int)std::lround(n * d), 1); };
  const int c64 = ch(64), c128 = ch(128), c256 = ch(256), c512 = ch(512), c1024 = ch(1024);
  const int n3 = rep(3), n6 = rep(6);
  e.nm = 32;
  const std::string M = "model.";
  // buffers (lvl = log2 of the downscale); concat inputs are slices of the concat buffer
  const int x0 = e.new_buf("model.0", 1, c64), x1 = e.new_buf("model.1", 2, c128), x2 = e.new_buf("model.2", 2, c128);
  const int x3 = e.new_buf("model.3", 3, c256);
  const int cat14 = e.new_buf("model.14", 3, c512 + c256);    // [up(12) | 4]
  const int x5 = e.new_buf("model.5", 4, c512);
  const int cat11 = e.new_buf("model.11", 4, c1024 + c512);   // [up(9) | 6]
  const int x7 = e.new_buf("model.7", 5, c1024), x8 = e.new_buf("model.8", 5, c1024);
  const int cat20 = e.new_buf("model.20", 5, c512 + c1024);   // [19 | 9]
  const int cat17 = e.new_buf("model.17", 4, c256 + c512);    // [16 | 12]
  const int x15 = e.new_buf("model.15", 3, c256), x18 = e.new_buf("model.18", 4, c512), x21 = e.new_buf("model.21", 5, c1024);
  const View v4 = e.slice(cat14, c512, c256), v6 = e.slice(cat11, c1024, c512), v9 = e.slice(cat20, c512, c1024);
  const View v12 = e.slice(cat17, c256, c512), v16 = e.slice(cat17, 0, c256), v19 = e.slice(cat20, 0, c512);

  {  // stem
    Op op;
    op.kind = OP_STEM; op.name = "model.0"; op.cin = 3; op.cout = c64; op.k = 3; op.s = 2; op.out = e.whole(x0);
    ConvSrc src{"model.0", SRC_CONV_BN, c64};
    e.add_conv_weights(src, 3, 3);
    op.srcs.push_back(src);
    e.ops.push_back(op);
  }
  e.conv1(M + "1", e.whole(x0), e.whole(x1), 3, 2);
  e.c2f(M + "2", e.whole(x1), e.whole(x2), n3, true, 2);
  e.conv1(M + "3", e.whole(x2), e.whole(x3), 3, 2);
  e.c2f(M + "4", e.whole(x3), v4, n6, true, 3);
  e.conv1(M + "5", v4, e.whole(x5), 3, 2);
  e.c2f(M + "6", e.whole(x5), v6, n6, true, 4);
  e.conv1(M + "7", v6, e.whole(x7), 3, 2);
  e.c2f(M + "8", e.whole(x7), e.whole(x8), n3, true, 5);
  e.sppf(M + "9", e.whole(x8), v9, 5);
  e.upsample(v9, e.slice(cat11, 0, c1024));
  e.c2f(M + "12", e.whole(cat11), v12, n3, false, 4);
  e.upsample(v12, e.slice(cat14, 0, c512));
  e.c2f(M + "15", e.whole(cat14), e.whole(x15), n3, false, 3);
  e.conv1(M + "16", e.whole(x15), v16, 3, 2);
  e.c2f(M + "18", e.whole(cat17), e.whole(x18), n3, false, 4);
  e.conv1(M + "19", e.whole(x18), v19, 3, 2);
  e.c2f(M + "21", e.whole(cat20), e.whole(x21), n3, false, 5);
  e.name_view("model.0", e.whole(x0)); e.name_view("model.1", e.whole(x1)); e.name_view("model.2", e.whole(x2));
  e.name_view("model.3", e.whole(x3)); e.name_view("model.4", v4); e.name_view("model.5", e.whole(x5));
  e.name_view("model.6", v6); e.name_view("model.7", e.whole(x7)); e.name_view("model.8", e.whole(x8));
  e.name_view("model.9", v9); e.name_view("model.11", e.whole(cat11)); e.name_view("model.12", v12);
  e.name_view("model.14", e.whole(cat14)); e.name_view("model.15", e.whole(x15)); e.name_view("model.16", v16);
  e.name_view("model.17", e.whole(cat17)); e.name_view("model.18", e.whole(x18)); e.name_view("model.19", v19);
  e.name_view("model.20", e.whole(cat20)); e.name_view("model.21", e.whole(x21));

  // Segment head (UPSTREAM head.py::Segment), module index 22
  const std::string Hd = "model.22.";
  const int chs[3] = {c256, c512, c1024};
  const View P[3] = {e.whole(x15), e.whole(x18), e.whole(x21)};
  const int hc2 = std::max(std::max(16, chs[0] / 4), 64), hc3 = std::max(chs[0], std::min(e.nc, 100));
  const int hc4 = std::max(chs[0] / 4, e.nm);
  // Proto first: it only needs P3, so in the branch-parallel graph it starts while the neck is still running
  e.add_weight(Hd + "dfl.conv.weight", {1, 16, 1, 1}, false);
  // Proto (UPSTREAM block.py::Proto)
  const int npr = ch(256);
  const int p1 = e.new_buf(Hd + "proto.cv1", 3, npr), p2 = e.new_buf(Hd + "proto.upsample", 2, npr);
  const int p3 = e.new_buf(Hd + "proto.cv2", 2, npr), pr = e.new_buf("proto", 2, e.nm, 1);
  e.conv1(Hd + "proto.cv1", P[0], e.whole(p1), 3, 1);
  {
    Op op;
    op.kind = OP_CONV; op.name = Hd + "proto.upsample";
    op.cin = npr; op.cout = 4 * npr; op.k = 1; op.s = 1; op.act = 0; op.out_mode = OUT_SHUFFLE2_BF16;
    op.in = e.whole(p1); op.out = e.whole(p2);
    ConvSrc src{Hd + "proto.upsample", SRC_CONVT, npr};
    e.add_conv_weights(src, npr, 2);
    op.srcs.push_back(src);
    e.ops.push_back(op);
  }
  e.conv1(Hd + "proto.cv2", e.whole(p2), e.whole(p3), 3, 1);
  {
    e.conv1(Hd + "proto.cv3", e.whole(p3), e.whole(pr), 1, 1);
    e.ops.back().out_mode = OUT_F32;
  }
  for (int i = 0; i < 3; ++i) {
    e.feat[i] = P[i];
    const std::string si = std::to_string(i);
    const int lvl = 3 + i;
    // the first 3x3 conv of the box / class / coef branches reads the same P_i: one fused GEMM, N = hc2+hc3+hc4.
    // N > 256 is split into equal tiles of a multiple of 16 channels; yolov8m-seg's 64+192+48 = 304 = 16 x 19 has no
    // such split above 16 (the planner used to fall back to NINETEEN 16-channel tiles: 0.15 of tensor peak), so the
    // fused output is padded with zero-weight channels to the next total that splits evenly (304 -> 320 = 2 x 160)
    const int fpad = fused_n_pad(hc2 + hc3 + hc4);
    const int f0 = e.new_buf(Hd + "lvl" + si + ".s0", lvl, hc2 + hc3 + hc4 + fpad);
    e.conv({Hd + "cv2." + si + ".0", Hd + "cv3." + si + ".0", Hd + "cv4." + si + ".0"}, {hc2, hc3, hc4}, P[i], e.whole(f0), 3, 1);
    if (fpad) { e.ops.back().srcs.back().slot = hc4 + fpad; e.ops.back().cout += fpad; }
    const int t2 = e.new_buf(Hd + "cv2." + si + ".t", lvl, hc2), t3 = e.new_buf(Hd + "cv3." + si + ".t", lvl, hc3);
    const int t4 = e.new_buf(Hd + "cv4." + si + ".t", lvl, hc4);
    e.conv1(Hd + "cv2." + si + ".1", e.slice(f0, 0, hc2), e.whole(t2), 3, 1);
    e.conv1(Hd + "cv3." + si + ".1", e.slice(f0, hc2, hc3), e.whole(t3), 3, 1);
    e.conv1(Hd + "cv4." + si + ".1", e.slice(f0, hc2 + hc3, hc4), e.whole(t4), 3, 1);
    e.head_out(Hd + "cv2." + si + ".2", e.whole(t2), 64, i, 0);
    e.head_out(Hd + "cv3." + si + ".2", e.whole(t3), e.nc, i, 64);
    e.head_out(Hd + "cv4." + si + ".2", e.whole(t4), e.nm, i, 64 + e.nc);
  }
  e.proto_buf = pr;
  e.name_view("proto", e.whole(pr));
  return true;
}


static bool build_v10n(ypb_engine& e) {
  const double w = 0.25;
  auto ch = [&](int c) { return make_div8(std::min(c, 1024) * w); };
  const int c64 = ch(64), c128 = ch(128), c256 = ch(256), c512 = ch(512), c1024 = ch(1024);
  e.nm = 0; e.end2end = true;
  const std::string M = "model.";
  const int x0 = e.new_buf("model.0", 1, c64), x1 = e.new_buf("model.1", 2, c128), x2 = e.new_buf("model.2", 2, c128);
  const int x3 = e.new_buf("model.3", 3, c256);
  const int cat15 = e.new_buf("model.15", 3, c512 + c256);   // [up(13) | 4]
  const int x5 = e.new_buf("model.5", 4, c512);
  const int cat12 = e.new_buf("model.12", 4, c1024 + c512);  // [up(10) | 6]
  const int x7 = e.new_buf("model.7", 5, c1024), x8 = e.new_buf("model.8", 5, c1024), x9 = e.new_buf("model.9", 5, c1024);
  const int cat21 = e.new_buf("model.21", 5, c512 + c1024);  // [20 | 10]
  const int cat18 = e.new_buf("model.18", 4, c256 + c512);   // [17 | 13]
  const int x16 = e.new_buf("model.16", 3, c256), x19 = e.new_buf("model.19", 4, c512), x22 = e.new_buf("model.22", 5, c1024);
  const View v4 = e.slice(cat15, c512, c256), v6 = e.slice(cat12, c1024, c512), v10 = e.slice(cat21, c512, c1024);
  const View v13 = e.slice(cat18, c256, c512), v17 = e.slice(cat18, 0, c256), v20 = e.slice(cat21, 0, c512);
  {
    Op op;
    op.kind = OP_STEM; op.name = "model.0"; op.cin = 3; op.cout = c64; op.k = 3; op.s = 2; op.out = e.whole(x0);
    ConvSrc src{"model.0", SRC_CONV_BN, c64};
    e.add_conv_weights(src, 3, 3);
    op.srcs.push_back(src);
    e.ops.push_back(op);
  }
  e.conv1(M + "1", e.whole(x0), e.whole(x1), 3, 2);
  e.c2f(M + "2", e.whole(x1), e.whole(x2), 1, true, 2);
  e.conv1(M + "3", e.whole(x2), e.whole(x3), 3, 2);
  e.c2f(M + "4", e.whole(x3), v4, 2, true, 3);
  e.scdown(M + "5", v4, e.whole(x5), 3);
  e.c2f(M + "6", e.whole(x5), v6, 2, true, 4);
  e.scdown(M + "7", v6, e.whole(x7), 4);
  e.c2f(M + "8", e.whole(x7), e.whole(x8), 1, true, 5);
  e.sppf(M + "9", e.whole(x8), e.whole(x9), 5);
  e.psa(M + "10", e.whole(x9), v10, 5);
  e.upsample(v10, e.slice(cat12, 0, c1024));
  e.c2f(M + "13", e.whole(cat12), v13, 1, false, 4);
  e.upsample(v13, e.slice(cat15, 0, c512));
  e.c2f(M + "16", e.whole(cat15), e.whole(x16), 1, false, 3);
  e.conv1(M + "17", e.whole(x16), v17, 3, 2);
  e.c2f(M + "19", e.whole(cat18), e.whole(x19), 1, false, 4);
  e.scdown(M + "20", e.whole(x19), v20, 4);
  e.c2fcib(M + "22", e.whole(cat21), e.whole(x22), 1, true, true, 5);
  e.name_view("model.0", e.whole(x0)); e.name_view("model.1", e.whole(x1)); e.name_view("model.2", e.whole(x2));
  e.name_view("model.3", e.whole(x3)); e.name_view("model.4", v4); e.name_view("model.5", e.whole(x5));
  e.name_view("model.6", v6); e.name_view("model.7", e.whole(x7)); e.name_view("model.8", e.whole(x8));
  e.name_view("model.9", e.whole(x9)); e.name_view("model.10", v10); e.name_view("model.12", e.whole(cat12));
  e.name_view("model.13", v13); e.name_view("model.15", e.whole(cat15)); e.name_view("model.16", e.whole(x16));
  e.name_view("model.17", v17); e.name_view("model.18", e.whole(cat18)); e.name_view("model.19", e.whole(x19));
  e.name_view("model.20", v20); e.name_view("model.21", e.whole(cat21)); e.name_view("model.22", e.whole(x22));

  // v10Detect (UPSTREAM head.py::v10Detect): only the one-to-one branches run at inference
  const std::string Hd = "model.23.";
  const int chs[3] = {c256, c512, c1024};
  const View P[3] = {e.whole(x16), e.whole(x19), e.whole(x22)};
  const int hc2 = std::max(std::max(16, chs[0] / 4), 64), hc3 = std::max(chs[0], std::min(e.nc, 100));
  for (int i = 0; i < 3; ++i) {
    e.feat[i] = P[i];
    const std::string si = std::to_string(i);
    const int lvl = 3 + i, x = chs[i];
    const std::string b2 = Hd + "one2one_cv2." + si + ".", b3 = Hd + "one2one_cv3." + si + ".";
    const int t0 = e.new_buf(b2 + "t0", lvl, hc2), t1 = e.new_buf(b2 + "t1", lvl, hc2);
    e.conv1(b2 + "0", P[i], e.whole(t0), 3, 1);
    e.conv1(b2 + "1", e.whole(t0), e.whole(t1), 3, 1);
    e.head_out(b2 + "2", e.whole(t1), 64, i, 0);
    const int u0 = e.new_buf(b3 + "u0", lvl, x), u1 = e.new_buf(b3 + "u1", lvl, hc3), u2 = e.new_buf(b3 + "u2", lvl, hc3);
    const int u3 = e.new_buf(b3 + "u3", lvl, hc3);
    e.dwconv({b3 + "0.0"}, {3}, P[i], e.whole(u0), 3, 1, true);
    e.conv1(b3 + "0.1", e.whole(u0), e.whole(u1), 1, 1);
    e.dwconv({b3 + "1.0"}, {3}, e.whole(u1), e.whole(u2), 3, 1, true);
    e.conv1(b3 + "1.1", e.whole(u2), e.whole(u3), 1, 1);
    e.head_out(b3 + "2", e.whole(u3), e.nc, i, 64);
    // one-to-many twins: in every checkpoint, never executed
    const std::string m2 = Hd + "cv2." + si + ".", m3 = Hd + "cv3." + si + ".";
    e.add_conv_weights(ConvSrc{m2 + "0", SRC_CONV_BN, hc2}, x, 3, false);
    e.add_conv_weights(ConvSrc{m2 + "1", SRC_CONV_BN, hc2}, hc2, 3, false);
    e.add_conv_weights(ConvSrc{m2 + "2", SRC_CONV_BIAS, 64}, hc2, 1, false);
    e.add_conv_weights(ConvSrc{m3 + "0.0", SRC_CONV_BN, x}, 1, 3, false);
    e.add_conv_weights(ConvSrc{m3 + "0.1", SRC_CONV_BN, hc3}, x, 1, false);
    e.add_conv_weights(ConvSrc{m3 + "1.0", SRC_CONV_BN, hc3}, 1, 3, false);
    e.add_conv_weights(ConvSrc{m3 + "1.1", SRC_CONV_BN, hc3}, hc3, 1, false);
    e.add_conv_weights(ConvSrc{m3 + "2", SRC_CONV_BIAS, e.nc}, hc3, 1, false);
  }
  e.add_weight(Hd + "dfl.conv.weight", {1, 16, 1, 1}, false);
  return true;
}

// UPSTREAM cfg/models/11/yolo11-seg.yaml: the app's default weights are yolo11{n,x}-seg
// (reference yolo_seg/app.py:216-223, yolo_with_deva.py:226).  Same skeleton as yolov10n (C2PSA at layer 10, head at
// 23); C3k2 blocks, non-legacy (depthwise) class branch, Segment head.
static bool build_v11seg(ypb_engine& e, char scale) {
  double d, w; int mc;
  switch (scale) {
    case 'n': d = 0.50; w = 0.25; mc = 1024; break;
    case 's': d = 0.50; w = 0.50; mc = 1024; break;
    case 'm': d = 0.50; w = 1.00; mc = 512; break;
    case 'l': d = 1.00; w = 1.00; mc = 512; break;
    case 'x': d = 1.00; w = 1.50; mc = 512; break;
    default: return false;
  }
  auto ch = [&](int c) { return make_div8(std::min(c, mc) * w); };
  const int n2 = std::max((int)std::lround(2 * d), 1);
  const bool big = scale == 'm' || scale == 'l' || scale == 'x';  // parse_model forces c3k=True for m/l/x
  const int c64 = ch(64), c128 = ch(128), c256 = ch(256), c512 = ch(512), c1024 = ch(1024);
  e.nm = 32;
  const std::string M = "model.";
  const int x0 = e.new_buf("model.0", 1, c64), x1 = e.new_buf("model.1", 2, c128), x2 = e.new_buf("model.2", 2, c256);
  const int x3 = e.new_buf("model.3", 3, c256);
  const int cat15 = e.new_buf("model.15", 3, c512 + c512);    // [up(13) | 4]
  const int x5 = e.new_buf("model.5", 4, c512);
  const int cat12 = e.new_buf("model.12", 4, c1024 + c512);   // [up(10) | 6]
  const int x7 = e.new_buf("model.7", 5, c1024), x8 = e.new_buf("model.8", 5, c1024), x9 = e.new_buf("model.9", 5, c1024);
  const int cat21 = e.new_buf("model.21", 5, c512 + c1024);   // [20 | 10]
  const int cat18 = e.new_buf("model.18", 4, c256 + c512);    // [17 | 13]
  const int x16 = e.new_buf("model.16", 3, c256), x19 = e.new_buf("model.19", 4, c512), x22 = e.new_buf("model.22", 5, c1024);
  const View v4 = e.slice(cat15, c512, c512), v6 = e.slice(cat12, c1024, c512), v10 = e.slice(cat21, c512, c1024);
  const View v13 = e.slice(cat18, c256, c512), v17 = e.slice(cat18, 0, c256), v20 = e.slice(cat21, 0, c512);
  {
    Op op;
    op.kind = OP_STEM; op.name = "model.0"; op.cin = 3; op.cout = c64; op.k = 3; op.s = 2; op.out = e.whole(x0);
    ConvSrc src{"model.0", SRC_CONV_BN, c64};
    e.add_conv_weights(src, 3, 3);
    op.srcs.push_back(src);
    e.ops.push_back(op);
  }
  e.conv1(M + "1", e.whole(x0), e.whole(x1), 3, 2);
  e.c3k2(M + "2", e.whole(x1), e.whole(x2), n2, big, 0.25, 2);       // C3k2(256, False, 0.25)
  e.conv1(M + "3", e.whole(x2), e.whole(x3), 3, 2);
  e.c3k2(M + "4", e.whole(x3), v4, n2, big, 0.25, 3);                // C3k2(512, False, 0.25)
  e.conv1(M + "5", v4, e.whole(x5), 3, 2);
  e.c3k2(M + "6", e.whole(x5), v6, n2, true, 0.5, 4);                // C3k2(512, True)
  e.conv1(M + "7", v6, e.whole(x7), 3, 2);
  e.c3k2(M + "8", e.whole(x7), e.whole(x8), n2, true, 0.5, 5);       // C3k2(1024, True)
  e.sppf(M + "9", e.whole(x8), e.whole(x9), 5);
  e.c2psa(M + "10", e.whole(x9), v10, n2, 5);
  e.upsample(v10, e.slice(cat12, 0, c1024));
  e.c3k2(M + "13", e.whole(cat12), v13, n2, big, 0.5, 4);            // C3k2(512, False)
  e.upsample(v13, e.slice(cat15, 0, c512));
  e.c3k2(M + "16", e.whole(cat15), e.whole(x16), n2, big, 0.5, 3);   // C3k2(256, False) -> P3
  e.conv1(M + "17", e.whole(x16), v17, 3, 2);
  e.c3k2(M + "19", e.whole(cat18), e.whole(x19), n2, big, 0.5, 4);   // C3k2(512, False) -> P4
  e.conv1(M + "20", e.whole(x19), v20, 3, 2);
  e.c3k2(M + "22", e.whole(cat21), e.whole(x22), n2, true, 0.5, 5);  // C3k2(1024, True) -> P5
  e.name_view("model.0", e.whole(x0)); e.name_view("model.1", e.whole(x1)); e.name_view("model.2", e.whole(x2));
  e.name_view("model.3", e.whole(x3)); e.name_view("model.4", v4); e.name_view("model.5", e.whole(x5));
  e.name_view("model.6", v6); e.name_view("model.7", e.whole(x7)); e.name_view("model.8", e.whole(x8));
  e.name_view("model.9", e.whole(x9)); e.name_view("model.10", v10); e.name_view("model.12", e.whole(cat12));
  e.name_view("model.13", v13); e.name_view("model.15", e.whole(cat15)); e.name_view("model.16", e.whole(x16));
  e.name_view("model.17", v17); e.name_view("model.18", e.whole(cat18)); e.name_view("model.19", e.whole(x19));
  e.name_view("model.20", v20); e.name_view("model.21", e.whole(cat21)); e.name_view("model.22", e.whole(x22));

  // Segment head (module 23): cv2 / cv4 as in yolov8-seg, cv3 = the depthwise class branch of UPSTREAM Detect(legacy=False)
  const std::string Hd = "model.23.";
  const int chs[3] = {c256, c512, c1024};
  const View P[3] = {e.whole(x16), e.whole(x19), e.whole(x22)};
  const int hc2 = std::max(std::max(16, chs[0] / 4), 64), hc3 = std::max(chs[0], std::min(e.nc, 100));
  const int hc4 = std::max(chs[0] / 4, e.nm);
  // Proto first (see build_v8seg)
  e.add_weight(Hd + "dfl.conv.weight", {1, 16, 1, 1}, false);
  // Proto (UPSTREAM block.py::Proto)
  const int npr = ch(256);
  const int p1 = e.new_buf(Hd + "proto.cv1", 3, npr), p2 = e.new_buf(Hd + "proto.upsample", 2, npr);
  const int p3 = e.new_buf(Hd + "proto.cv2", 2, npr), pr = e.new_buf("proto", 2, e.nm, 1);
  e.conv1(Hd + "proto.cv1", P[0], e.whole(p1), 3, 1);
  {
    Op op;
    op.kind = OP_CONV; op.name = Hd + "proto.upsample";
    op.cin = npr; op.cout = 4 * npr; op.k = 1; op.s = 1; op.act = 0; op.out_mode = OUT_SHUFFLE2_BF16;
    op.in = e.whole(p1); op.out = e.whole(p2);
    ConvSrc src{Hd + "proto.upsample", SRC_CONVT, npr};
    e.add_conv_weights(src, npr, 2);
    op.srcs.push_back(src);
    e.ops.push_back(op);
  }
  e.conv1(Hd + "proto.cv2", e.whole(p2), e.whole(p3), 3, 1);
  {
    e.conv1(Hd + "proto.cv3", e.whole(p3), e.whole(pr), 1, 1);
    e.ops.back().out_mode = OUT_F32;
  }
  for (int i = 0; i < 3; ++i) {
    e.feat[i] = P[i];
    const std::string si = std::to_string(i);
    const int lvl = 3 + i, x = chs[i];
    // the first 3x3 conv of the box and coefficient branches reads the same P_i: one fused GEMM, N = hc2 + hc4
    const int f0 = e.new_buf(Hd + "lvl" + si + ".s0", lvl, hc2 + hc4);
    e.conv({Hd + "cv2." + si + ".0", Hd + "cv4." + si + ".0"}, {hc2, hc4}, P[i], e.whole(f0), 3, 1);
    const int t2 = e.new_buf(Hd + "cv2." + si + ".t", lvl, hc2), t4 = e.new_buf(Hd + "cv4." + si + ".t", lvl, hc4);
    e.conv1(Hd + "cv2." + si + ".1", e.slice(f0, 0, hc2), e.whole(t2), 3, 1);
    e.conv1(Hd + "cv4." + si + ".1", e.slice(f0, hc2, hc4), e.whole(t4), 3, 1);
    const std::string b3 = Hd + "cv3." + si + ".";
    const int u0 = e.new_buf(b3 + "u0", lvl, x), u1 = e.new_buf(b3 + "u1", lvl, hc3), u2 = e.new_buf(b3 + "u2", lvl, hc3);
    const int u3 = e.new_buf(b3 + "u3", lvl, hc3);
    e.dwconv({b3 + "0.0"}, {3}, P[i], e.whole(u0), 3, 1, true);
    e.conv1(b3 + "0.1", e.whole(u0), e.whole(u1), 1, 1);
    e.dwconv({b3 + "1.0"}, {3}, e.whole(u1), e.whole(u2), 3, 1, true);
    e.conv1(b3 + "1.1", e.whole(u2), e.whole(u3), 1, 1);
    e.head_out(Hd + "cv2." + si + ".2", e.whole(t2), 64, i, 0);
    e.head_out(b3 + "2", e.whole(u3), e.nc, i, 64);
    e.head_out(Hd + "cv4." + si + ".2", e.whole(t4), e.nm, i, 64 + e.nc);
  }
  e.proto_buf = pr;
  e.name_view("proto", e.whole(pr));
  return true;
}

// ------------------------------------------------------------------------------------------------
// weights: BN fold + repack
// ------------------------------------------------------------------------------------------------
static uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// Folded fp32 weight [cout][cin][k][k] (ConvT: as stored [cin][cout][2][2]) and bias [cout] of one source module.
static int folded(const ypb_engine& e, const ConvSrc& s, std::vector<float>* w, std::vector<float>* b) {
  auto get = [&](const std::string& n) -> const WeightEntry* {
    auto it = e.windex.find(n);
    return it == e.windex.end() ? nullptr : &e.weights[it->second];
  };
  if (s.kind == SRC_CONV_BN) {
    const WeightEntry *cw = get(s.mod + ".conv.weight"), *g = get(s.mod + ".bn.weight"), *be = get(s.mod + ".bn.bias"),
                      *mu = get(s.mod + ".bn.running_mean"), *var = get(s.mod + ".bn.running_var");
    if (!cw || !g || !be || !mu || !var || !cw->loaded || !g->loaded || !be->loaded || !mu->loaded || !var->loaded)
      return fail(YPB_ERR_STATE, "weights of " + s.mod + " not loaded");
    *w = cw->data;
    b->assign(s.cout, 0.f);
    const int64_t per = cw->numel() / s.cout;
    for (int co = 0; co < s.cout; ++co) {
      const float sc = g->data[co] / std::sqrt(var->data[co] + 1e-3f);
      for (int64_t i = 0; i < per; ++i) (*w)[co * per + i] *= sc;
      (*b)[co] = be->data[co] - mu->data[co] * sc;
    }
  } else {
    const WeightEntry *cw = get(s.mod + ".weight"), *cb = get(s.mod + ".bias");
    if (!cw || !cb || !cw->loaded || !cb->loaded) return fail(YPB_ERR_STATE, "weights of " + s.mod + " not loaded");
    *w = cw->data;
    *b = cb->data;
  }
  return YPB_OK;
}


// Enqueue one layer op on `st`.
static int launch_op(ypb_engine* e, const Op& op, cudaStream_t st, const uint8_t* frames, int split = 0) {
  const uint8_t* wa = reinterpret_cast<const uint8_t*>(e->w_arena);
  const int B = e->sB[split], b0 = e->sb0[split];  // this launch covers images [b0, b0 + B)
  if (B <= 0) return YPB_OK;
  // first byte of image b0 in an activation buffer (every kernel below indexes images from its base pointer)
  auto base = [&](const BufDesc& b) { return e->ws + b.offset + (size_t)b0 * b.H * b.W * b.C * (b.dtype ? 4 : 2); };
  frames += (size_t)b0 * e->H * e->W * 3;
  switch (op.kind) {
    case OP_STEM: {
      const BufDesc& ob = e->bufs[op.out.buf];
      ConvParams sp;
      memset(&sp, 0, sizeof sp);
      sp.Cout = op.cout; sp.n_tile = op.cout; sp.act = 1; sp.out_mode = OUT_BF16;
      sp.img_HW = ob.H * ob.W; sp.img_W = ob.W;
      sp.out = base(ob); sp.out_img_stride = (long long)ob.H * ob.W * ob.C; sp.out_pix_stride = ob.C;
      sp.out_c_off = op.out.c_off; sp.bias = reinterpret_cast<const float*>(wa + op.b_off);
      const int tiles_w = (ob.W + kStemTW - 1) / kStemTW, tiles_h = (ob.H + kStemTH - 1) / kStemTH;
      sp.tiles_w = tiles_w; sp.tiles_h = tiles_h;
      conv_set_fastdiv(sp, 1);
      const int total = B * tiles_h * tiles_w, per_cta = 8;
      // the epilogue staging tile (32 rows x (min(2*C0,128)+16) B per warp) aliases the A tile when it fits in 4 KB per warp
      const int alias = 1;  // a pass stages 32 rows x (64 + 16) B per warp
      if (op.cout > kStemMaxC0) return fail(YPB_ERR_ARG, "stem: too many output channels");
      const size_t smem = 1024 + 16384 + ((op.cout * 128 + 1023) & ~1023) + 17 * kStemRowWords * 4 + 16 +
                          (alias ? 0 : 4 * kEpiStageBytes);
      stem_tc_kernel<<<(total + per_cta - 1) / per_cta, 128, smem, st>>>(frames, e->H, e->W, B,
                                                                        reinterpret_cast<const __nv_bfloat16*>(wa + op.w_off), sp,
                                                                        tiles_h * tiles_w, tiles_w, total, per_cta, alias);
      break;
    }
    case OP_CONV:
      CUDA_TRY(conv_launch(split ? op.L1 : op.L, st, e->conv_impl));
      break;
    case OP_UPSAMPLE: {
      const BufDesc &ib = e->bufs[op.in.buf], &ob = e->bufs[op.out.buf];
      const int row_vecs = ib.W * (op.in.C / 8);
      if ((long long)B * ib.H > 65535) return fail(YPB_ERR_ARG, "upsample: batch x rows exceeds the grid limit");
      upsample2x_kernel<<<dim3((unsigned)((row_vecs + 255) / 256), (unsigned)(B * ib.H)), 256, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(base(ib)), ib.C, op.in.c_off,
          reinterpret_cast<__nv_bfloat16*>(base(ob)), ob.C, op.out.c_off, B, ib.H, ib.W, op.in.C);
      break;
    }
    case OP_SPPF: {
      const BufDesc& ib = e->bufs[op.in.buf];
      const size_t smem = (size_t)2 * ib.H * ib.W * 16;
      if (smem > 200 * 1024) return fail(YPB_ERR_ARG, "sppf: feature map too large for the shared-memory pool kernel");
      if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        DeviceState& ds = device_state();
        if (smem > ds.sppf_smem) {
          CUDA_TRY(cudaFuncSetAttribute(sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          ds.sppf_smem = 200 * 1024;
        }
      }
      sppf_pool_kernel<<<dim3(op.in.C / 8, B), 256, smem, st>>>(reinterpret_cast<__nv_bfloat16*>(base(ib)), B, ib.H,
                                                                ib.W, op.in.C);
      break;
    }
    case OP_DW: {
      const BufDesc &ib = e->bufs[op.in.buf], &ob = e->bufs[op.out.buf];
      DwParams p{};
      p.in = reinterpret_cast<const __nv_bfloat16*>(base(ib));
      p.H = ib.H; p.W = ib.W; p.in_ctot = ib.C; p.in_c_off = op.in.c_off;
      p.wk = reinterpret_cast<const __nv_bfloat16*>(wa + op.w_off);
      p.bias = reinterpret_cast<const float*>(wa + op.b_off);
      p.C = op.cout; p.k = op.k; p.stride = op.s; p.act = op.act;
      p.out = reinterpret_cast<__nv_bfloat16*>(base(ob));
      p.oH = ob.H; p.oW = ob.W; p.out_ctot = ob.C; p.out_c_off = op.out.c_off;
      if (op.res.buf >= 0) {
        p.res = reinterpret_cast<const __nv_bfloat16*>(base(e->bufs[op.res.buf]));
        p.res_ctot = e->bufs[op.res.buf].C; p.res_c_off = op.res.c_off;
      }
      p.nB = B;
      if (op.heads > 0) {  // Attention.pe: one launch per head over that head's v channels of the qkv buffer
        const int hd = kAttnHD, hc = 2 * kAttnKD + kAttnHD;
        for (int h = 0; h < op.heads; ++h) {
          DwParams q = p;
          q.C = hd; q.in_c_off = h * hc + 2 * kAttnKD; q.out_c_off = h * hd;
          q.wk = p.wk + h * hd; q.bias = p.bias + h * hd;
          // weights are [tap][C_total]: give the kernel the full row stride through a strided view
          q.in = p.in; q.out = p.out;
          const long long total = (long long)B * ob.H * ob.W * (hd / 8);
          dwconv_strided_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(q, op.cout);
        }
      } else {
        const long long total = (long long)B * ob.H * ob.W * (op.cout / 8);
        dwconv_strided_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(p, op.cout);
      }
      break;
    }
    case OP_ATTN: {
      const BufDesc &qb = e->bufs[op.in.buf], &pb = e->bufs[op.res.buf], &ob = e->bufs[op.out.buf];
      const int N = qb.H * qb.W;
      dim3 grid((N + kAttnThreads - 1) / kAttnThreads, op.heads, B);
      psa_attention_kernel<<<grid, kAttnThreads, 0, st>>>(
          reinterpret_cast<const __nv_bfloat16*>(base(qb)), qb.C, op.in.c_off,
          reinterpret_cast<const __nv_bfloat16*>(base(pb)), pb.C, op.res.c_off,
          reinterpret_cast<__nv_bfloat16*>(base(ob)), ob.C, op.out.c_off, N, 1.0f / std::sqrt((float)kAttnKD));
      break;
    }
  }
  return YPB_OK;
}

// Selection stage: candidate filter + decode, then batched NMS.
static int launch_select(ypb_engine* e, cudaStream_t st, const float* xform, const ypb_infer_params* prm, float* det,
                         float* det_lb, int32_t* keep, float* coef, int32_t* count, cudaEvent_t mid, int b0 = 0, int nb = -1) {
  // images [b0, b0 + nb) of the batch (default: all of it); every per-image array is advanced to image b0
  const int B = nb < 0 ? e->B : nb;
  if (B <= 0) return YPB_OK;
  const HeadGeom& g = e->hg;
  int* cand_count = reinterpret_cast<int*>(e->ws + e->off_count) + b0;
  CUDA_TRY(cudaMemsetAsync(cand_count, 0, (size_t)B * 4, st));
  const float* head = reinterpret_cast<const float*>(e->ws + e->off_head) + (size_t)b0 * g.A * g.no;
  float4* dbox = reinterpret_cast<float4*>(e->ws + e->off_dbox) + (size_t)b0 * g.A;
  int* dcls = reinterpret_cast<int*>(e->ws + e->off_dcls) + (size_t)b0 * g.A;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(e->ws + e->off_keys) + (size_t)b0 * g.cand_stride;
  xform += (size_t)b0 * 5;
  det += (size_t)b0 * kNmsMaxDet * 6;
  det_lb += (size_t)b0 * kNmsMaxDet * 4;
  keep += (size_t)b0 * kNmsMaxDet;
  if (coef) coef += (size_t)b0 * kNmsMaxDet * g.nm;
  count += b0;
  const long long warps = (long long)B * g.A;
  bool pairs_done = false;
  if (g.nc % 4 == 0 && g.no % 4 == 0 && g.nc >= 16) {
    const long long w8 = (warps + 7) / 8;
    decode_filter8_kernel<<<(unsigned)((w8 * 32 + 255) / 256), 256, 0, st>>>(head, g, B, prm->conf, e->end2end ? 1 : 0,
                                                                            prm->class_mask, dbox, dcls, keys, cand_count);
    pairs_done = true;  // end-to-end heads: the (anchor, class) pairs were pushed in the same pass
  } else {
    decode_filter_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(head, g, B, prm->conf, e->end2end ? 1 : 0,
                                                                              prm->class_mask, dbox, dcls, keys, cand_count);
  }
  if (e->end2end && !pairs_done) {  // NMS-free head: candidates are (anchor, class) pairs; rebuild the key list from scratch
    CUDA_TRY(cudaMemsetAsync(cand_count, 0, (size_t)B * 4, st));
    pair_candidates_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(head, g, B, prm->conf, prm->class_mask, keys,
                                                                              cand_count);
  }
  if (mid) CUDA_TRY(cudaEventRecord(mid, st));
  nms_kernel<<<B, kNmsThreads, 0, st>>>(head, g, dbox, dcls, keys, cand_count, prm->iou, prm->max_det, 30000,
                                prm->agnostic_nms ? 0.0f : 7680.0f, e->end2end ? 1 : 0, e->end2end ? prm->class_mask : nullptr,
                                kNmsMaxDet, reinterpret_cast<const FrameXform*>(xform), det, det_lb, keep, coef, count);
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

static int check_infer_args(ypb_engine* e, const uint8_t* frames, const float* xform, const ypb_infer_params* prm,
                            float* det, float* det_lb, int32_t* keep, float* coef, int32_t* count) {
  if (!e || !frames || !xform || !prm || !det || !det_lb || !keep || !count) return fail(YPB_ERR_ARG, "bad argument");
  if (!e->bound) return fail(YPB_ERR_STATE, "infer: bind a workspace first");
  if (prm->max_det < 1 || prm->max_det > kNmsMaxDet) return fail(YPB_ERR_ARG, "max_det must be in [1,300]");
  if (e->nm > 0 && !coef) return fail(YPB_ERR_ARG, "coef buffer required for -seg models");
  return YPB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int ypb_version(void) { return 100; }
int ypb_is_diag_build(void) { return YPB_DIAG ? 1 : 0; }
const char* ypb_last_error(void) { return g_last_error.c_str(); }

int ypb_engine_create(const char* model_spec, int nc, ypb_engine** out) {
  if (!model_spec || !out || nc < 1) return fail(YPB_ERR_ARG, "bad argument");
  ypb_engine* e = new ypb_engine();
  e->spec = model_spec;
  e->nc = nc;
  bool ok = false;
  const std::string s = model_spec;
  if (s.size() == 11 && s.rfind("yolov8", 0) == 0 && s.substr(7) == "-seg") ok = build_v8seg(*e, s[6]);
  if (s == "yolov10n") ok = build_v10n(*e);
  if (s.size() == 11 && s.rfind("yolo11", 0) == 0 && s.substr(7) == "-seg") ok = build_v11seg(*e, s[6]);
  if (!ok) {
    delete e;
    return fail(YPB_ERR_ARG, "unknown model spec '" + s + "'");
  }
  e->branch_parallel = getenv("YPB_NO_BRANCHES") == nullptr;
  e->build_schedule(getenv("YPB_STREAMS") ? std::max(1, std::min(8, atoi(getenv("YPB_STREAMS")))) : 8);
  *out = e;
  return YPB_OK;
}

void ypb_engine_destroy(ypb_engine* e) {
  if (!e) return;
  {
    DeviceGuard guard(e->device);
    e->drop_graph();
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    for (cudaStream_t ss : e->side_streams) if (ss) cudaStreamDestroy(ss);
    for (cudaEvent_t ev : e->op_events) cudaEventDestroy(ev);
    if (e->w_arena) cudaFree(e->w_arena);
  }
  delete e;
}

int ypb_weight_count(const ypb_engine* e) { return e ? (int)e->weights.size() : 0; }

int ypb_weight_info(const ypb_engine* e, int i, const char** name, int* ndim, int64_t shape[4], int* used) {
  if (!e || i < 0 || i >= (int)e->weights.size()) return fail(YPB_ERR_ARG, "bad weight index");
  const WeightEntry& w = e->weights[i];
  if (name) *name = w.name.c_str();
  if (ndim) *ndim = (int)w.shape.size();
  if (shape) for (size_t d = 0; d < 4; ++d) shape[d] = d < w.shape.size() ? w.shape[d] : 1;
  if (used) *used = w.used ? 1 : 0;
  return YPB_OK;
}

int ypb_load_weight(ypb_engine* e, const char* name, const float* data, int64_t numel) {
  if (!e || !name || !data) return fail(YPB_ERR_ARG, "bad argument");
  auto it = e->windex.find(name);
  if (it == e->windex.end()) return fail(YPB_ERR_ARG, std::string("unknown weight '") + name + "'");
  WeightEntry& w = e->weights[it->second];
  if (numel != w.numel()) return fail(YPB_ERR_ARG, std::string("size mismatch for '") + name + "'");
  if (w.used) w.data.assign(data, data + numel);
  w.loaded = true;
  // the device copy is stale from here on: ypb_infer() must fail until finalize + bind have been redone
  e->finalized = false;
  e->bound = false;
  e->drop_graph();
  return YPB_OK;
}

int ypb_finalize_weights(ypb_engine* e, int device) {
  if (!e) return fail(YPB_ERR_ARG, "null engine");
  // layout pass
  size_t off = 0;
  auto align = [](size_t x) { return (x + 1023) & ~size_t(1023); };
  for (Op& op : e->ops) {
    if (op.kind == OP_STEM) {
      op.w_off = off; off = align(off + (size_t)64 * op.cout * 2);
      op.b_off = off; off = align(off + (size_t)op.cout * 4);
    } else if (op.kind == OP_CONV) {
      op.w_off = off; off = align(off + (size_t)op.k * op.k * op.cout * op.cin * 2);
      op.b_off = off; off = align(off + (size_t)op.cout * 4);
    } else if (op.kind == OP_DW) {
      op.w_off = off; off = align(off + (size_t)op.k * op.k * op.cout * 2);
      op.b_off = off; off = align(off + (size_t)op.cout * 4);
    }
  }
  std::vector<uint8_t> host(off, 0);
  for (Op& op : e->ops) {
    if (op.kind == OP_DW) {  // depthwise: [tap][C] bf16; RepVGGDW sums the centred 3x3 into the 7x7 in fp32 first
      const int C = op.cout, kk = op.k * op.k;
      std::vector<float> wsum((size_t)kk * C, 0.f);
      float* bias = reinterpret_cast<float*>(host.data() + op.b_off);
      for (const ConvSrc& s : op.srcs) {
        std::vector<float> w, b;
        int rc = folded(*e, s, &w, &b);
        if (rc) return rc;
        const int ks = (int)std::lround(std::sqrt((double)(w.size() / C)));
        const int o = (op.k - ks) / 2;
        for (int c = 0; c < C; ++c) {
          for (int kh = 0; kh < ks; ++kh)
            for (int kw = 0; kw < ks; ++kw) wsum[(size_t)((kh + o) * op.k + kw + o) * C + c] += w[((size_t)c * ks + kh) * ks + kw];
          bias[c] += b[c];
        }
      }
      uint16_t* wk = reinterpret_cast<uint16_t*>(host.data() + op.w_off);
      for (size_t i = 0; i < wsum.size(); ++i) wk[i] = f32_to_bf16(wsum[i]);
      continue;
    }
    if (op.kind != OP_STEM && op.kind != OP_CONV) continue;
    float* bias = reinterpret_cast<float*>(host.data() + op.b_off);
    int n_off = 0;
    for (const ConvSrc& s : op.srcs) {
      std::vector<float> w, b;
      int rc = folded(*e, s, &w, &b);
      if (rc) return rc;
      if (op.kind == OP_STEM) {
        // K-major rows [C0][64] bf16 = [hi(32) | lo(32)], k = (kh*3+kw)*3 + c_rgb: the /255 of preprocess is folded into
        // the weights (the kernel feeds raw pixel values, exact in bf16) and w/255 is split into two bf16 terms
        uint16_t* wq = reinterpret_cast<uint16_t*>(host.data() + op.w_off);
        for (int co = 0; co < s.cout; ++co)
          for (int c = 0; c < 3; ++c)
            for (int kh = 0; kh < 3; ++kh)
              for (int kw = 0; kw < 3; ++kw) {
                const float v = w[((co * 3 + c) * 3 + kh) * 3 + kw] / 255.0f;
                const uint16_t hi = f32_to_bf16(v);
                uint32_t hb = (uint32_t)hi << 16;
                float hf;
                memcpy(&hf, &hb, 4);
                wq[co * 64 + (kh * 3 + kw) * 3 + c] = hi;
                wq[co * 64 + 32 + (kh * 3 + kw) * 3 + c] = f32_to_bf16(v - hf);
              }
        for (int co = 0; co < s.cout; ++co) bias[co] = b[co];
      } else if (s.kind == SRC_CONVT) {
        uint16_t* wg = reinterpret_cast<uint16_t*>(host.data() + op.w_off);
        const int cq = s.cout;
        for (int ci = 0; ci < op.cin; ++ci)
          for (int co = 0; co < cq; ++co)
            for (int g = 0; g < 4; ++g)
              wg[((size_t)(g * cq + co)) * op.cin + ci] = f32_to_bf16(w[((size_t)(ci * cq + co)) * 4 + g]);
        for (int g = 0; g < 4; ++g)
          for (int co = 0; co < cq; ++co) bias[g * cq + co] = b[co];
      } else {
        uint16_t* wg = reinterpret_cast<uint16_t*>(host.data() + op.w_off);
        const int kk = op.k * op.k;
        const int cr = op.cin_real > 0 ? op.cin_real : op.cin;  // padded input / output channels keep zero weights
        for (int co = 0; co < s.cout; ++co)
          for (int ci = 0; ci < cr; ++ci)
            for (int t = 0; t < kk; ++t)
              wg[((size_t)t * op.cout + n_off + co) * op.cin + ci] = f32_to_bf16(w[((size_t)co * cr + ci) * kk + t]);
        for (int co = 0; co < s.cout; ++co) bias[n_off + co] = b[co];
      }
      n_off += s.width();
    }
  }
  if (e->device >= 0 && e->device != device) {
    // the engine moves to another GPU (YOLO.to('cuda:1')): graphs, capture streams, events and the weight arena belong
    // to the old device
    DeviceGuard old_dev(e->device);
    e->drop_graph();
    if (e->cap_stream) { cudaStreamDestroy(e->cap_stream); e->cap_stream = nullptr; }
    for (cudaStream_t& ss : e->side_streams) if (ss) { cudaStreamDestroy(ss); ss = nullptr; }
    for (cudaEvent_t ev : e->op_events) cudaEventDestroy(ev);
    e->op_events.clear();
    if (e->w_arena) { cudaFree(e->w_arena); e->w_arena = nullptr; }
  }
  DeviceGuard guard(device);  // the caller's current device is restored on return
  {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return fail(YPB_ERR_CUDA, "cannot select CUDA device " + std::to_string(device));
  }
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(YPB_ERR_CUDA, "device is not sm_100 (Blackwell B200); there is no fallback path");
  if (e->w_arena) { cudaFree(e->w_arena); e->w_arena = nullptr; }
  CUDA_TRY(cudaMalloc(&e->w_arena, off));
  CUDA_TRY(cudaMemcpy(e->w_arena, host.data(), off, cudaMemcpyHostToDevice));
  e->w_bytes = off;
  e->device = device;
  e->finalized = true;
  e->bound = false;
  return YPB_OK;
}

int ypb_plan(ypb_engine* e, int B, int H, int W, size_t* workspace_bytes) {
  if (!e || B < 1 || H < 32 || W < 32 || (H % 32) || (W % 32)) return fail(YPB_ERR_ARG, "plan: H and W must be multiples of 32");
  e->B = B; e->H = H; e->W = W;
  size_t off = 0;
  auto align = [](size_t x) { return (x + 1023) & ~size_t(1023); };
  for (BufDesc& b : e->bufs) {
    b.H = H >> b.lvl; b.W = W >> b.lvl;
    b.bytes = (size_t)B * b.H * b.W * b.C * (b.dtype ? 4 : 2);
    b.offset = off;
    off = align(off + b.bytes);
  }
  // head geometry
  HeadGeom& g = e->hg;
  int A = 0;
  for (int i = 0; i < 3; ++i) {
    const BufDesc& fb = e->bufs[e->feat[i].buf];
    g.lvl_start[i] = A; g.lvl_w[i] = fb.W; g.lvl_stride[i] = (float)(1 << fb.lvl);
    A += fb.H * fb.W;
  }
  g.lvl_start[3] = A;
  g.A = A; g.nc = e->nc; g.nm = e->nm; g.no = 64 + e->nc + e->nm;
  int cs = 1;
  while (cs < A) cs <<= 1;
  g.cand_stride = cs;
  e->A = A; e->no = g.no;
  e->off_head = off; off = align(off + (size_t)B * A * g.no * 4);
  e->off_dbox = off; off = align(off + (size_t)B * A * 16);
  e->off_dcls = off; off = align(off + (size_t)B * A * 4);
  e->off_keys = off; off = align(off + (size_t)B * cs * 8);
  e->off_count = off; off = align(off + (size_t)B * 4);
  e->off_moff = off; off = align(off + (size_t)(B + 1) * 4);
  e->ws_bytes = off;
  // batch halves (see ypb_engine::n_split)
  {
    // Measured on B200 (profiles/r2_ab_split_branches.md): with one persistent CTA per SM the second half's kernels
    // cannot co-run, only tail-overlap, and every half-sized launch loses more to wave quantisation than the overlap
    // recovers (yolov8s-seg B=64: 14 725 vs 15 061 frames/s) - so the split is opt-in (YPB_BATCH_SPLIT=2).
    int want = 1;
    if (const char* ev = getenv("YPB_BATCH_SPLIT")) want = atoi(ev);
    e->n_split = (want >= 2 && B >= 4) ? 2 : 1;
    e->sB[0] = e->n_split == 2 ? (B + 1) / 2 : B;
    e->sB[1] = B - e->sB[0];
    e->sb0[0] = 0; e->sb0[1] = e->sB[0];
  }
  // per-op geometry
  e->launches = 0; e->flops = 0;
  for (Op& op : e->ops) {
    e->launches += e->n_split;
    if (op.kind == OP_STEM) { e->flops += 2.0 * B * (H / 2) * (W / 2) * op.cout * 27; continue; }
    if (op.kind == OP_DW) {
      const BufDesc& ob = e->bufs[op.out.buf];
      e->flops += 2.0 * B * ob.H * ob.W * op.cout * op.k * op.k;
      if (op.heads > 1) e->launches += e->n_split * (op.heads - 1);
      continue;
    }
    if (op.kind != OP_CONV) continue;
    const BufDesc& ib = e->bufs[op.in.buf];
    ConvDesc d;
    d.B = B; d.Hin = ib.H; d.Win = ib.W; d.in_ctot = ib.C; d.in_c_off = op.in.c_off; d.cin = op.cin;
    d.cout = op.cout; d.k = op.k; d.stride = op.s; d.act = op.act; d.out_mode = op.out_mode;
    if (op.head_lvl >= 0) {
      d.out_img_stride = (long long)A * g.no; d.out_pix_stride = g.no;
      d.out_c_off = g.lvl_start[op.head_lvl] * g.no + op.head_coff;
    } else {
      const BufDesc& ob = e->bufs[op.out.buf];
      d.out_img_stride = (long long)ob.H * ob.W * ob.C; d.out_pix_stride = ob.C; d.out_c_off = op.out.c_off;
    }
    if (op.res.buf >= 0) {
      const BufDesc& rb = e->bufs[op.res.buf];
      d.res_img_stride = (long long)rb.H * rb.W * rb.C; d.res_pix_stride = rb.C; d.res_c_off = op.res.c_off;
    }
    std::string err;
    d.B = e->sB[0];
    if (!conv_plan_geometry(d, &op.L, &err)) return fail(YPB_ERR_ARG, op.name + ": " + err);
    if (e->n_split == 2) {
      d.B = e->sB[1];
      if (!conv_plan_geometry(d, &op.L1, &err)) return fail(YPB_ERR_ARG, op.name + ": " + err);
    }
    {  // algorithmic FLOPs: the module's real channels, not the zero-padded ones the tensor core multiplies
      int cout_real = 0;
      for (const ConvSrc& sc : op.srcs) cout_real += sc.kind == SRC_CONVT ? 4 * sc.cout : sc.cout;
      const int cin_r = op.cin_real > 0 ? op.cin_real : op.cin;
      op.flops = 2.0 * B * op.L.oH * op.L.oW * (double)cout_real * cin_r * op.k * op.k;
      op.L.flops = op.flops;
    }
    e->flops += op.flops;
    if (getenv("YPB_PLAN_DEBUG")) {
      if (op.L.use_halo)
        fprintf(stderr, "[plan] %-28s k%d s%d %4d->%4d %4dx%-4d halo msub %d %s a_slots %d b_slots %d tiles %d smem %d\n",
                op.name.c_str(), op.k, op.s, op.cin, op.cout, op.L.oH, op.L.oW, op.L.x3.msub,
                op.L.x3.b_stat ? "W-resident" : "W-stream", op.L.x3.a_slots, op.L.x3.b_slots, op.L.total_tiles3, op.L.smem3);
      else
        fprintf(stderr, "[plan] %-28s k%d s%d %4d->%4d %4dx%-4d taps tile %dx%d n_tile %d stages %d tiles %d smem %d\n",
                op.name.c_str(), op.k, op.s, op.cin, op.cout, op.L.oH, op.L.oW, op.L.p.TH, op.L.p.TW, op.L.p.n_tile,
                op.L.stages2, op.L.total_tiles, op.L.smem2);
    }
  }
  e->launches += e->n_split * 2;  // decode_filter + nms per batch half
  e->planned = true;
  e->bound = false;
  if (workspace_bytes) *workspace_bytes = off;
  return YPB_OK;
}

int ypb_bind_workspace(ypb_engine* e, void* workspace, size_t bytes) {
  if (!e || !workspace) return fail(YPB_ERR_ARG, "bad argument");
  if (!e->planned || !e->finalized) return fail(YPB_ERR_STATE, "bind: finalize weights and plan first");
  if (bytes < e->ws_bytes) return fail(YPB_ERR_ARG, "workspace too small");
  if ((uintptr_t)workspace & 1023) return fail(YPB_ERR_ARG, "workspace must be 1024-byte aligned");
  DeviceGuard guard(e->device);
  e->ws = reinterpret_cast<uint8_t*>(workspace);
  const uint8_t* wa = reinterpret_cast<const uint8_t*>(e->w_arena);
  const HeadGeom& g = e->hg;
  for (Op& op : e->ops) {
    if (op.kind != OP_CONV) continue;
    const BufDesc& ib = e->bufs[op.in.buf];
    for (int sp = 0; sp < e->n_split; ++sp) {
      const size_t b0 = (size_t)e->sb0[sp];
      auto img0 = [&](const BufDesc& b) { return e->ws + b.offset + b0 * b.H * b.W * b.C * (b.dtype ? 4 : 2); };
      ConvDesc d;
      d.in = img0(ib);
      d.B = e->sB[sp]; d.Hin = ib.H; d.Win = ib.W; d.in_ctot = ib.C; d.in_c_off = op.in.c_off; d.cin = op.cin;
      d.wg = wa + op.w_off; d.bias = reinterpret_cast<const float*>(wa + op.b_off);
      d.cout = op.cout; d.k = op.k; d.stride = op.s; d.act = op.act; d.out_mode = op.out_mode;
      d.out = op.head_lvl >= 0 ? (void*)(e->ws + e->off_head + b0 * g.A * g.no * 4) : (void*)img0(e->bufs[op.out.buf]);
      if (op.res.buf >= 0) d.res = img0(e->bufs[op.res.buf]);
      std::string err;
      if (!conv_bind(d, sp ? &op.L1 : &op.L, &err)) return fail(YPB_ERR_CUDA, op.name + ": " + err);
    }
  }
  (void)g;
  e->drop_graph();
  {
    cudaError_t ce = conv_launch_init();
    if (ce != cudaSuccess) return fail(YPB_ERR_CUDA, std::string("conv_launch_init: ") + cudaGetErrorString(ce));
  }
  e->bound = true;
  return YPB_OK;
}

int ypb_num_anchors(const ypb_engine* e) { return e ? e->A : 0; }
int ypb_num_classes(const ypb_engine* e) { return e ? e->nc : 0; }
int ypb_num_mask_coefs(const ypb_engine* e) { return e ? e->nm : 0; }
int ypb_kernel_launches(const ypb_engine* e) { return e ? e->launches : 0; }
double ypb_conv_flops(const ypb_engine* e) { return e ? e->flops : 0.0; }

int ypb_set_conv_impl(ypb_engine* e, int impl) {
  if (!e || impl < 0 || impl > 3) return fail(YPB_ERR_ARG, "bad argument");
#if !YPB_DIAG
  if (impl != 0) return fail(YPB_ERR_ARG, "conv impl 1-3 are debugging twins: load libypb200_diag.so (include/ypb200_diag.h)");
#endif
  e->conv_impl = impl;
  e->drop_graph();
  return YPB_OK;
}

int ypb_set_graph(ypb_engine* e, int on) {
  if (!e) return fail(YPB_ERR_ARG, "bad argument");
  e->use_graph = on != 0;
  e->drop_graph();
  return YPB_OK;
}

// Channel ranges of buffers an op reads / writes (head rows: pseudo buffer 1000 + level).
namespace {
struct Access { int buf, lo, hi; };
inline bool overlaps(const Access& a, const Access& b) { return a.buf == b.buf && a.lo < b.hi && b.lo < a.hi; }
void op_accesses(const Op& op, std::vector<Access>* rd, std::vector<Access>* wr) {
  auto add = [](std::vector<Access>* v, const View& w) { if (w.buf >= 0) v->push_back({w.buf, w.c_off, w.c_off + w.C}); };
  if (op.kind != OP_STEM) add(rd, op.in);
  add(rd, op.res);
  if (op.head_lvl >= 0) wr->push_back({1000 + op.head_lvl, op.head_coff, op.head_coff + op.cout});
  else add(wr, op.out);
  if (op.kind == OP_SPPF) add(rd, op.out);  // the pools read what earlier pools of the same launch wrote
}
}  // namespace

void ypb_engine::build_schedule(int max_streams) {
  const int n = (int)ops.size();
  sched.assign(n, OpSched());
  std::vector<std::vector<Access>> rd(n), wr(n);
  for (int j = 0; j < n; ++j) op_accesses(ops[j], &rd[j], &wr[j]);
  std::vector<int> tail;  // last op of every stream
  tail.push_back(-1);
  for (int j = 0; j < n; ++j) {
    std::vector<int> preds;
    for (int i = 0; i < j; ++i) {
      bool dep = false;
      for (const Access& w : wr[i]) {
        for (const Access& r : rd[j]) dep = dep || overlaps(w, r);   // read after write
        for (const Access& w2 : wr[j]) dep = dep || overlaps(w, w2);  // write after write
      }
      for (const Access& r : rd[i])
        for (const Access& w2 : wr[j]) dep = dep || overlaps(r, w2);  // write after read (in-place blocks)
      if (dep) preds.push_back(i);
    }
    int stream = -1;
    // continue the stream whose tail is one of my predecessors (the latest such predecessor wins)
    for (int k = (int)preds.size() - 1; k >= 0 && stream < 0; --k)
      if (tail[sched[preds[k]].stream] == preds[k]) stream = sched[preds[k]].stream;
    if (stream < 0) {
      if (preds.empty()) stream = 0;
      else if ((int)tail.size() < max_streams) { tail.push_back(-1); stream = (int)tail.size() - 1; }
      else stream = sched[preds.back()].stream;
    }
    sched[j].stream = stream;
    for (int i : preds)
      if (sched[i].stream != stream) { sched[j].waits.push_back(i); sched[i].record = true; }
    tail[stream] = j;
  }
  n_streams = (int)tail.size();
  if (getenv("YPB_PLAN_DEBUG"))
    for (int j = 0; j < n; ++j) {
      fprintf(stderr, "[sched] %-30s stream %d waits", ops[j].name.c_str(), sched[j].stream);
      for (int w : sched[j].waits) fprintf(stderr, " %s", ops[w].name.c_str());
      fprintf(stderr, "\n");
    }
  // the selection stage runs on stream 0 after every branch: the tails of the side streams are recorded too
  for (int st = 1; st < n_streams; ++st)
    if (tail[st] >= 0) sched[tail[st]].record = true;
}

static int enqueue_infer(ypb_engine* e, cudaStream_t st, const uint8_t* frames, const float* xform,
                         const ypb_infer_params* prm, float* det, float* det_lb, int32_t* keep, float* coef, int32_t* count,
                         bool capturing = false) {
  const int n = (int)e->ops.size();
  const bool par = capturing && e->branch_parallel && (int)e->sched.size() == n && (e->n_streams > 1 || e->n_split > 1);
  if (!par) {  // plain stream order: half 0 then half 1 of every op, then selection over the whole batch
    for (const Op& op : e->ops)
      for (int sp = 0; sp < e->n_split; ++sp) {
        int rc = launch_op(e, op, st, frames, sp);
        if (rc) return rc;
      }
    CUDA_TRY(cudaGetLastError());
    return launch_select(e, st, xform, prm, det, det_lb, keep, coef, count, nullptr);
  }
  // Inside a stream capture: every batch half is its own set of streams (stream 0 of half 0 is the capture's origin);
  // side streams join the capture through event waits and are joined back before the capture ends.  Each half runs
  // its own decode + NMS at the end of its chain, overlapping the other half's tensor work.
  const int ns = e->n_streams, total_streams = ns * e->n_split;
  if ((int)e->op_events.size() < n * e->n_split + 2) {
    const size_t old = e->op_events.size();
    e->op_events.resize(n * e->n_split + 2);
    for (size_t i = old; i < e->op_events.size(); ++i) CUDA_TRY(cudaEventCreateWithFlags(&e->op_events[i], cudaEventDisableTiming));
  }
  for (int k = 1; k < total_streams; ++k)
    if (!e->side_streams[k]) CUDA_TRY(cudaStreamCreateWithFlags(&e->side_streams[k], cudaStreamNonBlocking));
  auto stream_of = [&](int sp, int k) { return (sp == 0 && k == 0) ? st : e->side_streams[sp * ns + k]; };
  cudaEvent_t fork = e->op_events[n * e->n_split];
  if (e->n_split > 1) {
    CUDA_TRY(cudaEventRecord(fork, st));
    CUDA_TRY(cudaStreamWaitEvent(stream_of(1, 0), fork, 0));
  }
  // interleave the halves op by op so that the graph's node order (a scheduling hint) alternates between them
  std::vector<int> last(total_streams, -1);
  for (int j = 0; j < n; ++j) {
    const ypb_engine::OpSched& sc = e->sched[j];
    for (int sp = 0; sp < e->n_split; ++sp) {
      cudaStream_t sj = stream_of(sp, sc.stream);
      for (int w : sc.waits) CUDA_TRY(cudaStreamWaitEvent(sj, e->op_events[sp * n + w], 0));
      int rc = launch_op(e, e->ops[j], sj, frames, sp);
      if (rc) return rc;
      if (sc.record) CUDA_TRY(cudaEventRecord(e->op_events[sp * n + j], sj));
      last[sp * ns + sc.stream] = j;
    }
  }
  CUDA_TRY(cudaGetLastError());
  for (int sp = 0; sp < e->n_split; ++sp) {
    cudaStream_t s0 = stream_of(sp, 0);
    for (int k = 1; k < ns; ++k)
      if (last[sp * ns + k] >= 0) CUDA_TRY(cudaStreamWaitEvent(s0, e->op_events[sp * n + last[sp * ns + k]], 0));
    int rc = launch_select(e, s0, xform, prm, det, det_lb, keep, coef, count, nullptr, e->sb0[sp], e->sB[sp]);
    if (rc) return rc;
  }
  if (e->n_split > 1) {
    cudaEvent_t join = e->op_events[n * e->n_split + 1];
    CUDA_TRY(cudaEventRecord(join, stream_of(1, 0)));
    CUDA_TRY(cudaStreamWaitEvent(st, join, 0));
  }
  return YPB_OK;
}

int ypb_infer(ypb_engine* e, void* cuda_stream, const uint8_t* frames, const float* xform, const ypb_infer_params* prm,
              float* det, float* det_lb, int32_t* keep, float* coef, int32_t* count) {
  int rc0 = check_infer_args(e, frames, xform, prm, det, det_lb, keep, coef, count);
  if (rc0) return rc0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  DeviceGuard guard(e->device);
  if (!e->use_graph) return enqueue_infer(e, st, frames, xform, prm, det, det_lb, keep, coef, count);
  ypb_engine::GraphKey key;
  memset(&key, 0, sizeof key);
  key.frames = frames; key.xform = xform; key.det = det; key.det_lb = det_lb; key.keep = keep; key.coef = coef;
  key.count = count; key.cmask = prm->class_mask; key.conf = prm->conf; key.iou = prm->iou; key.max_det = prm->max_det;
  key.agnostic = prm->agnostic_nms; key.impl = e->conv_impl;
  cudaGraphExec_t exec = nullptr;
  for (auto& g : e->graphs)
    if (g.key == key) { exec = g.exec; g.stamp = ++e->graph_clock; break; }
  if (!exec) {
    // capture the ~80 launches of a forward pass on a private stream; replays cost one launch
    if (!e->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_infer(e, e->cap_stream, frames, xform, prm, det, det_lb, keep, coef, count, true);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(YPB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(YPB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
    if (e->graphs.size() >= 8) {  // evict the least recently used
      size_t lru = 0;
      for (size_t i = 1; i < e->graphs.size(); ++i) if (e->graphs[i].stamp < e->graphs[lru].stamp) lru = i;
      cudaGraphExecDestroy(e->graphs[lru].exec);
      e->graphs.erase(e->graphs.begin() + lru);
    }
    e->graphs.push_back({key, exec, ++e->graph_clock});
  }
  CUDA_TRY(cudaGraphLaunch(exec, st));
  return YPB_OK;
}

int ypb_op_count(const ypb_engine* e) { return e ? (int)e->ops.size() + 2 : 0; }

int ypb_op_info(const ypb_engine* e, int i, const char** name, int* kind, double* flops, double* bytes) {
  if (!e || !e->planned || i < 0 || i >= (int)e->ops.size() + 2) return fail(YPB_ERR_ARG, "bad op index / not planned");
  static const char* sel_names[2] = {"decode_filter", "nms"};
  const int n = (int)e->ops.size();
  if (i >= n) {
    *name = sel_names[i - n]; *kind = 100 + (i - n); *flops = 0;
    *bytes = (i == n) ? (double)e->B * e->A * e->no * 4 : (double)e->B * 300 * (6 + 4 + 1 + e->nm) * 4;
    return YPB_OK;
  }
  const Op& op = e->ops[i];
  *name = op.name.c_str(); *kind = (int)op.kind;
  const int B = e->B;
  if (op.kind == OP_STEM) {
    *flops = 2.0 * B * (e->H / 2) * (e->W / 2) * op.cout * 27;
    *bytes = (double)B * e->H * e->W * 3 + (double)B * (e->H / 2) * (e->W / 2) * op.cout * 2;
  } else if (op.kind == OP_CONV) {
    const BufDesc& ib = e->bufs[op.in.buf];
    const double M = (double)B * op.L.oH * op.L.oW;
    *flops = op.flops;
    *bytes = (double)B * ib.H * ib.W * op.cin * 2 + M * op.cout * (op.out_mode == OUT_F32 ? 4 : 2) +
             (double)op.k * op.k * op.cin * op.cout * 2 + (op.res.buf >= 0 ? M * op.cout * 2 : 0.0);
  } else if (op.kind == OP_UPSAMPLE) {
    const BufDesc& ib = e->bufs[op.in.buf];
    *flops = 0; *bytes = (double)B * ib.H * ib.W * op.in.C * 2 * 5;
  } else if (op.kind == OP_DW) {
    const BufDesc &ib = e->bufs[op.in.buf], &ob = e->bufs[op.out.buf];
    *flops = 2.0 * B * ob.H * ob.W * op.cout * op.k * op.k;
    *bytes = (double)B * ib.H * ib.W * op.cout * 2 + (double)B * ob.H * ob.W * op.cout * 2 * (op.res.buf >= 0 ? 2 : 1);
  } else if (op.kind == OP_ATTN) {
    const BufDesc& qb = e->bufs[op.in.buf];
    const double N = (double)qb.H * qb.W;
    *flops = 2.0 * B * op.heads * N * N * (kAttnKD + kAttnHD);
    *bytes = (double)B * N * (qb.C + 2 * op.heads * kAttnHD) * 2;
  } else {
    const BufDesc& ib = e->bufs[op.in.buf];
    *flops = 0; *bytes = (double)B * ib.H * ib.W * op.in.C * 2 * 4;
  }
  return YPB_OK;
}

int ypb_infer_profile(ypb_engine* e, void* cuda_stream, const uint8_t* frames, const float* xform,
                      const ypb_infer_params* prm, float* det, float* det_lb, int32_t* keep, float* coef, int32_t* count,
                      float* op_ms, int capacity) {
  int rc = check_infer_args(e, frames, xform, prm, det, det_lb, keep, coef, count);
  if (rc) return rc;
  const int n = (int)e->ops.size();
  if (!op_ms || capacity < n + 2) return fail(YPB_ERR_ARG, "op_ms too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  DeviceGuard guard(e->device);
  std::vector<cudaEvent_t> ev(n + 3);
  for (auto& x : ev) CUDA_TRY(cudaEventCreate(&x));
  CUDA_TRY(cudaEventRecord(ev[0], st));
  for (int i = 0; i < n; ++i) {
    for (int sp = 0; sp < e->n_split; ++sp) {  // both batch halves of the op, back to back, inside its event bracket
      rc = launch_op(e, e->ops[i], st, frames, sp);
      if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(ev[i + 1], st));
  }
  rc = launch_select(e, st, xform, prm, det, det_lb, keep, coef, count, ev[n + 1]);
  if (rc) return rc;
  CUDA_TRY(cudaEventRecord(ev[n + 2], st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i < n + 2; ++i) CUDA_TRY(cudaEventElapsedTime(op_ms + i, ev[i], ev[i + 1]));
  for (auto& x : ev) cudaEventDestroy(x);
  return YPB_OK;
}

int ypb_masks_ex(ypb_engine* e, void* cuda_stream, int retina, int out_h, int out_w, const float* det, const float* det_lb,
                 const float* coef, const int32_t* count, uint8_t* masks, int capacity, int32_t* status, const float* proto,
                 int32_t* offsets_scratch);

int ypb_masks(ypb_engine* e, void* cuda_stream, int retina, int out_h, int out_w, const float* det, const float* det_lb,
              const float* coef, const int32_t* count, uint8_t* masks, int capacity, int32_t* status) {
  return ypb_masks_ex(e, cuda_stream, retina, out_h, out_w, det, det_lb, coef, count, masks, capacity, status, nullptr, nullptr);
}

int ypb_proto_info(const ypb_engine* e, size_t* offset, size_t* bytes) {
  if (!e || !e->planned || e->proto_buf < 0 || !offset || !bytes) return fail(YPB_ERR_ARG, "no proto buffer / not planned");
  *offset = e->bufs[e->proto_buf].offset;
  *bytes = e->bufs[e->proto_buf].bytes;
  return YPB_OK;
}

// Diagnostics: where the selection stage keeps its per-image candidate counters inside the workspace.
int ypb_select_info(const ypb_engine* e, size_t* count_offset, int* cand_stride) {
  if (!e || !e->planned || !count_offset || !cand_stride) return fail(YPB_ERR_ARG, "bad argument");
  *count_offset = e->off_count;
  *cand_stride = e->hg.cand_stride;
  return YPB_OK;
}

int ypb_masks_ex(ypb_engine* e, void* cuda_stream, int retina, int out_h, int out_w, const float* det, const float* det_lb,
                 const float* coef, const int32_t* count, uint8_t* masks, int capacity, int32_t* status, const float* proto,
                 int32_t* offsets_scratch) {
  if (!e || !det || !det_lb || !coef || !count || !masks || !status || capacity < 1) return fail(YPB_ERR_ARG, "bad argument");
  if (!e->bound || e->nm == 0 || e->proto_buf < 0) return fail(YPB_ERR_STATE, "masks: not a bound -seg engine");
  if (e->nm != 32) return fail(YPB_ERR_ARG, "masks: the decode kernel is built for 32 mask coefficients");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  DeviceGuard guard(e->device);
  const BufDesc& pb = e->bufs[e->proto_buf];
  MaskGeom g{};
  g.mh = pb.H; g.mw = pb.W; g.nm = e->nm; g.max_det = kNmsMaxDet; g.retina = retina ? 1 : 0;
  if (retina) {
    if (out_h < 1 || out_w < 1) return fail(YPB_ERR_ARG, "masks: bad output size");
    // ops.scale_masks: gain/pad in proto space with int() truncation, then bilinear to (H0, W0)
    const double gain = std::min((double)g.mh / out_h, (double)g.mw / out_w);
    const double pad_w = (g.mw - out_w * gain) / 2, pad_h = (g.mh - out_h * gain) / 2;
    const int top = (int)pad_h, left = (int)pad_w, bottom = (int)(g.mh - pad_h), right = (int)(g.mw - pad_w);
    g.top = top; g.left = left; g.ch = bottom - top; g.cw = right - left;
    g.out_h = out_h; g.out_w = out_w;
  } else {
    g.top = 0; g.left = 0; g.ch = g.mh; g.cw = g.mw;
    g.out_h = e->H; g.out_w = e->W;
    g.ratio_w = (float)((double)g.mw / e->W); g.ratio_h = (float)((double)g.mh / e->H);
  }
  if (g.ch < 1 || g.cw < 1) return fail(YPB_ERR_ARG, "masks: empty proto window");
  g.scale_h = (float)g.ch / (float)g.out_h; g.scale_w = (float)g.cw / (float)g.out_w;
  int* offsets = offsets_scratch ? offsets_scratch : reinterpret_cast<int*>(e->ws + e->off_moff);
  mask_offsets_kernel<<<1, 32, 0, st>>>(count, e->B, capacity, offsets, status);
  if (capacity > 65535) return fail(YPB_ERR_ARG, "masks: capacity > 65535");
  if (g.ch > g.out_h || g.cw > g.out_w) {
    // the output is smaller than the proto window (a tiny frame with retina masks): scale_masks down-samples
    const int px = g.out_h * g.out_w;
    mask_decode_small_kernel<<<dim3((px + 255) / 256, capacity), 256, 0, st>>>(
        proto ? proto : reinterpret_cast<const float*>(e->ws + pb.offset), coef, det, det_lb, offsets, e->B, capacity, g, masks);
    CUDA_TRY(cudaGetLastError());
    return YPB_OK;
  }
  {
    // tile-stationary decode when a 64 x 64 output tile needs at most kMaskWin x kMaskWin prototypes (mask_kernels.cuh)
    const int need_h = std::min(g.ch, (int)(kMaskTile * g.scale_h) + 3), need_w = std::min(g.cw, (int)(kMaskTile * g.scale_w) + 3);
    // Measured (profiles/r2_ab_split_branches.md): correct to the bit but slower than the band kernel on the bench
    // workloads (boxes of the synthetic recipe are large: the per-pixel blend, not the prototype reads, is the bulk of
    // the work, and the band kernel blends each window row once per column) - opt-in with YPB_MASK_TILES=1.
    static const bool tiles = getenv("YPB_MASK_TILES") != nullptr;
    if (tiles && need_h <= kMaskWin && need_w <= kMaskWin) {
      const size_t smem = (size_t)(32 + 2) * kMaskWin * kMaskWin * sizeof(float);
      {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        DeviceState& ds = device_state();
        if (!ds.mask_tile_attr) {
          CUDA_TRY(cudaFuncSetAttribute(mask_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          ds.mask_tile_attr = true;
        }
      }
      g.prefilled = 1;
      mask_zero_kernel<<<device_state().num_sms * 8, 256, 0, st>>>(offsets, e->B, capacity, (long long)g.out_h * g.out_w, masks);
      dim3 tgrid((g.out_w + kMaskTile - 1) / kMaskTile, (g.out_h + kMaskTile - 1) / kMaskTile, e->B);
      mask_tile_kernel<<<tgrid, 256, smem, st>>>(proto ? proto : reinterpret_cast<const float*>(e->ws + pb.offset), coef, det, det_lb,
                                                 offsets, e->B, capacity, g, masks);
      CUDA_TRY(cudaGetLastError());
      return YPB_OK;
    }
  }
  const int bands = (g.out_h + kMaskTile - 1) / kMaskTile;
  dim3 grid(bands, capacity);
  if (capacity > 65535) return fail(YPB_ERR_ARG, "masks: capacity > 65535");
  // dynamic smem = the band's logit window: rows needed by 64 output rows x every column of the un-padded proto window
  const int band_rows = std::min(g.ch, (int)(kMaskTile * g.scale_h) + 3);
  const size_t band_smem = (size_t)band_rows * g.cw * sizeof(float);
  if (band_smem > 160 * 1024) return fail(YPB_ERR_ARG, "masks: proto window too large for the band buffer");
  if (band_smem > 30 * 1024) {
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceState& ds = device_state();
    if (band_smem > ds.mask_smem) {
      CUDA_TRY(cudaFuncSetAttribute(mask_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      CUDA_TRY(cudaFuncSetAttribute(mask_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      ds.mask_smem = 160 * 1024;
    }
  }
  // Measured (profiles/r2_ab_split_branches.md): the two-kernel form is SLOWER (0.40 vs 0.33 ms for 1 824 detections at
  // 640x640): the zero fill was never what bounds this kernel - the per-band logit window and the blends are - so the
  // one-kernel form stays the default and the streaming fill is an opt-in (YPB_MASK_PREFILL=1).
  static const bool prefill = getenv("YPB_MASK_PREFILL") != nullptr;
  g.prefilled = prefill ? 1 : 0;
  if (g.prefilled)
    mask_zero_kernel<<<device_state().num_sms * 8, 256, 0, st>>>(offsets, e->B, capacity, (long long)g.out_h * g.out_w, masks);
  const float* proto_p = proto ? proto : reinterpret_cast<const float*>(e->ws + pb.offset);
  if (g.retina) mask_decode_kernel<true><<<grid, 256, band_smem, st>>>(proto_p, coef, det, det_lb, offsets, e->B, capacity, g, masks);
  else mask_decode_kernel<false><<<grid, 256, band_smem, st>>>(proto_p, coef, det, det_lb, offsets, e->B, capacity, g, masks);
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

int ypb_device_error(ypb_engine* e, uint32_t* word) {
  if (!word) return fail(YPB_ERR_ARG, "bad argument");
  DeviceGuard guard(e ? e->device : -1);
  unsigned int v = 0;
  CUDA_TRY(cudaMemcpyFromSymbol(&v, g_dev_error, sizeof v));
  *word = v;
  if (v) {
    unsigned int z = 0;
    CUDA_TRY(cudaMemcpyToSymbol(g_dev_error, &z, sizeof z));
  }
  return YPB_OK;
}

int ypb_device_error_async(ypb_engine* e, void* cuda_stream, uint32_t* host_word) {
  if (!e || !host_word) return fail(YPB_ERR_ARG, "bad argument");
  DeviceGuard guard(e->device);
  CUDA_TRY(cudaMemcpyFromSymbolAsync(host_word, g_dev_error, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost,
                                     reinterpret_cast<cudaStream_t>(cuda_stream)));
  return YPB_OK;
}

int ypb_view_count(const ypb_engine* e) { return e ? (int)e->views.size() + 1 : 0; }

int ypb_view_info(const ypb_engine* e, int i, const char** name, size_t* offset, int* H, int* W, int* Ctot, int* c_off,
                  int* C, int* dtype) {
  if (!e || !e->planned || i < 0 || i > (int)e->views.size()) return fail(YPB_ERR_ARG, "bad view index / not planned");
  if (i == (int)e->views.size()) {  // the fp32 head rows, as a (B, 1, A, no) "image"
    static const char* hn = "head";
    *name = hn; *offset = e->off_head; *H = 1; *W = e->A; *Ctot = e->no; *c_off = 0; *C = e->no; *dtype = 1;
    return YPB_OK;
  }
  const NamedView& nv = e->views[i];
  const BufDesc& b = e->bufs[nv.v.buf];
  *name = nv.name.c_str(); *offset = b.offset; *H = b.H; *W = b.W; *Ctot = b.C; *c_off = nv.v.c_off; *C = nv.v.C;
  *dtype = b.dtype;
  return YPB_OK;
}

// ---- stand-alone kernels for parity tests -----------------------------------------------------
int ypb_conv2d_bf16(void* cuda_stream, const void* in, int B, int H, int W, int in_ctot, int in_c_off, int cin,
                    const void* wg, const float* bias, int cout, int k, int stride, int act, const void* res, void* out,
                    int out_ctot, int out_c_off, int out_fp32, int impl) {
  ConvDesc d;
  d.in = in; d.B = B; d.Hin = H; d.Win = W; d.in_ctot = in_ctot; d.in_c_off = in_c_off; d.cin = cin;
  d.wg = wg; d.bias = bias; d.cout = cout; d.k = k; d.stride = stride; d.act = act;
  d.out_mode = out_fp32 ? OUT_F32 : OUT_BF16;
  const int oH = H / stride, oW = W / stride;
  d.out = out; d.out_img_stride = (long long)oH * oW * out_ctot; d.out_pix_stride = out_ctot; d.out_c_off = out_c_off;
  if (res) { d.res = res; d.res_img_stride = d.out_img_stride; d.res_pix_stride = out_ctot; d.res_c_off = out_c_off; }
  ConvLaunch L;
  std::string err;
#if !YPB_DIAG
  if (impl != 0) return fail(YPB_ERR_ARG, "conv impl 1-3 are debugging twins: load libypb200_diag.so (include/ypb200_diag.h)");
#endif
  if (!conv_plan_geometry(d, &L, &err)) return fail(YPB_ERR_ARG, err);
  if (!conv_bind(d, &L, &err)) return fail(YPB_ERR_CUDA, err);
  CUDA_TRY(conv_launch(L, reinterpret_cast<cudaStream_t>(cuda_stream), impl));
  return YPB_OK;
}

#if YPB_DIAG  // ---- diagnostics and micro-benchmarks: libypb200_diag.so only (include/ypb200_diag.h) ----
// Diagnostics: time `iters` back-to-back launches of one conv (planned once) with CUDA events on `cuda_stream`.
// dbg >= 0 overrides the YPB_DBG experiment mask of the launch (see ConvParams::dbg).
int ypb_conv_bench(void* cuda_stream, const void* in, int B, int H, int W, int in_ctot, int in_c_off, int cin,
                   const void* wg, const float* bias, int cout, int k, int stride, int act, const void* res, void* out,
                   int out_ctot, int out_c_off, int out_fp32, int impl, int dbg, int iters, float* ms, char* desc,
                   int desc_len) {
  if (!ms || iters < 1) return fail(YPB_ERR_ARG, "bad argument");
  ConvDesc d;
  d.in = in; d.B = B; d.Hin = H; d.Win = W; d.in_ctot = in_ctot; d.in_c_off = in_c_off; d.cin = cin;
  d.wg = wg; d.bias = bias; d.cout = cout; d.k = k; d.stride = stride; d.act = act;
  d.out_mode = out_fp32 == 2 ? OUT_SHUFFLE2_BF16 : out_fp32 ? OUT_F32 : OUT_BF16;
  const int oH = H / stride, oW = W / stride;
  if (d.out_mode == OUT_SHUFFLE2_BF16) {
    d.out = out; d.out_img_stride = 4LL * oH * oW * out_ctot; d.out_pix_stride = out_ctot; d.out_c_off = out_c_off;
  } else {
    d.out = out; d.out_img_stride = (long long)oH * oW * out_ctot; d.out_pix_stride = out_ctot; d.out_c_off = out_c_off;
  }
  if (res) { d.res = res; d.res_img_stride = d.out_img_stride; d.res_pix_stride = out_ctot; d.res_c_off = out_c_off; }
  ConvLaunch L;
  std::string err;
  if (!conv_plan_geometry(d, &L, &err)) return fail(YPB_ERR_ARG, err);
  if (!conv_bind(d, &L, &err)) return fail(YPB_ERR_CUDA, err);
  if (dbg >= 0) { L.p.dbg = dbg; L.p1.dbg = dbg; }
  if (desc && desc_len > 0) conv_describe(L, impl, desc, desc_len);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) CUDA_TRY(conv_launch(L, st, impl));
  if (dbg >= 0 && (dbg & 8)) {
    unsigned long long z[16] = {0};
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemcpyToSymbol(g_conv_prof, z, sizeof z));
  }
  CUDA_TRY(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) CUDA_TRY(conv_launch(L, st, impl));
  CUDA_TRY(cudaEventRecord(e1, st));
  CUDA_TRY(cudaEventSynchronize(e1));
  CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
  *ms /= (float)iters;
  if (dbg >= 0 && (dbg & 8) && desc && desc_len > 0) {
    unsigned long long z[16];
    CUDA_TRY(cudaMemcpyFromSymbol(z, g_conv_prof, sizeof z));
    const double n = z[7] ? (double)z[7] : 1.0, life = (double)z[5] / n;
    const size_t len = strlen(desc);
    snprintf(desc + len, desc_len - len,
             " | per-CTA kcycles: life %.1f prod.wait_empty %.1f mma.wait_full %.1f mma.wait_tempty %.1f mma.wait_w %.1f "
             "epi.wait_tfull %.1f epi.drain %.1f mma.issue %.1f mma.commit %.1f",
             life / 1e3, z[0] / n / 1e3, z[1] / n / 1e3, z[2] / n / 1e3, z[6] / n / 1e3, z[3] / n / 1e3, z[4] / n / 1e3,
             z[15] / n / 1e3, z[14] / n / 1e3);
    const double np = z[13] ? (double)z[13] : 1.0;
    const size_t len2 = strlen(desc);
    snprintf(desc + len2, desc_len - len2, " | cycles per epilogue pass: ld+wait %.0f release %.0f math %.0f stage %.0f writeout %.0f",
             z[8] / np, z[9] / np, z[10] / np, z[11] / np, z[12] / np);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return YPB_OK;
}

// Diagnostics (profiling build): read and optionally clear the 16 device-side cycle counters.
int ypb_debug_prof(unsigned long long* out16, int reset) {
  if (!out16) return fail(YPB_ERR_ARG, "bad argument");
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out16, g_conv_prof, 16 * sizeof(unsigned long long)));
  if (reset) {
    unsigned long long z[16] = {0};
    CUDA_TRY(cudaMemcpyToSymbol(g_conv_prof, z, sizeof z));
  }
  return YPB_OK;
}

// Diagnostics: operand-fetch ceiling of the TMA path (see tma_bench.cuh).  buf: device, >= rows*128 bytes (mode 0/1)
// or B*H*W*128 bytes (mode 2).  Returns elapsed milliseconds in *ms and bytes moved in *bytes.
int ypb_tma_bench(void* buf, int mode, int stages, int iters, int rows, int W, int H, int B, float* ms, double* bytes) {
  if (!buf || !ms || !bytes || stages < 1 || stages > 12) return fail(YPB_ERR_ARG, "bad argument");
  CUtensorMap m2, m5;
  std::string err;
  {
    cuuint64_t dims[5] = {64, (cuuint64_t)(mode == 2 ? (long long)B * H * W : rows), 1, 1, 1};
    cuuint64_t str[4] = {128, dims[1] * 128, dims[1] * 128, dims[1] * 128};
    cuuint32_t box[5] = {64, (cuuint32_t)((mode != 2 && H == 256) ? 256 : 128), 1, 1, 1};
    if (!encode_bf16_map(&m2, buf, 5, dims, str, box, &err)) return fail(YPB_ERR_CUDA, err);
  }
  const int halo_R = (rows >> 16) & 0xff, halo_C = rows & 0xffff;
  if (mode == 4) {
    if (halo_R < 3 || halo_C < 8) return fail(YPB_ERR_ARG, "mode 4: rows = C_tot | R << 16 | producers << 24");
    cuuint64_t dims[5] = {(cuuint64_t)halo_C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 1};
    cuuint64_t str[4] = {(cuuint64_t)halo_C * 2, (cuuint64_t)halo_C * 2 * W, (cuuint64_t)halo_C * 2 * W * H, (cuuint64_t)halo_C * 2 * W * H * B};
    cuuint32_t box[5] = {64, 10, (cuuint32_t)halo_R, 1, 1};
    if (!encode_bf16_map(&m5, buf, 5, dims, str, box, &err)) return fail(YPB_ERR_CUDA, err);
  } else if (mode == 3) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(H == 256 ? 256 : 128)};
    if (!encode_bf16_map(&m5, buf, 2, dims, str, box, &err)) return fail(YPB_ERR_CUDA, err);
  } else {
    cuuint64_t dims[5] = {64, (cuuint64_t)(mode == 2 ? W : 16), (cuuint64_t)(mode == 2 ? H : 8), (cuuint64_t)(mode == 2 ? B : 1), 1};
    cuuint64_t str[4] = {128, dims[1] * 128, dims[1] * dims[2] * 128, dims[1] * dims[2] * dims[3] * 128};
    cuuint32_t box[5] = {64, 16, 8, 1, 1};
    if (!encode_bf16_map(&m5, buf, 5, dims, str, box, &err)) return fail(YPB_ERR_CUDA, err);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int stage_b = mode == 4 ? ((halo_R * 10 * 128 + 1023) & ~1023) : (mode != 2 && H == 256) ? 32768 : 16384;
  const int smem = stages * stage_b + 1024 + 256;
  if (smem > 227 * 1024) return fail(YPB_ERR_ARG, "too many stages");
  CUDA_TRY(cudaFuncSetAttribute(tma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  TmaBenchParams p{mode, stages, iters, rows, W, H, B};
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  tma_bench_kernel<<<sms, 256, smem>>>(m2, m5, p);  // warm-up
  CUDA_TRY(cudaEventRecord(e0));
  tma_bench_kernel<<<sms, 256, smem>>>(m2, m5, p);
  CUDA_TRY(cudaEventRecord(e1));
  CUDA_TRY(cudaEventSynchronize(e1));
  CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *bytes = (double)sms * iters * (double)stage_b;
  return YPB_OK;
}

// Diagnostics: tensor-pipe ceiling for M=128 x N MMAs fed from shared memory, optionally with concurrent TMA traffic.
int ypb_mma_bench(void* buf, int rows, int n, int iters, int shifted, int tma_iters, float* ms) {
  if (!buf || !ms || n < 16 || n > 256 || (n % 16)) return fail(YPB_ERR_ARG, "bad argument");
  CUtensorMap m2;
  std::string err;
  cuuint64_t dims[5] = {64, (cuuint64_t)rows, 1, 1, 1};
  cuuint64_t str[4] = {128, dims[1] * 128, dims[1] * 128, dims[1] * 128};
  cuuint32_t box[5] = {64, 128, 1, 1, 1};
  if (!encode_bf16_map(&m2, buf, 5, dims, str, box, &err)) return fail(YPB_ERR_CUDA, err);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 131072 + 1024 + 256;
  CUDA_TRY(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  MmaBenchParams p{n, iters, shifted, tma_iters, rows};
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  mma_bench_kernel<<<sms, 96, smem>>>(m2, p);
  CUDA_TRY(cudaEventRecord(e0));
  mma_bench_kernel<<<sms, 96, smem>>>(m2, p);
  CUDA_TRY(cudaEventRecord(e1));
  CUDA_TRY(cudaEventSynchronize(e1));
  CUDA_TRY(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return YPB_OK;
}

int ypb_latency_probe(long long* out_dev) {
  if (!out_dev) return fail(YPB_ERR_ARG, "bad argument");
  latency_probe_kernel<<<1, 64>>>(out_dev);
  CUDA_TRY(cudaDeviceSynchronize());
  return YPB_OK;
}
#endif  // YPB_DIAG

// Host-side helper: copy n frames into the pinned staging buffer with `nthreads` host threads (persistent pool,
// 256 KB pieces, optional non-temporal stores: csrc/host_stage.cpp).  The Python caller releases the GIL for the
// duration of the call.  Frame staging is the host stage of the predict() pipeline.
extern "C" int ypb_host_stage_frames(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads, int mode);

int ypb_stage_frames_ex(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads, int mode) {
  if (!dst || !src || !bytes || n < 0 || mode < 0 || mode > 1) return fail(YPB_ERR_ARG, "bad argument");
  return ypb_host_stage_frames(dst, src, bytes, n, nthreads, mode);
}

// Asynchronous staging: queue the copy of n frames into pinned memory (returns at once) and place a gate in `cuda_stream`:
// work enqueued on that stream after this call (the chunk's H2D copy) starts only when the frames are staged.  The source
// frames and both pointer targets must stay alive until the stream has passed the gate.
extern "C" void* ypb_host_stage_submit(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads, int mode);
extern "C" void ypb_host_stage_wait(void* ticket);
static void CUDART_CB stage_gate_fn(void* ticket) { ypb_host_stage_wait(ticket); }

int ypb_stage_frames_gated(void* cuda_stream, void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads) {
  if (!dst || !src || !bytes || n < 0) return fail(YPB_ERR_ARG, "bad argument");
  if (n == 0) return YPB_OK;
  static const int mode = getenv("YPB_STAGE_MEMCPY") ? 0 : 1;
  void* ticket = ypb_host_stage_submit(dst, src, bytes, n, nthreads, mode);
  cudaError_t ce = cudaLaunchHostFunc(reinterpret_cast<cudaStream_t>(cuda_stream), stage_gate_fn, ticket);
  if (ce != cudaSuccess) {
    ypb_host_stage_wait(ticket);  // no gate in the stream: finish the copy here so that nothing dangles
    return fail(YPB_ERR_CUDA, std::string("cudaLaunchHostFunc: ") + cudaGetErrorString(ce));
  }
  return YPB_OK;
}

int ypb_stage_frames(void* const* dst, const void* const* src, const size_t* bytes, int n, int nthreads) {
  static const int mode = getenv("YPB_STAGE_MEMCPY") ? 0 : 1;
  return ypb_stage_frames_ex(dst, src, bytes, n, nthreads, mode);
}

// Needle length on the device: minimum-area rectangle of every mask (mask_kernels.cuh).
int ypb_mask_min_rect(void* cuda_stream, const uint8_t* masks, int n, int H, int W, int32_t* row_extents /*(n,H,2)*/,
                      float* out /*(n,2): length, length/width*/) {
  if (n < 0 || H < 1 || W < 1 || (n > 0 && (!masks || !row_extents || !out))) return fail(YPB_ERR_ARG, "bad argument");
  if (2 * H > kRectMaxPts || W > 32767) return fail(YPB_ERR_ARG, "mask_min_rect: frame larger than 2304 rows / 32767 columns");
  if (n == 0) return YPB_OK;
  if (n > 65535) return fail(YPB_ERR_ARG, "mask_min_rect: more than 65535 masks per call");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  mask_row_extents_kernel<<<dim3((H + 3) / 4, n), 128, 0, st>>>(masks, H, W, reinterpret_cast<int2*>(row_extents));
  mask_min_rect_kernel<<<n, 256, 0, st>>>(reinterpret_cast<const int2*>(row_extents), H, out);
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

// LetterBox on the device (misc_kernels.cuh): all pointers device, tables built by the caller as cv2 builds them.
int ypb_letterbox_u8(void* cuda_stream, const uint8_t* src, int B, int H0, int W0, uint8_t* dst, int H, int W, int new_w,
                     int new_h, int top, int left, const int32_t* xofs, const int16_t* xa, const int32_t* yofs,
                     const int16_t* ya, int pad_value) {
  if (!src || !dst || !xofs || !xa || !yofs || !ya || B < 1 || H0 < 1 || W0 < 1 || H < 1 || W < 1 || new_w < 1 || new_h < 1 ||
      top < 0 || left < 0 || top + new_h > H || left + new_w > W)
    return fail(YPB_ERR_ARG, "bad argument");
  const long long total = (long long)B * H * W;
  letterbox_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      src, B, H0, W0, dst, H, W, new_w, new_h, top, left, xofs, xa, yofs, ya, pad_value);
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

// Index-mask hand-off to the tracker (reference yolo_seg/yolo_with_deva.py:54-88), see mask_kernels.cuh.
int ypb_index_masks(void* cuda_stream, const uint8_t* masks, const int32_t* offsets, int B, int n_total, int H, int W,
                    int min_area, int32_t* area, int32_t* ids, int64_t* index_map) {
  if (!offsets || !index_map || B < 1 || H < 1 || W < 1 || n_total < 0 || (n_total > 0 && (!masks || !area || !ids)))
    return fail(YPB_ERR_ARG, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const long long hw = (long long)H * W;
  if (n_total > 0) {
    if (n_total > 65535 || B > 65535) return fail(YPB_ERR_ARG, "index_masks: more than 65535 masks / frames per call");
    CUDA_TRY(cudaMemsetAsync(area, 0, (size_t)n_total * 4, st));
    const int slices = (int)std::max(1LL, std::min(64LL, hw / (256 * 16)));
    mask_area_kernel<<<dim3(slices, n_total), 256, 0, st>>>(masks, hw, area);
    mask_ids_kernel<<<(B + 127) / 128, 128, 0, st>>>(offsets, B, area, min_area, ids);
  }
  index_paint_kernel<<<dim3((unsigned)((hw + 2047) / 2048), B), 256, 0, st>>>(masks, offsets, ids, hw,
                                                                             reinterpret_cast<long long*>(index_map));
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

// The hand-off for masks that are zero outside rects[i] = (x0, y0, x1, y1) (exclusive ends): what predict() produces,
// every mask cropped to its box.  rects == nullptr, W % 16 != 0 or unaligned masks: the full-scan kernels above.
int ypb_index_masks_boxed(void* cuda_stream, const uint8_t* masks, const int32_t* offsets, const int32_t* rects, int B, int n_total,
                          int H, int W, int min_area, int32_t* area, int32_t* ids, int64_t* index_map) {
  if (!rects || (W & 15) != 0 || (reinterpret_cast<uintptr_t>(masks) & 15) != 0 || n_total == 0)
    return ypb_index_masks(cuda_stream, masks, offsets, B, n_total, H, W, min_area, area, ids, index_map);
  if (!offsets || !index_map || B < 1 || H < 1 || W < 1 || n_total < 0 || !masks || !area || !ids) return fail(YPB_ERR_ARG, "bad argument");
  if (n_total > 65535 || B > 65535 || (H + kPaintTH - 1) / kPaintTH > 65535)
    return fail(YPB_ERR_ARG, "index_masks: more than 65535 masks / frames / tile rows per call");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  CUDA_TRY(cudaMemsetAsync(area, 0, (size_t)n_total * 4, st));
  mask_area_rect_kernel<<<dim3(8, n_total), 256, 0, st>>>(masks, reinterpret_cast<const int4*>(rects), H, W, area);
  mask_ids_kernel<<<(B + 127) / 128, 128, 0, st>>>(offsets, B, area, min_area, ids);
  index_paint_rect_kernel<<<dim3((W + kPaintTW - 1) / kPaintTW, (H + kPaintTH - 1) / kPaintTH, B), 256, 0, st>>>(
      masks, offsets, ids, reinterpret_cast<const int4*>(rects), H, W, reinterpret_cast<long long*>(index_map));
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

// Same hand-off when predict() ran on a resized frame (the reference's `min_side`, yolo_with_deva.py:44-48,71-72): masks
// (n_total, h1, w1) are resized to (H, W) as torchvision F.resize does (antialiased bilinear), area_f receives the float
// sum of every resized mask (the reference filters on `mask.sum() < MIN_AREA`), bins (n_total, H, W) uint8 scratch.
int ypb_index_masks_resized(void* cuda_stream, const uint8_t* masks, const int32_t* offsets, int B, int n_total, int h1, int w1,
                            int H, int W, float min_area, uint8_t* bins, float* area_f, int32_t* ids, int64_t* index_map) {
  if (!offsets || !index_map || B < 1 || H < 1 || W < 1 || h1 < 1 || w1 < 1 || n_total < 0 ||
      (n_total > 0 && (!masks || !bins || !area_f || !ids)))
    return fail(YPB_ERR_ARG, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const long long hw = (long long)H * W;
  if (n_total > 0) {
    if (n_total > 65535 || B > 65535 || (H + 7) / 8 > 65535) return fail(YPB_ERR_ARG, "index_masks: too many masks / frames / rows per call");
    CUDA_TRY(cudaMemsetAsync(area_f, 0, (size_t)n_total * 4, st));
    mask_resize_aa_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, n_total), 256, 0, st>>>(masks, n_total, h1, w1, H, W, bins, area_f);
    mask_ids_f_kernel<<<(B + 127) / 128, 128, 0, st>>>(offsets, B, area_f, min_area, ids);
  }
  index_paint_kernel<<<dim3((unsigned)((hw + 2047) / 2048), B), 256, 0, st>>>(bins, offsets, ids, hw,
                                                                             reinterpret_cast<long long*>(index_map));
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

// ---- JPEG frames decoded on the device (csrc/jpeg_source.cpp: nvJPEG, loaded lazily) ----------------------------------
extern "C" int ypb_host_jpeg_info(const unsigned char* data, size_t n, int* h, int* w, char* err, int errlen);
extern "C" int ypb_host_jpeg_decode(void* stream, const unsigned char* data, size_t n, unsigned char* dst_dev, int H, int W, char* err,
                                    int errlen);

int ypb_jpeg_info(const uint8_t* jpeg_host, size_t nbytes, int* height, int* width) {
  if (!jpeg_host || nbytes == 0 || !height || !width) return fail(YPB_ERR_ARG, "bad argument");
  char err[256] = {0};
  if (ypb_host_jpeg_info(jpeg_host, nbytes, height, width, err, sizeof err) != 0) return fail(YPB_ERR_ARG, err);
  return YPB_OK;
}

int ypb_jpeg_decode_bgr(void* cuda_stream, const uint8_t* jpeg_host, size_t nbytes, uint8_t* dst_dev, int height, int width) {
  if (!jpeg_host || nbytes == 0 || !dst_dev || height < 1 || width < 1) return fail(YPB_ERR_ARG, "bad argument");
  char err[256] = {0};
  if (ypb_host_jpeg_decode(cuda_stream, jpeg_host, nbytes, dst_dev, height, width, err, sizeof err) != 0) return fail(YPB_ERR_CUDA, err);
  return YPB_OK;
}

// ---- point-to-point mask hand-off between the GPUs of one box (BASELINE config C5) ----------------
// The tracker (DEVA) is sequential and lives on ONE GPU; the detector replicas on the other GPUs push their per-frame
// index masks into a mailbox in the consumer GPU's memory over NVLink / NVSwitch: cudaMemcpyAsync between peer
// devices, exported across the one-process-per-GPU boundary with CUDA IPC.  No collective, no host bounce.
int ypb_mailbox_create(int device, size_t bytes, void** dev_ptr, unsigned char handle[64]) {
  if (!dev_ptr || !handle || bytes == 0) return fail(YPB_ERR_ARG, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  DeviceGuard guard(device);
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t ce = cudaIpcGetMemHandle(&h, p);
  if (ce != cudaSuccess) { cudaFree(p); return fail(YPB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(ce)); }
  memcpy(handle, &h, 64);
  *dev_ptr = p;
  return YPB_OK;
}

int ypb_mailbox_destroy(int device, void* dev_ptr) {
  if (!dev_ptr) return YPB_OK;
  DeviceGuard guard(device);
  CUDA_TRY(cudaFree(dev_ptr));
  return YPB_OK;
}

// Producer side: map the consumer's mailbox into this process (device = the producer's own GPU).
int ypb_mailbox_open(int device, const unsigned char handle[64], void** dev_ptr) {
  if (!handle || !dev_ptr) return fail(YPB_ERR_ARG, "bad argument");
  DeviceGuard guard(device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return YPB_OK;
}

int ypb_mailbox_close(int device, void* dev_ptr) {
  if (!dev_ptr) return YPB_OK;
  DeviceGuard guard(device);
  CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return YPB_OK;
}

// Enqueue a device-to-device copy (same GPU or a peer's mailbox) on the caller's stream.
int ypb_peer_copy(void* cuda_stream, void* dst, const void* src, size_t bytes) {
  if (!dst || !src) return fail(YPB_ERR_ARG, "bad argument");
  if (bytes == 0) return YPB_OK;
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(cuda_stream)));
  return YPB_OK;
}

// Host helpers of the zero-staging path: frames that already live in page-locked memory (a capture / decode ring, a
// pinned torch tensor) are copied to the device straight from where they are.
int ypb_host_is_pinned(const void* p, int* pinned) {
  if (!p || !pinned) return fail(YPB_ERR_ARG, "bad argument");
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *pinned = 0;
    return YPB_OK;
  }
  *pinned = a.type == cudaMemoryTypeHost ? 1 : 0;
  return YPB_OK;
}

int ypb_hosts_are_pinned(const void* const* ptrs, int n, int* all_pinned) {
  if (!ptrs || !all_pinned || n < 0) return fail(YPB_ERR_ARG, "bad argument");
  *all_pinned = 1;
  for (int i = 0; i < n && *all_pinned; ++i) {
    int one = 0;
    if (ptrs[i] == nullptr) return fail(YPB_ERR_ARG, "null frame pointer");
    const int rc = ypb_host_is_pinned(ptrs[i], &one);
    if (rc != YPB_OK) return rc;
    if (!one) *all_pinned = 0;
  }
  return YPB_OK;
}

// n frames of `bytes_each` bytes -> consecutive slots of dst_dev, one cudaMemcpyAsync per run of adjacent sources.
int ypb_h2d_frames(void* cuda_stream, void* dst_dev, const void* const* src, size_t bytes_each, int n) {
  if (!dst_dev || !src || n < 0) return fail(YPB_ERR_ARG, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  uint8_t* d = reinterpret_cast<uint8_t*>(dst_dev);
  int i = 0;
  while (i < n) {
    int j = i + 1;
    while (j < n && reinterpret_cast<const uint8_t*>(src[j]) == reinterpret_cast<const uint8_t*>(src[j - 1]) + bytes_each) ++j;
    CUDA_TRY(cudaMemcpyAsync(d + (size_t)i * bytes_each, src[i], (size_t)(j - i) * bytes_each, cudaMemcpyHostToDevice, st));
    i = j;
  }
  return YPB_OK;
}

size_t ypb_nms_scratch_bytes(int B, int N) {
  int cs = 1;
  while (cs < N) cs <<= 1;
  return (size_t)B * cs * 8 + (size_t)B * sizeof(FrameXform) + 1024;
}

__global__ void nms_test_pack_kernel(const float* boxes, const float* scores, const int* n_valid, int B, int N, int cs,
                                     unsigned long long* keys, FrameXform* xf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) xf[i] = FrameXform{0.f, 0.f, 1.f, 3.0e38f, 3.0e38f};
  if (i >= B * N) return;
  const int b = i / N, a = i - b * N;
  if (a < n_valid[b])
    keys[(long long)b * cs + a] = ((unsigned long long)__float_as_uint(scores[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a);
  (void)boxes;
}

int ypb_nms(void* cuda_stream, const float* boxes, const float* scores, const int32_t* cls, const int32_t* n_valid, int B,
            int N, float iou, int max_det, int agnostic, void* scratch, int32_t* keep, int32_t* count) {
  if (!boxes || !scores || !cls || !n_valid || !scratch || !keep || !count || max_det < 1 || max_det > kNmsMaxDet)
    return fail(YPB_ERR_ARG, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  HeadGeom g{};
  g.A = N; g.no = 0; g.nc = 0; g.nm = 0;
  int cs = 1;
  while (cs < N) cs <<= 1;
  g.cand_stride = cs;
  uint8_t* s = reinterpret_cast<uint8_t*>(scratch);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(s);
  FrameXform* xf = reinterpret_cast<FrameXform*>(s + (size_t)B * cs * 8);
  nms_test_pack_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(boxes, scores, n_valid, B, N, cs, keys, xf);
  nms_kernel<<<B, kNmsThreads, 0, st>>>(nullptr, g, reinterpret_cast<const float4*>(boxes), cls, keys, n_valid, iou, max_det, 30000,
                                agnostic ? 0.0f : 7680.0f, 0, nullptr, max_det, xf, nullptr, nullptr, keep, nullptr, count);
  CUDA_TRY(cudaGetLastError());
  return YPB_OK;
}

}  // extern "C"
