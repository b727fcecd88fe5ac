// The KeypointCompleter step engine: parameter-arena layout, activation workspace, and the
// forward (model.py:100-170) / backward call sequences over the kernels of this library.
// Everything is enqueued on the caller's stream; no allocation, no synchronisation.
#include <map>
#include <string>
#include <vector>

#include "attention.cuh"
#include "gemm_sm100.cuh"
#include "gemm_wgrad_group.cuh"
#include "ffn_fused.cuh"
#include "rowops.cuh"

namespace kit {

struct LinearW {
  int64_t w = -1, b = -1;  // fp32 arena offsets (floats)
  int rows = 0, cols = 0;  // W is [rows, cols] (out_features, in_features)
  int64_t wb = -1, wbT = -1;  // bf16 arena offsets (elements); wbT = W^T [cols, ldT]
  int ld = 0, ldT = 0;
  bool need_T = true;
};
struct LNW { int64_t g = -1, b = -1; };
struct AttnW { LinearW in, out; };
struct EncW { AttnW sa; LinearW l1, l2; LNW n1, n2; };
struct DecW { AttnW sa, ca; LinearW l1, l2; LNW n1, n2, n3; };
struct SwiW { LinearW fc12, fc3; };

struct Entry {
  std::string name;
  int64_t offset, numel, rows, cols;
  int is_buffer;
};

struct Layout {
  KitModelConfig cfg;
  std::vector<Entry> entries;
  int64_t trainable = 0, total = 0;
  int64_t wb_elems = 0;
  std::vector<std::pair<int64_t, int64_t>> buckets;  // in backward completion order
  std::vector<char> dec_cut, enc_cut;                // [layer]: backward reports a finished bucket after this layer
  int64_t learned_i, learned_f, pe_i, pe_f;
  LinearW emb_i, emb_f, fc_final;
  SwiW swi_i, swi_f, swi_d;
  std::vector<EncW> enc;
  std::vector<DecW> dec;
  LNW enc_norm, dec_norm;
  std::vector<WeightDesc> wdescs;
  // The cross-attention key / value projections of ALL decoder layers read the same tensor (the encoder output): their weights
  // are kept a second time as one [layers * 2H, H] B operand (and its transpose [H, layers * 2H]) so the forward is one GEMM with
  // N = layers * 2H and the gradient with respect to the encoder output one GEMM with K = layers * 2H.
  int64_t kvcat_wb = -1, kvcat_wbT = -1;
  int kvcat_n = 0;   // layers * 2H
};

static int64_t up8(int64_t x) { return round_up(x, 8); }

struct LayoutBuilder {
  Layout& L;
  int64_t cur = 0, wcur = 0;
  explicit LayoutBuilder(Layout& l) : L(l) {}
  int64_t add(const std::string& name, int64_t rows, int64_t cols, int is_buffer = 0) {
    const int64_t off = cur;
    L.entries.push_back({name, off, rows * cols, rows, cols, is_buffer});
    cur = up8(cur + rows * cols);
    return off;
  }
  void add_wb(LinearW& w) {
    w.ld = (int)up8(w.cols);
    w.wb = wcur;
    wcur = round_up(wcur + (int64_t)w.rows * w.ld, 64);
    if (w.need_T) {
      w.ldT = (int)up8(w.rows);
      w.wbT = wcur;
      wcur = round_up(wcur + (int64_t)w.cols * w.ldT, 64);
    }
    WeightDesc d;
    d.src_off = w.w; d.rows = w.rows; d.cols = w.cols;
    d.dst_off = w.wb; d.dst_ld = w.ld; d.rows_pad = w.rows; d.cols_pad = w.ld;
    d.dstT_off = w.wbT; d.dstT_ld = w.ldT;
    L.wdescs.push_back(d);
  }
  // weight and bias registered under the reference's names
  void linear(LinearW& w, const std::string& prefix, int rows, int cols, bool need_T = true) {
    w.rows = rows; w.cols = cols; w.need_T = need_T;
    w.w = add(prefix + ".weight", rows, cols);
  }
  void linear_bias(LinearW& w, const std::string& prefix) { w.b = add(prefix + ".bias", 1, w.rows); }
  void ln(LNW& n, const std::string& prefix, int H) {
    n.g = add(prefix + ".weight", 1, H);
    n.b = add(prefix + ".bias", 1, H);
  }
  // SwiGLU: fc1 | fc2 stored back to back so they run as ONE [2H,H] GEMM
  void swiglu(SwiW& s, const std::string& prefix, int H) {
    s.fc12.rows = 2 * H; s.fc12.cols = H; s.fc12.need_T = true;
    s.fc12.w = add(prefix + ".fc1.weight", H, H);
    add(prefix + ".fc2.weight", H, H);
    s.fc12.b = add(prefix + ".fc1.bias", 1, H);
    add(prefix + ".fc2.bias", 1, H);
    linear(s.fc3, prefix + ".fc3", H, H);
    linear_bias(s.fc3, prefix + ".fc3");
  }
  void attn(AttnW& a, const std::string& prefix, int H) {
    a.in.rows = 3 * H; a.in.cols = H;
    a.in.w = add(prefix + ".in_proj_weight", 3 * H, H);
    a.in.b = add(prefix + ".in_proj_bias", 1, 3 * H);
    linear(a.out, prefix + ".out_proj", H, H);
    linear_bias(a.out, prefix + ".out_proj");
  }
};

static int build_layout(const KitModelConfig* c, Layout& L) {
  KIT_REQUIRE(c != nullptr, "null model config");
  KIT_REQUIRE(c->hidden > 0 && c->hidden % 64 == 0 && c->hidden <= 1024, "hidden %d unsupported (multiple of 64, <= 1024)", c->hidden);
  KIT_REQUIRE(c->heads > 0 && c->hidden % c->heads == 0, "hidden %d not divisible by heads %d", c->hidden, c->heads);
  const int d = c->hidden / c->heads;
  KIT_REQUIRE(d == 16 || d == 32 || d == 64, "head dim %d unsupported (16, 32, 64)", d);
  KIT_REQUIRE(c->input_size > 0 && c->input_size % 2 == 0, "input_size must be 2*K");
  KIT_REQUIRE(c->layers > 0 && c->ff > 0 && c->ff % 8 == 0 && c->max_len > 0, "bad layers/ff/max_len");
  KIT_REQUIRE(c->variant == KIT_MODEL_COMPLETER || c->variant == KIT_MODEL_CYCLE, "unknown model variant %d", c->variant);
  L.cfg = *c;
  const int H = c->hidden, IN = c->input_size, FF = c->ff;
  LayoutBuilder lb(L);
  // ---- group E: embeddings, learned PEs, the two pre-transformer SwiGLUs (last bucket of backward)
  L.learned_i = lb.add("learned_input_positional_encoder", 1, H);
  L.learned_f = lb.add("learned_filled_positional_encoder", 1, H);
  lb.linear(L.emb_i, "input_embedding", H, IN, false);
  lb.linear_bias(L.emb_i, "input_embedding");
  lb.linear(L.emb_f, "filled_embedding", H, IN, false);
  lb.linear_bias(L.emb_f, "filled_embedding");
  lb.swiglu(L.swi_i, "swiGlu_input_prev", H);
  lb.swiglu(L.swi_f, "swiGlu_filled_prev", H);
  const int64_t end_group_e = lb.cur;
  L.enc.resize(c->layers);
  L.dec.resize(c->layers);
  std::vector<int64_t> enc_begin(c->layers), dec_begin(c->layers);
  for (int l = 0; l < c->layers; ++l) {
    enc_begin[l] = lb.cur;
    const std::string p = "transformer.encoder.layers." + std::to_string(l);
    EncW& e = L.enc[l];
    lb.attn(e.sa, p + ".self_attn", H);
    lb.linear(e.l1, p + ".linear1", FF, H); lb.linear_bias(e.l1, p + ".linear1");
    lb.linear(e.l2, p + ".linear2", H, FF); lb.linear_bias(e.l2, p + ".linear2");
    lb.ln(e.n1, p + ".norm1", H);
    lb.ln(e.n2, p + ".norm2", H);
  }
  lb.ln(L.enc_norm, "transformer.encoder.norm", H);
  const int64_t end_enc = lb.cur;
  for (int l = 0; l < c->layers; ++l) {
    dec_begin[l] = lb.cur;
    const std::string p = "transformer.decoder.layers." + std::to_string(l);
    DecW& e = L.dec[l];
    lb.attn(e.sa, p + ".self_attn", H);
    lb.attn(e.ca, p + ".multihead_attn", H);
    lb.linear(e.l1, p + ".linear1", FF, H); lb.linear_bias(e.l1, p + ".linear1");
    lb.linear(e.l2, p + ".linear2", H, FF); lb.linear_bias(e.l2, p + ".linear2");
    lb.ln(e.n1, p + ".norm1", H);
    lb.ln(e.n2, p + ".norm2", H);
    lb.ln(e.n3, p + ".norm3", H);
  }
  lb.ln(L.dec_norm, "transformer.decoder.norm", H);
  lb.swiglu(L.swi_d, "swiGlu_decoded", H);
  lb.linear(L.fc_final, "fc_final", IN, H);
  lb.linear_bias(L.fc_final, "fc_final");
  L.trainable = lb.cur;
  L.pe_i = lb.add("trig_input_positional_encoder.pos_encoding", c->max_len, H, 1);
  L.pe_f = lb.add("trig_filled_positional_encoder.pos_encoding", c->max_len, H, 1);
  L.total = lb.cur;
  (void)end_group_e;
  // Buckets in the order backward completes them: the decoder stack, encoder layers 1.., and a small tail -- the all-reduce of
  // the LAST bucket cannot overlap anything, so it holds only encoder layer 0 and the embedding / pre-transformer group (5.9 MB
  // of the 72 MB arena; it was a quarter of the arena with four equal buckets).  Finer buckets (KIT_BUCKET_LAYERS = layers per
  // bucket) measured slower on 2 and 8 GPUs: every cut is one more graph launch and one more collective (profiles/r02_dp.md).
  L.dec_cut.assign(c->layers, 0);
  L.enc_cut.assign(c->layers, 0);
  int per = 0;   // layers per bucket (KIT_BUCKET_LAYERS: A/B measurements; 0 = one bucket per stack + the tail)
  if (const char* v = getenv("KIT_BUCKET_LAYERS")) per = atoi(v);
  for (int l = 1; l < c->layers; ++l) {
    L.dec_cut[l] = (per > 0 && l % per == 0) ? 1 : 0;
    L.enc_cut[l] = ((per > 0 && l % per == 0) || l == 1) ? 1 : 0;
  }
  L.buckets.clear();
  {
    int64_t hi = L.trainable;
    for (int l = c->layers - 1; l >= 1; --l)
      if (L.dec_cut[l]) {
        L.buckets.push_back({dec_begin[l], hi});
        hi = dec_begin[l];
      }
    L.buckets.push_back({dec_begin[0], hi});   // decoder finished
    hi = end_enc;
    for (int l = c->layers - 1; l >= 1; --l)
      if (L.enc_cut[l]) {
        L.buckets.push_back({enc_begin[l], hi});
        hi = enc_begin[l];
      }
    L.buckets.push_back({0, hi});              // encoder layer 0 (and below) + embeddings, learned PEs, the two input SwiGLUs
  }
  // bf16 GEMM operands
  lb.add_wb(L.emb_i); lb.add_wb(L.emb_f);
  lb.add_wb(L.swi_i.fc12); lb.add_wb(L.swi_i.fc3);
  lb.add_wb(L.swi_f.fc12); lb.add_wb(L.swi_f.fc3);
  for (int l = 0; l < c->layers; ++l) {
    EncW& e = L.enc[l];
    lb.add_wb(e.sa.in); lb.add_wb(e.sa.out); lb.add_wb(e.l1); lb.add_wb(e.l2);
  }
  for (int l = 0; l < c->layers; ++l) {
    DecW& e = L.dec[l];
    lb.add_wb(e.sa.in); lb.add_wb(e.sa.out); lb.add_wb(e.ca.in); lb.add_wb(e.ca.out); lb.add_wb(e.l1); lb.add_wb(e.l2);
  }
  lb.add_wb(L.swi_d.fc12); lb.add_wb(L.swi_d.fc3);
  lb.add_wb(L.fc_final);
  L.kvcat_n = c->layers * 2 * H;
  L.kvcat_wb = lb.wcur;
  lb.wcur = round_up(lb.wcur + (int64_t)L.kvcat_n * H, 64);
  L.kvcat_wbT = lb.wcur;
  lb.wcur = round_up(lb.wcur + (int64_t)H * L.kvcat_n, 64);
  for (int l = 0; l < c->layers; ++l) {   // rows H .. 3H of in_proj_weight (torch/nn/functional.py:6360 _in_projection_packed)
    WeightDesc d;
    d.src_off = L.dec[l].ca.in.w + (int64_t)H * H; d.rows = 2 * H; d.cols = H;
    d.dst_off = L.kvcat_wb + (int64_t)l * 2 * H * H; d.dst_ld = H; d.rows_pad = 2 * H; d.cols_pad = H;
    d.dstT_off = L.kvcat_wbT + (int64_t)l * 2 * H; d.dstT_ld = L.kvcat_n;
    L.wdescs.push_back(d);
  }
  L.wb_elems = lb.wcur;
  return KIT_OK;
}

// ---------------------------------------------------------------- engine
struct Buf {
  int64_t off;  // bytes into the workspace
  int64_t elems;
  int esize;
  int64_t ld;
};

struct EncAct { bf16 *qkv, *ao, *s1, *x1, *z, *hh, *s2, *x2; float *lse, *st1, *st2; };
struct DecAct { bf16 *qkv, *ao, *s1, *y1, *qc, *kvc, *aoc, *s2, *y2, *z, *hh, *s3, *y3; float *lse, *lsec, *st1, *st2, *st3; };   // kvc: this layer's columns of kv_all

}  // namespace kit

using namespace kit;

struct KitEngine {
  Layout L;
  int B = 0, T = 0;
  int training = 1;  // 0: layers share one set of activation buffers (inference only)
  int64_t M = 0;
  int K2p = 0;
  std::map<std::string, Buf> bufs;
  int64_t ws_bytes = 0;
  float* params = nullptr;
  float* grads = nullptr;
  uint8_t* ws = nullptr;
  bool bound = false, weights_fresh = false;
  std::vector<GemmPlan> fwd_plans, bwd_plans;
  std::vector<WgradGroupPlan> group_plans;   // one per layer of the backward (grouped stream-K weight gradients)
  size_t group_cursor = 0;
  std::vector<FfnPlan> ffn_plans, ffn_bwd_plans;   // one per layer of the forward / backward (fused feed-forward block)
  size_t ffn_cursor = 0, ffn_bwd_cursor = 0;
  bool fuse_ffn = true;
  // LayerNorm backward in the epilogue of the kernel that produces its output gradient (EPI_ADD_LNBWD, ffn_kernel<true>).  Opt-in
  // (KIT_FUSE_LNBWD=1): it removes 28 of the 32 ln_bwd launches of a step, but with one accumulator row per lane the dgamma / dbeta
  // column sums need a warp transpose-reduce that costs more than the launches it saves (4.45 vs 4.34 ms per step at B = 256; 4.20
  // without the column sums) -- see profiles/r01c_summary.md.
  bool fuse_lnbwd = false;
  size_t cursor = 0;
  std::vector<GemmPlan>* active = nullptr;
  int64_t launches = 0;
  cudaStream_t st = nullptr;
  // optional per-category CUDA-event profiling (bench.py roofline leg)
  bool profiling = false;
  struct ProfRec { int cat; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pool;
  size_t ev_used = 0;
  double cat_flops[KIT_PROF_CATEGORIES] = {0};
  int64_t cat_launches[KIT_PROF_CATEGORIES] = {0};
  // saved masks (backward reuses the forward's)
  KitAttnMask enc_mask{}, dec_mask{};
  // resolved pointers
  bf16* wb = nullptr;
  WeightDesc* wdesc_dev = nullptr;
  int* tile_prefix_dev = nullptr;
  int total_tiles = 0;
  bf16 *xe, *xd, *ei_raw, *ei, *si12, *sig, *x0, *ef_raw, *ef, *sf12, *sfg, *y0;
  bf16 *mem, *dec_out, *sd12, *sdg, *sd, *zf, *sf, *dp;
  float *st_encn, *st_decn;
  std::vector<EncAct> ea;
  std::vector<DecAct> da;
  bf16 *g0, *g1, *g1b, *g1c, *g2, *g3, *gmem, *gff, *gqkv, *gqc, *gkv, *g2h;   // gkv: [M, layers * 2H], one column block per decoder layer
  bf16* kv_all;     // [M, layers * 2H]: cross-attention keys / values of every decoder layer (one GEMM over the encoder output)
  float* kv_bias;   // [layers * 2H]: the in_proj_bias[H:3H] slices gathered at weight-refresh time
  float* dq_acc;

  int64_t alloc(const std::string& name, int64_t elems, int esize, int64_t ld) {
    Buf b{ws_bytes, elems, esize, ld};
    bufs[name] = b;
    ws_bytes = round_up(ws_bytes + elems * esize, 256);
    return b.off;
  }
};

namespace kit {

template <typename T>
static T* wsptr(KitEngine* e, const std::string& name) {
  return reinterpret_cast<T*>(e->ws + e->bufs.at(name).off);
}

static void plan_workspace(KitEngine* e) {
  const Layout& L = e->L;
  const int64_t M = e->M;
  const int H = L.cfg.hidden, FF = L.cfg.ff, NH = L.cfg.heads;
  const int64_t lse_n = (int64_t)e->B * NH * e->T;
  e->alloc("wb", L.wb_elems, 2, 0);
  e->alloc("wdesc", (int64_t)L.wdescs.size() * sizeof(WeightDesc), 1, 0);
  e->alloc("tile_prefix", (int64_t)L.wdescs.size() * 4 + 4, 1, 0);
  e->alloc("xe", M * e->K2p, 2, e->K2p);
  e->alloc("xd", M * e->K2p, 2, e->K2p);
  for (const char* n : {"ei_raw", "ei", "sig", "x0", "ef_raw", "ef", "sfg", "y0", "mem", "dec_out", "sdg", "sd", "zf", "sf"})
    e->alloc(n, M * H, 2, H);
  for (const char* n : {"si12", "sf12", "sd12"}) e->alloc(n, M * 2 * H, 2, 2 * H);
  e->alloc("dp", M * e->K2p, 2, e->K2p);
  e->alloc("kv_all", M * L.kvcat_n, 2, L.kvcat_n);
  e->alloc("kv_bias", L.kvcat_n, 4, 0);
  e->alloc("st_encn", 2 * M, 4, 0);
  e->alloc("st_decn", 2 * M, 4, 0);
  const int n_saved = e->training ? L.cfg.layers : 1;
  // the [M, FF] hidden activation of the feed-forward block exists in HBM only for the backward pass (or the two-GEMM path)
  const int64_t zh_elems = (!e->training && e->fuse_ffn && ffn_fwd_supported(H, FF)) ? 8 : M * FF;
  for (int l = 0; l < n_saved; ++l) {
    const std::string p = "enc" + std::to_string(l) + ".";
    e->alloc(p + "qkv", M * 3 * H, 2, 3 * H);
    for (const char* n : {"ao", "s1", "x1", "s2", "x2"}) e->alloc(p + n, M * H, 2, H);
    e->alloc(p + "z", zh_elems, 2, FF);
    e->alloc(p + "hh", zh_elems, 2, FF);
    e->alloc(p + "lse", lse_n, 4, 0);
    e->alloc(p + "st1", 2 * M, 4, 0);
    e->alloc(p + "st2", 2 * M, 4, 0);
  }
  for (int l = 0; l < n_saved; ++l) {
    const std::string p = "dec" + std::to_string(l) + ".";
    e->alloc(p + "qkv", M * 3 * H, 2, 3 * H);
    for (const char* n : {"ao", "s1", "y1", "qc", "aoc", "s2", "y2", "s3", "y3"}) e->alloc(p + n, M * H, 2, H);
    e->alloc(p + "z", zh_elems, 2, FF);
    e->alloc(p + "hh", zh_elems, 2, FF);
    e->alloc(p + "lse", lse_n, 4, 0);
    e->alloc(p + "lsec", lse_n, 4, 0);
    for (const char* n : {"st1", "st2", "st3"}) e->alloc(p + n, 2 * M, 4, 0);
  }
  // backward scratch
  const int64_t gm = e->training ? M : 8;
  for (const char* n : {"g0", "g1", "g1b", "g1c", "g2", "g3", "gmem", "gqc"}) e->alloc(n, gm * H, 2, H);
  e->alloc("gff", gm * FF, 2, FF);
  e->alloc("gqkv", gm * 3 * H, 2, 3 * H);
  e->alloc("gkv", gm * L.kvcat_n, 2, L.kvcat_n);
  e->alloc("g2h", gm * 2 * H, 2, 2 * H);
  e->alloc("dq_acc", (e->training && e->T > 64) ? M * H + M * NH : 8, 4, H);
}

static void resolve_pointers(KitEngine* e) {
  const Layout& L = e->L;
  const int H = L.cfg.hidden;
  e->wb = wsptr<bf16>(e, "wb");
  e->wdesc_dev = wsptr<WeightDesc>(e, "wdesc");
  e->tile_prefix_dev = wsptr<int>(e, "tile_prefix");
#define KIT_P(n) e->n = wsptr<bf16>(e, #n)
  KIT_P(xe); KIT_P(xd); KIT_P(ei_raw); KIT_P(ei); KIT_P(si12); KIT_P(sig); KIT_P(x0); KIT_P(ef_raw); KIT_P(ef);
  KIT_P(sf12); KIT_P(sfg); KIT_P(y0); KIT_P(mem); KIT_P(dec_out); KIT_P(sd12); KIT_P(sdg); KIT_P(sd); KIT_P(zf); KIT_P(sf);
  KIT_P(dp); KIT_P(g0); KIT_P(g1); KIT_P(g1b); KIT_P(g1c); KIT_P(gqc); KIT_P(g2); KIT_P(g3); KIT_P(gmem); KIT_P(gff); KIT_P(gqkv); KIT_P(gkv); KIT_P(g2h);
#undef KIT_P
  e->dq_acc = wsptr<float>(e, "dq_acc");
  e->kv_all = wsptr<bf16>(e, "kv_all");
  e->kv_bias = wsptr<float>(e, "kv_bias");
  e->st_encn = wsptr<float>(e, "st_encn");
  e->st_decn = wsptr<float>(e, "st_decn");
  e->ea.resize(L.cfg.layers);
  e->da.resize(L.cfg.layers);
  for (int l = 0; l < L.cfg.layers; ++l) {
    const std::string p = "enc" + std::to_string(e->training ? l : 0) + ".";
    EncAct& a = e->ea[l];
    a.qkv = wsptr<bf16>(e, p + "qkv"); a.ao = wsptr<bf16>(e, p + "ao"); a.s1 = wsptr<bf16>(e, p + "s1");
    a.x1 = wsptr<bf16>(e, p + "x1"); a.z = wsptr<bf16>(e, p + "z"); a.hh = wsptr<bf16>(e, p + "hh");
    a.s2 = wsptr<bf16>(e, p + "s2"); a.x2 = wsptr<bf16>(e, p + "x2");
    a.lse = wsptr<float>(e, p + "lse"); a.st1 = wsptr<float>(e, p + "st1"); a.st2 = wsptr<float>(e, p + "st2");
    const std::string q = "dec" + std::to_string(e->training ? l : 0) + ".";
    DecAct& d = e->da[l];
    d.qkv = wsptr<bf16>(e, q + "qkv"); d.ao = wsptr<bf16>(e, q + "ao"); d.s1 = wsptr<bf16>(e, q + "s1");
    d.y1 = wsptr<bf16>(e, q + "y1"); d.qc = wsptr<bf16>(e, q + "qc"); d.kvc = e->kv_all + (int64_t)l * 2 * H;
    d.aoc = wsptr<bf16>(e, q + "aoc"); d.s2 = wsptr<bf16>(e, q + "s2"); d.y2 = wsptr<bf16>(e, q + "y2");
    d.z = wsptr<bf16>(e, q + "z"); d.hh = wsptr<bf16>(e, q + "hh"); d.s3 = wsptr<bf16>(e, q + "s3");
    d.y3 = wsptr<bf16>(e, q + "y3");
    d.lse = wsptr<float>(e, q + "lse"); d.lsec = wsptr<float>(e, q + "lsec");
    d.st1 = wsptr<float>(e, q + "st1"); d.st2 = wsptr<float>(e, q + "st2"); d.st3 = wsptr<float>(e, q + "st3");
  }
}

static void prof_begin(KitEngine* e, int cat, double flops) {
  e->cat_launches[cat]++;
  e->cat_flops[cat] += flops;
  if (!e->profiling) return;
  if (e->ev_used >= e->ev_pool.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    e->ev_pool.push_back({a, b});
  }
  auto& ev = e->ev_pool[e->ev_used++];
  e->prof.push_back({cat, ev.first, ev.second});
  cudaEventRecord(ev.first, e->st);
}
static void prof_end(KitEngine* e) {
  if (!e->profiling) return;
  cudaEventRecord(e->prof.back().b, e->st);
}
static void prof_reset(KitEngine* e) {
  e->prof.clear();
  e->ev_used = 0;
  for (int i = 0; i < KIT_PROF_CATEGORIES; ++i) {
    e->cat_flops[i] = 0;
    e->cat_launches[i] = 0;
  }
}

// GEMM through the per-engine plan cache (tensor maps are built once per call site).
static int eg(KitEngine* e, int mode, const bf16* A, int64_t lda, const bf16* Bm, int64_t ldb, void* C, int64_t ldc, int M,
              int N, int K, const float* bias, const bf16* addend, int64_t ld_add, int out_kind, int act, bf16* aux,
              int64_t ld_aux, float* bias_grad = nullptr, bool* bias_grad_fused = nullptr, const GemmLN* ln = nullptr,
              bool* ln_fused = nullptr, const LnBwdArgs* lnb = nullptr, bool* lnb_fused = nullptr) {
  std::vector<GemmPlan>& plans = *e->active;
  if (e->cursor >= plans.size()) {
    GemmPlan p;
    int rc = gemm_plan(&p, mode, A, lda, Bm, ldb, C, ldc, M, N, K, bias, addend, ld_add, out_kind, act, aux, ld_aux,
                       mode == 1 ? 0 : 1, bias_grad, ln, lnb);
    if (rc) return rc;
    plans.push_back(p);
  }
  GemmPlan& p = plans[e->cursor++];
  if (p.p.C != C) {  // caller memory moved (pred): the output tensor map must be rebuilt
    int rc = gemm_plan(&p, mode, A, lda, Bm, ldb, C, ldc, M, N, K, bias, addend, ld_add, out_kind, act, aux, ld_aux,
                       mode == 1 ? 0 : 1, bias_grad, ln, lnb);
    if (rc) return rc;
  }
  if (bias_grad_fused != nullptr) *bias_grad_fused = p.p.bias_grad != nullptr;
  if (ln_fused != nullptr) *ln_fused = p.epi == EPI_ADD_LN;
  if (lnb_fused != nullptr) *lnb_fused = p.epi == EPI_ADD_LNBWD;
  e->launches++;
  prof_begin(e, mode == 0 ? KIT_PROF_GEMM_TN : KIT_PROF_GEMM_WGRAD, 2.0 * (double)M * (double)N * (double)K);
  const int rc = gemm_launch(&p, e->st);
  prof_end(e);
  return rc;
}
// attention through the profiler: 4*S*S*d flops per head forward, 10*S*S*d backward
static int eattn_fwd(KitEngine* e, const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv,
                     bf16* out, int64_t ldo, float* lse, const KitAttnMask* mask) {
  const int H = e->L.cfg.hidden, NH = e->L.cfg.heads, d = H / NH;
  e->launches++;
  prof_begin(e, KIT_PROF_ATTN_FWD, 4.0 * e->B * NH * (double)e->T * e->T * d);
  const int rc = attention_fwd(q, ldq, k, ldk, v, ldv, out, ldo, lse, e->B, NH, e->T, e->T, d, mask, e->st);
  prof_end(e);
  return rc;
}
static int eattn_bwd(KitEngine* e, const bf16* q, int64_t ldq, const bf16* k, int64_t ldk, const bf16* v, int64_t ldv,
                     const bf16* o, int64_t ldo, const bf16* dout, int64_t ld_do, const float* lse, bf16* dq, int64_t ld_dq,
                     bf16* dk, int64_t ld_dk, bf16* dv, int64_t ld_dv, const KitAttnMask* mask) {
  const int H = e->L.cfg.hidden, NH = e->L.cfg.heads, d = H / NH;
  e->launches += e->T > 64 ? 3 : 1;   // one persistent tile kernel, or rowsum(dO * O) + streaming kernel + fp32 -> bf16 dQ
  prof_begin(e, KIT_PROF_ATTN_BWD, 10.0 * e->B * NH * (double)e->T * e->T * d);
  const int rc = attention_bwd(q, ldq, k, ldk, v, ldv, o, ldo, dout, ld_do, lse, dq, ld_dq, dk, ld_dk, dv, ld_dv,
                               e->T > 64 ? e->dq_acc : nullptr, e->B, NH, e->T, e->T, d, mask, e->st);
  prof_end(e);
  return rc;
}
#define KIT_TRY(x)      \
  do {                  \
    int _rc = (x);      \
    if (_rc) return _rc; \
  } while (0)

// y = x W^T + b (+ addend)
static int linear_fwd(KitEngine* e, const bf16* x, int64_t ldx, const LinearW& w, int row0, int nrows, bf16* y, int64_t ldy,
                      const bf16* addend, int64_t ld_add, int act = ACT_NONE, bf16* aux = nullptr, int64_t ld_aux = 0) {
  return eg(e, 0, x, ldx, e->wb + w.wb + (int64_t)row0 * w.ld, w.ld, y, ldy, (int)e->M, nrows, w.cols,
            e->params + w.b + row0, addend, ld_add, OUT_BF16, act, aux, ld_aux);
}
// s = x W^T + b + addend ; y = LN(s) (torch/nn/modules/transformer.py:956 post-norm): one GEMM with the LayerNorm in its
// epilogue when the row fits one tile (H = 256), else GEMM + row kernel
static int linear_add_ln_fwd(KitEngine* e, const bf16* x, int64_t ldx, const LinearW& w, bf16* s, const bf16* addend, const LNW& n,
                             bf16* y, float* stats) {
  const int H = e->L.cfg.hidden;
  const int64_t M = e->M;
  GemmLN ln{e->params + n.g, e->params + n.b, stats, stats + M, y, H, 1e-5f};
  bool fused = false;
  KIT_TRY(eg(e, 0, x, ldx, e->wb + w.wb, w.ld, s, H, (int)M, H, w.cols, e->params + w.b, addend, H, OUT_BF16, ACT_NONE, nullptr, 0,
             nullptr, nullptr, &ln, &fused));
  if (fused) return KIT_OK;
  e->launches++;
  return add_ln_fwd(s, nullptr, e->params + n.g, e->params + n.b, nullptr, y, stats, stats + M, M, H, e->st);
}
// dx = dy W[row0:row0+nrows, :] (+ addend)   (B operand = rows of W^T restricted to those columns)
static int linear_dgrad(KitEngine* e, const bf16* dy, int64_t ld_dy, const LinearW& w, int row0, int nrows, bf16* dx,
                        int64_t ld_dx, const bf16* addend, int64_t ld_add, int act = ACT_NONE, bf16* aux = nullptr,
                        int64_t ld_aux = 0, const LnBwdArgs* lnb = nullptr) {
  bool fused = false;
  KIT_TRY(eg(e, 0, dy, ld_dy, e->wb + w.wbT + row0, w.ldT, dx, ld_dx, (int)e->M, w.cols, nrows, nullptr, addend, ld_add,
             OUT_BF16, act, aux, ld_aux, nullptr, nullptr, nullptr, nullptr, lnb, &fused));
  KIT_REQUIRE(lnb == nullptr || fused, "linear_dgrad: the planner refused the fused LayerNorm backward");
  return KIT_OK;
}
// LayerNorm backward fused behind its producer is possible when a row is one 256-wide tile
static bool lnbwd_fusable(const KitEngine* e) { return e->fuse_lnbwd && e->L.cfg.hidden == 256; }
static LnBwdArgs lnbwd_args(KitEngine* e, const bf16* s, const float* stats, const LNW& n) {
  return LnBwdArgs{s, e->L.cfg.hidden, e->params + n.g, stats, stats + e->M, e->grads + n.g, e->grads + n.b};
}
// dx_ln = LayerNormBackward(dy) with dy = din W[row0:row0+nrows, :] + addend: in the GEMM epilogue when possible (then the
// bias gradient of the Linear that feeds the LayerNorm is left to its weight-gradient GEMM: *bias_done = false), else
// GEMM -> tmp -> ln_bwd kernel (which also sums dx's columns into dxsum: *bias_done = true).
static int dgrad_then_lnbwd(KitEngine* e, const bf16* din, int64_t ld_in, const LinearW& w, int row0, int nrows, const bf16* addend,
                            bf16* tmp, bf16* dx_ln, const bf16* s, const float* stats, const LNW& n, float* dxsum, bool* bias_done) {
  const int H = e->L.cfg.hidden;
  const int64_t M = e->M;
  if (lnbwd_fusable(e) && addend != nullptr && gemm_lnbwd_supported((int)M)) {
    const LnBwdArgs lb = lnbwd_args(e, s, stats, n);
    *bias_done = false;
    return linear_dgrad(e, din, ld_in, w, row0, nrows, dx_ln, H, addend, H, ACT_NONE, nullptr, 0, &lb);
  }
  KIT_TRY(linear_dgrad(e, din, ld_in, w, row0, nrows, tmp, H, addend, H));
  e->launches++;
  *bias_done = true;
  return ln_bwd(tmp, s, stats, stats + M, e->params + n.g, nullptr, dx_ln, e->grads + n.g, e->grads + n.b, dxsum, M, H, e->st);
}
// dW[row0:row0+nrows, :] += dy^T x ; db[row0:...] += colsum(dy)
static int linear_wgrad(KitEngine* e, const bf16* dy, int64_t ld_dy, const bf16* x, int64_t ldx, const LinearW& w, int row0,
                        int nrows, bool bias_done = false) {
  bool fused = false;   // db = column sums of dy: from the same GEMM (all-ones MMA) whenever its epilogue kind allows
  KIT_TRY(eg(e, 1, dy, ld_dy, x, ldx, e->grads + w.w + (int64_t)row0 * w.cols, w.cols, nrows, w.cols, (int)e->M, nullptr,
             nullptr, 0, OUT_F32_ATOMIC, ACT_NONE, nullptr, 0, bias_done ? nullptr : e->grads + w.b + row0, &fused));
  if (bias_done || fused) return KIT_OK;   // (bias_done: the LayerNorm backward that produced dy already summed its columns)
  e->launches++;
  return colsum(dy, ld_dy, e->grads + w.b + row0, e->M, (int)up8(nrows), e->st);
}

// A weight gradient whose launch is deferred to the end of its layer's backward (its dy buffer stays untouched until then).
struct PendingW {
  const bf16* dy;
  int64_t ld_dy;
  const bf16* x;
  int64_t ldx;
  const LinearW* w;
  int row0, nrows;
  bool bias_done;
};
// All weight (and bias) gradients of one layer in one grouped stream-K launch (gemm_wgrad_group.cuh); one by one if a
// pitch rules the tensor maps out.
static int flush_wgrads(KitEngine* e, std::vector<PendingW>& pend) {
  if (pend.empty()) return KIT_OK;
  WgradProblemDesc d[WG_MAX_PROBLEMS];
  bool ok = pend.size() <= (size_t)WG_MAX_PROBLEMS;
  double flops = 0.0;
  for (size_t i = 0; ok && i < pend.size(); ++i) {
    const PendingW& q = pend[i];
    d[i].A = q.dy; d[i].lda = q.ld_dy;
    d[i].B = q.x; d[i].ldb = q.ldx;
    d[i].C = e->grads + q.w->w + (int64_t)q.row0 * q.w->cols; d[i].ldc = q.w->cols;
    d[i].M = q.nrows; d[i].N = q.w->cols;
    d[i].bias_grad = q.bias_done ? nullptr : e->grads + q.w->b + q.row0;
    ok = wgrad_group_supported(d[i]);
    flops += 2.0 * (double)e->M * q.nrows * q.w->cols;
  }
  if (!ok) {
    for (const PendingW& q : pend) KIT_TRY(linear_wgrad(e, q.dy, q.ld_dy, q.x, q.ldx, *q.w, q.row0, q.nrows, q.bias_done));
    pend.clear();
    return KIT_OK;
  }
  if (e->group_cursor >= e->group_plans.size()) {
    WgradGroupPlan plan;
    KIT_TRY(wgrad_group_plan(&plan, d, (int)pend.size(), (int)e->M));
    e->group_plans.push_back(plan);
  }
  const WgradGroupPlan& plan = e->group_plans[e->group_cursor++];
  e->launches++;
  prof_begin(e, KIT_PROF_GEMM_WGRAD, flops);
  const int rc = wgrad_group_launch(&plan, e->st);
  prof_end(e);
  pend.clear();
  return rc;
}

// s = x + linear2(gelu(linear1(x))) ; y = LN(s): one fused kernel (ffn_fused.cuh) when the shape allows, else two GEMMs.
// z / hh (pre-activation / activation, [M, FF]) are written only when the engine keeps activations for a backward pass.
static int ffn_block_fwd(KitEngine* e, const bf16* x, const LinearW& l1, const LinearW& l2, bf16* z, bf16* hh, bf16* s,
                         const LNW& n, bf16* y, float* stats) {
  const int H = e->L.cfg.hidden, FF = e->L.cfg.ff;
  const int64_t M = e->M;
  if (e->fuse_ffn && ffn_fwd_supported(H, FF)) {
    if (e->ffn_cursor >= e->ffn_plans.size()) {
      FfnPlan plan;
      KIT_TRY(ffn_fwd_plan(&plan, x, H, e->wb + l1.wb, l1.ld, e->wb + l2.wb, l2.ld, e->params + l1.b, e->params + l2.b, z, hh, FF,
                           s, H, y, H, e->params + n.g, e->params + n.b, stats, stats + M, 1e-5f, (int)M, H, FF, e->training));
      e->ffn_plans.push_back(plan);
    }
    const FfnPlan& plan = e->ffn_plans[e->ffn_cursor++];
    e->launches++;
    prof_begin(e, KIT_PROF_FFN, 4.0 * (double)M * H * FF);
    const int rc = ffn_launch(&plan, e->st);
    prof_end(e);
    return rc;
  }
  KIT_TRY(linear_fwd(e, x, H, l1, 0, FF, hh, FF, nullptr, 0, ACT_GELU, z, FF));
  return linear_add_ln_fwd(e, hh, FF, l2, s, x, n, y, stats);
}

// Input gradients of the same block: dz = (g W2) * gelu'(z) -> gff (kept for the weight gradients), dx = dz W1 + g.
// Then, when ln_s != null, the backward of the LayerNorm in front of the block (its output gradient is dx): dx_ln, with the same
// fused / unfused choice and *bias_done meaning as dgrad_then_lnbwd.
static int ffn_block_bwd(KitEngine* e, const bf16* g, const LinearW& l1, const LinearW& l2, const bf16* z, bf16* gff, bf16* dx,
                         bf16* dx_ln, const bf16* ln_s, const float* ln_stats, const LNW& n, float* dxsum, bool* bias_done) {
  const int H = e->L.cfg.hidden, FF = e->L.cfg.ff;
  const int64_t M = e->M;
  if (e->fuse_ffn && ffn_fwd_supported(H, FF) && l2.wbT >= 0 && l1.wbT >= 0) {
    const bool fuse_ln = lnbwd_fusable(e);
    if (e->ffn_bwd_cursor >= e->ffn_bwd_plans.size()) {
      FfnPlan plan;
      const LnBwdArgs lb = lnbwd_args(e, ln_s, ln_stats, n);
      KIT_TRY(ffn_bwd_plan(&plan, g, H, e->wb + l2.wbT, l2.ldT, e->wb + l1.wbT, l1.ldT, z, gff, FF, fuse_ln ? dx_ln : dx, H, (int)M, H,
                           FF, fuse_ln ? &lb : nullptr));
      e->ffn_bwd_plans.push_back(plan);
    }
    const FfnPlan& plan = e->ffn_bwd_plans[e->ffn_bwd_cursor++];
    e->launches++;
    prof_begin(e, KIT_PROF_FFN, 4.0 * (double)M * H * FF);
    const int rc = ffn_launch(&plan, e->st);
    prof_end(e);
    if (rc) return rc;
    if (fuse_ln) {
      *bias_done = false;
      return KIT_OK;
    }
    e->launches++;
    *bias_done = true;
    return ln_bwd(dx, ln_s, ln_stats, ln_stats + M, e->params + n.g, nullptr, dx_ln, e->grads + n.g, e->grads + n.b, dxsum, M, H, e->st);
  }
  KIT_TRY(linear_dgrad(e, g, H, l2, 0, H, gff, FF, nullptr, 0, ACT_GELU_BWD, const_cast<bf16*>(z), FF));
  return dgrad_then_lnbwd(e, gff, FF, l1, 0, FF, g, dx, dx_ln, ln_s, ln_stats, n, dxsum, bias_done);
}

static int swiglu_fwd(KitEngine* e, const bf16* x, const SwiW& s, bf16* x12, bf16* g, bf16* out) {
  const int H = e->L.cfg.hidden;
  KIT_TRY(linear_fwd(e, x, H, s.fc12, 0, 2 * H, x12, 2 * H, nullptr, 0));
  e->launches++;
  KIT_TRY(swiglu_gate_fwd(x12, g, e->M, H, e->st));
  return linear_fwd(e, g, H, s.fc3, 0, H, out, H, nullptr, 0);
}
// dout -> dx (into dx_out); uses g2h / a scratch [M,H].  The two weight gradients are QUEUED (pend): the caller flushes them as one
// grouped launch while dout and g2h are still intact (each was a single 9.5 us launch on the critical path of the step).
static int swiglu_bwd(KitEngine* e, const bf16* dout, const bf16* x, const SwiW& s, const bf16* x12, const bf16* g,
                      bf16* scratch, bf16* dx_out, std::vector<PendingW>& pend) {
  const int H = e->L.cfg.hidden;
  pend.push_back({dout, H, g, H, &s.fc3, 0, H, false});
  KIT_TRY(linear_dgrad(e, dout, H, s.fc3, 0, H, scratch, H, nullptr, 0));
  e->launches++;
  KIT_TRY(swiglu_gate_bwd(scratch, x12, e->g2h, e->M, H, e->st));
  pend.push_back({e->g2h, 2 * H, x, H, &s.fc12, 0, 2 * H, false});
  return linear_dgrad(e, e->g2h, 2 * H, s.fc12, 0, 2 * H, dx_out, H, nullptr, 0);
}

// Scope during which no collective can be in flight: the persistent kernels planned / launched inside keep every SM.
struct ReserveOff {
  ReserveOff() { set_sm_reserve_active(false); }
  ~ReserveOff() { set_sm_reserve_active(true); }
  void end() { set_sm_reserve_active(true); }
};

static int engine_forward(KitEngine* e, const float* x_enc, int64_t xes, const float* x_dec, int64_t xds,
                          const KitAttnMask* em, const KitAttnMask* dm, int zero_masked, float* pred) {
  const Layout& L = e->L;
  const int H = L.cfg.hidden, NH = L.cfg.heads, d = H / NH, IN = L.cfg.input_size, FF = L.cfg.ff;
  const int B = e->B, T = e->T;
  const int64_t M = e->M;
  e->active = &e->fwd_plans;
  e->cursor = 0;
  e->ffn_cursor = 0;
  e->launches = 0;
  prof_reset(e);
  ReserveOff no_collective_in_flight;   // every all-reduce of the previous step was waited for by its optimiser step
  e->enc_mask = em ? *em : KitAttnMask{};
  e->dec_mask = dm ? *dm : KitAttnMask{};
  if (x_enc != nullptr) {   // else: the pre-pass wrote the bf16 operands straight into xe / xd (kit_engine_operands)
    const float* zm = (zero_masked && em) ? em->frame_mask : nullptr;
    KIT_REQUIRE(!zero_masked || zm != nullptr, "zero_masked_enc needs enc_mask.frame_mask");
    e->launches += 2;
    KIT_TRY(pack_frames(x_enc, xes, B, T, IN, zm, em ? em->frame_mask_stride : 0, e->xe, e->K2p, e->st));
    KIT_TRY(pack_frames(x_dec, xds, B, T, IN, nullptr, 0, e->xd, e->K2p, e->st));
  }
  // model.py:120-137 (embedding, token norm, PE, SwiGLU) for both branches
  KIT_TRY(eg(e, 0, e->xe, e->K2p, e->wb + L.emb_i.wb, L.emb_i.ld, e->ei_raw, H, (int)M, H, e->K2p, e->params + L.emb_i.b,
             nullptr, 0, OUT_BF16, ACT_NONE, nullptr, 0));
  KIT_TRY(eg(e, 0, e->xd, e->K2p, e->wb + L.emb_f.wb, L.emb_f.ld, e->ef_raw, H, (int)M, H, e->K2p, e->params + L.emb_f.b,
             nullptr, 0, OUT_BF16, ACT_NONE, nullptr, 0));
  e->launches += 2;
  const float ns = L.cfg.variant == KIT_MODEL_CYCLE ? 2.f : 1.f;   // model.py:283-284
  KIT_TRY(embed_post_fwd(e->ei_raw, e->params + L.pe_i, e->params + L.learned_i, e->ei, M, H, T, ns, e->st));
  KIT_TRY(embed_post_fwd(e->ef_raw, e->params + L.pe_f, e->params + L.learned_f, e->ef, M, H, T, ns, e->st));
  KIT_TRY(swiglu_fwd(e, e->ei, L.swi_i, e->si12, e->sig, e->x0));
  KIT_TRY(swiglu_fwd(e, e->ef, L.swi_f, e->sf12, e->sfg, e->y0));
  // encoder (torch/nn/modules/transformer.py:956 post-norm layers + final norm :137)
  const bf16* x = e->x0;
  for (int l = 0; l < L.cfg.layers; ++l) {
    const EncW& w = L.enc[l];
    EncAct& a = e->ea[l];
    KIT_TRY(linear_fwd(e, x, H, w.sa.in, 0, 3 * H, a.qkv, 3 * H, nullptr, 0));
    KIT_TRY(eattn_fwd(e, a.qkv, 3 * H, a.qkv + H, 3 * H, a.qkv + 2 * H, 3 * H, a.ao, H, a.lse, &e->enc_mask));
    KIT_TRY(linear_add_ln_fwd(e, a.ao, H, w.sa.out, a.s1, x, w.n1, a.x1, a.st1));
    KIT_TRY(ffn_block_fwd(e, a.x1, w.l1, w.l2, a.z, a.hh, a.s2, w.n2, a.x2, a.st2));
    x = a.x2;
  }
  e->launches++;
  KIT_TRY(add_ln_fwd(x, nullptr, e->params + L.enc_norm.g, e->params + L.enc_norm.b, nullptr, e->mem, e->st_encn,
                     e->st_encn + M, M, H, e->st));
  // cross-attention keys / values of all decoder layers: kv_all = mem Wkv_cat^T + b (N = layers * 2H)
  KIT_TRY(eg(e, 0, e->mem, H, e->wb + L.kvcat_wb, H, e->kv_all, L.kvcat_n, (int)M, L.kvcat_n, H, e->kv_bias, nullptr, 0, OUT_BF16,
             ACT_NONE, nullptr, 0));
  // decoder (transformer.py:1147-1153 + final norm :163); cross-attention is unmasked
  const bf16* y = e->y0;
  for (int l = 0; l < L.cfg.layers; ++l) {
    const DecW& w = L.dec[l];
    DecAct& a = e->da[l];
    KIT_TRY(linear_fwd(e, y, H, w.sa.in, 0, 3 * H, a.qkv, 3 * H, nullptr, 0));
    KIT_TRY(eattn_fwd(e, a.qkv, 3 * H, a.qkv + H, 3 * H, a.qkv + 2 * H, 3 * H, a.ao, H, a.lse, &e->dec_mask));
    KIT_TRY(linear_add_ln_fwd(e, a.ao, H, w.sa.out, a.s1, y, w.n1, a.y1, a.st1));
    KIT_TRY(linear_fwd(e, a.y1, H, w.ca.in, 0, H, a.qc, H, nullptr, 0));
    KIT_TRY(eattn_fwd(e, a.qc, H, a.kvc, L.kvcat_n, a.kvc + H, L.kvcat_n, a.aoc, H, a.lsec, nullptr));
    KIT_TRY(linear_add_ln_fwd(e, a.aoc, H, w.ca.out, a.s2, a.y1, w.n2, a.y2, a.st2));
    KIT_TRY(ffn_block_fwd(e, a.y2, w.l1, w.l2, a.z, a.hh, a.s3, w.n3, a.y3, a.st3));
    y = a.y3;
  }
  e->launches++;
  KIT_TRY(add_ln_fwd(y, nullptr, e->params + L.dec_norm.g, e->params + L.dec_norm.b, nullptr, e->dec_out, e->st_decn,
                     e->st_decn + M, M, H, e->st));
  // model.py:147-155 output head
  KIT_TRY(swiglu_fwd(e, e->dec_out, L.swi_d, e->sd12, e->sdg, e->sd));
  e->launches++;
  KIT_TRY(final_norm_silu_fwd(e->sd, e->ef_raw, e->zf, e->sf, M, H, e->st));
  KIT_TRY(eg(e, 0, e->sf, H, e->wb + L.fc_final.wb, L.fc_final.ld, pred, IN, (int)M, IN, H, e->params + L.fc_final.b, nullptr,
             0, OUT_F32, ACT_NONE, nullptr, 0));
  return KIT_OK;
}

static int engine_backward(KitEngine* e, const float* dpred, KitBucketCallback cb, void* user) {
  const Layout& L = e->L;
  const int H = L.cfg.hidden, NH = L.cfg.heads, d = H / NH, IN = L.cfg.input_size, FF = L.cfg.ff;
  const int B = e->B, T = e->T;
  const int64_t M = e->M;
  const int nl = L.cfg.layers;
  e->active = &e->bwd_plans;
  e->cursor = 0;
  e->group_cursor = 0;
  e->ffn_bwd_cursor = 0;
  e->launches = 0;
  std::vector<PendingW> pend;
  int bucket = 0;
  ReserveOff before_first_bucket;   // the first all-reduce starts at the first bucket callback: until then the kernels keep every SM
  auto done = [&]() {
    if (cb) cb(bucket, user);
    before_first_bucket.end();
    ++bucket;
  };
  // ---- output head
  e->launches++;
  KIT_TRY(cast_pad(dpred, M, IN, IN, e->dp, e->K2p, e->st));
  pend.push_back({e->dp, e->K2p, e->sf, H, &L.fc_final, 0, IN, false});
  KIT_TRY(eg(e, 0, e->dp, e->K2p, e->wb + L.fc_final.wbT, L.fc_final.ldT, e->g0, H, (int)M, H, e->K2p, nullptr, nullptr, 0,
             OUT_BF16, ACT_NONE, nullptr, 0));
  e->launches++;
  KIT_TRY(final_norm_silu_bwd(e->g0, e->zf, e->g3, M, H, e->st));  // g3 = d(zf): kept for the embedding residual
  KIT_TRY(swiglu_bwd(e, e->g3, e->dec_out, L.swi_d, e->sd12, e->sdg, e->g0, e->g1, pend));  // g1 = d dec_out
  KIT_TRY(flush_wgrads(e, pend));   // fc_final + the head's SwiGLU: dp, g3 and g2h are intact
  e->launches++;
  KIT_TRY(ln_bwd(e->g1, e->da[nl - 1].y3, e->st_decn, e->st_decn + M, e->params + L.dec_norm.g, nullptr, e->g0,
                 e->grads + L.dec_norm.g, e->grads + L.dec_norm.b, nullptr, M, H, e->st));
  bf16* dy = e->g0;  // gradient w.r.t. the current layer's output
  bool top_done = false;   // the LayerNorm backward at the top of this layer already ran in the previous layer's last GEMM (-> g1)
  bool bd = true;
  for (int l = nl - 1; l >= 0; --l) {
    const DecW& w = L.dec[l];
    DecAct& a = e->da[l];
    const bf16* y_in = (l == 0) ? e->y0 : e->da[l - 1].y3;
    // FFN block.  Weight gradients are queued (their dy buffers g1 / gff / g1b / gqc / gkv / g1c / gqkv stay untouched
    // until the end of the layer) and leave in one grouped launch.  Every LayerNorm backward runs in the epilogue of the
    // kernel that produces its output gradient (dgrad_then_lnbwd / ffn_block_bwd) when the row is one 256-wide tile.
    if (!top_done) {
      e->launches++;
      KIT_TRY(ln_bwd(dy, a.s3, a.st3, a.st3 + M, e->params + w.n3.g, nullptr, e->g1, e->grads + w.n3.g, e->grads + w.n3.b, e->grads + w.l2.b, M, H, e->st));
      bd = true;
    }
    pend.push_back({e->g1, H, a.hh, FF, &w.l2, 0, H, bd});
    KIT_TRY(ffn_block_bwd(e, e->g1, w.l1, w.l2, a.z, e->gff, e->g2, e->g1b, a.s2, a.st2, w.n2, e->grads + w.ca.out.b, &bd));  // g1b = d s2
    pend.push_back({e->gff, FF, a.y2, H, &w.l1, 0, FF, false});
    // cross-attention block
    pend.push_back({e->g1b, H, a.aoc, H, &w.ca.out, 0, H, bd});
    KIT_TRY(linear_dgrad(e, e->g1b, H, w.ca.out, 0, H, e->g2, H, nullptr, 0));  // g2 = d aoc
    bf16* gkv = e->gkv + (int64_t)l * 2 * H;   // this layer's columns of the [M, layers * 2H] key / value gradient
    KIT_TRY(eattn_bwd(e, a.qc, H, a.kvc, L.kvcat_n, a.kvc + H, L.kvcat_n, a.aoc, H, e->g2, H, a.lsec, e->gqc, H, gkv, L.kvcat_n,
                      gkv + H, L.kvcat_n, nullptr));
    pend.push_back({e->gqc, H, a.y1, H, &w.ca.in, 0, H, false});
    pend.push_back({gkv, L.kvcat_n, e->mem, H, &w.ca.in, H, 2 * H, false});
    // d y1 = gqc Wq + g1b, then the backward of norm1 -> g1c = d s1
    KIT_TRY(dgrad_then_lnbwd(e, e->gqc, H, w.ca.in, 0, H, e->g1b, e->g2, e->g1c, a.s1, a.st1, w.n1, e->grads + w.sa.out.b, &bd));
    // self-attention block
    pend.push_back({e->g1c, H, a.ao, H, &w.sa.out, 0, H, bd});
    KIT_TRY(linear_dgrad(e, e->g1c, H, w.sa.out, 0, H, e->g2, H, nullptr, 0));  // g2 = d ao
    KIT_TRY(eattn_bwd(e, a.qkv, 3 * H, a.qkv + H, 3 * H, a.qkv + 2 * H, 3 * H, a.ao, H, e->g2, H, a.lse, e->gqkv, 3 * H,
                      e->gqkv + H, 3 * H, e->gqkv + 2 * H, 3 * H, &e->dec_mask));
    pend.push_back({e->gqkv, 3 * H, y_in, H, &w.sa.in, 0, 3 * H, false});
    KIT_TRY(flush_wgrads(e, pend));   // before the last GEMM: fused with the next layer's norm3 backward it overwrites g1
    if (l > 0 && lnbwd_fusable(e) && gemm_lnbwd_supported((int)M)) {   // d y_in = gqkv Win + g1c, then the backward of the layer below's norm3 -> g1
      const DecW& wn = L.dec[l - 1];
      DecAct& an = e->da[l - 1];
      KIT_TRY(dgrad_then_lnbwd(e, e->gqkv, 3 * H, w.sa.in, 0, 3 * H, e->g1c, e->g0, e->g1, an.s3, an.st3, wn.n3, e->grads + wn.l2.b, &bd));
      top_done = true;
    } else {
      KIT_TRY(linear_dgrad(e, e->gqkv, 3 * H, w.sa.in, 0, 3 * H, e->g0, H, e->g1c, H));  // g0 = d y_in
      top_done = false;
    }
    dy = e->g0;
    if (L.dec_cut[l]) done();
  }
  done();  // decoder finished
  // filled branch: SwiGLU, token-norm/PE, embedding  (g3 = residual gradient from the head)
  KIT_TRY(swiglu_bwd(e, dy, e->ef, L.swi_f, e->sf12, e->sfg, e->g1, e->g2, pend));  // g2 = d ef
  KIT_TRY(flush_wgrads(e, pend));   // (dy = g0 and g2h are intact)
  e->launches++;
  const float ns = L.cfg.variant == KIT_MODEL_CYCLE ? 2.f : 1.f;
  KIT_TRY(embed_post_bwd(e->g2, e->ef_raw, e->g3, e->g1, e->grads + L.learned_f, M, H, ns, e->st));  // g1 = d ef_raw
  KIT_TRY(eg(e, 1, e->g1, H, e->xd, e->K2p, e->grads + L.emb_f.w, IN, H, IN, (int)M, nullptr, nullptr, 0, OUT_F32_ATOMIC,
             ACT_NONE, nullptr, 0));
  e->launches++;
  KIT_TRY(colsum(e->g1, H, e->grads + L.emb_f.b, M, H, e->st));
  // ---- encoder.  d mem = gkv Wkv_cat: the key / value gradients of all decoder layers in one GEMM with K = layers * 2H
  KIT_TRY(eg(e, 0, e->gkv, L.kvcat_n, e->wb + L.kvcat_wbT, L.kvcat_n, e->gmem, H, (int)M, H, L.kvcat_n, nullptr, nullptr, 0, OUT_BF16,
             ACT_NONE, nullptr, 0));
  e->launches++;
  KIT_TRY(ln_bwd(e->gmem, e->ea[nl - 1].x2, e->st_encn, e->st_encn + M, e->params + L.enc_norm.g, nullptr, e->g0,
                 e->grads + L.enc_norm.g, e->grads + L.enc_norm.b, nullptr, M, H, e->st));
  bf16* dx = e->g0;
  top_done = false;
  for (int l = nl - 1; l >= 0; --l) {
    const EncW& w = L.enc[l];
    EncAct& a = e->ea[l];
    const bf16* x_in = (l == 0) ? e->x0 : e->ea[l - 1].x2;
    if (!top_done) {
      e->launches++;
      KIT_TRY(ln_bwd(dx, a.s2, a.st2, a.st2 + M, e->params + w.n2.g, nullptr, e->g1, e->grads + w.n2.g, e->grads + w.n2.b, e->grads + w.l2.b, M, H, e->st));
      bd = true;
    }
    pend.push_back({e->g1, H, a.hh, FF, &w.l2, 0, H, bd});
    KIT_TRY(ffn_block_bwd(e, e->g1, w.l1, w.l2, a.z, e->gff, e->g2, e->g1b, a.s1, a.st1, w.n1, e->grads + w.sa.out.b, &bd));  // g1b = d s1
    pend.push_back({e->gff, FF, a.x1, H, &w.l1, 0, FF, false});
    pend.push_back({e->g1b, H, a.ao, H, &w.sa.out, 0, H, bd});
    KIT_TRY(linear_dgrad(e, e->g1b, H, w.sa.out, 0, H, e->g2, H, nullptr, 0));
    KIT_TRY(eattn_bwd(e, a.qkv, 3 * H, a.qkv + H, 3 * H, a.qkv + 2 * H, 3 * H, a.ao, H, e->g2, H, a.lse, e->gqkv, 3 * H,
                      e->gqkv + H, 3 * H, e->gqkv + 2 * H, 3 * H, &e->enc_mask));
    pend.push_back({e->gqkv, 3 * H, x_in, H, &w.sa.in, 0, 3 * H, false});
    KIT_TRY(flush_wgrads(e, pend));
    if (l > 0 && lnbwd_fusable(e) && gemm_lnbwd_supported((int)M)) {   // d x_in = gqkv Win + g1b, then the backward of the layer below's norm2 -> g1
      const EncW& wn = L.enc[l - 1];
      EncAct& an = e->ea[l - 1];
      KIT_TRY(dgrad_then_lnbwd(e, e->gqkv, 3 * H, w.sa.in, 0, 3 * H, e->g1b, e->g0, e->g1, an.s2, an.st2, wn.n2, e->grads + wn.l2.b, &bd));
      top_done = true;
    } else {
      KIT_TRY(linear_dgrad(e, e->gqkv, 3 * H, w.sa.in, 0, 3 * H, e->g0, H, e->g1b, H));
      top_done = false;
    }
    dx = e->g0;
    if (L.enc_cut[l]) done();
  }
  // input branch
  KIT_TRY(swiglu_bwd(e, dx, e->ei, L.swi_i, e->si12, e->sig, e->g1, e->g2, pend));
  KIT_TRY(flush_wgrads(e, pend));
  e->launches++;
  KIT_TRY(embed_post_bwd(e->g2, e->ei_raw, nullptr, e->g1, e->grads + L.learned_i, M, H, ns, e->st));
  KIT_TRY(eg(e, 1, e->g1, H, e->xe, e->K2p, e->grads + L.emb_i.w, IN, H, IN, (int)M, nullptr, nullptr, 0, OUT_F32_ATOMIC,
             ACT_NONE, nullptr, 0));
  e->launches++;
  KIT_TRY(colsum(e->g1, H, e->grads + L.emb_i.b, M, H, e->st));
  done();
  return KIT_OK;
}

// kv_bias[l * n + i] = params[base + l * stride + i]: the in_proj_bias[H:3H] slices of the decoder layers' cross-attention
__global__ void gather_kv_bias_kernel(const float* __restrict__ params, int64_t base, int64_t stride, int n, int layers,
                                      float* __restrict__ out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * layers) out[i] = params[base + (int64_t)(i / n) * stride + (i % n)];
}

__global__ void bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  pdl_grid_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}

}  // namespace kit

// ---------------------------------------------------------------- C ABI
extern "C" int32_t kit_layout_num_entries(const KitModelConfig* cfg) {
  Layout L;
  if (build_layout(cfg, L)) return -1;
  return (int32_t)L.entries.size();
}
extern "C" int kit_layout_entry(const KitModelConfig* cfg, int32_t index, char* name, int32_t name_cap, int64_t* offset,
                                int64_t* numel, int64_t* rows, int64_t* cols, int32_t* is_buffer) {
  Layout L;
  KIT_TRY(build_layout(cfg, L));
  KIT_REQUIRE(index >= 0 && index < (int)L.entries.size(), "layout entry %d out of range", index);
  const Entry& en = L.entries[index];
  if (name != nullptr && name_cap > 0) snprintf(name, name_cap, "%s", en.name.c_str());
  if (offset) *offset = en.offset;
  if (numel) *numel = en.numel;
  if (rows) *rows = en.rows;
  if (cols) *cols = en.cols;
  if (is_buffer) *is_buffer = en.is_buffer;
  return KIT_OK;
}
extern "C" int64_t kit_layout_trainable_floats(const KitModelConfig* cfg) {
  Layout L;
  if (build_layout(cfg, L)) return -1;
  return L.trainable;
}
extern "C" int64_t kit_layout_total_floats(const KitModelConfig* cfg) {
  Layout L;
  if (build_layout(cfg, L)) return -1;
  return L.total;
}
extern "C" int32_t kit_layout_num_buckets(const KitModelConfig* cfg) {
  Layout L;
  if (build_layout(cfg, L)) return -1;
  return (int32_t)L.buckets.size();
}
extern "C" int kit_layout_bucket(const KitModelConfig* cfg, int32_t bucket, int64_t* begin, int64_t* end) {
  Layout L;
  KIT_TRY(build_layout(cfg, L));
  KIT_REQUIRE(bucket >= 0 && bucket < (int)L.buckets.size(), "bucket %d out of range", bucket);
  *begin = L.buckets[bucket].first;
  *end = L.buckets[bucket].second;
  return KIT_OK;
}

extern "C" int kit_engine_create(const KitModelConfig* cfg, int32_t batch, int32_t seq_len, int32_t training,
                                 KitEngine** out) {
  KIT_REQUIRE(out != nullptr, "kit_engine_create: out is null");
  KIT_REQUIRE(batch > 0 && seq_len > 0, "kit_engine_create: batch and seq_len must be positive");
  KitEngine* e = new KitEngine();
  int rc = build_layout(cfg, e->L);
  if (rc) {
    delete e;
    return rc;
  }
  if (seq_len > cfg->max_len) {
    delete e;
    set_error("seq_len %d exceeds the positional table (%d rows, model.py:74-75)", seq_len, cfg->max_len);
    return KIT_ERR_INVALID;
  }
  e->B = batch;
  e->T = seq_len;
  e->training = training ? 1 : 0;
  e->M = (int64_t)batch * seq_len;
  e->K2p = (int)up8(cfg->input_size);
  {
    const char* v = getenv("KIT_FUSE_FFN");   // KIT_FUSE_FFN=0: the two-GEMM path (A/B measurements)
    e->fuse_ffn = !(v != nullptr && v[0] == '0');
    v = getenv("KIT_FUSE_LNBWD");
    e->fuse_lnbwd = v != nullptr && v[0] == '1';
  }
  plan_workspace(e);
  *out = e;
  return KIT_OK;
}
extern "C" int kit_engine_destroy(KitEngine* e) {
  if (e != nullptr) {
    for (auto& ev : e->ev_pool) {   // profiling events (kit_engine_set_profiling)
      cudaEventDestroy(ev.first);
      cudaEventDestroy(ev.second);
    }
    e->ev_pool.clear();
  }
  delete e;
  return KIT_OK;
}
extern "C" int64_t kit_engine_workspace_bytes(const KitEngine* e) { return e ? e->ws_bytes : -1; }

extern "C" int kit_engine_bind(KitEngine* e, float* params, float* grads, void* workspace, int64_t workspace_bytes) {
  KIT_REQUIRE(e && params && workspace, "kit_engine_bind: null argument");
  KIT_REQUIRE(workspace_bytes >= e->ws_bytes, "workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)e->ws_bytes);
  KIT_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)params & 31) == 0 && ((uintptr_t)grads & 31) == 0,
              "kit_engine_bind: workspace must be 256-byte aligned, arenas 32-byte aligned");
  e->params = params;
  e->grads = grads;
  e->ws = (uint8_t*)workspace;
  resolve_pointers(e);
  e->fwd_plans.clear();
  e->bwd_plans.clear();
  e->group_plans.clear();
  e->ffn_plans.clear();
  e->ffn_bwd_plans.clear();

  // upload the weight-refresh tables (synchronous, bind time only)
  std::vector<int> prefix(e->L.wdescs.size() + 1, 0);
  for (size_t i = 0; i < e->L.wdescs.size(); ++i) {
    const WeightDesc& d = e->L.wdescs[i];
    prefix[i + 1] = prefix[i] + ((d.rows + WR_TILE - 1) / WR_TILE) * ((d.cols_pad + WR_TILE - 1) / WR_TILE);
  }
  e->total_tiles = prefix.back();
  KIT_CHECK_CUDA(cudaMemcpy(e->wdesc_dev, e->L.wdescs.data(), e->L.wdescs.size() * sizeof(WeightDesc), cudaMemcpyHostToDevice));
  KIT_CHECK_CUDA(cudaMemcpy(e->tile_prefix_dev, prefix.data(), prefix.size() * sizeof(int), cudaMemcpyHostToDevice));
  e->bound = true;
  e->weights_fresh = false;
  return KIT_OK;
}

extern "C" int kit_engine_refresh_weights(KitEngine* e, void* stream) {
  KIT_REQUIRE(e && e->bound, "kit_engine_refresh_weights: engine not bound");
  KIT_TRY(weight_refresh(e->params, e->wb, e->wdesc_dev, e->tile_prefix_dev, (int)e->L.wdescs.size(), e->total_tiles,
                         (cudaStream_t)stream));
  {
    const Layout& L = e->L;
    const int H = L.cfg.hidden, nl = L.cfg.layers;
    const int64_t base = L.dec[0].ca.in.b + H, stride = nl > 1 ? L.dec[1].ca.in.b - L.dec[0].ca.in.b : 0;
    launch_kernel(gather_kv_bias_kernel, dim3((unsigned)ceil_div((int64_t)L.kvcat_n, 256)), dim3(256), 0, (cudaStream_t)stream,
                  (const float*)e->params, base, stride, 2 * H, nl, e->kv_bias);
    KIT_LAUNCH_CHECK();
  }
  e->weights_fresh = true;
  return KIT_OK;
}

extern "C" int kit_engine_forward(KitEngine* e, const float* x_enc, int64_t x_enc_batch_stride, const float* x_dec,
                                  int64_t x_dec_batch_stride, const KitAttnMask* enc_mask, const KitAttnMask* dec_mask,
                                  int32_t zero_masked_enc, float* pred, void* stream) {
  KIT_REQUIRE(e && e->bound, "kit_engine_forward: engine not bound");
  KIT_REQUIRE(e->weights_fresh, "kit_engine_forward: call kit_engine_refresh_weights after binding / updating parameters");
  KIT_REQUIRE(pred != nullptr && (x_enc == nullptr) == (x_dec == nullptr), "kit_engine_forward: pred is required; x_enc / x_dec are both given or both NULL");
  e->st = (cudaStream_t)stream;
  return engine_forward(e, x_enc, x_enc_batch_stride, x_dec, x_dec_batch_stride, enc_mask, dec_mask, zero_masked_enc, pred);
}

extern "C" int kit_engine_backward(KitEngine* e, const float* dpred, KitBucketCallback bucket_done, void* user, void* stream) {
  KIT_REQUIRE(e && e->bound && e->grads, "kit_engine_backward: engine not bound to a gradient arena");
  KIT_REQUIRE(e->training, "kit_engine_backward: engine was created for inference (training = 0)");
  KIT_REQUIRE(!e->fwd_plans.empty(), "kit_engine_backward: run kit_engine_forward first");
  KIT_REQUIRE(dpred != nullptr, "kit_engine_backward: dpred is null");
  e->st = (cudaStream_t)stream;
  return engine_backward(e, dpred, bucket_done, user);
}

extern "C" int kit_engine_operands(KitEngine* e, void** x_enc_bf16, void** x_dec_bf16, int32_t* k2p) {
  KIT_REQUIRE(e && e->bound, "kit_engine_operands: engine not bound");
  if (x_enc_bf16) *x_enc_bf16 = e->xe;
  if (x_dec_bf16) *x_dec_bf16 = e->xd;
  if (k2p) *k2p = e->K2p;
  return KIT_OK;
}

extern "C" int kit_engine_debug_read(KitEngine* e, const char* name, float* out, int64_t out_floats, void* stream) {
  KIT_REQUIRE(e && e->bound && name && out, "kit_engine_debug_read: bad arguments");
  auto it = e->bufs.find(name);
  KIT_REQUIRE(it != e->bufs.end(), "kit_engine_debug_read: no buffer named '%s'", name);
  const Buf& b = it->second;
  KIT_REQUIRE(out_floats >= b.elems, "kit_engine_debug_read: output too small (%lld < %lld)", (long long)out_floats, (long long)b.elems);
  if (b.esize == 4) {
    KIT_CHECK_CUDA(cudaMemcpyAsync(out, e->ws + b.off, b.elems * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  } else {
    KIT_REQUIRE(b.esize == 2, "kit_engine_debug_read: buffer '%s' is not a tensor", name);
    launch_kernel(bf16_to_f32_kernel, dim3((unsigned)ceil_div(b.elems, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)(e->ws + b.off), out, b.elems);
    KIT_LAUNCH_CHECK();
  }
  return KIT_OK;
}
extern "C" int64_t kit_engine_last_launches(const KitEngine* e) { return e ? e->launches : -1; }

extern "C" int kit_engine_set_profiling(KitEngine* e, int32_t on) {
  KIT_REQUIRE(e != nullptr, "kit_engine_set_profiling: null engine");
  e->profiling = on != 0;
  return KIT_OK;
}
extern "C" int kit_engine_profile_read(KitEngine* e, int32_t category, float* ms, int64_t* launches, double* flops) {
  KIT_REQUIRE(e != nullptr && category >= 0 && category < KIT_PROF_CATEGORIES, "kit_engine_profile_read: bad arguments");
  float total = 0.f;
  for (const auto& r : e->prof) {
    if (r.cat != category) continue;
    KIT_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    KIT_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    total += t;
  }
  if (ms) *ms = total;
  if (launches) *launches = e->cat_launches[category];
  if (flops) *flops = e->cat_flops[category];
  return KIT_OK;
}

// The fused feed-forward block on its own (tests): x [M,256] bf16, w1 [FF,256], w2 [256,FF] bf16 row-major,
// s / y [M,256] bf16, z / hh [M,FF] bf16 (written when store_zh), mean / rstd [M] fp32.
extern "C" int kit_ffn_fwd(const void* x, const void* w1, const void* w2, const float* b1, const float* b2, const float* gamma,
                           const float* beta, void* z, void* hh, void* s, void* y, float* mean, float* rstd, int32_t M,
                           int32_t H, int32_t FF, int32_t store_zh, void* stream) {
  FfnPlan plan;
  KIT_TRY(ffn_fwd_plan(&plan, (const bf16*)x, H, (const bf16*)w1, H, (const bf16*)w2, FF, b1, b2, (bf16*)z, (bf16*)hh, FF, (bf16*)s,
                       H, (bf16*)y, H, gamma, beta, mean, rstd, 1e-5f, M, H, FF, store_zh));
  return ffn_launch(&plan, (cudaStream_t)stream);
}

// Its input-gradient pass: dz = (g w2) * gelu'(z) -> dz [M,FF], dx = dz w1 + g -> dx [M,H].  w2t = w2^T [FF,H], w1t = w1^T [H,FF].
extern "C" int kit_ffn_bwd(const void* g, const void* w2t, const void* w1t, const void* z, void* dz, void* dx, int32_t M, int32_t H,
                           int32_t FF, void* stream) {
  FfnPlan plan;
  KIT_TRY(ffn_bwd_plan(&plan, (const bf16*)g, H, (const bf16*)w2t, H, (const bf16*)w1t, FF, (const bf16*)z, (bf16*)dz, FF, (bf16*)dx, H,
                       M, H, FF));
  return ffn_launch(&plan, (cudaStream_t)stream);
}

// Tests: dx = LayerNormBackward(dy = A B^T + addend; saved sum s, mean, rstd, gamma) in the GEMM epilogue (N must be 256).
extern "C" int kit_gemm_lnbwd(const void* A, const void* B, const void* addend, const void* s, const float* gamma, const float* mean,
                              const float* rstd, void* dx, float* dgamma, float* dbeta, int32_t M, int32_t K, void* stream) {
  GemmPlan p;
  LnBwdArgs lnb{(const bf16*)s, 256, gamma, mean, rstd, dgamma, dbeta};
  KIT_TRY(gemm_plan(&p, 0, (const bf16*)A, K, (const bf16*)B, K, dx, 256, M, 256, K, nullptr, (const bf16*)addend, 256, OUT_BF16,
                    ACT_NONE, nullptr, 0, 1, nullptr, nullptr, &lnb));
  KIT_REQUIRE(p.epi == EPI_ADD_LNBWD, "kit_gemm_lnbwd: the planner did not select the fused epilogue");
  return gemm_launch(&p, (cudaStream_t)stream);
}

extern "C" int kit_gemm_bf16(int32_t mode, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                             int32_t M, int32_t N, int32_t K, const float* bias, const void* addend, int64_t ld_addend,
                             int32_t out_kind, int32_t act, void* aux, int64_t ld_aux, int32_t split_k, void* stream) {
  GemmPlan p;
  KIT_TRY(gemm_plan(&p, mode, (const bf16*)A, lda, (const bf16*)B, ldb, C, ldc, M, N, K, bias, (const bf16*)addend, ld_addend,
                    out_kind, act, (bf16*)aux, ld_aux, split_k));
  return gemm_launch(&p, (cudaStream_t)stream);
}
