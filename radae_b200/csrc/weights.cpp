// Weight ingest for libradae_b200: RDW container (this repo's format, radae_b200/rdw.py) or the reference's DNNw
// blob (src/write_rade_weights.c:51-74; int8 matrices in 8x4 blocks, weight-exchange/wexchange/c_export/common.py:59-67,
// optional block index lists :156-170) -> host row-major matrices -> device layouts.
// Device layout: ONE byte stream per codec holding every weight of a 40 ms step as a sequence of <= 32 KB chunks in the
// order the kernel's layer walk consumes them (core_codec.cu); the kernel's producer warp TMA-bulk-copies chunk after
// chunk into a shared-memory ring.  int8 layers are stored in MMA B-fragment order [kb][nt][lane] = {b0, b1} with
//   b0 = W[nt*8 + lane/4][kb*32 + (lane%4)*4 .. +3],  b1 = same row, columns +16   (m16n8k32 .col B operand);
// float layers as the rows [j][out] of the concat segment a chunk covers.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "rade_common.h"
#include "rade_host.h"
#include "umma_layout.h"

namespace {

struct HostArray { int dtype; int rows, cols; std::vector<unsigned char> data; };
typedef std::map<std::string, HostArray> ArrayMap;

bool parse_rdw(const unsigned char *buf, size_t len, ArrayMap &out) {
  if (len < 64 || memcmp(buf, "RADEB200", 8) != 0) return false;
  uint32_t version, n;
  memcpy(&version, buf + 8, 4); memcpy(&n, buf + 12, 4);
  if (version != 1 || 64 + 80 * (size_t)n > len) return false;
  for (uint32_t i = 0; i < n; i++) {
    const unsigned char *e = buf + 64 + 80 * (size_t)i;
    char name[49]; memcpy(name, e, 48); name[48] = 0;
    uint32_t dtype, rows, cols; uint64_t off, nbytes;
    memcpy(&dtype, e + 48, 4); memcpy(&rows, e + 52, 4); memcpy(&cols, e + 56, 4);
    memcpy(&off, e + 64, 8); memcpy(&nbytes, e + 72, 8);
    if (off + nbytes > len) return false;
    HostArray a; a.dtype = (int)dtype; a.rows = (int)rows; a.cols = (int)cols;
    a.data.assign(buf + off, buf + off + nbytes);
    out[name] = a;
  }
  return true;
}

struct LayerSpec { const char *name; int nin, nout; int kind; };   // kind 0 f32, 1 int8 dense, 2 int8 block-indexed

void layer_specs(std::vector<LayerSpec> &v, std::vector<std::string> &names) {
  static const int enc_k[5] = {64, 224, 384, 544, 704}, enc_ck[5] = {256, 576, 896, 1216, 1536};
  static const int dec_k[5] = {96, 224, 352, 480, 608}, dec_ck[5] = {384, 640, 896, 1152, 1408};
  names.reserve(64);
  auto add = [&](const std::string &n, int a, int b, int k) { names.push_back(n); v.push_back({nullptr, a, b, k}); };
  add("enc_dense1", 84, 64, 0); add("enc_zdense", 864, 80, 0); add("dec_dense1", 80, 96, 0); add("dec_output", 736, 84, 0);
  for (int i = 0; i < 5; i++) {
    std::string n = std::to_string(i + 1);
    add("enc_gru" + n + "_input", enc_k[i], 192, 2); add("enc_gru" + n + "_recurrent", 64, 192, 1); add("enc_conv" + n, enc_ck[i], 96, 1);
    add("dec_gru" + n + "_input", dec_k[i], 288, 2); add("dec_gru" + n + "_recurrent", 96, 288, 1);
    add("dec_glu" + n, 96, 96, 1); add("dec_conv" + n, dec_ck[i], 32, 1);
  }
  for (size_t i = 0; i < v.size(); i++) v[i].name = names[i].c_str();
}

// DNNw blob -> the same flat map RDW gives (unblocking the 8x4 tiles)
bool parse_dnnw(const unsigned char *buf, size_t len, ArrayMap &out) {
  struct Raw { int type; const unsigned char *p; int size; };
  std::map<std::string, Raw> raw;
  size_t off = 0;
  while (off + 64 <= len) {
    if (memcmp(buf + off, "DNNw", 4) != 0) return false;
    int version, type, size, block;
    memcpy(&version, buf + off + 4, 4); memcpy(&type, buf + off + 8, 4);
    memcpy(&size, buf + off + 12, 4); memcpy(&block, buf + off + 16, 4);
    if (version != 0 || size <= 0 || block < size || off + 64 + (size_t)block > len) return false;
    char name[45]; memcpy(name, buf + off + 20, 44); name[44] = 0;
    raw[name] = {type, buf + off + 64, size};
    off += 64 + (size_t)block;
  }
  std::vector<LayerSpec> specs; std::vector<std::string> names;
  layer_specs(specs, names);
  for (auto &L : specs) {
    std::string n = L.name;
    auto need = [&](const std::string &k, size_t bytes) -> const unsigned char * {
      auto it = raw.find(k);
      return (it != raw.end() && (size_t)it->second.size == bytes) ? it->second.p : nullptr;
    };
    // models without the auxiliary symbol (model05, src/test_rade_enc.c:40): enc_dense1 has 80 inputs, dec_output 80 outputs;
    // they are ingested as they are and widened with zeros by normalise_io_width() below
    if (n == "enc_dense1" && !need(n + "_weights_float", 4 * (size_t)L.nin * L.nout)) L.nin = 80;
    if (n == "dec_output" && !need(n + "_bias", 4 * (size_t)L.nout)) L.nout = 80;
    const unsigned char *b = need(n + "_bias", 4 * (size_t)L.nout);
    if (!b) return false;
    HostArray bias; bias.dtype = 0; bias.rows = 1; bias.cols = L.nout; bias.data.assign(b, b + 4 * (size_t)L.nout);
    out[n + ".bias"] = bias;
    if (L.kind == 0) {
      const unsigned char *w = need(n + "_weights_float", 4 * (size_t)L.nin * L.nout);
      if (!w) return false;
      HostArray a; a.dtype = 0; a.rows = L.nin; a.cols = L.nout; a.data.assign(w, w + 4 * (size_t)L.nin * L.nout);
      out[n + ".wf"] = a;
    } else {
      const unsigned char *sc = need(n + "_scale", 4 * (size_t)L.nout);
      auto itw = raw.find(n + "_weights_int8");
      if (!sc || itw == raw.end()) return false;
      const int8_t *wb = (const int8_t *)itw->second.p;
      const int *idx = nullptr;
      if (L.kind == 2) { auto ii = raw.find(n + "_weights_idx"); if (ii == raw.end()) return false; idx = (const int *)ii->second.p; }
      HostArray a; a.dtype = 1; a.rows = L.nout; a.cols = L.nin; a.data.assign((size_t)L.nin * L.nout, 0);
      size_t p = 0;
      for (int ob = 0; ob < L.nout / 8; ob++) {
        int nblk = idx ? *idx++ : L.nin / 4;
        for (int j = 0; j < nblk; j++) {
          int pos = idx ? *idx++ : 4 * j;
          if (pos < 0 || pos + 3 >= L.nin || p + 32 > (size_t)itw->second.size) return false;
          for (int k = 0; k < 8; k++) for (int c = 0; c < 4; c++)
            a.data[(size_t)(ob * 8 + k) * L.nin + pos + c] = (unsigned char)wb[p + 4 * k + c];
          p += 32;
        }
      }
      out[n + ".w8"] = a;
      HostArray s; s.dtype = 0; s.rows = 1; s.cols = L.nout; s.data.assign(sc, sc + 4 * (size_t)L.nout);
      out[n + ".scale"] = s;
    }
  }
  return true;
}

// enc_dense1.wf [80][64] -> [84][64] (zero rows), dec_output.wf [736][80] -> [736][84], dec_output.bias [80] -> [84] (zeros):
// with zero weights the 84-wide kernels compute exactly what the 80-wide layers would (acc + 0 * x == acc)
bool normalise_io_width(ArrayMap &arrays, int *input_dim, int *output_dim) {
  *input_dim = 84; *output_dim = 84;
  auto d1 = arrays.find("enc_dense1.wf");
  if (d1 != arrays.end() && d1->second.rows == 80 && d1->second.cols == 64) {
    d1->second.data.resize((size_t)84 * 64 * 4, 0); d1->second.rows = 84; *input_dim = 80;
  }
  auto ow = arrays.find("dec_output.wf"); auto ob = arrays.find("dec_output.bias");
  if (ow != arrays.end() && ob != arrays.end() && ow->second.rows == 736 && ow->second.cols == 80 && ob->second.cols == 80) {
    std::vector<unsigned char> wide((size_t)736 * 84 * 4, 0);
    for (int r = 0; r < 736; r++) memcpy(&wide[(size_t)r * 84 * 4], &ow->second.data[(size_t)r * 80 * 4], 80 * 4);
    ow->second.data.swap(wide); ow->second.cols = 84;
    ob->second.data.resize(84 * 4, 0); ob->second.cols = 84;
    *output_dim = 80;
  }
  return true;
}

}  // namespace

void core_weights_free(CoreWeightsHolder *h) {
  if (!h) return;
  for (void *p : h->allocs) cudaFree(p);
  h->allocs.clear();
}

namespace {

struct StreamBuilder {
  std::vector<unsigned char> bytes;
  std::vector<ChunkDesc> chunks;
  bool ok = true;
  bool umma = false;       // int8 chunks in the tcgen05 operand layout instead of the mma.sync fragment order
  // tile_step / n_tiles: where the layer's M = 128 tiles start in the tcgen05 formulation (GRU: 0, units, 2 units; others: one tile)
  void add_layer_i8(const int8_t *W8, int N, int K, int kb_lo, int kb_hi, int tile_step, int n_tiles) {
    if (umma) add_i8_umma(W8, N, K, kb_lo, kb_hi, tile_step, n_tiles); else add_i8(W8, N, K, kb_lo, kb_hi);
  }
  void add(const void *p, size_t n) {
    if (n == 0 || n > CORE_STAGE_BYTES || (n & 15)) { ok = false; return; }
    ChunkDesc d; d.offset = (unsigned)bytes.size(); d.bytes = (unsigned)n;
    const unsigned char *b = (const unsigned char *)p;
    bytes.insert(bytes.end(), b, b + n);
    chunks.push_back(d);
  }
  // k-blocks [kb_lo, kb_hi) of an int8 [N][K] matrix in fragment order, chunked kbc k-blocks at a time
  void add_i8(const int8_t *W8, int N, int K, int kb_lo, int kb_hi) {
    const int NTL = N / 8;
    const int kbc = core_kbc(NTL);
    for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += kbc) {
      const int nk = std::min(kbc, kb_hi - kb0);
      std::vector<uint32_t> t((size_t)nk * NTL * 64);
      for (int kb = 0; kb < nk; kb++) for (int nt = 0; nt < NTL; nt++) for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, tig = lane & 3;
        const int8_t *row = W8 + (size_t)(nt * 8 + g) * K + (kb0 + kb) * 32 + tig * 4;
        uint32_t b0, b1; memcpy(&b0, row, 4); memcpy(&b1, row + 16, 4);
        const size_t o = (((size_t)kb * NTL + nt) * 32 + lane) * 2;
        t[o] = b0; t[o + 1] = b1;
      }
      add(t.data(), t.size() * 4);
    }
  }
  // tcgen05 formulation: k-blocks [kb_lo, kb_hi) of ALL rows in the canonical K-major operand layout, as many
  // k-blocks per chunk as a stage can hold for the rows the layer's M = 128 tiles may touch (tiles start at rows 0, tile_step, ...)
  void add_i8_umma(const int8_t *W8, int N, int K, int kb_lo, int kb_hi, int tile_step, int n_tiles) {
    const int nk_max = umma_kblocks_per_stage(umma_span_rows(N, tile_step, n_tiles), CORE_STAGE_BYTES);
    if (nk_max < 1) { ok = false; return; }
    for (int kb0 = kb_lo; kb0 < kb_hi; kb0 += nk_max) {
      std::vector<uint8_t> c = umma_bake_chunk(W8, N, K, kb0, std::min(nk_max, kb_hi - kb0));
      add(c.data(), c.size());
    }
  }
  // rows [j0, j0+nrows) of a float [K][NOUT] matrix, each row zero-padded to NOUTP floats, core_f32_rpc(NOUTP) rows per chunk
  void add_f32_rows(const float *Wf, int NOUT, int NOUTP, int j0, int nrows) {
    const int rpc = core_f32_rpc(NOUTP);
    for (int r0 = 0; r0 < nrows; r0 += rpc) {
      const int n = std::min(rpc, nrows - r0);
      std::vector<float> t((size_t)n * NOUTP, 0.f);
      for (int r = 0; r < n; r++) memcpy(&t[(size_t)r * NOUTP], Wf + (size_t)(j0 + r0 + r) * NOUT, NOUT * sizeof(float));
      add(t.data(), t.size() * 4);
    }
  }
};

// Host-only: the two per-step weight streams from row-major host matrices (p8: int8 [out][in], pf: float [in][out]).
void build_streams(const std::map<std::string, const int8_t *> &p8, const std::map<std::string, const float *> &pf, bool umma,
                   StreamBuilder &e, StreamBuilder &d, int &e_pro, int &d_pro) {
  auto I8 = [&](const std::string &n) { return p8.at(n); };
  auto F = [&](const std::string &n) { return pf.at(n); };
  // Order = the kernels' consumption order.  dense1 is needed once up front (prologue) and then, for the NEXT step, just before
  // the last conv layer, so that the F-warps can have the next step's first activation ready when the I-warps finish this one.
  e.umma = d.umma = umma;
  {
    e.add_f32_rows(F("enc_dense1"), 64, 64, 0, ENC_IN);
    e_pro = (int)e.chunks.size();
    e.add_f32_rows(F("enc_zdense"), RADE_LATENT, RADE_LATENT, 0, 64);
    int off = 64;
    for (int l = 0; l < 5; l++) {
      std::string n = std::to_string(l + 1);
      e.add_layer_i8(I8("enc_gru" + n + "_input"), 192, off, 0, off / 32, ENC_GRU, 3);
      e.add_layer_i8(I8("enc_gru" + n + "_recurrent"), 192, 64, 0, 2, ENC_GRU, 3);
      e.add_f32_rows(F("enc_zdense"), RADE_LATENT, RADE_LATENT, off, ENC_GRU);
      off += ENC_GRU;
      if (l == 4) e.add_f32_rows(F("enc_dense1"), 64, 64, 0, ENC_IN);             // for the next step
      e.add_layer_i8(I8("enc_conv" + n), 96, 2 * off, 0, off / 32, 0, 1);                    // tap 0 (oldest frame)
      e.add_layer_i8(I8("enc_conv" + n), 96, 2 * off, off / 32, 2 * off / 32, 0, 1);         // tap 1 (current frame)
      e.add_f32_rows(F("enc_zdense"), RADE_LATENT, RADE_LATENT, off, ENC_CONV);
      off += ENC_CONV;
    }
  }
  {
    d.add_f32_rows(F("dec_dense1"), 96, 96, 0, DEC_IN);
    d_pro = (int)d.chunks.size();
    d.add_f32_rows(F("dec_output"), DEC_OUT, DEC_OUTP, 0, 96);
    int off = 96;
    for (int l = 0; l < 5; l++) {
      std::string n = std::to_string(l + 1);
      d.add_layer_i8(I8("dec_gru" + n + "_input"), 288, off, 0, off / 32, DEC_GRU, 3);
      d.add_layer_i8(I8("dec_gru" + n + "_recurrent"), 288, 96, 0, 3, DEC_GRU, 3);
      d.add_layer_i8(I8("dec_glu" + n), 96, 96, 0, 3, 0, 1);
      d.add_f32_rows(F("dec_output"), DEC_OUT, DEC_OUTP, off, DEC_GRU);
      off += DEC_GRU;
      if (l == 4) d.add_f32_rows(F("dec_dense1"), 96, 96, 0, DEC_IN);             // for the next step
      d.add_layer_i8(I8("dec_conv" + n), 32, 2 * off, 0, off / 32, 0, 1);
      d.add_layer_i8(I8("dec_conv" + n), 32, 2 * off, off / 32, 2 * off / 32, 0, 1);
      d.add_f32_rows(F("dec_output"), DEC_OUT, DEC_OUTP, off, DEC_CONV);
      off += DEC_CONV;
    }
  }
}

}  // namespace

int core_weights_upload(const unsigned char *blob, size_t len, CoreWeightsHolder *h) {
  ArrayMap arrays;
  if (!parse_rdw(blob, len, arrays) && !parse_dnnw(blob, len, arrays)) {
    fprintf(stderr, "libradae_b200: weight blob is neither RDW v1 nor a DNNw blob (%zu bytes)\n", len);
    return -1;
  }
  normalise_io_width(arrays, &h->input_dim, &h->output_dim);
  std::vector<LayerSpec> specs; std::vector<std::string> names;
  layer_specs(specs, names);
  auto dev_copy = [&](const void *src, size_t bytes) -> void * {
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    h->allocs.push_back(d);
    return d;
  };
  std::map<std::string, I8LayerDev> i8; std::map<std::string, F32LayerDev> f32;
  std::map<std::string, const HostArray *> w8, wf;
  h->weight_bytes = 0;
  for (auto &L : specs) {
    std::string n = L.name;
    auto get = [&](const std::string &k, int dtype, int rows, int cols) -> const HostArray * {
      auto it = arrays.find(k);
      if (it == arrays.end() || it->second.dtype != dtype || it->second.rows != rows || it->second.cols != cols) return nullptr;
      return &it->second;
    };
    const HostArray *b = get(n + ".bias", 0, 1, L.nout);
    if (!b) { fprintf(stderr, "libradae_b200: missing %s.bias\n", L.name); return -1; }
    const float *dbias = (const float *)dev_copy(b->data.data(), b->data.size());
    if (!dbias) return -1;
    if (L.kind == 0) {
      const HostArray *w = get(n + ".wf", 0, L.nin, L.nout);
      if (!w) { fprintf(stderr, "libradae_b200: missing %s.wf\n", L.name); return -1; }
      wf[n] = w;
      f32[n] = {dbias, L.nin, L.nout};
      h->weight_bytes += w->data.size();
    } else {
      const HostArray *w = get(n + ".w8", 1, L.nout, L.nin), *s = get(n + ".scale", 0, 1, L.nout);
      if (!w || !s || (L.nin % 32) || (L.nout % 8)) { fprintf(stderr, "libradae_b200: missing/misshaped %s\n", L.name); return -1; }
      const float *ds = (const float *)dev_copy(s->data.data(), s->data.size());
      if (!ds) return -1;
      w8[n] = w;
      i8[n] = {ds, dbias, L.nin, L.nout};
      h->weight_bytes += w->data.size();
    }
  }
  CoreWeightsDev &W = h->dev;
  W.enc_z_tanh = 0;                      // set by the context from its flags (RADE_B200_BOTTLENECK_1)
  W.enc_dense1 = f32["enc_dense1"]; W.enc_zdense = f32["enc_zdense"]; W.dec_dense1 = f32["dec_dense1"]; W.dec_output = f32["dec_output"];
  for (int i = 0; i < 5; i++) {
    std::string n = std::to_string(i + 1);
    W.enc_gru_in[i] = i8["enc_gru" + n + "_input"]; W.enc_gru_rec[i] = i8["enc_gru" + n + "_recurrent"]; W.enc_conv[i] = i8["enc_conv" + n];
    W.dec_gru_in[i] = i8["dec_gru" + n + "_input"]; W.dec_gru_rec[i] = i8["dec_gru" + n + "_recurrent"];
    W.dec_glu[i] = i8["dec_glu" + n]; W.dec_conv[i] = i8["dec_conv" + n];
  }
  // ---- per-step weight streams, in the kernels' consumption order (keep in lock-step with core_codec.cu)
  StreamBuilder e, d;
  int e_pro = 0, d_pro = 0;
  {
    std::map<std::string, const int8_t *> p8; std::map<std::string, const float *> pf;
    for (auto &kv : w8) p8[kv.first] = (const int8_t *)kv.second->data.data();
    for (auto &kv : wf) pf[kv.first] = (const float *)kv.second->data.data();
    build_streams(p8, pf, false, e, d, e_pro, d_pro);
    if (core_codec_umma_enabled()) {             // the experimental tcgen05 encoder consumes its int8 chunks in the operand layout
      StreamBuilder e2, d2; int ep2 = 0, dp2 = 0;
      build_streams(p8, pf, true, e2, d2, ep2, dp2);
      e = e2; e_pro = ep2; d = d2; d_pro = dp2;
    }
  }
  if (!e.ok || !d.ok) { fprintf(stderr, "libradae_b200: internal error building the weight streams\n"); return -1; }
  auto up_stream = [&](StreamBuilder &sb, CodecStreamDev &out, int n_pro) -> int {
    out.n_prologue = n_pro;
    out.stream = (const unsigned char *)dev_copy(sb.bytes.data(), sb.bytes.size());
    out.chunks = (const ChunkDesc *)dev_copy(sb.chunks.data(), sb.chunks.size() * sizeof(ChunkDesc));
    out.n_chunks = (int)sb.chunks.size();
    return (out.stream && out.chunks) ? 0 : -1;
  };
  if (up_stream(e, W.enc_stream, e_pro) < 0 || up_stream(d, W.dec_stream, d_pro) < 0) return -1;
  if (core_codec_set_chunk_table(0, e.chunks.data(), (int)e.chunks.size()) < 0 ||
      core_codec_set_chunk_table(1, d.chunks.data(), (int)d.chunks.size()) < 0) return -1;
  h->enc_chunks_per_step = W.enc_stream.n_chunks; h->dec_chunks_per_step = W.dec_stream.n_chunks;
  return 0;
}

// Debug / test hook (no device involved): the weight stream of one codec as the host builds it, in either int8 chunk format.
// which: 0 encoder, 1 decoder.  Returns 0 and fills bytes / chunks / n_prologue, or -1.
int core_weights_debug_stream(const unsigned char *blob, size_t len, int which, int umma, std::vector<unsigned char> *bytes,
                              std::vector<ChunkDesc> *chunks, int *n_prologue) {
  ArrayMap arrays;
  if (!parse_rdw(blob, len, arrays) && !parse_dnnw(blob, len, arrays)) return -1;
  std::map<std::string, const int8_t *> p8; std::map<std::string, const float *> pf;
  for (auto &kv : arrays) {
    const std::string &k = kv.first;
    if (k.size() > 3 && k.compare(k.size() - 3, 3, ".w8") == 0) p8[k.substr(0, k.size() - 3)] = (const int8_t *)kv.second.data.data();
    if (k.size() > 3 && k.compare(k.size() - 3, 3, ".wf") == 0) pf[k.substr(0, k.size() - 3)] = (const float *)kv.second.data.data();
  }
  StreamBuilder e, d; int e_pro = 0, d_pro = 0;
  try { build_streams(p8, pf, umma != 0, e, d, e_pro, d_pro); } catch (...) { return -1; }
  if (!e.ok || !d.ok) return -1;
  StreamBuilder &sb = which ? d : e;
  *bytes = sb.bytes; *chunks = sb.chunks; *n_prologue = which ? d_pro : e_pro;
  return 0;
}
