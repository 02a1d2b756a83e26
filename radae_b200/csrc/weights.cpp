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
#include "umma_program.h"

namespace {

struct HostArray { int dtype; int rows, cols; std::vector<unsigned char> data; };
typedef std::map<std::string, HostArray> ArrayMap;

// Every length that comes from the file is checked against the buffer before it is used (overflow-safe: all comparisons are
// of the form `n <= len - off` with `off <= len` established first), and a payload must be exactly rows * cols * sizeof(dtype)
// bytes — a truncated or crafted container is rejected as a whole, nothing is read past `buf + len`.
bool parse_rdw(const unsigned char *buf, size_t len, ArrayMap &out) {
  if (len < 64 || memcmp(buf, "RADEB200", 8) != 0) return false;
  uint32_t version, n;
  memcpy(&version, buf + 8, 4); memcpy(&n, buf + 12, 4);
  if (version != 1 || (size_t)n > (len - 64) / 80) return false;
  const size_t table_end = 64 + 80 * (size_t)n;
  for (uint32_t i = 0; i < n; i++) {
    const unsigned char *e = buf + 64 + 80 * (size_t)i;
    char name[49]; memcpy(name, e, 48); name[48] = 0;
    uint32_t dtype, rows, cols; uint64_t off, nbytes;
    memcpy(&dtype, e + 48, 4); memcpy(&rows, e + 52, 4); memcpy(&cols, e + 56, 4);
    memcpy(&off, e + 64, 8); memcpy(&nbytes, e + 72, 8);
    if (dtype > 2 || rows == 0 || cols == 0 || rows > (1u << 20) || cols > (1u << 20)) return false;
    const uint64_t want = (uint64_t)rows * cols * (dtype == 1 ? 1u : 4u);   // 0 f32, 1 int8, 2 int32 (radae_b200/rdw.py)      // <= 2^42: cannot overflow
    if (nbytes != want || off < table_end || off > len || nbytes > len - off) return false;
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

// DNNw blob -> the same flat map RDW gives (unblocking the 8x4 tiles).  Record headers, record types / sizes and every entry of
// a block index list are bounded by what the file actually holds; the int8 payload and the index list must be consumed exactly.
bool parse_dnnw(const unsigned char *buf, size_t len, ArrayMap &out) {
  struct Raw { int type; const unsigned char *p; size_t size; };
  std::map<std::string, Raw> raw;
  size_t off = 0;
  while (off < len) {
    if (len - off < 64 || memcmp(buf + off, "DNNw", 4) != 0) return false;
    int version, type, size, block;
    memcpy(&version, buf + off + 4, 4); memcpy(&type, buf + off + 8, 4);
    memcpy(&size, buf + off + 12, 4); memcpy(&block, buf + off + 16, 4);
    if (version != 0 || size <= 0 || block < size || (size_t)block > len - off - 64) return false;
    char name[45]; memcpy(name, buf + off + 20, 44); name[44] = 0;
    raw[name] = {type, buf + off + 64, (size_t)size};
    off += 64 + (size_t)block;
  }
  std::vector<LayerSpec> specs; std::vector<std::string> names;
  layer_specs(specs, names);
  // record types of the reference's writer (src/write_rade_weights.c:56-72): 0 float, 1 int, 3 int8
  for (auto &L : specs) {
    std::string n = L.name;
    auto need = [&](const std::string &k, int type, size_t bytes) -> const unsigned char * {
      auto it = raw.find(k);
      return (it != raw.end() && it->second.type == type && it->second.size == bytes) ? it->second.p : nullptr;
    };
    // models without the auxiliary symbol (model05, src/test_rade_enc.c:40): enc_dense1 has 80 inputs, dec_output 80 outputs;
    // they are ingested as they are and widened with zeros by normalise_io_width() below
    if (n == "enc_dense1" && !need(n + "_weights_float", 0, 4 * (size_t)L.nin * L.nout)) L.nin = 80;
    if (n == "dec_output" && !need(n + "_bias", 0, 4 * (size_t)L.nout)) L.nout = 80;
    const unsigned char *b = need(n + "_bias", 0, 4 * (size_t)L.nout);
    if (!b) return false;
    HostArray bias; bias.dtype = 0; bias.rows = 1; bias.cols = L.nout; bias.data.assign(b, b + 4 * (size_t)L.nout);
    out[n + ".bias"] = bias;
    if (L.kind == 0) {
      const unsigned char *w = need(n + "_weights_float", 0, 4 * (size_t)L.nin * L.nout);
      if (!w) return false;
      HostArray a; a.dtype = 0; a.rows = L.nin; a.cols = L.nout; a.data.assign(w, w + 4 * (size_t)L.nin * L.nout);
      out[n + ".wf"] = a;
    } else {
      const unsigned char *sc = need(n + "_scale", 0, 4 * (size_t)L.nout);
      auto itw = raw.find(n + "_weights_int8");
      if (!sc || itw == raw.end() || itw->second.type != 3) return false;
      const int8_t *wb = (const int8_t *)itw->second.p;
      const size_t wsize = itw->second.size;
      const unsigned char *idx = nullptr; size_t idx_n = 0, ip = 0;      // index list as 4-byte ints, read with memcpy (alignment)
      if (L.kind == 2) {
        auto ii = raw.find(n + "_weights_idx");
        if (ii == raw.end() || ii->second.type != 1 || (ii->second.size & 3)) return false;
        idx = ii->second.p; idx_n = ii->second.size / 4;
      }
      auto next_idx = [&](int *v) -> bool { if (ip >= idx_n) return false; memcpy(v, idx + 4 * ip++, 4); return true; };
      HostArray a; a.dtype = 1; a.rows = L.nout; a.cols = L.nin; a.data.assign((size_t)L.nin * L.nout, 0);
      size_t p = 0;
      for (int ob = 0; ob < L.nout / 8; ob++) {
        int nblk = L.nin / 4;
        if (idx && (!next_idx(&nblk) || nblk < 0 || nblk > L.nin / 4)) return false;
        for (int j = 0; j < nblk; j++) {
          int pos = 4 * j;
          if (idx && !next_idx(&pos)) return false;
          if (pos < 0 || pos > L.nin - 4 || wsize < 32 || p > wsize - 32) return false;
          for (int k = 0; k < 8; k++) for (int c = 0; c < 4; c++)
            a.data[(size_t)(ob * 8 + k) * L.nin + pos + c] = (unsigned char)wb[p + 4 * k + c];
          p += 32;
        }
      }
      if (p != wsize || ip != idx_n) return false;       // payload and index list consumed exactly (as rdw.py's unblock_int8)
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
  void add_layer_i8(const int8_t *W8, int N, int K, int kb_lo, int kb_hi, int, int) { add_i8(W8, N, K, kb_lo, kb_hi); }
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
void build_streams(const std::map<std::string, const int8_t *> &p8, const std::map<std::string, const float *> &pf,
                   StreamBuilder &e, StreamBuilder &d, int &e_pro, int &d_pro) {
  auto I8 = [&](const std::string &n) { return p8.at(n); };
  auto F = [&](const std::string &n) { return pf.at(n); };
  // Order = the kernels' consumption order.  dense1 is needed once up front (prologue) and then, for the NEXT step, just before
  // the last conv layer, so that the F-warps can have the next step's first activation ready when the I-warps finish this one.
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


// ---- tcgen05 formulation (core_codec_umma.cu): per codec an int8 stream (weight images — k-blocks of ALL rows of a matrix in
// the canonical K-major operand layout, umma_layout.h — packed back to back into ring stages of <= 40 KB, one bulk copy each),
// laid out exactly as the compile-time MMA program says (umma_program.h: the device code is generated from the same table), and
// a float stream (rows of dense1 / zdense / output in the float warps' order, packed into stages of <= 22 KB).
struct UmmaBuilder {
  std::vector<unsigned char> i8, f32;
  std::vector<ChunkDesc> i8_chunks, f32_chunks;       // ring stages (= bulk copies)
  std::vector<UmmaRec> recs;                          // the program in the debug hook's format
  std::vector<unsigned char> f32_cur;                 // the float stage being filled
  int f32_prologue = 0;
  bool ok = true;
  void close_f32() {
    if (f32_cur.empty()) return;
    ChunkDesc d; d.offset = (unsigned)f32.size(); d.bytes = (unsigned)f32_cur.size();
    f32.insert(f32.end(), f32_cur.begin(), f32_cur.end());
    f32_chunks.push_back(d);
    f32_cur.clear();
  }
  // the int8 stream of one codec from its compile-time program
  void bake_i8(const UmmaProgC &P, const std::map<int, const int8_t *> &mats) {
    int stage = 0;
    std::vector<unsigned char> cur;
    for (int i = 0; i < P.n; i++) {
      const UmmaRecC &r = P.r[i];
      auto it = mats.find(r.mat);
      if (it == mats.end() || (size_t)r.a_off16 * 16 != cur.size()) { ok = false; return; }
      std::vector<uint8_t> c = umma_bake_chunk(it->second, r.n_rows, r.K, r.kb, r.nk);
      cur.insert(cur.end(), c.begin(), c.end());
      UmmaRec o; memset(&o, 0, sizeof(o));
      o.a_off16 = (unsigned short)r.a_off16; o.tile_step = (unsigned short)r.tile_step; o.b_kb = (unsigned short)r.b_kb;
      o.nk = (unsigned char)r.nk; o.n_tiles = (unsigned char)r.n_tiles; o.b_buf = (unsigned char)r.b_buf; o.flags = (unsigned char)r.flags;
      o.d_blk = (unsigned char)r.d_blk; o.d_tile_stride = (unsigned char)r.d_tile_stride; o.dep = (signed char)r.dep; o.commit = (signed char)r.commit;
      recs.push_back(o);
      if (r.flags & UR_STAGE_LAST) {
        if (stage >= P.n_stages || (int)cur.size() != P.stage_bytes[stage] || cur.size() > UMMA_I8_STAGE_BYTES) { ok = false; return; }
        ChunkDesc d; d.offset = (unsigned)i8.size(); d.bytes = (unsigned)cur.size();
        i8.insert(i8.end(), cur.begin(), cur.end());
        i8_chunks.push_back(d);
        cur.clear(); stage++;
      }
    }
    if (!cur.empty() || stage != P.n_stages) ok = false;
  }
  // rows [j0, j0 + nrows) of a float [K][NOUT] matrix, each zero-padded to NOUTP floats; a stage always holds a multiple of 4 rows
  // of a segment (the float warps read the activations four at a time)
  void add_f32(const float *Wf, int NOUT, int NOUTP, int j0, int nrows) {
    const int rb = NOUTP * 4;
    for (int r0 = 0; r0 < nrows;) {
      const int space = UMMA_F32_STAGE_BYTES - (int)f32_cur.size();
      const int n = std::min(nrows - r0, space / rb) & ~3;
      if (n < 4) { if (f32_cur.empty()) { ok = false; return; } close_f32(); continue; }
      std::vector<float> t((size_t)n * NOUTP, 0.f);
      for (int r = 0; r < n; r++) memcpy(&t[(size_t)r * NOUTP], Wf + (size_t)(j0 + r0 + r) * NOUT, NOUT * sizeof(float));
      const unsigned char *p = (const unsigned char *)t.data();
      f32_cur.insert(f32_cur.end(), p, p + (size_t)n * rb);
      r0 += n;
    }
  }
};

void build_umma(const std::map<std::string, const int8_t *> &p8, const std::map<std::string, const float *> &pf, UmmaBuilder &e, UmmaBuilder &d) {
  auto F = [&](const std::string &n) { return pf.at(n); };
  std::map<int, const int8_t *> mats;
  for (int l = 0; l < 5; l++) {
    std::string n = std::to_string(l + 1);
    mats[UM_ENC_GRU_IN + l] = p8.at("enc_gru" + n + "_input"); mats[UM_ENC_GRU_REC + l] = p8.at("enc_gru" + n + "_recurrent");
    mats[UM_ENC_CONV + l] = p8.at("enc_conv" + n);
    mats[UM_DEC_GRU_IN + l] = p8.at("dec_gru" + n + "_input"); mats[UM_DEC_GRU_REC + l] = p8.at("dec_gru" + n + "_recurrent");
    mats[UM_DEC_GLU + l] = p8.at("dec_glu" + n); mats[UM_DEC_CONV + l] = p8.at("dec_conv" + n);
  }
  e.bake_i8(kUmmaEncProg, mats);
  d.bake_i8(kUmmaDecProg, mats);
  {
    e.add_f32(F("enc_dense1"), 64, 64, 0, ENC_IN);                         // prologue: dense1 of step 0
    e.close_f32();
    e.f32_prologue = (int)e.f32_chunks.size();
    e.add_f32(F("enc_zdense"), RADE_LATENT, RADE_LATENT, 0, 64);           // concat segment 0 = dense1 output
    e.add_f32(F("enc_dense1"), 64, 64, 0, ENC_IN);                         // dense1 of the next step
    int off = 64;
    for (int l = 0; l < 5; l++) {
      e.add_f32(F("enc_zdense"), RADE_LATENT, RADE_LATENT, off, ENC_GRU); off += ENC_GRU;
      e.add_f32(F("enc_zdense"), RADE_LATENT, RADE_LATENT, off, ENC_CONV); off += ENC_CONV;
    }
    e.close_f32();
  }
  {
    d.add_f32(F("dec_dense1"), 96, 96, 0, DEC_IN);
    d.close_f32();
    d.f32_prologue = (int)d.f32_chunks.size();
    d.add_f32(F("dec_output"), DEC_OUT, DEC_OUTP, 0, 96);
    d.add_f32(F("dec_dense1"), 96, 96, 0, DEC_IN);
    int off = 96;
    for (int l = 0; l < 5; l++) {
      d.add_f32(F("dec_output"), DEC_OUT, DEC_OUTP, off, DEC_GRU); off += DEC_GRU;
      d.add_f32(F("dec_output"), DEC_OUT, DEC_OUTP, off, DEC_CONV); off += DEC_CONV;
    }
    d.close_f32();
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
      if (it == arrays.end() || it->second.dtype != dtype || it->second.rows != rows || it->second.cols != cols ||
          it->second.data.size() != (size_t)rows * cols * (dtype == 1 ? 1 : 4)) return nullptr;
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
  UmmaBuilder ue, ud;
  int e_pro = 0, d_pro = 0;
  {
    std::map<std::string, const int8_t *> p8; std::map<std::string, const float *> pf;
    for (auto &kv : w8) p8[kv.first] = (const int8_t *)kv.second->data.data();
    for (auto &kv : wf) pf[kv.first] = (const float *)kv.second->data.data();
    build_streams(p8, pf, e, d, e_pro, d_pro);
    build_umma(p8, pf, ue, ud);
  }
  if (!e.ok || !d.ok || !ue.ok || !ud.ok || ue.recs.size() > UMMA_MAX_RECS || ud.recs.size() > UMMA_MAX_RECS ||
      ue.i8_chunks.size() > UMMA_MAX_I8_CHUNKS || ud.i8_chunks.size() > UMMA_MAX_I8_CHUNKS ||
      ue.f32_chunks.size() > UMMA_MAX_F32_CHUNKS || ud.f32_chunks.size() > UMMA_MAX_F32_CHUNKS) {
    fprintf(stderr, "libradae_b200: internal error building the weight streams\n"); return -1;
  }
  auto up_umma = [&](UmmaBuilder &ub, UmmaCodecDev &out) -> int {
    out.i8_stream = (const unsigned char *)dev_copy(ub.i8.data(), ub.i8.size());
    out.i8_chunks = (const ChunkDesc *)dev_copy(ub.i8_chunks.data(), ub.i8_chunks.size() * sizeof(ChunkDesc));
    out.n_i8_chunks = (int)ub.i8_chunks.size();
    out.f32_stream = (const unsigned char *)dev_copy(ub.f32.data(), ub.f32.size());
    out.f32_chunks = (const ChunkDesc *)dev_copy(ub.f32_chunks.data(), ub.f32_chunks.size() * sizeof(ChunkDesc));
    out.n_f32_chunks = (int)ub.f32_chunks.size(); out.n_f32_prologue = ub.f32_prologue;
    return (out.i8_stream && out.i8_chunks && out.f32_stream && out.f32_chunks) ? 0 : -1;
  };
  if (up_umma(ue, W.enc_umma) < 0 || up_umma(ud, W.dec_umma) < 0) return -1;
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

// Host-only validation of a weight blob (no device involved): container parse + every array the layer table needs present with
// the right dtype and shape.  0 = core_weights_upload would accept it, -1 = rejected.  (rade_b200_debug_check_weights)
int core_weights_validate(const unsigned char *blob, size_t len, int *input_dim, int *output_dim) {
  ArrayMap arrays;
  if (!blob || (!parse_rdw(blob, len, arrays) && !parse_dnnw(blob, len, arrays))) return -1;
  int in_dim, out_dim;
  normalise_io_width(arrays, &in_dim, &out_dim);
  if (input_dim) *input_dim = in_dim;
  if (output_dim) *output_dim = out_dim;
  std::vector<LayerSpec> specs; std::vector<std::string> names;
  layer_specs(specs, names);
  auto ok = [&](const std::string &k, int dtype, int rows, int cols) {
    auto it = arrays.find(k);
    return it != arrays.end() && it->second.dtype == dtype && it->second.rows == rows && it->second.cols == cols &&
           it->second.data.size() == (size_t)rows * cols * (dtype == 1 ? 1 : 4);
  };
  for (auto &L : specs) {
    std::string n = L.name;
    if (!ok(n + ".bias", 0, 1, L.nout)) return -1;
    if (L.kind == 0) { if (!ok(n + ".wf", 0, L.nin, L.nout)) return -1; }
    else if (!ok(n + ".w8", 1, L.nout, L.nin) || !ok(n + ".scale", 0, 1, L.nout)) return -1;
  }
  return 0;
}

// Debug / test hook (no device involved): the weight stream of one codec as the host builds it, in either int8 chunk format.
// which: 0 encoder, 1 decoder.  Returns 0 and fills bytes / chunks / n_prologue, or -1.
int core_weights_debug_stream(const unsigned char *blob, size_t len, int which, int umma, std::vector<unsigned char> *bytes,
                              std::vector<ChunkDesc> *chunks, int *n_prologue, std::vector<UmmaRec> *ops) {
  ArrayMap arrays;
  if (!parse_rdw(blob, len, arrays) && !parse_dnnw(blob, len, arrays)) return -1;
  std::map<std::string, const int8_t *> p8; std::map<std::string, const float *> pf;
  for (auto &kv : arrays) {
    const std::string &k = kv.first;
    if (k.size() > 3 && k.compare(k.size() - 3, 3, ".w8") == 0) p8[k.substr(0, k.size() - 3)] = (const int8_t *)kv.second.data.data();
    if (k.size() > 3 && k.compare(k.size() - 3, 3, ".wf") == 0) pf[k.substr(0, k.size() - 3)] = (const float *)kv.second.data.data();
  }
  if (umma) {                                  // 1: int8 stream, 2: float stream of the tcgen05 formulation
    UmmaBuilder ue, ud;
    try { build_umma(p8, pf, ue, ud); } catch (...) { return -1; }
    if (!ue.ok || !ud.ok) return -1;
    UmmaBuilder &ub = which ? ud : ue;
    if (umma == 1) { *bytes = ub.i8; *chunks = ub.i8_chunks; *n_prologue = 0; }
    else { *bytes = ub.f32; *chunks = ub.f32_chunks; *n_prologue = ub.f32_prologue; }
    if (ops) *ops = ub.recs;
    return 0;
  }
  StreamBuilder e, d; int e_pro = 0, d_pro = 0;
  try { build_streams(p8, pf, e, d, e_pro, d_pro); } catch (...) { return -1; }
  if (!e.ok || !d.ok) return -1;
  StreamBuilder &sb = which ? d : e;
  *bytes = sb.bytes; *chunks = sb.chunks; *n_prologue = which ? d_pro : e_pro;
  return 0;
}
