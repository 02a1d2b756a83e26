// Host-side internal interfaces of libradae_b200 (not part of the C ABI).
#pragma once
#include <vector>
#include <stddef.h>
#include "rade_common.h"

struct CoreWeightsHolder {
  CoreWeightsDev dev;
  std::vector<void *> allocs;
  size_t weight_bytes = 0;
  int enc_chunks_per_step = 0, dec_chunks_per_step = 0;
  // 84 for model19_check3 (20 features + aux symbol per 10 ms vector), 80 for models without the aux symbol (model05): the
  // kernels always run the 84-wide layout, the missing inputs / outputs are zero weights (exact: + 0 * x)
  int input_dim = 84, output_dim = 84;
};
// k-blocks (32 inputs each) of an int8 layer with NTL n-tiles (8 outputs each) that fit one pipeline stage
static inline __host__ __device__ int core_kbc(int NTL) { int k = CORE_STAGE_BYTES / (NTL * 256); return k < 1 ? 1 : k; }
// rows of a float layer (padded width noutp) per pipeline stage, multiple of 4
static inline __host__ __device__ constexpr int core_f32_rpc(int noutp) { return (CORE_STAGE_BYTES / (noutp * 4)) & ~3; }
int core_weights_upload(const unsigned char *blob, size_t len, CoreWeightsHolder *h);
// codec kernel family: tcgen05 (core_codec_umma.cu, default) or mma.sync (core_codec.cu, RADE_B200_CODEC=mma)
int core_codec_use_umma();
int core_codec_umma_init_device();
int core_encoder_umma_launch(const CoreWeightsDev &W, EncStreamState *state, const float *in, int in_mode, float *z,
                             const uint8_t *active, int S, int T, cudaStream_t stream);
int core_decoder_umma_launch(const CoreWeightsDev &W, DecStreamState *state, const float *z, float *out, int out_mode,
                             int *uw_count, const uint8_t *active, int S, int T, cudaStream_t stream);
int core_weights_debug_stream(const unsigned char *blob, size_t len, int which, int umma, std::vector<unsigned char> *bytes,
                              std::vector<ChunkDesc> *chunks, int *n_prologue, std::vector<UmmaRec> *ops);
void core_weights_free(CoreWeightsHolder *h);
int core_weights_validate(const unsigned char *blob, size_t len, int *input_dim = nullptr, int *output_dim = nullptr);   // host only: 0 = acceptable, -1 = rejected

int core_encoder_launch(const CoreWeightsDev &W, EncStreamState *state, const float *in, int in_mode, float *z,
                        const uint8_t *active, int S, int T, cudaStream_t stream);
int core_decoder_launch(const CoreWeightsDev &W, DecStreamState *state, const float *z, float *out, int out_mode,
                        int *uw_count, const uint8_t *active, int S, int T, cudaStream_t stream);

extern "C" const unsigned char rade_b200_default_weights[];
extern "C" const unsigned char rade_b200_default_weights_end[];

#include <complex>
struct DspTablesHost {
  std::vector<float> w, bpf_h, fcoarse;
  std::vector<float> srch_basis, srch_expand;      // AcqTables::basis / ::expand (coarse grid in its low-rank basis)
  double srch_residual;                            // max |cos / sin - expansion| over the grid, checked at start-up
  std::vector<std::complex<float>> Winv, Wfwd, P, Pend, p, pend, p_w, cs_tab, Pmat, eq_rot, bpf_exp, eoo_base;
  double pilot_gain;
  float bpf_bw, bpf_centre, bpf_alpha;
};
void dsp_tables_host(DspTablesHost &T);
int dsp_tables_upload(const DspTablesHost &T, DspTables *D, std::vector<void *> &allocs);

// device buffers of the receiver (all [S] leading dimension)
struct RxBuffers {
  RxCtl *ctl;
  float2 *ring;            // [S][2112]
  float2 *bpf_mem;         // [S][102]
  float *rowsum;           // [S][2][960]
  int *uw_errors;          // [S]
  float *z_hat;            // [S][240]
  float *eoo;              // [S][180]
  DecStreamState *dec_state;
  unsigned char *dec_active;  // [S] valid_output of the last call
  int *nin;                // [S]
  int *search_list;        // [S] streams that need the coarse search this call (built by rx_bpf)
  int *track_list;         // [S] streams in sync this call (built by rx_bpf) -> rx_refresh / rx_track
  void *track_tmp;         // [S] TrackTmp (ofdm_rx.cu): refine + spot-correlation results handed from rx_track to rx_demod
  int *counters;           // [2][4] ping-pong by call parity: {search entries, search work-item counter, track entries, -};
                           //        rx_finish of call k zeroes the set call k+1 will use
  int parity;              // host-side: which counter set the next call uses
  cudaStream_t side_stream;   // the search branch (rx_detect -> rx_finish) runs here, concurrently with rx_track -> rx_demod
  cudaStream_t side2_stream;  // rx_track (fp64 refine) runs here, concurrently with rx_refresh (fp32) on the main stream
  cudaEvent_t ev_fork, ev_join, ev_join2;
};

int ofdm_mod_launch(const DspTables &T, const float *z, float2 *tx, int S, cudaStream_t stream);
int eoo_launch(const DspTables &T, const float *bits, const int *has_bits, float2 *tx, int S, cudaStream_t stream);
int tx_bpf_clip_launch(const DspTables &T, float2 *tx, size_t stride, int n, TxBpfState *st, int S, cudaStream_t stream);
int channel_apply_launch(float2 *rx, const float2 *tx, const float2 *G1, const float2 *G2, const float2 *noise, int S, int n,
                         int d, float mp_gain, float freq, float df_dt, float phase0, float sigma, float gain, cudaStream_t stream);
int channel_stream_launch(const DspTables &T, const float *z_mod, float2 *rx, const float2 *tx, ChanState *st, int S, float sigma, float freq0, float freq_spread,
                          float doppler, int d, float gain, unsigned long long seed, float2 *link_ring, long long *link_wr,
                          const long long *link_rd, int *link_overflow, cudaStream_t stream);
int link_push_launch(float2 *ring, long long *wr, const long long *rd, int *overflow, const float2 *in, int S, cudaStream_t stream);
int link_pop_launch(const float2 *ring, const long long *wr, long long *rd, const RxCtl *ctl, float2 *out,
                    unsigned char *active, int S, cudaStream_t stream);
int rx_init_launch(RxCtl *ctl, int *uw_errors, int S, double foff_err, cudaStream_t stream);
struct Profiler;
// link source (loop-back runs): when link_ring is non-null the band-pass kernel takes nin[s] samples from the stream's link FIFO
// (if it holds that many; otherwise the stream sits this call out) and writes the active flags itself into `active_out`
struct LinkSrc { const float2 *ring; const long long *wr; long long *rd; unsigned char *active_out; };
int rx_dsp_launch(const DspTables &T, RxBuffers &B, const float2 *rx_in, const unsigned char *active, const LinkSrc *link, int S,
                  int bpf_en, int reset_dec_on_sync, int *ret_out, cudaStream_t stream, Profiler *prof);

// ---- optional per-kernel timing with CUDA events on the context's stream (rade_b200_profile_*)
enum KernelId { K_CORE_ENC = 0, K_OFDM_MOD, K_EOO, K_CHANNEL, K_LINK_PUSH, K_LINK_POP, K_RX_BPF, K_RX_DETECT, K_RX_TRACK,
                K_RX_DEMOD, K_RX_FINISH, K_CORE_DEC, K_TX_BPF, K_RX_REFRESH, K_COUNT };
struct Profiler {
  bool on = false;                            // serialising mode: everything on the main stream, one event pair per launch
  bool timeline = false;                      // timeline mode: event pairs on the LAUNCHING stream, nothing is serialised (rade_b200_timeline_*)
  cudaStream_t stream = nullptr;
  cudaEvent_t t0 = nullptr;                   // timeline origin
  std::vector<cudaEvent_t> ev[K_COUNT];       // start,end pairs
  void begin(int k, cudaStream_t s = nullptr) { if (on || timeline) rec(k, s ? s : stream); }
  void end(int k, cudaStream_t s = nullptr) { if (on || timeline) rec(k, s ? s : stream); }
  void rec(int k, cudaStream_t s) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); ev[k].push_back(e); }
};

int core_codec_init_device();
int core_codec_set_chunk_table(int which, const ChunkDesc *d, int n);
int rx_dsp_init_device();
