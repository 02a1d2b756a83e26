// C ABI of libradae_b200: the batched context (include/rade_b200.h) and, on top of a 1-stream batch, the
// reference's single-stream surface (include/rade_api.h <-> /root/reference/src/rade_api.h:71-129, rade_api.c).
#include <cstring>
#include <vector>
#include <string>
#include "rade_common.h"
#include "rade_host.h"
#include "rade_b200.h"
#include <omp.h>

#define LINK_CAP 4096

struct rade_batch {
  int S, device, flags;
  cudaStream_t stream;
  // optional software pipeline: TX-side device calls (tx_dev, channel_link_dev) go to their own stream so that frame k+1's
  // transmitter runs concurrently with frame k's receiver; fork / join are explicit (rade_b200_pipeline_*)
  cudaStream_t tx_stream; cudaEvent_t ev_txfork, ev_txjoin; int pipelined;
  // rade_b200_loopback_step_dev replays the whole step (9 kernels on up to 3 streams) as ONE CUDA graph launch; two
  // instances because the receiver's work-list counters ping-pong between calls
  struct StepGraph { const void *key[5]; cudaGraphExec_t exec; int n_kernels; };
  std::vector<StepGraph> step_graphs;
  long long launches;
  Profiler prof;
  CoreWeightsHolder weights;
  std::vector<void *> allocs;
  DspTables tables;
  // state
  EncStreamState *enc_state;
  RxBuffers rx;
  ChanState *chan_state;
  rade_b200_channel_cfg chan_cfg;
  float2 *link_ring; long long *link_wr, *link_rd; int *link_overflow;    // device loop-back FIFOs; frames dropped because a FIFO was full
  // transmitter
  float *z_tx;                // [S][240]
  float *eoo_bits; int *has_eoo_bits;
  TxBpfState *tx_bpf; int tx_bpf_en;      // optional TX band-pass filter + clip (radae_tx(txbpf_en=True)), allocated on first enable
  // device staging for the host-pointer API
  float *d_feat_in, *d_feat_out, *d_z, *d_core_in, *d_core_out; size_t core_cap;
  float2 *d_tx, *d_tx_eoo, *d_rx_in;
  int *d_ret; unsigned char *d_active;
  // pinned host staging
  float *h_feat; float2 *h_cplx; int *h_int;
  // rade_b200_loopback_run: double-buffered device staging + copy streams
  float *lb_fin[2], *lb_fout[2]; int *lb_ret[2]; cudaStream_t lb_h2d, lb_d2h; cudaEvent_t lb_ev_in[2], lb_ev_step[2], lb_ev_out[2];
};

namespace {

template <typename T> int dalloc(rade_batch *b, T **p, size_t n) {
  void *d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, n * sizeof(T)));
  CUDA_CHECK(cudaMemset(d, 0, n * sizeof(T)));
  b->allocs.push_back(d);
  *p = (T *)d;
  return 0;
}

int reset_state(rade_batch *b) {
  const int S = b->S;
  CUDA_CHECK(cudaMemsetAsync(b->enc_state, 0, sizeof(EncStreamState) * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.dec_state, 0, sizeof(DecStreamState) * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.ring, 0, sizeof(float2) * RADE_RXBUF * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.bpf_mem, 0, sizeof(float2) * RADE_BPF_MEM * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.rowsum, 0, sizeof(float) * 2 * RADE_NMF * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.z_hat, 0, sizeof(float) * 240 * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.eoo, 0, sizeof(float) * RADE_NEOO_BITS * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.dec_active, 0, S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->rx.counters, 0, 8 * sizeof(int), b->stream));
  b->rx.parity = 0;
  CUDA_CHECK(cudaMemsetAsync(b->chan_state, 0, sizeof(ChanState) * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->link_wr, 0, sizeof(long long) * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->link_rd, 0, sizeof(long long) * S, b->stream));
  CUDA_CHECK(cudaMemsetAsync(b->link_overflow, 0, sizeof(int), b->stream));
  for (auto &g : b->step_graphs) cudaGraphExecDestroy(g.exec);      // captured graphs hold the old state's arguments
  b->step_graphs.clear();
  CUDA_CHECK(cudaMemsetAsync(b->has_eoo_bits, 0, sizeof(int) * S, b->stream));
  if (b->tx_bpf) CUDA_CHECK(cudaMemsetAsync(b->tx_bpf, 0, sizeof(TxBpfState) * S, b->stream));
  const double foff_err = (b->flags & RADE_FOFF_TEST) ? 10.0 : 0.0;     // src/rade_api.c:263-264
  if (rx_init_launch(b->rx.ctl, b->rx.uw_errors, S, foff_err, b->stream) < 0) return -1;
  std::vector<int> nin(S, RADE_NMF);
  CUDA_CHECK(cudaMemcpyAsync(b->rx.nin, nin.data(), sizeof(int) * S, cudaMemcpyHostToDevice, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  b->launches += 1;
  return 0;
}

// Device-visible alias of a pinned (cudaMallocHost / cudaHostRegister) host buffer, or nullptr.
// RADE_B200_ZERO_COPY: 0 (default) every host array is staged through cudaMemcpyAsync — the copy engines move 7.9 MB in 0.15 ms
// either way without occupying an SM; 1: kernels read and write pinned buffers in place over PCIe (round 1's default: a kernel
// that writes in place reaches 48 GB/s but holds its SMs for the whole transfer, a kernel that reads in place gets 19-27 GB/s —
// with transmitter, channel and receiver contexts sharing one GPU that serialised them, tools/e2e_breakdown.py); 2: writes in
// place, reads staged.
int zero_copy_mode() {
  static const int mode = getenv("RADE_B200_ZERO_COPY") ? atoi(getenv("RADE_B200_ZERO_COPY")) : 0;
  return mode;
}
void *pinned_alias(const void *p, bool for_read = false) {
  const int mode = zero_copy_mode();
  if (mode == 0 || (mode == 2 && for_read)) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer) return a.devicePointer;
  cudaGetLastError();
  return nullptr;
}

int ensure_core_staging(rade_batch *b, size_t floats) {
  if (floats <= b->core_cap) return 0;
  if (b->d_core_in) { cudaFree(b->d_core_in); cudaFree(b->d_core_out); }
  CUDA_CHECK(cudaMalloc((void **)&b->d_core_in, floats * sizeof(float)));
  CUDA_CHECK(cudaMalloc((void **)&b->d_core_out, floats * sizeof(float)));
  b->core_cap = floats;
  return 0;
}

}  // namespace

extern "C" {

RADE_EXPORT const void *rade_b200_default_weights_blob(size_t *len) {
  if (len) *len = (size_t)(rade_b200_default_weights_end - rade_b200_default_weights);
  return rade_b200_default_weights;
}

RADE_EXPORT rade_batch *rade_b200_open(int n_streams, int device, int flags, const void *weights, size_t weights_len) {
  if (n_streams <= 0) { fprintf(stderr, "libradae_b200: n_streams must be positive\n"); return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fprintf(stderr, "libradae_b200: no CUDA device available (%s) — this library has no CPU fallback\n", cudaGetErrorString(e));
    return nullptr;
  }
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) { fprintf(stderr, "libradae_b200: cannot select device %d\n", device); return nullptr; }
  if (device < 0) cudaGetDevice(&device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    fprintf(stderr, "libradae_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
    return nullptr;
  }
  rade_batch *b = new rade_batch();
  b->S = n_streams; b->device = device; b->flags = flags; b->launches = 0; b->core_cap = 0;
  b->tx_stream = nullptr; b->pipelined = 0;
  b->d_core_in = b->d_core_out = nullptr;
  b->lb_fin[0] = nullptr; b->lb_h2d = b->lb_d2h = nullptr;
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) { delete b; return nullptr; }
  b->prof.stream = b->stream;
  if (core_codec_init_device() < 0 || rx_dsp_init_device() < 0) { delete b; return nullptr; }
  if (cudaStreamCreateWithFlags(&b->rx.side_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&b->rx.side2_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&b->rx.ev_join2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&b->rx.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&b->rx.ev_join, cudaEventDisableTiming) != cudaSuccess) { delete b; return nullptr; }
  if (!weights) { weights = rade_b200_default_weights_blob(&weights_len); }
  if (core_weights_upload((const unsigned char *)weights, weights_len, &b->weights) < 0) { delete b; return nullptr; }
  b->weights.dev.trace = nullptr;
  b->weights.dev.one = 1.0f;
  b->weights.dev.full_tiles = 0;
  b->weights.dev.float_fma = 0;            // measurement switch only (tools/gpu_fma_ab.sh sets it through RADE_B200_DEBUG_FLOAT_FMA): never on in the product
  if (getenv("RADE_B200_DEBUG_FLOAT_FMA") && atoi(getenv("RADE_B200_DEBUG_FLOAT_FMA")) == 1) b->weights.dev.float_fma = 1;
  b->weights.dev.enc_z_tanh = (flags & RADE_B200_BOTTLENECK_1) ? 1 : 0;      // src/rade_enc.c:107-113
  DspTablesHost th; dsp_tables_host(th);
  if (dsp_tables_upload(th, &b->tables, b->allocs) < 0) { delete b; return nullptr; }
  const size_t S = n_streams;
  int bad = 0;
  bad |= dalloc(b, &b->enc_state, S);
  bad |= dalloc(b, &b->rx.dec_state, S);
  bad |= dalloc(b, &b->rx.ctl, S);
  bad |= dalloc(b, &b->rx.ring, S * RADE_RXBUF);
  bad |= dalloc(b, &b->rx.bpf_mem, S * RADE_BPF_MEM);
  bad |= dalloc(b, &b->rx.rowsum, S * 2 * RADE_NMF);
  bad |= dalloc(b, &b->rx.uw_errors, S);
  bad |= dalloc(b, &b->rx.z_hat, S * 240);
  bad |= dalloc(b, &b->rx.eoo, S * RADE_NEOO_BITS);
  bad |= dalloc(b, &b->rx.dec_active, S);
  bad |= dalloc(b, &b->rx.nin, S);
  bad |= dalloc(b, &b->rx.search_list, S);
  bad |= dalloc(b, &b->rx.track_list, S);
  { double *p = nullptr; bad |= dalloc(b, &p, (size_t)S * 8); b->rx.track_tmp = p; }      // TrackTmp = 64 bytes
  bad |= dalloc(b, &b->rx.counters, 8);
  bad |= dalloc(b, &b->chan_state, S);
  bad |= dalloc(b, &b->link_ring, S * LINK_CAP);
  bad |= dalloc(b, &b->link_wr, S);
  bad |= dalloc(b, &b->link_rd, S);
  bad |= dalloc(b, &b->link_overflow, (size_t)1);
  bad |= dalloc(b, &b->z_tx, S * 240);
  bad |= dalloc(b, &b->eoo_bits, S * RADE_NEOO_BITS);
  bad |= dalloc(b, &b->has_eoo_bits, S);
  bad |= dalloc(b, &b->d_feat_in, S * RADE_NFEAT);
  bad |= dalloc(b, &b->d_feat_out, S * RADE_NFEAT);
  bad |= dalloc(b, &b->d_z, S * 240);
  bad |= dalloc(b, &b->d_tx, S * RADE_NMF);
  bad |= dalloc(b, &b->d_tx_eoo, S * RADE_NEOO);
  bad |= dalloc(b, &b->d_rx_in, S * RADE_NIN_MAX);
  bad |= dalloc(b, &b->d_ret, S);
  bad |= dalloc(b, &b->d_active, S);
  if (bad) { rade_b200_close(b); return nullptr; }
  if (cudaMallocHost((void **)&b->h_feat, S * RADE_NFEAT * sizeof(float)) != cudaSuccess ||
      cudaMallocHost((void **)&b->h_cplx, S * RADE_NEOO * sizeof(float2)) != cudaSuccess ||
      cudaMallocHost((void **)&b->h_int, S * 4 * sizeof(int)) != cudaSuccess) { rade_b200_close(b); return nullptr; }
  memset(&b->chan_cfg, 0, sizeof(b->chan_cfg));
  b->chan_cfg.EbNodB = 100.f; b->chan_cfg.gain = 1.f; b->chan_cfg.delay_samples = 16; b->chan_cfg.seed = 1;
  if (reset_state(b) < 0) { rade_b200_close(b); return nullptr; }
  return b;
}

RADE_EXPORT void rade_b200_close(rade_batch *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  cudaStreamSynchronize(b->stream);
  if (b->rx.side_stream) cudaStreamSynchronize(b->rx.side_stream);
  if (b->rx.side2_stream) cudaStreamSynchronize(b->rx.side2_stream);
  for (void *p : b->allocs) cudaFree(p);
  core_weights_free(&b->weights);
  if (b->d_core_in) { cudaFree(b->d_core_in); cudaFree(b->d_core_out); }
  if (b->h_feat) cudaFreeHost(b->h_feat);
  if (b->h_cplx) cudaFreeHost(b->h_cplx);
  if (b->h_int) cudaFreeHost(b->h_int);
  if (b->lb_h2d) {
    cudaStreamDestroy(b->lb_h2d); cudaStreamDestroy(b->lb_d2h);
    for (int i = 0; i < 2; i++) { cudaEventDestroy(b->lb_ev_in[i]); cudaEventDestroy(b->lb_ev_step[i]); cudaEventDestroy(b->lb_ev_out[i]); }
  }
  for (auto &g : b->step_graphs) cudaGraphExecDestroy(g.exec);
  cudaStreamDestroy(b->stream);
  if (b->tx_stream) { cudaStreamDestroy(b->tx_stream); cudaEventDestroy(b->ev_txfork); cudaEventDestroy(b->ev_txjoin); }
  if (b->rx.side2_stream) { cudaStreamDestroy(b->rx.side2_stream); cudaEventDestroy(b->rx.ev_join2); }
  if (b->rx.side_stream) { cudaStreamDestroy(b->rx.side_stream); cudaEventDestroy(b->rx.ev_fork); cudaEventDestroy(b->rx.ev_join); }
  delete b;
}

RADE_EXPORT int rade_b200_n_streams(rade_batch *b) { return b->S; }
RADE_EXPORT void *rade_b200_cuda_stream(rade_batch *b) { return (void *)b->stream; }
RADE_EXPORT int rade_b200_synchronize(rade_batch *b) {
  if (b->tx_stream) CUDA_CHECK(cudaStreamSynchronize(b->tx_stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
// ---- software pipeline over frames for device-pointer callers: after pipeline_enable(1), rade_b200_tx_dev and
// rade_b200_channel_link_dev enqueue on a second stream.  pipeline_fork makes that stream wait for everything enqueued so far
// on the main stream, pipeline_join makes the main stream wait for it.  Typical step: fork; tx_dev(frame k+1);
// channel_link_dev; rx_link_dev(frame k, main stream); join.
RADE_EXPORT int rade_b200_pipeline_enable(rade_batch *b, int enable) {
  cudaSetDevice(b->device);
  if (enable && !b->tx_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&b->tx_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_txfork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_txjoin, cudaEventDisableTiming));
  }
  if (rade_b200_synchronize(b) < 0) return -1;
  b->pipelined = enable ? 1 : 0;
  b->weights.dev.full_tiles = b->pipelined;      // kernels of the other side of the pipeline run beside the codec
  return 0;
}
RADE_EXPORT int rade_b200_pipeline_fork(rade_batch *b) {
  if (!b->pipelined) return 0;
  CUDA_CHECK(cudaEventRecord(b->ev_txfork, b->stream));
  CUDA_CHECK(cudaStreamWaitEvent(b->tx_stream, b->ev_txfork, 0));
  return 0;
}
RADE_EXPORT int rade_b200_pipeline_join(rade_batch *b) {
  if (!b->pipelined) return 0;
  CUDA_CHECK(cudaEventRecord(b->ev_txjoin, b->tx_stream));
  CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->ev_txjoin, 0));
  return 0;
}
static cudaStream_t tx_side(rade_batch *b) { return (b->pipelined && !b->prof.on) ? b->tx_stream : b->stream; }
RADE_EXPORT long long rade_b200_launch_count(rade_batch *b) { return b->launches; }
RADE_EXPORT int rade_b200_reset(rade_batch *b) { return reset_state(b); }

// ------------------------------------------------------------------ core codec
RADE_EXPORT int rade_b200_core_encode_dev(rade_batch *b, float *d_z, const float *d_features, int n_steps) {
  cudaSetDevice(b->device);        // the current device is per host thread
  b->prof.begin(K_CORE_ENC);
  if (core_encoder_launch(b->weights.dev, b->enc_state, d_features, 0, d_z, nullptr, b->S, n_steps, b->stream) < 0) return -1;
  b->prof.end(K_CORE_ENC);
  b->launches += 1;
  return 0;
}
RADE_EXPORT int rade_b200_core_decode_dev(rade_batch *b, float *d_features, const float *d_z, int n_steps) {
  cudaSetDevice(b->device);        // the current device is per host thread
  b->prof.begin(K_CORE_DEC);
  if (core_decoder_launch(b->weights.dev, b->rx.dec_state, d_z, d_features, 0, nullptr, nullptr, b->S, n_steps, b->stream) < 0) return -1;
  b->prof.end(K_CORE_DEC);
  b->launches += 1;
  return 0;
}
RADE_EXPORT int rade_b200_core_dims(rade_batch *b, int *input_dim, int *output_dim) {
  if (input_dim) *input_dim = b->weights.input_dim;
  if (output_dim) *output_dim = b->weights.output_dim;
  return 0;
}
RADE_EXPORT int rade_b200_core_encode(rade_batch *b, float *z, const float *features, int n_steps) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t rows = (size_t)b->S * n_steps, nin = rows * ENC_IN, nout = rows * RADE_LATENT;
  const int w = b->weights.input_dim;                 // caller's row width: 84, or 80 for a model without the aux symbol
  if (ensure_core_staging(b, nin) < 0) return -1;
  if (w == ENC_IN) {
    CUDA_CHECK(cudaMemcpyAsync(b->d_core_in, features, nin * sizeof(float), cudaMemcpyHostToDevice, b->stream));
  } else {                                            // widen to the kernels' 84-float rows, missing inputs = 0
    CUDA_CHECK(cudaMemsetAsync(b->d_core_in, 0, nin * sizeof(float), b->stream));
    CUDA_CHECK(cudaMemcpy2DAsync(b->d_core_in, ENC_IN * sizeof(float), features, w * sizeof(float), w * sizeof(float), rows,
                                 cudaMemcpyHostToDevice, b->stream));
  }
  if (rade_b200_core_encode_dev(b, b->d_core_out, b->d_core_in, n_steps) < 0) return -1;
  CUDA_CHECK(cudaMemcpyAsync(z, b->d_core_out, nout * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
RADE_EXPORT int rade_b200_core_decode(rade_batch *b, float *features, const float *z, int n_steps) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t nin = (size_t)b->S * n_steps * RADE_LATENT, nout = (size_t)b->S * n_steps * DEC_OUT;
  if (ensure_core_staging(b, nout) < 0) return -1;
  CUDA_CHECK(cudaMemcpyAsync(b->d_core_in, z, nin * sizeof(float), cudaMemcpyHostToDevice, b->stream));
  if (rade_b200_core_decode_dev(b, b->d_core_out, b->d_core_in, n_steps) < 0) return -1;
  const int w = b->weights.output_dim;
  if (w == DEC_OUT) {
    CUDA_CHECK(cudaMemcpyAsync(features, b->d_core_out, nout * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  } else {                                            // hand back the model's own 80-float rows
    CUDA_CHECK(cudaMemcpy2DAsync(features, w * sizeof(float), b->d_core_out, DEC_OUT * sizeof(float), w * sizeof(float),
                                 (size_t)b->S * n_steps, cudaMemcpyDeviceToHost, b->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}

// ------------------------------------------------------------------ transmitter
RADE_EXPORT int rade_b200_ofdm_mod_dev(rade_batch *b, RADE_COMP *d_tx_out, const float *d_z) {
  cudaSetDevice(b->device);        // the current device is per host thread
  b->prof.begin(K_OFDM_MOD);
  if (ofdm_mod_launch(b->tables, d_z, (float2 *)d_tx_out, b->S, b->stream) < 0) return -1;
  b->prof.end(K_OFDM_MOD);
  b->launches += 1;
  return 0;
}
// optional TX band-pass filter + clip on the frame the modulator has just written (radae_txe.py:130-132, :141-143)
static int tx_bpf_stage(rade_batch *b, float2 *d_frames, int n, cudaStream_t stream) {
  if (!b->tx_bpf_en) return 0;
  b->prof.begin(K_TX_BPF);
  if (tx_bpf_clip_launch(b->tables, d_frames, (size_t)n, n, b->tx_bpf, b->S, stream) < 0) return -1;
  b->prof.end(K_TX_BPF);
  b->launches += 1;
  return 0;
}
RADE_EXPORT int rade_b200_tx_bpf_enable(rade_batch *b, int enable) {
  cudaSetDevice(b->device);        // the current device is per host thread
  if (enable && !b->tx_bpf && dalloc(b, &b->tx_bpf, (size_t)b->S) < 0) return -1;
  if (b->tx_bpf) {                 // switching the filter on or off starts from a new filter object
    if (rade_b200_synchronize(b) < 0) return -1;
    CUDA_CHECK(cudaMemsetAsync(b->tx_bpf, 0, sizeof(TxBpfState) * b->S, b->stream));
    CUDA_CHECK(cudaStreamSynchronize(b->stream));
  }
  b->tx_bpf_en = enable != 0;
  return 0;
}
RADE_EXPORT int rade_b200_tx_dev(rade_batch *b, RADE_COMP *d_tx_out, const float *d_features_in) {
  cudaSetDevice(b->device);        // the current device is per host thread
  // 3 core-encoder steps on the API feature layout (src/rade_api.c:411-434) then transmitter_one (radae_txe.py:127)
  b->prof.begin(K_CORE_ENC);
  if (core_encoder_launch(b->weights.dev, b->enc_state, d_features_in, 1, b->z_tx, nullptr, b->S, RADE_NZMF, tx_side(b)) < 0) return -1;
  b->prof.end(K_CORE_ENC); b->prof.begin(K_OFDM_MOD);
  if (ofdm_mod_launch(b->tables, b->z_tx, (float2 *)d_tx_out, b->S, tx_side(b)) < 0) return -1;
  b->prof.end(K_OFDM_MOD);
  b->launches += 2;
  if (tx_bpf_stage(b, (float2 *)d_tx_out, RADE_NMF, tx_side(b)) < 0) return -1;
  return RADE_NMF;
}
// radae_tx(bypass_enc=True).do_radae_tx (radae_txe.py:122-132): the caller ran the core encoder and hands over the
// 3 x 80 latents of one modem frame per stream; modulator, then the optional TX filter
RADE_EXPORT int rade_b200_tx_z_dev(rade_batch *b, RADE_COMP *d_tx_out, const float *d_z) {
  cudaSetDevice(b->device);        // the current device is per host thread
  b->prof.begin(K_OFDM_MOD);
  if (ofdm_mod_launch(b->tables, d_z, (float2 *)d_tx_out, b->S, b->stream) < 0) return -1;
  b->prof.end(K_OFDM_MOD);
  b->launches += 1;
  if (tx_bpf_stage(b, (float2 *)d_tx_out, RADE_NMF, b->stream) < 0) return -1;
  return RADE_NMF;
}
RADE_EXPORT int rade_b200_tx_z(rade_batch *b, RADE_COMP *tx_out, const float *z) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t S = b->S;
  CUDA_CHECK(cudaMemcpyAsync(b->d_z, z, S * RADE_NZMF * RADE_LATENT * sizeof(float), cudaMemcpyHostToDevice, b->stream));
  if (rade_b200_tx_z_dev(b, (RADE_COMP *)b->d_tx, b->d_z) < 0) return -1;
  CUDA_CHECK(cudaMemcpyAsync(tx_out, b->d_tx, S * RADE_NMF * sizeof(float2), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return RADE_NMF;
}
RADE_EXPORT int rade_b200_tx(rade_batch *b, RADE_COMP *tx_out, const float *features_in) {
  cudaSetDevice(b->device);        // the current device is per host thread
  // copies go straight from / to the caller's buffers: truly asynchronous when they are pinned (cudaHostAlloc /
  // cudaHostRegister), staged by the driver when they are pageable
  const size_t S = b->S;
  CUDA_CHECK(cudaMemcpyAsync(b->d_feat_in, features_in, S * RADE_NFEAT * sizeof(float), cudaMemcpyHostToDevice, b->stream));
  if (void *alias = b->tx_bpf_en ? nullptr : pinned_alias(tx_out)) {     // modulator writes the samples straight into the caller's pinned buffer
    if (rade_b200_pipeline_fork(b) < 0 || rade_b200_tx_dev(b, (RADE_COMP *)alias, b->d_feat_in) < 0 || rade_b200_pipeline_join(b) < 0) return -1;
  } else {
    if (rade_b200_pipeline_fork(b) < 0 || rade_b200_tx_dev(b, (RADE_COMP *)b->d_tx, b->d_feat_in) < 0 || rade_b200_pipeline_join(b) < 0) return -1;
    CUDA_CHECK(cudaMemcpyAsync(tx_out, b->d_tx, S * RADE_NMF * sizeof(float2), cudaMemcpyDeviceToHost, b->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return RADE_NMF;
}
RADE_EXPORT int rade_b200_tx_set_eoo_bits(rade_batch *b, const float *eoo_bits) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t S = b->S;
  std::vector<int> ones(S, 1);
  CUDA_CHECK(cudaMemcpyAsync(b->eoo_bits, eoo_bits, S * RADE_NEOO_BITS * sizeof(float), cudaMemcpyHostToDevice, b->stream));
  CUDA_CHECK(cudaMemcpyAsync(b->has_eoo_bits, ones.data(), S * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
RADE_EXPORT int rade_b200_tx_eoo(rade_batch *b, RADE_COMP *tx_eoo_out) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t S = b->S;
  b->prof.begin(K_EOO);
  if (eoo_launch(b->tables, b->eoo_bits, b->has_eoo_bits, b->d_tx_eoo, b->S, b->stream) < 0) return -1;
  b->prof.end(K_EOO);
  b->launches += 1;
  if (tx_bpf_stage(b, b->d_tx_eoo, RADE_NEOO, b->stream) < 0) return -1;
  CUDA_CHECK(cudaMemcpyAsync(tx_eoo_out, b->d_tx_eoo, S * RADE_NEOO * sizeof(float2), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return RADE_NEOO;
}

// ------------------------------------------------------------------ receiver
RADE_EXPORT int rade_b200_rx_dev(rade_batch *b, float *d_features_out, int *d_ret, float *d_eoo_out, const RADE_COMP *d_rx_in,
                                 const unsigned char *d_active) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const int reset_dec = (b->flags & RADE_USE_C_DECODER) ? 0 : 1;
  int n = rx_dsp_launch(b->tables, b->rx, (const float2 *)d_rx_in, d_active, nullptr, b->S, 1, reset_dec, d_ret, b->stream, &b->prof);
  if (n < 0) return -1;
  b->prof.begin(K_CORE_DEC);
  // decoder last, only for streams with valid output; counts aux-symbol errors (src/rade_api.c:494-513)
  if (core_decoder_launch(b->weights.dev, b->rx.dec_state, b->rx.z_hat, d_features_out, 1, b->rx.uw_errors, b->rx.dec_active,
                          b->S, RADE_NZMF, b->stream) < 0) return -1;
  b->prof.end(K_CORE_DEC);
  b->launches += n + 1;
  if (d_eoo_out && d_eoo_out != b->rx.eoo) {
    CUDA_CHECK(cudaMemcpyAsync(d_eoo_out, b->rx.eoo, (size_t)b->S * RADE_NEOO_BITS * sizeof(float), cudaMemcpyDeviceToDevice, b->stream));
  }
  return 0;
}
RADE_EXPORT const int *rade_b200_nin_dev(rade_batch *b) { return b->rx.nin; }
RADE_EXPORT int rade_b200_nin(rade_batch *b, int *nin) {
  cudaSetDevice(b->device);        // the current device is per host thread
  CUDA_CHECK(cudaMemcpyAsync(nin, b->rx.nin, sizeof(int) * b->S, cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
RADE_EXPORT int rade_b200_rx(rade_batch *b, float *features_out, int *ret, float *eoo_out, const RADE_COMP *rx_in,
                             const unsigned char *active) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t S = b->S;
  CUDA_CHECK(cudaMemcpyAsync(b->d_rx_in, rx_in, S * RADE_NIN_MAX * sizeof(float2), cudaMemcpyHostToDevice, b->stream));
  if (active) CUDA_CHECK(cudaMemcpyAsync(b->d_active, active, S, cudaMemcpyHostToDevice, b->stream));
  if (rade_b200_rx_dev(b, b->d_feat_out, b->d_ret, nullptr, (const RADE_COMP *)b->d_rx_in, active ? b->d_active : nullptr) < 0) return -1;
  CUDA_CHECK(cudaMemcpyAsync(features_out, b->d_feat_out, S * RADE_NFEAT * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaMemcpyAsync(ret, b->d_ret, S * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (eoo_out) CUDA_CHECK(cudaMemcpyAsync(eoo_out, b->rx.eoo, S * RADE_NEOO_BITS * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
RADE_EXPORT int rade_b200_rx_get_status(rade_batch *b, rade_b200_rx_status *status) {
  cudaSetDevice(b->device);        // the current device is per host thread
  std::vector<RxCtl> ctl(b->S); std::vector<int> uw(b->S);
  CUDA_CHECK(cudaMemcpyAsync(ctl.data(), b->rx.ctl, sizeof(RxCtl) * b->S, cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaMemcpyAsync(uw.data(), b->rx.uw_errors, sizeof(int) * b->S, cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  for (int s = 0; s < b->S; s++) {
    const RxCtl &c = ctl[s]; rade_b200_rx_status &o = status[s];
    o.state = c.state; o.nin = c.nin; o.tmax = c.tmax; o.valid_count = c.valid_count; o.uw_errors = uw[s];
    o.synced_count = c.synced_count; o.snrdB_3k_est = (int)c.snr_est; o.snrdB_3k_est_f = (float)c.snr_est;
    o.fmax = c.fmax; o.Dthresh = c.Dthresh; o.Dtmax12 = c.Dtmax12; o.Dtmax12_eoo = c.Dtmax12_eoo;
  }
  return 0;
}
RADE_EXPORT int rade_b200_rx_get_z_hat(rade_batch *b, float *z_hat) {
  cudaSetDevice(b->device);        // the current device is per host thread
  CUDA_CHECK(cudaMemcpyAsync(z_hat, b->rx.z_hat, (size_t)b->S * 240 * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}

// ------------------------------------------------------------------ channel + link
RADE_EXPORT int rade_b200_channel_apply_dev(rade_batch *b, RADE_COMP *d_rx, const RADE_COMP *d_tx, const RADE_COMP *d_G1,
                                            const RADE_COMP *d_G2, const RADE_COMP *d_noise, int n, int delay,
                                            float mp_gain, float freq_offset_hz, float phase0, float sigma, float gain) {
  cudaSetDevice(b->device);        // the current device is per host thread
  if (channel_apply_launch((float2 *)d_rx, (const float2 *)d_tx, (const float2 *)d_G1, (const float2 *)d_G2, (const float2 *)d_noise,
                           b->S, n, delay, mp_gain, freq_offset_hz, 0.f, phase0, sigma, gain, b->stream) < 0) return -1;
  b->launches += 1;
  return 0;
}
// host-pointer form of the explicit channel: tx, G1, G2, noise are [S][n] complex64 host arrays (G from a fading file,
// radae_b200/gfile.py), rx [S][n] comes back
RADE_EXPORT int rade_b200_channel_apply_drift(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx, const RADE_COMP *G1, const RADE_COMP *G2,
                                              const RADE_COMP *noise, int n, int delay, float mp_gain, float freq_offset_hz, float df_dt,
                                              float phase0, float sigma, float gain);
RADE_EXPORT int rade_b200_channel_apply(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx, const RADE_COMP *G1, const RADE_COMP *G2,
                                        const RADE_COMP *noise, int n, int delay, float mp_gain, float freq_offset_hz, float phase0,
                                        float sigma, float gain) {
  return rade_b200_channel_apply_drift(b, rx, tx, G1, G2, noise, n, delay, mp_gain, freq_offset_hz, 0.f, phase0, sigma, gain);
}
// ... with a frequency drift df_dt (Hz/s): phase = phase0 + cumsum(2 pi (f + df_dt i / Fs) / Fs), radae.py:546-550
RADE_EXPORT int rade_b200_channel_apply_drift(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx, const RADE_COMP *G1, const RADE_COMP *G2,
                                              const RADE_COMP *noise, int n, int delay, float mp_gain, float freq_offset_hz, float df_dt,
                                              float phase0, float sigma, float gain) {
  cudaSetDevice(b->device);
  const size_t bytes = (size_t)b->S * n * sizeof(float2);
  float2 *d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const void *src[4] = {tx, G1, G2, noise};
  int rc = 0;
  for (int i = 0; i < 5 && rc == 0; i++) if (cudaMalloc((void **)&d[i], bytes) != cudaSuccess) rc = -1;
  for (int i = 0; i < 4 && rc == 0; i++) if (cudaMemcpyAsync(d[i], src[i], bytes, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) rc = -1;
  if (rc == 0 && channel_apply_launch(d[4], d[0], d[1], d[2], d[3], b->S, n, delay, mp_gain, freq_offset_hz, df_dt, phase0, sigma, gain, b->stream) < 0) rc = -1;
  if (rc == 0 && cudaMemcpyAsync(rx, d[4], bytes, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess) rc = -1;
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) rc = -1;
  for (int i = 0; i < 5; i++) if (d[i]) cudaFree(d[i]);
  if (rc == 0) b->launches += 1;
  return rc;
}
RADE_EXPORT int rade_b200_channel_config(rade_batch *b, const rade_b200_channel_cfg *cfg) {
  cudaSetDevice(b->device);        // the current device is per host thread
  if (cfg->delay_samples < 0 || cfg->delay_samples > 64) return -1;
  b->chan_cfg = *cfg;
  CUDA_CHECK(cudaMemsetAsync(b->chan_state, 0, sizeof(ChanState) * b->S, b->stream));
  for (auto &g : b->step_graphs) cudaGraphExecDestroy(g.exec);      // the channel parameters are baked into captured graphs
  b->step_graphs.clear();
  return 0;
}
RADE_EXPORT int rade_b200_channel_dev(rade_batch *b, RADE_COMP *d_rx, const RADE_COMP *d_tx) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const rade_b200_channel_cfg &c = b->chan_cfg;
  const float sigma = sqrtf((float)RADE_FS / (powf(10.f, c.EbNodB / 10.f) * 2000.f));       // radae.py:570-574, Rb = 80/0.04
  b->prof.begin(K_CHANNEL);
  if (channel_stream_launch(b->tables, nullptr, (float2 *)d_rx, (const float2 *)d_tx, b->chan_state, b->S, sigma, c.freq_offset_hz,
                            c.freq_offset_spread_hz, c.doppler_spread_hz, c.delay_samples, c.gain, c.seed, nullptr, nullptr, nullptr, nullptr, b->stream) < 0) return -1;
  b->prof.end(K_CHANNEL);
  b->launches += 1;
  return 0;
}
RADE_EXPORT int rade_b200_channel(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx) {
  cudaSetDevice(b->device);        // the current device is per host thread
  const size_t S = b->S;
  const void *src = pinned_alias(tx, true); void *dst = pinned_alias(rx);
  if (!src) { CUDA_CHECK(cudaMemcpyAsync(b->d_tx, tx, S * RADE_NMF * sizeof(float2), cudaMemcpyHostToDevice, b->stream)); src = b->d_tx; }
  if (rade_b200_channel_dev(b, (RADE_COMP *)(dst ? dst : (void *)b->d_rx_in), (const RADE_COMP *)src) < 0) return -1;
  if (!dst) CUDA_CHECK(cudaMemcpyAsync(rx, b->d_rx_in, S * RADE_NMF * sizeof(float2), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
// loop-back runs without the two copy kernels: the channel writes straight into the per-stream link FIFOs, and the
// receiver's band-pass kernel pops nin[s] samples per stream from them (streams without enough samples sit the call out)
RADE_EXPORT int rade_b200_channel_link_dev(rade_batch *b, const RADE_COMP *d_tx) {
  cudaSetDevice(b->device);
  const rade_b200_channel_cfg &c = b->chan_cfg;
  const float sigma = sqrtf((float)RADE_FS / (powf(10.f, c.EbNodB / 10.f) * 2000.f));
  b->prof.begin(K_CHANNEL);
  if (channel_stream_launch(b->tables, nullptr, nullptr, (const float2 *)d_tx, b->chan_state, b->S, sigma, c.freq_offset_hz, c.freq_offset_spread_hz,
                            c.doppler_spread_hz, c.delay_samples, c.gain, c.seed, b->link_ring, b->link_wr, b->link_rd, b->link_overflow, tx_side(b)) < 0) return -1;
  b->prof.end(K_CHANNEL);
  b->launches += 1;
  return 0;
}
// transmitter side of the loop-back in two launches: core encoder, then ONE kernel that modulates the frame and pushes it
// through the channel into the link FIFOs (the 960 tx samples per stream stay in shared memory)
RADE_EXPORT int rade_b200_tx_channel_link_dev(rade_batch *b, const float *d_features_in) {
  cudaSetDevice(b->device);
  if (b->tx_bpf_en) {      // the fused modulator+channel kernel has no filter stage: use tx_dev + channel_link_dev instead
    fprintf(stderr, "libradae_b200: rade_b200_tx_channel_link_dev is not available while the TX band-pass filter is enabled\n");
    return -1;
  }
  const rade_b200_channel_cfg &c = b->chan_cfg;
  const float sigma = sqrtf((float)RADE_FS / (powf(10.f, c.EbNodB / 10.f) * 2000.f));
  b->prof.begin(K_CORE_ENC, tx_side(b));
  if (core_encoder_launch(b->weights.dev, b->enc_state, d_features_in, 1, b->z_tx, nullptr, b->S, RADE_NZMF, tx_side(b)) < 0) return -1;
  b->prof.end(K_CORE_ENC, tx_side(b)); b->prof.begin(K_CHANNEL, tx_side(b));
  if (channel_stream_launch(b->tables, b->z_tx, nullptr, nullptr, b->chan_state, b->S, sigma, c.freq_offset_hz, c.freq_offset_spread_hz,
                            c.doppler_spread_hz, c.delay_samples, c.gain, c.seed, b->link_ring, b->link_wr, b->link_rd, b->link_overflow, tx_side(b)) < 0) return -1;
  b->prof.end(K_CHANNEL, tx_side(b));
  b->launches += 2;
  return 0;
}
RADE_EXPORT int rade_b200_rx_link_dev(rade_batch *b, float *d_features_out, int *d_ret, float *d_eoo_out) {
  cudaSetDevice(b->device);
  const int reset_dec = (b->flags & RADE_USE_C_DECODER) ? 0 : 1;
  LinkSrc ls = {b->link_ring, b->link_wr, b->link_rd, b->d_active};
  int n = rx_dsp_launch(b->tables, b->rx, nullptr, nullptr, &ls, b->S, 1, reset_dec, d_ret, b->stream, &b->prof);
  if (n < 0) return -1;
  b->prof.begin(K_CORE_DEC);
  if (core_decoder_launch(b->weights.dev, b->rx.dec_state, b->rx.z_hat, d_features_out, 1, b->rx.uw_errors, b->rx.dec_active,
                          b->S, RADE_NZMF, b->stream) < 0) return -1;
  b->prof.end(K_CORE_DEC);
  b->launches += n + 1;
  if (d_eoo_out && d_eoo_out != b->rx.eoo) {
    CUDA_CHECK(cudaMemcpyAsync(d_eoo_out, b->rx.eoo, (size_t)b->S * RADE_NEOO_BITS * sizeof(float), cudaMemcpyDeviceToDevice, b->stream));
  }
  return 0;
}
// One loop-back step for device-pointer callers: transmitter side of the NEXT frame (core encoder, modulator, channel -> link
// FIFOs) and receiver side of the frame already in the FIFOs (pop nin[s], DSP, core decoder), on two streams when the frame
// pipeline is enabled.  The sequence is captured once into a CUDA graph (per counter parity) and replayed with a single launch
// per step when RADE_B200_GRAPH=1 (for hosts that cannot keep up with ~25 stream operations per step).  Default: plain
// launches — measured on the B200 box the step is not launch-bound and the graph launch latency costs 1.5 % (0.395 vs 0.389 ms).
static int loopback_step_body(rade_batch *b, const float *d_features, float *d_features_out, int *d_ret, float *d_eoo_out) {
  if (rade_b200_pipeline_fork(b) < 0) return -1;
  if (rade_b200_tx_channel_link_dev(b, d_features) < 0) return -1;
  if (rade_b200_rx_link_dev(b, d_features_out, d_ret, d_eoo_out) < 0) return -1;
  return rade_b200_pipeline_join(b);
}
RADE_EXPORT int rade_b200_loopback_step_dev(rade_batch *b, const float *d_features_next, float *d_features_out, int *d_ret,
                                            float *d_eoo_out) {
  cudaSetDevice(b->device);
  static const bool graphs = getenv("RADE_B200_GRAPH") && atoi(getenv("RADE_B200_GRAPH")) != 0;
  if (!graphs || b->prof.on) return loopback_step_body(b, d_features_next, d_features_out, d_ret, d_eoo_out);
  // one instantiated graph per (caller pointers, pipeline mode, counter parity): callers cycle through a few buffers
  const void *key[5] = {d_features_next, d_features_out, d_ret, d_eoo_out, (const void *)(size_t)(2 * b->pipelined + b->rx.parity + 1)};
  cudaGraphExec_t exec = nullptr; int n_kernels = 0;
  for (auto &g : b->step_graphs) if (memcmp(g.key, key, sizeof(key)) == 0) { exec = g.exec; n_kernels = g.n_kernels; break; }
  if (!exec) {
    if (b->step_graphs.size() >= 64) { for (auto &g : b->step_graphs) cudaGraphExecDestroy(g.exec); b->step_graphs.clear(); }
    cudaGraph_t g = nullptr;
    CUDA_CHECK(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
    const long long launches0 = b->launches;
    const int rc = loopback_step_body(b, d_features_next, d_features_out, d_ret, d_eoo_out);  // toggles the parity once
    n_kernels = (int)(b->launches - launches0);             // kernels inside the graph: counted per replay below
    b->launches = launches0;
    cudaError_t e = cudaStreamEndCapture(b->stream, &g);
    if (rc < 0 || e != cudaSuccess || !g) { fprintf(stderr, "libradae_b200: graph capture of the loop-back step failed (%s)\n", cudaGetErrorString(e)); return -1; }
    CUDA_CHECK(cudaGraphInstantiate(&exec, g, 0));
    cudaGraphDestroy(g);
    rade_batch::StepGraph sg; memcpy(sg.key, key, sizeof(key)); sg.exec = exec; sg.n_kernels = n_kernels;
    b->step_graphs.push_back(sg);
  } else {
    b->rx.parity ^= 1;
  }
  CUDA_CHECK(cudaGraphLaunch(exec, b->stream));
  b->launches += n_kernels;
  return 0;
}
// The whole loop-back pipeline (features -> core encoder -> OFDM modulator -> HF channel -> receiver -> core decoder -> features)
// for n_frames modem frames with HOST buffers at both ends: per frame the S x 432 input features go up from (pinned) host memory,
// the S x 432 recovered features and S return codes come back; the modem samples stay on the device (this is the call for a
// simulation on one machine — the three-program pipe of the reference moves every sample through host memory four times,
// rade_b200_duplex_run).  Copies run on their own streams, double-buffered, so the upload of frame k + 1 and the download of
// frame k - 1 overlap the kernels of frame k.  features_in: n_in frames of [S][432], cycled over; features_out [S][432] / ret [S]:
// overwritten every frame (the last frame's values remain; rows of streams whose ret lacks RADE_B200_VALID are unspecified);
// valid_frames [S] (optional): += frames that returned features.
RADE_EXPORT int rade_b200_loopback_run(rade_batch *b, const float *features_in, int n_in, int n_frames, float *features_out, int *ret,
                                       long long *valid_frames) {
  if (!b || !features_in || n_in < 1 || n_frames < 0 || !features_out || !ret) return -1;
  cudaSetDevice(b->device);
  const size_t S = b->S, fb = S * RADE_NFEAT * sizeof(float);
  if (!b->lb_h2d) {
    for (int i = 0; i < 2; i++) {
      if (dalloc(b, &b->lb_fin[i], S * RADE_NFEAT) < 0 || dalloc(b, &b->lb_fout[i], S * RADE_NFEAT) < 0 || dalloc(b, &b->lb_ret[i], S) < 0) return -1;
      CUDA_CHECK(cudaEventCreateWithFlags(&b->lb_ev_in[i], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&b->lb_ev_step[i], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&b->lb_ev_out[i], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaStreamCreateWithFlags(&b->lb_h2d, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&b->lb_d2h, cudaStreamNonBlocking));
  }
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  for (int k = 0; k < n_frames; k++) {
    const int sl = k & 1;
    if (k >= 2) CUDA_CHECK(cudaStreamWaitEvent(b->lb_h2d, b->lb_ev_step[sl], 0));            // frame k - 2 has consumed this input buffer
    CUDA_CHECK(cudaMemcpyAsync(b->lb_fin[sl], features_in + (size_t)(k % n_in) * S * RADE_NFEAT, fb, cudaMemcpyHostToDevice, b->lb_h2d));
    CUDA_CHECK(cudaEventRecord(b->lb_ev_in[sl], b->lb_h2d));
    CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->lb_ev_in[sl], 0));
    if (k >= 2) CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->lb_ev_out[sl], 0));              // its outputs have left this output buffer
    if (rade_b200_loopback_step_dev(b, b->lb_fin[sl], b->lb_fout[sl], b->lb_ret[sl], nullptr) < 0) return -1;
    CUDA_CHECK(cudaEventRecord(b->lb_ev_step[sl], b->stream));
    if (k >= 1) {                                                                           // frame k - 1 is on the host: count it
      CUDA_CHECK(cudaEventSynchronize(b->lb_ev_out[sl ^ 1]));
      if (valid_frames) for (size_t s = 0; s < S; s++) valid_frames[s] += (ret[s] & 1);
    }
    CUDA_CHECK(cudaStreamWaitEvent(b->lb_d2h, b->lb_ev_step[sl], 0));
    CUDA_CHECK(cudaMemcpyAsync(features_out, b->lb_fout[sl], fb, cudaMemcpyDeviceToHost, b->lb_d2h));
    CUDA_CHECK(cudaMemcpyAsync(ret, b->lb_ret[sl], S * sizeof(int), cudaMemcpyDeviceToHost, b->lb_d2h));
    CUDA_CHECK(cudaEventRecord(b->lb_ev_out[sl], b->lb_d2h));
  }
  if (n_frames > 0) {
    CUDA_CHECK(cudaEventSynchronize(b->lb_ev_out[(n_frames - 1) & 1]));
    if (valid_frames) for (size_t s = 0; s < S; s++) valid_frames[s] += (ret[s] & 1);
  }
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  return 0;
}
RADE_EXPORT int rade_b200_link_push_dev(rade_batch *b, const RADE_COMP *d_samples) {
  cudaSetDevice(b->device);        // the current device is per host thread
  b->prof.begin(K_LINK_PUSH);
  if (link_push_launch(b->link_ring, b->link_wr, b->link_rd, b->link_overflow, (const float2 *)d_samples, b->S, b->stream) < 0) return -1;
  b->prof.end(K_LINK_PUSH);
  b->launches += 1;
  return 0;
}
RADE_EXPORT int rade_b200_link_pop_dev(rade_batch *b, RADE_COMP *d_rx_in, unsigned char *d_active) {
  cudaSetDevice(b->device);        // the current device is per host thread
  b->prof.begin(K_LINK_POP);
  if (link_pop_launch(b->link_ring, b->link_wr, b->link_rd, b->rx.ctl, (float2 *)d_rx_in, d_active, b->S, b->stream) < 0) return -1;
  b->prof.end(K_LINK_POP);
  b->launches += 1;
  return 0;
}

// ---- host-side sample link (batched host I/O around the C ABI) between whatever produces receive samples
// and rade_b200_rx, which wants nin[s] in {800, 960, 1120} fresh samples per stream per call.
// Host side: a ring of HL_SLOTS modem-frame slots [S][960] in pinned memory (what a producer fills: rade_b200_hostlink_push by
// host memcpy, rade_b200_channel_hostlink by letting the channel kernel write its output there in place).  Device side: the
// receiver context's per-stream sample rings (the loop-back link).  rade_b200_hostlink_rx moves queued frames host -> device
// with the COPY ENGINE (one contiguous 7.9 MB cudaMemcpyAsync per frame of 1024 streams, issued ahead on its own stream so it
// overlaps the previous call's kernels), appends them to the per-stream rings on the device and runs the receiver on the rings.
// (Measured, tools/e2e_breakdown.py: a kernel READING pinned host memory in place gets 19-27 GB/s over PCIe — the band-pass
// kernel took 0.41 ms instead of 0.025 — while the copy engine moves the same bytes in 0.155 ms; kernel WRITES in place reach
// 48 GB/s, which is why producers on the device still write their frames straight into the host slots.)
// Single producer / single consumer per link; they may be different host threads with a context each.
#define HL_SLOTS 4
struct rade_b200_hostlink {
  rade_batch *b; int nthreads;
  float2 *frames, *d_frames;                                 // pinned [HL_SLOTS][S][960] + its device alias
  volatile long long pushed, popped;                         // frames queued by the producer / released by the consumer
  long long staged, appended, rx_calls;                      // consumer side: frames whose H2D copy was issued / appended to the device rings
  float2 *d_stage[2];                                        // device staging of a frame
  cudaStream_t copy_stream; cudaEvent_t ev_copied[2], ev_appended[2];
  unsigned char *active;                                     // pinned: which streams the last rx call advanced
  long long dropped;                                         // frames refused because all slots were full
};
RADE_EXPORT rade_b200_hostlink *rade_b200_hostlink_open(rade_batch *b, int capacity_samples) {
  cudaSetDevice(b->device);
  (void)capacity_samples;                                    // fixed: HL_SLOTS frames on the host, 4096 samples per stream on the device
  rade_b200_hostlink *h = new rade_b200_hostlink();
  memset((void *)h, 0, sizeof(*h));
  h->b = b;
  // OpenMP team of rade_b200_hostlink_push: RADE_B200_HOST_THREADS, else OpenMP's default (launchers such as torchrun export
  // OMP_NUM_THREADS=1, which would serialise 8 MB of copies per modem frame of 1024 streams)
  h->nthreads = getenv("RADE_B200_HOST_THREADS") ? atoi(getenv("RADE_B200_HOST_THREADS")) : 0;
  if (h->nthreads < 1) h->nthreads = 0;
  const size_t S = b->S, fb = S * RADE_NMF * sizeof(float2);
  const unsigned fl = cudaHostAllocMapped | cudaHostAllocPortable;
  bool ok = cudaHostAlloc((void **)&h->frames, HL_SLOTS * fb, fl) == cudaSuccess &&
            cudaHostGetDevicePointer((void **)&h->d_frames, h->frames, 0) == cudaSuccess &&
            cudaHostAlloc((void **)&h->active, S, fl) == cudaSuccess &&
            cudaMalloc((void **)&h->d_stage[0], fb) == cudaSuccess && cudaMalloc((void **)&h->d_stage[1], fb) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; i++)
    ok = cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&h->ev_appended[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { fprintf(stderr, "libradae_b200: cannot allocate the host sample link\n"); rade_b200_hostlink_close(h); return nullptr; }
  memset(h->frames, 0, HL_SLOTS * fb); memset(h->active, 0, S);
  return h;
}
RADE_EXPORT void rade_b200_hostlink_close(rade_b200_hostlink *h) {
  if (!h) return;
  cudaSetDevice(h->b->device);
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (int i = 0; i < 2; i++) {
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_appended[i]) cudaEventDestroy(h->ev_appended[i]);
    if (h->d_stage[i]) cudaFree(h->d_stage[i]);
  }
  if (h->frames) cudaFreeHost(h->frames);
  if (h->active) cudaFreeHost(h->active);
  delete h;
}
// samples [S][960] in ordinary host memory -> the next frame slot (host memcpy, OpenMP over streams).  Returns the number of
// streams whose frame was DROPPED because all slots were full (0 = queued, S = dropped), < 0 on error.
RADE_EXPORT int rade_b200_hostlink_push(rade_b200_hostlink *h, const RADE_COMP *samples) {
  const int S = h->b->S;
  if (h->pushed - h->popped >= HL_SLOTS) { h->dropped += S; return S; }       // the consumer is not keeping up
  float2 *dst = h->frames + (size_t)(h->pushed % HL_SLOTS) * S * RADE_NMF;
#pragma omp parallel for schedule(static) num_threads(h->nthreads ? h->nthreads : omp_get_max_threads())
  for (int s = 0; s < S; s++) memcpy(dst + (size_t)s * RADE_NMF, (const float2 *)samples + (size_t)s * RADE_NMF, RADE_NMF * sizeof(float2));
  __sync_synchronize();
  h->pushed = h->pushed + 1;
  return 0;
}
// transmit samples tx [S][960] (host) -> channel simulator of context `bch` (any context on the link's device with the same S; its
// channel configuration and state are used) -> the next frame slot: tx goes up with the copy engine, the channel kernel writes
// its output straight into the pinned slot.  Returns 0, or S when the frame was dropped because all slots were full; < 0 on error.
RADE_EXPORT int rade_b200_channel_hostlink(rade_batch *bch, rade_b200_hostlink *h, const RADE_COMP *tx) {
  cudaSetDevice(bch->device);
  if (bch->S != h->b->S || bch->device != h->b->device) { fprintf(stderr, "libradae_b200: channel_hostlink: contexts do not match\n"); return -1; }
  const size_t S = bch->S;
  if (h->pushed - h->popped >= HL_SLOTS) { h->dropped += (long long)S; return (int)S; }
  const rade_b200_channel_cfg &c = bch->chan_cfg;
  const float sigma = sqrtf((float)RADE_FS / (powf(10.f, c.EbNodB / 10.f) * 2000.f));
  CUDA_CHECK(cudaMemcpyAsync(bch->d_tx, tx, S * RADE_NMF * sizeof(float2), cudaMemcpyHostToDevice, bch->stream));
  const size_t slot_off = (size_t)(h->pushed % HL_SLOTS) * S * RADE_NMF;
  float2 *out = zero_copy_mode() ? h->d_frames + slot_off : bch->d_rx_in;      // in place over PCIe, or device buffer + copy engine
  bch->prof.begin(K_CHANNEL);
  if (channel_stream_launch(bch->tables, nullptr, out, bch->d_tx, bch->chan_state, bch->S, sigma, c.freq_offset_hz, c.freq_offset_spread_hz,
                            c.doppler_spread_hz, c.delay_samples, c.gain, c.seed, nullptr, nullptr, nullptr, nullptr, bch->stream) < 0) return -1;
  bch->prof.end(K_CHANNEL);
  bch->launches += 1;
  if (!zero_copy_mode()) CUDA_CHECK(cudaMemcpyAsync(h->frames + slot_off, bch->d_rx_in, S * RADE_NMF * sizeof(float2), cudaMemcpyDeviceToHost, bch->stream));
  CUDA_CHECK(cudaStreamSynchronize(bch->stream));
  __sync_synchronize();
  h->pushed = h->pushed + 1;
  return 0;
}
// every stream with >= nin[s] samples queued is advanced by exactly one rade_rx call (the others are left untouched, `active` = 0)
RADE_EXPORT int rade_b200_hostlink_rx(rade_b200_hostlink *h, float *features_out, int *ret, float *eoo_out) {
  rade_batch *b = h->b;
  cudaSetDevice(b->device);
  const size_t Sz = b->S, fb = Sz * RADE_NMF * sizeof(float2);
  const long long pushed = h->pushed;
  __sync_synchronize();
  // (1) host -> device copies of every queued frame not yet on its way (at most two in flight: two staging buffers)
  while (h->staged < pushed && h->staged - h->appended < 2) {
    const int sb = (int)(h->staged & 1);
    if (h->staged >= 2) CUDA_CHECK(cudaStreamWaitEvent(h->copy_stream, h->ev_appended[sb], 0));   // the staging buffer has been consumed
    CUDA_CHECK(cudaMemcpyAsync(h->d_stage[sb], h->frames + (size_t)(h->staged % HL_SLOTS) * Sz * RADE_NMF, fb, cudaMemcpyHostToDevice, h->copy_stream));
    CUDA_CHECK(cudaEventRecord(h->ev_copied[sb], h->copy_stream));
    h->staged++;
  }
  // (2) append ONE staged frame per call to the per-stream device rings (what a call consumes on average); the frames behind it
  // are already on their way up, so the next call finds its frame on the device
  long long newly = 0;
  while (h->appended < h->staged && h->appended < h->rx_calls + 1) {
    const int sb = (int)(h->appended & 1);
    CUDA_CHECK(cudaStreamWaitEvent(b->stream, h->ev_copied[sb], 0));
    b->prof.begin(K_LINK_PUSH);
    if (link_push_launch(b->link_ring, b->link_wr, b->link_rd, b->link_overflow, h->d_stage[sb], b->S, b->stream) < 0) return -1;
    b->prof.end(K_LINK_PUSH);
    CUDA_CHECK(cudaEventRecord(h->ev_appended[sb], b->stream));
    b->launches += 1;
    h->appended++; newly++;
  }
  // (3) the receiver on the rings
  const int reset_dec = (b->flags & RADE_USE_C_DECODER) ? 0 : 1;
  LinkSrc ls = {b->link_ring, b->link_wr, b->link_rd, b->d_active};
  int n = rx_dsp_launch(b->tables, b->rx, nullptr, nullptr, &ls, b->S, 1, reset_dec, b->d_ret, b->stream, &b->prof);
  if (n < 0) return -1;
  b->prof.begin(K_CORE_DEC);
  if (core_decoder_launch(b->weights.dev, b->rx.dec_state, b->rx.z_hat, b->d_feat_out, 1, b->rx.uw_errors, b->rx.dec_active,
                          b->S, RADE_NZMF, b->stream) < 0) return -1;
  b->prof.end(K_CORE_DEC);
  b->launches += n + 1;
  CUDA_CHECK(cudaMemcpyAsync(features_out, b->d_feat_out, Sz * RADE_NFEAT * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaMemcpyAsync(ret, b->d_ret, Sz * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (eoo_out) CUDA_CHECK(cudaMemcpyAsync(eoo_out, b->rx.eoo, Sz * RADE_NEOO_BITS * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaMemcpyAsync(h->active, b->d_active, Sz, cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  h->rx_calls++;
  __sync_synchronize();
  h->popped = h->appended;                                   // their copies have completed: the producer may reuse those slots
  return 0;
}
// pinned host memory for callers without a CUDA runtime of their own (C hosts, ctypes): buffers handed to the host-pointer entry
// points are copied by the copy engines asynchronously only when they are pinned — a pageable array is staged by the driver
RADE_EXPORT void *rade_b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  memset(p, 0, bytes);
  return p;
}
RADE_EXPORT void rade_b200_host_free(void *p) { if (p) cudaFreeHost(p); }
RADE_EXPORT const unsigned char *rade_b200_hostlink_active(rade_b200_hostlink *h) { return h->active; }
RADE_EXPORT long long rade_b200_hostlink_dropped(rade_b200_hostlink *h) { return h->dropped; }

// ---- the reference's `radae_tx | ch | radae_rx` pipe (three programs joined by pipes: src/radae_tx.c:14-55, the channel
// simulator, src/radae_rx.c:14-58) for S streams as ONE C call: three host threads, one context each — rade_b200_tx on `btx`,
// rade_b200_channel_hostlink on `bch`, rade_b200_hostlink_rx (the calling thread) on the link's receiver context — every call the
// synchronous host-buffer call a C host would make itself; the pipes are the n_tx_bufs pinned tx buffers and the link.
// features_in: n_in frames of [S][432] (cycled over); tx_bufs: n_tx_bufs (2..4) x [S][960], pinned memory is written in place;
// features_out [S][432] / ret [S]: the receiver's outputs for the LAST frame; valid_frames [S] (optional): += calls that returned
// features.  bch may equal btx: transmitter and channel then share a thread (two stages).
}  // extern "C"
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
namespace {
struct Sem {
  std::mutex m; std::condition_variable cv; int n;
  explicit Sem(int v) : n(v) {}
  void post() { { std::lock_guard<std::mutex> l(m); n++; } cv.notify_one(); }
  void wait() { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return n > 0; }); n--; }
};
}  // namespace
extern "C" {
RADE_EXPORT int rade_b200_duplex_run(rade_batch *btx, rade_batch *bch, rade_b200_hostlink *link, const float *features_in, int n_in,
                                     int n_frames, RADE_COMP *tx_bufs, int n_tx_bufs, float *features_out, int *ret, long long *valid_frames) {
  if (!btx || !bch || !link || !features_in || n_in < 1 || n_frames < 0 || !tx_bufs || n_tx_bufs < 1 || !features_out || !ret) return -1;
  if (n_tx_bufs > 4) n_tx_bufs = 4;
  const size_t S = btx->S;
  const bool three = bch != btx && n_tx_bufs >= 2;
  Sem tx_free(three ? n_tx_bufs : 1), tx_full(0), frames_free(HL_SLOTS - 1), frames_full(0);
  int tx_rc = 0, ch_rc = 0;
  auto tx_stage = [&](int k) {
    RADE_COMP *buf = tx_bufs + (size_t)(three ? k % n_tx_bufs : 0) * S * RADE_NMF;
    if (tx_rc == 0 && rade_b200_tx(btx, buf, features_in + (size_t)(k % n_in) * S * RADE_NFEAT) < 0) tx_rc = -1;
  };
  auto ch_stage = [&](int k) {
    RADE_COMP *buf = tx_bufs + (size_t)(three ? k % n_tx_bufs : 0) * S * RADE_NMF;
    if (ch_rc == 0 && tx_rc == 0) { const int r = rade_b200_channel_hostlink(bch, link, buf); if (r != 0) ch_rc = r < 0 ? -1 : -2; }
  };
  std::thread t_tx, t_ch;
  if (three) {
    t_tx = std::thread([&] { for (int k = 0; k < n_frames; k++) { tx_free.wait(); tx_stage(k); tx_full.post(); } });
    t_ch = std::thread([&] { for (int k = 0; k < n_frames; k++) { tx_full.wait(); frames_free.wait(); ch_stage(k); tx_free.post(); frames_full.post(); } });
  } else {
    t_ch = std::thread([&] { for (int k = 0; k < n_frames; k++) { frames_free.wait(); tx_stage(k); ch_stage(k); frames_full.post(); } });
  }
  int rc = 0;
  const bool trace = getenv("RADE_B200_DUPLEX_TRACE") != nullptr;
  double t_wait = 0, t_rx = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int k = 0; k < n_frames; k++) {
    const double t0 = now();
    frames_full.wait();
    const double t1 = now();
    if (rc == 0 && rade_b200_hostlink_rx(link, features_out, ret, nullptr) < 0) rc = -1;
    t_wait += t1 - t0; t_rx += now() - t1;
    if (rc == 0 && valid_frames)
      for (size_t s = 0; s < S; s++) valid_frames[s] += (ret[s] & 1);
    frames_free.post();
  }
  if (three) t_tx.join();
  t_ch.join();
  if (trace && n_frames > 0)
    fprintf(stderr, "rade_b200_duplex_run: receiver thread per frame: %.3f ms waiting for a frame, %.3f ms in rade_b200_hostlink_rx\n",
            1e3 * t_wait / n_frames, 1e3 * t_rx / n_frames);
  return (rc < 0 || tx_rc < 0 || ch_rc < 0) ? -1 : 0;
}

// ---- per-kernel timing (CUDA events on the context's stream).  enable=1 starts recording an event pair around every
// kernel launch; rade_b200_profile_read synchronises, returns per-kernel-class total milliseconds and launch counts
// (arrays of rade_b200_profile_n_kernels() entries, names via rade_b200_profile_kernel_name) and clears the record.
RADE_EXPORT int rade_b200_profile_enable(rade_batch *b, int enable) { b->prof.on = enable != 0; return 0; }
RADE_EXPORT int rade_b200_profile_n_kernels(void) { return K_COUNT; }
RADE_EXPORT const char *rade_b200_profile_kernel_name(int k) {
  static const char *names[K_COUNT] = {"core_encoder_kernel", "ofdm_mod_kernel", "eoo_kernel", "channel_stream_kernel",
                                       "link_push_kernel", "link_pop_kernel", "rx_bpf_kernel", "rx_detect_kernel",
                                       "rx_track_kernel", "rx_demod_kernel", "rx_finish_kernel", "core_decoder_kernel",
                                       "tx_bpf_clip_kernel", "rx_refresh_kernel"};
  return (k >= 0 && k < K_COUNT) ? names[k] : "";
}
RADE_EXPORT int rade_b200_profile_read(rade_batch *b, float *total_ms, int *counts) {
  cudaSetDevice(b->device);        // the current device is per host thread
  CUDA_CHECK(cudaStreamSynchronize(b->stream));
  for (int k = 0; k < K_COUNT; k++) {
    float tot = 0.f; int n = 0;
    std::vector<cudaEvent_t> &v = b->prof.ev[k];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, v[i], v[i + 1]) == cudaSuccess) { tot += ms; n++; }
    }
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
    total_ms[k] = tot; counts[k] = n;
  }
  return K_COUNT;
}

// ---- timeline of one or more steps: like the profiler above, but the event pairs are recorded on whichever stream launches the
// kernel and nothing is serialised, so the offsets show what really overlaps.  rade_b200_timeline_begin marks the origin;
// rade_b200_timeline_read synchronises and returns up to cap records {kernel id, start ms, end ms} (ms since the origin).
RADE_EXPORT int rade_b200_timeline_begin(rade_batch *b) {
  cudaSetDevice(b->device);
  for (int k = 0; k < K_COUNT; k++) { for (cudaEvent_t e : b->prof.ev[k]) cudaEventDestroy(e); b->prof.ev[k].clear(); }
  if (!b->prof.t0) CUDA_CHECK(cudaEventCreate(&b->prof.t0));
  CUDA_CHECK(cudaEventRecord(b->prof.t0, b->stream));
  b->prof.on = false; b->prof.timeline = true;
  return 0;
}
RADE_EXPORT int rade_b200_timeline_read(rade_batch *b, int *kernel, float *start_ms, float *end_ms, int cap) {
  cudaSetDevice(b->device);
  CUDA_CHECK(cudaDeviceSynchronize());
  b->prof.timeline = false;
  int n = 0;
  for (int k = 0; k < K_COUNT; k++) {
    auto &v = b->prof.ev[k];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
      if (n < cap) {
        kernel[n] = k;
        cudaEventElapsedTime(&start_ms[n], b->prof.t0, v[i]); cudaEventElapsedTime(&end_ms[n], b->prof.t0, v[i + 1]);
        n++;
      }
    }
    for (cudaEvent_t e : v) cudaEventDestroy(e);
    v.clear();
  }
  return n;
}

// debug/test hook: DSP tables as built on the host (no device needed) — lets the CPU test-suite compare them with
// the oracle's.  which: 0 Winv, 1 Wfwd, 2 p, 3 pend, 4 p_w, 5 Pmat, 6 eq_rot, 7 bpf_exp, 8 eoo_base (complex64);
// 16 bpf_h, 17 fcoarse, 18 {pilot_gain, bpf_bw, bpf_centre, bpf_alpha, residual of the coarse-grid basis}, 19 coarse-grid basis
// [80][16], 20 its expansion coefficients [21][12] (float32; layouts: AcqTables in rade_common.h).  Returns element count.
RADE_EXPORT int rade_b200_debug_tables(int which, float *out, int cap_floats) {
  DspTablesHost T; dsp_tables_host(T);
  const std::vector<std::complex<float>> *cv = nullptr; std::vector<float> fv;
  switch (which) {
    case 0: cv = &T.Winv; break; case 1: cv = &T.Wfwd; break; case 2: cv = &T.p; break; case 3: cv = &T.pend; break;
    case 4: cv = &T.p_w; break; case 5: cv = &T.Pmat; break; case 6: cv = &T.eq_rot; break; case 7: cv = &T.bpf_exp; break;
    case 8: cv = &T.eoo_base; break;
    case 16: fv = T.bpf_h; break; case 17: fv = T.fcoarse; break;
    case 18: fv = {(float)T.pilot_gain, T.bpf_bw, T.bpf_centre, T.bpf_alpha, (float)T.srch_residual}; break;
    case 19: fv = T.srch_basis; break; case 20: fv = T.srch_expand; break;
    default: return -1;
  }
  if (cv) {
    if ((int)cv->size() * 2 > cap_floats) return -1;
    memcpy(out, cv->data(), cv->size() * sizeof(std::complex<float>));
    return (int)cv->size();
  }
  if ((int)fv.size() > cap_floats) return -1;
  memcpy(out, fv.data(), fv.size() * sizeof(float));
  return (int)fv.size();
}

// debug/test hook (host only): the per-step weight streams of the encoder (which = 0) or decoder (1) built from the embedded
// weights: umma = 0 the mma.sync kernels' stream (fragment order), 1 the tcgen05 kernels' int8 stream (operand layout),
// 2 the tcgen05 kernels' float stream.  chunks_out receives (offset, bytes) pairs.  Returns the stream length in bytes
// (call with cap_bytes = 0 to size the buffers; *n_chunks is always set) or -1.
RADE_EXPORT long long rade_b200_debug_codec_stream(int which, int umma, unsigned char *bytes_out, long long cap_bytes,
                                                   unsigned int *chunks_out, int cap_chunks, int *n_chunks, int *n_prologue) {
  size_t len = 0;
  const void *blob = rade_b200_default_weights_blob(&len);
  std::vector<unsigned char> bytes; std::vector<ChunkDesc> chunks; int pro = 0;
  if (core_weights_debug_stream((const unsigned char *)blob, len, which, umma, &bytes, &chunks, &pro, nullptr) < 0) return -1;
  if (n_chunks) *n_chunks = (int)chunks.size();
  if (n_prologue) *n_prologue = pro;
  if (bytes_out && cap_bytes >= (long long)bytes.size()) memcpy(bytes_out, bytes.data(), bytes.size());
  if (chunks_out && cap_chunks >= (int)chunks.size())
    for (size_t i = 0; i < chunks.size(); i++) { chunks_out[2 * i] = chunks[i].offset; chunks_out[2 * i + 1] = chunks[i].bytes; }
  return (long long)bytes.size();
}
// the per-step MMA program of the tcgen05 kernels (one record of 13 ints per weight image, in issue order): a_off16, tile_step,
// b_kb, nk, n_tiles, b_buf, flags, d_blk, d_tile_stride, dep, commit, 0, 0.  Returns the record count or -1.
RADE_EXPORT int rade_b200_debug_codec_program(int which, int *ops_out, int cap_ops) {
  size_t len = 0;
  const void *blob = rade_b200_default_weights_blob(&len);
  std::vector<unsigned char> bytes; std::vector<ChunkDesc> chunks; std::vector<UmmaRec> recs; int pro = 0;
  if (core_weights_debug_stream((const unsigned char *)blob, len, which, 1, &bytes, &chunks, &pro, &recs) < 0) return -1;
  if (ops_out && cap_ops >= (int)recs.size())
    for (size_t i = 0; i < recs.size(); i++) {
      const UmmaRec &o = recs[i];
      const int v[13] = {o.a_off16, o.tile_step, o.b_kb, o.nk, o.n_tiles, o.b_buf, o.flags, o.d_blk, o.d_tile_stride, o.dep, o.commit, 0, 0};
      memcpy(ops_out + 13 * i, v, sizeof(v));
    }
  return (int)recs.size();
}

// debug: timeline of CTA 0 of the next codec launches (clock64 stamps, slot map in core_codec_umma.cu).  enable allocates and
// zeroes an n-slot device buffer; read copies it back (after a synchronize) and disables tracing.
RADE_EXPORT int rade_b200_debug_trace_enable(rade_batch *b, int n) {
  cudaSetDevice(b->device);
  long long *d = nullptr;
  CUDA_CHECK(cudaMalloc((void **)&d, sizeof(long long) * n));
  CUDA_CHECK(cudaMemset(d, 0, sizeof(long long) * n));
  b->allocs.push_back(d);
  b->weights.dev.trace = d;
  return 0;
}
RADE_EXPORT int rade_b200_debug_trace_read(rade_batch *b, long long *out, int n) {
  cudaSetDevice(b->device);
  if (!b->weights.dev.trace) return -1;
  CUDA_CHECK(cudaDeviceSynchronize());
  CUDA_CHECK(cudaMemcpy(out, b->weights.dev.trace, sizeof(long long) * n, cudaMemcpyDeviceToHost));
  b->weights.dev.trace = nullptr;
  return 0;
}

// debug/test hook (host only): would rade_b200_open accept this weight blob?  0 yes, -1 no (truncated / crafted / wrong shapes)
RADE_EXPORT int rade_b200_debug_check_weights(const void *weights, size_t weights_len) {
  return core_weights_validate((const unsigned char *)weights, weights_len);
}

// ================================================================== several GPUs from ONE C host
// The reference API is "single context only" (src/rade_api.h:87); streams are independent, so n_streams are split into
// contiguous blocks, one rade_batch per device, and every call is routed by block with one host thread per device.  The only
// data every device needs are the weights: the caller's host blob goes to each device once at open (inside one process that
// is a host -> device upload per GPU; across processes it is the one NCCL broadcast of the blob before rade_b200_open, see
// radae_b200/multigpu.py).  No per-frame exchange between devices.
struct rade_multi {
  int S, n;
  std::vector<rade_batch *> ctx;
  std::vector<int> first, count;
};
RADE_EXPORT rade_multi *rade_b200_open_devices(int n_streams, const int *devices, int n_devices, int flags, const void *weights, size_t weights_len) {
  if (n_streams <= 0 || !devices || n_devices <= 0 || n_devices > 64 || n_streams < n_devices) return nullptr;
  rade_multi *m = new rade_multi();
  m->S = n_streams; m->n = n_devices;
  for (int i = 0; i < n_devices; i++) {
    const int lo = (int)((long long)n_streams * i / n_devices), hi = (int)((long long)n_streams * (i + 1) / n_devices);
    rade_batch *b = rade_b200_open(hi - lo, devices[i], flags, weights, weights_len);
    if (!b) { for (rade_batch *c : m->ctx) rade_b200_close(c); delete m; return nullptr; }
    m->ctx.push_back(b); m->first.push_back(lo); m->count.push_back(hi - lo);
  }
  return m;
}
// device_mask: bit i = CUDA device i (0 = every visible device)
RADE_EXPORT rade_multi *rade_b200_open_multi(int n_streams, unsigned long long device_mask, int flags, const void *weights, size_t weights_len) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "libradae_b200: no CUDA device available — this library has no CPU fallback\n");
    return nullptr;
  }
  std::vector<int> devs;
  for (int i = 0; i < ndev && i < 64; i++) if (!device_mask || (device_mask >> i) & 1) devs.push_back(i);
  if (devs.empty()) { fprintf(stderr, "libradae_b200: device mask 0x%llx selects no visible device\n", device_mask); return nullptr; }
  return rade_b200_open_devices(n_streams, devs.data(), (int)devs.size(), flags, weights, weights_len);
}
RADE_EXPORT void rade_b200_close_multi(rade_multi *m) {
  if (!m) return;
  for (rade_batch *c : m->ctx) rade_b200_close(c);
  delete m;
}
RADE_EXPORT int rade_b200_multi_n_devices(rade_multi *m) { return m->n; }
RADE_EXPORT int rade_b200_multi_n_streams(rade_multi *m) { return m->S; }
RADE_EXPORT rade_batch *rade_b200_multi_context(rade_multi *m, int i, int *first_stream, int *n_streams) {
  if (i < 0 || i >= m->n) return nullptr;
  if (first_stream) *first_stream = m->first[i];
  if (n_streams) *n_streams = m->count[i];
  return m->ctx[i];
}
}  // extern "C"
namespace {
// run f(i) for every device on its own host thread (the calls are synchronous: this is what overlaps the devices)
template <typename F> int multi_each(rade_multi *m, F f) {
  std::vector<int> rc(m->n, 0);
  std::vector<std::thread> th;
  for (int i = 1; i < m->n; i++) th.emplace_back([&, i] { rc[i] = f(i); });
  rc[0] = f(0);
  for (auto &t : th) t.join();
  for (int r : rc) if (r < 0) return -1;
  return 0;
}
}  // namespace
extern "C" {
// rade_tx / rade_nin / rade_rx for all S streams: arrays as in the single-device calls ([S][432], [S][960], [S][1120], [S] ...)
RADE_EXPORT int rade_b200_multi_tx(rade_multi *m, RADE_COMP *tx_out, const float *features_in) {
  const int rc = multi_each(m, [&](int i) {
    return rade_b200_tx(m->ctx[i], tx_out + (size_t)m->first[i] * RADE_NMF, features_in + (size_t)m->first[i] * RADE_NFEAT);
  });
  return rc < 0 ? -1 : RADE_NMF;
}
RADE_EXPORT int rade_b200_multi_nin(rade_multi *m, int *nin) {
  return multi_each(m, [&](int i) { return rade_b200_nin(m->ctx[i], nin + m->first[i]); });
}
RADE_EXPORT int rade_b200_multi_rx(rade_multi *m, float *features_out, int *ret, float *eoo_out, const RADE_COMP *rx_in, const unsigned char *active) {
  return multi_each(m, [&](int i) {
    const size_t o = m->first[i];
    return rade_b200_rx(m->ctx[i], features_out + o * RADE_NFEAT, ret + o, eoo_out ? eoo_out + o * RADE_NEOO_BITS : nullptr,
                        rx_in + o * RADE_NIN_MAX, active ? active + o : nullptr);
  });
}
RADE_EXPORT int rade_b200_multi_rx_get_status(rade_multi *m, rade_b200_rx_status *status) {
  return multi_each(m, [&](int i) { return rade_b200_rx_get_status(m->ctx[i], status + m->first[i]); });
}

// ================================================================== rade_api.h: the reference's single-stream surface
struct rade {
  rade_batch *b;
  int flags;
  int nin, sync, snr;
  float freq_offset;
};

RADE_EXPORT void rade_initialize(void) {}      /* reference: starts CPython (src/rade_api.c:329-332); nothing to start here */
RADE_EXPORT void rade_finalize(void) {}

RADE_EXPORT struct rade *rade_open(char model_file[], int flags) {
  struct rade *r = new rade();
  r->flags = flags;
  std::vector<unsigned char> blob;
  if (model_file) {                             // honoured when it is a readable RDW / DNNw file, else embedded weights
    FILE *f = fopen(model_file, "rb");
    if (f) {
      fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
      if (n > 8) { blob.resize(n); if (fread(blob.data(), 1, n, f) != (size_t)n) blob.clear(); }
      fclose(f);
      if (blob.size() > 8 && memcmp(blob.data(), "RADEB200", 8) != 0 && memcmp(blob.data(), "DNNw", 4) != 0) blob.clear();
    }
  }
  if (!blob.empty()) {
    // a file with the right magic is used only if it is a complete RADE V1 model (a truncated file, or the reference's
    // model05 — no auxiliary symbol, bottleneck 1: a core-codec test model, not a V1 waveform — used to end in exit(1) with a
    // misleading message); anything else falls back to the embedded weights, as rade_api.h promises
    int in_dim = 0, out_dim = 0;
    const int ok = core_weights_validate(blob.data(), blob.size(), &in_dim, &out_dim);
    if (ok < 0 || in_dim != ENC_IN || out_dim != DEC_OUT) {
      fprintf(stderr, "libradae_b200: %s is %s; using the embedded model19_check3 weights\n", model_file,
              ok < 0 ? "not a complete RDW / DNNw weight file" : "not a RADE V1 model (no auxiliary symbol: rade_b200_core_encode / _decode can run it, the modem cannot)");
      blob.clear();
    }
  }
  if (!(flags & RADE_VERBOSE_0))
    fprintf(stderr, "libradae_b200: model: %s\n", blob.empty() ? "embedded model19_check3" : model_file);
  r->b = rade_b200_open(1, -1, flags, blob.empty() ? nullptr : blob.data(), blob.size());
  if (!r->b) {
    fprintf(stderr, "Error: libradae_b200 could not create a CUDA context; there is no CPU fallback\n");
    exit(1);                                    // same contract as the reference's check_error (src/rade_api.c:93-102)
  }
  r->nin = RADE_NMF; r->sync = 0; r->snr = 0; r->freq_offset = 0.f;
  return r;
}

RADE_EXPORT void rade_close(struct rade *r) {
  if (!r) return;
  rade_b200_close(r->b);
  delete r;
}

RADE_EXPORT int rade_version(void) { return 1; }                                      /* src/rade_api.c:37 */
RADE_EXPORT int rade_n_tx_out(struct rade *r) { (void)r; return RADE_NMF; }
RADE_EXPORT int rade_n_tx_eoo_out(struct rade *r) { (void)r; return RADE_NEOO; }
RADE_EXPORT int rade_nin_max(struct rade *r) { (void)r; return RADE_NIN_MAX; }
RADE_EXPORT int rade_n_features_in_out(struct rade *r) { (void)r; return RADE_NFEAT; }
RADE_EXPORT int rade_n_eoo_bits(struct rade *r) { (void)r; return RADE_NEOO_BITS; }

RADE_EXPORT int rade_tx(struct rade *r, RADE_COMP tx_out[], float features_in[]) {
  if (!r || !tx_out || !features_in) { fprintf(stderr, "rade_tx: NULL argument\n"); abort(); }   /* assert()s in the reference */
  if (rade_b200_tx(r->b, tx_out, features_in) < 0) { fprintf(stderr, "Error: rade_tx failed on the device\n"); exit(1); }
  return RADE_NMF;
}
RADE_EXPORT void rade_tx_set_eoo_bits(struct rade *r, float eoo_bits[]) {
  if (!r || !eoo_bits) { fprintf(stderr, "rade_tx_set_eoo_bits: NULL argument\n"); abort(); }
  if (rade_b200_tx_set_eoo_bits(r->b, eoo_bits) < 0) exit(1);
}
RADE_EXPORT int rade_tx_eoo(struct rade *r, RADE_COMP tx_eoo_out[]) {
  if (!r || !tx_eoo_out) { fprintf(stderr, "rade_tx_eoo: NULL argument\n"); abort(); }
  if (rade_b200_tx_eoo(r->b, tx_eoo_out) < 0) exit(1);
  return RADE_NEOO;
}
RADE_EXPORT int rade_nin(struct rade *r) { return r->nin; }

RADE_EXPORT int rade_rx(struct rade *r, float features_out[], int *has_eoo_out, float eoo_out[], RADE_COMP rx_in[]) {
  if (!r || !features_out || !rx_in) { fprintf(stderr, "rade_rx: NULL argument\n"); abort(); }
  rade_batch *b = r->b;
  // only nin samples are valid in the caller's array (src/rade_api.c:472)
  CUDA_CHECK_FATAL(cudaMemcpyAsync(b->d_rx_in, rx_in, (size_t)r->nin * sizeof(float2), cudaMemcpyHostToDevice, b->stream));
  if (rade_b200_rx_dev(b, b->d_feat_out, b->d_ret, nullptr, (const RADE_COMP *)b->d_rx_in, nullptr) < 0) exit(1);
  int ret = 0; RxCtl c;
  CUDA_CHECK_FATAL(cudaMemcpyAsync(&ret, b->d_ret, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK_FATAL(cudaMemcpyAsync(&c, b->rx.ctl, sizeof(RxCtl), cudaMemcpyDeviceToHost, b->stream));
  CUDA_CHECK_FATAL(cudaStreamSynchronize(b->stream));
  const int valid = ret & 1, eoo = ret & 2;
  if (valid) CUDA_CHECK_FATAL(cudaMemcpy(features_out, b->d_feat_out, RADE_NFEAT * sizeof(float), cudaMemcpyDeviceToHost));
  if (has_eoo_out) *has_eoo_out = 0;
  if (eoo) {
    if (eoo_out) CUDA_CHECK_FATAL(cudaMemcpy(eoo_out, b->rx.eoo, RADE_NEOO_BITS * sizeof(float), cudaMemcpyDeviceToHost));
    if (has_eoo_out) *has_eoo_out = 1;
  }
  r->nin = c.nin; r->sync = (c.state == 2); r->snr = (int)c.snr_est; r->freq_offset = (float)c.fmax;   /* src/rade_api.c:528-530 */
  return valid ? RADE_NFEAT : 0;
}
RADE_EXPORT int rade_sync(struct rade *r) { return r->sync; }
/* the reference returns 0 here ("TODO: we need a float getter", src/rade_api.c:547-550); we return the tracked offset */
RADE_EXPORT float rade_freq_offset(struct rade *r) { return r->freq_offset; }
RADE_EXPORT int rade_snrdB_3k_est(struct rade *r) { return r->snr; }

}  // extern "C"
