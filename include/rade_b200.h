/* rade_b200.h — batched multi-stream extension of the rade_api.h C ABI (libradae_b200).
 *
 * The reference API is "single context only" (src/rade_api.h:87); one B200 carries thousands of independent
 * streams, so this header adds a context that owns S streams on one CUDA device.  Semantics per stream are
 * exactly those of rade_tx / rade_rx / rade_nin / ... (src/rade_api.c:403-555), arrays simply gain a leading
 * stream dimension.  Plain pointers and sizes only; `_dev` entry points take DEVICE pointers and enqueue work
 * on the context's CUDA stream without synchronising, the others take HOST pointers, copy in/out and synchronise.
 * All functions return 0 (or a count) on success and a negative value on error, after printing the CUDA error.
 */
#ifndef RADE_B200_H
#define RADE_B200_H
#include <stddef.h>
#include "rade_api.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rade_batch rade_batch;

/* weights: NULL/0 -> the RDW blob embedded in the library; else an RDW v1 or DNNw blob in host memory
 * (what a caller receives from the one-off ncclBroadcast at start-up).  device < 0 -> current device.
 * flags: the rade_api.h flags, plus RADE_B200_BOTTLENECK_1 for core-codec weights trained with bottleneck 1 (tanh on the
 * latents, src/rade_enc.c:107-113): the reference's second shipped model, bin/model05.bin, as driven by
 * `test_rade_enc 1 0 model05.bin` / `test_rade_dec 0 model05.bin` (CMakeLists.txt:519-545).  Such a model has no auxiliary
 * symbol, so rade_b200_core_encode / _decode exchange 80-float rows with the host (rade_b200_core_dims); the modem entry
 * points (rade_b200_tx / _rx ...) are defined for the RADE V1 waveform (model19_check3) only. */
#define RADE_B200_BOTTLENECK_1 0x100
RADE_EXPORT rade_batch *rade_b200_open(int n_streams, int device, int flags, const void *weights, size_t weights_len);
RADE_EXPORT void rade_b200_close(rade_batch *b);
RADE_EXPORT int rade_b200_n_streams(rade_batch *b);
RADE_EXPORT void *rade_b200_cuda_stream(rade_batch *b);                 /* cudaStream_t */
RADE_EXPORT int rade_b200_synchronize(rade_batch *b);
RADE_EXPORT long long rade_b200_launch_count(rade_batch *b);            /* kernels launched so far by this context */
RADE_EXPORT const void *rade_b200_default_weights_blob(size_t *len);    /* the embedded RDW blob (host memory) */
RADE_EXPORT int rade_b200_reset(rade_batch *b);                         /* all streams back to the rade_open() state */

/* --- core codec only (rade_core_encoder / rade_core_decoder, src/rade_core.h:42-49), n_steps 40 ms steps ---
 * features [S][n_steps][84] (4 x (20 features + aux)), z [S][n_steps][80] */
/* row widths of the loaded model on the HOST side of rade_b200_core_encode / _decode: 84 / 84 (model19_check3) or 80 / 80
 * (a model without the aux symbol); the _dev entry points always use 84-float rows (unused inputs zero, unused outputs 0) */
RADE_EXPORT int rade_b200_core_dims(rade_batch *b, int *input_dim, int *output_dim);
RADE_EXPORT int rade_b200_core_encode_dev(rade_batch *b, float *d_z, const float *d_features, int n_steps);
RADE_EXPORT int rade_b200_core_decode_dev(rade_batch *b, float *d_features, const float *d_z, int n_steps);
RADE_EXPORT int rade_b200_core_encode(rade_batch *b, float *z, const float *features, int n_steps);
RADE_EXPORT int rade_b200_core_decode(rade_batch *b, float *features, const float *z, int n_steps);

/* --- transmitter: rade_tx / rade_tx_set_eoo_bits / rade_tx_eoo per stream ---
 * features_in [S][432], tx_out [S][960], eoo_bits [S][180] (+-1), tx_eoo_out [S][1152] */
RADE_EXPORT int rade_b200_tx_dev(rade_batch *b, RADE_COMP *d_tx_out, const float *d_features_in);
RADE_EXPORT int rade_b200_tx(rade_batch *b, RADE_COMP *tx_out, const float *features_in);
RADE_EXPORT int rade_b200_tx_set_eoo_bits(rade_batch *b, const float *eoo_bits);
RADE_EXPORT int rade_b200_tx_eoo(rade_batch *b, RADE_COMP *tx_eoo_out);
/* OFDM modulator alone (transmitter_one, radae/dsp.py:340-378): z [S][3][80] -> tx [S][960] */
RADE_EXPORT int rade_b200_ofdm_mod_dev(rade_batch *b, RADE_COMP *d_tx_out, const float *d_z);
/* radae_tx(bypass_enc=True).do_radae_tx (radae_txe.py:122-132; what src/rade_api.c:411-436 calls after its own C encoder):
 * z [S][3][80] from the caller's core encoder -> modulator (+ the TX filter below when enabled) -> tx [S][960] */
RADE_EXPORT int rade_b200_tx_z_dev(rade_batch *b, RADE_COMP *d_tx_out, const float *d_z);
RADE_EXPORT int rade_b200_tx_z(rade_batch *b, RADE_COMP *tx_out, const float *z);
/* radae_tx(txbpf_en=True) (radae_txe.py:74-81, :130-132, :141-143): 101-tap complex band-pass filter (the receive
 * filter's band, state carried from call to call) + clip to unit magnitude on every frame rade_b200_tx[_dev],
 * rade_b200_tx_z[_dev] and rade_b200_tx_eoo produce.  Off by default (src/rade_api.c never sets it); enabling or
 * disabling restarts the filter.  The fused loop-back calls (tx_channel_link_dev, loopback_step_dev) refuse to run
 * while it is enabled. */
RADE_EXPORT int rade_b200_tx_bpf_enable(rade_batch *b, int enable);

/* --- receiver: rade_nin / rade_rx / rade_sync / rade_snrdB_3k_est per stream ---
 * rx_in [S][1120]: row s holds nin[s] fresh samples (800 | 960 | 1120);  features_out [S][432];
 * ret [S] = valid | eoo << 1 (radae_rxe.py:330);  eoo_out [S][180];  active [S] (optional, NULL = all):
 * streams with active == 0 are not advanced at all. */
RADE_EXPORT int rade_b200_nin(rade_batch *b, int *nin);
RADE_EXPORT int rade_b200_rx(rade_batch *b, float *features_out, int *ret, float *eoo_out, const RADE_COMP *rx_in,
                             const unsigned char *active);
RADE_EXPORT int rade_b200_rx_dev(rade_batch *b, float *d_features_out, int *d_ret, float *d_eoo_out,
                                 const RADE_COMP *d_rx_in, const unsigned char *d_active);
RADE_EXPORT const int *rade_b200_nin_dev(rade_batch *b);                /* device array [S], valid after rx_dev */

typedef struct {              /* one per stream; what the reference prints per frame at -v 2 (radae_rxe.py:239-246) */
  int state;                  /* 0 search, 1 candidate, 2 sync */
  int nin, tmax, valid_count, uw_errors, synced_count;
  int snrdB_3k_est;           /* int(), like get_snrdB_3k_est (radae_rxe.py:162-163) */
  float snrdB_3k_est_f;
  double fmax;
  float Dthresh, Dtmax12, Dtmax12_eoo;
} rade_b200_rx_status;
RADE_EXPORT int rade_b200_rx_get_status(rade_batch *b, rade_b200_rx_status *status /* [S] host */);
/* z_hat of the last rade_b200_rx call, [S][240] (what bypass_dec hands to the C decoder, radae_rxe.py:318) */
RADE_EXPORT int rade_b200_rx_get_z_hat(rade_batch *b, float *z_hat);

/* --- channel simulator: rate-Fs branch of RADAE.forward (radae/radae.py:529-599), per stream ---
 * explicit form (parity tests): rx = gain*(mp_gain*(tx*G1 + delay_d(tx*G2))*exp(j(phase0+2*pi*f*(n+1)/Fs)) + sigma*noise)
 * all arrays [S][n] complex64 on the device. */
RADE_EXPORT int rade_b200_channel_apply_dev(rade_batch *b, RADE_COMP *d_rx, const RADE_COMP *d_tx, const RADE_COMP *d_G1,
                                            const RADE_COMP *d_G2, const RADE_COMP *d_noise, int n, int delay,
                                            float mp_gain, float freq_offset_hz, float phase0, float sigma, float gain);
/* the same with host arrays [S][n] (e.g. G1, G2 read from one of the reference's rate-Fs fading files, g_mpp.f32 ...) */
RADE_EXPORT int rade_b200_channel_apply(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx, const RADE_COMP *G1, const RADE_COMP *G2,
                                        const RADE_COMP *noise, int n, int delay, float mp_gain, float freq_offset_hz, float phase0,
                                        float sigma, float gain);
typedef struct {
  float EbNodB;               /* sigma = sqrt(Fs/(EbNo*Rb)), Rb = 2000 (radae.py:570-574) */
  float freq_offset_hz;       /* per stream: freq_offset_hz + U(-1,1)*freq_offset_spread_hz */
  float freq_offset_spread_hz;
  float doppler_spread_hz;    /* 0 = AWGN only (G1 = 1, G2 = 0); MPP = 1.0 */
  int delay_samples;          /* 16 = 2 ms (MPP) */
  float gain;
  unsigned long long seed;
} rade_b200_channel_cfg;
/* streaming generator form: consumes tx [S][960] (one modem frame per stream), Philox AWGN + two-path Watterson
 * fading with Gaussian Doppler spectrum generated in-kernel; keeps per-stream phase / delay-line / time state */
RADE_EXPORT int rade_b200_channel_config(rade_batch *b, const rade_b200_channel_cfg *cfg);
RADE_EXPORT int rade_b200_channel_dev(rade_batch *b, RADE_COMP *d_rx, const RADE_COMP *d_tx);

/* ... with the reference's frequency drift df_dt in Hz/s (radae.py:546-550): phase[k] = phase0 + 2 pi / Fs (f (k + 1) + df_dt k (k + 1) / (2 Fs)) */
RADE_EXPORT int rade_b200_channel_apply_drift(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx, const RADE_COMP *G1, const RADE_COMP *G2,
                                              const RADE_COMP *noise, int n, int delay, float mp_gain, float freq_offset_hz, float df_dt,
                                              float phase0, float sigma, float gain);

/* --- loop-back link between channel output and receiver input (per-stream sample FIFO on the device): the
 * receiver consumes nin[s] in {800, 960, 1120} samples per call while the transmitter produces 960 --- */
RADE_EXPORT int rade_b200_link_push_dev(rade_batch *b, const RADE_COMP *d_samples /* [S][960] */);
/* the same loop-back without the two copy kernels: channel output goes straight into the link FIFOs, the receiver pops
   nin[s] samples per stream from them (a stream without nin[s] queued samples sits the call out: ret = 0) */
RADE_EXPORT int rade_b200_channel_link_dev(rade_batch *b, const RADE_COMP *d_tx /* [S][960] */);
/* software pipeline over frames for device-pointer callers: with it enabled, rade_b200_tx_dev and rade_b200_channel_link_dev
   enqueue on a second stream, so frame k+1's transmitter runs concurrently with frame k's receiver.  A step is
   fork; tx_dev(k+1); channel_link_dev; rx_link_dev(k); join.  fork: the TX stream waits for the main stream; join: the reverse. */
RADE_EXPORT int rade_b200_pipeline_enable(rade_batch *b, int enable);
RADE_EXPORT int rade_b200_pipeline_fork(rade_batch *b);
RADE_EXPORT int rade_b200_pipeline_join(rade_batch *b);
/* the whole loop-back step in one call: tx_dev(d_features_next) -> channel_link_dev -> rx_link_dev (+ fork / join when the frame
   pipeline is enabled); with RADE_B200_GRAPH=1 in the environment it is replayed as a single CUDA graph launch per step */
RADE_EXPORT int rade_b200_loopback_step_dev(rade_batch *b, const float *d_features_next /* [S][432] */, float *d_features_out,
                                            int *d_ret, float *d_eoo_out);
/* the same pipeline for n_frames modem frames with HOST buffers at both ends: per frame S x 432 input features go up from (pinned) host
   memory, S x 432 recovered features and S return codes come back; the modem samples stay on the device.  Uploads / downloads run on
   their own streams, double-buffered, and overlap the kernels of the neighbouring frames.  features_in: n_in frames of [S][432], cycled
   over; features_out [S][432] / ret [S]: overwritten every frame; valid_frames [S] (optional): += frames that returned features. */
RADE_EXPORT int rade_b200_loopback_run(rade_batch *b, const float *features_in, int n_in, int n_frames, float *features_out, int *ret,
                                       long long *valid_frames);
/* core encoder, then one kernel that modulates the frame and sends it through the channel into the link FIFOs */
RADE_EXPORT int rade_b200_tx_channel_link_dev(rade_batch *b, const float *d_features_in /* [S][432] */);
RADE_EXPORT int rade_b200_rx_link_dev(rade_batch *b, float *d_features_out, int *d_ret, float *d_eoo_out);
RADE_EXPORT int rade_b200_link_pop_dev(rade_batch *b, RADE_COMP *d_rx_in /* [S][1120] */, unsigned char *d_active /* [S] */);

/* --- per-kernel device timing (CUDA events on the context's stream; used by bench.py for the roofline line) --- */
RADE_EXPORT int rade_b200_profile_enable(rade_batch *b, int enable);
RADE_EXPORT int rade_b200_profile_n_kernels(void);
RADE_EXPORT const char *rade_b200_profile_kernel_name(int k);
RADE_EXPORT int rade_b200_profile_read(rade_batch *b, float *total_ms, int *counts);
/* timeline of the kernels of the steps issued between _begin and _read: event pairs on the launching streams, nothing serialised;
 * records {kernel id (rade_b200_profile_kernel_name), start, end} in ms since _begin; returns the number of records */
RADE_EXPORT int rade_b200_timeline_begin(rade_batch *b);
RADE_EXPORT int rade_b200_timeline_read(rade_batch *b, int *kernel, float *start_ms, float *end_ms, int cap);

/* --- host-side sample link in front of rade_b200_rx: a ring of 4 modem-frame slots [S][960] in pinned host
 * memory that producers fill, per-stream sample rings on the device that the receiver consumes; queued frames go up with the
 * copy engine, issued ahead so that they overlap the previous call's kernels.
 * push: samples [S][960] from an ordinary host array (host memcpy); returns 0, or S when the frame was DROPPED because all slots
 * were full.  channel_hostlink: tx [S][960] (host) -> channel simulator of `bch` (any context on the same device with the same
 * S) -> the next slot, written in place by the kernel; returns like push.  rx: every stream with >= nin[s] samples queued is
 * advanced by exactly one rade_rx call (the others are left untouched, `active` = 0); outputs as rade_b200_rx.
 * One producer and one consumer per link (they may be different host threads with a context each). --- */
typedef struct rade_b200_hostlink rade_b200_hostlink;
RADE_EXPORT rade_b200_hostlink *rade_b200_hostlink_open(rade_batch *b, int capacity_samples /* ignored */);
RADE_EXPORT void rade_b200_hostlink_close(rade_b200_hostlink *h);
RADE_EXPORT int rade_b200_hostlink_push(rade_b200_hostlink *h, const RADE_COMP *samples);
RADE_EXPORT int rade_b200_channel_hostlink(rade_batch *bch, rade_b200_hostlink *h, const RADE_COMP *tx);
RADE_EXPORT int rade_b200_hostlink_rx(rade_b200_hostlink *h, float *features_out, int *ret, float *eoo_out);
RADE_EXPORT const unsigned char *rade_b200_hostlink_active(rade_b200_hostlink *h);
RADE_EXPORT long long rade_b200_hostlink_dropped(rade_b200_hostlink *h);
/* pinned (page-locked) host memory for the arrays handed to the host-pointer entry points: only pinned arrays are moved by the copy
 * engines asynchronously, pageable ones are staged by the driver */
RADE_EXPORT void *rade_b200_host_alloc(size_t bytes);
RADE_EXPORT void rade_b200_host_free(void *p);
/* the reference's `radae_tx | ch | radae_rx` pipe (src/radae_tx.c:14-55, src/radae_rx.c:14-58) for S streams as one C call: three
 * host threads with a context each — rade_b200_tx on `btx`, rade_b200_channel_hostlink on `bch`, rade_b200_hostlink_rx (the calling
 * thread) on the link's receiver context; every call is the synchronous host-buffer call, the pipes are the tx buffers and the link.
 * features_in: n_in frames of [S][432] (cycled over); tx_bufs: n_tx_bufs (2..4) x [S][960] (pinned memory is written in place by the
 * modulator); features_out [S][432] / ret [S]: the receiver's outputs for the last frame; valid_frames [S] (optional): += calls
 * that returned features.  With bch == btx transmitter and channel share one thread. */
RADE_EXPORT int rade_b200_duplex_run(rade_batch *btx, rade_batch *bch, rade_b200_hostlink *link, const float *features_in, int n_in,
                                     int n_frames, RADE_COMP *tx_bufs, int n_tx_bufs, float *features_out, int *ret, long long *valid_frames);

/* --- host-buffer channel call (for end-to-end measurements through host memory): tx, rx [S][960] --- */
RADE_EXPORT int rade_b200_channel(rade_batch *b, RADE_COMP *rx, const RADE_COMP *tx);

/* --- several GPUs from ONE C host (replaces "single context only", src/rade_api.h:87): n_streams are split
 * into contiguous blocks, one context per device, every call is routed by block with one host thread per device.  The weights
 * are the only data every device needs: the host blob goes to each device once at open (across processes: one NCCL broadcast
 * of the blob before rade_b200_open, radae_b200/multigpu.py); there is no per-frame exchange between devices.
 * device_mask: bit i = CUDA device i, 0 = every visible device.  rade_b200_open_devices takes an explicit device list (a device
 * may appear more than once: several contexts on one GPU).  Arrays are those of the single-device calls with S = n_streams. */
typedef struct rade_multi rade_multi;
RADE_EXPORT rade_multi *rade_b200_open_multi(int n_streams, unsigned long long device_mask, int flags, const void *weights, size_t weights_len);
RADE_EXPORT rade_multi *rade_b200_open_devices(int n_streams, const int *devices, int n_devices, int flags, const void *weights, size_t weights_len);
RADE_EXPORT void rade_b200_close_multi(rade_multi *m);
RADE_EXPORT int rade_b200_multi_n_devices(rade_multi *m);
RADE_EXPORT int rade_b200_multi_n_streams(rade_multi *m);
RADE_EXPORT rade_batch *rade_b200_multi_context(rade_multi *m, int i, int *first_stream, int *n_streams);   /* the i-th device's context */
RADE_EXPORT int rade_b200_multi_tx(rade_multi *m, RADE_COMP *tx_out, const float *features_in);
RADE_EXPORT int rade_b200_multi_nin(rade_multi *m, int *nin);
RADE_EXPORT int rade_b200_multi_rx(rade_multi *m, float *features_out, int *ret, float *eoo_out, const RADE_COMP *rx_in, const unsigned char *active);
RADE_EXPORT int rade_b200_multi_rx_get_status(rade_multi *m, rade_b200_rx_status *status);

#ifdef __cplusplus
}
#endif
#endif
