/* sedb200.h -- C ABI of libsedb200.so: the sm_100a (B200) implementation of the sed-crnn hot path.
 *
 * The reference (noamzilo/sed-crnn) is pure Python and has NO native ABI, plugin registry or FFI of
 * its own.  Each entry point below therefore cites the reference *Python*
 * interface it stands behind; INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every function returns 0 (SEDB200_OK) or a negative SEDB200_E* code; the message for the last
 *     failure on the calling thread is returned by sedb200_last_error().
 *   - pointers named *_dev are DEVICE pointers on the current CUDA device; the library never
 *     allocates or frees caller-visible memory (the host framework -- PyTorch -- owns every buffer).
 *     The only state it keeps is a set of immutable constant tables per device (Hann window,
 *     twiddles, mel weights), created lazily.
 *   - `stream` is a cudaStream_t / CUstream passed as void*; all work is enqueued on it and the call
 *     returns without synchronising (except the *_host convenience calls, documented below).
 *   - there is NO CPU fallback: on a device that is not compute capability 10.x every compute call
 *     returns SEDB200_EARCH.
 */
#ifndef SEDB200_H
#define SEDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEDB200_VERSION 100            /* 0.1.0 */

#define SEDB200_OK          0
#define SEDB200_EINVAL     -1          /* bad argument */
#define SEDB200_ESHAPE     -2          /* shape the kernels do not support */
#define SEDB200_EWORKSPACE -3          /* workspace too small */
#define SEDB200_ECUDA      -4          /* CUDA runtime error (message has the cudaError string) */
#define SEDB200_EARCH      -5          /* device is not sm_100 */

/* feature.py:29-32 */
#define SEDB200_NFFT   2048
#define SEDB200_HOP    1024
#define SEDB200_NMEL   40

#define SEDB200_PAD_CONSTANT 0         /* librosa >= 0.10 stft default */
#define SEDB200_PAD_REFLECT  1         /* librosa <  0.10 stft default */

int         sedb200_version(void);
const char* sedb200_last_error(void);
/* SEDB200_OK if `device` (or the current device when -1) is sm_100; SEDB200_EARCH otherwise. */
int         sedb200_device_check(int device);
/* number of CUDA kernels this library has launched in this process (for the bench's gpu_launches). */
long        sedb200_launch_count(void);
/* Phase profiler: when enabled, every phase of the CRNN calls is bracketed by CUDA events on the
 * launching stream.  sedb200_prof_report synchronises those events and writes one line per phase,
 * "<name> <total_ms> <count>\n", into buf.  sedb200_prof_enable (on or off) clears the record. */
int         sedb200_prof_enable(int on);
int         sedb200_prof_report(char* buf, size_t buf_bytes);

/* ------------------------------------------------------------------------------------------------
 * Log-mel front end.   Replaces  feature._mbe(y, sr)  (/root/reference/feature.py:55-59):
 *   librosa.stft(n_fft=2048, hop=1024, centre-padded, periodic Hann) -> |X|^2 -> 40-band Slaney mel
 *   -> natural log -> [frames, 40].
 *
 * pcm_dev : float32 [n_clips][n_ch][n_samples]   (mono reference call: n_clips = n_ch = 1)
 * out_dev : float32 [n_clips][frames][n_ch*40],  frames = 1 + n_samples/1024; channel-major on the
 *           feature axis (the layout /root/reference/utils.py:15-25 split_multi_channels undoes).
 * sr      : sample rate used for the mel filterbank (feature.py:58 passes sr=44100).
 * Silence gives log(0) = -inf exactly like feature.py:59 (no floor).
 */
long   sedb200_logmel_frames(long n_samples);
int    sedb200_logmel_f32(const float* pcm_dev, int n_clips, int n_ch, long n_samples,
                          int sr, int pad_mode, float* out_dev, void* stream);

/* Same, HOST buffers: H2D copy of pcm, kernel, D2H copy of the result, all on `stream`, then a
 * stream synchronise (this is the call feature.py:84 would make per clip).  `scratch_dev` is a
 * caller-owned device buffer of at least sedb200_logmel_host_scratch() bytes. */
size_t sedb200_logmel_host_scratch(int n_clips, int n_ch, long n_samples);
int    sedb200_logmel_host_f32(const float* pcm_host, int n_clips, int n_ch, long n_samples,
                               int sr, int pad_mode, float* out_host,
                               void* scratch_dev, size_t scratch_bytes, void* stream);

/* int16 PCM ingest ("int16/f32 PCM ingest so a real decoder can feed the kernel"): the same
 * kernel reading 16-bit samples as a WAV file or `ffmpeg -f s16le` delivers them (feature.py:40-50 asks ffmpeg for
 * f32le).  A sample s stands for the float32 value s / 32768 -- exact in fp32 -- so the result is bit-identical to
 * sedb200_logmel_f32 on the converted signal; host -> device traffic is halved. */
int    sedb200_logmel_i16(const short* pcm_dev, int n_clips, int n_ch, long n_samples,
                          int sr, int pad_mode, float* out_dev, void* stream);
size_t sedb200_logmel_host_scratch_i16(int n_clips, int n_ch, long n_samples);
int    sedb200_logmel_host_i16(const short* pcm_host, int n_clips, int n_ch, long n_samples,
                               int sr, int pad_mode, float* out_host,
                               void* scratch_dev, size_t scratch_bytes, void* stream);

/* The same two calls with the kernel named.  Two kernels implement the front end:
 *   SEDB200_LOGMEL_FP32 : one warp per frame, 32 x 32 register FFT in fp32 on the CUDA cores (csrc/logmel.cu);
 *   SEDB200_LOGMEL_TC   : the 2048-point real DFT as two batched tcgen05 GEMMs (64-point real stage over rows
 *                         (frame, n2), 32-point complex stage over rows (frame, k1)) on fp16 hi / lo operand planes
 *                         with fp32 accumulation in TMEM (csrc/logmel_tc.cu); same result to ~1e-6 relative.
 * SEDB200_LOGMEL_AUTO picks the build's default (sedb200_logmel_default_kernel(); the environment variable
 * SEDB200_LOGMEL_KERNEL = "fp32" | "tc" overrides it), which is what sedb200_logmel_f32 / _i16 / _host_* run. */
#define SEDB200_LOGMEL_AUTO 0
#define SEDB200_LOGMEL_FP32 1
#define SEDB200_LOGMEL_TC   2
#define SEDB200_LOGMEL_DEFAULT_KERNEL SEDB200_LOGMEL_FP32
int    sedb200_logmel_f32_k(const float* pcm_dev, int n_clips, int n_ch, long n_samples,
                            int sr, int pad_mode, float* out_dev, void* stream, int kernel);
int    sedb200_logmel_i16_k(const short* pcm_dev, int n_clips, int n_ch, long n_samples,
                            int sr, int pad_mode, float* out_dev, void* stream, int kernel);
int    sedb200_logmel_default_kernel(void);

/* Per-bin standardisation -- the step between the two halves of the hot path (feature.py:127-129:
 * sklearn.preprocessing.StandardScaler().fit_transform(X_train) / .transform(X_test)).
 * fit:   mean[c], var[c] (population variance), scale[c] = sqrt(var) (1.0 for constant columns) of x [rows][cols],
 *        accumulated deterministically, float64 results (sklearn's mean_ / var_ / scale_ dtypes).
 * apply: out = (x - mean) / scale, float32. */
size_t sedb200_standardize_scratch_bytes(long rows, int cols);
int    sedb200_standardize_fit(const float* x_dev, long rows, int cols, double* mean_dev, double* var_dev,
                               double* scale_dev, void* scratch_dev, size_t scratch_bytes, void* stream);
int    sedb200_standardize_apply(const float* x_dev, long rows, int cols, const double* mean_dev,
                                 const double* scale_dev, float* out_dev, void* stream);

/* Copy of the float32 [40][1025] mel filterbank the kernel uses (== librosa.filters.mel(sr=sr,
 * n_fft=2048, n_mels=40), feature.py:58) into a HOST buffer; for inspection / tests. */
int    sedb200_mel_filterbank(int sr, float* out_host /* [40*1025] */);


/* ------------------------------------------------------------------------------------------------
 * CRNN (conv blocks -> BiGRU stack -> per-frame dense -> logits), forward / loss / backward / Adam.
 *
 * Replaces the arithmetic of
 *   crnn_lightning.py:41-73   TimePooledCRNN.forward          (Lightning variant)
 *   sed.py:82-112             TimePooledCRNN.forward          (plain-torch variant)
 *   crnn_lightning.py:27-35   FocalBCELoss.forward            sed.py:160 BCEWithLogitsLoss
 *   loss.backward()           crnn_lightning.py:157-163 (Lightning), sed.py:137
 *   clip_grad_norm_(1.0)+Adam train_lightning.py:50, crnn_lightning.py:195-197, sed.py:159
 * for the hyper-parameters of train_constants.py:6-28 and of the BASELINE.json configs.
 *
 * Layouts
 *   x        float32 [B][in_ch][H][W]  (NCHW, exactly what the reference feeds its Conv2d).
 *            mode 0 ("fork",   crnn_lightning.py:66-70): H = mel, W = time; pooling (1,p) shrinks TIME;
 *                     sequence axis = pooled W, features flattened as c*H + h.
 *            mode 1 ("sednet", the BASELINE configs)   : H = time, W = mel; pooling (1,p) shrinks MEL;
 *                     sequence axis = H, features flattened as c*W' + w.
 *   logits   float32 [B][T][n_classes]
 *   params   one flat float32 buffer; tensor i lives at offsets[i] (sedb200_crnn_param_layout), in
 *            PyTorch's own shapes: conv weight [C][Cin][3][3], GRU weight_ih [3H][in] with gate order
 *            r,z,n, Linear weight [out][in].  Order: per conv block {conv.weight, conv.bias, bn.weight,
 *            bn.bias}; per GRU layer {w_ih[2 dirs], w_hh[2], b_ih[2], b_hh[2]}; per dense {weight, bias}.
 *   bn_state float32: per conv block {running_mean[C], running_var[C]}, contiguous.
 *   grads    same layout as params.
 */
#define SEDB200_MAX_CONV  4
#define SEDB200_MAX_GRU   4
#define SEDB200_MAX_DENSE 3

#define SEDB200_LOSS_BCE   0           /* sed.py:160 */
#define SEDB200_LOSS_FOCAL 1           /* crnn_lightning.py:27-35 */

typedef struct sedb200_crnn_desc {
    int   mode;                        /* 0 fork, 1 sednet */
    int   in_ch, H, W;
    int   n_conv, conv_ch;
    int   pool[SEDB200_MAX_CONV];
    int   n_gru;
    int   gru_units[SEDB200_MAX_GRU];
    int   n_dense;                     /* including the final n_classes layer */
    int   dense_units[SEDB200_MAX_DENSE];
    int   dense_relu;                  /* ReLU after every dense layer but the last (crnn_lightning.py:72) */
    float dropout;                     /* crnn_lightning.py:52 / sed.py:92 */
    int   dropout_each_block;          /* sed.py:107 applies it after every block */
    float bn_eps, bn_momentum;         /* nn.BatchNorm2d defaults 1e-5 / 0.1 */
    int   tensor_cores;                /* 1: conv contractions on tcgen05 (3-term bf16 split, fp32-grade) wherever the
                                          shape allows (channels % 128 == 0, W | 128; the first block with in_ch <= 2
                                          and pool 2 or 5); 0: fp32 CUDA cores everywhere */
} sedb200_crnn_desc;

/* geometry helpers (host only, no GPU needed) */
int    sedb200_crnn_validate(const sedb200_crnn_desc* d);
int    sedb200_crnn_seq_len(const sedb200_crnn_desc* d);            /* T  */
int    sedb200_crnn_flat(const sedb200_crnn_desc* d);               /* GRU-1 input width */
int    sedb200_crnn_n_tensors(const sedb200_crnn_desc* d);
long   sedb200_crnn_param_layout(const sedb200_crnn_desc* d, long* offsets /* [n_tensors] or NULL */);
long   sedb200_crnn_bn_state_floats(const sedb200_crnn_desc* d);
size_t sedb200_crnn_workspace_bytes(const sedb200_crnn_desc* d, int batch);

/* Forward.  training != 0: batch-statistics BatchNorm (running stats in bn_state_dev are updated),
 * dropout with the counter-based generator keyed by `seed`, activations saved in `ws_dev` for
 * sedb200_crnn_backward.  training == 0: running statistics, no dropout. */
int sedb200_crnn_forward(const sedb200_crnn_desc* d, const float* params_dev, float* bn_state_dev,
                         const float* x_dev, int batch, int training, unsigned long long seed,
                         void* ws_dev, size_t ws_bytes, float* logits_dev, void* stream);

/* Loss + its gradient in one pass over the logits.  loss_dev[0] = mean loss; probs_dev (optional) =
 * sigmoid(logits) (crnn_lightning.py:98); dlogits_dev (optional) = d(mean loss)/d(logits) * grad_scale. */
int sedb200_loss_fwd_bwd(int loss_kind, float alpha, float gamma, const float* logits_dev,
                         const float* targets_dev, long n, float grad_scale, float* loss_dev,
                         float* probs_dev, float* dlogits_dev, void* scratch_dev, size_t scratch_bytes,
                         void* stream);
size_t sedb200_loss_scratch_bytes(long n);

/* Fused per-frame head (crnn_lightning.py:63-64,72-73 + the loss): for models with exactly two dense layers
 * (sedb200_crnn_head_supported != 0), ONE kernel computes relu(d1) -> d2 -> sigmoid -> loss from the GRU output saved
 * in ws_dev and the whole backward of that head.  Call order for a training step:
 *   sedb200_crnn_forward(..., logits_dev = NULL)       conv blocks + GRU stack only
 *   sedb200_crnn_head_fwd_bwd(...)                      zero-fills grads_dev, writes the dense gradients, loss_dev[0],
 *                                                       probs_dev / logits_dev (optional), d(GRU output) into ws_dev
 *   sedb200_crnn_backward(..., dlogits_dev = NULL)      GRU stack + conv blocks
 * Same arithmetic as forward + sedb200_loss_fwd_bwd + backward (summation order of the loss / weight-gradient
 * partials differs). */
int sedb200_crnn_head_supported(const sedb200_crnn_desc* d);
int sedb200_crnn_head_fwd_bwd(const sedb200_crnn_desc* d, const float* params_dev, int batch, void* ws_dev,
                              size_t ws_bytes, const float* targets_dev, int loss_kind, float alpha, float gamma,
                              float grad_scale, float* logits_dev, float* probs_dev, float* loss_dev,
                              float* grads_dev, void* stream);

/* Backward of the last training forward held in ws_dev: grads_dev (same layout as params) is
 * OVERWRITTEN with d(loss)/d(params); dx_dev (optional, [B][in_ch][H][W]) receives d(loss)/d(x).
 * dlogits_dev == NULL: continue after sedb200_crnn_head_fwd_bwd (see above).
 * Stream semantics: everything is ordered after the work already enqueued on `stream`, and everything enqueued on
 * `stream` after the call returns is ordered after the whole backward pass.  Internally the GRU weight-gradient GEMMs
 * run on a helper stream the library creates once per device (forked from and rejoined to `stream` with events; no
 * host synchronisation).  The first conv block keeps no conv output ("lean block 0"): with dx_dev != NULL
 * it is rebuilt from x_dev, which must therefore still hold the batch given to sedb200_crnn_forward. */
int sedb200_crnn_backward(const sedb200_crnn_desc* d, const float* params_dev, const float* x_dev,
                          int batch, unsigned long long seed, void* ws_dev, size_t ws_bytes,
                          const float* dlogits_dev, float* grads_dev, float* dx_dev, void* stream);

/* The keep-mask of conv block `block`'s dropout for (seed, batch) exactly as the fused kernels evaluate it (counter-
 * based generator, nothing is stored by the forward pass), laid out as the reference's nn.Dropout would see the
 * tensor: bytes [batch][conv_ch][H][W_out], 1 = kept (crnn_lightning.py:52, sed.py:92,107).  Blocks that carry no
 * dropout in this configuration give all ones.  Test hook: lets an oracle apply the SAME mask. */
int sedb200_crnn_dropout_mask(const sedb200_crnn_desc* d, int batch, unsigned long long seed, int block,
                              unsigned char* mask_dev, void* stream);

/* The recurrent half of nn.GRU(bidirectional=True, batch_first=True) on its own (crnn_lightning.py:61-62,71;
 * sed.py:101,111), exposed for unit tests.  Gate order r, z, n; h0 = 0.
 *   gi    [B][T][2][3H]  x W_ih^T + b_ih per direction (0 = forward in time, 1 = reverse)
 *   whh   [2][3H][H], bhh [2][3H]
 *   out   [B][T][2H]     forward half | reverse half
 *   gates [B][T][2][4H]  r, z, n, q = W_hn h_{t-1} + b_hn (saved for the backward scan)
 *   dout  [B][T][2H] -> dgi [B][T][2][3H] (gradient w.r.t. gi), dgh [B][T][2][3H] (w.r.t. W_hh h_{t-1} + b_hh);
 *   part_b [B][2][2][3H]: per-batch-row sums over t of dgi ([.][0]) and dgh ([.][1]) when
 *   sedb200_gru_scan_fused_bias_grads(H) != 0 (then it must be given), untouched otherwise. */
int sedb200_gru_scan_fused_bias_grads(int H);
int sedb200_gru_scan_fwd(const float* gi_dev, const float* whh_dev, const float* bhh_dev, float* out_dev,
                         float* gates_dev, int B, int T, int H, void* stream);
int sedb200_gru_scan_bwd(const float* dout_dev, const float* out_dev, const float* gates_dev, const float* whh_dev,
                         float* dgi_dev, float* dgh_dev, float* part_b_dev, int B, int T, int H, void* stream);

/* Global-norm clip (clip_grad_norm_ semantics: coef = min(1, max_norm/(norm+1e-6)); max_norm <= 0
 * disables) followed by torch.optim.Adam with coupled L2 weight decay, over flat buffers.
 * grad_prescale multiplies every gradient first (1/world_size after a sum all-reduce).
 * gnorm_dev[0] receives the pre-clip global norm. */
size_t sedb200_clip_adam_scratch_bytes(long n);
int sedb200_clip_adam(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                      long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                      long step /* 1-based */, float max_norm, float grad_prescale, float* gnorm_dev,
                      void* scratch_dev, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole training steps in a CUDA graph.  Two things change from step to step besides the data: the dropout seed and
 * Adam's bias corrections (crnn_lightning.py:195-197: torch.optim.Adam's `step`).  Passed by value they would be frozen
 * into a captured graph, so the *_s variants read them from a 32-byte device-side step state instead:
 *   { u64 seed; i64 step; f32 bias_correction1; f32 sqrt(bias_correction2); 8 pad bytes }
 * sedb200_step_state_init sets the number of completed steps; sedb200_step_advance (the first node of a captured step)
 * does  step += 1; seed = base_seed + step - 1; corrections for `step`  -- exactly the values the by-value entry points
 * are given by the host loop, so a replayed graph and an eager step are bit-identical.
 * Everything the step enqueues (the library's helper stream included: it is forked from and joined to `stream` with
 * events) is capturable; no entry point synchronises or allocates. */
size_t sedb200_step_state_bytes(void);
int sedb200_step_state_init(void* step_state_dev, long completed_steps, void* stream);
int sedb200_step_advance(void* step_state_dev, unsigned long long base_seed, float beta1, float beta2, void* stream);
int sedb200_crnn_forward_s(const sedb200_crnn_desc* d, const float* params_dev, float* bn_state_dev,
                           const float* x_dev, int batch, int training, const void* step_state_dev,
                           void* ws_dev, size_t ws_bytes, float* logits_dev, void* stream);
int sedb200_crnn_backward_s(const sedb200_crnn_desc* d, const float* params_dev, const float* x_dev,
                            int batch, const void* step_state_dev, void* ws_dev, size_t ws_bytes,
                            const float* dlogits_dev, float* grads_dev, float* dx_dev, void* stream);
int sedb200_clip_adam_s(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                        long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                        const void* step_state_dev, float max_norm, float grad_prescale, float* gnorm_dev,
                        void* scratch_dev, size_t scratch_bytes, void* stream);
/* the fused NVLink exchange step (below) for captured steps: exchange number == step of the state, `parity` = that
 * step & 1 is fixed per captured graph (capture one graph per parity and alternate them) */
int sedb200_p2p_allreduce_clip_adam_s(void* const* regions_host, int world, int rank, long n, int parity,
                                      const void* step_state_dev, float* params_dev, float* exp_avg_dev,
                                      float* exp_avg_sq_dev, float* reduced_dev, float lr, float beta1, float beta2,
                                      float eps, float weight_decay, float max_norm, float grad_prescale,
                                      float* gnorm_dev, void* scratch_dev, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel exchange step as ONE kernel over NVLink peer memory (single node, one process per GPU):
 * sum all-reduce of the flat gradient buffers of all ranks + global-norm clip + Adam, replacing
 * `dist.all_reduce(grads)` followed by sedb200_clip_adam (train_lightning.py:50, crnn_lightning.py:195-197).
 *
 * Each rank allocates one exchange region (sedb200_p2p_region_alloc: cudaMalloc + CUDA IPC handle -- the one place
 * the library allocates, because IPC needs a whole cudaMalloc allocation), sends the 64-byte handle to its peers
 * (any transport: torch.distributed.all_gather_object) and maps theirs (sedb200_p2p_region_open).  Region layout:
 * flag row + status word in the first 1024 B, then two gradient buffers selected by `seq & 1`; a rank's backward
 * pass for exchange number `seq` must write its gradients at region + sedb200_p2p_grad_offset_bytes(n, seq & 1).
 * `seq` starts at 1 and increases by exactly 1 per call on every rank; `step` is Adam's bias-correction step.
 * reduced_dev [n] receives the summed gradients; results are bit-identical on all ranks (fixed rank order).
 * Failure handling: if a peer does not publish within the wait bound (60 s; environment variable
 * SEDB200_P2P_TIMEOUT_MS overrides) the kernel raises the region's status word, which is STICKY, and this and every
 * later exchange aborts grid-uniformly: params / exp_avg / exp_avg_sq are left untouched and gnorm_dev[0] = NaN (a
 * stale peer buffer is never applied).  sedb200_p2p_status reads the word back (synchronising copy); the word sits at
 * byte offset sedb200_p2p_status_offset_bytes() of the region for callers that prefer an asynchronous copy. */
long   sedb200_p2p_status_offset_bytes(void);
size_t sedb200_p2p_region_bytes(long n);
long   sedb200_p2p_grad_offset_bytes(long n, int parity);
int    sedb200_p2p_region_alloc(size_t bytes, void** region_dev, unsigned char* ipc_handle /* [64] */);
int    sedb200_p2p_region_open(const unsigned char* ipc_handle /* [64] */, void** region_dev);
int    sedb200_p2p_region_close(void* region_dev);
int    sedb200_p2p_region_free(void* region_dev);
int    sedb200_p2p_status(const void* region_dev, unsigned int* status_host);
size_t sedb200_p2p_scratch_bytes(void);
int    sedb200_p2p_allreduce_clip_adam(void* const* regions_host /* [world], rank order */, int world, int rank, long n,
                                       long seq, long step, float* params_dev, float* exp_avg_dev,
                                       float* exp_avg_sq_dev, float* reduced_dev, float lr, float beta1, float beta2,
                                       float eps, float weight_decay, float max_norm, float grad_prescale,
                                       float* gnorm_dev, void* scratch_dev, size_t scratch_bytes, void* stream);

/* sigmoid -> strict > 0.5 -> integer counts behind metrics.py:20-68 (crnn_lightning.py:112-126).
 * probs/targets: float32 [n_rows][n_cls] (the [N,T,C] arrays flattened by utils.reshape_3Dto2D).
 * counts_dev: 12 uint64 = {TP,Nsys,Nref,S,D,I} framewise, then the same six after block-max with
 * `block` rows per block; ceil(n_rows/block) blocks feed TP/Nsys/Nref (metrics.py:50),
 * floor(n_rows/block) blocks feed S/D/I and a second Nref stored in slot 12 (metrics.py:62).
 * counts_dev therefore has 13 entries. */
int sedb200_threshold_counts(const float* probs_dev, const float* targets_dev, long n_rows, int n_cls,
                             int block, float threshold, unsigned long long* counts_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Window sampler: HitWindowDataset.__getitem__ for a whole batch (decorte_datamodule.py:88-111, sed.py:72-76).
 * The fold's feature matrix mel_dev [n_frames][n_ch*n_feat] and label matrix lab_dev [n_frames][n_lab] stay in
 * HBM; per step only the window starts (and SpecAugment offsets) are uploaded.
 *   layout 0 (fork, crnn_lightning.py:66):  x[b][c][f][t] = mel[starts[b]+t][c*n_feat+f]   -> [batch][n_ch][n_feat][seq_in]
 *   layout 1 (SEDnet, time-major):          x[b][c][t][f] = same element                    -> [batch][n_ch][seq_in][n_feat]
 *   y[b][j][k] = max_{p < seq_in/seq_out} lab[starts[b] + j*seq_in/seq_out + p][k]          (decorte_datamodule.py:101)
 * SpecAugment (decorte_datamodule.py:39-49): for each of n_masks masks, frames [tmask[b][i], +time_mask_w) and mel
 * bins [fmask[b][i], +freq_mask_w) are zeroed; an offset < 0 disables that mask; tmask_dev / fmask_dev may be NULL.
 * y_dev may be NULL (features only). */
int sedb200_window_batch_f32(const float* mel_dev, const float* lab_dev, long n_frames, int n_ch, int n_feat,
                             int n_lab, const long* starts_dev, int batch, int seq_in, int seq_out,
                             const int* tmask_dev, const int* fmask_dev, int n_masks, int time_mask_w,
                             int freq_mask_w, int layout, float* x_dev, float* y_dev, void* stream);
/* _find_clean_negatives (decorte_datamodule.py:19-23, sed.py:48-52): flag_dev[s] = 1 iff none of the frames
 * [s, s+seq_in) has lab[.][0] == 1, for s in [0, n_frames-seq_in]. */
int sedb200_clean_negatives(const float* lab_dev, long n_frames, int n_lab, int seq_in, unsigned char* flag_dev,
                            void* stream);
/* Label rasterisation (feature.py:89-93): for every event i, lab[s:e][col] = 1.0 with
 * s = int(floor(start_s[i]*sr/hop)), e = int(ceil(end_s[i]*sr/hop)) and Python slice clamping.  lab_dev
 * [n_frames][n_lab] is NOT cleared by the call (feature.py:88 zero-fills it first). */
int sedb200_rasterize_labels(const double* start_s_dev, const double* end_s_dev, int n_events, int sr, int hop,
                             long n_frames, int n_lab, int col, float* lab_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tensor-core building block, exposed for unit tests: one 3x3 / pad 1 convolution (nn.Conv2d(.., 3,
 * padding=1), crnn_lightning.py:47) or its data gradient on channels-last fp32 tensors, computed on
 * tcgen05 with the 3-term bf16 split.  in [B][H][W][Cin] -> out [B][H][W][Cout] (dgrad: in = dY
 * [B][H][W][Cout] -> out = dX [B][H][W][Cin]); weight in PyTorch layout [Cout][Cin][3][3].
 * Supported: K channels % 64 == 0, N channels % 128 == 0, W a divisor of 128. */
size_t sedb200_conv3x3_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout);
int    sedb200_conv3x3_tc(const float* in_dev, const float* weight_dev, const float* bias_dev, float* out_dev,
                          int B, int H, int W, int Cin, int Cout, int dgrad, void* scratch_dev,
                          size_t scratch_bytes, void* stream);
/* weight gradient of the same convolution: dw [Cout][Cin][3][3] from dy [B][H][W][Cout] and in
 * [B][H][W][Cin].  Supported: Cin % 128 == 0, Cout % 128 == 0, W a divisor of 32. */
size_t sedb200_conv3x3_wgrad_tc_scratch_bytes(int B, int H, int W, int Cin, int Cout);
int    sedb200_conv3x3_wgrad_tc(const float* dy_dev, const float* in_dev, float* dw_dev, int B, int H, int W,
                                int Cin, int Cout, void* scratch_dev, size_t scratch_bytes, void* stream);

/* Test hook for the PLANE-NATIVE launches the CRNN flow uses (fp16 operand planes; forward = fp16 pass + e4m3
 * correction pass, gradients = one fp16 pass, halo boxes, tile pairs): the planes are produced from the fp32 tensors
 * given here with the helpers of the pool kernels, then exactly the launches of sedb200_crnn_forward / _backward run.
 *   mode 0: out [B][H][W][Cout] = conv(in [B][H][W][Cin], weight)                       (in2 unused)
 *   mode 1: out [B][H][W][Cin]  = conv_transposed(in = dY [B][H][W][Cout], weight)      (in2 unused)
 *   mode 2: out [Cout][Cin][3][3] = wgrad(in = dY [B][H][W][Cout], in2 = x [B][H][W][Cin])   (weight unused)
 * Reference call sites: nn.Conv2d(.., 3, padding=1) and its autograd backward, crnn_lightning.py:47 / sed.py:88. */
size_t sedb200_conv3x3_planes_test_scratch_bytes(int B, int H, int W, int Cin, int Cout);
int    sedb200_conv3x3_planes_test(const float* in_dev, const float* in2_dev, const float* weight_dev, float* out_dev,
                                   int B, int H, int W, int Cin, int Cout, int mode, void* scratch_dev,
                                   size_t scratch_bytes, void* stream);

/* Small-channel direct 3x3 convolutions (<= 64 channels: the reference's shipped CONV_DEPTH = 16,
 * train_constants.py), exposed for unit tests.  `in` is read through strides (element b*sB + h*sH + w*sW + k*sC), so
 * NCHW user input and channels-last activations both work; outputs are channels-last.
 *   forward (dgrad = 0): out [B][H][W][N] = conv(in [K channels], weight [N][K][3][3]) + bias
 *   data gradient (dgrad = 1): in = dY with K = Cout channels, weight = the forward weight [K][N][3][3], out = dX
 *   weight gradient: dw [N][K][3][3] from dy [B][H][W][N] and the conv input (K channels). */
int    sedb200_conv3x3_small_supported(int K, int N, int W, int wgrad);
int    sedb200_conv3x3_small(const float* in_dev, long sB, long sH, long sW, long sC, int K, int B, int H, int W,
                             const float* weight_dev, const float* bias_dev, int N, int dgrad, float* out_dev,
                             void* stream);
size_t sedb200_conv3x3_small_wgrad_scratch_bytes(int K, int N, int B, int H);
int    sedb200_conv3x3_small_wgrad(const float* dy_dev, const float* in_dev, long sB, long sH, long sW, long sC, int K,
                                   int N, int B, int H, int W, float* dw_dev, void* scratch_dev, size_t scratch_bytes,
                                   void* stream);

/* Plain GEMM on tcgen05 with the same 3-term bf16 split (unit-test entry for the kernel behind the GRU input
 * projections, crnn_lightning.py:61-62 / sed.py:101, and their gradients):
 *   out[M][N] (+bias[N]) = sum_k A(m,k) * B(n,k),   a_mn / b_mn: 0 = operand stored [rows][K], 1 = stored [K][rows].
 * split_k != 0 splits K over the SMs and reduces the partials in a fixed order.  M, N, K multiples of 8. */
size_t sedb200_gemm_tc_scratch_bytes(int M, int N, int K);
int    sedb200_gemm_tc(const float* a_dev, int a_mn, const float* b_dev, int b_mn, int M, int N, int K,
                       const float* bias_dev, float* out_dev, int split_k, void* scratch_dev, size_t scratch_bytes,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEDB200_H */
