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

/* Copy of the float32 [40][1025] mel filterbank the kernel uses (== librosa.filters.mel(sr=sr,
 * n_fft=2048, n_mels=40), feature.py:58) into a HOST buffer; for inspection / tests. */
int    sedb200_mel_filterbank(int sr, float* out_host /* [40*1025] */);

#ifdef __cplusplus
}
#endif
#endif /* SEDB200_H */
