// logmel.cuh -- what the two log-mel kernels share: geometry constants, the constant tables (mel projection walk,
// Hann half window, FFT twiddles), PCM sample loads and the centre-padding rule (feature.py:55-59, librosa.stft).
#pragma once
#include "common.cuh"

namespace sedb200 {

constexpr int kNfft = SEDB200_NFFT;      // 2048
constexpr int kHop = SEDB200_HOP;        // 1024
constexpr int kBins = kNfft / 2 + 1;     // 1025
constexpr int kMel = SEDB200_NMEL;       // 40
constexpr int kM = kNfft / 2;            // complex FFT length 1024
constexpr int kBinStride = 33;           // bins walked per lane in the mel stage
constexpr int kMaxSlots = 96;            // (lane, segment) partial sums
constexpr int kMaxTerms = 12;            // max partial sums feeding one mel band
constexpr int kBandsRound1 = 32;         // bands kMel-32 .. kMel-1 are closed by lanes 0..31, the rest in a second round

// Constant tables, built on the host in double precision, one copy per (device, sr).
struct LogmelTables {
    float2 tw1[32 * 32];        // [a][t] = exp(-2 pi i t a / 1024)
    float  win[kNfft / 2];      // first half of the periodic Hann window; w[n + 1024] = 1 - w[n]
    float2 tw2[kM / 2 + 8];     // exp(-2 pi i k / 2048), k = 0..512
    float2 coef[kMel * kMaxTerms];          // per band: (A, B) of each partial sum: band += A * S0 + B * S1
    unsigned long long lanemask[32];        // bit i: the band-edge segment steps up at the lane's i-th bin
    unsigned char gslot[kMel * kMaxTerms];  // per band: the partial-sum slot of each term (padding: slot 0, A = B = 0)
    unsigned char lanebase[32];             // first slot of each lane (its segments take consecutive slots)
    int terms_round1, terms_round2;         // loop trip counts of the two closing rounds
    int pad_[2];
};
static_assert(sizeof(LogmelTables) % 16 == 0, "tables are copied as uint4");


// device copy of the tables for (current device, sr), built once in double precision
int logmel_get_tables(int sr, cudaStream_t stream, const LogmelTables** out);

// PCM sample types: float32 as the reference decodes it (feature.py:45-50, ffmpeg -f f32le), or int16 as a WAV /
// `-f s16le` decoder delivers it; an int16 sample s stands for the float32 value s / 32768 (exact), so both
// ingest paths feed the same arithmetic
__device__ __forceinline__ float ld_sample(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_sample(const short* p) { return (float)__ldg(p) * (1.0f / 32768.0f); }
__device__ __forceinline__ float2 ld_pair(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld_pair(const short* p) {
    const short2 v = __ldg(reinterpret_cast<const short2*>(p));
    return make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f));
}

// sample `i` of a clip of S samples under librosa's centre padding
template <typename T>
__device__ __forceinline__ float padded_sample(const T* __restrict__ x, long S, long i, int pad_mode) {
    if (i >= 0 && i < S) return ld_sample(x + i);
    if (pad_mode == SEDB200_PAD_CONSTANT) return 0.0f;
    if (S == 1) return ld_sample(x);
    const long period = 2 * (S - 1);
    long m = i % period;
    if (m < 0) m += period;
    return ld_sample(x + (m < S ? m : period - m));
}


// logmel_tc.cu: the tcgen05 formulation (DFT as two batched GEMMs on fp16 hi / lo planes)
template <typename T>
int logmel_tc_launch(const T* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode, float* out_dev,
                     cudaStream_t st);

}  // namespace sedb200
