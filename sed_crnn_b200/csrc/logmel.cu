// logmel.cu -- fused frame + Hann window + real FFT(2048) + |X|^2 + 40-band Slaney mel + log.
//
// Replaces feature._mbe (/root/reference/feature.py:55-59).  One WARP owns one frame end to end:
//
//   global --LDG.64--> 32 complex values / lane  (z[n] = x[2n] + i x[2n+1], n = lane + 32 j)
//     * Hann (shared table)                                              feature.py:56 (window)
//     -> 32-point DFT in registers over j            } complex FFT-1024 as 32 x 32
//     -> twiddle W_1024^(lane*a), transpose via smem }   (one shared-memory exchange)
//     -> 32-point DFT in registers over lane         }
//     -> real-FFT untangle of bins (k, 1024-k), |X|^2                    feature.py:57
//     -> sparse mel projection (<= 2 bands per bin), deterministic order feature.py:58-59
//     -> logf, 160 B coalesced store per (frame, channel)                feature.py:59
//
// The frame never touches HBM between the PCM load and the 40 output floats, so algorithmic HBM
// traffic is 4 B/sample in + 160 B/frame out; the 50 % overlap between neighbouring frames is
// served by L1/L2 because consecutive warps of a CTA take consecutive frames.
//
// Arithmetic is fp32 on the CUDA cores, on purpose: the 1e-4 log-mel gate needs ~2^-16 operand
// precision, a DFT-as-GEMM at that precision needs 3 bf16 MMAs per product (>= 1.6 MFLOP/frame,
// i.e. more than the whole tensor peak at the HBM roofline), while the factored fp32 FFT is
// ~45 kFLOP/frame ("log-mel kernel").
#include "common.cuh"

#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace sedb200 {
namespace {

constexpr int kNfft = SEDB200_NFFT;      // 2048
constexpr int kHop = SEDB200_HOP;        // 1024
constexpr int kBins = kNfft / 2 + 1;     // 1025
constexpr int kMel = SEDB200_NMEL;       // 40
constexpr int kM = kNfft / 2;            // complex FFT length 1024
constexpr int kWarps = 16;               // warps (= frames in flight) per CTA
constexpr int kBinStride = 33;           // bins walked per lane in the mel stage
constexpr int kMaxSlots = 192;
constexpr int kMaxTerms = 16;            // max lanes contributing to one mel band

// Constant tables, built on the host in double precision, one copy per (device, sr).
struct LogmelTables {
    float2 tw1[32 * 32];        // [a][t] = exp(-2 pi i t a / 1024)
    float  win[kNfft];          // periodic Hann
    float2 tw2[kM / 2 + 8];     // exp(-2 pi i k / 2048), k = 0..512
    float2 binw[kBinStride * 32];   // per bin: weights for band binband[f] and binband[f]+1; the SIGN BIT of .y says
                                    // "the band index steps up at this bin" (weights are >= 0); zero padded
    unsigned char gather[(kMel + 1) * kMaxTerms];   // per band: the partial-sum slots to add, in lane order
    unsigned char count[kMel + 1 + 7];              // number of partial sums of each band
    unsigned char lanebase[32];                     // first slot of each lane (its bands take consecutive slots)
};
static_assert(sizeof(LogmelTables) % 16 == 0, "tables are copied as uint4");

constexpr int kBufBytes = 32 * 33 * 8;                 // per-warp exchange buffer (8448 B)
constexpr int kPartBytes = kMaxSlots * 4;              // per-warp mel partial sums
constexpr int kSmemBytes = sizeof(LogmelTables) + kWarps * (kBufBytes + kPartBytes);

// ------------------------------------------------------------------------------ device: FFT-32
// complex add / subtract as ONE packed fp32x2 instruction each (sm_100 add.f32x2 / fma.f32x2): the (re, im)
// pair of a float2 is exactly the register pair the packed pipe wants
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
// (a + ib)(c - is)
__device__ __forceinline__ float2 cmul_conjtw(float2 d, float c, float s) {
    return make_float2(fmaf(d.y, s, d.x * c), fmaf(-d.x, s, d.y * c));
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.x, b.y, a.y * b.x));
}

// d * W_32^m, W_32 = exp(-2 pi i / 32), m a compile-time constant after unrolling.
__device__ __forceinline__ float2 mul_w32(float2 d, int m) {
    constexpr float C[16] = {1.0f,           0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654752f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f,
                             0.0f,           -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f,
                             -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
    constexpr float S[16] = {0.0f,           0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                             0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                             1.0f,           0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654752f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f};
    if (m == 0) return d;
    if (m == 8) return make_float2(d.y, -d.x);
    if (m == 4) return make_float2((d.x + d.y) * 0.70710678118654752f, (d.y - d.x) * 0.70710678118654752f);
    if (m == 12) return make_float2((d.y - d.x) * 0.70710678118654752f, -(d.x + d.y) * 0.70710678118654752f);
    return cmul_conjtw(d, C[m], S[m]);
}

__host__ __device__ constexpr int bitrev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// In-place radix-2 decimation-in-frequency DFT of 32 register values; on return v[i] holds the
// output with index bitrev5(i).
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;
#pragma unroll
        for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
            for (int k = 0; k < half; ++k) {
                const float2 u = v[g + k], w = v[g + k + half];
                v[g + k] = cadd(u, w);
                v[g + k + half] = mul_w32(csub(u, w), k << s);
            }
        }
    }
}

// sample `i` of a clip of S samples under librosa's centre padding
__device__ __forceinline__ float padded_sample(const float* __restrict__ x, long S, long i, int pad_mode) {
    if (i >= 0 && i < S) return __ldg(x + i);
    if (pad_mode == SEDB200_PAD_CONSTANT) return 0.0f;
    if (S == 1) return __ldg(x);
    const long period = 2 * (S - 1);
    long m = i % period;
    if (m < 0) m += period;
    return __ldg(x + (m < S ? m : period - m));
}

// ------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kWarps * 32, 1)
logmel_kernel(const float* __restrict__ pcm, float* __restrict__ out, int n_ch, long S, int n_frames,
              long total_frames, int pad_mode, const LogmelTables* __restrict__ gtab) {
    extern __shared__ __align__(16) unsigned char smem[];
    LogmelTables& tab = *reinterpret_cast<LogmelTables*>(smem);
    {
        const uint4* src = reinterpret_cast<const uint4*>(gtab);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(LogmelTables) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* buf = reinterpret_cast<float2*>(smem + sizeof(LogmelTables) + warp * kBufBytes);
    float* part = reinterpret_cast<float*>(smem + sizeof(LogmelTables) + kWarps * kBufBytes + warp * kPartBytes);
    float* P = reinterpret_cast<float*>(buf);

    for (long q = (long)blockIdx.x * kWarps + warp; q < total_frames; q += (long)gridDim.x * kWarps) {
        const int frame = (int)(q % n_frames);
        const long cc = q / n_frames;                      // clip * n_ch + ch
        const float* __restrict__ xb = pcm + cc * S;
        const long start = (long)(frame - 1) * kHop;       // first sample of the frame (may be < 0)

        // ---- load + window: v[j] = z[lane + 32 j]
        float2 v[32];
        if (start >= 0 && start + kNfft <= S) {
            const float* xs = xb + start;
            if ((reinterpret_cast<uintptr_t>(xs) & 7) == 0) {
                const float2* x2 = reinterpret_cast<const float2*>(xs);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __ldg(x2 + lane + 32 * j);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j].x = __ldg(xs + 2 * (lane + 32 * j));
                    v[j].y = __ldg(xs + 2 * (lane + 32 * j) + 1);
                }
            }
        } else {
            // boundary frame (first / last of a clip): gather through the padding rule into the
            // exchange buffer with a rolled loop, then pick the values up like the fast path
            float* stage = reinterpret_cast<float*>(buf);
#pragma unroll 1
            for (int i = lane; i < kNfft; i += 32) stage[i] = padded_sample(xb, S, start + i, pad_mode);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = buf[lane + 32 * j];
            __syncwarp();
        }
        {
            const float2* w2 = reinterpret_cast<const float2*>(tab.win);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float2 w = w2[lane + 32 * j];
                v[j].x *= w.x;
                v[j].y *= w.y;
            }
        }

        // ---- stage 1: DFT over j, twiddle, transpose through shared memory
        fft32(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int a = bitrev5(i);
            const float2 t = (a == 0) ? v[i] : cmul(v[i], tab.tw1[a * 32 + lane]);
            buf[lane * 33 + a] = t;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = buf[t * 33 + lane];
        __syncwarp();

        // ---- stage 2: DFT over t; lane a now owns Z[a + 32 b]
        fft32(v);
#pragma unroll
        for (int i = 0; i < 32; ++i) buf[lane + 32 * bitrev5(i)] = v[i];
        __syncwarp();

        // ---- real-FFT untangle + power spectrum; bins k and 1024-k are produced together
        {
            float2 zk[16], zp[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int k = lane + 32 * m;
                zk[m] = buf[k];
                zp[m] = buf[(kM - k) & (kM - 1)];
            }
            const float2 zmid = buf[kM / 2];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int k = lane + 32 * m;
                const float2 w = tab.tw2[k];                                     // (cos, -sin)
                const float2 e = __ffma2_rn(zp[m], make_float2(1.0f, -1.0f), zk[m]);   // 2E = (zk.x+zp.x, zk.y-zp.y)
                const float2 o = make_float2(zk[m].y + zp[m].y, zp[m].x - zk[m].x);   // 2O
                const float2 t = cmul(o, w);                                     // 2 W^k O
                const float2 xa = cadd(e, t), xb2 = csub(e, t);
                P[k] = 0.25f * fmaf(xa.x, xa.x, xa.y * xa.y);
                P[kM - k] = 0.25f * fmaf(xb2.x, xb2.x, xb2.y * xb2.y);
            }
            if (lane == 0) P[kM / 2] = fmaf(zmid.x, zmid.x, zmid.y * zmid.y);
        }
        __syncwarp();

        // ---- mel projection: each lane walks 33 consecutive bins with two running sums (band cur, cur+1); when
        //      the band index steps up (static flag stored in the weight's sign bit) the finished sum goes to the
        //      lane's next slot.  No data-dependent loop, fully unrolled, deterministic summation order.
        {
            const int f0 = lane * kBinStride;
            int k = tab.lanebase[lane];
            float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
            for (int i = 0; i < kBinStride; ++i) {
                const int f = f0 + i;
                const float2 w = tab.binw[f];
                const float p = f < kBins ? P[f] : 0.0f;
                if (i > 0 && (__float_as_uint(w.y) >> 31)) {
                    part[k++] = a0;
                    a0 = a1;
                    a1 = 0.0f;
                }
                a0 = fmaf(w.x, p, a0);
                a1 = fmaf(fabsf(w.y), p, a1);
            }
            part[k] = a0;
            part[k + 1] = a1;
        }
        __syncwarp();
        {
            const long clip = cc / n_ch;
            const int ch = (int)(cc % n_ch);
            float* o = out + ((clip * n_frames + frame) * n_ch + ch) * kMel;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int b = lane + 32 * r;
                if (b < kMel) {
                    const int n = tab.count[b];
                    float acc = 0.0f;
                    for (int i = 0; i < n; ++i) acc += part[tab.gather[b * kMaxTerms + i]];
                    o[b] = logf(acc);
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------ host: tables
double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3;
    const double min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3;
    const double min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// librosa.filters.mel(sr, n_fft=2048, n_mels=40): Slaney scale, slaney norm, float32 [40][1025]
void build_mel(int sr, std::vector<float>& fb) {
    fb.assign((size_t)kMel * kBins, 0.0f);
    std::vector<double> mel_f(kMel + 2), fftf(kBins);
    const double lo = hz_to_mel(0.0), hi = hz_to_mel(sr / 2.0);
    const double step = (hi - lo) / (kMel + 1);
    for (int i = 0; i < kMel + 2; ++i) mel_f[i] = mel_to_hz(i == kMel + 1 ? hi : lo + step * i);
    const double val = 1.0 / (kNfft * (1.0 / sr));
    for (int k = 0; k < kBins; ++k) fftf[k] = k * val;
    for (int i = 0; i < kMel; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < kBins; ++k) {
            const double lower = -(mel_f[i] - fftf[k]) / fd0;
            const double upper = (mel_f[i + 2] - fftf[k]) / fd1;
            const float w = (float)std::fmax(0.0, std::fmin(lower, upper));   // float32 store
            fb[(size_t)i * kBins + k] = (float)((double)w * enorm);            // in-place *= in numpy
        }
    }
}

int build_tables(int sr, LogmelTables& t) {
    std::memset(&t, 0, sizeof(t));
    const double two_pi = 6.283185307179586476925286766559;
    for (int a = 0; a < 32; ++a)
        for (int l = 0; l < 32; ++l) {
            const double ang = -two_pi * (double)(l * a) / kM;
            t.tw1[a * 32 + l] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
    for (int n = 0; n < kNfft; ++n) t.win[n] = (float)(0.5 - 0.5 * std::cos(two_pi * n / kNfft));
    for (int k = 0; k <= kM / 2; ++k) {
        const double ang = -two_pi * (double)k / kNfft;
        t.tw2[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    std::vector<float> fb;
    build_mel(sr, fb);
    int prev = 0;
    unsigned char band[kBinStride * 32];
    for (int f = 0; f < kBinStride * 32; ++f) {
        if (f >= kBins) { band[f] = (unsigned char)prev; t.binw[f] = make_float2(0.0f, 0.0f); continue; }
        int first = -1, last = -1;
        for (int b = 0; b < kMel; ++b)
            if (fb[(size_t)b * kBins + f] != 0.0f) {
                if (first < 0) first = b;
                last = b;
            }
        if (first < 0) first = last = prev;
        if (last > first + 1 || first < prev || first > prev + 1)
            return fail(SEDB200_ESHAPE, "mel filterbank for sr=%d: bin %d does not fit the 2-bands-per-bin walk", sr, f);
        band[f] = (unsigned char)first;
        float w1 = first + 1 < kMel ? fb[(size_t)(first + 1) * kBins + f] : 0.0f;
        if (first > prev) w1 = -w1;                       // sign bit = "band index steps up here" (-0.0f works too)
        if (first > prev && w1 == 0.0f) w1 = -0.0f;
        t.binw[f] = make_float2(fb[(size_t)first * kBins + f], w1);
        prev = first;
    }
    // slots: lane l covers bands [band[f0], band[f_last] + 1], one slot each, consecutive
    int bmin[32], bmax[32], slots = 0;
    for (int l = 0; l < 32; ++l) {
        const int f0 = l * kBinStride, f1 = std::min(f0 + kBinStride, kBins);
        if (f0 >= kBins) return fail(SEDB200_ESHAPE, "bin walk layout broken");
        bmin[l] = band[f0];
        bmax[l] = band[f1 - 1] + 1;
        t.lanebase[l] = (unsigned char)slots;
        slots += bmax[l] - bmin[l] + 1;
    }
    if (slots > kMaxSlots || slots > 255) return fail(SEDB200_ESHAPE, "mel partial-sum slots %d too many", slots);
    for (int b = 0; b <= kMel; ++b) {
        int n = 0;
        for (int l = 0; l < 32; ++l)
            if (b >= bmin[l] && b <= bmax[l]) {
                if (n >= kMaxTerms) return fail(SEDB200_ESHAPE, "mel band %d has more than %d partial sums", b, kMaxTerms);
                t.gather[b * kMaxTerms + n++] = (unsigned char)(t.lanebase[l] + b - bmin[l]);
            }
        t.count[b] = (unsigned char)n;
    }
    return SEDB200_OK;
}

std::mutex g_tab_mu;
std::map<std::pair<int, int>, LogmelTables*> g_tabs;   // (device, sr) -> device copy
bool g_attr_set[64] = {false};

int get_tables(int sr, cudaStream_t stream, const LogmelTables** out) {
    int dev = 0;
    SED_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto it = g_tabs.find({dev, sr});
    if (it == g_tabs.end()) {
        static LogmelTables host;   // guarded by g_tab_mu
        int rc = build_tables(sr, host);
        if (rc) return rc;
        LogmelTables* d = nullptr;
        SED_CUDA_OK(cudaMalloc(&d, sizeof(LogmelTables)));
        SED_CUDA_OK(cudaMemcpyAsync(d, &host, sizeof(LogmelTables), cudaMemcpyHostToDevice, stream));
        SED_CUDA_OK(cudaStreamSynchronize(stream));   // one-off; later calls are fully asynchronous
        it = g_tabs.emplace(std::make_pair(dev, sr), d).first;
    }
    if (dev < 64 && !g_attr_set[dev]) {
        SED_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        g_attr_set[dev] = true;
    }
    *out = it->second;
    return SEDB200_OK;
}

}  // namespace
}  // namespace sedb200

using namespace sedb200;

extern "C" {

long sedb200_logmel_frames(long n_samples) { return n_samples <= 0 ? 0 : 1 + n_samples / kHop; }

int sedb200_logmel_f32(const float* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                       float* out_dev, void* stream) {
    SED_REQUIRE(n_clips >= 0 && n_ch >= 1, SEDB200_EINVAL, "logmel: n_clips=%d n_ch=%d", n_clips, n_ch);
    SED_REQUIRE(n_samples >= 1, SEDB200_EINVAL, "logmel: empty signal (n_samples=%ld)", n_samples);
    SED_REQUIRE(sr > 0, SEDB200_EINVAL, "logmel: sr=%d", sr);
    SED_REQUIRE(pad_mode == SEDB200_PAD_CONSTANT || pad_mode == SEDB200_PAD_REFLECT, SEDB200_EINVAL,
                "logmel: pad_mode=%d", pad_mode);
    if (n_clips == 0) return SEDB200_OK;
    SED_REQUIRE(pcm_dev && out_dev, SEDB200_EINVAL, "logmel: null buffer");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const LogmelTables* tab = nullptr;
    rc = get_tables(sr, st, &tab);
    if (rc) return rc;
    const long nfr = sedb200_logmel_frames(n_samples);
    SED_REQUIRE(nfr < (1L << 31), SEDB200_ESHAPE, "logmel: %ld frames per clip", nfr);
    const long total = (long)n_clips * n_ch * nfr;
    const long want = (total + kWarps - 1) / kWarps;
    const int grid = (int)std::min<long>(want, sm_count());
    logmel_kernel<<<grid, kWarps * 32, kSmemBytes, st>>>(pcm_dev, out_dev, n_ch, n_samples, (int)nfr, total,
                                                         pad_mode, tab);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

size_t sedb200_logmel_host_scratch(int n_clips, int n_ch, long n_samples) {
    if (n_clips <= 0 || n_ch <= 0 || n_samples <= 0) return 0;
    const size_t in = (size_t)n_clips * n_ch * n_samples * 4;
    const size_t outb = (size_t)n_clips * sedb200_logmel_frames(n_samples) * n_ch * kMel * 4;
    return ((in + 255) & ~(size_t)255) + outb;
}

int sedb200_logmel_host_f32(const float* pcm_host, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                            float* out_host, void* scratch_dev, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(n_clips >= 0 && n_ch >= 1 && n_samples >= 1, SEDB200_EINVAL, "logmel_host: bad shape");
    if (n_clips == 0) return SEDB200_OK;
    SED_REQUIRE(pcm_host && out_host && scratch_dev, SEDB200_EINVAL, "logmel_host: null buffer");
    const size_t need = sedb200_logmel_host_scratch(n_clips, n_ch, n_samples);
    SED_REQUIRE(scratch_bytes >= need, SEDB200_EWORKSPACE, "logmel_host: scratch %zu < %zu bytes", scratch_bytes, need);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const size_t in = (size_t)n_clips * n_ch * n_samples * 4;
    const size_t outb = (size_t)n_clips * sedb200_logmel_frames(n_samples) * n_ch * kMel * 4;
    float* d_in = reinterpret_cast<float*>(scratch_dev);
    float* d_out = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch_dev) + ((in + 255) & ~(size_t)255));
    SED_CUDA_OK(cudaMemcpyAsync(d_in, pcm_host, in, cudaMemcpyHostToDevice, st));
    rc = sedb200_logmel_f32(d_in, n_clips, n_ch, n_samples, sr, pad_mode, d_out, stream);
    if (rc) return rc;
    SED_CUDA_OK(cudaMemcpyAsync(out_host, d_out, outb, cudaMemcpyDeviceToHost, st));
    SED_CUDA_OK(cudaStreamSynchronize(st));
    return SEDB200_OK;
}

int sedb200_mel_filterbank(int sr, float* out_host) {
    SED_REQUIRE(sr > 0 && out_host, SEDB200_EINVAL, "mel_filterbank: bad argument");
    std::vector<float> fb;
    build_mel(sr, fb);
    std::memcpy(out_host, fb.data(), fb.size() * sizeof(float));
    return SEDB200_OK;
}

}  // extern "C"
