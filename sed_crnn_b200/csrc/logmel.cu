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
//        (each lane keeps its own half of every pair in registers; only the partner half crosses shared memory)
//     -> mel projection: the triangular filters are linear in the bin index between band edges, so each lane
//        only accumulates (sum P, sum i*P) per band-edge segment of its 33 bins -- no per-bin weight loads --
//        and 40 bands are closed with <= 8 (A, B) coefficient pairs each, fixed order   feature.py:58-59
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
#include "logmel.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace sedb200 {
namespace {

constexpr int kWarps = 16;               // warps (= frames in flight) per CTA

constexpr int kBufBytes = 32 * 33 * 8;                 // per-warp exchange buffer (8448 B)
constexpr int kExchBytes = 512 * 8;                    // untangle exchange: Z[512..1023]; P starts behind it
constexpr int kPartBytes = kMaxSlots * 8;              // per-warp mel partial sums (S0, S1)
constexpr int kSmemBytes = sizeof(LogmelTables) + kWarps * (kBufBytes + kPartBytes);
static_assert(kExchBytes + kBinStride * 32 * 4 <= kBufBytes, "P (33 x 32 floats) must fit behind the exchange half");

// ------------------------------------------------------------------------------ device: FFT-32
// complex add / subtract as ONE packed fp32x2 instruction each (sm_100 add.f32x2 / fma.f32x2): the (re, im)
// pair of a float2 is exactly the register pair the packed pipe wants
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
// Complex products as TWO packed instructions: d * (c, c), then (swapped, sign-flipped d) * (s, s) added on top --
// the swap / negate of the halves is an operand modifier of the packed pipe (F32x2.LO_HI.NP in SASS), not a MOV.
// (a + ib)(c - is)
__device__ __forceinline__ float2 cmul_conjtw(float2 d, float c, float s) {
    return __ffma2_rn(make_float2(d.y, -d.x), make_float2(s, s), __fmul2_rn(d, make_float2(c, c)));
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
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
    constexpr float h = 0.70710678118654752f;
    if (m == 4) return __fmul2_rn(__fadd2_rn(d, make_float2(d.y, -d.x)), make_float2(h, h));
    if (m == 12) return __fmul2_rn(__fadd2_rn(make_float2(-d.x, -d.y), make_float2(d.y, -d.x)), make_float2(h, h));
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

// ------------------------------------------------------------------------------ the kernel
template <typename T>
__global__ void __launch_bounds__(kWarps * 32, 1)
logmel_kernel(const T* __restrict__ pcm, float* __restrict__ out, int n_ch, long S, int n_frames,
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
    float2* part = reinterpret_cast<float2*>(smem + sizeof(LogmelTables) + kWarps * kBufBytes + warp * kPartBytes);
    float* P = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(buf) + kExchBytes);   // power spectrum, 33 x 32

    // (clip*n_ch + ch, frame) of this warp's first frame, then advanced by the grid stride without any division
    const long q0 = (long)blockIdx.x * kWarps + warp;
    const long stride = (long)gridDim.x * kWarps;
    long cc = q0 / n_frames;                               // clip * n_ch + ch
    int frame = (int)(q0 - cc * n_frames);
    const long stride_cc = stride / n_frames;
    const int stride_fr = (int)(stride - stride_cc * n_frames);
    for (long q = q0; q < total_frames; q += stride, cc += stride_cc, frame += stride_fr) {
        if (frame >= n_frames) {
            frame -= n_frames;
            ++cc;
        }
        const T* __restrict__ xb = pcm + cc * S;
        const long start = (long)(frame - 1) * kHop;       // first sample of the frame (may be < 0)

        // ---- load + window: v[j] = z[lane + 32 j]
        float2 v[32];
        if (start >= 0 && start + kNfft <= S) {
            const T* xs = xb + start;
            if ((reinterpret_cast<uintptr_t>(xs) & (2 * sizeof(T) - 1)) == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ld_pair(xs + 2 * (lane + 32 * j));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    v[j].x = ld_sample(xs + 2 * (lane + 32 * j));
                    v[j].y = ld_sample(xs + 2 * (lane + 32 * j) + 1);
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
            // periodic Hann: w[n + 1024] = 1 - w[n], so v[j + 16] (sample n + 1024) takes x - x * w[n]
            const float2* w2 = reinterpret_cast<const float2*>(tab.win);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 w = w2[lane + 32 * j];
                v[j] = __fmul2_rn(v[j], w);
                v[j + 16] = __ffma2_rn(make_float2(-v[j + 16].x, -v[j + 16].y), w, v[j + 16]);
            }
        }

        // ---- stage 1: DFT over j, twiddle, transpose through shared memory
        fft32(v);
        {
            // twiddles W_1024^(lane*a): every fourth one comes from the table, the three after it by multiplying with
            // W_1024^lane (two packed instructions instead of a shared-memory load -- this kernel is bound by the
            // shared-memory pipe, not by issue)
            const float2 w1 = tab.tw1[32 + lane];
            float2 tw = w1;
#pragma unroll
            for (int a = 0; a < 32; ++a) {
                const int i = bitrev5(a);
                if (a > 1) tw = (a % 4 == 0) ? tab.tw1[a * 32 + lane] : cmul(tw, w1);
                buf[lane * 33 + a] = (a == 0) ? v[i] : cmul(v[i], tw);
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = buf[t * 33 + lane];
        __syncwarp();

        // ---- stage 2: DFT over t; lane a now owns Z[a + 32 b] (register v[i] holds b = bitrev5(i)).
        //      Bins k and 1024-k are untangled together by the lane that owns the one with b < 16; its partner
        //      (lane (32 - a) % 32, b' >= 16) passes the other half through shared memory: ex[k - 512] = Z[k].
        fft32(v);
        float2* ex = buf;
#pragma unroll
        for (int i = 1; i < 32; i += 2) ex[lane + 32 * (bitrev5(i) - 16)] = v[i];
        __syncwarp();

        // ---- real-FFT untangle + power spectrum
        {
            constexpr float kC64[16] = {1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f, 0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f, 0.38268343236508984f, 0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f};
            constexpr float kS64[16] = {0.0f, 0.098017140329560604f, 0.19509032201612825f, 0.29028467725446233f, 0.38268343236508978f, 0.47139673682599764f, 0.55557023301960218f, 0.63439328416364549f, 0.70710678118654746f, 0.77301045336273699f, 0.83146961230254524f, 0.88192126434835494f, 0.92387953251128674f, 0.95694033573220894f, 0.98078528040323043f, 0.99518472667219682f};
            const float2 wl = tab.tw2[lane];                                     // W_2048^lane = (cos, -sin)
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int k = lane + 32 * m;
                const float2 zk = v[bitrev5(m)];
                float2 zp = ex[(m == 0 && lane == 0) ? 0 : 512 - k];
                if (m == 0 && lane == 0) zp = zk;                                // k = 0 pairs with itself
                const float2 e = __ffma2_rn(zp, make_float2(1.0f, -1.0f), zk);   // 2E = (zk.x+zp.x, zk.y-zp.y)
                const float2 o = make_float2(zk.y + zp.y, zp.x - zk.x);          // 2O
                // W_2048^k = W_2048^lane * W_64^m: per-lane factor from the table, per-m factor an immediate
                const float2 t = cmul(m == 0 ? o : cmul_conjtw(o, kC64[m], kS64[m]), wl);   // 2 W^k O
                const float2 xa = cadd(e, t), xb2 = csub(e, t);
                P[k] = fmaf(xa.x, xa.x, xa.y * xa.y);                           // 4 |X[k]|^2: the 1/4 lives in the
                P[kM - k] = fmaf(xb2.x, xb2.x, xb2.y * xb2.y);                  // mel coefficients
            }
            const float2 zmid = ex[0];                                           // Z[512]
            P[lane == 0 ? kM / 2 : kM + lane] = lane == 0 ? 4.0f * fmaf(zmid.x, zmid.x, zmid.y * zmid.y) : 0.0f;
        }
        __syncwarp();

        // ---- mel projection, part 1: each lane walks 33 consecutive bins.  Between two band edges every
        //      triangular weight is linear in the bin index, so the lane only keeps S0 = sum P and S1 = sum i*P of
        //      the current edge-to-edge segment; when the segment steps up (static per-lane bit mask) the pair goes
        //      to the lane's next slot.  No weight loads, no data-dependent control flow, fixed summation order.
        {
            const float* Pl = P + lane * kBinStride;
            const unsigned long long msk = tab.lanemask[lane];
            const unsigned mlo = (unsigned)msk, mhi = (unsigned)(msk >> 32);
            unsigned slot = (unsigned)__cvta_generic_to_shared(part + tab.lanebase[lane]);
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int i = 0; i < kBinStride; ++i) {
                const float p = Pl[i];
                if (i > 0) {
                    // one predicate per bin; store / advance / reset are predicated, nothing branches
                    asm volatile(
                        "{\n\t.reg .pred q;\n\t"
                        "setp.ne.u32 q, %3, 0;\n\t"
                        "@q st.shared.v2.f32 [%0], {%1, %2};\n\t"
                        "@q add.u32 %0, %0, 8;\n\t"
                        "@q mov.f32 %1, 0f00000000;\n\t"
                        "@q mov.f32 %2, 0f00000000;\n\t}"
                        : "+r"(slot), "+f"(s0), "+f"(s1)
                        : "r"((i < 32 ? mlo : mhi) & (1u << (i & 31)))
                        : "memory");
                }
                s0 += p;
                s1 = fmaf((float)i, p, s1);
            }
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(slot), "f"(s0), "f"(s1) : "memory");
        }
        __syncwarp();
        // ---- part 2: close the bands.  band = sum over its partial sums of A * S0 + B * S1 (<= 8 terms at 44.1 kHz)
        {
            const unsigned clip = (unsigned)cc / (unsigned)n_ch;       // cc < 2^31 (checked by the launcher)
            const int ch = (int)((unsigned)cc - clip * (unsigned)n_ch);
            float* o = out + (((long)clip * n_frames + frame) * n_ch + ch) * kMel;
            {
                const int b = kMel - kBandsRound1 + lane;
                float acc = 0.0f;
                for (int i = 0; i < tab.terms_round1; ++i) {
                    const float2 c = tab.coef[b * kMaxTerms + i];
                    const float2 sv = part[tab.gslot[b * kMaxTerms + i]];
                    acc = fmaf(c.x, sv.x, acc);
                    acc = fmaf(c.y, sv.y, acc);
                }
                o[b] = logf(acc);
            }
            if (lane < kMel - kBandsRound1) {
                const int b = lane;
                float acc = 0.0f;
                for (int i = 0; i < tab.terms_round2; ++i) {
                    const float2 c = tab.coef[b * kMaxTerms + i];
                    const float2 sv = part[tab.gslot[b * kMaxTerms + i]];
                    acc = fmaf(c.x, sv.x, acc);
                    acc = fmaf(c.y, sv.y, acc);
                }
                o[b] = logf(acc);
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

// band-edge frequencies mel_f[0 .. kMel+1] of librosa.filters.mel (Slaney scale, fmin 0, fmax sr/2)
void mel_edges(int sr, std::vector<double>& mel_f) {
    mel_f.resize(kMel + 2);
    const double lo = hz_to_mel(0.0), hi = hz_to_mel(sr / 2.0);
    const double step = (hi - lo) / (kMel + 1);
    for (int i = 0; i < kMel + 2; ++i) mel_f[i] = mel_to_hz(i == kMel + 1 ? hi : lo + step * i);
}

// librosa.filters.mel(sr, n_fft=2048, n_mels=40): Slaney scale, slaney norm, float32 [40][1025]
void build_mel(int sr, std::vector<float>& fb) {
    fb.assign((size_t)kMel * kBins, 0.0f);
    std::vector<double> mel_f, fftf(kBins);
    mel_edges(sr, mel_f);
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
    for (int n = 0; n < kNfft / 2; ++n) t.win[n] = (float)(0.5 - 0.5 * std::cos(two_pi * n / kNfft));
    for (int k = 0; k <= kM / 2; ++k) {
        const double ang = -two_pi * (double)k / kNfft;
        t.tw2[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    // Mel projection tables.  Segment j = [mel_f[j], mel_f[j+1]) holds the bins whose weight into band j rises
    // linearly, enorm_j (f - mel_f[j]) / D_j, and whose weight into band j-1 falls linearly,
    // enorm_{j-1} (mel_f[j+1] - f) / D_j  (the two branches of librosa's min(lower, upper) ramp).
    std::vector<double> mel_f;
    mel_edges(sr, mel_f);
    const double df = (double)sr / kNfft;
    int seg[kBinStride * 32];
    for (int f = 0; f < kBinStride * 32; ++f) {
        const double fr = std::min(f, kBins - 1) * df;
        int j = 0;
        while (j < kMel && fr >= mel_f[j + 1]) ++j;
        seg[f] = j;                                        // 0 .. kMel
    }
    int slots = 0;
    int slot_lane[kMaxSlots], slot_seg[kMaxSlots];
    for (int l = 0; l < 32; ++l) {
        const int f0 = l * kBinStride;
        if (f0 >= kBins) return fail(SEDB200_ESHAPE, "bin walk layout broken");
        t.lanebase[l] = (unsigned char)slots;
        unsigned long long m = 0;
        for (int i = 0; i < kBinStride; ++i) {
            if (i > 0 && seg[f0 + i] != seg[f0 + i - 1]) {
                if (seg[f0 + i] != seg[f0 + i - 1] + 1)
                    return fail(SEDB200_ESHAPE, "mel filterbank for sr=%d: a band edge segment holds no bin", sr);
                m |= 1ull << i;
            }
            if (i == 0 || (m >> i) & 1) {
                if (slots >= kMaxSlots) return fail(SEDB200_ESHAPE, "mel partial-sum slots exceed %d", kMaxSlots);
                slot_lane[slots] = l;
                slot_seg[slots] = seg[f0 + i];
                ++slots;
            }
        }
        t.lanemask[l] = m;
    }
    int n_terms[kMel] = {0};
    auto add_term = [&](int band, int slot, double A, double B) -> int {
        if (n_terms[band] >= kMaxTerms)
            return fail(SEDB200_ESHAPE, "mel band %d has more than %d partial sums", band, kMaxTerms);
        t.gslot[band * kMaxTerms + n_terms[band]] = (unsigned char)slot;
        t.coef[band * kMaxTerms + n_terms[band]] = make_float2((float)(0.25 * A), (float)(0.25 * B));   // P holds 4|X|^2
        ++n_terms[band];
        return SEDB200_OK;
    };
    for (int s = 0; s < slots; ++s) {
        const int j = slot_seg[s];
        const double f0 = slot_lane[s] * kBinStride * df, D = mel_f[j + 1] - mel_f[j];
        if (j <= kMel - 1) {                               // rising edge of band j
            const double en = 2.0 / (mel_f[j + 2] - mel_f[j]);
            int rc = add_term(j, s, en * (f0 - mel_f[j]) / D, en * df / D);
            if (rc) return rc;
        }
        if (j >= 1) {                                      // falling edge of band j - 1
            const double en = 2.0 / (mel_f[j + 1] - mel_f[j - 1]);
            int rc = add_term(j - 1, s, en * (mel_f[j + 1] - f0) / D, -en * df / D);
            if (rc) return rc;
        }
    }
    t.terms_round1 = t.terms_round2 = 0;
    for (int b = 0; b < kMel; ++b) {
        int& r = b >= kMel - kBandsRound1 ? t.terms_round1 : t.terms_round2;
        r = std::max(r, n_terms[b]);
    }
    return SEDB200_OK;
}

std::mutex g_tab_mu;
std::map<std::pair<int, int>, LogmelTables*> g_tabs;   // (device, sr) -> device copy

}  // namespace

int logmel_get_tables(int sr, cudaStream_t stream, const LogmelTables** out) {
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
    *out = it->second;
    return SEDB200_OK;
}

namespace {

// which kernel sedb200_logmel_* runs when the caller does not say: SEDB200_LOGMEL_KERNEL = "fp32" | "tc" overrides
int default_kernel() {
    static const int k = [] {
        const char* e = std::getenv("SEDB200_LOGMEL_KERNEL");
        if (e && std::strcmp(e, "fp32") == 0) return SEDB200_LOGMEL_FP32;
        if (e && std::strcmp(e, "tc") == 0) return SEDB200_LOGMEL_TC;
        return SEDB200_LOGMEL_DEFAULT_KERNEL;
    }();
    return k;
}

template <typename T>
int logmel_launch(const T* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode, float* out_dev,
                  void* stream, int kernel = SEDB200_LOGMEL_AUTO) {
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
    SED_REQUIRE(kernel == SEDB200_LOGMEL_AUTO || kernel == SEDB200_LOGMEL_FP32 || kernel == SEDB200_LOGMEL_TC,
                SEDB200_EINVAL, "logmel: kernel=%d", kernel);
    if (kernel == SEDB200_LOGMEL_AUTO) kernel = default_kernel();
    if (kernel == SEDB200_LOGMEL_TC) return logmel_tc_launch<T>(pcm_dev, n_clips, n_ch, n_samples, sr, pad_mode, out_dev, st);
    const LogmelTables* tab = nullptr;
    rc = logmel_get_tables(sr, st, &tab);
    if (rc) return rc;
    const long nfr = n_samples <= 0 ? 0 : 1 + n_samples / kHop;
    SED_REQUIRE(nfr < (1L << 31), SEDB200_ESHAPE, "logmel: %ld frames per clip", nfr);
    SED_REQUIRE((long)n_clips * n_ch < (1L << 31), SEDB200_ESHAPE, "logmel: %ld channel-clips", (long)n_clips * n_ch);
    const long total = (long)n_clips * n_ch * nfr;
    const long want = (total + kWarps - 1) / kWarps;
    const int grid = (int)std::min<long>(want, sm_count());
    rc = ensure_dyn_smem((const void*)logmel_kernel<T>, kSmemBytes);
    if (rc) return rc;
    logmel_kernel<T><<<grid, kWarps * 32, kSmemBytes, st>>>(pcm_dev, out_dev, n_ch, n_samples, (int)nfr, total,
                                                            pad_mode, tab);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

inline size_t logmel_scratch_bytes(int n_clips, int n_ch, long n_samples, int elem) {
    if (n_clips <= 0 || n_ch <= 0 || n_samples <= 0) return 0;
    const size_t in = (size_t)n_clips * n_ch * n_samples * elem;
    const size_t outb = (size_t)n_clips * (1 + n_samples / kHop) * n_ch * kMel * 4;
    return ((in + 255) & ~(size_t)255) + outb;
}

template <typename T>
int logmel_host(const T* pcm_host, int n_clips, int n_ch, long n_samples, int sr, int pad_mode, float* out_host,
                void* scratch_dev, size_t scratch_bytes, void* stream) {
    SED_REQUIRE(n_clips >= 0 && n_ch >= 1 && n_samples >= 1, SEDB200_EINVAL, "logmel_host: bad shape");
    if (n_clips == 0) return SEDB200_OK;
    SED_REQUIRE(pcm_host && out_host && scratch_dev, SEDB200_EINVAL, "logmel_host: null buffer");
    const size_t need = logmel_scratch_bytes(n_clips, n_ch, n_samples, (int)sizeof(T));
    SED_REQUIRE(scratch_bytes >= need, SEDB200_EWORKSPACE, "logmel_host: scratch %zu < %zu bytes", scratch_bytes, need);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const size_t in = (size_t)n_clips * n_ch * n_samples * sizeof(T);
    const size_t outb = (size_t)n_clips * (1 + n_samples / kHop) * n_ch * kMel * 4;
    T* d_in = reinterpret_cast<T*>(scratch_dev);
    float* d_out = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch_dev) + ((in + 255) & ~(size_t)255));
    SED_CUDA_OK(cudaMemcpyAsync(d_in, pcm_host, in, cudaMemcpyHostToDevice, st));
    rc = logmel_launch<T>(d_in, n_clips, n_ch, n_samples, sr, pad_mode, d_out, stream);
    if (rc) return rc;
    SED_CUDA_OK(cudaMemcpyAsync(out_host, d_out, outb, cudaMemcpyDeviceToHost, st));
    SED_CUDA_OK(cudaStreamSynchronize(st));
    return SEDB200_OK;
}

}  // namespace
}  // namespace sedb200

using namespace sedb200;

extern "C" {

long sedb200_logmel_frames(long n_samples) { return n_samples <= 0 ? 0 : 1 + n_samples / kHop; }

int sedb200_logmel_f32(const float* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                       float* out_dev, void* stream) {
    return logmel_launch<float>(pcm_dev, n_clips, n_ch, n_samples, sr, pad_mode, out_dev, stream);
}

int sedb200_logmel_i16(const short* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                       float* out_dev, void* stream) {
    return logmel_launch<short>(pcm_dev, n_clips, n_ch, n_samples, sr, pad_mode, out_dev, stream);
}

int sedb200_logmel_f32_k(const float* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                         float* out_dev, void* stream, int kernel) {
    return logmel_launch<float>(pcm_dev, n_clips, n_ch, n_samples, sr, pad_mode, out_dev, stream, kernel);
}

int sedb200_logmel_i16_k(const short* pcm_dev, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                         float* out_dev, void* stream, int kernel) {
    return logmel_launch<short>(pcm_dev, n_clips, n_ch, n_samples, sr, pad_mode, out_dev, stream, kernel);
}

int sedb200_logmel_default_kernel(void) { return default_kernel(); }

size_t sedb200_logmel_host_scratch(int n_clips, int n_ch, long n_samples) {
    return logmel_scratch_bytes(n_clips, n_ch, n_samples, 4);
}
size_t sedb200_logmel_host_scratch_i16(int n_clips, int n_ch, long n_samples) {
    return logmel_scratch_bytes(n_clips, n_ch, n_samples, 2);
}

int sedb200_logmel_host_f32(const float* pcm_host, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                            float* out_host, void* scratch_dev, size_t scratch_bytes, void* stream) {
    return logmel_host<float>(pcm_host, n_clips, n_ch, n_samples, sr, pad_mode, out_host, scratch_dev, scratch_bytes, stream);
}

int sedb200_logmel_host_i16(const short* pcm_host, int n_clips, int n_ch, long n_samples, int sr, int pad_mode,
                            float* out_host, void* scratch_dev, size_t scratch_bytes, void* stream) {
    return logmel_host<short>(pcm_host, n_clips, n_ch, n_samples, sr, pad_mode, out_host, scratch_dev, scratch_bytes, stream);
}

int sedb200_mel_filterbank(int sr, float* out_host) {
    SED_REQUIRE(sr > 0 && out_host, SEDB200_EINVAL, "mel_filterbank: bad argument");
    std::vector<float> fb;
    build_mel(sr, fb);
    std::memcpy(out_host, fb.data(), fb.size() * sizeof(float));
    return SEDB200_OK;
}

}  // extern "C"
