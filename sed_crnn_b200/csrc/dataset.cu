// dataset.cu -- the callers either side of the hot path that feed it per step:
//
//   window_batch_kernel   HitWindowDataset.__getitem__ for a whole batch (decorte_datamodule.py:88-111,
//                         sed.py:72-76): slice SEQ_LEN_IN frames of the fold's feature matrix, transpose to
//                         (mel, time), SpecAugment zero masks (decorte_datamodule.py:39-49), max-pool the labels to
//                         SEQ_LEN_OUT steps (decorte_datamodule.py:101).  The feature matrix stays resident in HBM;
//                         only the window starts / mask offsets (a few hundred bytes) cross PCIe per step.
//   clean_negatives_kernel  _find_clean_negatives (decorte_datamodule.py:19-23): starts whose window holds no hit.
//   rasterize_labels_kernel event list -> frame labels (feature.py:89-93), Python slice semantics.
//
// Pure data movement / integer work: bit-exact against the reference by construction, HBM-bound.
#include "common.cuh"

#include <algorithm>

namespace sedb200 {
namespace {

constexpr int kTile = 32;

// x layout 0 ("fork", crnn_lightning.py:66): x[b][c][f][t] = mel[start_b + t][c*F + f]   (transpose per window)
// x layout 1 ("sednet", time-major):          x[b][c][t][f] = mel[start_b + t][c*F + f]
__global__ void __launch_bounds__(kTile * 8)
window_batch_kernel(const float* __restrict__ mel, const float* __restrict__ lab, long n_frames, int n_ch, int F,
                    int K, const long* __restrict__ starts, int L, int seq_out, const int* __restrict__ tmask,
                    const int* __restrict__ fmask, int n_masks, int tw, int fw, int layout, float* __restrict__ x,
                    float* __restrict__ y) {
    __shared__ float tile[kTile][kTile + 1];
    const int b = blockIdx.z;
    const int tiles_f = (F + kTile - 1) / kTile;
    const int c = blockIdx.y / tiles_f, f0 = (blockIdx.y % tiles_f) * kTile, t0 = blockIdx.x * kTile;
    const long start = __ldg(starts + b);
    const int cols = n_ch * F;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

    auto masked = [&](int f, int t) -> bool {
        bool m = false;
        for (int i = 0; i < n_masks; ++i) {
            if (tmask) { const int a = __ldg(tmask + (long)b * n_masks + i); m |= (a >= 0 && t >= a && t < a + tw); }
            if (fmask) { const int a = __ldg(fmask + (long)b * n_masks + i); m |= (a >= 0 && f >= a && f < a + fw); }
        }
        return m;
    };

    if (layout == 0) {
#pragma unroll
        for (int r = ty; r < kTile; r += 8) {             // coalesced along the feature axis
            const int t = t0 + r, f = f0 + tx;
            float v = 0.0f;
            if (t < L && f < F && start + t >= 0 && start + t < n_frames) v = __ldg(mel + (start + t) * cols + c * F + f);
            tile[r][tx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < kTile; r += 8) {             // coalesced along time
            const int f = f0 + r, t = t0 + tx;
            if (f < F && t < L) {
                float v = tile[tx][r];
                if (n_masks > 0 && masked(f, t)) v = 0.0f;
                x[(((long)b * n_ch + c) * F + f) * L + t] = v;
            }
        }
    } else {
#pragma unroll
        for (int r = ty; r < kTile; r += 8) {
            const int t = t0 + r, f = f0 + tx;
            if (t < L && f < F) {
                float v = 0.0f;
                if (start + t >= 0 && start + t < n_frames) v = __ldg(mel + (start + t) * cols + c * F + f);
                if (n_masks > 0 && masked(f, t)) v = 0.0f;
                x[(((long)b * n_ch + c) * L + t) * F + f] = v;
            }
        }
    }

    // labels: one block per window does the (tiny) max-pool  decorte_datamodule.py:101
    if (y != nullptr && blockIdx.x == 0 && blockIdx.y == 0) {
        const int pool = L / seq_out;
        for (int i = threadIdx.x; i < seq_out * K; i += blockDim.x) {
            const int j = i / K, k = i % K;
            float m = -INFINITY;
            for (int p = 0; p < pool; ++p) {
                const long fr = start + (long)j * pool + p;
                if (fr >= 0 && fr < n_frames) m = fmaxf(m, __ldg(lab + fr * K + k));
            }
            y[((long)b * seq_out + j) * K + k] = m;
        }
    }
}

// flag[s] = 1 iff no frame of [s, s+L) has lab[.][0] == 1      (np.convolve(mask, ones(L), 'valid') == 0)
__global__ void clean_negatives_kernel(const float* __restrict__ lab, long n_frames, int K, int L,
                                       unsigned char* __restrict__ flag, long n_starts) {
    for (long s = (long)blockIdx.x * blockDim.x + threadIdx.x; s < n_starts; s += (long)gridDim.x * blockDim.x) {
        bool hit = false;
        for (int i = 0; i < L; ++i) hit |= (__ldg(lab + (s + i) * K) == 1.0f);
        flag[s] = hit ? 0 : 1;
    }
}

// Python `lbl[s:e, 0] = 1.0` with s = int(floor(start*sr/hop)), e = int(ceil(end*sr/hop))  (feature.py:89-93)
__device__ __forceinline__ long py_slice_index(long i, long n) {
    if (i < 0) { i += n; return i < 0 ? 0 : i; }
    return i > n ? n : i;
}

__global__ void rasterize_labels_kernel(const double* __restrict__ start_s, const double* __restrict__ end_s,
                                        int n_events, double sr, double hop, long n_frames, int K, int col,
                                        float* __restrict__ lab) {
    for (long f = (long)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += (long)gridDim.x * blockDim.x) {
        bool on = false;
        for (int i = 0; i < n_events; ++i) {
            // two separately rounded double operations, exactly like `hit["start"] * SR / HOP`
            const double a = __ddiv_rn(__dmul_rn(__ldg(start_s + i), sr), hop);
            const double z = __ddiv_rn(__dmul_rn(__ldg(end_s + i), sr), hop);
            const long s = py_slice_index((long)floor(a), n_frames), e = py_slice_index((long)ceil(z), n_frames);
            on |= (f >= s && f < e);
        }
        if (on) lab[f * K + col] = 1.0f;
    }
}

}  // namespace
}  // namespace sedb200

using namespace sedb200;

extern "C" {

int sedb200_window_batch_f32(const float* mel_dev, const float* lab_dev, long n_frames, int n_ch, int n_feat,
                             int n_lab, const long* starts_dev, int batch, int seq_in, int seq_out,
                             const int* tmask_dev, const int* fmask_dev, int n_masks, int time_mask_w,
                             int freq_mask_w, int layout, float* x_dev, float* y_dev, void* stream) {
    SED_REQUIRE(batch >= 0 && n_ch >= 1 && n_feat >= 1 && seq_in >= 1, SEDB200_EINVAL,
                "window_batch: batch=%d n_ch=%d n_feat=%d seq_in=%d", batch, n_ch, n_feat, seq_in);
    SED_REQUIRE(n_frames >= seq_in, SEDB200_ESHAPE, "window_batch: %ld frames < window of %d", n_frames, seq_in);
    SED_REQUIRE(layout == 0 || layout == 1, SEDB200_EINVAL, "window_batch: layout=%d", layout);
    SED_REQUIRE(n_masks >= 0 && n_masks <= 16, SEDB200_EINVAL, "window_batch: n_masks=%d", n_masks);
    SED_REQUIRE(n_masks == 0 || tmask_dev || fmask_dev, SEDB200_EINVAL, "window_batch: masks requested, no offsets");
    if (y_dev) {
        SED_REQUIRE(lab_dev && n_lab >= 1, SEDB200_EINVAL, "window_batch: labels requested without a label matrix");
        SED_REQUIRE(seq_out >= 1 && seq_in % seq_out == 0, SEDB200_ESHAPE,
                    "window_batch: seq_in=%d is not a multiple of seq_out=%d", seq_in, seq_out);
    }
    if (batch == 0) return SEDB200_OK;
    SED_REQUIRE(mel_dev && starts_dev && x_dev, SEDB200_EINVAL, "window_batch: null buffer");
    SED_REQUIRE(batch <= 65535, SEDB200_ESHAPE, "window_batch: batch=%d", batch);
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    dim3 grid((seq_in + kTile - 1) / kTile, n_ch * ((n_feat + kTile - 1) / kTile), batch);
    window_batch_kernel<<<grid, kTile * 8, 0, st>>>(mel_dev, lab_dev, n_frames, n_ch, n_feat, n_lab, starts_dev,
                                                    seq_in, seq_out, n_masks ? tmask_dev : nullptr,
                                                    n_masks ? fmask_dev : nullptr, n_masks, time_mask_w, freq_mask_w,
                                                    layout, x_dev, y_dev);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int sedb200_clean_negatives(const float* lab_dev, long n_frames, int n_lab, int seq_in, unsigned char* flag_dev,
                            void* stream) {
    SED_REQUIRE(n_lab >= 1 && seq_in >= 1, SEDB200_EINVAL, "clean_negatives: n_lab=%d seq_in=%d", n_lab, seq_in);
    const long n_starts = n_frames - seq_in + 1;
    if (n_starts <= 0) return SEDB200_OK;
    SED_REQUIRE(lab_dev && flag_dev, SEDB200_EINVAL, "clean_negatives: null buffer");
    int rc = require_sm100();
    if (rc) return rc;
    const int blocks = (int)std::min<long>((n_starts + 255) / 256, 4L * sm_count());
    clean_negatives_kernel<<<blocks, 256, 0, as_stream(stream)>>>(lab_dev, n_frames, n_lab, seq_in, flag_dev, n_starts);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

int sedb200_rasterize_labels(const double* start_s_dev, const double* end_s_dev, int n_events, int sr, int hop,
                             long n_frames, int n_lab, int col, float* lab_dev, void* stream) {
    SED_REQUIRE(n_events >= 0 && sr > 0 && hop > 0 && n_lab >= 1 && col >= 0 && col < n_lab, SEDB200_EINVAL,
                "rasterize_labels: n_events=%d sr=%d hop=%d n_lab=%d col=%d", n_events, sr, hop, n_lab, col);
    if (n_events == 0 || n_frames <= 0) return SEDB200_OK;
    SED_REQUIRE(start_s_dev && end_s_dev && lab_dev, SEDB200_EINVAL, "rasterize_labels: null buffer");
    int rc = require_sm100();
    if (rc) return rc;
    const int blocks = (int)std::min<long>((n_frames + 255) / 256, 4L * sm_count());
    rasterize_labels_kernel<<<blocks, 256, 0, as_stream(stream)>>>(start_s_dev, end_s_dev, n_events, (double)sr,
                                                                   (double)hop, n_frames, n_lab, col, lab_dev);
    SED_POST_LAUNCH();
    return SEDB200_OK;
}

}  // extern "C"
