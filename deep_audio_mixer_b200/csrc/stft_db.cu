// K1: fused framing + reflect pad + window + 2048-point real FFT + |X|^2 + dB.
//
// Replaces MultitrackAudioDataset.compute_features (reference data/dataset.py:132-162:
// torch.stft -> torch.abs -> amplitude_to_DB), with _stereo_to_mono (data/dataset.py:181-183)
// and _augment_audio (data/dataset.py:164-168) folded into the staging prologue.  The complex
// spectrogram never touches HBM: bytes in = the samples, bytes out = the dB map.
//
// Decomposition of one 2048-sample frame (n = 32a + b, a < 64, b < 32):
//   stage 1  thread (frame, b): 64 real samples x[32a+b]*w[32a+b], in registers
//            -> 64-point real DFT Y_b[k'], k' = 0..32 (packed 32-point complex codelet)
//   exchange through shared memory (the only one)
//   stage 2  thread (frame, k'): 32 values Y_b[k'] * W_2048^(b k'), in registers
//            -> 32-point complex DFT over b -> X[k' + 64 q], q = 0..31
//            bins above 1024 are conjugate mirrors (|X| is what we need), so the 33 tasks
//            k' = 0..32 produce each of the 1025 bins exactly once.
//   Lanes of a stage-2 warp hold 8 consecutive frames x 4 consecutive k', so every store of a
//   bin row is 8 consecutive t (32 B) and no output transposition buffer is needed.
#include "dam_common.cuh"
#include "fft_codelets.cuh"

#include <math.h>
#include <mutex>
#include <vector>

namespace {

constexpr int kNfft = 2048;
constexpr int kBins = kNfft / 2 + 1;
constexpr int kFT = 8;            // frames per tile
constexpr int kThreads = 256;     // 8 warps: stage 1 one warp per frame; stage 2 runs 264 (k', frame) tasks
constexpr int kRow = 34;          // complex elements per k' row in the exchange / twiddle tables
constexpr int kKp = 33;           // stage-2 tasks per frame

template <typename R> struct Cx;
template <> struct Cx<float>  { using type = float2;  static constexpr int kXch = 32 * kRow + 2; };  // /2 odd
template <> struct Cx<double> { using type = double2; static constexpr int kXch = 32 * kRow + 1; };  // odd

struct StftParams {
    const void* audio;
    int audio_dtype;
    int channels;
    long long n_rows, n_samples, row_stride, chan_stride, samp_stride;
    double input_scale;
    const double* row_gain;
    const float* window;
    const void* twiddle;          // device table, Cx<R>::type [33][34]
    int hop;
    long long T;                  // frames per row
    long long tiles_per_row, n_tiles;
    float amin2;                  // amin^2 (power floor)
    float log_scale;              // multiplier/2 * log10(2)
    double log_scale_d;
    void* out;
    double* stats;
};

template <typename P>
__device__ __forceinline__ P load_sample(const void* base, int dtype, long long idx) {
    if (dtype == DAM_F32) return (P)__ldg(reinterpret_cast<const float*>(base) + idx);
    if (dtype == DAM_I16) return (P)__ldg(reinterpret_cast<const short*>(base) + idx);
    return (P)__ldg(reinterpret_cast<const double*>(base) + idx);
}

// One staged sample: scale, channel mean and row gain in the source precision P (what numpy
// does on the reference path), then converted to the compute type.
template <typename P>
__device__ __forceinline__ P staged_value(const StftParams& p, long long row_off, long long g, P gain,
                                          bool has_gain) {
    const long long off = row_off + g * p.samp_stride;
    P acc = load_sample<P>(p.audio, p.audio_dtype, off) * (P)p.input_scale;
    if (p.channels > 1) {
        for (int c = 1; c < p.channels; ++c)
            acc += load_sample<P>(p.audio, p.audio_dtype, off + c * p.chan_stride) * (P)p.input_scale;
        acc = acc / (P)p.channels;
    }
    if (has_gain) acc = gain * acc;
    return acc;
}

template <typename R, typename OUT>
__global__ void __launch_bounds__(kThreads, sizeof(R) == 4 ? 2 : 1)
stft_db_kernel(const StftParams p) {
    using C = typename Cx<R>::type;
    constexpr int kXch = Cx<R>::kXch;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* s_win = reinterpret_cast<R*>(smem_raw);                    // [2048]
    C* s_tw = reinterpret_cast<C*>(s_win + kNfft);                // [33][34]
    C* s_xch = s_tw + kKp * kRow;                                 // [kFT][kXch]
    R* s_stage = reinterpret_cast<R*>(s_xch);                     // aliases s_xch: [(kFT-1)*hop + 2048]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < kNfft; i += kThreads) s_win[i] = (R)__ldg(p.window + i);
    {
        const C* tw = reinterpret_cast<const C*>(p.twiddle);
        for (int i = tid; i < kKp * kRow; i += kThreads) s_tw[i] = tw[i];
    }

    const int hop = p.hop;
    const long long T = p.T;
    double st_sum = 0.0, st_sq = 0.0, st_cnt = 0.0;

    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long row = tile / p.tiles_per_row;
        const long long t0 = (tile - row * p.tiles_per_row) * kFT;
        const int nf = (int)((T - t0) < kFT ? (T - t0) : kFT);

        __syncthreads();     // previous tile's stage 2 has finished reading s_xch (and tables are loaded)

        // ---- staging: samples [t0*hop - 1024, (t0+nf-1)*hop + 1024) with reflection at the row ends
        {
            const long long row_off = row * p.row_stride;
            const long long g0 = t0 * hop - kNfft / 2;
            const int span = (nf - 1) * hop + kNfft;
            const long long n = p.n_samples;
            const bool has_gain = p.row_gain != nullptr;
            const double gain_d = has_gain ? p.row_gain[row] : 1.0;
            if (p.audio_dtype == DAM_F64) {
                for (int i = tid; i < span; i += kThreads) {
                    long long g = g0 + i;
                    g = g < 0 ? -g : g;
                    g = g >= n ? 2 * (n - 1) - g : g;
                    s_stage[i] = (R)staged_value<double>(p, row_off, g, gain_d, has_gain);
                }
            } else {
                const float gain_f = (float)gain_d;
                for (int i = tid; i < span; i += kThreads) {
                    long long g = g0 + i;
                    g = g < 0 ? -g : g;
                    g = g >= n ? 2 * (n - 1) - g : g;
                    s_stage[i] = (R)staged_value<float>(p, row_off, g, gain_f, has_gain);
                }
            }
        }
        __syncthreads();

        // ---- stage 1: warp = frame, lane = b
        R sr[32], si[32];
        {
            R x[64];
            if (warp < nf) {
                const R* f = s_stage + warp * hop + lane;
#pragma unroll
                for (int a = 0; a < 64; ++a) x[a] = f[32 * a] * s_win[32 * a + lane];
            }
            __syncthreads();     // all frames are in registers; s_stage may now be overwritten
            if (warp < nf) {
                dam_rfft64_packed<R>(x, sr, si);
                C* dst = s_xch + warp * kXch + lane;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    C v; v.x = sr[k]; v.y = si[k];
                    dst[k * kRow] = v;
                }
            }
        }
        __syncthreads();

        // ---- stage 2: task = (k' = task / 8, frame = task % 8); tasks 256..263 (k' = 32) take a
        //      second round on the first 8 threads
        for (int task = tid; task < kKp * kFT; task += kThreads) {
            const int kp = task >> 3, t = task & 7;
            if (t >= nf) continue;
            const C* src = s_xch + t * kXch + (kp == 32 ? 0 : kp) * kRow;
            const C* tw = s_tw + kp * kRow;
            const bool real_lo = (kp == 0), real_hi = (kp == 32);
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const C v = src[b];
                const C w = tw[b];
                const R vr = real_hi ? v.y : v.x;
                const R vi = (real_lo || real_hi) ? R(0) : v.y;
                sr[b] = vr * w.x - vi * w.y;
                si[b] = vr * w.y + vi * w.x;
            }
            dam_cfft32<R>(sr, si);

            OUT* o = reinterpret_cast<OUT*>(p.out) + (row * kBins) * T + t0 + t;
            const bool mirror_ok = !(real_lo || real_hi);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int f = q < 16 ? kp + 64 * q : kNfft - kp - 64 * q;
                const bool valid = q < 16 ? true : (q == 16 ? (mirror_ok || real_lo) : mirror_ok);
                float db;
                if (sizeof(R) == 4) {
                    float pw = fmaf((float)sr[q], (float)sr[q], (float)si[q] * (float)si[q]);
                    pw = fmaxf(pw, p.amin2);
                    db = p.log_scale * __log2f(pw);
                    if (valid) o[(long long)f * T] = (OUT)db;
                } else {
                    // |X|^2 in double, log2 split into exponent + mantissa so the only fp32 error is
                    // lg2.approx on [1, 2) (2^-22 absolute); recombined in double.
                    float pw = (float)(sr[q] * sr[q] + si[q] * si[q]);
                    pw = fmaxf(pw, p.amin2);
                    const int bits = __float_as_int(pw);
                    const int e = (bits >> 23) - 127;
                    const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
                    const double l2 = (double)e + (double)__log2f(m);
                    const double dbd = p.log_scale_d * l2;
                    db = (float)dbd;
                    if (valid) o[(long long)f * T] = (OUT)dbd;
                }
                if (p.stats != nullptr && valid) {
                    st_sum += db; st_sq += (double)db * db; st_cnt += 1.0;
                }
            }
        }
    }

    if (p.stats != nullptr) {
        st_sum = dam_warp_sum(st_sum); st_sq = dam_warp_sum(st_sq); st_cnt = dam_warp_sum(st_cnt);
        if (lane == 0 && st_cnt > 0.0) {
            atomicAdd(p.stats + 0, st_sum);
            atomicAdd(p.stats + 1, st_sq);
            atomicAdd(p.stats + 2, st_cnt);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Per-device twiddle tables T[k'][b] = s(k') * exp(-2 pi i b k' / 2048), s = 1 for k' in {0, 32},
// 1/2 otherwise (the stage-1 codelet emits 2*Y for those slots).  Built once, in long double.
struct TwiddleTables {
    float2* f32 = nullptr;
    double2* f64 = nullptr;
};

int get_twiddles(TwiddleTables** out) {
    static std::mutex mu;
    static TwiddleTables tables[64];
    int dev = 0;
    DAM_CHECK_CUDA(cudaGetDevice(&dev));
    DAM_CHECK_ARG(dev >= 0 && dev < 64, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    TwiddleTables& t = tables[dev];
    if (t.f32 == nullptr) {
        std::vector<float2> hf(kKp * kRow, make_float2(0.f, 0.f));
        std::vector<double2> hd(kKp * kRow, make_double2(0.0, 0.0));
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int kp = 0; kp < kKp; ++kp) {
            const long double s = (kp == 0 || kp == 32) ? 1.0L : 0.5L;
            for (int b = 0; b < 32; ++b) {
                const int j = (b * kp) % kNfft;                 // angle index, exact
                long double c, sn;
                // octant-exact values for multiples of 256
                if (j % 512 == 0) {
                    const int q = j / 512;
                    c = (q == 0) ? 1.0L : (q == 2 ? -1.0L : 0.0L);
                    sn = (q == 1) ? 1.0L : (q == 3 ? -1.0L : 0.0L);
                } else {
                    const long double ang = two_pi * (long double)j / (long double)kNfft;
                    c = cosl(ang);
                    sn = sinl(ang);
                }
                hd[kp * kRow + b] = make_double2((double)(s * c), (double)(-s * sn));
                hf[kp * kRow + b] = make_float2((float)(s * c), (float)(-s * sn));
            }
        }
        float2* df = nullptr;
        double2* dd = nullptr;
        DAM_CHECK_CUDA(cudaMalloc(&df, hf.size() * sizeof(float2)));
        DAM_CHECK_CUDA(cudaMalloc(&dd, hd.size() * sizeof(double2)));
        DAM_CHECK_CUDA(cudaMemcpy(df, hf.data(), hf.size() * sizeof(float2), cudaMemcpyHostToDevice));
        DAM_CHECK_CUDA(cudaMemcpy(dd, hd.data(), hd.size() * sizeof(double2), cudaMemcpyHostToDevice));
        t.f32 = df;
        t.f64 = dd;
    }
    *out = &t;
    return DAM_OK;
}

template <typename R>
size_t smem_bytes(int hop) {
    using C = typename Cx<R>::type;
    const size_t xch = (size_t)kFT * Cx<R>::kXch * sizeof(C);
    const size_t stage = ((size_t)(kFT - 1) * hop + kNfft) * sizeof(R);
    return kNfft * sizeof(R) + (size_t)kKp * kRow * sizeof(C) + (xch > stage ? xch : stage);
}

template <typename R, typename OUT>
int launch(const StftParams& p, cudaStream_t stream) {
    const size_t smem = smem_bytes<R>(p.hop);
    auto kern = stft_db_kernel<R, OUT>;
    static std::mutex mu;
    static bool configured[64] = {};
    static int blocks_per_sm[64] = {};
    int dev = 0;
    DAM_CHECK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[dev]) {
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured[dev] = true;
        }
    }
    int occ = 0;
    DAM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    DAM_CHECK_ARG(occ > 0, "stft kernel does not fit on an SM (hop=%d needs %zu B shared memory)", p.hop, smem);
    (void)blocks_per_sm;
    long long grid = (long long)dam_sm_count() * occ;
    if (grid > p.n_tiles) grid = p.n_tiles;
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(p);
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

}  // namespace

extern "C" int dam_stft_db(const void* audio, int audio_dtype, int channels,
                           int64_t n_rows, int64_t n_samples,
                           int64_t row_stride, int64_t chan_stride, int64_t samp_stride,
                           double input_scale, const double* row_gain,
                           const float* window, int n_fft, int hop,
                           double amin, double multiplier, int compute_dtype,
                           void* out, int out_dtype, double* stats, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(audio != nullptr && window != nullptr && out != nullptr, "null pointer argument");
    DAM_CHECK_ARG(audio_dtype == DAM_I16 || audio_dtype == DAM_F32 || audio_dtype == DAM_F64,
                  "audio_dtype %d is not one of I16/F32/F64", audio_dtype);
    DAM_CHECK_ARG(compute_dtype == DAM_F32 || compute_dtype == DAM_F64, "compute_dtype must be F32 or F64");
    DAM_CHECK_ARG(out_dtype == DAM_F32 || (out_dtype == DAM_F64 && compute_dtype == DAM_F64),
                  "out_dtype must be F32, or F64 together with compute_dtype F64");
    DAM_CHECK_ARG(channels >= 1 && channels <= 64, "channels=%d out of range [1, 64]", channels);
    DAM_CHECK_ARG(n_rows >= 0 && n_samples >= 0, "negative size");
    if (n_fft != kNfft) {
        dam_set_error("n_fft=%d is not supported (this build instantiates n_fft=2048, the only size the "
                      "reference uses: data/dataset.py:132)", n_fft);
        return DAM_ERR_UNSUPPORTED;
    }
    DAM_CHECK_ARG(hop >= 1 && hop <= n_fft, "hop=%d must be in [1, n_fft]", hop);
    DAM_CHECK_ARG(amin >= 0.0, "amin must be non-negative");
    if (n_rows == 0) return DAM_OK;
    DAM_CHECK_ARG(n_samples > n_fft / 2,
                  "reflect padding needs n_samples > n_fft/2 (got %lld)", (long long)n_samples);

    TwiddleTables* tw = nullptr;
    int rc = get_twiddles(&tw);
    if (rc != DAM_OK) return rc;

    StftParams p;
    p.audio = audio; p.audio_dtype = audio_dtype; p.channels = channels;
    p.n_rows = n_rows; p.n_samples = n_samples;
    p.row_stride = row_stride; p.chan_stride = chan_stride; p.samp_stride = samp_stride;
    p.input_scale = input_scale; p.row_gain = row_gain; p.window = window;
    p.hop = hop; p.T = 1 + n_samples / hop;
    p.tiles_per_row = (p.T + kFT - 1) / kFT;
    p.n_tiles = p.tiles_per_row * n_rows;
    p.amin2 = (float)(amin * amin);
    p.log_scale_d = multiplier * 0.5 * 0.30102999566398119521373889472449302676819;
    p.log_scale = (float)p.log_scale_d;
    p.out = out; p.stats = stats;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (compute_dtype == DAM_F32) {
        p.twiddle = tw->f32;
        return launch<float, float>(p, s);
    }
    p.twiddle = tw->f64;
    if (out_dtype == DAM_F32) return launch<double, float>(p, s);
    return launch<double, double>(p, s);
}
