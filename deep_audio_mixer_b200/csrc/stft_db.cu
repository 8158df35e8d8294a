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


#include "stft_common.cuh"

#include <map>

namespace {

template <typename R, typename OUT, bool kStats, int kFT, bool kPc = false>
__global__ void __launch_bounds__(32 * kFT, sizeof(R) == 4 ? 16 / kFT : 1)
stft_db_kernel(const StftParams p) {
    using C = typename Cx<R>::type;
    constexpr int kXch = XchStride<R, kFT>::v;
    constexpr int kThreads = 32 * kFT;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    R* s_win = reinterpret_cast<R*>(smem_raw);                    // [2048]
    C* s_tw = reinterpret_cast<C*>(s_win + kNfft);                // [33][34]
    R* s_nq = reinterpret_cast<R*>(s_tw + kKp * kRow);            // [8][32]
    C* s_xch = reinterpret_cast<C*>(s_nq + kNyqTab);              // [kFT][kXch]
    R* s_stage = reinterpret_cast<R*>(s_xch);                     // aliases s_xch: [(kFT-1)*hop + 2048]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < kNfft; i += kThreads) s_win[i] = (R)__ldg(p.window + i);
    {
        const C* tw = reinterpret_cast<const C*>(p.twiddle);
        for (int i = tid; i < kKp * kRow; i += kThreads) s_tw[i] = tw[i];
    }
    nyq_table_init<R>(s_nq, tid);
    const R* nq = s_nq;

    const int hop = p.hop;
    const long long T = p.T;
    const int elem_bytes = p.audio_dtype == DAM_I16 ? 2 : (p.audio_dtype == DAM_F32 ? 4 : 8);
    const unsigned tiles_per_row = (unsigned)p.tiles_per_row, n_tiles = (unsigned)p.n_tiles;
    StatAcc st;

    // Tiles are dealt round-robin, so at any moment the resident CTAs work on ADJACENT tiles: their 32-byte
    // output row segments complete whole DRAM sectors / lines in L2 together and their input halos hit in
    // L2.  (A contiguous tile range per CTA was measured 5-40 % slower: every CTA then scatters 1025
    // isolated 32-byte segments per tile with nobody writing their neighbours.)
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned urow = tile / tiles_per_row;
        const long long row = urow;
        const long long t0 = (long long)(tile - urow * tiles_per_row) * kFT;
        const int nf = (int)((T - t0) < kFT ? (T - t0) : kFT);
        OUT* out_row = reinterpret_cast<OUT*>(p.out) + row * p.out_row_stride + t0;

        __syncthreads();     // previous tile's stage 2 has finished reading s_xch (and tables are loaded)

        // ---- staging: samples [t0*hop - 1024, (t0+nf-1)*hop + 1024) with reflection at the row ends
        {
            const long long g0 = t0 * hop - kNfft / 2;
            const int span = (nf - 1) * hop + kNfft;
            if (p.audio_dtype == DAM_F32)      stage_tile<R, float, float, kThreads>(p, s_stage, row, g0, span, tid);
            else if (p.audio_dtype == DAM_I16) {
                // PCM16 prologue in the COMPUTE precision: x / 32768 and the mean of two channels are exact in
                // float32, so float32 compute keeps the float32 prologue (librosa.load semantics); float64 compute
                // is the soundfile path (sf.read yields float64), where the augmentation gain multiplies a double
                if (sizeof(R) == 8) stage_tile<R, double, short, kThreads>(p, s_stage, row, g0, span, tid);
                else stage_tile<R, float, short, kThreads>(p, s_stage, row, g0, span, tid);
            }
            else                               stage_tile<R, double, double, kThreads>(p, s_stage, row, g0, span, tid);
        }
        __syncthreads();

        // pull the next tile's samples towards L2 while this one is transformed
        if (p.l2_prefetch) {
            const unsigned nt = tile + gridDim.x;
            if (nt < n_tiles) {
                const unsigned nrow = nt / tiles_per_row;
                const long long nt0 = (long long)(nt - nrow * tiles_per_row) * kFT;
                const int nnf = (int)((T - nt0) < kFT ? (T - nt0) : kFT);
                prefetch_tile<kThreads>(p, nrow, nt0 * hop - kNfft / 2, (nnf - 1) * hop + kNfft, elem_bytes, tid);
            }
        }

        // ---- stage 1: warp = frame, lane = b
        if constexpr (kPc) {
            // (x[2m], x[2m+1]) of the lane's 64 samples x[a] = frame[32a + lane].  Loaded by every warp (the
            // idle warps of a row's last tile re-read its last frame): assigning them under `warp < nf` makes
            // the registers loop-carried and ptxas spills part of them to local memory.
            pk2 xp[32];
            {
                const float* f = reinterpret_cast<const float*>(s_stage) + (warp < nf ? warp : nf - 1) * hop + lane;
                const float* w = reinterpret_cast<const float*>(s_win) + lane;
#pragma unroll
                for (int m = 0; m < 32; ++m)
                    xp[m] = pk_mul(pk_make(f[64 * m], f[64 * m + 32]), pk_make(w[64 * m], w[64 * m + 32]));
            }
            __syncthreads();     // all frames are in registers; s_stage may now be overwritten
            if (warp < nf)
                stage1_frame_pc<OUT, kStats>(xp, reinterpret_cast<float2*>(s_xch) + warp * kXch + lane, out_row + warp, p.bin_stride,
                                             nq, p, lane, st);
        } else {
            R x[64];
            {
                const R* f = s_stage + (warp < nf ? warp : nf - 1) * hop + lane;
#pragma unroll
                for (int a = 0; a < 64; ++a) x[a] = f[32 * a] * s_win[32 * a + lane];
            }
            __syncthreads();     // all frames are in registers; s_stage may now be overwritten
            if (warp < nf)
                stage1_frame<R, OUT, kStats>(x, s_xch + warp * kXch + lane, out_row + warp, p.bin_stride, nq, p, lane, st);
        }
        __syncthreads();

        // ---- stage 2: thread = (k' = tid / FT in 0..31, frame = tid % FT)
        {
            const int kp = tid / kFT, t = tid % kFT;
            if (t < nf) {
                if constexpr (kPc)
                    stage2_task_pc<OUT, kStats>(reinterpret_cast<const float2*>(s_xch) + t * kXch + kp * kRow,
                                                reinterpret_cast<const float2*>(s_tw) + kp * kRow, out_row + t, p.bin_stride, kp, p, st);
                else
                    stage2_task<R, OUT, kStats>(s_xch + t * kXch + kp * kRow, s_tw + kp * kRow, out_row + t, p.bin_stride, kp, p, st);
            }
        }
    }
    if (kStats) st.flush(p.stats, lane);
}

// ==========================================================================================
// Generic window sizes (power of two, 64..8192 except 2048): a plain shared-memory radix-2 transform,
// one CTA per (row, tile of 8 frames).  The reference only ever uses window_size = 2048
// (data/dataset.py:132), so this path exists for API completeness (window_size is a runtime argument
// of compute_features), not for speed: same staging semantics, same dB epilogue, same output layout.
constexpr int kGenThreads = 256;     // frames per tile: 8, or fewer when the dB staging would not fit

template <typename R>
size_t generic_smem_bytes(int n_fft, int ft) {
    using C = typename Cx<R>::type;
    return (size_t)n_fft * sizeof(C) + (size_t)(n_fft / 2) * sizeof(C) + (size_t)n_fft * sizeof(R) +
           (size_t)(n_fft / 2 + 1) * ft * sizeof(float);
}

template <typename R, typename OUT>
__global__ void __launch_bounds__(kGenThreads)
stft_db_generic_kernel(const StftParams p, int n_fft, int log2n, int kGenFT) {
    using C = typename Cx<R>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* s_buf = reinterpret_cast<C*>(smem_raw);                    // [n_fft] in-place transform
    C* s_twd = s_buf + n_fft;                                     // [n_fft/2] exp(-2 pi i k / n_fft)
    R* s_win = reinterpret_cast<R*>(s_twd + n_fft / 2);           // [n_fft]
    float* s_db = reinterpret_cast<float*>(s_win + n_fft);        // [n_fft/2+1][kGenFT]
    const int tid = threadIdx.x;
    const int bins = n_fft / 2 + 1;
    for (int i = tid; i < n_fft; i += kGenThreads) s_win[i] = (R)__ldg(p.window + i);
    for (int i = tid; i < n_fft / 2; i += kGenThreads) {
        double sn, cs;
        sincospi(-2.0 * (double)i / (double)n_fft, &sn, &cs);
        C w; w.x = (R)cs; w.y = (R)sn;
        s_twd[i] = w;
    }
    const long long row = blockIdx.y;
    const long long t0 = (long long)blockIdx.x * kGenFT;
    const long long T = p.T, n = p.n_samples;
    const int nf = (int)((T - t0) < kGenFT ? (T - t0) : kGenFT);
    const long long row_off = row_offset(p, row);
    const bool f64src = p.audio_dtype == DAM_F64;
    const double gain = p.row_gain != nullptr ? p.row_gain[row] : 1.0;
    __syncthreads();
    for (int t = 0; t < nf; ++t) {
        // windowed frame in bit-reversed order (prologue in the source precision, like the fast path)
        for (int i = tid; i < n_fft; i += kGenThreads) {
            long long g = (t0 + t) * p.hop - n_fft / 2 + i;
            g = g < 0 ? -g : g;
            g = g >= n ? 2 * (n - 1) - g : g;
            const long long off = row_off + g * p.samp_stride;
            R v;
            if (f64src) {
                double acc = __ldg(reinterpret_cast<const double*>(p.audio) + off) * p.input_scale;
                for (int c = 1; c < p.channels; ++c)
                    acc += __ldg(reinterpret_cast<const double*>(p.audio) + off + c * p.chan_stride) * p.input_scale;
                if (p.channels > 1) acc = acc / (double)p.channels;
                if (p.row_gain != nullptr) acc = gain * acc;
                v = (R)acc;
            } else {
                const float sc = (float)p.input_scale;
                float acc = p.audio_dtype == DAM_F32 ? __ldg(reinterpret_cast<const float*>(p.audio) + off) * sc
                                                     : (float)__ldg(reinterpret_cast<const short*>(p.audio) + off) * sc;
                for (int c = 1; c < p.channels; ++c) {
                    const long long o2 = off + c * p.chan_stride;
                    acc += p.audio_dtype == DAM_F32 ? __ldg(reinterpret_cast<const float*>(p.audio) + o2) * sc
                                                    : (float)__ldg(reinterpret_cast<const short*>(p.audio) + o2) * sc;
                }
                if (p.channels > 1) acc = acc / (float)p.channels;
                if (p.row_gain != nullptr) acc = (float)gain * acc;
                v = (R)acc;
            }
            C z; z.x = v * s_win[i]; z.y = R(0);
            s_buf[__brev((unsigned)i) >> (32 - log2n)] = z;
        }
        __syncthreads();
        for (int st = 0; st < log2n; ++st) {
            const int half = 1 << st;
            for (int j = tid; j < n_fft / 2; j += kGenThreads) {
                const int k = j & (half - 1);
                const int i0 = ((j >> st) << (st + 1)) + k;
                const C w = s_twd[k << (log2n - 1 - st)];
                const C a = s_buf[i0], b = s_buf[i0 + half];
                C tw; tw.x = b.x * w.x - b.y * w.y; tw.y = b.x * w.y + b.y * w.x;
                C u, d; u.x = a.x + tw.x; u.y = a.y + tw.y; d.x = a.x - tw.x; d.y = a.y - tw.y;
                s_buf[i0] = u; s_buf[i0 + half] = d;
            }
            __syncthreads();
        }
        for (int f = tid; f < bins; f += kGenThreads)
            s_db[f * kGenFT + t] = (float)DbOut<R>::eval(s_buf[f].x, s_buf[f].y, p);
        __syncthreads();
    }
    // coalesced-ish store: runs of nf consecutive frames per bin row
    OUT* out_row = reinterpret_cast<OUT*>(p.out) + row * p.out_row_stride + t0;
    for (int i = tid; i < bins * kGenFT; i += kGenThreads) {
        const int f = i / kGenFT, t = i % kGenFT;
        if (t < nf) out_row[(long long)f * p.bin_stride + t] = (OUT)s_db[i];
    }
    if (p.stats != nullptr) {
        double sm = 0.0, sq = 0.0, cnt = 0.0;
        for (int i = tid; i < bins * kGenFT; i += kGenThreads)
            if (i % kGenFT < nf) { const double v = s_db[i]; sm += v; sq += v * v; cnt += 1.0; }
        sm = dam_warp_sum(sm); sq = dam_warp_sum(sq); cnt = dam_warp_sum(cnt);
        if ((tid & 31) == 0 && cnt > 0.0) { atomicAdd(p.stats, sm); atomicAdd(p.stats + 1, sq); atomicAdd(p.stats + 2, cnt); }
    }
}

template <typename R, typename OUT>
int launch_generic(const StftParams& p, int n_fft, cudaStream_t stream) {
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    int kGenFT = 8;
    while (kGenFT > 1 && generic_smem_bytes<R>(n_fft, kGenFT) > 227 * 1024) kGenFT >>= 1;
    const size_t smem = generic_smem_bytes<R>(n_fft, kGenFT);
    if (smem > 227 * 1024) {
        dam_set_error("n_fft=%d with %s arithmetic needs %zu B of shared memory (float64 compute supports n_fft <= 4096)",
                      n_fft, sizeof(R) == 8 ? "float64" : "float32", smem);
        return DAM_ERR_UNSUPPORTED;
    }
    auto kern = stft_db_generic_kernel<R, OUT>;
    DAM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tiles = (p.T + kGenFT - 1) / kGenFT;
    DAM_CHECK_ARG(tiles < 0x7fffffffLL && p.n_rows <= 65535, "generic path: too many tiles or rows (> 65535 rows)");
    dim3 grid((unsigned)tiles, (unsigned)p.n_rows);
    { kern<<<grid, kGenThreads, smem, stream>>>(p, n_fft, log2n, kGenFT); dam_note_launch(); }
    dam_k1_last_kernel_name = sizeof(R) == 8 ? "stft_db_generic_kernel<float64>" : "stft_db_generic_kernel<float32>";
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
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
        // [8][32] per-lane constants of the cross-lane DFT (nyq_table_init) behind the float table, for the
        // kernel variant that keeps its tables in global memory
        {
            const long double pi = 3.141592653589793238462643383279502884L;
            auto snap = [](long double v) { return fabsl(v) < 1e-17L ? 0.0L : v; };     // cos/sin(pi/2 k) exactly
            std::vector<float> nq(kNyqTab, 0.f);
            for (int lane = 0; lane < 32; ++lane) {
                nq[0 * 32 + lane] = (float)snap(cosl(-pi * lane / 32.0L));
                nq[1 * 32 + lane] = (float)snap(sinl(-pi * lane / 32.0L));
                for (int i = 0; i < 3; ++i) {
                    const int dist = 16 >> i;
                    const bool upper = (lane & dist) != 0;
                    const long double a = upper ? -pi * (long double)((lane & (dist - 1)) * (16 / dist)) / 16.0L : 0.0L;
                    nq[(2 + 2 * i) * 32 + lane] = (float)snap(cosl(a));
                    nq[(3 + 2 * i) * 32 + lane] = (float)snap(sinl(a));
                }
            }
            for (int i = 0; i < kNyqTab / 2; ++i) hf.push_back(make_float2(nq[2 * i], nq[2 * i + 1]));
        }
        float2* df = nullptr;
        double2* dd = nullptr;
        DAM_CHECK_CUDA(cudaMalloc(&df, hf.size() * sizeof(float2)));
        cudaError_t e = cudaMalloc(&dd, hd.size() * sizeof(double2));
        if (e == cudaSuccess) e = cudaMemcpy(df, hf.data(), hf.size() * sizeof(float2), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(dd, hd.data(), hd.size() * sizeof(double2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(df);
            cudaFree(dd);
            dam_set_error("building the twiddle tables failed: %s", cudaGetErrorString(e));
            return DAM_ERR_CUDA;
        }
        t.f32 = df;
        t.f64 = dd;
    }
    *out = &t;
    return DAM_OK;
}

template <typename R, int kFT>
size_t smem_bytes(int hop) {
    using C = typename Cx<R>::type;
    const size_t xch = (size_t)kFT * XchStride<R, kFT>::v * sizeof(C);
    const size_t stage = ((size_t)(kFT - 1) * hop + kNfft) * sizeof(R);
    return kNfft * sizeof(R) + (size_t)kKp * kRow * sizeof(C) + kNyqTab * sizeof(R) + (xch > stage ? xch : stage);
}

template <typename R, typename OUT, bool kStats, int kFT, bool kPc = false>
int launch(StftParams p, cudaStream_t stream) {
    constexpr int kThreads = 32 * kFT;
    p.tiles_per_row = (p.T + kFT - 1) / kFT;
    p.n_tiles = p.tiles_per_row * p.n_rows;
    DAM_CHECK_ARG(p.n_tiles < 0x7fffffffLL, "too many frame tiles for one launch (%lld); split the batch",
                  (long long)p.n_tiles);
    static const int pf_env = [] { const char* e = getenv("DAM_K1_PREFETCH"); return e ? atoi(e) : 1; }();
    p.l2_prefetch = pf_env;
    const size_t smem = smem_bytes<R, kFT>(p.hop);
    auto kern = stft_db_kernel<R, OUT, kStats, kFT, kPc>;
    static std::mutex mu;
    static bool configured[64] = {};
    int dev = 0;
    DAM_CHECK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[dev]) {
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            configured[dev] = true;
        }
    }
    // resident CTAs per SM for this shared-memory size (depends on the hop only): queried once per device and size
    int occ = 0;
    {
        static std::map<size_t, int> occ_cache[64];
        std::lock_guard<std::mutex> lock(mu);
        auto it = occ_cache[dev].find(smem);
        if (it == occ_cache[dev].end()) {
            DAM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
            occ_cache[dev][smem] = occ;
        } else {
            occ = it->second;
        }
    }
    DAM_CHECK_ARG(occ > 0, "stft kernel does not fit on an SM (hop=%d needs %zu B shared memory)", p.hop, smem);
    long long grid = (long long)dam_sm_count() * occ;
    if (grid > p.n_tiles) grid = p.n_tiles;
    static const bool trace = getenv("DAM_K1_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[dam] stft_db_kernel FT=%d pc=%d f%d smem=%zu occ=%d hop=%d C=%d T=%lld\n", kFT, (int)kPc,
                       (int)sizeof(R) * 8, smem, occ, p.hop, p.channels, (long long)p.T);
    { kern<<<(unsigned)grid, kThreads, smem, stream>>>(p); dam_note_launch(); }
    dam_k1_last_kernel_name = sizeof(R) == 8 ? "stft_db_kernel<float64, FT=8> (phase kernel)"
                                             : (kFT == 16 ? "stft_db_kernel<float32, FT=16> (phase kernel)" : "stft_db_kernel<float32, FT=8> (phase kernel)");
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

}  // namespace

// Creates the per-device constant tables of K1 (the only allocation + synchronous copy the library ever does for
// K1).  dam_stft_db does the same on its first call per device; call dam_init first when that first call would
// sit inside a CUDA-graph stream capture.
extern "C" int dam_init(void) {
    dam_clear_error();
    TwiddleTables* tw = nullptr;
    return get_twiddles(&tw);
}

extern "C" int dam_stft_db(const void* audio, int audio_dtype, int channels,
                           int64_t n_rows, int64_t n_samples,
                           int64_t row_stride, int64_t chan_stride, int64_t samp_stride,
                           int64_t n_inner, int64_t outer_stride,
                           double input_scale, const double* row_gain,
                           const float* window, int n_fft, int hop,
                           double amin, double multiplier, int compute_dtype,
                           void* out, int out_dtype, int64_t out_row_stride, int64_t out_bin_stride, double* stats,
                           void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(audio != nullptr && window != nullptr && out != nullptr, "null pointer argument");
    DAM_CHECK_ARG(audio_dtype == DAM_I16 || audio_dtype == DAM_F32 || audio_dtype == DAM_F64,
                  "audio_dtype %d is not one of I16/F32/F64", audio_dtype);
    DAM_CHECK_ARG(compute_dtype == DAM_F32 || compute_dtype == DAM_F64, "compute_dtype must be F32 or F64");
    DAM_CHECK_ARG(out_dtype == DAM_F32 || (out_dtype == DAM_F64 && compute_dtype == DAM_F64),
                  "out_dtype must be F32, or F64 together with compute_dtype F64");
    DAM_CHECK_ARG(channels >= 1 && channels <= 64, "channels=%d out of range [1, 64]", channels);
    DAM_CHECK_ARG(n_rows >= 0 && n_samples >= 0, "negative size");
    if (n_fft < 64 || n_fft > 8192 || (n_fft & (n_fft - 1)) != 0) {
        dam_set_error("n_fft=%d is not supported: window sizes are powers of two in [64, 8192] "
                      "(2048, the only size the reference uses, has the fused fast path)", n_fft);
        return DAM_ERR_UNSUPPORTED;
    }
    DAM_CHECK_ARG(hop >= 1 && hop <= n_fft, "hop=%d must be in [1, n_fft]", hop);
    DAM_CHECK_ARG(amin >= 0.0, "amin must be non-negative");
    DAM_CHECK_ARG(out_row_stride >= 0, "out_row_stride must be non-negative");
    DAM_CHECK_ARG((out_bin_stride > 0 ? out_bin_stride : 1 + n_samples / hop) < (1LL << 22),
                  "rows of more than 4M frames are not supported (split the signal)");
    DAM_CHECK_ARG(out_bin_stride == 0 || out_bin_stride >= 1 + n_samples / hop,
                  "out_bin_stride=%lld is smaller than the %lld frames of a row", (long long)out_bin_stride,
                  (long long)(1 + n_samples / hop));
    DAM_CHECK_ARG(n_inner >= 0, "n_inner must be non-negative");
    if (n_rows == 0) return DAM_OK;
    DAM_CHECK_ARG(n_samples > n_fft / 2,
                  "reflect padding needs n_samples > n_fft/2 (got %lld)", (long long)n_samples);

    StftParams p;
    p.audio = audio; p.audio_dtype = audio_dtype; p.channels = channels;
    p.n_rows = n_rows; p.n_samples = n_samples;
    p.row_stride = row_stride; p.chan_stride = chan_stride; p.samp_stride = samp_stride;
    p.n_inner = (n_inner > 0 && n_inner < n_rows) ? n_inner : 0; p.outer_stride = outer_stride;
    p.input_scale = input_scale; p.row_gain = row_gain; p.window = window;
    p.hop = hop; p.T = 1 + n_samples / hop;
    p.log_scale_d = multiplier * 0.5 * 0.30102999566398119521373889472449302676819;
    p.log_scale = (float)p.log_scale_d;
    p.floor_db_d = amin > 0.0 ? multiplier * log10(amin) : -INFINITY;
    p.floor_db = (float)p.floor_db_d;
    p.out = out; p.stats = stats;
    p.bin_stride = out_bin_stride > 0 ? out_bin_stride : p.T;
    DAM_CHECK_ARG(out_row_stride == 0 || out_row_stride >= (int64_t)(n_fft / 2 + 1) * p.bin_stride,
                  "out_row_stride=%lld is smaller than one plane (%lld elements)", (long long)out_row_stride,
                  (long long)(n_fft / 2 + 1) * p.bin_stride);
    p.out_row_stride = out_row_stride > 0 ? out_row_stride : (long long)(n_fft / 2 + 1) * p.bin_stride;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (n_fft != kNfft) {
        p.twiddle = nullptr;
        p.tiles_per_row = 0; p.n_tiles = 0; p.l2_prefetch = 0;
        if (compute_dtype == DAM_F32) return launch_generic<float, float>(p, n_fft, s);
        return out_dtype == DAM_F32 ? launch_generic<double, float>(p, n_fft, s) : launch_generic<double, double>(p, n_fft, s);
    }
    TwiddleTables* tw = nullptr;
    int rc = get_twiddles(&tw);
    if (rc != DAM_OK) return rc;
    const bool st = stats != nullptr;
    if (compute_dtype == DAM_F32) {
        p.twiddle = tw->f32;
        // TMA-staged kernel (stft_tma.cu) wherever the shape qualifies: float32 / PCM16 source, mono or two
        // channels, hop a multiple of 8 up to 1024, no row gain.  DAM_K1_ASYNC: 0 = never (phase kernel only),
        // 1 (default) = wherever it qualifies, 2 / 3 = with the tables forced into shared / global memory.
        static const int async_env = [] { const char* e = getenv("DAM_K1_ASYNC"); return e ? atoi(e) : 1; }();
        static const int ft_env = [] { const char* e = getenv("DAM_FT_F32"); return e ? atoi(e) : 0; }();
        if (async_env >= 1 && async_env <= 4) {
            const int force = async_env == 2 ? 1 : (async_env == 3 ? 0 : -1);
            // 16 frames per tile (one 16-warp CTA per SM, 64-byte store segments): +30 % when the output rows are not
            // a multiple of 8 frames, +3.6 % on long aligned rows at hop 1024; 8 frames per tile everywhere else
            const bool long_rows = p.T >= 160;
            const bool ft16 = ft_env == 16 || (ft_env == 0 && long_rows && ((p.bin_stride % 8) != 0 || hop >= 1024));
            int rc2 = ft16 ? dam_k1_launch_tma16(p, s, force) : dam_k1_launch_tma(p, s, force);
            if (rc2 == 1 && ft16) rc2 = dam_k1_launch_tma(p, s, force);
            if (rc2 <= 0) return rc2;
        }
        // Phase kernel (everything else: row gains, odd hops, strided samples, more than two channels).  Frames per
        // tile: 8 (two 8-warp CTAs per SM); 16 (one 16-warp CTA per SM) writes 64-byte row segments and measured
        // +15 % when rows are not sector aligned (T % 8 != 0) at hop >= 1024, -17 % when they are, -4 % at hop 512.
        // DAM_FT_F32 = 8 | 16 overrides the choice.
        int ft = 8;
        if (ft_env == 8 || ft_env == 16) ft = ft_env;
        else if (hop >= 1024 && (p.bin_stride % 8) != 0 && p.T >= 160) ft = 16;
        if (ft == 16) return st ? launch<float, float, true, 16, true>(p, s) : launch<float, float, false, 16, true>(p, s);
        return st ? launch<float, float, true, 8, true>(p, s) : launch<float, float, false, 8, true>(p, s);
    }
    p.twiddle = tw->f64;
    if (out_dtype == DAM_F32) return st ? launch<double, float, true, 8>(p, s) : launch<double, float, false, 8>(p, s);
    return st ? launch<double, double, true, 8>(p, s) : launch<double, double, false, 8>(p, s);
}
