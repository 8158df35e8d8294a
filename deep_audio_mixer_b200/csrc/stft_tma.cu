// K1, TMA-staged kernel (float32 arithmetic, float32 or PCM16 source, mono / planar stereo / interleaved stereo).
//
// Same transform as stft_db_kernel (stft_db.cu: two-stage 64 x 32 decomposition of the 2048-point frame,
// packed-complex codelets, dB epilogue in registers), with everything around the transform stripped down:
//
// * Every tile of 8 frames -- interior or at a row end -- arrives through the TMA unit: one thread hands the RAW
//   bytes of the next tile (float32 or int16, one contiguous run per channel plane, or the interleaved frames)
//   to cp.async.bulk while both transform stages of the current tile run.  The in-range part of a row-end tile
//   is copied and the reflected samples (center=True, pad_mode='reflect' of torch.stft, reference
//   data/dataset.py:145-149) are filled in by a short shared->shared pass.
// * PCM16 and stereo sources are converted in place, once per tile, by a two-phase shared->registers->shared
//   pass that reproduces the numpy prologue of the reference bit for bit (x / 32768 in float32 is exact; the
//   channel mean of two channels is (L + R) * 0.5): N1 of SURVEY section 8(f) -- the bytes of a WAV file go
//   straight from HBM to shared memory, 2 bytes per sample.
// * The tile walk needs no division: (row, tile-in-row) advance by a per-launch constant step.
// * The window table sits TRANSPOSED in shared memory (wt[b][a] = w[32 a + b], 68 floats per lane), so a stage-1
//   thread reads its 64 window values with 16 conflict-free 128-bit loads instead of 64 scalar ones.
// * The bin pitch of the output plane is a template parameter for the common shapes (88, 176, 216 frames), so
//   the 1025 stores of a frame use immediate offsets.
//
// Results are bit-identical to the phase kernel (tests/test_stft_gpu.py).
#include "stft_common.cuh"

// A/B switches (tools/ab_variants.py builds one library per setting; the defaults are what ships)
#ifndef DAM_TMA_SPLIT
#define DAM_TMA_SPLIT 1        // 1: split-phase mbarrier hand-over of the staging / exchange buffers, 0: __syncthreads
#endif
#ifndef DAM_TMA_BLOCKING_ISSUE
#define DAM_TMA_BLOCKING_ISSUE 0   // 1: thread 0 waits for all warps' loads before its own transform (A/B only)
#endif
#ifndef DAM_TMA_NYQ_LATE
#define DAM_TMA_NYQ_LATE 1     // 1: cross-lane DFT of Y_b[32] behind the stage-2 transform, 0: at the end of stage 1
#endif

namespace {

#ifndef DAM_TMA_FT
#define DAM_TMA_FT 8           // frames per tile = warps per CTA: 8 (two CTAs per SM); stft_tma16.cu compiles this file
#endif                         // again with 16 (one 16-warp CTA per SM, 64-byte store segments) as dam_k1_launch_tma16
constexpr int kFT = DAM_TMA_FT;             // frames per tile = warps per CTA
constexpr int kCtasPerSm = 16 / kFT;
constexpr int kThreads = 32 * kFT;
constexpr int kXch = XchStride<float, kFT>::v;
constexpr int kWtStride = 68;               // floats per lane of the transposed window table
constexpr int kRawSlack = 64;               // bytes behind a raw run for the 16-byte rounding of a skewed copy
// Samples staged in front of a tile's first frame.  When n_samples is a multiple of hop the last frame of a row is
// centred on x[n] and its reflected half reaches back to x[c - 1025], one sample before the frame itself: if that
// frame opens its tile the sample would not be staged.  Eight samples keep every copy 16-byte aligned.
constexpr int kPre = 8;

__host__ __device__ inline int span_max_of(int hop) { return (kFT - 1) * hop + kNfft + kPre; }
// bytes of one raw region: holds span_max float32 samples, or both int16 planes of a stereo tile
__host__ __device__ inline int region_bytes_of(int hop) { return span_max_of(hop) * 4 + 2 * kRawSlack; }

inline size_t tma_smem_bytes(int hop, int mode, bool tab_smem) {
    const size_t tables = tab_smem ? (size_t)32 * kWtStride * sizeof(float) + (size_t)kKp * kRow * sizeof(float2) + kNyqTab * sizeof(float) : 0;
    const size_t xch = (size_t)kFT * kXch * sizeof(float2);
    const int regions = (mode == TMA_F32_PLANAR2 || mode == TMA_F32_INTER2) ? 2 : 1;
    return tables + xch + (size_t)regions * region_bytes_of(hop) + 48;
}

// offset-binary trick: int16 v -> float32 v exactly, two samples per 32-bit word
//   bits(0x4B000000 | u16) = 2^23 + u16;  with u16 = v ^ 0x8000 (offset binary) the value is 2^23 + 32768 + v
//   With the exponent lowered by s the same two instructions yield v * 2^-s (the x / 32768 of PCM16, s = 15: the
//   float 256 + u16 / 32768 minus 257), so a power-of-two input scale costs nothing.
__device__ __forceinline__ void cvt2_i16(unsigned word, unsigned magic_bits, float magic_off, float& lo, float& hi) {
    const unsigned w = word ^ 0x80008000u;
    lo = __uint_as_float(__byte_perm(w, magic_bits, 0x7610)) - magic_off;
    hi = __uint_as_float(__byte_perm(w, magic_bits, 0x7632)) - magic_off;
}

// threadIdx.x through a volatile asm: values derived from it cannot be hoisted out of the tile loop (ptxas otherwise
// keeps the per-thread item indices / addresses of the conversion pass live across both transform stages and
// spills them around the codelets)
__device__ __forceinline__ int tid_now() {
    int t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    return t;
}

struct ConvArgs {
    float scale;
    bool unit;
    unsigned magic_bits;         // int16 -> float: exponent field of 2^(23 - s), s = 0 or the folded power-of-two scale
    float magic_off;             // 2^-s * (2^23 + 2^15)
    // PCM16 source: fold an input scale that is an exact power of two into the conversion itself
    __device__ __forceinline__ void init(double input_scale, bool is_i16) {
        scale = (float)input_scale;
        unit = input_scale == 1.0;
        int s = 0;
        if (is_i16 && input_scale > 0.0) {
            int e;
            if (frexp(input_scale, &e) == 0.5 && 1 - e >= 0 && 1 - e <= 100) { s = 1 - e; unit = true; }
        }
        magic_bits = (unsigned)(150 - s) << 23;
        magic_off = (float)ldexp(8421376.0, -s);
    }
    __device__ __forceinline__ float in(float v) const { return unit ? v : __fmul_rn(v, scale); }
    __device__ __forceinline__ float mean2(float l, float r) const { return (in(l) + in(r)) * 0.5f; }
};

// In-place conversion of one raw tile to float32 mono samples s[0, span): all loads, a barrier, all stores.
// One instantiation per source mode, so that each is a straight run of loads, one barrier and a straight run of
// stores (a single body switching on the mode per item made ptxas spill around BOTH transform codelets).
template <int kMode, int kMaxIt>
__device__ __forceinline__ void convert_tile_mode(const unsigned char* raw, const unsigned char* raw2, float* s,
                                                  int span, const ConvArgs& cv, int tid) {
    const int n_items = span >> 2;                       // 4 output samples per item
    constexpr bool kTwo = kMode == TMA_F32_PLANAR2 || kMode == TMA_F32_INTER2;
    uint4 a[kMaxIt], b[kTwo ? kMaxIt : 1];
    // loads are unconditional (items past the end re-read item 0): registers assigned under a predicate are
    // live from the top of the tile loop as far as ptxas is concerned
#pragma unroll
    for (int k = 0; k < kMaxIt; ++k) {
        const int it0 = tid + k * kThreads;
        const int it = it0 < n_items ? it0 : 0;
        if (kMode == TMA_I16_MONO) {
            const uint2 t = *reinterpret_cast<const uint2*>(raw + 8 * it);
            a[k] = make_uint4(t.x, t.y, 0, 0);
        } else if (kMode == TMA_I16_PLANAR2) {
            const uint2 t = *reinterpret_cast<const uint2*>(raw + 8 * it);
            const uint2 u = *reinterpret_cast<const uint2*>(raw2 + 8 * it);
            a[k] = make_uint4(t.x, t.y, u.x, u.y);
        } else if (kMode == TMA_I16_INTER2) {
            a[k] = *reinterpret_cast<const uint4*>(raw + 16 * it);
        } else if (kMode == TMA_F32_PLANAR2) {
            a[k] = *reinterpret_cast<const uint4*>(raw + 16 * it);
            b[kTwo ? k : 0] = *reinterpret_cast<const uint4*>(raw2 + 16 * it);
        } else {
            a[k] = *reinterpret_cast<const uint4*>(raw + 32 * it);
            b[kTwo ? k : 0] = *reinterpret_cast<const uint4*>(raw + 32 * it + 16);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kMaxIt; ++k) {
        const int it = tid + k * kThreads;
        if (it < n_items) {
            float4 o;
            if (kMode == TMA_I16_MONO) {
                float v0, v1, v2, v3;
                cvt2_i16(a[k].x, cv.magic_bits, cv.magic_off, v0, v1); cvt2_i16(a[k].y, cv.magic_bits, cv.magic_off, v2, v3);
                o = make_float4(cv.in(v0), cv.in(v1), cv.in(v2), cv.in(v3));
            } else if (kMode == TMA_I16_PLANAR2) {
                float l0, l1, l2, l3, r0, r1, r2, r3;
                cvt2_i16(a[k].x, cv.magic_bits, cv.magic_off, l0, l1); cvt2_i16(a[k].y, cv.magic_bits, cv.magic_off, l2, l3);
                cvt2_i16(a[k].z, cv.magic_bits, cv.magic_off, r0, r1); cvt2_i16(a[k].w, cv.magic_bits, cv.magic_off, r2, r3);
                o = make_float4(cv.mean2(l0, r0), cv.mean2(l1, r1), cv.mean2(l2, r2), cv.mean2(l3, r3));
            } else if (kMode == TMA_I16_INTER2) {
                float l0, r0, l1, r1, l2, r2, l3, r3;
                cvt2_i16(a[k].x, cv.magic_bits, cv.magic_off, l0, r0); cvt2_i16(a[k].y, cv.magic_bits, cv.magic_off, l1, r1);
                cvt2_i16(a[k].z, cv.magic_bits, cv.magic_off, l2, r2); cvt2_i16(a[k].w, cv.magic_bits, cv.magic_off, l3, r3);
                o = make_float4(cv.mean2(l0, r0), cv.mean2(l1, r1), cv.mean2(l2, r2), cv.mean2(l3, r3));
            } else if (kMode == TMA_F32_PLANAR2) {
                const uint4 c = b[kTwo ? k : 0];
                o = make_float4(cv.mean2(__uint_as_float(a[k].x), __uint_as_float(c.x)), cv.mean2(__uint_as_float(a[k].y), __uint_as_float(c.y)),
                                cv.mean2(__uint_as_float(a[k].z), __uint_as_float(c.z)), cv.mean2(__uint_as_float(a[k].w), __uint_as_float(c.w)));
            } else {
                const uint4 c = b[kTwo ? k : 0];
                o = make_float4(cv.mean2(__uint_as_float(a[k].x), __uint_as_float(a[k].y)), cv.mean2(__uint_as_float(a[k].z), __uint_as_float(a[k].w)),
                                cv.mean2(__uint_as_float(c.x), __uint_as_float(c.y)), cv.mean2(__uint_as_float(c.z), __uint_as_float(c.w)));
            }
            *reinterpret_cast<float4*>(s + 4 * it) = o;
        }
    }
}

template <int kMaxIt>
__device__ __forceinline__ void convert_tile(int mode, const unsigned char* raw, const unsigned char* raw2, float* s,
                                             int span, const ConvArgs& cv, int tid) {
    switch (mode) {
        case TMA_I16_MONO:    convert_tile_mode<TMA_I16_MONO, kMaxIt>(raw, raw2, s, span, cv, tid); break;
        case TMA_I16_PLANAR2: convert_tile_mode<TMA_I16_PLANAR2, kMaxIt>(raw, raw2, s, span, cv, tid); break;
        case TMA_I16_INTER2:  convert_tile_mode<TMA_I16_INTER2, kMaxIt>(raw, raw2, s, span, cv, tid); break;
        case TMA_F32_PLANAR2: convert_tile_mode<TMA_F32_PLANAR2, kMaxIt>(raw, raw2, s, span, cv, tid); break;
        default:              convert_tile_mode<TMA_F32_INTER2, kMaxIt>(raw, raw2, s, span, cv, tid); break;
    }
}


// ---- stage 2 of the TMA kernel, in two parts so that the dB epilogue of one tile can be interleaved with the
// stage-1 loads of the next (both are light on the FMA pipe and heavy on the LSU / MUFU / store path) ----------
template <bool kStats, bool kGlobalTw>
__device__ __forceinline__ void stage2_transform(pk2 (&z)[32], const float2* src, const float2* tw, unsigned long long* xch_free) {
    const ulonglong2* s2 = reinterpret_cast<const ulonglong2*>(src);     // two complex values per 128-bit load
    const float4* w4 = reinterpret_cast<const float4*>(tw);
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
        ulonglong2 v[4];
        float4 w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { v[u] = s2[grp * 4 + u]; w[u] = kGlobalTw ? __ldg(w4 + grp * 4 + u) : w4[grp * 4 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int b2 = grp * 4 + u;
            z[2 * b2] = pk_fma(pk_swap(v[u].x), pk_make(-w[u].y, w[u].y), pk_mul(v[u].x, pk_c(w[u].x)));
            z[2 * b2 + 1] = pk_fma(pk_swap(v[u].y), pk_make(-w[u].w, w[u].w), pk_mul(v[u].y, pk_c(w[u].z)));
        }
        asm volatile("" ::: "memory");
    }
    if (xch_free != nullptr) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(xch_free);  // this warp's exchange rows may be overwritten
    }
    dam_cfft32_pc(z);
}

// dB + store of bin slot q of a task (q < 16: bin k' + 64 q, q >= 16: mirror bin 2048 - k' - 64 q)
template <bool kStats, int kPitch>
struct Epilogue {
    float* o_lo;
    float* o_hi;
    int sb;
    bool lo_ok, hi_ok;
    __device__ __forceinline__ void init(float* out_frame, long long T, int kp, bool active) {
        const long long pitch = kPitch > 0 ? (long long)kPitch : T;
        o_lo = out_frame + kp * pitch;
        o_hi = out_frame + (kNfft / 2 - kp) * pitch;
        sb = (int)(64 * pitch * (long long)sizeof(float));
        lo_ok = active;
        hi_ok = active && kp != 0;
    }
    __device__ __forceinline__ void put(int q, pk2 zq, const StftParams& p, StatAcc& st) const {
        float re, im;
        pk_split(zq, re, im);
        const float db = DbOut<float>::eval(re, im, p);
        const bool valid = q <= 16 ? lo_ok : hi_ok;
        float* o;
        if (kPitch > 0) o = q < 16 ? o_lo + q * (64 * kPitch) : o_hi - (q - 16) * (64 * kPitch);
        else o = reinterpret_cast<float*>(q < 16 ? reinterpret_cast<char*>(o_lo) + (long long)q * (long long)sb
                                                 : reinterpret_cast<char*>(o_hi) + (long long)(16 - q) * (long long)sb);
        if (valid) *o = db;
        if (kStats && valid) st.add(db);
    }
};

// stage-1 loads of one frame: xp[m] = (x[64m + b] w[64m + b], x[64m + 32 + b] w[64m + 32 + b]), group j = xp[2j], xp[2j+1]
template <bool kTabSmem>
__device__ __forceinline__ void load_group(pk2 (&xp)[32], int j, const float* f, const float* s_wt_lane, const float* gw) {
    if (kTabSmem) {
        const ulonglong2 w = reinterpret_cast<const ulonglong2*>(s_wt_lane)[j];
        xp[2 * j] = pk_mul(pk_make(f[128 * j], f[128 * j + 32]), w.x);
        xp[2 * j + 1] = pk_mul(pk_make(f[128 * j + 64], f[128 * j + 96]), w.y);
    } else {
        xp[2 * j] = pk_mul(pk_make(f[128 * j], f[128 * j + 32]), pk_make(__ldg(gw + 128 * j), __ldg(gw + 128 * j + 32)));
        xp[2 * j + 1] = pk_mul(pk_make(f[128 * j + 64], f[128 * j + 96]), pk_make(__ldg(gw + 128 * j + 64), __ldg(gw + 128 * j + 96)));
    }
}

// geometry of one tile
struct Tile {
    int nf, span, lo, hi;        // frames, samples staged, in-range part [lo, hi) of the staged span
    long long g0;                // first staged sample (may be negative)
};
__device__ __forceinline__ Tile tile_of(unsigned tin, int hop, int T, long long n) {
    Tile t;
    const int f0 = (int)tin * kFT;
    t.nf = (T - f0) < kFT ? (T - f0) : kFT;
    t.g0 = (long long)f0 * hop - kNfft / 2 - kPre;
    t.span = (t.nf - 1) * hop + kNfft + kPre;
    t.lo = t.g0 < 0 ? (int)(-t.g0) : 0;
    const long long rem = n - t.g0;
    t.hi = rem < (long long)t.span ? (int)rem : t.span;
    return t;
}

template <bool kStats, bool kTabSmem, int kPitch, int kMaxIt>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
stft_db_tma_kernel(const StftParams p) {
    using OUT = float;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_wt = reinterpret_cast<float*>(smem_raw);                                     // [32][68]   (kTabSmem)
    float2* s_tw = reinterpret_cast<float2*>(s_wt + (kTabSmem ? 32 * kWtStride : 0));     // [33][34]   (kTabSmem)
    float* s_nq = reinterpret_cast<float*>(s_tw + (kTabSmem ? kKp * kRow : 0));           // [8][32]    (kTabSmem)
    float2* s_xch = reinterpret_cast<float2*>(s_nq + (kTabSmem ? kNyqTab : 0));           // [kFT][kXch]
    unsigned char* s_region = reinterpret_cast<unsigned char*>(s_xch + kFT * kXch);       // raw tile / float32 samples
    const int hop = p.hop;
    const int mode = p.tma_mode;
    const int region_bytes = region_bytes_of(hop);
    const int regions = (mode == TMA_F32_PLANAR2 || mode == TMA_F32_INTER2) ? 2 : 1;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(s_region + (size_t)regions * region_bytes);   // tile landed
    unsigned long long* bar_stage = bar + 1;     // 8 arrivals: every warp holds its frame in registers
    unsigned long long* bar_xch = bar + 2;       // 8 arrivals: every warp has consumed its exchange rows
    volatile int* s_skew = reinterpret_cast<volatile int*>(bar + 3);                      // [2], by tile parity

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (kTabSmem) {
        for (int i = tid; i < kNfft; i += kThreads) s_wt[(i & 31) * kWtStride + (i >> 5)] = __ldg(p.window + i);
        const float2* tw = reinterpret_cast<const float2*>(p.twiddle);
        for (int i = tid; i < kKp * kRow; i += kThreads) s_tw[i] = tw[i];
        nyq_table_init<float>(s_nq, tid);
    }
    const float* nq = kTabSmem ? s_nq : reinterpret_cast<const float*>(p.twiddle) + 2 * kKp * kRow;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar_stage, kFT);
        mbar_init(bar_xch, kFT);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const int T = (int)p.T;
    const long long n = p.n_samples;
    const unsigned tpr = (unsigned)p.tiles_per_row, n_rows = (unsigned)p.n_rows;
    const float2* twp = kTabSmem ? s_tw : reinterpret_cast<const float2*>(p.twiddle);
    const int eb = mode >= TMA_I16_MONO ? 2 : 4;                                           // bytes per source element
    const int frame_bytes = (mode == TMA_F32_INTER2 || mode == TMA_I16_INTER2) ? 2 * eb : eb;
    // second plane of a planar stereo tile: behind the first int16 plane inside the region, or the second region
    const int plane2_off = mode == TMA_I16_PLANAR2 ? span_max_of(hop) * 2 + kRawSlack : region_bytes;
    ConvArgs cv;
    cv.init(p.input_scale, mode >= TMA_I16_MONO);
    StatAcc st;

    // Thread 0 hands the in-range bytes of tile (r, tin) to the TMA unit.  A source run that is not 16-byte
    // aligned is copied from the aligned address below it; the skew (bytes) is left in s_skew[parity].
    auto issue = [&](unsigned r, unsigned tin, unsigned parity) {
        const Tile t = tile_of(tin, hop, T, n);
        const long long elem = row_offset(p, (long long)r) + (t.g0 + t.lo) * (frame_bytes / eb);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.audio) + elem * eb;
        const unsigned skew = (unsigned)(reinterpret_cast<unsigned long long>(src) & 15ull);
        const unsigned bytes = ((unsigned)(t.hi - t.lo) * (unsigned)frame_bytes + skew + 15u) & ~15u;
        unsigned char* dst = s_region + (size_t)t.lo * frame_bytes;
        s_skew[parity] = (int)skew;
        const bool planar2 = mode == TMA_F32_PLANAR2 || mode == TMA_I16_PLANAR2;
        mbar_expect_tx(bar, planar2 ? 2u * bytes : bytes);
        bulk_g2s(dst, src - skew, bytes, bar);
        if (planar2) bulk_g2s(dst + plane2_off, src - skew + p.chan_stride * eb, bytes, bar);
    };

    // Prepares the staged tile (conversion to float32 mono, reflection at a row end) and returns the pointer to its
    // samples, s_x[i] = x[g0 + i].  Called by all threads after the tile has landed; contains CTA barriers.
    auto prepare = [&](const Tile& t, unsigned par) -> const float* {
        const int skew = s_skew[par];
        float* s_x = reinterpret_cast<float*>(s_region);
        if (mode == TMA_F32_MONO) {
            s_x = reinterpret_cast<float*>(s_region + skew);
        } else {
            // six items per thread cover a tile up to hop 512 (the unrolled, guarded iterations beyond cost issue slots)
            if (t.span <= 6 * 4 * kThreads) convert_tile<6>(mode, s_region + skew, s_region + skew + plane2_off, s_x, t.span, cv, tid_now());
            else convert_tile<kMaxIt>(mode, s_region + skew, s_region + skew + plane2_off, s_x, t.span, cv, tid_now());
            fence_proxy_async();                                  // generic writes before the next TMA write to the region
            __syncthreads();
        }
        if (t.lo > 0 || t.hi < t.span) {
            // row end: reflected samples from the in-range part (x[-k] = x[k], x[n-1+k] = x[n-1-k])
            const int tv = tid_now();
            for (int i = tv; i < t.lo; i += kThreads) s_x[i] = s_x[2 * t.lo - i];
            for (int i = t.hi + tv; i < t.span; i += kThreads) s_x[i] = s_x[2 * t.hi - 2 - i];
            fence_proxy_async();
            __syncthreads();
        }
        return s_x;
    };
    const float* wt_lane = s_wt + lane * kWtStride;
    const float* gw_lane = p.window + lane;

    // tile walk: (row, tin) advance by (step_rows, step_tiles) = gridDim.x in mixed radix
    unsigned row = blockIdx.x / tpr, tin = blockIdx.x - row * tpr;
    unsigned parity = 0;
    if (tid == 0 && row < n_rows) issue(row, tin, 0);

    while (row < n_rows) {
        unsigned nrow = row + p.step_rows, ntin = tin + p.step_tiles;
        if (ntin >= tpr) { ntin -= tpr; ++nrow; }
        const Tile t = tile_of(tin, hop, T, n);
        const int nf = t.nf;
        OUT* out_row = reinterpret_cast<OUT*>(p.out) + (long long)row * p.out_row_stride + (int)tin * kFT;

        mbar_wait(bar, parity);                                   // the raw bytes of this tile have landed
        const float* s_x = prepare(t, parity);

        // ---- stage 1 loads: warp = frame, lane = b
        pk2 xp[32];
        {
            const float* f = s_x + kPre + (warp < nf ? warp : nf - 1) * hop + lane;
#pragma unroll
            for (int j = 0; j < 16; ++j) load_group<kTabSmem>(xp, j, f, wt_lane, gw_lane);
        }

        // ---- split-phase hand-over of the two shared buffers (no CTA-wide barrier here: the warps drift apart, so
        // the shared-memory bursts of one overlap the transforms of another)
        // (a) staged samples: a warp arrives on bar_stage once its frame is in registers; thread 0 waits for all
        //     eight and only then lets the TMA unit overwrite the region with the next tile
#if DAM_TMA_SPLIT
        // (a) staged samples: a warp arrives on bar_stage once its frame is in registers.  Thread 0 hands the region
        //     to the TMA unit as soon as all eight have arrived, but never WAITS for them (its warp would start its
        //     transform last and keep the whole CTA at the stage-1 -> stage-2 barrier): it probes now, probes again
        //     after its own transform, and issues behind that barrier at the latest -- every warp has passed its
        //     loads by then.
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_stage);
        bool issued = nrow >= n_rows;
#if DAM_TMA_BLOCKING_ISSUE
        if (tid == 0 && !issued) mbar_wait(bar_stage, parity);
#endif
        if (tid == 0 && !issued && mbar_test(bar_stage, parity)) {
            issue(nrow, ntin, parity ^ 1u);
            issued = true;
        }
#else
        __syncthreads();
        if (tid == 0 && nrow < n_rows) issue(nrow, ntin, parity ^ 1u);
#endif

        // ---- stage 1: 64-point real DFT of the lane's samples in registers; the exchange rows are written once
        // (b) every warp has consumed the rows of the previous tile (arrivals on bar_xch in stage 2; by now every
        //     warp is past them, so waiting before the transform costs nothing and lets the exchange stores
        //     interleave with the last butterflies)
        float y32 = 0.f;
        if (warp < nf) {
#if DAM_TMA_SPLIT
            mbar_wait(bar_xch, parity ^ 1u);
#endif
            pk2 y0n, s[32];
            dam_rfft64_pc(xp, y0n, s);
            float y0;
            pk_split(y0n, y0, y32);
            float2* dst = s_xch + warp * kXch + lane;
            dst[0] = make_float2(y0, 0.f);                       // row 0: Y[0] (real)
#pragma unroll
            for (int k = 1; k < 32; ++k) *reinterpret_cast<pk2*>(dst + k * kRow) = s[k];
#if DAM_TMA_SPLIT
            if (tid == 0 && !issued && mbar_test(bar_stage, parity)) {
                issue(nrow, ntin, parity ^ 1u);
                issued = true;
            }
#endif
#if !DAM_TMA_NYQ_LATE
            nyq_bins<float, OUT, kStats>(y32, out_row + warp, p.bin_stride, nq, p, lane, st);
#endif
        }
        __syncthreads();         // stage 1 -> stage 2: all rows of the tile are in the exchange buffer
#if DAM_TMA_SPLIT
        if (tid == 0 && !issued) issue(nrow, ntin, parity ^ 1u);
#endif

        // ---- stage 2: thread = (k' = tid / FT in 0..31, frame = tid % FT); lanes past the end of a ragged tile
        // transform a duplicate of its last frame and store nothing (no divergence around the arrive)
        {
            const int kp = tid / kFT, fr = tid % kFT;
            pk2 z[32];
            stage2_transform<kStats, !kTabSmem>(z, s_xch + (fr < nf ? fr : nf - 1) * kXch + kp * kRow, twp + kp * kRow,
                                                DAM_TMA_SPLIT ? bar_xch : nullptr);
            // bins 32 + 64 q of the frame this WARP transformed in stage 1 (Y_b[32] kept in a register): the shuffle
            // chain is issued here so that its latency hides behind the epilogue
#if DAM_TMA_NYQ_LATE
            nyq_bins<float, OUT, kStats>(y32, out_row + warp, p.bin_stride, nq, p, lane, st, warp < nf);
#endif
            Epilogue<kStats, kPitch> ep;
            ep.init(out_row + fr, p.bin_stride, kp, fr < nf);
#pragma unroll
            for (int q = 0; q < 32; ++q) ep.put(q, z[q], p, st);
            if (kStats) st.fold();
        }
        row = nrow; tin = ntin; parity ^= 1u;
    }
    if (kStats) st.flush(p.stats, lane);
}

template <bool kStats, bool kTabSmem, int kPitch, int kMaxIt>
int launch_one(const StftParams& p, size_t smem, unsigned grid, cudaStream_t stream) {
    auto kern = stft_db_tma_kernel<kStats, kTabSmem, kPitch, kMaxIt>;
    static std::mutex mu;
    static bool configured[64] = {};
    int dev = 0;
    DAM_CHECK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[dev]) {
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            configured[dev] = true;
        }
    }
    { kern<<<grid, kThreads, smem, stream>>>(p); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

template <bool kTabSmem, int kMaxIt>
int launch_pitch(const StftParams& p, size_t smem, unsigned grid, cudaStream_t stream) {
    if (p.stats != nullptr) return launch_one<true, kTabSmem, 0, kMaxIt>(p, smem, grid, stream);
    switch ((int)p.bin_stride) {
        case 88:  return launch_one<false, kTabSmem, 88, kMaxIt>(p, smem, grid, stream);
        case 176: return launch_one<false, kTabSmem, 176, kMaxIt>(p, smem, grid, stream);
        case 216: return launch_one<false, kTabSmem, 216, kMaxIt>(p, smem, grid, stream);
        default:  return launch_one<false, kTabSmem, 0, kMaxIt>(p, smem, grid, stream);
    }
}

}  // namespace

#if DAM_TMA_FT == 8
const char* dam_k1_last_kernel_name = "";
#define DAM_TMA_ENTRY dam_k1_launch_tma
#else
#define DAM_TMA_ENTRY dam_k1_launch_tma16
#endif

int DAM_TMA_ENTRY(StftParams p, cudaStream_t stream, int force_tab) {
    // ---- does the shape qualify?  (anything else runs the phase kernel, same results)
    const int hop = p.hop;
    if (hop % 8 != 0 || hop < 64 || hop > 1024) return 1;
    if (p.row_gain != nullptr) return 1;
    if (p.n_samples < 2 * kNfft || p.T >= (1 << 24)) return 1;
    const bool f32 = p.audio_dtype == DAM_F32, i16 = p.audio_dtype == DAM_I16;
    if (!f32 && !i16) return 1;
    const int eb = f32 ? 4 : 2;
    int mode;
    if (p.channels == 1 && p.samp_stride == 1) mode = f32 ? TMA_F32_MONO : TMA_I16_MONO;
    else if (p.channels == 2 && p.samp_stride == 1) mode = f32 ? TMA_F32_PLANAR2 : TMA_I16_PLANAR2;
    else if (p.channels == 2 && p.samp_stride == 2 && p.chan_stride == 1) mode = f32 ? TMA_F32_INTER2 : TMA_I16_INTER2;
    else return 1;
    if (mode == TMA_F32_MONO && p.input_scale != 1.0) return 1;      // no conversion pass to apply the scale in
    // alignment of every row start: the conversion pass reads 8-byte (PCM16 mono / planar) or 16-byte vectors from
    // (16-byte aligned address + skew), float32 mono rows are read element-wise and only need their natural 4 bytes
    const long long align = mode == TMA_F32_MONO ? 4 : ((mode == TMA_I16_MONO || mode == TMA_I16_PLANAR2) ? 8 : 16);
    const unsigned long long base = reinterpret_cast<unsigned long long>(p.audio);
    if (base % align != 0 || (p.row_stride * eb) % align != 0) return 1;
    if (p.n_inner > 0 && (p.outer_stride * eb) % align != 0) return 1;
    if ((mode == TMA_F32_PLANAR2 || mode == TMA_I16_PLANAR2) && (p.chan_stride * eb) % 16 != 0) return 1;
    // the in-range part of a row-end tile must end on a conversion vector (4 frames)
    if (mode != TMA_F32_MONO && p.n_samples % 4 != 0) return 1;

    p.tma_mode = mode;
    p.tiles_per_row = (p.T + kFT - 1) / kFT;
    p.n_tiles = p.tiles_per_row * p.n_rows;
    DAM_CHECK_ARG(p.n_tiles < 0x7fffffffLL && p.n_rows < 0x7fffffffLL, "too many frame tiles for one launch (%lld); split the batch",
                  (long long)p.n_tiles);
    p.l2_prefetch = 0;
    // Output rows that are not a multiple of 8 frames (32 bytes) make every 8-frame store straddle two sectors; the
    // variant with its tables in global memory leaves the SM twice the L1 to merge them in and measured 10-15 %
    // faster there (T = 173 dense: 0.87 vs 1.00 ms, hop 256 T = 345: 0.89 vs 0.98 ms), and slower everywhere else.
    if (force_tab < 0 && p.bin_stride % 8 != 0) force_tab = 0;
    for (int tab = 1; tab >= 0; --tab) {
        if (force_tab >= 0 && tab != force_tab) continue;
        const size_t smem = tma_smem_bytes(hop, mode, tab != 0);
        if (kCtasPerSm * (smem + 1024) > 228 * 1024) continue;       // two CTAs per SM (one of 16 warps) or not at all
        long long grid = (long long)dam_sm_count() * kCtasPerSm;
        if (grid > p.n_tiles) grid = p.n_tiles;
        p.step_rows = (unsigned)(grid / p.tiles_per_row);
        p.step_tiles = (unsigned)(grid % p.tiles_per_row);
        static const bool trace = getenv("DAM_K1_TRACE") != nullptr;
        if (trace) fprintf(stderr, "[dam] stft_db_tma_kernel mode=%d tables=%s smem=%zu hop=%d pitch=%lld T=%lld\n", mode,
                           tab ? "smem" : "global", smem, hop, (long long)p.bin_stride, (long long)p.T);
        dam_k1_last_kernel_name = kFT == 8 ? (tab ? "stft_db_tma_kernel<tables in shared memory>" : "stft_db_tma_kernel<tables in global memory>")
                                           : (tab ? "stft_db_tma_kernel<16 frames per tile, tables in shared memory>"
                                                  : "stft_db_tma_kernel<16 frames per tile, tables in global memory>");
        // conversion items per thread: span_max / 4 / 256 rounded up = 10 at hop 1024, the largest hop accepted
        static_assert(((kFT - 1) * 1024 + kNfft + kPre + 3) / 4 <= 10 * kThreads, "kMaxIt");
        return tab ? launch_pitch<true, 10>(p, smem, (unsigned)grid, stream) : launch_pitch<false, 10>(p, smem, (unsigned)grid, stream);
    }
    return 1;
}
