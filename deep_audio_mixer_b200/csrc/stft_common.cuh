// Device code shared by the K1 translation units (stft_db.cu: phase / generic kernels and the C entry point;
// stft_tma.cu: the TMA-staged kernel).  Everything lives in an anonymous namespace of the including file.
#pragma once
#include "dam_common.cuh"
#include "fft_codelets.cuh"

#include <math.h>
#include <stdlib.h>
#include <cstdio>
#include <mutex>
#include <vector>

struct StftParams {
    const void* audio;
    int audio_dtype;
    int channels;
    long long n_rows, n_samples, row_stride, chan_stride, samp_stride;
    long long n_inner, outer_stride;   // row r starts at (r / n_inner)*outer_stride + (r % n_inner)*row_stride
    double input_scale;
    const double* row_gain;
    const float* window;
    const void* twiddle;          // device table, Cx<R>::type [33][34]
    int hop;
    long long T;                  // frames per row
    long long tiles_per_row, n_tiles;
    int l2_prefetch;              // 1: prefetch the CTA's next tile into L2 while transforming this one
    float log_scale;              // multiplier/2 * log10(2): dB = log_scale * log2(|X|^2)
    float floor_db;               // multiplier * log10(amin), the clamp applied in the dB domain
    double log_scale_d, floor_db_d;
    void* out;
    long long out_row_stride;
    long long bin_stride;         // elements between consecutive bins of a plane (>= T; T when dense)
    double* stats;
    // ---- TMA-staged kernel only (stft_tma.cu) ----
    int tma_mode;                 // TmaMode: source dtype x channel layout
    unsigned step_rows, step_tiles;   // gridDim.x = step_rows * tiles_per_row + step_tiles (tile walk without divisions)
};

// source dtype x channel layout of the TMA-staged kernel
enum TmaMode { TMA_F32_MONO = 0, TMA_F32_PLANAR2 = 1, TMA_F32_INTER2 = 2, TMA_I16_MONO = 3, TMA_I16_PLANAR2 = 4, TMA_I16_INTER2 = 5 };

// stft_tma.cu: launches the TMA-staged kernel if the shape qualifies.  Returns DAM_OK (launched), 1 (shape does
// not qualify: the caller uses the phase kernel) or a negative error code.  p.twiddle must be the float table.
int dam_k1_launch_tma(StftParams p, cudaStream_t stream, int force_tab);
int dam_k1_launch_tma16(StftParams p, cudaStream_t stream, int force_tab);     // 16 frames per tile (stft_tma16.cu)
// name of the K1 kernel the last dam_stft_db call of this process dispatched (dam_last_stft_kernel())
extern const char* dam_k1_last_kernel_name;


namespace {

constexpr int kNfft = 2048;
constexpr int kBins = kNfft / 2 + 1;
// Frames per tile FT is a template parameter: a CTA has FT warps (stage 1: one warp per frame;
// stage 2: 32 k' x FT frames = 32*FT thread tasks).  float32 uses FT = 8 with two CTAs per SM,
// float64 FT = 8 with one; FT = 16 (one 16-warp CTA per SM) is used for unaligned long rows at hop >= 1024.
constexpr int kRow = 34;          // complex elements per k' row in the exchange / twiddle tables
constexpr int kKp = 33;           // stage-2 tasks per frame

template <typename R> struct Cx;
template <> struct Cx<float>  { using type = float2; };
template <> struct Cx<double> { using type = double2; };
// Exchange buffer: per frame 32 rows (k' = 0..31) of kRow complex values (k' = 32 never leaves
// stage 1).  The per-frame stride, in complex elements, is chosen so the 128-bit stage-2 loads
// (lane = frame + FT * (k' mod 32/FT)) are bank-conflict free:
//   float  (two complex per load, row = 17 x 16 B): stride/2 odd
//   double (one complex per load):                  FT = 8 needs stride odd
template <typename R, int FT> struct XchStride;
template <> struct XchStride<float, 8>  { static constexpr int v = 32 * kRow + 2; };   // 1090
template <> struct XchStride<float, 16> { static constexpr int v = 32 * kRow + 2; };   // 1090
template <> struct XchStride<double, 8> { static constexpr int v = 32 * kRow + 1; };   // 1089


__device__ __forceinline__ long long row_offset(const StftParams& p, long long row) {
    if (p.n_inner <= 0) return row * p.row_stride;
    const long long outer = row / p.n_inner;
    return outer * p.outer_stride + (row - outer * p.n_inner) * p.row_stride;
}

// ---- staging --------------------------------------------------------------------------------
// Samples [g0, g0 + span) of one row go to shared memory as the compute type R after the
// prologue the reference performs in numpy, in the SOURCE precision P (float for int16/float32
// sources, double for float64): x*input_scale, channel mean (sum in channel order, then a
// division by C -- a multiply by 1/C when C is a power of two, which is bit-identical), row gain.
// Loads are issued in batches so every thread keeps several global requests in flight; interior,
// unit-stride, 16-byte-aligned tiles use 128-bit loads and vector shared-memory stores.
template <typename SRC, int VB> struct VecOf { static constexpr int kN = VB / (int)sizeof(SRC); };

// VB-byte (16 or 8) read-only vector load; an 8-byte load fills the low half
template <int VB>
__device__ __forceinline__ uint4 ldg_vec(const void* ptr, int v) {
    if (VB == 16) return __ldg(reinterpret_cast<const uint4*>(ptr) + v);
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(ptr) + v);
    return make_uint4(t.x, t.y, 0u, 0u);
}

template <typename P, typename SRC, int N>
__device__ __forceinline__ void unpack_vec(const uint4& raw, P (&v)[N]) {
    const SRC* e = reinterpret_cast<const SRC*>(&raw);
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = (P)e[j];
}

template <typename R, int V>
__device__ __forceinline__ void store_vec(R* dst, const R (&v)[V]) {
    constexpr int kBytes = V * (int)sizeof(R);
    if (kBytes == 4) {
        *reinterpret_cast<unsigned*>(dst) = *reinterpret_cast<const unsigned*>(v);
    } else if (kBytes == 8) {
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(v);
    } else {
#pragma unroll
        for (int k = 0; k < kBytes / 16; ++k)
            reinterpret_cast<uint4*>(dst)[k] = reinterpret_cast<const uint4*>(v)[k];
    }
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

template <typename P>
struct Prologue {
    P scale, gain, inv_c, f_c;
    int C;
    bool unit_scale, has_gain;
    // x * input_scale is rounded on its own (numpy: `x / 32768` is an array before the channels are added);
    // the intrinsic keeps ptxas from contracting it into the channel sum
    __device__ __forceinline__ P load(P raw) const { return unit_scale ? raw : mul_rn(raw, scale); }
    __device__ __forceinline__ P finish(P acc) const {
        if (C > 1) acc = (inv_c != (P)0) ? acc * inv_c : acc / f_c;
        return has_gain ? gain * acc : acc;
    }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Interior tile of unit-stride channels (mono or planar), every channel VB-byte aligned.  Returns the
// number of samples staged (a multiple of the vector length; the caller finishes the tail).
template <typename R, typename P, typename SRC, int kThreads, int VB>
__device__ __forceinline__ int stage_planar(const StftParams& p, const Prologue<P>& pro, R* __restrict__ s_stage,
                                            const SRC* __restrict__ src, int span, int tid) {
    constexpr int V = VecOf<SRC, VB>::kN;
    constexpr int U = 4;                         // vector loads in flight per thread and channel
    const int C = p.channels;
    const int nvec = span / V;
    if (VB == 16 && sizeof(R) == sizeof(SRC) && sizeof(P) == sizeof(SRC) && C == 1 && pro.unit_scale && !pro.has_gain) {
        // plain copy (same type, mono, no scaling): asynchronous 16-byte global->shared copies
        for (int v = tid; v < nvec; v += kThreads) cp_async16(s_stage + v * V, src + v * V);
        cp_async_wait_all();
    } else if (C == 2 && pro.unit_scale && !pro.has_gain) {
        // stereo source, no scaling (the inference path with the mono shim): both channels of
        // six vectors in flight per thread, (L + R) * 0.5 == numpy's float mean over two channels
        constexpr int U2 = 6;
        const SRC* s1 = src + p.chan_stride;
        for (int v0 = tid; v0 < nvec; v0 += kThreads * U2) {
            uint4 r0[U2], r1[U2];
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int v = v0 + u * kThreads;
                r0[u] = v < nvec ? ldg_vec<VB>(src, v) : make_uint4(0, 0, 0, 0);
                r1[u] = v < nvec ? ldg_vec<VB>(s1, v) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U2; ++u) {
                const int v = v0 + u * kThreads;
                if (v < nvec) {
                    P a[V], c[V];
                    R r[V];
                    unpack_vec<P, SRC, V>(r0[u], a);
                    unpack_vec<P, SRC, V>(r1[u], c);
#pragma unroll
                    for (int j = 0; j < V; ++j) r[j] = (R)((a[j] + c[j]) * (P)0.5);
                    store_vec<R, V>(s_stage + v * V, r);
                }
            }
        }
    } else {
        for (int v0 = tid; v0 < nvec; v0 += kThreads * U) {
            P acc[U][V];
            uint4 raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int v = v0 + u * kThreads;
                raw[u] = v < nvec ? ldg_vec<VB>(src, v) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                unpack_vec<P, SRC, V>(raw[u], acc[u]);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[u][j] = pro.load(acc[u][j]);
            }
            for (int c = 1; c < C; ++c) {
                const SRC* sc = src + c * p.chan_stride;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int v = v0 + u * kThreads;
                    raw[u] = v < nvec ? ldg_vec<VB>(sc, v) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    P t[V];
                    unpack_vec<P, SRC, V>(raw[u], t);
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[u][j] += pro.load(t[j]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int v = v0 + u * kThreads;
                if (v < nvec) {
                    R r[V];
#pragma unroll
                    for (int j = 0; j < V; ++j) r[j] = (R)pro.finish(acc[u][j]);
                    store_vec<R, V>(s_stage + v * V, r);
                }
            }
        }
    }
    return nvec * V;
}

// Interior tile of an INTERLEAVED stereo source (L R L R ..., the frame order of a WAV file and of
// soundfile.read): a VB-byte vector holds VB / (2 sizeof(SRC)) frames; L + R in the source precision, then
// the same mean / gain as the planar path.
template <typename R, typename P, typename SRC, int kThreads, int VB>
__device__ __forceinline__ int stage_interleaved2(const Prologue<P>& pro, R* __restrict__ s_stage,
                                                  const SRC* __restrict__ src, int span, int tid) {
    constexpr int VE = VecOf<SRC, VB>::kN;       // elements per vector
    constexpr int F = VE / 2;                    // frames per vector
    constexpr int U = 6;
    const int nvec = span / F;
    for (int v0 = tid; v0 < nvec; v0 += kThreads * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = v0 + u * kThreads;
            raw[u] = v < nvec ? ldg_vec<VB>(src, v) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int v = v0 + u * kThreads;
            if (v < nvec) {
                P e[VE];
                R r[F];
                unpack_vec<P, SRC, VE>(raw[u], e);
#pragma unroll
                for (int j = 0; j < F; ++j) r[j] = (R)pro.finish(pro.load(e[2 * j]) + pro.load(e[2 * j + 1]));
                store_vec<R, F>(s_stage + v * F, r);
            }
        }
    }
    return nvec * F;
}

template <typename R, typename P, typename SRC, int kThreads>
__device__ __forceinline__ void stage_tile(const StftParams& p, R* __restrict__ s_stage, long long row,
                                           long long g0, int span, int tid) {
    const SRC* __restrict__ base = reinterpret_cast<const SRC*>(p.audio) + row_offset(p, row);
    const long long n = p.n_samples;
    const int C = p.channels;
    Prologue<P> pro;
    pro.C = C;
    pro.scale = (P)p.input_scale;
    pro.has_gain = p.row_gain != nullptr;
    pro.gain = pro.has_gain ? (P)p.row_gain[row] : (P)1;
    pro.unit_scale = p.input_scale == 1.0;
    pro.f_c = (P)C;
    pro.inv_c = ((C & (C - 1)) == 0) ? (P)1 / (P)C : (P)0;

    int done = 0;
    const bool interior = g0 >= 0 && g0 + span <= n;
    if (interior && p.samp_stride == 1) {
        // unit-stride channels: 16-byte vectors when every channel start is 16-byte aligned, else 8-byte
        // vectors (PCM16 rows of an odd number of 8-byte groups, e.g. 220500 samples, alternate)
        const SRC* src = base + g0;
        const unsigned long long bits = reinterpret_cast<unsigned long long>(src) |
                                        (C == 1 ? 0ull : (unsigned long long)(p.chan_stride * (long long)sizeof(SRC)));
        if ((bits & 15ull) == 0) done = stage_planar<R, P, SRC, kThreads, 16>(p, pro, s_stage, src, span, tid);
        else if constexpr (sizeof(SRC) <= 4) {
            if ((bits & 7ull) == 0) done = stage_planar<R, P, SRC, kThreads, 8>(p, pro, s_stage, src, span, tid);
        }
    } else if (interior && C == 2 && p.samp_stride == 2 && p.chan_stride == 1) {
        const SRC* src = base + 2 * g0;
        const unsigned long long bits = reinterpret_cast<unsigned long long>(src);
        if ((bits & 15ull) == 0) done = stage_interleaved2<R, P, SRC, kThreads, 16>(pro, s_stage, src, span, tid);
        else if constexpr (sizeof(SRC) <= 4) {
            if ((bits & 7ull) == 0) done = stage_interleaved2<R, P, SRC, kThreads, 8>(pro, s_stage, src, span, tid);
        }
    }
    // general path: edges (reflection), strided or unaligned sources, vector tail
    constexpr int UG = 8;
    for (int i0 = done + tid; i0 < span; i0 += kThreads * UG) {
        P acc[UG];
        long long off[UG];
#pragma unroll
        for (int u = 0; u < UG; ++u) {
            const int i = i0 + u * kThreads;
            long long g = g0 + i;
            g = g < 0 ? -g : g;
            g = g >= n ? 2 * (n - 1) - g : g;
            g = i < span ? g : 0;
            off[u] = g * p.samp_stride;
            acc[u] = pro.load((P)__ldg(base + off[u]));
        }
        for (int c = 1; c < C; ++c) {
            P t[UG];
#pragma unroll
            for (int u = 0; u < UG; ++u) t[u] = (P)__ldg(base + off[u] + c * p.chan_stride);
#pragma unroll
            for (int u = 0; u < UG; ++u) acc[u] += pro.load(t[u]);
        }
#pragma unroll
        for (int u = 0; u < UG; ++u) {
            const int i = i0 + u * kThreads;
            if (i < span) s_stage[i] = (R)pro.finish(acc[u]);
        }
    }
}

// ---- mbarrier / bulk-copy (TMA) primitives ---------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(
                     (unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// one non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// L2 prefetch of the samples the NEXT tile of this CTA will stage (no registers, no smem).
template <int kThreads>
__device__ __forceinline__ void prefetch_tile(const StftParams& p, long long row, long long g0, int span,
                                              int elem_bytes, int tid) {
    if (p.samp_stride != 1) return;
    long long lo = g0 < 0 ? 0 : g0;
    long long hi = g0 + span > p.n_samples ? p.n_samples : g0 + span;
    const char* base = reinterpret_cast<const char*>(p.audio) + row_offset(p, row) * elem_bytes;
    const long long bytes = (hi - lo) * elem_bytes;
    for (int c = 0; c < p.channels; ++c) {
        const char* q = base + (c * p.chan_stride + lo) * elem_bytes;
        for (long long o = (long long)tid * 128; o < bytes; o += (long long)kThreads * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
    }
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// dB of one bin from its complex value.  float path: 10*log10(|X|^2) through MUFU.LG2, clamped in
// the dB domain (so the floor is exactly multiplier*log10(amin)).  double path: |X|^2 in double,
// log2 split into exponent + mantissa so the only fp32 error is lg2.approx on [1, 2) (2^-22
// absolute), recombined and clamped in double.
template <typename R>
struct DbOut;
template <>
struct DbOut<float> {
    using type = float;
    static __device__ __forceinline__ float eval(float re, float im, const StftParams& p) {
        const float pw = fmaf(re, re, im * im);
        return fmaxf(p.log_scale * lg2_approx(pw), p.floor_db);
    }
};
template <>
struct DbOut<double> {
    using type = double;
    static __device__ __forceinline__ double eval(double re, double im, const StftParams& p) {
        const float pw = (float)(re * re + im * im);
        const int bits = __float_as_int(pw);
        const int e = (bits >> 23) - 127;
        const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
        double db = p.log_scale_d * ((double)e + (double)lg2_approx(m));
        db = (pw >= 1.17549435e-38f) ? db : -INFINITY;
        return fmax(db, p.floor_db_d);
    }
};

template <typename R> __device__ __forceinline__ R shfl_xor_r(R v, int m);
template <> __device__ __forceinline__ float shfl_xor_r<float>(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ double shfl_xor_r<double>(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// Per-lane constants of the cross-lane 32-point DFT that turns Y_b[32] into bins 32 + 64 q:
// pre-twiddle W_64^lane and the decimation-in-frequency twiddles of the distance-16/8/4 stages
// (1 for the lower lane of a pair).  Lane r ends up holding q = bitrev5(r).
// They live in a [8][32] shared-memory table, not in registers: as eight loop-invariant registers
// ptxas spilled them to local memory around the codelets, and the reloads missed L1 93 % of the
// time (profiles/: 18 % of warp stall samples were L2 round trips of those reloads).
constexpr int kNyqTab = 8 * 32;
template <typename R>
__device__ __forceinline__ void nyq_table_init(R* tab, int tid) {
    if (tid < 32) {
        const int lane = tid;
        double sn, cs;
        sincospi(-(double)lane / 32.0, &sn, &cs);
        tab[0 * 32 + lane] = (R)cs; tab[1 * 32 + lane] = (R)sn;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int dist = 16 >> i;
            const bool upper = (lane & dist) != 0;
            sincospi(upper ? -(double)((lane & (dist - 1)) * (16 / dist)) / 16.0 : 0.0, &sn, &cs);
            tab[(2 + 2 * i) * 32 + lane] = (R)cs; tab[(3 + 2 * i) * 32 + lane] = (R)sn;
        }
    }
}

struct StatAcc {
    float sum = 0.f, sq = 0.f;
    double sum_d = 0.0, sq_d = 0.0, cnt = 0.0;
    __device__ __forceinline__ void add(float db) { sum += db; sq = fmaf(db, db, sq); cnt += 1.0; }
    __device__ __forceinline__ void fold() { sum_d += sum; sq_d += sq; sum = 0.f; sq = 0.f; }
    __device__ __forceinline__ void flush(double* stats, int lane) {
        fold();
        sum_d = dam_warp_sum(sum_d); sq_d = dam_warp_sum(sq_d); cnt = dam_warp_sum(cnt);
        if (lane == 0 && cnt > 0.0) {
            atomicAdd(stats + 0, sum_d);
            atomicAdd(stats + 1, sq_d);
            atomicAdd(stats + 2, cnt);
        }
    }
};

// Output stores are plain stores: streaming (st.global.cs) stores measured 2 % slower on the phase kernel.
template <typename OUT> __device__ __forceinline__ void store_out(OUT* o, OUT v) { *o = v; }

// k' = 32: y32 = Y_b[32] of lane b is real; bins 32 + 64 q = sum_b Y_b[32] W_64^b W_32^(bq) are a 32-point
// DFT ACROSS the lanes of the warp (radix-2 DIF over shuffles; lane r ends with q = bitrev5(r), q < 16 are
// the distinct bins) and go straight to global memory (out_frame = &out[row][0][t]).
template <typename R, typename OUT, bool kStats>
__device__ __forceinline__ void nyq_bins(R y32, OUT* out_frame, long long T, const R* nq,
                                         const StftParams& p, int lane, StatAcc& st, bool active = true) {
    nq += lane;
    R vr = y32 * nq[0], vi = y32 * nq[32];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int dist = 16 >> i;
        const R pr = shfl_xor_r<R>(vr, dist), pi = shfl_xor_r<R>(vi, dist);
        const bool upper = (lane & dist) != 0;
        const R tr = upper ? pr - vr : vr + pr;
        const R ti = upper ? pi - vi : vi + pi;
        if (i < 3) {
            const R wr = nq[(2 + 2 * i) * 32], wi = nq[(3 + 2 * i) * 32];
            vr = tr * wr - ti * wi;
            vi = tr * wi + ti * wr;
        } else if (i == 3) {                 // W_4^j: j = 1 on the upper lane of odd pairs -> -i
            const bool rot = upper && (lane & 1);
            vr = rot ? ti : tr;
            vi = rot ? -tr : ti;
        } else {
            vr = tr; vi = ti;
        }
    }
    const typename DbOut<R>::type db = DbOut<R>::eval(vr, vi, p);
    if ((lane & 1) == 0 && active) {
        const int bin = 32 + 64 * (int)(__brev((unsigned)lane) >> 27);
        store_out<OUT>(out_frame + (long long)bin * T, (OUT)db);
        if (kStats) st.add((float)db);
    }
}

// ---- stage 1 of one frame, executed by one warp (lane = b) -------------------------------------
// x: the lane's 64 windowed samples.  Writes the 32 exchange rows of this frame (dst = row 0 of the
// frame + lane) and, through the cross-lane DFT, the 16 bins 32 + 64 q straight to global memory
// (out_frame = &out[row][0][t]).
template <typename R, typename OUT, bool kStats>
__device__ __forceinline__ void stage1_frame(R (&x)[64], typename Cx<R>::type* dst, OUT* out_frame, long long T,
                                             const R* nq, const StftParams& p, int lane, StatAcc& st) {
    using C = typename Cx<R>::type;
    R sr[32], si[32];
    dam_rfft64_packed<R>(x, sr, si);
    {
        C v; v.x = sr[0]; v.y = R(0);
        dst[0] = v;                        // row 0: Y[0] (real)
    }
#pragma unroll
    for (int k = 1; k < 32; ++k) {
        C v; v.x = sr[k]; v.y = si[k];
        dst[k * kRow] = v;
    }
    nyq_bins<R, OUT, kStats>(si[0], out_frame, T, nq, p, lane, st);
}

// ---- stage 2 of one (k', frame) pair, executed by one thread ------------------------------------
// src: exchange row k' of the frame, tw: twiddle row k', out_frame = &out[row][0][t].
template <typename R, typename OUT, bool kStats>
__device__ __forceinline__ void stage2_task(const typename Cx<R>::type* src, const typename Cx<R>::type* tw,
                                            OUT* out_frame, long long T, int kp, const StftParams& p, StatAcc& st) {
    using C = typename Cx<R>::type;
    R sr[32], si[32];
    if (sizeof(R) == 4) {
        // two complex values per 128-bit load (rows are 16-byte aligned by construction)
        const float4* s4 = reinterpret_cast<const float4*>(src);
        const float4* w4 = reinterpret_cast<const float4*>(tw);
#pragma unroll
        for (int grp = 0; grp < 4; ++grp) {
            // four groups of four 128-bit loads each: keeps at most 32 load registers live next to the
            // 64 result registers (hoisting all 32 loads spills under the 128-register cap)
            float4 v[4], w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { v[u] = s4[grp * 4 + u]; w[u] = w4[grp * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int b2 = grp * 4 + u;
                sr[2 * b2] = v[u].x * w[u].x - v[u].y * w[u].y;
                si[2 * b2] = v[u].x * w[u].y + v[u].y * w[u].x;
                sr[2 * b2 + 1] = v[u].z * w[u].z - v[u].w * w[u].w;
                si[2 * b2 + 1] = v[u].z * w[u].w + v[u].w * w[u].z;
            }
            asm volatile("" ::: "memory");
        }
    } else {
#pragma unroll
        for (int b = 0; b < 32; ++b) {
            const C v = src[b];
            const C w = tw[b];
            sr[b] = v.x * w.x - v.y * w.y;
            si[b] = v.x * w.y + v.y * w.x;
        }
    }
    dam_cfft32<R>(sr, si);

    // bins k'+64q (q < 16) walk up from row k'; mirrors 2048-k'-64q (q >= 16) walk down from row
    // 1024-k'.  k' = 0 owns only q <= 16 (bins 0, 64, ..., 1024).
    const int sb = (int)(64 * T * (long long)sizeof(OUT));      // one IMAD.WIDE per store (see stage2_task_pc)
    char* o_lo = reinterpret_cast<char*>(out_frame + kp * T);
    char* o_hi = reinterpret_cast<char*>(out_frame + (kNfft / 2 - kp) * T);
    const bool mirror_ok = kp != 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        const bool valid = q <= 16 ? true : mirror_ok;
        OUT* o = reinterpret_cast<OUT*>(q < 16 ? o_lo + (long long)q * (long long)sb : o_hi + (long long)(16 - q) * (long long)sb);
        const typename DbOut<R>::type db = DbOut<R>::eval(sr[q], si[q], p);
        if (valid) store_out<OUT>(o, (OUT)db);
        if (kStats && valid) st.add((float)db);
    }
    if (kStats) st.fold();
}

// ---- packed-complex forms of the two stages (float32): one complex value per 64-bit register,
// arithmetic on sm_100's packed FP32 instructions (FADD2 / FMUL2 / FFMA2).  ptxas folds half swaps,
// per-half signs and scalar broadcasts into operand modifiers, so a complex add is 1 instruction, a
// twiddle butterfly 3 and a complex multiply 2 -- half the issue slots of the scalar codelets with the
// same 64 data registers per thread (tools/gen_codelets.py, dam_cfft32_pc / dam_rfft64_pc).
template <typename OUT, bool kStats>
__device__ __forceinline__ void stage1_frame_pc(pk2 (&xp)[32], float2* dst, OUT* out_frame, long long T,
                                                const float* nq, const StftParams& p, int lane, StatAcc& st) {
    pk2 y0n, s[32];
    dam_rfft64_pc(xp, y0n, s);
    float y0, y32;
    pk_split(y0n, y0, y32);
    dst[0] = make_float2(y0, 0.f);                       // row 0: Y[0] (real)
#pragma unroll
    for (int k = 1; k < 32; ++k) *reinterpret_cast<pk2*>(dst + k * kRow) = s[k];
    nyq_bins<float, OUT, kStats>(y32, out_frame, T, nq, p, lane, st);
}

// kPitch > 0: the bin pitch of the output plane (elements) is a compile-time constant, so the 32 stores of a task
// are two base registers plus immediate offsets (STG [R.64 + imm]) instead of one IMAD.WIDE each.
// xch_free (optional): an mbarrier on which lane 0 of the warp arrives as soon as the warp has consumed its
// exchange rows, so that the next tile's stage 1 may overwrite them while this task is still transforming.
// active = false: the task runs on a duplicate of a valid frame (ragged last tile) and stores nothing.
template <typename OUT, bool kStats, bool kGlobalTw = false, int kPitch = 0>
__device__ __forceinline__ void stage2_task_pc(const float2* src, const float2* tw, OUT* out_frame, long long T, int kp,
                                               const StftParams& p, StatAcc& st, unsigned long long* xch_free = nullptr,
                                               bool active = true, const float* nq = nullptr, float y32 = 0.f,
                                               OUT* nyq_frame = nullptr, bool nyq_active = false) {
    pk2 z[32];
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
            // v * w = (vr wr - vi wi, vi wr + vr wi) = v * (wr, wr) + swap(v) * (-wi, wi)
            z[2 * b2] = pk_fma(pk_swap(v[u].x), pk_make(-w[u].y, w[u].y), pk_mul(v[u].x, pk_c(w[u].x)));
            z[2 * b2 + 1] = pk_fma(pk_swap(v[u].y), pk_make(-w[u].w, w[u].w), pk_mul(v[u].y, pk_c(w[u].z)));
        }
        asm volatile("" ::: "memory");
    }
    if (xch_free != nullptr) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(xch_free);
    }
    dam_cfft32_pc(z);
    // bins 32 + 64 q of the frame this WARP transformed in stage 1 (Y_b[32] kept in a register): the shuffle chain
    // is issued here so that its latency hides behind the epilogue below (same basic block, no divergence)
    if (nq != nullptr) nyq_bins<float, OUT, kStats>(y32, nyq_frame, T, nq, p, (int)(threadIdx.x & 31), st, nyq_active);
    // bins k'+64q (q < 16) walk up from row k'; mirrors 2048-k'-64q (q >= 16) walk down from row 1024-k'.
    // k' = 0 owns only q <= 16 (bins 0, 64, ..., 1024).
    const bool mirror_ok = kp != 0 && active;
    if (kPitch > 0) {
        OUT* o_lo = out_frame + kp * kPitch;
        OUT* o_hi = out_frame + (kNfft / 2 - kp) * kPitch;
        // four dB values at a time into distinct registers, then their four stores: a single result register
        // makes every FMNMX wait for the previous store to have read it (profiles/: long-scoreboard stalls)
#pragma unroll
        for (int q0 = 0; q0 < 32; q0 += 4) {
            float db[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float re, im;
                pk_split(z[q0 + u], re, im);
                db[u] = DbOut<float>::eval(re, im, p);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = q0 + u;
                const bool valid = q <= 16 ? active : mirror_ok;
                OUT* o = q < 16 ? o_lo + q * (64 * kPitch) : o_hi - (q - 16) * (64 * kPitch);
                if (valid) store_out<OUT>(o, (OUT)db[u]);
                if (kStats && valid) st.add(db[u]);
            }
        }
    } else {
        // one IMAD.WIDE per store: base + (compile-time q) * (32-bit byte stride); the launcher keeps 64 rows < 2 GiB
        const int sb = (int)(64 * T * (long long)sizeof(OUT));
        char* o_lo = reinterpret_cast<char*>(out_frame + kp * T);
        char* o_hi = reinterpret_cast<char*>(out_frame + (kNfft / 2 - kp) * T);
#pragma unroll
        for (int q0 = 0; q0 < 32; q0 += 4) {
            float db[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float re, im;
                pk_split(z[q0 + u], re, im);
                db[u] = DbOut<float>::eval(re, im, p);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = q0 + u;
                const bool valid = q <= 16 ? active : mirror_ok;
                OUT* o = reinterpret_cast<OUT*>(q < 16 ? o_lo + (long long)q * (long long)sb : o_hi + (long long)(16 - q) * (long long)sb);
                if (valid) store_out<OUT>(o, (OUT)db[u]);
                if (kStats && valid) st.add(db[u]);
            }
        }
    }
    if (kStats) st.fold();
}

}  // namespace
