// K2 / K3 / K5 / K6 / N2: the bandwidth-bound elementwise and reduction kernels of the mix path.
// All of them stream their operands once with 128-bit accesses; grids are sized in multiples of
// the SM count and grid-stride over the data.
#include "dam_common.cuh"
#include "kweight_common.cuh"
#include <type_traits>

#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int kBlock = 256;

inline int grid_for(long long work_items, int per_sm = 8) {
    long long blocks = (work_items + kBlock - 1) / kBlock;
    long long cap = (long long)dam_sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* s_buf) {
    v = dam_warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? s_buf[lane] : T(0);
        v = dam_warp_sum(v);
    }
    __syncthreads();
    return v;
}
template <typename T>
__device__ __forceinline__ T block_max(T v, T* s_buf) {
    v = dam_warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? s_buf[lane] : T(0);
        v = dam_warp_max(v);
    }
    __syncthreads();
    return v;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    // valid for non-negative doubles: their bit patterns order like signed integers
    atomicMax(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
}

template <typename T, int V> struct alignas(V * sizeof(T) >= 16 ? 16 : V * sizeof(T)) Vec { T v[V]; };
template <typename T, int V>
__device__ __forceinline__ Vec<T, V> load_vec(const T* p) {
    Vec<T, V> r;
    constexpr int kBytes = V * (int)sizeof(T);
    static_assert(V == 1 || kBytes == 8 || kBytes == 16 || kBytes == 32, "load_vec: unsupported vector width");
    if (V == 1) {
        r.v[0] = __ldg(p);
    } else if (kBytes == 8) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
        *reinterpret_cast<uint2*>(&r) = raw;
    } else if (kBytes == 16) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
        *reinterpret_cast<uint4*>(&r) = raw;
    } else {      // 32 bytes
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
        reinterpret_cast<uint4*>(&r)[0] = a;
        reinterpret_cast<uint4*>(&r)[1] = b;
    }
    return r;
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const Vec<T, V>& r) {
    constexpr int kBytes = V * (int)sizeof(T);
    static_assert(V == 1 || kBytes == 8 || kBytes == 16 || kBytes == 32, "store_vec: unsupported vector width");
    if (V == 1) {
        *p = r.v[0];
    } else if (kBytes == 8) {
        *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(&r);
    } else if (kBytes == 16) {
        *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&r);
    } else {
        reinterpret_cast<uint4*>(p)[0] = reinterpret_cast<const uint4*>(&r)[0];
        reinterpret_cast<uint4*>(p)[1] = reinterpret_cast<const uint4*>(&r)[1];
    }
}

// Product and sum with SEPARATE roundings (numpy: `track * mask` is a rounded array, then np.sum adds): the
// intrinsics are never contracted into an FMA, which would change the last bit of the mix.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// ------------------------------------------------------------------------------------------
// K2 gain-mask apply + stem sum
struct GainMixParams {
    const void* stems;
    long long n_batch, batch_stride;
    long long n_stems, n_channels, n_samples, stem_stride, chan_stride;
    const double* gains;
    long long n_gains, interp_coef;
    int zero_mask;                 // n_gains == 1 without hold_single: mask is all zeros
    void* mix_out;
    void* stems_out;
    double* peak;
    double* sumsq;
};

// Per-plane constants of one (song, channel) of the mixdown.
template <typename TS, typename TO>
struct GainMixPlane {
    const TS* base;
    TO* mix;
    TO* stems_out;                 // plane of stem 0 (stems out_stem_stride apart), or null
    const double* gains;
    long long n, K, coef;
    long long stem_stride, out_stem_stride;
    int n_stems;
    bool zero_mask;
};

template <typename TS, typename TO>
__device__ __forceinline__ GainMixPlane<TS, TO> gain_mix_plane(const GainMixParams& p, long long c, long long bt) {
    GainMixPlane<TS, TO> q;
    q.n = p.n_samples;
    q.base = reinterpret_cast<const TS*>(p.stems) + bt * p.batch_stride + c * p.chan_stride;
    q.mix = p.mix_out ? reinterpret_cast<TO*>(p.mix_out) + (bt * p.n_channels + c) * q.n : nullptr;
    q.K = p.n_gains;
    q.coef = p.interp_coef;
    q.gains = p.gains + bt * p.n_stems * q.K;
    q.stems_out = p.stems_out ? reinterpret_cast<TO*>(p.stems_out) + bt * p.n_stems * p.n_channels * q.n + c * q.n : nullptr;
    q.stem_stride = p.stem_stride;
    q.out_stem_stride = p.n_channels * q.n;
    q.n_stems = (int)p.n_stems;
    q.zero_mask = p.zero_mask != 0;
    return q;
}

// The run of samples that share the gain of sample i0: k0 = min(i0 / coef, K-1) and the first sample of the next
// gain (LLONG_MAX on the last one).  A thread walks its samples in ascending order, so this (with its 64-bit
// division) runs once per gain run, not once per vector.
struct GainRun { int k0; long long end; };

template <typename TS, typename TO>
__device__ __forceinline__ GainRun gain_run(const GainMixPlane<TS, TO>& q, long long i0) {
    GainRun r;
    r.k0 = (int)q.K - 1;
    r.end = 0x7fffffffffffffffLL;
    if (q.coef > 0) {
        const long long kk = i0 / q.coef;
        if (kk < q.K - 1) { r.k0 = (int)kk; r.end = (kk + 1) * q.coef; }
    }
    return r;
}

// One stem of NV vectors (vgap samples apart) that lie inside one gain run: product rounded, added to the running
// sum in stem order (numpy: `track * mask` is a rounded array, np.sum starts from the first one), optionally stored.
template <typename TS, typename TC, typename TO, int V, int NV, bool kFirst>
__device__ __forceinline__ void mix_stem(const Vec<TS, V> (&x)[NV], TC g, TC (&acc)[NV][V], TO* po, long long vgap) {
#pragma unroll
    for (int h = 0; h < NV; ++h) {
        Vec<TO, V> y;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const TC prod = mul_rn((TC)x[h].v[j], g);
            acc[h][j] = kFirst ? prod : add_rn(acc[h][j], prod);
            y.v[j] = (TO)prod;
        }
        if (po) store_vec<TO, V>(po + h * vgap, y);
    }
}

// All stems of NV vectors inside one gain run.  Stems are taken SU at a time (their loads are in flight together);
// source, gain and output pointers advance by additions -- the 64-bit multiplies and the per-stem `s < n_stems`
// predicates of the first version were most of its 343 instructions per vector (ncu: 96 IMAD, 32 LDC, 27 ISETP,
// 20 BRA against 45 floating-point instructions).
template <typename TS> struct StemsInFlight { static constexpr int v = sizeof(TS) == 4 ? 4 : 2; };

// SU stems of NV vectors with their gains: what a thread keeps in flight
template <typename TS, typename TC, int V, int NV>
struct StemGroup {
    static constexpr int SU = StemsInFlight<TS>::v;
    Vec<TS, V> x[SU][NV];
    TC g[SU];
    template <typename TO>
    __device__ __forceinline__ void load(const GainMixPlane<TS, TO>& q, const TS* ps, const double* pg, long long vgap) {
#pragma unroll
        for (int u = 0; u < SU; ++u) {
#pragma unroll
            for (int h = 0; h < NV; ++h) x[u][h] = load_vec<TS, V>(ps + u * q.stem_stride + h * vgap);
            g[u] = q.zero_mask ? TC(0) : (TC)__ldg(pg + u * q.K);
        }
    }
};

template <typename TS, typename TC, typename TO, int V, int NV>
__device__ __forceinline__ void mix_stems(const GainMixPlane<TS, TO>& q, const TS* ps, const double* pg, TO* po,
                                          long long vgap, TC (&acc)[NV][V], StemGroup<TS, TC, V, NV>& grp, bool loaded) {
    constexpr int SU = StemsInFlight<TS>::v;
    const long long ss = q.stem_stride, os = q.out_stem_stride, K = q.K;
    const int n_stems = q.n_stems;
    int s = 0;
    if (n_stems >= SU) {
        if (!loaded) grp.load(q, ps, pg, vgap);            // `loaded`: the caller prefetched the first SU stems
        mix_stem<TS, TC, TO, V, NV, true>(grp.x[0], grp.g[0], acc, po, vgap);
#pragma unroll
        for (int u = 1; u < SU; ++u) mix_stem<TS, TC, TO, V, NV, false>(grp.x[u], grp.g[u], acc, po ? po + u * os : nullptr, vgap);
        s = SU; ps += SU * ss; pg += SU * K; if (po) po += SU * os;
        for (; s + SU <= n_stems; s += SU) {
            grp.load(q, ps, pg, vgap);
#pragma unroll
            for (int u = 0; u < SU; ++u) mix_stem<TS, TC, TO, V, NV, false>(grp.x[u], grp.g[u], acc, po ? po + u * os : nullptr, vgap);
            ps += SU * ss; pg += SU * K; if (po) po += SU * os;
        }
    }
    for (; s < n_stems; ++s) {
        Vec<TS, V> x[NV];
#pragma unroll
        for (int h = 0; h < NV; ++h) x[h] = load_vec<TS, V>(ps + h * vgap);
        const TC g = q.zero_mask ? TC(0) : (TC)__ldg(pg);
        if (s == 0) mix_stem<TS, TC, TO, V, NV, true>(x, g, acc, po, vgap);
        else mix_stem<TS, TC, TO, V, NV, false>(x, g, acc, po, vgap);
        ps += ss; pg += K; if (po) po += os;
    }
}

// A vector with a gain boundary inside it (once per interp_coef samples): per-sample gain look-up.
template <typename TS, typename TC, typename TO, int V>
__device__ __forceinline__ void mix_boundary_vector(const GainMixPlane<TS, TO>& q, long long i0, const GainRun& run, TC (&acc)[V]) {
    const int cross = (int)(run.end - i0);                 // samples of this vector that still belong to gain k0
    const TS* src = q.base + i0;
    TO* sout = q.stems_out ? q.stems_out + i0 : nullptr;
    for (int s = 0; s < q.n_stems; ++s) {
        const Vec<TS, V> x = load_vec<TS, V>(src + (long long)s * q.stem_stride);
        const double g0 = __ldg(q.gains + (long long)s * q.K + run.k0);
        Vec<TO, V> y;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            double gd = g0;
            if (j >= cross) {
                long long k = (i0 + j) / q.coef;
                if (k > q.K - 1) k = q.K - 1;
                gd = __ldg(q.gains + (long long)s * q.K + k);
            }
            const TC g = q.zero_mask ? TC(0) : (TC)gd;
            const TC prod = mul_rn((TC)x.v[j], g);
            acc[j] = s == 0 ? prod : add_rn(acc[j], prod);
            y.v[j] = (TO)prod;
        }
        if (sout) store_vec<TO, V>(sout + (long long)s * q.out_stem_stride, y);
    }
}

// the mix in its output dtype, running peak and sum of squares, optional store
template <typename TC, typename TO, int V>
__device__ __forceinline__ void finish_vector(const TC (&acc)[V], Vec<TO, V>& o, TO* mix_at, double& loc_peak, double& loc_sq) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
        o.v[j] = (TO)acc[j];
        const double a = fabs((double)acc[j]);
        loc_peak = fmax(loc_peak, a);
        loc_sq += a * a;
    }
    if (mix_at) store_vec<TO, V>(mix_at, o);
}

// Samples [v * V, (v + 1) * V) of one plane (the general entry: locates the gain run itself).
template <typename TS, typename TC, typename TO, int V>
__device__ __forceinline__ void gain_mix_vector(const GainMixPlane<TS, TO>& q, long long v, GainRun& run,
                                                Vec<TO, V>& o, double& loc_peak, double& loc_sq) {
    const long long i0 = v * V;
    if (i0 >= run.end || run.k0 < 0) run = gain_run(q, i0);
    TC acc[1][V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[0][j] = TC(0);
    if (i0 + V <= run.end) {
        StemGroup<TS, TC, V, 1> grp;
        mix_stems<TS, TC, TO, V, 1>(q, q.base + i0, q.gains + run.k0, q.stems_out ? q.stems_out + i0 : nullptr, 0, acc, grp, false);
    } else
        mix_boundary_vector<TS, TC, TO, V>(q, i0, run, acc[0]);
    finish_vector<TC, TO, V>(acc[0], o, q.mix ? q.mix + i0 : nullptr, loc_peak, loc_sq);
}

template <typename TS, typename TC, typename TO, int V>
__global__ void __launch_bounds__(kBlock, 4) gain_mix_kernel(const GainMixParams p) {
    __shared__ double s_red[kBlock / 32];
    const long long c = blockIdx.y;
    const long long bt = blockIdx.z;
    const GainMixPlane<TS, TO> q = gain_mix_plane<TS, TO>(p, c, bt);
    const long long nvec = (q.n + V - 1) / V;
    double loc_peak = 0.0, loc_sq = 0.0;
    GainRun run;
    run.k0 = -1; run.end = 0;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        Vec<TO, V> o;
        gain_mix_vector<TS, TC, TO, V>(q, v, run, o, loc_peak, loc_sq);
    }
    if (p.peak) {
        const double m = block_max(loc_peak, s_red);
        if (threadIdx.x == 0) atomic_max_nonneg(p.peak + bt * p.n_channels + c, m);
    }
    if (p.sumsq) {
        const double m = block_sum(loc_sq, s_red);
        if (threadIdx.x == 0) atomicAdd(p.sumsq + bt * p.n_channels + c, m);
    }
}

// K2 + pass A of K4 in one kernel (dam_gain_mix_kweight): the mix is about to be measured by the BS.1770 meter, whose
// first pass over it computes, for every 256-sample chunk, the end state of the two K-weighting biquads from a zero
// start state -- a dot product of the chunk with a response table.  Here a WARP owns a chunk (lane l holds samples
// 4 l .. 4 l + 3 and 128 + 4 l .. 128 + 4 l + 3), multiplies its 8 mix samples with its 8 x 4 table entries (shared
// memory, lane-major: conflict-free 16-byte loads) and the 32 x 2 group sums are added in the order of
// kweight_common.cuh by a transposed butterfly: a lane sends half of what it holds and keeps the other half, so the
// 8 values per lane cost 20 shuffles instead of 80.  The meter then starts at its scan; the mix is read once less.
struct KwFuseLite {                // what the chunk kernel needs without its meter part
    KwStateLayout lay;
    long long n_chunks;
    unsigned cpp_inv, super_inv;   // floor(2^32 / cpp), floor(2^32 / super) (0xffffffff for a divisor of 1)
};
struct KwFuseParams : KwFuseLite {
    double G[kKwChunk * 4];        // response table, in the kernel parameters like K4's own
};

// x / d for x < 2^32 with inv = floor(2^32 / d): the estimate is at most one short
__device__ __forceinline__ unsigned div_u32(unsigned x, unsigned d, unsigned inv) {
    unsigned qt = __umulhi(x, inv);
    if (x - qt * d >= d) ++qt;
    return qt;
}

// kw_state_ptr without its two 64-bit divisions (a chunk index fits 32 bits: they were ~100 of the instructions a warp
// spends per chunk)
__device__ __forceinline__ double* kw_state_ptr_fast(const KwFuseLite& kw, long long row, unsigned chunk) {
    const unsigned cpp = (unsigned)kw.lay.cpp, super = (unsigned)kw.lay.super;
    const unsigned part = div_u32(chunk, cpp, kw.cpp_inv);
    const unsigned cl = chunk - part * cpp;
    const unsigned u = div_u32(cl, super, kw.super_inv);
    const unsigned k = cl - u * super;
    return kw.lay.states + ((row * kw.lay.parts + part) * kw.lay.cpp + (long long)k * kKwScanBlock + u) * 4;
}

#ifndef DAM_K2KW_CTAS
#define DAM_K2KW_CTAS 2
#endif
template <typename TS, typename TC, typename TO, bool kKw>
__global__ void __launch_bounds__(kBlock, DAM_K2KW_CTAS)
gain_mix_chunks_kernel(const GainMixParams p, const typename std::conditional<kKw, KwFuseParams, KwFuseLite>::type kw) {
    constexpr int V = 4;
    // [half][2 j + (slot >> 1)][lane] = G[(half * 32 + lane) * 4 + j][slot ^ (lane & 3)] for slot = 2 (k & 1), + 1: a lane
    // keeps its four sums in SLOTS ordered by when they leave it in the butterfly (slot 0 = the component the lane
    // ends up with, lane & 3; slots 1 and 3 go to lane ^ 1, slot 2 to lane ^ 2), so no step has to select
    __shared__ double2 s_G[kKw ? 2 : 1][kKw ? 8 : 1][kKw ? 32 : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if constexpr (kKw) {
        for (int e = threadIdx.x; e < 2 * 8 * 32; e += kBlock) {
            const int l = e & 31, k = (e >> 5) & 7, h = e >> 8;
            const int base = ((h * 32 + l) * 4 + (k >> 1)) * 4, slot = (k & 1) * 2, cs = l & 3;
            s_G[h][k][l] = make_double2(kw.G[base + (slot ^ cs)], kw.G[base + ((slot + 1) ^ cs)]);
        }
        __syncthreads();
    }
    // PERSISTENT grid, one contiguous range of (plane, chunk) items per WARP (planes = (song, channel) in row order):
    // with one CTA per slice of a plane the ~5 us a CTA spends on this table, its launch and its block reduction came
    // to 20 us per bench step and the grid had to be sized around whole waves of CTAs.  A warp's range touches one or
    // two planes; it leaves its peak / sum of squares at every plane change (shuffle reduction + two atomics).
    const long long n = p.n_samples;
    const long long nvec = n / V;              // the launcher guarantees n % V == 0
    const long long planes = p.n_batch * p.n_channels;
    const long long total = planes * kw.n_chunks;
    const long long wid = (long long)blockIdx.x * (kBlock / 32) + warp, nw = (long long)gridDim.x * (kBlock / 32);
    const long long t0 = total / nw * wid + (total % nw) * wid / nw;             // = floor(total * wid / nw) without overflow
    const long long t1 = total / nw * (wid + 1) + (total % nw) * (wid + 1) / nw;
    long long row = t0 / kw.n_chunks;
    long long ch = t0 - row * kw.n_chunks;
    GainMixPlane<TS, TO> q = gain_mix_plane<TS, TO>(p, row % p.n_channels, row / p.n_channels);
    double loc_peak = 0.0, loc_sq = 0.0;
    GainRun run;
    run.k0 = -1; run.end = 0;
    // The first SU stems of the NEXT chunk are requested before the epilogue of this one (rounding, statistics,
    // store and, with the meter part, ~130 instructions of table products and butterfly that touch no global memory).
    StemGroup<TS, TC, V, 2> grp;
    bool prefetched = false;                   // warp-uniform: grp holds the first stems of chunk `ch` of this plane
    for (long long item = t0; item < t1; ++item, ++ch) {
        if (ch == kw.n_chunks) {               // next plane
            const double m = dam_warp_max(loc_peak), sq = dam_warp_sum(loc_sq);
            if (lane == 0 && p.peak) atomic_max_nonneg(p.peak + row, m);
            if (lane == 0 && p.sumsq) atomicAdd(p.sumsq + row, sq);
            loc_peak = 0.0; loc_sq = 0.0;
            prefetched = false;
            ++row; ch = 0;
            q = gain_mix_plane<TS, TO>(p, row % p.n_channels, row / p.n_channels);
            run.k0 = -1; run.end = 0;
        }
        Vec<TO, V> o[2];
        // warp-uniform: is the chunk whole and under ONE gain per stem?  Then the loads of both of a lane's vectors
        // (2 x SU stems) are in flight together.
        const long long i0c = ch * kKwChunk;
        if (i0c >= run.end || run.k0 < 0) run = gain_run(q, i0c);
        if (i0c + kKwChunk <= run.end && i0c + kKwChunk <= n) {
            TC acc[2][V];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < V; ++j) acc[h][j] = TC(0);
            const long long at = i0c + lane * V;
            mix_stems<TS, TC, TO, V, 2>(q, q.base + at, q.gains + run.k0, q.stems_out ? q.stems_out + at : nullptr,
                                        kKwChunk / 2, acc, grp, prefetched);
            // the first stems of the NEXT chunk are requested now: they fly while this chunk is rounded, stored,
            // measured (peak, sum of squares) and, with the meter part, run through the butterfly
            prefetched = false;
            {
                const long long nxt = i0c + 2 * kKwChunk;      // end of the next chunk
                if (q.n_stems >= StemsInFlight<TS>::v && item + 1 < t1 && ch + 1 < kw.n_chunks && nxt <= run.end && nxt <= n) {
                    grp.load(q, q.base + at + kKwChunk, q.gains + run.k0, kKwChunk / 2);
                    prefetched = true;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
                finish_vector<TC, TO, V>(acc[h], o[h], q.mix ? q.mix + at + h * (kKwChunk / 2) : nullptr, loc_peak, loc_sq);
        } else {
            prefetched = false;
            GainRun r2 = run;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long v = ch * (kKwChunk / V) + h * 32 + lane;
#pragma unroll
                for (int j = 0; j < V; ++j) o[h].v[j] = TO(0);
                if (v < nvec) gain_mix_vector<TS, TC, TO, V>(q, v, r2, o[h], loc_peak, loc_sq);
            }
        }
        if constexpr (kKw) {
            double val[8];
    #pragma unroll
            for (int h = 0; h < 2; ++h) {
                double g[16];
    #pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const double2 t = s_G[h][k][lane];
                    g[(k >> 1) * 4 + (k & 1) * 2] = t.x;
                    g[(k >> 1) * 4 + (k & 1) * 2 + 1] = t.y;
                }
                kw_group_sum(g, (double)o[h].v[0], (double)o[h].v[1], (double)o[h].v[2], (double)o[h].v[3], val + h * 4);
            }
            // val[0..3]: group `lane` (first half A of the chunk), val[4..7]: group 32 + lane (second half B); slot s of
            // either holds component s ^ (lane & 3).  Transposed butterfly: a lane hands over half of what it holds.
            const unsigned full = 0xffffffffu;
            const double w0 = __dadd_rn(val[0], __shfl_xor_sync(full, val[1], 1));
            const double w1 = __dadd_rn(val[2], __shfl_xor_sync(full, val[3], 1));
            const double w2 = __dadd_rn(val[4], __shfl_xor_sync(full, val[5], 1));
            const double w3 = __dadd_rn(val[6], __shfl_xor_sync(full, val[7], 1));
            const double u0 = __dadd_rn(w0, __shfl_xor_sync(full, w1, 2));       // A, component lane & 3, over 4 lanes
            const double u1 = __dadd_rn(w2, __shfl_xor_sync(full, w3, 2));       // B
            const bool upper = (lane & 4) != 0;                                  // lanes with bit 2 set carry B from here on
            double t = __dadd_rn(upper ? u1 : u0, __shfl_xor_sync(full, upper ? u0 : u1, 4));
            t = __dadd_rn(t, __shfl_xor_sync(full, t, 8));
            t = __dadd_rn(t, __shfl_xor_sync(full, t, 16));
            t = __dadd_rn(t, __shfl_xor_sync(full, t, 4));                       // A + B
            if (lane < 4) {
                // the last chunk of a row (possibly partial) has no successor: its end state is never used
                kw_state_ptr_fast(kw, row, (unsigned)ch)[lane] = ch == kw.n_chunks - 1 ? 0.0 : t;
            }
        }
    }
    if (t1 > t0) {
        const double m = dam_warp_max(loc_peak), sq = dam_warp_sum(loc_sq);
        if (lane == 0 && p.peak) atomic_max_nonneg(p.peak + row, m);
        if (lane == 0 && p.sumsq) atomicAdd(p.sumsq + row, sq);
    }
}

template <typename TS, typename TC, typename TO>
int launch_gain_mix_kweight(const GainMixParams& p, const KwFuseParams& kw, cudaStream_t st);

template <typename TS, typename TC, typename TO>
int launch_gain_mix(const GainMixParams& p, bool vec_ok, cudaStream_t st) {
    if (vec_ok) {
        // the persistent chunk kernel without its meter part (states == nullptr): 6 % faster than one CTA per slice
        // of a plane (no wave quantisation, no per-CTA block reduction, two vectors per lane in flight)
        KwFuseParams kw;
        kw.lay.states = nullptr;
        kw.n_chunks = (p.n_samples + kKwChunk - 1) / kKwChunk;
        return launch_gain_mix_kweight<TS, TC, TO>(p, kw, st);
    }
    // unaligned planes or n % 4 != 0: one sample per thread
    dim3 grid(grid_for(p.n_samples, 8), (unsigned)p.n_channels, (unsigned)p.n_batch);
    { gain_mix_kernel<TS, TC, TO, 1><<<grid, kBlock, 0, st>>>(p); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

template <typename TS, typename TC, typename TO>
int launch_gain_mix_kweight(const GainMixParams& p, const KwFuseParams& kw, cudaStream_t st) {
    // persistent: DAM_K2KW_CTAS resident CTAs per SM (two: with both of a lane's vectors in flight and the butterfly the
    // kernel needs ~100 registers; at 80 or 64 it spills and is 35-60 % slower, measured); a warp takes a contiguous
    // range of whole chunks
    const long long total = p.n_channels * p.n_batch * kw.n_chunks;
    long long blocks = (long long)DAM_K2KW_CTAS * dam_sm_count();
    const long long max_blocks = (total + kBlock / 32 - 1) / (kBlock / 32);
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    dim3 grid((unsigned)blocks);
    if (kw.lay.states) { gain_mix_chunks_kernel<TS, TC, TO, true><<<grid, kBlock, 0, st>>>(p, kw); dam_note_launch(); }
    else { gain_mix_chunks_kernel<TS, TC, TO, false><<<grid, kBlock, 0, st>>>(p, static_cast<const KwFuseLite&>(kw)); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

// ------------------------------------------------------------------------------------------
// N2 bus sum
struct BusParams {
    const void* stems;
    long long batch_stride;
    long long n_stems, n_elems, stem_stride;
    int n_bus;
    void* out;
    int group[64];
    // optional channel-mean output (the mono down-mix K1 would otherwise compute from the stereo bus)
    void* mono_out;                // [n_batch, n_bus, n_elems / n_channels] or null
    int n_channels;                // n_elems = n_channels * samples per channel
    unsigned char order[64];       // stems sorted by bus (stable); bus g owns order[start[g] .. start[g+1])
    unsigned char start[65];
};

template <typename T, int V>
__global__ void __launch_bounds__(kBlock) bus_sum_kernel(const BusParams p) {
    const long long nvec = (p.n_elems + V - 1) / V;
    const T* in = reinterpret_cast<const T*>(p.stems) + (long long)blockIdx.y * p.batch_stride;
    T* outp = reinterpret_cast<T*>(p.out) + (long long)blockIdx.y * p.n_bus * p.n_elems;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const long long i0 = v * V;
        for (int g = 0; g < p.n_bus; ++g) {
            Vec<T, V> acc;
#pragma unroll
            for (int j = 0; j < V; ++j) acc.v[j] = T(0);
            for (long long s = 0; s < p.n_stems; ++s) {
                if (p.group[s] != g) continue;
                const Vec<T, V> x = load_vec<T, V>(in + s * p.stem_stride + i0);
#pragma unroll
                for (int j = 0; j < V; ++j) acc.v[j] += x.v[j];
            }
            store_vec<T, V>(outp + (long long)g * p.n_elems + i0, acc);
        }
    }
}


// Bus sum that also writes the channel mean of every bus (data/dataset.py:181-183 applied to the bus): K1 then
// reads a mono float32 row through its plain asynchronous staging path instead of averaging two channels per
// tile.  Arithmetic per sample is exactly K1's: channels added in order in float32, then * 1/C (C a power of
// two) or / C, so the features are bit-identical (tested).  kI16: PCM16 stems, x * scale first.
template <int V, bool kI16>
__global__ void __launch_bounds__(kBlock) bus_sum_mono_kernel(const BusParams p, float scale) {
    using TS = typename std::conditional<kI16, short, float>::type;
    const int C = p.n_channels;
    const long long n = p.n_elems / C;                     // samples per channel
    const long long nvec = (n + V - 1) / V;
    const TS* in = reinterpret_cast<const TS*>(p.stems) + (long long)blockIdx.y * p.batch_stride;
    float* outp = reinterpret_cast<float*>(p.out) + (long long)blockIdx.y * p.n_bus * p.n_elems;
    float* monop = reinterpret_cast<float*>(p.mono_out) + (long long)blockIdx.y * p.n_bus * n;
    const float f_c = (float)C;
    const float inv_c = ((C & (C - 1)) == 0) ? 1.0f / f_c : 0.0f;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const long long i0 = v * V;
        for (int g = 0; g < p.n_bus; ++g) {
            const int s0 = p.start[g], s1 = p.start[g + 1];
            float mono[V];
#pragma unroll
            for (int j = 0; j < V; ++j) mono[j] = 0.f;
            if (C == 2) {
                // both channels of up to three stems in flight (six independent loads), then the adds in stem order
                float acc[2][V];
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[c][j] = 0.f;
                for (int q0 = s0; q0 < s1; q0 += 3) {
                    Vec<TS, V> x[2][3];
#pragma unroll
                    for (int u = 0; u < 3; ++u)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (q0 + u < s1) x[c][u] = load_vec<TS, V>(in + (long long)p.order[q0 + u] * p.stem_stride + c * n + i0);
#pragma unroll
                    for (int u = 0; u < 3; ++u)
                        if (q0 + u < s1) {
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int j = 0; j < V; ++j)
                                    acc[c][j] = kI16 ? __fadd_rn(acc[c][j], __fmul_rn((float)x[c][u].v[j], scale))
                                                     : acc[c][j] + (float)x[c][u].v[j];
                        }
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    Vec<float, V> o;
#pragma unroll
                    for (int j = 0; j < V; ++j) o.v[j] = acc[c][j];
                    store_vec<float, V>(outp + (long long)g * p.n_elems + c * n + i0, o);
                }
#pragma unroll
                for (int j = 0; j < V; ++j) mono[j] = acc[0][j] + acc[1][j];
            } else {
                for (int c = 0; c < C; ++c) {
                    float acc[V];
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] = 0.f;
                    for (int q = s0; q < s1; ++q) {
                        const Vec<TS, V> x = load_vec<TS, V>(in + (long long)p.order[q] * p.stem_stride + c * n + i0);
#pragma unroll
                        for (int j = 0; j < V; ++j)
                            acc[j] = kI16 ? __fadd_rn(acc[j], __fmul_rn((float)x.v[j], scale)) : acc[j] + (float)x.v[j];
                    }
                    Vec<float, V> o;
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        o.v[j] = acc[j];
                        mono[j] = c == 0 ? acc[j] : mono[j] + acc[j];
                    }
                    store_vec<float, V>(outp + (long long)g * p.n_elems + c * n + i0, o);
                }
            }
            Vec<float, V> m;
#pragma unroll
            for (int j = 0; j < V; ++j) m.v[j] = C == 1 ? mono[j] : (inv_c != 0.f ? mono[j] * inv_c : mono[j] / f_c);
            store_vec<float, V>(monop + (long long)g * n + i0, m);
        }
    }
}

// PCM16 ingest form of the bus sum: int16 stems are converted with x * scale (1/32768 reproduces what
// librosa.load yields, exactly) and the buses are accumulated and written in float32, so 16-bit audio
// crosses PCIe and HBM at half the bytes.
template <int V>
__global__ void __launch_bounds__(kBlock) bus_sum_i16_kernel(const BusParams p, float scale) {
    const long long nvec = (p.n_elems + V - 1) / V;
    const short* in = reinterpret_cast<const short*>(p.stems) + (long long)blockIdx.y * p.batch_stride;
    float* outp = reinterpret_cast<float*>(p.out) + (long long)blockIdx.y * p.n_bus * p.n_elems;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const long long i0 = v * V;
        for (int g = 0; g < p.n_bus; ++g) {
            float acc[V];
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = 0.f;
            for (long long s = 0; s < p.n_stems; ++s) {
                if (p.group[s] != g) continue;
                const Vec<short, V> x = load_vec<short, V>(in + s * p.stem_stride + i0);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn((float)x.v[j], scale));   // never an FMA
            }
            if (V == 8) {
                reinterpret_cast<float4*>(outp + (long long)g * p.n_elems + i0)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                reinterpret_cast<float4*>(outp + (long long)g * p.n_elems + i0)[1] = make_float4(acc[4 % V], acc[5 % V], acc[6 % V], acc[7 % V]);
            } else {
                outp[(long long)g * p.n_elems + i0] = acc[0];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K3 feature-domain gain-weighted sum and gradients (float32)
constexpr int kMaxS = 8;

template <int V>
__global__ void __launch_bounds__(kBlock) feature_mix_fwd_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                                 long long S, long long P, float* __restrict__ out) {
    const long long b = blockIdx.y;
    const float* xb = x + b * S * P;
    float mb[kMaxS];
#pragma unroll
    for (int i = 0; i < kMaxS; ++i) mb[i] = i < S ? __ldg(m + b * S + i) : 0.f;
    const long long nvec = (P + V - 1) / V;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const long long i0 = v * V;
        Vec<float, V> acc;
#pragma unroll
        for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxS; ++i) {
            if (i < S) {
                const Vec<float, V> xv = load_vec<float, V>(xb + i * P + i0);
#pragma unroll
                for (int j = 0; j < V; ++j) acc.v[j] = __fadd_rn(acc.v[j], __fmul_rn(mb[i], xv.v[j]));
            }
        }
        store_vec<float, V>(out + b * P + i0, acc);
    }
}

template <int V>
__global__ void __launch_bounds__(kBlock) feature_mix_bwd_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                                                 const float* __restrict__ go, long long S, long long P,
                                                                 float* __restrict__ grad_m, float* __restrict__ grad_x) {
    __shared__ double s_red[kBlock / 32];
    const long long b = blockIdx.y;
    const float* xb = x + b * S * P;
    float mb[kMaxS];
    float part[kMaxS];
#pragma unroll
    for (int i = 0; i < kMaxS; ++i) { mb[i] = i < S ? __ldg(m + b * S + i) : 0.f; part[i] = 0.f; }
    double tot[kMaxS];
#pragma unroll
    for (int i = 0; i < kMaxS; ++i) tot[i] = 0.0;
    const long long nvec = (P + V - 1) / V;
    int since = 0;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const long long i0 = v * V;
        const Vec<float, V> g = load_vec<float, V>(go + b * P + i0);
#pragma unroll
        for (int i = 0; i < kMaxS; ++i) {
            if (i < S) {
                const Vec<float, V> xv = load_vec<float, V>(xb + i * P + i0);
                if (grad_x) {
                    Vec<float, V> gx;
#pragma unroll
                    for (int j = 0; j < V; ++j) gx.v[j] = mb[i] * g.v[j];
                    store_vec<float, V>(grad_x + (b * S + i) * P + i0, gx);
                }
#pragma unroll
                for (int j = 0; j < V; ++j) part[i] = fmaf(g.v[j], xv.v[j], part[i]);
            }
        }
        if (++since == 16) {       // flush fp32 partials into fp64 before they lose bits
#pragma unroll
            for (int i = 0; i < kMaxS; ++i) { tot[i] += part[i]; part[i] = 0.f; }
            since = 0;
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxS; ++i) {
        if (i < S) {
            const double t = block_sum(tot[i] + (double)part[i], s_red);
            if (threadIdx.x == 0) atomicAdd(grad_m + b * S + i, (float)t);
        }
    }
}

// K6 DummyModel: 20*log10(sum_i 10^(0.5 x_i))
template <int V>
__global__ void __launch_bounds__(kBlock) db_sum_kernel(const float* __restrict__ x, long long S, long long P,
                                                        float* __restrict__ out) {
    const long long b = blockIdx.y;
    const float* xb = x + b * S * P;
    const long long nvec = (P + V - 1) / V;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const long long i0 = v * V;
        Vec<float, V> acc;
#pragma unroll
        for (int j = 0; j < V; ++j) acc.v[j] = 0.f;
        for (long long i = 0; i < S; ++i) {
            const Vec<float, V> xv = load_vec<float, V>(xb + i * P + i0);
#pragma unroll
            for (int j = 0; j < V; ++j) acc.v[j] = __fadd_rn(acc.v[j], powf(10.0f, __fmul_rn(0.5f, xv.v[j])));
        }
#pragma unroll
        for (int j = 0; j < V; ++j) acc.v[j] = __fmul_rn(20.0f, log10f(acc.v[j]));
        store_vec<float, V>(out + b * P + i0, acc);
    }
}

// ------------------------------------------------------------------------------------------
// K5 row scale and row reductions
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(kBlock) scale_rows_kernel(const TI* __restrict__ x, long long n, long long row_stride,
                                                            const double* __restrict__ factor, int divide,
                                                            TO* __restrict__ out) {
    const long long r = blockIdx.y;
    const double f = __ldg(factor + r);
    const TI* xr = x + r * row_stride;
    TO* o = out + r * n;
    const long long nvec = (n + V - 1) / V;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const Vec<TI, V> xv = load_vec<TI, V>(xr + v * V);
        Vec<TO, V> y;
#pragma unroll
        for (int j = 0; j < V; ++j) y.v[j] = (TO)(divide ? (double)xv.v[j] / f : (double)xv.v[j] * f);
        store_vec<TO, V>(o + v * V, y);
    }
}

template <typename TI, int V>
__global__ void __launch_bounds__(kBlock) row_reduce_kernel(const TI* __restrict__ x, long long n, long long row_stride,
                                                            double* absmax, double* sumsq) {
    __shared__ double s_red[kBlock / 32];
    const long long r = blockIdx.y;
    const TI* xr = x + r * row_stride;
    double mx = 0.0, sq = 0.0;
    const long long nvec = n / V;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const Vec<TI, V> xv = load_vec<TI, V>(xr + v * V);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const double a = fabs((double)xv.v[j]);
            mx = fmax(mx, a);
            sq += a * a;
        }
    }
    for (long long i = nvec * V + (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const double a = fabs((double)__ldg(xr + i));
        mx = fmax(mx, a);
        sq += a * a;
    }
    if (absmax) {
        const double m = block_max(mx, s_red);
        if (threadIdx.x == 0) atomic_max_nonneg(absmax + r, m);
    }
    if (sumsq) {
        const double s = block_sum(sq, s_red);
        if (threadIdx.x == 0) atomicAdd(sumsq + r, s);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// ==========================================================================================
// Shared body of dam_gain_mix and dam_gain_mix_kweight (kw_coeffs_host == nullptr: plain mixdown).
static int gain_mix_entry(const void* stems, int stems_dtype, int64_t n_batch, int64_t batch_stride,
                          int64_t n_stems, int64_t n_channels,
                          int64_t n_samples, int64_t stem_stride, int64_t chan_stride,
                          const double* gains, int64_t n_gains, int64_t interp_coef, int hold_single,
                          int compute_dtype, void* mix_out, void* stems_out, int out_dtype,
                          double* peak, double* sumsq,
                          const double* kw_coeffs_host, void* kw_workspace, size_t kw_workspace_bytes, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(stems && gains, "null pointer argument");
    DAM_CHECK_ARG(mix_out || stems_out || peak || sumsq, "no output requested");
    DAM_CHECK_ARG(stems_dtype == DAM_F32 || stems_dtype == DAM_F64, "stems_dtype must be F32 or F64");
    DAM_CHECK_ARG(compute_dtype == DAM_F32 || compute_dtype == DAM_F64, "compute_dtype must be F32 or F64");
    DAM_CHECK_ARG(out_dtype == DAM_F32 || out_dtype == DAM_F64, "out_dtype must be F32 or F64");
    DAM_CHECK_ARG(!(stems_dtype == DAM_F64 && compute_dtype == DAM_F32), "F64 stems need F64 compute");
    DAM_CHECK_ARG(n_stems >= 1 && n_channels >= 1 && n_channels <= 65535 && n_samples >= 0, "bad sizes");
    DAM_CHECK_ARG(n_batch >= 0 && n_batch <= 65535, "n_batch must be in [0, 65535]");
    const bool want_kw = kw_coeffs_host != nullptr;
    if (want_kw) {
        DAM_CHECK_ARG(mix_out && kw_workspace, "the K-weighting states need mix_out and a workspace");
        DAM_CHECK_ARG(n_batch >= 1 && n_samples >= 1 && n_batch * n_channels <= 65535, "bad sizes for the K-weighting states");
        DAM_CHECK_ARG(kw_workspace_bytes >= dam_kweight_workspace_bytes(n_batch * n_channels, n_samples),
                      "workspace too small: need %zu bytes", dam_kweight_workspace_bytes(n_batch * n_channels, n_samples));
        DAM_CHECK_ARG((reinterpret_cast<uintptr_t>(kw_workspace) & 15u) == 0, "workspace must be 16-byte aligned");
    }
    if (n_batch == 0) return DAM_OK;
    DAM_CHECK_ARG(n_gains >= 1, "need at least one gain per stem");
    DAM_CHECK_ARG(n_gains <= n_samples || n_samples == 0, "Target mask should be longer than the initial one");
    DAM_CHECK_ARG(interp_coef >= 0, "interp_coef must be non-negative");
    if (n_samples == 0) return DAM_OK;
    GainMixParams p;
    p.stems = stems; p.n_stems = n_stems; p.n_channels = n_channels; p.n_samples = n_samples;
    p.n_batch = n_batch; p.batch_stride = batch_stride;
    p.stem_stride = stem_stride; p.chan_stride = chan_stride;
    p.gains = gains; p.n_gains = n_gains; p.interp_coef = interp_coef;
    p.zero_mask = (n_gains == 1 && !hold_single) ? 1 : 0;
    p.mix_out = mix_out; p.stems_out = stems_out; p.peak = peak; p.sumsq = sumsq;
    const size_t es = stems_dtype == DAM_F32 ? 4 : 8;
    bool vec_ok = (n_samples % 4 == 0) && (stem_stride % 4 == 0) && (chan_stride % 4 == 0) && (batch_stride % 4 == 0) &&
                  ((reinterpret_cast<uintptr_t>(stems) % (4 * es)) == 0) &&
                  (!mix_out || (reinterpret_cast<uintptr_t>(mix_out) % 32) == 0) &&
                  (!stems_out || (reinterpret_cast<uintptr_t>(stems_out) % 32) == 0);
    if (stems_dtype == DAM_F64) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(stems) % 32 == 0);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (want_kw && vec_ok) {
        KwFuseParams kw;
        const double* G = nullptr;
        const int rc = dam_kw_response_table(kw_coeffs_host, &G);
        if (rc != DAM_OK) return rc;
        memcpy(kw.G, G, sizeof(kw.G));
        kw.n_chunks = (n_samples + kKwChunk - 1) / kKwChunk;
        int parts;
        int64_t super;
        dam_kw_scan_partition(kw.n_chunks, &parts, &super);
        kw.lay.states = reinterpret_cast<double*>(kw_workspace);
        kw.lay.parts = parts; kw.lay.super = super; kw.lay.cpp = super * kKwScanBlock;
        DAM_CHECK_ARG(kw.n_chunks < (1LL << 31), "n_samples too large");
        kw.cpp_inv = kw.lay.cpp == 1 ? 0xffffffffu : (unsigned)((1ULL << 32) / (unsigned long long)kw.lay.cpp);
        kw.super_inv = kw.lay.super == 1 ? 0xffffffffu : (unsigned)((1ULL << 32) / (unsigned long long)kw.lay.super);
        if (stems_dtype == DAM_F32) {
            if (compute_dtype == DAM_F64)
                return out_dtype == DAM_F64 ? launch_gain_mix_kweight<float, double, double>(p, kw, st)
                                            : launch_gain_mix_kweight<float, double, float>(p, kw, st);
            return out_dtype == DAM_F64 ? launch_gain_mix_kweight<float, float, double>(p, kw, st)
                                        : launch_gain_mix_kweight<float, float, float>(p, kw, st);
        }
        return out_dtype == DAM_F64 ? launch_gain_mix_kweight<double, double, double>(p, kw, st)
                                    : launch_gain_mix_kweight<double, double, float>(p, kw, st);
    }
    int rc;
    if (stems_dtype == DAM_F32) {
        if (compute_dtype == DAM_F64)
            rc = out_dtype == DAM_F64 ? launch_gain_mix<float, double, double>(p, vec_ok, st)
                                      : launch_gain_mix<float, double, float>(p, vec_ok, st);
        else
            rc = out_dtype == DAM_F64 ? launch_gain_mix<float, float, double>(p, vec_ok, st)
                                      : launch_gain_mix<float, float, float>(p, vec_ok, st);
    } else {
        rc = out_dtype == DAM_F64 ? launch_gain_mix<double, double, double>(p, vec_ok, st)
                                  : launch_gain_mix<double, double, float>(p, vec_ok, st);
    }
    if (rc != DAM_OK || !want_kw) return rc;
    // a layout the fused kernel does not take (unaligned or n % 4 != 0): the states come from the stand-alone pass
    // over the mix that was just written -- the same bits, one more read
    return dam_kw_chunk_states(mix_out, out_dtype, n_batch * n_channels, n_samples, n_samples, kw_coeffs_host,
                               kw_workspace, kw_workspace_bytes, stream);
}

extern "C" int dam_gain_mix(const void* stems, int stems_dtype, int64_t n_batch, int64_t batch_stride,
                            int64_t n_stems, int64_t n_channels,
                            int64_t n_samples, int64_t stem_stride, int64_t chan_stride,
                            const double* gains, int64_t n_gains, int64_t interp_coef, int hold_single,
                            int compute_dtype, void* mix_out, void* stems_out, int out_dtype,
                            double* peak, double* sumsq, void* stream) {
    return gain_mix_entry(stems, stems_dtype, n_batch, batch_stride, n_stems, n_channels, n_samples, stem_stride, chan_stride,
                          gains, n_gains, interp_coef, hold_single, compute_dtype, mix_out, stems_out, out_dtype, peak, sumsq,
                          nullptr, nullptr, 0, stream);
}

extern "C" int dam_gain_mix_kweight(const void* stems, int stems_dtype, int64_t n_batch, int64_t batch_stride,
                                    int64_t n_stems, int64_t n_channels,
                                    int64_t n_samples, int64_t stem_stride, int64_t chan_stride,
                                    const double* gains, int64_t n_gains, int64_t interp_coef, int hold_single,
                                    int compute_dtype, void* mix_out, void* stems_out, int out_dtype,
                                    double* peak, double* sumsq,
                                    const double* kw_coeffs_host, void* kw_workspace, size_t kw_workspace_bytes,
                                    void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(kw_coeffs_host != nullptr, "null pointer argument");
    return gain_mix_entry(stems, stems_dtype, n_batch, batch_stride, n_stems, n_channels, n_samples, stem_stride, chan_stride,
                          gains, n_gains, interp_coef, hold_single, compute_dtype, mix_out, stems_out, out_dtype, peak, sumsq,
                          kw_coeffs_host, kw_workspace, kw_workspace_bytes, stream);
}

// N3: head outputs of one batch of chunks -> per-chunk amplitude gains in the layout K2 reads.
//   gains[b][s][c] = (double) powf(10, 0.5 * m[(b * n_chunks + c) * n_stems + s])
// i.e. data/dataset_utils.py:46-50 (10 ** (0.5 m) in float32) widened for the float64 mask -- what three elementwise
// launches (scale, power, widening transposition) did before.
__global__ void __launch_bounds__(kBlock) head_gains_kernel(const float* __restrict__ m, long long n_batch, long long n_chunks,
                                                            long long n_stems, double* __restrict__ out) {
    const long long total = n_batch * n_stems * n_chunks;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < total; i += (long long)gridDim.x * kBlock) {
        const long long c = i % n_chunks, bs = i / n_chunks;
        const long long s = bs % n_stems, b = bs / n_stems;
        out[i] = (double)powf(10.0f, __fmul_rn(0.5f, m[(b * n_chunks + c) * n_stems + s]));
    }
}

extern "C" int dam_head_gains(const float* m, int64_t n_batch, int64_t n_chunks, int64_t n_stems, double* gains_out,
                              void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(m && gains_out, "null pointer argument");
    DAM_CHECK_ARG(n_batch >= 0 && n_chunks >= 1 && n_stems >= 1, "bad sizes");
    if (n_batch == 0) return DAM_OK;
    { head_gains_kernel<<<grid_for(n_batch * n_chunks * n_stems, 4), kBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
          m, n_batch, n_chunks, n_stems, gains_out); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_bus_sum(const void* stems, int dtype, int64_t n_batch, int64_t batch_stride,
                           int64_t n_stems, int64_t n_elems, int64_t stem_stride,
                           const int32_t* group_host, int n_bus, double input_scale, void* out, void* mono_out,
                           int n_channels, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(stems && group_host && out, "null pointer argument");
    DAM_CHECK_ARG(dtype == DAM_F32 || dtype == DAM_F64 || dtype == DAM_I16, "dtype must be I16, F32 or F64");
    DAM_CHECK_ARG(n_stems >= 1 && n_stems <= 64, "n_stems must be in [1, 64]");
    DAM_CHECK_ARG(n_bus >= 1 && n_bus <= 64 && n_elems >= 0, "bad sizes");
    DAM_CHECK_ARG(n_batch >= 0 && n_batch <= 65535, "n_batch must be in [0, 65535]");
    if (n_elems == 0 || n_batch == 0) return DAM_OK;
    BusParams p;
    p.stems = stems; p.n_stems = n_stems; p.n_elems = n_elems; p.stem_stride = stem_stride;
    p.batch_stride = batch_stride;
    p.n_bus = n_bus; p.out = out;
    p.mono_out = mono_out; p.n_channels = n_channels;
    for (int i = 0; i < 64; ++i) p.group[i] = -1;
    for (int64_t s = 0; s < n_stems; ++s) {
        DAM_CHECK_ARG(group_host[s] >= -1 && group_host[s] < n_bus, "group[%lld]=%d out of range", (long long)s, group_host[s]);
        p.group[s] = group_host[s];
    }
    {
        int pos = 0;
        for (int g = 0; g < n_bus; ++g) {
            p.start[g] = (unsigned char)pos;
            for (int64_t s2 = 0; s2 < n_stems; ++s2)
                if (group_host[s2] == g) p.order[pos++] = (unsigned char)s2;
        }
        p.start[n_bus] = (unsigned char)pos;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (mono_out != nullptr) {
        DAM_CHECK_ARG(dtype == DAM_F32 || dtype == DAM_I16, "the mono output is implemented for F32 and I16 stems");
        DAM_CHECK_ARG(n_channels >= 1 && n_channels <= 64 && n_elems % n_channels == 0,
                      "n_channels=%d must divide n_elems=%lld", n_channels, (long long)n_elems);
        const long long n = n_elems / n_channels;
        const size_t es = dtype == DAM_F32 ? 4 : 2;
        const bool vec_ok = (n % 4 == 0) && (stem_stride % 4 == 0) && (batch_stride % 4 == 0) &&
                            (reinterpret_cast<uintptr_t>(stems) % (4 * es) == 0) &&
                            (reinterpret_cast<uintptr_t>(out) % 16 == 0) && (reinterpret_cast<uintptr_t>(mono_out) % 16 == 0);
        const float scale = (float)input_scale;
        const dim3 grid(grid_for(vec_ok ? n / 4 : n, n_batch >= 8 ? 1 : 8), (unsigned)n_batch);
        if (dtype == DAM_F32) {
            if (vec_ok) { bus_sum_mono_kernel<4, false><<<grid, kBlock, 0, st>>>(p, 1.0f); dam_note_launch(); }
            else { bus_sum_mono_kernel<1, false><<<grid, kBlock, 0, st>>>(p, 1.0f); dam_note_launch(); }
        } else {
            if (vec_ok) { bus_sum_mono_kernel<4, true><<<grid, kBlock, 0, st>>>(p, scale); dam_note_launch(); }
            else { bus_sum_mono_kernel<1, true><<<grid, kBlock, 0, st>>>(p, scale); dam_note_launch(); }
        }
        DAM_CHECK_CUDA(cudaGetLastError());
        return DAM_OK;
    }
    if (dtype == DAM_I16) {
        const bool v8 = (n_elems % 8 == 0) && (stem_stride % 8 == 0) && (batch_stride % 8 == 0) &&
                        (reinterpret_cast<uintptr_t>(stems) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
        const float scale = (float)input_scale;
        if (v8) { bus_sum_i16_kernel<8><<<dim3(grid_for((n_elems + 7) / 8, n_batch >= 8 ? 1 : 8), (unsigned)n_batch), kBlock, 0, st>>>(p, scale); dam_note_launch(); }
        else { bus_sum_i16_kernel<1><<<dim3(grid_for(n_elems, n_batch >= 8 ? 1 : 8), (unsigned)n_batch), kBlock, 0, st>>>(p, scale); dam_note_launch(); }
        DAM_CHECK_CUDA(cudaGetLastError());
        return DAM_OK;
    }
    const size_t es = dtype == DAM_F32 ? 4 : 8;
    const bool vec_ok = (n_elems % 4 == 0) && (stem_stride % 4 == 0) && (batch_stride % 4 == 0) &&
                        (reinterpret_cast<uintptr_t>(stems) % (4 * es) == 0) &&
                        (reinterpret_cast<uintptr_t>(out) % (4 * es) == 0);
    if (dtype == DAM_F32) {
        if (vec_ok) { bus_sum_kernel<float, 4><<<dim3(grid_for((n_elems + 3) / 4, n_batch >= 8 ? 1 : 8), (unsigned)n_batch), kBlock, 0, st>>>(p); dam_note_launch(); }
        else { bus_sum_kernel<float, 1><<<dim3(grid_for(n_elems, n_batch >= 8 ? 1 : 8), (unsigned)n_batch), kBlock, 0, st>>>(p); dam_note_launch(); }
    } else {
        if (vec_ok) { bus_sum_kernel<double, 4><<<dim3(grid_for((n_elems + 3) / 4, n_batch >= 8 ? 1 : 8), (unsigned)n_batch), kBlock, 0, st>>>(p); dam_note_launch(); }
        else { bus_sum_kernel<double, 1><<<dim3(grid_for(n_elems, n_batch >= 8 ? 1 : 8), (unsigned)n_batch), kBlock, 0, st>>>(p); dam_note_launch(); }
    }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

static int check_bsp(int64_t B, int64_t S, int64_t P) {
    DAM_CHECK_ARG(B >= 0 && B <= 65535 && P >= 0, "bad sizes");
    DAM_CHECK_ARG(S >= 1 && S <= kMaxS, "S must be in [1, %d]", kMaxS);
    return DAM_OK;
}

extern "C" int dam_feature_mix_fwd(const float* x, const float* m, int64_t B, int64_t S, int64_t P,
                                   float* out, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && m && out, "null pointer argument");
    int rc = check_bsp(B, S, P);
    if (rc != DAM_OK) return rc;
    if (B == 0 || P == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec_ok = (P % 4 == 0) && aligned16(x) && aligned16(out);
    if (vec_ok) {
        dim3 grid(grid_for((P + 3) / 4, 2), (unsigned)B);
        { feature_mix_fwd_kernel<4><<<grid, kBlock, 0, st>>>(x, m, S, P, out); dam_note_launch(); }
    } else {
        dim3 grid(grid_for(P, 2), (unsigned)B);
        { feature_mix_fwd_kernel<1><<<grid, kBlock, 0, st>>>(x, m, S, P, out); dam_note_launch(); }
    }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_feature_mix_bwd(const float* x, const float* m, const float* grad_out,
                                   int64_t B, int64_t S, int64_t P,
                                   float* grad_m, float* grad_x, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && m && grad_out && grad_m, "null pointer argument");
    int rc = check_bsp(B, S, P);
    if (rc != DAM_OK) return rc;
    if (B == 0 || P == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec_ok = (P % 4 == 0) && aligned16(x) && aligned16(grad_out) && (!grad_x || aligned16(grad_x));
    if (vec_ok) {
        dim3 grid(grid_for((P + 3) / 4, 2) > 64 ? 64 : grid_for((P + 3) / 4, 2), (unsigned)B);
        { feature_mix_bwd_kernel<4><<<grid, kBlock, 0, st>>>(x, m, grad_out, S, P, grad_m, grad_x); dam_note_launch(); }
    } else {
        dim3 grid(grid_for(P, 2) > 64 ? 64 : grid_for(P, 2), (unsigned)B);
        { feature_mix_bwd_kernel<1><<<grid, kBlock, 0, st>>>(x, m, grad_out, S, P, grad_m, grad_x); dam_note_launch(); }
    }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_db_sum(const float* x, int64_t B, int64_t S, int64_t P, float* out, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && out, "null pointer argument");
    DAM_CHECK_ARG(B >= 0 && B <= 65535 && P >= 0 && S >= 1, "bad sizes");
    if (B == 0 || P == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec_ok = (P % 4 == 0) && aligned16(x) && aligned16(out);
    if (vec_ok) {
        dim3 grid(grid_for((P + 3) / 4, 2), (unsigned)B);
        { db_sum_kernel<4><<<grid, kBlock, 0, st>>>(x, S, P, out); dam_note_launch(); }
    } else {
        dim3 grid(grid_for(P, 2), (unsigned)B);
        { db_sum_kernel<1><<<grid, kBlock, 0, st>>>(x, S, P, out); dam_note_launch(); }
    }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_scale_rows(const void* x, int dtype, int64_t n_rows, int64_t n, int64_t row_stride,
                              const double* factor, int divide, void* out, int out_dtype, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && factor && out, "null pointer argument");
    DAM_CHECK_ARG((dtype == DAM_F32 || dtype == DAM_F64) && (out_dtype == DAM_F32 || out_dtype == DAM_F64),
                  "dtypes must be F32 or F64");
    DAM_CHECK_ARG(n_rows >= 0 && n_rows <= 65535 && n >= 0, "bad sizes");
    if (n_rows == 0 || n == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec_ok = (n % 4 == 0) && (row_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 32 == 0) &&
                        (reinterpret_cast<uintptr_t>(out) % 32 == 0);
    dim3 grid(grid_for(vec_ok ? (n + 3) / 4 : n, 4), (unsigned)n_rows);
#define DAM_SCALE(TI, TO)                                                                                         \
    do {                                                                                                          \
        if (vec_ok) scale_rows_kernel<TI, TO, 4><<<grid, kBlock, 0, st>>>((const TI*)x, n, row_stride, factor, divide, (TO*)out); \
        else scale_rows_kernel<TI, TO, 1><<<grid, kBlock, 0, st>>>((const TI*)x, n, row_stride, factor, divide, (TO*)out);        \
        dam_note_launch();                                                                                                        \
    } while (0)
    if (dtype == DAM_F32 && out_dtype == DAM_F32) DAM_SCALE(float, float);
    else if (dtype == DAM_F32) DAM_SCALE(float, double);
    else if (out_dtype == DAM_F32) DAM_SCALE(double, float);
    else DAM_SCALE(double, double);
#undef DAM_SCALE
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_row_reduce(const void* x, int dtype, int64_t n_rows, int64_t n, int64_t row_stride,
                              double* absmax, double* sumsq, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && (absmax || sumsq), "null pointer argument");
    DAM_CHECK_ARG(dtype == DAM_F32 || dtype == DAM_F64, "dtype must be F32 or F64");
    DAM_CHECK_ARG(n_rows >= 0 && n_rows <= 65535 && n >= 0, "bad sizes");
    if (n_rows == 0 || n == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec_ok = (row_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 32 == 0);
    dim3 grid(grid_for(vec_ok ? (n + 3) / 4 : n, 4), (unsigned)n_rows);
    if (dtype == DAM_F32) {
        if (vec_ok) { row_reduce_kernel<float, 4><<<grid, kBlock, 0, st>>>((const float*)x, n, row_stride, absmax, sumsq); dam_note_launch(); }
        else { row_reduce_kernel<float, 1><<<grid, kBlock, 0, st>>>((const float*)x, n, row_stride, absmax, sumsq); dam_note_launch(); }
    } else {
        if (vec_ok) { row_reduce_kernel<double, 4><<<grid, kBlock, 0, st>>>((const double*)x, n, row_stride, absmax, sumsq); dam_note_launch(); }
        else { row_reduce_kernel<double, 1><<<grid, kBlock, 0, st>>>((const double*)x, n, row_stride, absmax, sumsq); dam_note_launch(); }
    }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

// dam_scale: one signal times a HOST scalar: pyln.normalize.loudness with a gain the caller already
// holds, MeanLoudnessModel / RandomModel per-stem gains, peak normalisation with a known peak.
namespace {
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(kBlock) scale_scalar_kernel(const TI* __restrict__ x, long long n, double gain,
                                                              TO* __restrict__ out) {
    const long long nvec = (n + V - 1) / V;
    for (long long v = (long long)blockIdx.x * kBlock + threadIdx.x; v < nvec; v += (long long)gridDim.x * kBlock) {
        const Vec<TI, V> xv = load_vec<TI, V>(x + v * V);
        Vec<TO, V> y;
#pragma unroll
        for (int j = 0; j < V; ++j) y.v[j] = (TO)((double)xv.v[j] * gain);
        store_vec<TO, V>(out + v * V, y);
    }
}

// ------------------------------------------------------------------------------------------
// N1: packed little-endian PCM bytes -> float samples (x * scale), element for element (the channel
// interleave is kept; K1 / the bus sum take interleaved rows).  24-bit is the case no tensor dtype covers.
template <int kBytes, typename TO>
__global__ void __launch_bounds__(kBlock) pcm_unpack_kernel(const unsigned char* __restrict__ src, long long n, double scale,
                                                            TO* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const unsigned char* b = src + i * kBytes;
        int v;
        if (kBytes == 2) v = (int)(short)(b[0] | (b[1] << 8));
        else if (kBytes == 3) v = ((int)(b[0] | (b[1] << 8) | (b[2] << 16)) << 8) >> 8;      // sign-extend 24 bits
        else v = (int)(b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24));
        out[i] = (TO)((double)v * scale);
    }
}

}  // namespace

extern "C" int dam_pcm_unpack(const void* src, int bits, int64_t n, double scale, void* out, int out_dtype, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(src && out, "null pointer argument");
    DAM_CHECK_ARG(bits == 16 || bits == 24 || bits == 32, "bits must be 16, 24 or 32");
    DAM_CHECK_ARG(out_dtype == DAM_F32 || out_dtype == DAM_F64, "out_dtype must be F32 or F64");
    DAM_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(n, 8);
    const unsigned char* s = reinterpret_cast<const unsigned char*>(src);
#define DAM_UNPACK(B)                                                                                        \
    do {                                                                                                     \
        if (out_dtype == DAM_F32) pcm_unpack_kernel<B, float><<<grid, kBlock, 0, st>>>(s, n, scale, (float*)out);   \
        else pcm_unpack_kernel<B, double><<<grid, kBlock, 0, st>>>(s, n, scale, (double*)out);                      \
        dam_note_launch();                                                                                   \
    } while (0)
    if (bits == 16) DAM_UNPACK(2);
    else if (bits == 24) DAM_UNPACK(3);
    else DAM_UNPACK(4);
#undef DAM_UNPACK
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_scale(const void* x, int dtype, int64_t n, double gain, void* out, int out_dtype, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && out, "null pointer argument");
    DAM_CHECK_ARG((dtype == DAM_F32 || dtype == DAM_F64) && (out_dtype == DAM_F32 || out_dtype == DAM_F64),
                  "dtypes must be F32 or F64");
    DAM_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return DAM_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec_ok = (n % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 32 == 0) && (reinterpret_cast<uintptr_t>(out) % 32 == 0);
    const int grid = grid_for(vec_ok ? n / 4 : n, 8);
#define DAM_SCALE1(TI, TO)                                                                              \
    do {                                                                                               \
        if (vec_ok) scale_scalar_kernel<TI, TO, 4><<<grid, kBlock, 0, st>>>((const TI*)x, n, gain, (TO*)out); \
        else scale_scalar_kernel<TI, TO, 1><<<grid, kBlock, 0, st>>>((const TI*)x, n, gain, (TO*)out);        \
        dam_note_launch();                                                                                    \
    } while (0)
    if (dtype == DAM_F32 && out_dtype == DAM_F32) DAM_SCALE1(float, float);
    else if (dtype == DAM_F32) DAM_SCALE1(float, double);
    else if (out_dtype == DAM_F32) DAM_SCALE1(double, float);
    else DAM_SCALE1(double, double);
#undef DAM_SCALE1
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}
