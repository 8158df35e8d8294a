// K4: ITU-R BS.1770 K-weighting (two cascaded biquads, float64) and mean-square partial sums over
// caller-defined sample segments, for many signals at once.
//
// Replaces the inside of pyloudnorm.Meter.integrated_loudness as the reference calls it
// (data/dataset.py:118,126; evaluation.py:32,40,63; models/baselines/mean_loudness_model.py:7,14;
// data/listening_test_data_preparation.py:16,23): scipy.signal.lfilter x2 per channel followed by a
// Python loop over 400 ms blocks.  Gating and the final log are done by the caller on the tiny
// per-segment array.
//
// The IIR recurrence is serial in time, so the signal is cut into chunks of kChunk samples:
//   pass A  (parallel)  zero-state end state of every chunk = sum_i G[:, i] * x[i]  (a dot product
//                       with the precomputed decaying response table G; no serial dependency)
//   scan    (tiny)      true start state of every chunk: s_{c+1} = M^L s_c + z_c, two-level
//   pass C  (parallel)  every chunk re-runs the exact DF2T recurrence from its true start state,
//                       squares the output and adds it to the segment sums
// This is the exact state hand-off (no warm-up approximation); all filter arithmetic is float64.
#include "dam_common.cuh"
#include "kweight_common.cuh"

#include <math.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <mutex>
#include <string.h>
#include <vector>

namespace {

constexpr int kChunk = kKwChunk;    // samples per chunk (L)
constexpr int kBlock = 256;        // chunks (= threads) per CTA in passes A and C
constexpr int kSub = 16;           // samples per chunk staged per pipeline step
constexpr int kPitch = kSub + 1;   // shared-memory row pitch: thread c walks row c conflict-free
constexpr int kScanBlock = kKwScanBlock;   // threads per row in the scan (long rows: 41 k chunks per 4-minute stem)
constexpr int kScanLevels = 10;     // log2(kScanBlock) Kogge-Stone levels

struct Coeffs { double b0[2], b1[2], b2[2], a1[2], a2[2]; };

struct KwTables {                  // device-resident, per coefficient set
    Coeffs c;
    double G[kChunk * 4];          // response table (host; copied into the kernel parameters)
    double ML[16];                 // M^L   (row-major 4x4)
};

__host__ __device__ inline void step(const Coeffs& c, double x, double s[4], double* y_out) {
    const double y1 = fma(c.b0[0], x, s[0]);
    s[0] = fma(-c.a1[0], y1, fma(c.b1[0], x, s[1]));
    s[1] = fma(-c.a2[0], y1, c.b2[0] * x);
    const double y2 = fma(c.b0[1], y1, s[2]);
    s[2] = fma(-c.a1[1], y2, fma(c.b1[1], y1, s[3]));
    s[3] = fma(-c.a2[1], y2, c.b2[1] * y1);
    if (y_out) *y_out = y2;
}

void matmul4(const double* a, const double* b, double* out) {
    double t[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0;
            for (int k = 0; k < 4; ++k) acc += a[i * 4 + k] * b[k * 4 + j];
            t[i * 4 + j] = acc;
        }
    memcpy(out, t, sizeof(t));
}

int get_tables(const Coeffs& c, KwTables** out) {
    static std::mutex mu;
    static std::vector<KwTables*> cache[64];
    int dev = 0;
    DAM_CHECK_CUDA(cudaGetDevice(&dev));
    DAM_CHECK_ARG(dev >= 0 && dev < 64, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    for (KwTables* t : cache[dev])
        if (memcmp(&t->c, &c, sizeof(Coeffs)) == 0) { *out = t; return DAM_OK; }
    KwTables* t = new KwTables();
    t->c = c;
    // one-step zero-input matrix M (columns = images of unit states)
    double M[16];
    for (int j = 0; j < 4; ++j) {
        double s[4] = {0, 0, 0, 0};
        s[j] = 1.0;
        step(c, 0.0, s, nullptr);
        for (int i = 0; i < 4; ++i) M[i * 4 + j] = s[i];
    }
    // G[:, L-1-k] = M^k * B, B = state after a unit impulse from rest
    double* G = t->G;
    {
        double v[4] = {0, 0, 0, 0};
        step(c, 1.0, v, nullptr);
        for (int k = 0; k < kChunk; ++k) {
            for (int i = 0; i < 4; ++i) G[(kChunk - 1 - k) * 4 + i] = v[i];
            step(c, 0.0, v, nullptr);
        }
    }
    // M^L by running L zero-input steps on unit states
    for (int j = 0; j < 4; ++j) {
        double s[4] = {0, 0, 0, 0};
        s[j] = 1.0;
        for (int k = 0; k < kChunk; ++k) step(c, 0.0, s, nullptr);
        for (int i = 0; i < 4; ++i) t->ML[i * 4 + j] = s[i];
    }
    (void)M;
    cache[dev].push_back(t);
    *out = t;
    return DAM_OK;
}

struct KwParams {
    const void* x;
    long long n_rows, n, row_stride, samp_stride;
    long long n_chunks;            // ceil(n / kChunk)
    long long n_items;             // ceil(n_rows * n_chunks / kBlock): work items of passes A and C
    Coeffs c;
    double G[kChunk * 4];          // response table, IN the kernel parameters: pass A's multiplicands are constant-bank
                                   // operands of its DFMAs (as shared-memory loads they kept the LSU pipe 84 % busy)
    double ML[16];                 // M^L (one chunk)
    double P[kScanLevels][16];     // (M^(L*super))^(2^j): Kogge-Stone levels of the scan
    long long super;               // chunks per scan thread
    long long cpp;                 // chunks per part = super * kScanBlock; a row is scanned by `parts` CTAs of one cluster
    int parts;
    double Q[16];                  // (M^L)^cpp: carries a start state across one whole part
    const long long* seg_bounds;   // [n_seg + 1]
    long long n_seg;
    double* seg_sumsq;             // [n_rows][n_seg]
    double* states;                // workspace [n_rows][parts * cpp][4] (transposed per part, see state_ptr)
    double* partials;              // workspace [n_rows * n_chunks][2]: pass C's per-chunk sums for the (at most two) segments it touches
    int zero_seg;                  // the scan clears seg_sumsq first (dam_kweight_segments_resume OVERWRITES its output)
};

// Filter state of chunk c of a row (layout: kweight_common.cuh).
__device__ __forceinline__ double* state_ptr(const KwParams& p, long long row, long long chunk) {
    const KwStateLayout lay = {p.states, p.super, p.cpp, p.parts};
    return kw_state_ptr(lay, row, chunk);
}

// Coalesced, software-pipelined, WARP-PRIVATE tile loader shared by passes A and C.  The unit of work
// is kBlock consecutive GLOBAL chunk ids (row-major over all rows, so a work item may straddle two
// rows and no CTA is padded per row: 220 544 chunks of the bench batch are 862 items, 2.91 rounds of
// 296 resident CTAs, instead of 896 per-row items = 4 rounds).  Lane l of warp w owns chunk 32 w + l of
// the item and publishes its base offset and length to the warp's slice of shared memory; per step the
// warp stages [32 chunks x kSub samples] into ITS OWN double-buffered tile: one load instruction
// reads two rows of 16 consecutive samples (full sectors), registers hold the NEXT step's samples
// while the current ones are consumed from shared memory, and lane l then walks row l (pitch
// kSub+1: conflict-free).  Only __syncwarp() separates the steps: with a CTA-wide tile and a
// __syncthreads() per 16 samples, 41 % of pass C's warp stall samples were barrier waits (profiles/).
// Thread-per-chunk loads straight from global memory touch 32 different lines per instruction and
// are bound by the L1 tag rate (measured: 70 % l1tex, 10 % issue).
template <typename T>
struct TileLoader {
    T reg[kSub];
    // w_base / w_len / tile: this warp's slices
    __device__ __forceinline__ void load(const T* __restrict__ x, const long long* __restrict__ w_base,
                                         const int* __restrict__ w_len, int sb, long long samp_stride, int lane) {
#pragma unroll
        for (int k = 0; k < kSub; ++k) {
            const int e = lane + k * 32;
            const int r = e >> 4, idx = sb * kSub + (e & 15);
            reg[k] = idx < w_len[r] ? x[w_base[r] + idx * samp_stride] : T(0);
        }
    }
    // The warp's 32 chunks are full and consecutive in one unit-stride row (all but the ~1/54 warps that
    // straddle a row end): chunk r starts at q0 + r * kChunk, so the 16 loads of a step are one base pointer
    // plus compile-time offsets -- no per-element base / length look-ups (they were 6 of the 40 instructions
    // per sample that made pass C issue bound).
    __device__ __forceinline__ void load_fast(const T* __restrict__ q0, int sb, int lane) {
        const T* q = q0 + (lane >> 4) * kChunk + (lane & 15) + sb * kSub;
#pragma unroll
        for (int k = 0; k < kSub; ++k) reg[k] = q[k * 2 * kChunk];
    }
    __device__ __forceinline__ void store(T* __restrict__ tile, int lane) const {
#pragma unroll
        for (int k = 0; k < kSub; ++k) {
            const int e = lane + k * 32;
            tile[(e >> 4) * kPitch + (e & 15)] = reg[k];
        }
    }
};

template <typename T>
constexpr size_t kw_smem_bytes() {
    return kBlock * sizeof(long long) + kBlock * sizeof(int) + 2 * kBlock * kPitch * sizeof(T);
}

struct ChunkId { long long g, row, chunk; int len; bool active; };

// thread c's chunk of work item `item`, published to shared memory for its warp's loader
__device__ __forceinline__ ChunkId publish_chunk(const KwParams& p, long long item, long long* s_base, int* s_len) {
    ChunkId id;
    id.g = item * kBlock + threadIdx.x;
    id.active = id.g < p.n_rows * p.n_chunks;
    id.row = id.active ? id.g / p.n_chunks : 0;
    id.chunk = id.g - id.row * p.n_chunks;
    const long long left = p.n - id.chunk * kChunk;
    id.len = !id.active ? 0 : (left < kChunk ? (int)left : kChunk);
    s_base[threadIdx.x] = id.row * p.row_stride + id.chunk * kChunk * p.samp_stride;
    s_len[threadIdx.x] = id.len;
    return id;
}

// warp-uniform: are this warp's 32 chunks full, consecutive and unit-stride?  q0 = first sample of the warp's first chunk
template <typename T>
__device__ __forceinline__ bool warp_fast_path(const KwParams& p, const ChunkId& id, const T* x, const T** q0) {
    const long long row0 = __shfl_sync(0xffffffffu, id.row, 0);
    const long long chunk0 = __shfl_sync(0xffffffffu, id.chunk, 0);
    const bool ok = id.active && id.len == kChunk && id.row == row0 && p.samp_stride == 1;
    *q0 = x + row0 * p.row_stride + chunk0 * kChunk;
    return __all_sync(0xffffffffu, ok);
}

// pass A: zero-state end state of each full chunk (the last chunk's end state is never needed)
template <typename T>
__global__ void __launch_bounds__(kBlock) kw_chunk_state_kernel(const KwParams p) {
    extern __shared__ __align__(16) unsigned char kw_smem[];
    long long* s_base = reinterpret_cast<long long*>(kw_smem);
    int* s_len = reinterpret_cast<int*>(s_base + kBlock);
    T* s_tile = reinterpret_cast<T*>(s_len + kBlock);
    const T* x = reinterpret_cast<const T*>(p.x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long* w_base = s_base + warp * 32;
    const int* w_len = s_len + warp * 32;
    T* w_tile = s_tile + warp * 32 * kPitch;               // [2][32][kPitch] with buffer stride kBlock * kPitch
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        __syncwarp();                         // the warp has consumed its previous item
        const ChunkId id = publish_chunk(p, item, s_base, s_len);
        __syncwarp();
        TileLoader<T> ld;
        const T* q0;
        const bool fast = warp_fast_path<T>(p, id, x, &q0);
        if (fast) ld.load_fast(q0, 0, lane); else ld.load(x, w_base, w_len, 0, p.samp_stride, lane);
        ld.store(w_tile, lane);
        __syncwarp();
        // Fully unrolled over the chunk: every index into p.G is a compile-time constant, so the table entries are
        // constant-bank operands of the DFMAs.  Summation order of kweight_common.cuh (groups of 4 samples, pairwise
        // over the groups): lvl[b] holds the finished subtree of 16 << b groups waiting for its right-hand sibling.
        double lvl[4][4];
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int sb = 0; sb < kChunk / kSub; ++sb) {
            const bool more = sb + 1 < kChunk / kSub;
            if (more) { if (fast) ld.load_fast(q0, sb + 1, lane); else ld.load(x, w_base, w_len, sb + 1, p.samp_stride, lane); }
            const T* mine = w_tile + (sb & 1) * kBlock * kPitch + lane * kPitch;
            double S[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                kw_group_sum(p.G + (sb * kSub + q * 4) * 4, (double)mine[q * 4], (double)mine[q * 4 + 1],
                             (double)mine[q * 4 + 2], (double)mine[q * 4 + 3], S[q]);
            double carry[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) carry[c] = __dadd_rn(__dadd_rn(S[0][c], S[1][c]), __dadd_rn(S[2][c], S[3][c]));
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if ((sb >> b) & 1) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) carry[c] = __dadd_rn(lvl[b][c], carry[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) lvl[b][c] = carry[c];
                    break;
                }
            }
            if (!more) { a0 = carry[0]; a1 = carry[1]; a2 = carry[2]; a3 = carry[3]; }
            if (more) ld.store(w_tile + ((sb + 1) & 1) * kBlock * kPitch, lane);
            __syncwarp();
        }
        if (id.active) {
            const bool last = id.chunk == p.n_chunks - 1;      // possibly partial: its end state is unused
            double* st = state_ptr(p, id.row, id.chunk);
            *reinterpret_cast<double2*>(st) = last ? make_double2(0.0, 0.0) : make_double2(a0, a1);
            *reinterpret_cast<double2*>(st + 2) = last ? make_double2(0.0, 0.0) : make_double2(a2, a3);
        }
    }
}

__device__ __forceinline__ void matvec_add(const double* M, const double s[4], const double z[4], double out[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        out[i] = fma(M[i * 4 + 0], s[0], fma(M[i * 4 + 1], s[1], fma(M[i * 4 + 2], s[2], fma(M[i * 4 + 3], s[3], z[i]))));
}

// scan: turn zero-state end states into true start states, in place.  One CTA per row; thread u owns
// a run of p.super (= ceil(n_chunks / kScanBlock)) consecutive chunks.
//   1. fold the run from a zero start state (loads are independent of the recurrence),
//   2. Kogge-Stone scan of the affine maps s -> P s + e across the CTA: level j combines with the
//      thread 2^j to the left using P_j = (M^(L*super))^(2^j), precomputed on the host,
//   3. replay the run from its true start state, storing the start state of every chunk.
__device__ __forceinline__ void scan_levels(const KwParams& p, double (&s)[4], double (*s_e)[4], long long u) {
    // inclusive Kogge-Stone scan: after level j thread u holds the end state of runs (u - 2^(j+1), u]
    s_e[u][0] = s[0]; s_e[u][1] = s[1]; s_e[u][2] = s[2]; s_e[u][3] = s[3];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kScanLevels; ++j) {
        const int d = 1 << j;
        double l[4] = {0, 0, 0, 0};
        if (u >= d) { l[0] = s_e[u - d][0]; l[1] = s_e[u - d][1]; l[2] = s_e[u - d][2]; l[3] = s_e[u - d][3]; }
        __syncthreads();                       // everyone has read its left neighbour of this level
        if (u >= d) {
            double o[4];
            matvec_add(p.P[j], l, s, o);
            s[0] = o[0]; s[1] = o[1]; s[2] = o[2]; s[3] = o[3];
        }
        s_e[u][0] = s[0]; s_e[u][1] = s[1]; s_e[u][2] = s[2]; s_e[u][3] = s[3];
        __syncthreads();
    }
}

// One CLUSTER of p.parts CTAs per row (one CTA for short rows): a single CTA scanning a 4-minute stem was bound by the
// memory bandwidth of its one SM (47 % long-scoreboard + 34 % LG-throttle stalls, 8 SMs busy for 8 rows).  Every CTA
// scans its part from a zero start state, the parts' end states are exchanged through distributed shared memory,
// every CTA folds the ones in front of it into its true start state and scans again with that seed.
__global__ void __launch_bounds__(kScanBlock) kw_scan_kernel(const KwParams p) {
    namespace cg = cooperative_groups;
    __shared__ double s_e[kScanBlock][4];
    __shared__ double s_E[4];                  // end state of this CTA's part from a zero start (read by the cluster)
    __shared__ double s_S[4];                  // true start state of this CTA's part
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned part = p.parts > 1 ? cluster.block_rank() : 0;
    const long long row = blockIdx.x / p.parts;
    if (p.zero_seg && part == 0)
        for (long long k = threadIdx.x; k < p.n_seg; k += kScanBlock) p.seg_sumsq[row * p.n_seg + k] = 0.0;
    double* st = p.states + ((row * p.parts + part) * p.cpp + threadIdx.x) * 4;   // slot of step k: st + k * kScanBlock * 4
    const long long u = threadIdx.x;
    const long long c0 = (long long)part * p.cpp + u * p.super;
    long long c1 = c0 + p.super;
    if (c1 > p.n_chunks) c1 = p.n_chunks;
    const long long steps = c1 > c0 ? c1 - c0 : 0;
    double e[4] = {0, 0, 0, 0};                // end state of this thread's run from a zero start
#pragma unroll 4
    for (long long k = 0; k < steps; ++k) {
        const double2 za = *reinterpret_cast<const double2*>(st + k * kScanBlock * 4);
        const double2 zb = *reinterpret_cast<const double2*>(st + k * kScanBlock * 4 + 2);
        double z[4] = {za.x, za.y, zb.x, zb.y}, o[4];
        matvec_add(p.ML, e, z, o);
        e[0] = o[0]; e[1] = o[1]; e[2] = o[2]; e[3] = o[3];
    }
    double s[4] = {e[0], e[1], e[2], e[3]};
    if (p.parts > 1) {
        scan_levels(p, s, s_e, u);
        if (u == kScanBlock - 1) { s_E[0] = s[0]; s_E[1] = s[1]; s_E[2] = s[2]; s_E[3] = s[3]; }
        cluster.sync();                        // every part's end state is published
        if (u == 0) {
            double S[4] = {0, 0, 0, 0};
            for (unsigned i = 0; i < part; ++i) {          // parts in front of this one are full: Q carries S across each
                const double* r = cluster.map_shared_rank(s_E, i);
                const double z[4] = {r[0], r[1], r[2], r[3]};
                double o[4];
                matvec_add(p.Q, S, z, o);
                S[0] = o[0]; S[1] = o[1]; S[2] = o[2]; S[3] = o[3];
            }
            s_S[0] = S[0]; s_S[1] = S[1]; s_S[2] = S[2]; s_S[3] = S[3];
        }
        cluster.sync();                        // remote reads done before any CTA may exit; s_S visible to the CTA
        // second scan, seeded: thread 0's run now starts from S
        s[0] = e[0]; s[1] = e[1]; s[2] = e[2]; s[3] = e[3];
        if (u == 0) {
            const double S[4] = {s_S[0], s_S[1], s_S[2], s_S[3]};
            double o[4];
            matvec_add(p.P[0], S, e, o);
            s[0] = o[0]; s[1] = o[1]; s[2] = o[2]; s[3] = o[3];
        }
    } else if (u == 0) {
        s_S[0] = s_S[1] = s_S[2] = s_S[3] = 0.0;
    }
    scan_levels(p, s, s_e, u);
    // start state of run u = inclusive result of run u-1 (the part's start state for run 0)
    if (u == 0) { s[0] = s_S[0]; s[1] = s_S[1]; s[2] = s_S[2]; s[3] = s_S[3]; }
    else { s[0] = s_e[u - 1][0]; s[1] = s_e[u - 1][1]; s[2] = s_e[u - 1][2]; s[3] = s_e[u - 1][3]; }
    // The zero-state end states of four steps are loaded BEFORE the first of their slots is overwritten with a start
    // state: written as load / store pairs per step the compiler must assume the store aliases the next load.
    for (long long k0 = 0; k0 < steps; k0 += 4) {
        double2 za[4], zb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long k = k0 + j < steps ? k0 + j : steps - 1;
            za[j] = *reinterpret_cast<const double2*>(st + k * kScanBlock * 4);
            zb[j] = *reinterpret_cast<const double2*>(st + k * kScanBlock * 4 + 2);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k0 + j < steps) {
                double* q = st + (k0 + j) * kScanBlock * 4;
                double z[4] = {za[j].x, za[j].y, zb[j].x, zb[j].y}, o[4];
                *reinterpret_cast<double2*>(q) = make_double2(s[0], s[1]);
                *reinterpret_cast<double2*>(q + 2) = make_double2(s[2], s[3]);
                matvec_add(p.ML, s, z, o);
                s[0] = o[0]; s[1] = o[1]; s[2] = o[2]; s[3] = o[3];
            }
        }
    }
}

// pass C: exact recurrence per chunk from its true start state; squared output -> segment sums
template <typename T>
__global__ void __launch_bounds__(kBlock) kw_apply_kernel(const KwParams p) {
    extern __shared__ __align__(16) unsigned char kw_smem[];
    long long* s_base = reinterpret_cast<long long*>(kw_smem);
    int* s_len = reinterpret_cast<int*>(s_base + kBlock);
    T* s_tile = reinterpret_cast<T*>(s_len + kBlock);
    const T* x = reinterpret_cast<const T*>(p.x);
    const Coeffs c = p.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long* w_base = s_base + warp * 32;
    const int* w_len = s_len + warp * 32;
    T* w_tile = s_tile + warp * 32 * kPitch;
    for (long long item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        __syncwarp();                         // the warp has consumed its previous item
        const ChunkId id = publish_chunk(p, item, s_base, s_len);
        __syncwarp();
        TileLoader<T> ld;
        const T* q0;
        const bool fast = warp_fast_path<T>(p, id, x, &q0);
        if (fast) ld.load_fast(q0, 0, lane); else ld.load(x, w_base, w_len, 0, p.samp_stride, lane);
        const long long i0 = id.chunk * kChunk;
        const long long len = id.len;
        double s[4] = {0, 0, 0, 0};
        if (id.active) {
            const double* st = state_ptr(p, id.row, id.chunk);
            s[0] = st[0]; s[1] = st[1]; s[2] = st[2]; s[3] = st[3];
        }
        // segment containing sample i0: largest k with bounds[k] <= i0 (k = -1 before the first bound)
        long long seg = -1, next = 0x7fffffffffffffffLL;
        if (id.active) {
            long long lo = 0, hi = p.n_seg + 1;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (p.seg_bounds[mid] <= i0) lo = mid + 1; else hi = mid;
            }
            seg = lo - 1;                             // -1: before first bound; n_seg: after last
            next = seg + 1 <= p.n_seg ? p.seg_bounds[seg + 1] : (long long)0x7fffffffffffffffLL;
        }
        double* out = p.seg_sumsq + id.row * p.n_seg;
        double acc = 0.0;
        // A chunk touches at most two segments unless segments are shorter than a chunk: the sums of the segment the
        // chunk starts in and of the next one go to a per-chunk slot and are added up in chunk order by
        // kw_segment_reduce_kernel (atomics would make the last bit of the loudness depend on the scheduling);
        // only a third and later segment inside one chunk falls back to atomicAdd.
        int flushes = 0;
        double part0 = 0.0, part1 = 0.0;
        ld.store(w_tile, lane);
        __syncwarp();

        for (int sb = 0; sb < kChunk / kSub; ++sb) {
            const bool more = sb + 1 < kChunk / kSub;
            if (more) { if (fast) ld.load_fast(q0, sb + 1, lane); else ld.load(x, w_base, w_len, sb + 1, p.samp_stride, lane); }
            const T* mine = w_tile + (sb & 1) * kBlock * kPitch + lane * kPitch;
            const long long base = (long long)sb * kSub;
            if (base < len) {
                const long long g_last = i0 + (base + kSub < len ? base + kSub : len) - 1;
                if (g_last < next && base + kSub <= len) {
                    // no boundary inside this step and a full step (the common case): straight-line recurrence
#pragma unroll
                    for (int j = 0; j < kSub; ++j) {
                        double y;
                        step(c, (double)mine[j], s, &y);
                        acc = fma(y, y, acc);
                    }
                } else {
                    for (int j = 0; j < kSub && base + j < len; ++j) {
                        const long long gi = i0 + base + j;
                        while (gi >= next) {          // crossed a segment boundary: flush
                            if (flushes == 0) part0 = acc;
                            else if (flushes == 1) part1 = acc;
                            else if (seg >= 0 && seg < p.n_seg && acc != 0.0) atomicAdd(out + seg, acc);
                            ++flushes;
                            acc = 0.0;
                            ++seg;
                            next = seg + 1 <= p.n_seg ? p.seg_bounds[seg + 1] : (long long)0x7fffffffffffffffLL;
                        }
                        double y;
                        step(c, (double)mine[j], s, &y);
                        acc = fma(y, y, acc);
                    }
                }
            }
            if (more) ld.store(w_tile + ((sb + 1) & 1) * kBlock * kPitch, lane);
            __syncwarp();
        }
        if (flushes == 0) part0 = acc;
        else if (flushes == 1) part1 = acc;
        else if (seg >= 0 && seg < p.n_seg && acc != 0.0) atomicAdd(out + seg, acc);
        if (id.active) *reinterpret_cast<double2*>(p.partials + id.g * 2) = make_double2(part0, part1);
    }
}

// seg_sumsq[row][k] += the per-chunk partial sums of segment k: one WARP per segment, lane l takes chunks c_lo + l,
// c_lo + l + 32, ... in order and the 32 lane sums are combined by a fixed shuffle tree -- the same order in every
// run, whatever the scheduling.
__global__ void __launch_bounds__(256) kw_segment_reduce_kernel(const KwParams p) {
    const long long idx = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (idx >= p.n_rows * p.n_seg) return;                                        // warp-uniform
    const long long row = idx / p.n_seg, k = idx - row * p.n_seg;
    const long long b0 = p.seg_bounds[k], b1 = p.seg_bounds[k + 1];
    if (b0 >= p.n || b1 <= b0) return;
    const long long c_lo = b0 / kChunk;
    long long c_hi = (b1 - 1) / kChunk;
    if (c_hi > p.n_chunks - 1) c_hi = p.n_chunks - 1;
    const double* part = p.partials + row * p.n_chunks * 2;
    const bool second_ok = k == 0 || p.seg_bounds[k - 1] <= c_lo * kChunk;       // segment k starts at chunk c_lo's FIRST boundary
    double total = 0.0;
    for (long long c = c_lo + lane; c <= c_hi; c += 32) {
        const double2 v = *reinterpret_cast<const double2*>(part + c * 2);
        if (c * kChunk >= b0) total += v.x;                                       // the chunk starts inside segment k
        else if (second_ok) total += v.y;
        // else: a third or later segment of that chunk, added atomically by pass C
    }
    total = dam_warp_sum(total);
    if (lane == 0) p.seg_sumsq[row * p.n_seg + k] += total;
}

// Gating + LUFS on the device (one CTA per signal), so batch pipelines never synchronise with the host.
constexpr int kGateBlock = 256;
constexpr int kMaxCh = 5;

struct GateParams {
    const double* seg;             // [n_signals * C][n_seg]
    long long n_seg, n_blocks;
    int C;
    const long long* first;        // [n_blocks] first segment of block j
    const long long* last;         // [n_blocks] one past the last segment
    double inv_len;                // 1 / (block_size * rate)
    double* out;                   // [n_signals]
    double target;                 // gain_out[sig] = 10^((target - L) / 20) (evaluation.py:64), when gain_out is set
    double* gain_out;              // [n_signals] or null
};

__device__ __forceinline__ double gate_block_sum(double v, double* s_buf) {
    v = dam_warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < kGateBlock / 32; ++i) t += s_buf[i];
    return t;
}

__global__ void __launch_bounds__(kGateBlock) gated_loudness_kernel(const GateParams p) {
    __shared__ double s_buf[kGateBlock / 32];
    const double G[kMaxCh] = {1.0, 1.0, 1.0, 1.41, 1.41};
    const long long sig = blockIdx.x;
    const double* seg = p.seg + sig * p.C * p.n_seg;
    double gamma = -70.0;
    double zavg[kMaxCh];
    for (int pass = 0; pass < 2; ++pass) {
        double acc[kMaxCh] = {0, 0, 0, 0, 0};
        double cnt = 0.0;
        for (long long j = threadIdx.x; j < p.n_blocks; j += kGateBlock) {
            double z[kMaxCh];
            double w = 0.0;
            for (int c = 0; c < p.C; ++c) {
                double a = 0.0;
                for (long long k = p.first[j]; k < p.last[j]; ++k) a += seg[c * p.n_seg + k];
                z[c] = a * p.inv_len;
                w += G[c] * z[c];
            }
            const double l = -0.691 + 10.0 * log10(w);
            const bool keep = pass == 0 ? (l >= -70.0) : (l > gamma && l > -70.0);
            if (keep) {
                for (int c = 0; c < p.C; ++c) acc[c] += z[c];
                cnt += 1.0;
            }
        }
        cnt = gate_block_sum(cnt, s_buf);
        double w = 0.0;
        for (int c = 0; c < p.C; ++c) {
            const double t = gate_block_sum(acc[c], s_buf);
            zavg[c] = cnt > 0.0 ? t / cnt : (pass == 0 ? NAN : 0.0);
            w += G[c] * zavg[c];
        }
        if (pass == 0) gamma = -0.691 + 10.0 * log10(w) - 10.0;      // NaN if nothing passed: pass 2 keeps nothing
        else if (threadIdx.x == 0) {
            const double L = -0.691 + 10.0 * log10(w);
            p.out[sig] = L;
            if (p.gain_out) p.gain_out[sig] = pow(10.0, (p.target - L) / 20.0);
        }
    }
}

}  // namespace

// Scan partition of a row of n_chunks chunks: `parts` CTAs of one thread-block cluster (1 for rows up to 8192 chunks,
// up to 8 for the 41 k chunks of a 4-minute stem), every CTA 1024 threads x `super` consecutive chunks.
void dam_kw_scan_partition(int64_t n_chunks, int* parts, int64_t* super) {
    int p = (int)((n_chunks + 8191) / 8192);
    p = p < 1 ? 1 : (p > 8 ? 8 : p);
    if (p == 3) p = 4;                       // cluster sizes 1, 2, 4, 8
    else if (p > 4 && p < 8) p = 8;
    const int64_t per_part = (n_chunks + p - 1) / p;
    *parts = p;
    *super = (per_part + kScanBlock - 1) / kScanBlock;
}

extern "C" size_t dam_kweight_workspace_bytes(int64_t n_rows, int64_t n_samples) {
    if (n_rows <= 0 || n_samples <= 0) return 0;
    const int64_t n_chunks = (n_samples + kChunk - 1) / kChunk;
    int parts;
    int64_t super;
    dam_kw_scan_partition(n_chunks, &parts, &super);                         // slots per row = parts * super * kScanBlock
    return (size_t)n_rows * (size_t)parts * (size_t)super * kScanBlock * 4 * sizeof(double) +
           (size_t)n_rows * (size_t)n_chunks * 2 * sizeof(double);                // + per-chunk partial segment sums
}

// Pre-computes the response table of one filter-coefficient set (host side; it travels in the kernel parameters,
// so the loudness kernels allocate nothing on the device).  Optional: dam_kweight_* does the same on first use.
static Coeffs parse_coeffs(const double* coeffs_host) {
    Coeffs c;
    for (int s = 0; s < 2; ++s) {
        c.b0[s] = coeffs_host[s * 5 + 0]; c.b1[s] = coeffs_host[s * 5 + 1]; c.b2[s] = coeffs_host[s * 5 + 2];
        c.a1[s] = coeffs_host[s * 5 + 3]; c.a2[s] = coeffs_host[s * 5 + 4];
    }
    return c;
}

int dam_kw_response_table(const double* coeffs_host, const double** G_host) {
    KwTables* t = nullptr;
    const int rc = get_tables(parse_coeffs(coeffs_host), &t);
    if (rc == DAM_OK) *G_host = t->G;
    return rc;
}

extern "C" int dam_kweight_init(const double* coeffs_host) {
    dam_clear_error();
    DAM_CHECK_ARG(coeffs_host != nullptr, "null pointer argument");
    KwTables* t = nullptr;
    return get_tables(parse_coeffs(coeffs_host), &t);
}

// KW_RESUME: the workspace already holds every chunk's zero-state end state (dam_gain_mix_kweight left them there
// while it wrote x), so pass A -- one of the two reads of x -- is skipped.  KW_STATES_ONLY: pass A alone (what
// dam_gain_mix_kweight runs after a mixdown whose layout its fused kernel does not take).
enum KwMode { KW_FULL = 0, KW_RESUME = 1, KW_STATES_ONLY = 2 };

static int kweight_segments(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                            int64_t row_stride, int64_t samp_stride,
                            const double* coeffs_host, const int64_t* seg_bounds, int64_t n_seg,
                            double* seg_sumsq, void* workspace, size_t workspace_bytes, void* stream, KwMode mode) {
    dam_clear_error();
    DAM_CHECK_ARG(x && coeffs_host && workspace, "null pointer argument");
    DAM_CHECK_ARG(mode == KW_STATES_ONLY || (seg_bounds && seg_sumsq), "null pointer argument");
    DAM_CHECK_ARG(dtype == DAM_F32 || dtype == DAM_F64, "dtype must be F32 or F64");
    DAM_CHECK_ARG(n_rows >= 1 && n_rows <= 65535 && n_samples >= 1 && (n_seg >= 1 || mode == KW_STATES_ONLY), "bad sizes");
    DAM_CHECK_ARG(workspace_bytes >= dam_kweight_workspace_bytes(n_rows, n_samples),
                  "workspace too small: need %zu bytes", dam_kweight_workspace_bytes(n_rows, n_samples));
    DAM_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
    const Coeffs c = parse_coeffs(coeffs_host);
    KwTables* t = nullptr;
    int rc = get_tables(c, &t);
    if (rc != DAM_OK) return rc;
    KwParams p;
    p.x = x; p.n_rows = n_rows; p.n = n_samples; p.row_stride = row_stride; p.samp_stride = samp_stride;
    p.n_chunks = (n_samples + kChunk - 1) / kChunk;
    p.n_items = (n_rows * p.n_chunks + kBlock - 1) / kBlock;
    p.c = c;
    memcpy(p.G, t->G, sizeof(p.G));
    memcpy(p.ML, t->ML, sizeof(p.ML));
    {
        int parts;
        int64_t super;
        dam_kw_scan_partition(p.n_chunks, &parts, &super);
        p.parts = parts; p.super = super; p.cpp = super * kScanBlock;
    }
    {
        // P_0 = (M^L)^super by binary exponentiation, P_{j+1} = P_j^2
        double base[16], acc[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        memcpy(base, t->ML, sizeof(base));
        for (long long e = p.super; e > 0; e >>= 1) {
            if (e & 1) matmul4(acc, base, acc);
            matmul4(base, base, base);
        }
        memcpy(p.P[0], acc, sizeof(acc));
        for (int j = 1; j < kScanLevels; ++j) matmul4(p.P[j - 1], p.P[j - 1], p.P[j]);
        matmul4(p.P[kScanLevels - 1], p.P[kScanLevels - 1], p.Q);          // P_0^1024 = (M^L)^cpp
    }
    p.seg_bounds = reinterpret_cast<const long long*>(seg_bounds); p.n_seg = n_seg;
    p.seg_sumsq = seg_sumsq; p.states = reinterpret_cast<double*>(workspace);
    p.zero_seg = mode == KW_RESUME ? 1 : 0;
    p.partials = p.states + (size_t)n_rows * (size_t)p.parts * (size_t)p.cpp * 4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static std::mutex mu;
    static bool configured[64] = {};
    static int occ_f32[64] = {}, occ_f64[64] = {};
    int dev = 0;
    DAM_CHECK_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[dev]) {
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kw_chunk_state_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_smem_bytes<double>()));
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kw_apply_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_smem_bytes<double>()));
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kw_chunk_state_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_smem_bytes<float>()));
            DAM_CHECK_CUDA(cudaFuncSetAttribute(kw_apply_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw_smem_bytes<float>()));
            DAM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f32[dev], kw_apply_kernel<float>, kBlock, kw_smem_bytes<float>()));
            DAM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f64[dev], kw_apply_kernel<double>, kBlock, kw_smem_bytes<double>()));
            configured[dev] = true;
        }
    }
    // persistent grid: a whole number of CTAs per SM, work items dealt round-robin
    const int occ = dtype == DAM_F32 ? occ_f32[dev] : occ_f64[dev];
    const long long n_items = p.n_items;
    long long grid_ll = (long long)dam_sm_count() * (occ > 0 ? occ : 1);
    if (grid_ll > n_items) grid_ll = n_items;
    const unsigned grid = (unsigned)grid_ll;
    if (mode != KW_RESUME) {
        if (dtype == DAM_F32) { kw_chunk_state_kernel<float><<<grid, kBlock, kw_smem_bytes<float>(), st>>>(p); dam_note_launch(); }
        else { kw_chunk_state_kernel<double><<<grid, kBlock, kw_smem_bytes<double>(), st>>>(p); dam_note_launch(); }
        DAM_CHECK_CUDA(cudaGetLastError());
    }
    if (mode == KW_STATES_ONLY) return DAM_OK;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_rows * p.parts));
        cfg.blockDim = dim3(kScanBlock);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.parts;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        DAM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kw_scan_kernel, p));
        dam_note_launch();
    }
    DAM_CHECK_CUDA(cudaGetLastError());
    if (dtype == DAM_F32) { kw_apply_kernel<float><<<grid, kBlock, kw_smem_bytes<float>(), st>>>(p); dam_note_launch(); }
    else { kw_apply_kernel<double><<<grid, kBlock, kw_smem_bytes<double>(), st>>>(p); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    { kw_segment_reduce_kernel<<<(unsigned)((n_rows * n_seg * 32 + 255) / 256), 256, 0, st>>>(p); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_kweight_segments(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                                    int64_t row_stride, int64_t samp_stride,
                                    const double* coeffs_host, const int64_t* seg_bounds, int64_t n_seg,
                                    double* seg_sumsq, void* workspace, size_t workspace_bytes, void* stream) {
    return kweight_segments(x, dtype, n_rows, n_samples, row_stride, samp_stride, coeffs_host, seg_bounds, n_seg, seg_sumsq,
                            workspace, workspace_bytes, stream, KW_FULL);
}

extern "C" int dam_kweight_segments_resume(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                                           int64_t row_stride, int64_t samp_stride,
                                           const double* coeffs_host, const int64_t* seg_bounds, int64_t n_seg,
                                           double* seg_sumsq, void* workspace, size_t workspace_bytes, void* stream) {
    return kweight_segments(x, dtype, n_rows, n_samples, row_stride, samp_stride, coeffs_host, seg_bounds, n_seg, seg_sumsq,
                            workspace, workspace_bytes, stream, KW_RESUME);
}

int dam_kw_chunk_states(const void* x, int dtype, int64_t n_rows, int64_t n_samples, int64_t row_stride,
                        const double* coeffs_host, void* workspace, size_t workspace_bytes, void* stream) {
    return kweight_segments(x, dtype, n_rows, n_samples, row_stride, 1, coeffs_host, nullptr, 0, nullptr, workspace,
                            workspace_bytes, stream, KW_STATES_ONLY);
}

static int gated_loudness(const double* seg_sumsq, int64_t n_signals, int n_channels, int64_t n_seg,
                          const int64_t* block_first, const int64_t* block_last, int64_t n_blocks,
                          double inv_block_len, double target_lufs, double* lufs_out, double* gain_out, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(seg_sumsq && block_first && block_last && lufs_out, "null pointer argument");
    DAM_CHECK_ARG(n_channels >= 1 && n_channels <= kMaxCh, "n_channels must be in [1, 5]");
    DAM_CHECK_ARG(n_signals >= 0 && n_seg >= 1 && n_blocks >= 1, "bad sizes");
    if (n_signals == 0) return DAM_OK;
    GateParams p;
    p.seg = seg_sumsq; p.n_seg = n_seg; p.n_blocks = n_blocks; p.C = n_channels;
    p.first = reinterpret_cast<const long long*>(block_first);
    p.last = reinterpret_cast<const long long*>(block_last);
    p.inv_len = inv_block_len; p.out = lufs_out;
    p.target = target_lufs; p.gain_out = gain_out;
    { gated_loudness_kernel<<<(unsigned)n_signals, kGateBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}

extern "C" int dam_gated_loudness(const double* seg_sumsq, int64_t n_signals, int n_channels, int64_t n_seg,
                                  const int64_t* block_first, const int64_t* block_last, int64_t n_blocks,
                                  double inv_block_len, double* lufs_out, void* stream) {
    return gated_loudness(seg_sumsq, n_signals, n_channels, n_seg, block_first, block_last, n_blocks, inv_block_len, 0.0,
                          lufs_out, nullptr, stream);
}

extern "C" int dam_gated_loudness_gain(const double* seg_sumsq, int64_t n_signals, int n_channels, int64_t n_seg,
                                       const int64_t* block_first, const int64_t* block_last, int64_t n_blocks,
                                       double inv_block_len, double target_lufs, double* lufs_out, double* gain_out,
                                       void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(gain_out != nullptr, "null pointer argument");
    return gated_loudness(seg_sumsq, n_signals, n_channels, n_seg, block_first, block_last, n_blocks, inv_block_len,
                          target_lufs, lufs_out, gain_out, stream);
}

// ------------------------------------------------------------------------------------------
// dam_kweight_blocks: K-weighting + mean square per 400 ms gating block in one call.
// The block geometry is evaluated on the host with pyloudnorm's own floating-point expressions
//   T = n / rate;  num_blocks = int(round((T - T_g) / (T_g * step)) + 1),  step = 1 - overlap
//   l_j = int(T_g * (j * step) * rate),  u_j = int(T_g * (j * step + 1) * rate)
// (meter.py:69-78 of pyloudnorm 0.1.0), turned into disjoint segments for dam_kweight_segments, and
// a small kernel adds the run of segments of every block.
namespace {

struct BlockGeometry {
    std::vector<long long> bounds, first, last;
    long long n_blocks = 0, n_seg = 0;
};

bool block_geometry(int64_t n, double rate, double T_g, double overlap, BlockGeometry* g) {
    if (n <= 0 || rate <= 0 || T_g <= 0 || overlap < 0 || overlap >= 1) return false;
    const double step = 1.0 - overlap;
    const double T = (double)n / rate;
    if ((double)n < T_g * rate) return false;
    const long long nb = (long long)(nearbyint((T - T_g) / (T_g * step)) + 1);
    if (nb < 1) return false;
    std::vector<long long> lo(nb), hi(nb);
    for (long long j = 0; j < nb; ++j) {
        lo[j] = (long long)(T_g * ((double)j * step) * rate);
        hi[j] = (long long)(T_g * ((double)j * step + 1) * rate);
        if (hi[j] > n) hi[j] = n;
    }
    g->bounds = lo;
    g->bounds.insert(g->bounds.end(), hi.begin(), hi.end());
    std::sort(g->bounds.begin(), g->bounds.end());
    g->bounds.erase(std::unique(g->bounds.begin(), g->bounds.end()), g->bounds.end());
    g->first.resize(nb); g->last.resize(nb);
    for (long long j = 0; j < nb; ++j) {
        g->first[j] = std::lower_bound(g->bounds.begin(), g->bounds.end(), lo[j]) - g->bounds.begin();
        g->last[j] = std::lower_bound(g->bounds.begin(), g->bounds.end(), hi[j]) - g->bounds.begin();
    }
    g->n_blocks = nb;
    g->n_seg = (long long)g->bounds.size() - 1;
    return g->n_seg >= 1;
}

inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

__global__ void block_sum_kernel(const double* __restrict__ seg, long long n_seg, const long long* __restrict__ first,
                                 const long long* __restrict__ last, long long n_blocks, double inv_len,
                                 double* __restrict__ z) {
    const long long r = blockIdx.y;
    const double* s = seg + r * n_seg;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_blocks; j += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (long long k = first[j]; k < last[j]; ++k) acc += s[k];
        z[r * n_blocks + j] = acc * inv_len;
    }
}

}  // namespace

extern "C" int64_t dam_kweight_num_blocks(int64_t n_samples, double rate, double block_seconds, double overlap) {
    BlockGeometry g;
    return block_geometry(n_samples, rate, block_seconds, overlap, &g) ? (int64_t)g.n_blocks : 0;
}

extern "C" size_t dam_kweight_blocks_workspace_bytes(int64_t n_rows, int64_t n_samples, double rate,
                                                      double block_seconds, double overlap) {
    BlockGeometry g;
    if (n_rows <= 0 || !block_geometry(n_samples, rate, block_seconds, overlap, &g)) return 0;
    return align16(dam_kweight_workspace_bytes(n_rows, n_samples)) + align16((g.bounds.size() + 2 * g.n_blocks) * sizeof(long long)) +
           align16((size_t)n_rows * g.n_seg * sizeof(double));
}

extern "C" int dam_kweight_blocks(const void* x, int dtype, int64_t n_rows, int64_t n_samples,
                                  int64_t row_stride, int64_t samp_stride, double rate,
                                  const double* coeffs_host, double block_seconds, double overlap,
                                  double* z_out, void* workspace, size_t workspace_bytes, void* stream) {
    dam_clear_error();
    DAM_CHECK_ARG(x && coeffs_host && z_out && workspace, "null pointer argument");
    BlockGeometry g;
    DAM_CHECK_ARG(block_geometry(n_samples, rate, block_seconds, overlap, &g),
                  "audio must be longer than the gating block (n=%lld, rate=%g, block=%g s)", (long long)n_samples, rate,
                  block_seconds);
    const size_t need = dam_kweight_blocks_workspace_bytes(n_rows, n_samples, rate, block_seconds, overlap);
    DAM_CHECK_ARG(workspace_bytes >= need, "workspace too small: need %zu bytes", need);
    DAM_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "workspace must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    const size_t kw_bytes = align16(dam_kweight_workspace_bytes(n_rows, n_samples));
    long long* d_geo = reinterpret_cast<long long*>(w + kw_bytes);
    const size_t geo_count = g.bounds.size() + 2 * (size_t)g.n_blocks;
    double* d_seg = reinterpret_cast<double*>(w + kw_bytes + align16(geo_count * sizeof(long long)));
    std::vector<long long> host(g.bounds);
    host.insert(host.end(), g.first.begin(), g.first.end());
    host.insert(host.end(), g.last.begin(), g.last.end());
    // pageable source: the runtime stages it before returning, so `host` may go out of scope
    DAM_CHECK_CUDA(cudaMemcpyAsync(d_geo, host.data(), geo_count * sizeof(long long), cudaMemcpyHostToDevice, st));
    DAM_CHECK_CUDA(cudaMemsetAsync(d_seg, 0, (size_t)n_rows * g.n_seg * sizeof(double), st));
    const int rc = dam_kweight_segments(x, dtype, n_rows, n_samples, row_stride, samp_stride, coeffs_host,
                                        reinterpret_cast<const int64_t*>(d_geo), g.n_seg, d_seg, workspace, kw_bytes, stream);
    if (rc != DAM_OK) return rc;
    const long long* d_first = d_geo + g.bounds.size();
    const long long* d_last = d_first + g.n_blocks;
    dim3 grid((unsigned)((g.n_blocks + 127) / 128), (unsigned)n_rows);
    { block_sum_kernel<<<grid, 128, 0, st>>>(d_seg, g.n_seg, d_first, d_last, g.n_blocks, 1.0 / (block_seconds * rate), z_out); dam_note_launch(); }
    DAM_CHECK_CUDA(cudaGetLastError());
    return DAM_OK;
}
