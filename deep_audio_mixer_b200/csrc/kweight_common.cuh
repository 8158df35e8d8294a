// Pieces of K4 (loudness.cu) that the mixdown kernel K2 (mix_kernels.cu) shares when it produces the chunk end
// states of the K-weighting filters while the mix is still in its registers (dam_gain_mix_kweight): the chunk
// size, where the state of a chunk lives in the workspace, and the ORDER in which a chunk's 256 response-table
// products are added -- both producers must leave the same bits.
#pragma once

#include <stdint.h>

constexpr int kKwChunk = 256;          // samples per chunk (L)
constexpr int kKwScanBlock = 1024;     // threads per CTA of the scan
constexpr int kKwGroups = kKwChunk / 4;

// Where the filter states of the chunks of all rows live.  The states of a row are stored TRANSPOSED for the scan --
// thread u of the scan owns the run of chunks [u * super, (u + 1) * super), and slot
// (c % super) * kKwScanBlock + c / super makes step k of all threads one coalesced access (with the natural order
// every scan load touched 32 separate sectors per warp and a 4-minute stem spent more time in the scan than in
// either pass over the samples).  A row is scanned by `parts` CTAs of one cluster, cpp = super * kKwScanBlock chunks each.
struct KwStateLayout {
    double* states;                // [n_rows][parts * cpp][4]
    long long super, cpp;
    int parts;
};

__device__ __forceinline__ double* kw_state_ptr(const KwStateLayout& s, long long row, long long chunk) {
    const long long part = chunk / s.cpp;
    const long long cl = chunk - part * s.cpp;
    const long long u = cl / s.super;
    const long long k = cl - u * s.super;
    return s.states + ((row * s.parts + part) * s.cpp + k * kKwScanBlock + u) * 4;
}

// Zero-state end state of a chunk = sum_i G[i][:] x[i].  Summation order (the same wherever it is evaluated): the
// chunk is cut into 64 groups of 4 consecutive samples, a group is one product followed by three FMAs, and the 64
// group sums are added PAIRWISE (neighbours first: a balanced binary tree over the groups in sample order).  A
// thread walking a chunk sample by sample builds that tree with a binary counter; 32 lanes holding two groups each
// build it with xor-shuffles 1, 2, 4, 8, 16 and one last addition of the two halves.
// g: the 16 table entries of the group, g[j * 4 + c] = G[4 q + j][c].
__device__ __forceinline__ void kw_group_sum(const double* g, double x0, double x1, double x2, double x3, double* S) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double s = __dmul_rn(g[c], x0);
        s = fma(g[4 + c], x1, s);
        s = fma(g[8 + c], x2, s);
        S[c] = fma(g[12 + c], x3, s);
    }
}

// host side (loudness.cu)
void dam_kw_scan_partition(int64_t n_chunks, int* parts, int64_t* super);
// response table G[kKwChunk][4] of a coefficient set (HOST double[10]); the pointer stays valid for the process
int dam_kw_response_table(const double* coeffs_host, const double** G_host);
// pass A of K4 alone: chunk end states of x [n_rows][n_samples] (unit sample stride) into the workspace
int dam_kw_chunk_states(const void* x, int dtype, int64_t n_rows, int64_t n_samples, int64_t row_stride,
                        const double* coeffs_host, void* workspace, size_t workspace_bytes, void* stream);
