// Shared helpers for the fastproject_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fastproject_b200.h"

namespace fp {

// thread-local last-error text returned by fp_last_error()
void set_error(const char *fmt, ...);
void clear_error();
void count_launch(int n = 1);

// B200: 148 SMs.  Queried once; used to size persistent grids.
int sm_count();

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

#define FP_CHECK_ARG(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            fp::set_error(__VA_ARGS__);          \
            return FP_EINVAL;                    \
        }                                        \
    } while (0)

#define FP_CHECK_CUDA(expr)                                                        \
    do {                                                                           \
        cudaError_t err__ = (expr);                                                \
        if (err__ != cudaSuccess) {                                                \
            fp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                          __FILE__, __LINE__);                                     \
            return FP_ECUDA;                                                       \
        }                                                                          \
    } while (0)

#define FP_CHECK_LAUNCH(name)                                                      \
    do {                                                                           \
        cudaError_t err__ = cudaGetLastError();                                    \
        if (err__ != cudaSuccess) {                                                \
            fp::set_error("launch of %s failed: %s", name, cudaGetErrorString(err__)); \
            return FP_ECUDA;                                                       \
        }                                                                          \
        fp::count_launch();                                                        \
    } while (0)

// NumPy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
// pairwise_sum_DOUBLE): < 8 elements sequential; <= 128 elements eight
// interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then
// a sequential tail; otherwise split at n/2 rounded down to a multiple of 8.
// Reproduced so that row reductions (axis = contiguous axis) are bit-identical.
// Written without recursion (explicit stack of pending halves): a recursive device function
// needs stack the compiler cannot size, and overflows the default 1 KB per thread silently when
// the caller's frame is large or the row is long.
template <typename Load>
__device__ double np_pairwise_sum(Load load, long lo, long n) {
    long st_lo[40], st_n[40];
    signed char st_seen[40];
    double val[40];
    int sp = 1, vp = 0;
    st_lo[0] = lo; st_n[0] = n; st_seen[0] = 0;
    while (sp) {
        const long flo = st_lo[sp - 1], fn = st_n[sp - 1];
        if (fn <= 128) {
            double res;
            if (fn < 8) {
                res = -0.0;
                for (long i = 0; i < fn; i++) res = __dadd_rn(res, load(flo + i));
            } else {
                double r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = load(flo + k);
                long i;
                for (i = 8; i < fn - (fn % 8); i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], load(flo + i + k));
                }
                res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; i < fn; i++) res = __dadd_rn(res, load(flo + i));
            }
            val[vp++] = res;
            --sp;
        } else if (!st_seen[sp - 1]) {
            st_seen[sp - 1] = 1;
            long n2 = fn / 2;
            n2 -= n2 % 8;
            st_lo[sp] = flo + n2; st_n[sp] = fn - n2; st_seen[sp] = 0; ++sp;     // right half: summed second
            st_lo[sp] = flo; st_n[sp] = n2; st_seen[sp] = 0; ++sp;               // left half: summed first
        } else {
            const double b = val[--vp], a = val[--vp];
            val[vp++] = __dadd_rn(a, b);
            --sp;
        }
    }
    return val[0];
}

// One leaf of that tree (8 <= n <= 128) summed by EIGHT consecutive lanes: lane k owns the
// accumulator r_k of the reference loop (elements k, k + 8, ... in order), the eight partial
// sums are combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) by an xor-shuffle tree (IEEE
// addition is commutative, so both partners of a shuffle compute the same bits), and the
// n % 8 tail is added sequentially.  Same operations in the same order as np_pairwise_sum on
// the leaf, with 64-byte coalesced loads and eight times the threads at work.  All 32 lanes of
// the warp must call it (lanes whose leaf is a dummy pass any valid leaf).
template <typename Load>
__device__ __forceinline__ double np_leaf_sum_octet(Load load, long lo, long n) {
    const int k = threadIdx.x & 7;
    const long body = n - (n % 8);
    double r = load(lo + k);
    long i = 8;
    for (; i + 24 < body; i += 32) {                 // four loads in flight, added in order
        const double a = load(lo + i + k), b = load(lo + i + 8 + k), c = load(lo + i + 16 + k),
                     d = load(lo + i + 24 + k);
        r = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r, a), b), c), d);
    }
    for (; i < body; i += 8) r = __dadd_rn(r, load(lo + i + k));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    for (i = body; i < n; ++i) r = __dadd_rn(r, load(lo + i));
    return r;
}

}  // namespace fp
