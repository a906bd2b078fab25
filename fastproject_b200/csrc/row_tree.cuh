// NumPy's pairwise summation tree over one contiguous row, shared by the kernels that must
// reproduce add.reduce(axis=1) bit for bit (normalize.cu, weights.cu).
//
// add.reduce along the contiguous axis sums blocks of <= 128 elements with eight interleaved
// accumulators and combines halves recursively (numpy/_core/src/umath/loops_utils.h.src:
// n > 128 splits at n/2 rounded down to a multiple of 8).  The tree depends on N only: the host
// enumerates its leaves once per call; on the device one thread sums a leaf
// (fp::np_pairwise_sum) and one thread folds the leaf sums with the tree's post-order program.
#pragma once
#include <vector>

#include "fp_common.cuh"

namespace fp {
namespace rowtree {

struct Leaf {
    int lo, n;
};

constexpr int MAX_DEPTH = 48;

// post-order program: entry >= 0 pushes the sum of that leaf, -1 pops two and pushes their sum
__device__ inline double fold_tree(const double *leaf_sum, const int *__restrict__ prog, int n_prog, double *stack) {
    int sp = 0;
    for (int p = 0; p < n_prog; ++p) {
        const int op = prog[p];
        if (op >= 0) {
            stack[sp++] = leaf_sum[op];
        } else {
            const double b = stack[--sp], a = stack[--sp];
            stack[sp++] = __dadd_rn(a, b);
        }
    }
    return stack[0];
}

// leaf sums of one row for the whole CTA: eight lanes per leaf (blockDim.x a multiple of 32);
// a row shorter than 8 elements is a single short leaf summed by one thread
template <typename Load>
__device__ __forceinline__ void leaf_sums(Load load, const Leaf *__restrict__ leaves, int n_leaves, double *leaf_sum) {
    if (n_leaves == 1 && leaves[0].n < 8) {
        if (threadIdx.x == 0) {                  // short loop, from -0.0 like NumPy's
            double res = -0.0;
            for (int i = 0; i < leaves[0].n; ++i) res = __dadd_rn(res, load((long)leaves[0].lo + i));
            leaf_sum[0] = res;
        }
        return;
    }
    const int per_pass = blockDim.x >> 3;
    for (int base = 0; base < n_leaves; base += per_pass) {
        const int l = base + (threadIdx.x >> 3);
        const bool live = l < n_leaves;
        const Leaf leaf = leaves[live ? l : 0];
        const double v = fp::np_leaf_sum_octet(load, leaf.lo, leaf.n);
        if (live && (threadIdx.x & 7) == 0) leaf_sum[l] = v;
    }
}

// The same enumeration on the device (iterative: explicit stack of pending halves), so that the
// tree reaches the kernels without a host -> device copy -- a copy-engine transfer would queue
// behind whatever large upload is in flight, and its pageable source forces a stream sync.
__device__ inline void build_tree_device(long lo, long n, Leaf *leaves, int *n_leaves, int *prog, int *n_prog) {
    long st_lo[MAX_DEPTH], st_n[MAX_DEPTH];
    signed char st_seen[MAX_DEPTH];
    int sp = 1, nl = 0, np = 0;
    st_lo[0] = lo; st_n[0] = n; st_seen[0] = 0;
    while (sp) {
        const long flo = st_lo[sp - 1], fn = st_n[sp - 1];
        if (fn <= 128) {
            prog[np++] = nl;
            leaves[nl].lo = (int)flo;
            leaves[nl].n = (int)fn;
            ++nl;
            --sp;
        } else if (!st_seen[sp - 1]) {
            st_seen[sp - 1] = 1;
            long n2 = fn / 2;
            n2 -= n2 % 8;
            st_lo[sp] = flo + n2; st_n[sp] = fn - n2; st_seen[sp] = 0; ++sp;     // right half: visited second
            st_lo[sp] = flo; st_n[sp] = n2; st_seen[sp] = 0; ++sp;               // left half: visited first
        } else {
            prog[np++] = -1;                     // both halves summed: add them
            --sp;
        }
    }
    if (n_leaves) *n_leaves = nl;
    if (n_prog) *n_prog = np;
}

// one thread fills the leaf table and the program of rows of N elements in the workspace
static __global__ void build_tree_kernel(long N, Leaf *leaves, int *prog) {
    if (blockIdx.x == 0 && threadIdx.x == 0) build_tree_device(0, N, leaves, nullptr, prog, nullptr);
}

inline void build_tree(long lo, long n, std::vector<Leaf> &leaves, std::vector<int> &prog, int depth, int &max_depth) {
    if (depth > max_depth) max_depth = depth;
    if (n <= 128) {
        prog.push_back((int)leaves.size());
        leaves.push_back(Leaf{(int)lo, (int)n});
        return;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    build_tree(lo, n2, leaves, prog, depth + 1, max_depth);
    build_tree(lo + n2, n - n2, leaves, prog, depth + 1, max_depth);
    prog.push_back(-1);
}

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// bytes of device workspace for the leaves + program of rows of N elements
inline size_t workspace_bytes(long N) {
    const size_t max_leaves = (size_t)(N / 32 + 2);
    return align_up(max_leaves * sizeof(Leaf)) + align_up(2 * max_leaves * sizeof(int)) + 256;
}

}  // namespace rowtree
}  // namespace fp
