// NumPy's pairwise summation tree over one contiguous row, shared by the kernels that must
// reproduce add.reduce(axis=1) bit for bit (normalize.cu, weights.cu).
//
// add.reduce along the contiguous axis sums blocks of <= 128 elements with eight interleaved
// accumulators and combines halves recursively (numpy/_core/src/umath/loops_utils.h.src:
// n > 128 splits at n/2 rounded down to a multiple of 8).  The tree depends on N only: the host
// enumerates its leaves once per call; on the device one thread sums a leaf
// (fp::np_pairwise_sum) and one thread folds the leaf sums with the tree's post-order program.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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

// The fold as a table of internal nodes for a whole CTA: node k (post-order) adds the value
// slots l and r into slot n_leaves + k; h is its height (1 = both children are leaves).  Nodes
// of one height are independent, so the CTA folds level by level (fold_levels) instead of one
// thread walking the post-order program: same additions, same operands, same order of operands.
struct Node {
    int l, r, h, pad;
};

__device__ inline void build_nodes_device(const int *prog, int n_prog, int n_leaves, Node *nodes) {
    int sid[MAX_DEPTH], sh[MAX_DEPTH];
    int sp = 0, k = 0;
    for (int p = 0; p < n_prog; ++p) {
        const int op = prog[p];
        if (op >= 0) {
            sid[sp] = op; sh[sp] = 0; ++sp;
        } else {
            const int b = --sp, a = --sp;
            const int h = (sh[a] > sh[b] ? sh[a] : sh[b]) + 1;
            nodes[k] = Node{sid[a], sid[b], h, 0};
            sid[sp] = n_leaves + k; sh[sp] = h; ++sp;
            ++k;
        }
    }
}

// val[0 .. n_leaves) holds the leaf sums (written before a __syncthreads()); returns the root to
// every thread.  All threads of the CTA must call it; val needs 2 * n_leaves slots.
__device__ __forceinline__ double fold_levels(double *val, const Node *__restrict__ nodes, int n_leaves, int height) {
    const int n_int = n_leaves - 1;
    const Node mine = (int)threadIdx.x < n_int ? nodes[threadIdx.x] : Node{0, 0, 0, 0};
    for (int h = 1; h <= height; ++h) {
        if (mine.h == h) val[n_leaves + threadIdx.x] = __dadd_rn(val[mine.l], val[mine.r]);
        for (int k = threadIdx.x + blockDim.x; k < n_int; k += blockDim.x) {
            const Node nd = nodes[k];
            if (nd.h == h) val[n_leaves + k] = __dadd_rn(val[nd.l], val[nd.r]);
        }
        __syncthreads();
    }
    const double root = n_int ? val[n_leaves + n_int - 1] : val[0];
    __syncthreads();             // everybody has the root: val may be refilled (a one-leaf tree's root IS slot 0)
    return root;
}

// one thread fills the leaf table, the program and the node table of rows of N elements
static __global__ void build_tree_kernel(long N, Leaf *leaves, int *prog, Node *nodes) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int n_leaves = 0, n_prog = 0;
        build_tree_device(0, N, leaves, &n_leaves, prog, &n_prog);
        if (nodes) build_nodes_device(prog, n_prog, n_leaves, nodes);
    }
}

// ---- a whole row staged in shared memory (rows that fit: one CTA per SM, one DRAM read) ------
// One thread queues the row as bulk asynchronous copies (TMA, 1-D) that complete on an mbarrier;
// the NEXT row of the CTA is pulled into L2 meanwhile.  Rows must start on 16-byte boundaries.
constexpr uint32_t STAGE_CHUNK = 32768;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// thread-0 side: bytes is a multiple of 16
__device__ __forceinline__ void stage_row_async(double *dst, const double *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    for (uint32_t off = 0; off < bytes; off += STAGE_CHUNK) {
        const uint32_t n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_addr(dst) + off),
                     "l"(reinterpret_cast<const char *>(src) + off), "r"(n), "r"(smem_addr(bar))
                     : "memory");
    }
}
__device__ __forceinline__ void prefetch_row_l2(const double *src, uint32_t bytes) {
    for (uint32_t off = 0; off < bytes; off += STAGE_CHUNK) {
        const uint32_t n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char *>(src) + off),
                     "r"(n)
                     : "memory");
    }
}
__device__ __forceinline__ void stage_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FP_STAGE_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FP_STAGE_DONE;\n\t"
        "bra FP_STAGE_WAIT;\n\t"
        "FP_STAGE_DONE:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// host side: does the staged kernel apply?  (FASTPROJECT_B200_ROW_STAGE=off keeps the kernels
// that read the rows from global memory -- both give the same bits; tests run both)
inline bool stage_applies(const void *X, long ld, size_t smem_bytes) {
    const char *env = getenv("FASTPROJECT_B200_ROW_STAGE");
    if (env && strcmp(env, "off") == 0) return false;
    return (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (ld % 2 == 0) && smem_bytes <= 224 * 1024;
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
    return align_up(max_leaves * sizeof(Leaf)) + align_up(2 * max_leaves * sizeof(int)) +
           align_up(max_leaves * sizeof(Node)) + 256;
}

}  // namespace rowtree
}  // namespace fp
