// Host-side label matching: gene names of many signatures -> rows of the expression matrix.
//
// Replaces the per-signature Python loop of Signature.sig_indices
// (/root/reference/FastProject/Signatures.py:738-753), which probes the signature's dictionary
// with every row label of the matrix (O(genes) per signature), by one hash table over the row
// labels probed once per signature gene.  Pure host code: the strings never go to the device; the
// Python mirror passes both sides as NUL-separated UTF-8 blobs ("\0".join(...)), so no Python
// object is touched per gene on the way in or out.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "fp_common.cuh"

namespace {

inline uint64_t fnv1a(const char *s, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) {
        h ^= (unsigned char)s[i];
        h *= 1099511628211ull;
    }
    return h;
}

struct Slot {
    uint64_t hash;
    int64_t off;     // -1: empty
    int32_t len;
    int32_t row;
};

}  // namespace

extern "C" int fp_match_labels(const char *labels, int64_t labels_bytes, int64_t n_labels, const char *queries,
                               int64_t queries_bytes, int64_t n_queries, int32_t *out_rows_host) {
    fp::clear_error();
    FP_CHECK_ARG(n_labels >= 0 && n_queries >= 0 && labels_bytes >= 0 && queries_bytes >= 0 &&
                     n_labels < (1ll << 31),
                 "fp_match_labels: bad sizes");
    FP_CHECK_ARG((labels || labels_bytes == 0) && (queries || queries_bytes == 0) && (out_rows_host || n_queries == 0),
                 "fp_match_labels: null pointer");
    if (n_queries == 0) return FP_OK;
    size_t cap = 16;
    while (cap < (size_t)n_labels * 2 + 1) cap <<= 1;
    std::vector<Slot> table(cap);
    for (size_t i = 0; i < cap; ++i) table[i].off = -1;
    const size_t mask = cap - 1;
    // labels: n_labels tokens separated by n_labels - 1 NULs
    int64_t pos = 0;
    int duplicates = 0;
    for (int64_t r = 0; r < n_labels; ++r) {
        const void *nul = pos <= labels_bytes ? memchr(labels + pos, 0, (size_t)(labels_bytes - pos)) : nullptr;
        const int64_t end = nul ? (const char *)nul - labels : labels_bytes;
        if (!nul && r != n_labels - 1) {
            fp::set_error("fp_match_labels: label blob holds fewer than %ld labels", (long)n_labels);
            return FP_EINVAL;
        }
        const size_t len = (size_t)(end - pos);
        const uint64_t h = fnv1a(labels + pos, len);
        size_t s = (size_t)h & mask;
        while (table[s].off >= 0) {
            if (table[s].hash == h && (size_t)table[s].len == len && !memcmp(labels + table[s].off, labels + pos, len)) {
                duplicates = 1;          // the reference's loop matches every duplicated row: caller's slow path
                break;
            }
            s = (s + 1) & mask;
        }
        if (table[s].off < 0) {
            table[s].hash = h;
            table[s].off = pos;
            table[s].len = (int32_t)len;
            table[s].row = (int32_t)r;
        }
        pos = end + 1;
    }
    if (n_labels > 0 && pos <= labels_bytes) {
        fp::set_error("fp_match_labels: label blob holds more than %ld labels (a label contains NUL?)", (long)n_labels);
        return FP_EINVAL;
    }
    if (duplicates) return 1;
    pos = 0;
    for (int64_t q = 0; q < n_queries; ++q) {
        const void *nul = pos <= queries_bytes ? memchr(queries + pos, 0, (size_t)(queries_bytes - pos)) : nullptr;
        const int64_t end = nul ? (const char *)nul - queries : queries_bytes;
        if (!nul && q != n_queries - 1) {
            fp::set_error("fp_match_labels: query blob holds fewer than %ld names", (long)n_queries);
            return FP_EINVAL;
        }
        const size_t len = (size_t)(end - pos);
        const uint64_t h = fnv1a(queries + pos, len);
        size_t s = (size_t)h & mask;
        int32_t row = -1;
        while (table[s].off >= 0) {
            if (table[s].hash == h && (size_t)table[s].len == len && !memcmp(labels + table[s].off, queries + pos, len)) {
                row = table[s].row;
                break;
            }
            s = (s + 1) & mask;
        }
        out_rows_host[q] = row;
        pos = end + 1;
    }
    if (pos <= queries_bytes) {
        fp::set_error("fp_match_labels: query blob holds more than %ld names (a name contains NUL?)", (long)n_queries);
        return FP_EINVAL;
    }
    return FP_OK;
}
