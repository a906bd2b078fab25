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

// One pass over a NUL-separated blob: the token's end is found while it is hashed (FNV-1a), so
// a name costs neither a memchr nor a memcpy call.  Returns the position of the token's end
// (the NUL, or `end`).
inline const char *scan_token(const char *p, const char *end, uint64_t *hash) {
    uint64_t h = 1469598103934665603ull;
    while (p < end && *p) {
        h = (h ^ (unsigned char)*p) * 1099511628211ull;
        ++p;
    }
    *hash = h;
    return p;
}

inline bool same_bytes(const char *a, const char *b, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (a[i] != b[i]) return false;
    return true;
}

struct Slot {            // 16 bytes: four slots per cache line
    uint32_t hash;
    int32_t off;         // -1: empty
    int32_t len;
    int32_t row;
};

}  // namespace

extern "C" int fp_match_labels(const char *labels, int64_t labels_bytes, int64_t n_labels, const char *queries,
                               int64_t queries_bytes, int64_t n_queries, int32_t *out_rows_host) {
    fp::clear_error();
    FP_CHECK_ARG(n_labels >= 0 && n_queries >= 0 && labels_bytes >= 0 && queries_bytes >= 0 &&
                     n_labels < (1ll << 30) && labels_bytes < (1ll << 31),
                 "fp_match_labels: bad sizes");
    FP_CHECK_ARG((labels || labels_bytes == 0) && (queries || queries_bytes == 0) && (out_rows_host || n_queries == 0),
                 "fp_match_labels: null pointer");
    if (n_queries == 0) return FP_OK;
    size_t cap = 16;
    int shift = 60;                     // table index = top bits of hash * golden ratio
    while (cap < (size_t)n_labels * 2 + 1) {
        cap <<= 1;
        --shift;
    }
    std::vector<Slot> table(cap);
    for (size_t i = 0; i < cap; ++i) table[i].off = -1;
    const size_t mask = cap - 1;
    // labels: n_labels tokens separated by n_labels - 1 NULs
    const char *p = labels, *lend = labels + labels_bytes;
    int duplicates = 0;
    for (int64_t r = 0; r < n_labels; ++r) {
        if (p > lend) {
            fp::set_error("fp_match_labels: label blob holds fewer than %ld labels", (long)n_labels);
            return FP_EINVAL;
        }
        uint64_t h64;
        const char *tok = p;
        p = scan_token(p, lend, &h64);
        const size_t len = (size_t)(p - tok);
        const uint32_t h = (uint32_t)h64;
        size_t s = (size_t)((h64 * 0x9E3779B97F4A7C15ull) >> shift);
        while (table[s].off >= 0) {
            if (table[s].hash == h && (size_t)table[s].len == len && same_bytes(labels + table[s].off, tok, len)) {
                duplicates = 1;          // the reference's loop matches every duplicated row: caller's slow path
                break;
            }
            s = (s + 1) & mask;
        }
        if (table[s].off < 0) {
            table[s].hash = h;
            table[s].off = (int32_t)(tok - labels);
            table[s].len = (int32_t)len;
            table[s].row = (int32_t)r;
        }
        ++p;                             // past the separator (or one past the end after the last label)
    }
    if (n_labels > 0 && p <= lend) {
        fp::set_error("fp_match_labels: label blob holds more than %ld labels (a label contains NUL?)", (long)n_labels);
        return FP_EINVAL;
    }
    if (duplicates) return 1;
    const char *q = queries, *qend = queries + queries_bytes;
    for (int64_t k = 0; k < n_queries; ++k) {
        if (q > qend) {
            fp::set_error("fp_match_labels: query blob holds fewer than %ld names", (long)n_queries);
            return FP_EINVAL;
        }
        uint64_t h64;
        const char *tok = q;
        q = scan_token(q, qend, &h64);
        const size_t len = (size_t)(q - tok);
        const uint32_t h = (uint32_t)h64;
        size_t s = (size_t)((h64 * 0x9E3779B97F4A7C15ull) >> shift);
        int32_t row = -1;
        while (table[s].off >= 0) {
            if (table[s].hash == h && (size_t)table[s].len == len && same_bytes(labels + table[s].off, tok, len)) {
                row = table[s].row;
                break;
            }
            s = (s + 1) & mask;
        }
        out_rows_host[k] = row;
        ++q;
    }
    if (q <= qend) {
        fp::set_error("fp_match_labels: query blob holds more than %ld names (a name contains NUL?)", (long)n_queries);
        return FP_EINVAL;
    }
    return FP_OK;
}
