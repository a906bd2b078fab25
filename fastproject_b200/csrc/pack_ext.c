/* Signature -> CSR packing on the host, straight from the Python dicts.
 *
 * Replaces the per-signature loop of Signature.sig_indices
 * (/root/reference/FastProject/Signatures.py:738-753) for a whole list of signatures at once:
 * every (gene, value) item of every sig_dict is looked up in a table of the row labels of the
 * expression matrix (open addressing on the strings' cached hashes, hit by object identity or by
 * string equality: no re-hashing, no copies of the names, two cache lines per lookup) and
 * written as (row, sign) into the CSR arrays the scoring kernel reads, rows ascending inside a
 * signature.  A dict value equal to 1 or 0 counts as +1, equal to -1 as -1, anything else is
 * ignored (:750-753); signatures with no matched gene or fewer than `min_genes` are rejected
 * (SigScoreMethods.py:60-64).  Anything unusual -- duplicated or non-string row labels (no table
 * is built), values that are neither int nor float -- returns None and the caller takes the
 * per-signature Python path.
 *
 * A CPython extension (not part of the C ABI of include/fastproject_b200.h): it walks Python
 * objects.  fp_match_labels is the same matcher for callers that hold plain label buffers.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MATCH_THREADS 4

static int cmp_i64(const void *a, const void *b) {
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* keys = row * 2 + (sign > 0), rows distinct and < n_rows: ascending.  Short lists by insertion;
 * longer ones through a bitmap of the rows (set, then swept in order: O(n + n_rows / 64)) when
 * that is cheaper than comparing. */
static void sort_keys(int64_t *k, Py_ssize_t n, uint64_t *seen, uint64_t *positive, Py_ssize_t n_rows) {
    if (n > 32 && seen && (n_rows >> 6) < 8 * n) {
        Py_ssize_t lo = n_rows >> 6, hi = 0;
        for (Py_ssize_t i = 0; i < n; ++i) {
            const int64_t row = k[i] >> 1;
            const Py_ssize_t w = (Py_ssize_t)(row >> 6);
            seen[w] |= 1ull << (row & 63);
            if (k[i] & 1) positive[w] |= 1ull << (row & 63);
            if (w < lo) lo = w;
            if (w > hi) hi = w;
        }
        Py_ssize_t out = 0;
        for (Py_ssize_t w = lo; w <= hi; ++w) {
            uint64_t bits = seen[w];
            const uint64_t pos = positive[w];
            while (bits) {
                const int b = __builtin_ctzll(bits);
                bits &= bits - 1;
                k[out++] = (((int64_t)w << 6) + b) * 2 + (int64_t)((pos >> b) & 1);
            }
            seen[w] = 0;
            positive[w] = 0;
        }
        return;
    }
    if (n <= 48) {                                  /* the usual signature: insertion sort */
        for (Py_ssize_t i = 1; i < n; ++i) {
            const int64_t v = k[i];
            Py_ssize_t j = i;
            while (j > 0 && k[j - 1] > v) {
                k[j] = k[j - 1];
                --j;
            }
            k[j] = v;
        }
    } else {
        qsort(k, (size_t)n, sizeof(int64_t), cmp_i64);
    }
}

/* ---- label table ------------------------------------------------------------------------- */
typedef struct {
    PyObject *key;          /* strong reference; NULL = empty slot */
    Py_hash_t hash;
    int64_t row;
} Slot;

typedef struct {
    Slot *slots;
    size_t mask;
    Py_ssize_t n_rows;
} Table;

static void table_free(PyObject *capsule) {
    Table *t = (Table *)PyCapsule_GetPointer(capsule, "fp_label_table");
    if (!t) return;
    if (t->slots) {
        for (size_t i = 0; i <= t->mask; ++i) Py_XDECREF(t->slots[i].key);
        free(t->slots);
    }
    free(t);
}

/* row of a name, -1: not on the matrix, -3: not for this table (read-only: safe without the GIL) */
static inline int64_t table_find(const Table *t, PyObject *name) {
    if (!PyUnicode_CheckExact(name)) {
        /* other hashables can equal a str only through their own __eq__ (str subclasses, ...):
         * the dict-based path decides */
        return -3;
    }
    const Py_hash_t h = ((PyASCIIObject *)name)->hash;
    if (h == -1) return -3;      /* not cached: cannot happen for a dict key; the Python path decides */
    for (size_t i = (size_t)h & t->mask;; i = (i + 1) & t->mask) {
        const Slot *s = &t->slots[i];
        if (!s->key) return -1;
        if (s->key == name) return s->row;
        if (s->hash == h && PyUnicode_Compare(s->key, name) == 0) return s->row;
    }
}

/* label_table(labels: list[str]) -> capsule | None (duplicated or non-str labels) */
static PyObject *label_table(PyObject *self, PyObject *arg) {
    if (!PyList_CheckExact(arg)) Py_RETURN_NONE;
    const Py_ssize_t n = PyList_GET_SIZE(arg);
    size_t cap = 16;
    while (cap < (size_t)n * 2 + 2) cap <<= 1;
    Table *t = (Table *)calloc(1, sizeof(Table));
    if (!t) return PyErr_NoMemory();
    t->slots = (Slot *)calloc(cap, sizeof(Slot));
    t->mask = cap - 1;
    t->n_rows = n;
    PyObject *capsule = t->slots ? PyCapsule_New(t, "fp_label_table", table_free) : NULL;
    if (!capsule) {
        free(t->slots);
        free(t);
        return PyErr_NoMemory();
    }
    for (Py_ssize_t r = 0; r < n; ++r) {
        PyObject *name = PyList_GET_ITEM(arg, r);
        if (!PyUnicode_CheckExact(name)) goto decline;
        const Py_hash_t h = PyObject_Hash(name);
        if (h == -1) {
            Py_DECREF(capsule);
            return NULL;
        }
        size_t i = (size_t)h & t->mask;
        for (;; i = (i + 1) & t->mask) {
            Slot *s = &t->slots[i];
            if (!s->key) break;
            if (s->key == name || (s->hash == h && PyUnicode_Compare(s->key, name) == 0)) goto decline;   /* duplicate */
        }
        Py_INCREF(name);
        t->slots[i].key = name;
        t->slots[i].hash = h;
        t->slots[i].row = r;
    }
    return capsule;
decline:
    Py_DECREF(capsule);
    Py_RETURN_NONE;
}

/* sign of a dict value: 1, -1, 0 (ignored); -2: not a plain number */
static int sign_of(PyObject *v) {
    if (PyLong_Check(v)) {
        int overflow = 0;
        const long x = PyLong_AsLongAndOverflow(v, &overflow);
        if (overflow) return 0;
        return (x == 1 || x == 0) ? 1 : (x == -1 ? -1 : 0);
    }
    if (PyFloat_Check(v)) {
        const double x = PyFloat_AS_DOUBLE(v);
        return (x == 1.0 || x == 0.0) ? 1 : (x == -1.0 ? -1 : 0);
    }
    return -2;
}

/* ---- matching: a few threads, read-only --------------------------------------------------
 * The caller holds the GIL for the whole call, so no Python code can change the dicts meanwhile;
 * the workers only READ Python objects (PyDict_Next, cached str hashes, int / float payloads):
 * no reference counts, no allocation, no error state.  Whatever would need more than that (a
 * name whose hash is not cached, a value that is not a plain number) marks the job "unusual"
 * and the caller takes the Python path. */
typedef struct {
    PyObject *dicts;
    const Table *table;
    const Py_ssize_t *offset;     /* first slot of signature s in keys (prefix sums of the dict sizes) */
    int64_t *keys;                /* matched row * 2 + (sign > 0), ascending inside a signature */
    int32_t *n_matched;           /* per signature */
    Py_ssize_t n_sigs, longest;
    int n_threads;
    volatile int unusual;
} Job;

typedef struct {
    Job *job;
    int tid;
} Worker;

static void *match_worker(void *arg) {
    Worker *w = (Worker *)arg;
    Job *job = w->job;
    const Table *table = job->table;
    const Py_ssize_t n_rows = table->n_rows;
    PyObject **items = (PyObject **)malloc(sizeof(PyObject *) * 2 * (size_t)(job->longest + 1));
    uint64_t *seen = (uint64_t *)calloc((size_t)(n_rows / 64 + 1) * 2, sizeof(uint64_t));
    uint64_t *positive = seen ? seen + (n_rows / 64 + 1) : NULL;
    if (!items || !seen) job->unusual = 1;
    for (Py_ssize_t s = w->tid; s < job->n_sigs && !job->unusual; s += job->n_threads) {
        PyObject *d = PyList_GET_ITEM(job->dicts, s), *gene, *value;
        int64_t *keys = job->keys + job->offset[s];
        Py_ssize_t pos = 0, n = 0, n_items = 0;
        /* two passes over the items: the first only collects them and asks for the cache lines
         * of the name objects (their cached hashes) */
        while (PyDict_Next(d, &pos, &gene, &value)) {
            __builtin_prefetch(gene);
            items[2 * n_items] = gene;
            items[2 * n_items + 1] = value;
            ++n_items;
        }
        for (Py_ssize_t it = 0; it < n_items; ++it) {
            const int64_t row = table_find(table, items[2 * it]);
            if (row < 0) {
                if (row == -1) continue;                 /* not on the matrix */
                job->unusual = 1;                        /* not a plain str / hash not cached */
                break;
            }
            const int sg = sign_of(items[2 * it + 1]);
            if (sg == -2) {
                job->unusual = 1;
                break;
            }
            if (sg == 0) continue;
            keys[n++] = row * 2 + (sg > 0);
        }
        if (job->unusual) break;
        sort_keys(keys, n, seen, positive, n_rows);
        job->n_matched[s] = (int32_t)n;
    }
    free(items);
    free(seen);
    return NULL;
}

/* pack(dicts: list[dict], table: capsule of label_table(), min_genes: float)
 *   -> None | (kept int64[], ptr int32[], rows int32[], signs int8[], counts int64[]) as bytearrays */
static PyObject *pack(PyObject *self, PyObject *args) {
    PyObject *dicts, *capsule;
    double min_genes;
    if (!PyArg_ParseTuple(args, "O!Od", &PyList_Type, &dicts, &capsule, &min_genes)) return NULL;
    const Table *table = (const Table *)PyCapsule_GetPointer(capsule, "fp_label_table");
    if (!table) return NULL;
    const Py_ssize_t n_sigs = PyList_GET_SIZE(dicts);
    Py_ssize_t *offset = (Py_ssize_t *)malloc(sizeof(Py_ssize_t) * (size_t)(n_sigs + 1));
    if (!offset) return PyErr_NoMemory();
    Py_ssize_t n_entries = 0, longest = 0;
    for (Py_ssize_t s = 0; s < n_sigs; ++s) {
        PyObject *d = PyList_GET_ITEM(dicts, s);
        if (!PyDict_CheckExact(d)) {
            free(offset);
            Py_RETURN_NONE;
        }
        const Py_ssize_t n = PyDict_GET_SIZE(d);
        offset[s] = n_entries;
        n_entries += n;
        if (n > longest) longest = n;
    }
    offset[n_sigs] = n_entries;
    if (n_entries >= (Py_ssize_t)INT32_MAX) {
        free(offset);
        PyErr_SetString(PyExc_ValueError, "signature membership table exceeds int32 indexing");
        return NULL;
    }
    Job job;
    job.dicts = dicts;
    job.table = table;
    job.offset = offset;
    job.keys = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_entries + 1));
    job.n_matched = (int32_t *)calloc((size_t)(n_sigs + 1), sizeof(int32_t));
    job.n_sigs = n_sigs;
    job.longest = longest;
    job.n_threads = n_entries >= 16384 ? MATCH_THREADS : 1;
    job.unusual = 0;
    PyObject *b_kept = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)sizeof(int64_t) * n_sigs);
    PyObject *b_ptr = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)sizeof(int32_t) * (n_sigs + 1));
    PyObject *b_rows = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)sizeof(int32_t) * n_entries);
    PyObject *b_signs = PyByteArray_FromStringAndSize(NULL, n_entries);
    PyObject *b_counts = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)sizeof(int64_t) * n_sigs);
    PyObject *result = NULL;
    if (!job.keys || !job.n_matched || !b_kept || !b_ptr || !b_rows || !b_signs || !b_counts) {
        PyErr_NoMemory();
        goto done;
    }
    {
        Worker workers[MATCH_THREADS];
        pthread_t threads[MATCH_THREADS];
        int started[MATCH_THREADS] = {0};
        for (int t = 0; t < job.n_threads; ++t) {
            workers[t].job = &job;
            workers[t].tid = t;
        }
        for (int t = 1; t < job.n_threads; ++t)
            started[t] = pthread_create(&threads[t], NULL, match_worker, &workers[t]) == 0;
        match_worker(&workers[0]);
        for (int t = 1; t < job.n_threads; ++t) {
            if (started[t]) {
                pthread_join(threads[t], NULL);
            } else {
                match_worker(&workers[t]);               /* no thread to be had: do its share here */
            }
        }
        if (job.unusual) goto done;
        int64_t *kept = (int64_t *)PyByteArray_AS_STRING(b_kept);
        int32_t *ptr = (int32_t *)PyByteArray_AS_STRING(b_ptr);
        int32_t *rows = (int32_t *)PyByteArray_AS_STRING(b_rows);
        int8_t *signs = (int8_t *)PyByteArray_AS_STRING(b_signs);
        int64_t *counts = (int64_t *)PyByteArray_AS_STRING(b_counts);
        Py_ssize_t n_kept = 0, fill = 0;
        ptr[0] = 0;
        for (Py_ssize_t s = 0; s < n_sigs; ++s) {
            const Py_ssize_t n = job.n_matched[s];
            if (n == 0 || (double)n < min_genes) continue;
            const int64_t *keys = job.keys + offset[s];
            for (Py_ssize_t i = 0; i < n; ++i) {
                rows[fill + i] = (int32_t)(keys[i] >> 1);
                signs[fill + i] = (keys[i] & 1) ? 1 : -1;
            }
            fill += n;
            kept[n_kept] = s;
            counts[n_kept] = n;
            ptr[++n_kept] = (int32_t)fill;
        }
        if (PyByteArray_Resize(b_kept, (Py_ssize_t)sizeof(int64_t) * n_kept) < 0 ||
            PyByteArray_Resize(b_ptr, (Py_ssize_t)sizeof(int32_t) * (n_kept + 1)) < 0 ||
            PyByteArray_Resize(b_rows, (Py_ssize_t)sizeof(int32_t) * fill) < 0 ||
            PyByteArray_Resize(b_signs, fill) < 0 ||
            PyByteArray_Resize(b_counts, (Py_ssize_t)sizeof(int64_t) * n_kept) < 0)
            goto done;
        result = PyTuple_Pack(5, b_kept, b_ptr, b_rows, b_signs, b_counts);
    }
done:
    free(offset);
    free(job.keys);
    free(job.n_matched);
    Py_XDECREF(b_kept);
    Py_XDECREF(b_ptr);
    Py_XDECREF(b_rows);
    Py_XDECREF(b_signs);
    Py_XDECREF(b_counts);
    if (!result && !PyErr_Occurred()) Py_RETURN_NONE;                         /* unusual input */
    return result;
}

static PyMethodDef methods[] = {
    {"label_table", label_table, METH_O, "label_table(labels) -> capsule | None (duplicated / non-str labels)"},
    {"pack", pack, METH_VARARGS, "pack(dicts, table, min_genes) -> None | (kept, ptr, rows, signs, counts) bytearrays"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fp_pack", "signature -> CSR packing", -1, methods};

PyMODINIT_FUNC PyInit__fp_pack(void) { return PyModule_Create(&module); }
