# -*- coding: utf-8 -*-
"""Device plumbing between the Python mirror of the reference API and the
C ABI: PyTorch tensors carry the buffers (allocation, H2D/D2H copies, the
current stream); every computation is a call into libfastproject_b200.so.

There is no CPU path here: a missing GPU or library raises.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from . import _native

_ROW_ALIGN = 32          # leading dimension of every [signatures x cells] matrix, in elements


def require_cuda():
    if not torch.cuda.is_available():
        raise _native.NativeError("fastproject_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def padded(n):
    return (int(n) + _ROW_ALIGN - 1) // _ROW_ALIGN * _ROW_ALIGN


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return 0 if t is None else t.data_ptr()


def alloc_matrix(g, n, device):
    """[g x n] float64 view of a buffer whose row pitch is a multiple of 16 bytes (an even number of
    doubles): what the TMA-staged scoring kernel needs of the matrices it streams."""
    ld = int(n) + (int(n) & 1)
    return torch.empty((int(g), ld), dtype=torch.float64, device=device)[:, :int(n)]


def row_major(t):
    """CUDA float64 matrix with unit column stride (any row pitch): as is; anything else: a
    compact copy.  (``.contiguous()`` would repack a padded matrix to an odd pitch.)"""
    t = t.to(torch.float64)
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1]:
        return t
    return t.contiguous()


TC_MAX_CELLS = 4000000   # FP_PRECISION_TC: three balanced base-256 digits of the centred values


def default_precision():
    return os.environ.get("FASTPROJECT_B200_PRECISION", "auto")


def resolve_precision(precision, n):
    """"auto": the tensor-core path when the shape is supported, else FP64."""
    precision = precision or default_precision()
    if precision == "auto":
        return "tc" if n <= TC_MAX_CELLS else "fp64"
    return precision


# --------------------------------------------------------------------------
# host -> device
# --------------------------------------------------------------------------
_SMALL_UPLOAD = 16 << 20       # host arrays up to this size bypass the copy engine
_small_pool = []               # [pinned uint8 tensor, event of its last use]


def _pinned_slot(nbytes):
    """A pinned staging buffer no kernel is reading any more (or a new one)."""
    for slot in _small_pool:
        if slot[0].numel() >= nbytes and slot[1].query():
            return slot
    size = max(1 << 20, 1 << (int(nbytes) - 1).bit_length())
    slot = [torch.empty((size,), dtype=torch.uint8, pin_memory=True), torch.cuda.Event()]
    _small_pool.append(slot)
    return slot


def _pull_small(a, dev):
    """NumPy array (C-contiguous, <= _SMALL_UPLOAD bytes) -> CUDA tensor of the same dtype,
    fetched by a kernel from a pinned staging buffer (``fp_pull_host``): the per-call tables
    of the path do not queue behind a multi-gigabyte prefetch on the copy engine."""
    nbytes = a.nbytes
    out = torch.empty(tuple(a.shape), dtype=torch.from_numpy(np.empty(0, dtype=a.dtype)).dtype, device=dev)
    if nbytes == 0:
        return out
    slot = _pinned_slot(nbytes)
    slot[0].numpy()[:nbytes] = a.reshape(-1).view(np.uint8)
    _native.call("fp_pull_host", ptr(out), ptr(slot[0]), nbytes, stream_ptr())
    slot[1].record()
    return out


def to_device(arr, dtype=torch.float64):
    """NumPy array or CPU/GPU torch tensor -> contiguous CUDA tensor."""
    dev = require_cuda()
    if isinstance(arr, torch.Tensor):
        t = arr
    else:
        a = np.asarray(arr)
        if a.dtype == np.bool_:
            a = a.astype(np.float64)
        if not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a)
        if a.nbytes <= _SMALL_UPLOAD and a.dtype.kind in "fiu" and a.dtype.byteorder in "=|<":
            return _pull_small(a, dev).to(dtype=dtype).contiguous()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)     # read-only arrays: we only read them
            t = torch.from_numpy(a)
    if t.dim() == 2 and dtype == torch.float64 and not t.is_cuda and (t.shape[1] & 1):
        out = alloc_matrix(t.shape[0], t.shape[1], dev)           # odd row length: even pitch on the device
        out.copy_(t, non_blocking=True)
        return out
    t = t.to(device=dev, dtype=dtype, non_blocking=True)
    return t.contiguous()


def to_device_replicated(arr):
    """Upload a matrix that is identical on every rank: rank r copies rows
    [r*rows_per, (r+1)*rows_per) host -> device, then one NCCL all-gather assembles the whole
    matrix on every GPU.  The PCIe traffic per GPU is 1/world of the matrix."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return to_device(arr)
    dev = require_cuda()
    world, rank = dist.get_world_size(), dist.get_rank()
    if isinstance(arr, torch.Tensor):
        host = arr
    else:
        a = np.asarray(arr)
        if a.dtype == np.bool_:
            a = a.astype(np.float64)
        if not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a)
        host = torch.from_numpy(a)
    rows, cols = host.shape
    rows_per = (rows + world - 1) // world
    ld = cols + (cols & 1)                                  # even row pitch (TMA-staged scoring)
    full = torch.empty((world * rows_per, ld), dtype=torch.float64, device=dev)
    lo, hi = min(rows, rank * rows_per), min(rows, (rank + 1) * rows_per)
    mine = torch.zeros((rows_per, ld), dtype=torch.float64, device=dev)
    if hi > lo:
        mine[:hi - lo, :cols].copy_(host[lo:hi], non_blocking=True)
    from . import distributed as D
    D.all_gather_into(full, mine)
    return full[:rows, :cols]


_PINNED_CHUNK = 32 << 20
_pinned = []
_d2h_stream = None


def to_host_async(dev_tensor):
    """Start the copy of a CUDA float64 matrix into a new pinned host tensor on a side stream
    (it waits for what the current stream has queued so far, then runs beside whatever comes
    next).  Returns (host tensor, event): the host tensor is valid after ``event.synchronize()``."""
    global _d2h_stream
    if _d2h_stream is None:
        _d2h_stream = torch.cuda.Stream(device=dev_tensor.device)
    ready = torch.cuda.Event()
    ready.record()
    _d2h_stream.wait_event(ready)
    host = torch.empty(tuple(dev_tensor.shape), dtype=dev_tensor.dtype, pin_memory=True)
    done = torch.cuda.Event()
    with torch.cuda.stream(_d2h_stream):
        host.copy_(dev_tensor, non_blocking=True)
        done.record()
    dev_tensor.record_stream(_d2h_stream)
    return host, done


_PINNED_RESULT_MAX = 1 << 30


def to_host(dev_tensor):
    """CUDA tensor (any strides) -> NumPy array.  Results up to 1 GiB land directly in a pinned
    host tensor that the returned array views (one DMA at PCIe speed, no second host copy; the
    pinned block goes back to torch's host allocator when the array is released).  Larger ones
    are staged through two reusable pinned buffers into pageable memory -- a plain ``.cpu()``
    into pageable memory runs at a few GB/s."""
    if dev_tensor.dtype != torch.float64 or dev_tensor.numel() * 8 < (1 << 20):
        return dev_tensor.cpu().numpy()
    if dev_tensor.numel() * 8 <= _PINNED_RESULT_MAX:
        host = torch.empty(tuple(dev_tensor.shape), dtype=torch.float64, pin_memory=True)
        host.copy_(dev_tensor, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy()
    t = dev_tensor.contiguous()
    out = np.empty(tuple(t.shape), dtype=np.float64)
    flat_dev = t.view(-1)
    flat_out = out.reshape(-1)
    n = flat_dev.numel()
    step = _PINNED_CHUNK // 8
    while len(_pinned) < 2:
        _pinned.append((torch.empty((step,), dtype=torch.float64, pin_memory=True), torch.cuda.Event()))
    pending = []
    for k, lo in enumerate(range(0, n, step)):
        hi = min(n, lo + step)
        buf, ev = _pinned[k & 1]
        if len(pending) == 2:                       # buffer k&1 still holds chunk k-2: drain it first
            plo, phi, pbuf, pev = pending.pop(0)
            pev.synchronize()
            flat_out[plo:phi] = pbuf[:phi - plo].numpy()
        buf[:hi - lo].copy_(flat_dev[lo:hi], non_blocking=True)
        ev.record()
        pending.append((lo, hi, buf, ev))
    for plo, phi, pbuf, pev in pending:
        pev.synchronize()
        flat_out[plo:phi] = pbuf[:phi - plo].numpy()
    return out


class _ExpressionCache(object):
    """Explicit residency of genes x cells host matrices in HBM.

    Nothing is cached implicitly: a host matrix handed to ``calculate_sig_scores`` is
    uploaded by that call and dropped with it, unless the caller has declared it resident --
    ``fastproject_b200.prefetch(x, w, ...)`` (asynchronous upload, stays until ``evict``) or
    the ``fastproject_b200.resident(x, w, ...)`` context manager.  The declaration is the
    caller's promise that the host array is not modified while it is resident; entries are
    found by the identity of the array object (the object is kept alive, so its address
    cannot be recycled) plus its data pointer, shape, strides and dtype.  (Round 1 guessed
    identity from the address and ~4 000 sampled elements, which missed in-place edits and
    recycled addresses.)  In a multi-process run every rank must make the same declarations:
    a replicated upload contains a collective."""

    def __init__(self, capacity=8):
        self.capacity = capacity
        self.entries = []          # list of [key, host object, device tensor, ready event | None]

    @staticmethod
    def _key(a):
        if isinstance(a, torch.Tensor):
            return ("t", id(a), a.data_ptr(), tuple(a.shape), tuple(a.stride()), str(a.dtype))
        if not isinstance(a, np.ndarray):
            return None
        return ("n", id(a), a.__array_interface__["data"][0], a.shape, a.strides, a.dtype.str)

    def _find(self, key):
        if key is None:
            return -1
        for i, entry in enumerate(self.entries):
            if entry[0] == key:
                return i
        return -1

    def get(self, a, replicated=False):
        """Device copy of ``a``: the resident one if the caller declared it, else a fresh
        upload that nobody remembers.  ``replicated``: every rank of the process group holds
        the same host matrix; each rank then uploads 1/world of the rows over its own PCIe
        link and the rows are all-gathered over NVLink."""
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return row_major(a)
        i = self._find(self._key(a))
        if i >= 0:
            entry = self.entries[i]
            if entry[3] is not None:                 # uploaded on the copy stream: order after it
                torch.cuda.current_stream().wait_event(entry[3])
                entry[3] = None
            return entry[2]
        return (to_device_replicated if replicated else to_device)(a)

    def prefetch(self, a, replicated=False):
        """Declare a host matrix resident and start its upload on a side stream; returns at
        once.  A later ``get`` of the same array object waits for the copy instead of issuing
        one.  Lets the upload of the NEXT data set overlap the kernels of the current one
        (pinned host memory).

        The host -> device copy engine serves transfers in submission order (measured: the
        kilobyte-sized uploads of a running step -- index tables, coordinates -- waited ~90 ms
        behind a queued 4.8 GB prefetch and stalled the compute stream; feeding the engine in
        paced pieces from a Python thread halved the copy rate).  The small tables of the path
        therefore do not use the copy engine at all (``_pull_small``), and a prefetch can be
        issued at any point -- best at the very start of the step it overlaps."""
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return
        dev = require_cuda()
        key = self._key(a)
        if key is None:
            raise TypeError("prefetch / resident need a NumPy array or a torch tensor (the object itself "
                            "identifies the resident copy), got %r" % type(a))
        if self._find(key) >= 0:
            return
        if isinstance(a, torch.Tensor):
            host = a
        else:
            arr = a
            if arr.dtype == np.bool_:
                arr = arr.astype(np.float64)
            if not arr.flags["C_CONTIGUOUS"]:
                arr = np.ascontiguousarray(arr)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", UserWarning)
                host = torch.from_numpy(arr)
        global _copy_stream
        if _copy_stream is None:
            _copy_stream = torch.cuda.Stream(device=dev)
        _copy_stream.wait_stream(torch.cuda.current_stream())      # the buffer may be recycled memory
        with torch.cuda.stream(_copy_stream):
            if replicated:
                # row shard over this rank's PCIe link + NCCL all-gather, all on the copy stream;
                # every rank issues its collectives in the same program order
                t = to_device_replicated(host)
            else:
                t = alloc_matrix(host.shape[0], host.shape[1], dev)
                t.copy_(host, non_blocking=True)
            done = torch.cuda.Event()
            done.record(_copy_stream)
        t.record_stream(torch.cuda.current_stream())
        self.entries.append([key, a, t, done])
        while len(self.entries) > self.capacity:
            self.entries.pop(0)

    def evict(self, a):
        i = self._find(self._key(a))
        if i >= 0:
            self.entries.pop(i)

    def clear(self):
        self.entries = []


_copy_stream = None
expression_cache = _ExpressionCache()


def prefetch(*arrays, **kw):
    """Public hook: declare host matrices (expression, weights, zero mask) resident and start
    uploading them; later ``calculate_sig_scores`` calls on the same array OBJECTS use the
    device copy.  The caller must not modify the arrays until ``evict`` (or ``resident``'s
    exit).  ``replicated=True``: the cooperative multi-GPU upload (same arrays on every rank;
    every rank must make the same call)."""
    for a in arrays:
        if a is not None:
            expression_cache.prefetch(a, kw.get("replicated", False))


def evict(*arrays):
    """Drop the resident device copies of these host matrices."""
    for a in arrays:
        if a is not None:
            expression_cache.evict(a)


class resident(object):
    """``with fastproject_b200.resident(data.base, data.weights, zero_locations): ...`` -- the
    matrices are uploaded once for all scoring calls inside the block (the pipeline scores the
    real and the random signatures on the same data, Pipelines.py:204,229) and released at
    its end."""

    def __init__(self, *arrays, **kw):
        self.arrays = [a for a in arrays if a is not None]
        self.kw = kw

    def __enter__(self):
        prefetch(*self.arrays, **self.kw)
        return self

    def __exit__(self, *exc):
        evict(*self.arrays)
        return False


_clear_hooks = []       # other modules' device-side state (Signatures' permutation background)


def clear_device_cache():
    expression_cache.clear()
    _ws_cache.clear()
    for hook in _clear_hooks:
        hook()


# --------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------
def score_signatures(method, x, w, z, sig_ptr, sig_gene, sig_sign, n_sigs, nnz=None):
    """x, w, z: G x N CUDA float64 (w / z may be None); CSR arrays on device.
    ``nnz``: number of CSR entries when the host knows it (it always does: it built the CSR);
    lets the library pick the TMA-staged kernel without reading ``sig_ptr`` back.
    Returns a [n_sigs x ld] CUDA tensor (ld = padded N)."""
    g, n = x.shape
    ld = padded(n)
    out = torch.zeros((n_sigs, ld), dtype=torch.float64, device=x.device)
    if n_sigs == 0:
        return out
    # the kernels address X, W and Z with ONE row pitch
    pitch = max(m.stride(0) for m in (x, w, z) if m is not None)
    pitch += pitch & 1

    def repitch(m):
        if m is None or m.stride(0) == pitch:
            return m
        buf = torch.empty((g, pitch), dtype=torch.float64, device=m.device)[:, :n]
        buf.copy_(m)
        return buf
    x, w, z = repitch(x), repitch(w), repitch(z)
    mu = None
    if method == "imputed":
        mu = torch.empty((g,), dtype=torch.float64, device=x.device)
        _native.call("fp_gene_detected_mean", ptr(x), ptr(z), g, n, x.stride(0), ptr(mu), stream_ptr())
    nnz = int(sig_gene.shape[0]) if nnz is None else int(nnz)
    nbytes = _native.query("fp_score_workspace_bytes", g, n_sigs, nnz)
    ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=x.device)
    _native.call("fp_score_signatures", _native.METHOD_CODES[method], ptr(x), ptr(w), ptr(z),
                 g, n, x.stride(0), ptr(sig_ptr), ptr(sig_gene), ptr(sig_sign), n_sigs, nnz, ptr(mu), ptr(out), ld,
                 ptr(ws), nbytes, stream_ptr())
    return out


def rank_rows(scores, n, method=_native.RANK_AVERAGE):
    """scores: [K x ld] CUDA float64 -> ranks of every row (average or min tie rule), same shape."""
    k, ld = scores.shape
    out = torch.zeros_like(scores)
    if k == 0:
        return out
    nbytes = _native.query("fp_rank_workspace_bytes", k, n, ld)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=scores.device)
    _native.call("fp_rank_columns_ex", ptr(scores), k, n, ld, ptr(out), ptr(ws), nbytes, method, stream_ptr())
    return out


def normalize_rows(x):
    """[G x N] CUDA float64 -> per-row z-scores (NumPy summation order), new tensor."""
    g, n = x.shape
    out = alloc_matrix(g, n, x.device)
    nbytes = _native.query("fp_normalize_rows_workspace_bytes", n)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    _native.call("fp_normalize_rows", ptr(x), g, n, x.stride(0), ptr(out), out.stride(0), ptr(ws), nbytes,
                 stream_ptr())
    return out


def compute_weights(x, params, p_nd=None):
    """[G x N] CUDA float64 expression, [4 x N] logistic parameters (x0, a, L, S), optional
    [G x N] probability of non-detection (None: the zero mask) -> [G x N] weights."""
    g, n = x.shape
    out = alloc_matrix(g, n, x.device)
    params = params.contiguous()
    nbytes = _native.query("fp_compute_weights_workspace_bytes", n)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    _native.call("fp_compute_weights", ptr(x), ptr(p_nd), ptr(params), g, n, x.stride(0),
                 0 if p_nd is None else p_nd.stride(0), ptr(out), out.stride(0), ptr(ws), nbytes, stream_ptr())
    return out


def normalize_cols(x):
    """[G x N] CUDA float64 -> per-column z-scores (sequential row order), new tensor."""
    g, n = x.shape
    out = alloc_matrix(g, n, x.device)
    _native.call("fp_normalize_cols", ptr(x), g, n, x.stride(0), ptr(out), out.stride(0), stream_ptr())
    return out


class PackedValues(object):
    """The values of ``consistency`` in the form the tensor-core mode contracts them
    (``fp_values_pack``: value range, digit planes in the caller's cell order, refusal flag).
    ``sigs_vs_projections`` tests one rank matrix against every projection: packed once, a
    projection only permutes the planes into its own cell order."""

    PACK_MAX_BYTES = 2 << 30

    def __init__(self, buf, k, n):
        self.buf, self.k, self.n = buf, int(k), int(n)


def pack_values(ranks, n, precision=None):
    """``PackedValues`` of the [K x ld] CUDA float64 matrix, or None when the mode in force is not
    the tensor-core one (or the planes would be too large to keep beside the matrix)."""
    k = ranks.shape[0]
    if not k or resolve_precision(precision, n) != "tc":
        return None
    nbytes = _native.query("fp_values_pack_bytes", n, k)
    if not nbytes or nbytes > PackedValues.PACK_MAX_BYTES:
        return None
    buf = torch.empty((nbytes,), dtype=torch.uint8, device=ranks.device)
    _native.call("fp_values_pack", ptr(ranks), k, n, ranks.stride(0), ptr(buf), nbytes, stream_ptr())
    return PackedValues(buf, k, n)


def consistency(coords, sigma_sq, ranks, n, out=None, output_kind=_native.OUT_ABS_ERROR,
                precision=None, slot="default", any_cell_order=False, packed=None):
    """coords: [2 x N] CUDA float64; ranks [K x ld]; sigma_sq float or [N] tensor.
    ``any_cell_order``: the caller only reduces the rows over the cells (median): the columns may
    come in the library's own cell order (saves scattered writes on the tensor-core path).
    ``packed``: ``pack_values(ranks, n)`` of the same matrix -- the call then reads the packed
    form instead of ``ranks`` (same result, bit for bit).
    Asynchronous.  The tensor-core mode refuses values it cannot represent exactly (the output
    is then NaN): ``consistency_refused`` tells, after the calls on a workspace ``slot``."""
    k, ld = ranks.shape
    if out is None:
        out = torch.empty((k, ld), dtype=torch.float64, device=ranks.device)
    mode = _native.PRECISION_CODES[resolve_precision(precision, n)]
    vec = sigma_sq if isinstance(sigma_sq, torch.Tensor) else None
    scalar = 0.0 if vec is not None else float(sigma_sq)
    nbytes = _native.query("fp_consistency_workspace_bytes", n, k, mode)
    ws, ws_ptr = _workspace(nbytes, ranks.device, slot)
    if any_cell_order:
        output_kind |= _native.OUT_ANY_CELL_ORDER
    if packed is not None and mode == _native.PRECISION_CODES["tc"]:
        if (packed.k, packed.n) != (k, int(n)):
            raise ValueError("packed values of a %d x %d matrix given for %d x %d" % (packed.k, packed.n, k, n))
        _native.call("fp_consistency_packed", ptr(coords), ptr(vec), scalar, ptr(packed.buf), n, k,
                     ptr(out), out.stride(0), output_kind, ws_ptr, nbytes, stream_ptr())
        return out
    _native.call("fp_consistency", ptr(coords), ptr(vec), scalar, ptr(ranks), n, k, ld,
                 ptr(out), out.stride(0), output_kind, mode, ws_ptr, nbytes, stream_ptr())
    return out


def consistency_refused(ranks, n, precision=None, slot="default"):
    """True if the last ``consistency`` call on this workspace ``slot`` (same shape, same
    precision) refused its values -- FP_PRECISION_TC holds half-integers of bounded range
    only.  Synchronises the stream (call it once, after the projection loop)."""
    k = ranks.shape[0]
    mode = _native.PRECISION_CODES[resolve_precision(precision, n)]
    if not k or mode != _native.PRECISION_CODES["tc"]:
        return False
    nbytes = _native.query("fp_consistency_workspace_bytes", n, k, mode)
    ws, ws_ptr = _workspace(nbytes, ranks.device, slot)
    rc = _native.load().fp_consistency_status(ws_ptr, n, k, mode, stream_ptr())
    if rc == _native.FP_EUNSUPPORTED:
        return True
    if rc != 0:
        raise _native.NativeError("fp_consistency_status failed (%d): %s"
                                  % (rc, (_native.load().fp_last_error() or b"").decode()))
    return False


_WS_ALIGN = 1024
_ws_cache = dict()


def _workspace(nbytes, device, slot="default"):
    """Scratch buffer for one call (1024-byte aligned: the TMA / UMMA tiles of
    the tensor-core path need it).  The largest buffer per device is kept and
    reused; all work of a process goes to one stream at a time."""
    if not nbytes:
        return None, 0
    key = (str(device), slot)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes + _WS_ALIGN:
        buf = torch.empty((nbytes + _WS_ALIGN,), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    base = buf.data_ptr()
    return buf, (base + _WS_ALIGN - 1) // _WS_ALIGN * _WS_ALIGN


def row_medians(values, n):
    k, ld = values.shape
    out = torch.empty((k,), dtype=torch.float64, device=values.device)
    _native.call("fp_row_medians", ptr(values), k, n, ld, ptr(out), stream_ptr())
    return out


_median_stream = None


class MedianOverlap(object):
    """Runs ``fp_row_medians`` of one projection on a side stream while the main stream already
    generates the weight planes of the next one (the median kernel is a bandwidth-bound sweep,
    the weight planes are FP64-bound: they share the SMs well).  Two dissimilarity buffers
    alternate; a buffer is handed back to the contraction only after the median that read it.
    Same kernels, same arguments, same results -- only the stream differs.  Above
    ``DOUBLE_BUFFER_MAX`` bytes per buffer (a million cells x thousands of signatures) there is
    ONE buffer: the next contraction then waits for the median, which is a few per cent of it."""

    DOUBLE_BUFFER_MAX = 8 << 30

    def __init__(self, rows, ld, device):
        global _median_stream
        if _median_stream is None:
            _median_stream = torch.cuda.Stream(device=device)
        self.side = _median_stream
        n_bufs = 2 if int(rows) * int(ld) * 8 <= self.DOUBLE_BUFFER_MAX else 1
        self.bufs = [torch.empty((rows, ld), dtype=torch.float64, device=device) for _ in range(n_bufs)]
        self.done = [None] * n_bufs
        self.turn = 0

    def buffer(self, rows):
        """The dissimilarity buffer for the next contraction (main stream)."""
        k = self.turn % len(self.bufs)
        if self.done[k] is not None:
            torch.cuda.current_stream().wait_event(self.done[k])
        return self.bufs[k][:rows]

    def median_into(self, dis, n, dest):
        """dest[...] = row medians of ``dis`` (just written by the main stream), on the side stream."""
        k = self.turn % len(self.bufs)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dest.copy_(row_medians(dis, n))
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[k] = ev
        self.turn += 1

    def join(self):
        """Main stream waits for every median issued so far."""
        torch.cuda.current_stream().wait_stream(self.side)


def background_pvalues(med, sig_group, rnd_med, rnd_order, group_ptr):
    """All arguments CUDA tensors (float64 / int32).  Returns (p, mu, sigma)."""
    k = med.shape[0]
    n_groups = group_ptr.shape[0] - 1
    p = torch.empty((k,), dtype=torch.float64, device=med.device)
    mu = torch.empty((max(n_groups, 1),), dtype=torch.float64, device=med.device)
    sd = torch.empty((max(n_groups, 1),), dtype=torch.float64, device=med.device)
    _native.call("fp_background_pvalues", ptr(med), ptr(sig_group), k, ptr(rnd_med), ptr(rnd_order),
                 ptr(group_ptr), n_groups, ptr(p), ptr(mu), ptr(sd), stream_ptr())
    return p, mu, sd
