"""Host side of the B200 clustering step: owns buffers through PyTorch, calls the C ABI.

`Engine.cluster(table, params)` is the fused entry the benchmark and the drop-in functions of
fslr_b200.cluster use (SURVEY.md §8b `cluster_table`).  Threshold tables that depend on floating point
(`umax`, `1 - qlen_diff`, `1 - n_alignment_diff`) are computed here with Python floats — the reference's own
arithmetic (cluster.py:170,179,181,218-219) — so the device only ever compares integers.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _native
from .table import ClusterParams, ColumnarTable

_COLS = ("read_id", "chrom", "rstart", "rend", "aln_size", "qstart", "qend", "n_alignments")


def umax_table(cutoffs, n_max=_native.MAX_FILLINGS):
    """umax[n] = largest union u >= n with n/u >= cutoff(n) in float64; n-1 when none (cluster.py:165-170,218-219)."""
    out = [0]
    for n in range(1, n_max + 1):
        t = cutoffs[n - 1] if n - 1 < len(cutoffs) else cutoffs[-1]
        if t <= 0:
            out.append(2**31 - 1)
            continue
        u = min(int(n / t) + 2, 2**31 - 1)
        while u >= n and not (n / u >= t):
            u -= 1
        out.append(u if u >= n else n - 1)
    return out


@dataclass
class ClusterResult:
    cluster: np.ndarray        # int32 per read id: the `cluster` value main.py:334-342 writes for that read's rows
    n_reads: np.ndarray        # int32 per read id
    no_clusters: bool          # main.py:247-249 early return ("No clusters were found.")
    stats: dict


class DeviceTable:
    """The table's columns resident in HBM (torch tensors own the memory)."""

    def __init__(self, table: ColumnarTable, device, order=None):
        self.n_rows, self.n_reads = table.n_rows, table.n_reads
        self.cols = {k: torch.from_numpy(np.ascontiguousarray(getattr(table, k), dtype=np.int32)).to(device) for k in _COLS}
        self.order = None
        if order is not None:
            self.order = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int32)).to(device)
        self.out_cluster = torch.empty(max(self.n_reads, 1), dtype=torch.int32, device=device)
        self.out_n_reads = torch.empty(max(self.n_reads, 1), dtype=torch.int32, device=device)


class PinnedTable:
    """The table's columns in pinned host memory (the e2e path copies them in every call).  compact=True keeps `chrom` as
    uint8 and `n_alignments` as uint16 when they fit, and leaves out `aln_size` when it equals qend - qstart on every row (what
    collect_mapping_info.py:88 writes), and `read_id` when the rows of every read are contiguous and in id order (run lengths
    travel instead): 19 instead of 32 bytes per row over PCIe, widened / derived on the device."""

    def __init__(self, table: ColumnarTable, order=None, compact=False):
        self.n_rows, self.n_reads = table.n_rows, table.n_reads
        self.cols, self.narrow = {}, {}
        if compact and self.n_rows > 0 and table.n_chrom <= 256 and int(np.max(table.n_alignments)) < 65536 and int(np.min(table.n_alignments)) >= 0:
            self.narrow = {"chrom": torch.from_numpy(np.ascontiguousarray(table.chrom, dtype=np.uint8)).pin_memory(),
                           "n_alignments": torch.from_numpy(np.ascontiguousarray(table.n_alignments).astype(np.uint16).view(np.int16)).pin_memory()}
        self.aln_is_qspan = bool(compact and self.n_rows > 0 and np.array_equal(
            np.asarray(table.aln_size, dtype=np.int64), np.asarray(table.qend, dtype=np.int64) - np.asarray(table.qstart, dtype=np.int64)))
        self.rows_per_read = None
        if compact and self.n_rows > 0:                      # rows of a read contiguous, reads in id order: send run lengths
            rid = np.asarray(table.read_id)
            cnt = np.bincount(rid, minlength=self.n_reads)
            if cnt.min(initial=1) >= 1 and cnt.max(initial=0) <= 255 and np.array_equal(rid, np.repeat(np.arange(self.n_reads, dtype=rid.dtype), cnt)):
                self.rows_per_read = torch.from_numpy(cnt.astype(np.uint8)).pin_memory()
        for k in _COLS:
            if k in self.narrow or (k == "aln_size" and self.aln_is_qspan) or (k == "read_id" and self.rows_per_read is not None):
                continue
            t = torch.empty(max(self.n_rows, 1), dtype=torch.int32).pin_memory()
            t[:self.n_rows] = torch.from_numpy(np.ascontiguousarray(getattr(table, k), dtype=np.int32))
            self.cols[k] = t
        self.order = None
        if order is not None:
            self.order = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int32)).pin_memory()
        self.out_cluster = torch.empty(max(self.n_reads, 1), dtype=torch.int32).pin_memory()
        self.out_n_reads = torch.empty(max(self.n_reads, 1), dtype=torch.int32).pin_memory()

    @property
    def h2d_bytes(self):
        per_row = 4 * len(self.cols) + (3 if self.narrow else 0)
        return per_row * self.n_rows + (4 * int(self.order.numel()) if self.order is not None else 0) + \
            (self.n_reads if self.rows_per_read is not None else 0)

    @property
    def d2h_bytes(self):
        return 8 * self.n_reads


class _DevView:
    """Device memory owned by the library, exposed to torch through __cuda_array_interface__ (int32)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


class Engine:
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("fslr_b200: no CUDA device — this package has no CPU fallback")
        self.lib = _native.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)            # make sure the primary context exists
        self.ctx = C.c_void_p()
        rc = self.lib.fslrc_create(device, C.byref(self.ctx))
        if rc != 0:
            raise _native.FslrError(rc, "fslrc_create failed")

    def pinned_bytes(self, n):
        """A pinned uint8 staging buffer of at least n bytes, kept and reused (pinning hundreds of MB costs more than the
        copy it speeds up).  The returned view is valid until the next call."""
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < n:
            self._pinned = buf = torch.empty(max(int(n * 1.25), 1 << 20), dtype=torch.uint8).pin_memory()
        return buf[:max(n, 1)]

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.fslrc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- argument marshalling
    def _params(self, table, params: ClusterParams):
        p = _native.Params()
        p.overlap = float(params.overlap)
        p.qlen_c = 1 - params.qlen_diff                 # cluster.py:179
        p.naln_c = 1 - params.n_alignment_diff          # cluster.py:181
        um = umax_table(list(params.jaccard_cutoffs))
        for i, v in enumerate(um):
            p.umax[i] = int(v)
        p.edge_threshold = int(max(min(params.edge_threshold, 2**62), -2**62))
        self._clen = np.ascontiguousarray(table_chrom_len(table), dtype=np.int64)
        m = params.chrom_masked if params.chrom_masked is not None else np.zeros(len(self._clen), np.uint8)
        self._cmask = np.ascontiguousarray(m, dtype=np.uint8)
        p.n_chrom = len(self._clen)
        p.chrom_len = self._clen.ctypes.data if len(self._clen) else None
        p.chrom_masked = self._cmask.ctypes.data if len(self._cmask) else None
        p.mask_subtelomere = int(bool(params.mask_subtelomere))
        p.subtel = int(params.subtel)
        return p

    @staticmethod
    def _table(buf):
        t = _native.Table()
        t.n_rows, t.n_reads = buf.n_rows, buf.n_reads
        for k in _COLS:
            setattr(t, k, buf.cols[k].data_ptr() if k in buf.cols else None)
        narrow = getattr(buf, "narrow", None) or {}
        t.chrom_u8 = narrow["chrom"].data_ptr() if "chrom" in narrow else None
        t.n_alignments_u16 = narrow["n_alignments"].data_ptr() if "n_alignments" in narrow else None
        t.aln_size_is_qspan = int(bool(getattr(buf, "aln_is_qspan", False)))
        rpr = getattr(buf, "rows_per_read", None)
        t.rows_per_read_u8 = rpr.data_ptr() if rpr is not None else None
        if buf.order is not None:
            t.order, t.n_order = buf.order.data_ptr(), int(buf.order.numel())
        else:
            t.order, t.n_order = None, 0
        return t

    def _check(self, rc):
        if rc != 0:
            raise _native.FslrError(rc, self.lib.fslrc_last_error(self.ctx).decode())

    # ---- compute entry points
    def run_resident(self, dtab: DeviceTable, chrom_table, params: ClusterParams):
        """Inputs already in HBM; outputs stay in HBM (dtab.out_cluster / out_n_reads).  Returns stats dict."""
        p = self._params(chrom_table, params)
        t = self._table(dtab)
        st = _native.Stats()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.fslrc_cluster_device(self.ctx, C.byref(t), C.byref(p), dtab.out_cluster.data_ptr(),
                                                  dtab.out_n_reads.data_ptr(), C.byref(st), C.c_void_p(stream)))
        return st.as_dict(self.lib)

    def run_host(self, ptab: PinnedTable, chrom_table, params: ClusterParams):
        """Host buffers in, host buffers out (H2D + D2H inside the call).  Returns stats dict."""
        p = self._params(chrom_table, params)
        t = self._table(ptab)
        st = _native.Stats()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.fslrc_cluster_host(self.ctx, C.byref(t), C.byref(p), ptab.out_cluster.data_ptr(),
                                                ptab.out_n_reads.data_ptr(), C.byref(st), C.c_void_p(stream)))
        return st.as_dict(self.lib)

    def cluster(self, table: ColumnarTable, params: ClusterParams, order=None) -> ClusterResult:
        ptab = PinnedTable(table, order)
        stats = self.run_host(ptab, table, params)
        n = table.n_reads
        return ClusterResult(ptab.out_cluster[:n].numpy().copy(), ptab.out_n_reads[:n].numpy().copy(),
                             bool(stats["no_clusters"]), stats)

    # ---- multi-GPU: pair space sharded over ranks, three exchange steps
    def run_sharded(self, dtab: DeviceTable, chrom_table, params: ClusterParams, rank, world, group=None):
        """Every rank holds the whole table in HBM (dtab); candidate generation and the pair tests run on shard `rank` of `world`.
        Exchange steps (torch.distributed, NCCL on GPUs): sum-all-reduce of the per-read partner counters (4 bytes per query
        read), all-gather of the recorded pairs of saturating reads (8 bytes per pair: the replay needs them on every
        rank), all-gather of each rank's spanning forest.  Every rank ends with the full result in dtab.out_*."""
        from .sharded import exchange_counts, exchange_forests
        p = self._params(chrom_table, params)
        t = self._table(dtab)
        st = _native.Stats()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.fslrc_mg_prepare(self.ctx, C.byref(t), C.byref(p), C.c_void_p(stream)))
        ptr, n = C.c_void_p(), C.c_int64()
        self._check(self.lib.fslrc_mg_pair(self.ctx, rank, world, C.byref(ptr), C.byref(n)))
        if n.value > 0 and world > 1:
            exchange_counts(torch.as_tensor(_DevView(ptr.value, (n.value,)), device=self.device), group)
        self._check(self.lib.fslrc_mg_partners(self.ctx, rank, world, C.byref(ptr), C.byref(n)))
        npairs = int(n.value)
        if world > 1:
            local = (torch.as_tensor(_DevView(ptr.value, (2 * npairs,)), device=self.device) if npairs > 0
                     else torch.zeros(0, dtype=torch.int32, device=self.device))
            pairs, per_rank = exchange_forests(local, group)                   # (same variable-length all-gather)
            npairs = int(sum(per_rank))
            pptr = pairs.data_ptr() if npairs > 0 else None
        else:
            pptr = ptr.value if npairs > 0 else None
        self._check(self.lib.fslrc_mg_replay(self.ctx, rank, world, pptr, npairs, C.byref(ptr), C.byref(n)))
        ne = int(n.value)
        if world > 1:
            local = (torch.as_tensor(_DevView(ptr.value, (2 * ne,)), device=self.device) if ne > 0
                     else torch.zeros(0, dtype=torch.int32, device=self.device))
            forest, per_rank = exchange_forests(local, group)
            tot = int(sum(per_rank))
            fptr = forest.data_ptr() if tot > 0 else None
        else:
            tot, fptr = ne, ptr.value if ne > 0 else None
        self._check(self.lib.fslrc_mg_finish(self.ctx, fptr, tot, dtab.out_cluster.data_ptr(), dtab.out_n_reads.data_ptr(),
                                             C.byref(st)))
        return st.as_dict(self.lib)

    def upload_sharded(self, ptab: PinnedTable, dtab: DeviceTable, rank, world, group=None):
        """Multi-GPU ingest of HOST columns: every rank copies only its 1/world slice of the rows over its own PCIe link and
        the slices are all-gathered over NVLink (fslr_b200.sharded.gather_columns), so the host->device time shrinks with
        the number of GPUs instead of being paid in full by every rank.  Narrow columns of a compact PinnedTable travel
        narrow (host->device and over NVLink) and are widened on the device."""
        from .sharded import gather_columns
        n = dtab.n_rows
        gather_columns({k: ptab.cols[k] for k in ptab.cols}, {k: dtab.cols[k] for k in ptab.cols}, n, rank, world, group)
        if ptab.narrow:
            if not hasattr(dtab, "_narrow"):
                dtab._narrow = {k: torch.empty(max(n, 1), dtype=v.dtype, device=self.device) for k, v in ptab.narrow.items()}
            gather_columns(ptab.narrow, dtab._narrow, n, rank, world, group)
            dtab.cols["chrom"][:n].copy_(dtab._narrow["chrom"][:n])
            dtab.cols["n_alignments"][:n].copy_(dtab._narrow["n_alignments"][:n])
            dtab.cols["n_alignments"][:n].bitwise_and_(0xffff)
        if getattr(ptab, "aln_is_qspan", False):
            torch.sub(dtab.cols["qend"][:n], dtab.cols["qstart"][:n], out=dtab.cols["aln_size"][:n])
        if getattr(ptab, "rows_per_read", None) is not None:   # run lengths (one byte per read, sent by every rank) -> read ids
            cnt = ptab.rows_per_read.to(self.device, non_blocking=True).to(torch.int64)
            dtab.cols["read_id"][:n].copy_(torch.repeat_interleave(torch.arange(cnt.numel(), device=self.device, dtype=torch.int32), cnt))

    def choose_alignment(self, read_id, alignment_score, cluster, n_clusters):
        """cluster.py:237-254 on the GPU.  read_id/alignment_score: int32 per table row; cluster: int32 per read (dense ids).
        Returns (is_rep uint8 [n_reads], rep_read int32 [n_clusters])."""
        rid = np.ascontiguousarray(read_id, dtype=np.int32)
        sc = np.ascontiguousarray(alignment_score, dtype=np.int32)
        cl = np.ascontiguousarray(cluster, dtype=np.int32)
        is_rep = np.zeros(max(cl.shape[0], 1), dtype=np.uint8)
        rep = np.full(max(int(n_clusters), 1), -1, dtype=np.int32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.fslrc_choose_alignment_host(self.ctx, rid.shape[0], cl.shape[0], int(n_clusters), rid.ctypes.data,
                                                         sc.ctypes.data, cl.ctypes.data, is_rep.ctypes.data, rep.ctypes.data,
                                                         C.c_void_p(stream)))
        return is_rep[:cl.shape[0]], rep[:int(n_clusters)]

    def launch_count(self):
        return int(self.lib.fslrc_launch_count(self.ctx))

    def int_peak(self):
        v = C.c_double()
        self._check(self.lib.fslrc_int_peak(self.ctx, C.byref(v)))
        return v.value


class HostPipeline:
    """Host-buffer clustering of a stream of tables with the upload of one table overlapping the kernels of the previous
    one: `depth` library contexts, each with its own CUDA stream and worker thread (the C ABI call blocks its thread, not the
    interpreter).  This is how a service clusters sample after sample; `bench.py` uses it for the end-to-end number."""

    def __init__(self, device=0, depth=2, blocking_sync=None):
        import os
        from concurrent.futures import ThreadPoolExecutor
        self.engines = [Engine(device) for _ in range(depth)]
        if blocking_sync is None:                          # default: sleep in the waits — spinning worker threads starve each other
            blocking_sync = os.environ.get("FSLR_B200_BLOCKING_SYNC", "1") == "1"   # on hosts that give the process few cores
        for e in self.engines:                            # worker threads sleep in their waits instead of spinning
            e.lib.fslrc_set_blocking_sync(e.ctx, int(bool(blocking_sync)))
        self.streams = [torch.cuda.Stream(device=self.engines[0].device) for _ in range(depth)]
        self.pool = ThreadPoolExecutor(max_workers=depth)
        self.depth, self._next = depth, 0
        import threading
        self._locks = [threading.Lock() for _ in range(depth)]    # a library context is not re-entrant

    def _run(self, slot, ptab, chrom_table, params):
        eng, stream = self.engines[slot], self.streams[slot]
        torch.cuda.set_device(eng.device)
        with self._locks[slot]:
            p = eng._params(chrom_table, params)
            t = eng._table(ptab)
            st = _native.Stats()
            eng._check(eng.lib.fslrc_cluster_host(eng.ctx, C.byref(t), C.byref(p), ptab.out_cluster.data_ptr(),
                                                  ptab.out_n_reads.data_ptr(), C.byref(st), C.c_void_p(stream.cuda_stream)))
            return st.as_dict(eng.lib)

    def _run_resident(self, slot, dtab, chrom_table, params):
        eng, stream = self.engines[slot], self.streams[slot]
        torch.cuda.set_device(eng.device)
        with self._locks[slot]:
            p = eng._params(chrom_table, params)
            t = eng._table(dtab)
            st = _native.Stats()
            eng._check(eng.lib.fslrc_cluster_device(eng.ctx, C.byref(t), C.byref(p), dtab.out_cluster.data_ptr(),
                                                    dtab.out_n_reads.data_ptr(), C.byref(st), C.c_void_p(stream.cuda_stream)))
            return st.as_dict(eng.lib)

    def submit_resident(self, dtab: DeviceTable, chrom_table, params):
        """Same for a table that already lives in HBM (results stay in dtab.out_*): the latency-bound replay of one table
        runs beside the issue-bound kernels of the next."""
        slot = self._next
        self._next = (self._next + 1) % self.depth
        return self.pool.submit(self._run_resident, slot, dtab, chrom_table, params)

    def submit(self, ptab: PinnedTable, chrom_table, params):
        """ptab must not be reused by another submit before this one's future is done (its out_* buffers receive the result)."""
        slot = self._next
        self._next = (self._next + 1) % self.depth
        return self.pool.submit(self._run, slot, ptab, chrom_table, params)

    def launch_count(self):
        return sum(e.launch_count() for e in self.engines)

    def close(self):
        self.pool.shutdown(wait=True)
        for e in self.engines:
            e.close()


def table_chrom_len(table):
    return table.chrom_len


_engines = {}


def get_engine(device=0):
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
