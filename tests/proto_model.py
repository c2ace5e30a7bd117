"""Executable model (numpy/pure Python, small inputs only) of the PARALLEL formulation the CUDA path
implements, used by the CPU tests to check the decomposition itself against the oracle/goldens.

The reference loop (cluster.py:187-227) is sequential because of `seen_edges` and the per-query
`edges >= edge_threshold` break.  The model removes the sequence like this:

  * Every filling scan covers a contiguous run of sorted positions [stop_f, ub_f] (descending), so
    the whole run state is one integer `stop` per filling; "X visited A" and "A met B before" are
    pure functions of the stops and the interval table.
  * phase A (all reads, order free): relation pass(A->B) over all candidate pairs; reads whose
    count of passing candidates is < edge_threshold can never break (set P = the others).
  * phase B (P only): fixed point over the stops; a read is final once every earlier-ranked P
    candidate it inspected was final.  Rounds may process reads in ANY order.
  * phase C: edges.  A pair {A<B} is tested from A if A visited B, else from B if B visited A.
"""
import numpy as np


def _thr(a, p):
    """min{o >= 0 : float(o)/float(a) >= p} in the reference's double arithmetic (cluster.py:133-136)."""
    if p <= 0:
        return 0
    o = int(np.ceil(p * a))
    while o > 0 and (o - 1) / a >= p:
        o -= 1
    while o / a < p:
        o += 1
    return o


def _umax(n, t):
    """max union u >= n with n/u >= t (cluster.py:165-170,218-219); n-1 if none."""
    if t <= 0:
        return 2**31 - 1
    u = int(n / t) + 2
    while u >= n and not (n / u >= t):
        u -= 1
    return u if u >= n else n - 1


class Model:
    def __init__(self, table, params, order=None):
        t, pr = table, params
        A, R = t.n_rows, t.n_reads
        rid = t.read_id
        idx = np.arange(A)
        first = np.full(R, A); last = np.full(R, -1)
        np.minimum.at(first, rid, idx); np.maximum.at(last, rid, idx)
        keep = (idx != first[rid]) & (idx != last[rid])
        rows = idx[keep]                                                 # fillings, bed order
        qmax = np.full(R, -2**62); qmin = np.full(R, 2**62)
        np.maximum.at(qmax, rid[rows], t.qend[rows].astype(np.int64)); np.minimum.at(qmin, rid[rows], t.qstart[rows].astype(np.int64))
        start = np.minimum(t.rstart, t.rend)[rows].astype(np.int64); end = np.maximum(t.rstart, t.rend)[rows].astype(np.int64)
        o = np.asarray(order) if order is not None else np.argsort(start, kind="stable")
        rows, start, end = rows[o], start[o], end[o]
        chrom = t.chrom[rows]
        m = np.ones(rows.shape[0], bool)
        if pr.chrom_masked is not None:
            m &= pr.chrom_masked[chrom] == 0
        if pr.mask_subtelomere:
            cl = t.chrom_len[chrom]
            m &= ~((cl > 1_000_000) & ((start < pr.subtel) | (cl - end < pr.subtel)))
        rows, start, end, chrom = rows[m], start[m], end[m], chrom[m]
        self.D = D = rows.shape[0]
        self.R = R
        self.start, self.end, self.chrom = start, end, chrom
        self.aln = t.aln_size[rows]
        irid = rid[rows]
        # query rank = order of first appearance in data
        _, fpos = np.unique(irid, return_index=True)
        qreads = irid[np.sort(fpos)]
        self.Q = Q = qreads.shape[0]
        self.rid_of_q = qreads
        q_of_rid = np.full(R, -1); q_of_rid[qreads] = np.arange(Q)
        self.q = q_of_rid[irid]
        self.lists = [[] for _ in range(Q)]
        for d in range(D):
            self.lists[self.q[d]].append(d)
        self.qlen2 = (qmax - qmin)[qreads]
        self.naln = np.array([t.n_alignments[rows[l[0]]] for l in self.lists], dtype=np.int64) if Q else np.zeros(0, np.int64)
        # sorted positions per chrom: (chrom, start asc, end desc, data order)
        self.sidx = np.lexsort((np.arange(D), -end, start, chrom))
        self.pos = np.empty(D, np.int64); self.pos[self.sidx] = np.arange(D)
        sc, ss = chrom[self.sidx], start[self.sidx]
        self.chrom_lo = {c: int(np.searchsorted(sc, c, "left")) for c in np.unique(chrom)}
        self.ub = np.empty(D, np.int64)
        for d in range(D):
            c = chrom[d]; lo = self.chrom_lo[c]; hi = int(np.searchsorted(sc, c, "right"))
            self.ub[d] = lo + int(np.searchsorted(ss[lo:hi], end[d], "right")) - 1
        self.T = np.array([_thr(int(a), pr.overlap) for a in self.aln], dtype=np.int64)
        cq, cn = 1 - pr.qlen_diff, 1 - pr.n_alignment_diff
        self.Lq = np.array([_thr(int(x), cq) for x in self.qlen2], dtype=np.int64)
        self.Ln = np.array([_thr(int(x), cn) for x in self.naln], dtype=np.int64)
        cut = pr.jaccard_cutoffs
        Lmax = max([len(l) for l in self.lists] + [1])
        self.umax = [0] + [_umax(n, cut[n - 1] if n - 1 < len(cut) else cut[-1]) for n in range(1, Lmax + 1)]
        self.Tedge = pr.edge_threshold

    # ---- pair-level pieces
    def difflen_ok(self, a, b):
        q_ok = min(self.qlen2[a], self.qlen2[b]) >= max(self.Lq[a], self.Lq[b])
        n_ok = min(self.naln[a], self.naln[b]) >= max(self.Ln[a], self.Ln[b])
        return q_ok or n_ok

    def match(self, i, j):
        if self.chrom[i] != self.chrom[j]:
            return False
        ov = max(0, min(self.end[i], self.end[j]) - max(self.start[i], self.start[j]))
        return ov >= max(self.T[i], self.T[j])

    def greedy(self, a, b):
        used, n = set(), 0
        for i in self.lists[a]:
            for j in self.lists[b]:
                if j in used:
                    continue
                if self.match(i, j):
                    used.add(j); n += 1
                    break
        return n

    def passes(self, a, b):
        """a queries b: reaches line 218 (n>0) and j >= target."""
        if not self.difflen_ok(a, b):
            return False, False
        n = self.greedy(a, b)
        if n == 0:
            return False, False
        u = len(self.lists[a]) + len(self.lists[b]) - n
        return True, u <= self.umax[n]

    def candidates_at(self, f):
        """sorted positions of closed-overlap candidates of item f, descending (search_values order)."""
        lo = self.chrom_lo[self.chrom[f]]
        return [p for p in range(int(self.ub[f]), lo - 1, -1) if self.end[self.sidx[p]] >= self.start[f]]

    # ---- phase A
    def phase_a(self):
        Q = self.Q
        self.degub = np.zeros(Q, np.int64)
        self.later, self.cond = [], []
        for a in range(Q):
            done = set()
            for f in self.lists[a]:
                for p in self.candidates_at(f):
                    b = self.q[self.sidx[p]]
                    if b == a or b in done:
                        continue
                    done.add(b)
                    _, ok = self.passes(a, b)
                    if ok:
                        self.degub[a] += 1
                        (self.later if b > a else self.cond).append((a, b))
        self.isP = self.degub >= self.Tedge

    # ---- stop-based predicates
    def visited(self, b, a):
        """did b's own query scan reach one of a's intervals (given b's current stops)?"""
        for f in self.lists[b]:
            for g in self.lists[a]:
                if self.chrom[f] == self.chrom[g] and self.stop[f] <= self.pos[g] <= self.ub[f] and self.end[g] >= self.start[f]:
                    return True
        return False

    def replay(self, a, emit=None):
        """a's own query given the current stops of earlier reads.  Returns (stops, depends_on_nonfinal)."""
        edges, my, dep = 0, {}, False
        la = self.lists[a]
        for fi, f in enumerate(la):
            stop = self.chrom_lo[self.chrom[f]]
            for p in self.candidates_at(f):
                o = self.sidx[p]; b = self.q[o]
                if b == a:
                    continue
                if b < a and self.isP[b] and not self.final[b]:
                    dep = True
                met = False                                     # met b earlier in this very query?
                for g in self.lists[b]:
                    for f2 in la[:fi]:
                        if self.chrom[f2] == self.chrom[g] and my[f2] <= self.pos[g] <= self.ub[f2] and self.end[g] >= self.start[f2]:
                            met = True
                    if self.chrom[g] == self.chrom[f] and p < self.pos[g] <= self.ub[f] and self.end[g] >= self.start[f]:
                        met = True
                if met:
                    continue
                if b < a and (not self.isP[b] or self.visited(b, a)):
                    continue                                    # seen: b's query got here first
                reach, ok = self.passes(a, b)
                if not reach:
                    continue
                if ok:
                    edges += 1
                    if emit is not None:
                        emit.append((a, b))
                if edges >= self.Tedge:
                    stop = p
                    break
            my[f] = stop
        return my, dep

    def run(self, rng=None):
        self.phase_a()
        Q = self.Q
        self.stop = np.array([self.chrom_lo[c] for c in self.chrom], dtype=np.int64) if self.D else np.zeros(0, np.int64)
        self.final = ~self.isP
        self.rounds = 0
        P = [a for a in range(Q) if self.isP[a]]
        while not all(self.final[a] for a in P):
            self.rounds += 1
            todo = [a for a in P if not self.final[a]]
            if rng is not None:
                rng.shuffle(todo)
            for a in todo:
                my, dep = self.replay(a)
                for f, s in my.items():
                    self.stop[f] = s
                if not dep:
                    self.final[a] = True
        edges = [(a, b) for (a, b) in self.later if not self.isP[a]]
        edges += [(a, b) for (a, b) in self.cond if not self.isP[a] and self.isP[b] and not self.visited(b, a)]
        for a in P:
            self.replay(a, emit=edges)
        self.edges = edges
        parent = list(range(Q))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]; x = parent[x]
            return x
        ing = np.zeros(Q, bool)
        for a, b in edges:
            ing[a] = ing[b] = True
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        root = np.array([find(x) for x in range(Q)], dtype=np.int64)
        isroot = ing & (root == np.arange(Q))
        cidx = np.cumsum(isroot) - 1
        ncl = int(isroot.sum())
        size = np.bincount(root[ing], minlength=Q) if Q else np.zeros(0, np.int64)
        cluster = np.full(self.R, -1, np.int64); nreads = np.ones(self.R, np.int64)
        inG_rid = np.zeros(self.R, bool)
        for x in range(Q):
            if ing[x]:
                r = self.rid_of_q[x]
                inG_rid[r] = True; cluster[r] = cidx[root[x]]; nreads[r] = size[root[x]]
        cluster[~inG_rid] = ncl + np.arange(int((~inG_rid).sum()))
        return cluster.astype(np.int32), nreads.astype(np.int32), ncl == 0
