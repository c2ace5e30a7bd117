"""TEST INFRASTRUCTURE ONLY — pure-Python restatement of the one class of the
third-party `superintervals` package (kcleal/superintervals, required
`>= 0.2.10` by /root/reference/setup.py:17, unpinned, source NOT present in this
container) that /root/reference/fslr/cluster.py uses:

    IntervalMap(with_data=True)   cluster.py:125
    .add(start, end, value)       cluster.py:127
    .build()                      cluster.py:129
    .search_values(start, end)    cluster.py:201

PARITY UNPINNED at this boundary: the reference ships no test that fixes the
result order of `search_values`, and the library is not installed, so the
behaviour below is the published algorithm as restated in SURVEY.md §8c:

  * intervals are closed on both ends;
  * build() orders by (start asc, end desc), ties keep insertion order;
  * branch[i] = nearest k < i with end_k >= end_i, else -1;
  * search_values(s, e) finds the last index with start <= e and walks DOWN,
    emitting every i with end_i >= s, skipping through `branch` otherwise, so
    results come in DESCENDING sorted position.

The order only matters through the reference's `edge_threshold` break
(cluster.py:223-224); the edge relation itself is order independent.
"""
from bisect import bisect_right


class IntervalMap:
    def __init__(self, with_data=False):
        self.with_data = with_data
        self._items = []          # (start, end, insertion_no, value)
        self.starts = []
        self.ends = []
        self.data = []
        self.branch = []

    def add(self, start, end, value=None):
        self._items.append((start, end, len(self._items), value))

    def build(self):
        self._items.sort(key=lambda t: (t[0], -t[1], t[2]))
        self.starts = [t[0] for t in self._items]
        self.ends = [t[1] for t in self._items]
        self.data = [t[3] for t in self._items]
        branch = [-1] * len(self._items)
        stack = []                # indices with strictly decreasing... ends (monotone stack)
        for i, e in enumerate(self.ends):
            while stack and self.ends[stack[-1]] < e:
                stack.pop()
            branch[i] = stack[-1] if stack else -1
            stack.append(i)
        self.branch = branch

    def search_values(self, start, end):
        out = []
        i = bisect_right(self.starts, end) - 1
        ends, data, branch = self.ends, self.data, self.branch
        while i >= 0:
            if ends[i] >= start:
                out.append(data[i])
                i -= 1
            else:
                i = branch[i]
        return out
