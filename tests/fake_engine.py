"""numpy test double of the engine's reduce stages (TEST INFRASTRUCTURE).

Implements the same stage contract as ``traceml_b200.engine.Engine`` on CPU
tensors so the multi-process orchestration in ``traceml_b200.reduce`` /
``sections`` (collectives, sharding, partial-sum merging) can run under
``gloo`` with world_size 2 on a box without a GPU.  It is NOT a fallback: the
product never imports it.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

RF_USABLE, RF_HAS_MEM, RF_IN_TIME, RF_CAND_T, RF_CAND_M = 1, 2, 4, 8, 16


class FakeEngine:
    def __init__(self, records, procs=None, device=0, dense_ok=False):
        self.dense_ok = dense_ok
        self.records = records
        self.procs = procs
        self.device = device
        self.xrows = {0: None, 1: None}

    # ---- stage 1
    def win_prepare(self, window, stream=0):
        r = self.records
        n = len(r)
        self.n = n
        self.t_start = max(0, n - window)
        ms = r["dur_ns"].astype(np.float64) / 1.0e6
        self.rows = np.zeros((n, 8))
        self.rows[:, :6] = ms
        self.rows[:, 6] = r["peak_alloc"].astype(np.float64)
        self.rows[:, 7] = r["peak_resv"].astype(np.float64)
        self.steps = r["step"].astype(np.int64)
        usable = (ms[:, [0, 2, 3, 4, 5]] > 0).any(axis=1) if n else np.zeros(0, bool)
        has_mem = (r["flags"] & 1) != 0
        idx = np.arange(n)
        in_time = idx >= self.t_start
        prev_diff = np.ones(n, bool)
        prev_diff[1:] = self.steps[1:] != self.steps[:-1]
        first = prev_diff | (idx == self.t_start)
        nxt_diff = np.ones(n, bool)
        if n > 1:
            nxt_diff[:-1] = (self.steps[1:] != self.steps[:-1]) | (~has_mem[1:])
        self.cand = {0: usable & in_time & first, 1: has_mem & nxt_diff}
        sel = usable & in_time
        fwd, bwd, opt, wall, dl = ms[:, 2], ms[:, 3], ms[:, 4], ms[:, 5], ms[:, 0]
        comp = (fwd + bwd) + opt
        traced = np.maximum(wall, comp)
        sums = [0.0] * 7
        for i in range(n - 1, self.t_start - 1, -1):  # reference order
            if sel[i]:
                for k, v in enumerate((dl[i], fwd[i], bwd[i], opt[i], wall[i], traced[i], dl[i] + traced[i])):
                    sums[k] += v
        out = SimpleNamespace(
            n_retained=n, latest_step=int(self.steps.max()) if n else 0, monotone=1,
            dup_rows=int((~prev_diff).sum()) if n else 0,
            n_rows=[int(in_time.sum()), int(has_mem.sum())],
            n_cand=[int(self.cand[0].sum()), int(self.cand[1].sum())],
            lo=[int(self.steps[self.cand[k]].min()) if self.cand[k].any() else 0 for k in (0, 1)],
            hi=[int(self.steps[self.cand[k]].max()) if self.cand[k].any() else 0 for k in (0, 1)],
            t_sums=sums, t_count=int(sel.sum()), n_both=int((self.cand[0] & self.cand[1]).sum()), dense=(0, 0))
        if self.dense_ok:
            rows_in = (n - self.t_start, n)
            out.dense = tuple(int(out.n_cand[k] > 0 and out.n_cand[k] == rows_in[k]
                                  and out.hi[k] - out.lo[k] + 1 == out.n_cand[k]) for k in (0, 1))
        self.lo = out.lo
        return out

    def win_select_dense(self, kind, first, n_common, stream=0):
        """Dense window: row(step) = first_row + (step - lo); same outputs as win_select."""
        span = int(n_common)
        presence = torch.ones(span, dtype=torch.uint8)
        rowof = (self.t_start if kind == 0 else 0) + (int(first) - self.lo[kind]) + np.arange(span)
        if not hasattr(self, "rowof"):
            self.rowof = {}
        self.rowof[kind] = rowof
        return self.win_select(kind, int(first), span, presence, span, stream)

    # ---- stage 2
    def win_presence(self, kind, glo, span, presence, stream=0):
        if self.n == 0 or not self.cand[kind].any():
            presence.fill_(1)
            return
        presence.zero_()
        self.rowof = {} if not hasattr(self, "rowof") else self.rowof
        rowof = np.full(span, -1, dtype=np.int64)
        for i in np.nonzero(self.cand[kind])[0]:
            s = int(self.steps[i]) - glo
            if 0 <= s < span:
                presence[s] = 1
                rowof[s] = i
        self.rowof[kind] = rowof

    # ---- stage 3
    def win_select(self, kind, glo, span, presence, window, stream=0):
        p = presence.numpy().astype(bool)
        idx = np.nonzero(p)[0][-window:]
        out = SimpleNamespace(n_common=len(idx), start_step=0, end_step=0, n_rows=0,
                              t_sums=[0.0] * 7, m_sums=[0.0] * 4)
        if len(idx) == 0 or self.n == 0 or not self.cand[kind].any():
            self.xrows[kind] = None
            return out
        rows = self.rows[self.rowof[kind][idx]]
        self.xrows[kind] = rows.copy()
        dl, fwd, bwd, opt, wall = rows[:, 0], rows[:, 2], rows[:, 3], rows[:, 4], rows[:, 5]
        comp = (fwd + bwd) + opt
        traced = np.maximum(wall, comp)
        sums = [0.0] * 7
        for j in range(len(idx) - 1, -1, -1):
            for k, v in enumerate((dl[j], fwd[j], bwd[j], opt[j], max(0.0, traced[j]), traced[j], dl[j] + traced[j])):
                sums[k] += v
        out.t_sums = sums
        out.m_sums = [float(rows[:, 6].sum()), float(rows[:, 7].sum()),
                      float(rows[:, 6].max()), float(rows[:, 7].max())]
        out.start_step, out.end_step, out.n_rows = int(glo + idx[0]), int(glo + idx[-1]), len(idx)
        return out

    def win_rows_tensor(self, kind, n):
        if self.xrows[kind] is None:
            return torch.zeros(n * 8, dtype=torch.float64)
        return torch.from_numpy(self.xrows[kind].reshape(-1).copy())

    # ---- stage 4
    def win_reduce(self, rows, mask, n, lo, hi, series, stream=0):
        R = len(rows)

        def shard(r):
            if isinstance(r, int):  # virtual base pointer (a2a exchange): rows lo..hi live at r + j*64
                import ctypes as C
                m = hi - lo
                return np.ctypeslib.as_array((C.c_double * (m * 8)).from_address(r + lo * 64)).reshape(m, 8)
            return r.numpy().reshape(n, 8)[lo:hi]

        a = np.stack([shard(r) for r in rows])  # [R, m, 8]
        dl, fwd, bwd, opt, wall = a[..., 0], a[..., 2], a[..., 3], a[..., 4], a[..., 5]
        comp = (fwd + bwd) + opt
        traced = np.maximum(wall, comp)
        wait = np.maximum(0.0, traced - comp)
        S = series.numpy().reshape(16, n)
        mets = [dl, fwd, bwd, opt, traced, wait, a[..., 6], a[..., 7]]
        for mi, v in enumerate(mets):
            if (mi < 6 and not mask & 1) or (mi >= 6 and not mask & 2):
                continue
            srt = np.sort(v, axis=0)
            med = srt[R // 2] if R % 2 else (srt[R // 2 - 1] + srt[R // 2]) * 0.5
            S[2 * mi, lo:hi] = med
            S[2 * mi + 1, lo:hi] = srt[-1]

    # ---- stage 5
    def win_bands(self, series, args, stream=0):
        n = int(args.n_common)
        S = series.numpy().reshape(16, n)
        lo_s, hi_s = int(args.shard_lo), int(args.shard_hi)
        out = SimpleNamespace(sum=[[0.0] * 3 for _ in range(16)], cnt=[[0] * 3 for _ in range(16)],
                              tail_first=[float("nan")] * 16, tail_last=[float("nan")] * 16)
        for s in range(16):
            k = 1 if s >= 12 else 0
            for b in range(3):
                lo, hi = max(int(args.band_lo[k][b]), lo_s), min(int(args.band_hi[k][b]), hi_s)
                if hi > lo:
                    out.sum[s][b] = float(S[s, lo:hi].sum())
                    out.cnt[s][b] = hi - lo
            f, l = int(args.tail_first[k]), n - 1
            if n and lo_s <= f < hi_s:
                out.tail_first[s] = float(S[s, f])
            if n and lo_s <= l < hi_s:
                out.tail_last[s] = float(S[s, l])
        return out

    def proc_reduce(self, max_rows, stream=0):
        from test_native_diag_cpu import _agg_from_records

        if self.procs is None:
            from traceml_b200 import _abi

            a = _abi.ProcAgg()
            a.max_ratio = -1.0
            return a
        return _agg_from_records(self.procs, max_rows)

    # ---- live tick (tml_combined_*): numpy double of csrc/tml_combined.cuh
    def combined_prepare(self, kind, lookback, stream=0):
        total = len(self.records)
        r = self.records[-int(lookback):] if total else self.records
        n = len(r)
        c = SimpleNamespace()
        c.rows = np.zeros((n, 8))
        c.rows[:, :6] = r["dur_ns"].astype(np.float64) / 1.0e6
        c.rows[:, 6] = r["peak_alloc"].astype(np.float64)
        c.rows[:, 7] = r["peak_resv"].astype(np.float64)
        c.steps = r["step"].astype(np.int64)
        has_mem = (r["flags"] & 1) != 0
        last = np.ones(n, bool)
        if n > 1:
            last[:-1] = c.steps[1:] != c.steps[:-1]
            if kind == 1:
                last[:-1] |= ~has_mem[1:]
        if kind == 1:
            last &= has_mem
        c.cand, c.n, c.inrange = last, n, 0
        if not hasattr(self, "comb"):
            self.comb = {}
        self.comb[kind] = c
        return SimpleNamespace(n_rows=n, n_cand=int(last.sum()),
                               lo=int(c.steps[last].min()) if last.any() else 0,
                               hi=int(c.steps[last].max()) if last.any() else 0,
                               latest_step=int(c.steps.max()) if n else 0,
                               first_step=int(c.steps[0]) if n else 0,
                               truncated=int(total > n), monotone=1)

    def combined_presence(self, kind, glo, span, presence, stream=0):
        c = self.comb[kind]
        if c.n == 0:
            presence.fill_(1)
            return
        inr = [i for i in np.nonzero(c.cand)[0] if 0 <= int(c.steps[i]) - glo < span]
        c.inrange = len(inr)
        if not inr and kind == 1:
            presence.fill_(1)
            return
        presence.zero_()
        c.rowof = np.full(span, -1, dtype=np.int64)
        for i in inr:
            presence[int(c.steps[i]) - glo] = 1
            c.rowof[int(c.steps[i]) - glo] = i

    def combined_select(self, kind, glo, span, presence, window, stream=0):
        c = self.comb[kind]
        idx = np.nonzero(presence.cpu().numpy())[0][-int(window):]
        n = len(idx)
        out = SimpleNamespace(n_common=n, n_rows=0, sums=[0.0] * 6, peaks=[0.0, 0.0])
        c.x = None
        if n == 0 or c.n == 0 or c.inrange == 0:
            return out
        c.x = torch.from_numpy(np.ascontiguousarray(c.rows[c.rowof[idx]]))
        c.sel_steps = [int(glo + i) for i in idx]
        sums = [0.0] * 6
        for row in c.x.numpy():  # ascending step order
            for k in range(6):
                sums[k] += float(row[k])
        out.n_rows, out.sums = n, sums
        out.peaks = [float(c.x[:, 6].max()), float(c.x[:, 7].max())]
        return out

    def combined_rows_tensor(self, kind, n):
        c = self.comb[kind]
        return c.x.reshape(-1) if c.x is not None else torch.empty(0, dtype=torch.float64)

    def combined_steps(self, kind, n, stream=0):
        return list(self.comb[kind].sel_steps)

    def combined_series(self, ptrs, n, first_col, n_cols, series, stream=0):
        import ctypes as C

        rows = [np.ctypeslib.as_array((C.c_double * (n * 8)).from_address(int(p))).reshape(n, 8)
                for p in ptrs]
        out = series.numpy()
        for m in range(n_cols):
            vals = np.stack([r[:, first_col + m] for r in rows], axis=1)  # [n, R]
            for j in range(n):
                v = np.ascontiguousarray(vals[j])
                out[m * 3, j] = np.median(v)
                out[m * 3 + 1, j] = np.max(v)
                out[m * 3 + 2, j] = np.sum(v)
