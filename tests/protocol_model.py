#!/usr/bin/env python
"""Discrete-event model of the barrier protocol of the fused split kernels (mlp_fused_split.cu): the TMA producer, the MMA
issuer and the eight epilogue warps as coroutines over mbarriers with the hardware's phase/parity semantics, a FIFO
tensor pipe, and hazard checks on tensor memory (ACT columns: in-place rewrite vs MMAs still reading the old layer;
accumulator quarters: overwritten only after they were loaded, loaded only after their MMAs completed).

    python tests/protocol_model.py                       # shipped schedule vs the quarter schedule of the draft patch

It answers two questions without a GPU: (1) is the protocol of `k_mlp_fused_split_q` (mlp_fused_split.cu) deadlock- and hazard-free over
several tiles of the forward (8 trunk layers, skip at 4, condition layer) and of the dgrad chain, with exactly the parity
expressions of the code; (2) what period per layer does each schedule give under the cycle costs measured with the
in-kernel clock64 stamps (N=128 MMA 73 cycles -> 3.5 k per half of 48; one 32-column epilogue pass 1.45 k).
The model mirrors the control flow of the kernels line by line; it does not model shared-memory capacity beyond the
ring's stage count, nor the TMEM port contention between MMAs and tcgen05.ld/st."""
from __future__ import annotations

import heapq
import sys

T_MMA128, T_MMA64 = 73.0, 36.5      # cycles per tcgen05.mma M=128, N=128 / N=64 (K=16) as issued back to back
T_TMA = 400.0                       # ring stage load latency (L2-resident weights / encodings)
T_LD = 230.0                        # tcgen05.ld + wait, 8-warp effective (B300_MICROARCH.md)
T_CHUNK = 1220.0                    # bias/ReLU/split/stores/ship of one 32-column pass after the load
T_ISSUE = 4.0                       # issuer cost per MMA
N_EPI = 8


class Deadlock(Exception):
    pass


class Hazard(Exception):
    pass


class Bar:
    def __init__(self, name, count):
        self.name, self.init, self.pending, self.completed = name, count, count, 0
        self.waiters = []  # (parity, thread)

    def ok(self, parity):
        return (self.completed & 1) != parity


class Step:
    def __init__(self, n_act_kb, n_enc_kb, n_halves, produces):
        self.n_act_kb, self.n_enc_kb, self.n_halves, self.produces = n_act_kb, n_enc_kb, n_halves, produces


class Sim:
    def __init__(self, steps, n_tiles, NS, quarters):
        self.steps, self.n_tiles, self.NS, self.quarters = steps, n_tiles, NS, quarters
        self.t = 0.0
        self.q = []
        self.seq = 0
        self.threads = {}
        self.blocked = {}
        self.w_full = [Bar(f"w_full{i}", 1) for i in range(NS)]
        self.w_empty = [Bar(f"w_empty{i}", 1) for i in range(NS)]
        nacc = 4 if quarters else 2
        self.acc_full = [Bar(f"acc_full{i}", 1) for i in range(nacc)]
        self.acc_empty = [Bar(f"acc_empty{i}", N_EPI) for i in range(nacc)]
        self.act_ready, self.act_lo_ready = Bar("act_ready", N_EPI), Bar("act_lo_ready", N_EPI)
        self.fifo_free = 0.0
        # hazard state.  ACT is tracked in 4 column groups g = K/64 (k-block g reads group g); version = (tile, step) that wrote it
        self.act_ver = [None] * 4
        self.act_readers = [[] for _ in range(4)]     # (version expected, completion time)
        self.acc_state = {}                            # quarter-or-half -> dict(writer=(tile, step), done=time, loaded=set(warps))
        self.layer_end = []                            # completion time of each layer's last MMA (tile 0..)
        self.mma_busy = 0.0

    # ---- engine
    def at(self, t, fn):
        self.seq += 1
        heapq.heappush(self.q, (t, self.seq, fn))

    def spawn(self, name, gen):
        self.threads[name] = gen
        self.at(0.0, lambda n=name: self.step_thread(n))

    def step_thread(self, name):
        gen = self.threads[name]
        try:
            op = next(gen)
        except StopIteration:
            del self.threads[name]
            return
        if op[0] == "delay":
            self.at(self.t + op[1], lambda n=name: self.step_thread(n))
        elif op[0] == "wait":
            bar, parity = op[1], op[2]
            if bar.ok(parity):
                self.at(self.t, lambda n=name: self.step_thread(n))
            else:
                bar.waiters.append((parity, name))
                self.blocked[name] = (bar.name, parity)

    def arrive(self, bar):
        bar.pending -= 1
        if bar.pending == 0:
            bar.pending, bar.completed = bar.init, bar.completed + 1
            still = []
            for parity, name in bar.waiters:
                if bar.ok(parity):
                    self.blocked.pop(name, None)
                    self.at(self.t, lambda n=name: self.step_thread(n))
                else:
                    still.append((parity, name))
            bar.waiters = still

    def run(self):
        while self.q:
            self.t, _, fn = heapq.heappop(self.q)
            fn()
        if self.threads:
            raise Deadlock(f"t={self.t:.0f}: blocked {self.blocked}")

    # ---- tensor pipe
    def mma(self, dur, reads_act=None, expect=None, acc=None, first=False, who=None):
        """issued now; executes in FIFO order.  reads_act: ACT column group; expect: version it must hold; acc: accumulator id."""
        start = max(self.t, self.fifo_free)
        end = start + dur
        self.fifo_free = end
        self.mma_busy += dur
        if reads_act is not None:
            if self.act_ver[reads_act] != expect:
                raise Hazard(f"t={self.t:.0f}: MMA of {who} reads ACT group {reads_act} = {self.act_ver[reads_act]}, expected {expect}")
            self.act_readers[reads_act].append((expect, end))
        if acc is not None:
            st = self.acc_state.get(acc)
            if first:
                if st is not None and len(st["loaded"]) < N_EPI:
                    raise Hazard(f"t={self.t:.0f}: accumulator {acc} of {st['writer']} overwritten by {who} before all warps loaded it")
                self.acc_state[acc] = dict(writer=who, done=end, loaded=set())
            else:
                if st is None or st["writer"] != who:
                    raise Hazard(f"t={self.t:.0f}: accumulate into {acc} of {st and st['writer']} by {who}")
                st["done"] = end
        return end

    def commit(self, bar):
        t = max(self.t, self.fifo_free)
        self.at(t, lambda b=bar: self.arrive(b))

    def write_act(self, group_halves, ver, warp):
        """an epilogue warp's tcgen05.st into ACT: group_halves = list of column groups it touches (each warp writes its lanes)."""
        for g in group_halves:
            for expect, end in self.act_readers[g]:
                if end > self.t + 1e-9 and expect != ver:
                    raise Hazard(f"t={self.t:.0f}: warp {warp} rewrites ACT group {g} with {ver} while an MMA reading {expect} runs until {end:.0f}")
            self.act_pending = getattr(self, "act_pending", {})
            key = (g, ver)
            self.act_pending.setdefault(key, set()).add(warp)
            if len(self.act_pending[key]) == N_EPI * 1:  # all warps wrote their lanes / columns of this group
                pass
        return

    def publish_act(self, groups, ver):
        for g in groups:
            self.act_ver[g] = ver
            self.act_readers[g] = [(e, t) for e, t in self.act_readers[g] if t > self.t]

    def load_acc(self, acc, who, warp):
        st = self.acc_state.get(acc)
        if st is None or st["writer"] != who:
            raise Hazard(f"t={self.t:.0f}: warp {warp} loads accumulator {acc} expecting {who}, holds {st and st['writer']}")
        if st["done"] > self.t + 1e-9:
            raise Hazard(f"t={self.t:.0f}: warp {warp} loads accumulator {acc} of {who} before its MMAs complete ({st['done']:.0f})")
        st["loaded"].add(warp)


# ------------------------------------------------------------------------------------------------ schedules
def two_phase(st):
    return st.n_halves == 2 and st.n_act_kb == 4


def q_item(st, i):
    n_kb = st.n_act_kb + st.n_enc_kb
    if not two_phase(st):
        return i // n_kb, i % n_kb
    a = st.n_enc_kb + 2
    if i < 2 * a:
        h, r = i // a, i % a
        return h, (4 + r if r < st.n_enc_kb else r - st.n_enc_kb)
    j = i - 2 * a
    return j >> 1, 2 + (j & 1)


def items(st, quarters):
    n = st.n_halves * (st.n_act_kb + st.n_enc_kb)
    if quarters:
        return [q_item(st, i) for i in range(n)]
    return [(h, kb) for h in range(st.n_halves) for kb in range(st.n_act_kb + st.n_enc_kb)]


def producer(S: Sim):
    it = 0
    for tile in range(S.n_tiles):
        for s, st in enumerate(S.steps):
            for h, kb in items(st, S.quarters):
                for _ in range(2 if kb >= st.n_act_kb else 1):  # encoding k-block: its A stage first, then the W stage
                    ws = it % S.NS
                    yield ("wait", S.w_empty[ws], ((it // S.NS) & 1) ^ 1)
                    S.at(S.t + T_TMA, lambda b=S.w_full[ws]: S.arrive(b))
                    it += 1
                    yield ("delay", 10.0)


def act_version(tile, s):
    return (tile, s - 1)  # step s reads the output of step s-1 of the same tile


def issuer_halves(S: Sim):
    """the shipped kernel (two-instalment ACT readiness)"""
    it, n_acc, n_act = 0, [0, 0], 0
    for tile in range(S.n_tiles):
        for s, st in enumerate(S.steps):
            n_kb = st.n_act_kb + st.n_enc_kb
            for h in range(st.n_halves):
                yield ("wait", S.acc_empty[h], (n_acc[h] & 1) ^ 1)
                n_acc[h] += 1
                wait_act = h == 0 and st.n_act_kb > 0
                if wait_act:
                    yield ("wait", S.act_lo_ready, n_act & 1)
                for kb in range(n_kb):
                    from_act = kb < st.n_act_kb
                    if wait_act and kb == st.n_act_kb // 2:
                        yield ("wait", S.act_ready, n_act & 1)
                    a_stage = None
                    if not from_act:
                        a_stage = it % S.NS
                        yield ("wait", S.w_full[a_stage], (it // S.NS) & 1)
                        it += 1
                    ws = it % S.NS
                    yield ("wait", S.w_full[ws], (it // S.NS) & 1)
                    for k in range(4):
                        for term in range(3):
                            S.mma(T_MMA128, reads_act=kb if from_act else None, expect=act_version(tile, s), acc=h,
                                  first=(kb == 0 and k == 0 and term == 0), who=(tile, s))
                            yield ("delay", T_ISSUE)
                    if a_stage is not None:
                        S.commit(S.w_empty[a_stage])
                    S.commit(S.w_empty[ws])
                    it += 1
                if wait_act:
                    n_act += 1
                S.commit(S.acc_full[h])
            S.layer_end.append((tile, s, S.fifo_free))


def epilogue_halves(S: Sim, warp):
    n_full = [0, 0]
    for tile in range(S.n_tiles):
        for s, st in enumerate(S.steps):
            for h in range(st.n_halves):
                last = h == st.n_halves - 1
                yield ("wait", S.acc_full[h], n_full[h] & 1)
                n_full[h] += 1
                unpark = last and st.n_halves == 2 and st.produces
                if unpark:
                    S.write_act([0, 1], (tile, s), warp)
                yield ("delay", T_LD)
                S.load_acc(h, (tile, s), warp)
                if unpark and warp == 0:
                    pass
                S.arrive(S.acc_empty[h])
                if unpark:
                    if S.act_lo_ready.pending == 1:
                        S.publish_act([0, 1], (tile, s))
                    S.arrive(S.act_lo_ready)
                yield ("delay", 2 * T_CHUNK)
                if last and st.produces:
                    S.write_act([2, 3] if st.n_halves == 2 else [0, 1], (tile, s), warp)
                    if S.act_ready.pending == 1:
                        S.publish_act([2, 3], (tile, s))
                    S.arrive(S.act_ready)


def issuer_quarters(S: Sim):
    """fused_split_quarters.patch, line by line"""
    it, n_acc, n_act = 0, [0, 0, 0, 0], 0
    for tile in range(S.n_tiles):
        for s, st in enumerate(S.steps):
            n_kb = st.n_act_kb + st.n_enc_kb
            two = two_phase(st)
            n_a = 2 * (st.n_enc_kb + 2) if two else st.n_halves * n_kb
            act_lo_seen = act_hi_seen = False
            cur_h, n_in_half = -1, 0
            for i in range(n_a):
                h, kb = q_item(st, i)
                if h != cur_h:
                    cur_h, n_in_half = h, 0
                    for q in (2 * h, 2 * h + 1):
                        yield ("wait", S.acc_empty[q], (n_acc[q] & 1) ^ 1)
                        n_acc[q] += 1
                from_act = kb < st.n_act_kb
                if from_act and not act_lo_seen:
                    yield ("wait", S.act_lo_ready, n_act & 1)
                    act_lo_seen = True
                if from_act and kb >= 2 and not act_hi_seen:
                    yield ("wait", S.act_ready, n_act & 1)
                    act_hi_seen = True
                a_stage = None
                if not from_act:
                    a_stage = it % S.NS
                    yield ("wait", S.w_full[a_stage], (it // S.NS) & 1)
                    it += 1
                ws = it % S.NS
                yield ("wait", S.w_full[ws], (it // S.NS) & 1)
                for k in range(4):
                    for term in range(3):
                        first = n_in_half == 0 and k == 0 and term == 0
                        for q in (2 * h, 2 * h + 1):  # an N=128 MMA covers both quarters of the half
                            S.mma(0.0, acc=q, first=first, who=(tile, s))
                        S.mma(T_MMA128, reads_act=kb if from_act else None, expect=act_version(tile, s), who=(tile, s))
                        for q in (2 * h, 2 * h + 1):
                            S.acc_state[q]["done"] = S.fifo_free
                        yield ("delay", T_ISSUE)
                if a_stage is not None:
                    S.commit(S.w_empty[a_stage])
                S.commit(S.w_empty[ws])
                it += 1
                n_in_half += 1
                if not two and n_in_half == n_kb:
                    S.commit(S.acc_full[2 * h])
                    S.commit(S.acc_full[2 * h + 1])
            if two:
                if not act_hi_seen:
                    yield ("wait", S.act_ready, n_act & 1)
                    act_hi_seen = True
                for h in range(2):
                    ws0, ws1 = it % S.NS, (it + 1) % S.NS
                    yield ("wait", S.w_full[ws0], (it // S.NS) & 1)
                    yield ("wait", S.w_full[ws1], ((it + 1) // S.NS) & 1)
                    for j in range(2):
                        for b in range(2):
                            for k in range(4):
                                for term in range(3):
                                    S.mma(T_MMA64, reads_act=2 + b, expect=act_version(tile, s), acc=2 * h + j, who=(tile, s))
                                    yield ("delay", T_ISSUE)
                        S.commit(S.acc_full[2 * h + j])
                    S.commit(S.w_empty[ws0])
                    S.commit(S.w_empty[ws1])
                    it += 2
            if st.n_act_kb > 0:
                n_act += 1
            S.layer_end.append((tile, s, S.fifo_free))


def epilogue_quarters(S: Sim, warp):
    n_full = [0, 0, 0, 0]
    for tile in range(S.n_tiles):
        for s, st in enumerate(S.steps):
            nq = 2 * st.n_halves
            waited3 = False
            for q in range(nq):
                if not (q == 3 and waited3):
                    yield ("wait", S.acc_full[q], n_full[q] & 1)
                    n_full[q] += 1
                yield ("delay", T_LD)
                S.load_acc(q, (tile, s), warp)
                S.arrive(S.acc_empty[q])
                yield ("delay", T_CHUNK * 0.6)      # bias / ReLU / split
                if st.produces:
                    if q == 2 and not waited3:
                        yield ("wait", S.acc_full[3], n_full[3] & 1)
                        n_full[3] += 1
                        waited3 = True
                    S.write_act([q], (tile, s), warp)
                    if q in (1, 3):
                        bar = S.act_lo_ready if q == 1 else S.act_ready
                        if bar.pending == 1:
                            S.publish_act([0, 1] if q == 1 else [2, 3], (tile, s))
                        S.arrive(bar)
                yield ("delay", T_CHUNK * 0.4)      # ship + bit plane


def network(kind):
    if kind == "forward":  # 8 trunk layers (skip connection into layer 4), condition layer
        st = [Step(0, 2, 2, 1)] + [Step(4, 2 if i == 4 else 0, 2, 1) for i in range(1, 8)] + [Step(4, 1, 1, 0)]
    else:                   # dgrad chain: step 0 streams dZ of the condition layer, then 7 trunk steps
        st = [Step(0, 2, 2, 1)] + [Step(4, 0, 2, 1 if i < 7 else 0) for i in range(1, 8)]
    return st


def run(kind, quarters, n_tiles=4, NS=5):
    S = Sim(network(kind), n_tiles, NS, quarters)
    S.spawn("producer", producer(S))
    S.spawn("issuer", (issuer_quarters if quarters else issuer_halves)(S))
    for w in range(N_EPI):
        S.spawn(f"epi{w}", (epilogue_quarters if quarters else epilogue_halves)(S, w))
    S.run()
    ends = [t for _, _, t in S.layer_end]
    per_tile = (ends[-1] - ends[len(S.steps) - 1]) / (n_tiles - 1)  # steady state: tiles 1..n-1
    return S.t, per_tile, S.mma_busy / S.t


if __name__ == "__main__":
    for kind in ("forward", "dgrad"):
        for quarters in (False, True):
            try:
                total, per_tile, util = run(kind, quarters)
                n_layers = len(network(kind))
                print(f"{kind:8s} {'quarters' if quarters else 'halves  '}: {per_tile:8.0f} cycles per tile ({per_tile / n_layers:6.0f} per layer), "
                      f"tensor pipe busy {100 * util:4.1f} %  -- no deadlock, no hazard")
            except (Deadlock, Hazard) as e:
                print(f"{kind:8s} {'quarters' if quarters else 'halves  '}: {type(e).__name__}: {e}")
                sys.exit(1)
