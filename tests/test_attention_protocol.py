"""Discrete-event model of the synchronisation protocol of ``attention_tcgen05_v3_kernel`` (memvul_b200/csrc/
attention_tcgen05_v3.cuh): one CTA = TMA producer warp, MMA issuer warp and four soft-max warps that share ONE S
accumulator, ONE P buffer, two K and two V stages through eleven mbarriers whose wait parities are derived from block /
item counters.  The GPU tests show that the kernel computes the right numbers on the shapes they run; this model checks
the protocol itself -- for random work lists (skipped items, one-block items, ragged lengths) and random latencies of
every agent:
  * no deadlock (every role finishes), with the hardware's parity semantics (``try_wait.parity P`` succeeds iff the
    barrier's current phase has parity != P, so a barrier that runs two phases ahead of a waiter hangs or mis-fires);
  * no hazard: S is not overwritten before all four warps have loaded it, P not before the previous P V has retired and
    not while the ctx TMA store still reads the staging rows, a K / V stage not before the products that read it have
    retired, Q not before the item's last Q K^T, O not before the previous item's read-out, and every product finds the
    operands of ITS block in place.
The role programs below restate the kernel's control flow line by line (same barriers, same parity expressions)."""
import heapq
import random

import pytest

KV = 2                      # Attn3Cfg::KV_STAGES
BREAK = None                # set by the self-test of the checker: drop one wait of the protocol


class Bar:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, f"{self.name}: more arrivals than its count in one phase"
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase + 1

    def done(self, parity):                 # mbarrier.try_wait.parity semantics
        return (self.phase & 1) != parity


class Sim:
    """Cooperative scheduler: roles are generators yielding ('wait', bar, parity) or ('delay', cycles)."""

    def __init__(self, rng):
        self.rng, self.now, self.seq = rng, 0.0, 0
        self.ready, self.blocked, self.alive = [], [], 0

    def spawn(self, gen, name):
        self.alive += 1
        self._push(0.0, gen, name)

    def _push(self, t, gen, name):
        self.seq += 1
        heapq.heappush(self.ready, (t, self.seq, gen, name))

    def run(self, limit=5_000_000):
        steps = 0
        while self.alive:
            steps += 1
            assert steps < limit, "model did not terminate"
            still = []
            for gen, name, bar, par in self.blocked:          # wake waiters whose phase has completed
                if bar.done(par):
                    self._push(self.now, gen, name)
                else:
                    still.append((gen, name, bar, par))
            self.blocked = still
            if not self.ready:
                waiting = [(n, b.name, p, b.phase) for _, n, b, p in self.blocked]
                raise AssertionError(f"deadlock at t={self.now:.0f}: {waiting}")
            t, _, gen, name = heapq.heappop(self.ready)
            self.now = max(self.now, t)
            try:
                op = next(gen)
            except StopIteration:
                self.alive -= 1
                continue
            if op[0] == "delay":
                self._push(self.now + op[1], gen, name)
            else:
                _, bar, par = op
                if bar.done(par):
                    self._push(self.now + self.rng.uniform(1, 30), gen, name)
                else:
                    self.blocked.append((gen, name, bar, par))


def simulate(items, seed):
    """items: list of (q0, len) in the order the CTA walks them (q0 >= len: the item is skipped by every role)."""
    rng = random.Random(seed)
    sim = Sim(rng)
    lat = lambda lo, hi: ("delay", rng.uniform(lo, hi))
    B = {n: Bar(n, c) for n, c in dict(q_full=1, q_empty=1, s_full=1, s_free=4, p_full=4, pv_done=1, o_free=4).items()}
    for i in range(KV):
        for n in ("k_full", "v_full", "k_empty", "v_empty"):
            B[f"{n}{i}"] = Bar(f"{n}{i}", 1)
    # ---- shared state the hazards are checked on ----
    st = {"S_block": None, "S_read": 4, "P_written": {}, "P_block_ready": None, "pv_retired": -1, "qk_retired": -1,
          "K": [None] * KV, "V": [None] * KV, "Q_item": None, "O_item": None, "O_read": 4, "stg_busy": [False] * 4,
          "last_qk_of_item": {}, "n_pv": 0}
    live = [(i, q0, ln) for i, (q0, ln) in enumerate(items) if q0 < ln]
    nkb_of = {i: (ln + 63) // 64 for i, _, ln in live}
    first_block, g = {}, 0
    for i, _, _ in live:
        first_block[i] = g
        g += nkb_of[i]
    total_blocks = g
    pipe = []                                                   # the tensor pipe: in-order queue of issued operations

    def tensor_pipe():
        idle = 0
        while st["n_pv"] < total_blocks or pipe:
            if not pipe:
                idle += 1
                assert idle < 2_000_000, "tensor pipe starved"
                yield ("delay", 5)
                continue
            idle = 0
            op = pipe.pop(0)
            if op[0] == "commit":
                op[1].arrive()
                continue
            kind, gb, item, j = op
            s = gb % KV
            if kind == "qk":
                assert st["S_read"] == 4, f"Q K^T of block {gb} overwrites S before all warps loaded block {st['S_block']}"
                assert st["K"][s] == gb, f"Q K^T of block {gb} finds K of block {st['K'][s]} in its stage"
                assert st["Q_item"] == item, f"Q K^T of item {item} finds the Q tile of item {st['Q_item']}"
                yield lat(20, 200)
                st["S_block"], st["S_read"], st["qk_retired"] = gb, 0, gb
            else:
                assert st["P_block_ready"] == gb, f"P V of block {gb} issued before its P is complete"
                assert st["V"][s] == gb, f"P V of block {gb} finds V of block {st['V'][s]}"
                if j == 0:
                    assert st["O_read"] == 4, f"first P V of item {item} overwrites O before the previous read-out"
                    st["O_item"], st["O_read"] = item, 0
                yield lat(20, 200)
                st["pv_retired"] = gb
                st["n_pv"] += 1

    def landed(key, slot, value, bar):                      # a TMA load in flight: completes on its own
        yield lat(200, 1500)
        if slot is None:
            st[key] = value
        else:
            st[key][slot] = value
        bar.arrive()

    def tma():
        gcnt = it = 0
        for i, q0, ln in live:
            yield ("wait", B["q_empty"], (it & 1) ^ 1)
            if it > 0:
                prev = live[it - 1][0]
                assert st["qk_retired"] >= first_block[prev] + nkb_of[prev] - 1, "Q overwritten before the item's last Q K^T"
            yield lat(2, 20)
            sim.spawn(landed("Q_item", None, i, B["q_full"]), "ldQ")
            for j in range(nkb_of[i]):
                s, par = gcnt % KV, ((gcnt // KV) & 1) ^ 1
                yield ("wait", B[f"k_empty{s}"], par)
                assert st["K"][s] is None or st["qk_retired"] >= st["K"][s], "K stage overwritten before its Q K^T retired"
                yield lat(2, 20)
                sim.spawn(landed("K", s, gcnt, B[f"k_full{s}"]), "ldK")
                yield ("wait", B[f"v_empty{s}"], par)
                assert st["V"][s] is None or st["pv_retired"] >= st["V"][s], "V stage overwritten before its P V retired"
                yield lat(2, 20)
                sim.spawn(landed("V", s, gcnt, B[f"v_full{s}"]), "ldV")
                gcnt += 1
            it += 1

    def mma():
        g0 = it = 0

        def issue_qk(i, j, nkb):
            gb = g0 + j
            s = gb % KV
            if gb > 0 and BREAK != "no_s_free":
                yield ("wait", B["s_free"], (gb - 1) & 1)
            yield ("wait", B[f"k_full{s}"], (gb // KV) & 1)
            yield lat(5, 60)
            pipe.append(("qk", gb, i, j))
            pipe.append(("commit", B["s_full"]))
            pipe.append(("commit", B[f"k_empty{s}"]))
            if j == nkb - 1:
                pipe.append(("commit", B["q_empty"]))

        for i, q0, ln in live:
            nkb = nkb_of[i]
            yield ("wait", B["q_full"], it & 1)
            yield from issue_qk(i, 0, nkb)
            for j in range(nkb):
                gb = g0 + j
                s = gb % KV
                if j + 1 < nkb:
                    yield from issue_qk(i, j + 1, nkb)
                yield ("wait", B["p_full"], gb & 1)
                if j == 0:
                    yield ("wait", B["o_free"], (it & 1) ^ 1)
                yield ("wait", B[f"v_full{s}"], (gb // KV) & 1)
                yield lat(5, 60)
                pipe.append(("pv", gb, i, j))
                pipe.append(("commit", B["pv_done"]))
                pipe.append(("commit", B[f"v_empty{s}"]))
            g0 += nkb
            it += 1

    def softmax(w):
        gb = 0
        store_pending = False
        for i, (q0, ln) in enumerate(items):
            if q0 >= ln:
                yield lat(10, 100)                              # zero-fill of a padded tile: no barrier traffic
                continue
            nkb = (ln + 63) // 64
            full_tile = q0 + 128 <= ln or rng.random() < 0.5    # padded layout: S rows exist; packed: only len rows
            for j in range(nkb):
                yield ("wait", B["s_full"], gb & 1)
                assert st["S_block"] == gb, f"warp {w} loads S of block {st['S_block']} for block {gb}"
                yield lat(20, 150)                              # tcgen05.ld + wait::ld
                st["S_read"] += 1
                B["s_free"].arrive()                            # one arrival per warp (count 4)
                yield lat(100, 900)                             # row maximum, exponentials into registers
                if j == 0:
                    if store_pending:
                        yield lat(0, 400)                       # cp.async.bulk.wait_group.read
                        st["stg_busy"][w] = False
                        store_pending = False
                elif BREAK != "no_pv_done":
                    yield ("wait", B["pv_done"], (gb - 1) & 1)
                assert st["pv_retired"] >= gb - 1, f"warp {w} stores P of block {gb} before P V of block {gb - 1} retired"
                assert not st["stg_busy"][w], f"warp {w} stores P while its ctx store still reads the staging rows"
                yield lat(10, 80)                               # STS of the packed row (+ rare O rescale)
                st["P_written"][gb] = st["P_written"].get(gb, 0) + 1
                if st["P_written"][gb] == 4:
                    st["P_block_ready"] = gb
                B["p_full"].arrive()
                gb += 1
            yield ("wait", B["pv_done"], (gb - 1) & 1)
            assert st["pv_retired"] >= gb - 1 and st["O_item"] == i
            yield lat(20, 150)                                  # O read-out
            st["O_read"] += 1
            B["o_free"].arrive()
            if full_tile:
                yield lat(10, 80)                               # staging + TMA store issue
                st["stg_busy"][w] = True
                store_pending = True

    sim.spawn(tensor_pipe(), "pipe")
    sim.spawn(tma(), "tma")
    sim.spawn(mma(), "mma")
    for w in range(4):
        sim.spawn(softmax(w), f"softmax{w}")
    sim.run()
    assert st["n_pv"] == total_blocks and not pipe
    return total_blocks


def _work_list(rng, n_items):
    items = []
    for _ in range(n_items):
        ln = rng.choice([1, 5, 63, 64, 65, 128, 129, 200, 256, 300, 511, 512])
        items.append((rng.choice([0, 128, 256, 384]), ln))
    return items


@pytest.mark.parametrize("seed", range(32))
def test_three_stream_attention_protocol_random_work_lists(seed):
    rng = random.Random(1000 + seed)
    blocks = simulate(_work_list(rng, rng.randint(1, 14)), seed)
    assert blocks >= 0


def test_three_stream_attention_protocol_corner_lists():
    assert simulate([(0, 512)] * 7, 1) == 56                   # C2: seven 8-block items per CTA
    assert simulate([(0, 1)], 2) == 1                           # one key, one block, one item
    assert simulate([(0, 64), (128, 100), (0, 64), (0, 1)], 3) == 3   # one-block items around a skipped one
    assert simulate([(384, 300), (256, 200)], 4) == 0           # every item skipped
    assert simulate([(0, 65), (128, 129), (256, 512)], 5) == 2 + 3 + 8


@pytest.mark.parametrize("broken", ["no_s_free", "no_pv_done"])
def test_the_model_catches_a_broken_protocol(monkeypatch, broken):
    """Sanity of the checker itself: without the MMA warp's `s_free` wait (S overwritten under the soft-max) or without
    the soft-max's `pv_done` wait before it stores P (P overwritten under the tensor core) a hazard check must fire."""
    import sys
    monkeypatch.setattr(sys.modules[__name__], "BREAK", broken)
    with pytest.raises(AssertionError):
        for seed in range(8):
            simulate([(0, 512)] * 3, seed)
