"""The synchronisation protocol of concat_cost_cluster_kernel (knn_svc_b200/csrc/concat_cost_sm100.cu), replayed on the
CPU as a discrete-event model under random schedules.

compute-sanitizer's racecheck does not follow mbarrier arrive / try_wait pairs across warps, so it reports every
producer -> consumer pair of that kernel as a hazard (profiles/r2b_sanitizer.txt).  This model is the other half of the
argument: the kernel's agents (8 compute warps, 2 producer warps, 4 fetch warps, 1 output warp in each of the 8 CTAs of a
cluster), its barriers (full / empty per ring buffer, the exchange barriers per step parity and candidate, the cost
barrier) with the hardware's phase-parity semantics, and its asynchronous copies (cp.async.bulk, st.async) that land at
arbitrary later times.  Every buffer carries the tag of what was written into it; every read asserts the tag it expects.
A protocol error shows up as a wrong tag (a buffer overwritten too early or read too soon), as a parity wait that
aliases (a stale tag again), or as a deadlock.  It checks the PROTOCOL as written in the kernel's comments, not the CUDA
code itself: the GPU tests compare that against the one-CTA kernel bit for bit."""
import random

import pytest

N_CTA, N_CAND, K, GENS = 8, 8, 4, 3


class MBar:
    """mbarrier: pending arrivals + transaction bytes of the current phase; try_wait.parity(P) is true when the phase of
    parity P has completed, i.e. the current phase has the other parity"""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _settle(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, tx=0):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.tx += tx
        self.pending -= 1
        self._settle()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._settle()

    def done(self, parity):
        return (self.phase & 1) != parity


class Cta:
    def __init__(self):
        self.rows = [[None] * 13 for _ in range(GENS)]          # tag: (generation, row)
        self.meta = [dict() for _ in range(GENS)]               # 'A': gen, 'B': gen, ('F', r): gen
        self.xch = [[[None] * N_CTA for _ in range(N_CAND)] for _ in range(2)]   # [parity][candidate][source] -> step
        self.cost = [[None] * N_CAND for _ in range(2)]         # [parity][candidate] -> step
        self.full = [MBar(2 + K) for _ in range(GENS)]
        self.empty = [MBar(1) for _ in range(GENS)]
        self.xbar = [[MBar(1) for _ in range(N_CAND)] for _ in range(2)]
        self.cbar = MBar(N_CAND + 1 + K)


def selection(step):
    """the (deterministic, identical in all CTAs) selection of a step: 4 distinct candidate slots; "step 0" is idx[0]"""
    if step == 0:
        return list(range(K))
    rs = random.Random(1000 + step)
    return rs.sample(range(N_CAND), K)


def ring_row(step, m):
    """row of candidate m of `step` in its generation's ring buffer: the rows of idx[step] for m < 4, else the row after
    selection m-4 of the step before — speculative slot = that selection's candidate slot"""
    return m if m < K else K + selection(step - 1)[m - K]


def run(n_steps, seed, output_arrives=True, output_delay=3, max_async=300):
    rng = random.Random(seed)
    ctas = [Cta() for _ in range(N_CTA)]
    if not output_arrives:                      # (the deliberately broken variant of the last test)
        for c in ctas:
            c.cbar = MBar(N_CAND + K)
    pending_async = []      # (due tick, fn)
    tick = [0]

    def later(fn):
        pending_async.append((tick[0] + rng.randint(1, max_async), fn))

    def tma(cta, g, row, tag, bar):           # cp.async.bulk: lands later, then complete_tx
        def land():
            cta.rows[g][row] = tag
            bar.complete_tx(1)
        later(land)

    # ---- agents (generators: `yield cond` blocks until cond() is true)
    def producer_a(c):
        for r in range(K):
            tma(c, 0, r, (0, r), c.full[0])
        c.meta[0]['A'] = 0
        c.full[0].arrive(tx=K)
        for t in range(1, n_steps):
            g = t % GENS
            if t >= GENS:
                yield lambda: c.empty[g].done((t // GENS - 1) & 1)
            for r in list(range(K)) + [12]:
                tma(c, g, r, (t, r), c.full[g])
            c.meta[g]['A'] = t
            c.full[g].arrive(tx=K + 1)
            yield lambda: True

    def producer_b(c):
        c.full[0].arrive()
        for t in range(1, n_steps):
            g = t % GENS
            if t >= GENS:
                yield lambda: c.empty[g].done((t // GENS - 1) & 1)
            for r in range(K):
                tma(c, g, K + r, (t, K + r), c.full[g])
            c.meta[g]['B'] = t
            c.full[g].arrive(tx=K)
            yield lambda: True

    def fetch(c, r):
        c.full[0].arrive()
        if n_steps > 1:
            c.full[1].arrive()
        for s in range(1, n_steps):
            if s >= 2:
                yield lambda: c.cbar.done(s & 1)                       # costs of step s-1 (phase s-2)
                assert all(v == s - 1 for v in c.cost[(s & 1) ^ 1]), "fetch warp read costs of the wrong step"
            c.cbar.arrive()
            if s + 1 < n_steps:
                g = (s + 1) % GENS
                tma(c, g, 2 * K + r, (s + 1, 2 * K + r), c.full[g])
                c.meta[g][('F', r)] = s + 1
                c.full[g].arrive(tx=1)
            yield lambda: True

    def output(c):
        for s in range(1, n_steps):
            if output_arrives:
                c.cbar.arrive()
            for _ in range(rng.randint(0, output_delay)):              # (slow to come back to the barrier)
                yield lambda: True
            yield lambda: c.cbar.done((s - 1) & 1)
            for _ in range(rng.randint(0, output_delay)):              # (reading takes a while)
                yield lambda: True
            assert all(v == s for v in c.cost[s & 1]), "output warp read costs of the wrong step"
            c.empty[(s - 1) % GENS].arrive()
            yield lambda: True

    def compute(ci, m):
        c = ctas[ci]
        yield lambda: c.full[0].done(0)
        full_phase = 1
        g, gp = 1, 0
        for s in range(1, n_steps):
            par = s & 1
            yield lambda: c.full[g].done((full_phase >> g) & 1)
            full_phase ^= 1 << g
            c.xbar[par][m].arrive(tx=N_CTA)
            assert c.rows[g][12] == (s, 12), f"query row of generation {s} not in place: {c.rows[g][12]}"
            assert c.meta[g].get('A') == s and (s < 1 or c.meta[g].get('B') == s)
            if s >= 2:
                yield lambda: c.cbar.done(par)                        # phase s-2
                assert all(v == s - 1 for v in c.cost[par ^ 1]), "compute warp read costs of the wrong step"
            sel_prev_rows = [ring_row(s - 1, sl) for sl in selection(s - 1)]   # ring rows of the previous selection
            crow = ring_row(s, m)
            # the candidate's row and the four previously selected rows must hold what the step expects
            assert c.rows[g][crow] == (s, crow), f"step {s} candidate {m}: row {crow} holds {c.rows[g][crow]}"
            if crow >= 2 * K:
                assert c.meta[g].get(('F', crow - 2 * K)) == s
            for pr in sel_prev_rows:
                assert c.rows[gp][pr] == (s - 1, pr), f"step {s}: previous row {pr} holds {c.rows[gp][pr]}"
            # st.async of the partial sums to warp m of every CTA
            for d in range(N_CTA):
                def land(d=d, s=s, par=par):
                    ctas[d].xch[par][m][ci] = s
                    ctas[d].xbar[par][m].complete_tx(1)
                later(land)
            yield lambda: c.xbar[par][m].done(((s - 1) >> 1) & 1)
            assert all(v == s for v in c.xch[par][m]), f"step {s}: exchange buffer holds {c.xch[par][m]}"
            c.cost[par][m] = s
            c.cbar.arrive()
            gp, g = g, (0 if g == GENS - 1 else g + 1)
            yield lambda: True

    agents = []
    for ci, c in enumerate(ctas):
        agents += [producer_a(c), producer_b(c), output(c)] + [fetch(c, r) for r in range(K)]
        agents += [compute(ci, m) for m in range(N_CAND)]
    waiting = {}
    for a in agents:
        try:
            waiting[a] = next(a)
        except StopIteration:
            pass
    while waiting or pending_async:
        tick[0] += 1
        due = [x for x in pending_async if x[0] <= tick[0]]
        if due:
            x = rng.choice(due)
            pending_async.remove(x)
            x[1]()
        ready = [a for a, cond in waiting.items() if cond()]
        if ready:
            a = rng.choice(ready)
            try:
                waiting[a] = next(a)
            except StopIteration:
                del waiting[a]
        elif not pending_async:
            raise AssertionError(f"deadlock at tick {tick[0]}: {len(waiting)} agents blocked")
        else:
            tick[0] = max(tick[0], min(x[0] for x in pending_async) - 1)      # nothing can move before the next landing
    return tick[0]


@pytest.mark.parametrize("seed", range(3))
def test_cluster_protocol_has_no_stale_read_and_no_deadlock(seed):
    """n = 1, 2, 3 (the prologue cases) and a run long enough to wrap every ring and parity several times; copies that
    land almost at once or very late; an output warp that dawdles (the compute warps must wait for it)"""
    for n_steps in (1, 2, 3, 4):
        run(n_steps, seed * 100 + n_steps)
    run(14, seed, max_async=20)
    run(14, seed + 50, max_async=3000)
    run(14, seed + 90, output_delay=400)


def test_the_model_catches_a_protocol_error():
    """sanity of the model itself: without the output warp's own arrival on the cost barrier, that barrier can complete
    TWO phases while the output warp is away, its parity wait then aliases (it waits for a phase that needs a ring buffer
    only the output warp can free), and the cluster deadlocks — the model must see it"""
    with pytest.raises(AssertionError):
        for seed in range(5):
            run(14, seed, output_arrives=False, output_delay=3000)
