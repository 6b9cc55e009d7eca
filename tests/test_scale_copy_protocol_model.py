"""CPU: model check of the scale-factor pipeline of the block-scaled GEMM (csrc/gemm_tcgen05.cu, SCALED kinds).

Actors per CTA pair, as in the kernel (one k-block = one pipeline stage use):
  * TMA producer: waits on empty[s] (parity ph ^ 1), then issues the stage's SCALE loads (complete sf_ld[s]) and its operand loads
    (complete full[s]).  Loads are asynchronous: they land later, in any order.
  * scale-copy thread (warp 2): waits on sf_ld[s] (parity ph) -- and, when the TMEM scale buffers are a ring SHORTER than the
    stage ring (SF_NB < STAGES), on sf_empty[b] (parity bph ^ 1) -- then issues the tcgen05.cp copies smem[s] -> TMEM buffer b and
    a commit on sf_full[b].  Its queue executes asynchronously, in issue order.
  * MMA thread: waits on full[s] (parity ph) and sf_full[b] (parity ph, or bph for the ring), issues the MMAs (they read
    operands smem[s] and scale buffer b), then a commit on empty[s] (and on sf_empty[b] for the ring).  Its queue executes
    asynchronously, in issue order, independently of the copy thread's queue.
  * buffer index b = s when SF_NB == STAGES ("atoms landed" implies "buffer free": the producer only reloads stage s after the
    MMAs of its previous round retired), else a separate ring index.
mbarrier parity semantics as in test_pair_tile_protocol_model.py.

Checked under random interleavings: a copy never writes a TMEM scale buffer that an unretired MMA still reads, an MMA never
executes before the copy of ITS k-block completed, a scale load never overwrites smem a pending copy still reads, every k-block
is multiplied with its own scales, and nobody deadlocks.  Seeded bugs show the model can fail: a ring shorter than the stage
ring WITHOUT the sf_empty handshake, and an MMA thread that does not wait for sf_full.
"""
import random
from collections import deque

import pytest


class Barrier:
    def __init__(self):
        self.phase = 0

    def passed(self, parity):
        return (self.phase & 1) != parity

    def complete(self):
        self.phase += 1


def run_model(n_kb, stages, sf_nb, seed, bug=None):
    rng = random.Random(seed)
    per_stage = sf_nb == stages
    use_sf_empty = (not per_stage) and bug != "ring_without_sf_empty"
    full, empty, sf_ld = ([Barrier() for _ in range(stages)] for _ in range(3))
    sf_full, sf_empty = [Barrier() for _ in range(sf_nb)], [Barrier() for _ in range(sf_nb)]
    smem_sf = [None] * stages          # k-block whose scale atoms the stage's smem currently holds (None while a load is in flight)
    smem_ops = [None] * stages
    tmem_sf = [None] * sf_nb           # k-block whose scales the TMEM buffer holds
    inflight_loads = []                # asynchronous TMA loads: ("sf"|"ops", stage, kb)
    cp_q, mma_q = deque(), deque()     # issued, unexecuted work of the copy thread / the MMA thread
    pending_cp_reads = [0] * stages    # copies issued but not executed that read smem_sf[s]
    pending_mma_reads = [0] * sf_nb    # MMAs issued but not executed that read TMEM buffer b
    done = []                          # (kb, scales used) in execution order
    prod = {"kb": 0}
    cpt = {"kb": 0}
    mma = {"kb": 0}

    def st(kb):
        return kb % stages, (kb // stages) & 1

    def sb(kb):
        return (kb % stages, (kb // stages) & 1) if per_stage else (kb % sf_nb, (kb // sf_nb) & 1)

    guard = 0
    while len(done) < n_kb:
        guard += 1
        if guard > 400000:
            return "deadlock"
        actor = rng.choice(["prod", "land", "cpt", "cpq", "mma", "mmaq"])
        if actor == "prod" and prod["kb"] < n_kb:
            s, ph = st(prod["kb"])
            if not empty[s].passed(ph ^ 1):
                continue
            if pending_cp_reads[s]:
                return f"scale load of k-block {prod['kb']} overwrites smem a pending copy still reads"
            smem_sf[s] = smem_ops[s] = None
            inflight_loads.append(("sf", s, prod["kb"]))
            inflight_loads.append(("ops", s, prod["kb"]))
            prod["kb"] += 1
        elif actor == "land" and inflight_loads:
            kind, s, kb = inflight_loads.pop(rng.randrange(len(inflight_loads)))
            if kind == "sf":
                smem_sf[s] = kb
                sf_ld[s].complete()
            else:
                smem_ops[s] = kb
                full[s].complete()
        elif actor == "cpt" and cpt["kb"] < n_kb:
            kb = cpt["kb"]
            s, ph = st(kb)
            b, bph = sb(kb)
            if not sf_ld[s].passed(ph):
                continue
            if use_sf_empty and not sf_empty[b].passed(bph ^ 1):
                continue
            cp_q.append(("cp", s, b, kb))
            cp_q.append(("commit", b))
            pending_cp_reads[s] += 1
            cpt["kb"] += 1
        elif actor == "cpq" and cp_q:
            op = cp_q.popleft()
            if op[0] == "cp":
                _, s, b, kb = op
                if smem_sf[s] != kb:
                    return f"copy of k-block {kb} read smem holding {smem_sf[s]}"
                if pending_mma_reads[b]:
                    return f"copy of k-block {kb} overwrote TMEM scale buffer {b} under an unretired MMA"
                tmem_sf[b] = kb
                pending_cp_reads[s] -= 1
            else:
                sf_full[op[1]].complete()
        elif actor == "mma" and mma["kb"] < n_kb:
            kb = mma["kb"]
            s, ph = st(kb)
            b, bph = sb(kb)
            if not full[s].passed(ph):
                continue
            if bug != "mma_skips_sf_full" and not sf_full[b].passed(bph):
                continue
            mma_q.append(("mma", s, b, kb))
            mma_q.append(("commit", s, b))
            pending_mma_reads[b] += 1
            mma["kb"] += 1
        elif actor == "mmaq" and mma_q:
            op = mma_q.popleft()
            if op[0] == "mma":
                _, s, b, kb = op
                if smem_ops[s] != kb:
                    return f"MMA of k-block {kb} read operands of {smem_ops[s]}"
                if tmem_sf[b] != kb:
                    return f"MMA of k-block {kb} multiplied with the scales of k-block {tmem_sf[b]}"
                done.append(kb)
                pending_mma_reads[b] -= 1
            else:
                empty[op[1]].complete()
                if use_sf_empty:
                    sf_empty[op[2]].complete()
    return "ok" if done == list(range(n_kb)) else "out of order"


@pytest.mark.parametrize("n_kb,stages,sf_nb", [(1, 6, 6), (40, 6, 6), (40, 5, 5), (23, 3, 3), (40, 6, 2), (40, 6, 5), (17, 8, 4)])
def test_scale_pipeline_is_hazard_free_and_live(n_kb, stages, sf_nb):
    assert all(run_model(n_kb, stages, sf_nb, seed) == "ok" for seed in range(120))


def test_model_catches_a_short_ring_without_its_empty_barrier():
    # fewer TMEM scale buffers than stages and no sf_empty: a copy for k-block kb + SF_NB lands under the MMAs of k-block kb
    outcomes = {run_model(40, 6, 2, seed, bug="ring_without_sf_empty") for seed in range(120)}
    assert outcomes - {"ok"}


def test_model_catches_an_mma_thread_that_ignores_sf_full():
    outcomes = {run_model(40, 6, 6, seed, bug="mma_skips_sf_full") for seed in range(120)}
    assert outcomes - {"ok"}
