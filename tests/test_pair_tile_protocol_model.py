"""CPU: model check of the accumulator-unit protocol of the 512 x 256 pair tile (csrc/gemm_tcgen05.cu, MT = 2).

Actors per CTA pair, as in the kernel:
  * the MMA issuer walks tiles; per k-block it issues the MMAs of unit 0 then unit 1.  Before a unit's FIRST MMA of a tile it
    waits on tempty[unit] (parity aph ^ 1); after a unit's LAST MMA of a tile it issues a commit on tfull[unit].  Issued work
    goes into a FIFO: the tensor pipe executes MMAs and commits asynchronously, in issue order.
  * one epilogue warpgroup per unit: waits on tfull[unit] (parity aph), drains the accumulator into registers, arrives on
    tempty[unit], then runs its stores (which touch neither TMEM nor the barriers).
  * aph flips once per tile on every actor (one accumulator stage).
mbarrier parity semantics: a barrier starts in phase 0; try_wait.parity(p) succeeds once the phase of parity p has completed,
so waiting on parity 1 succeeds on a fresh barrier.

Checked under random interleavings: the pipe never accumulates into a unit whose previous tile has not been drained, every
epilogue sees exactly its own tile's k-blocks, and nobody deadlocks.  Two seeded bugs show the model can fail: waiting on
unit 0's barrier for both units, and committing a unit before its last MMA.
"""
import random
from collections import deque

import pytest


class Barrier:
    def __init__(self):
        self.phase = 0                      # number of completed phases

    def passed(self, parity):               # try_wait.parity
        return (self.phase & 1) != parity

    def complete(self):
        self.phase += 1


def run_model(tiles, num_kb, seed, bug=None):
    rng = random.Random(seed)
    tfull, tempty = [Barrier(), Barrier()], [Barrier(), Barrier()]
    acc = [{"tile": None, "kbs": 0, "draining": False} for _ in range(2)]   # TMEM contents of each unit
    fifo = deque()                                                           # issued, not yet executed tensor-pipe work
    drained = {0: [], 1: []}
    # MMA issuer program: list of micro-steps
    issuer = {"tile": 0, "kb": 0, "unit": 0, "stage": "wait", "aph": 0}
    epi = [{"tile": 0, "aph": 0, "stage": "wait", "stores": 0} for _ in range(2)]
    guard = 0
    while (issuer["tile"] < tiles or fifo or any(e["tile"] < tiles for e in epi)):
        guard += 1
        if guard > 200000:
            return "deadlock"
        actor = rng.choice(["issuer", "pipe", "epi0", "epi1"])
        if actor == "issuer" and issuer["tile"] < tiles:
            u, kb = issuer["unit"], issuer["kb"]
            if issuer["stage"] == "wait":
                if kb == 0:
                    wait_on = 0 if bug == "wait_unit0_only" else u
                    if not tempty[wait_on].passed(issuer["aph"] ^ 1):
                        continue                                             # spin
                issuer["stage"] = "issue"
            else:
                last = kb + 1 == num_kb
                if bug == "commit_before_last_mma" and last:
                    fifo.append(("commit", u, issuer["tile"]))
                    fifo.append(("mma", u, issuer["tile"], kb))
                else:
                    fifo.append(("mma", u, issuer["tile"], kb))
                    if last:
                        fifo.append(("commit", u, issuer["tile"]))
                issuer["stage"] = "wait"
                if u == 0:
                    issuer["unit"] = 1
                else:
                    issuer["unit"] = 0
                    issuer["kb"] += 1
                    if issuer["kb"] == num_kb:
                        issuer.update(kb=0, tile=issuer["tile"] + 1, aph=issuer["aph"] ^ 1)
        elif actor == "pipe" and fifo:
            op = fifo.popleft()
            if op[0] == "mma":
                _, u, t, kb = op
                a = acc[u]
                if a["draining"]:
                    return f"mma into unit {u} while its epilogue is reading it"
                if kb == 0:
                    if a["tile"] is not None and a["tile"] not in drained[u]:
                        return f"tile {a['tile']} of unit {u} overwritten before it was drained"
                    a.update(tile=t, kbs=0)
                elif a["tile"] != t:
                    return "accumulating into a foreign tile"
                a["kbs"] += 1
            else:
                tfull[op[1]].complete()
        elif actor in ("epi0", "epi1"):
            u = int(actor[-1])
            e = epi[u]
            if e["tile"] >= tiles:
                continue
            if e["stage"] == "wait":
                if not tfull[u].passed(e["aph"]):
                    continue
                acc[u]["draining"] = True
                e["stage"] = "drain"
            elif e["stage"] == "drain":                                      # tcgen05.ld of the whole unit into registers
                a = acc[u]
                if a["tile"] != e["tile"] or a["kbs"] != num_kb:
                    return f"epilogue of tile {e['tile']} unit {u} read tile {a['tile']} with {a['kbs']}/{num_kb} k-blocks"
                drained[u].append(e["tile"])
                a["draining"] = False
                tempty[u].complete()                                         # (4 warps x 2 CTAs arrive; one phase completion)
                e.update(stage="store", stores=rng.randint(0, 6))
            else:                                                            # staging + TMA stores: no TMEM, no barriers
                if e["stores"] > 0:
                    e["stores"] -= 1
                else:
                    e.update(stage="wait", tile=e["tile"] + 1, aph=e["aph"] ^ 1)
    return "ok" if drained[0] == list(range(tiles)) and drained[1] == list(range(tiles)) else "incomplete"


@pytest.mark.parametrize("tiles,num_kb", [(1, 1), (1, 5), (2, 1), (7, 3), (14, 4)])
def test_two_unit_protocol_is_hazard_free_and_live(tiles, num_kb):
    assert all(run_model(tiles, num_kb, seed) == "ok" for seed in range(150))


def test_model_catches_a_shared_empty_barrier():
    # waiting on unit 0's hand-back for both units lets the next tile's unit-1 MMAs run over an undrained accumulator
    outcomes = {run_model(6, 2, seed, bug="wait_unit0_only") for seed in range(150)}
    assert outcomes - {"ok"}


def test_model_catches_an_early_commit():
    outcomes = {run_model(4, 3, seed, bug="commit_before_last_mma") for seed in range(150)}
    assert outcomes - {"ok"}
