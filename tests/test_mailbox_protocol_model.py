"""CPU: model check of the cross-GPU mailbox protocol used by reduce_all_sum_f32_xgpu (csrc/reduce.cu, XgpuParams).

Every rank, per collective call (epoch e): (1) stores (e, value) into slot[e & 1][rank] of EVERY rank's mailbox, one store
per peer in arbitrary order; (2) for every peer r, spins on its OWN slot[e & 1][r] until the tag equals e, then reads the
value; (3) sums in rank order.  A rank may start epoch e+1 as soon as it has gathered epoch e.  The model runs random
interleavings of those micro-steps and checks that every rank computes the right sum in every epoch; with a single slot
(no parity double-buffering) a fast rank overwrites a value a slow rank has not read yet, which the model also shows.
"""
import random

import pytest


def run_model(nranks, epochs, depth, seed):
    rng = random.Random(seed)
    mail = [[[(0, 0.0)] * nranks for _ in range(depth)] for _ in range(nranks)]   # mail[owner][parity][src] = (tag, value)
    value = lambda r, e: float(r * 1000 + e)                                       # noqa: E731
    # per-rank program counter: epoch, phase ("store"/"gather"), pending peers, gathered values
    st = [{"e": 1, "phase": "store", "todo": list(range(nranks)), "got": {}} for _ in range(nranks)]
    results = [[] for _ in range(nranks)]
    for r in range(nranks):
        rng.shuffle(st[r]["todo"])
    while any(s["e"] <= epochs for s in st):
        r = rng.choice([i for i, s in enumerate(st) if s["e"] <= epochs])
        s = st[r]
        e, par = s["e"], s["e"] % depth
        if s["phase"] == "store":
            peer = s["todo"].pop()
            mail[peer][par][r] = (e, value(r, e))                                  # one 64-bit store: tag and value together
            if not s["todo"]:
                s["phase"], s["todo"] = "gather", list(range(nranks))
                rng.shuffle(s["todo"])
        else:
            peer = s["todo"][-1]
            tag, v = mail[r][par][peer]
            if tag == e:                                                           # spin otherwise (try again later)
                s["got"][peer] = v
                s["todo"].pop()
            elif tag > e:
                return False                                                       # a newer epoch overwrote an unread value
            if not s["todo"]:
                results[r].append(sum(s["got"][p] for p in range(nranks)))
                s.update(e=e + 1, phase="store", todo=list(range(nranks)), got={})
                rng.shuffle(s["todo"])
    expect = [sum(value(r, e) for r in range(nranks)) for e in range(1, epochs + 1)]
    return all(res == expect for res in results)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_parity_double_buffering_is_sufficient(nranks):
    assert all(run_model(nranks, epochs=12, depth=2, seed=s) for s in range(200))


def test_a_single_slot_is_not():
    # without the parity bit a rank one epoch ahead clobbers a slot a slower rank has not consumed yet
    assert not all(run_model(4, epochs=12, depth=1, seed=s) for s in range(200))
