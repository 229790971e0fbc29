"""Event-level model of the producer / consumer protocol of the warp-specialised tridiagonal sweep kernel
(csrc/tridiag.cu: k_tri_sweep_ws).  The kernel's two roles and the asynchronous TMA engine are three actors that a
random scheduler interleaves; the mbarriers (`full[st]`: one expect_tx arrival + the bytes of three tensor loads,
`empty[st]`: one arrival of the consumer's lane 0) are modelled with their phase bit exactly as the code uses them
(try_wait on the parity of use i / NS, the producer's wait on use i / NS - 1).

Checked for every interleaving tried: no deadlock; the consumer fetches box i from a stage that holds box i completely
and has no load in flight; the producer never overwrites a stage before the consumer has released it; every box is
consumed once, in order.  `python tools/tri_ws_model.py` runs a sweep of (boxes, ring) pairs; tests/test_tri_ws_model_cpu.py
runs a smaller one.
"""
import random


class MBarrier:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier expects in this phase"
        self.tx += expect_tx
        self.pending -= 1
        self._maybe_flip()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_flip()

    def done(self, parity):          # mbarrier.try_wait.parity: has the phase with this parity completed?
        return (self.phase & 1) != parity


def simulate(nb, ns, seed, loads_per_box=3):
    assert ns >= 2, "the double-buffered consumer needs at least two boxes in the ring"
    rng = random.Random(seed)
    full = [MBarrier(1) for _ in range(ns)]
    empty = [MBarrier(1) for _ in range(ns)]
    stage_box = [[None] * loads_per_box for _ in range(ns)]     # what each of the three arrays of a stage holds
    inflight = []                                               # (stage, array, box) tensor loads not yet landed
    released = [True] * ns                                      # the consumer has taken the stage's last content
    prod_i = 0                                                  # next box the producer issues
    cons_state, cons_i = "wait_first", 0                        # consumer program counter
    fetched, consumed = [], []
    have = {}                                                   # register sets: box -> contents
    steps = 0
    while len(consumed) < nb:
        steps += 1
        assert steps < 200 * (nb + ns) + 1000, "no progress: deadlock"
        actors = []
        # producer: for box i >= ns wait for use (i / ns) - 1 of the stage to be released, then expect_tx + 3 loads
        if prod_i < nb:
            st = prod_i % ns
            if prod_i < ns or empty[st].done(((prod_i // ns) - 1) & 1):
                actors.append("produce")
        if inflight:
            actors.append("tma")
        # consumer (k_tri_sweep_ws: step lambda): fetch box i + 1 after the chain of box i, then store + release box i
        if cons_state == "wait_first":
            if full[0].done(0):
                actors.append("consume")
        elif cons_state == "chain":
            actors.append("consume")
        elif cons_state == "fetch_next":
            nxt = cons_i + 1
            if nxt >= nb or full[nxt % ns].done((nxt // ns) & 1):
                actors.append("consume")
        elif cons_state == "release":
            actors.append("consume")
        assert actors, "every actor is blocked: deadlock"
        a = rng.choice(actors)
        if a == "produce":
            st = prod_i % ns
            assert released[st], f"box {prod_i} issued into stage {st} before its previous box was fetched"
            assert not any(s == st for s, _, _ in inflight), "two boxes in flight into one stage"
            released[st] = False
            full[st].arrive(expect_tx=loads_per_box)
            for arr in range(loads_per_box):
                inflight.append((st, arr, prod_i))
            prod_i += 1
        elif a == "tma":
            st, arr, box = inflight.pop(rng.randrange(len(inflight)))
            stage_box[st][arr] = box
            full[st].complete_tx(1)
        else:
            if cons_state == "wait_first":
                have[0] = list(stage_box[0])
                fetched.append(0)
                assert have[0] == [0] * loads_per_box
                cons_state = "chain"
            elif cons_state == "chain":
                assert have[cons_i] == [cons_i] * loads_per_box, f"chain of box {cons_i} ran on {have[cons_i]}"
                cons_state = "fetch_next"
            elif cons_state == "fetch_next":
                nxt = cons_i + 1
                if nxt < nb:
                    st = nxt % ns
                    assert not any(s == st for s, _, _ in inflight), f"box {nxt} fetched with a load in flight"
                    have[nxt] = list(stage_box[st])
                    assert have[nxt] == [nxt] * loads_per_box, f"stage {st} holds {have[nxt]}, expected box {nxt}"
                    fetched.append(nxt)
                cons_state = "release"
            else:                                               # rows stored; lane 0 arrives on empty[i % ns]
                st = cons_i % ns
                released[st] = True
                empty[st].arrive()
                consumed.append(cons_i)
                del have[cons_i]
                cons_i += 1
                cons_state = "chain"
    assert consumed == list(range(nb)) and fetched == list(range(nb))
    assert not inflight and prod_i == nb
    return steps


if __name__ == "__main__":
    n = 0
    for ns in (2, 3, 8, 16, 32):        # the consumer fetches box i + 1 before it releases box i: a ring of one would deadlock
        for nb in (1, 2, 3, 7, 8, 9, 31, 33, 64, 129, 513):
            for seed in range(25):
                simulate(nb, ns, seed)
                n += 1
    print(f"{n} interleavings: no deadlock, every box consumed once and in order")
